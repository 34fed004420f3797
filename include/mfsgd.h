/*
 * mfsgd.h -- C ABI of libmfsgd.so, the B200-native matrix-factorization SGD engine.
 *
 * This is the drop-in boundary for the factorization path of MatrixFactorizationSGD.java
 * (/root/reference/README.md:1 names the class; the reference ships no source, so the interface
 * replaced is that of the committed stand-in baseline/java/MatrixFactorizationSGD.java). A Java host
 * binds these symbols with Panama FFM (java/MatrixFactorizationSGDGpu.java); see INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; every pointer argument is caller-owned, read or written only during the
 *    call, never retained or freed by the library (the opaque handle excepted);
 *  - return value 0 = MFSGD_OK, negative = error code; text via mfsgd_last_error() (thread-local);
 *  - no exceptions, no abort, no CPU fallback: without a usable sm_100 device every compute entry
 *    point fails with MFSGD_E_CUDA;
 *  - a handle is used by one caller thread at a time; distinct handles are independent.
 */
#ifndef MFSGD_H
#define MFSGD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFSGD_ABI_VERSION 3

#if defined(__GNUC__)
#define MFSGD_API __attribute__((visibility("default")))
#else
#define MFSGD_API
#endif

#define MFSGD_OK             0
#define MFSGD_E_INVALID_ARG -1
#define MFSGD_E_CUDA        -2
#define MFSGD_E_NCCL        -3
#define MFSGD_E_OOM         -4
#define MFSGD_E_STATE       -5

/* mfsgd_config.mode */
#define MFSGD_MODE_DETERMINISTIC 0 /* one warp, records strictly in stand-in shuffle order (parity mode) */
#define MFSGD_MODE_HOGWILD       1 /* one GPU, lock-free, full grid                                     */
#define MFSGD_MODE_DSGD          2 /* n_gpus ring members: P striped, Q shard groups rotate             */

/* mfsgd_config.scatter */
#define MFSGD_SCATTER_STORE  0 /* st.global of the updated rows (Hogwild, last writer wins)  */
#define MFSGD_SCATTER_ATOMIC 1 /* red.global.add.v4.f32 of the row deltas (no lost updates)  */
#define MFSGD_SCATTER_ATOMIC_Q 2 /* store p_u, red q_i */
#define MFSGD_SCATTER_ATOMIC_P 3 /* red p_u, store q_i */

/* mfsgd_config.model -- model extension (stand-in factorizeModel :305): r ~ mu + b_u + b_i + p_u . q_i */
#define MFSGD_MODEL_GLOBAL_MEAN 1u /* mu = mean of the training ratings (stand-in globalMean :272), subtracted from every
                                      rating as it is loaded; mfsgd_get_model returns it                               */
#define MFSGD_MODEL_BIASES      2u /* user and item biases, b <- b + lr * (e - lambda * b), initialised to 0            */

/* mfsgd_config.p_storage -- mixed-precision factor storage: how the rows of P are KEPT in device memory.
 * Arithmetic is binary32 either way; P, Q cross this interface as binary32 (get: widened exactly; set: rounded to nearest). */
#define MFSGD_STORAGE_F32 0      /* the reference's: binary32                                                              */
#define MFSGD_STORAGE_F16 1      /* binary16 rows, narrowed with stochastic rounding from a counter hash of (seed, epoch, u, i,
                                    chunk): half the bytes of the update's dominant stream. Needs MFSGD_SCATTER_STORE and, outside
                                    DETERMINISTIC mode, the FMA arrangement (no MFSGD_FLAG_EXACT_ARITH). Q stays binary32.    */

/* mfsgd_config.flags */
#define MFSGD_FLAG_TIME_KERNELS   1u /* bracket every update launch with events -> stats.update_kernel_ms */
#define MFSGD_FLAG_VIRTUAL_RING   2u /* place all n_gpus ring members on one device (scheduler test mode) */
#define MFSGD_FLAG_NO_SHUFFLE     4u /* skip the per-epoch reshuffle (measurement aid)                    */
#define MFSGD_FLAG_SPLIT_SHARDS  16u /* one launch pair per item sub-shard (shards_per_gpu > 1) instead of one per
                                        shard group: a single GPU then replays the launch sizes of a larger ring */
#define MFSGD_FLAG_MATERIALIZE_SHUFFLE 32u /* run the reshuffle kernel every epoch (two record buffers) instead of letting the
                                        update kernels read each bucket through its per-epoch permutation (the default)  */
#define MFSGD_FLAG_EXACT_ARITH    8u /* HOGWILD/DSGD kernels apply the reference rule operation by operation
                                        (no FMA) instead of the FFMA2 arrangement (a few ulp apart, 3x the issue slots);
                                        DETERMINISTIC mode is always exact                                   */

typedef struct mfsgd_handle mfsgd_handle;

/* POD, caller-owned, copied by mfsgd_create. Zero-fill then set; mfsgd_config_default() does that. */
typedef struct mfsgd_config {
    int32_t  n_users;          /* rows of P                                                        */
    int32_t  n_items;          /* rows of Q                                                        */
    int32_t  k;                /* rank; k % 4 == 0, 4 <= k <= 512                                   */
    float    lr;               /* learning rate (constant)                                         */
    float    lambda;           /* L2 regularisation                                                */
    float    init_scale;       /* <= 0 -> 1/sqrt(k) (MatrixFactorizationSGD.java:63)                */
    uint64_t seed;             /* init + shuffle seed (MatrixFactorizationSGD.java:109 `seed`)      */
    int32_t  mode;             /* MFSGD_MODE_*                                                     */
    int32_t  n_gpus;           /* ring size G: 1 for DETERMINISTIC/HOGWILD; 1,2,4,8 for DSGD        */
    int32_t  stripes_per_gpu;  /* P sub-stripes per ring member, 0 = auto (sized to stay L2-resident) */
    int32_t  shards_per_gpu;   /* Q sub-shards per ring member, 0 = auto (1)                         */
    int32_t  scatter;          /* MFSGD_SCATTER_*                                                  */
    uint32_t flags;            /* MFSGD_FLAG_*                                                     */
    int32_t  device;           /* first CUDA ordinal to use; single-process ring uses device..device+G-1 */
    int32_t  world_size;       /* 1 = this process drives all ring members; G = one process per member */
    int32_t  rank;             /* ring member driven by this process when world_size == n_gpus       */
    uint8_t  nccl_id[128];     /* world_size > 1: ncclUniqueId from mfsgd_nccl_unique_id() of rank 0 */
    int32_t  ctas_per_sm;      /* 0 = auto; update-kernel CTAs per SM (tuning aid)                   */
    int32_t  rounds;           /* 0 = auto (>= 8 launches per epoch where the data allows); each sub-epoch visits its
                                  P sub-stripes in `rounds` interleaved passes */
    float    hot_share;        /* items rated by >= this share of the training set -- and by >= 16 ratings per
                                  (sub-stripe, ring member) bucket -- take the run path (q_i register-resident for a
                                  run of its ratings, model-averaged over concurrent runs); 0 = default 1e-6, < 0 = off */
    int32_t  hot_chunk;        /* max records per run (one sub-warp); 0 = auto: the per-sub-warp share of a launch,
                                  64..1024                                                                  */
    float    merge_boost;      /* runs of one item that share a launch are merged with weight min(1, merge_boost / runs);
                                  0 = default 1.25, 1 = plain model averaging; must be < 2                      */
    uint32_t model;            /* MFSGD_MODEL_* : 0 = the reference model r ~ p_u . q_i                          */
    float    p_atomic_threshold; /* run kernel: a user expected to have >= this many ratings in flight at once (its share of a
                                  launch's records x the ratings the resident sub-warps hold in flight) is a HEAVY user: its row is
                                  updated in memory with red.global.add instead of a store, so concurrent updates are not lost.
                                  0 = default 0.25, < 0 = never (round-1 behaviour); MFSGD_SCATTER_ATOMIC_P = every user       */
    /* model extension, all off at 0 */
    float    lr_decay;         /* learning-rate schedule: epoch e runs at lr_e, lr_0 = lr, lr_(e+1) = lr_e * lr_decay (one binary32
                                  multiply per epoch: MatrixFactorizationSGD.java learningRate); 0 = constant rate, else in (0, 1] */
    int32_t  early_stop_patience; /* > 0: mfsgd_train evaluates the held-out set after every epoch and returns once its RMSE has
                                  failed `patience` times in a row to fall below best * (1 - early_stop_min_delta)
                                  (MatrixFactorizationSGD.java factorizeEarlyStop); needs a held-out set; mfsgd_get_progress reports */
    float    early_stop_min_delta; /* relative improvement that counts, in [0, 1)                                 */
    int32_t  p_storage;        /* MFSGD_STORAGE_*                                                                  */
    int32_t  reserved[2];
} mfsgd_config;

/* One entry per epoch, filled by mfsgd_train when `stats` is non-null. Times are device times (CUDA
 * events on the engine's own streams), max over the ring members this process drives. */
typedef struct mfsgd_epoch_stats {
    int64_t updates;            /* rating updates applied by this process's ring members             */
    double  epoch_ms;           /* shuffle + all sub-epochs + rotations                              */
    double  shuffle_ms;         /* the reshuffle kernel alone                                        */
    double  update_kernel_ms;   /* MFSGD_FLAG_TIME_KERNELS: time inside the update kernels (per sub-epoch span
                                   of the concurrent cold + hot-item launches), else 0                */
    int32_t update_launches;    /* update-kernel launches in the epoch                               */
    int32_t total_launches;     /* every kernel this library launched in the epoch                   */
    double  heldout_rmse;       /* NaN unless a held-out set is loaded and eval_every_epoch is on     */
    /* MFSGD_FLAG_TIME_KERNELS breakdown, summed over the epoch's sub-epochs (the two kernels overlap): */
    double  cold_ms;            /* fork -> last cold (full-grid) launch done                          */
    double  hot_ms;             /* fork -> last run-kernel launch done                                */
    double  exchange_ms;        /* join -> Q rotation enqueued on the compute stream done (NCCL path) */
} mfsgd_epoch_stats;

/* Synthetic power-law ratings (MatrixFactorizationSGD.java:220 syntheticRecord). Record n in
 * [0, n_total) is a pure function of (seed, n); those with hash64(seed,6,n) % 10 == 0 are held out. */
typedef struct mfsgd_synth_params {
    int64_t  n_total;
    uint64_t seed;
    int32_t  log2_alpha_user;   /* 2  (alpha 4)                                                      */
    int32_t  log2_alpha_item;   /* 3  (alpha 8; top 1 % of items ~30 %), 4 = heavy (alpha 16; ~60 %)  */
    double   c_user;            /* 0.25                                                              */
    double   c_item;            /* 0.375                                                             */
    float    planted_amplitude; /* 0 = 0.8660254 (planted dot of unit variance x 0.25: noise-dominant sets);
                                   1.7320508 = the signal-dominant variant                */
    float    noise_scale;       /* 0 = 0.5; 0.125 = the signal-dominant variant                         */
} mfsgd_synth_params;

/* Layout report for tests and tools (all counts are for this process's ring members). */
typedef struct mfsgd_layout_info {
    int32_t n_gpus, stripes_per_gpu, shards_per_gpu;
    int32_t user_blocks;        /* n_gpus * stripes_per_gpu                                          */
    int32_t item_blocks;        /* n_gpus * shards_per_gpu                                           */
    int64_t n_train_local;      /* records held by this process                                      */
    int64_t n_heldout_local;
    int64_t n_train_total;      /* records in the whole data set (all processes)                     */
    int32_t rounds;             /* interleaved passes per sub-epoch (see mfsgd_config.rounds)        */
    int32_t n_hot_items;        /* items on the run path (whole data set)                            */
    int32_t n_heavy_users;      /* users whose rows the run kernel updates with red.global.add (whole data set) */
    int32_t run_length;         /* longest run of the plan (mfsgd_config.hot_chunk resolved)         */
} mfsgd_layout_info;

MFSGD_API int  mfsgd_abi_version(void);
MFSGD_API const char* mfsgd_last_error(void);   /* library-owned, valid until this thread's next mfsgd call */
MFSGD_API int  mfsgd_device_count(int32_t* count);
MFSGD_API int  mfsgd_config_default(mfsgd_config* cfg);

MFSGD_API int  mfsgd_create(const mfsgd_config* cfg, mfsgd_handle** out);
MFSGD_API void mfsgd_destroy(mfsgd_handle* h);

/* Training triplets (MatrixFactorizationSGD.java:109 users/items/ratings). Host pointers. In a
 * multi-process ring every rank passes the same full arrays; each keeps its own user stripe. */
MFSGD_API int  mfsgd_load_ratings(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings,
                        int64_t n);
/* Multi-process ring: every rank passes ITS OWN disjoint slice of the training triplets (any split; slices may be empty).
 * The ranks sum their per-row counts (ncclAllReduce), derive the same stripe bounds, and exchange the records by stripe owner
 * on the device (grouped ncclSend/ncclRecv): host-to-device traffic is 12 B per rating in total instead of per rank.
 * Collective: every rank of the ring must call it. With world_size == 1 it is mfsgd_load_ratings. */
MFSGD_API int  mfsgd_load_ratings_sharded(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings,
                                int64_t n_local);
/* Optional held-out triplets kept on the device for mfsgd_rmse_heldout / per-epoch evaluation. */
MFSGD_API int  mfsgd_load_heldout(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings,
                        int64_t n);
/* Extension: generate, split and bucket the synthetic data set on the device (no host arrays). */
MFSGD_API int  mfsgd_generate_synthetic(mfsgd_handle* h, const mfsgd_synth_params* sp, int64_t* n_train, int64_t* n_heldout);

MFSGD_API int  mfsgd_init_factors(mfsgd_handle* h);                        /* MatrixFactorizationSGD.java:53 */
MFSGD_API int  mfsgd_set_factors(mfsgd_handle* h, const float* P, const float* Q);  /* full nU*k, nI*k host arrays */
MFSGD_API int  mfsgd_get_factors(mfsgd_handle* h, float* P, float* Q);     /* multi-process: only this rank's rows are written */
/* Model extension: the global mean (0 when off) and the biases (full n_users / n_items host arrays, nullable; multi-process:
 * only this rank's rows are written). MFSGD_E_STATE for the bias arrays when MFSGD_MODEL_BIASES is off. */
MFSGD_API int  mfsgd_get_model(mfsgd_handle* h, float* global_mean, float* user_bias, float* item_bias);
MFSGD_API int  mfsgd_set_biases(mfsgd_handle* h, const float* user_bias, const float* item_bias);
/* Epochs trained since the ratings were loaded, the learning rate the next epoch would use (lr_decay), and whether the last
 * mfsgd_train call returned on the early-stopping rule. Every out pointer is nullable. */
MFSGD_API int  mfsgd_get_progress(mfsgd_handle* h, int32_t* epochs_done, float* next_lr, int32_t* stopped_early);
/* Row ranges this process owns (users [u_lo,u_hi), items [i_lo,i_hi)) -- whole matrices when world_size==1. */
MFSGD_API int  mfsgd_get_partition(mfsgd_handle* h, int32_t* u_lo, int32_t* u_hi, int32_t* i_lo, int32_t* i_hi);

/* `epochs` more epochs of MatrixFactorizationSGD.java:127-133; stats nullable, length epochs. */
MFSGD_API int  mfsgd_train(mfsgd_handle* h, int32_t epochs, mfsgd_epoch_stats* stats);
/* DETERMINISTIC mode only: also returns the error e of every update in visiting order
 * (err_trace length = epochs * n_train). */
MFSGD_API int  mfsgd_train_traced(mfsgd_handle* h, int32_t epochs, mfsgd_epoch_stats* stats, float* err_trace);
MFSGD_API int  mfsgd_set_eval_every_epoch(mfsgd_handle* h, int32_t on);

/* MatrixFactorizationSGD.java:169 rmse. Multi-process: sse_out and n_out receive this rank's partial
 * sums (nullable); rmse_out is the rank-local value -- reduce the partials across ranks for the total. */
MFSGD_API int  mfsgd_rmse(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n,
                double* rmse_out);
MFSGD_API int  mfsgd_rmse_heldout(mfsgd_handle* h, double* rmse_out, double* sse_out, int64_t* n_out);
MFSGD_API int  mfsgd_rmse_train(mfsgd_handle* h, double* rmse_out, double* sse_out, int64_t* n_out);

/* One-shot form = the Java entry point 1:1 (init from cfg->seed, train `epochs`, copy P and Q out). */
MFSGD_API int  mfsgd_factorize(const int32_t* users, const int32_t* items, const float* ratings, int64_t n,
                     const mfsgd_config* cfg, int32_t epochs, float* P_out, float* Q_out);

/* Introspection for tests and tools. */
MFSGD_API int  mfsgd_get_layout_info(mfsgd_handle* h, mfsgd_layout_info* out);
/* user_bounds[user_blocks+1], item_bounds[item_blocks+1] (global row ids). */
MFSGD_API int  mfsgd_get_bounds(mfsgd_handle* h, int32_t* user_bounds, int32_t* item_bounds);
/* Ring member `member`'s current record layout: recs = 3*n int32 words (u,i,r-bits per record; bit 31 of u marks a heavy
 * user, see p_atomic_threshold -- mask with 0x7fffffff for the id);
 * block_offsets[stripes_per_gpu*(item_blocks+n_hot_items)+1]: the cold blocks (stripe-major), then one
 * bucket per (stripe, hot item). Pass recs=NULL to query *n only. */
MFSGD_API int  mfsgd_get_records(mfsgd_handle* h, int32_t member, int32_t* recs, int64_t* block_offsets, int64_t* n);
/* Runs the reshuffle kernel once for `epoch` (HOGWILD/DSGD) or the stand-in order sort (DETERMINISTIC). */
MFSGD_API int  mfsgd_shuffle_once(mfsgd_handle* h, int32_t epoch);

/* Test hook, host-only: plans the run kernel's work for caller-provided bucket offsets (csrc/run_plan.hpp).
 * block_off[stripes * (item_blocks + n_hot) + 1] as mfsgd_get_records reports them; *n_units: in = slots in the
 * unit_* arrays (each nullable), out = runs planned; visit_units[stripes * rounds * item_blocks + 1]. */
MFSGD_API int  mfsgd_plan_runs(const int64_t* block_off, int32_t stripes, int32_t n_hot, int32_t item_blocks,
                     const int32_t* hot_block_lo, const int32_t* hot_items, int32_t rounds, int32_t chunk, uint64_t seed,
                     int32_t member, float merge_boost, int64_t* unit_start, int32_t* unit_count, int32_t* unit_item,
                     float* unit_weight, int64_t* n_units, int32_t* visit_units);

/* Test hook, host-only: the automatic layout (csrc/run_plan.hpp) for a configuration, an L2 size and one ring member's
 * share of the data: P sub-stripes and Q sub-shards per member, interleaved passes, longest run. */
MFSGD_API int  mfsgd_plan_layout(const mfsgd_config* cfg, int64_t l2_bytes, int64_t member_records, int32_t member_users,
                       int64_t run_records, int32_t resident_ctas, int32_t* stripes, int32_t* shards, int32_t* rounds,
                       int32_t* run_length);

/* Teacher-forced per-update check: applies the update rule to n independent
 * (pre_p[j], pre_q[j], r[j]) row pairs on the device, returns post rows and errors. Host pointers. */
MFSGD_API int  mfsgd_apply_updates_forced(int32_t device, int32_t k, float lr, float lambda, int64_t n, const float* pre_p,
                                const float* pre_q, const float* r, float* post_p, float* post_q, float* err);
/* Device twin of MatrixFactorizationSGD.java:220 for records [start,start+count): host output arrays. */
MFSGD_API int  mfsgd_generate_to_host(int32_t device, const mfsgd_synth_params* sp, int32_t n_users, int32_t n_items,
                            int64_t start, int64_t count, int32_t* users, int32_t* items, float* ratings,
                            uint8_t* held);

/* Multi-process ring bootstrap: rank 0 calls this, ships the 128 bytes to every rank (any channel),
 * every rank puts them in mfsgd_config.nccl_id. */
MFSGD_API int  mfsgd_nccl_unique_id(uint8_t out[128]);

/* Ratings-file ingest: text file -> the triplet arrays of mfsgd_load_ratings / mfsgd_factorize, sparse file
 * ids compacted to dense rows (ascending file id). Host-only: works without a GPU. */
#define MFSGD_FORMAT_AUTO          0 /* NETFLIX_PRIZE if the first data line is "<digits>:", else TRIPLETS           */
#define MFSGD_FORMAT_TRIPLETS      1 /* "user item rating [...]" per line; separators tab blank , ; : | (MovieLens u.data,
                                        ratings.csv with its header line, ratings.dat with "::"); non-numeric lines skipped */
#define MFSGD_FORMAT_NETFLIX_PRIZE 2 /* "movie:" lines, each followed by its "customer,rating[,date]" lines              */
typedef struct mfsgd_ratings {
    int32_t* users;      /* n dense user rows                                   (library-owned: mfsgd_free_ratings)   */
    int32_t* items;      /* n dense item rows                                                                         */
    float*   ratings;    /* n ratings                                                                                 */
    int64_t  n;
    int32_t  n_users;    /* distinct users = rows of P                                                                */
    int32_t  n_items;    /* distinct items = rows of Q                                                                */
    int64_t* user_ids;   /* n_users: dense row -> id in the file (ascending)                                          */
    int64_t* item_ids;   /* n_items                                                                                   */
    int32_t  format;     /* the format that was parsed (AUTO resolved)                                                */
    int32_t  reserved;
} mfsgd_ratings;
MFSGD_API int  mfsgd_read_ratings(const char* path, int32_t format, mfsgd_ratings* out);
MFSGD_API void mfsgd_free_ratings(mfsgd_ratings* r);

/* Diagnostic (bench.py): measured ceilings of the update path's access pattern on `device`, taken in the caller's process:
 * random 512-B row gather + scatter inside a buffer_mb-sized (L2-resident: 61 MB is one Netflix-shaped P sub-stripe) buffer
 * counting 1024 B per row, the same rows read only (512 B per row), and a 1 GB streaming copy (read + write bytes). ~50 ms. */
typedef struct mfsgd_ceilings {
    double  row_gather_scatter_gbs;
    double  row_gather_only_gbs;
    double  hbm_stream_copy_gbs;
    double  buffer_mb;
    double  l2_mb;             /* cudaDeviceProp::l2CacheSize */
    int32_t sm_count;
    int32_t reserved;
} mfsgd_ceilings;
MFSGD_API int  mfsgd_measure_ceilings(int32_t device, double buffer_mb, mfsgd_ceilings* out);

/* Device buffers of destroyed handles are kept in a process-wide cache and reused by later handles (cudaFree of gigabytes
 * costs up to a second and synchronises the device); this returns them to the driver. MFSGD_NO_CACHE=1 disables the cache. */
MFSGD_API int  mfsgd_release_cached_memory(void);

/* Pinned host staging helpers (optional; plain host memory works too, at lower H2D bandwidth). */
MFSGD_API int  mfsgd_host_alloc(void** out, int64_t bytes);
MFSGD_API int  mfsgd_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* MFSGD_H */
