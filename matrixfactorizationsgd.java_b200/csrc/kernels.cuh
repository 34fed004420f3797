// kernels.cuh -- host-callable launchers of the sm_100a kernels (definitions in kernels_*.cu).
// Every launcher enqueues on `stream`, returns cudaGetLastError(), and adds the number of kernels it
// launched to *launches (nullable) -- that count feeds mfsgd_epoch_stats.total_launches.
#pragma once
#include "common.cuh"

namespace mfsgd {

// Sub-warp geometry for rank k: `lanes` lanes cooperate on one rating, each holding
// `vec` float4 chunks of the row; chunk c belongs to lane c % lanes.
struct Geometry {
    int lanes, vec;
    bool full;  // lanes * vec == k/4 -> no tail predicate
};
inline Geometry geometry_for(int k) {
    int chunks = k / 4, lanes = 1;
    while (lanes < chunks && lanes < 32) lanes <<= 1;
    Geometry g;
    g.lanes = lanes;
    g.vec = (chunks + lanes - 1) / lanes;
    g.full = (lanes * g.vec == chunks);
    return g;
}
inline bool rank_supported(int k) { return k >= 4 && k <= 512 && (k % 4) == 0; }

// Expands CALL(LANES, VEC, FULL) for the geometry g (compile-time sub-warp shapes).
#define MFSGD_DISPATCH_GEOMETRY(g, CALL)                                   \
    do {                                                                   \
        if ((g).vec == 1) {                                                \
            switch ((g).lanes) {                                           \
                case 1:  if ((g).full) { CALL(1, 1, true); }  else { CALL(1, 1, false); }  break;  \
                case 2:  if ((g).full) { CALL(2, 1, true); }  else { CALL(2, 1, false); }  break;  \
                case 4:  if ((g).full) { CALL(4, 1, true); }  else { CALL(4, 1, false); }  break;  \
                case 8:  if ((g).full) { CALL(8, 1, true); }  else { CALL(8, 1, false); }  break;  \
                case 16: if ((g).full) { CALL(16, 1, true); } else { CALL(16, 1, false); } break;  \
                default: if ((g).full) { CALL(32, 1, true); } else { CALL(32, 1, false); } break;  \
            }                                                              \
        } else if ((g).vec == 2) {                                         \
            if ((g).full) { CALL(32, 2, true); } else { CALL(32, 2, false); }  \
        } else if ((g).vec == 3) {                                         \
            if ((g).full) { CALL(32, 3, true); } else { CALL(32, 3, false); }  \
        } else {                                                           \
            if ((g).full) { CALL(32, 4, true); } else { CALL(32, 4, false); }  \
        }                                                                  \
    } while (0)

struct UpdateArgs {
    const Rec* recs;   // the member's record array (current layout)
    int64_t first;     // cold / deterministic launches: positions [first, first + n) of that array
    int64_t n;
    float* P;          // row (u - u_base) of the local P stripe
    float* Q;          // row (i - i_base) of the held Q shard group
    int32_t k, u_base, i_base;
    float lr, lambda;
    // virtual reshuffle (virt != 0): position j of a bucket reads record bucket_start + perm(j - bucket_start);
    // for cold launches the range lies inside the single block described here, hot units carry their own bucket.
    int32_t virt;
    uint32_t epoch, blk_id;
    uint64_t seed;
    int64_t blk_start, blk_n;
    // model extension (MatrixFactorizationSGD.java:282 sgdUpdateModel): user / item biases, indexed like P / Q; both null = off
    float* BU;
    float* BI;
    // mixed-precision factor storage (mfsgd_config.p_storage): != 0 -> P points at binary16 rows (common.cuh)
    int32_t p_half;
    // run kernel: CTAs that land on an SM with %smid >= sm_limit leave at once (their runs are claimed by the others), which keeps
    // those SMs free for the NCCL kernels of the pipelined Q rotation; 0 = use every SM
    int32_t sm_limit;
};

// (2) the SGD update kernel, Hogwild: full grid, one sub-warp per rating, software-pipelined gathers.
// min_windows: lower bound on (records in the launch) / (ratings in flight at once), see kernels_update.cu.
// fast: FMA2 arithmetic (kernels_update.cu); false = the reference rule op by op.
cudaError_t launch_sgd_update_hogwild(const UpdateArgs& a, int scatter, bool fast, int grid, int min_windows,
                                      cudaStream_t stream, int* launches);
// Resident CTAs per SM of the Hogwild kernel for rank k (occupancy query; sizes the grid).
cudaError_t hogwild_max_ctas_per_sm(int k, int scatter, bool fast, bool p_half, int* ctas);
cudaError_t hot_max_ctas_per_sm(int k, bool fast, bool p_red, bool p_half, int* ctas);
// Lanes per rating of the run kernel at rank k (kernels_hot.cu): a warp walks 32 / lanes runs side by side.
int run_kernel_lanes(int k);
// Deterministic parity mode: one warp, records strictly in order; err_trace nullable (n floats).
cudaError_t launch_sgd_update_deterministic(const UpdateArgs& a, float* err_trace, cudaStream_t stream, int* launches);
// Teacher-forced check: n independent row pairs.
cudaError_t launch_sgd_update_forced(int k, float lr, float lambda, int64_t n, const float* pre_p, const float* pre_q,
                                     const float* r, float* post_p, float* post_q, float* err, cudaStream_t stream);

// (2b) run path (kernels_hot.cu): one sub-warp per unit = a run of records that all rate the same item. q_i lives in
// the sub-warp's registers for the whole run (no L2 round trips, no contention on the row); the run's net
// change is merged into Q at the end, scaled by `weight` (model averaging across the units of one item).
struct HotUnit {
    int64_t start;     // first position (index into the member's record array)
    int64_t bstart;    // the (stripe, item) bucket the run lies in: [bstart, bstart + bn)
    int32_t count;
    int32_t item;      // global item id
    float   weight;    // 1 / (units of this item in the launch)
    int32_t bn;
    uint32_t bid;      // bucket id keying the per-epoch permutation
    int32_t pad;
};
// follows_hot_launch: the previous operation on `stream` is a hot launch of the same sub-epoch (it may then be overlapped
// by programmatic dependent launch, see kernels_update.cu).
// overlaps_previous: launch with programmatic stream serialisation (it follows a run launch on `stream` and
// hot_launch_overlaps says so); overlapped_by_next: the NEXT launch on the chain will overlap this one's tail.
// p_red: p_u is updated in memory by red.global.add of its increment (no lost updates between concurrent writers of a row).
cudaError_t launch_sgd_update_hot(const UpdateArgs& a, const HotUnit* units, int n_units, unsigned int* counter, bool fast, bool p_red,
                                  int grid, int max_sub_warps, bool overlaps_previous, bool overlapped_by_next, cudaStream_t stream,
                                  int* launches);
bool hot_launch_overlaps(int k, int n_units, int64_t n_records, int longest_run, int full_grid, int max_sub_warps);

// (4) ring window: wait on `stream` until *flag >= value (cyclic); fallback of the stream memory operation (kernels_ring.cu)
cudaError_t launch_ring_wait_flag(const uint32_t* flag, uint32_t value, cudaStream_t stream, int* launches);

// (3) held-out RMSE: adds sum (r - p_u.q_i)^2 over the records to *sse_accum (double, device).
// scratch: >= rmse_scratch_doubles() doubles of device memory owned by the caller.
int rmse_scratch_doubles();
// BU / BI: biases of the model extension (nullable together), indexed like P / Q.
// p_half: P points at binary16 rows (mfsgd_config.p_storage).
cudaError_t launch_rmse_sse(const Rec* recs, int64_t n, const float* P, bool p_half, const float* Q, const float* BU, const float* BI,
                            int32_t k, int32_t u_base, int32_t i_base, double* scratch, double* sse_accum, int n_sms,
                            cudaStream_t stream, int* launches);

// factor init (MatrixFactorizationSGD.java:53) for local rows [row_lo, row_lo + n_rows).
// half: the rows are binary16 (the binary32 value rounded to nearest even).
cudaError_t launch_init_factors(void* rows, bool half, int64_t n_rows, int32_t k, int64_t row_lo, uint64_t seed, uint64_t stream_id,
                                float scale, cudaStream_t stream, int* launches);
// binary16 rows <-> binary32 rows (n values): exact / round to nearest even
cudaError_t launch_widen_rows(const void* half_rows, int64_t n, float* out, cudaStream_t stream, int* launches);
cudaError_t launch_narrow_rows(const float* rows, int64_t n, void* half_out, cudaStream_t stream, int* launches);

// synthetic records [start, start+count) -> SoA + held flag (MatrixFactorizationSGD.java:220).
struct SynthArgs {
    uint64_t seed;
    int32_t n_users, n_items, l2au, l2ai;
    double cu, ci;
    float amplitude, noise_scale;   // planted amplitude and noise scale (stand-in defaults: PLANTED_AMPLITUDE, 0.5)
};
cudaError_t launch_generate(const SynthArgs& s, int64_t start, int64_t count, int32_t* u, int32_t* i, float* r,
                            uint8_t* held, cudaStream_t stream, int* launches);

// (1) bucketing + shuffle
// per-row rating counts of the training records (held != 0 records are skipped when held != nullptr)
// r / rating_sum (nullable together): *rating_sum += sum floor(r * 2^20) over the counted records (MatrixFactorizationSGD.java:272)
cudaError_t launch_count_rows(const int32_t* u, const int32_t* i, const uint8_t* held, int64_t n, uint32_t* user_cnt,
                              uint32_t* item_cnt, int32_t n_users, int32_t n_items, int* bad_flag, const float* r,
                              unsigned long long* rating_sum, cudaStream_t stream, int* launches);
// bounds[b] = first row whose exclusive cumulative count reaches b * total / nblocks; bounds[nblocks] = n_rows.
// cum = exclusive prefix sums of the counts (n_rows + 1 entries, cum[n_rows] = total).
cudaError_t launch_balanced_bounds(const uint64_t* cum, int32_t n_rows, int32_t nblocks, int32_t* bounds,
                                   cudaStream_t stream, int* launches);
// owner[row] = block b with bounds[b] <= row < bounds[b+1]
cudaError_t launch_fill_owner(const int32_t* bounds, int32_t nblocks, int32_t n_rows, uint16_t* owner,
                              cudaStream_t stream, int* launches);
// Block of a record for the ring member owning user blocks [ub_lo, ub_hi):
//   row = (owner_u[u] - ub_lo) / row_div, col = owner_i[i] / col_div, block = row * n_cols + col.
// Records of other members, or whose held flag differs from want_held, are skipped.
// Training layout: row_div = col_div = 1, n_cols = item_blocks. Held-out sets: row_div = stripes_per_gpu
// (one row), col_div = shards_per_gpu, n_cols = n_gpus (one block per Q shard group).
struct BucketArgs {
    const int32_t* u;
    const int32_t* i;
    const float* r;
    const uint8_t* held;   // nullable
    int64_t n;
    const uint16_t* owner_u;
    const uint16_t* owner_i;
    int32_t ub_lo, ub_hi, row_div, col_div, n_cols;
    int32_t want_held;     // 0: training records, 1: held-out records
    // hot items (training layout only; hot_index == nullptr disables): a record whose item has
    // hot_index[i] >= 0 goes to bucket hot_base + row * n_hot + hot_index[i] instead of its cold block.
    const int32_t* hot_index;
    int32_t hot_base, n_hot;
    const uint32_t* heavy_bits;   // nullable: bit u set -> the scattered record carries the heavy-user mark (common.cuh REC_USER_MASK)
    float center;                 // subtracted from every rating as it is scattered (the model extension's global mean; else 0)
};
inline int bucket_rows(const BucketArgs& b) { return (b.ub_hi - b.ub_lo + b.row_div - 1) / b.row_div; }
inline int bucket_block_count(const BucketArgs& b) {
    return bucket_rows(b) * b.n_cols + (b.hot_index ? bucket_rows(b) * b.n_hot : 0);
}
constexpr int MAX_BUCKETS = 1 << 20;      // per ring member: cold blocks + run (stripe, item) buckets. (2^23 was tried on the two largest
                                          // shapes squeezed onto ONE GPU: 7x more items on the run path, no faster -- their buckets are tiny.)
constexpr int MAX_SMEM_BUCKETS = 16384;   // up to here histogram/scatter keep per-CTA counters in shared memory
cudaError_t launch_block_histogram(const BucketArgs& b, unsigned long long* block_cnt, cudaStream_t stream, int* launches);
// cursors start as the exclusive offsets; each record claims a slot with atomicAdd.
cudaError_t launch_block_scatter(const BucketArgs& b, unsigned long long* cursors, Rec* out, cudaStream_t stream,
                                 int* launches);
// out[off[b] + j] = in[off[b] + perm_b(j)], perm_b = keyed Feistel bijection on [0, n_b) (epoch, block).
cudaError_t launch_block_shuffle(const Rec* in, Rec* out, const int64_t* block_off, int32_t nblocks, int64_t n,
                                 uint64_t seed, uint32_t epoch, uint32_t block_id_base, cudaStream_t stream, int* launches);
// SoA -> AoS in input order (deterministic mode keeps the caller's record order).
cudaError_t launch_pack_records(const int32_t* u, const int32_t* i, const float* r, int64_t n, float center, Rec* out,
                                cudaStream_t stream, int* launches);
// AoS -> SoA
cudaError_t launch_unpack_records(const Rec* in, int64_t n, int32_t* u, int32_t* i, float* r, cudaStream_t stream, int* launches);
// Stand-in visiting order (MatrixFactorizationSGD.java:72): out[j] = in[order_j]; device radix sort of the
// packed (key31<<32 | idx) words. temp/temp_bytes: caller-owned scratch, query with temp == nullptr.
cudaError_t deterministic_order_gather(const Rec* in, Rec* out, int32_t n, uint64_t seed, uint32_t epoch,
                                       uint64_t* keys_a, uint64_t* keys_b, void* temp, size_t* temp_bytes,
                                       cudaStream_t stream, int* launches);
// exclusive prefix sum of uint32 counts into uint64 cum[n+1] (cub::DeviceScan); query with temp == nullptr.
cudaError_t exclusive_cumsum_u32(const uint32_t* cnt, uint64_t* cum, int32_t n, void* temp, size_t* temp_bytes,
                                 cudaStream_t stream, int* launches);

}  // namespace mfsgd
