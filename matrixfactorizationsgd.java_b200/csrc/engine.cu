// engine.cu -- host side of libmfsgd.so: handle lifecycle, data staging/bucketing, and subsystem (4),
// the single-box DSGD scheduler (P row stripes stay put, Q shard groups rotate round the ring).
//
// Replaces the loop of baseline/java/MatrixFactorizationSGD.java:109-135 (factorize). C ABI in
// include/mfsgd.h. No CPU fallback: every compute entry point needs an sm_100 device.
#include <cuda.h>      // types of the stream memory operations only; libcuda is reached through cudaGetDriverEntryPoint
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <unordered_map>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mfsgd.h"
#include "kernels.cuh"
#include "run_plan.hpp"

using namespace mfsgd;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

namespace mfsgd {
// the same message slot for the library's other translation units (ratings_io.cpp)
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace mfsgd

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(e__ == cudaErrorMemoryAllocation ? MFSGD_E_OOM : MFSGD_E_CUDA, "%s:%d %s: %s", \
                        __FILE__, __LINE__, #call, cudaGetErrorString(e__));                          \
    } while (0)

#define CKRC(call)               \
    do {                         \
        int rc__ = (call);       \
        if (rc__ != MFSGD_OK) return rc__; \
    } while (0)

// No C++ exception crosses the C boundary (include/mfsgd.h: "no exceptions, no abort"): host allocations that fail
// (std::vector, new) surface as MFSGD_E_OOM.
template <typename F>
static int guarded(F&& body) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return fail(MFSGD_E_OOM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(MFSGD_E_STATE, "internal error: %s", e.what());
    } catch (...) {
        return fail(MFSGD_E_STATE, "internal error");
    }
}

// MFSGD_TRACE=1: phase timings on stderr (diagnostic aid)
static inline bool trace_on() { return getenv("MFSGD_TRACE") != nullptr; }     // read every time: callers switch it on mid-process
static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct PhaseTimer {
    const char* what;
    double t0;
    explicit PhaseTimer(const char* w) : what(w), t0(trace_on() ? now_s() : 0.0) {}
    ~PhaseTimer() {
        if (trace_on()) fprintf(stderr, "[mfsgd]   %s %.1f ms\n", what, (now_s() - t0) * 1e3);
    }
};

// The engine keeps up to ten streams per ring member busy (main, copy, two lanes of two, reshuffle, ...) and orders them with
// events and flag waits. With the default of 8 hardware work queues two of those streams can share a queue, and a wait that
// blocks one then holds the other one's launches back (false dependency). Ask for 32 queues unless the caller has chosen;
// this runs when the library is loaded, i.e. before the process creates its CUDA context through it.
__attribute__((constructor)) static void mfsgd_process_defaults() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

// ------------------------------------------------------------------------------------------------
// NCCL, loaded lazily (only a multi-process ring needs it; libmfsgd.so has no link-time dependency)
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.lib) return MFSGD_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) {
        lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(MFSGD_E_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                        \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                                \
    if (!g_nccl.field) return fail(MFSGD_E_NCCL, "libnccl lacks symbol %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = lib;
    return MFSGD_OK;
}

#define CKN(call)                                                                                          \
    do {                                                                                                   \
        ncclResult_t r__ = (call);                                                                         \
        if (r__ != ncclSuccess)                                                                            \
            return fail(MFSGD_E_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------------
struct EvalSet {            // held-out records of one ring member, bucketed per Q shard group
    Rec* recs = nullptr;
    int64_t n = 0;
    std::vector<int64_t> group_off;  // [G + 1]
};

struct Lane {                // a pair of streams that carries the update launches of an item sub-shard (lane mode, see train_impl)
    cudaStream_t cold = nullptr, hot = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, done = nullptr;
    bool used = false;       // work enqueued since the last join into the member's main stream
};

struct Member {              // one ring member ("GPU g")
    int g = 0;               // ring index
    std::vector<Lane> lanes; // created on first use
    int device = 0;          // CUDA ordinal
    int n_sms = 148;
    size_t l2_bytes = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, hot_stream = nullptr, shuffle_stream = nullptr;
    std::vector<cudaEvent_t> ev_part_done, ev_part_recv;   // pipelined rotation: per item sub-shard of the held group
    std::vector<char> part_recv_pending;
    // Ring window (one process per GPU, one node; see "ring window" below): Q[0], Q[1], BQ[0], BQ[1] and the sequence flags of
    // this member live in ONE allocation that the two ring neighbours map through CUDA IPC. The rotation then is a copy-engine
    // write into the neighbour's window over NVLink plus a flag, with no kernel and no SM involved on either side.
    struct Window {
        void* base = nullptr;
        size_t bytes = 0;
        int64_t cap_rows = 0;
        int k = 0;
        bool biases = false, active = false;
        bool can_flush = false;       // the device can flush outstanding remote writes behind a flag wait (CU_STREAM_WAIT_VALUE_FLUSH)
        size_t off_q[2] = {0, 0}, off_bq[2] = {0, 0};
        char* to_base = nullptr;      // window of member g - 1 (we write its Q buffers and its arrival flags)
        char* from_base = nullptr;    // window of member g + 1 (we write its credit flags)
        uint32_t* d_seq = nullptr;    // d_seq[j] = seq_base + j: the source of the 4-byte flag writes
        uint32_t seq_base = 0;
        int seq_n = 0;
        std::vector<uint32_t> sent;   // per item sub-shard: sequence number of the last hand-over
    } win;
    int ahead_epoch = -1;    // epoch whose layout the background reshuffle has written (is writing) into recs[rcur ^ 1]
    int32_t u_lo = 0, u_hi = 0;      // owned P rows
    float* P = nullptr;
    float* Q[2] = {nullptr, nullptr};
    float* BU = nullptr;              // model extension: biases of the owned users ...
    float* BQ[2] = {nullptr, nullptr}; // ... and of the held item group (travels with Q)
    int64_t q_cap_rows = 0;
    int cur = 0;             // Q buffer holding the current shard group
    int held_group = 0;      // which shard group that is
    Rec* recs[2] = {nullptr, nullptr};
    int rcur = 0;
    int64_t n_recs = 0;
    std::vector<int64_t> block_off;  // host copy, [mu * IB + 1]
    int64_t* d_block_off = nullptr;
    uint16_t* d_owner_u = nullptr;
    uint16_t* d_owner_i = nullptr;
    int32_t* d_hot_index = nullptr;  // item -> index into handle.hot_items, or -1
    uint32_t* d_heavy_bits = nullptr; // bit u: heavy user (rows updated with red.global.add by the run kernel)
    HotUnit* d_units = nullptr;      // hot-item units of all visits, grouped by visit
    std::vector<int> visit_units;    // [(a * rounds + rnd) * IB + item block] -> first unit; one extra entry at the end
    std::vector<int64_t> unit_recs_cum;   // records of units [0, j): prefix sums over the unit list
    int run_chunk = 0;               // longest run of the plan
    unsigned int* d_counters = nullptr;   // one unit-claim counter per hot launch; COUNTER_EPOCHS epochs' worth, zeroed per batch
    int n_counters = 0, counter_next = 0;
    cudaEvent_t ev_epoch_go = nullptr;
    int hot_grid = 148;
    EvalSet heldout;
    double* d_scratch = nullptr;     // rmse partials + 1 accumulator at the end
    double* d_sse = nullptr;
    // deterministic mode
    Rec* recs_orig = nullptr;
    uint64_t* keys_a = nullptr;
    uint64_t* keys_b = nullptr;
    void* sort_temp = nullptr;
    size_t sort_temp_bytes = 0;
    // events
    cudaEvent_t ev_compute = nullptr, ev_sent = nullptr, ev_fork = nullptr, ev_join = nullptr, ev_shuffle_go = nullptr, ev_shuffle_done = nullptr;
    std::vector<cudaEvent_t> evpool; // timing events, handed out per train call
    int ev_used = 0;
    struct EpochRec {                // one per epoch of the running train call, resolved after a sync
        cudaEvent_t start, shuffled, end;
        int kbeg, kend;              // range in kev of (fork, join, cold_end, hot_end, exchange_end) 5-tuples, one per sub-epoch
        int launches, update_launches;
    };
    std::vector<EpochRec> pending;
    std::vector<cudaEvent_t> kev;
    int grid = 148;
    int launches = 0, update_launches = 0;
};

struct mfsgd_handle {
    mfsgd_config cfg;
    int G = 1, mu = 1, mi = 1, UB = 1, IB = 1;
    int rounds = 1;          // interleaved passes over the P sub-stripes per sub-epoch
    bool virtual_shuffle = false;   // update kernels read records through the per-epoch permutation (no reshuffle pass)
    std::vector<int32_t> hot_items;     // global ids of the hot items, ascending (so grouped by item block)
    std::vector<int32_t> hot_block_lo;  // [IB + 1]: hot_items[hot_block_lo[b] .. hot_block_lo[b+1]) lie in item block b
    int H = 0;
    std::vector<uint32_t> heavy_bits;   // bit u set: heavy user (empty: none)
    int n_heavy = 0;
    float center = 0.f;      // model extension: global mean, subtracted from every rating at load (0 when off)
    bool biases = false;
    bool p_half = false;     // mixed-precision factor storage: P rows kept as binary16 (mfsgd_config.p_storage)
    float scale = 0.f;
    std::vector<int32_t> user_bounds, item_bounds;  // [UB + 1], [IB + 1]
    std::vector<Member> members;                    // the ring members this process drives
    bool loaded = false, factors_ready = false;
    int epoch = 0;
    int eval_every = 0;
    // model extension: learning-rate schedule and early stopping
    float lr_decay = 1.f;    // lr of epoch e+1 = lr of epoch e * lr_decay, one binary32 multiply (MatrixFactorizationSGD.java learningRate)
    float lr_now = 0.f;      // lr of the next epoch to run
    int es_patience = 0;     // 0 = off
    float es_min_delta = 0.f;
    double es_best = std::numeric_limits<double>::infinity();
    int es_bad = 0;
    bool es_stopped = false; // the last mfsgd_train call ended on the early-stopping rule
    double* d_allreduce = nullptr;   // 2 doubles on member 0's device: (sse, n) summed over a multi-process ring
    int64_t n_train_total = 0;
    ncclComm_t comm = nullptr;
    bool multi_process = false;
    int min_windows = 128;   // MFSGD_MIN_WINDOWS overrides (tuning aid)
    int ring_transport = 1;  // pipelined rotation of a multi-process ring: 1 = ring window (peer writes + flags), 0 = ncclSend/ncclRecv
                             // (MFSGD_RING_TRANSPORT = window | nccl; falls back to NCCL when the window cannot be mapped)
    int ring_wait_kernel = 0;   // MFSGD_RING_WAIT = kernel: poll the flags with a one-thread kernel instead of cuStreamWaitValue32
    int ring_signal_copy = 0;   // MFSGD_RING_SIGNAL = copy: set the flags with a 4-byte copy instead of cuStreamWriteValue32 (experiment only)
    int reserve_sms = 0;     // multi-process ring: SMs the run kernel leaves to the rotation's NCCL kernels (MFSGD_RESERVE_SMS)
    int n_lanes = 2;         // stream lanes of the pipelined rotation (MFSGD_LANES = 1 | 2)
    int sub_warp_div = 4;    // run kernel: at most (users of a P sub-stripe) / sub_warp_div runs in flight; MFSGD_SUBWARP_DIV overrides
};

static inline int group_lo(const mfsgd_handle* h, int grp) { return h->item_bounds[(size_t)grp * h->mi]; }
static inline int group_hi(const mfsgd_handle* h, int grp) { return h->item_bounds[(size_t)(grp + 1) * h->mi]; }

// ------------------------------------------------------------------------------------------------
// device memory: a process-wide cache of freed blocks
// ------------------------------------------------------------------------------------------------
// cudaFree of multi-gigabyte buffers synchronises the device and unmaps the pages: measured 0.06-0.9 s inside mfsgd_destroy of
// a Netflix-shaped handle (round-2 e2e trace: the "first-call stall" of round 1), and cudaMalloc pays again on the next
// handle. A resident caller (the JVM) creates and destroys handles repeatedly, so freed blocks are kept per device and
// handed out again (first block of >= the requested size and <= 1.25 x + 1 MB); on an allocation failure the cache is
// emptied and the allocation retried. mfsgd_release_cached_memory() returns everything to the driver.
struct DeviceBlockCache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void*> free_blocks;      // (device, bytes) -> block
    std::unordered_map<void*, std::pair<int, size_t>> live;        // blocks handed out
    size_t cached_bytes = 0;

    cudaError_t alloc(void** out, size_t bytes) {
        *out = nullptr;
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        bytes = (bytes + 511) & ~(size_t)511;
        {
            std::lock_guard<std::mutex> lock(mu);
            auto it = free_blocks.lower_bound({dev, bytes});
            if (it != free_blocks.end() && it->first.first == dev && it->first.second <= bytes + bytes / 4 + ((size_t)1 << 20)) {
                *out = it->second;
                live[*out] = it->first;
                cached_bytes -= it->first.second;
                free_blocks.erase(it);
                return cudaSuccess;
            }
        }
        e = cudaMalloc(out, bytes);
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            release_all();
            e = cudaMalloc(out, bytes);
        }
        if (e == cudaSuccess) {
            std::lock_guard<std::mutex> lock(mu);
            live[*out] = {dev, bytes};
        }
        return e;
    }
    void free(void* p) {
        if (!p) return;
        std::lock_guard<std::mutex> lock(mu);
        auto it = live.find(p);
        if (it == live.end()) {      // not ours (never happens): hand it to the driver
            cudaFree(p);
            return;
        }
        free_blocks.insert({it->second, p});
        cached_bytes += it->second.second;
        live.erase(it);
    }
    void release_all() {
        std::lock_guard<std::mutex> lock(mu);
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto& kv : free_blocks) {
            cudaSetDevice(kv.first.first);
            cudaFree(kv.second);
        }
        cudaSetDevice(cur);
        free_blocks.clear();
        cached_bytes = 0;
    }
};
static DeviceBlockCache g_cache;
static inline bool cache_enabled() {
    static int on = -1;
    if (on < 0) on = (getenv("MFSGD_NO_CACHE") == nullptr) ? 1 : 0;
    return on == 1;
}
static cudaError_t raw_alloc(void** p, size_t bytes) {
    if (bytes == 0) bytes = 1;
    return cache_enabled() ? g_cache.alloc(p, bytes) : cudaMalloc(p, bytes);
}
static void raw_free(void* p) {
    if (!p) return;
    if (cache_enabled()) g_cache.free(p);
    else cudaFree(p);
}

template <typename T>
static cudaError_t dev_alloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    return raw_alloc((void**)p, count * sizeof(T));
}
template <typename T>
static void dev_free(T*& p) {
    if (p) raw_free(p);
    p = nullptr;
}

extern "C" int mfsgd_release_cached_memory(void) {
    g_cache.release_all();
    return MFSGD_OK;
}

static void free_eval(EvalSet& e) {
    dev_free(e.recs);
    e.n = 0;
    e.group_off.clear();
}

// Q and the item biases either live in the member's ring window (which outlives a reload) or are blocks of their own
static void release_q(Member& m) {
    if (m.win.active) {
        m.Q[0] = m.Q[1] = m.BQ[0] = m.BQ[1] = nullptr;
        return;
    }
    dev_free(m.Q[0]);
    dev_free(m.Q[1]);
    dev_free(m.BQ[0]);
    dev_free(m.BQ[1]);
}

static void free_member_data(Member& m) {
    cudaSetDevice(m.device);
    if (m.shuffle_stream) cudaStreamSynchronize(m.shuffle_stream);
    m.ahead_epoch = -1;
    dev_free(m.P);
    release_q(m);
    dev_free(m.BU);
    dev_free(m.recs[0]);
    dev_free(m.recs[1]);
    dev_free(m.d_block_off);
    dev_free(m.d_owner_u);
    dev_free(m.d_owner_i);
    dev_free(m.d_hot_index);
    dev_free(m.d_heavy_bits);
    dev_free(m.d_units);
    dev_free(m.d_counters);
    m.visit_units.clear();
    dev_free(m.recs_orig);
    dev_free(m.keys_a);
    dev_free(m.keys_b);
    dev_free(m.sort_temp);
    free_eval(m.heldout);
    m.n_recs = 0;
}

// ------------------------------------------------------------------------------------------------
// record sources: device SoA chunks of the input (host arrays or the synthetic generator)
// ------------------------------------------------------------------------------------------------
struct Source {
    const int32_t* hu = nullptr;   // host arrays (training set, or explicit held-out set)
    const int32_t* hi = nullptr;
    const float* hr = nullptr;
    bool synthetic = false;
    SynthArgs synth{};
    int64_t total = 0;
    bool all_held = false;         // host arrays are a held-out set
    bool sharded = false;          // multi-process ring: the host arrays are THIS rank's slice of the training set
    const Rec* drecs = nullptr;    // device records (what the other members sent this one), instead of host arrays
};

struct Chunk {                     // per-device staging buffers
    int32_t* u = nullptr;
    int32_t* i = nullptr;
    float* r = nullptr;
    uint8_t* held = nullptr;       // only for synthetic sources
    int64_t cap = 0;
    int64_t start = -1, count = 0; // what is currently staged
};

static const int COUNTER_EPOCHS = 64;          // epochs whose run-claim counters are zeroed by one memset (lanes run across epoch boundaries)
static const int64_t CHUNK_MAX = 256LL << 20;  // records per staging chunk (3.3 GB of device buffers)

static void chunk_free(Chunk& c) {
    dev_free(c.u);
    dev_free(c.i);
    dev_free(c.r);
    dev_free(c.held);
    c.start = -1;
}
static int chunk_alloc(Chunk& c, int64_t cap, bool with_held) {
    c.cap = cap;
    cudaError_t e = dev_alloc(&c.u, (size_t)cap);
    if (e == cudaSuccess) e = dev_alloc(&c.i, (size_t)cap);
    if (e == cudaSuccess) e = dev_alloc(&c.r, (size_t)cap);
    if (e == cudaSuccess && with_held) e = dev_alloc(&c.held, (size_t)cap);
    if (e != cudaSuccess) chunk_free(c);        // nothing half-allocated is left behind
    CK(e);
    return MFSGD_OK;
}
// Stage records [start, start+count) on the current device (no-op when already staged).
static int chunk_stage(Chunk& c, const Source& s, int64_t start, int64_t count, cudaStream_t stream, int* launches) {
    if (c.start == start && c.count == count) return MFSGD_OK;
    if (s.synthetic) {
        CK(launch_generate(s.synth, start, count, c.u, c.i, c.r, c.held, stream, launches));
    } else if (s.drecs) {
        CK(launch_unpack_records(s.drecs + start, count, c.u, c.i, c.r, stream, launches));
    } else {
        CK(cudaMemcpyAsync(c.u, s.hu + start, (size_t)count * 4, cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(c.i, s.hi + start, (size_t)count * 4, cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(c.r, s.hr + start, (size_t)count * 4, cudaMemcpyHostToDevice, stream));
    }
    c.start = start;
    c.count = count;
    return MFSGD_OK;
}

// ------------------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------------------
extern "C" int mfsgd_abi_version(void) { return MFSGD_ABI_VERSION; }
extern "C" const char* mfsgd_last_error(void) { return g_err; }

extern "C" int mfsgd_device_count(int32_t* count) {
    if (!count) return fail(MFSGD_E_INVALID_ARG, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(MFSGD_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return MFSGD_OK;
}

extern "C" int mfsgd_config_default(mfsgd_config* cfg) {
    if (!cfg) return fail(MFSGD_E_INVALID_ARG, "cfg is null");
    memset(cfg, 0, sizeof(*cfg));
    cfg->k = 128;
    cfg->lr = 0.005f;
    cfg->lambda = 0.05f;
    cfg->seed = 20261018ULL;
    cfg->mode = MFSGD_MODE_HOGWILD;
    cfg->n_gpus = 1;
    cfg->world_size = 1;
    return MFSGD_OK;
}

static int validate_config(const mfsgd_config* c) {
    if (!c) return fail(MFSGD_E_INVALID_ARG, "cfg is null");
    if (c->n_users <= 0 || c->n_items <= 0) return fail(MFSGD_E_INVALID_ARG, "n_users and n_items must be positive");
    if (!rank_supported(c->k)) return fail(MFSGD_E_INVALID_ARG, "k=%d unsupported: need k %% 4 == 0 and 4 <= k <= 512", c->k);
    if (!(c->lr > 0.f) || !(c->lambda >= 0.f)) return fail(MFSGD_E_INVALID_ARG, "need lr > 0 and lambda >= 0");
    if (c->mode < MFSGD_MODE_DETERMINISTIC || c->mode > MFSGD_MODE_DSGD) return fail(MFSGD_E_INVALID_ARG, "bad mode %d", c->mode);
    if (c->n_gpus < 1 || c->n_gpus > 64) return fail(MFSGD_E_INVALID_ARG, "n_gpus=%d out of range", c->n_gpus);
    if (c->mode != MFSGD_MODE_DSGD && c->n_gpus != 1) return fail(MFSGD_E_INVALID_ARG, "DETERMINISTIC/HOGWILD modes need n_gpus == 1");
    if (c->scatter < MFSGD_SCATTER_STORE || c->scatter > 6) return fail(MFSGD_E_INVALID_ARG, "bad scatter %d", c->scatter);
    if (c->stripes_per_gpu < 0 || c->stripes_per_gpu > 256 || c->shards_per_gpu < 0 || c->shards_per_gpu > 64)
        return fail(MFSGD_E_INVALID_ARG, "stripes_per_gpu/shards_per_gpu out of range");
    if (c->world_size != 1 && c->world_size != c->n_gpus) return fail(MFSGD_E_INVALID_ARG, "world_size must be 1 or n_gpus");
    if (c->world_size > 1 && (c->rank < 0 || c->rank >= c->world_size)) return fail(MFSGD_E_INVALID_ARG, "bad rank %d", c->rank);
    if (c->world_size > 1 && (c->flags & MFSGD_FLAG_VIRTUAL_RING)) return fail(MFSGD_E_INVALID_ARG, "virtual ring is single-process only");
    if (c->mode == MFSGD_MODE_DETERMINISTIC && (c->stripes_per_gpu > 1 || c->shards_per_gpu > 1))
        return fail(MFSGD_E_INVALID_ARG, "DETERMINISTIC mode keeps the caller's record order: no blocking");
    if (c->device < 0) return fail(MFSGD_E_INVALID_ARG, "bad device %d", c->device);
    if (c->hot_chunk < 0 || c->hot_chunk > 65536) return fail(MFSGD_E_INVALID_ARG, "hot_chunk=%d out of range (0..65536)", c->hot_chunk);
    if (c->model & ~(MFSGD_MODEL_GLOBAL_MEAN | MFSGD_MODEL_BIASES)) return fail(MFSGD_E_INVALID_ARG, "unknown model bits 0x%x", c->model);
    if (c->p_storage != MFSGD_STORAGE_F32 && c->p_storage != MFSGD_STORAGE_F16) return fail(MFSGD_E_INVALID_ARG, "bad p_storage %d", c->p_storage);
    if (c->p_storage == MFSGD_STORAGE_F16 && c->scatter != MFSGD_SCATTER_STORE)
        return fail(MFSGD_E_INVALID_ARG, "binary16 P storage needs scatter = MFSGD_SCATTER_STORE");
    if (c->p_storage == MFSGD_STORAGE_F16 && c->mode != MFSGD_MODE_DETERMINISTIC && (c->flags & MFSGD_FLAG_EXACT_ARITH))
        return fail(MFSGD_E_INVALID_ARG, "binary16 P storage runs the FMA arrangement (MFSGD_FLAG_EXACT_ARITH: DETERMINISTIC mode only)");
    if (!(c->lr_decay >= 0.f) || c->lr_decay > 1.f) return fail(MFSGD_E_INVALID_ARG, "lr_decay must be 0 (constant rate) or in (0, 1]");
    if (c->early_stop_patience < 0) return fail(MFSGD_E_INVALID_ARG, "early_stop_patience < 0");
    if (!(c->early_stop_min_delta >= 0.f) || c->early_stop_min_delta >= 1.f) return fail(MFSGD_E_INVALID_ARG, "early_stop_min_delta must be in [0, 1)");
    if (c->p_atomic_threshold != c->p_atomic_threshold) return fail(MFSGD_E_INVALID_ARG, "p_atomic_threshold is NaN");
    if (!(c->merge_boost >= 0.f) || c->merge_boost >= 2.f) return fail(MFSGD_E_INVALID_ARG, "merge_boost must be 0 (default) or in [1, 2)");
    if (c->merge_boost > 0.f && c->merge_boost < 1.f) return fail(MFSGD_E_INVALID_ARG, "merge_boost must be 0 (default) or in [1, 2)");
    if (c->rounds < 0 || c->rounds > 256) return fail(MFSGD_E_INVALID_ARG, "rounds=%d out of range", c->rounds);
    return MFSGD_OK;
}

// ------------------------------------------------------------------------------------------------
// ring window: the Q rotation of a one-process-per-GPU ring over NVLink peer memory
// ------------------------------------------------------------------------------------------------
// Round 2 moved the pipelined rotation off ncclSend/ncclRecv: their kernels need SM slots, and the run kernel's CTAs are
// persistent, so with two stream lanes keeping the machine full a slice only left when the OTHER lane's launch drained -- the
// pipeline serialised and 8 GPUs ran at 3.9x one (profiles/r02_experiments.md section 11). Here every member keeps both Q
// buffers, the item biases and a page of flags in one allocation, its ring neighbours map that allocation (CUDA IPC; NVLink /
// NVSwitch peer access), and a hand-over of item sub-shard `part` is, on the sender's copy stream:
//     wait   credit[part]  >= n - 1      (own window: the receiver's own send out of the destination buffer is done)
//     copy   slice  -> the receiver's other Q buffer            (copy engine, peer write)
//     write  arrival[part] = n           into the receiver's window (after the slice: same stream, same engine order)
//     write  credit[part]  = n           into the window of the member that sends to us (our buffer is free again)
// and on the receiver's lane:  wait arrival[part] >= n  before the sub-shard's next launches. n counts the hand-overs of a
// part since the window was built; it is the same number on every rank. Waits are stream memory operations
// (cuStreamWaitValue32, with the remote-write flush where the device offers it; MFSGD_RING_WAIT=kernel polls with one thread
// instead), flag writes are cuStreamWriteValue32 with its default memory fence, which orders the flag behind the slice the
// same stream copied before it. (MFSGD_RING_SIGNAL=copy writes the flag as a 4-byte copy out of a table of sequence numbers
// instead: no fence between two copies -- the one bit-exactness failure seen on 8 GPUs was measured with that variant.)
// No SM, no kernel, no host thread on the data path.
static const int RING_FLAG_STRIDE = 128;      // one flag per 128-byte line
static const int RING_MAX_PARTS = 64;         // = the largest shards_per_gpu validate_config accepts
static const size_t RING_FLAGS_BYTES = (size_t)2 * RING_MAX_PARTS * RING_FLAG_STRIDE;
static inline size_t ring_arrival_off(int part) { return (size_t)part * RING_FLAG_STRIDE; }
static inline size_t ring_credit_off(int part) { return (size_t)(RING_MAX_PARTS + part) * RING_FLAG_STRIDE; }

typedef CUresult (*StreamValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamValue32Fn g_wait_value32 = nullptr, g_write_value32 = nullptr;
static void stream_mem_ops_load() {
    static bool tried = false;
    if (tried) return;
    tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
        g_wait_value32 = (StreamValue32Fn)fn;
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
        g_write_value32 = (StreamValue32Fn)fn;
    cudaGetLastError();
}

// stream waits until the flag (own window) has reached `value`
static int ring_wait(mfsgd_handle* h, Member& m, cudaStream_t stream, size_t flag_off, uint32_t value) {
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(static_cast<char*>(m.win.base) + flag_off);
    if (!h->ring_wait_kernel && g_wait_value32) {
        CUresult r = g_wait_value32((CUstream)stream, (CUdeviceptr)(uintptr_t)flag, value,
                                    CU_STREAM_WAIT_VALUE_GEQ | (m.win.can_flush ? CU_STREAM_WAIT_VALUE_FLUSH : 0u));
        if (r != CUDA_SUCCESS) return fail(MFSGD_E_CUDA, "cuStreamWaitValue32 failed (%d); set MFSGD_RING_WAIT=kernel or MFSGD_RING_TRANSPORT=nccl", (int)r);
        return MFSGD_OK;
    }
    CK(launch_ring_wait_flag(flag, value, stream, &m.launches));
    return MFSGD_OK;
}

// stream sets a flag in a neighbour's window to `value`, after everything enqueued on it so far
static int ring_signal(mfsgd_handle* h, Member& m, cudaStream_t stream, char* peer_base, size_t flag_off, uint32_t value) {
    if (!h->ring_signal_copy) {
        // default flags: the write is preceded by a system-scope memory fence over everything the stream has done before it, so
        // the slice (copied by the operation before this one) is in the neighbour's memory when the flag becomes visible there.
        CUresult r = g_write_value32((CUstream)stream, (CUdeviceptr)(uintptr_t)(peer_base + flag_off), value, CU_STREAM_WRITE_VALUE_DEFAULT);
        if (r != CUDA_SUCCESS) return fail(MFSGD_E_CUDA, "cuStreamWriteValue32 failed (%d); set MFSGD_RING_TRANSPORT=nccl", (int)r);
        return MFSGD_OK;
    }
    if (value < m.win.seq_base || value >= m.win.seq_base + (uint32_t)m.win.seq_n) return fail(MFSGD_E_STATE, "ring window: sequence table does not cover %u", value);
    CK(cudaMemcpyAsync(peer_base + flag_off, m.win.d_seq + (value - m.win.seq_base), 4, cudaMemcpyDefault, stream));
    return MFSGD_OK;
}

// the table of sequence numbers covers [first, first + count) (grown between train calls, never while the copy stream reads it)
static int ring_seq_table(Member& m, uint32_t first, uint32_t count) {
    Member::Window& w = m.win;
    if (w.d_seq && first >= w.seq_base && first + count <= w.seq_base + (uint32_t)w.seq_n) return MFSGD_OK;
    CK(cudaStreamSynchronize(m.copy_stream));
    const uint32_t n = std::max<uint32_t>(count, 1u << 16);
    if ((int)n > w.seq_n) {
        if (w.d_seq) cudaFree(w.d_seq);
        w.d_seq = nullptr;
        w.seq_n = 0;
        CK(cudaMalloc((void**)&w.d_seq, (size_t)n * 4));
        w.seq_n = (int)n;
    }
    std::vector<uint32_t> host((size_t)w.seq_n);
    for (int j = 0; j < w.seq_n; j++) host[(size_t)j] = first + (uint32_t)j;
    CK(cudaMemcpy(w.d_seq, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
    w.seq_base = first;
    return MFSGD_OK;
}

static void window_close_peers(Member& m) {
    Member::Window& w = m.win;
    if (w.to_base) cudaIpcCloseMemHandle(w.to_base);
    if (w.from_base && w.from_base != w.to_base) cudaIpcCloseMemHandle(w.from_base);
    w.to_base = w.from_base = nullptr;
    cudaGetLastError();
}

// destroy: no barrier here -- every write into this window has landed before the last train call returned (it waits for
// its arrivals AND its credits), and a neighbour that still maps the window only ever writes to it inside a train call
static void window_release(Member& m) {
    window_close_peers(m);
    if (m.win.active) m.Q[0] = m.Q[1] = m.BQ[0] = m.BQ[1] = nullptr;
    if (m.win.base) cudaFree(m.win.base);
    if (m.win.d_seq) cudaFree(m.win.d_seq);
    m.win = Member::Window();
}

// all ranks agree on a flag: min over the ring (also a barrier)
static int ring_all_min(mfsgd_handle* h, Member& m, int* flag) {
    uint32_t* d = nullptr;
    CK(dev_alloc(&d, 1));
    uint32_t v = (uint32_t)(*flag != 0);
    cudaError_t e = cudaMemcpyAsync(d, &v, 4, cudaMemcpyHostToDevice, m.stream);
    ncclResult_t r = ncclSuccess;
    if (e == cudaSuccess) r = g_nccl.AllReduce(d, d, 1, ncclUint32, ncclMin, h->comm, m.stream);
    if (e == cudaSuccess && r == ncclSuccess) e = cudaMemcpyAsync(&v, d, 4, cudaMemcpyDeviceToHost, m.stream);
    if (e == cudaSuccess && r == ncclSuccess) e = cudaStreamSynchronize(m.stream);
    dev_free(d);
    if (r != ncclSuccess) return fail(MFSGD_E_NCCL, "ring window agreement: %s", g_nccl.GetErrorString(r));
    CK(e);
    *flag = (int)v;
    return MFSGD_OK;
}

// Build (or keep) the member's window for `cap` rows per Q buffer and map the neighbours'. Collective over the ring: cap, k
// and the model are the same on every rank, so every rank takes the same branch. On success m.Q / m.BQ point into the
// window; when the window cannot be set up on ANY rank, every rank falls back to separate buffers + NCCL (win.active false).
static int window_setup(mfsgd_handle* h, Member& m, int64_t cap) {
    Member::Window& w = m.win;
    const int k = h->cfg.k, G = h->G;
    const bool want = h->multi_process && G > 1 && h->ring_transport == 1 && h->cfg.mode != MFSGD_MODE_DETERMINISTIC;
    if (!want) {
        if (w.base) window_release(m);
        return MFSGD_OK;
    }
    auto point_into = [&]() {
        char* b = static_cast<char*>(w.base);
        m.Q[0] = reinterpret_cast<float*>(b + w.off_q[0]);
        m.Q[1] = reinterpret_cast<float*>(b + w.off_q[1]);
        m.BQ[0] = h->biases ? reinterpret_cast<float*>(b + w.off_bq[0]) : nullptr;
        m.BQ[1] = h->biases ? reinterpret_cast<float*>(b + w.off_bq[1]) : nullptr;
    };
    if (w.active && w.cap_rows >= cap && w.k == k && w.biases == h->biases) {     // reload with the same shape: keep everything
        point_into();
        return MFSGD_OK;
    }
    stream_mem_ops_load();
    int ok = 1;
    if (w.base) {                        // rebuild: nobody may still map the old window when it is freed
        window_close_peers(m);
        CKRC(ring_all_min(h, m, &ok));
        cudaFree(w.base);
        w.base = nullptr;                // (the table of sequence numbers stays)
        w.active = false;
        w.sent.clear();
        m.Q[0] = m.Q[1] = m.BQ[0] = m.BQ[1] = nullptr;
    }
    auto up = [](size_t v) { return (v + 511) & ~(size_t)511; };
    w.cap_rows = cap;
    w.k = k;
    w.biases = h->biases;
    w.off_q[0] = up(RING_FLAGS_BYTES);
    w.off_q[1] = w.off_q[0] + up((size_t)cap * k * 4);
    w.off_bq[0] = w.off_q[1] + up((size_t)cap * k * 4);
    w.off_bq[1] = w.off_bq[0] + up(h->biases ? (size_t)cap * 4 : 0);
    w.bytes = w.off_bq[1] + up(h->biases ? (size_t)cap * 4 : 0);
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    ok = ok && (h->ring_wait_kernel || g_wait_value32 != nullptr) && g_write_value32 != nullptr;
    {
        int flush = 0;
        if (cudaDeviceGetAttribute(&flush, cudaDevAttrCanFlushRemoteWrites, m.device) != cudaSuccess) cudaGetLastError();
        w.can_flush = flush != 0;
    }
    if (cudaMalloc(&w.base, w.bytes) != cudaSuccess) {      // a plain allocation of its own: an IPC handle names a whole cudaMalloc block
        cudaGetLastError();
        w.base = nullptr;
        ok = 0;
    }
    if (w.base && (cudaMemsetAsync(w.base, 0, RING_FLAGS_BYTES, m.stream) != cudaSuccess || cudaStreamSynchronize(m.stream) != cudaSuccess ||
                   cudaIpcGetMemHandle(&mine, w.base) != cudaSuccess)) {
        cudaGetLastError();
        ok = 0;
    }
    // every rank's handle to every rank (in place all-gather of 64 bytes per rank)
    std::vector<cudaIpcMemHandle_t> all((size_t)G);
    {
        uint8_t* d = nullptr;
        CK(dev_alloc(&d, (size_t)G * 64));
        cudaError_t e = cudaMemcpyAsync(d + (size_t)m.g * 64, &mine, 64, cudaMemcpyHostToDevice, m.stream);
        ncclResult_t r = ncclSuccess;
        if (e == cudaSuccess) r = g_nccl.AllGather(d + (size_t)m.g * 64, d, 64, ncclUint8, h->comm, m.stream);
        if (e == cudaSuccess && r == ncclSuccess) e = cudaMemcpyAsync(all.data(), d, (size_t)G * 64, cudaMemcpyDeviceToHost, m.stream);
        if (e == cudaSuccess && r == ncclSuccess) e = cudaStreamSynchronize(m.stream);
        dev_free(d);
        if (r != ncclSuccess) return fail(MFSGD_E_NCCL, "ring window handles: %s", g_nccl.GetErrorString(r));
        CK(e);
    }
    const int to = (m.g - 1 + G) % G, from = (m.g + 1) % G;
    if (ok) {
        void* pt = nullptr;
        if (cudaIpcOpenMemHandle(&pt, all[(size_t)to], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
        }
        w.to_base = static_cast<char*>(pt);
        if (ok && from == to) w.from_base = w.to_base;
        else if (ok) {
            void* pf = nullptr;
            if (cudaIpcOpenMemHandle(&pf, all[(size_t)from], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
            }
            w.from_base = static_cast<char*>(pf);
        }
    }
    CKRC(ring_all_min(h, m, &ok));       // also: every flag page is zeroed before anybody writes to one
    if (!ok) {                           // some rank could not: everybody uses separate buffers + NCCL from here on
        if (trace_on()) fprintf(stderr, "[mfsgd] ring window unavailable, rotation falls back to ncclSend/ncclRecv\n");
        window_release(m);
        h->ring_transport = 0;
        return MFSGD_OK;
    }
    w.active = true;
    w.sent.assign((size_t)RING_MAX_PARTS, 0u);
    point_into();
    return MFSGD_OK;
}

// Run kernel: p_u updated in memory by red.global.add (MFSGD_SCATTER_ATOMIC_P / MFSGD_SCATTER_ATOMIC) or stored (last writer wins)
static inline bool run_p_red(const mfsgd_config& c) { return c.scatter == MFSGD_SCATTER_ATOMIC_P || c.scatter == MFSGD_SCATTER_ATOMIC; }

static int member_setup(mfsgd_handle* h, Member& m) {
    CK(cudaSetDevice(m.device));
    int cc_major = 0, cc_minor = 0, n_sms = 0, l2 = 0;     // three attributes instead of cudaGetDeviceProperties (milliseconds per call)
    CK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, m.device));
    CK(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, m.device));
    CK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, m.device));
    CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, m.device));
    if (cc_major < 10) return fail(MFSGD_E_CUDA, "device %d is sm_%d%d; libmfsgd.so carries sm_100a code only", m.device, cc_major, cc_minor);
    m.n_sms = n_sms;
    m.l2_bytes = (size_t)l2;
    CK(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&m.copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&m.hot_stream, cudaStreamNonBlocking));
    {
        int lo_prio = 0, hi_prio = 0;   // the background reshuffle yields to the update kernels
        CK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        CK(cudaStreamCreateWithPriority(&m.shuffle_stream, cudaStreamNonBlocking, lo_prio));
    }
    CK(cudaEventCreateWithFlags(&m.ev_compute, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m.ev_sent, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m.ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m.ev_join, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m.ev_shuffle_go, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m.ev_shuffle_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m.ev_epoch_go, cudaEventDisableTiming));
    CK(dev_alloc(&m.d_scratch, (size_t)rmse_scratch_doubles() + 1));
    m.d_sse = m.d_scratch + rmse_scratch_doubles();
    int ctas = 0;
    const bool fast = !(h->cfg.flags & MFSGD_FLAG_EXACT_ARITH);
    const bool p_half = h->cfg.p_storage == MFSGD_STORAGE_F16;
    CK(hogwild_max_ctas_per_sm(h->cfg.k, h->cfg.scatter, fast, p_half, &ctas));
    if (ctas < 1) ctas = 1;
    if (h->cfg.ctas_per_sm > 0 && h->cfg.ctas_per_sm < ctas) ctas = h->cfg.ctas_per_sm;
    m.grid = m.n_sms * ctas;
    int hot_ctas = 0;
    CK(hot_max_ctas_per_sm(h->cfg.k, fast, run_p_red(h->cfg), p_half, &hot_ctas));
    m.hot_grid = m.n_sms * std::max(1, hot_ctas);
    return MFSGD_OK;
}

extern "C" void mfsgd_destroy(mfsgd_handle* h) {
    if (!h) return;
    for (Member& m : h->members) {
        cudaSetDevice(m.device);
        if (m.stream) cudaStreamSynchronize(m.stream);
        if (m.copy_stream) cudaStreamSynchronize(m.copy_stream);
        if (m.hot_stream) cudaStreamSynchronize(m.hot_stream);
        if (m.shuffle_stream) cudaStreamSynchronize(m.shuffle_stream);
        for (Lane& l : m.lanes) {
            if (l.cold) cudaStreamSynchronize(l.cold);
            if (l.hot) cudaStreamSynchronize(l.hot);
        }
    }
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    if (h->d_allreduce && !h->members.empty()) {
        cudaSetDevice(h->members[0].device);
        dev_free(h->d_allreduce);
    }
    for (Member& m : h->members) {
        cudaSetDevice(m.device);
        window_release(m);
        free_member_data(m);
        dev_free(m.d_scratch);
        for (cudaEvent_t e : m.evpool) cudaEventDestroy(e);
        for (cudaEvent_t e : m.ev_part_done) cudaEventDestroy(e);
        for (cudaEvent_t e : m.ev_part_recv) cudaEventDestroy(e);
        for (Lane& l : m.lanes) {
            cudaEvent_t levs[] = {l.fork, l.join, l.done};
            for (cudaEvent_t e : levs)
                if (e) cudaEventDestroy(e);
            if (l.cold) cudaStreamDestroy(l.cold);
            if (l.hot) cudaStreamDestroy(l.hot);
        }
        cudaEvent_t evs[] = {m.ev_compute, m.ev_sent, m.ev_fork, m.ev_join, m.ev_shuffle_go, m.ev_shuffle_done, m.ev_epoch_go};
        for (cudaEvent_t e : evs)
            if (e) cudaEventDestroy(e);
        if (m.stream) cudaStreamDestroy(m.stream);
        if (m.copy_stream) cudaStreamDestroy(m.copy_stream);
        if (m.hot_stream) cudaStreamDestroy(m.hot_stream);
        if (m.shuffle_stream) cudaStreamDestroy(m.shuffle_stream);
    }
    delete h;
}

static int mfsgd_create_body(const mfsgd_config* cfg, mfsgd_handle** out);
extern "C" int mfsgd_create(const mfsgd_config* cfg, mfsgd_handle** out) {
    return guarded([&]() { return mfsgd_create_body(cfg, out); });
}
static int mfsgd_create_body(const mfsgd_config* cfg, mfsgd_handle** out) {
    if (!out) return fail(MFSGD_E_INVALID_ARG, "out is null");
    *out = nullptr;
    CKRC(validate_config(cfg));
    int ndev = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            return fail(MFSGD_E_CUDA, "no CUDA device: %s (libmfsgd.so has no CPU fallback)",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    mfsgd_handle* h = new mfsgd_handle();
    h->cfg = *cfg;
    h->G = cfg->n_gpus;
    h->multi_process = cfg->world_size > 1;
    if (const char* mw = getenv("MFSGD_MIN_WINDOWS")) h->min_windows = std::max(1, atoi(mw));
    if (const char* sd = getenv("MFSGD_SUBWARP_DIV")) h->sub_warp_div = std::max(1, atoi(sd));
    if (const char* rs = getenv("MFSGD_RESERVE_SMS")) h->reserve_sms = std::max(0, std::min(32, atoi(rs)));
    if (const char* nl = getenv("MFSGD_LANES")) h->n_lanes = atoi(nl) == 1 ? 1 : 2;
    if (const char* rt = getenv("MFSGD_RING_TRANSPORT")) h->ring_transport = strcmp(rt, "nccl") == 0 ? 0 : 1;
    if (const char* rw = getenv("MFSGD_RING_WAIT")) h->ring_wait_kernel = strcmp(rw, "kernel") == 0 ? 1 : 0;
    if (const char* rs2 = getenv("MFSGD_RING_SIGNAL")) h->ring_signal_copy = strcmp(rs2, "copy") == 0 ? 1 : 0;
    h->scale = cfg->init_scale > 0.f ? cfg->init_scale : (float)(1.0 / std::sqrt((double)cfg->k));
    h->biases = (cfg->model & MFSGD_MODEL_BIASES) != 0;
    h->p_half = cfg->p_storage == MFSGD_STORAGE_F16;
    h->lr_decay = cfg->lr_decay > 0.f ? cfg->lr_decay : 1.f;
    h->lr_now = cfg->lr;
    h->es_patience = cfg->early_stop_patience;
    h->es_min_delta = cfg->early_stop_min_delta;
    const bool virtual_ring = (cfg->flags & MFSGD_FLAG_VIRTUAL_RING) != 0;
    const int n_local = h->multi_process ? 1 : h->G;
    if (!virtual_ring && !h->multi_process && cfg->device + h->G > ndev) {
        delete h;
        return fail(MFSGD_E_INVALID_ARG, "need devices %d..%d but only %d visible", cfg->device, cfg->device + h->G - 1, ndev);
    }
    if (cfg->device >= ndev) {
        delete h;
        return fail(MFSGD_E_INVALID_ARG, "device %d not visible (%d devices)", cfg->device, ndev);
    }
    h->members.resize((size_t)n_local);
    for (int j = 0; j < n_local; j++) {
        Member& m = h->members[(size_t)j];
        m.g = h->multi_process ? cfg->rank : j;
        m.device = (virtual_ring || h->multi_process) ? cfg->device : cfg->device + j;
        m.held_group = m.g;
        int rc = member_setup(h, m);
        if (rc != MFSGD_OK) {
            mfsgd_destroy(h);
            return rc;
        }
    }
    if (!h->multi_process && !virtual_ring && h->G > 1) {
        for (Member& a : h->members)
            for (Member& b : h->members)
                if (a.device != b.device) {
                    cudaSetDevice(a.device);
                    int can = 0;
                    cudaDeviceCanAccessPeer(&can, a.device, b.device);
                    if (can) {
                        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                        if (e != cudaSuccess) cudaGetLastError();  // already enabled is fine
                    }
                }
    }
    if (h->multi_process) {
        int rc = nccl_load();
        if (rc != MFSGD_OK) {
            mfsgd_destroy(h);
            return rc;
        }
        ncclUniqueId id;
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
        memcpy(&id, cfg->nccl_id, sizeof(id));
        cudaSetDevice(h->members[0].device);
        ncclResult_t r = g_nccl.CommInitRank(&h->comm, cfg->world_size, id, cfg->rank);
        if (r != ncclSuccess) {
            int rc2 = fail(MFSGD_E_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
            h->comm = nullptr;
            mfsgd_destroy(h);
            return rc2;
        }
        // NCCL connects two ranks lazily, on their first exchange (tens of milliseconds per pair): do it now, as part of
        // the ring's bootstrap, for every pair the engine will use -- ring neighbours (Q rotation) and all pairs (record
        // exchange of mfsgd_load_ratings_sharded, all-reduce of the row counts).
        {
            Member& m = h->members[0];
            uint32_t* warm = nullptr;
            cudaError_t e = dev_alloc(&warm, (size_t)2 * cfg->world_size + 2);
            ncclResult_t nr = ncclSuccess;
            if (e == cudaSuccess) e = cudaMemsetAsync(warm, 0, ((size_t)2 * cfg->world_size + 2) * 4, m.stream);
            if (e == cudaSuccess) {
                nr = g_nccl.AllReduce(warm, warm, 1, ncclUint32, ncclSum, h->comm, m.stream);
                if (nr == ncclSuccess) nr = g_nccl.GroupStart();
                for (int p = 0; p < cfg->world_size && nr == ncclSuccess; p++) {
                    if (p == cfg->rank) continue;
                    nr = g_nccl.Send(warm + 1, 1, ncclUint32, p, h->comm, m.stream);
                    if (nr == ncclSuccess) nr = g_nccl.Recv(warm + 2 + p, 1, ncclUint32, p, h->comm, m.stream);
                }
                ncclResult_t ne = g_nccl.GroupEnd();
                if (nr == ncclSuccess) nr = ne;
                if (nr == ncclSuccess) e = cudaStreamSynchronize(m.stream);
            }
            dev_free(warm);
            if (e != cudaSuccess || nr != ncclSuccess) {
                int rc2 = nr != ncclSuccess ? fail(MFSGD_E_NCCL, "ring warm-up: %s", g_nccl.GetErrorString(nr))
                                            : fail(MFSGD_E_CUDA, "ring warm-up: %s", cudaGetErrorString(e));
                mfsgd_destroy(h);
                return rc2;
            }
        }
    }
    *out = h;
    return MFSGD_OK;
}

static int mfsgd_nccl_unique_id_body(uint8_t out[128]);
extern "C" int mfsgd_nccl_unique_id(uint8_t out[128]) {
    return guarded([&]() { return mfsgd_nccl_unique_id_body(out); });
}
static int mfsgd_nccl_unique_id_body(uint8_t out[128]) {
    if (!out) return fail(MFSGD_E_INVALID_ARG, "out is null");
    CKRC(nccl_load());
    ncclUniqueId id;
    CKN(g_nccl.GetUniqueId(&id));
    memcpy(out, &id, 128);
    return MFSGD_OK;
}

extern "C" int mfsgd_host_alloc(void** out, int64_t bytes) {
    if (!out || bytes < 0) return fail(MFSGD_E_INVALID_ARG, "bad arguments");
    CK(cudaHostAlloc(out, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocDefault));
    return MFSGD_OK;
}
extern "C" int mfsgd_host_free(void* p) {
    if (p) CK(cudaFreeHost(p));
    return MFSGD_OK;
}

// ------------------------------------------------------------------------------------------------
// layout: bounds, owner tables, bucketing
// ------------------------------------------------------------------------------------------------
static void choose_blocking(mfsgd_handle* h) {
    const mfsgd_config& c = h->cfg;
    const Blocking b = plan_blocking(c.n_users, c.n_items, c.k, h->G, c.mode, c.stripes_per_gpu, c.shards_per_gpu, h->multi_process,
                                     (double)h->members[0].l2_bytes);     // run_plan.hpp
    h->mu = b.mu;
    h->mi = b.mi;
    h->UB = h->G * h->mu;
    h->IB = h->G * h->mi;
}

// Balanced (by rating count) bounds from the source, computed on member 0's device.
static int compute_bounds(mfsgd_handle* h, const Source& src, Chunk& ch) {
    Member& m = h->members[0];
    const mfsgd_config& c = h->cfg;
    CK(cudaSetDevice(m.device));
    uint32_t *ucnt = nullptr, *icnt = nullptr;
    uint64_t *ucum = nullptr, *icum = nullptr;
    int32_t *ub = nullptr, *ib = nullptr;
    int* bad = nullptr;
    unsigned long long* rsum = nullptr;
    void* temp = nullptr;
    int rc = MFSGD_OK;
    const bool want_mean = (c.model & MFSGD_MODEL_GLOBAL_MEAN) != 0;
    auto cleanup = [&]() {
        dev_free(ucnt); dev_free(icnt); dev_free(ucum); dev_free(icum); dev_free(ub); dev_free(ib); dev_free(bad); dev_free(rsum);
        if (temp) raw_free(temp);
    };
#define CKC(call)                                                                                           \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess) {                                                                           \
            rc = fail(e__ == cudaErrorMemoryAllocation ? MFSGD_E_OOM : MFSGD_E_CUDA, "%s:%d %s: %s", __FILE__, \
                      __LINE__, #call, cudaGetErrorString(e__));                                            \
            cleanup();                                                                                      \
            return rc;                                                                                      \
        }                                                                                                   \
    } while (0)
    CKC(dev_alloc(&ucnt, (size_t)c.n_users + 1));
    CKC(dev_alloc(&icnt, (size_t)c.n_items + 1));
    CKC(dev_alloc(&ucum, (size_t)c.n_users + 1));
    CKC(dev_alloc(&icum, (size_t)c.n_items + 1));
    CKC(dev_alloc(&ub, (size_t)h->UB + 1));
    CKC(dev_alloc(&ib, (size_t)h->IB + 1));
    CKC(dev_alloc(&bad, 1));
    CKC(dev_alloc(&rsum, 1));
    CKC(cudaMemsetAsync(rsum, 0, sizeof(unsigned long long), m.stream));
    CKC(cudaMemsetAsync(ucnt, 0, ((size_t)c.n_users + 1) * 4, m.stream));
    CKC(cudaMemsetAsync(icnt, 0, ((size_t)c.n_items + 1) * 4, m.stream));
    CKC(cudaMemsetAsync(bad, 0, sizeof(int), m.stream));
    for (int64_t start = 0; start < src.total; start += ch.cap) {
        const int64_t count = std::min(ch.cap, src.total - start);
        rc = chunk_stage(ch, src, start, count, m.stream, &m.launches);
        if (rc != MFSGD_OK) { cleanup(); return rc; }
        CKC(launch_count_rows(ch.u, ch.i, src.synthetic ? ch.held : nullptr, count, ucnt, icnt, c.n_users, c.n_items, bad,
                              ch.r, want_mean ? rsum : nullptr, m.stream, &m.launches));
    }
    if (src.sharded) {     // every rank counted its own slice: sum the per-row counts (and the bad-id flag) over the ring
        ncclResult_t r1 = g_nccl.AllReduce(ucnt, ucnt, (size_t)c.n_users, ncclUint32, ncclSum, h->comm, m.stream);
        ncclResult_t r2 = r1 == ncclSuccess ? g_nccl.AllReduce(icnt, icnt, (size_t)c.n_items, ncclUint32, ncclSum, h->comm, m.stream) : r1;
        ncclResult_t r3 = r2 == ncclSuccess ? g_nccl.AllReduce(bad, bad, 1, ncclInt32, ncclMax, h->comm, m.stream) : r2;
        if (r3 == ncclSuccess) r3 = g_nccl.AllReduce(rsum, rsum, 1, ncclUint64, ncclSum, h->comm, m.stream);
        if (r3 != ncclSuccess) {
            cleanup();
            return fail(MFSGD_E_NCCL, "all-reduce of the row counts: %s", g_nccl.GetErrorString(r3));
        }
        m.launches += 3;
    }
    size_t tb_u = 0, tb_i = 0;
    CKC(exclusive_cumsum_u32(ucnt, ucum, c.n_users, nullptr, &tb_u, m.stream, nullptr));
    CKC(exclusive_cumsum_u32(icnt, icum, c.n_items, nullptr, &tb_i, m.stream, nullptr));
    size_t tb = std::max(tb_u, tb_i);
    CKC(raw_alloc(&temp, tb ? tb : 1));
    CKC(exclusive_cumsum_u32(ucnt, ucum, c.n_users, temp, &tb, m.stream, &m.launches));
    CKC(launch_balanced_bounds(ucum, c.n_users, h->UB, ub, m.stream, &m.launches));
    CKC(exclusive_cumsum_u32(icnt, icum, c.n_items, temp, &tb, m.stream, &m.launches));
    CKC(launch_balanced_bounds(icum, c.n_items, h->IB, ib, m.stream, &m.launches));
    h->user_bounds.assign((size_t)h->UB + 1, 0);
    h->item_bounds.assign((size_t)h->IB + 1, 0);
    int bad_host = 0;
    uint64_t total_train = 0;
    unsigned long long rsum_host = 0;
    CKC(cudaMemcpyAsync(&rsum_host, rsum, sizeof(rsum_host), cudaMemcpyDeviceToHost, m.stream));
    CKC(cudaMemcpyAsync(h->user_bounds.data(), ub, ((size_t)h->UB + 1) * 4, cudaMemcpyDeviceToHost, m.stream));
    CKC(cudaMemcpyAsync(h->item_bounds.data(), ib, ((size_t)h->IB + 1) * 4, cudaMemcpyDeviceToHost, m.stream));
    CKC(cudaMemcpyAsync(&bad_host, bad, sizeof(int), cudaMemcpyDeviceToHost, m.stream));
    CKC(cudaMemcpyAsync(&total_train, ucum + c.n_users, sizeof(uint64_t), cudaMemcpyDeviceToHost, m.stream));
    CKC(cudaStreamSynchronize(m.stream));
#undef CKC
    // hot items: rated by at least hot_share of the training set (and often enough to fill a warp's run)
    h->hot_items.clear();
    const float share = c.hot_share == 0.f ? 1e-6f : c.hot_share;
    if (bad_host == 0 && share > 0.f && c.mode != MFSGD_MODE_DETERMINISTIC && total_train > 0) {
        std::vector<uint32_t> icnt_host((size_t)c.n_items);
        cudaError_t e = cudaMemcpy(icnt_host.data(), icnt, (size_t)c.n_items * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { cleanup(); return fail(MFSGD_E_CUDA, "copying item counts: %s", cudaGetErrorString(e)); }
        // An item's ratings are spread over the mu * G (sub-stripe, ring member) buckets; a run should hold >= ~16 of
        // them, or the q_i load + merge (1 KB per run) is not amortised and the item is better served by the cold
        // kernel. (Rounds do not enter: build_hot_units spreads a bucket over as many rounds as it has full runs for.)
        const double min_count = (double)MIN_RUN * h->mu * h->G;
        const double thr = std::max((double)share * (double)total_train, min_count);
        std::vector<std::pair<uint32_t, int32_t>> cand;
        for (int32_t it = 0; it < c.n_items; it++)
            if ((double)icnt_host[(size_t)it] >= thr) cand.push_back({icnt_host[(size_t)it], it});
        const size_t hmax = (size_t)std::max(0, MAX_BUCKETS / h->mu - h->IB - 1);
        if (cand.size() > hmax) {
            std::sort(cand.begin(), cand.end(), [](const std::pair<uint32_t, int32_t>& x, const std::pair<uint32_t, int32_t>& y) { return x.first > y.first; });
            cand.resize(hmax);
        }
        for (auto& pr : cand) h->hot_items.push_back(pr.second);
        std::sort(h->hot_items.begin(), h->hot_items.end());
    }
    h->H = (int)h->hot_items.size();
    // model extension: the training mean from the exact integer sum (MatrixFactorizationSGD.java:272 globalMean)
    h->center = (want_mean && total_train > 0) ? (float)((double)(long long)rsum_host / (double)total_train / 1048576.0) : 0.f;
    // heavy users: expected ratings in flight at once >= p_atomic_threshold. A launch of the run kernel walks one P sub-stripe
    // (1 / (G * mu) of the ratings, evenly over the rounds) with every resident sub-warp holding ~2 ratings between the
    // gather of p_u and its scatter, so user u has about cnt_u * G * mu / total * (2 * resident sub-warps) ratings in flight.
    h->heavy_bits.clear();
    h->n_heavy = 0;
    const float tau = c.p_atomic_threshold == 0.f ? 0.25f : c.p_atomic_threshold;
    if (bad_host == 0 && tau > 0.f && c.mode != MFSGD_MODE_DETERMINISTIC && total_train > 0 && h->H > 0) {
        std::vector<uint32_t> ucnt_host((size_t)c.n_users);
        cudaError_t e = cudaMemcpy(ucnt_host.data(), ucnt, (size_t)c.n_users * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { cleanup(); return fail(MFSGD_E_CUDA, "copying user counts: %s", cudaGetErrorString(e)); }
        const double in_flight = 2.0 * (double)m.hot_grid * 8.0 * (32 / run_kernel_lanes(c.k));
        const double thr = (double)tau * (double)total_train / (in_flight * (double)h->G * (double)h->mu);
        h->heavy_bits.assign(((size_t)c.n_users + 31) / 32, 0u);
        for (int32_t u = 0; u < c.n_users; u++)
            if ((double)ucnt_host[(size_t)u] >= thr) {
                h->heavy_bits[(size_t)u >> 5] |= 1u << (u & 31);
                h->n_heavy++;
            }
        if (h->n_heavy == 0) h->heavy_bits.clear();
    }
    cleanup();
    if (bad_host) return fail(MFSGD_E_INVALID_ARG, "a rating has a user or item id outside [0,n_users) x [0,n_items)");
    h->n_train_total = (int64_t)total_train;
    return MFSGD_OK;
}

static void uniform_bounds(mfsgd_handle* h) {
    h->hot_items.clear();
    h->H = 0;
    h->heavy_bits.clear();
    h->n_heavy = 0;
    h->center = 0.f;
    h->user_bounds.resize((size_t)h->UB + 1);
    h->item_bounds.resize((size_t)h->IB + 1);
    for (int b = 0; b <= h->UB; b++) h->user_bounds[(size_t)b] = (int32_t)((int64_t)h->cfg.n_users * b / h->UB);
    for (int b = 0; b <= h->IB; b++) h->item_bounds[(size_t)b] = (int32_t)((int64_t)h->cfg.n_items * b / h->IB);
}

// Allocate factors + owner tables of a member for the current bounds.
static int member_alloc_factors(mfsgd_handle* h, Member& m) {
    const mfsgd_config& c = h->cfg;
    CK(cudaSetDevice(m.device));
    m.u_lo = h->user_bounds[(size_t)m.g * h->mu];
    m.u_hi = h->user_bounds[(size_t)(m.g + 1) * h->mu];
    int64_t cap = 0;
    for (int grp = 0; grp < h->G; grp++) cap = std::max<int64_t>(cap, group_hi(h, grp) - group_lo(h, grp));
    m.q_cap_rows = cap;
    dev_free(m.P); release_q(m); dev_free(m.d_owner_u); dev_free(m.d_owner_i);
    CK(dev_alloc(&m.P, h->p_half ? ((size_t)(m.u_hi - m.u_lo) * c.k + 1) / 2 : (size_t)(m.u_hi - m.u_lo) * c.k));   // binary16 rows: half the bytes
    CKRC(window_setup(h, m, cap));           // one process per GPU: Q and the item biases live in the member's ring window
    if (!m.win.active) {
        CK(dev_alloc(&m.Q[0], (size_t)cap * c.k));
        if (h->G > 1) CK(dev_alloc(&m.Q[1], (size_t)cap * c.k));
    }
    dev_free(m.BU);
    if (h->biases) {
        CK(dev_alloc(&m.BU, (size_t)(m.u_hi - m.u_lo)));
        if (!m.win.active) {
            CK(dev_alloc(&m.BQ[0], (size_t)cap));
            if (h->G > 1) CK(dev_alloc(&m.BQ[1], (size_t)cap));
        }
    }
    m.cur = 0;
    m.held_group = m.g;
    CK(dev_alloc(&m.d_owner_u, (size_t)c.n_users));
    CK(dev_alloc(&m.d_owner_i, (size_t)c.n_items));
    int32_t* d_b = nullptr;
    CK(dev_alloc(&d_b, (size_t)std::max(h->UB, h->IB) + 1));
    cudaError_t e = cudaMemcpyAsync(d_b, h->user_bounds.data(), ((size_t)h->UB + 1) * 4, cudaMemcpyHostToDevice, m.stream);
    if (e == cudaSuccess) e = launch_fill_owner(d_b, h->UB, c.n_users, m.d_owner_u, m.stream, &m.launches);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_b, h->item_bounds.data(), ((size_t)h->IB + 1) * 4, cudaMemcpyHostToDevice, m.stream);
    if (e == cudaSuccess) e = launch_fill_owner(d_b, h->IB, c.n_items, m.d_owner_i, m.stream, &m.launches);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m.stream);
    raw_free(d_b);
    CK(e);
    // hot-item lookup table and the per-group ranges of the (ascending) hot list
    h->hot_block_lo.assign((size_t)h->IB + 1, 0);
    for (int ib = 0; ib <= h->IB; ib++) {
        const int32_t bound = h->item_bounds[(size_t)ib];
        h->hot_block_lo[(size_t)ib] = (int32_t)(std::lower_bound(h->hot_items.begin(), h->hot_items.end(), bound) - h->hot_items.begin());
    }
    dev_free(m.d_heavy_bits);
    if (!h->heavy_bits.empty()) {
        CK(dev_alloc(&m.d_heavy_bits, h->heavy_bits.size()));
        CK(cudaMemcpy(m.d_heavy_bits, h->heavy_bits.data(), h->heavy_bits.size() * 4, cudaMemcpyHostToDevice));
    }
    dev_free(m.d_hot_index);
    if (h->H > 0) {
        std::vector<int32_t> idx((size_t)c.n_items, -1);
        for (int j = 0; j < h->H; j++) idx[(size_t)h->hot_items[(size_t)j]] = j;
        CK(dev_alloc(&m.d_hot_index, (size_t)c.n_items));
        CK(cudaMemcpy(m.d_hot_index, idx.data(), (size_t)c.n_items * 4, cudaMemcpyHostToDevice));
    }
    return MFSGD_OK;
}

// Bucket the member's share of the source into `nblk` blocks. On success *out_recs (device, owned by the
// caller) holds the records sorted by block and off[nblk+1] the offsets.
static int bucket_with(Member& m, const Source& src, Chunk& ch, BucketArgs b, Rec** out_recs, std::vector<int64_t>& off);

static int bucket_member(mfsgd_handle* h, Member& m, const Source& src, Chunk& ch, int want_held, int row_div, int col_div,
                         int n_cols, bool split_hot, Rec** out_recs, std::vector<int64_t>& off) {
    BucketArgs b{};
    b.owner_u = m.d_owner_u;
    b.owner_i = m.d_owner_i;
    b.ub_lo = m.g * h->mu;
    b.ub_hi = (m.g + 1) * h->mu;
    b.row_div = row_div;
    b.col_div = col_div;
    b.n_cols = n_cols;
    b.want_held = want_held;
    b.center = h->center;          // model extension: ratings are stored centred (training and held-out sets alike)
    if (split_hot && h->H > 0) {
        b.hot_index = m.d_hot_index;
        b.n_hot = h->H;
        b.hot_base = bucket_rows(b) * n_cols;
        b.heavy_bits = m.d_heavy_bits;      // training layout: mark the heavy users' records for the run kernel
    }
    return bucket_with(m, src, ch, b, out_recs, off);
}

// The source's records that fall into the blocks `b` describes, sorted by block: histogram pass, scatter pass.
static int bucket_with(Member& m, const Source& src, Chunk& ch, BucketArgs b, Rec** out_recs, std::vector<int64_t>& off) {
    CK(cudaSetDevice(m.device));
    const int nblk = bucket_block_count(b);
    if (nblk > MAX_BUCKETS) return fail(MFSGD_E_INVALID_ARG, "%d blocks per ring member exceed the %d limit", nblk, MAX_BUCKETS);
    unsigned long long* d_cnt = nullptr;
    CK(dev_alloc(&d_cnt, (size_t)nblk));
    int rc = MFSGD_OK;
    Rec* recs = nullptr;
    std::vector<unsigned long long> cnt((size_t)nblk, 0ULL);
    cudaError_t e = cudaMemsetAsync(d_cnt, 0, (size_t)nblk * 8, m.stream);
    for (int64_t start = 0; e == cudaSuccess && start < src.total; start += ch.cap) {
        const int64_t count = std::min(ch.cap, src.total - start);
        rc = chunk_stage(ch, src, start, count, m.stream, &m.launches);
        if (rc != MFSGD_OK) break;
        b.u = ch.u; b.i = ch.i; b.r = ch.r; b.n = count;
        b.held = src.synthetic ? ch.held : nullptr;
        if (src.all_held) { b.held = nullptr; b.want_held = 0; }   // explicit held-out arrays: take every record
        e = launch_block_histogram(b, d_cnt, m.stream, &m.launches);
    }
    if (rc == MFSGD_OK && e == cudaSuccess) e = cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)nblk * 8, cudaMemcpyDeviceToHost, m.stream);
    if (rc == MFSGD_OK && e == cudaSuccess) e = cudaStreamSynchronize(m.stream);
    if (rc == MFSGD_OK && e == cudaSuccess) {
        off.assign((size_t)nblk + 1, 0);
        for (int j = 0; j < nblk; j++) off[(size_t)j + 1] = off[(size_t)j] + (int64_t)cnt[(size_t)j];
        std::vector<unsigned long long> cursors((size_t)nblk);
        for (int j = 0; j < nblk; j++) cursors[(size_t)j] = (unsigned long long)off[(size_t)j];
        e = dev_alloc(&recs, (size_t)off[(size_t)nblk]);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_cnt, cursors.data(), (size_t)nblk * 8, cudaMemcpyHostToDevice, m.stream);
        for (int64_t start = 0; e == cudaSuccess && start < src.total; start += ch.cap) {
            const int64_t count = std::min(ch.cap, src.total - start);
            rc = chunk_stage(ch, src, start, count, m.stream, &m.launches);
            if (rc != MFSGD_OK) break;
            b.u = ch.u; b.i = ch.i; b.r = ch.r; b.n = count;
            b.held = src.synthetic ? ch.held : nullptr;
            if (src.all_held) { b.held = nullptr; b.want_held = 0; }
            e = launch_block_scatter(b, d_cnt, recs, m.stream, &m.launches);
        }
        if (rc == MFSGD_OK && e == cudaSuccess) e = cudaStreamSynchronize(m.stream);
    }
    raw_free(d_cnt);
    if (rc != MFSGD_OK || e != cudaSuccess) {
        if (recs) raw_free(recs);
        if (rc != MFSGD_OK) return rc;
        CK(e);
    }
    *out_recs = recs;
    return MFSGD_OK;
}

// Multi-process ring, sharded input: this rank's slice is sorted by the ring member that owns each record's user stripe and
// the slices travel member to member in one grouped ncclSend / ncclRecv round (an all-to-all of 12-byte records over NVLink).
// On return *recv holds every training record of this member's stripe (device, caller frees), *n_recv their number.
static int exchange_slices(mfsgd_handle* h, Member& m, const Source& src, Chunk& ch, Rec** recv, int64_t* n_recv) {
    const int G = h->G;
    *recv = nullptr;
    *n_recv = 0;
    BucketArgs b{};                       // one block per destination member: row = owner_u / mu, a single column
    b.owner_u = m.d_owner_u;
    b.owner_i = m.d_owner_i;
    b.ub_lo = 0;
    b.ub_hi = h->UB;
    b.row_div = h->mu;
    b.col_div = h->IB;
    b.n_cols = 1;
    b.want_held = 0;
    Rec* send = nullptr;
    std::vector<int64_t> soff;
    CKRC(bucket_with(m, src, ch, b, &send, soff));
    std::vector<unsigned long long> cnt((size_t)G), all((size_t)G * G);
    for (int p = 0; p < G; p++) cnt[(size_t)p] = (unsigned long long)(soff[(size_t)p + 1] - soff[(size_t)p]);
    unsigned long long *d_cnt = nullptr, *d_all = nullptr;
    int rc = MFSGD_OK;
    cudaError_t e = dev_alloc(&d_cnt, (size_t)G);
    if (e == cudaSuccess) e = dev_alloc(&d_all, (size_t)G * G);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cnt, cnt.data(), (size_t)G * 8, cudaMemcpyHostToDevice, m.stream);
    if (e == cudaSuccess) {
        ncclResult_t r = g_nccl.AllGather(d_cnt, d_all, (size_t)G, ncclUint64, h->comm, m.stream);
        if (r != ncclSuccess) rc = fail(MFSGD_E_NCCL, "all-gather of the slice sizes: %s", g_nccl.GetErrorString(r));
    }
    if (rc == MFSGD_OK && e == cudaSuccess) e = cudaMemcpyAsync(all.data(), d_all, (size_t)G * G * 8, cudaMemcpyDeviceToHost, m.stream);
    if (rc == MFSGD_OK && e == cudaSuccess) e = cudaStreamSynchronize(m.stream);
    std::vector<int64_t> roff((size_t)G + 1, 0);
    if (rc == MFSGD_OK && e == cudaSuccess) {
        for (int p = 0; p < G; p++) roff[(size_t)p + 1] = roff[(size_t)p] + (int64_t)all[(size_t)p * G + (size_t)m.g];   // what p sends to me
        e = dev_alloc(recv, (size_t)roff[(size_t)G]);
    }
    if (rc == MFSGD_OK && e == cudaSuccess) {
        ncclResult_t r = g_nccl.GroupStart();
        for (int p = 0; p < G && r == ncclSuccess; p++) {
            const int64_t ns = soff[(size_t)p + 1] - soff[(size_t)p], nr = roff[(size_t)p + 1] - roff[(size_t)p];
            if (ns > 0) r = g_nccl.Send(send + soff[(size_t)p], (size_t)ns * sizeof(Rec), ncclChar, p, h->comm, m.stream);
            if (r == ncclSuccess && nr > 0) r = g_nccl.Recv(*recv + roff[(size_t)p], (size_t)nr * sizeof(Rec), ncclChar, p, h->comm, m.stream);
        }
        ncclResult_t r2 = g_nccl.GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) rc = fail(MFSGD_E_NCCL, "record exchange: %s", g_nccl.GetErrorString(r));
        else e = cudaStreamSynchronize(m.stream);
        m.launches += 1;
    }
    raw_free(d_cnt);
    raw_free(d_all);
    raw_free(send);
    if (rc == MFSGD_OK && e != cudaSuccess) rc = fail(e == cudaErrorMemoryAllocation ? MFSGD_E_OOM : MFSGD_E_CUDA, "record exchange: %s", cudaGetErrorString(e));
    if (rc != MFSGD_OK) {
        if (*recv) raw_free(*recv);
        *recv = nullptr;
        return rc;
    }
    *n_recv = roff[(size_t)G];
    return MFSGD_OK;
}

// Record range of visit (sub-stripe sa, round rnd) inside bucket `blk` (cold block or hot bucket).
static inline void slice_of(const Member& m, size_t blk_lo, size_t blk_hi, int rnd, int rounds, int64_t* lo, int64_t* hi) {
    const int64_t blo = m.block_off[blk_lo], bhi = m.block_off[blk_hi];
    *lo = blo + (bhi - blo) * rnd / rounds;
    *hi = blo + (bhi - blo) * (rnd + 1) / rounds;
}

// host threads of the run planner: 0 = its own choice (one for small plans, up to 8 for large ones); MFSGD_PLAN_THREADS overrides
// (tests force 1, 3 and 8 and compare the plans bit for bit)
static int plan_threads_env() {
    const char* e = getenv("MFSGD_PLAN_THREADS");
    if (!e || !*e) return 0;
    const int t = atoi(e);
    return t < 0 ? 0 : std::min(t, 64);
}

// Hot-item units of every visit (sa, grp, rnd): the round's slice of each (sa, hot item) bucket, cut into
// runs of <= hot_chunk records. Bucket sizes never change, so this is built once per load.
static int build_hot_units(mfsgd_handle* h, Member& m) {
    dev_free(m.d_units);
    dev_free(m.d_counters);
    m.visit_units.assign((size_t)h->mu * h->rounds * h->IB + 1, 0);
    if (h->H == 0 || h->cfg.mode == MFSGD_MODE_DETERMINISTIC) return MFSGD_OK;
    CK(cudaSetDevice(m.device));
    const size_t hb = (size_t)h->mu * h->IB;
    const bool laned = (h->multi_process && h->G > 1 && h->mi > 1) || ((h->cfg.flags & MFSGD_FLAG_SPLIT_SHARDS) && h->mi > 1);
    const int chunk = plan_run_length(h->cfg.hot_chunk, laned ? 2 : 1, h->mu, h->rounds, h->IB, m.block_off.back() - m.block_off[hb],
                                      m.hot_grid, 32 / run_kernel_lanes(h->cfg.k));      // run_plan.hpp
    std::vector<HotUnit> units;
    RunPlanArgs pa{};
    pa.block_off = m.block_off.data();
    pa.n_blocks = m.block_off.size() - 1;
    pa.mu = h->mu; pa.H = h->H; pa.IB = h->IB; pa.rounds = h->rounds; pa.chunk = chunk;
    pa.member = m.g;
    pa.seed = h->cfg.seed;
    pa.boost = h->cfg.merge_boost > 0.f ? h->cfg.merge_boost : DEFAULT_MERGE_BOOST;
    pa.hot_block_lo = h->hot_block_lo.data();
    pa.hot_items = h->hot_items.data();
    plan_runs(pa, units, m.visit_units, plan_threads_env());
    m.run_chunk = chunk;
    m.unit_recs_cum.assign(units.size() + 1, 0);
    for (size_t j = 0; j < units.size(); j++) m.unit_recs_cum[j + 1] = m.unit_recs_cum[j] + units[j].count;
    m.n_counters = h->mu * h->rounds * h->IB;
    CK(dev_alloc(&m.d_units, units.size()));
    CK(dev_alloc(&m.d_counters, (size_t)m.n_counters * COUNTER_EPOCHS));
    if (!units.empty()) CK(cudaMemcpy(m.d_units, units.data(), units.size() * sizeof(HotUnit), cudaMemcpyHostToDevice));
    return MFSGD_OK;
}

static int load_training(mfsgd_handle* h, const Source& src, bool with_heldout) {
    const mfsgd_config& c = h->cfg;
    for (Member& m : h->members) free_member_data(m);
    h->loaded = false;
    h->factors_ready = false;
    h->epoch = 0;
    h->lr_now = c.lr;
    h->es_best = std::numeric_limits<double>::infinity();
    h->es_bad = 0;
    h->es_stopped = false;
    choose_blocking(h);
    if (h->UB > 65535 || h->IB > 65535) return fail(MFSGD_E_INVALID_ARG, "too many blocks");
    if (c.mode == MFSGD_MODE_DETERMINISTIC && src.total > CHUNK_MAX)     // staged and sorted as one chunk (one warp walks it anyway)
        return fail(MFSGD_E_INVALID_ARG, "DETERMINISTIC mode takes at most 2^28 records");
    const int64_t cap = std::max<int64_t>(1, std::min(src.total, CHUNK_MAX));
    int rc = MFSGD_OK;
    for (size_t mi_ = 0; mi_ < h->members.size() && rc == MFSGD_OK; mi_++) {
        Member& m = h->members[mi_];
        cudaSetDevice(m.device);
        Chunk ch;
        { PhaseTimer pt("load: staging buffers"); rc = chunk_alloc(ch, cap, src.synthetic); }
        if (rc == MFSGD_OK && mi_ == 0) {
            PhaseTimer pt("load: H2D/generate + row counts + balanced bounds");
            if (src.total > 0) rc = compute_bounds(h, src, ch);
            else { uniform_bounds(h); h->n_train_total = 0; }
        }
        if (rc == MFSGD_OK) { PhaseTimer pt("load: factor + owner tables"); rc = member_alloc_factors(h, m); }
        if (rc == MFSGD_OK && c.mode == MFSGD_MODE_DETERMINISTIC) {
            // keep the caller's order; validate ids happened in compute_bounds; synthetic sources drop held records
            if (src.synthetic) rc = fail(MFSGD_E_INVALID_ARG, "DETERMINISTIC mode takes host triplets (use mfsgd_generate_to_host)");
            if (rc == MFSGD_OK && src.total > 0) {
                rc = chunk_stage(ch, src, 0, src.total, m.stream, &m.launches);
                cudaError_t e = cudaSuccess;
                if (rc == MFSGD_OK) e = dev_alloc(&m.recs_orig, (size_t)src.total);
                if (e == cudaSuccess) e = dev_alloc(&m.recs[0], (size_t)src.total);
                if (e == cudaSuccess) e = dev_alloc(&m.keys_a, (size_t)src.total);
                if (e == cudaSuccess) e = dev_alloc(&m.keys_b, (size_t)src.total);
                if (e == cudaSuccess) e = launch_pack_records(ch.u, ch.i, ch.r, src.total, h->center, m.recs_orig, m.stream, &m.launches);
                if (e == cudaSuccess) e = deterministic_order_gather(nullptr, nullptr, (int32_t)src.total, 0, 0, m.keys_a, m.keys_b, nullptr, &m.sort_temp_bytes, m.stream, nullptr);
                if (e == cudaSuccess) e = raw_alloc(&m.sort_temp, m.sort_temp_bytes ? m.sort_temp_bytes : 1);
                if (e == cudaSuccess) e = cudaStreamSynchronize(m.stream);
                if (rc == MFSGD_OK && e != cudaSuccess) rc = fail(e == cudaErrorMemoryAllocation ? MFSGD_E_OOM : MFSGD_E_CUDA, "deterministic staging: %s", cudaGetErrorString(e));
            }
            m.n_recs = src.total;
            m.block_off.assign(2, 0);
            m.block_off[1] = src.total;
        } else if (rc == MFSGD_OK) {
            Rec* mine = nullptr;          // sharded input: the records the ring sent this member
            if (src.sharded) {
                int64_t n_mine = 0;
                { PhaseTimer pt("load: record exchange (route by stripe owner + all-to-all)"); rc = exchange_slices(h, m, src, ch, &mine, &n_mine); }
                if (rc == MFSGD_OK) {
                    Source own;
                    own.drecs = mine;
                    own.total = n_mine;
                    Chunk ch2;
                    rc = chunk_alloc(ch2, std::max<int64_t>(1, std::min(n_mine, CHUNK_MAX)), false);
                    if (rc == MFSGD_OK) { PhaseTimer pt("load: bucketing (histogram + scatter)"); rc = bucket_member(h, m, own, ch2, 0, 1, 1, h->IB, true, &m.recs[0], m.block_off); }
                    chunk_free(ch2);
                }
                if (mine) raw_free(mine);
            } else
            { PhaseTimer pt("load: bucketing (histogram + scatter)"); rc = bucket_member(h, m, src, ch, 0, 1, 1, h->IB, true, &m.recs[0], m.block_off); }
            if (rc == MFSGD_OK) {
                m.n_recs = m.block_off.back();
                m.rcur = 0;
                cudaError_t e = dev_alloc(&m.d_block_off, m.block_off.size());   // recs[1] is allocated by the first materialising reshuffle
                if (e == cudaSuccess) e = cudaMemcpy(m.d_block_off, m.block_off.data(), m.block_off.size() * 8, cudaMemcpyHostToDevice);
                if (e != cudaSuccess) rc = fail(e == cudaErrorMemoryAllocation ? MFSGD_E_OOM : MFSGD_E_CUDA, "record buffers: %s", cudaGetErrorString(e));
            }
            if (rc == MFSGD_OK && with_heldout) {
                free_eval(m.heldout);
                rc = bucket_member(h, m, src, ch, 1, h->mu, h->mi, h->G, false, &m.heldout.recs, m.heldout.group_off);
                if (rc == MFSGD_OK) m.heldout.n = m.heldout.group_off.back();
            }
        }
        chunk_free(ch);
    }
    if (rc != MFSGD_OK) {
        for (Member& m : h->members) free_member_data(m);
        return rc;
    }
    // interleaved passes per sub-epoch (run_plan.hpp: plan_rounds)
    h->rounds = plan_rounds(c.rounds, c.mode, h->G, h->mu, h->members[0].n_recs, h->members[0].u_hi - h->members[0].u_lo);
    {
        const bool pipelined = h->multi_process && h->G > 1 && h->mi > 1;
        const int parts = (pipelined || (c.flags & MFSGD_FLAG_SPLIT_SHARDS)) ? h->mi : 1;
        h->virtual_shuffle = c.mode != MFSGD_MODE_DETERMINISTIC && !(c.flags & (MFSGD_FLAG_NO_SHUFFLE | MFSGD_FLAG_MATERIALIZE_SHUFFLE)) &&
                             h->mi / parts == 1;   // every cold launch must lie inside one block
    }
    for (Member& m : h->members) {
        PhaseTimer pt("load: hot-unit lists");
        rc = build_hot_units(h, m);
        if (rc != MFSGD_OK) {
            for (Member& mm : h->members) free_member_data(mm);
            return rc;
        }
    }
    h->loaded = true;
    return MFSGD_OK;
}

static int mfsgd_load_ratings_body(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n);
extern "C" int mfsgd_load_ratings(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n) {
    return guarded([&]() { return mfsgd_load_ratings_body(h, users, items, ratings, n); });
}
static int mfsgd_load_ratings_body(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n) {
    if (!h) return fail(MFSGD_E_INVALID_ARG, "handle is null");
    if (n < 0 || (n > 0 && (!users || !items || !ratings))) return fail(MFSGD_E_INVALID_ARG, "null triplet arrays or negative n");
    Source s;
    s.hu = users; s.hi = items; s.hr = ratings; s.total = n;
    return load_training(h, s, false);
}

static int mfsgd_load_ratings_sharded_body(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n);
extern "C" int mfsgd_load_ratings_sharded(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n) {
    return guarded([&]() { return mfsgd_load_ratings_sharded_body(h, users, items, ratings, n); });
}
static int mfsgd_load_ratings_sharded_body(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n) {
    if (!h) return fail(MFSGD_E_INVALID_ARG, "handle is null");
    if (n < 0 || (n > 0 && (!users || !items || !ratings))) return fail(MFSGD_E_INVALID_ARG, "null triplet arrays or negative n");
    if (!h->multi_process) return mfsgd_load_ratings(h, users, items, ratings, n);     // one process holds everything anyway
    Source s;
    s.hu = users; s.hi = items; s.hr = ratings; s.total = n;
    s.sharded = true;
    return load_training(h, s, false);
}

static int mfsgd_generate_synthetic_body(mfsgd_handle* h, const mfsgd_synth_params* sp, int64_t* n_train, int64_t* n_heldout);
extern "C" int mfsgd_generate_synthetic(mfsgd_handle* h, const mfsgd_synth_params* sp, int64_t* n_train, int64_t* n_heldout) {
    return guarded([&]() { return mfsgd_generate_synthetic_body(h, sp, n_train, n_heldout); });
}
static int mfsgd_generate_synthetic_body(mfsgd_handle* h, const mfsgd_synth_params* sp, int64_t* n_train, int64_t* n_heldout) {
    if (!h || !sp) return fail(MFSGD_E_INVALID_ARG, "null argument");
    if (sp->n_total < 0 || sp->log2_alpha_user < 0 || sp->log2_alpha_user > 6 || sp->log2_alpha_item < 0 || sp->log2_alpha_item > 6 ||
        !(sp->c_user >= 0.0 && sp->c_user < 1.0) || !(sp->c_item >= 0.0 && sp->c_item < 1.0) || !(sp->planted_amplitude >= 0.f) ||
        !(sp->noise_scale >= 0.f))
        return fail(MFSGD_E_INVALID_ARG, "bad synthetic parameters");
    Source s;
    s.synthetic = true;
    s.total = sp->n_total;
    s.synth.seed = sp->seed;
    s.synth.n_users = h->cfg.n_users;
    s.synth.n_items = h->cfg.n_items;
    s.synth.l2au = sp->log2_alpha_user;
    s.synth.l2ai = sp->log2_alpha_item;
    s.synth.cu = sp->c_user;
    s.synth.ci = sp->c_item;
    s.synth.amplitude = sp->planted_amplitude > 0.f ? sp->planted_amplitude : 0.8660254f;
    s.synth.noise_scale = sp->noise_scale > 0.f ? sp->noise_scale : 0.5f;
    CKRC(load_training(h, s, true));
    int64_t nt = 0, nh = 0;
    for (Member& m : h->members) { nt += m.n_recs; nh += m.heldout.n; }
    if (n_train) *n_train = nt;
    if (n_heldout) *n_heldout = nh;
    return MFSGD_OK;
}

static int build_eval_set(mfsgd_handle* h, Member& m, const int32_t* users, const int32_t* items, const float* ratings,
                          int64_t n, EvalSet& out) {
    Source s;
    s.hu = users; s.hi = items; s.hr = ratings; s.total = n; s.all_held = true;
    CK(cudaSetDevice(m.device));
    // validate ids on the device first (count kernel with throw-away counters would cost memory: do it on the host)
    for (int64_t t = 0; t < n; t++)
        if (users[t] < 0 || users[t] >= h->cfg.n_users || items[t] < 0 || items[t] >= h->cfg.n_items)
            return fail(MFSGD_E_INVALID_ARG, "record %lld has an id outside [0,n_users) x [0,n_items)", (long long)t);
    Chunk ch;
    int rc = chunk_alloc(ch, std::max<int64_t>(1, std::min(n, CHUNK_MAX)), false);
    if (rc == MFSGD_OK) {
        free_eval(out);
        rc = bucket_member(h, m, s, ch, 0, h->mu, h->mi, h->G, false, &out.recs, out.group_off);
        if (rc == MFSGD_OK) out.n = out.group_off.back();
    }
    chunk_free(ch);
    return rc;
}

static int mfsgd_load_heldout_body(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n);
extern "C" int mfsgd_load_heldout(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n) {
    return guarded([&]() { return mfsgd_load_heldout_body(h, users, items, ratings, n); });
}
static int mfsgd_load_heldout_body(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n) {
    if (!h) return fail(MFSGD_E_INVALID_ARG, "handle is null");
    if (!h->loaded) return fail(MFSGD_E_STATE, "call mfsgd_load_ratings first (it fixes the stripe bounds)");
    if (n < 0 || (n > 0 && (!users || !items || !ratings))) return fail(MFSGD_E_INVALID_ARG, "null triplet arrays or negative n");
    for (Member& m : h->members) CKRC(build_eval_set(h, m, users, items, ratings, n, m.heldout));
    return MFSGD_OK;
}

// ------------------------------------------------------------------------------------------------
// factors
// ------------------------------------------------------------------------------------------------
static int mfsgd_init_factors_body(mfsgd_handle* h);
extern "C" int mfsgd_init_factors(mfsgd_handle* h) {
    return guarded([&]() { return mfsgd_init_factors_body(h); });
}
static int mfsgd_init_factors_body(mfsgd_handle* h) {
    if (!h) return fail(MFSGD_E_INVALID_ARG, "handle is null");
    if (!h->loaded) return fail(MFSGD_E_STATE, "load ratings before initialising factors");
    for (Member& m : h->members) {
        CK(cudaSetDevice(m.device));
        CK(launch_init_factors(m.P, h->p_half, m.u_hi - m.u_lo, h->cfg.k, m.u_lo, h->cfg.seed, STREAM_P_INIT, h->scale, m.stream, &m.launches));
        const int lo = group_lo(h, m.held_group), hi = group_hi(h, m.held_group);
        CK(launch_init_factors(m.Q[m.cur], false, hi - lo, h->cfg.k, lo, h->cfg.seed, STREAM_Q_INIT, h->scale, m.stream, &m.launches));
        if (h->biases) {          // MatrixFactorizationSGD.java:305: biases start at 0
            CK(cudaMemsetAsync(m.BU, 0, (size_t)(m.u_hi - m.u_lo) * 4, m.stream));
            CK(cudaMemsetAsync(m.BQ[m.cur], 0, (size_t)(hi - lo) * 4, m.stream));
        }
    }
    for (Member& m : h->members) { CK(cudaSetDevice(m.device)); CK(cudaStreamSynchronize(m.stream)); }
    h->factors_ready = true;
    return MFSGD_OK;
}

static int mfsgd_set_factors_body(mfsgd_handle* h, const float* P, const float* Q);
extern "C" int mfsgd_set_factors(mfsgd_handle* h, const float* P, const float* Q) {
    return guarded([&]() { return mfsgd_set_factors_body(h, P, Q); });
}
static int mfsgd_set_factors_body(mfsgd_handle* h, const float* P, const float* Q) {
    if (!h || !P || !Q) return fail(MFSGD_E_INVALID_ARG, "null argument");
    if (!h->loaded) return fail(MFSGD_E_STATE, "load ratings before setting factors");
    const int k = h->cfg.k;
    for (Member& m : h->members) {
        CK(cudaSetDevice(m.device));
        const size_t pn = (size_t)(m.u_hi - m.u_lo) * k;
        if (h->p_half) {           // binary16 rows: stage the caller's binary32 rows, narrow (round to nearest even) on the device
            float* tmp = nullptr;
            CK(dev_alloc(&tmp, std::max<size_t>(pn, 1)));
            cudaError_t e = cudaMemcpyAsync(tmp, P + (size_t)m.u_lo * k, pn * 4, cudaMemcpyHostToDevice, m.stream);
            if (e == cudaSuccess) e = launch_narrow_rows(tmp, (int64_t)pn, m.P, m.stream, &m.launches);
            if (e == cudaSuccess) e = cudaStreamSynchronize(m.stream);
            dev_free(tmp);
            CK(e);
        } else {
            CK(cudaMemcpyAsync(m.P, P + (size_t)m.u_lo * k, pn * 4, cudaMemcpyHostToDevice, m.stream));
        }
        const int lo = group_lo(h, m.held_group), hi = group_hi(h, m.held_group);
        CK(cudaMemcpyAsync(m.Q[m.cur], Q + (size_t)lo * k, (size_t)(hi - lo) * k * 4, cudaMemcpyHostToDevice, m.stream));
        if (h->biases && !h->factors_ready) {      // fresh factors: biases start at 0 (mfsgd_set_biases overrides)
            CK(cudaMemsetAsync(m.BU, 0, (size_t)(m.u_hi - m.u_lo) * 4, m.stream));
            CK(cudaMemsetAsync(m.BQ[m.cur], 0, (size_t)(hi - lo) * 4, m.stream));
        }
    }
    for (Member& m : h->members) { CK(cudaSetDevice(m.device)); CK(cudaStreamSynchronize(m.stream)); }
    h->factors_ready = true;
    return MFSGD_OK;
}

static int mfsgd_get_factors_body(mfsgd_handle* h, float* P, float* Q);
extern "C" int mfsgd_get_factors(mfsgd_handle* h, float* P, float* Q) {
    return guarded([&]() { return mfsgd_get_factors_body(h, P, Q); });
}
static int mfsgd_get_factors_body(mfsgd_handle* h, float* P, float* Q) {
    if (!h || !P || !Q) return fail(MFSGD_E_INVALID_ARG, "null argument");
    if (!h->factors_ready) return fail(MFSGD_E_STATE, "factors are not initialised");
    const int k = h->cfg.k;
    for (Member& m : h->members) {
        CK(cudaSetDevice(m.device));
        const size_t pn = (size_t)(m.u_hi - m.u_lo) * k;
        if (h->p_half) {           // binary16 rows come back widened (exact)
            float* tmp = nullptr;
            CK(dev_alloc(&tmp, std::max<size_t>(pn, 1)));
            cudaError_t e = launch_widen_rows(m.P, (int64_t)pn, tmp, m.stream, &m.launches);
            if (e == cudaSuccess) e = cudaMemcpyAsync(P + (size_t)m.u_lo * k, tmp, pn * 4, cudaMemcpyDeviceToHost, m.stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(m.stream);
            dev_free(tmp);
            CK(e);
        } else {
            CK(cudaMemcpyAsync(P + (size_t)m.u_lo * k, m.P, pn * 4, cudaMemcpyDeviceToHost, m.stream));
        }
        const int lo = group_lo(h, m.held_group), hi = group_hi(h, m.held_group);
        CK(cudaMemcpyAsync(Q + (size_t)lo * k, m.Q[m.cur], (size_t)(hi - lo) * k * 4, cudaMemcpyDeviceToHost, m.stream));
    }
    for (Member& m : h->members) { CK(cudaSetDevice(m.device)); CK(cudaStreamSynchronize(m.stream)); }
    return MFSGD_OK;
}

static int mfsgd_get_model_body(mfsgd_handle* h, float* global_mean, float* user_bias, float* item_bias);
extern "C" int mfsgd_get_model(mfsgd_handle* h, float* global_mean, float* user_bias, float* item_bias) {
    return guarded([&]() { return mfsgd_get_model_body(h, global_mean, user_bias, item_bias); });
}
static int mfsgd_get_model_body(mfsgd_handle* h, float* global_mean, float* user_bias, float* item_bias) {
    if (!h) return fail(MFSGD_E_INVALID_ARG, "handle is null");
    if (!h->loaded) return fail(MFSGD_E_STATE, "no ratings loaded (the global mean is the training set's)");
    if (global_mean) *global_mean = h->center;
    if (!user_bias && !item_bias) return MFSGD_OK;
    if (!h->biases) return fail(MFSGD_E_STATE, "MFSGD_MODEL_BIASES is off");
    if (!h->factors_ready) return fail(MFSGD_E_STATE, "factors are not initialised");
    for (Member& m : h->members) {
        CK(cudaSetDevice(m.device));
        if (user_bias) CK(cudaMemcpyAsync(user_bias + m.u_lo, m.BU, (size_t)(m.u_hi - m.u_lo) * 4, cudaMemcpyDeviceToHost, m.stream));
        const int lo = group_lo(h, m.held_group), hi = group_hi(h, m.held_group);
        if (item_bias) CK(cudaMemcpyAsync(item_bias + lo, m.BQ[m.cur], (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, m.stream));
    }
    for (Member& m : h->members) { CK(cudaSetDevice(m.device)); CK(cudaStreamSynchronize(m.stream)); }
    return MFSGD_OK;
}

static int mfsgd_set_biases_body(mfsgd_handle* h, const float* user_bias, const float* item_bias);
extern "C" int mfsgd_set_biases(mfsgd_handle* h, const float* user_bias, const float* item_bias) {
    return guarded([&]() { return mfsgd_set_biases_body(h, user_bias, item_bias); });
}
static int mfsgd_set_biases_body(mfsgd_handle* h, const float* user_bias, const float* item_bias) {
    if (!h || !user_bias || !item_bias) return fail(MFSGD_E_INVALID_ARG, "null argument");
    if (!h->biases) return fail(MFSGD_E_STATE, "MFSGD_MODEL_BIASES is off");
    if (!h->loaded) return fail(MFSGD_E_STATE, "load ratings before setting biases");
    for (Member& m : h->members) {
        CK(cudaSetDevice(m.device));
        CK(cudaMemcpyAsync(m.BU, user_bias + m.u_lo, (size_t)(m.u_hi - m.u_lo) * 4, cudaMemcpyHostToDevice, m.stream));
        const int lo = group_lo(h, m.held_group), hi = group_hi(h, m.held_group);
        CK(cudaMemcpyAsync(m.BQ[m.cur], item_bias + lo, (size_t)(hi - lo) * 4, cudaMemcpyHostToDevice, m.stream));
    }
    for (Member& m : h->members) { CK(cudaSetDevice(m.device)); CK(cudaStreamSynchronize(m.stream)); }
    return MFSGD_OK;
}

static int mfsgd_get_partition_body(mfsgd_handle* h, int32_t* u_lo, int32_t* u_hi, int32_t* i_lo, int32_t* i_hi);
extern "C" int mfsgd_get_partition(mfsgd_handle* h, int32_t* u_lo, int32_t* u_hi, int32_t* i_lo, int32_t* i_hi) {
    return guarded([&]() { return mfsgd_get_partition_body(h, u_lo, u_hi, i_lo, i_hi); });
}
static int mfsgd_get_partition_body(mfsgd_handle* h, int32_t* u_lo, int32_t* u_hi, int32_t* i_lo, int32_t* i_hi) {
    if (!h) return fail(MFSGD_E_INVALID_ARG, "handle is null");
    if (!h->loaded) return fail(MFSGD_E_STATE, "no ratings loaded");
    int32_t ul = h->cfg.n_users, uh = 0, il = h->cfg.n_items, ih = 0;
    for (Member& m : h->members) {
        ul = std::min(ul, m.u_lo); uh = std::max(uh, m.u_hi);
        il = std::min(il, (int32_t)group_lo(h, m.held_group)); ih = std::max(ih, (int32_t)group_hi(h, m.held_group));
    }
    if (u_lo) *u_lo = ul;
    if (u_hi) *u_hi = uh;
    if (i_lo) *i_lo = il;
    if (i_hi) *i_hi = ih;
    return MFSGD_OK;
}

// ------------------------------------------------------------------------------------------------
// subsystem (4): ring rotation of the Q shard groups
// ------------------------------------------------------------------------------------------------
// Pipelined rotation (one process per GPU, shards_per_gpu > 1): as soon as the launches of item sub-shard `part` are
// done, that slice of the held Q group travels on the copy stream (ncclSend to g-1 / ncclRecv from g+1) while the
// compute stream trains the next sub-shard; the next sub-epoch waits per sub-shard, not for the whole group.
static int ensure_part_events(mfsgd_handle* h, Member& m) {
    while ((int)m.ev_part_done.size() < h->mi) {
        cudaEvent_t a, b;
        CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        m.ev_part_done.push_back(a);
        m.ev_part_recv.push_back(b);
        m.part_recv_pending.push_back(0);
    }
    return MFSGD_OK;
}

// `stream` waits for the slice of the held group that item sub-shard `part` trains next
static int wait_part(mfsgd_handle* h, Member& m, int part, cudaStream_t stream) {
    if (m.win.active) return ring_wait(h, m, stream, ring_arrival_off(part), m.win.sent[(size_t)part]);
    CK(cudaStreamWaitEvent(stream, m.ev_part_recv[(size_t)part], 0));
    return MFSGD_OK;
}

static int rotate_part(mfsgd_handle* h, Member& m, int part, cudaStream_t after) {
    const int G = h->G, k = h->cfg.k;
    const int to = (m.g - 1 + G) % G, from = (m.g + 1) % G;
    const int send_grp = m.held_group, recv_grp = (m.held_group + 1) % G;
    const int32_t s_lo = h->item_bounds[(size_t)send_grp * h->mi + part], s_hi = h->item_bounds[(size_t)send_grp * h->mi + part + 1];
    const int32_t r_lo = h->item_bounds[(size_t)recv_grp * h->mi + part], r_hi = h->item_bounds[(size_t)recv_grp * h->mi + part + 1];
    float* sptr = m.Q[m.cur] + (size_t)(s_lo - group_lo(h, send_grp)) * k;
    float* rptr = m.Q[m.cur ^ 1] + (size_t)(r_lo - group_lo(h, recv_grp)) * k;
    CK(cudaEventRecord(m.ev_part_done[(size_t)part], after));
    CK(cudaStreamWaitEvent(m.copy_stream, m.ev_part_done[(size_t)part], 0));
    if (m.win.active) {          // ring window: peer write + flags (see "ring window" above); every rank flips cur in step
        Member::Window& w = m.win;
        const uint32_t n = ++w.sent[(size_t)part];
        const size_t row0 = (size_t)(s_lo - group_lo(h, send_grp));
        if (n > 1) CKRC(ring_wait(h, m, m.copy_stream, ring_credit_off(part), n - 1));
        CK(cudaMemcpyAsync(w.to_base + w.off_q[m.cur ^ 1] + row0 * k * 4, sptr, (size_t)(s_hi - s_lo) * k * 4, cudaMemcpyDefault, m.copy_stream));
        if (h->biases)
            CK(cudaMemcpyAsync(w.to_base + w.off_bq[m.cur ^ 1] + row0 * 4, m.BQ[m.cur] + row0, (size_t)(s_hi - s_lo) * 4, cudaMemcpyDefault, m.copy_stream));
        CKRC(ring_signal(h, m, m.copy_stream, w.to_base, ring_arrival_off(part), n));
        CKRC(ring_signal(h, m, m.copy_stream, w.from_base, ring_credit_off(part), n));
        m.part_recv_pending[(size_t)part] = 1;
        (void)rptr;
        return MFSGD_OK;
    }
    CKN(g_nccl.GroupStart());
    CKN(g_nccl.Send(sptr, (size_t)(s_hi - s_lo) * k, ncclFloat, to, h->comm, m.copy_stream));
    CKN(g_nccl.Recv(rptr, (size_t)(r_hi - r_lo) * k, ncclFloat, from, h->comm, m.copy_stream));
    if (h->biases) {         // the slice's item biases travel with it
        CKN(g_nccl.Send(m.BQ[m.cur] + (s_lo - group_lo(h, send_grp)), (size_t)(s_hi - s_lo), ncclFloat, to, h->comm, m.copy_stream));
        CKN(g_nccl.Recv(m.BQ[m.cur ^ 1] + (r_lo - group_lo(h, recv_grp)), (size_t)(r_hi - r_lo), ncclFloat, from, h->comm, m.copy_stream));
    }
    CKN(g_nccl.GroupEnd());
    CK(cudaEventRecord(m.ev_part_recv[(size_t)part], m.copy_stream));
    m.part_recv_pending[(size_t)part] = 1;
    m.launches += 1;
    return MFSGD_OK;
}

// After the kernels of a sub-epoch: member g hands the group it holds to member g-1 and takes the next
// one from member g+1. Everything is stream/event ordered; the host never blocks inside an epoch.
static int rotate_q(mfsgd_handle* h) {
    const int G = h->G, k = h->cfg.k;
    if (G == 1) return MFSGD_OK;
    if (h->multi_process) {
        Member& m = h->members[0];
        for (size_t part = 0; part < m.part_recv_pending.size(); part++)
            if (m.part_recv_pending[part]) {
                CKRC(wait_part(h, m, (int)part, m.stream));
                m.part_recv_pending[part] = 0;
            }
        const int to = (m.g - 1 + G) % G, from = (m.g + 1) % G;
        const int send_grp = m.held_group, recv_grp = (m.held_group + 1) % G;
        const size_t send_n = (size_t)(group_hi(h, send_grp) - group_lo(h, send_grp)) * k;
        const size_t recv_n = (size_t)(group_hi(h, recv_grp) - group_lo(h, recv_grp)) * k;
        CK(cudaSetDevice(m.device));
        CKN(g_nccl.GroupStart());
        CKN(g_nccl.Send(m.Q[m.cur], send_n, ncclFloat, to, h->comm, m.stream));
        CKN(g_nccl.Recv(m.Q[m.cur ^ 1], recv_n, ncclFloat, from, h->comm, m.stream));
        if (h->biases) {
            CKN(g_nccl.Send(m.BQ[m.cur], send_n / k, ncclFloat, to, h->comm, m.stream));
            CKN(g_nccl.Recv(m.BQ[m.cur ^ 1], recv_n / k, ncclFloat, from, h->comm, m.stream));
        }
        CKN(g_nccl.GroupEnd());
        m.launches += 1;
        m.cur ^= 1;
        m.held_group = recv_grp;
        return MFSGD_OK;
    }
    // single process: peer copies on the copy streams
    for (Member& m : h->members) {
        Member& dst = h->members[(size_t)((m.g - 1 + G) % G)];
        CK(cudaSetDevice(m.device));
        CK(cudaEventRecord(m.ev_compute, m.stream));
        CK(cudaStreamWaitEvent(m.copy_stream, m.ev_compute, 0));
        CK(cudaStreamWaitEvent(m.copy_stream, dst.ev_sent, 0));   // dst's previous send out of the buffer we overwrite
    }
    for (Member& m : h->members) {
        Member& dst = h->members[(size_t)((m.g - 1 + G) % G)];
        const size_t bytes = (size_t)(group_hi(h, m.held_group) - group_lo(h, m.held_group)) * k * 4;
        CK(cudaSetDevice(m.device));
        if (dst.device == m.device)
            CK(cudaMemcpyAsync(dst.Q[dst.cur ^ 1], m.Q[m.cur], bytes, cudaMemcpyDeviceToDevice, m.copy_stream));
        else
            CK(cudaMemcpyPeerAsync(dst.Q[dst.cur ^ 1], dst.device, m.Q[m.cur], m.device, bytes, m.copy_stream));
        if (h->biases) {
            if (dst.device == m.device)
                CK(cudaMemcpyAsync(dst.BQ[dst.cur ^ 1], m.BQ[m.cur], bytes / k, cudaMemcpyDeviceToDevice, m.copy_stream));
            else
                CK(cudaMemcpyPeerAsync(dst.BQ[dst.cur ^ 1], dst.device, m.BQ[m.cur], m.device, bytes / k, m.copy_stream));
        }
        CK(cudaEventRecord(m.ev_sent, m.copy_stream));   // also the receiver's "data arrived" signal
    }
    for (Member& m : h->members) {
        Member& src = h->members[(size_t)((m.g + 1) % G)];
        CK(cudaSetDevice(m.device));
        CK(cudaStreamWaitEvent(m.stream, src.ev_sent, 0));        // cross-device waits are legal; records are not
        m.cur ^= 1;
        m.held_group = (m.held_group + 1) % G;
    }
    return MFSGD_OK;
}

// ------------------------------------------------------------------------------------------------
// training
// ------------------------------------------------------------------------------------------------
static int timing_event(Member& m, cudaEvent_t* out) {
    if (m.ev_used == (int)m.evpool.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        m.evpool.push_back(e);
    }
    *out = m.evpool[(size_t)m.ev_used++];
    return MFSGD_OK;
}

// Lay out the records for `epoch` (subsystem 1). Deterministic mode: the stand-in's visiting order. Otherwise an
// in-block reshuffle into the spare buffer; with prefetch_next the layout of epoch + 1 is then produced in the
// background (low-priority stream, DRAM-bound) while the update kernels (L2-bound) consume this one.
static int shuffle_member(mfsgd_handle* h, Member& m, int epoch, bool prefetch_next) {
    CK(cudaSetDevice(m.device));
    if (h->cfg.mode == MFSGD_MODE_DETERMINISTIC) {
        if (m.n_recs > 0)
            CK(deterministic_order_gather(m.recs_orig, m.recs[0], (int32_t)m.n_recs, h->cfg.seed, (uint32_t)epoch, m.keys_a,
                                          m.keys_b, m.sort_temp, &m.sort_temp_bytes, m.stream, &m.launches));
        m.rcur = 0;
        return MFSGD_OK;
    }
    if (h->virtual_shuffle && prefetch_next) return MFSGD_OK;   // training: the kernels apply the permutation themselves
    const int nblk = (int)m.block_off.size() - 1;   // cold blocks + hot (stripe, item) buckets
    if (!m.recs[1]) CK(dev_alloc(&m.recs[1], (size_t)m.n_recs));
    if (m.ahead_epoch == epoch) {
        CK(cudaStreamWaitEvent(m.stream, m.ev_shuffle_done, 0));      // prepared during the previous epoch
    } else {
        if (m.ahead_epoch >= 0) CK(cudaStreamWaitEvent(m.stream, m.ev_shuffle_done, 0));   // a stale prefetch owns the spare buffer
        CK(launch_block_shuffle(m.recs[m.rcur], m.recs[m.rcur ^ 1], m.d_block_off, nblk, m.n_recs, h->cfg.seed, (uint32_t)epoch,
                                (uint32_t)(m.g * nblk), m.stream, &m.launches));
    }
    m.rcur ^= 1;
    m.ahead_epoch = -1;
    if (prefetch_next) {
        // recs[rcur] is read-only until the next reshuffle; recs[rcur ^ 1] is free once everything enqueued so far is done
        CK(cudaEventRecord(m.ev_shuffle_go, m.stream));
        CK(cudaStreamWaitEvent(m.shuffle_stream, m.ev_shuffle_go, 0));
        CK(launch_block_shuffle(m.recs[m.rcur], m.recs[m.rcur ^ 1], m.d_block_off, nblk, m.n_recs, h->cfg.seed, (uint32_t)(epoch + 1),
                                (uint32_t)(m.g * nblk), m.shuffle_stream, &m.launches));
        CK(cudaEventRecord(m.ev_shuffle_done, m.shuffle_stream));
        m.ahead_epoch = epoch + 1;
    }
    return MFSGD_OK;
}

static int rmse_pass(mfsgd_handle* h, bool heldout, double* sse_out, int64_t* n_out);

// Lane mode: the item sub-shards ("parts") of a sub-epoch are launched separately (pipelined rotation of a multi-process
// ring, or MFSGD_FLAG_SPLIT_SHARDS). Part p's launches go to lane p % N_LANES, a stream pair of its own: consecutive parts
// work on disjoint item rows (and Hogwild-style on the same P stripe), so they may run side by side -- the next part's
// runs fill the SMs the previous part's last runs leave idle, its Q slice travels while the other lane computes, and the
// chain of a lane runs ahead across sub-epoch and epoch boundaries as far as its own slices allow. The member's main
// stream only coordinates (timing events, joins). Round 1 launched all parts on one stream with an event in between,
// which cost ~25 % of an 8-GPU epoch in launch tails (profiles/r01_bench.md).
static const int N_LANES = 2;

static int ensure_lanes(Member& m, int n_lanes) {
    while ((int)m.lanes.size() < n_lanes) {
        Lane l;
        CK(cudaStreamCreateWithFlags(&l.cold, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&l.hot, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
        m.lanes.push_back(l);
    }
    return MFSGD_OK;
}

// main stream waits for everything the lanes have been given so far
static int join_lanes(Member& m) {
    for (Lane& l : m.lanes)
        if (l.used) {
            CK(cudaEventRecord(l.done, l.cold));
            CK(cudaStreamWaitEvent(m.stream, l.done, 0));
            l.used = false;
        }
    return MFSGD_OK;
}

// The update launches of item blocks [ib_lo, ib_hi) of sub-epoch s: every (sub-stripe, round) visit's cold records on
// `cold_s` (full-grid Hogwild kernel) and its runs on `hot_s` (run kernel); hot_s forks from cold_s before and joins after.
// chain_overlap: consecutive run launches of this chain may overlap at their tails (programmatic dependent launch). Never for a
// lane: the dependent launch takes whatever SM slots come free, and with a second lane's CTAs retiring all the time that is not
// the predecessor's tail any more -- two visits of the SAME items run side by side, both add their increments from one base, and
// the factors diverge (Yahoo-shaped on 8 real GPUs: NaN in the second epoch, on either transport; with plain stream order inside
// a lane -0.17 / -0.23 / -0.10 % of the oracle and 14 % faster, since the other lane fills the tails anyway).
static int enqueue_visits(mfsgd_handle* h, Member& m, UpdateArgs a, int s, size_t ib_lo, size_t ib_hi, bool fast_arith, bool chain_overlap,
                          cudaStream_t cold_s, cudaStream_t hot_s, cudaEvent_t ev_fork, cudaEvent_t ev_join, cudaEvent_t mark_cold,
                          cudaEvent_t mark_hot, bool* any_cold, bool* any_hot) {
    const mfsgd_config& c = h->cfg;
    struct Visit { int64_t lo, hi; size_t cblk; int unit_lo, unit_hi; };
    std::vector<Visit> visits;
    visits.reserve((size_t)h->mu * h->rounds);
    bool has_hot = false;
    for (int vis = 0; vis < h->mu * h->rounds; vis++) {
        const int rnd = vis / h->mu;
        // sub-stripe order of this round: a fresh rotation + direction per (epoch, sub-epoch, round)
        const uint64_t hv = hash64(c.seed, 10, ((uint64_t)h->epoch << 24) ^ ((uint64_t)s << 12) ^ (uint64_t)rnd);
        const int pos = vis % h->mu;
        const int sa = (int)(((hv >> 1) + (uint64_t)((hv & 1) ? pos : h->mu - 1 - pos)) % (uint64_t)h->mu);
        Visit v;
        slice_of(m, (size_t)sa * h->IB + ib_lo, (size_t)sa * h->IB + ib_hi, rnd, h->rounds, &v.lo, &v.hi);
        v.cblk = (size_t)sa * h->IB + ib_lo;
        const size_t vkey = ((size_t)sa * h->rounds + rnd) * h->IB;
        v.unit_lo = m.d_units ? m.visit_units[vkey + ib_lo] : 0;
        v.unit_hi = m.d_units ? m.visit_units[vkey + ib_hi] : 0;
        has_hot = has_hot || v.unit_hi > v.unit_lo;
        visits.push_back(v);
    }
    if (has_hot) {
        CK(cudaEventRecord(ev_fork, cold_s));
        CK(cudaStreamWaitEvent(hot_s, ev_fork, 0));
    }
    bool hot_chained = false;       // the previous operation on hot_s is a run launch of this part
    for (size_t vi = 0; vi < visits.size(); vi++) {
        const Visit& v = visits[vi];
        if (v.hi > v.lo) {          // cold records: full-grid Hogwild kernel
            a.recs = m.recs[m.rcur];
            a.first = v.lo;
            a.n = v.hi - v.lo;
            a.blk_start = m.block_off[v.cblk];
            a.blk_n = m.block_off[v.cblk + 1] - m.block_off[v.cblk];
            a.blk_id = (uint32_t)((size_t)m.g * (m.block_off.size() - 1) + v.cblk);
            CK(launch_sgd_update_hogwild(a, c.scatter, fast_arith, m.grid, h->min_windows, cold_s, &m.launches));
            m.update_launches++;
            *any_cold = true;
        }
        if (v.unit_hi > v.unit_lo) {   // runs: one sub-warp per run, q_i in registers
            a.recs = m.recs[m.rcur];
            a.first = 0;
            a.n = m.n_recs;
            if (m.counter_next >= m.n_counters * COUNTER_EPOCHS) return fail(MFSGD_E_STATE, "run launch counters exhausted");
            const int sub_warp_cap = (int)std::min<int64_t>(1 << 30, (int64_t)(m.u_hi - m.u_lo) / ((int64_t)h->sub_warp_div * h->mu));   // a fraction of the sub-stripe's users
            auto overlaps = [&](const Visit& w) {
                return chain_overlap && hot_launch_overlaps(c.k, w.unit_hi - w.unit_lo, m.unit_recs_cum[(size_t)w.unit_hi] - m.unit_recs_cum[(size_t)w.unit_lo],
                                           m.run_chunk, m.hot_grid, sub_warp_cap);
            };
            bool next_overlaps = false;    // will the next run launch of this chain overlap this one's tail?
            for (size_t vj = vi + 1; vj < visits.size(); vj++)
                if (visits[vj].unit_hi > visits[vj].unit_lo) {
                    next_overlaps = overlaps(visits[vj]);
                    break;
                }
            CK(launch_sgd_update_hot(a, m.d_units + v.unit_lo, v.unit_hi - v.unit_lo, m.d_counters + m.counter_next++, fast_arith,
                                     run_p_red(c), m.hot_grid, sub_warp_cap, hot_chained && overlaps(v), next_overlaps, hot_s, &m.launches));
            hot_chained = true;
            m.update_launches++;
            *any_hot = true;
        }
    }
    if (mark_cold) CK(cudaEventRecord(mark_cold, cold_s));
    if (mark_hot) CK(cudaEventRecord(mark_hot, has_hot ? hot_s : cold_s));
    if (has_hot) {
        CK(cudaEventRecord(ev_join, hot_s));
        CK(cudaStreamWaitEvent(cold_s, ev_join, 0));
    }
    return MFSGD_OK;
}

// Multi-process ring: (sse, n) of a held-out pass summed over the ranks (ncclAllReduce of two doubles on member 0's stream).
// Collective: every rank evaluates in the same epochs (same configuration), so every rank gets here together.
static int allreduce_sse(mfsgd_handle* h, double* sse, int64_t* n) {
    Member& m = h->members[0];
    CK(cudaSetDevice(m.device));
    if (!h->d_allreduce) CK(dev_alloc(&h->d_allreduce, 2));
    double v[2] = {*sse, (double)*n};
    CK(cudaMemcpyAsync(h->d_allreduce, v, sizeof(v), cudaMemcpyHostToDevice, m.stream));
    ncclResult_t r = g_nccl.AllReduce(h->d_allreduce, h->d_allreduce, 2, ncclFloat64, ncclSum, h->comm, m.stream);
    if (r != ncclSuccess) return fail(MFSGD_E_NCCL, "all-reduce of the held-out error: %s", g_nccl.GetErrorString(r));
    CK(cudaMemcpyAsync(v, h->d_allreduce, sizeof(v), cudaMemcpyDeviceToHost, m.stream));
    CK(cudaStreamSynchronize(m.stream));
    *sse = v[0];
    *n = (int64_t)(v[1] + 0.5);
    return MFSGD_OK;
}

static int train_impl(mfsgd_handle* h, int32_t epochs, mfsgd_epoch_stats* stats, float* err_trace) {
    if (!h) return fail(MFSGD_E_INVALID_ARG, "handle is null");
    if (epochs < 0) return fail(MFSGD_E_INVALID_ARG, "epochs < 0");
    if (!h->loaded) return fail(MFSGD_E_STATE, "no ratings loaded");
    if (!h->factors_ready) return fail(MFSGD_E_STATE, "factors are not initialised");
    const mfsgd_config& c = h->cfg;
    if (err_trace && c.mode != MFSGD_MODE_DETERMINISTIC) return fail(MFSGD_E_STATE, "error traces exist in DETERMINISTIC mode only");
    const bool time_kernels = (c.flags & MFSGD_FLAG_TIME_KERNELS) != 0;
    const bool fast_arith = !(c.flags & MFSGD_FLAG_EXACT_ARITH);
    struct TraceBuf {            // freed on every exit path
        float* p = nullptr;
        int device = 0;
        ~TraceBuf() {
            if (p) {
                cudaSetDevice(device);
                raw_free(p);
            }
        }
    } trace;
    if (err_trace && h->members[0].n_recs > 0) {
        trace.device = h->members[0].device;
        CK(cudaSetDevice(trace.device));
        CK(dev_alloc(&trace.p, (size_t)h->members[0].n_recs));
    }
    float* const d_trace = trace.p;
    int rc = MFSGD_OK;
    for (Member& m : h->members) {
        m.ev_used = 0;
        m.pending.clear();
        m.kev.clear();
    }
    // parts: item sub-shards launched separately -- always when the rotation is pipelined, else on request
    // (MFSGD_FLAG_SPLIT_SHARDS: a single GPU then replays the launch sizes of a larger ring)
    const bool pipelined = h->multi_process && h->G > 1 && h->mi > 1 && c.mode != MFSGD_MODE_DETERMINISTIC;
    const int parts = (pipelined || (c.flags & MFSGD_FLAG_SPLIT_SHARDS)) ? h->mi : 1;
    const int blocks_per_part = h->mi / parts;
    const bool lane_mode = parts > 1 && c.mode != MFSGD_MODE_DETERMINISTIC;
    if (pipelined && h->members[0].win.active && epochs > 0) {      // the flag values this call will write: one per sub-epoch
        Member& m0 = h->members[0];
        CK(cudaSetDevice(m0.device));
        if ((uint64_t)m0.win.sent[0] + (uint64_t)h->G * (uint64_t)epochs >= 0x7fffffffULL) return fail(MFSGD_E_STATE, "ring window: sequence numbers exhausted");
        if (h->ring_signal_copy) CKRC(ring_seq_table(m0, m0.win.sent[0] + 1, (uint32_t)h->G * (uint32_t)epochs));
    }
    int resolved = 0;   // epochs of this call whose stats are final
    // Epochs are enqueued back to back (no host sync in between) unless per-epoch evaluation is on;
    // their event records are resolved after the next synchronisation point.
    auto resolve = [&](int upto) -> int {
        for (Member& m : h->members) {
            CK(cudaSetDevice(m.device));
            CK(cudaStreamSynchronize(m.stream));
            CK(cudaStreamSynchronize(m.copy_stream));
        }
        for (int ep = resolved; ep < upto; ep++) {
            mfsgd_epoch_stats st{};
            st.heldout_rmse = std::numeric_limits<double>::quiet_NaN();
            for (Member& m : h->members) {
                const Member::EpochRec& er = m.pending[(size_t)ep];
                float ms = 0.f, sms = 0.f;
                CK(cudaEventElapsedTime(&ms, er.start, er.end));
                CK(cudaEventElapsedTime(&sms, er.start, er.shuffled));
                double kms = 0.0, cms = 0.0, hms = 0.0, xms = 0.0;
                for (int j = er.kbeg; j + 4 < er.kend; j += 5) {
                    float t = 0.f;
                    CK(cudaEventElapsedTime(&t, m.kev[(size_t)j], m.kev[(size_t)j + 1]));
                    kms += t;
                    if (!lane_mode) {     // lanes overlap parts and sub-epochs: only the span is meaningful there
                        CK(cudaEventElapsedTime(&t, m.kev[(size_t)j], m.kev[(size_t)j + 2]));
                        cms += t;
                        CK(cudaEventElapsedTime(&t, m.kev[(size_t)j], m.kev[(size_t)j + 3]));
                        hms += t;
                        CK(cudaEventElapsedTime(&t, m.kev[(size_t)j + 1], m.kev[(size_t)j + 4]));
                        xms += t;
                    }
                }
                st.cold_ms = std::max(st.cold_ms, cms);
                st.hot_ms = std::max(st.hot_ms, hms);
                st.exchange_ms = std::max(st.exchange_ms, xms);
                st.updates += m.n_recs;
                st.epoch_ms = std::max(st.epoch_ms, (double)ms);
                st.shuffle_ms = std::max(st.shuffle_ms, (double)sms);
                st.update_kernel_ms = std::max(st.update_kernel_ms, kms);
                st.update_launches = std::max(st.update_launches, er.update_launches);
                st.total_launches += er.launches;
            }
            if (stats) stats[ep] = st;
        }
        resolved = upto;
        return MFSGD_OK;
    };
    const double t_train0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    const bool have_heldout = !h->members[0].heldout.group_off.empty();
    if (h->es_patience > 0 && !have_heldout) return fail(MFSGD_E_STATE, "early stopping needs a held-out set (mfsgd_load_heldout / mfsgd_generate_synthetic)");
    const bool want_eval = (h->eval_every || h->es_patience > 0) && have_heldout;
    h->es_bad = 0;
    h->es_stopped = false;
    int ran = epochs;        // epochs this call actually runs (early stopping may cut it short)
    for (int ep = 0; ep < epochs && rc == MFSGD_OK; ep++) {
        for (Member& m : h->members) {
            m.launches = 0;
            m.update_launches = 0;
            Member::EpochRec er{};
            CK(cudaSetDevice(m.device));
            if (lane_mode) CKRC(ensure_lanes(m, h->n_lanes));
            CKRC(timing_event(m, &er.start));
            CKRC(timing_event(m, &er.shuffled));
            CKRC(timing_event(m, &er.end));
            er.kbeg = er.kend = (int)m.kev.size();
            m.pending.push_back(er);
            CK(cudaEventRecord(er.start, m.stream));
            bool gate = false;   // lanes must wait for something the main stream does at this epoch's start
            if (m.d_counters && ep % COUNTER_EPOCHS == 0) {
                if (lane_mode) CKRC(join_lanes(m));     // no launch of an earlier batch may still be claiming runs
                CK(cudaMemsetAsync(m.d_counters, 0, (size_t)m.n_counters * COUNTER_EPOCHS * sizeof(unsigned int), m.stream));
                m.counter_next = 0;
                gate = true;
            }
            const bool reshuffles = c.mode == MFSGD_MODE_DETERMINISTIC || (!(c.flags & MFSGD_FLAG_NO_SHUFFLE) && !h->virtual_shuffle);
            if (lane_mode && reshuffles) {
                CKRC(join_lanes(m));                    // the reshuffle rewrites the record buffer the lanes read
                gate = true;
            }
            if (!(c.flags & MFSGD_FLAG_NO_SHUFFLE) || c.mode == MFSGD_MODE_DETERMINISTIC) CKRC(shuffle_member(h, m, h->epoch, true));
            CK(cudaEventRecord(er.shuffled, m.stream));
            if (lane_mode && gate) {
                CK(cudaEventRecord(m.ev_epoch_go, m.stream));
                for (Lane& l : m.lanes) CK(cudaStreamWaitEvent(l.cold, m.ev_epoch_go, 0));
            }
        }
        for (int s = 0; s < h->G; s++) {
            for (Member& m : h->members) {
                CK(cudaSetDevice(m.device));
                const int grp = m.held_group;
                UpdateArgs a{};
                a.P = m.P;
                a.Q = m.Q[m.cur];
                a.k = c.k;
                a.u_base = m.u_lo;
                a.i_base = group_lo(h, grp);
                a.lr = h->lr_now;
                a.lambda = c.lambda;
                a.seed = c.seed;
                a.epoch = (uint32_t)h->epoch;
                a.virt = h->virtual_shuffle ? 1 : 0;
                a.BU = h->biases ? m.BU : nullptr;
                a.BI = h->biases ? m.BQ[m.cur] : nullptr;
                a.p_half = h->p_half ? 1 : 0;
                a.sm_limit = (pipelined && h->reserve_sms > 0 && h->reserve_sms < m.n_sms) ? m.n_sms - h->reserve_sms : 0;
                if (c.mode == MFSGD_MODE_DETERMINISTIC) {
                    a.recs = m.recs[0];
                    a.first = 0;
                    a.n = m.n_recs;
                    CK(launch_sgd_update_deterministic(a, d_trace, m.stream, &m.launches));
                    m.update_launches++;
                    if (d_trace)
                        CK(cudaMemcpyAsync(err_trace + (size_t)ep * m.n_recs, d_trace, (size_t)m.n_recs * 4, cudaMemcpyDeviceToHost, m.stream));
                    continue;
                }
                cudaEvent_t e0 = nullptr, e1 = nullptr, e_cold = nullptr, e_hot = nullptr, e_xchg = nullptr;
                if (time_kernels) {
                    CKRC(timing_event(m, &e0));
                    CKRC(timing_event(m, &e1));
                    CKRC(timing_event(m, &e_cold));
                    CKRC(timing_event(m, &e_hot));
                    CKRC(timing_event(m, &e_xchg));
                    CK(cudaEventRecord(e0, m.stream));
                }
                bool any_cold = false, any_hot = false;
                if (!lane_mode) {
                    // The cold (full-grid Hogwild) launches go to m.stream, the run launches to m.hot_stream: they touch the
                    // same P sub-stripes Hogwild-style and fill each other's tails. Fork here, join before the Q rotation.
                    const size_t ib_lo = (size_t)grp * h->mi, ib_hi = ib_lo + (size_t)h->mi;
                    // (the time marks are recorded on the two streams before they join: cold = last Hogwild launch done,
                    // hot = last run launch done, both measured from the sub-epoch's start)
                    CKRC(enqueue_visits(h, m, a, s, ib_lo, ib_hi, fast_arith, true, m.stream, m.hot_stream, m.ev_fork, m.ev_join, e_cold, e_hot,
                                        &any_cold, &any_hot));
                } else {
                    if (pipelined) CKRC(ensure_part_events(h, m));
                    for (int part = 0; part < parts; part++) {
                        Lane& l = m.lanes[(size_t)(part % h->n_lanes)];
                        if (pipelined && m.part_recv_pending[(size_t)part]) {      // this slice of the group has to have arrived
                            CKRC(wait_part(h, m, part, l.cold));
                            m.part_recv_pending[(size_t)part] = 0;
                        }
                        const size_t ib_lo = (size_t)grp * h->mi + (size_t)part * blocks_per_part, ib_hi = ib_lo + blocks_per_part;
                        CKRC(enqueue_visits(h, m, a, s, ib_lo, ib_hi, fast_arith, false, l.cold, l.hot, l.fork, l.join, nullptr, nullptr, &any_cold,
                                            &any_hot));
                        l.used = true;
                        if (pipelined) CKRC(rotate_part(h, m, part, l.cold));
                    }
                    if (pipelined) {          // all slices are on their way: the other buffer holds the next group
                        m.cur ^= 1;
                        m.held_group = (m.held_group + 1) % h->G;
                    }
                    if (time_kernels) {       // span of this sub-epoch's launches on the coordinating stream (does not hold the lanes back)
                        for (Lane& l : m.lanes) {
                            CK(cudaEventRecord(l.done, l.cold));
                            CK(cudaStreamWaitEvent(m.stream, l.done, 0));
                        }
                        CK(cudaEventRecord(e_cold, m.stream));
                        CK(cudaEventRecord(e_hot, m.stream));
                    }
                }
                if (time_kernels) {
                    CK(cudaEventRecord(e1, m.stream));
                    m.kev.push_back(e0);
                    m.kev.push_back(e1);
                    m.kev.push_back(e_cold);
                    m.kev.push_back(e_hot);
                    m.kev.push_back(e_xchg);
                }
            }
            if (!pipelined) {
                const bool around = lane_mode && h->G > 1;     // whole-group rotation between laned sub-epochs: lanes -> main -> lanes
                if (around)
                    for (Member& m : h->members) {
                        CK(cudaSetDevice(m.device));
                        CKRC(join_lanes(m));
                    }
                CKRC(rotate_q(h));
                if (around)
                    for (Member& m : h->members) {
                        CK(cudaSetDevice(m.device));
                        CK(cudaEventRecord(m.ev_epoch_go, m.stream));
                        for (Lane& l : m.lanes) CK(cudaStreamWaitEvent(l.cold, m.ev_epoch_go, 0));
                    }
            }
            if (time_kernels && c.mode != MFSGD_MODE_DETERMINISTIC)
                for (Member& m : h->members) {
                    CK(cudaSetDevice(m.device));
                    CK(cudaEventRecord(m.kev.back(), m.stream));     // e_xchg of this sub-epoch
                }
        }
        for (Member& m : h->members) {
            Member::EpochRec& er = m.pending.back();
            CK(cudaSetDevice(m.device));
            if (lane_mode) CKRC(join_lanes(m));        // the main stream only observes: the lanes are not held back by this
            // Q is home again once the last slices have arrived. Between back-to-back epochs of one call that wait is left
            // to the next epoch's first launches (per slice), so the pipeline keeps running across the epoch boundary.
            const bool drain = (ep + 1 == epochs) || want_eval;
            for (size_t part = 0; drain && part < m.part_recv_pending.size(); part++)
                if (m.part_recv_pending[part]) {
                    CKRC(wait_part(h, m, (int)part, m.stream));
                    m.part_recv_pending[part] = 0;
                }
            if (drain && pipelined && m.win.active) {
                // ring window: the call may only return (and the buffers be reused by anything else) once our own hand-overs
                // have left and every write a neighbour addresses to this window has landed -- the arrivals above and the credits
                CK(cudaEventRecord(m.ev_sent, m.copy_stream));
                CK(cudaStreamWaitEvent(m.stream, m.ev_sent, 0));
                for (int part = 0; part < parts; part++)
                    if (m.win.sent[(size_t)part] > 0) CKRC(ring_wait(h, m, m.stream, ring_credit_off(part), m.win.sent[(size_t)part]));
            }
            CK(cudaEventRecord(er.end, m.stream));
            er.kend = (int)m.kev.size();
            er.launches = m.launches;
            er.update_launches = m.update_launches;
        }
        h->epoch++;
        h->lr_now = h->lr_now * h->lr_decay;          // binary32 multiply, epoch by epoch: the same value on every host
        if (want_eval) {
            CKRC(resolve(ep + 1));
            double sse = 0.0;
            int64_t n = 0;
            rc = rmse_pass(h, true, &sse, &n);
            if (rc == MFSGD_OK && h->multi_process) rc = allreduce_sse(h, &sse, &n);   // every rank sees the ring's RMSE
            const double v = n > 0 ? std::sqrt(sse / (double)n) : 0.0;
            if (rc == MFSGD_OK && stats) stats[ep].heldout_rmse = v;
            if (rc == MFSGD_OK && h->es_patience > 0) {     // MatrixFactorizationSGD.java factorizeEarlyStop: the rule, verbatim
                if (v < h->es_best * (1.0 - (double)h->es_min_delta)) {
                    h->es_best = v;
                    h->es_bad = 0;
                } else if (++h->es_bad >= h->es_patience) {
                    h->es_stopped = true;
                    ran = ep + 1;
                    break;
                }
            }
        }
    }
    const double t_enq = trace_on() ? std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() : 0.0;
    if (rc == MFSGD_OK) rc = resolve(ran);
    if (trace_on()) {
        const double t_end = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        fprintf(stderr, "[mfsgd] train: %d epochs, host enqueue %.3f ms per epoch, enqueue + drain %.3f ms per epoch\n", ran,
                (t_enq - t_train0) * 1e3 / std::max(1, ran), (t_end - t_train0) * 1e3 / std::max(1, ran));
    }
    if (rc == MFSGD_OK && stats)
        for (int ep = ran; ep < epochs; ep++) {       // epochs the early-stopping rule skipped: updates == 0
            stats[ep] = mfsgd_epoch_stats{};
            stats[ep].heldout_rmse = std::numeric_limits<double>::quiet_NaN();
        }
    return rc;
}

extern "C" int mfsgd_train(mfsgd_handle* h, int32_t epochs, mfsgd_epoch_stats* stats) {
    return guarded([&]() { return train_impl(h, epochs, stats, nullptr); });
}
extern "C" int mfsgd_train_traced(mfsgd_handle* h, int32_t epochs, mfsgd_epoch_stats* stats, float* err_trace) {
    if (!err_trace) return fail(MFSGD_E_INVALID_ARG, "err_trace is null");
    return guarded([&]() { return train_impl(h, epochs, stats, err_trace); });
}
extern "C" int mfsgd_get_progress(mfsgd_handle* h, int32_t* epochs_done, float* next_lr, int32_t* stopped_early) {
    if (!h) return fail(MFSGD_E_INVALID_ARG, "handle is null");
    if (epochs_done) *epochs_done = (int32_t)h->epoch;
    if (next_lr) *next_lr = h->lr_now;
    if (stopped_early) *stopped_early = h->es_stopped ? 1 : 0;
    return MFSGD_OK;
}
extern "C" int mfsgd_set_eval_every_epoch(mfsgd_handle* h, int32_t on) {
    if (!h) return fail(MFSGD_E_INVALID_ARG, "handle is null");
    h->eval_every = on ? 1 : 0;
    return MFSGD_OK;
}

// ------------------------------------------------------------------------------------------------
// subsystem (3) driver: RMSE over bucketed record sets, Q groups rotating once round the ring
// ------------------------------------------------------------------------------------------------
static int rmse_sets(mfsgd_handle* h, std::vector<EvalSet*>& sets, bool train_set, double* sse_out, int64_t* n_out) {
    const int k = h->cfg.k;
    for (Member& m : h->members) {
        CK(cudaSetDevice(m.device));
        CK(cudaMemsetAsync(m.d_sse, 0, sizeof(double), m.stream));
    }
    for (int s = 0; s < h->G; s++) {
        for (size_t j = 0; j < h->members.size(); j++) {
            Member& m = h->members[j];
            CK(cudaSetDevice(m.device));
            const int grp = m.held_group;
            if (train_set) {
                for (int sa = 0; sa < h->mu; sa++) {
                    const int64_t lo = m.block_off[(size_t)sa * h->IB + (size_t)grp * h->mi];
                    const int64_t hi = m.block_off[(size_t)sa * h->IB + (size_t)(grp + 1) * h->mi];
                    CK(launch_rmse_sse(m.recs[m.rcur] + lo, hi - lo, m.P, h->p_half, m.Q[m.cur], m.BU, m.BQ[m.cur], k, m.u_lo, group_lo(h, grp), m.d_scratch,
                                       m.d_sse, m.n_sms, m.stream, &m.launches));
                    if (h->H > 0) {   // the hot buckets of (sa, grp) are contiguous too
                        const size_t hb = (size_t)h->mu * h->IB + (size_t)sa * h->H;
                        const int64_t hlo = m.block_off[hb + (size_t)h->hot_block_lo[(size_t)grp * h->mi]];
                        const int64_t hhi = m.block_off[hb + (size_t)h->hot_block_lo[(size_t)(grp + 1) * h->mi]];
                        CK(launch_rmse_sse(m.recs[m.rcur] + hlo, hhi - hlo, m.P, h->p_half, m.Q[m.cur], m.BU, m.BQ[m.cur], k, m.u_lo, group_lo(h, grp), m.d_scratch,
                                           m.d_sse, m.n_sms, m.stream, &m.launches));
                    }
                }
            } else {
                EvalSet* e = sets[j];
                const int64_t lo = e->group_off[(size_t)grp], hi = e->group_off[(size_t)grp + 1];
                CK(launch_rmse_sse(e->recs + lo, hi - lo, m.P, h->p_half, m.Q[m.cur], m.BU, m.BQ[m.cur], k, m.u_lo, group_lo(h, grp), m.d_scratch, m.d_sse,
                                   m.n_sms, m.stream, &m.launches));
            }
        }
        CKRC(rotate_q(h));
    }
    double sse = 0.0;
    int64_t n = 0;
    for (size_t j = 0; j < h->members.size(); j++) {
        Member& m = h->members[j];
        double part = 0.0;
        CK(cudaSetDevice(m.device));
        CK(cudaMemcpyAsync(&part, m.d_sse, sizeof(double), cudaMemcpyDeviceToHost, m.stream));
        CK(cudaStreamSynchronize(m.stream));
        CK(cudaStreamSynchronize(m.copy_stream));
        sse += part;
        n += train_set ? m.n_recs : sets[j]->n;
    }
    *sse_out = sse;
    *n_out = n;
    return MFSGD_OK;
}

static int rmse_pass(mfsgd_handle* h, bool heldout, double* sse_out, int64_t* n_out) {
    std::vector<EvalSet*> sets;
    if (heldout)
        for (Member& m : h->members) sets.push_back(&m.heldout);
    return rmse_sets(h, sets, !heldout, sse_out, n_out);
}

static int finish_rmse(double sse, int64_t n, double* rmse_out, double* sse_out, int64_t* n_out) {
    if (rmse_out) *rmse_out = n > 0 ? std::sqrt(sse / (double)n) : 0.0;
    if (sse_out) *sse_out = sse;
    if (n_out) *n_out = n;
    return MFSGD_OK;
}

static int mfsgd_rmse_heldout_body(mfsgd_handle* h, double* rmse_out, double* sse_out, int64_t* n_out);
extern "C" int mfsgd_rmse_heldout(mfsgd_handle* h, double* rmse_out, double* sse_out, int64_t* n_out) {
    return guarded([&]() { return mfsgd_rmse_heldout_body(h, rmse_out, sse_out, n_out); });
}
static int mfsgd_rmse_heldout_body(mfsgd_handle* h, double* rmse_out, double* sse_out, int64_t* n_out) {
    if (!h) return fail(MFSGD_E_INVALID_ARG, "handle is null");
    if (!h->factors_ready) return fail(MFSGD_E_STATE, "factors are not initialised");
    for (Member& m : h->members)
        if (m.heldout.group_off.empty()) return fail(MFSGD_E_STATE, "no held-out set loaded");
    double sse = 0.0;
    int64_t n = 0;
    CKRC(rmse_pass(h, true, &sse, &n));
    return finish_rmse(sse, n, rmse_out, sse_out, n_out);
}

static int mfsgd_rmse_train_body(mfsgd_handle* h, double* rmse_out, double* sse_out, int64_t* n_out);
extern "C" int mfsgd_rmse_train(mfsgd_handle* h, double* rmse_out, double* sse_out, int64_t* n_out) {
    return guarded([&]() { return mfsgd_rmse_train_body(h, rmse_out, sse_out, n_out); });
}
static int mfsgd_rmse_train_body(mfsgd_handle* h, double* rmse_out, double* sse_out, int64_t* n_out) {
    if (!h) return fail(MFSGD_E_INVALID_ARG, "handle is null");
    if (!h->factors_ready || !h->loaded) return fail(MFSGD_E_STATE, "need loaded ratings and factors");
    if (h->cfg.mode == MFSGD_MODE_DETERMINISTIC) {
        // records live in caller order (recs_orig); evaluate them as one block
        Member& m = h->members[0];
        CK(cudaSetDevice(m.device));
        CK(cudaMemsetAsync(m.d_sse, 0, sizeof(double), m.stream));
        CK(launch_rmse_sse(m.recs_orig, m.n_recs, m.P, h->p_half, m.Q[m.cur], m.BU, m.BQ[m.cur], h->cfg.k, m.u_lo, 0, m.d_scratch, m.d_sse, m.n_sms, m.stream, &m.launches));
        double sse = 0.0;
        CK(cudaMemcpyAsync(&sse, m.d_sse, sizeof(double), cudaMemcpyDeviceToHost, m.stream));
        CK(cudaStreamSynchronize(m.stream));
        return finish_rmse(sse, m.n_recs, rmse_out, sse_out, n_out);
    }
    double sse = 0.0;
    int64_t n = 0;
    CKRC(rmse_pass(h, false, &sse, &n));
    return finish_rmse(sse, n, rmse_out, sse_out, n_out);
}

static int mfsgd_rmse_body(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n, double* rmse_out);
extern "C" int mfsgd_rmse(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n, double* rmse_out) {
    return guarded([&]() { return mfsgd_rmse_body(h, users, items, ratings, n, rmse_out); });
}
static int mfsgd_rmse_body(mfsgd_handle* h, const int32_t* users, const int32_t* items, const float* ratings, int64_t n, double* rmse_out) {
    if (!h || !rmse_out) return fail(MFSGD_E_INVALID_ARG, "null argument");
    if (n < 0 || (n > 0 && (!users || !items || !ratings))) return fail(MFSGD_E_INVALID_ARG, "null triplet arrays or negative n");
    if (!h->loaded || !h->factors_ready) return fail(MFSGD_E_STATE, "need loaded ratings and factors");
    std::vector<EvalSet> tmp(h->members.size());
    std::vector<EvalSet*> sets;
    int rc = MFSGD_OK;
    for (size_t j = 0; j < h->members.size() && rc == MFSGD_OK; j++) {
        rc = build_eval_set(h, h->members[j], users, items, ratings, n, tmp[j]);
        sets.push_back(&tmp[j]);
    }
    double sse = 0.0;
    int64_t cnt = 0;
    if (rc == MFSGD_OK) rc = rmse_sets(h, sets, false, &sse, &cnt);
    for (size_t j = 0; j < tmp.size(); j++) {
        cudaSetDevice(h->members[j].device);
        free_eval(tmp[j]);
    }
    if (rc != MFSGD_OK) return rc;
    *rmse_out = cnt > 0 ? std::sqrt(sse / (double)cnt) : 0.0;
    return MFSGD_OK;
}

// ------------------------------------------------------------------------------------------------
// one-shot entry point (MatrixFactorizationSGD.java:109 factorize)
// ------------------------------------------------------------------------------------------------
static int mfsgd_factorize_body(const int32_t* users, const int32_t* items, const float* ratings, int64_t n,
                               const mfsgd_config* cfg, int32_t epochs, float* P_out, float* Q_out);
extern "C" int mfsgd_factorize(const int32_t* users, const int32_t* items, const float* ratings, int64_t n,
                               const mfsgd_config* cfg, int32_t epochs, float* P_out, float* Q_out) {
    return guarded([&]() { return mfsgd_factorize_body(users, items, ratings, n, cfg, epochs, P_out, Q_out); });
}
static int mfsgd_factorize_body(const int32_t* users, const int32_t* items, const float* ratings, int64_t n,
                               const mfsgd_config* cfg, int32_t epochs, float* P_out, float* Q_out) {
    if (!P_out || !Q_out) return fail(MFSGD_E_INVALID_ARG, "output arrays are null");
    if (epochs < 0) return fail(MFSGD_E_INVALID_ARG, "epochs < 0");
    const bool trace = getenv("MFSGD_TRACE") != nullptr;   // phase timings on stderr (diagnostic aid)
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now(), t1;
    mfsgd_handle* h = nullptr;
    CKRC(mfsgd_create(cfg, &h));
    if (trace) { t1 = now(); fprintf(stderr, "[mfsgd] create %.1f ms\n", (t1 - t0) * 1e3); t0 = t1; }
    int rc = mfsgd_load_ratings(h, users, items, ratings, n);
    if (trace) { t1 = now(); fprintf(stderr, "[mfsgd] load_ratings %.1f ms\n", (t1 - t0) * 1e3); t0 = t1; }
    if (rc == MFSGD_OK) rc = mfsgd_init_factors(h);
    if (trace) { t1 = now(); fprintf(stderr, "[mfsgd] init_factors %.1f ms\n", (t1 - t0) * 1e3); t0 = t1; }
    if (rc == MFSGD_OK) rc = mfsgd_train(h, epochs, nullptr);
    if (trace) { t1 = now(); fprintf(stderr, "[mfsgd] train %.1f ms\n", (t1 - t0) * 1e3); t0 = t1; }
    if (rc == MFSGD_OK) rc = mfsgd_get_factors(h, P_out, Q_out);
    if (trace) { t1 = now(); fprintf(stderr, "[mfsgd] get_factors %.1f ms\n", (t1 - t0) * 1e3); t0 = t1; }
    mfsgd_destroy(h);
    if (trace) { t1 = now(); fprintf(stderr, "[mfsgd] destroy %.1f ms\n", (t1 - t0) * 1e3); t0 = t1; }
    return rc;
}

// ------------------------------------------------------------------------------------------------
// introspection + test hooks
// ------------------------------------------------------------------------------------------------
// Test hook: the layout planner (run_plan.hpp) for a configuration and a data size. Host-only.
static int mfsgd_plan_layout_body(const mfsgd_config* cfg, int64_t l2_bytes, int64_t member_records, int32_t member_users,
                                 int64_t run_records, int32_t resident_ctas, int32_t* stripes, int32_t* shards, int32_t* rounds,
                                 int32_t* run_length);
extern "C" int mfsgd_plan_layout(const mfsgd_config* cfg, int64_t l2_bytes, int64_t member_records, int32_t member_users,
                                 int64_t run_records, int32_t resident_ctas, int32_t* stripes, int32_t* shards, int32_t* rounds,
                                 int32_t* run_length) {
    return guarded([&]() { return mfsgd_plan_layout_body(cfg, l2_bytes, member_records, member_users, run_records, resident_ctas, stripes, shards, rounds, run_length); });
}
static int mfsgd_plan_layout_body(const mfsgd_config* cfg, int64_t l2_bytes, int64_t member_records, int32_t member_users,
                                 int64_t run_records, int32_t resident_ctas, int32_t* stripes, int32_t* shards, int32_t* rounds,
                                 int32_t* run_length) {
    if (!cfg || !stripes || !shards || !rounds || !run_length) return fail(MFSGD_E_INVALID_ARG, "null argument");
    int rc = validate_config(cfg);
    if (rc != MFSGD_OK) return rc;
    const int G = cfg->mode == MFSGD_MODE_DSGD ? cfg->n_gpus : 1;
    const Blocking b = plan_blocking(cfg->n_users, cfg->n_items, cfg->k, G, cfg->mode, cfg->stripes_per_gpu, cfg->shards_per_gpu,
                                     cfg->world_size > 1, (double)l2_bytes);
    *stripes = b.mu;
    *shards = b.mi;
    *rounds = plan_rounds(cfg->rounds, cfg->mode, G, b.mu, member_records, member_users);
    const bool laned = (cfg->world_size > 1 && G > 1 && b.mi > 1) || ((cfg->flags & MFSGD_FLAG_SPLIT_SHARDS) && b.mi > 1);
    *run_length = plan_run_length(cfg->hot_chunk, laned ? 2 : 1, b.mu, *rounds, G * b.mi, run_records, resident_ctas, 32 / run_kernel_lanes(cfg->k));
    return MFSGD_OK;
}

// Test hook: the run planner (run_plan.hpp) on caller-provided bucket offsets. Host-only.
static int mfsgd_plan_runs_body(const int64_t* block_off, int32_t stripes, int32_t n_hot, int32_t item_blocks, const int32_t* hot_block_lo,
                               const int32_t* hot_items, int32_t rounds, int32_t chunk, uint64_t seed, int32_t member,
                               float merge_boost, int64_t* unit_start, int32_t* unit_count, int32_t* unit_item, float* unit_weight,
                               int64_t* n_units, int32_t* visit_units);
extern "C" int mfsgd_plan_runs(const int64_t* block_off, int32_t stripes, int32_t n_hot, int32_t item_blocks, const int32_t* hot_block_lo,
                               const int32_t* hot_items, int32_t rounds, int32_t chunk, uint64_t seed, int32_t member,
                               float merge_boost, int64_t* unit_start, int32_t* unit_count, int32_t* unit_item, float* unit_weight,
                               int64_t* n_units, int32_t* visit_units) {
    return guarded([&]() { return mfsgd_plan_runs_body(block_off, stripes, n_hot, item_blocks, hot_block_lo, hot_items, rounds, chunk, seed, member, merge_boost, unit_start, unit_count, unit_item, unit_weight, n_units, visit_units); });
}
static int mfsgd_plan_runs_body(const int64_t* block_off, int32_t stripes, int32_t n_hot, int32_t item_blocks, const int32_t* hot_block_lo,
                               const int32_t* hot_items, int32_t rounds, int32_t chunk, uint64_t seed, int32_t member,
                               float merge_boost, int64_t* unit_start, int32_t* unit_count, int32_t* unit_item, float* unit_weight,
                               int64_t* n_units, int32_t* visit_units) {
    if (!block_off || !hot_block_lo || (n_hot > 0 && !hot_items) || !n_units || !visit_units) return fail(MFSGD_E_INVALID_ARG, "null argument");
    if (stripes < 1 || n_hot < 0 || item_blocks < 1 || rounds < 1 || chunk < 1 || chunk > 65536 || member < 0 || !(merge_boost >= 0.f) ||
        merge_boost >= 2.f)
        return fail(MFSGD_E_INVALID_ARG, "bad shape");
    RunPlanArgs pa{};
    pa.block_off = block_off;
    pa.n_blocks = (size_t)stripes * ((size_t)item_blocks + (size_t)n_hot);
    pa.mu = stripes; pa.H = n_hot; pa.IB = item_blocks; pa.rounds = rounds; pa.chunk = chunk;
    pa.member = member;
    pa.seed = seed;
    pa.boost = merge_boost > 0.f ? merge_boost : DEFAULT_MERGE_BOOST;
    pa.hot_block_lo = hot_block_lo;
    pa.hot_items = hot_items;
    std::vector<HotUnit> units;
    std::vector<int> visits;
    plan_runs(pa, units, visits, plan_threads_env());
    const int64_t cap = *n_units;
    *n_units = (int64_t)units.size();
    for (size_t v = 0; v < visits.size(); v++) visit_units[v] = visits[v];
    if ((int64_t)units.size() > cap) return fail(MFSGD_E_INVALID_ARG, "%zu runs do not fit the %lld provided slots", units.size(), (long long)cap);
    for (size_t j = 0; j < units.size(); j++) {
        if (unit_start) unit_start[j] = units[j].start;
        if (unit_count) unit_count[j] = units[j].count;
        if (unit_item) unit_item[j] = units[j].item;
        if (unit_weight) unit_weight[j] = units[j].weight;
    }
    return MFSGD_OK;
}

static int mfsgd_get_layout_info_body(mfsgd_handle* h, mfsgd_layout_info* out);
extern "C" int mfsgd_get_layout_info(mfsgd_handle* h, mfsgd_layout_info* out) {
    return guarded([&]() { return mfsgd_get_layout_info_body(h, out); });
}
static int mfsgd_get_layout_info_body(mfsgd_handle* h, mfsgd_layout_info* out) {
    if (!h || !out) return fail(MFSGD_E_INVALID_ARG, "null argument");
    if (!h->loaded) return fail(MFSGD_E_STATE, "no ratings loaded");
    memset(out, 0, sizeof(*out));
    out->n_gpus = h->G;
    out->stripes_per_gpu = h->mu;
    out->shards_per_gpu = h->mi;
    out->user_blocks = h->UB;
    out->item_blocks = h->IB;
    for (Member& m : h->members) {
        out->n_train_local += m.n_recs;
        out->n_heldout_local += m.heldout.n;
    }
    out->n_train_total = h->n_train_total;
    out->rounds = h->rounds;
    out->n_hot_items = h->H;
    out->n_heavy_users = h->n_heavy;
    out->run_length = h->members[0].run_chunk;
    return MFSGD_OK;
}

static int mfsgd_get_bounds_body(mfsgd_handle* h, int32_t* user_bounds, int32_t* item_bounds);
extern "C" int mfsgd_get_bounds(mfsgd_handle* h, int32_t* user_bounds, int32_t* item_bounds) {
    return guarded([&]() { return mfsgd_get_bounds_body(h, user_bounds, item_bounds); });
}
static int mfsgd_get_bounds_body(mfsgd_handle* h, int32_t* user_bounds, int32_t* item_bounds) {
    if (!h || !user_bounds || !item_bounds) return fail(MFSGD_E_INVALID_ARG, "null argument");
    if (!h->loaded) return fail(MFSGD_E_STATE, "no ratings loaded");
    memcpy(user_bounds, h->user_bounds.data(), h->user_bounds.size() * 4);
    memcpy(item_bounds, h->item_bounds.data(), h->item_bounds.size() * 4);
    return MFSGD_OK;
}

static int mfsgd_get_records_body(mfsgd_handle* h, int32_t member, int32_t* recs, int64_t* block_offsets, int64_t* n);
extern "C" int mfsgd_get_records(mfsgd_handle* h, int32_t member, int32_t* recs, int64_t* block_offsets, int64_t* n) {
    return guarded([&]() { return mfsgd_get_records_body(h, member, recs, block_offsets, n); });
}
static int mfsgd_get_records_body(mfsgd_handle* h, int32_t member, int32_t* recs, int64_t* block_offsets, int64_t* n) {
    if (!h || !n) return fail(MFSGD_E_INVALID_ARG, "null argument");
    if (!h->loaded) return fail(MFSGD_E_STATE, "no ratings loaded");
    if (member < 0 || member >= (int)h->members.size()) return fail(MFSGD_E_INVALID_ARG, "bad member index");
    Member& m = h->members[(size_t)member];
    *n = m.n_recs;
    if (block_offsets) memcpy(block_offsets, m.block_off.data(), m.block_off.size() * 8);
    if (recs && m.n_recs > 0) {
        CK(cudaSetDevice(m.device));
        CK(cudaStreamSynchronize(m.stream));
        CK(cudaMemcpy(recs, m.recs[m.rcur], (size_t)m.n_recs * sizeof(Rec), cudaMemcpyDeviceToHost));
    }
    return MFSGD_OK;
}

static int mfsgd_shuffle_once_body(mfsgd_handle* h, int32_t epoch);
extern "C" int mfsgd_shuffle_once(mfsgd_handle* h, int32_t epoch) {
    return guarded([&]() { return mfsgd_shuffle_once_body(h, epoch); });
}
static int mfsgd_shuffle_once_body(mfsgd_handle* h, int32_t epoch) {
    if (!h) return fail(MFSGD_E_INVALID_ARG, "handle is null");
    if (!h->loaded) return fail(MFSGD_E_STATE, "no ratings loaded");
    for (Member& m : h->members) {
        CKRC(shuffle_member(h, m, epoch, false));
        CK(cudaStreamSynchronize(m.stream));
    }
    return MFSGD_OK;
}

static int mfsgd_apply_updates_forced_body(int32_t device, int32_t k, float lr, float lambda, int64_t n, const float* pre_p,
                                          const float* pre_q, const float* r, float* post_p, float* post_q, float* err);
extern "C" int mfsgd_apply_updates_forced(int32_t device, int32_t k, float lr, float lambda, int64_t n, const float* pre_p,
                                          const float* pre_q, const float* r, float* post_p, float* post_q, float* err) {
    return guarded([&]() { return mfsgd_apply_updates_forced_body(device, k, lr, lambda, n, pre_p, pre_q, r, post_p, post_q, err); });
}
static int mfsgd_apply_updates_forced_body(int32_t device, int32_t k, float lr, float lambda, int64_t n, const float* pre_p,
                                          const float* pre_q, const float* r, float* post_p, float* post_q, float* err) {
    if (!rank_supported(k)) return fail(MFSGD_E_INVALID_ARG, "k=%d unsupported", k);
    if (n < 0 || (n > 0 && (!pre_p || !pre_q || !r || !post_p || !post_q || !err))) return fail(MFSGD_E_INVALID_ARG, "null argument");
    if (n == 0) return MFSGD_OK;
    CK(cudaSetDevice(device));
    float *dp = nullptr, *dq = nullptr, *dr = nullptr, *op = nullptr, *oq = nullptr, *de = nullptr;
    const size_t rows = (size_t)n * k;
    cudaError_t e = dev_alloc(&dp, rows);
    if (e == cudaSuccess) e = dev_alloc(&dq, rows);
    if (e == cudaSuccess) e = dev_alloc(&dr, (size_t)n);
    if (e == cudaSuccess) e = dev_alloc(&op, rows);
    if (e == cudaSuccess) e = dev_alloc(&oq, rows);
    if (e == cudaSuccess) e = dev_alloc(&de, (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpy(dp, pre_p, rows * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dq, pre_q, rows * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dr, r, (size_t)n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_sgd_update_forced(k, lr, lambda, n, dp, dq, dr, op, oq, de, nullptr);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(post_p, op, rows * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(post_q, oq, rows * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(err, de, (size_t)n * 4, cudaMemcpyDeviceToHost);
    raw_free(dp); raw_free(dq); raw_free(dr); raw_free(op); raw_free(oq); raw_free(de);
    CK(e);
    return MFSGD_OK;
}

static int mfsgd_generate_to_host_body(int32_t device, const mfsgd_synth_params* sp, int32_t n_users, int32_t n_items,
                                      int64_t start, int64_t count, int32_t* users, int32_t* items, float* ratings, uint8_t* held);
extern "C" int mfsgd_generate_to_host(int32_t device, const mfsgd_synth_params* sp, int32_t n_users, int32_t n_items,
                                      int64_t start, int64_t count, int32_t* users, int32_t* items, float* ratings, uint8_t* held) {
    return guarded([&]() { return mfsgd_generate_to_host_body(device, sp, n_users, n_items, start, count, users, items, ratings, held); });
}
static int mfsgd_generate_to_host_body(int32_t device, const mfsgd_synth_params* sp, int32_t n_users, int32_t n_items,
                                      int64_t start, int64_t count, int32_t* users, int32_t* items, float* ratings, uint8_t* held) {
    if (!sp || count < 0 || start < 0 || n_users <= 0 || n_items <= 0) return fail(MFSGD_E_INVALID_ARG, "bad arguments");
    if (count > 0 && (!users || !items || !ratings || !held)) return fail(MFSGD_E_INVALID_ARG, "null output arrays");
    if (count == 0) return MFSGD_OK;
    CK(cudaSetDevice(device));
    SynthArgs s{};
    s.seed = sp->seed; s.n_users = n_users; s.n_items = n_items;
    s.l2au = sp->log2_alpha_user; s.l2ai = sp->log2_alpha_item; s.cu = sp->c_user; s.ci = sp->c_item;
    s.amplitude = sp->planted_amplitude > 0.f ? sp->planted_amplitude : 0.8660254f;
    s.noise_scale = sp->noise_scale > 0.f ? sp->noise_scale : 0.5f;
    int32_t *du = nullptr, *di = nullptr;
    float* dr = nullptr;
    uint8_t* dh = nullptr;
    cudaError_t e = dev_alloc(&du, (size_t)count);
    if (e == cudaSuccess) e = dev_alloc(&di, (size_t)count);
    if (e == cudaSuccess) e = dev_alloc(&dr, (size_t)count);
    if (e == cudaSuccess) e = dev_alloc(&dh, (size_t)count);
    if (e == cudaSuccess) e = launch_generate(s, start, count, du, di, dr, dh, nullptr, nullptr);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(users, du, (size_t)count * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(items, di, (size_t)count * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(ratings, dr, (size_t)count * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(held, dh, (size_t)count, cudaMemcpyDeviceToHost);
    raw_free(du); raw_free(di); raw_free(dr); raw_free(dh);
    CK(e);
    return MFSGD_OK;
}
