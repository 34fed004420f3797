// kernels_hot.cu -- subsystem (2b): the run kernel of the SGD update path (sm_100a), the dominant kernel.
//
// Update rule = baseline/java/MatrixFactorizationSGD.java:89-105 (sgdUpdate). A unit is a run of records that all
// rate one item. A LANES-wide sub-warp walks the run strictly in order with q_i in its registers: gather p_u,
// dot, scatter p_u, update q_i -- the item row costs no L2 traffic and sees no concurrent writer inside the run.
// At the end the run's net change is merged into Q scaled by unit.weight (model averaging over the item's
// concurrent runs); a run that is alone on its item in the launch (weight 1) stores q_i outright, which makes
// the path exactly sequential -- unless the launch overlaps its predecessor (programmatic dependent launch): then a
// run of the same item from the previous visit's tail may still be merging, and the weight-1 run adds its net
// change with red.global.add as well instead of overwriting the row (always_add).
//
// Geometry: LANES lanes per rating, each holding VEC <= 4 float4 chunks of the row (chunk c belongs to lane c % LANES,
// so every load/store instruction moves >= 128 contiguous bytes per sub-warp = full sectors); the warp's 32/LANES
// sub-warps walk as many runs side by side, each strictly sequentially. Default: one chunk per lane up to k = 128
// (one warp per rating at k = 128), never fewer than 8 lanes. Narrower sub-warps (8 lanes x 4 chunks at k = 128:
// 3 shuffles per dot product serving 4 ratings, 28 instead of 43 warp instructions per update) were measured and
// run no faster: the launch is bound by L2 sector throughput (profiles/r01_experiments.md section 7).
//
// Pipeline: the run's (u, r) pairs are staged in shared memory by cp.async two LANES-record tiles ahead (ring of
// 4 tiles per sub-warp; under the virtual reshuffle position j of a bucket reads record bucket_start + perm(j));
// the p_u gathers go to a D-slot register ring, D-1 steps ahead, and the pipeline runs through the whole run.
// Steps are unrolled by D, so ring slots are compile-time; a tile (LANES steps, a multiple of D) always opens at
// the top of an unrolled block. Units are claimed from a per-launch counter (32/LANES at a time), longest first.
#include <cstdlib>

#include "kernels.cuh"
#include "update_math.cuh"

namespace mfsgd {

namespace {

__device__ __forceinline__ void cp_async_4_stream(void* smem_dst, const void* gsrc, uint64_t pol) {
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// Programmatic dependent launch: lets the next launch on the stream (if it asked for programmatic stream
// serialisation) become resident as this grid's CTAs retire, instead of after the grid has drained. The update
// launches of consecutive visits only ever meet Hogwild-style, so nothing waits on the other side ...
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// ... but a grid must not COMPLETE before the grid it was allowed to overtake: the stream's later operations (events,
// the next sub-epoch, the Q rotation) take this grid's completion for the completion of everything before it.
// Every thread therefore waits for the prerequisite grid as its last action (a no-op without the launch attribute).
__device__ __forceinline__ void pdl_wait_prerequisites() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// PH: P rows kept as binary16 (mfsgd_config.p_storage; common.cuh "mixed-precision factor storage"): a.P then points at binary16
// rows, a lane's chunk is 8 bytes, the ring holds the raw chunks and widens them at use, the scatter narrows with stochastic
// rounding (heavy users: the narrowed difference, one f16x4 red).
template <int LANES, int VEC, bool FULL, bool FAST, int D, bool PRED, bool BIAS, bool PH>
__global__ void __launch_bounds__(256, VEC >= 3 ? 2 : 3) sgd_update_runs_kernel(UpdateArgs a, const HotUnit* __restrict__ units, int n_units,
                                                                                 unsigned int* __restrict__ counter, int always_add) {
    constexpr int GPW = 32 / LANES;                   // runs walked side by side by one warp
    constexpr int S = LANES;                          // steps per record tile (a sub-warp stages LANES records at a time)
    constexpr int RING = 4 * LANES;                   // records of a run resident in shared memory (4 tiles)
    static_assert((D & (D - 1)) == 0 && S >= D && S % D == 0, "pipeline depth");
    __shared__ int2 srec[8][GPW][RING];               // per sub-warp: (u, r bits), record j of the run at j % RING
    __shared__ __align__(16) float4 sq0[VEC][8][32];  // q_i as the run found it (only the merge needs it again)
    pdl_launch_dependents();
    if (a.sm_limit > 0) {                             // SMs kept free for the rotation's NCCL kernels (engine.cu reserve_sms)
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (smid >= (unsigned int)a.sm_limit) {
            pdl_wait_prerequisites();
            return;
        }
    }
    const int lane = threadIdx.x & 31;
    const int wic = threadIdx.x >> 5;
    const int gl = lane & (LANES - 1);
    const int grp = lane / LANES;
    const int64_t k = FULL ? (int64_t)(4 * LANES * VEC) : (int64_t)a.k;
    const int chunks = (int)(k >> 2);
    using pchunk_t = typename PChunk<PH>::type;
    constexpr int ES = p_elem_bytes<PH>();                // bytes per stored value of P
    const int64_t prow_bytes = k * ES;
    char* const Pl = reinterpret_cast<char*>(a.P) + ((int64_t)4 * gl - (int64_t)a.u_base * k) * ES;   // lane-adjusted, indexed by global user id
    const uint32_t s32 = sr_seed32(a.seed);
    float* const BUl = BIAS ? a.BU - a.u_base : nullptr;     // model extension: user biases by global id
    const Coef cf = {a.lr, a.lambda, __fsub_rn(1.0f, __fmul_rn(a.lr, a.lambda))};
    const float ccoef = -__fmul_rn(a.lr, a.lambda);   // PRED: p_u += b * q_i + ccoef * p_u, added in memory
    const int32_t* __restrict__ words = reinterpret_cast<const int32_t*>(a.recs);
    const uint64_t pol = l2_policy_evict_first();
    int2* const recs = srec[wic][grp];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (;;) {
        unsigned int first = 0;
        if (lane == 0) first = atomicAdd(counter, (unsigned int)GPW);
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= (unsigned int)n_units) break;
        const bool has = first + (unsigned int)grp < (unsigned int)n_units;     // the sub-warp has a run of its own
        const HotUnit hu = units[has ? first + grp : first];
        const int count = has ? hu.count : 0;
        const int steps = GPW == 1 ? count : __reduce_max_sync(0xffffffffu, count);   // warp-uniform trip count
        const bool virt = a.virt != 0 && hu.bn > 1;
        const int vhb = virt ? perm_half_bits((uint64_t)hu.bn) : 0;
        const uint64_t vkey = virt ? bucket_perm_key(a.seed, a.epoch, hu.bid) : 0;
        float* const qrow = a.Q + (int64_t)(hu.item - a.i_base) * k + 4 * gl;
        float4 q[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            q[v] = (has && (FULL || gl + v * LANES < chunks)) ? ld_row4(qrow + 4 * v * LANES) : zero4;
            sq0[v][wic][lane] = q[v];
        }
        float* const birow = BIAS ? a.BI + (hu.item - a.i_base) : nullptr;
        float bi = (BIAS && has) ? __ldcg(birow) : 0.0f;          // the run's private b_i, merged like q_i at the end
        const float bi0 = bi;
        auto stage_tile = [&](int tile) {               // lane gl copies (u, r) of record LANES*tile + gl of its run
            const int j = tile * LANES + gl;
            if (j < count) {
                const int64_t pos = hu.start + j;
                const int64_t idx = virt ? hu.bstart + (int64_t)block_perm((uint64_t)(pos - hu.bstart), (uint64_t)hu.bn, vhb, vkey) : pos;
                cp_async_4_stream(&recs[j & (RING - 1)].x, words + 3 * idx, pol);
                cp_async_4_stream(&recs[j & (RING - 1)].y, words + 3 * idx + 2, pol);
            }
            cp_async_commit();
        };
        auto gather = [&](pchunk_t (&slot)[VEC], float& bslot, int t) {   // p_u (and b_u) of step t -> a ring slot (nothing past the run's end)
            if (t < count) {
                const int32_t uu = recs[t & (RING - 1)].x & REC_USER_MASK;
                if (BIAS) bslot = __ldcg(BUl + uu);
                const char* xp = Pl + (int64_t)uu * prow_bytes;
#pragma unroll
                for (int v = 0; v < VEC; v++)
                    if (FULL || gl + v * LANES < chunks) slot[v] = ld_pchunk<PH>(xp + 4 * v * LANES * ES);
            }
        };
        // prologue: tiles 0 and 1 staged and visible, D-1 gathers in flight
        stage_tile(0);
        stage_tile(1);
        cp_async_wait<0>();
        __syncwarp();
        pchunk_t ring[D][VEC];
        float bring[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            bring[d] = 0.0f;
#pragma unroll
            for (int v = 0; v < VEC; v++) ring[d][v] = zero_pchunk<PH>();
        }
#pragma unroll
        for (int d = 0; d < D - 1; d++) gather(ring[d], bring[d], d);
        for (int t0 = 0; t0 < steps; t0 += D) {
            if ((t0 & (S - 1)) == 0) {                  // warp-uniform: a tile opens
                stage_tile(t0 / S + 2);
                cp_async_wait<1>();                     // the tile staged one tile ago has landed ...
                __syncwarp();                           // ... and is visible to every lane of the sub-warp
            }
#pragma unroll
            for (int d = 0; d < D; d++) {
                const int t = t0 + d;
                if (t >= steps) break;                  // warp-uniform
                gather(ring[(d + D - 1) & (D - 1)], bring[(d + D - 1) & (D - 1)], t + D - 1);   // into the slot step t - 1 has just released
                const int2 rec = recs[t & (RING - 1)];
                float4 pw[VEC];                          // the gathered row in binary32 (PH: widened here, exactly)
#pragma unroll
                for (int v = 0; v < VEC; v++) pw[v] = widen4(ring[d][v]);
                float pred = rows_dot<LANES, VEC, FAST>(pw, q);
                if (BIAS) pred = __fadd_rn(__fadd_rn(pred, bring[d]), bi);
                const float e = __fsub_rn(__int_as_float(rec.y), pred);
                const float b = __fmul_rn(cf.lr, e);
                if (t < count) {
                    char* const cp = Pl + (int64_t)(rec.x & REC_USER_MASK) * prow_bytes;
                    const bool p_red = PRED || rec.x < 0;       // heavy user (or every user): add the increment in memory
                    if (BIAS) {
                        const float du = bias_delta(bring[d], e, cf.lr, cf.lambda);
                        if (gl == 0) {
                            if (p_red) atomicAdd(BUl + (rec.x & REC_USER_MASK), du);
                            else __stcg(BUl + (rec.x & REC_USER_MASK), __fadd_rn(bring[d], du));
                        }
                        bi = __fadd_rn(bi, bias_delta(bi, e, cf.lr, cf.lambda));
                    }
#pragma unroll
                    for (int v = 0; v < VEC; v++) {
                        if (FULL || gl + v * LANES < chunks) {
                            char* const dst = cp + 4 * v * LANES * ES;
                            if constexpr (PH) {
                                const uint32_t w = sr_word(s32, a.epoch, (uint32_t)(rec.x & REC_USER_MASK), (uint32_t)hu.item, (uint32_t)(gl + v * LANES));
                                const float4 np = new_chunk<FAST>(pw[v], q[v], e, cf.lr, cf.lambda, cf.acoef, b);
                                if (p_red) red_pchunk_f16(dst, np, ring[d][v], w);
                                else st_pchunk(dst, np, ring[d][v], w);
                            } else {
                                if (p_red) red_add_row4(reinterpret_cast<float*>(dst), delta_chunk<FAST>(pw[v], q[v], e, cf.lr, cf.lambda, ccoef, b));
                                else st_row4(reinterpret_cast<float*>(dst), new_chunk<FAST>(pw[v], q[v], e, cf.lr, cf.lambda, cf.acoef, b));
                            }
                            q[v] = new_chunk<FAST>(q[v], pw[v], e, cf.lr, cf.lambda, cf.acoef, b);
                        }
                    }
                }
            }
        }
        cp_async_wait<0>();
        __syncwarp();                                   // every lane is done with the runs' record tiles
        if (BIAS && has && gl == 0) {
            if (hu.weight == 1.0f && !always_add) __stcg(birow, bi);
            else atomicAdd(birow, __fmul_rn(__fsub_rn(bi, bi0), hu.weight));
        }
        if (has) {
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                if (FULL || gl + v * LANES < chunks) {
                    if (hu.weight == 1.0f && !always_add) {
                        st_row4(qrow + 4 * v * LANES, q[v]);
                    } else {
                        const float w = hu.weight;
                        const float4 q0 = sq0[v][wic][lane];
                        red_add_row4(qrow + 4 * v * LANES,
                                     make_float4(__fmul_rn(__fsub_rn(q[v].x, q0.x), w), __fmul_rn(__fsub_rn(q[v].y, q0.y), w),
                                                 __fmul_rn(__fsub_rn(q[v].z, q0.z), w), __fmul_rn(__fsub_rn(q[v].w, q0.w), w)));
                    }
                }
            }
        }
    }
    pdl_wait_prerequisites();
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

}  // namespace

// Lanes per rating of the run kernel at rank k: one float4 chunk per lane up to k = 128 (32 lanes = one warp per
// rating there, 16 at k = 64, never fewer than 8), up to 4 chunks per lane beyond. MFSGD_HOT_LANES = 8 | 16 | 32
// overrides it where that needs at most 4 chunks per lane (tuning aid: on the Netflix-shaped workload 8, 16 and 32
// lanes at k = 128 run within 4 % of each other -- the launch is bound by L2 sector throughput, not by issue slots).
Geometry run_geometry_for(int k) {
    const int chunks = k / 4;
    int lanes = chunks <= 8 ? 8 : (chunks <= 16 ? 16 : 32);
    static int forced = -1;
    if (forced < 0) forced = env_int("MFSGD_HOT_LANES", 0);
    if ((forced == 8 || forced == 16 || forced == 32) && (chunks + forced - 1) / forced <= 4) lanes = forced;
    Geometry g;
    g.lanes = lanes;
    g.vec = (chunks + lanes - 1) / lanes;
    g.full = (lanes * g.vec == chunks);
    return g;
}

int run_kernel_lanes(int k) { return run_geometry_for(k).lanes; }

// MFSGD_HDEPTH = 2 | 4: slots of the register ring of gathered rows (depth - 1 gathers in flight per run). Default 2 since
// round 2: between the gather of p_u and its scatter another sub-warp's update of the same row is lost (store) or applied to
// a stale row (red.add); gathering one step ahead instead of three halves that window. On the signal-dominant Netflix-shaped
// set the held-out RMSE after epochs 1 / 2 / 3 is 10.8 / 3.1 / 1.2 % above the sequential oracle with depth 2 against
// 27.5 / 8.5 / 4.0 % with depth 4, for 2.5 % of the epoch time (profiles/r02_experiments.md).
static int run_depth() {
    static int dep = 0;
    if (dep == 0) dep = env_int("MFSGD_HDEPTH", 2) == 4 ? 4 : 2;
    return dep;
}
// MFSGD_PDL = 0: plain stream order between the run launches of consecutive visits. Default: programmatic
// dependent launch, so a visit's first runs fill the SMs the previous visit's last runs no longer occupy.
static int run_pdl() {
    static int on = -1;
    if (on < 0) on = env_int("MFSGD_PDL", 1);
    return on;
}

// CALL(LANES, VEC, FULL) for the run geometry g
#define MFSGD_DISPATCH_RUN_GEOMETRY(g, CALL)                                                              \
    do {                                                                                                  \
        if ((g).lanes == 8) {                                                                             \
            switch ((g).vec) {                                                                            \
                case 1:  if ((g).full) { CALL(8, 1, true); } else { CALL(8, 1, false); } break;           \
                case 2:  if ((g).full) { CALL(8, 2, true); } else { CALL(8, 2, false); } break;           \
                case 3:  if ((g).full) { CALL(8, 3, true); } else { CALL(8, 3, false); } break;           \
                default: if ((g).full) { CALL(8, 4, true); } else { CALL(8, 4, false); } break;           \
            }                                                                                             \
        } else if ((g).lanes == 16) {                                                                     \
            switch ((g).vec) {                                                                            \
                case 1:  if ((g).full) { CALL(16, 1, true); } else { CALL(16, 1, false); } break;         \
                case 2:  if ((g).full) { CALL(16, 2, true); } else { CALL(16, 2, false); } break;         \
                case 3:  if ((g).full) { CALL(16, 3, true); } else { CALL(16, 3, false); } break;         \
                default: if ((g).full) { CALL(16, 4, true); } else { CALL(16, 4, false); } break;         \
            }                                                                                             \
        } else {                                                                                          \
            switch ((g).vec) {                                                                            \
                case 1:  if ((g).full) { CALL(32, 1, true); } else { CALL(32, 1, false); } break;         \
                case 2:  if ((g).full) { CALL(32, 2, true); } else { CALL(32, 2, false); } break;         \
                case 3:  if ((g).full) { CALL(32, 3, true); } else { CALL(32, 3, false); } break;         \
                default: if ((g).full) { CALL(32, 4, true); } else { CALL(32, 4, false); } break;         \
            }                                                                                             \
        }                                                                                                 \
    } while (0)

template <typename Kernel>
static cudaError_t launch_runs(Kernel kernel, int grid, cudaStream_t stream, bool pdl, int always_add, const UpdateArgs& a,
                               const HotUnit* units, int n_units, unsigned int* counter) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, a, units, n_units, counter, always_add);
}

// Does a run launch on a chain of run launches overlap its predecessor's tail (programmatic dependent launch)? Only where
// that is a tail effect: the launch fills the machine (>= 2 runs per resident sub-warp; the resident grid is the occupancy
// limit, so the dependent's CTAs only get SM slots as the predecessor's retire -- MFSGD_HOT_CTAS below that limit would make
// whole launches co-resident, which diverges: round-2 experiment, NaN in the first epoch) and its longest run is no longer
// than twice a sub-warp's share of the launch (runs are handed out longest first, so long runs end well before the launch
// does; a launch whose hot-item runs outlast everything else would still be merging when the next launch reads q_i).
bool hot_launch_overlaps(int k, int n_units, int64_t n_records, int longest_run, int full_grid, int max_sub_warps) {
    const int per_warp = 32 / run_geometry_for(k).lanes;
    if (run_pdl() == 2) return true;                       // tuning aid: force it
    if (run_pdl() == 0 || env_int("MFSGD_HOT_CTAS", 0) > 0) return false;
    const int64_t sub_warps = (int64_t)full_grid * 8 * per_warp;
    if (max_sub_warps > 0 && max_sub_warps < sub_warps) return false;      // a capped grid leaves SM slots free: no co-resident launches
    return (int64_t)n_units >= 2 * sub_warps && (int64_t)longest_run * sub_warps <= 2 * n_records;
}

cudaError_t launch_sgd_update_hot(const UpdateArgs& a, const HotUnit* units, int n_units, unsigned int* counter, bool fast, bool p_red,
                                  int grid, int max_sub_warps, bool overlaps_previous, bool overlapped_by_next, cudaStream_t stream,
                                  int* launches) {
    if (n_units <= 0) return cudaSuccess;
    const Geometry g = run_geometry_for(a.k);
    const int per_warp = 32 / g.lanes;            // runs a warp walks side by side
    const int full_grid = grid;
    // 8 warps per CTA -- and at least two waves of runs per launch: the runs of an item that are in flight together
    // start from the same q_i and are averaged, the next wave builds on their result.
    const int max_grid = (n_units + 16 * per_warp - 1) / (16 * per_warp);
    if (grid > max_grid) grid = max_grid;
    // ... and no more sub-warps than the stripe has users to spare: with more ratings in flight than half the users of the P
    // sub-stripe, every row has concurrent writers all the time and parallel SGD stops resembling the sequential rule
    // (the engine passes a quarter of the sub-stripe's users: at most ~2 ratings in flight per 4 users)
    if (max_sub_warps > 0 && grid > max_sub_warps / (8 * per_warp)) grid = max_sub_warps / (8 * per_warp);
    if (grid < 1) grid = 1;
    // Overlap with the previous visit's launch only where it is a tail effect: both launches fill the machine (this
    // one offers >= 2 runs per resident sub-warp), so this grid's CTAs become resident as the other's retire. Letting
    // the small launches of a small data set all run at once makes SGD diverge (ML-100K-shaped: NaN after 4 epochs):
    // with a whole epoch in flight p_u and q_i are both corrected from the same stale state and the product overshoots
    // (profiles/r01_experiments.md section 8).
    // MFSGD_PDL = 2 overlaps every chained launch (tuning aid).
    (void)full_grid;
    const bool pdl = overlaps_previous;      // decided by the caller (hot_launch_overlaps of this launch, and it follows a run launch)
    // a launch that overlaps a neighbour (either side) never overwrites an item row: its weight-1 runs add their net change too
    const int always_add = (pdl || overlapped_by_next) ? 1 : 0;
    const bool d2 = run_depth() == 2;
    const bool bias = a.BU != nullptr;            // model extension: biases (always one gather ahead)
    const bool p_half = a.p_half != 0;
    if (p_half && (!fast || p_red)) return cudaErrorInvalidValue;      // binary16 P: FMA arrangement, scatter = store (validate_config)
    cudaError_t err = cudaSuccess;
#define RK(L_, V_, F_, FAST_, D_, PRED_, BIAS_, PH_) launch_runs(sgd_update_runs_kernel<L_, V_, F_, FAST_, D_, PRED_, BIAS_, PH_>, grid, stream, pdl, always_add, a, units, n_units, counter)
#define CALL(L, V, F)                                                            \
    err = p_half ? (bias ? RK(L, V, F, true, 2, false, true, true) : RK(L, V, F, true, 2, false, false, true))   \
        : bias ? ((fast && p_red) ? RK(L, V, F, true, 2, true, true, false)               \
                  : fast          ? RK(L, V, F, true, 2, false, true, false)              \
                  : p_red         ? RK(L, V, F, false, 2, true, true, false)              \
                                  : RK(L, V, F, false, 2, false, true, false))            \
        : (fast && p_red && d2) ? RK(L, V, F, true, 2, true, false, false)                \
          : (fast && p_red)     ? RK(L, V, F, true, 4, true, false, false)                \
          : (fast && d2)        ? RK(L, V, F, true, 2, false, false, false)               \
          : fast                ? RK(L, V, F, true, 4, false, false, false)               \
          : p_red               ? RK(L, V, F, false, 2, true, false, false)               \
                                : RK(L, V, F, false, 2, false, false, false)
    MFSGD_DISPATCH_RUN_GEOMETRY(g, CALL);
#undef CALL
#undef RK
    if (launches) *launches += 1;
    return err != cudaSuccess ? err : cudaGetLastError();
}

cudaError_t hot_max_ctas_per_sm(int k, bool fast, bool p_red, bool p_half, int* ctas) {
    const Geometry g = run_geometry_for(k);
    const bool d2 = run_depth() == 2;
    cudaError_t err = cudaSuccess;
#define OCC(L_, V_, F_, FAST_, D_, PRED_, PH_) cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_runs_kernel<L_, V_, F_, FAST_, D_, PRED_, false, PH_>, 256, 0)
#define CALL(L, V, F)                                                \
    err = p_half                ? OCC(L, V, F, true, 2, false, true)          \
          : (fast && p_red && d2) ? OCC(L, V, F, true, 2, true, false)        \
          : (fast && p_red)     ? OCC(L, V, F, true, 4, true, false)          \
          : (fast && d2)        ? OCC(L, V, F, true, 2, false, false)         \
          : fast                ? OCC(L, V, F, true, 4, false, false)         \
          : p_red               ? OCC(L, V, F, false, 2, true, false)         \
                                : OCC(L, V, F, false, 2, false, false)
    MFSGD_DISPATCH_RUN_GEOMETRY(g, CALL);
#undef CALL
#undef OCC
    const int cap = env_int("MFSGD_HOT_CTAS", 0);    // tuning aid: resident run-kernel CTAs per SM
    if (err == cudaSuccess && cap > 0 && *ctas > cap) *ctas = cap;
    return err;
}

}  // namespace mfsgd
