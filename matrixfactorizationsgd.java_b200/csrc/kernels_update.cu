// kernels_update.cu -- subsystem (2): the per-rating SGD update kernels (sm_100a).
//
// Update rule = baseline/java/MatrixFactorizationSGD.java:89-105 (sgdUpdate). One LANES-wide sub-warp
// per rating (LANES = 32 at k = 128: "one warp per rating"), float4 gathers of p_u and q_i, xor-
// butterfly dot product, lock-free scatter. Bound by L2/HBM bandwidth (12 + 16k algorithmic bytes per
// update); tensor cores are deliberately unused (gather-dot-scatter, not a contraction).
#include <cstdlib>

#include "kernels.cuh"
#include "update_math.cuh"

namespace mfsgd {

namespace {

template <int LANES, int VEC>
struct RowPair {
    float4 p[VEC];
    float4 q[VEC];
};

template <int LANES, int VEC, bool FULL>
__device__ __forceinline__ void load_rows(RowPair<LANES, VEC>& rp, const float* __restrict__ prow,
                                          const float* __restrict__ qrow, int gl, int chunks, bool active) {
#pragma unroll
    for (int v = 0; v < VEC; v++) {
        const int c = gl + v * LANES;
        if (active && (FULL || c < chunks)) {
            rp.p[v] = ld_row4(prow + 4 * c);
            rp.q[v] = ld_row4(qrow + 4 * c);
        } else {
            rp.p[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            rp.q[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

template <int LANES, int VEC>
__device__ __forceinline__ float row_dot(const RowPair<LANES, VEC>& rp) {
    float s = 0.0f;
#pragma unroll
    for (int v = 0; v < VEC; v++) s = dot4_acc(s, rp.p[v], rp.q[v]);
    return group_sum<LANES>(s);
}

// SC: 0 = st/st, 1 = red/red, 2 = st P + red Q, 3 = red P + st Q  (mfsgd.h MFSGD_SCATTER_*)
template <int LANES, int VEC, bool FULL, int SC>
__device__ __forceinline__ void scatter_rows(const RowPair<LANES, VEC>& rp, float* prow, float* qrow, int gl,
                                             int chunks, float e, float lr, float lambda) {
#pragma unroll
    for (int v = 0; v < VEC; v++) {
        const int c = gl + v * LANES;
        if (FULL || c < chunks) {
            if (SC == 1 || SC == 3) red_add_row4(prow + 4 * c, delta4(rp.p[v], rp.q[v], e, lr, lambda));
            else st_row4(prow + 4 * c, upd4(rp.p[v], rp.q[v], e, lr, lambda));
            if (SC == 1 || SC == 2) red_add_row4(qrow + 4 * c, delta4(rp.q[v], rp.p[v], e, lr, lambda));
            else st_row4(qrow + 4 * c, upd4(rp.q[v], rp.p[v], e, lr, lambda));
        }
    }
}

// One tile (<= 32 records, staged in shared memory as (u, i, r-bits, -) quads) walked by the warp's 32/LANES
// sub-warps with a DEPTH-deep software pipeline: the row gathers of the next DEPTH-1 ratings are in flight while
// the current one is reduced and scattered. FULLTILE: all 32 records present -> no activity predicates.
// SC: 0 = st/st, 1 = red/red, 2 = st P + red Q, 3 = red P + st Q  (mfsgd.h MFSGD_SCATTER_*; FAST needs SC == 0)
// PH: P rows kept as binary16 (mfsgd_config.p_storage; common.cuh): the slot holds the raw 8-byte chunks of p_u.
template <int VEC, bool PH>
struct Slot {          // only the gathered rows live in registers; ids and rating are re-read from the tile (one LDS)
    typename PChunk<PH>::type p[VEC];
    float4 q[VEC];
    float bu, bi;      // biases of the model extension (BIAS instantiations only; never touched otherwise)
};

template <int LANES, int VEC, bool FULL, bool FULLTILE, bool BIAS, bool PH>
__device__ __forceinline__ void slot_load(Slot<VEC, PH>& sl, const int4* __restrict__ tile, int j, int cnt, const char* __restrict__ Pl,
                                          const float* __restrict__ Ql, int64_t k, int lane_chunk, int chunks,
                                          const float* __restrict__ BUl, const float* __restrict__ BIl) {
    constexpr int ES = p_elem_bytes<PH>();
    const int4 rec = tile[j & 31];
    const int32_t uu = rec.x & REC_USER_MASK;          // bit 31 is the heavy-user mark
    const bool act = FULLTILE || j < cnt;
    if (BIAS) {
        sl.bu = act ? __ldcg(BUl + uu) : 0.0f;
        sl.bi = act ? __ldcg(BIl + rec.y) : 0.0f;
    }
    const char* pp = Pl + (int64_t)uu * k * ES;
    const float* qq = Ql + (int64_t)rec.y * k;
#pragma unroll
    for (int v = 0; v < VEC; v++) {
        const bool on = act && (FULL || lane_chunk + v * LANES < chunks);
        sl.p[v] = on ? ld_pchunk<PH>(pp + 4 * v * LANES * ES) : zero_pchunk<PH>();
        sl.q[v] = on ? ld_row4(qq + 4 * v * LANES) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int LANES, int VEC, bool FULL, int SC, bool FAST, bool FULLTILE, int DEPTH, bool BIAS, bool PH>
__device__ __forceinline__ void walk_tile(const int4* __restrict__ tile, int cnt, char* __restrict__ Pl,
                                          float* __restrict__ Ql, int64_t k, int grp, int lane_chunk, int chunks, Coef cf,
                                          float* __restrict__ BUl, float* __restrict__ BIl, uint32_t s32, uint32_t epoch) {
    static_assert(!PH || (SC == 0 && FAST), "binary16 P: FMA arrangement, scatter = store");
    constexpr int GPW = 32 / LANES;
    constexpr int FULL_STEPS = 32 / GPW;
    constexpr int ES = p_elem_bytes<PH>();
    const int steps = FULLTILE ? FULL_STEPS : (cnt + GPW - 1) / GPW;
    const float ccoef = -__fmul_rn(cf.lr, cf.lambda);      // heavy users: p_u += b * q_i + ccoef * p_u, added in memory
    Slot<VEC, PH> ring[DEPTH];
#pragma unroll
    for (int d = 0; d < DEPTH - 1; d++)
        slot_load<LANES, VEC, FULL, FULLTILE, BIAS, PH>(ring[d], tile, d * GPW + grp, cnt, Pl, Ql, k, lane_chunk, chunks, BUl, BIl);
    for (int t0 = 0; t0 < steps; t0 += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; d++) {
            const int t = t0 + d;
            if ((!FULLTILE || FULL_STEPS % DEPTH != 0) && t >= steps) break;   // warp-uniform; compiled out for full tiles
            // keep DEPTH-1 gathers ahead: step t + DEPTH - 1 goes into the slot that step t - 1 just freed. Past the end of
            // a partial tile slot_load sees j >= cnt and loads nothing; a full tile has no such predicate, so the look-ahead
            // itself is skipped there (warp-uniform) -- it would gather rows of wrapped records that are never used.
            if (!FULLTILE || t + DEPTH - 1 < FULL_STEPS)
                slot_load<LANES, VEC, FULL, FULLTILE, BIAS, PH>(ring[(d + DEPTH - 1) % DEPTH], tile, (t + DEPTH - 1) * GPW + grp, cnt, Pl, Ql, k,
                                                      lane_chunk, chunks, BUl, BIl);
            const int j = t * GPW + grp;
            const int4 rec = tile[j & 31];
            const bool act = FULLTILE ? true : j < cnt;
            const int32_t uu = rec.x & REC_USER_MASK;
            // heavy user (common.cuh REC_USER_MASK): p_u moves by its increment, added in memory, as in the run kernel -- a plain
            // store here would wipe out the run kernel's concurrent red.adds on the same row (ML-100K-shaped signal set: the third
            // of the items that take this path cost 0.3 % of held-out RMSE that way, profiles/r02_experiments.md section 10)
            const bool p_red = SC == 1 || SC == 3 || rec.x < 0;
            char* const cp = Pl + (int64_t)uu * k * ES;
            float* const cq = Ql + (int64_t)rec.y * k;
            const Slot<VEC, PH>& c = ring[d];
            float4 pw[VEC];                        // p_u in binary32 (PH: widened here, exactly)
#pragma unroll
            for (int v = 0; v < VEC; v++) pw[v] = widen4(c.p[v]);
            float pred = rows_dot<LANES, VEC, FAST>(pw, c.q);
            if (BIAS) pred = __fadd_rn(__fadd_rn(pred, c.bu), c.bi);
            const float e = __fsub_rn(__int_as_float(rec.z), pred);
            const float b = __fmul_rn(cf.lr, e);
            if (BIAS && act && lane_chunk == 0) {      // one lane of the sub-warp owns the two bias entries
                if (p_red) atomicAdd(BUl + uu, bias_delta(c.bu, e, cf.lr, cf.lambda));
                else __stcg(BUl + uu, __fadd_rn(c.bu, bias_delta(c.bu, e, cf.lr, cf.lambda)));
                if (SC == 1 || SC == 2) atomicAdd(BIl + rec.y, bias_delta(c.bi, e, cf.lr, cf.lambda));
                else __stcg(BIl + rec.y, __fadd_rn(c.bi, bias_delta(c.bi, e, cf.lr, cf.lambda)));
            }
            if (act) {
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    if (FULL || lane_chunk + v * LANES < chunks) {
                        char* const dst = cp + 4 * v * LANES * ES;
                        if constexpr (PH) {
                            const uint32_t w = sr_word(s32, epoch, (uint32_t)uu, (uint32_t)rec.y, (uint32_t)(lane_chunk + v * LANES));
                            const float4 np = new_chunk<FAST>(pw[v], c.q[v], e, cf.lr, cf.lambda, cf.acoef, b);
                            if (p_red) red_pchunk_f16(dst, np, c.p[v], w);
                            else st_pchunk(dst, np, c.p[v], w);
                        } else {
                            if (p_red) red_add_row4(reinterpret_cast<float*>(dst), delta_chunk<FAST>(pw[v], c.q[v], e, cf.lr, cf.lambda, ccoef, b));
                            else st_row4(reinterpret_cast<float*>(dst), new_chunk<FAST>(pw[v], c.q[v], e, cf.lr, cf.lambda, cf.acoef, b));
                        }
                        if (SC == 1 || SC == 2) red_add_row4(cq + 4 * v * LANES, delta4(c.q[v], pw[v], e, cf.lr, cf.lambda));
                        else st_row4(cq + 4 * v * LANES, new_chunk<FAST>(c.q[v], pw[v], e, cf.lr, cf.lambda, cf.acoef, b));
                    }
                }
            }
        }
    }
}

// Hogwild kernel (cold records). Work unit = tile of 32 consecutive records per warp: lane l streams record l
// (12 B, past L1, evict-first) one tile ahead and parks it in the warp's shared-memory slot, from where every
// step reads its (u, i, r) with one broadcast LDS.128 instead of three shuffles.
template <int LANES, int VEC, bool FULL, int SC, bool FAST, int DEPTH, bool BIAS, bool PH>
__global__ void __launch_bounds__(256, VEC == 1 ? 4 : (VEC == 2 ? 2 : 1)) sgd_update_hogwild_kernel(UpdateArgs a) {
    __shared__ int4 srec[8][2][32];
    const int lane = threadIdx.x & 31;
    const int wic = threadIdx.x >> 5;
    const int gl = lane & (LANES - 1);
    const int grp = lane / LANES;
    const int64_t k = FULL ? (int64_t)(4 * LANES * VEC) : (int64_t)a.k;   // compile-time row length for the common ranks
    const int chunks = (int)(k >> 2);
    char* const Pl = reinterpret_cast<char*>(a.P) + ((int64_t)4 * gl - (int64_t)a.u_base * k) * p_elem_bytes<PH>();   // lane-adjusted bases,
    float* const Ql = a.Q - (int64_t)a.i_base * k + 4 * gl;                                                              // indexed by global ids
    const uint32_t s32 = sr_seed32(a.seed);
    float* const BUl = BIAS ? a.BU - a.u_base : nullptr;                  // biases of the model extension, indexed by global ids
    float* const BIl = BIAS ? a.BI - a.i_base : nullptr;
    const Coef cf = {a.lr, a.lambda, __fsub_rn(1.0f, __fmul_rn(a.lr, a.lambda))};
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_tiles = (a.n + 31) >> 5;
    const int32_t* __restrict__ words = reinterpret_cast<const int32_t*>(a.recs);
    const uint64_t pol = l2_policy_evict_first();

    // position -> record index: identity, or the block's per-epoch permutation (virtual reshuffle)
    const bool virt = a.virt != 0 && a.blk_n > 1;
    const int vhb = virt ? perm_half_bits((uint64_t)a.blk_n) : 0;
    const uint64_t vkey = virt ? bucket_perm_key(a.seed, a.epoch, a.blk_id) : 0;
    auto record_index = [&](int64_t j) -> int64_t {   // j relative to a.first
        const int64_t pos = a.first + j;
        return virt ? a.blk_start + (int64_t)block_perm((uint64_t)(pos - a.blk_start), (uint64_t)a.blk_n, vhb, vkey) : pos;
    };
    int64_t tile = warp;
    int buf = 0;
    if (tile < n_tiles) {
        const int64_t j = tile * 32 + lane;
        int4 rec = make_int4(0, 0, 0, 0);
        if (j < a.n) {
            const int64_t idx = record_index(j);
            rec.x = ld_stream_i32(words + 3 * idx, pol);                    // with the heavy-user mark (bit 31)
            rec.y = ld_stream_i32(words + 3 * idx + 1, pol);
            rec.z = ld_stream_i32(words + 3 * idx + 2, pol);
        }
        srec[wic][0][lane] = rec;
    }
    __syncwarp();
    for (; tile < n_tiles; tile += n_warps) {
        const int64_t base = tile * 32;
        const int cnt = (a.n - base) < 32 ? (int)(a.n - base) : 32;
        int4 nrec = make_int4(0, 0, 0, 0);            // the warp's next tile, fetched now, parked after this tile
        {
            const int64_t j = (tile + n_warps) * 32 + lane;
            if (j < a.n) {
                const int64_t idx = record_index(j);
                nrec.x = ld_stream_i32(words + 3 * idx, pol);
                nrec.y = ld_stream_i32(words + 3 * idx + 1, pol);
                nrec.z = ld_stream_i32(words + 3 * idx + 2, pol);
            }
        }
        if (cnt == 32) walk_tile<LANES, VEC, FULL, SC, FAST, true, DEPTH, BIAS, PH>(srec[wic][buf], 32, Pl, Ql, k, grp, gl, chunks, cf, BUl, BIl, s32, a.epoch);
        else walk_tile<LANES, VEC, FULL, SC, FAST, false, DEPTH, BIAS, PH>(srec[wic][buf], cnt, Pl, Ql, k, grp, gl, chunks, cf, BUl, BIl, s32, a.epoch);
        buf ^= 1;
        srec[wic][buf][lane] = nrec;
        __syncwarp();
    }
}

// Deterministic parity mode: a single warp applies the records strictly in array order; sub-warp 0
// holds the rows (the other lanes carry zeros through the shuffles). Each lane re-reads only
// addresses it wrote itself, so program order makes every update see its predecessor's result.
template <int LANES, int VEC, bool FULL, bool PH>
__global__ void __launch_bounds__(32) sgd_update_deterministic_kernel(UpdateArgs a, float* __restrict__ err_trace) {
    const int lane = threadIdx.x & 31;
    const int gl = lane & (LANES - 1);
    const bool act = lane < LANES;
    const int chunks = a.k >> 2;
    const uint32_t s32 = sr_seed32(a.seed);
    for (int64_t j = 0; j < a.n; j++) {
        const Rec rec = a.recs[a.first + j];
        float* prow = a.P + (int64_t)(rec.u - a.u_base) * a.k;
        float* qrow = a.Q + (int64_t)(rec.i - a.i_base) * a.k;
        RowPair<LANES, VEC> rp;
        if constexpr (PH) {              // binary16 P: widen p_u, the exact rule in binary32, narrow with the update's random words
            const char* ph = reinterpret_cast<const char*>(a.P) + (int64_t)(rec.u - a.u_base) * a.k * 2;
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                const int c = gl + v * LANES;
                const bool on = act && (FULL || c < chunks);
                rp.p[v] = on ? widen4(ld_pchunk<true>(ph + 8 * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                rp.q[v] = on ? ld_row4(qrow + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            load_rows<LANES, VEC, FULL>(rp, prow, qrow, gl, chunks, act);
        }
        float pred = row_dot<LANES, VEC>(rp);
        float bu = 0.0f, bi = 0.0f;
        if (a.BU != nullptr) {                 // model extension: every lane reads the two entries, lane 0 writes them
            bu = __ldcg(a.BU + (rec.u - a.u_base));
            bi = __ldcg(a.BI + (rec.i - a.i_base));
            pred = __fadd_rn(__fadd_rn(pred, bu), bi);
        }
        const float e = __fsub_rn(rec.r, pred);
        __syncwarp();                          // all lanes have read the biases before lane 0 overwrites them
        if (a.BU != nullptr && lane == 0) {
            __stcg(a.BU + (rec.u - a.u_base), __fadd_rn(bu, bias_delta(bu, e, a.lr, a.lambda)));
            __stcg(a.BI + (rec.i - a.i_base), __fadd_rn(bi, bias_delta(bi, e, a.lr, a.lambda)));
        }
        __syncwarp();
        if constexpr (PH) {
            char* ph = reinterpret_cast<char*>(a.P) + (int64_t)(rec.u - a.u_base) * a.k * 2;
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                const int c = gl + v * LANES;
                if (act && (FULL || c < chunks)) {
                    st_pchunk(ph + 8 * c, upd4(rp.p[v], rp.q[v], e, a.lr, a.lambda), make_uint2(0u, 0u),
                              sr_word(s32, a.epoch, (uint32_t)rec.u, (uint32_t)rec.i, (uint32_t)c));
                    st_row4(qrow + 4 * c, upd4(rp.q[v], rp.p[v], e, a.lr, a.lambda));
                }
            }
        } else {
            if (act) scatter_rows<LANES, VEC, FULL, 0>(rp, prow, qrow, gl, chunks, e, a.lr, a.lambda);
        }
        if (err_trace != nullptr && lane == 0) err_trace[j] = e;
    }
}

// Teacher-forced: update j reads pre_p[j], pre_q[j] (dense [n,k]) and writes post rows; one sub-warp each.
template <int LANES, int VEC, bool FULL>
__global__ void __launch_bounds__(256) sgd_update_forced_kernel(int k, float lr, float lambda, int64_t n,
                                                                const float* __restrict__ pre_p,
                                                                const float* __restrict__ pre_q,
                                                                const float* __restrict__ r, float* __restrict__ post_p,
                                                                float* __restrict__ post_q, float* __restrict__ err) {
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (LANES - 1);
    const int grp = lane / LANES;
    const int chunks = k >> 2;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * GPW; base < n; base += n_warps * GPW) {   // warp-uniform trip count
        const int64_t j = base + grp;
        const bool act = j < n;
        const int64_t jj = act ? j : 0;
        RowPair<LANES, VEC> rp;
        load_rows<LANES, VEC, FULL>(rp, pre_p + jj * k, pre_q + jj * k, gl, chunks, act);
        const float e = __fsub_rn(act ? r[jj] : 0.f, row_dot<LANES, VEC>(rp));
        if (act) {
            scatter_rows<LANES, VEC, FULL, 0>(rp, post_p + jj * k, post_q + jj * k, gl, chunks, e, lr, lambda);
            if (gl == 0) err[jj] = e;
        }
    }
}


}  // namespace

static int pipeline_depth() {   // MFSGD_DEPTH = 2 | 4 (tuning aid); gathers in flight per sub-warp = depth - 1
    static int depth = 0;
    if (depth == 0) {
        const char* e = getenv("MFSGD_DEPTH");
        depth = (e && atoi(e) == 4) ? 4 : 2;
    }
    return depth;
}

cudaError_t launch_sgd_update_hogwild(const UpdateArgs& a, int scatter, bool fast, int grid, int min_windows,
                                      cudaStream_t stream, int* launches) {
    if (min_windows < 1) min_windows = 1;
    if (a.n <= 0) return cudaSuccess;
    const Geometry g = geometry_for(a.k);
    // Concurrency cap for small blocks: every resident sub-warp has ~2 ratings in flight that all read
    // the same stale snapshot, so a launch should still span >= min_windows such snapshots, or the
    // last-writer-wins scatter throws most of the block's updates away (tools/staleness_sim.py).
    const int gpw = 32 / g.lanes;
    const int64_t max_warps = a.n / ((int64_t)min_windows * 2 * gpw);
    int64_t max_grid = (max_warps + 7) / 8;   // 8 warps per CTA
    if (max_grid < 1) max_grid = 1;
    if (grid > max_grid) grid = (int)max_grid;
    if (grid < 1) grid = 1;
    if (scatter != 0) fast = false;           // the atomic scatter variants add exact-rule deltas
    const bool deep = pipeline_depth() == 4 && g.vec == 1;
    const bool bias = a.BU != nullptr;        // model extension: its own instantiations (one gather ahead), none of its registers otherwise
    const bool p_half = a.p_half != 0;        // binary16 P rows: FMA arrangement and plain stores only (validate_config)
    if (p_half && (!fast || scatter != 0)) return cudaErrorInvalidValue;
#define HK(L_, V_, F_, SC_, FAST_, D_, BIAS_, PH_) sgd_update_hogwild_kernel<L_, V_, F_, SC_, FAST_, D_, BIAS_, PH_><<<grid, 256, 0, stream>>>(a)
#define CALL(L, V, F)                                                  \
    if (p_half) { if (bias) HK(L, V, F, 0, true, 2, true, true); else HK(L, V, F, 0, true, 2, false, true); }   \
    else if (bias && fast) HK(L, V, F, 0, true, 2, true, false);                \
    else if (bias) switch (scatter) {                                  \
        case 1: HK(L, V, F, 1, false, 2, true, false); break;                   \
        case 2: HK(L, V, F, 2, false, 2, true, false); break;                   \
        case 3: HK(L, V, F, 3, false, 2, true, false); break;                   \
        default: HK(L, V, F, 0, false, 2, true, false); break;                  \
    }                                                                  \
    else if (fast && deep) HK(L, V, F, 0, true, 4, false, false);               \
    else if (fast) HK(L, V, F, 0, true, 2, false, false);                       \
    else switch (scatter) {                                            \
        case 1: HK(L, V, F, 1, false, 2, false, false); break;                  \
        case 2: HK(L, V, F, 2, false, 2, false, false); break;                  \
        case 3: HK(L, V, F, 3, false, 2, false, false); break;                  \
        default: HK(L, V, F, 0, false, 2, false, false); break;                 \
    }
    MFSGD_DISPATCH_GEOMETRY(g, CALL);
#undef CALL
#undef HK
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t hogwild_max_ctas_per_sm(int k, int scatter, bool fast, bool p_half, int* ctas) {
    const Geometry g = geometry_for(k);
    cudaError_t err = cudaSuccess;
    if (scatter != 0) fast = false;
    const bool deep = pipeline_depth() == 4 && g.vec == 1;
#define CALL(L, V, F)                                                                                                          \
    err = p_half ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hogwild_kernel<L, V, F, 0, true, 2, false, true>, 256, 0)         \
          : (fast && deep) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hogwild_kernel<L, V, F, 0, true, 4, false, false>, 256, 0) \
          : fast ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hogwild_kernel<L, V, F, 0, true, 2, false, false>, 256, 0)       \
                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hogwild_kernel<L, V, F, 0, false, 2, false, false>, 256, 0)
    MFSGD_DISPATCH_GEOMETRY(g, CALL);
#undef CALL
    return err;
}

cudaError_t launch_sgd_update_deterministic(const UpdateArgs& a, float* err_trace, cudaStream_t stream, int* launches) {
    if (a.n <= 0) return cudaSuccess;
    const Geometry g = geometry_for(a.k);
#define CALL(L, V, F)                                                                                  \
    if (a.p_half) sgd_update_deterministic_kernel<L, V, F, true><<<1, 32, 0, stream>>>(a, err_trace);  \
    else sgd_update_deterministic_kernel<L, V, F, false><<<1, 32, 0, stream>>>(a, err_trace)
    MFSGD_DISPATCH_GEOMETRY(g, CALL);
#undef CALL
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_sgd_update_forced(int k, float lr, float lambda, int64_t n, const float* pre_p, const float* pre_q,
                                     const float* r, float* post_p, float* post_q, float* err, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const Geometry g = geometry_for(k);
    const int gpw = 32 / g.lanes;
    int64_t ctas = (n + (int64_t)gpw * 8 - 1) / ((int64_t)gpw * 8);
    if (ctas > 148 * 8) ctas = 148 * 8;
    const int grid = (int)ctas;
#define CALL(L, V, F) \
    sgd_update_forced_kernel<L, V, F><<<grid, 256, 0, stream>>>(k, lr, lambda, n, pre_p, pre_q, r, post_p, post_q, err)
    MFSGD_DISPATCH_GEOMETRY(g, CALL);
#undef CALL
    return cudaGetLastError();
}

}  // namespace mfsgd
