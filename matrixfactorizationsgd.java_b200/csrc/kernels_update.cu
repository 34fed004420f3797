// kernels_update.cu -- subsystem (2): the per-rating SGD update kernels (sm_100a).
//
// Update rule = baseline/java/MatrixFactorizationSGD.java:89-105 (sgdUpdate). One LANES-wide sub-warp
// per rating (LANES = 32 at k = 128: "one warp per rating"), float4 gathers of p_u and q_i, xor-
// butterfly dot product, lock-free scatter. Bound by L2/HBM bandwidth (12 + 16k algorithmic bytes per
// update); tensor cores are deliberately unused (gather-dot-scatter, not a contraction).
#include "kernels.cuh"

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
            if (SC >= 4) {   // store cache-operator experiments: 4 = default (wb), 5 = wt, 6 = cs
                const float4 np_ = upd4(rp.p[v], rp.q[v], e, lr, lambda), nq_ = upd4(rp.q[v], rp.p[v], e, lr, lambda);
                if (SC == 4) { st_row4_wb(prow + 4 * c, np_); st_row4_wb(qrow + 4 * c, nq_); }
                else if (SC == 5) { st_row4_wt(prow + 4 * c, np_); st_row4_wt(qrow + 4 * c, nq_); }
                else { st_row4_cs(prow + 4 * c, np_); st_row4_cs(qrow + 4 * c, nq_); }
                continue;
            }
            if (SC == 1 || SC == 3) red_add_row4(prow + 4 * c, delta4(rp.p[v], rp.q[v], e, lr, lambda));
            else st_row4(prow + 4 * c, upd4(rp.p[v], rp.q[v], e, lr, lambda));
            if (SC == 1 || SC == 2) red_add_row4(qrow + 4 * c, delta4(rp.q[v], rp.p[v], e, lr, lambda));
            else st_row4(qrow + 4 * c, upd4(rp.q[v], rp.p[v], e, lr, lambda));
        }
    }
}

// Hogwild kernel. Work unit = tile of 32 consecutive records per warp (lane l loads record l, 12 B
// each, streamed past L1); the warp's 32/LANES sub-warps walk the tile, each step gathering the rows
// of the NEXT rating before reducing the current one (2 ratings in flight per sub-warp), and the
// next tile's records are fetched a whole tile ahead.
template <int LANES, int VEC, bool FULL, int SC>
__global__ void __launch_bounds__(256) sgd_update_hogwild_kernel(UpdateArgs a) {
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (LANES - 1);
    const int grp = lane / LANES;
    const int chunks = a.k >> 2;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_tiles = (a.n + 31) >> 5;
    const int32_t* __restrict__ words = reinterpret_cast<const int32_t*>(a.recs);
    const uint64_t pol = l2_policy_evict_first();

    int32_t ru = 0, ri = 0, rr = 0;  // this lane's record of the current tile (r as bits)
    int64_t tile = warp;
    if (tile < n_tiles) {
        const int64_t idx = tile * 32 + lane;
        if (idx < a.n) {
            ru = ld_stream_i32(words + 3 * idx, pol);
            ri = ld_stream_i32(words + 3 * idx + 1, pol);
            rr = ld_stream_i32(words + 3 * idx + 2, pol);
        }
    }
    for (; tile < n_tiles; tile += n_warps) {
        const int64_t base = tile * 32;
        const int cnt = (a.n - base) < 32 ? (int)(a.n - base) : 32;
        const int steps = (cnt + GPW - 1) / GPW;
        // records of the warp's next tile, one tile ahead
        int32_t nu = 0, ni = 0, nr = 0;
        {
            const int64_t idx = (tile + n_warps) * 32 + lane;
            if (idx < a.n) {
                nu = ld_stream_i32(words + 3 * idx, pol);
                ni = ld_stream_i32(words + 3 * idx + 1, pol);
                nr = ld_stream_i32(words + 3 * idx + 2, pol);
            }
        }
        RowPair<LANES, VEC> nxt;
        int32_t xu = __shfl_sync(0xffffffffu, ru, grp);
        int32_t xi = __shfl_sync(0xffffffffu, ri, grp);
        int32_t xr = __shfl_sync(0xffffffffu, rr, grp);
        bool xact = grp < cnt;
        float* xp = a.P + (int64_t)(xu - a.u_base) * a.k;
        float* xq = a.Q + (int64_t)(xi - a.i_base) * a.k;
        load_rows<LANES, VEC, FULL>(nxt, xp, xq, gl, chunks, xact);
#pragma unroll 2
        for (int t = 0; t < steps; t++) {
            const RowPair<LANES, VEC> cur = nxt;
            float* const cp = xp;
            float* const cq = xq;
            const float cr = __int_as_float(xr);
            const bool cact = xact;
            if (t + 1 < steps) {
                const int j = (t + 1) * GPW + grp;
                xu = __shfl_sync(0xffffffffu, ru, j);
                xi = __shfl_sync(0xffffffffu, ri, j);
                xr = __shfl_sync(0xffffffffu, rr, j);
                xact = j < cnt;
                xp = a.P + (int64_t)(xu - a.u_base) * a.k;
                xq = a.Q + (int64_t)(xi - a.i_base) * a.k;
                load_rows<LANES, VEC, FULL>(nxt, xp, xq, gl, chunks, xact);
            }
            const float e = __fsub_rn(cr, row_dot<LANES, VEC>(cur));
            if (cact) scatter_rows<LANES, VEC, FULL, SC>(cur, cp, cq, gl, chunks, e, a.lr, a.lambda);
        }
        ru = nu; ri = ni; rr = nr;
    }
}

// (2b) Hot-item kernel. A unit is a run of records that all rate one hot item. The warp keeps q_i in
// registers (each of its 32/LANES sub-warps a private copy, taking alternate records), streams the
// run's users: gather p_u, dot, scatter p_u, update q_i in registers -- the item row costs no L2
// traffic and sees no concurrent writer. At the end the run's net change is merged into Q scaled by
// unit.weight (model averaging over the item's concurrent units); a unit that is alone on its item
// (weight 1) stores q_i outright, which makes the path exactly sequential. Units are claimed from a
// per-launch counter so uneven runs balance themselves.
template <int LANES, int VEC, bool FULL>
__global__ void __launch_bounds__(256) sgd_update_hot_kernel(UpdateArgs a, const HotUnit* __restrict__ units, int n_units,
                                                             unsigned int* __restrict__ counter) {
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (LANES - 1);
    const int grp = lane / LANES;
    const int chunks = a.k >> 2;
    const int32_t* __restrict__ words = reinterpret_cast<const int32_t*>(a.recs);
    const uint64_t pol = l2_policy_evict_first();
    for (;;) {
        unsigned int unit = 0;
        if (lane == 0) unit = atomicAdd(counter, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= (unsigned int)n_units) break;
        const HotUnit hu = units[unit];
        float* const qrow = a.Q + (int64_t)(hu.item - a.i_base) * a.k;
        float4 q0[VEC], q[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            const int c = gl + v * LANES;
            q0[v] = (FULL || c < chunks) ? ld_row4(qrow + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            q[v] = q0[v];
        }
        for (int base = 0; base < hu.count; base += 32) {
            const int cnt = (hu.count - base) < 32 ? (hu.count - base) : 32;
            int32_t ru = 0, rr = 0;
            if (lane < cnt) {
                const int64_t idx = hu.start + base + lane;
                ru = ld_stream_i32(words + 3 * idx, pol);
                rr = ld_stream_i32(words + 3 * idx + 2, pol);
            }
            const int steps = (cnt + GPW - 1) / GPW;
            float4 pn[VEC];
            int32_t xu = __shfl_sync(0xffffffffu, ru, grp);
            int32_t xr = __shfl_sync(0xffffffffu, rr, grp);
            bool xact = grp < cnt;
            float* xp = a.P + (int64_t)(xu - a.u_base) * a.k;
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                const int c = gl + v * LANES;
                pn[v] = (xact && (FULL || c < chunks)) ? ld_row4(xp + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll 2
            for (int t = 0; t < steps; t++) {
                float4 p[VEC];
#pragma unroll
                for (int v = 0; v < VEC; v++) p[v] = pn[v];
                float* const cp = xp;
                const float cr = __int_as_float(xr);
                const bool cact = xact;
                if (t + 1 < steps) {
                    const int j = (t + 1) * GPW + grp;
                    xu = __shfl_sync(0xffffffffu, ru, j);
                    xr = __shfl_sync(0xffffffffu, rr, j);
                    xact = j < cnt;
                    xp = a.P + (int64_t)(xu - a.u_base) * a.k;
#pragma unroll
                    for (int v = 0; v < VEC; v++) {
                        const int c = gl + v * LANES;
                        pn[v] = (xact && (FULL || c < chunks)) ? ld_row4(xp + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                float s = 0.0f;
#pragma unroll
                for (int v = 0; v < VEC; v++) s = dot4_acc(s, p[v], q[v]);
                s = group_sum<LANES>(s);
                const float e = __fsub_rn(cr, s);
                if (cact) {
#pragma unroll
                    for (int v = 0; v < VEC; v++) {
                        const int c = gl + v * LANES;
                        if (FULL || c < chunks) {
                            st_row4(cp + 4 * c, upd4(p[v], q[v], e, a.lr, a.lambda));
                            q[v] = upd4(q[v], p[v], e, a.lr, a.lambda);
                        }
                    }
                }
            }
        }
        // merge the run's result into Q
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            const int c = gl + v * LANES;
            if (FULL || c < chunks) {
                if (GPW == 1 && hu.weight == 1.0f) {
                    st_row4(qrow + 4 * c, q[v]);
                } else {
                    const float w = hu.weight;
                    red_add_row4(qrow + 4 * c, make_float4(__fmul_rn(__fsub_rn(q[v].x, q0[v].x), w), __fmul_rn(__fsub_rn(q[v].y, q0[v].y), w),
                                                           __fmul_rn(__fsub_rn(q[v].z, q0[v].z), w), __fmul_rn(__fsub_rn(q[v].w, q0[v].w), w)));
                }
            }
        }
    }
}

// Deterministic parity mode: a single warp applies the records strictly in array order; sub-warp 0
// holds the rows (the other lanes carry zeros through the shuffles). Each lane re-reads only
// addresses it wrote itself, so program order makes every update see its predecessor's result.
template <int LANES, int VEC, bool FULL>
__global__ void __launch_bounds__(32) sgd_update_deterministic_kernel(UpdateArgs a, float* __restrict__ err_trace) {
    const int lane = threadIdx.x & 31;
    const int gl = lane & (LANES - 1);
    const bool act = lane < LANES;
    const int chunks = a.k >> 2;
    for (int64_t j = 0; j < a.n; j++) {
        const Rec rec = a.recs[j];
        float* prow = a.P + (int64_t)(rec.u - a.u_base) * a.k;
        float* qrow = a.Q + (int64_t)(rec.i - a.i_base) * a.k;
        RowPair<LANES, VEC> rp;
        load_rows<LANES, VEC, FULL>(rp, prow, qrow, gl, chunks, act);
        const float e = __fsub_rn(rec.r, row_dot<LANES, VEC>(rp));
        if (act) scatter_rows<LANES, VEC, FULL, 0>(rp, prow, qrow, gl, chunks, e, a.lr, a.lambda);
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

cudaError_t launch_sgd_update_hogwild(const UpdateArgs& a, int scatter, int grid, int min_windows,
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
#define CALL(L, V, F)                                                                         \
    switch (scatter) {                                                                        \
        case 1: sgd_update_hogwild_kernel<L, V, F, 1><<<grid, 256, 0, stream>>>(a); break;    \
        case 2: sgd_update_hogwild_kernel<L, V, F, 2><<<grid, 256, 0, stream>>>(a); break;    \
        case 3: sgd_update_hogwild_kernel<L, V, F, 3><<<grid, 256, 0, stream>>>(a); break;    \
        case 4: sgd_update_hogwild_kernel<L, V, F, 4><<<grid, 256, 0, stream>>>(a); break;    \
        case 5: sgd_update_hogwild_kernel<L, V, F, 5><<<grid, 256, 0, stream>>>(a); break;    \
        case 6: sgd_update_hogwild_kernel<L, V, F, 6><<<grid, 256, 0, stream>>>(a); break;    \
        default: sgd_update_hogwild_kernel<L, V, F, 0><<<grid, 256, 0, stream>>>(a); break;   \
    }
    MFSGD_DISPATCH_GEOMETRY(g, CALL);
#undef CALL
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_sgd_update_hot(const UpdateArgs& a, const HotUnit* units, int n_units, unsigned int* counter,
                                  int grid, cudaStream_t stream, int* launches) {
    if (n_units <= 0) return cudaSuccess;
    const Geometry g = geometry_for(a.k);
    const int max_grid = (n_units + 7) / 8;   // 8 warps per CTA, one unit per warp at a time
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
#define CALL(L, V, F) sgd_update_hot_kernel<L, V, F><<<grid, 256, 0, stream>>>(a, units, n_units, counter)
    MFSGD_DISPATCH_GEOMETRY(g, CALL);
#undef CALL
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t hogwild_max_ctas_per_sm(int k, int scatter, int* ctas) {
    const Geometry g = geometry_for(k);
    cudaError_t err = cudaSuccess;
#define CALL(L, V, F)                                                                                          \
    err = (scatter == 1)                                                                                       \
              ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hogwild_kernel<L, V, F, 1>, 256, 0) \
              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hogwild_kernel<L, V, F, 0>, 256, 0)
    MFSGD_DISPATCH_GEOMETRY(g, CALL);
#undef CALL
    return err;
}

cudaError_t launch_sgd_update_deterministic(const UpdateArgs& a, float* err_trace, cudaStream_t stream, int* launches) {
    if (a.n <= 0) return cudaSuccess;
    const Geometry g = geometry_for(a.k);
#define CALL(L, V, F) sgd_update_deterministic_kernel<L, V, F><<<1, 32, 0, stream>>>(a, err_trace)
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
