// kernels_update.cu -- subsystem (2): the per-rating SGD update kernels (sm_100a).
//
// Update rule = baseline/java/MatrixFactorizationSGD.java:89-105 (sgdUpdate). One LANES-wide sub-warp
// per rating (LANES = 32 at k = 128: "one warp per rating"), float4 gathers of p_u and q_i, xor-
// butterfly dot product, lock-free scatter. Bound by L2/HBM bandwidth (12 + 16k algorithmic bytes per
// update); tensor cores are deliberately unused (gather-dot-scatter, not a contraction).
#include <cstdlib>

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
            if (SC == 1 || SC == 3) red_add_row4(prow + 4 * c, delta4(rp.p[v], rp.q[v], e, lr, lambda));
            else st_row4(prow + 4 * c, upd4(rp.p[v], rp.q[v], e, lr, lambda));
            if (SC == 1 || SC == 2) red_add_row4(qrow + 4 * c, delta4(rp.q[v], rp.p[v], e, lr, lambda));
            else st_row4(qrow + 4 * c, upd4(rp.q[v], rp.p[v], e, lr, lambda));
        }
    }
}

// ---- arithmetic of the full-grid modes -------------------------------------------------------------
// EXACT (FAST = false): the reference rule operation by operation, no FMA -- identical to the
// deterministic kernel and to oracle.cpp ORC_ORDER_WARP_TREE.
// FAST (FAST = true): the same algebra arranged for Blackwell's packed FP32 pipe (FFMA2):
//     lane partial : (lo, hi) = fma2((p.z,p.w),(q.z,q.w), (p.x*q.x, p.y*q.y)), ... ; s = lo + hi
//     update       : p' = fma(b, q, a*p),  q' = fma(b, p, a*q),  a = 1 - lr*lambda,  b = lr*e
// A per-update deviation of a few ulp from the reference rule (<< the 1e-5 bar); oracle.cpp
// ORC_ORDER_WARP_TREE_FMA reproduces it bit for bit with fmaf. 3.3x fewer issue slots per update.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int LANES, int VEC, bool FAST>
__device__ __forceinline__ float rows_dot(const float4 (&p)[VEC], const float4 (&q)[VEC]) {
    float s;
    if (FAST) {
        uint64_t acc = mul2(pk2(p[0].x, p[0].y), pk2(q[0].x, q[0].y));
        acc = fma2(pk2(p[0].z, p[0].w), pk2(q[0].z, q[0].w), acc);
#pragma unroll
        for (int v = 1; v < VEC; v++) {
            acc = fma2(pk2(p[v].x, p[v].y), pk2(q[v].x, q[v].y), acc);
            acc = fma2(pk2(p[v].z, p[v].w), pk2(q[v].z, q[v].w), acc);
        }
        float lo, hi;
        upk2(acc, lo, hi);
        s = __fadd_rn(lo, hi);
    } else {
        s = 0.0f;
#pragma unroll
        for (int v = 0; v < VEC; v++) s = dot4_acc(s, p[v], q[v]);
    }
    return group_sum<LANES>(s);
}

// new value of row chunk `o` given the other row's chunk `x`
template <bool FAST>
__device__ __forceinline__ float4 new_chunk(float4 o, float4 x, float e, float lr, float lambda, float acoef, float b) {
    if (FAST) {
        const uint64_t a2 = pk2(acoef, acoef), b2 = pk2(b, b);
        const uint64_t lo = fma2(b2, pk2(x.x, x.y), mul2(a2, pk2(o.x, o.y)));
        const uint64_t hi = fma2(b2, pk2(x.z, x.w), mul2(a2, pk2(o.z, o.w)));
        float4 r;
        upk2(lo, r.x, r.y);
        upk2(hi, r.z, r.w);
        return r;
    }
    return upd4(o, x, e, lr, lambda);
}

struct Coef {
    float lr, lambda, acoef;   // acoef = 1 - lr * lambda (FAST arithmetic)
};

// One tile (<= 32 records, staged in shared memory as (u, i, r-bits, -) quads) walked by the warp's 32/LANES
// sub-warps with a DEPTH-deep software pipeline: the row gathers of the next DEPTH-1 ratings are in flight while
// the current one is reduced and scattered. FULLTILE: all 32 records present -> no activity predicates.
// SC: 0 = st/st, 1 = red/red, 2 = st P + red Q, 3 = red P + st Q  (mfsgd.h MFSGD_SCATTER_*; FAST needs SC == 0)
template <int VEC>
struct Slot {          // only the gathered rows live in registers; ids and rating are re-read from the tile (one LDS)
    float4 p[VEC], q[VEC];
};

template <int LANES, int VEC, bool FULL, bool FULLTILE>
__device__ __forceinline__ void slot_load(Slot<VEC>& sl, const int4* __restrict__ tile, int j, int cnt, const float* __restrict__ Pl,
                                          const float* __restrict__ Ql, int64_t k, int lane_chunk, int chunks) {
    const int4 rec = tile[j & 31];
    const bool act = FULLTILE || j < cnt;
    const float* pp = Pl + (int64_t)rec.x * k;
    const float* qq = Ql + (int64_t)rec.y * k;
#pragma unroll
    for (int v = 0; v < VEC; v++) {
        const bool on = act && (FULL || lane_chunk + v * LANES < chunks);
        sl.p[v] = on ? ld_row4(pp + 4 * v * LANES) : make_float4(0.f, 0.f, 0.f, 0.f);
        sl.q[v] = on ? ld_row4(qq + 4 * v * LANES) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int LANES, int VEC, bool FULL, int SC, bool FAST, bool FULLTILE, int DEPTH>
__device__ __forceinline__ void walk_tile(const int4* __restrict__ tile, int cnt, float* __restrict__ Pl,
                                          float* __restrict__ Ql, int64_t k, int grp, int lane_chunk, int chunks, Coef cf) {
    constexpr int GPW = 32 / LANES;
    constexpr int FULL_STEPS = 32 / GPW;
    const int steps = FULLTILE ? FULL_STEPS : (cnt + GPW - 1) / GPW;
    Slot<VEC> ring[DEPTH];
#pragma unroll
    for (int d = 0; d < DEPTH - 1; d++)
        slot_load<LANES, VEC, FULL, FULLTILE>(ring[d], tile, d * GPW + grp, cnt, Pl, Ql, k, lane_chunk, chunks);
    for (int t0 = 0; t0 < steps; t0 += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; d++) {
            const int t = t0 + d;
            if ((!FULLTILE || FULL_STEPS % DEPTH != 0) && t >= steps) break;   // warp-uniform; compiled out for full tiles
            // keep DEPTH-1 gathers ahead: step t + DEPTH - 1 goes into the slot that step t - 1 just freed
            // (past the end of a partial tile slot_load sees j >= cnt and loads nothing)
            slot_load<LANES, VEC, FULL, FULLTILE>(ring[(d + DEPTH - 1) % DEPTH], tile, (t + DEPTH - 1) * GPW + grp,
                                                  (FULLTILE && t + DEPTH - 1 >= FULL_STEPS) ? 0 : cnt, Pl, Ql, k, lane_chunk, chunks);
            const int j = t * GPW + grp;
            const int4 rec = tile[j & 31];
            const bool act = FULLTILE ? true : j < cnt;
            float* const cp = Pl + (int64_t)rec.x * k;
            float* const cq = Ql + (int64_t)rec.y * k;
            const Slot<VEC>& c = ring[d];
            const float e = __fsub_rn(__int_as_float(rec.z), rows_dot<LANES, VEC, FAST>(c.p, c.q));
            const float b = __fmul_rn(cf.lr, e);
            if (act) {
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    if (FULL || lane_chunk + v * LANES < chunks) {
                        if (SC == 1 || SC == 3) red_add_row4(cp + 4 * v * LANES, delta4(c.p[v], c.q[v], e, cf.lr, cf.lambda));
                        else st_row4(cp + 4 * v * LANES, new_chunk<FAST>(c.p[v], c.q[v], e, cf.lr, cf.lambda, cf.acoef, b));
                        if (SC == 1 || SC == 2) red_add_row4(cq + 4 * v * LANES, delta4(c.q[v], c.p[v], e, cf.lr, cf.lambda));
                        else st_row4(cq + 4 * v * LANES, new_chunk<FAST>(c.q[v], c.p[v], e, cf.lr, cf.lambda, cf.acoef, b));
                    }
                }
            }
        }
    }
}

// Hogwild kernel (cold records). Work unit = tile of 32 consecutive records per warp: lane l streams record l
// (12 B, past L1, evict-first) one tile ahead and parks it in the warp's shared-memory slot, from where every
// step reads its (u, i, r) with one broadcast LDS.128 instead of three shuffles.
template <int LANES, int VEC, bool FULL, int SC, bool FAST, int DEPTH>
__global__ void __launch_bounds__(256, VEC == 1 ? 4 : (VEC == 2 ? 2 : 1)) sgd_update_hogwild_kernel(UpdateArgs a) {
    __shared__ int4 srec[8][2][32];
    const int lane = threadIdx.x & 31;
    const int wic = threadIdx.x >> 5;
    const int gl = lane & (LANES - 1);
    const int grp = lane / LANES;
    const int64_t k = FULL ? (int64_t)(4 * LANES * VEC) : (int64_t)a.k;   // compile-time row length for the common ranks
    const int chunks = (int)(k >> 2);
    float* const Pl = a.P - (int64_t)a.u_base * k + 4 * gl;               // lane-adjusted bases, indexed by global ids
    float* const Ql = a.Q - (int64_t)a.i_base * k + 4 * gl;
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
            rec.x = ld_stream_i32(words + 3 * idx, pol);
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
        if (cnt == 32) walk_tile<LANES, VEC, FULL, SC, FAST, true, DEPTH>(srec[wic][buf], 32, Pl, Ql, k, grp, gl, chunks, cf);
        else walk_tile<LANES, VEC, FULL, SC, FAST, false, DEPTH>(srec[wic][buf], cnt, Pl, Ql, k, grp, gl, chunks, cf);
        buf ^= 1;
        srec[wic][buf][lane] = nrec;
        __syncwarp();
    }
}

// (2b) Hot-item kernel. A unit is a run of records that all rate one hot item. The warp keeps q_i in
// registers (each of its 32/LANES sub-warps a private copy, taking alternate records), streams the
// run's users: gather p_u, dot, scatter p_u, update q_i in registers -- the item row costs no L2
// traffic and sees no concurrent writer. At the end the run's net change is merged into Q scaled by
// unit.weight (model averaging over the item's concurrent units); a unit that is alone on its item
// (weight 1) stores q_i outright, which makes the path exactly sequential. Units are claimed from a
// per-launch counter so uneven runs balance themselves.
template <int LANES, int VEC, bool FULL, bool FAST, int HD>
__global__ void __launch_bounds__(256) sgd_update_hot_kernel(UpdateArgs a, const HotUnit* __restrict__ units, int n_units,
                                                             unsigned int* __restrict__ counter) {
    constexpr int GPW = 32 / LANES;
    constexpr int HDEPTH = VEC == 1 ? HD : 2;
    __shared__ int2 srec[8][32];
    const int lane = threadIdx.x & 31;
    const int wic = threadIdx.x >> 5;
    const int gl = lane & (LANES - 1);
    const int grp = lane / LANES;
    const int64_t k = FULL ? (int64_t)(4 * LANES * VEC) : (int64_t)a.k;
    const int chunks = (int)(k >> 2);
    float* const Pl = a.P - (int64_t)a.u_base * k + 4 * gl;
    const Coef cf = {a.lr, a.lambda, __fsub_rn(1.0f, __fmul_rn(a.lr, a.lambda))};
    const int32_t* __restrict__ words = reinterpret_cast<const int32_t*>(a.recs);
    const uint64_t pol = l2_policy_evict_first();
    for (;;) {
        unsigned int unit = 0;
        if (lane == 0) unit = atomicAdd(counter, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= (unsigned int)n_units) break;
        const HotUnit hu = units[unit];
        const bool virt = a.virt != 0 && hu.bn > 1;
        const int vhb = virt ? perm_half_bits((uint64_t)hu.bn) : 0;
        const uint64_t vkey = virt ? bucket_perm_key(a.seed, a.epoch, hu.bid) : 0;
        float* const qrow = a.Q + (int64_t)(hu.item - a.i_base) * k + 4 * gl;
        float4 q0[VEC], q[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            q0[v] = (FULL || gl + v * LANES < chunks) ? ld_row4(qrow + 4 * v * LANES) : make_float4(0.f, 0.f, 0.f, 0.f);
            q[v] = q0[v];
        }
        for (int base = 0; base < hu.count; base += 32) {
            const int cnt = (hu.count - base) < 32 ? (hu.count - base) : 32;
            __syncwarp();                                   // the previous tile's readers are done with the slot
            if (lane < cnt) {
                const int64_t pos = hu.start + base + lane;
                const int64_t idx = virt ? hu.bstart + (int64_t)block_perm((uint64_t)(pos - hu.bstart), (uint64_t)hu.bn, vhb, vkey) : pos;
                srec[wic][lane] = make_int2(ld_stream_i32(words + 3 * idx, pol), ld_stream_i32(words + 3 * idx + 2, pol));
            }
            __syncwarp();
            const int steps = (cnt + GPW - 1) / GPW;
            // HDEPTH-deep ring of gathered p rows: the gathers do not depend on q_i, so several ratings' rows are
            // in flight while the (serial, q_i-dependent) dot -> update chain walks the run.
            float4 ring[HDEPTH][VEC];
#pragma unroll
            for (int d = 0; d < HDEPTH - 1; d++) {
                const int j = d * GPW + grp;
                const int2 rec = srec[wic][j & 31];
                const float* xp = Pl + (int64_t)rec.x * k;
#pragma unroll
                for (int v = 0; v < VEC; v++)
                    ring[d][v] = (j < cnt && (FULL || gl + v * LANES < chunks)) ? ld_row4(xp + 4 * v * LANES) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for (int t0 = 0; t0 < steps; t0 += HDEPTH) {
#pragma unroll
                for (int d = 0; d < HDEPTH; d++) {
                    const int t = t0 + d;
                    if (t >= steps) break;                                   // warp-uniform
                    {
                        const int j = (t + HDEPTH - 1) * GPW + grp;
                        const int2 rec = srec[wic][j & 31];
                        const float* xp = Pl + (int64_t)rec.x * k;
#pragma unroll
                        for (int v = 0; v < VEC; v++)
                            ring[(d + HDEPTH - 1) % HDEPTH][v] =
                                (j < cnt && (FULL || gl + v * LANES < chunks)) ? ld_row4(xp + 4 * v * LANES) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    const int j = t * GPW + grp;
                    const int2 rec = srec[wic][j & 31];
                    float* const cp = Pl + (int64_t)rec.x * k;
                    const bool cact = j < cnt;
                    float4 p[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; v++) p[v] = ring[d][v];
                    const float e = __fsub_rn(__int_as_float(rec.y), rows_dot<LANES, VEC, FAST>(p, q));
                    const float b = __fmul_rn(cf.lr, e);
                    if (cact) {
#pragma unroll
                        for (int v = 0; v < VEC; v++) {
                            if (FULL || gl + v * LANES < chunks) {
                                st_row4(cp + 4 * v * LANES, new_chunk<FAST>(p[v], q[v], e, cf.lr, cf.lambda, cf.acoef, b));
                                q[v] = new_chunk<FAST>(q[v], p[v], e, cf.lr, cf.lambda, cf.acoef, b);
                            }
                        }
                    }
                }
            }
        }
        // merge the run's result into Q
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            if (FULL || gl + v * LANES < chunks) {
                if (GPW == 1 && hu.weight == 1.0f) {
                    st_row4(qrow + 4 * v * LANES, q[v]);
                } else {
                    const float w = hu.weight;
                    red_add_row4(qrow + 4 * v * LANES,
                                 make_float4(__fmul_rn(__fsub_rn(q[v].x, q0[v].x), w), __fmul_rn(__fsub_rn(q[v].y, q0[v].y), w),
                                             __fmul_rn(__fsub_rn(q[v].z, q0[v].z), w), __fmul_rn(__fsub_rn(q[v].w, q0[v].w), w)));
                }
            }
        }
    }
}

// (2b') Hot-item kernel, asynchronous-copy pipeline (ranks with one float4 chunk per lane and >= 8 lanes per rating:
// 32 <= k <= 128). Same arithmetic and visiting order as sgd_update_hot_kernel, but
//   * every LANES-wide sub-warp walks a run OF ITS OWN (the warp claims 32/LANES consecutive units), so a run is
//     applied strictly sequentially at every rank -- no averaging inside a warp;
//   * the p_u gathers no longer pass through registers: every step issues one cp.async (LDGSTS, 16 B per lane =
//     one row per sub-warp, L2 -> shared memory) for the rating D-1 steps ahead, and the pipeline runs through the
//     whole run instead of restarting every 32 records: the run's (u, r) pairs are staged by cp.async too, two
//     LANES-record tiles ahead, in a ring of 4 tiles per sub-warp.
//   group accounting: step t commits group t = {row of step t+D-1, records of tile T+2 when t opens tile T};
//   wait_group D-1 then guarantees groups <= t-(D-1), i.e. the row of step t and every record tile opened
//   >= D-1 steps ago. Each lane reads back only the 16 row bytes it copied itself; the record tiles are read by
//   all lanes of the sub-warp, hence the __syncwarp after the wait at the top of every tile. Steps are unrolled by D,
//   so slots are compile-time and a tile (LANES steps, a multiple of D) always opens at the top of an unrolled block.
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_4_stream(void* smem_dst, const void* gsrc, uint64_t pol) {
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// Programmatic dependent launch: lets the next launch on the stream (if it asked for programmatic stream
// serialisation) become resident as this grid's CTAs retire, instead of after the grid has drained. The update
// launches of consecutive visits only ever meet Hogwild-style, so nothing waits on the other side.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// ... but a grid must not COMPLETE before the grid it was allowed to overtake: the stream's later operations (events,
// the next sub-epoch, the Q rotation) take this grid's completion for the completion of everything before it.
// Every thread therefore waits for the prerequisite grid as its last action (a no-op without the launch attribute).
__device__ __forceinline__ void pdl_wait_prerequisites() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int LANES, bool FULL, bool FAST, int D>
__global__ void __launch_bounds__(256) sgd_update_hot_async_kernel(UpdateArgs a, const HotUnit* __restrict__ units, int n_units,
                                                                   unsigned int* __restrict__ counter) {
    constexpr int GPW = 32 / LANES;                   // runs walked side by side by one warp
    constexpr int S = LANES;                          // steps per record tile (a sub-warp stages LANES records at a time)
    constexpr int RING = 4 * LANES;                   // records of a run resident in shared memory (4 tiles)
    static_assert((D & (D - 1)) == 0 && S >= D && S % D == 0, "pipeline depth");
    __shared__ __align__(16) float4 srow[8][D][32];   // per warp: D slots of one row per sub-warp (512 B each)
    __shared__ int2 srec[8][GPW][RING];               // per sub-warp: ring of (u, r bits), record j of the run at j % RING
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int wic = threadIdx.x >> 5;
    const int gl = lane & (LANES - 1);
    const int grp = lane / LANES;
    const int64_t k = FULL ? (int64_t)(4 * LANES) : (int64_t)a.k;
    const int chunks = (int)(k >> 2);
    const bool lane_on = FULL || gl < chunks;
    float* const Pl = a.P - (int64_t)a.u_base * k + 4 * gl;
    const Coef cf = {a.lr, a.lambda, __fsub_rn(1.0f, __fmul_rn(a.lr, a.lambda))};
    const int32_t* __restrict__ words = reinterpret_cast<const int32_t*>(a.recs);
    const uint64_t pol = l2_policy_evict_first();
    float4(*const rows)[32] = srow[wic];
    int2* const recs = srec[wic][grp];
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
        const float4 q0 = (has && lane_on) ? ld_row4(qrow) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 q = q0;
        auto stage_tile = [&](int tile) {               // lane gl copies (u, r) of record LANES*tile + gl of its run
            const int j = tile * LANES + gl;
            if (j < count) {
                const int64_t pos = hu.start + j;
                const int64_t idx = virt ? hu.bstart + (int64_t)block_perm((uint64_t)(pos - hu.bstart), (uint64_t)hu.bn, vhb, vkey) : pos;
                cp_async_4_stream(&recs[j & (RING - 1)].x, words + 3 * idx, pol);
                cp_async_4_stream(&recs[j & (RING - 1)].y, words + 3 * idx + 2, pol);
            }
        };
        // prologue: tiles 0 and 1 staged and visible, then D-1 rows in flight (one group each)
        stage_tile(0);
        stage_tile(1);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
#pragma unroll
        for (int d = 0; d < D - 1; d++) {
            if (d < count && lane_on) cp_async_16(&rows[d][lane], Pl + (int64_t)recs[d].x * k);
            cp_async_commit();
        }
        for (int t0 = 0; t0 < steps; t0 += D) {
            const bool tile_top = (t0 & (S - 1)) == 0;  // warp-uniform
            if (tile_top) stage_tile(t0 / S + 2);       // joins the group of step t0
#pragma unroll
            for (int d = 0; d < D; d++) {
                const int t = t0 + d;
                if (t >= steps) break;                  // warp-uniform
                {   // gather the row of step t + D - 1 into the slot step t - 1 has just released
                    const int tp = t + D - 1;
                    if (tp < count && lane_on) cp_async_16(&rows[(d + D - 1) & (D - 1)][lane], Pl + (int64_t)recs[tp & (RING - 1)].x * k);
                }
                cp_async_commit();
                cp_async_wait<D - 1>();
                if (d == 0 && tile_top) __syncwarp();   // record tiles staged >= D steps ago: visible to every lane
                const int2 rec = recs[t & (RING - 1)];
                float4 p = rows[d][lane];
                if ((!FULL && !lane_on) || (GPW > 1 && t >= count)) p = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 pa[1] = {p}, qa[1] = {q};
                const float e = __fsub_rn(__int_as_float(rec.y), rows_dot<LANES, 1, FAST>(pa, qa));
                const float b = __fmul_rn(cf.lr, e);
                if (t < count && lane_on) {
                    st_row4(Pl + (int64_t)rec.x * k, new_chunk<FAST>(p, q, e, cf.lr, cf.lambda, cf.acoef, b));
                    q = new_chunk<FAST>(q, p, e, cf.lr, cf.lambda, cf.acoef, b);
                }
            }
        }
        cp_async_wait<0>();
        __syncwarp();                                   // every lane is done with the runs' tiles and slots
        if (has && lane_on) {
            if (hu.weight == 1.0f) {
                st_row4(qrow, q);
            } else {
                const float w = hu.weight;
                red_add_row4(qrow, make_float4(__fmul_rn(__fsub_rn(q.x, q0.x), w), __fmul_rn(__fsub_rn(q.y, q0.y), w),
                                               __fmul_rn(__fsub_rn(q.z, q0.z), w), __fmul_rn(__fsub_rn(q.w, q0.w), w)));
            }
        }
    }
    pdl_wait_prerequisites();
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
        const Rec rec = a.recs[a.first + j];
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

static int pipeline_depth() {   // MFSGD_DEPTH = 2 | 4 (tuning aid); gathers in flight per sub-warp = depth - 1
    static int depth = 0;
    if (depth == 0) {
        const char* e = getenv("MFSGD_DEPTH");
        depth = (e && atoi(e) == 4) ? 4 : 2;
    }
    return depth;
}

static int hot_depth() {   // MFSGD_HDEPTH = 4 | 8: p_u gathers kept in flight per run (depth - 1)
    static int depth = 0;
    if (depth == 0) {
        const char* e = getenv("MFSGD_HDEPTH");
        depth = (e && atoi(e) == 8) ? 8 : 4;
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
#define CALL(L, V, F)                                                                                    \
    if (fast && deep) sgd_update_hogwild_kernel<L, V, F, 0, true, 4><<<grid, 256, 0, stream>>>(a);       \
    else if (fast) sgd_update_hogwild_kernel<L, V, F, 0, true, 2><<<grid, 256, 0, stream>>>(a);          \
    else switch (scatter) {                                                                              \
        case 1: sgd_update_hogwild_kernel<L, V, F, 1, false, 2><<<grid, 256, 0, stream>>>(a); break;     \
        case 2: sgd_update_hogwild_kernel<L, V, F, 2, false, 2><<<grid, 256, 0, stream>>>(a); break;     \
        case 3: sgd_update_hogwild_kernel<L, V, F, 3, false, 2><<<grid, 256, 0, stream>>>(a); break;     \
        default: sgd_update_hogwild_kernel<L, V, F, 0, false, 2><<<grid, 256, 0, stream>>>(a); break;    \
    }
    MFSGD_DISPATCH_GEOMETRY(g, CALL);
#undef CALL
    if (launches) *launches += 1;
    return cudaGetLastError();
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}
// MFSGD_HOT_ASYNC = 0 keeps the register-ring kernel for every rank (tuning aid); default: cp.async pipeline for
// ranks with one float4 per lane and >= 8 lanes per rating (32 <= k <= 128).
static bool hot_async_for(const Geometry& g) {
    static int on = -1;
    if (on < 0) on = env_int("MFSGD_HOT_ASYNC", 1);
    return on != 0 && g.vec == 1 && g.lanes >= 8;
}
// MFSGD_PDL = 0: plain stream order between the hot launches of consecutive visits. Default: programmatic
// dependent launch, so a visit's first runs fill the SMs the previous visit's last runs no longer occupy.
// MFSGD_HDEPTH = 2 | 4 | 8: slots of the cp.async row ring (depth - 1 gathers in flight per warp). Measured on the
// Netflix-shaped workload: 4 beats 8 (the launch is L2-throughput-bound, deeper queues only add pressure).
static int hot_async_depth() {
    static int dep = 0;
    if (dep == 0) {
        dep = env_int("MFSGD_HDEPTH", 4);
        if (dep != 2 && dep != 8) dep = 4;
    }
    return dep;
}
static bool hot_pdl() {
    static int on = -1;
    if (on < 0) on = env_int("MFSGD_PDL", 1);
    return on != 0;
}

template <typename Kernel>
static cudaError_t launch_hot(Kernel kernel, int grid, cudaStream_t stream, bool pdl, const UpdateArgs& a, const HotUnit* units,
                              int n_units, unsigned int* counter) {
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
    return cudaLaunchKernelEx(&cfg, kernel, a, units, n_units, counter);
}

cudaError_t launch_sgd_update_hot(const UpdateArgs& a, const HotUnit* units, int n_units, unsigned int* counter, bool fast,
                                  int grid, bool follows_hot_launch, cudaStream_t stream, int* launches) {
    if (n_units <= 0) return cudaSuccess;
    const Geometry g = geometry_for(a.k);
    // 8 warps per CTA, one unit per warp at a time -- and at least two waves of units per launch: the runs of an item that
    // are in flight together start from the same q_i and are averaged, the next wave builds on their result.
    const int per_warp = hot_async_for(g) ? 32 / g.lanes : 1;     // runs a warp walks side by side
    const int full_grid = grid;
    const int max_grid = (n_units + 16 * per_warp - 1) / (16 * per_warp);
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    const bool deep = hot_depth() == 8;
    cudaError_t err = cudaSuccess;
    if (hot_async_for(g)) {
        // Overlap with the previous visit's launch only where it is a tail effect: both launches fill the machine (this
        // one offers >= 2 runs per resident sub-warp), so this grid's CTAs become resident as the other's retire.
        // MFSGD_PDL = 2 overlaps every chained launch (tuning aid).
        const bool pdl = follows_hot_launch && hot_pdl() &&
                         (env_int("MFSGD_PDL", 1) == 2 || (int64_t)n_units >= 2LL * full_grid * 8 * per_warp);
        const int dep = hot_async_depth();
#define CALLA(L, F)                                                                                                                 \
    err = (fast && dep == 2) ? launch_hot(sgd_update_hot_async_kernel<L, F, true, 2>, grid, stream, pdl, a, units, n_units, counter)  \
          : (fast && dep == 8) ? launch_hot(sgd_update_hot_async_kernel<L, F, true, 8>, grid, stream, pdl, a, units, n_units, counter) \
          : fast       ? launch_hot(sgd_update_hot_async_kernel<L, F, true, 4>, grid, stream, pdl, a, units, n_units, counter)      \
                       : launch_hot(sgd_update_hot_async_kernel<L, F, false, 4>, grid, stream, pdl, a, units, n_units, counter)
        switch (g.lanes) {
            case 8:  if (g.full) { CALLA(8, true); }  else { CALLA(8, false); }  break;
            case 16: if (g.full) { CALLA(16, true); } else { CALLA(16, false); } break;
            default: if (g.full) { CALLA(32, true); } else { CALLA(32, false); } break;
        }
#undef CALLA
        if (launches) *launches += 1;
        return err != cudaSuccess ? err : cudaGetLastError();
    }
#define CALL(L, V, F)                                                                                                  \
    if (fast && deep) sgd_update_hot_kernel<L, V, F, true, 8><<<grid, 256, 0, stream>>>(a, units, n_units, counter);   \
    else if (fast) sgd_update_hot_kernel<L, V, F, true, 4><<<grid, 256, 0, stream>>>(a, units, n_units, counter);      \
    else sgd_update_hot_kernel<L, V, F, false, 4><<<grid, 256, 0, stream>>>(a, units, n_units, counter)
    MFSGD_DISPATCH_GEOMETRY(g, CALL);
#undef CALL
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t hogwild_max_ctas_per_sm(int k, int scatter, bool fast, int* ctas) {
    const Geometry g = geometry_for(k);
    cudaError_t err = cudaSuccess;
    if (scatter != 0) fast = false;
    const bool deep = pipeline_depth() == 4 && g.vec == 1;
#define CALL(L, V, F)                                                                                                          \
    err = (fast && deep) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hogwild_kernel<L, V, F, 0, true, 4>, 256, 0) \
          : fast ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hogwild_kernel<L, V, F, 0, true, 2>, 256, 0)       \
                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hogwild_kernel<L, V, F, 0, false, 2>, 256, 0)
    MFSGD_DISPATCH_GEOMETRY(g, CALL);
#undef CALL
    return err;
}

int hot_sub_warps_per_run(int k) {
    const Geometry g = geometry_for(k);
    return hot_async_for(g) ? 1 : 32 / g.lanes;
}

cudaError_t hot_max_ctas_per_sm(int k, bool fast, int* ctas) {
    const Geometry g = geometry_for(k);
    cudaError_t err = cudaSuccess;
    const bool deep = hot_depth() == 8;
    if (hot_async_for(g)) {
        const int dep = hot_async_depth();
#define CALLA(L, F)                                                                                                                    \
    err = (fast && dep == 2) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hot_async_kernel<L, F, true, 2>, 256, 0)   \
          : (fast && dep == 8) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hot_async_kernel<L, F, true, 8>, 256, 0) \
          : fast       ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hot_async_kernel<L, F, true, 4>, 256, 0)       \
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hot_async_kernel<L, F, false, 4>, 256, 0)
        switch (g.lanes) {
            case 8:  if (g.full) { CALLA(8, true); }  else { CALLA(8, false); }  break;
            case 16: if (g.full) { CALLA(16, true); } else { CALLA(16, false); } break;
            default: if (g.full) { CALLA(32, true); } else { CALLA(32, false); } break;
        }
#undef CALLA
        const int cap = env_int("MFSGD_HOT_CTAS", 0);    // tuning aid: resident hot-kernel CTAs per SM
        if (err == cudaSuccess && cap > 0 && *ctas > cap) *ctas = cap;
        return err;
    }
#define CALL(L, V, F)                                                                                                        \
    err = (fast && deep) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hot_kernel<L, V, F, true, 8>, 256, 0)  \
          : fast ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hot_kernel<L, V, F, true, 4>, 256, 0)          \
                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, sgd_update_hot_kernel<L, V, F, false, 4>, 256, 0)
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
