// update_math.cuh -- arithmetic shared by the full-grid update kernels (kernels_update.cu, kernels_hot.cu).
#pragma once
#include "common.cuh"

namespace mfsgd {

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

// the increment of row chunk `o` alone: new_chunk - o = b * x + (a - 1) * o, with ccoef = a - 1 = -(lr * lambda) exactly.
// Used where the row is updated in memory by red.global.add (no lost updates between concurrent writers of a row).
template <bool FAST>
__device__ __forceinline__ float4 delta_chunk(float4 o, float4 x, float e, float lr, float lambda, float ccoef, float b) {
    if (FAST) {
        const uint64_t c2 = pk2(ccoef, ccoef), b2 = pk2(b, b);
        const uint64_t lo = fma2(b2, pk2(x.x, x.y), mul2(c2, pk2(o.x, o.y)));
        const uint64_t hi = fma2(b2, pk2(x.z, x.w), mul2(c2, pk2(o.z, o.w)));
        float4 r;
        upk2(lo, r.x, r.y);
        upk2(hi, r.z, r.w);
        return r;
    }
    return delta4(o, x, e, lr, lambda);
}

// Model extension (MatrixFactorizationSGD.java:282 sgdUpdateModel): the increment of a bias, lr * (e - lambda * b),
// one rounding per operation in every arithmetic.
__device__ __forceinline__ float bias_delta(float b, float e, float lr, float lambda) {
    return __fmul_rn(lr, __fsub_rn(e, __fmul_rn(lambda, b)));
}

struct Coef {
    float lr, lambda, acoef;   // acoef = 1 - lr * lambda (FAST arithmetic)
};

}  // namespace mfsgd
