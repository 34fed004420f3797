// common.cuh -- shared device helpers of libmfsgd.so (sm_100a only).
//
// Arithmetic contract: every float op of the update rule is an explicit round-to-
// nearest intrinsic (__fmul_rn/__fadd_rn/__fsub_rn), so nvcc can never contract a*b+c into an FMA;
// this mirrors the strict binary32 semantics of baseline/java/MatrixFactorizationSGD.java:89-105.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mfsgd {

// 12-byte rating record (BASELINE.json north_star: "12-byte (u, i, r) records").
struct Rec {
    int32_t u;
    int32_t i;
    float   r;
};
static_assert(sizeof(Rec) == 12, "Rec must be 12 bytes");
// Bit 31 of Rec::u in the bucketed (Hogwild / DSGD) layouts marks a HEAVY user: one whose ratings are so many that several of
// them are in flight at once, so that its row is updated in memory with red.global.add (no lost updates) instead of a store
// (engine.cu heavy_user_threshold). Ids are < 2^31, every reader masks the bit.
constexpr int32_t REC_USER_MASK = 0x7fffffff;

enum : uint64_t {
    STREAM_P_INIT = 0, STREAM_Q_INIT = 1, STREAM_SHUFFLE = 2, STREAM_USER = 3, STREAM_ITEM = 4,
    STREAM_NOISE = 5, STREAM_HELDOUT = 6, STREAM_PSTAR = 7, STREAM_QSTAR = 8,
    STREAM_BLOCK_SHUFFLE = 9   // Hogwild/DSGD in-block reshuffle keys (not in the stand-in)
};

// MatrixFactorizationSGD.java:39 hash64 (SplitMix64 finaliser over seed, stream, counter).
__host__ __device__ __forceinline__ uint64_t hash64(uint64_t seed, uint64_t stream, uint64_t ctr) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (ctr + 1ULL) + 0xD1B54A32D192ED03ULL * stream;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return z;
}

// MatrixFactorizationSGD.java:48 uniform: 24 random bits -> [0,1), exact in binary32.
__device__ __forceinline__ float uniform24(uint64_t seed, uint64_t stream, uint64_t ctr) {
    return __fmul_rn((float)(uint32_t)(hash64(seed, stream, ctr) >> 40), 0x1.0p-24f);
}

// ---- cache-policy loads/stores -------------------------------------------------------------
// Factor rows are read and written by every SM: keep them out of the (incoherent) L1 and resident
// in L2 (ld/st .cg). Records are streamed once per epoch: no L1 allocation, evict-first in L2.
__device__ __forceinline__ float4 ld_row4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_row4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }

// experiment variants of the row store (selected by scatter codes 4..6, see kernels_update.cu)
__device__ __forceinline__ void st_row4_wb(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_row4_wt(float* p, float4 v) { __stwt(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st_row4_cs(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

__device__ __forceinline__ void red_add_row4(float* p, float4 d) {
    // sm_90+: vectorised fire-and-forget float atomic add, resolved in L2.
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(d.x), "f"(d.y), "f"(d.z), "f"(d.w)
                 : "memory");
}

// ---- mixed-precision factor storage ---------------------------
// mfsgd_config.p_storage = MFSGD_STORAGE_F16: the rows of P are KEPT as binary16 (half the L2 / HBM bytes of the update's
// dominant stream), every operation of the rule stays binary32. A chunk of 4 values is widened exactly on load and narrowed on
// store with stochastic rounding: 8 random bits per value decide among the 13 dropped mantissa bits, drawn from a counter hash
// of (seed, epoch, u, i, chunk) -- stateless, the same bits in any visiting order, restated by the oracle bit for bit.
template <bool PH> struct PChunk { using type = float4; };     // what one lane holds of a P row: 4 consecutive values
template <> struct PChunk<true> { using type = uint2; };       // ... as 4 binary16
template <bool PH> __host__ __device__ constexpr int p_elem_bytes() { return PH ? 2 : 4; }

__device__ __forceinline__ float4 widen4(float4 v) { return v; }
__device__ __forceinline__ float4 widen4(uint2 h) {
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
template <bool PH>
__device__ __forceinline__ typename PChunk<PH>::type ld_pchunk(const char* p) {
    return __ldcg(reinterpret_cast<const typename PChunk<PH>::type*>(p));
}
template <bool PH>
__device__ __forceinline__ typename PChunk<PH>::type zero_pchunk() {
    if constexpr (PH) return make_uint2(0u, 0u);
    else return make_float4(0.f, 0.f, 0.f, 0.f);
}
__host__ __device__ __forceinline__ uint32_t sr_seed32(uint64_t seed) { return (uint32_t)(seed ^ (seed >> 32)); }
// the 32 random bits of chunk c of row u at the update (u, i) of `epoch` (lowbias32 finaliser)
__host__ __device__ __forceinline__ uint32_t sr_word(uint32_t s32, uint32_t epoch, uint32_t u, uint32_t i, uint32_t c) {
    uint32_t x = s32 + u * 0x9E3779B1u + i * 0x85EBCA77u + (epoch * 0x10001u + c) * 0xC2B2AE3Du;
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}
// value j of a chunk, pushed up by its random offset and cut to the 10 mantissa bits binary16 keeps;
// the conversion that follows is exact for normal results and rounds to nearest in binary16's subnormal range
__device__ __forceinline__ float sr_prepare(float v, uint32_t w, int j) {
    const uint32_t rho = (((w >> (8 * j)) & 0xFFu) << 5) | 0x10u;
    return __uint_as_float((__float_as_uint(v) + rho) & 0xFFFFE000u);
}
__device__ __forceinline__ uint2 narrow4_sr(float4 v, uint32_t w) {
    const __half2 lo = __floats2half2_rn(sr_prepare(v.x, w, 0), sr_prepare(v.y, w, 1));
    const __half2 hi = __floats2half2_rn(sr_prepare(v.z, w, 2), sr_prepare(v.w, w, 3));
    return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
__device__ __forceinline__ uint2 narrow4_rn(float4 v) {
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
// store the new value of a P chunk (binary32: as is; binary16: stochastic rounding with the update's random word)
__device__ __forceinline__ void st_pchunk(char* p, float4 v, float4, uint32_t) { __stcg(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st_pchunk(char* p, float4 v, uint2, uint32_t w) { __stcg(reinterpret_cast<uint2*>(p), narrow4_sr(v, w)); }
// heavy users, binary16: the row in memory moves by (narrowed new value - the binary16 value this update started from) -- an
// exact binary16 difference (neighbouring values), added in L2 by one vector red; with no concurrent writer the row ends on
// the narrowed new value exactly, like the store
__device__ __forceinline__ void red_pchunk_f16(char* p, float4 newv, uint2 old, uint32_t w) {
    const uint2 nv = narrow4_sr(newv, w);
    const __half2 dlo = __hsub2(*reinterpret_cast<const __half2*>(&nv.x), *reinterpret_cast<const __half2*>(&old.x));
    const __half2 dhi = __hsub2(*reinterpret_cast<const __half2*>(&nv.y), *reinterpret_cast<const __half2*>(&old.y));
    asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(p), "r"(*reinterpret_cast<const uint32_t*>(&dlo)),
                 "r"(*reinterpret_cast<const uint32_t*>(&dhi))
                 : "memory");
}

// L2 eviction policy (createpolicy): records are read once per epoch -> evict-first.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int32_t ld_stream_i32(const int32_t* p, uint64_t pol) {
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ void st_stream_i32(int32_t* p, int32_t v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

// ---- in-block reshuffle: tile-coherent keyed bijection (subsystem 1) -----------------------------
// perm_b(j) for bucket b of n records in epoch e: the materialising kernel (kernels_layout.cu) writes
// out[off+j] = in[off+perm(j)]; the update kernels can instead read record off+perm(j) directly.
//
// Round 2: the permutation keeps every aligned group of 32 consecutive positions inside ONE aligned group of 32
// records (384 contiguous bytes = 12 sectors), so a warp staging 32 positions of a bucket reads whole sectors instead
// of 32 scattered 12-byte records (round 1: one or two 32-B sectors from DRAM per 12-B record, 10.7 GB of DRAM
// traffic per Netflix-shaped epoch against 1.08 GB of records). Per epoch the ORDER OF THE TILES is a keyed 4-round
// balanced Feistel bijection with cycle walking over the n/32 full tiles, the order INSIDE a tile a keyed affine
// bijection of the 5 lane bits, and the < 32 records behind the last full tile are rotated by a keyed offset.
// The CPU checker restates this function (tests/test_gpu_parity.py compares the two bit for bit).
__host__ __device__ __forceinline__ uint32_t feistel_round(uint32_t x, uint32_t key) {
    uint32_t h = (x + key) * 0x9E3779B1u;
    h ^= h >> 15;
    h *= 0x85EBCA77u;
    h ^= h >> 13;
    return h;
}
// bijection on [0, n), n > 1; hb = perm_half_bits(n)
__host__ __device__ __forceinline__ uint64_t block_perm(uint64_t x, uint64_t n, int hb, uint64_t key) {
    const uint64_t tiles = n >> 5, full = tiles << 5;
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    if (x >= full) {                                   // the partial tile at the end: rotate
        const uint32_t rem = (uint32_t)(n - full);
        return full + (uint64_t)(((uint32_t)(x - full) + (k1 >> 8) % rem) % rem);
    }
    uint64_t t = x >> 5;
    if (tiles > 1) {                                   // tile order: Feistel over 2*hb bits (2^(2hb) >= tiles), cycle walking
        const uint32_t mask = (hb >= 32) ? 0xffffffffu : ((1u << hb) - 1u);
        do {
            uint32_t l = (uint32_t)(t >> hb) & mask, r = (uint32_t)t & mask;
#pragma unroll
            for (int round = 0; round < 4; round++) {
                const uint32_t f = feistel_round(r, (round & 1) ? (k1 + round) : (k0 + round)) & mask;
                const uint32_t nl = r;
                r = l ^ f;
                l = nl;
            }
            t = ((uint64_t)l << hb) | (uint64_t)r;
        } while (t >= tiles);
    }
    const uint32_t h = feistel_round((uint32_t)t ^ k1, k0);   // lane order inside the tile: ((l ^ c) * a + b) mod 32, a odd
    const uint32_t lane = ((((uint32_t)x & 31u) ^ ((h >> 16) & 31u)) * ((h & 31u) | 1u) + ((h >> 8) & 31u)) & 31u;
    return (t << 5) | (uint64_t)lane;
}

__host__ __device__ __forceinline__ int perm_half_bits(uint64_t n) {   // half the bits of the Feistel domain over the n/32 tiles
    const uint64_t tiles = n >> 5;
    if (tiles <= 1) return 1;
#ifdef __CUDA_ARCH__
    int bits = 64 - __clzll((long long)(tiles - 1));
#else
    int bits = 0;
    while (bits < 63 && (1ULL << bits) < tiles) bits++;
#endif
    if (bits < 2) bits = 2;
    return (bits + 1) >> 1;
}
__host__ __device__ __forceinline__ uint64_t bucket_perm_key(uint64_t seed, uint32_t epoch, uint32_t bucket_id) {
    return hash64(seed, STREAM_BLOCK_SHUFFLE, ((uint64_t)epoch << 32) | (uint64_t)bucket_id);
}

// ---- the update rule -----------------------------------------------------------------------
// Partial dot of one float4 chunk, continuing a running binary32 sum, element order x,y,z,w.
__device__ __forceinline__ float dot4_acc(float acc, float4 a, float4 b) {
    acc = __fadd_rn(acc, __fmul_rn(a.x, b.x));
    acc = __fadd_rn(acc, __fmul_rn(a.y, b.y));
    acc = __fadd_rn(acc, __fmul_rn(a.z, b.z));
    acc = __fadd_rn(acc, __fmul_rn(a.w, b.w));
    return acc;
}

// xor butterfly over a LANES-wide sub-warp, masks LANES/2 .. 1 (oracle.cpp
// dot_warp_tree reproduces this order bit for bit). All 32 lanes of the warp must call it.
template <int LANES>
__device__ __forceinline__ float group_sum(float s) {
#pragma unroll
    for (int m = LANES >> 1; m >= 1; m >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, m));
    return s;
}

// new = old + lr * (e * other - lambda * old)      MatrixFactorizationSGD.java:100-101
__device__ __forceinline__ float upd1(float old_, float other, float e, float lr, float lambda) {
    return __fadd_rn(old_, __fmul_rn(lr, __fsub_rn(__fmul_rn(e, other), __fmul_rn(lambda, old_))));
}
// the increment alone (for the atomic scatter): lr * (e * other - lambda * old)
__device__ __forceinline__ float delta1(float old_, float other, float e, float lr, float lambda) {
    return __fmul_rn(lr, __fsub_rn(__fmul_rn(e, other), __fmul_rn(lambda, old_)));
}
__device__ __forceinline__ float4 upd4(float4 o, float4 x, float e, float lr, float lambda) {
    return make_float4(upd1(o.x, x.x, e, lr, lambda), upd1(o.y, x.y, e, lr, lambda),
                       upd1(o.z, x.z, e, lr, lambda), upd1(o.w, x.w, e, lr, lambda));
}
__device__ __forceinline__ float4 delta4(float4 o, float4 x, float e, float lr, float lambda) {
    return make_float4(delta1(o.x, x.x, e, lr, lambda), delta1(o.y, x.y, e, lr, lambda),
                       delta1(o.z, x.z, e, lr, lambda), delta1(o.w, x.w, e, lr, lambda));
}

}  // namespace mfsgd
