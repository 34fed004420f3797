// tile_kernel.cu -- DRAFT of the user-tile SGD kernel planned in DESIGN.md section 9. NOT part of libmfsgd.so, not built by
// `make`, and NEVER RUN ON HARDWARE in round 1 (it was written after the round's GPU minutes were spent; it only compiles:
// nvcc -gencode arch=compute_100a,code=sm_100a). It exists so that round 2 can start from a concrete candidate whose
// semantics are exactly those of the CPU study tools/tile_sim (profiles/r01_tile_sim.md); tools/tile_kernel_draft/check.py
// is the harness that has to pass before any of it moves into the product.
//
// Kernel (k = 128 only in this draft): persistent CTAs of 16 warps, one per SM. A CTA claims a tile (a contiguous range of
// <= TILE_USERS users, visiting order given per epoch), copies the tile's P rows into shared memory (128 KB), then its
// warps claim chunks of CHUNK records of the tile (sorted by item on the host side) and walk them run by run: q_i in
// registers (the NEXT run's q_i is prefetched while the current run is applied), p_u in shared memory (LDS.128 / STS.128,
// one float4 per lane), FFMA2 dot + xor butterfly as in kernels_hot.cu, and at the end of a run
// red.global.add.v4.f32 of weight[i] * (q_i' - q_i). Finally the tile is written back. Races between the warps of a CTA on
// one p_u row are Hogwild in shared memory.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <vector>

namespace {

constexpr int K = 128;
constexpr int TILE_USERS = 256;          // rows of P resident per CTA: 256 * 512 B = 128 KB
constexpr int WARPS = 16;
constexpr int CHUNK = 64;                // records a warp claims at a time

struct Rec {
    int32_t u, i;
    float r;
};

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
// the FAST arithmetic of update_math.cuh at one float4 chunk per lane (oracle twin: ORC_ORDER_WARP_TREE_FMA, 32 lanes)
__device__ __forceinline__ float dot128(float4 p, float4 q) {
    uint64_t acc = mul2(pk2(p.x, p.y), pk2(q.x, q.y));
    acc = fma2(pk2(p.z, p.w), pk2(q.z, q.w), acc);
    float lo, hi;
    upk2(acc, lo, hi);
    float s = __fadd_rn(lo, hi);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, m));
    return s;
}
__device__ __forceinline__ float4 new_chunk(float4 o, float4 x, float acoef, float b) {
    const uint64_t a2 = pk2(acoef, acoef), b2 = pk2(b, b);
    const uint64_t lo = fma2(b2, pk2(x.x, x.y), mul2(a2, pk2(o.x, o.y)));
    const uint64_t hi = fma2(b2, pk2(x.z, x.w), mul2(a2, pk2(o.z, o.w)));
    float4 r;
    upk2(lo, r.x, r.y);
    upk2(hi, r.z, r.w);
    return r;
}
__device__ __forceinline__ void red_add4(float* p, float4 d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(d.x), "f"(d.y), "f"(d.z), "f"(d.w) : "memory");
}

struct TileArgs {
    const Rec* recs;               // sorted by (visit, item order); visit = (pass, user tile)
    const int64_t* visit_off;      // n_visits + 1
    const int32_t* visit_user_lo;  // n_visits: first user row of the visit's tile
    const int32_t* visit_users;    // n_visits: rows in the tile (<= TILE_USERS)
    const int32_t* visit_order;    // n_visits: this epoch's visiting order
    int32_t n_visits;
    float* P;
    float* Q;
    const float* weight;           // per item: merge weight
    float lr, lambda;
    int32_t active_warps;          // 16; 1 = strictly sequential (checking mode, with a grid of one CTA)
    unsigned int* visit_counter;   // zeroed before the launch
};

__global__ void __launch_bounds__(WARPS * 32, 1) sgd_update_tile_kernel(TileArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* const sP = reinterpret_cast<float4*>(smem_raw);                        // TILE_USERS * 32 float4
    int4* const sRec = reinterpret_cast<int4*>(smem_raw + (size_t)TILE_USERS * 512);  // WARPS * CHUNK records (u_local, i, r bits, -)
    __shared__ int s_visit;
    __shared__ unsigned long long s_next;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float acoef = __fsub_rn(1.0f, __fmul_rn(a.lr, a.lambda));
    int4* const myrec = sRec + warp * CHUNK;
    for (;;) {
        if (tid == 0) s_visit = (int)atomicAdd(a.visit_counter, 1u);
        __syncthreads();
        const int slot = s_visit;
        if (slot >= a.n_visits) break;
        const int v = a.visit_order[slot];
        const int u_lo = a.visit_user_lo[v], nu = a.visit_users[v];
        const int64_t r_lo = a.visit_off[v];
        const long long n_rec = (long long)(a.visit_off[v + 1] - r_lo);
        // (1) the tile's rows -> shared memory
        {
            const float4* src = reinterpret_cast<const float4*>(a.P + (int64_t)u_lo * K);
            for (int idx = tid; idx < nu * 32; idx += WARPS * 32) sP[idx] = __ldcg(src + idx);
        }
        if (tid == 0) s_next = 0ULL;
        __syncthreads();
        // (2) warps claim chunks and walk them run by run
        if (warp < a.active_warps) {
            for (;;) {
                unsigned long long c = 0;
                if (lane == 0) c = atomicAdd(&s_next, (unsigned long long)CHUNK);
                c = __shfl_sync(0xffffffffu, c, 0);
                if ((long long)c >= n_rec) break;
                const int cnt = (int)min((long long)CHUNK, n_rec - (long long)c);
                __syncwarp();
                for (int j = lane; j < cnt; j += 32) {
                    const Rec rc = a.recs[r_lo + (int64_t)c + j];
                    myrec[j] = make_int4(rc.u - u_lo, rc.i, __float_as_int(rc.r), 0);
                }
                __syncwarp();
                // run state: current item, its factor and snapshot; next run's factor prefetched
                int item = myrec[0].y;
                float4 q = __ldcg(reinterpret_cast<const float4*>(a.Q + (int64_t)item * K) + lane);
                float4 q0 = q;
                float w = a.weight[item];
                int j = 0;
                while (j < cnt) {
                    // end of the current run: first record at or after j whose item differs (warp-cooperative scan)
                    int run_end = cnt;
                    for (int base = j; base < cnt; base += 32) {
                        const int jj = base + lane;
                        const bool diff = jj < cnt && myrec[jj].y != item;
                        const unsigned m = __ballot_sync(0xffffffffu, diff);
                        if (m) { run_end = base + __ffs(m) - 1; break; }
                    }
                    // prefetch the next run's factor while this run is applied
                    const int next_item = run_end < cnt ? myrec[run_end].y : -1;
                    float4 qn = make_float4(0.f, 0.f, 0.f, 0.f);
                    float wn = 0.f;
                    if (next_item >= 0) {
                        qn = __ldcg(reinterpret_cast<const float4*>(a.Q + (int64_t)next_item * K) + lane);
                        wn = a.weight[next_item];
                    }
                    for (; j < run_end; j++) {
                        const int4 rc = myrec[j];
                        float4* const prow = sP + rc.x * 32 + lane;
                        const float4 p = *prow;
                        const float e = __fsub_rn(__int_as_float(rc.z), dot128(p, q));
                        const float b = __fmul_rn(a.lr, e);
                        *prow = new_chunk(p, q, acoef, b);
                        q = new_chunk(q, p, acoef, b);
                    }
                    // merge the run
                    red_add4(a.Q + (int64_t)item * K + 4 * lane,
                             make_float4(__fmul_rn(__fsub_rn(q.x, q0.x), w), __fmul_rn(__fsub_rn(q.y, q0.y), w),
                                         __fmul_rn(__fsub_rn(q.z, q0.z), w), __fmul_rn(__fsub_rn(q.w, q0.w), w)));
                    if (next_item >= 0) {
                        // NOTE for round 2: if next_item == item cannot happen (runs are maximal); but the prefetched qn
                        // predates this run's merge when another run of next_item is in flight -- that is the model.
                        item = next_item;
                        q = qn;
                        q0 = qn;
                        w = wn;
                    }
                }
            }
        }
        __syncthreads();
        // (3) write the tile back
        {
            float4* dst = reinterpret_cast<float4*>(a.P + (int64_t)u_lo * K);
            for (int idx = tid; idx < nu * 32; idx += WARPS * 32) __stcg(dst + idx, sP[idx]);
        }
        __syncthreads();
    }
}

#define CK(x)                                                                                    \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess) {                                                                 \
            fprintf(stderr, "tile draft: %s: %s\n", #x, cudaGetErrorString(e_));                 \
            return -1;                                                                           \
        }                                                                                        \
    } while (0)

}  // namespace

// Host harness entry (ctypes): runs `epochs` epochs on device 0 over a host-built layout; P and Q are updated in place.
// visit_order: epochs * n_visits entries. grid = 0 -> one CTA per SM. Returns the mean epoch time in ms through *ms_per_epoch.
extern "C" int tiledraft_train(const int32_t* recs3, int64_t n_rec, const int64_t* visit_off, const int32_t* visit_user_lo,
                               const int32_t* visit_users, int32_t n_visits, const int32_t* visit_order, int32_t epochs, float* P,
                               int64_t n_users, float* Q, int64_t n_items, const float* weight, float lr, float lambda, int32_t grid,
                               int32_t active_warps, double* ms_per_epoch) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    if (grid <= 0) grid = prop.multiProcessorCount;
    const size_t smem = (size_t)TILE_USERS * 512 + (size_t)WARPS * CHUNK * sizeof(int4);
    CK(cudaFuncSetAttribute(sgd_update_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Rec* d_recs = nullptr;
    int64_t* d_off = nullptr;
    int32_t *d_lo = nullptr, *d_nu = nullptr, *d_order = nullptr;
    float *d_P = nullptr, *d_Q = nullptr, *d_w = nullptr;
    unsigned int* d_counter = nullptr;
    CK(cudaMalloc(&d_recs, (size_t)n_rec * sizeof(Rec)));
    CK(cudaMalloc(&d_off, ((size_t)n_visits + 1) * 8));
    CK(cudaMalloc(&d_lo, (size_t)n_visits * 4));
    CK(cudaMalloc(&d_nu, (size_t)n_visits * 4));
    CK(cudaMalloc(&d_order, (size_t)n_visits * 4 * (size_t)epochs));
    CK(cudaMalloc(&d_P, (size_t)n_users * K * 4));
    CK(cudaMalloc(&d_Q, (size_t)n_items * K * 4));
    CK(cudaMalloc(&d_w, (size_t)n_items * 4));
    CK(cudaMalloc(&d_counter, 4));
    CK(cudaMemcpy(d_recs, recs3, (size_t)n_rec * sizeof(Rec), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_off, visit_off, ((size_t)n_visits + 1) * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_lo, visit_user_lo, (size_t)n_visits * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_nu, visit_users, (size_t)n_visits * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_order, visit_order, (size_t)n_visits * 4 * (size_t)epochs, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_P, P, (size_t)n_users * K * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_Q, Q, (size_t)n_items * K * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, weight, (size_t)n_items * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int ep = 0; ep < epochs; ep++) {
        CK(cudaMemsetAsync(d_counter, 0, 4));
        TileArgs a{};
        a.recs = d_recs; a.visit_off = d_off; a.visit_user_lo = d_lo; a.visit_users = d_nu;
        a.visit_order = d_order + (size_t)ep * n_visits;
        a.n_visits = n_visits;
        a.P = d_P; a.Q = d_Q; a.weight = d_w; a.lr = lr; a.lambda = lambda;
        a.active_warps = active_warps > 0 ? active_warps : WARPS;
        a.visit_counter = d_counter;
        sgd_update_tile_kernel<<<grid, WARPS * 32, smem>>>(a);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_per_epoch) *ms_per_epoch = epochs > 0 ? ms / epochs : 0.0;
    CK(cudaMemcpy(P, d_P, (size_t)n_users * K * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(Q, d_Q, (size_t)n_items * K * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_recs); cudaFree(d_off); cudaFree(d_lo); cudaFree(d_nu); cudaFree(d_order);
    cudaFree(d_P); cudaFree(d_Q); cudaFree(d_w); cudaFree(d_counter);
    return 0;
}
