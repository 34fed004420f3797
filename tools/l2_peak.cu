// l2_peak.cu -- measured ceilings for the update path's access pattern on this GPU (test/bench infrastructure, not
// product code): (1) random 512-B row gather + scatter inside an L2-resident buffer (what one SGD update does to
// p_u: 16 sectors read, 16 written), (2) the same rows read only, (3) a plain streaming copy out of HBM.
// usage: l2_peak [buffer MB = 61] [row bytes = 512]   -> one JSON line
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// one warp per row (32 lanes x float4 = 512 B), DEPTH independent rows in flight per warp
template <int DEPTH, bool WRITE>
__global__ void __launch_bounds__(256) row_gather_scatter(float4* __restrict__ buf, uint32_t n_rows, int iters, uint64_t seed) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t state = (uint32_t)mix(seed + warp * 0x9E3779B97F4A7C15ULL) | 1u;   // per-warp PCG-style stream: a few ALU ops per row
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        float4 v[DEPTH];
        uint32_t row[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; d++) {
            state = state * 747796405u + 2891336453u;
            const uint32_t x = ((state >> ((state >> 28) + 4u)) ^ state) * 277803737u;
            row[d] = __umulhi((x >> 22) ^ x, n_rows);                            // uniform in [0, n_rows)
            v[d] = __ldcg(buf + (size_t)row[d] * 32 + lane);
        }
#pragma unroll
        for (int d = 0; d < DEPTH; d++) {
            if (WRITE) {
                v[d].x += 1.0f;
                __stcg(buf + (size_t)row[d] * 32 + lane, v[d]);
            } else {
                acc += v[d].x + v[d].w;
            }
        }
    }
    if (!WRITE && acc == 123.456f) buf[0].x = acc;
}

__global__ void __launch_bounds__(256) stream_copy(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

int main(int argc, char** argv) {
    const double mb = argc > 1 ? atof(argv[1]) : 61.0;
    const uint32_t n_rows = (uint32_t)(mb * 1e6 / 512.0);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    float4* buf = nullptr;
    CK(cudaMalloc(&buf, (size_t)n_rows * 512));
    CK(cudaMemset(buf, 0, (size_t)n_rows * 512));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int grid = prop.multiProcessorCount * 8;
    const int iters = 256;
    const double rows = (double)grid * 8 * iters * 4;
    double best_rw = 0, best_ro = 0;
    for (int rep = 0; rep < 5; rep++) {
        float ms = 0;
        CK(cudaEventRecord(e0));
        row_gather_scatter<4, true><<<grid, 256>>>(buf, n_rows, iters, 1234 + rep);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && rows * 1024.0 / (ms * 1e-3) / 1e9 > best_rw) best_rw = rows * 1024.0 / (ms * 1e-3) / 1e9;
        CK(cudaEventRecord(e0));
        row_gather_scatter<4, false><<<grid, 256>>>(buf, n_rows, iters, 99 + rep);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && rows * 512.0 / (ms * 1e-3) / 1e9 > best_ro) best_ro = rows * 512.0 / (ms * 1e-3) / 1e9;
    }
    // streaming copy, 2 GB in + 2 GB out (HBM)
    const size_t n4 = (size_t)1 << 27;
    float4 *a = nullptr, *b = nullptr;
    CK(cudaMalloc(&a, n4 * 16));
    CK(cudaMalloc(&b, n4 * 16));
    CK(cudaMemset(a, 1, n4 * 16));
    double best_copy = 0;
    for (int rep = 0; rep < 4; rep++) {
        float ms = 0;
        CK(cudaEventRecord(e0));
        stream_copy<<<prop.multiProcessorCount * 16, 256>>>(a, b, n4);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && 2.0 * n4 * 16 / (ms * 1e-3) / 1e9 > best_copy) best_copy = 2.0 * n4 * 16 / (ms * 1e-3) / 1e9;
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"buffer_mb\": %.1f, \"l2_mb\": %.1f, \"row_bytes\": 512, "
           "\"row_gather_scatter_gbs\": %.1f, \"row_gather_scatter_rows_per_s\": %.4g, \"row_gather_only_gbs\": %.1f, "
           "\"hbm_stream_copy_gbs\": %.1f, \"note\": \"gather+scatter counts 1024 B per row (512 read + 512 written), "
           "random rows of an L2-resident buffer, one warp per row, 4 rows in flight per warp, 64 warps per SM\"}\n",
           prop.name, mb, prop.l2CacheSize / 1e6, best_rw, best_rw * 1e9 / 1024.0, best_ro, best_copy);
    return 0;
}
