// kernels_diag.cu -- measured ceilings of the update path's access pattern on the device this process runs on
// (mfsgd_measure_ceilings, include/mfsgd.h). Diagnostic entry point: bench.py calls it inside its own process, on the
// box it reports from, so that roofline.l2_bound is measured beside the number it explains instead of being a
// committed constant (round-1 review). Not on any training path.
//   (1) random 512-B row gather + scatter inside an L2-resident buffer -- what one SGD update of the run kernel does to
//       p_u: 16 sectors read, 16 written, no arithmetic to speak of;
//   (2) the same rows read only;
//   (3) a plain streaming copy out of HBM (read + write bytes).
#include <cstdint>

#include "../../include/mfsgd.h"
#include "common.cuh"

namespace mfsgd {
int set_error(int code, const char* fmt, ...);
}

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// one warp per row (32 lanes x float4 = 512 B), DEPTH independent rows in flight per warp
template <int DEPTH, bool WRITE>
__global__ void __launch_bounds__(256) ceiling_row_gather_scatter_kernel(float4* __restrict__ buf, uint32_t n_rows, int iters, uint64_t seed) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t state = (uint32_t)mix64(seed + warp * 0x9E3779B97F4A7C15ULL) | 1u;   // per-warp PCG-style stream: a few ALU ops per row
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

__global__ void __launch_bounds__(256) ceiling_stream_copy_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

}  // namespace

#define DCK(call)                                                                                                   \
    do {                                                                                                            \
        cudaError_t e__ = (call);                                                                                   \
        if (e__ != cudaSuccess) {                                                                                   \
            if (buf) cudaFree(buf);                                                                                 \
            if (a) cudaFree(a);                                                                                     \
            if (b) cudaFree(b);                                                                                     \
            if (e0) cudaEventDestroy(e0);                                                                           \
            if (e1) cudaEventDestroy(e1);                                                                           \
            return mfsgd::set_error(e__ == cudaErrorMemoryAllocation ? MFSGD_E_OOM : MFSGD_E_CUDA, "%s:%d %s: %s", __FILE__, \
                                    __LINE__, #call, cudaGetErrorString(e__));                                      \
        }                                                                                                           \
    } while (0)

extern "C" int mfsgd_measure_ceilings(int32_t device, double buffer_mb, mfsgd_ceilings* out) {
    if (!out || !(buffer_mb > 0.0) || buffer_mb > 65536.0) return mfsgd::set_error(MFSGD_E_INVALID_ARG, "bad arguments");
    float4 *buf = nullptr, *a = nullptr, *b = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    DCK(cudaSetDevice(device));
    cudaDeviceProp prop;
    DCK(cudaGetDeviceProperties(&prop, device));
    const uint32_t n_rows = (uint32_t)(buffer_mb * 1e6 / 512.0);
    if (n_rows == 0) return mfsgd::set_error(MFSGD_E_INVALID_ARG, "buffer too small");
    DCK(cudaMalloc(&buf, (size_t)n_rows * 512));
    DCK(cudaMemset(buf, 0, (size_t)n_rows * 512));
    DCK(cudaEventCreate(&e0));
    DCK(cudaEventCreate(&e1));
    const int grid = prop.multiProcessorCount * 8;      // 64 warps per SM
    const int iters = 256;
    const double rows = (double)grid * 8 * iters * 4;
    double best_rw = 0, best_ro = 0;
    for (int rep = 0; rep < 5; rep++) {
        float ms = 0;
        DCK(cudaEventRecord(e0));
        ceiling_row_gather_scatter_kernel<4, true><<<grid, 256>>>(buf, n_rows, iters, 1234 + rep);
        DCK(cudaEventRecord(e1));
        DCK(cudaEventSynchronize(e1));
        DCK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && rows * 1024.0 / (ms * 1e-3) / 1e9 > best_rw) best_rw = rows * 1024.0 / (ms * 1e-3) / 1e9;
        DCK(cudaEventRecord(e0));
        ceiling_row_gather_scatter_kernel<4, false><<<grid, 256>>>(buf, n_rows, iters, 99 + rep);
        DCK(cudaEventRecord(e1));
        DCK(cudaEventSynchronize(e1));
        DCK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && rows * 512.0 / (ms * 1e-3) / 1e9 > best_ro) best_ro = rows * 512.0 / (ms * 1e-3) / 1e9;
    }
    const size_t n4 = (size_t)1 << 26;                   // 1 GB in + 1 GB out
    DCK(cudaMalloc(&a, n4 * 16));
    DCK(cudaMalloc(&b, n4 * 16));
    DCK(cudaMemset(a, 1, n4 * 16));
    double best_copy = 0;
    for (int rep = 0; rep < 4; rep++) {
        float ms = 0;
        DCK(cudaEventRecord(e0));
        ceiling_stream_copy_kernel<<<prop.multiProcessorCount * 16, 256>>>(a, b, n4);
        DCK(cudaEventRecord(e1));
        DCK(cudaEventSynchronize(e1));
        DCK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && 2.0 * n4 * 16 / (ms * 1e-3) / 1e9 > best_copy) best_copy = 2.0 * n4 * 16 / (ms * 1e-3) / 1e9;
    }
    cudaFree(buf);
    cudaFree(a);
    cudaFree(b);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    out->row_gather_scatter_gbs = best_rw;
    out->row_gather_only_gbs = best_ro;
    out->hbm_stream_copy_gbs = best_copy;
    out->buffer_mb = buffer_mb;
    out->l2_mb = prop.l2CacheSize / 1e6;
    out->sm_count = prop.multiProcessorCount;
    out->reserved = 0;
    return MFSGD_OK;
}
