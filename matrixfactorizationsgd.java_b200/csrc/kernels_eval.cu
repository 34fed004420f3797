// kernels_eval.cu -- subsystem (3) held-out RMSE, plus factor init and the synthetic-data twin.
#include "kernels.cuh"

namespace mfsgd {

namespace {

constexpr int RMSE_THREADS = 256;
constexpr int RMSE_MAX_CTAS = 148 * 8;

// sum (r - p_u.q_i)^2: baseline/java/MatrixFactorizationSGD.java:169 (rmse). Dot in binary32 (same
// sub-warp tree as the update kernel), squared error accumulated in binary64 per sub-warp, then a
// warp shuffle reduction, a shared-memory block reduction and one partial per CTA (fixed order ->
// the result is reproducible for a given grid).
template <int LANES, int VEC, bool FULL, bool PH>
__global__ void __launch_bounds__(RMSE_THREADS) rmse_sse_kernel(const Rec* __restrict__ recs, int64_t n,
                                                                const float* __restrict__ P,
                                                                const float* __restrict__ Q, const float* __restrict__ BU,
                                                                const float* __restrict__ BI, int k, int u_base,
                                                                int i_base, double* __restrict__ partial) {
    constexpr int GPW = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (LANES - 1);
    const int grp = lane / LANES;
    const int chunks = k >> 2;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_tiles = (n + 31) >> 5;
    const int32_t* __restrict__ words = reinterpret_cast<const int32_t*>(recs);
    const uint64_t pol = l2_policy_evict_first();
    double acc = 0.0;
    for (int64_t tile = warp; tile < n_tiles; tile += n_warps) {
        const int64_t base = tile * 32;
        const int cnt = (n - base) < 32 ? (int)(n - base) : 32;
        int32_t ru = 0, ri = 0, rr = 0;
        if (lane < cnt) {
            ru = ld_stream_i32(words + 3 * (base + lane), pol) & REC_USER_MASK;
            ri = ld_stream_i32(words + 3 * (base + lane) + 1, pol);
            rr = ld_stream_i32(words + 3 * (base + lane) + 2, pol);
        }
        const int steps = (cnt + GPW - 1) / GPW;
#pragma unroll 4
        for (int t = 0; t < steps; t++) {
            const int j = t * GPW + grp;
            const int32_t u = __shfl_sync(0xffffffffu, ru, j);
            const int32_t i = __shfl_sync(0xffffffffu, ri, j);
            const float r = __int_as_float(__shfl_sync(0xffffffffu, rr, j));
            const bool act = j < cnt;
            const char* prow = reinterpret_cast<const char*>(P) + (int64_t)(u - u_base) * k * p_elem_bytes<PH>();   // PH: binary16 rows
            const float* qrow = Q + (int64_t)(i - i_base) * k;
            float s = 0.0f;
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                const int c = gl + v * LANES;
                if (act && (FULL || c < chunks))
                    s = dot4_acc(s, widen4(ld_pchunk<PH>(prow + 4 * c * p_elem_bytes<PH>())), ld_row4(qrow + 4 * c));
            }
            s = group_sum<LANES>(s);
            if (BU != nullptr && act) s = __fadd_rn(__fadd_rn(s, __ldcg(BU + (u - u_base))), __ldcg(BI + (i - i_base)));   // rmseModel :389
            const float e = __fsub_rn(r, s);
            if (act && gl == 0) acc += (double)e * (double)e;
        }
    }
    // warp reduction (fixed xor order), then block reduction
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    __shared__ double warp_sums[RMSE_THREADS / 32];
    if (lane == 0) warp_sums[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < RMSE_THREADS / 32; w++) s += warp_sums[w];
        partial[blockIdx.x] = s;
    }
}

// *accum += sum(partial[0..n)) in a fixed order (single CTA).
__global__ void __launch_bounds__(256) rmse_finalize_kernel(const double* __restrict__ partial, int n,
                                                            double* __restrict__ accum) {
    __shared__ double sm[256];
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += 256) s += partial[j];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w >= 1; w >>= 1) {
        if ((int)threadIdx.x < w) sm[threadIdx.x] += sm[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *accum += sm[0];
}

// MatrixFactorizationSGD.java:53 initFactors; counters are GLOBAL row ids so stripes agree with the whole.
// HALF: the rows are kept as binary16 -- the same values, rounded to nearest even once.
template <bool HALF>
__global__ void __launch_bounds__(256) init_factors_kernel(void* __restrict__ rows, int64_t n_elems, int k,
                                                           int64_t row_lo, uint64_t seed, uint64_t stream_id,
                                                           float scale) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += stride) {
        const uint64_t ctr = (uint64_t)(row_lo * k + e);
        const float v = __fmul_rn(uniform24(seed, stream_id, ctr), scale);
        if (HALF) static_cast<__half*>(rows)[e] = __float2half_rn(v);
        else static_cast<float*>(rows)[e] = v;
    }
}

// binary16 rows <-> binary32 (mfsgd_get_factors / mfsgd_set_factors under MFSGD_STORAGE_F16): widening is exact, narrowing
// rounds to nearest even
__global__ void __launch_bounds__(256) widen_rows_kernel(const __half* __restrict__ in, int64_t n, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) out[e] = __half2float(in[e]);
}
__global__ void __launch_bounds__(256) narrow_rows_kernel(const float* __restrict__ in, int64_t n, __half* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) out[e] = __float2half_rn(in[e]);
}

// ---- synthetic records: MatrixFactorizationSGD.java:186-239, binary64/binary32 ops spelled out ----
constexpr int PLANTED_RANK = 16;
constexpr float PLANTED_AMPLITUDE = 0.8660254f;
constexpr uint64_t ID_MULT = 2654435761ULL;

__device__ __forceinline__ double uniform53(uint64_t seed, uint64_t stream, uint64_t ctr) {
    return __dmul_rn((double)(hash64(seed, stream, ctr) >> 11), 0x1.0p-53);
}
__device__ __forceinline__ int32_t skewed_rank(double x, int32_t count, int log2_alpha, double c) {
    double y = __dadd_rn(c, __dmul_rn(__dsub_rn(1.0, c), x));
    double ca = c;
    for (int s = 0; s < log2_alpha; s++) {
        y = __dmul_rn(y, y);
        ca = __dmul_rn(ca, ca);
    }
    const double t = __ddiv_rn(__dsub_rn(y, ca), __dsub_rn(1.0, ca));
    long long rank = (long long)floor(__dmul_rn((double)count, t));
    if (rank < 0) rank = 0;
    if (rank > count - 1) rank = count - 1;
    return (int32_t)rank;
}
__device__ __forceinline__ int32_t scatter_id(int32_t rank, int32_t count) {
    return (int32_t)((((uint64_t)rank * ID_MULT) + (uint64_t)(count / 2)) % (uint64_t)count);
}
__device__ __forceinline__ float planted_entry(uint64_t seed, uint64_t stream, int32_t row, int f, float amplitude) {
    return __fmul_rn(__fsub_rn(uniform24(seed, stream, (uint64_t)row * PLANTED_RANK + (uint64_t)f), 0.5f), amplitude);
}

__global__ void __launch_bounds__(256) generate_kernel(SynthArgs s, int64_t start, int64_t count,
                                                       int32_t* __restrict__ u, int32_t* __restrict__ i,
                                                       float* __restrict__ r, uint8_t* __restrict__ held) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const uint64_t n = (uint64_t)(start + t);
        const int32_t uu = scatter_id(skewed_rank(uniform53(s.seed, STREAM_USER, n), s.n_users, s.l2au, s.cu), s.n_users);
        const int32_t ii = scatter_id(skewed_rank(uniform53(s.seed, STREAM_ITEM, n), s.n_items, s.l2ai, s.ci), s.n_items);
        float dot = 0.0f;
#pragma unroll
        for (int f = 0; f < PLANTED_RANK; f++)
            dot = __fadd_rn(dot, __fmul_rn(planted_entry(s.seed, STREAM_PSTAR, uu, f, s.amplitude),
                                           planted_entry(s.seed, STREAM_QSTAR, ii, f, s.amplitude)));
        float noise = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; j++) noise = __fadd_rn(noise, uniform24(s.seed, STREAM_NOISE, 4ULL * n + (uint64_t)j));
        noise = __fsub_rn(noise, 2.0f);
        float rating = __fadd_rn(3.5f, dot);
        rating = __fadd_rn(rating, __fmul_rn(s.noise_scale, noise));
        if (rating < 1.0f) rating = 1.0f;
        if (rating > 5.0f) rating = 5.0f;
        u[t] = uu;
        i[t] = ii;
        r[t] = rating;
        if (held != nullptr) held[t] = (hash64(s.seed, STREAM_HELDOUT, n) % 10ULL == 0ULL) ? 1 : 0;
    }
}



inline int grid_for(int64_t n, int threads, int max_ctas) {
    int64_t g = (n + threads - 1) / threads;
    if (g > max_ctas) g = max_ctas;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

int rmse_scratch_doubles() { return RMSE_MAX_CTAS; }

cudaError_t launch_rmse_sse(const Rec* recs, int64_t n, const float* P, bool p_half, const float* Q, const float* BU, const float* BI,
                            int32_t k, int32_t u_base, int32_t i_base, double* scratch, double* sse_accum, int n_sms,
                            cudaStream_t stream, int* launches) {
    if (n <= 0) return cudaSuccess;
    const Geometry g = geometry_for(k);
    int64_t tiles = (n + 31) / 32;
    int grid = n_sms * 8;
    if (grid > RMSE_MAX_CTAS) grid = RMSE_MAX_CTAS;
    if (grid > (tiles + 7) / 8) grid = (int)((tiles + 7) / 8);
    if (grid < 1) grid = 1;
#define CALL(L, V, F)                                                                                                          \
    if (p_half) rmse_sse_kernel<L, V, F, true><<<grid, RMSE_THREADS, 0, stream>>>(recs, n, P, Q, BU, BI, k, u_base, i_base, scratch); \
    else rmse_sse_kernel<L, V, F, false><<<grid, RMSE_THREADS, 0, stream>>>(recs, n, P, Q, BU, BI, k, u_base, i_base, scratch)
    MFSGD_DISPATCH_GEOMETRY(g, CALL);
#undef CALL
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    rmse_finalize_kernel<<<1, 256, 0, stream>>>(scratch, grid, sse_accum);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_init_factors(void* rows, bool half, int64_t n_rows, int32_t k, int64_t row_lo, uint64_t seed, uint64_t stream_id,
                                float scale, cudaStream_t stream, int* launches) {
    const int64_t n = n_rows * k;
    if (n <= 0) return cudaSuccess;
    if (half) init_factors_kernel<true><<<grid_for(n, 256, 148 * 16), 256, 0, stream>>>(rows, n, k, row_lo, seed, stream_id, scale);
    else init_factors_kernel<false><<<grid_for(n, 256, 148 * 16), 256, 0, stream>>>(rows, n, k, row_lo, seed, stream_id, scale);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_widen_rows(const void* half_rows, int64_t n, float* out, cudaStream_t stream, int* launches) {
    if (n <= 0) return cudaSuccess;
    widen_rows_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, stream>>>(static_cast<const __half*>(half_rows), n, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_narrow_rows(const float* rows, int64_t n, void* half_out, cudaStream_t stream, int* launches) {
    if (n <= 0) return cudaSuccess;
    narrow_rows_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, stream>>>(rows, n, static_cast<__half*>(half_out));
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_generate(const SynthArgs& s, int64_t start, int64_t count, int32_t* u, int32_t* i, float* r,
                            uint8_t* held, cudaStream_t stream, int* launches) {
    if (count <= 0) return cudaSuccess;
    generate_kernel<<<grid_for(count, 256, 148 * 16), 256, 0, stream>>>(s, start, count, u, i, r, held);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace mfsgd
