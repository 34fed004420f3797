// kernels_layout.cu -- subsystem (1): on-device bucketing into stratified blocks and the per-epoch
// in-block reshuffle of the 12-byte (u, i, r) records. HBM-bound integer work: coalesced streams,
// shared-memory histograms/cursors, grids sized in multiples of the SM count.
//
// Stand-in counterpart: baseline/java/MatrixFactorizationSGD.java:72 (shuffle) defines the visiting
// order of the sequential path; the deterministic mode reproduces it exactly (device radix sort of the
// same packed keys); Hogwild/DSGD modes only need *a* fresh permutation inside every block.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include "kernels.cuh"

namespace mfsgd {

namespace {

constexpr int SHUFFLE_SMEM_BLOCKS = 4096;   // offsets above this count are read from global memory

inline int grid_for(int64_t n, int threads, int max_ctas) {
    int64_t g = (n + threads - 1) / threads;
    if (g > max_ctas) g = max_ctas;
    if (g < 1) g = 1;
    return (int)g;
}

__global__ void __launch_bounds__(256) count_rows_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ i,
                                                         const uint8_t* __restrict__ held, int64_t n,
                                                         uint32_t* __restrict__ user_cnt, uint32_t* __restrict__ item_cnt,
                                                         int32_t n_users, int32_t n_items, int* __restrict__ bad_flag,
                                                         const float* __restrict__ r, unsigned long long* __restrict__ rating_sum) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    long long fixed = 0;      // sum floor(r * 2^20): exact integers, so the training mean does not depend on the summation order
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        const int32_t uu = u[t], ii = i[t];
        if (uu < 0 || uu >= n_users || ii < 0 || ii >= n_items) {
            *bad_flag = 1;
            continue;
        }
        if (held != nullptr && held[t] != 0) continue;
        atomicAdd(user_cnt + uu, 1u);
        atomicAdd(item_cnt + ii, 1u);
        if (r != nullptr) fixed += __double2ll_rd(__dmul_rn((double)r[t], 1048576.0));
    }
    if (rating_sum != nullptr) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) fixed += __shfl_xor_sync(0xffffffffu, fixed, m);
        if ((threadIdx.x & 31) == 0 && fixed != 0) atomicAdd(rating_sum, (unsigned long long)fixed);
    }
}

struct U32ToU64 {
    __host__ __device__ __forceinline__ uint64_t operator()(const uint32_t& x) const { return (uint64_t)x; }
};

__global__ void balanced_bounds_kernel(const uint64_t* __restrict__ cum, int32_t n_rows, int32_t nblocks,
                                       int32_t* __restrict__ bounds) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nblocks) return;
    if (b == 0) { bounds[0] = 0; return; }
    if (b == nblocks) { bounds[b] = n_rows; return; }
    const uint64_t total = cum[n_rows];
    const uint64_t target = (total / (uint64_t)nblocks) * (uint64_t)b + ((total % (uint64_t)nblocks) * (uint64_t)b) / (uint64_t)nblocks;
    int32_t lo = 0, hi = n_rows;  // first row with cum[row] >= target
    while (lo < hi) {
        const int32_t mid = lo + ((hi - lo) >> 1);
        if (cum[mid] >= target) hi = mid; else lo = mid + 1;
    }
    bounds[b] = lo;
}

__device__ __forceinline__ int upper_block(const int32_t* bounds, int nblocks, int32_t row) {
    int lo = 0, hi = nblocks;  // largest b with bounds[b] <= row (bounds non-decreasing, empty blocks allowed)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (bounds[mid] <= row) lo = mid; else hi = mid;
    }
    return lo;
}

constexpr size_t FILL_OWNER_SMEM_MAX = 48 * 1024;

__global__ void __launch_bounds__(256) fill_owner_kernel(const int32_t* __restrict__ bounds, int32_t nblocks,
                                                         int32_t n_rows, uint16_t* __restrict__ owner) {
    extern __shared__ int32_t sb[];
    const int32_t* __restrict__ tab = bounds;            // more bounds than fit the default 48 KB: search them in global memory
    if ((size_t)(nblocks + 1) * sizeof(int32_t) <= FILL_OWNER_SMEM_MAX) {
        for (int j = threadIdx.x; j <= nblocks; j += blockDim.x) sb[j] = bounds[j];
        __syncthreads();
        tab = sb;
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += stride)
        owner[row] = (uint16_t)upper_block(tab, nblocks, (int32_t)row);
}

__device__ __forceinline__ int record_block(const BucketArgs& b, int64_t t) {
    if (b.held != nullptr) {
        if ((b.held[t] != 0) != (b.want_held != 0)) return -1;
    } else if (b.want_held != 0) {
        return -1;
    }
    const int ou = (int)b.owner_u[b.u[t]];
    if (ou < b.ub_lo || ou >= b.ub_hi) return -1;
    const int row = (ou - b.ub_lo) / b.row_div;
    const int32_t item = b.i[t];
    if (b.hot_index != nullptr) {
        const int32_t hx = b.hot_index[item];
        if (hx >= 0) return b.hot_base + row * b.n_hot + hx;
    }
    return row * b.n_cols + (int)b.owner_i[item] / b.col_div;
}

__device__ __forceinline__ int32_t marked_user(const BucketArgs& b, int32_t u) {
    if (b.heavy_bits != nullptr && ((b.heavy_bits[u >> 5] >> (u & 31)) & 1u)) return u | (int32_t)0x80000000;
    return u;
}

constexpr int BUCKET_THREADS = 256;
constexpr int BUCKET_ITEMS = 8;
constexpr int BUCKET_CHUNK = BUCKET_THREADS * BUCKET_ITEMS;

// per-CTA shared-memory histogram, one global atomic per (CTA, non-empty block)
__global__ void __launch_bounds__(BUCKET_THREADS) block_histogram_kernel(BucketArgs b, int nblk,
                                                                         unsigned long long* __restrict__ block_cnt) {
    extern __shared__ uint32_t hist[];
    for (int j = threadIdx.x; j < nblk; j += blockDim.x) hist[j] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < b.n; t += stride) {
        const int blk = record_block(b, t);
        if (blk >= 0) atomicAdd(hist + blk, 1u);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nblk; j += blockDim.x)
        if (hist[j] != 0) atomicAdd(block_cnt + j, (unsigned long long)hist[j]);
}

// Each CTA takes chunks of BUCKET_CHUNK records: counts them per block in shared memory, reserves one
// contiguous range per non-empty block with a single global atomic, then places its records.
__global__ void __launch_bounds__(BUCKET_THREADS) block_scatter_kernel(BucketArgs b, int nblk,
                                                                       unsigned long long* __restrict__ cursors,
                                                                       Rec* __restrict__ out) {
    extern __shared__ uint32_t sm[];
    uint32_t* cnt = sm;                                                   // [nblk]
    unsigned long long* base = reinterpret_cast<unsigned long long*>(sm + ((nblk + 1) & ~1));  // [nblk]
    const int64_t n_chunks = (b.n + BUCKET_CHUNK - 1) / BUCKET_CHUNK;
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        for (int j = threadIdx.x; j < nblk; j += blockDim.x) cnt[j] = 0;
        __syncthreads();
        int blk[BUCKET_ITEMS];
        uint32_t slot[BUCKET_ITEMS];
#pragma unroll
        for (int it = 0; it < BUCKET_ITEMS; it++) {
            const int64_t t = chunk * BUCKET_CHUNK + it * BUCKET_THREADS + threadIdx.x;
            blk[it] = (t < b.n) ? record_block(b, t) : -1;
            slot[it] = (blk[it] >= 0) ? atomicAdd(cnt + blk[it], 1u) : 0u;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < nblk; j += blockDim.x)
            if (cnt[j] != 0) base[j] = atomicAdd(cursors + j, (unsigned long long)cnt[j]);
        __syncthreads();
#pragma unroll
        for (int it = 0; it < BUCKET_ITEMS; it++) {
            if (blk[it] >= 0) {
                const int64_t t = chunk * BUCKET_CHUNK + it * BUCKET_THREADS + threadIdx.x;
                Rec rec;
                rec.u = marked_user(b, b.u[t]);
                rec.i = b.i[t];
                rec.r = __fsub_rn(b.r[t], b.center);
                out[base[blk[it]] + slot[it]] = rec;
            }
        }
        __syncthreads();
    }
}

// Many-bucket variants (hot-item sets of tens of thousands of (stripe, item) buckets): counters live in global
// memory; with that many buckets the atomics spread out, and a warp first merges its lanes that hit the same bucket.
__global__ void __launch_bounds__(256) block_histogram_global_kernel(BucketArgs b, unsigned long long* __restrict__ block_cnt) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (b.n + 31) & ~(int64_t)31;   // whole warps stay in the loop for the match below
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += stride) {
        const int blk = (t < b.n) ? record_block(b, t) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, blk);
        if (blk >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(block_cnt + blk, (unsigned long long)__popc(peers));
    }
}

__global__ void __launch_bounds__(256) block_scatter_global_kernel(BucketArgs b, unsigned long long* __restrict__ cursors,
                                                                   Rec* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (b.n + 31) & ~(int64_t)31;
    const int lane = threadIdx.x & 31;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += stride) {
        const int blk = (t < b.n) ? record_block(b, t) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, blk);
        const int leader = __ffs(peers) - 1;
        unsigned long long base = 0;
        if (blk >= 0 && lane == leader) base = atomicAdd(cursors + blk, (unsigned long long)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (blk >= 0) {
            Rec rec;
            rec.u = marked_user(b, b.u[t]);
            rec.i = b.i[t];
            rec.r = __fsub_rn(b.r[t], b.center);
            out[base + (unsigned long long)__popc(peers & ((1u << lane) - 1u))] = rec;
        }
    }
}

// ---- in-block reshuffle (the permutation itself lives in common.cuh: the update kernels apply it on the fly) ----
__global__ void __launch_bounds__(256) block_shuffle_kernel(const Rec* __restrict__ in, Rec* __restrict__ out,
                                                            const int64_t* __restrict__ block_off, int nblocks,
                                                            int64_t n, uint64_t seed, uint32_t epoch,
                                                            uint32_t block_id_base) {
    extern __shared__ int64_t soff_smem[];
    const int64_t* __restrict__ soff = block_off;
    if (nblocks <= SHUFFLE_SMEM_BLOCKS) {
        for (int j = threadIdx.x; j <= nblocks; j += blockDim.x) soff_smem[j] = block_off[j];
        __syncthreads();
        soff = soff_smem;
    }
    const int32_t* __restrict__ win = reinterpret_cast<const int32_t*>(in);
    int32_t* __restrict__ wout = reinterpret_cast<int32_t*>(out);
    // The reshuffle runs in the background of the update kernels, whose factors live in L2: stream the
    // records past it (evict-first both ways) so 2 x 12 B x N of one-touch data does not displace them.
    const uint64_t pol = l2_policy_evict_first();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        int lo = 0, hi = nblocks;  // largest b with soff[b] <= j
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (soff[mid] <= j) lo = mid; else hi = mid;
        }
        const int64_t off = soff[lo];
        const uint64_t nb = (uint64_t)(soff[lo + 1] - off);
        uint64_t src = (uint64_t)(j - off);
        if (nb > 1) src = block_perm(src, nb, perm_half_bits(nb), bucket_perm_key(seed, epoch, block_id_base + (uint32_t)lo));
        const int64_t s3 = 3 * (off + (int64_t)src);
        const int32_t a = ld_stream_i32(win + s3, pol), bb = ld_stream_i32(win + s3 + 1, pol), c = ld_stream_i32(win + s3 + 2, pol);
        st_stream_i32(wout + 3 * j, a, pol);
        st_stream_i32(wout + 3 * j + 1, bb, pol);
        st_stream_i32(wout + 3 * j + 2, c, pol);
    }
}

__global__ void __launch_bounds__(256) pack_records_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ i,
                                                           const float* __restrict__ r, int64_t n, float center, Rec* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        Rec rec;
        rec.u = u[t];
        rec.i = i[t];
        rec.r = __fsub_rn(r[t], center);
        out[t] = rec;
    }
}

// AoS -> SoA (records received from other ring members are re-bucketed through the same SoA staging as host input)
__global__ void __launch_bounds__(256) unpack_records_kernel(const Rec* __restrict__ in, int64_t n, int32_t* __restrict__ u,
                                                             int32_t* __restrict__ i, float* __restrict__ r) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        const Rec rec = in[t];
        u[t] = rec.u & REC_USER_MASK;
        i[t] = rec.i;
        r[t] = rec.r;
    }
}

// MatrixFactorizationSGD.java:75-78: packed = (hash64(seed,2,(epoch<<32)|idx) >>> 33) << 32 | idx
__global__ void __launch_bounds__(256) order_keys_kernel(uint64_t* __restrict__ keys, int32_t n, uint64_t seed,
                                                         uint32_t epoch) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        const uint64_t key = hash64(seed, STREAM_SHUFFLE, ((uint64_t)epoch << 32) | (uint64_t)t) >> 33;
        keys[t] = (key << 32) | (uint64_t)t;
    }
}
__global__ void __launch_bounds__(256) order_gather_kernel(const Rec* __restrict__ in, Rec* __restrict__ out,
                                                           const uint64_t* __restrict__ sorted, int32_t n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = in[(uint32_t)(sorted[t] & 0xFFFFFFFFULL)];
}

}  // namespace

cudaError_t launch_count_rows(const int32_t* u, const int32_t* i, const uint8_t* held, int64_t n, uint32_t* user_cnt,
                              uint32_t* item_cnt, int32_t n_users, int32_t n_items, int* bad_flag, const float* r,
                              unsigned long long* rating_sum, cudaStream_t stream, int* launches) {
    if (n <= 0) return cudaSuccess;
    count_rows_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, stream>>>(u, i, held, n, user_cnt, item_cnt, n_users, n_items,
                                                                    bad_flag, rating_sum ? r : nullptr, rating_sum);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t exclusive_cumsum_u32(const uint32_t* cnt, uint64_t* cum, int32_t n, void* temp, size_t* temp_bytes,
                                 cudaStream_t stream, int* launches) {
    // cnt has n + 1 entries (the last one zero), so cum[n] is the total.
    auto in = thrust::make_transform_iterator(cnt, U32ToU64());
    cudaError_t err = cub::DeviceScan::ExclusiveSum(temp, *temp_bytes, in, cum, n + 1, stream);
    if (temp != nullptr && launches) *launches += 2;
    return err;
}

cudaError_t launch_balanced_bounds(const uint64_t* cum, int32_t n_rows, int32_t nblocks, int32_t* bounds,
                                   cudaStream_t stream, int* launches) {
    balanced_bounds_kernel<<<(nblocks + 1 + 127) / 128, 128, 0, stream>>>(cum, n_rows, nblocks, bounds);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_fill_owner(const int32_t* bounds, int32_t nblocks, int32_t n_rows, uint16_t* owner,
                              cudaStream_t stream, int* launches) {
    size_t smem = (size_t)(nblocks + 1) * sizeof(int32_t);
    if (smem > FILL_OWNER_SMEM_MAX) smem = 0;            // the kernel then reads the bounds from global memory
    fill_owner_kernel<<<grid_for(n_rows, 256, 148 * 8), 256, smem, stream>>>(bounds, nblocks, n_rows, owner);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_block_histogram(const BucketArgs& b, unsigned long long* block_cnt, cudaStream_t stream, int* launches) {
    if (b.n <= 0) return cudaSuccess;
    const int nblk = bucket_block_count(b);
    if (nblk > MAX_BUCKETS) return cudaErrorInvalidValue;
    if (nblk > MAX_SMEM_BUCKETS) {
        block_histogram_global_kernel<<<grid_for(b.n, 256 * 8, 148 * 8), 256, 0, stream>>>(b, block_cnt);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    const size_t smem = (size_t)nblk * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(block_histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    block_histogram_kernel<<<grid_for(b.n, BUCKET_THREADS * 16, 148 * 4), BUCKET_THREADS, smem, stream>>>(b, nblk, block_cnt);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_block_scatter(const BucketArgs& b, unsigned long long* cursors, Rec* out, cudaStream_t stream,
                                 int* launches) {
    if (b.n <= 0) return cudaSuccess;
    const int nblk = bucket_block_count(b);
    if (nblk > MAX_BUCKETS) return cudaErrorInvalidValue;
    if (nblk > MAX_SMEM_BUCKETS) {
        block_scatter_global_kernel<<<grid_for(b.n, 256 * 8, 148 * 8), 256, 0, stream>>>(b, cursors, out);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    const size_t smem = (size_t)((nblk + 1) & ~1) * sizeof(uint32_t) + (size_t)nblk * sizeof(unsigned long long);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(block_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    block_scatter_kernel<<<grid_for(b.n, BUCKET_CHUNK, 148 * 4), BUCKET_THREADS, smem, stream>>>(b, nblk, cursors, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_block_shuffle(const Rec* in, Rec* out, const int64_t* block_off, int32_t nblocks, int64_t n,
                                 uint64_t seed, uint32_t epoch, uint32_t block_id_base, cudaStream_t stream, int* launches) {
    if (n <= 0) return cudaSuccess;
    const size_t smem = nblocks <= SHUFFLE_SMEM_BLOCKS ? (size_t)(nblocks + 1) * sizeof(int64_t) : 0;
    block_shuffle_kernel<<<grid_for(n, 256 * 4, 148 * 8), 256, smem, stream>>>(
        in, out, block_off, nblocks, n, seed, epoch, block_id_base);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_pack_records(const int32_t* u, const int32_t* i, const float* r, int64_t n, float center, Rec* out,
                                cudaStream_t stream, int* launches) {
    if (n <= 0) return cudaSuccess;
    pack_records_kernel<<<grid_for(n, 256 * 4, 148 * 8), 256, 0, stream>>>(u, i, r, n, center, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_unpack_records(const Rec* in, int64_t n, int32_t* u, int32_t* i, float* r, cudaStream_t stream, int* launches) {
    if (n <= 0) return cudaSuccess;
    unpack_records_kernel<<<grid_for(n, 256 * 4, 148 * 8), 256, 0, stream>>>(in, n, u, i, r);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t deterministic_order_gather(const Rec* in, Rec* out, int32_t n, uint64_t seed, uint32_t epoch,
                                       uint64_t* keys_a, uint64_t* keys_b, void* temp, size_t* temp_bytes,
                                       cudaStream_t stream, int* launches) {
    if (temp == nullptr) return cub::DeviceRadixSort::SortKeys(nullptr, *temp_bytes, keys_a, keys_b, n, 0, 63, stream);
    if (n <= 0) return cudaSuccess;
    order_keys_kernel<<<(n + 255) / 256, 256, 0, stream>>>(keys_a, n, seed, epoch);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    err = cub::DeviceRadixSort::SortKeys(temp, *temp_bytes, keys_a, keys_b, n, 0, 63, stream);
    if (err != cudaSuccess) return err;
    order_gather_kernel<<<(n + 255) / 256, 256, 0, stream>>>(in, out, keys_b, n);
    if (launches) *launches += 10;  // 2 own kernels + the radix sort's passes (8 digit passes at 8 bits)
    return cudaGetLastError();
}

}  // namespace mfsgd
