/*
 * oracle.cpp -- CPU ORACLE for the factorization path. TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library. libmfsgd.so never links, loads or calls it.
 *
 * PARITY UNPINNED BY THE REFERENCE: /root/reference holds README.md:1-2 only (no source, no tests,
 * no golden vectors). This file restates, function by function, the stand-in
 * baseline/java/MatrixFactorizationSGD.java that BASELINE.json's north_star tells us to commit; each
 * function cites the stand-in's line. The pins that exist are project-made: the hand-computed KAT
 * (tests/golden/kat.json), SplitMix64 known answers, and an independent NumPy restatement
 * (tests/np_restatement.py) whose outputs are committed under tests/golden/.
 *
 * Build: g++ -O2 -ffp-contract=off -fno-fast-math (Java float arithmetic is strict binary32, no FMA).
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <limits>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

enum { ORC_ORDER_SEQ = 0, ORC_ORDER_WARP_TREE = 1, ORC_ORDER_WARP_TREE_FMA = 2, ORC_ORDER_WARP_TREE_FMA_PDELTA = 3 };

static const uint64_t STREAM_P_INIT = 0, STREAM_Q_INIT = 1, STREAM_SHUFFLE = 2, STREAM_USER = 3,
                      STREAM_ITEM = 4, STREAM_NOISE = 5, STREAM_HELDOUT = 6, STREAM_PSTAR = 7,
                      STREAM_QSTAR = 8;

/* MatrixFactorizationSGD.java:39 hash64 */
ORC_API uint64_t orc_hash64(uint64_t seed, uint64_t stream, uint64_t ctr) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (ctr + 1ULL) + 0xD1B54A32D192ED03ULL * stream;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return z;
}

/* MatrixFactorizationSGD.java:48 uniform */
ORC_API float orc_uniform(uint64_t seed, uint64_t stream, uint64_t ctr) {
    return (float)(orc_hash64(seed, stream, ctr) >> 40) * 0x1.0p-24f;
}

/* MatrixFactorizationSGD.java:63 defaultInitScale */
ORC_API float orc_default_init_scale(int k) { return (float)(1.0 / std::sqrt((double)k)); }

/* MatrixFactorizationSGD.java:53 initFactors */
ORC_API void orc_init_factors(float* rows, int64_t n_rows, int k, uint64_t seed, uint64_t stream,
                              float scale) {
    for (int64_t r = 0; r < n_rows; r++)
        for (int f = 0; f < k; f++) {
            uint64_t ctr = (uint64_t)r * (uint64_t)k + (uint64_t)f;
            rows[ctr] = orc_uniform(seed, stream, ctr) * scale;
        }
}

/* MatrixFactorizationSGD.java:72 shuffle */
ORC_API void orc_shuffle(uint64_t seed, int epoch, int n, int32_t* order) {
    std::vector<uint64_t> packed((size_t)n);
    for (int idx = 0; idx < n; idx++) {
        uint64_t key = orc_hash64(seed, STREAM_SHUFFLE, ((uint64_t)(uint32_t)epoch << 32) | (uint64_t)idx) >> 33;
        packed[idx] = (key << 32) | (uint64_t)idx;
    }
    std::sort(packed.begin(), packed.end());
    for (int j = 0; j < n; j++) order[j] = (int32_t)(packed[j] & 0xFFFFFFFFULL);
}

/* The same permutation as orc_shuffle, produced by `threads` host threads (the threaded variant may use every core the box
 * has, also for its per-epoch order): keys are uniform 31-bit values, so the packed words are range-partitioned by their
 * top bits into 16 * threads buckets (counting pass, scatter pass), and the buckets are sorted side by side. */
ORC_API void orc_shuffle_mt(uint64_t seed, int epoch, int n, int32_t* order, int threads) {
    if (threads <= 1 || n < (1 << 16)) { orc_shuffle(seed, epoch, n, order); return; }
    const int B = 16 * threads;                                  /* buckets by key range */
    std::vector<uint64_t> packed((size_t)n), sorted((size_t)n);
    std::vector<std::vector<int64_t>> cnt((size_t)threads, std::vector<int64_t>((size_t)B, 0));
    auto bucket_of = [B](uint64_t w) { return (int)(((w >> 32) * (uint64_t)B) >> 31); };   /* key < 2^31 */
    auto run = [&](auto&& f) {
        std::vector<std::thread> pool;
        for (int w = 0; w < threads; w++) pool.emplace_back(f, w);
        for (auto& th : pool) th.join();
    };
    run([&](int w) {
        const int64_t lo = (int64_t)n * w / threads, hi = (int64_t)n * (w + 1) / threads;
        for (int64_t idx = lo; idx < hi; idx++) {
            uint64_t key = orc_hash64(seed, STREAM_SHUFFLE, ((uint64_t)(uint32_t)epoch << 32) | (uint64_t)idx) >> 33;
            packed[(size_t)idx] = (key << 32) | (uint64_t)idx;
            cnt[(size_t)w][(size_t)bucket_of(packed[(size_t)idx])]++;
        }
    });
    std::vector<int64_t> start((size_t)B + 1, 0);
    for (int b = 0; b < B; b++) {
        int64_t c = 0;
        for (int w = 0; w < threads; w++) { int64_t t = cnt[(size_t)w][(size_t)b]; cnt[(size_t)w][(size_t)b] = start[(size_t)b] + c; c += t; }
        start[(size_t)b + 1] = start[(size_t)b] + c;
    }
    run([&](int w) {
        const int64_t lo = (int64_t)n * w / threads, hi = (int64_t)n * (w + 1) / threads;
        for (int64_t idx = lo; idx < hi; idx++) sorted[(size_t)cnt[(size_t)w][(size_t)bucket_of(packed[(size_t)idx])]++] = packed[(size_t)idx];
    });
    run([&](int w) {
        for (int b = w; b < B; b += threads) std::sort(sorted.begin() + start[(size_t)b], sorted.begin() + start[(size_t)b + 1]);
    });
    run([&](int w) {
        const int64_t lo = (int64_t)n * w / threads, hi = (int64_t)n * (w + 1) / threads;
        for (int64_t j = lo; j < hi; j++) order[j] = (int32_t)(sorted[(size_t)j] & 0xFFFFFFFFULL);
    });
}

/* The dot product of MatrixFactorizationSGD.java:91-94 (f ascending, binary32 accumulate). */
static inline float dot_seq(const float* p, const float* q, int k) {
    float dot = 0.0f;
    for (int f = 0; f < k; f++) dot = dot + p[f] * q[f];
    return dot;
}

/*
 * ORC_ORDER_WARP_TREE: the same products summed in the order the CUDA kernels use:
 * L = min(32, pow2ceil(k/4)) lanes; lane l adds, starting from 0, the 4 products of each chunk
 * c = l, l+L, l+2L, ... (< k/4), ascending; then an xor butterfly over masks L/2 .. 1.
 * Not in the stand-in: it exists so the GPU's deterministic mode can be checked bit for bit.
 */
/* Lanes per rating: min(32, pow2ceil(k/4)) as in the GPU's cold / deterministic / RMSE kernels, or the value set by
 * orc_set_tree_lanes (the GPU's run kernel, kernels_hot.cu, puts 8 lanes on a rating up to k = 128). */
static int g_tree_lanes = 0;
ORC_API void orc_set_tree_lanes(int lanes) { g_tree_lanes = (lanes == 8 || lanes == 16 || lanes == 32) ? lanes : 0; }
static inline int tree_lanes(int chunks) {
    if (g_tree_lanes > 0) return g_tree_lanes;
    int L = 1;
    while (L < chunks && L < 32) L <<= 1;
    return L;
}

static inline float dot_warp_tree(const float* p, const float* q, int k) {
    const int chunks = k / 4, L = tree_lanes(chunks);
    float s[32];
    for (int l = 0; l < L; l++) {
        float acc = 0.0f;
        for (int c = l; c < chunks; c += L)
            for (int j = 0; j < 4; j++) acc = acc + p[4 * c + j] * q[4 * c + j];
        s[l] = acc;
    }
    for (int m = L >> 1; m >= 1; m >>= 1) {
        float t[32];
        for (int l = 0; l < L; l++) t[l] = s[l] + s[l ^ m];
        for (int l = 0; l < L; l++) s[l] = t[l];
    }
    return s[0];
}

/*
 * ORC_ORDER_WARP_TREE_FMA: the arithmetic of the GPU's full-grid kernels (kernels_update.cu FAST):
 * per lane a (lo, hi) pair accumulated with fused multiply-adds over (x, z) and (y, w) of its chunks, lo + hi,
 * then the same butterfly; update p' = fma(b, q, a*p), q' = fma(b, p, a*q) with a = 1 - lr*lambda, b = lr*e.
 * Algebraically the stand-in's rule (MatrixFactorizationSGD.java:95-101), a few ulp apart per update. Not in the
 * stand-in: it exists so the Hogwild/DSGD kernels can be checked bit for bit on conflict-free data.
 */
static inline float dot_warp_tree_fma(const float* p, const float* q, int k) {
    const int chunks = k / 4, L = tree_lanes(chunks);
    float s[32];
    for (int l = 0; l < L; l++) {
        float lo = 0.0f, hi = 0.0f;
        bool first = true;
        for (int c = l; c < chunks; c += L) {
            const float* pp = p + 4 * c;
            const float* qq = q + 4 * c;
            if (first) { lo = pp[0] * qq[0]; hi = pp[1] * qq[1]; first = false; }
            else { lo = std::fmaf(pp[0], qq[0], lo); hi = std::fmaf(pp[1], qq[1], hi); }
            lo = std::fmaf(pp[2], qq[2], lo);
            hi = std::fmaf(pp[3], qq[3], hi);
        }
        s[l] = lo + hi;
    }
    for (int m = L >> 1; m >= 1; m >>= 1) {
        float t[32];
        for (int l = 0; l < L; l++) t[l] = s[l] + s[l ^ m];
        for (int l = 0; l < L; l++) s[l] = t[l];
    }
    return s[0];
}

static inline float dot_ordered(const float* p, const float* q, int k, int order_mode) {
    if (order_mode == ORC_ORDER_WARP_TREE_FMA || order_mode == ORC_ORDER_WARP_TREE_FMA_PDELTA) return dot_warp_tree_fma(p, q, k);
    return order_mode == ORC_ORDER_WARP_TREE ? dot_warp_tree(p, q, k) : dot_seq(p, q, k);
}

/* MatrixFactorizationSGD.java:89 sgdUpdate */
ORC_API float orc_sgd_update(float* p, float* q, int k, float r, float lr, float lambda, int order_mode) {
    float e = r - dot_ordered(p, q, k, order_mode);
    if (order_mode == ORC_ORDER_WARP_TREE_FMA_PDELTA) {
        /* the GPU run kernel with p_u updated in memory by red.global.add (MFSGD_SCATTER_ATOMIC_P): the increment
         * b * q + c * p (c = -(lr * lambda), one fused multiply-add) is added to the row; q_i, in registers, as in FMA mode */
        const float a = 1.0f - lr * lambda, b = lr * e, c = -(lr * lambda);
        for (int f = 0; f < k; f++) {
            float pf = p[f], qf = q[f];
            p[f] = pf + std::fmaf(b, qf, c * pf);
            q[f] = std::fmaf(b, pf, a * qf);
        }
        return e;
    }
    if (order_mode == ORC_ORDER_WARP_TREE_FMA) {
        const float a = 1.0f - lr * lambda, b = lr * e;
        for (int f = 0; f < k; f++) {
            float pf = p[f], qf = q[f];
            p[f] = std::fmaf(b, qf, a * pf);
            q[f] = std::fmaf(b, pf, a * qf);
        }
        return e;
    }
    for (int f = 0; f < k; f++) {
        float pf = p[f], qf = q[f];
        p[f] = pf + lr * (e * qf - lambda * pf);
        q[f] = qf + lr * (e * pf - lambda * qf);
    }
    return e;
}

/* ------------------------------------------------------------------------------------------------
 * Mixed-precision factor storage (MatrixFactorizationSGD.java:403-460: srWord :413, storeF16Sr :420,
 * sgdUpdateMixed :426, factorizeMixed :439): P rows are KEPT as binary16,
 * every operation of the update rule stays binary32. A row is widened exactly on load and narrowed on store with
 * stochastic rounding driven by a counter hash of (seed, epoch, u, i, chunk) -- no state, any visiting order.
 * ------------------------------------------------------------------------------------------------ */
static inline uint32_t f32_bits(float v) { uint32_t b; std::memcpy(&b, &v, 4); return b; }
static inline float bits_f32(uint32_t b) { float v; std::memcpy(&v, &b, 4); return v; }

/* binary16 -> binary32, exact (stand-in Float.float16ToFloat) */
ORC_API float orc_f16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, ex = (h >> 10) & 0x1Fu, man = h & 0x3FFu;
    if (ex == 0) {
        if (man == 0) return bits_f32(sign);
        float v = (float)man * 0x1.0p-24f;                   /* subnormal: man * 2^-24, exact */
        return sign ? -v : v;
    }
    if (ex == 31) return bits_f32(sign | 0x7F800000u | (man << 13));
    return bits_f32(sign | ((ex + 112u) << 23) | (man << 13));
}

/* binary32 -> binary16, round to nearest even (stand-in Float.floatToFloat16; GPU cvt.rn.f16.f32) */
ORC_API uint16_t orc_f32_to_f16_rn(float v) {
    const uint32_t b = f32_bits(v), sign = (b >> 16) & 0x8000u, a = b & 0x7FFFFFFFu;
    if (a >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (a > 0x7F800000u ? 0x200u : 0u));   /* inf / nan */
    if (a >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                                        /* rounds to inf (>= 65520) */
    if (a < 0x33000001u) return (uint16_t)sign;                                                    /* <= 2^-25: rounds to 0 */
    int ex = (int)(a >> 23) - 127;
    uint32_t man = (a & 0x7FFFFFu) | 0x800000u;               /* 24-bit significand */
    int shift;                                                /* bits to drop */
    uint32_t base;
    if (ex < -14) { shift = 13 + (-14 - ex); base = 0; }      /* subnormal result */
    else { shift = 13; base = (uint32_t)(ex + 15) << 10; man &= 0x7FFFFFu; }
    uint32_t q = man >> shift, rem = man & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    return (uint16_t)(sign | (base + q));                     /* a carry out of the mantissa bumps the exponent: intended */
}

/* the 32 random bits of one float4 chunk c of row u at the update (u, i) of `epoch` (lowbias32 finaliser) */
ORC_API uint32_t orc_sr_word(uint64_t seed, uint32_t epoch, uint32_t u, uint32_t i, uint32_t c) {
    uint32_t x = (uint32_t)(seed ^ (seed >> 32)) + u * 0x9E3779B1u + i * 0x85EBCA77u + (epoch * 0x10001u + c) * 0xC2B2AE3Du;
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

/* narrow element j (0..3) of a chunk with stochastic rounding: 8 random bits decide among the 13 dropped mantissa bits */
ORC_API uint16_t orc_store_f16_sr(float v, uint32_t word, int j) {
    const uint32_t rho = (((word >> (8 * j)) & 0xFFu) << 5) | 0x10u;
    return orc_f32_to_f16_rn(bits_f32((f32_bits(v) + rho) & 0xFFFFE000u));
}

/* binary16 a + b (sub != 0: a - b), one rounding to nearest even -- the GPU's HSUB2 / red.global.add.noftz.f16x2 on finite values.
 * Every finite binary16 value is an integer multiple of 2^-24, so the exact result is an integer in that unit: round it to 11
 * significant bits (or to the subnormal grid, which is that unit itself). Engine-side twin, not in the stand-in. */
static inline int64_t f16_scaled(uint16_t h) {
    const int ex = (h >> 10) & 0x1F, man = h & 0x3FF;
    const int64_t mag = ex == 0 ? (int64_t)man : ((int64_t)(man | 0x400) << (ex - 1));
    return (h & 0x8000u) ? -mag : mag;
}
ORC_API uint16_t orc_f16_add_rn(uint16_t a, uint16_t b, int sub) {
    int64_t v = f16_scaled(a) + (sub ? -f16_scaled(b) : f16_scaled(b));
    if (v == 0) return (uint16_t)((a & (sub ? (b ^ 0x8000u) : b)) & 0x8000u);        /* -0 only from (-0) + (-0); x - x = +0 */
    const uint16_t sign = v < 0 ? 0x8000u : 0u;
    uint64_t m = (uint64_t)(v < 0 ? -v : v);
    int top = 63;
    while (!((m >> top) & 1u)) top--;                      /* position of the leading bit: value = m * 2^-24 */
    if (top <= 10) return (uint16_t)(sign | m);            /* subnormal, or the first normal binades: exact (ex = 1 when bit 10 is set) */
    const int shift = top - 10;
    uint64_t q = m >> shift;
    const uint64_t rem = m & ((1ULL << shift) - 1ULL), half = 1ULL << (shift - 1);
    if (rem > half || (rem == half && (q & 1ULL))) q++;
    int ex = shift + 1;
    if (q == 0x800ULL) { q = 0x400ULL; ex++; }
    if (ex >= 31) return (uint16_t)(sign | 0x7C00u);
    return (uint16_t)(sign | ((uint16_t)ex << 10) | (uint16_t)(q & 0x3FFULL));
}

ORC_API void orc_init_factors_f16(uint16_t* rows, int64_t n_rows, int k, uint64_t seed, uint64_t stream, float scale) {
    for (int64_t e = 0; e < n_rows * (int64_t)k; e++) rows[e] = orc_f32_to_f16_rn(orc_uniform(seed, stream, (uint64_t)e) * scale);
}

/* sgdUpdate on a binary16 p_u: widen, the rule in `order_mode`, narrow (sr != 0: stochastic, else round to nearest even).
 * sr == 2: the engine's heavy-user rows (kernels_hot.cu red_pchunk_f16) -- the row in memory moves by the binary16 difference
 * between the narrowed new value and the value the update started from, two binary16 roundings (equal to the plain store
 * whenever that difference is exact, i.e. almost always; small values that more than double are the exception). */
ORC_API float orc_sgd_update_mixed(uint16_t* p16, float* q, int k, float r, float lr, float lambda, int order_mode, uint64_t seed,
                                   uint32_t epoch, int32_t u, int32_t i, int sr) {
    std::vector<float> p((size_t)k);
    for (int f = 0; f < k; f++) p[(size_t)f] = orc_f16_to_f32(p16[f]);
    const float e = orc_sgd_update(p.data(), q, k, r, lr, lambda, order_mode);
    for (int c = 0; c < k / 4; c++) {
        const uint32_t w = orc_sr_word(seed, epoch, (uint32_t)u, (uint32_t)i, (uint32_t)c);
        for (int j = 0; j < 4; j++) {
            const uint16_t nv = sr ? orc_store_f16_sr(p[(size_t)(4 * c + j)], w, j) : orc_f32_to_f16_rn(p[(size_t)(4 * c + j)]);
            p16[4 * c + j] = sr == 2 ? orc_f16_add_rn(p16[4 * c + j], orc_f16_add_rn(nv, p16[4 * c + j], 1), 0) : nv;
        }
    }
    return e;
}

ORC_API int orc_train_mixed(const int32_t* u, const int32_t* i, const float* r, int64_t n, uint16_t* P16, float* Q, int nU, int nI,
                            int k, float lr, float lambda, int epoch_begin, int epoch_end, uint64_t seed, int order_mode, int shuffled,
                            int sr) {
    if (n > 0x7fffffffLL) return -1;
    for (int64_t t = 0; t < n; t++)
        if (u[t] < 0 || u[t] >= nU || i[t] < 0 || i[t] >= nI) return -1;
    std::vector<int32_t> order((size_t)n);
    for (int epoch = epoch_begin; epoch < epoch_end; epoch++) {
        if (shuffled) orc_shuffle(seed, epoch, (int)n, order.data());
        else for (int64_t j = 0; j < n; j++) order[(size_t)j] = (int32_t)j;
        for (int64_t j = 0; j < n; j++) {
            const int32_t t = order[(size_t)j];
            orc_sgd_update_mixed(P16 + (int64_t)u[t] * k, Q + (int64_t)i[t] * k, k, r[t], lr, lambda, order_mode, seed, (uint32_t)epoch,
                                 u[t], i[t], sr);
        }
    }
    return 0;
}

ORC_API void orc_widen_f16(const uint16_t* in, int64_t n, float* out) {
    for (int64_t e = 0; e < n; e++) out[e] = orc_f16_to_f32(in[e]);
}

/* MatrixFactorizationSGD.java:272 globalMean: exact integer sum, hence order-independent */
ORC_API float orc_global_mean(const float* r, int64_t n) {
    int64_t s = 0;
    for (int64_t t = 0; t < n; t++) s += (int64_t)std::floor((double)r[t] * 1048576.0);
    return n == 0 ? 0.0f : (float)((double)s / (double)n / 1048576.0);
}

/* MatrixFactorizationSGD.java:282 sgdUpdateModel on the centred rating rc; bu / bi point at the two bias entries (or are null).
 * The factor part is orc_sgd_update in any order mode; the bias part is the plain rule, one rounding per operation, in every mode. */
ORC_API float orc_sgd_update_model(float* p, float* q, int k, float* bu, float* bi, float rc, float lr, float lambda, int order_mode) {
    if (!bu) return orc_sgd_update(p, q, k, rc, lr, lambda, order_mode);
    const float b0 = *bu, b1 = *bi;
    /* e = rc - ((dot + b_u) + b_i): feed the factor update with the rating reduced by the biases -- algebraically the same, but
     * NOT the same roundings; so restate: compute the dot in the requested order, then the error exactly as the stand-in does */
    float pred = dot_ordered(p, q, k, order_mode) + b0;
    pred = pred + b1;
    const float e = rc - pred;
    *bu = b0 + lr * (e - lambda * b0);
    *bi = b1 + lr * (e - lambda * b1);
    if (order_mode == ORC_ORDER_WARP_TREE_FMA_PDELTA) {
        const float a = 1.0f - lr * lambda, b = lr * e, c = -(lr * lambda);
        for (int f = 0; f < k; f++) {
            float pf = p[f], qf = q[f];
            p[f] = pf + std::fmaf(b, qf, c * pf);
            q[f] = std::fmaf(b, pf, a * qf);
        }
    } else if (order_mode == ORC_ORDER_WARP_TREE_FMA) {
        const float a = 1.0f - lr * lambda, b = lr * e;
        for (int f = 0; f < k; f++) {
            float pf = p[f], qf = q[f];
            p[f] = std::fmaf(b, qf, a * pf);
            q[f] = std::fmaf(b, pf, a * qf);
        }
    } else {
        for (int f = 0; f < k; f++) {
            float pf = p[f], qf = q[f];
            p[f] = pf + lr * (e * qf - lambda * pf);
            q[f] = qf + lr * (e * pf - lambda * qf);
        }
    }
    return e;
}

static int check_triplets(const int32_t* u, const int32_t* i, int64_t n, int nU, int nI) {
    for (int64_t t = 0; t < n; t++)
        if (u[t] < 0 || u[t] >= nU || i[t] < 0 || i[t] >= nI) return -1;
    return 0;
}

/*
 * Epochs [epoch_begin, epoch_end) of the loop in MatrixFactorizationSGD.java:127-133 on existing
 * P, Q. shuffled=0 visits records in array order (used by tests that feed a pre-ordered list).
 * trace_e (nullable): the error of every update, in visiting order.
 */
ORC_API int orc_train(const int32_t* u, const int32_t* i, const float* r, int64_t n, float* P, float* Q,
                      int nU, int nI, int k, float lr, float lambda, int epoch_begin, int epoch_end,
                      uint64_t seed, int order_mode, int shuffled, float* trace_e) {
    if (n > 0x7fffffffLL || check_triplets(u, i, n, nU, nI)) return -1;
    std::vector<int32_t> order((size_t)n);
    int64_t w = 0;
    for (int epoch = epoch_begin; epoch < epoch_end; epoch++) {
        if (shuffled) orc_shuffle(seed, epoch, (int)n, order.data());
        else for (int64_t j = 0; j < n; j++) order[j] = (int32_t)j;
        for (int64_t j = 0; j < n; j++) {
            int32_t t = order[j];
            float e = orc_sgd_update(P + (int64_t)u[t] * k, Q + (int64_t)i[t] * k, k, r[t], lr, lambda, order_mode);
            if (trace_e) trace_e[w++] = e;
        }
    }
    return 0;
}

/* MatrixFactorizationSGD.java:305 factorizeModel's loop on existing P, Q, biases (bu, bi nullable together); rc = centred ratings. */
ORC_API int orc_train_model(const int32_t* u, const int32_t* i, const float* rc, int64_t n, float* P, float* Q, float* bu, float* bi,
                            int nU, int nI, int k, float lr, float lambda, int epoch_begin, int epoch_end, uint64_t seed,
                            int order_mode, int shuffled) {
    if (n > 0x7fffffffLL || check_triplets(u, i, n, nU, nI)) return -1;
    std::vector<int32_t> order((size_t)n);
    for (int epoch = epoch_begin; epoch < epoch_end; epoch++) {
        if (shuffled) orc_shuffle(seed, epoch, (int)n, order.data());
        else for (int64_t j = 0; j < n; j++) order[j] = (int32_t)j;
        for (int64_t j = 0; j < n; j++) {
            int32_t t = order[j];
            orc_sgd_update_model(P + (int64_t)u[t] * k, Q + (int64_t)i[t] * k, k, bu ? bu + u[t] : nullptr, bi ? bi + i[t] : nullptr,
                                 rc[t], lr, lambda, order_mode);
        }
    }
    return 0;
}

/* MatrixFactorizationSGD.java:389 rmseModel on centred ratings */
ORC_API double orc_rmse_model(const float* P, const float* Q, const float* bu, const float* bi, int k, const int32_t* u, const int32_t* i,
                              const float* rc, int64_t n, int order_mode) {
    double sse = 0.0;
    for (int64_t t = 0; t < n; t++) {
        float pred = dot_ordered(P + (int64_t)u[t] * k, Q + (int64_t)i[t] * k, k, order_mode);
        if (bu) { pred = pred + bu[u[t]]; pred = pred + bi[i[t]]; }
        float e = rc[t] - pred;
        sse += (double)e * (double)e;
    }
    return n == 0 ? 0.0 : std::sqrt(sse / (double)n);
}

/* MatrixFactorizationSGD.java:331 learningRate: lr_0 = lr, lr_(e+1) = lr_e * decay, one binary32 multiply per epoch */
ORC_API float orc_learning_rate(float lr, float decay, int epoch) {
    float l = lr;
    for (int e = 0; e < epoch; e++) l = l * decay;
    return l;
}

/* MatrixFactorizationSGD.java:350 factorizeEarlyStop's loop on existing P, Q, biases; rc / vrc = centred training / validation ratings.
 * curve[e] = validation RMSE after epoch e (max_epochs entries). Returns the epochs run, or -1 on bad triplets. */
ORC_API int orc_train_early_stop(const int32_t* u, const int32_t* i, const float* rc, int64_t n, const int32_t* vu, const int32_t* vi,
                                 const float* vrc, int64_t vn, float* P, float* Q, float* bu, float* bi, int nU, int nI, int k, float lr,
                                 float lambda, float lr_decay, int patience, float min_delta, int max_epochs, uint64_t seed,
                                 int order_mode, double* curve) {
    if (n > 0x7fffffffLL || check_triplets(u, i, n, nU, nI) || check_triplets(vu, vi, vn, nU, nI)) return -1;
    std::vector<int32_t> order((size_t)n);
    double best = std::numeric_limits<double>::infinity();
    int strikes = 0, ran = 0;
    float lr_now = lr;
    for (int epoch = 0; epoch < max_epochs; epoch++) {
        orc_shuffle(seed, epoch, (int)n, order.data());
        for (int64_t j = 0; j < n; j++) {
            int32_t t = order[j];
            orc_sgd_update_model(P + (int64_t)u[t] * k, Q + (int64_t)i[t] * k, k, bu ? bu + u[t] : nullptr, bi ? bi + i[t] : nullptr,
                                 rc[t], lr_now, lambda, order_mode);
        }
        lr_now = lr_now * lr_decay;
        ran = epoch + 1;
        const double v = orc_rmse_model(P, Q, bu, bi, k, vu, vi, vrc, vn, order_mode);
        if (curve) curve[epoch] = v;
        if (patience > 0) {
            if (v < best * (1.0 - (double)min_delta)) { best = v; strikes = 0; }
            else if (++strikes >= patience) break;
        }
    }
    return ran;
}

/*
 * One epoch like orc_train, additionally recording for every update (visiting order) the rows
 * before and after it: pre_p, pre_q, post_p, post_q are [n*k]. This is the teacher-forcing tape
 * for the per-update parity test.
 */
ORC_API int orc_train_tape(const int32_t* u, const int32_t* i, const float* r, int64_t n, float* P, float* Q,
                           int nU, int nI, int k, float lr, float lambda, int epoch, uint64_t seed,
                           int order_mode, int32_t* order_out, float* pre_p, float* pre_q,
                           float* post_p, float* post_q, float* err) {
    if (n > 0x7fffffffLL || check_triplets(u, i, n, nU, nI)) return -1;
    orc_shuffle(seed, epoch, (int)n, order_out);
    for (int64_t j = 0; j < n; j++) {
        int32_t t = order_out[j];
        float* p = P + (int64_t)u[t] * k;
        float* q = Q + (int64_t)i[t] * k;
        std::memcpy(pre_p + j * k, p, sizeof(float) * k);
        std::memcpy(pre_q + j * k, q, sizeof(float) * k);
        err[j] = orc_sgd_update(p, q, k, r[t], lr, lambda, order_mode);
        std::memcpy(post_p + j * k, p, sizeof(float) * k);
        std::memcpy(post_q + j * k, q, sizeof(float) * k);
    }
    return 0;
}

/* MatrixFactorizationSGD.java:109 factorize */
ORC_API int orc_factorize(const int32_t* u, const int32_t* i, const float* r, int64_t n, int nU, int nI, int k,
                          float lr, float lambda, int epochs, uint64_t seed, int order_mode,
                          float* P_out, float* Q_out) {
    if (k <= 0 || nU <= 0 || nI <= 0 || epochs < 0) return -1;
    float scale = orc_default_init_scale(k);
    orc_init_factors(P_out, nU, k, seed, STREAM_P_INIT, scale);
    orc_init_factors(Q_out, nI, k, seed, STREAM_Q_INIT, scale);
    return orc_train(u, i, r, n, P_out, Q_out, nU, nI, k, lr, lambda, 0, epochs, seed, order_mode, 1, nullptr);
}

/*
 * MatrixFactorizationSGD.java:140 factorizeThreaded -- Hogwild, T threads, thread w takes positions
 * w, w+T, ... of the epoch order; join per epoch. Races on P and Q are intended. shuffled=0 skips
 * the per-epoch sort (for bounded timing samples of very large inputs). Returns seconds spent in
 * the update loops (sort excluded) through *seconds.
 */
ORC_API int orc_train_hogwild(const int32_t* u, const int32_t* i, const float* r, int64_t n, float* P, float* Q,
                              int nU, int nI, int k, float lr, float lambda, int epoch_begin, int epoch_end,
                              uint64_t seed, int threads, int shuffled, double* seconds) {
    if (n > 0x7fffffffLL || threads < 1 || check_triplets(u, i, n, nU, nI)) return -1;
    std::vector<int32_t> order((size_t)n);
    double total = 0.0;
    for (int epoch = epoch_begin; epoch < epoch_end; epoch++) {
        if (shuffled) orc_shuffle_mt(seed, epoch, (int)n, order.data(), threads);
        else for (int64_t j = 0; j < n; j++) order[j] = (int32_t)j;
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        const int32_t* ord = order.data();
        for (int w = 0; w < threads; w++)
            pool.emplace_back([=]() {
                for (int64_t j = w; j < n; j += threads) {
                    int32_t t = ord[j];
                    orc_sgd_update(P + (int64_t)u[t] * k, Q + (int64_t)i[t] * k, k, r[t], lr, lambda, ORC_ORDER_SEQ);
                }
            });
        for (auto& th : pool) th.join();
        total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    if (seconds) *seconds = total;
    return 0;
}

/* MatrixFactorizationSGD.java:169 rmse */
ORC_API double orc_rmse(const float* P, const float* Q, int k, const int32_t* u, const int32_t* i,
                        const float* r, int64_t n, int order_mode) {
    double sse = 0.0;
    for (int64_t t = 0; t < n; t++) {
        float e = r[t] - dot_ordered(P + (int64_t)u[t] * k, Q + (int64_t)i[t] * k, k, order_mode);
        sse += (double)e * (double)e;
    }
    return n == 0 ? 0.0 : std::sqrt(sse / (double)n);
}

/* ---- synthetic power-law ratings: MatrixFactorizationSGD.java:186-239 ---- */

static const int PLANTED_RANK = 16;
static const float PLANTED_AMPLITUDE = 0.8660254f;   /* default planted amplitude (noise-dominant sets) */
static const float NOISE_SCALE = 0.5f;               /* default noise scale */
static const uint64_t ID_MULT = 2654435761ULL;

/* MatrixFactorizationSGD.java:191 uniform53 */
static inline double uniform53(uint64_t seed, uint64_t stream, uint64_t ctr) {
    return (double)(orc_hash64(seed, stream, ctr) >> 11) * 0x1.0p-53;
}

/* MatrixFactorizationSGD.java:199 skewedRank */
ORC_API int32_t orc_skewed_rank(double x, int32_t count, int log2_alpha, double c) {
    double y = c + (1.0 - c) * x;
    double ca = c;
    for (int s = 0; s < log2_alpha; s++) { y = y * y; ca = ca * ca; }
    double t = (y - ca) / (1.0 - ca);
    int64_t rank = (int64_t)std::floor((double)count * t);
    if (rank < 0) rank = 0;
    if (rank > count - 1) rank = count - 1;
    return (int32_t)rank;
}

/* MatrixFactorizationSGD.java:211 scatterId */
ORC_API int32_t orc_scatter_id(int32_t rank, int32_t count) {
    return (int32_t)((((uint64_t)rank * ID_MULT) + (uint64_t)(count / 2)) % (uint64_t)count);
}

/* MatrixFactorizationSGD.java:215 plantedEntry */
static inline float planted_entry(uint64_t seed, uint64_t stream, int32_t row, int f, float amplitude) {
    return (orc_uniform(seed, stream, (uint64_t)row * PLANTED_RANK + (uint64_t)f) - 0.5f) * amplitude;
}

/* MatrixFactorizationSGD.java:220 syntheticRecord */
static inline bool synthetic_record(uint64_t seed, uint64_t n, int nU, int nI, int l2au, double cu, int l2ai,
                                    double ci, float amplitude, float noise_scale, int32_t* u, int32_t* i, float* r) {
    int32_t uu = orc_scatter_id(orc_skewed_rank(uniform53(seed, STREAM_USER, n), nU, l2au, cu), nU);
    int32_t ii = orc_scatter_id(orc_skewed_rank(uniform53(seed, STREAM_ITEM, n), nI, l2ai, ci), nI);
    float dot = 0.0f;
    for (int f = 0; f < PLANTED_RANK; f++)
        dot = dot + planted_entry(seed, STREAM_PSTAR, uu, f, amplitude) * planted_entry(seed, STREAM_QSTAR, ii, f, amplitude);
    float noise = 0.0f;
    for (int j = 0; j < 4; j++) noise = noise + orc_uniform(seed, STREAM_NOISE, 4ULL * n + (uint64_t)j);
    noise = noise - 2.0f;
    float rating = 3.5f + dot;
    rating = rating + noise_scale * noise;
    if (rating < 1.0f) rating = 1.0f;
    if (rating > 5.0f) rating = 5.0f;
    *u = uu; *i = ii; *r = rating;
    return orc_hash64(seed, STREAM_HELDOUT, n) % 10ULL == 0ULL;
}

/*
 * Records [start, start+count) of the synthetic set, in record order; held[t] = 1 when record
 * start+t belongs to the held-out tenth. Split across `threads` host threads (pure function of n).
 */
ORC_API void orc_generate2(uint64_t seed, int64_t start, int64_t count, int nU, int nI, int l2au, double cu,
                           int l2ai, double ci, float amplitude, float noise_scale, int32_t* u, int32_t* i, float* r,
                           uint8_t* held, int threads) {
    if (threads < 1) threads = 1;
    if (!(amplitude > 0.0f)) amplitude = PLANTED_AMPLITUDE;      /* 0 = the default (noise-dominant) set */
    if (!(noise_scale > 0.0f)) noise_scale = NOISE_SCALE;
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; w++)
        pool.emplace_back([=]() {
            int64_t lo = count * w / threads, hi = count * (w + 1) / threads;
            for (int64_t t = lo; t < hi; t++)
                held[t] = synthetic_record(seed, (uint64_t)(start + t), nU, nI, l2au, cu, l2ai, ci, amplitude, noise_scale,
                                           u + t, i + t, r + t) ? 1 : 0;
        });
    for (auto& th : pool) th.join();
}

ORC_API void orc_generate(uint64_t seed, int64_t start, int64_t count, int nU, int nI, int l2au, double cu,
                          int l2ai, double ci, int32_t* u, int32_t* i, float* r, uint8_t* held, int threads) {
    orc_generate2(seed, start, count, nU, nI, l2au, cu, l2ai, ci, PLANTED_AMPLITUDE, NOISE_SCALE, u, i, r, held, threads);
}

ORC_API int orc_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

/* ---- twins of the GPU engine's run path (no counterpart in the stand-in) ------------------------------------------
 * The stand-in applies every rating sequentially (MatrixFactorizationSGD.java:127-133). The GPU's run kernel
 * (csrc/kernels_hot.cu) applies RUNS of one item's ratings sequentially with q_i in registers and merges the runs of an
 * item that share a launch by a weighted sum of their net changes. The functions below restate exactly that, on the
 * CPU, from the same plan (mfsgd_plan_runs) -- so that the averaged-merge branch has an exact check, and so that the
 * convergence of the merge rule can be studied without a GPU (tools/run_sim.py).
 */

/* csrc/common.cuh feistel_round / block_perm / perm_half_bits / bucket_perm_key, restated. */
static inline uint32_t feistel_round(uint32_t x, uint32_t key) {
    uint32_t h = (x + key) * 0x9E3779B1u;
    h ^= h >> 15;
    h *= 0x85EBCA77u;
    h ^= h >> 13;
    return h;
}

ORC_API uint64_t orc_bucket_perm_key(uint64_t seed, uint32_t epoch, uint32_t bucket_id) {
    return orc_hash64(seed, 9 /* STREAM_BLOCK_SHUFFLE */, ((uint64_t)epoch << 32) | (uint64_t)bucket_id);
}

/* position x of a bucket of n records reads record orc_block_perm(x, n, key) of that bucket (n > 1) */
ORC_API uint64_t orc_block_perm(uint64_t x, uint64_t n, uint64_t key) {
    const uint64_t tiles = n >> 5, full = tiles << 5;
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    if (x >= full) {
        const uint32_t rem = (uint32_t)(n - full);
        return full + (uint64_t)(((uint32_t)(x - full) + (k1 >> 8) % rem) % rem);
    }
    uint64_t t = x >> 5;
    if (tiles > 1) {
        int bits = 0;
        while (bits < 63 && (1ULL << bits) < tiles) bits++;
        if (bits < 2) bits = 2;
        const int hb = (bits + 1) >> 1;
        const uint32_t mask = (hb >= 32) ? 0xffffffffu : ((1u << hb) - 1u);
        do {
            uint32_t l = (uint32_t)(t >> hb) & mask, r = (uint32_t)t & mask;
            for (int round = 0; round < 4; round++) {
                const uint32_t f = feistel_round(r, (round & 1) ? (k1 + round) : (k0 + round)) & mask;
                const uint32_t nl = r;
                r = l ^ f;
                l = nl;
            }
            t = ((uint64_t)l << hb) | (uint64_t)r;
        } while (t >= tiles);
    }
    const uint32_t h = feistel_round((uint32_t)t ^ k1, k0);
    const uint32_t lane = ((((uint32_t)x & 31u) ^ ((h >> 16) & 31u)) * ((h & 31u) | 1u) + ((h >> 8) & 31u)) & 31u;
    return (t << 5) | (uint64_t)lane;
}

/* out[j] = record index (relative to the bucket) read at position j, j in [0, n) */
ORC_API void orc_block_perm_fill(uint64_t n, uint64_t seed, uint32_t epoch, uint32_t bucket_id, int64_t* out) {
    const uint64_t key = orc_bucket_perm_key(seed, epoch, bucket_id);
    for (uint64_t j = 0; j < n; j++) out[j] = n > 1 ? (int64_t)orc_block_perm(j, n, key) : (int64_t)j;
}

/*
 * One launch of the run kernel. Units [0, n_units) in claim order (mfsgd_plan_runs: longest first). `resident`
 * sub-warps walk runs side by side, one rating per tick each; sub-warps come in groups of `gpw` (the runs one warp
 * walks side by side): a group claims gpw units at a time when all its runs are done. A run loads q_i when it starts,
 * applies its ratings strictly in order (p_u read and written in place, q_i private), then merges:
 * weight == 1 (and !always_add): Q[i] = q; else Q[i] += (q - q_at_start) * weight  (weight from the planner:
 * min(1, merge_boost / runs of the slice), csrc/run_plan.hpp merge_weight).
 * rec_u / rec_r: the member's record array in layout order (SoA); run position p reads record
 * bstart + perm(p - bstart) when virt != 0, else record p.  P row = u - u_base, Q row = item - i_base.
 */
ORC_API int orc_train_runs_launch(const int32_t* rec_u, const float* rec_r, int64_t n_recs,
                                  const int64_t* unit_start, const int32_t* unit_count, const int32_t* unit_item,
                                  const float* unit_weight, const int64_t* unit_bstart, const int32_t* unit_bn,
                                  const uint32_t* unit_bid, int64_t n_units, int virt, uint64_t seed, uint32_t epoch,
                                  float* P, int32_t u_base, float* Q, int32_t i_base, int k, float lr, float lambda,
                                  int order_mode, int resident, int gpw, int always_add, float* BU, float* BI) {
    if (resident < 1 || gpw < 1 || resident % gpw) return -1;
    struct Slot { int64_t unit; int step; std::vector<float> q, q0; uint64_t key; float b = 0.f, b0 = 0.f; };   /* b: the run's private b_i */
    std::vector<Slot> slots((size_t)resident);
    for (auto& s : slots) { s.unit = -1; s.step = 0; s.q.resize((size_t)k); s.q0.resize((size_t)k); s.key = 0; }
    int64_t next = 0, done = 0;
    const int groups = resident / gpw;
    auto claim = [&](int g) {
        for (int j = 0; j < gpw; j++) {
            Slot& s = slots[(size_t)g * gpw + j];
            if (next < n_units) {
                s.unit = next++;
                s.step = 0;
                const float* qr = Q + (int64_t)(unit_item[s.unit] - i_base) * k;
                std::memcpy(s.q.data(), qr, sizeof(float) * k);
                std::memcpy(s.q0.data(), qr, sizeof(float) * k);
                if (BI) s.b = s.b0 = BI[unit_item[s.unit] - i_base];
                s.key = (virt && unit_bn[s.unit] > 1) ? orc_bucket_perm_key(seed, epoch, unit_bid[s.unit]) : 0;
            } else s.unit = -1;
        }
    };
    for (int g = 0; g < groups; g++) claim(g);
    while (done < n_units) {
        for (auto& s : slots) {                            /* one rating per active run */
            if (s.unit < 0 || s.step >= unit_count[s.unit]) continue;
            const int64_t pos = unit_start[s.unit] + s.step;
            int64_t idx = pos;
            if (virt && unit_bn[s.unit] > 1)
                idx = unit_bstart[s.unit] + (int64_t)orc_block_perm((uint64_t)(pos - unit_bstart[s.unit]), (uint64_t)unit_bn[s.unit], s.key);
            if (idx < 0 || idx >= n_recs) return -2;
            /* bit 31 of u marks a heavy user (csrc/common.cuh REC_USER_MASK): the kernel adds that row's increment in memory,
             * which in the FMA arrangement rounds differently from storing the new value (ORC_ORDER_WARP_TREE_FMA_PDELTA) */
            const int32_t uid = rec_u[idx] & 0x7fffffff;
            const int mode = (rec_u[idx] < 0 && order_mode == ORC_ORDER_WARP_TREE_FMA) ? ORC_ORDER_WARP_TREE_FMA_PDELTA : order_mode;
            orc_sgd_update_model(P + (int64_t)(uid - u_base) * k, s.q.data(), k, BU ? BU + (uid - u_base) : nullptr, BI ? &s.b : nullptr,
                                 rec_r[idx], lr, lambda, mode);
            s.step++;
            static const char* senv = getenv("ORC_SYNC_EVERY");          /* PROTOTYPE: Hogwild-style exchange inside a run */
            if (senv && unit_weight[s.unit] < 1.0f && s.step < unit_count[s.unit] && s.step % atoi(senv) == 0) {
                static const float mstar = getenv("ORC_SYNC_MSTAR") ? (float)atof(getenv("ORC_SYNC_MSTAR")) : 8.0f;
                const float m = std::nearbyintf(1.25f / unit_weight[s.unit]);
                const float w = std::min(1.0f, mstar / m);
                float* qr = Q + (int64_t)(unit_item[s.unit] - i_base) * k;
                for (int f = 0; f < k; f++) { qr[f] = qr[f] + (s.q[f] - s.q0[f]) * w; s.q[f] = qr[f]; s.q0[f] = qr[f]; }
            }
        }
        for (int g = 0; g < groups; g++) {                 /* a warp merges its runs when the longest is done, then claims again */
            bool all_done = true, any = false;
            for (int j = 0; j < gpw; j++) {
                const Slot& s = slots[(size_t)g * gpw + j];
                if (s.unit < 0) continue;
                any = true;
                if (s.step < unit_count[s.unit]) all_done = false;
            }
            if (!any || !all_done) continue;
            for (int j = 0; j < gpw; j++) {
                Slot& s = slots[(size_t)g * gpw + j];
                if (s.unit < 0) continue;
                float* qr = Q + (int64_t)(unit_item[s.unit] - i_base) * k;
                const float planned = unit_weight[s.unit];
                if (BI) {        /* the run's b_i merges exactly like its q_i */
                    float* br = BI + (unit_item[s.unit] - i_base);
                    if (planned == 1.0f && !always_add) *br = s.b;
                    else *br = *br + (s.b - s.b0) * planned;
                }
                if (planned == 1.0f && !always_add) {
                    std::memcpy(qr, s.q.data(), sizeof(float) * k);
                } else {
                    const float w = planned;               /* the planner's weight: min(1, merge_boost / runs of the slice) */
                    static const char* senv2 = getenv("ORC_SYNC_EVERY");
                    if (senv2) {
                        static const float mstar = getenv("ORC_SYNC_MSTAR") ? (float)atof(getenv("ORC_SYNC_MSTAR")) : 8.0f;
                        const float m = std::nearbyintf(1.25f / w);
                        const float ws = std::min(1.0f, mstar / m);
                        for (int f = 0; f < k; f++) qr[f] = qr[f] + (s.q[f] - s.q0[f]) * ws;
                        done++;
                        continue;
                    }
                    static const char* kenv = getenv("ORC_MERGE_KAPPA");      /* PROTOTYPE: direction-split merge */
                    if (kenv) {
                        const float kappa = (float)atof(kenv);
                        const float m = std::nearbyintf(1.25f / w);
                        double n0 = 0, a = 0;
                        for (int f = 0; f < k; f++) { n0 += (double)s.q0[f] * s.q0[f]; a += (double)(s.q[f] - s.q0[f]) * s.q0[f]; }
                        const float wperp = std::min(1.0f, kappa / m);
                        for (int f = 0; f < k; f++) {
                            const float par = n0 > 0 ? (float)(a / n0) * s.q0[f] : 0.0f;
                            const float d = s.q[f] - s.q0[f];
                            qr[f] = qr[f] + par * w + (d - par) * wperp;
                        }
                        done++;
                        continue;
                    }
                    for (int f = 0; f < k; f++) qr[f] = qr[f] + (s.q[f] - s.q0[f]) * w;
                }
                done++;
            }
            claim(g);
        }
    }
    return 0;
}
