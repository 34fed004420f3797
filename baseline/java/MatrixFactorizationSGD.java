/*
 * MatrixFactorizationSGD.java -- REFERENCE STAND-IN (not the upstream source).
 *
 * The mounted reference (/root/reference) holds only README.md:1-2 ("# MatrixFactorizationSGD.java",
 * a concurrent-programming coursework at UFRN) and no Java source. BASELINE.json's north_star
 * instructs: "If no Java SGD source is present, commit a plain sequential Java SGD with the
 * identical update rule as the reference stand-in". This file is that stand-in. The class name
 * comes from README.md:1. Everything the README leaves open is pinned in SURVEY.md section 8
 * (rows a1-a6, 8d) and restated in DESIGN.md section 2.
 *
 * Zero dependencies, Java 17 source level. It cannot be compiled in this project's image (no JDK);
 * its executable proxy is oracle/oracle.cpp, which restates it function by function and cites the
 * line numbers of this file.
 *
 * Arithmetic contract: every float operation below is a separate IEEE-754 binary32 operation,
 * rounded to nearest-even, in exactly the order written (Java never contracts a*b+c into an FMA
 * unless Math.fma is called, and since JDK 17 all float arithmetic is strict).
 */
import java.util.Arrays;

public final class MatrixFactorizationSGD {

    /** Output of factorize: row-major P (nUsers x k) and Q (nItems x k). */
    public static final class Factors {
        public final float[] P;
        public final float[] Q;
        public final int nUsers, nItems, k;
        Factors(float[] P, float[] Q, int nUsers, int nItems, int k) {
            this.P = P; this.Q = Q; this.nUsers = nUsers; this.nItems = nItems; this.k = k;
        }
    }

    /* hash streams (SURVEY.md 8a row a6) */
    public static final long STREAM_P_INIT = 0, STREAM_Q_INIT = 1, STREAM_SHUFFLE = 2,
                             STREAM_USER = 3, STREAM_ITEM = 4, STREAM_NOISE = 5,
                             STREAM_HELDOUT = 6, STREAM_PSTAR = 7, STREAM_QSTAR = 8;

    /** a6: counter hash, SplitMix64 finaliser over (seed, stream, ctr); arithmetic mod 2^64. */
    public static long hash64(long seed, long stream, long ctr) {
        long z = seed + 0x9E3779B97F4A7C15L * (ctr + 1L) + 0xD1B54A32D192ED03L * stream;
        z = (z ^ (z >>> 30)) * 0xBF58476D1CE4E5B9L;
        z = (z ^ (z >>> 27)) * 0x94D049BB133111EBL;
        z = z ^ (z >>> 31);
        return z;
    }

    /** a6: uniform in [0,1) with 24 random bits; exact in binary32. */
    public static float uniform(long seed, long stream, long ctr) {
        return (float) (hash64(seed, stream, ctr) >>> 40) * 0x1.0p-24f;
    }

    /** a3: rows[r*k+f] = uniform(seed, stream, r*k+f) * scale. */
    public static void initFactors(float[] rows, int nRows, int k, long seed, long stream, float scale) {
        for (int r = 0; r < nRows; r++) {
            for (int f = 0; f < k; f++) {
                long ctr = (long) r * k + f;
                rows[(int) ctr] = uniform(seed, stream, ctr) * scale;
            }
        }
    }

    /** Default init scale 1/sqrt(k), rounded once to binary32 (sqrt in double, divide in double). */
    public static float defaultInitScale(int k) {
        return (float) (1.0 / Math.sqrt((double) k));
    }

    /**
     * a4: visiting order for one epoch. Record idx gets the 31-bit key hash64(seed,2,(epoch<<32)|idx)>>>33;
     * the order is ascending (key, idx). Packing key and idx into one non-negative long makes that a
     * plain primitive sort, identical in Java, C++ (sort of uint64) and CUDA (64-bit radix sort).
     */
    public static int[] shuffle(long seed, int epoch, int n) {
        long[] packed = new long[n];
        for (int idx = 0; idx < n; idx++) {
            long key = hash64(seed, STREAM_SHUFFLE, ((long) epoch << 32) | (long) idx) >>> 33;
            packed[idx] = (key << 32) | (long) idx;
        }
        Arrays.sort(packed);
        int[] order = new int[n];
        for (int j = 0; j < n; j++) order[j] = (int) (packed[j] & 0xFFFFFFFFL);
        return order;
    }

    /**
     * a2: one SGD update on (p_u, q_i) for rating r. Returns the error e = r - p.q used for it.
     * Dot product: f ascending, binary32 accumulate starting from 0. Both rows are updated from the
     * PRE-update values (simultaneous update).
     */
    public static float sgdUpdate(float[] P, int pOff, float[] Q, int qOff, int k,
                                  float r, float lr, float lambda) {
        float dot = 0.0f;
        for (int f = 0; f < k; f++) {
            dot = dot + P[pOff + f] * Q[qOff + f];
        }
        float e = r - dot;
        for (int f = 0; f < k; f++) {
            float pf = P[pOff + f];
            float qf = Q[qOff + f];
            P[pOff + f] = pf + lr * (e * qf - lambda * pf);
            Q[qOff + f] = qf + lr * (e * pf - lambda * qf);
        }
        return e;
    }

    /**
     * a1: the entry point. Ratings triplets, rank k, learning rate, lambda, epochs in; P and Q out.
     * Sequential: epoch e visits the records in shuffle(seed, e, n) order.
     */
    public static Factors factorize(int[] users, int[] items, float[] ratings,
                                    int nUsers, int nItems, int k,
                                    float lr, float lambda, int epochs, long seed) {
        if (users.length != items.length || users.length != ratings.length)
            throw new IllegalArgumentException("triplet arrays differ in length");
        if (k <= 0 || nUsers <= 0 || nItems <= 0 || epochs < 0)
            throw new IllegalArgumentException("bad shape");
        final int n = ratings.length;
        for (int t = 0; t < n; t++) {
            if (users[t] < 0 || users[t] >= nUsers || items[t] < 0 || items[t] >= nItems)
                throw new IllegalArgumentException("index out of range at record " + t);
        }
        float[] P = new float[nUsers * k];
        float[] Q = new float[nItems * k];
        float scale = defaultInitScale(k);
        initFactors(P, nUsers, k, seed, STREAM_P_INIT, scale);
        initFactors(Q, nItems, k, seed, STREAM_Q_INIT, scale);
        for (int epoch = 0; epoch < epochs; epoch++) {
            int[] order = shuffle(seed, epoch, n);
            for (int j = 0; j < n; j++) {
                int t = order[j];
                sgdUpdate(P, users[t] * k, Q, items[t] * k, k, ratings[t], lr, lambda);
            }
        }
        return new Factors(P, Q, nUsers, nItems, k);
    }

    /**
     * Thread-parallel (Hogwild) variant: T threads, thread w visits positions j = w, w+T, ... of the
     * same per-epoch order, no locks; threads join at every epoch end. Nondeterministic by design.
     */
    public static Factors factorizeThreaded(int[] users, int[] items, float[] ratings,
                                            int nUsers, int nItems, int k,
                                            float lr, float lambda, int epochs, long seed,
                                            int threads) throws InterruptedException {
        final int n = ratings.length;
        final float[] P = new float[nUsers * k];
        final float[] Q = new float[nItems * k];
        float scale = defaultInitScale(k);
        initFactors(P, nUsers, k, seed, STREAM_P_INIT, scale);
        initFactors(Q, nItems, k, seed, STREAM_Q_INIT, scale);
        for (int epoch = 0; epoch < epochs; epoch++) {
            final int[] order = shuffle(seed, epoch, n);
            Thread[] pool = new Thread[threads];
            for (int w = 0; w < threads; w++) {
                final int w0 = w;
                pool[w] = new Thread(() -> {
                    for (int j = w0; j < n; j += threads) {
                        int t = order[j];
                        sgdUpdate(P, users[t] * k, Q, items[t] * k, k, ratings[t], lr, lambda);
                    }
                });
                pool[w].start();
            }
            for (Thread th : pool) th.join();
        }
        return new Factors(P, Q, nUsers, nItems, k);
    }

    /** a5: sqrt( sum (r - p_u.q_i)^2 / n ); dot in binary32 (f ascending), sum in double. */
    public static double rmse(float[] P, float[] Q, int k, int[] users, int[] items, float[] ratings) {
        double sse = 0.0;
        final int n = ratings.length;
        for (int t = 0; t < n; t++) {
            int pOff = users[t] * k, qOff = items[t] * k;
            float dot = 0.0f;
            for (int f = 0; f < k; f++) dot = dot + P[pOff + f] * Q[qOff + f];
            float e = ratings[t] - dot;
            sse += (double) e * (double) e;
        }
        return n == 0 ? 0.0 : Math.sqrt(sse / (double) n);
    }

    /* ------------------------------------------------------------------------------------------
     * Synthetic power-law ratings (SURVEY.md 8d). Every value is a pure function of (seed, n).
     * ------------------------------------------------------------------------------------------ */

    public static final int PLANTED_RANK = 16;
    public static final float PLANTED_AMPLITUDE = 0.8660254f;   /* default: sqrt(0.75), planted dot of std 0.25 (noise-dominant sets) */
    public static final float NOISE_SCALE = 0.5f;               /* default noise scale (noise std 0.289)                              */
    public static final float SIGNAL_AMPLITUDE = 1.7320508f;    /* signal-dominant variant (SURVEY.md 8d): planted dot of std 1.0 ... */
    public static final float SIGNAL_NOISE_SCALE = 0.125f;      /* ... and noise std 0.072                                            */
    public static final long ID_MULT = 2654435761L;              /* prime > 2^31, so coprime to every id count */

    /** 53-bit uniform double in [0,1). */
    public static double uniform53(long seed, long stream, long ctr) {
        return (double) (hash64(seed, stream, ctr) >>> 11) * 0x1.0p-53;
    }

    /**
     * Shifted power-law rank in [0, count): y = c + (1-c)x, rank = floor(count * (y^a - c^a)/(1 - c^a)),
     * a = 2^log2Alpha computed by repeated squaring (only + - * / in binary64: bit-identical anywhere).
     */
    public static int skewedRank(double x, int count, int log2Alpha, double c) {
        double y = c + (1.0 - c) * x;
        double ca = c;
        for (int s = 0; s < log2Alpha; s++) { y = y * y; ca = ca * ca; }
        double t = (y - ca) / (1.0 - ca);
        long rank = (long) Math.floor((double) count * t);
        if (rank < 0) rank = 0;
        if (rank > count - 1) rank = count - 1;
        return (int) rank;
    }

    /** Fixed bijection on [0,count) so hot ids are not contiguous: (rank*ID_MULT + count/2) mod count. */
    public static int scatterId(int rank, int count) {
        return (int) ((((long) rank * ID_MULT) + (long) (count / 2)) % (long) count);
    }

    public static float plantedEntry(long seed, long stream, int row, int f, float amplitude) {
        return (uniform(seed, stream, (long) row * PLANTED_RANK + f) - 0.5f) * amplitude;
    }

    /** Record n of the default (noise-dominant) synthetic data set. */
    public static boolean syntheticRecord(long seed, long n, int nUsers, int nItems,
                                          int log2AlphaU, double cU, int log2AlphaI, double cI,
                                          int[] u, int[] i, float[] r) {
        return syntheticRecord(seed, n, nUsers, nItems, log2AlphaU, cU, log2AlphaI, cI, PLANTED_AMPLITUDE, NOISE_SCALE, u, i, r);
    }

    /** Record n of the synthetic data set: fills u[0], i[0], r[0]; returns true when n is held out. */
    public static boolean syntheticRecord(long seed, long n, int nUsers, int nItems,
                                          int log2AlphaU, double cU, int log2AlphaI, double cI,
                                          float amplitude, float noiseScale,
                                          int[] u, int[] i, float[] r) {
        int uu = scatterId(skewedRank(uniform53(seed, STREAM_USER, n), nUsers, log2AlphaU, cU), nUsers);
        int ii = scatterId(skewedRank(uniform53(seed, STREAM_ITEM, n), nItems, log2AlphaI, cI), nItems);
        float dot = 0.0f;
        for (int f = 0; f < PLANTED_RANK; f++) {
            dot = dot + plantedEntry(seed, STREAM_PSTAR, uu, f, amplitude) * plantedEntry(seed, STREAM_QSTAR, ii, f, amplitude);
        }
        float noise = 0.0f;
        for (int j = 0; j < 4; j++) noise = noise + uniform(seed, STREAM_NOISE, 4L * n + j);
        noise = noise - 2.0f;
        float rating = 3.5f + dot;
        rating = rating + noiseScale * noise;
        if (rating < 1.0f) rating = 1.0f;
        if (rating > 5.0f) rating = 5.0f;
        u[0] = uu; i[0] = ii; r[0] = rating;
        return Long.remainderUnsigned(hash64(seed, STREAM_HELDOUT, n), 10L) == 0L;
    }

    /* ------------------------------------------------------------------------------------------
     * Model extension (SURVEY.md 8f.4): global mean and user / item biases. Same loop, same order;
     * prediction mu + b_u + b_i + p_u.q_i. With both switches off this is factorize().
     * ------------------------------------------------------------------------------------------ */

    /** Factors plus the extension's terms. */
    public static final class Model {
        public final float[] P, Q, userBias, itemBias;
        public final float globalMean;
        public final int nUsers, nItems, k;
        Model(float[] P, float[] Q, float[] bu, float[] bi, float mu, int nUsers, int nItems, int k) {
            this.P = P; this.Q = Q; this.userBias = bu; this.itemBias = bi; this.globalMean = mu;
            this.nUsers = nUsers; this.nItems = nItems; this.k = k;
        }
    }

    /**
     * Training mean, defined on exact integers so that every implementation (sequential, threaded, GPU) gets the same
     * binary32 value whatever its summation order: S = sum floor(r * 2^20) (the product is exact in binary64),
     * mu = (float) (S / n / 2^20).
     */
    public static float globalMean(float[] ratings) {
        long s = 0L;
        for (float r : ratings) s += (long) Math.floor((double) r * 1048576.0);
        return ratings.length == 0 ? 0.0f : (float) ((double) s / (double) ratings.length / 1048576.0);
    }

    /**
     * One update of the extended model on the CENTRED rating rc = r - mu. pred = (dot + b_u) + b_i, e = rc - pred;
     * biases (when present): b <- b + lr * (e - lambda * b), from the pre-update values; factors as in sgdUpdate.
     */
    public static float sgdUpdateModel(float[] P, int pOff, float[] Q, int qOff, int k, float[] bu, int u, float[] bi, int i,
                                       float rc, float lr, float lambda) {
        float dot = 0.0f;
        for (int f = 0; f < k; f++) dot = dot + P[pOff + f] * Q[qOff + f];
        float e;
        if (bu != null) {
            float pred = dot + bu[u];
            pred = pred + bi[i];
            e = rc - pred;
            float b0 = bu[u], b1 = bi[i];
            bu[u] = b0 + lr * (e - lambda * b0);
            bi[i] = b1 + lr * (e - lambda * b1);
        } else {
            e = rc - dot;
        }
        for (int f = 0; f < k; f++) {
            float pf = P[pOff + f], qf = Q[qOff + f];
            P[pOff + f] = pf + lr * (e * qf - lambda * pf);
            Q[qOff + f] = qf + lr * (e * pf - lambda * qf);
        }
        return e;
    }

    public static Model factorizeModel(int[] users, int[] items, float[] ratings, int nUsers, int nItems, int k,
                                       float lr, float lambda, int epochs, long seed, boolean useGlobalMean, boolean useBiases) {
        if (users.length != items.length || users.length != ratings.length)
            throw new IllegalArgumentException("triplet arrays differ in length");
        if (k <= 0 || nUsers <= 0 || nItems <= 0 || epochs < 0) throw new IllegalArgumentException("bad shape");
        final int n = ratings.length;
        float[] P = new float[nUsers * k], Q = new float[nItems * k];
        float scale = defaultInitScale(k);
        initFactors(P, nUsers, k, seed, STREAM_P_INIT, scale);
        initFactors(Q, nItems, k, seed, STREAM_Q_INIT, scale);
        float[] bu = useBiases ? new float[nUsers] : null, bi = useBiases ? new float[nItems] : null;    // biases start at 0
        final float mu = useGlobalMean ? globalMean(ratings) : 0.0f;
        for (int epoch = 0; epoch < epochs; epoch++) {
            int[] order = shuffle(seed, epoch, n);
            for (int j = 0; j < n; j++) {
                int t = order[j];
                sgdUpdateModel(P, users[t] * k, Q, items[t] * k, k, bu, users[t], bi, items[t], ratings[t] - mu, lr, lambda);
            }
        }
        return new Model(P, Q, bu, bi, mu, nUsers, nItems, k);
    }

    /**
     * Learning-rate schedule of the extension: lr_0 = lr, lr_(e+1) = lr_e * decay, one binary32 multiply per epoch
     * (so every implementation gets the same rate; Math.pow would not be bit-reproducible). decay = 1: the constant rate.
     */
    public static float learningRate(float lr, float decay, int epoch) {
        float l = lr;
        for (int e = 0; e < epoch; e++) l = l * decay;
        return l;
    }

    /** Result of factorizeEarlyStop: the model after the last epoch run, the epochs run and the validation curve. */
    public static final class EarlyStopResult {
        public final Model model;
        public final int epochsRun;
        public final double[] validationRmse;
        EarlyStopResult(Model model, int epochsRun, double[] curve) { this.model = model; this.epochsRun = epochsRun; this.validationRmse = curve; }
    }

    /**
     * factorizeModel with the schedule and early stopping: after every epoch the RMSE v on the validation triplets is
     * taken; v < best * (1 - minDelta) makes it the new best and clears the strike count, anything else is a strike, and
     * `patience` strikes in a row end the training (patience = 0: never). The model returned is the one after the last epoch run.
     */
    public static EarlyStopResult factorizeEarlyStop(int[] users, int[] items, float[] ratings, int[] vUsers, int[] vItems, float[] vRatings,
                                                     int nUsers, int nItems, int k, float lr, float lambda, int maxEpochs, long seed,
                                                     boolean useGlobalMean, boolean useBiases, float lrDecay, int patience, float minDelta) {
        if (users.length != items.length || users.length != ratings.length)
            throw new IllegalArgumentException("triplet arrays differ in length");
        if (k <= 0 || nUsers <= 0 || nItems <= 0 || maxEpochs < 0) throw new IllegalArgumentException("bad shape");
        if (!(lrDecay > 0.0f) || lrDecay > 1.0f || patience < 0 || !(minDelta >= 0.0f) || minDelta >= 1.0f)
            throw new IllegalArgumentException("bad schedule");
        final int n = ratings.length;
        float[] P = new float[nUsers * k], Q = new float[nItems * k];
        float scale = defaultInitScale(k);
        initFactors(P, nUsers, k, seed, STREAM_P_INIT, scale);
        initFactors(Q, nItems, k, seed, STREAM_Q_INIT, scale);
        float[] bu = useBiases ? new float[nUsers] : null, bi = useBiases ? new float[nItems] : null;
        final float mu = useGlobalMean ? globalMean(ratings) : 0.0f;
        Model m = new Model(P, Q, bu, bi, mu, nUsers, nItems, k);
        double[] curve = new double[maxEpochs];
        double best = Double.POSITIVE_INFINITY;
        int strikes = 0, ran = 0;
        float lrNow = lr;
        for (int epoch = 0; epoch < maxEpochs; epoch++) {
            int[] order = shuffle(seed, epoch, n);
            for (int j = 0; j < n; j++) {
                int t = order[j];
                sgdUpdateModel(P, users[t] * k, Q, items[t] * k, k, bu, users[t], bi, items[t], ratings[t] - mu, lrNow, lambda);
            }
            lrNow = lrNow * lrDecay;
            ran = epoch + 1;
            double v = rmseModel(m, vUsers, vItems, vRatings);
            curve[epoch] = v;
            if (patience > 0) {
                if (v < best * (1.0 - (double) minDelta)) { best = v; strikes = 0; }
                else if (++strikes >= patience) break;
            }
        }
        return new EarlyStopResult(m, ran, java.util.Arrays.copyOf(curve, ran));
    }

    /** RMSE of the extended model: e = (r - mu) - ((dot + b_u) + b_i). */
    public static double rmseModel(Model m, int[] users, int[] items, float[] ratings) {
        double sse = 0.0;
        final int n = ratings.length, k = m.k;
        for (int t = 0; t < n; t++) {
            int pOff = users[t] * k, qOff = items[t] * k;
            float dot = 0.0f;
            for (int f = 0; f < k; f++) dot = dot + m.P[pOff + f] * m.Q[qOff + f];
            float pred = dot;
            if (m.userBias != null) { pred = pred + m.userBias[users[t]]; pred = pred + m.itemBias[items[t]]; }
            float e = (ratings[t] - m.globalMean) - pred;
            sse += (double) e * (double) e;
        }
        return n == 0 ? 0.0 : Math.sqrt(sse / (double) n);
    }

    /* ------------------------------------------------------------------------------------------
     * Mixed-precision storage (SURVEY.md 8f.3): the rows of P are KEPT as binary16 (short bit patterns), every
     * operation of the update rule stays binary32. A row is widened exactly before an update and narrowed after it with
     * stochastic rounding: 8 random bits per value decide among the 13 mantissa bits binary16 drops; they come from a
     * counter hash of (seed, epoch, u, i, chunk of 4 values) -- no state, the same bits in any visiting order.
     * Q stays binary32. (Float.float16ToFloat / floatToFloat16: JDK 20+, round to nearest even.)
     * ------------------------------------------------------------------------------------------ */

    /** The 32 random bits of chunk c (values 4c .. 4c+3) of row u at the update (u, i) of `epoch`: lowbias32 finaliser. */
    public static int srWord(long seed, int epoch, int u, int i, int c) {
        int x = (int) (seed ^ (seed >>> 32)) + u * 0x9E3779B1 + i * 0x85EBCA77 + (epoch * 0x10001 + c) * 0xC2B2AE3D;
        x ^= x >>> 16; x *= 0x7FEB352D; x ^= x >>> 15; x *= 0x846CA68B; x ^= x >>> 16;
        return x;
    }

    /** Narrow value j (0..3) of a chunk: push the magnitude up by (byte j of word) * 32 + 16, cut the 13 low mantissa bits, convert. */
    public static short storeF16Sr(float v, int word, int j) {
        int rho = (((word >>> (8 * j)) & 0xFF) << 5) | 0x10;
        return Float.floatToFloat16(Float.intBitsToFloat((Float.floatToRawIntBits(v) + rho) & 0xFFFFE000));
    }

    /** sgdUpdate on a binary16 p_u (P16 holds bit patterns); returns the error. */
    public static float sgdUpdateMixed(short[] P16, int pOff, float[] Q, int qOff, int k, float r, float lr, float lambda,
                                       long seed, int epoch, int u, int i) {
        float[] p = new float[k];
        for (int f = 0; f < k; f++) p[f] = Float.float16ToFloat(P16[pOff + f]);
        float e = sgdUpdate(p, 0, Q, qOff, k, r, lr, lambda);
        for (int c = 0; c < k / 4; c++) {
            int w = srWord(seed, epoch, u, i, c);
            for (int j = 0; j < 4; j++) P16[pOff + 4 * c + j] = storeF16Sr(p[4 * c + j], w, j);
        }
        return e;
    }

    /** factorize with P kept in binary16 (k a multiple of 4); P comes back widened (exactly). */
    public static Factors factorizeMixed(int[] users, int[] items, float[] ratings, int nUsers, int nItems, int k,
                                         float lr, float lambda, int epochs, long seed) {
        if (users.length != items.length || users.length != ratings.length)
            throw new IllegalArgumentException("triplet arrays differ in length");
        if (k <= 0 || k % 4 != 0 || nUsers <= 0 || nItems <= 0 || epochs < 0) throw new IllegalArgumentException("bad shape");
        final int n = ratings.length;
        float[] P = new float[nUsers * k], Q = new float[nItems * k];
        float scale = defaultInitScale(k);
        initFactors(P, nUsers, k, seed, STREAM_P_INIT, scale);
        initFactors(Q, nItems, k, seed, STREAM_Q_INIT, scale);
        short[] P16 = new short[nUsers * k];
        for (int e = 0; e < P16.length; e++) P16[e] = Float.floatToFloat16(P[e]);      // round to nearest even, once
        for (int epoch = 0; epoch < epochs; epoch++) {
            int[] order = shuffle(seed, epoch, n);
            for (int j = 0; j < n; j++) {
                int t = order[j];
                sgdUpdateMixed(P16, users[t] * k, Q, items[t] * k, k, ratings[t], lr, lambda, seed, epoch, users[t], items[t]);
            }
        }
        for (int e = 0; e < P16.length; e++) P[e] = Float.float16ToFloat(P16[e]);
        return new Factors(P, Q, nUsers, nItems, k);
    }

    /** ML-100K-shaped demo: sequential and threaded, updates/s and held-out RMSE. */
    public static void main(String[] args) throws Exception {
        final long seed = 20261018L;
        final int nUsers = 943, nItems = 1682, total = 100_000, k = 32, epochs = 20;
        final float lr = 0.01f, lambda = 0.05f;
        int[] tu = new int[total], ti = new int[total]; float[] tr = new float[total];
        int[] hu = new int[total], hi = new int[total]; float[] hr = new float[total];
        int nt = 0, nh = 0;
        int[] u = new int[1], i = new int[1]; float[] r = new float[1];
        for (long n = 0; n < total; n++) {
            boolean held = syntheticRecord(seed, n, nUsers, nItems, 2, 0.25, 3, 0.375, u, i, r);
            if (held) { hu[nh] = u[0]; hi[nh] = i[0]; hr[nh] = r[0]; nh++; }
            else      { tu[nt] = u[0]; ti[nt] = i[0]; tr[nt] = r[0]; nt++; }
        }
        tu = Arrays.copyOf(tu, nt); ti = Arrays.copyOf(ti, nt); tr = Arrays.copyOf(tr, nt);
        hu = Arrays.copyOf(hu, nh); hi = Arrays.copyOf(hi, nh); hr = Arrays.copyOf(hr, nh);
        long t0 = System.nanoTime();
        Factors seq = factorize(tu, ti, tr, nUsers, nItems, k, lr, lambda, epochs, seed);
        double sSeq = (System.nanoTime() - t0) * 1e-9;
        System.out.printf("sequential: %.3e updates/s, held-out RMSE %.6f%n",
                (double) nt * epochs / sSeq, rmse(seq.P, seq.Q, k, hu, hi, hr));
        int threads = Runtime.getRuntime().availableProcessors();
        t0 = System.nanoTime();
        Factors par = factorizeThreaded(tu, ti, tr, nUsers, nItems, k, lr, lambda, epochs, seed, threads);
        double sPar = (System.nanoTime() - t0) * 1e-9;
        System.out.printf("threaded(%d): %.3e updates/s, held-out RMSE %.6f%n", threads,
                (double) nt * epochs / sPar, rmse(par.P, par.Q, k, hu, hi, hr));
    }
}
