/*
 * MatrixFactorizationSGDGpu.java -- the Java host of the B200-native engine (Panama FFM, JDK 22+).
 *
 * Drop-in for the factorization path of MatrixFactorizationSGD (README.md:1; stand-in
 * baseline/java/MatrixFactorizationSGD.java:109 factorize): same arguments in, P and Q out, but the
 * epochs run in libmfsgd.so on the GPU(s). Binds the C ABI of include/mfsgd.h with plain downcalls;
 * no JNI glue, no native code on the Java side. A non-zero return becomes IllegalStateException with
 * the library's message; there is no CPU fallback.
 *
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK there); tests/c/abi_harness.c and the ctypes binding
 * (matrixfactorizationsgd.java_b200/_capi.py) exercise exactly the same symbols, struct layout
 * (tests/test_abi_cpu.py::test_struct_sizes_match_c_layout pins the offsets used below) and call order.
 *
 *   javac --release 22 java/MatrixFactorizationSGDGpu.java baseline/java/MatrixFactorizationSGD.java
 *   java --enable-native-access=ALL-UNNAMED -Dmfsgd.lib=/path/to/libmfsgd.so MatrixFactorizationSGDGpu
 */
import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.SymbolLookup;
import java.lang.foreign.ValueLayout;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

public final class MatrixFactorizationSGDGpu {

    public static final int MODE_DETERMINISTIC = 0, MODE_HOGWILD = 1, MODE_DSGD = 2;

    /** struct mfsgd_config (include/mfsgd.h, ABI 3), 248 bytes; field order and padding as laid out by the C compiler. */
    static final StructLayout CONFIG = MemoryLayout.structLayout(
            JAVA_INT.withName("n_users"), JAVA_INT.withName("n_items"), JAVA_INT.withName("k"),
            JAVA_FLOAT.withName("lr"), JAVA_FLOAT.withName("lambda"), JAVA_FLOAT.withName("init_scale"),
            JAVA_LONG.withName("seed"),                                   /* offset 24 */
            JAVA_INT.withName("mode"), JAVA_INT.withName("n_gpus"), JAVA_INT.withName("stripes_per_gpu"),
            JAVA_INT.withName("shards_per_gpu"), JAVA_INT.withName("scatter"), JAVA_INT.withName("flags"),
            JAVA_INT.withName("device"), JAVA_INT.withName("world_size"), JAVA_INT.withName("rank"),
            MemoryLayout.sequenceLayout(128, JAVA_BYTE).withName("nccl_id"),   /* offset 68 */
            JAVA_INT.withName("ctas_per_sm"),                             /* offset 196 */
            JAVA_INT.withName("rounds"), JAVA_FLOAT.withName("hot_share"), JAVA_INT.withName("hot_chunk"),
            JAVA_FLOAT.withName("merge_boost"), JAVA_INT.withName("model"),  /* offsets 212, 216 */
            JAVA_FLOAT.withName("p_atomic_threshold"), JAVA_FLOAT.withName("lr_decay"),
            JAVA_INT.withName("early_stop_patience"), JAVA_FLOAT.withName("early_stop_min_delta"),   /* offsets 228, 232 */
            JAVA_INT.withName("p_storage"),                                /* offset 236 */
            MemoryLayout.sequenceLayout(2, JAVA_INT).withName("reserved"));  /* 248 bytes, no tail padding */

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(
            System.getProperty("mfsgd.lib", "libmfsgd.so"), Arena.global());

    private static MethodHandle down(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(
                () -> new UnsatisfiedLinkError("libmfsgd.so lacks " + name)), fd);
    }

    /* int mfsgd_config_default(mfsgd_config*) */
    private static final MethodHandle CONFIG_DEFAULT = down("mfsgd_config_default", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    /* const char* mfsgd_last_error(void) */
    private static final MethodHandle LAST_ERROR = down("mfsgd_last_error", FunctionDescriptor.of(ADDRESS));
    /* int mfsgd_factorize(const int32_t*, const int32_t*, const float*, int64_t, const mfsgd_config*, int32_t, float*, float*) */
    private static final MethodHandle FACTORIZE = down("mfsgd_factorize", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    /* handle API, for callers that keep the data resident across calls */
    private static final MethodHandle CREATE = down("mfsgd_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle DESTROY = down("mfsgd_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle LOAD_RATINGS = down("mfsgd_load_ratings", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG));
    private static final MethodHandle INIT_FACTORS = down("mfsgd_init_factors", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle TRAIN = down("mfsgd_train", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    private static final MethodHandle GET_FACTORS = down("mfsgd_get_factors", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle RMSE = down("mfsgd_rmse", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));

    /* model extension / schedule / mixed storage (SURVEY.md 8f.3, 8f.4): the calls the handle API adds for them */
    private static final MethodHandle SET_FACTORS = down("mfsgd_set_factors", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle LOAD_HELDOUT = down("mfsgd_load_heldout", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG));
    private static final MethodHandle GET_MODEL = down("mfsgd_get_model", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle SET_BIASES = down("mfsgd_set_biases", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle GET_PROGRESS = down("mfsgd_get_progress", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle SET_EVAL_EVERY_EPOCH = down("mfsgd_set_eval_every_epoch", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));

    /* mfsgd_config members the extension rows set (offsets: CONFIG above; tests/test_java_cpu.py checks them against the C struct) */
    static final long OFF_MODEL = 216, OFF_LR_DECAY = 224, OFF_ES_PATIENCE = 228, OFF_ES_MIN_DELTA = 232, OFF_P_STORAGE = 236;
    public static final int MODEL_GLOBAL_MEAN = 1, MODEL_BIASES = 2;       /* mfsgd_config.model bits */
    public static final int STORAGE_F32 = 0, STORAGE_F16 = 1;               /* mfsgd_config.p_storage */
    /* struct mfsgd_epoch_stats (include/mfsgd.h): 72 bytes, heldout_rmse (double) at offset 40 */
    static final long EPOCH_STATS_BYTES = 72, OFF_STATS_HELDOUT_RMSE = 40;

    private static final MethodHandle READ_RATINGS = down("mfsgd_read_ratings", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    private static final MethodHandle FREE_RATINGS = down("mfsgd_free_ratings", FunctionDescriptor.ofVoid(ADDRESS));

    /** A parsed ratings file: dense triplets for factorize plus the file's id of every dense row. */
    public static final class RatingsFile {
        public int[] users, items;
        public float[] ratings;
        public long[] userIds, itemIds;
        public int nUsers, nItems;
    }

    /**
     * mfsgd_read_ratings: MovieLens u.data / ratings.csv / ratings.dat or Netflix-Prize text (format 0 = detect)
     * into the triplet arrays factorize takes. struct mfsgd_ratings (include/mfsgd.h, 64 bytes): users@0 items@8
     * ratings@16 (pointers), n@24 (int64), n_users@32, n_items@36 (int32), user_ids@40, item_ids@48 (pointers).
     */
    public static RatingsFile readRatings(String path, int format) {
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment out = arena.allocate(64, 8);
            check((int) READ_RATINGS.invokeExact(arena.allocateFrom(path), format, out));
            try {
                RatingsFile f = new RatingsFile();
                final long n = out.get(JAVA_LONG, 24);
                f.nUsers = out.get(JAVA_INT, 32);
                f.nItems = out.get(JAVA_INT, 36);
                f.users = out.get(ADDRESS, 0).reinterpret(4 * n).toArray(JAVA_INT);
                f.items = out.get(ADDRESS, 8).reinterpret(4 * n).toArray(JAVA_INT);
                f.ratings = out.get(ADDRESS, 16).reinterpret(4 * n).toArray(JAVA_FLOAT);
                f.userIds = out.get(ADDRESS, 40).reinterpret(8L * f.nUsers).toArray(JAVA_LONG);
                f.itemIds = out.get(ADDRESS, 48).reinterpret(8L * f.nItems).toArray(JAVA_LONG);
                return f;
            } finally {
                FREE_RATINGS.invokeExact(out);
            }
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    private static void check(int rc) throws Throwable {
        if (rc != 0) {
            MemorySegment msg = ((MemorySegment) LAST_ERROR.invokeExact()).reinterpret(512);
            throw new IllegalStateException("mfsgd error " + rc + ": " + msg.getString(0));
        }
    }

    private static MemorySegment config(Arena arena, int nUsers, int nItems, int k, float lr, float lambda,
                                        long seed, int mode, int nGpus) throws Throwable {
        MemorySegment cfg = arena.allocate(CONFIG);
        check((int) CONFIG_DEFAULT.invokeExact(cfg));
        cfg.set(JAVA_INT, 0, nUsers);
        cfg.set(JAVA_INT, 4, nItems);
        cfg.set(JAVA_INT, 8, k);
        cfg.set(JAVA_FLOAT, 12, lr);
        cfg.set(JAVA_FLOAT, 16, lambda);
        cfg.set(JAVA_LONG, 24, seed);
        cfg.set(JAVA_INT, 32, mode);
        cfg.set(JAVA_INT, 36, nGpus);
        return cfg;
    }

    /**
     * Same contract as MatrixFactorizationSGD.factorize (stand-in line 109): triplets, rank, learning rate,
     * lambda, epochs, seed in; row-major P (nUsers x k) and Q (nItems x k) out. One GPU, Hogwild.
     */
    public static MatrixFactorizationSGD.Factors factorize(int[] users, int[] items, float[] ratings,
                                                          int nUsers, int nItems, int k,
                                                          float lr, float lambda, int epochs, long seed) {
        return factorize(users, items, ratings, nUsers, nItems, k, lr, lambda, epochs, seed, MODE_HOGWILD, 1);
    }

    /** mode = MODE_DETERMINISTIC reproduces the sequential stand-in update for update; MODE_DSGD uses nGpus GPUs. */
    public static MatrixFactorizationSGD.Factors factorize(int[] users, int[] items, float[] ratings,
                                                          int nUsers, int nItems, int k,
                                                          float lr, float lambda, int epochs, long seed,
                                                          int mode, int nGpus) {
        if (users.length != items.length || users.length != ratings.length)
            throw new IllegalArgumentException("triplet arrays differ in length");
        if (k <= 0 || nUsers <= 0 || nItems <= 0 || epochs < 0)
            throw new IllegalArgumentException("bad shape");
        final long n = ratings.length;
        try (Arena arena = Arena.ofConfined()) {
            /* off-heap copies: the library reads them only during the call and never retains them */
            MemorySegment u = arena.allocateFrom(JAVA_INT, users);
            MemorySegment i = arena.allocateFrom(JAVA_INT, items);
            MemorySegment r = arena.allocateFrom(JAVA_FLOAT, ratings);
            MemorySegment p = arena.allocate(JAVA_FLOAT, (long) nUsers * k);
            MemorySegment q = arena.allocate(JAVA_FLOAT, (long) nItems * k);
            MemorySegment cfg = config(arena, nUsers, nItems, k, lr, lambda, seed, mode, nGpus);
            check((int) FACTORIZE.invokeExact(u, i, r, n, cfg, epochs, p, q));
            return newFactors(p.toArray(JAVA_FLOAT), q.toArray(JAVA_FLOAT), nUsers, nItems, k);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /** Stand-in line 169 (rmse) evaluated by the GPU RMSE kernel on factors the caller holds. */
    public static double rmse(float[] P, float[] Q, int k, int[] users, int[] items, float[] ratings) {
        final int nUsers = P.length / k, nItems = Q.length / k;
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment cfg = config(arena, nUsers, nItems, k, 1e-3f, 0.0f, 0L, MODE_HOGWILD, 1);
            MemorySegment hp = arena.allocate(ADDRESS);
            check((int) CREATE.invokeExact(cfg, hp));
            MemorySegment h = hp.get(ADDRESS, 0);
            try {
                check((int) LOAD_RATINGS.invokeExact(h, MemorySegment.NULL, MemorySegment.NULL, MemorySegment.NULL, 0L));
                check((int) SET_FACTORS.invokeExact(h, arena.allocateFrom(JAVA_FLOAT, P), arena.allocateFrom(JAVA_FLOAT, Q)));
                MemorySegment out = arena.allocate(ValueLayout.JAVA_DOUBLE);
                check((int) RMSE.invokeExact(h, arena.allocateFrom(JAVA_INT, users), arena.allocateFrom(JAVA_INT, items),
                        arena.allocateFrom(JAVA_FLOAT, ratings), (long) ratings.length, out));
                return out.get(ValueLayout.JAVA_DOUBLE, 0);
            } finally {
                DESTROY.invokeExact(h);
            }
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /**
     * Stand-in factorizeMixed (:439): the rows of P are KEPT as binary16 on the device (narrowed with stochastic rounding from a
     * counter hash), every operation of the rule stays binary32; P comes back widened exactly. One GPU, Hogwild.
     */
    public static MatrixFactorizationSGD.Factors factorizeMixed(int[] users, int[] items, float[] ratings, int nUsers, int nItems, int k,
                                                               float lr, float lambda, int epochs, long seed) {
        if (users.length != items.length || users.length != ratings.length)
            throw new IllegalArgumentException("triplet arrays differ in length");
        if (k <= 0 || k % 4 != 0 || nUsers <= 0 || nItems <= 0 || epochs < 0) throw new IllegalArgumentException("bad shape");
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment p = arena.allocate(JAVA_FLOAT, (long) nUsers * k);
            MemorySegment q = arena.allocate(JAVA_FLOAT, (long) nItems * k);
            MemorySegment cfg = config(arena, nUsers, nItems, k, lr, lambda, seed, MODE_HOGWILD, 1);
            cfg.set(JAVA_INT, OFF_P_STORAGE, STORAGE_F16);
            check((int) FACTORIZE.invokeExact(arena.allocateFrom(JAVA_INT, users), arena.allocateFrom(JAVA_INT, items),
                    arena.allocateFrom(JAVA_FLOAT, ratings), (long) ratings.length, cfg, epochs, p, q));
            return newFactors(p.toArray(JAVA_FLOAT), q.toArray(JAVA_FLOAT), nUsers, nItems, k);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /**
     * Stand-in factorizeModel (:305): r ~ mu + b_u + b_i + p_u . q_i. The mean is taken on the device while the ratings are
     * counted, ratings are stored centred, the biases ride through the update kernels beside their rows.
     */
    public static MatrixFactorizationSGD.Model factorizeModel(int[] users, int[] items, float[] ratings, int nUsers, int nItems, int k,
                                                             float lr, float lambda, int epochs, long seed,
                                                             boolean useGlobalMean, boolean useBiases) {
        return trainModel(users, items, ratings, null, null, null, nUsers, nItems, k, lr, lambda, epochs, seed,
                useGlobalMean, useBiases, 1.0f, 0, 0.0f).model;
    }

    /**
     * Stand-in factorizeEarlyStop (:350): the schedule lr_(e+1) = lr_e * lrDecay and the early-stopping rule on the validation
     * RMSE, both evaluated inside mfsgd_train (the validation triplets live on the device).
     */
    public static MatrixFactorizationSGD.EarlyStopResult factorizeEarlyStop(int[] users, int[] items, float[] ratings,
                                                                           int[] vUsers, int[] vItems, float[] vRatings,
                                                                           int nUsers, int nItems, int k, float lr, float lambda, int maxEpochs, long seed,
                                                                           boolean useGlobalMean, boolean useBiases,
                                                                           float lrDecay, int patience, float minDelta) {
        if (!(lrDecay > 0.0f) || lrDecay > 1.0f || patience < 0 || !(minDelta >= 0.0f) || minDelta >= 1.0f)
            throw new IllegalArgumentException("bad schedule");
        if (vUsers.length != vItems.length || vUsers.length != vRatings.length)
            throw new IllegalArgumentException("triplet arrays differ in length");
        return trainModel(users, items, ratings, vUsers, vItems, vRatings, nUsers, nItems, k, lr, lambda, maxEpochs, seed,
                useGlobalMean, useBiases, lrDecay, patience, minDelta);
    }

    private static MatrixFactorizationSGD.EarlyStopResult trainModel(int[] users, int[] items, float[] ratings,
                                                                    int[] vUsers, int[] vItems, float[] vRatings,
                                                                    int nUsers, int nItems, int k, float lr, float lambda, int epochs, long seed,
                                                                    boolean useGlobalMean, boolean useBiases,
                                                                    float lrDecay, int patience, float minDelta) {
        if (users.length != items.length || users.length != ratings.length)
            throw new IllegalArgumentException("triplet arrays differ in length");
        if (k <= 0 || nUsers <= 0 || nItems <= 0 || epochs < 0) throw new IllegalArgumentException("bad shape");
        final boolean validated = vRatings != null;
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment cfg = config(arena, nUsers, nItems, k, lr, lambda, seed, MODE_HOGWILD, 1);
            cfg.set(JAVA_INT, OFF_MODEL, (useGlobalMean ? MODEL_GLOBAL_MEAN : 0) | (useBiases ? MODEL_BIASES : 0));
            cfg.set(JAVA_FLOAT, OFF_LR_DECAY, lrDecay);
            cfg.set(JAVA_INT, OFF_ES_PATIENCE, patience);
            cfg.set(JAVA_FLOAT, OFF_ES_MIN_DELTA, minDelta);
            MemorySegment hp = arena.allocate(ADDRESS);
            check((int) CREATE.invokeExact(cfg, hp));
            MemorySegment h = hp.get(ADDRESS, 0);
            try {
                check((int) LOAD_RATINGS.invokeExact(h, arena.allocateFrom(JAVA_INT, users), arena.allocateFrom(JAVA_INT, items),
                        arena.allocateFrom(JAVA_FLOAT, ratings), (long) ratings.length));
                if (validated) {
                    check((int) LOAD_HELDOUT.invokeExact(h, arena.allocateFrom(JAVA_INT, vUsers), arena.allocateFrom(JAVA_INT, vItems),
                            arena.allocateFrom(JAVA_FLOAT, vRatings), (long) vRatings.length));
                    check((int) SET_EVAL_EVERY_EPOCH.invokeExact(h, 1));
                }
                check((int) INIT_FACTORS.invokeExact(h));
                MemorySegment stats = validated && epochs > 0 ? arena.allocate(EPOCH_STATS_BYTES * epochs, 8) : MemorySegment.NULL;
                if (epochs > 0) check((int) TRAIN.invokeExact(h, epochs, stats));
                MemorySegment ran = arena.allocate(JAVA_INT);
                check((int) GET_PROGRESS.invokeExact(h, ran, MemorySegment.NULL, MemorySegment.NULL));
                final int epochsRun = ran.get(JAVA_INT, 0);
                MemorySegment p = arena.allocate(JAVA_FLOAT, (long) nUsers * k);
                MemorySegment q = arena.allocate(JAVA_FLOAT, (long) nItems * k);
                check((int) GET_FACTORS.invokeExact(h, p, q));
                MemorySegment mu = arena.allocate(JAVA_FLOAT);
                MemorySegment bu = useBiases ? arena.allocate(JAVA_FLOAT, (long) nUsers) : MemorySegment.NULL;
                MemorySegment bi = useBiases ? arena.allocate(JAVA_FLOAT, (long) nItems) : MemorySegment.NULL;
                check((int) GET_MODEL.invokeExact(h, mu, bu, bi));
                double[] curve = new double[validated ? epochsRun : 0];
                for (int e = 0; e < curve.length; e++) curve[e] = stats.get(ValueLayout.JAVA_DOUBLE, EPOCH_STATS_BYTES * e + OFF_STATS_HELDOUT_RMSE);
                /* Model's and EarlyStopResult's constructors are package-private: both classes live in the default package. */
                MatrixFactorizationSGD.Model m = new MatrixFactorizationSGD.Model(p.toArray(JAVA_FLOAT), q.toArray(JAVA_FLOAT),
                        useBiases ? bu.toArray(JAVA_FLOAT) : null, useBiases ? bi.toArray(JAVA_FLOAT) : null,
                        mu.get(JAVA_FLOAT, 0), nUsers, nItems, k);
                return new MatrixFactorizationSGD.EarlyStopResult(m, epochsRun, curve);
            } finally {
                DESTROY.invokeExact(h);
            }
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /** Stand-in rmseModel (:389): e = (r - mu) - ((p_u . q_i + b_u) + b_i), evaluated by the GPU RMSE kernel. */
    public static double rmseModel(MatrixFactorizationSGD.Model m, int[] users, int[] items, float[] ratings) {
        final boolean biased = m.userBias != null;
        float[] centred = new float[ratings.length];
        for (int t = 0; t < ratings.length; t++) centred[t] = ratings[t] - m.globalMean;      /* one binary32 subtraction, as :399 */
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment cfg = config(arena, m.nUsers, m.nItems, m.k, 1e-3f, 0.0f, 0L, MODE_HOGWILD, 1);
            cfg.set(JAVA_INT, OFF_MODEL, biased ? MODEL_BIASES : 0);
            MemorySegment hp = arena.allocate(ADDRESS);
            check((int) CREATE.invokeExact(cfg, hp));
            MemorySegment h = hp.get(ADDRESS, 0);
            try {
                check((int) LOAD_RATINGS.invokeExact(h, MemorySegment.NULL, MemorySegment.NULL, MemorySegment.NULL, 0L));
                check((int) SET_FACTORS.invokeExact(h, arena.allocateFrom(JAVA_FLOAT, m.P), arena.allocateFrom(JAVA_FLOAT, m.Q)));
                if (biased)
                    check((int) SET_BIASES.invokeExact(h, arena.allocateFrom(JAVA_FLOAT, m.userBias), arena.allocateFrom(JAVA_FLOAT, m.itemBias)));
                MemorySegment out = arena.allocate(ValueLayout.JAVA_DOUBLE);
                check((int) RMSE.invokeExact(h, arena.allocateFrom(JAVA_INT, users), arena.allocateFrom(JAVA_INT, items),
                        arena.allocateFrom(JAVA_FLOAT, centred), (long) centred.length, out));
                return out.get(ValueLayout.JAVA_DOUBLE, 0);
            } finally {
                DESTROY.invokeExact(h);
            }
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    private static MatrixFactorizationSGD.Factors newFactors(float[] P, float[] Q, int nUsers, int nItems, int k) {
        /* Factors' constructor is package-private: both classes live in the default package. */
        return new MatrixFactorizationSGD.Factors(P, Q, nUsers, nItems, k);
    }

    /** ML-100K-shaped demo next to the stand-in's own main(): GPU deterministic mode vs the sequential Java path. */
    public static void main(String[] args) {
        final long seed = 20261018L;
        final int nUsers = 943, nItems = 1682, total = 100_000, k = 32, epochs = 20;
        final float lr = 0.01f, lambda = 0.05f;
        int[] tu = new int[total], ti = new int[total];
        float[] tr = new float[total];
        int nt = 0;
        int[] u = new int[1], i = new int[1];
        float[] r = new float[1];
        for (long n = 0; n < total; n++) {
            if (!MatrixFactorizationSGD.syntheticRecord(seed, n, nUsers, nItems, 2, 0.25, 3, 0.375, u, i, r)) {
                tu[nt] = u[0]; ti[nt] = i[0]; tr[nt] = r[0]; nt++;
            }
        }
        tu = java.util.Arrays.copyOf(tu, nt); ti = java.util.Arrays.copyOf(ti, nt); tr = java.util.Arrays.copyOf(tr, nt);
        MatrixFactorizationSGD.Factors cpu = MatrixFactorizationSGD.factorize(tu, ti, tr, nUsers, nItems, k, lr, lambda, epochs, seed);
        MatrixFactorizationSGD.Factors gpu = factorize(tu, ti, tr, nUsers, nItems, k, lr, lambda, epochs, seed, MODE_DETERMINISTIC, 1);
        double worst = 0.0;
        for (int j = 0; j < cpu.P.length; j++) worst = Math.max(worst, Math.abs(cpu.P[j] - gpu.P[j]));
        for (int j = 0; j < cpu.Q.length; j++) worst = Math.max(worst, Math.abs(cpu.Q[j] - gpu.Q[j]));
        System.out.printf("max |cpu - gpu| over P and Q after %d epochs: %.3e%n", epochs, worst);
        System.out.printf("train RMSE cpu %.6f gpu %.6f%n",
                MatrixFactorizationSGD.rmse(cpu.P, cpu.Q, k, tu, ti, tr), rmse(gpu.P, gpu.Q, k, tu, ti, tr));
    }
}
