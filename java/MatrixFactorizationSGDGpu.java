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
                MethodHandle setFactors = down("mfsgd_set_factors", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
                check((int) setFactors.invokeExact(h, arena.allocateFrom(JAVA_FLOAT, P), arena.allocateFrom(JAVA_FLOAT, Q)));
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
