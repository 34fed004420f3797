"""Independent NumPy restatement of the stand-in's arithmetic (test infrastructure).

Third implementation of SURVEY.md section 8a rows a2-a6 (after the Java stand-in and the C++ oracle),
written from the written pins, not from oracle.cpp. Used by tests/golden/make_golden.py to produce
the committed fixtures and by tests/test_oracle.py to cross-check the oracle.
Cites baseline/java/MatrixFactorizationSGD.java line numbers.
"""
import numpy as np

U64 = np.uint64
GOLDEN = U64(0x9E3779B97F4A7C15)
STREAM_MUL = U64(0xD1B54A32D192ED03)
M1 = U64(0xBF58476D1CE4E5B9)
M2 = U64(0x94D049BB133111EB)
PLANTED_RANK = 16
PLANTED_AMPLITUDE = np.float32(0.8660254)
ID_MULT = 2654435761


def hash64(seed, stream, ctr):
    """MatrixFactorizationSGD.java:39 (vectorised over ctr)."""
    with np.errstate(over="ignore"):
        ctr = np.asarray(ctr, dtype=U64)
        z = U64(seed) + GOLDEN * (ctr + U64(1)) + STREAM_MUL * U64(stream)
        z = (z ^ (z >> U64(30))) * M1
        z = (z ^ (z >> U64(27))) * M2
        z = z ^ (z >> U64(31))
    return z


def uniform(seed, stream, ctr):
    """MatrixFactorizationSGD.java:48."""
    return (hash64(seed, stream, ctr) >> U64(40)).astype(np.float32) * np.float32(2.0 ** -24)


def default_init_scale(k):
    """MatrixFactorizationSGD.java:63."""
    return np.float32(1.0 / np.sqrt(np.float64(k)))


def init_factors(n_rows, k, seed, stream, scale):
    """MatrixFactorizationSGD.java:53."""
    ctr = np.arange(n_rows * k, dtype=U64)
    return (uniform(seed, stream, ctr) * np.float32(scale)).reshape(n_rows, k)


def shuffle(seed, epoch, n):
    """MatrixFactorizationSGD.java:72."""
    idx = np.arange(n, dtype=U64)
    key = hash64(seed, 2, (U64(epoch) << U64(32)) | idx) >> U64(33)
    packed = (key << U64(32)) | idx
    packed.sort()
    return (packed & U64(0xFFFFFFFF)).astype(np.int32)


def dot_seq(p, q):
    dot = np.float32(0.0)
    for f in range(len(p)):
        dot = np.float32(dot + np.float32(p[f] * q[f]))
    return dot


def dot_warp_tree(p, q, lanes=0):
    """DESIGN.md 4.2 summation order (not in the stand-in). lanes = 0: the cold / deterministic / RMSE kernels'
    min(32, pow2ceil(k/4)) lanes per rating; 8 / 16 / 32: the run kernel's geometry."""
    k = len(p)
    chunks = k // 4
    if lanes == 0:
        lanes = 1
        while lanes < chunks and lanes < 32:
            lanes *= 2
    s = np.zeros(lanes, dtype=np.float32)
    for l in range(lanes):
        acc = np.float32(0.0)
        for c in range(l, chunks, lanes):
            for j in range(4):
                acc = np.float32(acc + np.float32(p[4 * c + j] * q[4 * c + j]))
        s[l] = acc
    m = lanes // 2
    while m >= 1:
        s = (s + s[np.arange(lanes) ^ m]).astype(np.float32)
        m //= 2
    return s[0]


def sgd_update(p, q, r, lr, lam, tree=False):
    """MatrixFactorizationSGD.java:89. p, q: float32 views, updated in place. Returns e."""
    lr = np.float32(lr)
    lam = np.float32(lam)
    e = np.float32(np.float32(r) - (dot_warp_tree(p, q) if tree else dot_seq(p, q)))
    pf = p.copy()
    qf = q.copy()
    p[:] = pf + lr * (e * qf - lam * pf)   # elementwise float32 ops, one rounding each
    q[:] = qf + lr * (e * pf - lam * qf)
    return e


def factorize(u, i, r, n_users, n_items, k, lr, lam, epochs, seed, tree=False):
    """MatrixFactorizationSGD.java:109."""
    scale = default_init_scale(k)
    P = init_factors(n_users, k, seed, 0, scale)
    Q = init_factors(n_items, k, seed, 1, scale)
    for epoch in range(epochs):
        for t in shuffle(seed, epoch, len(r)):
            sgd_update(P[u[t]], Q[i[t]], r[t], lr, lam, tree)
    return P, Q


def rmse(P, Q, u, i, r):
    """MatrixFactorizationSGD.java:169."""
    sse = 0.0
    for t in range(len(r)):
        e = np.float32(np.float32(r[t]) - dot_seq(P[u[t]], Q[i[t]]))
        sse += float(e) * float(e)
    return float(np.sqrt(sse / len(r))) if len(r) else 0.0


def uniform53(seed, stream, ctr):
    """MatrixFactorizationSGD.java:191."""
    return (hash64(seed, stream, ctr) >> U64(11)).astype(np.float64) * (2.0 ** -53)


def skewed_rank(x, count, log2_alpha, c):
    """MatrixFactorizationSGD.java:199."""
    y = c + (1.0 - c) * np.asarray(x, dtype=np.float64)
    ca = np.float64(c)
    for _ in range(log2_alpha):
        y = y * y
        ca = ca * ca
    t = (y - ca) / (1.0 - ca)
    rank = np.floor(np.float64(count) * t).astype(np.int64)
    return np.clip(rank, 0, count - 1)


def scatter_id(rank, count):
    """MatrixFactorizationSGD.java:211 (python ints: no overflow)."""
    return np.array([(int(x) * ID_MULT + count // 2) % count for x in np.atleast_1d(rank)], dtype=np.int32)


def planted(seed, stream, rows, amplitude=PLANTED_AMPLITUDE):
    """MatrixFactorizationSGD.java:215; returns [len(rows), 16]."""
    rows = np.asarray(rows, dtype=U64)
    ctr = rows[:, None] * U64(PLANTED_RANK) + np.arange(PLANTED_RANK, dtype=U64)[None, :]
    return (uniform(seed, stream, ctr) - np.float32(0.5)) * np.float32(amplitude)


def generate(seed, start, count, n_users, n_items, l2au, cu, l2ai, ci, amplitude=PLANTED_AMPLITUDE, noise_scale=0.5):
    """MatrixFactorizationSGD.java:220; returns u, i, r, held. (amplitude, noise_scale) = (1.7320508, 0.125) is the
    signal-dominant variant of SURVEY.md 8d."""
    n = np.arange(start, start + count, dtype=U64)
    u = scatter_id(skewed_rank(uniform53(seed, 3, n), n_users, l2au, cu), n_users)
    i = scatter_id(skewed_rank(uniform53(seed, 4, n), n_items, l2ai, ci), n_items)
    ps = planted(seed, 7, u, amplitude)
    qs = planted(seed, 8, i, amplitude)
    dot = np.zeros(count, dtype=np.float32)
    for f in range(PLANTED_RANK):
        dot = (dot + (ps[:, f] * qs[:, f]).astype(np.float32)).astype(np.float32)
    noise = np.zeros(count, dtype=np.float32)
    for j in range(4):
        noise = (noise + uniform(seed, 5, U64(4) * n + U64(j))).astype(np.float32)
    noise = noise - np.float32(2.0)
    rating = np.float32(3.5) + dot
    rating = rating + np.float32(noise_scale) * noise
    rating = np.minimum(np.maximum(rating, np.float32(1.0)), np.float32(5.0)).astype(np.float32)
    held = (hash64(seed, 6, n) % U64(10)) == U64(0)
    return u, i, rating, held
