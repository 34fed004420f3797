"""Harness for the DRAFT user-tile kernel (tools/tile_kernel_draft/tile_kernel.cu). Needs a B200.
    python tools/tile_kernel_draft/check.py exact            # one CTA, one warp == tools/tile_sim with GPU arithmetic, bit for bit
    python tools/tile_kernel_draft/check.py netflix [rounds=2] [epochs]   # full grid: ms per epoch, held-out RMSE vs the oracle curve
Never run in round 1 (no GPU minutes were left); the layout is built on the host here, a real integration builds it on the device."""
import ctypes as C, json, os, subprocess, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as orc
SEED, U, K, CHUNK, WARPS = 20261018, 256, 128, 64, 16


def build(lib, src, cmd):
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.check_call(cmd + ["-o", lib, src])
    return C.CDLL(lib)


kern = build(os.path.join(HERE, "libtiledraft.so"), os.path.join(HERE, "tile_kernel.cu"),
             ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"])
sim = build(os.path.join(ROOT, "tools", "tile_sim", "libtilesim.so"), os.path.join(ROOT, "tools", "tile_sim", "tile_sim.cpp"),
            ["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC"])


def layout(u, it, r, n_users, rounds):
    """records sorted by (pass, user tile, per-tile item order); one visit per (pass, tile)"""
    n_tiles = (n_users + U - 1) // U
    pass_of = (np.arange(len(u), dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) >> np.uint64(40)) % np.uint64(rounds)
    visit = pass_of.astype(np.int64) * n_tiles + (u // U).astype(np.int64)
    key = ((it.astype(np.uint64) * np.uint64(2654435761) + visit.astype(np.uint64) * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF))
    order = np.argsort((visit.astype(np.uint64) << np.uint64(32)) | key, kind="stable")
    u, it, r, visit = u[order], it[order], r[order], visit[order]
    n_visits = rounds * n_tiles
    off = np.zeros(n_visits + 1, np.int64)
    np.cumsum(np.bincount(visit, minlength=n_visits), out=off[1:])
    lo = (np.arange(n_visits) % n_tiles * U).astype(np.int32)
    nu = np.minimum(U, n_users - lo).astype(np.int32)
    recs = np.empty((len(u), 3), np.int32)
    recs[:, 0], recs[:, 1], recs[:, 2] = u, it, r.view(np.int32)
    return recs, np.ascontiguousarray(u), np.ascontiguousarray(it), np.ascontiguousarray(r), off, lo, nu, n_tiles


def visit_orders(rng, n_tiles, rounds, epochs):
    return np.concatenate([np.concatenate([p * n_tiles + rng.permutation(n_tiles) for p in range(rounds)]) for _ in range(epochs)]).astype(np.int32)


def gpu_train(recs, off, lo, nu, order, epochs, P, Q, wts, lr, lam, grid=0, active_warps=0):
    ms = C.c_double(0.0)
    rc = kern.tiledraft_train(C.c_void_p(recs.ctypes.data), C.c_int64(len(recs)), C.c_void_p(off.ctypes.data), C.c_void_p(lo.ctypes.data),
                              C.c_void_p(nu.ctypes.data), C.c_int32(len(lo)), C.c_void_p(order.ctypes.data), C.c_int32(epochs),
                              C.c_void_p(P.ctypes.data), C.c_int64(P.shape[0]), C.c_void_p(Q.ctypes.data), C.c_int64(Q.shape[0]),
                              C.c_void_p(wts.ctypes.data), C.c_float(lr), C.c_float(lam), C.c_int32(grid), C.c_int32(active_warps), C.byref(ms))
    assert rc == 0, "tile draft failed"
    return ms.value


def cpu_sim_epoch(u, it, r, off, order, P, Q, wts, lr, lam, ctas, warps):
    stats = np.zeros(2)
    sim.tilesim_epoch(C.c_void_p(u.ctypes.data), C.c_void_p(it.ctypes.data), C.c_void_p(r.ctypes.data), C.c_void_p(off.ctypes.data),
                      C.c_int32(len(off) - 1), C.c_void_p(order.ctypes.data), C.c_void_p(P.ctypes.data), C.c_void_p(Q.ctypes.data), C.c_int32(K),
                      C.c_float(lr), C.c_float(lam), C.c_int32(ctas), C.c_int32(warps), C.c_int32(CHUNK), C.c_double(8.0),
                      C.c_void_p(wts.ctypes.data), C.c_int32(0), C.c_void_p(stats.ctypes.data))


mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
if mode == "exact":
    nu_, ni_, n = 1000, 300, 50_000
    u, it, r, h = orc.generate(SEED, 0, n, nu_, ni_)
    u, it, r = u[~h].copy(), it[~h].copy(), r[~h].copy()
    for rounds in (1, 2):
        recs, su, si, sr, off, lo, nu, n_tiles = layout(u, it, r, nu_, rounds)
        wts = (0.25 + 0.75 * np.random.default_rng(1).random(ni_)).astype(np.float32)      # arbitrary weights: the merge arithmetic is checked too
        order = visit_orders(np.random.default_rng(2), n_tiles, rounds, 2)
        P, Q = orc.init_factors(nu_, K, SEED, 0), orc.init_factors(ni_, K, SEED, 1)
        Ps, Qs = P.copy(), Q.copy()
        gpu_train(recs, off, lo, nu, order, 2, P, Q, wts, 0.01, 0.05, grid=1, active_warps=1)
        sim.tilesim_set_gpu_arithmetic(1)
        for ep in range(2):
            cpu_sim_epoch(su, si, sr, off, order[ep * len(lo):(ep + 1) * len(lo)].copy(), Ps, Qs, wts, 0.01, 0.05, 1, 1)
        sim.tilesim_set_gpu_arithmetic(0)
        print("rounds %d: P equal %s, Q equal %s (max |dP| %.3g, max |dQ| %.3g)" % (rounds, np.array_equal(P, Ps), np.array_equal(Q, Qs),
                                                                                 np.abs(P - Ps).max(), np.abs(Q - Qs).max()))
        assert np.array_equal(P, Ps) and np.array_equal(Q, Qs)
    print("exact: OK")
else:
    wl = {}
    exec(open(os.path.join(ROOT, "matrixfactorizationsgd.java_b200", "workloads.py")).read(), wl)
    w = wl["WORKLOADS"][mode]
    assert w.k == K
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    epochs = int(sys.argv[3]) if len(sys.argv) > 3 else w.epochs
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_rmse_%s.json" % mode)))["heldout_rmse_per_epoch"]
    us, its, rs, hus, his, hrs = [], [], [], [], [], []
    for start in range(0, w.n_ratings, 25_000_000):
        u, i, r, h = orc.generate(SEED, start, min(25_000_000, w.n_ratings - start), w.n_users, w.n_items, w.log2_alpha_user, w.c_user,
                                  w.log2_alpha_item, w.c_item)
        us.append(u[~h]); its.append(i[~h]); rs.append(r[~h]); hus.append(u[h]); his.append(i[h]); hrs.append(r[h])
    u, it, r = np.concatenate(us), np.concatenate(its), np.concatenate(rs)
    hu, hi, hr = np.concatenate(hus), np.concatenate(his), np.concatenate(hrs)
    t0 = time.time()
    recs, su, si, sr, off, lo, nu, n_tiles = layout(u, it, r, w.n_users, rounds)
    print("host layout %.1f s, %d visits" % (time.time() - t0, len(lo)), flush=True)
    sms = 148
    share = np.bincount(it, minlength=w.n_items) / float(len(it))
    wts = (1.0 / np.maximum(1.0, share * sms * WARPS)).astype(np.float32)
    P, Q = orc.init_factors(w.n_users, K, SEED, 0), orc.init_factors(w.n_items, K, SEED, 1)
    rng = np.random.default_rng(SEED)
    for ep in range(epochs):
        ms = gpu_train(recs, off, lo, nu, visit_orders(rng, n_tiles, rounds, 1), 1, P, Q, wts, w.lr, w.lambda_)
        rm = orc.rmse(P, Q, hu, hi, hr)
        print("epoch %d: %.2f ms = %.2f G updates/s, held-out RMSE %.5f (oracle %.5f, %+.2f %%)"
              % (ep + 1, ms, len(r) / ms / 1e6, rm, ref[min(ep, len(ref) - 1)], 100 * (rm / ref[min(ep, len(ref) - 1)] - 1)), flush=True)
