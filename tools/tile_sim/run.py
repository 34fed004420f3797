"""Driver of tools/tile_sim/tile_sim.cpp (design study for DESIGN.md section 8.1, CPU only): held-out RMSE per epoch of the
simulated user-tile kernel against the sequential oracle, for a merge rule and a machine size.
usage: python tools/tile_sim/run.py <midsize|ml20m|netflix> [U=256] [ctas=148] [warps=16] [chunk=64] [rule=expected|dynamic|one|store|sqrt] [epochs] [rounds=1]
Appends one JSON line per run to profiles/r01_tile_sim.jsonl."""
import ctypes as C, json, os, subprocess, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as orc
LIB = os.path.join(HERE, "libtilesim.so")
if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "tile_sim.cpp")):
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, os.path.join(HERE, "tile_sim.cpp")])
lib = C.CDLL(LIB)
SEED = 20261018
name = sys.argv[1]
U = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ctas = int(sys.argv[3]) if len(sys.argv) > 3 else 148
warps = int(sys.argv[4]) if len(sys.argv) > 4 else 16
chunk = int(sys.argv[5]) if len(sys.argv) > 5 else 64
rule = sys.argv[6] if len(sys.argv) > 6 else "expected"
if name == "midsize":      # the mid-size set of tests/test_gpu_parity.py
    nu, ni, n, k, lr, lam, epochs = 13_800, 2_700, 2_000_000, 32, 0.005, 0.05, 8
    gen = (2, 0.25, 3, 0.375)
    ref_curve = None
else:
    wl = {}
    exec(open(os.path.join(ROOT, "matrixfactorizationsgd.java_b200", "workloads.py")).read(), wl)
    w = wl["WORKLOADS"][name]
    nu, ni, n, k, lr, lam, epochs = w.n_users, w.n_items, w.n_ratings, w.k, w.lr, w.lambda_, w.epochs
    gen = (w.log2_alpha_user, w.c_user, w.log2_alpha_item, w.c_item)
    ref_curve = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_rmse_%s.json" % name)))["heldout_rmse_per_epoch"]
if len(sys.argv) > 7:
    epochs = int(sys.argv[7])
rounds = int(sys.argv[8]) if len(sys.argv) > 8 else 1     # interleaved passes: every tile is visited `rounds` times per epoch
us, its, rs, hus, his, hrs = [], [], [], [], [], []
for start in range(0, n, 25_000_000):
    u, i, r, h = orc.generate(SEED, start, min(25_000_000, n - start), nu, ni, *gen)
    us.append(u[~h]); its.append(i[~h]); rs.append(r[~h]); hus.append(u[h]); his.append(i[h]); hrs.append(r[h])
u, it, r = np.concatenate(us), np.concatenate(its), np.concatenate(rs)
hu, hi, hr = np.concatenate(hus), np.concatenate(his), np.concatenate(hrs)
if ref_curve is None:
    Po, Qo = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    ref_curve = []
    for ep in range(epochs):
        orc.train(u, it, r, Po, Qo, lr, lam, ep, ep + 1, SEED)
        ref_curve.append(orc.rmse(Po, Qo, hu, hi, hr))
# layout: records sorted by (tile, a per-tile pseudo-random item order) -- concurrent CTAs then sit on different items
n_user_tiles = int(u.max() // U) + 1
pass_of = (np.arange(len(u), dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) >> np.uint64(40)) % np.uint64(rounds)
tile = pass_of.astype(np.int64) * n_user_tiles + (u // U).astype(np.int64)      # a (pass, user tile) visit
n_tiles = rounds * n_user_tiles
item_key = ((it.astype(np.uint64) * np.uint64(2654435761) + tile.astype(np.uint64) * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF))
order = np.argsort((tile.astype(np.uint64) << np.uint64(32)) | item_key, kind="stable")
u, it, r, tile = u[order].copy(), it[order].copy(), r[order].copy(), tile[order]
tile_off = np.zeros(n_tiles + 1, np.int64)
np.cumsum(np.bincount(tile, minlength=n_tiles), out=tile_off[1:])
W = ctas * warps
share = np.bincount(it, minlength=ni) / float(len(it))
if rule == "expected":      # average over the runs expected to be in flight on the item: share of all in-flight warp time
    wts = (1.0 / np.maximum(1.0, share * W)).astype(np.float32)
elif rule == "sqrt":
    wts = (1.0 / np.sqrt(np.maximum(1.0, share * W))).astype(np.float32)
else:
    wts = np.ones(ni, np.float32)
P, Q = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
rng = np.random.default_rng(SEED)
curve, stats = [], np.zeros(2)
t0 = time.time()
for ep in range(epochs):
    tile_order = np.concatenate([p * n_user_tiles + rng.permutation(n_user_tiles) for p in range(rounds)]).astype(np.int32)
    lib.tilesim_epoch(C.c_void_p(u.ctypes.data), C.c_void_p(it.ctypes.data), C.c_void_p(r.ctypes.data), C.c_void_p(tile_off.ctypes.data),
                      C.c_int32(n_tiles), C.c_void_p(tile_order.ctypes.data), C.c_void_p(P.ctypes.data), C.c_void_p(Q.ctypes.data), C.c_int32(k),
                      C.c_float(lr), C.c_float(lam), C.c_int32(ctas), C.c_int32(warps), C.c_int32(chunk), C.c_double(8.0),
                      C.c_void_p(wts.ctypes.data), C.c_int32(1 if rule == "store" else (2 if rule == "dynamic" else 0)), C.c_void_p(stats.ctypes.data))
    curve.append(orc.rmse(P, Q, hu, hi, hr))
    print("epoch %d: sim %.5f oracle %.5f (%+.2f %%)  runs %.0f (mean %.2f ratings), runs in flight on the merged item %.2f  [%.0f s]"
          % (ep + 1, curve[-1], ref_curve[ep], 100 * (curve[-1] / ref_curve[ep] - 1), stats[0], len(r) / max(stats[0], 1), stats[1],
             time.time() - t0), flush=True)
out = {"workload": name, "users_per_tile": U, "ctas": ctas, "warps_per_cta": warps, "chunk": chunk, "merge_rule": rule, "rounds": rounds, "k": k,
       "sim_rmse_per_epoch": curve, "oracle_rmse_per_epoch": ref_curve[:epochs], "rel_diff": [c / o - 1 for c, o in zip(curve, ref_curve)],
       "mean_run": len(r) / max(stats[0], 1), "mean_runs_in_flight_on_item": stats[1]}
with open(os.path.join(ROOT, "profiles", "r01_tile_sim.jsonl"), "a") as f:
    f.write(json.dumps(out) + "\n")
