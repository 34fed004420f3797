"""What a user-tile kernel would move through L2 per update: for tiles of U users (P rows resident in
shared memory), every (tile, item) run costs one q_i read + merge (1 KB at k = 128), the tile's rows are read and written once
per visit. Counts the runs of 40 random tiles of the Netflix-shaped training set (CPU oracle generator).
usage: python tools/tile_traffic_estimate.py [workload]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as orc
wl = {}
exec(open(os.path.join(ROOT, "matrixfactorizationsgd.java_b200", "workloads.py")).read(), wl)
w = wl["WORKLOADS"][sys.argv[1] if len(sys.argv) > 1 else "netflix"]
seed = wl["SEED"]
us, its = [], []
for start in range(0, w.n_ratings, 25_000_000):
    u, i, r, h = orc.generate(seed, start, min(25_000_000, w.n_ratings - start), w.n_users, w.n_items, w.log2_alpha_user, w.c_user,
                              w.log2_alpha_item, w.c_item)
    us.append(u[~h]); its.append(i[~h])
u = np.concatenate(us); it = np.concatenate(its)
row_bytes = 4 * w.k
rng = np.random.default_rng(0)
for U in (128, 256, 400, 800):
    tiles = rng.choice(w.n_users // U, 40, replace=False)
    tile_of = u // U
    n_r = n_runs = 0
    for t in tiles:
        cnt = np.bincount(it[tile_of == t], minlength=w.n_items)
        n_r += int(cnt.sum()); n_runs += int((cnt > 0).sum())
    print("U=%d users per tile (%d KB of rows): %.0f ratings and %.0f runs per tile, mean run %.2f -> L2 bytes per update: "
          "q_i %.0f + p_u %.1f + record ~40 (run kernel today: %d)" % (U, U * row_bytes // 1024, n_r / 40, n_runs / 40, n_r / n_runs,
                                                                      2.0 * row_bytes * n_runs / n_r, 2.0 * row_bytes * U * 40 / n_r, 2 * row_bytes + 64))
