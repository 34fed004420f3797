"""Per-epoch held-out RMSE of the sequential CPU oracle on a BASELINE workload -> tests/golden/.

The GPU box cannot afford minutes of single-thread CPU per test run, so the oracle's RMSE curve for the
big shapes is computed once here (python tools/oracle_reference_rmse.py netflix) and committed as a
fixture; tests/test_gpu_workloads.py and bench.py compare the GPU's held-out RMSE at equal epochs to it.
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import pyoracle as orc
from importlib import import_module

def main():
    name = sys.argv[1]
    wl = {}
    exec(open(os.path.join(ROOT, "matrixfactorizationsgd.java_b200", "workloads.py")).read(), wl)
    w = wl["WORKLOADS"][name]; seed = wl["SEED"]
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else w.epochs
    t0 = time.time()
    u, i, r, held = orc.generate(seed, 0, w.n_ratings, w.n_users, w.n_items, w.log2_alpha_user, w.c_user, w.log2_alpha_item, w.c_item)
    tu, ti, tr = u[~held].copy(), i[~held].copy(), r[~held].copy()
    hu, hi, hr = u[held].copy(), i[held].copy(), r[held].copy()
    del u, i, r
    P = orc.init_factors(w.n_users, w.k, seed, 0); Q = orc.init_factors(w.n_items, w.k, seed, 1)
    curve = []
    for ep in range(epochs):
        orc.train(tu, ti, tr, P, Q, w.lr, w.lambda_, ep, ep + 1, seed)
        curve.append(orc.rmse(P, Q, hu, hi, hr))
        print(name, "epoch", ep + 1, "heldout rmse %.6f" % curve[-1], "%.0fs" % (time.time() - t0), flush=True)
        out = {"workload": name, "n_train": int(len(tr)), "n_heldout": int(len(hr)), "k": w.k, "lr": w.lr, "lambda": w.lambda_,
               "seed": seed, "oracle": "sequential (ORDER_SEQ), oracle/oracle.cpp", "heldout_rmse_per_epoch": curve}
        json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_rmse_%s.json" % name), "w"), indent=1)

if __name__ == "__main__":
    main()
