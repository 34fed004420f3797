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
    # generated in chunks straight into the train / held-out arrays: the 2 B-record shape must stay inside host RAM
    step = 50_000_000
    tu = np.empty(w.n_ratings, np.int32); ti = np.empty(w.n_ratings, np.int32); tr = np.empty(w.n_ratings, np.float32)
    hus, his, hrs = [], [], []
    nt = 0
    for start in range(0, w.n_ratings, step):
        u, i, r, held = orc.generate(seed, start, min(step, w.n_ratings - start), w.n_users, w.n_items,
                                     w.log2_alpha_user, w.c_user, w.log2_alpha_item, w.c_item)
        keep = ~held
        m = int(keep.sum())
        tu[nt:nt + m] = u[keep]; ti[nt:nt + m] = i[keep]; tr[nt:nt + m] = r[keep]
        nt += m
        hus.append(u[held]); his.append(i[held]); hrs.append(r[held])
    tu, ti, tr = tu[:nt], ti[:nt], tr[:nt]
    hu, hi, hr = np.concatenate(hus), np.concatenate(his), np.concatenate(hrs)
    del u, i, r, held, hus, his, hrs
    print(name, "generated", nt, "train", len(hr), "held-out", "%.0fs" % (time.time() - t0), flush=True)
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
