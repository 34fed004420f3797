"""How far apart are two SEQUENTIAL executions of the reference rule that differ only in the visiting order (the shuffle seed)?
On small sets that spread is of the size of the 0.5 % parity bar, so a parallel execution is compared with the band of sequential
executions, not with one of them. Writes tests/golden/order_spread.json (CPU oracle only).
  python tools/order_spread.py [orders]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import pyoracle as orc   # noqa: E402
from matrixfactorizationsgd.java_b200 import workloads as W   # noqa: E402   (pure data, no native library)

SEED = W.SEED
ORDER_SEED0 = SEED + 1000      # visiting-order seeds ORDER_SEED0 + j; init stays on SEED


def split(u, i, r, held):
    return (u[~held].copy(), i[~held].copy(), r[~held].copy()), (u[held].copy(), i[held].copy(), r[held].copy())


def ml100k_signal(orders):
    w = W.WORKLOADS["ml100k_signal"]
    u, i, r, held = orc.generate(SEED, 0, w.n_ratings, w.n_users, w.n_items, amplitude=w.amplitude, noise_scale=w.noise_scale)
    tr, ho = split(u, i, r, held)
    out = []
    for j in range(orders):
        P, Q = orc.init_factors(w.n_users, w.k, SEED, 0), orc.init_factors(w.n_items, w.k, SEED, 1)
        orc.train(*tr, P, Q, w.lr, w.lambda_, 0, w.epochs, ORDER_SEED0 + j)
        out.append(orc.rmse(P, Q, *ho))
    return {"epochs": w.epochs, "final_rmse_per_order": out}


def midsize_schedule(orders, lr_scale=2.0, decay=0.85):
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_rmse_midsize.json")))
    par = fx["params"]["signal"]
    nu, ni, n, k = fx["n_users"], fx["n_items"], fx["n_ratings"], fx["k"]
    u, i, r, held = orc.generate(SEED, 0, n, nu, ni, amplitude=par["amplitude"], noise_scale=par["noise_scale"])
    tr, ho = split(u, i, r, held)
    mu = np.float32(orc.global_mean(tr[2]))
    rc, hc = (tr[2] - mu).astype(np.float32), (ho[2] - mu).astype(np.float32)
    out = []
    for j in range(orders):
        P, Q = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
        bu, bi = np.zeros(nu, np.float32), np.zeros(ni, np.float32)
        # the stand-in's loop draws its order from (seed, epoch); a different order = a different seed for the loop only
        _, curve = orc.train_early_stop(tr[0], tr[1], rc, ho[0], ho[1], hc, P, Q, bu, bi, lr_scale * par["lr"], par["lam"], decay, 0, 0.0,
                                        par["epochs"], SEED if j == 0 else ORDER_SEED0 + j)
        out.append(curve[-1])
    return {"epochs": par["epochs"], "lr_scale": lr_scale, "decay": decay, "final_rmse_per_order": out,
            "note": "order 0 is the stand-in's own (seed = SEED)"}


def main():
    orders = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    t0 = time.time()
    res = {"seed": SEED, "order_seed0": ORDER_SEED0, "ml100k_signal": ml100k_signal(orders), "midsize_signal_schedule": midsize_schedule(orders)}
    for k in ("ml100k_signal", "midsize_signal_schedule"):
        v = np.array(res[k]["final_rmse_per_order"])
        res[k].update(min=float(v.min()), max=float(v.max()), mean=float(v.mean()))
        print(k, "spread %.3f %% of the mean" % (100 * (v.max() - v.min()) / v.mean()), np.round(v, 5))
    res["cpu_seconds"] = time.time() - t0
    with open(os.path.join(ROOT, "tests", "golden", "order_spread.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
