"""Oracle curves of the mid-size parity sets (tests/test_gpu_parity.py MidSet, tests/test_gpu_model.py ModelMidSet) as a committed
fixture, so the GPU box does not spend minutes of CPU on them: tests/golden/oracle_rmse_midsize.json.
Per data variant (default / signal) and model (plain / extended): held-out RMSE after every epoch for the stand-in's shuffled
order and for the DSGD schedule's own block order (pyoracle.dsgd_order) with the rating-count-balanced strata of G = 2, 4, 8.
  python tools/midsize_oracle_curves.py          (about 5 minutes on 8 cores)"""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import pyoracle as orc   # noqa: E402

SEED = 20261018
NU, NI, N, K = 13_800, 2_700, 2_000_000, 32
SIGNAL_AMPLITUDE, SIGNAL_NOISE_SCALE = 1.7320508, 0.125       # workloads.py (not imported: it would load libmfsgd.so's package)
PARAMS = {"default": dict(lr=0.005, lam=0.05, epochs=8, amplitude=0.0, noise_scale=0.0),
          "signal": dict(lr=0.02, lam=0.02, epochs=20, amplitude=SIGNAL_AMPLITUDE, noise_scale=SIGNAL_NOISE_SCALE)}


def data(variant):
    p = PARAMS[variant]
    u, i, r, held = orc.generate(SEED, 0, N, NU, NI, amplitude=p["amplitude"], noise_scale=p["noise_scale"])
    tr = (u[~held].copy(), i[~held].copy(), r[~held].copy())
    ho = (u[held].copy(), i[held].copy(), r[held].copy())
    return p, tr, ho


def curve(job):
    variant, model, G = job
    p, (tu, ti, tr), (hu, hi, hr) = data(variant)
    mu = np.float32(orc.global_mean(tr)) if model else np.float32(0.0)
    rc, hc = (tr - mu).astype(np.float32), (hr - mu).astype(np.float32)
    P, Q = orc.init_factors(NU, K, SEED, 0), orc.init_factors(NI, K, SEED, 1)
    bu, bi = (np.zeros(NU, np.float32), np.zeros(NI, np.float32)) if model else (None, None)
    ub = orc.balanced_bounds(tu, NU, G) if G else None
    ib = orc.balanced_bounds(ti, NI, G) if G else None
    out = []
    for e in range(p["epochs"]):
        if G:
            o = orc.dsgd_order(tu, ti, ub, ib, SEED, e)
            orc.train_model(tu[o], ti[o], rc[o], P, Q, bu, bi, p["lr"], p["lam"], e, e + 1, SEED, shuffled=False)
        else:
            orc.train_model(tu, ti, rc, P, Q, bu, bi, p["lr"], p["lam"], e, e + 1, SEED)
        out.append(orc.rmse_model(P, Q, bu, bi, hu, hi, hc))
    return variant, model, G, out


def main():
    jobs = [(v, m, G) for v in PARAMS for m in (False, True) for G in (0, 2, 4, 8)]
    res = {"seed": SEED, "n_users": NU, "n_items": NI, "n_ratings": N, "k": K, "params": PARAMS,
           "generated_by": "tools/midsize_oracle_curves.py (oracle/oracle.cpp: orc_train_model with bu = bi = null is orc_train)"}
    for v in PARAMS:
        _, tr, ho = data(v)
        res[v] = {"n_train": int(len(tr[2])), "n_heldout": int(len(ho[2])),
                  "constant_predictor_rmse": float(np.sqrt(np.mean((ho[2] - tr[2].mean()) ** 2)))}
    with ProcessPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for v, m, G, c in ex.map(curve, jobs):
            res[v].setdefault("model" if m else "plain", {})["dsgd%d" % G if G else "shuffled"] = c
            print(v, "model" if m else "plain", G, [round(x, 4) for x in c[-3:]], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "oracle_rmse_midsize.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
