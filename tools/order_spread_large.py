"""Order-to-order spread of the SEQUENTIAL reference rule on a full-size workload: the same training, the same initial factors,
only the visiting order differs (the shuffle seed). One order per invocation (minutes to an hour of one core each; run several
side by side), results collected into tests/golden/order_spread_large.json.
  python tools/order_spread_large.py netflix_signal 1        # order 1 (order 0 = the stand-in's own = the committed oracle curve)
  python tools/order_spread_large.py --collect"""
import glob
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
OUT = os.path.join(ROOT, "tests", "golden", "order_spread_large.json")
PART = os.environ.get("ORDER_SPREAD_DIR", "/tmp/order_spread")


def collect():
    res = {}
    for f in sorted(glob.glob(os.path.join(PART, "*.json"))):
        d = json.load(open(f))
        res.setdefault(d["workload"], {})[str(d["order"])] = d["heldout_rmse_per_epoch"]
    out = {"note": "sequential oracle (oracle.cpp, ORDER_SEQ), init seed 20261018, visiting-order seed 20262018 + order; order 0 = the "
                   "stand-in's own order = tests/golden/oracle_rmse_<workload>.json", "workloads": {}}
    for wname, orders in res.items():
        fx = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_rmse_%s.json" % wname)))
        epochs = min(len(c) for c in orders.values())
        curves = {"0": fx["heldout_rmse_per_epoch"][:epochs]}
        curves.update({k: v[:epochs] for k, v in orders.items()})
        final = np.array([c[epochs - 1] for c in curves.values()])
        out["workloads"][wname] = {"epochs": epochs, "heldout_rmse_per_epoch_per_order": curves, "final_min": float(final.min()),
                                   "final_max": float(final.max()), "final_spread_rel": float((final.max() - final.min()) / final.mean())}
        print(wname, "orders", sorted(curves), "final", np.round(final, 6), "spread %.4f %%" % (100 * (final.max() - final.min()) / final.mean()))
    json.dump(out, open(OUT, "w"), indent=1)


def main():
    if sys.argv[1] == "--collect":
        return collect()
    import pyoracle as orc
    from oracle_reference_rmse import generate_split, load_workload
    wname, order = sys.argv[1], int(sys.argv[2])
    epochs = int(sys.argv[3]) if len(sys.argv) > 3 else None
    w, seed = load_workload(wname)
    epochs = epochs or w.epochs
    os.makedirs(PART, exist_ok=True)
    t0 = time.time()
    (tu, ti, tr), (hu, hi, hr) = generate_split(w, seed)
    P, Q = orc.init_factors(w.n_users, w.k, seed, 0), orc.init_factors(w.n_items, w.k, seed, 1)
    curve = []
    for ep in range(epochs):
        orc.train(tu, ti, tr, P, Q, w.lr, w.lambda_, ep, ep + 1, seed + 1000 + order)      # the loop's seed only keys the order
        curve.append(orc.rmse(P, Q, hu, hi, hr))
        print(wname, "order", order, "epoch", ep + 1, "%.6f" % curve[-1], "%.0fs" % (time.time() - t0), flush=True)
        json.dump({"workload": wname, "order": order, "heldout_rmse_per_epoch": curve}, open(os.path.join(PART, "%s_%d.json" % (wname, order)), "w"))


if __name__ == "__main__":
    main()
