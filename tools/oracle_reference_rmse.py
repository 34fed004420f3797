"""Per-epoch held-out RMSE of the sequential CPU oracle on a workload -> tests/golden/oracle_rmse_<name>.json.

The GPU box cannot afford minutes of single-thread CPU per test run, so the oracle's curves for the big shapes are
computed once here and committed as fixtures; tests/test_gpu_workloads.py and bench.py compare the GPU's held-out
RMSE at equal epochs to them.

  python tools/oracle_reference_rmse.py netflix_signal [epochs] [--dsgd 2,4,8]

Every fixture carries the constant predictor's RMSE (predicting the training mean) beside the curve, so a reader can see
how much of the distance to it the training covers. --dsgd G adds, under "dsgd<G>", the curve of the SAME sequential rule
walking the ratings in the DSGD schedule's order (G x G strata by rating-count-balanced bounds, sub-epoch after
sub-epoch, member after member, the stand-in's shuffled order inside a block): the part of a ring's deviation from the
shuffled oracle that belongs to the schedule itself, not to the GPU's parallelism.
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import pyoracle as orc


def load_workload(name):
    wl = {}
    exec(open(os.path.join(ROOT, "matrixfactorizationsgd.java_b200", "workloads.py")).read(), wl)
    return wl["WORKLOADS"][name], wl["SEED"]


def generate_split(w, seed):
    # generated in chunks straight into the train / held-out arrays: the 2 B-record shape must stay inside host RAM
    step = 50_000_000
    tu = np.empty(w.n_ratings, np.int32); ti = np.empty(w.n_ratings, np.int32); tr = np.empty(w.n_ratings, np.float32)
    hus, his, hrs = [], [], []
    nt = 0
    for start in range(0, w.n_ratings, step):
        u, i, r, held = orc.generate(seed, start, min(step, w.n_ratings - start), w.n_users, w.n_items,
                                     w.log2_alpha_user, w.c_user, w.log2_alpha_item, w.c_item, amplitude=w.amplitude,
                                     noise_scale=w.noise_scale)
        keep = ~held
        m = int(keep.sum())
        tu[nt:nt + m] = u[keep]; ti[nt:nt + m] = i[keep]; tr[nt:nt + m] = r[keep]
        nt += m
        hus.append(u[held]); his.append(i[held]); hrs.append(r[held])
    return (tu[:nt], ti[:nt], tr[:nt]), (np.concatenate(hus), np.concatenate(his), np.concatenate(hrs))


def main():
    name = sys.argv[1]
    args = sys.argv[2:]
    dsgd = []
    if "--dsgd" in args:
        j = args.index("--dsgd")
        dsgd = [int(x) for x in args[j + 1].split(",")]
        args = args[:j] + args[j + 2:]
    w, seed = load_workload(name)
    epochs = int(args[0]) if args else w.epochs
    path = os.path.join(ROOT, "tests", "golden", "oracle_rmse_%s.json" % name)
    t0 = time.time()
    (tu, ti, tr), (hu, hi, hr) = generate_split(w, seed)
    print(name, "generated", len(tr), "train", len(hr), "held-out", "%.0fs" % (time.time() - t0), flush=True)
    mean = float(tr.astype(np.float64).mean())
    out = {"workload": name, "n_train": int(len(tr)), "n_heldout": int(len(hr)), "k": w.k, "lr": w.lr, "lambda": w.lambda_,
           "seed": seed, "amplitude": w.amplitude, "noise_scale": w.noise_scale,
           "oracle": "sequential (ORDER_SEQ), oracle/oracle.cpp", "train_mean": mean,
           "constant_predictor_rmse": float(np.sqrt(np.mean((hr.astype(np.float64) - mean) ** 2)))}
    if os.path.exists(path):                   # keep what an earlier invocation computed
        old = json.load(open(path))
        out.update({k: v for k, v in old.items() if k.startswith("dsgd") or k == "heldout_rmse_per_epoch"})
    if "heldout_rmse_per_epoch" not in out or len(out["heldout_rmse_per_epoch"]) < epochs:
        P = orc.init_factors(w.n_users, w.k, seed, 0); Q = orc.init_factors(w.n_items, w.k, seed, 1)
        curve = []
        for ep in range(epochs):
            orc.train(tu, ti, tr, P, Q, w.lr, w.lambda_, ep, ep + 1, seed)
            curve.append(orc.rmse(P, Q, hu, hi, hr))
            print(name, "epoch", ep + 1, "heldout rmse %.6f" % curve[-1], "%.0fs" % (time.time() - t0), flush=True)
            out["heldout_rmse_per_epoch"] = curve
            json.dump(out, open(path, "w"), indent=1)
    for G in dsgd:
        ub, ib = orc.balanced_bounds(tu, w.n_users, G), orc.balanced_bounds(ti, w.n_items, G)
        P = orc.init_factors(w.n_users, w.k, seed, 0); Q = orc.init_factors(w.n_items, w.k, seed, 1)
        curve = []
        for ep in range(epochs):
            order = orc.dsgd_order(tu, ti, ub, ib, seed, ep)
            orc.train(tu[order], ti[order], tr[order], P, Q, w.lr, w.lambda_, ep, ep + 1, seed, shuffled=False)
            curve.append(orc.rmse(P, Q, hu, hi, hr))
            print(name, "dsgd%d" % G, "epoch", ep + 1, "heldout rmse %.6f" % curve[-1], "%.0fs" % (time.time() - t0), flush=True)
            out["dsgd%d" % G] = {"user_bounds": ub.tolist(), "item_bounds": ib.tolist(), "heldout_rmse_per_epoch": curve,
                                 "oracle": "sequential rule in DSGD block order (pyoracle.dsgd_order)"}
            json.dump(out, open(path, "w"), indent=1)
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
