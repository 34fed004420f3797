"""GPU experiments (design aid): convergence vs blocking/concurrency, and throughput sweeps.
usage: python tools/gpu_experiments.py conv|sweep|curve [workload]   -> JSON lines on stdout"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import matrixfactorizationsgd.java_b200 as mf
capi = mf.capi
SEED = mf.SEED

def conv():
    import pyoracle as orc
    nu, ni, n, k, lr, lam, epochs = 13_800, 2_700, 2_000_000, 32, 0.005, 0.05, 8
    u, i, r, held = orc.generate(SEED, 0, n, nu, ni)
    tr = (u[~held].copy(), i[~held].copy(), r[~held].copy()); ho = (u[held].copy(), i[held].copy(), r[held].copy())
    P = orc.init_factors(nu, k, SEED, 0); Q = orc.init_factors(ni, k, SEED, 1)
    curve = []
    for ep in range(epochs):
        orc.train(*tr, P, Q, lr, lam, ep, ep + 1, SEED); curve.append(orc.rmse(P, Q, *ho))
    print(json.dumps({"oracle": curve}), flush=True)
    for mw in (128,):
        os.environ["MFSGD_MIN_WINDOWS"] = str(mw)
        for mode, G, mu, rounds in ((1, 1, 1, 1), (1, 1, 2, 1), (1, 1, 2, 4), (1, 1, 2, 16), (1, 1, 4, 1), (1, 1, 4, 4), (1, 1, 4, 16), (1, 1, 4, 0),
                                   (1, 1, 8, 1), (1, 1, 8, 4), (1, 1, 8, 16), (1, 1, 8, 0), (2, 4, 2, 0), (2, 4, 2, 4)):
            for rep in range(2):
                cfg = mf.make_config(nu, ni, k, lr, lam, seed=SEED + rep, mode=mode, n_gpus=G, stripes_per_gpu=mu, rounds=rounds,
                                     flags=capi.FLAG_VIRTUAL_RING if G > 1 else 0)
                with mf.Engine(cfg) as eng:
                    eng.load_ratings(*tr); eng.load_heldout(*ho)
                    eng.set_factors(orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1))
                    eng.set_eval_every_epoch(True)
                    st = eng.train(epochs)
                    rr = eng.layout_info().rounds
                c = [s.heldout_rmse for s in st]
                print(json.dumps({"min_windows": mw, "mode": mode, "G": G, "mu": mu, "rounds": rr, "rep": rep, "rel_end_pct": 100 * (c[-1] - curve[-1]) / curve[-1],
                                  "curve": c}), flush=True)

def sweep(wname):
    w = mf.WORKLOADS[wname]
    sp = mf.synth_params_of(w)
    for scatter in (0, 1, 2, 3):
        for stripes, rounds in ((1, 1), (7, 1), (7, 0), (7, 32), (14, 0)):
            cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=SEED, mode=capi.MODE_HOGWILD,
                                 stripes_per_gpu=stripes, scatter=scatter, rounds=rounds, flags=capi.FLAG_TIME_KERNELS)
            with mf.Engine(cfg) as eng:
                eng.generate_synthetic(sp); eng.init_factors(); eng.train(1)
                st = eng.train(3)
                rm = eng.rmse_heldout()[0]
                rr = eng.layout_info().rounds
            ms = np.median([s.epoch_ms for s in st]); kms = np.median([s.update_kernel_ms for s in st])
            print(json.dumps({"workload": wname, "scatter": scatter, "stripes": stripes, "rounds": rr, "epoch_ms": ms, "kernel_ms": kms,
                              "shuffle_ms": st[-1].shuffle_ms, "gupdates_s": st[0].updates / ms / 1e6, "rmse_after4": rm}), flush=True)

def curve(wname, stripes=0, rounds=0, scatter=0, hot=0.0, chunk=0):
    w = mf.WORKLOADS[wname]
    sp = mf.synth_params_of(w)
    cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=SEED, mode=capi.MODE_HOGWILD, stripes_per_gpu=stripes, rounds=rounds, scatter=scatter, hot_share=hot, hot_chunk=chunk)
    with mf.Engine(cfg) as eng:
        eng.generate_synthetic(sp); eng.init_factors(); eng.set_eval_every_epoch(True)
        st = eng.train(w.epochs)
        info = eng.layout_info()
    print(json.dumps({"workload": wname, "stripes": info.stripes_per_gpu, "rounds": info.rounds, "scatter": scatter, "n_hot": info.n_hot_items, "heldout_rmse_per_epoch": [s.heldout_rmse for s in st],
                      "epoch_ms": [s.epoch_ms for s in st]}), flush=True)

def skew(wname):
    w = mf.WORKLOADS[wname]
    for l2ai, ci in ((3, 0.375), (4, 0.375)):
        sp = mf.synth_params(w.n_ratings, SEED, w.log2_alpha_user, w.c_user, l2ai, ci)
        for scatter, hot, chunk, rounds in ((0, 0.0, 0, 1), (0, 0.0, 0, 0), (0, 0.0, 128, 0), (0, 0.0, 64, 0), (0, 0.0, 1024, 0), (0, 5e-5, 0, 0), (0, 1e-4, 0, 0), (0, 1e-4, 128, 0), (0, 0.0, 0, 4)):
            cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=SEED, mode=capi.MODE_HOGWILD,
                                 stripes_per_gpu=7, scatter=scatter, rounds=rounds, hot_share=hot, hot_chunk=chunk, flags=capi.FLAG_TIME_KERNELS)
            with mf.Engine(cfg) as eng:
                eng.generate_synthetic(sp); eng.init_factors(); eng.train(1)
                st = eng.train(3)
                info = eng.layout_info(); rm = eng.rmse_heldout()[0]
            ms = np.median([s.epoch_ms for s in st])
            print(json.dumps({"workload": wname, "l2ai": l2ai, "c_item": ci, "scatter": scatter, "hot_share": hot, "hot_chunk": chunk, "rounds": info.rounds, "n_hot": info.n_hot_items,
                              "epoch_ms": ms, "gupdates_s": st[0].updates / ms / 1e6, "rmse4": rm}), flush=True)

if __name__ == "__main__":
    what = sys.argv[1]
    if what == "conv": conv()
    elif what == "sweep": sweep(sys.argv[2])
    elif what == "skew": skew(sys.argv[2])
    elif what == "curve": curve(sys.argv[2], *[float(x) if "." in x or "e" in x else int(x) for x in sys.argv[3:]])


def smallblocks(wname):
    """Proxy for the per-GPU work of an 8-member ring on one GPU: 64 user sub-stripes, rounds 1 -> every launch
    holds 1/64 of the ratings, like one (member, sub-epoch) block at G = 8."""
    w = mf.WORKLOADS[wname]
    sp = mf.synth_params_of(w)
    for mw in (128, 32, 8):
        os.environ["MFSGD_MIN_WINDOWS"] = str(mw)
        for chunk in (256, 128):
            cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=SEED, mode=capi.MODE_HOGWILD,
                                 stripes_per_gpu=64, rounds=1, hot_chunk=chunk, flags=capi.FLAG_TIME_KERNELS)
            with mf.Engine(cfg) as eng:
                eng.generate_synthetic(sp); eng.init_factors(); eng.train(1)
                st = eng.train(3)
                rm = eng.rmse_heldout()[0]
            ms = np.median([s.epoch_ms for s in st])
            print(json.dumps({"min_windows": mw, "hot_chunk": chunk, "epoch_ms": ms, "per_launch_pair_us": 1e3 * ms / 64,
                              "gupdates_s": st[0].updates / ms / 1e6, "rmse4": rm}), flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "smallblocks":
    smallblocks(sys.argv[2])


def stripes_sweep(wname):
    w = mf.WORKLOADS[wname]
    sp = mf.synth_params_of(w)
    for stripes in (2, 3, 4, 5, 7):
        for rounds in (1, 2, 4):
            cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=SEED, mode=capi.MODE_HOGWILD,
                                 stripes_per_gpu=stripes, rounds=rounds)
            with mf.Engine(cfg) as eng:
                eng.generate_synthetic(sp); eng.init_factors(); eng.train(1, want_stats=False)
                st = eng.train(5)
                rm = eng.rmse_heldout()[0]
            ms = float(np.median([s.epoch_ms for s in st]))
            print(json.dumps({"workload": wname, "stripes": stripes, "rounds": rounds, "epoch_ms": ms, "gupdates_s": st[0].updates / ms / 1e6, "rmse6": rm}), flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "stripes":
    stripes_sweep(sys.argv[2])


def ringcurve(wname, G, chunk, shards=0):
    """Held-out RMSE per epoch of the G-member DSGD schedule (virtual ring on one GPU) for a given run length."""
    w = mf.WORKLOADS[wname]
    sp = mf.synth_params_of(w)
    cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=SEED, mode=capi.MODE_DSGD, n_gpus=G, hot_chunk=chunk,
                         shards_per_gpu=shards, flags=capi.FLAG_VIRTUAL_RING)
    with mf.Engine(cfg) as eng:
        eng.generate_synthetic(sp); eng.init_factors(); eng.set_eval_every_epoch(True)
        st = eng.train(w.epochs)
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_rmse_%s.json" % wname)))["heldout_rmse_per_epoch"]
    c = [s.heldout_rmse for s in st]
    print(json.dumps({"workload": wname, "G": G, "hot_chunk": chunk, "shards": shards, "rel_pct_vs_oracle": [round(100 * (a - b) / b, 3) for a, b in zip(c, ref)],
                      "final": c[-1]}), flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "ringcurve":
    ringcurve(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]) if len(sys.argv) > 5 else 0)
