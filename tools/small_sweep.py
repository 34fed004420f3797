"""Which engine knob moves the held-out RMSE of the SMALL parity sets (ML-100K-shaped signal variant; the mid-size signal set under
a decaying learning rate)? Hogwild is nondeterministic, so every variant is repeated; prints rel = got / sequential oracle - 1.
  python tools/small_sweep.py [reps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import matrixfactorizationsgd.java_b200 as mf   # noqa: E402
from matrixfactorizationsgd.java_b200 import _capi as capi   # noqa: E402
import pyoracle as orc   # noqa: E402

VARIANTS = [("base", {}, {}), ("div16", {}, {"MFSGD_SUBWARP_DIV": "16"}), ("div64", {}, {"MFSGD_SUBWARP_DIV": "64"}),
            ("rounds1", {"rounds": 1}, {}), ("rounds2", {"rounds": 2}, {}), ("rounds16", {"rounds": 16}, {}),
            ("chunk64", {"hot_chunk": 64}, {}), ("chunk1024", {"hot_chunk": 1024}, {}), ("boost1", {"merge_boost": 1.0}, {}),
            ("atomic_p", {"scatter": capi.SCATTER_ATOMIC_P}, {}), ("all_cold", {"hot_share": 0.9}, {}),
            ("div16_rounds16", {"rounds": 16}, {"MFSGD_SUBWARP_DIV": "16"})]


def with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def ml100k_case(kw):
    w = mf.WORKLOADS["ml100k_signal"]
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_rmse_ml100k_signal.json")))
    cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=mf.SEED, mode=capi.MODE_HOGWILD, **kw)
    with mf.Engine(cfg) as eng:
        eng.generate_synthetic(mf.synth_params_of(w))
        li = eng.layout_info()
        eng.init_factors()
        st = eng.train(w.epochs)
        got = eng.rmse_heldout()[0]
    return got / ref["heldout_rmse_per_epoch"][w.epochs - 1] - 1.0, (li.rounds, li.run_length, li.n_hot_items, li.n_heavy_users), float(np.median([s.epoch_ms for s in st]))


_sched = {}


def schedule_case(kw):
    from test_gpu_model import ModelMidSet, SEED
    if not _sched:
        m = ModelMidSet(signal=True)
        P, Q = orc.init_factors(m.nu, m.k, SEED, 0), orc.init_factors(m.ni, m.k, SEED, 1)
        bu, bi = np.zeros(m.nu, np.float32), np.zeros(m.ni, np.float32)
        _, curve = orc.train_early_stop(m.train[0], m.train[1], m.rc, m.held[0], m.held[1], m.hc, P, Q, bu, bi, 2 * m.lr, m.lam, 0.85, 0, 0.0,
                                        m.epochs, SEED)
        _sched.update(m=m, want=curve[-1])
    m = _sched["m"]
    cfg = mf.make_config(m.nu, m.ni, m.k, 2 * m.lr, m.lam, seed=SEED, mode=capi.MODE_HOGWILD, model=3, lr_decay=0.85, **kw)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(*m.train)
        eng.load_heldout(*m.held)
        li = eng.layout_info()
        eng.init_factors()
        st = eng.train(m.epochs)
        got = eng.rmse(*m.held)
    return got / _sched["want"] - 1.0, (li.rounds, li.run_length, li.n_hot_items, li.n_heavy_users), float(np.median([s.epoch_ms for s in st]))


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    for case_name, case in (("ml100k_signal", ml100k_case), ("midsize_signal_schedule", schedule_case)):
        for tag, kw, env in VARIANTS:
            try:
                res = [with_env(env, lambda: case(kw)) for _ in range(reps)]
            except Exception as e:       # a variant the engine refuses is a result too
                print(json.dumps({"case": case_name, "variant": tag, "error": str(e)[:200]}), flush=True)
                continue
            rels = [100 * r[0] for r in res]
            print(json.dumps({"case": case_name, "variant": tag, "layout": res[0][1], "epoch_ms": round(res[0][2], 3),
                              "rel_pct": [round(x, 3) for x in rels], "mean_pct": round(float(np.mean(rels)), 3)}), flush=True)


if __name__ == "__main__":
    main()
