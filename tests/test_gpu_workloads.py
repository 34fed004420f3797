"""Full-size BASELINE.json workloads on the GPU against the committed oracle RMSE curves
(tests/golden/oracle_rmse_*.json, produced by tools/oracle_reference_rmse.py with the sequential CPU oracle:
minutes of CPU per curve, so computed once). Same synthetic data (generated on the device, bit-identical to the
oracle's generator -- tests/test_gpu_parity.py::test_generator_bit_exact), same init, same epoch count."""
import json
import os

import numpy as np
import pytest

import matrixfactorizationsgd.java_b200 as mf
from matrixfactorizationsgd.java_b200 import _capi as capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RMSE_TOL = 0.005   # north_star: within 0.5 % of the reference's RMSE after the same epoch count


def oracle_curve(name):
    return json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_rmse_%s.json" % name)))


@pytest.mark.parametrize("name", ["ml20m", "netflix"])
def test_hogwild_reaches_oracle_rmse_at_equal_epochs(name):
    w = mf.WORKLOADS[name]
    ref = oracle_curve(name)
    cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=mf.SEED, mode=capi.MODE_HOGWILD)
    with mf.Engine(cfg) as eng:
        nt, nh = eng.generate_synthetic(mf.synth_params_of(w))
        assert (nt, nh) == (ref["n_train"], ref["n_heldout"])           # identical split
        eng.init_factors()
        eng.set_eval_every_epoch(True)
        stats = eng.train(w.epochs)
    curve = [s.heldout_rmse for s in stats]
    want = ref["heldout_rmse_per_epoch"]
    assert len(want) == w.epochs
    assert curve[-1] <= want[-1] * (1 + RMSE_TOL), (curve, want)
    assert curve[-1] >= want[-1] * (1 - 0.02), (curve, want)
    # and it gets there at a comparable pace: from the 3rd epoch on never more than 1 % behind the oracle
    for e in range(2, w.epochs):
        assert curve[e] <= want[e] * 1.01, (e, curve[e], want[e])


def test_ml100k_shaped_hogwild_close_to_oracle():
    """configs[0] (the reference's own CPU-sized case) through the full-grid Hogwild path. 90 K ratings are far fewer than
    the ratings a B200 keeps in flight, so parallel SGD trails the sequential oracle in the first epochs; after the
    workload's 20 epochs it is within 1 % (measured +0.4 %). The 0.5 % bar of the north star is held on the ML-20M- and
    Netflix-shaped workloads above; bit-level parity on this config is the deterministic mode's job (test_gpu_parity.py)."""
    w = mf.WORKLOADS["ml100k"]
    ref = oracle_curve("ml100k")
    cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=mf.SEED, mode=capi.MODE_HOGWILD)
    with mf.Engine(cfg) as eng:
        nt, nh = eng.generate_synthetic(mf.synth_params_of(w))
        assert (nt, nh) == (ref["n_train"], ref["n_heldout"])
        eng.init_factors()
        eng.train(w.epochs, want_stats=False)
        got = eng.rmse_heldout()[0]
    want = ref["heldout_rmse_per_epoch"][-1]
    assert np.isfinite(got) and got <= want * 1.01 and got >= want * 0.98, (got, want)


def test_dsgd_virtual_ring_netflix_shaped_reaches_oracle_rmse():
    """The 8-member DSGD schedule (virtual ring on one GPU) on the full Netflix-shaped workload."""
    w = mf.WORKLOADS["netflix"]
    ref = oracle_curve("netflix")
    cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=mf.SEED, mode=capi.MODE_DSGD, n_gpus=8,
                         flags=capi.FLAG_VIRTUAL_RING)
    with mf.Engine(cfg) as eng:
        eng.generate_synthetic(mf.synth_params_of(w))
        eng.init_factors()
        eng.train(w.epochs, want_stats=False)
        got = eng.rmse_heldout()[0]
    want = ref["heldout_rmse_per_epoch"][-1]
    assert got <= want * (1 + RMSE_TOL) and got >= want * (1 - 0.02), (got, want)


@pytest.mark.parametrize("name", ["yahoo", "powerlaw"])
def test_dsgd_virtual_ring_large_shapes_reach_oracle_rmse(name):
    """configs[3] and [4] at full size (700 M / 2 B ratings, generated on the device) under the 8-member DSGD schedule,
    executed by one GPU (the shapes fit in 180 GB): held-out RMSE within 0.5 % of the sequential oracle at equal epochs.
    The oracle curves took 70 / 120 CPU-minutes (tools/oracle_reference_rmse.py) and are committed fixtures."""
    path = os.path.join(ROOT, "tests", "golden", "oracle_rmse_%s.json" % name)
    if not os.path.exists(path):
        pytest.skip("no committed oracle curve for %s yet" % name)
    w = mf.WORKLOADS[name]
    ref = oracle_curve(name)
    if len(ref["heldout_rmse_per_epoch"]) < w.epochs:
        pytest.skip("oracle curve for %s is still incomplete" % name)
    cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=mf.SEED, mode=capi.MODE_DSGD, n_gpus=8,
                         flags=capi.FLAG_VIRTUAL_RING)
    with mf.Engine(cfg) as eng:
        nt, nh = eng.generate_synthetic(mf.synth_params_of(w))
        assert (nt, nh) == (ref["n_train"], ref["n_heldout"])
        eng.init_factors()
        eng.train(w.epochs, want_stats=False)
        got = eng.rmse_heldout()[0]
    want = ref["heldout_rmse_per_epoch"][w.epochs - 1]
    assert got <= want * (1 + RMSE_TOL) and got >= want * (1 - 0.02), (got, want)


def test_heavy_skew_does_not_collapse_throughput():
    """config 5's shape (top item ~8 % of the ratings) scaled to one GPU: the hot-item path must keep the
    update rate within 2x of the uniform-item rate (the plain kernel drops 25x)."""
    nu, ni, n, k = 480_000, 17_800, 40_000_000, 64
    rates = {}
    for label, l2ai, ci in (("uniform", 0, 0.0), ("heavy", 4, 0.375)):
        cfg = mf.make_config(nu, ni, k, 0.005, 0.05, seed=mf.SEED, mode=capi.MODE_HOGWILD)
        with mf.Engine(cfg) as eng:
            eng.generate_synthetic(mf.synth_params(n, mf.SEED, 2, 0.25, l2ai, ci))
            eng.init_factors()
            eng.train(1, want_stats=False)
            st = eng.train(3)
            rm = eng.rmse_heldout()[0]
        rates[label] = st[0].updates / (np.median([s.epoch_ms for s in st]) * 1e-3)
        assert np.isfinite(rm) and rm < 1.0
    assert rates["heavy"] > 0.5 * rates["uniform"], rates
