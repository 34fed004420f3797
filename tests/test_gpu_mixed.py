"""GPU parity of the mixed-precision factor storage: rows of P kept as binary16, arithmetic in binary32, narrowing
with stochastic rounding from a counter hash of (seed, epoch, u, i, chunk) -- stand-in srWord :413, storeF16Sr :420,
sgdUpdateMixed :426, factorizeMixed :439; oracle.cpp orc_train_mixed. The random bits do not depend on the visiting order, so every
conflict-free schedule is held to the oracle bit for bit; Hogwild / DSGD to 0.5 % of the BINARY32 sequential oracle's held-out RMSE."""
import numpy as np
import pytest

import matrixfactorizationsgd.java_b200 as mf
from matrixfactorizationsgd.java_b200 import _capi as capi
import pyoracle as orc
from test_gpu_parity import SEED, assert_curve_parity, assert_ring_rmse_parity, median_run, split  # noqa: F401

pytestmark = pytest.mark.gpu
F16 = capi.STORAGE_F16


@pytest.mark.parametrize("k", [8, 32, 128])
def test_binary16_rows_init_set_get(k):
    nu, ni = 257, 129
    with mf.Engine(mf.make_config(nu, ni, k, 0.01, 0.05, seed=SEED, p_storage=F16)) as eng:
        eng.load_ratings(np.zeros(1, np.int32), np.zeros(1, np.int32), np.ones(1, np.float32))
        eng.init_factors()
        P, Q = eng.get_factors()
        assert np.array_equal(P, orc.widen(orc.init_factors_f16(nu, k, SEED, 0)))       # initFactors, rounded to nearest even once
        assert np.array_equal(Q, orc.init_factors(ni, k, SEED, 1))                      # Q stays binary32
        rng = np.random.default_rng(k)
        P1 = (rng.standard_normal((nu, k)) * 0.3).astype(np.float32)
        eng.set_factors(P1, Q)
        P2, Q2 = eng.get_factors()
        assert np.array_equal(P2, P1.astype(np.float16).astype(np.float32)) and np.array_equal(Q2, Q)
        # the RMSE kernel reads the binary16 rows
        u = rng.integers(0, nu, 5000).astype(np.int32)
        i = rng.integers(0, ni, 5000).astype(np.int32)
        r = (1 + 4 * rng.random(5000)).astype(np.float32)
        want = orc.rmse(P2, Q2, u, i, r)
        assert abs(eng.rmse(u, i, r) - want) / want < 1e-6


@pytest.mark.parametrize("k", [8, 32, 100, 128, 256])
def test_mixed_deterministic_mode_bit_exact(k):
    nu, ni, n = 300, 200, 6000
    u, i, r, _ = orc.generate(SEED + k, 0, n, nu, ni)
    got = mf.MatrixFactorizationSGD.factorizeMixed(u, i, r, nu, ni, k, 0.02, 0.03, 3, SEED, mode=capi.MODE_DETERMINISTIC)
    P16, Q = orc.init_factors_f16(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    orc.train_mixed(u, i, r, P16, Q, 0.02, 0.03, 0, 3, SEED, orc.ORDER_WARP_TREE)
    assert np.array_equal(got.P, orc.widen(P16)) and np.array_equal(got.Q, Q)
    # the binary32 run of the same rule is a different trajectory, a rounding step away
    Pf, Qf = orc.factorize(u, i, r, nu, ni, k, 0.02, 0.03, 3, SEED, orc.ORDER_WARP_TREE)
    assert not np.array_equal(got.Q, Qf) and np.abs(got.Q - Qf).max() < 0.02 * np.abs(Qf).max()   # (binary16 ulp at 1.0 is 1e-3)


@pytest.mark.parametrize("k", [8, 32, 64, 100, 128, 256, 512])
def test_mixed_hogwild_kernel_bit_exact_on_conflict_free_data(k):
    """Pairwise distinct users and items: the full-grid kernel (and a virtual ring of 4) must reproduce the oracle bit for bit --
    including every stochastic rounding decision, which depends on (seed, epoch, u, i, chunk) only."""
    n = 5003
    rng = np.random.default_rng(k)
    u = rng.permutation(n).astype(np.int32)
    i = rng.permutation(n).astype(np.int32)
    r = (1 + 4 * rng.random(n)).astype(np.float32)
    P16, Q = orc.init_factors_f16(n, k, SEED, 0), orc.init_factors(n, k, SEED, 1)
    orc.train_mixed(u, i, r, P16, Q, 0.02, 0.03, 0, 3, SEED, orc.ORDER_WARP_TREE_FMA)
    P = orc.widen(P16)
    for kw in (dict(stripes_per_gpu=1, shards_per_gpu=1), dict(stripes_per_gpu=3, shards_per_gpu=2),
               dict(flags=capi.FLAG_MATERIALIZE_SHUFFLE), dict(mode=capi.MODE_DSGD, n_gpus=4, stripes_per_gpu=2, flags=capi.FLAG_VIRTUAL_RING)):
        got = mf.MatrixFactorizationSGD.factorizeMixed(u, i, r, n, n, k, 0.02, 0.03, 3, SEED, **kw)
        assert np.array_equal(got.P, P) and np.array_equal(got.Q, Q), kw
    assert len(np.unique(P16)) > 100


@pytest.mark.parametrize("arith", ["store", "heavy"])
@pytest.mark.parametrize("k", [8, 32, 64, 100, 128, 256])
def test_mixed_run_kernel_exact_sequential_runs(k, arith):
    """Run path with binary16 P rows, one run per item: gathers widen, stores narrow with the update's random word; a heavy user's
    row moves by the binary16 difference between the narrowed new value and the value the update read (one f16x4 red; the oracle
    restates its two binary16 roundings, orc_f16_add_rn -- the difference is exact unless a small value more than doubles)."""
    n_hot, per_hot, n_cold = 5, 3000, 5003
    n = n_hot * per_hot + n_cold
    rng = np.random.default_rng(7)
    items = np.concatenate([np.repeat(np.arange(n_hot), per_hot), n_hot + np.arange(n_cold)]).astype(np.int32)
    i = items[rng.permutation(n)]
    u = rng.permutation(n).astype(np.int32)
    r = (1 + 4 * rng.random(n)).astype(np.float32)
    ni = n_hot + n_cold
    cfg = mf.make_config(n, ni, k, 0.01, 0.03, seed=SEED, mode=capi.MODE_HOGWILD, stripes_per_gpu=1, rounds=1, hot_chunk=4096,
                         flags=capi.FLAG_NO_SHUFFLE, p_atomic_threshold=1e-9 if arith == "heavy" else -1.0, p_storage=F16)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(u, i, r)
        assert eng.layout_info().n_hot_items == n_hot
        assert eng.layout_info().n_heavy_users == (n if arith == "heavy" else 0)
        ou, oi, orr, off = eng.records()
        eng.init_factors()
        eng.train(3)
        P, Q = eng.get_factors()
    P16, Qo = orc.init_factors_f16(n, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    hot = oi < n_hot
    with orc.tree_lanes(orc.run_lanes(k)):
        orc.train_mixed(ou[hot].copy(), oi[hot].copy(), orr[hot].copy(), P16, Qo, 0.01, 0.03, 0, 3, SEED, orc.ORDER_WARP_TREE_FMA, shuffled=False,
                        sr=2 if arith == "heavy" else 1)      # heavy rows: moved by the binary16 difference (two binary16 roundings)
    orc.train_mixed(ou[~hot].copy(), oi[~hot].copy(), orr[~hot].copy(), P16, Qo, 0.01, 0.03, 0, 3, SEED, orc.ORDER_WARP_TREE_FMA, shuffled=False,
                    sr=2 if arith == "heavy" else 1)
    assert np.array_equal(P, orc.widen(P16)) and np.array_equal(Q, Qo)


@pytest.fixture(scope="module")
def midsets():
    from test_gpu_parity import MidSet
    return {"default": MidSet(signal=False), "signal": MidSet(signal=True)}


@pytest.mark.parametrize("variant", ["default", "signal"])
def test_mixed_hogwild_rmse_parity(midsets, variant):
    """Hogwild with binary16 P rows against the BINARY32 sequential oracle: final held-out RMSE within 0.5 %, both sides, and the
    half-epoch lag bound on the way (tests/test_gpu_parity.py assert_curve_parity), median of three runs (median_run)."""
    m = midsets[variant]

    def run():
        cfg = mf.make_config(m.nu, m.ni, m.k, m.lr, m.lam, seed=SEED, mode=capi.MODE_HOGWILD, p_storage=F16)
        with mf.Engine(cfg) as eng:
            eng.load_ratings(*m.train)
            eng.load_heldout(*m.held)
            eng.init_factors()
            eng.set_eval_every_epoch(True)
            stats = eng.train(m.epochs)
            P, Q = eng.get_factors()
        assert abs(stats[-1].heldout_rmse - orc.rmse(P, Q, *m.held)) / stats[-1].heldout_rmse < 1e-6
        assert np.array_equal(P, P.astype(np.float16).astype(np.float32))                  # the rows really are binary16 values
        return stats[-1].heldout_rmse, [s.heldout_rmse for s in stats]

    _, curve = median_run(run)
    assert_curve_parity(curve, m.curve)


@pytest.mark.parametrize("variant", ["default", "signal"])
def test_mixed_dsgd_virtual_ring_rmse_parity(midsets, variant):
    m = midsets[variant]
    cfg = mf.make_config(m.nu, m.ni, m.k, m.lr, m.lam, seed=SEED, mode=capi.MODE_DSGD, n_gpus=4, stripes_per_gpu=2, shards_per_gpu=2,
                         flags=capi.FLAG_VIRTUAL_RING, p_storage=F16)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(*m.train)
        ub, ib = eng.bounds()
        eng.init_factors()
        eng.train(m.epochs, want_stats=False)
        got = eng.rmse(*m.held)
    assert_ring_rmse_parity(got, m.oracle_rmse, m.dsgd_oracle_rmse(ub[::2], ib[::2]))


def test_mixed_with_the_model_extension(midsets):
    """binary16 P rows under the extended model (biases stay binary32): signal-dominant set, 0.5 % of the extended model's oracle."""
    from test_gpu_model import ModelMidSet
    m = ModelMidSet(signal=True)
    cfg = mf.make_config(m.nu, m.ni, m.k, m.lr, m.lam, seed=SEED, mode=capi.MODE_HOGWILD, p_storage=F16,
                         model=capi.MODEL_GLOBAL_MEAN | capi.MODEL_BIASES)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(*m.train)
        eng.load_heldout(*m.held)
        eng.init_factors()
        eng.set_eval_every_epoch(True)
        stats = eng.train(m.epochs)
    assert_curve_parity([s.heldout_rmse for s in stats], m.curve)


def test_mixed_argument_errors():
    for bad in (dict(p_storage=2), dict(p_storage=-1), dict(p_storage=F16, scatter=capi.SCATTER_ATOMIC),
                dict(p_storage=F16, scatter=capi.SCATTER_ATOMIC_P), dict(p_storage=F16, flags=capi.FLAG_EXACT_ARITH)):
        with pytest.raises(mf.MfsgdError) as ei:
            mf.Engine(mf.make_config(10, 10, 8, 0.1, 0.1, **bad))
        assert ei.value.code == capi.E_INVALID_ARG, bad
    mf.Engine(mf.make_config(10, 10, 8, 0.1, 0.1, p_storage=F16, mode=capi.MODE_DETERMINISTIC, flags=capi.FLAG_EXACT_ARITH)).close()
    with pytest.raises(ValueError):
        mf.MatrixFactorizationSGD.factorizeMixed(np.zeros(1, np.int32), np.zeros(1, np.int32), np.ones(1, np.float32), 4, 4, 6, 0.1, 0.1, 1, SEED)
