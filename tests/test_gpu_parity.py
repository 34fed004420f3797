"""GPU parity tests (run with -m gpu on a B200). Every call goes through the C ABI of libmfsgd.so
(ctypes, the same symbols FFM binds); the CPU oracle (oracle/) is only the checker.

Bars (BASELINE.json north_star):
  * deterministic single-warp mode vs the reference's per-update results: <= 1e-5 relative fp32
    (and bit-exact against the oracle's warp-tree summation order);
  * integer/index work (generator ids, shuffle order, bucketing) bit-exact;
  * Hogwild / DSGD held-out RMSE within 0.5 % of the oracle's at equal epochs on the same data.
"""
import numpy as np
import pytest

import matrixfactorizationsgd.java_b200 as mf
from matrixfactorizationsgd.java_b200 import _capi as capi
import pyoracle as orc

pytestmark = pytest.mark.gpu

SEED = 20261018
REL_TOL = 1e-5      # north_star: "within 1e-5 relative fp32"
RMSE_TOL = 0.005    # north_star: "held-out RMSE within 0.5 %"


def assert_rmse_parity(got, want):
    """Two-sided: held-out RMSE within 0.5 % of the sequential oracle's at equal epochs (north_star)."""
    assert abs(got / want - 1.0) <= RMSE_TOL, (got, want, got / want - 1.0)


def assert_ring_rmse_parity(got, shuffled, dsgd_ordered):
    """A DSGD ring is held between two sequential executions of the reference rule on the same data: the stand-in's
    shuffled order and the DSGD schedule's own block order (pyoracle.dsgd_order: same strata, same sub-epoch sequence,
    one rating at a time). On noise-dominant sets the two are < 0.3 % apart and this is the 0.5 % bar; on signal-dominant
    sets the schedule itself costs several per cent at a constant learning rate and the GPU's run path,
    which averages an item's concurrent runs, lands between the two. Each side gets the 0.5 % tolerance."""
    lo, hi = min(shuffled, dsgd_ordered), max(shuffled, dsgd_ordered)
    assert lo * (1 - RMSE_TOL) <= got <= hi * (1 + RMSE_TOL), (got, shuffled, dsgd_ordered)


def train_runs_then_cold(ou, oi, orr, n_hot, Po, Qo, k, lr, lam, e0, e1, order, heavy=False):
    """Oracle twin of a layout whose items < n_hot go through the run kernel (orc.run_lanes(k) lanes per rating) and
    the others through the cold kernel (default lanes). Valid when the two sets share no P or Q row.
    heavy: every user is marked heavy, i.e. both kernels add p_u's increment in memory (FMA arrangement: PDELTA rounding)."""
    hot = oi < n_hot
    run_order = orc.ORDER_WARP_TREE_FMA_PDELTA if heavy and order == orc.ORDER_WARP_TREE_FMA else order
    with orc.tree_lanes(orc.run_lanes(k)):
        orc.train(ou[hot].copy(), oi[hot].copy(), orr[hot].copy(), Po, Qo, lr, lam, e0, e1, SEED, run_order, shuffled=False)
    orc.train(ou[~hot].copy(), oi[~hot].copy(), orr[~hot].copy(), Po, Qo, lr, lam, e0, e1, SEED, run_order, shuffled=False)


def split(u, i, r, held):
    return (u[~held].copy(), i[~held].copy(), r[~held].copy()), (u[held].copy(), i[held].copy(), r[held].copy())


@pytest.fixture(scope="module")
def ml100k():
    w = mf.WORKLOADS["ml100k"]
    u, i, r, held = orc.generate(SEED, 0, w.n_ratings, w.n_users, w.n_items)
    return w, split(u, i, r, held)


# ------------------------------------------------------------------------------------------------
# integer / index paths: bit-exact
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nu,ni,l2ai,start", [(943, 1682, 3, 0), (480_000, 17_800, 3, 0),
                                              (10_000_000, 1_000_000, 4, 0), (10_000_000, 1_000_000, 4, 1_999_000_000)])
def test_generator_bit_exact(nu, ni, l2ai, start):
    n = 200_000
    sp = mf.synth_params(2_000_000_000, SEED, 2, 0.25, l2ai, 0.375)
    gu, gi, gr, gh = mf.generate_to_host(sp, nu, ni, start, n)
    ou, oi, or_, oh = orc.generate(SEED, start, n, nu, ni, 2, 0.25, l2ai, 0.375)
    assert np.array_equal(gu, ou) and np.array_equal(gi, oi)
    assert np.array_equal(gr.view(np.uint32), or_.view(np.uint32))
    assert np.array_equal(gh, oh)


def test_generator_golden(golden):
    g = golden["gen_ml100k"]
    sp = mf.synth_params(100_000, SEED)
    u, i, r, h = mf.generate_to_host(sp, 943, 1682, 0, 256)
    assert u.tolist() == g["u"] and i.tolist() == g["i"] and r.view(np.uint32).tolist() == g["r"]
    assert h.astype(int).tolist() == g["held"]


@pytest.mark.parametrize("k", [8, 32, 128])
def test_init_factors_bit_exact(k):
    nu, ni = 1000, 700
    with mf.Engine(mf.make_config(nu, ni, k, 0.01, 0.05, seed=SEED)) as eng:
        eng.load_ratings(np.zeros(1, np.int32), np.zeros(1, np.int32), np.ones(1, np.float32))
        eng.init_factors()
        P, Q = eng.get_factors()
    assert np.array_equal(P, orc.init_factors(nu, k, SEED, 0))
    assert np.array_equal(Q, orc.init_factors(ni, k, SEED, 1))


def test_deterministic_shuffle_order_is_the_stand_ins(ml100k):
    w, ((u, i, r), _) = ml100k
    cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=SEED, mode=capi.MODE_DETERMINISTIC)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(u, i, r)
        for epoch in (0, 7):
            eng.shuffle_once(epoch)
            gu, gi, gr, _ = eng.records()
            order = orc.shuffle(SEED, epoch, len(r))
            assert np.array_equal(gu, u[order]) and np.array_equal(gi, i[order])
            assert np.array_equal(gr.view(np.uint32), r[order].view(np.uint32))


# ------------------------------------------------------------------------------------------------
# the update rule
# ------------------------------------------------------------------------------------------------
def test_kat_hand_computed(kat):
    """k=2 hand KAT, padded to k=4 with zeros (they add exactly)."""
    p = np.array([kat["p"] + [0, 0]], dtype=np.float32)
    q = np.array([kat["q"] + [0, 0]], dtype=np.float32)
    pp, qq, e = mf.apply_updates_forced(4, kat["lr"], kat["lambda"], p, q, np.array([kat["r"]], np.float32))
    assert abs(e[0] - kat["e"]) < 1e-6
    np.testing.assert_allclose(pp[0, :2], kat["p_new"], rtol=1e-6)
    np.testing.assert_allclose(qq[0, :2], kat["q_new"], rtol=1e-6)
    assert pp[0, 2] == 0 and qq[0, 3] == 0


@pytest.mark.parametrize("epoch", [0, 19])
def test_per_update_parity_teacher_forced(ml100k, epoch):
    """Every update of an epoch, both sides starting from the ORACLE's pre-update rows (sequential
    summation = the stand-in's order): GPU result within 1e-5 relative; and bit-identical to the
    oracle when it sums in the kernel's warp-tree order."""
    w, ((u, i, r), _) = ml100k
    P = orc.init_factors(w.n_users, w.k, SEED, 0)
    Q = orc.init_factors(w.n_items, w.k, SEED, 1)
    if epoch:
        orc.train(u, i, r, P, Q, w.lr, w.lambda_, 0, epoch, SEED)
    order, pre_p, pre_q, post_p, post_q, err = orc.train_tape(u, i, r, P, Q, w.lr, w.lambda_, epoch, SEED)
    gp, gq, ge = mf.apply_updates_forced(w.k, w.lr, w.lambda_, pre_p, pre_q, r[order])
    for got, want in ((gp, post_p), (gq, post_q)):
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
        assert rel.max() <= REL_TOL, rel.max()
    assert np.abs(ge - err).max() <= 1e-5 * np.maximum(np.abs(err), 1.0).max()
    # stronger: the tree-order oracle is reproduced bit for bit
    tp, tq = pre_p.copy(), pre_q.copy()
    te = np.empty_like(err)
    for j in range(0, len(err), 11):
        te[j] = orc.lib.orc_sgd_update(tp[j], tq[j], w.k, r[order[j]], w.lr, w.lambda_, orc.ORDER_WARP_TREE)
        assert np.array_equal(tp[j], gp[j]) and np.array_equal(tq[j], gq[j]) and te[j] == ge[j]


def test_deterministic_mode_free_running_ml100k(ml100k):
    """configs[0] on the GPU in deterministic single-warp mode, 20 epochs free-running."""
    w, ((u, i, r), (hu, hi, hr)) = ml100k
    cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=SEED, mode=capi.MODE_DETERMINISTIC)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(u, i, r)
        eng.init_factors()
        stats, trace = eng.train_traced(w.epochs, len(r))
        P, Q = eng.get_factors()
        gpu_rmse = eng.rmse(hu, hi, hr)
    assert all(s.updates == len(r) for s in stats)
    # bit-exact against the oracle summing in the kernel's order, including every per-update error
    Pt = orc.init_factors(w.n_users, w.k, SEED, 0)
    Qt = orc.init_factors(w.n_items, w.k, SEED, 1)
    tt = orc.train(u, i, r, Pt, Qt, w.lr, w.lambda_, 0, w.epochs, SEED, orc.ORDER_WARP_TREE, trace=True)
    assert np.array_equal(trace, tt)
    assert np.array_equal(P, Pt) and np.array_equal(Q, Qt)
    # against the stand-in's sequential order: trajectories drift in the last ulps only
    Ps, Qs = orc.factorize(u, i, r, w.n_users, w.n_items, w.k, w.lr, w.lambda_, w.epochs, SEED)
    seq_rmse = orc.rmse(Ps, Qs, hu, hi, hr)
    assert abs(gpu_rmse - seq_rmse) / seq_rmse < 1e-4
    assert np.abs(P - Ps).max() < 1e-3 and np.abs(Q - Qs).max() < 1e-3


@pytest.mark.parametrize("k", [4, 8, 20, 64, 100, 128, 256, 320, 512])
def test_deterministic_mode_all_ranks_bit_exact(k):
    nu, ni, n = 300, 200, 6000
    u, i, r, _ = orc.generate(SEED + k, 0, n, nu, ni)
    got = mf.MatrixFactorizationSGD.factorize(u, i, r, nu, ni, k, 0.02, 0.03, 2, SEED, mode=capi.MODE_DETERMINISTIC)
    P, Q = orc.factorize(u, i, r, nu, ni, k, 0.02, 0.03, 2, SEED, orc.ORDER_WARP_TREE)
    assert np.array_equal(got.P, P) and np.array_equal(got.Q, Q)
    Ps, Qs = orc.factorize(u, i, r, nu, ni, k, 0.02, 0.03, 2, SEED, orc.ORDER_SEQ)
    assert np.abs(got.P - Ps).max() < 1e-4 and np.abs(got.Q - Qs).max() < 1e-4


def test_sgd_small_run_golden(golden):
    s = golden["sgd_small_shape"]
    u, i, r, _ = orc.generate(SEED, 0, s["n"], s["n_users"], s["n_items"])
    got = mf.MatrixFactorizationSGD.factorize(u, i, r, s["n_users"], s["n_items"], s["k"], s["lr"], s["lambda"],
                                              s["epochs"], SEED, mode=capi.MODE_DETERMINISTIC)
    assert got.P.view(np.uint32).ravel().tolist() == golden["sgd_small_tree"]["P"]
    assert got.Q.view(np.uint32).ravel().tolist() == golden["sgd_small_tree"]["Q"]


def test_empty_and_single_record():
    e = np.empty(0, np.int32)
    got = mf.MatrixFactorizationSGD.factorize(e, e, np.empty(0, np.float32), 5, 7, 8, 0.1, 0.1, 3, SEED,
                                              mode=capi.MODE_DETERMINISTIC)
    assert np.array_equal(got.P, orc.init_factors(5, 8, SEED, 0)) and np.array_equal(got.Q, orc.init_factors(7, 8, SEED, 1))
    got = mf.MatrixFactorizationSGD.factorize(e, e, np.empty(0, np.float32), 5, 7, 8, 0.1, 0.1, 3, SEED)   # hogwild
    assert np.array_equal(got.P, orc.init_factors(5, 8, SEED, 0))
    one = mf.MatrixFactorizationSGD.factorize(np.array([4], np.int32), np.array([6], np.int32), np.array([2.5], np.float32),
                                              5, 7, 8, 0.1, 0.1, 3, SEED)   # hogwild, one record: order is forced
    P, Q = orc.factorize(np.array([4], np.int32), np.array([6], np.int32), np.array([2.5], np.float32), 5, 7, 8,
                         0.1, 0.1, 3, SEED, orc.ORDER_WARP_TREE_FMA)
    assert np.array_equal(one.P, P) and np.array_equal(one.Q, Q)


def test_out_of_range_ids_rejected():
    with pytest.raises(mf.MfsgdError) as ei:
        mf.MatrixFactorizationSGD.factorize(np.array([0, 5], np.int32), np.array([0, 0], np.int32),
                                            np.ones(2, np.float32), 5, 3, 8, 0.1, 0.1, 1, SEED)
    assert ei.value.code == capi.E_INVALID_ARG


# ------------------------------------------------------------------------------------------------
# subsystem (3): RMSE kernel
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [32, 64, 128])
def test_rmse_kernel_vs_float64(k):
    nu, ni, n = 20_000, 5_000, 1_000_000
    u, i, r, _ = orc.generate(SEED + 3, 0, n, nu, ni)
    P = orc.init_factors(nu, k, SEED, 0, 0.3)
    Q = orc.init_factors(ni, k, SEED, 1, 0.3)
    got = mf.MatrixFactorizationSGD.rmse(P, Q, k, u, i, r)
    pred = np.einsum("ij,ij->i", P[u].astype(np.float64), Q[i].astype(np.float64))
    want = float(np.sqrt(np.mean((r.astype(np.float64) - pred) ** 2)))
    assert abs(got - want) / want <= 1e-6
    assert abs(got - orc.rmse(P, Q, u, i, r)) / want <= 1e-6
    assert mf.MatrixFactorizationSGD.rmse(P, Q, k, u[:0], i[:0], r[:0]) == 0.0


# ------------------------------------------------------------------------------------------------
# subsystem (1): bucketing + shuffle
# ------------------------------------------------------------------------------------------------
def rec_keys(u, i, r):
    return np.sort((u.astype(np.uint64) << np.uint64(40)) ^ (i.astype(np.uint64) << np.uint64(20)) ^
                   r.view(np.uint32).astype(np.uint64))


@pytest.mark.parametrize("mode,G,mu,mi", [(capi.MODE_HOGWILD, 1, 1, 1), (capi.MODE_HOGWILD, 1, 5, 3),
                                          (capi.MODE_DSGD, 4, 2, 2), (capi.MODE_DSGD, 8, 1, 1)])
def test_bucketing_layout_and_shuffle(mode, G, mu, mi):
    nu, ni, n = 30_000, 4_000, 1_500_000
    u, i, r, held = orc.generate(SEED + 5, 0, n, nu, ni)
    cfg = mf.make_config(nu, ni, 32, 0.01, 0.05, seed=SEED, mode=mode, n_gpus=G, stripes_per_gpu=mu, shards_per_gpu=mi,
                         flags=capi.FLAG_VIRTUAL_RING if G > 1 else 0)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(u, i, r)
        info = eng.layout_info()
        assert (info.user_blocks, info.item_blocks, info.n_train_local, info.n_train_total) == (G * mu, G * mi, n, n)
        ub, ib = eng.bounds()
        assert ub[0] == 0 and ub[-1] == nu and ib[0] == 0 and ib[-1] == ni
        assert np.all(np.diff(ub) >= 0) and np.all(np.diff(ib) >= 0)
        # stripes balanced by rating count (not by row count)
        ucnt = np.add.reduceat(np.bincount(u, minlength=nu), ub[:-1])
        assert ucnt.max() <= 1.1 * n / (G * mu) + np.bincount(u).max()
        IB = G * mi
        H = info.n_hot_items
        icount = np.bincount(i, minlength=ni)
        hot_ids = np.flatnonzero(icount >= max(np.float32(1e-6) * np.float64(n), 16.0 * mu * G))   # the default rule
        assert H == len(hot_ids) and H > 0
        all_keys = []
        for g in range(G):
            gu, gi, gr, off = eng.records(g)
            assert len(off) == mu * (IB + H) + 1
            assert off[0] == 0 and off[-1] == len(gu) and np.all(np.diff(off) >= 0)
            for b in range(mu * IB):                                   # cold blocks: right stripe, right shard, no hot item
                a, c = divmod(b, IB)
                su, si = gu[off[b]:off[b + 1]], gi[off[b]:off[b + 1]]
                if len(su):
                    assert su.min() >= ub[g * mu + a] and su.max() < ub[g * mu + a + 1]
                    assert si.min() >= ib[c] and si.max() < ib[c + 1]
                    assert not np.isin(si, hot_ids).any()
            for a in range(mu):                                        # hot buckets: one item each, right stripe
                for hx in range(H):
                    b = mu * IB + a * H + hx
                    su, si = gu[off[b]:off[b + 1]], gi[off[b]:off[b + 1]]
                    if len(su):
                        assert np.all(si == hot_ids[hx])
                        assert su.min() >= ub[g * mu + a] and su.max() < ub[g * mu + a + 1]
            all_keys.append(rec_keys(gu, gi, gr))
        assert np.array_equal(np.sort(np.concatenate(all_keys)), rec_keys(u, i, r))     # multiset preserved
        # shuffle: a permutation inside every block, different per epoch
        before = eng.records(0)
        eng.shuffle_once(0)
        e0 = eng.records(0)
        eng.shuffle_once(1)
        e1 = eng.records(0)
        off = before[3]
        assert np.array_equal(off, e0[3])
        moved = 0
        nblk = len(off) - 1
        for b in range(nblk):
            s = slice(off[b], off[b + 1])
            nb = int(off[b + 1] - off[b])
            if nb > 0 and b % 7 == 0:        # the materialised order is the oracle's restatement of the permutation, bit for bit
                for ep, src, got in ((0, before, e0), (1, e0, e1)):      # each materialising pass permutes the current layout
                    perm = orc.block_perm(nb, SEED, ep, 0 * nblk + b)   # bucket id = member * blocks + block; member 0 here
                    assert np.array_equal(got[0][s], src[0][s][perm]) and np.array_equal(got[1][s], src[1][s][perm])
                    assert np.array_equal(got[2][s].view(np.uint32), src[2][s][perm].view(np.uint32))
            kb = rec_keys(before[0][s], before[1][s], before[2][s])
            assert np.array_equal(kb, rec_keys(e0[0][s], e0[1][s], e0[2][s]))
            assert np.array_equal(kb, rec_keys(e1[0][s], e1[1][s], e1[2][s]))
            moved += int(np.sum(before[0][s] != e0[0][s]))
        assert moved > 0.5 * len(before[0])
        assert not np.array_equal(e0[0], e1[0])


@pytest.mark.parametrize("marks", ["off", "all"])
def test_virtual_reshuffle_visits_the_materialised_order(marks):
    """The permutation the update kernels apply on the fly is the one block_shuffle_kernel materialises: with one
    run per hot item (sequential, hence order-sensitive) an epoch trained through the virtual reshuffle must equal
    the oracle walking the order that mfsgd_shuffle_once materialises from the same layout."""
    n_hot, per_hot, n_cold = 4, 2500, 3001
    n = n_hot * per_hot + n_cold
    rng = np.random.default_rng(11)
    items = np.concatenate([np.repeat(np.arange(n_hot), per_hot), n_hot + np.arange(n_cold)]).astype(np.int32)
    i = items[rng.permutation(n)]
    u = rng.permutation(n).astype(np.int32)
    r = (1 + 4 * rng.random(n)).astype(np.float32)
    ni, k = n_hot + n_cold, 128
    cfg = mf.make_config(n, ni, k, 0.01, 0.03, seed=SEED, mode=capi.MODE_HOGWILD, stripes_per_gpu=1, rounds=1, hot_chunk=4096,
                         p_atomic_threshold=-1.0 if marks == "off" else 1e-9)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(u, i, r)                 # bucketing order inside a bucket is arbitrary: stay on this one layout
        assert eng.layout_info().n_heavy_users == (0 if marks == "off" else n)
        eng.init_factors()
        eng.train(1)                              # epoch 0, records read through the permutation, layout untouched
        P, Q = eng.get_factors()
        eng.shuffle_once(0)                       # now materialise epoch 0's permutation of that same layout
        ou, oi, orr, _ = eng.records()
    Po, Qo = orc.init_factors(n, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    train_runs_then_cold(ou, oi, orr, n_hot, Po, Qo, k, 0.01, 0.03, 0, 1, orc.ORDER_WARP_TREE_FMA, heavy=marks == "all")
    assert np.array_equal(Q, Qo) and np.array_equal(P, Po)


def test_synthetic_on_device_matches_oracle_split():
    nu, ni, n = 20_000, 3_000, 1_000_000
    ou, oi, or_, oh = orc.generate(SEED, 0, n, nu, ni)
    with mf.Engine(mf.make_config(nu, ni, 32, 0.01, 0.05, seed=SEED, stripes_per_gpu=2)) as eng:
        nt, nh = eng.generate_synthetic(mf.synth_params(n, SEED))
        assert (nt, nh) == (int((~oh).sum()), int(oh.sum()))
        gu, gi, gr, _ = eng.records(0)
        assert np.array_equal(rec_keys(gu, gi, gr), rec_keys(ou[~oh], oi[~oh], or_[~oh]))
        eng.init_factors()
        rm, sse, cnt = eng.rmse_heldout()
        assert cnt == nh
        P, Q = eng.get_factors()
        assert abs(rm - orc.rmse(P, Q, ou[oh].copy(), oi[oh].copy(), or_[oh].copy())) / rm < 1e-6
        rt, _, ct = eng.rmse_train()
        assert ct == nt and abs(rt - orc.rmse(P, Q, ou[~oh].copy(), oi[~oh].copy(), or_[~oh].copy())) / rt < 1e-6


# ------------------------------------------------------------------------------------------------
# Hogwild and DSGD: convergence parity (held-out RMSE within 0.5 % of the oracle at equal epochs)
# ------------------------------------------------------------------------------------------------
def midsize_fixture():
    """tests/golden/oracle_rmse_midsize.json (tools/midsize_oracle_curves.py; tests/test_oracle.py re-derives its first epochs)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_rmse_midsize.json")) as f:
        return json.load(f)


class MidSet:
    """ML-20M-shaped scaled 1:10 (13.8K x 2.7K, 2M ratings), k=32. `signal` selects the signal-dominant variant
    (workloads.SIGNAL_*; lr 0.02, lambda 0.02, 20 epochs: the sequential oracle ends 79 % below the constant predictor; its curve
    falls 25 % per epoch at first and 0.5 % per epoch at the end). The oracle's curves are a committed fixture (minutes of CPU)."""
    MODEL = "plain"

    def __init__(self, signal):
        fx = midsize_fixture()
        self.variant = "signal" if signal else "default"
        par, self.fx = fx["params"][self.variant], fx[self.variant]
        self.nu, self.ni, n, self.k = fx["n_users"], fx["n_items"], fx["n_ratings"], fx["k"]
        assert fx["seed"] == SEED
        self.lr, self.lam, self.epochs = par["lr"], par["lam"], par["epochs"]
        self.synth = dict(amplitude=par["amplitude"], noise_scale=par["noise_scale"])
        u, i, r, held = orc.generate(SEED, 0, n, self.nu, self.ni, **self.synth)
        self.train, self.held = split(u, i, r, held)
        assert (len(self.train[2]), len(self.held[2])) == (self.fx["n_train"], self.fx["n_heldout"])
        self.curve = self.fx[self.MODEL]["shuffled"]
        self.oracle_rmse = self.curve[-1]
        self.const_rmse = self.fx["constant_predictor_rmse"]

    def __getitem__(self, key):          # the round-1 tests index the fixture like a dict
        return {"nu": self.nu, "ni": self.ni, "k": self.k, "lr": self.lr, "lam": self.lam, "epochs": self.epochs,
                "train": self.train, "held": self.held, "oracle_rmse": self.oracle_rmse}[key]

    def dsgd_oracle_rmse(self, user_bounds, item_bounds):
        """The sequential rule walking the DSGD schedule's block order for these strata (pyoracle.dsgd_order); the fixture holds
        the curves for the rating-count-balanced strata of G = 2, 4, 8, which are the engine's."""
        G = len(user_bounds) - 1
        assert np.array_equal(user_bounds, orc.balanced_bounds(self.train[0], self.nu, G))
        assert np.array_equal(item_bounds, orc.balanced_bounds(self.train[1], self.ni, G))
        return self.fx[self.MODEL]["dsgd%d" % G][-1]


def assert_curve_parity(curve, want):
    """Hogwild against the sequential oracle at equal epochs: the final held-out RMSE within 0.5 %, both sides (north_star); on the
    way there (from the 3rd epoch on) never more than half an epoch behind -- on a steep curve an epoch is worth 2-25 %, so this is
    the sharper of the two where it applies -- and never more than 1 % ahead."""
    assert len(curve) == len(want)
    assert_rmse_parity(curve[-1], want[-1])
    for e in range(2, len(want)):
        assert want[e] * (1 - 2 * RMSE_TOL) <= curve[e] <= max(want[e], 0.5 * (want[e] + want[e - 1])) * (1 + RMSE_TOL), (e, curve[e], want[e - 1], want[e])


def median_run(run, n=3):
    """Hogwild is not deterministic: from run to run the final held-out RMSE of a mid-size set moves by +-0.2 % around a
    configuration's own mean (tools/rmse_spread.py, profiles/r02_experiments.md section 12: e.g. -0.10 ... -0.37 % over 8 runs of
    the signal-dominant set on the planned layout, one run in ~30 at -0.50 %), the same size as the bar itself. The parity bars are
    therefore applied to the MEDIAN of n runs. `run` returns a tuple whose first element is the final held-out RMSE."""
    outs = sorted((run() for _ in range(n)), key=lambda o: o[0])
    return outs[n // 2]


@pytest.fixture(scope="module")
def midsize():
    return MidSet(signal=False)


@pytest.fixture(scope="module")
def midsize_signal():
    m = MidSet(signal=True)
    assert m.oracle_rmse < 0.7 * m.const_rmse          # the variant's reason to exist: training beats the mean by far
    return m


@pytest.mark.parametrize("k", [8, 32, 64, 100, 128, 256, 512])
@pytest.mark.parametrize("arith", ["fast", "exact", "exact-atomic", "fast-materialized"])
def test_hogwild_kernel_bit_exact_on_conflict_free_data(k, arith):
    """Records with pairwise distinct users and items commute exactly, so the full-grid Hogwild kernel
    (tiles, sub-warps, prefetch, tails, blocking) must reproduce the oracle bit for bit in any order:
    the FFMA2 arrangement against ORDER_WARP_TREE_FMA, MFSGD_FLAG_EXACT_ARITH against ORDER_WARP_TREE."""
    n = 5003                                           # not a multiple of 32
    rng = np.random.default_rng(k)
    u = rng.permutation(n).astype(np.int32)
    i = rng.permutation(n).astype(np.int32)
    r = (1 + 4 * rng.random(n)).astype(np.float32)
    order = orc.ORDER_WARP_TREE_FMA if arith.startswith("fast") else orc.ORDER_WARP_TREE
    flags = {"fast": 0, "fast-materialized": capi.FLAG_MATERIALIZE_SHUFFLE}.get(arith, capi.FLAG_EXACT_ARITH)
    scatter = capi.SCATTER_ATOMIC if arith == "exact-atomic" else capi.SCATTER_STORE
    P, Q = orc.factorize(u, i, r, n, n, k, 0.02, 0.03, 3, SEED, order)
    for mu, mi in ((1, 1), (3, 2)):
        got = mf.MatrixFactorizationSGD.factorize(u, i, r, n, n, k, 0.02, 0.03, 3, SEED, mode=capi.MODE_HOGWILD,
                                                  stripes_per_gpu=mu, shards_per_gpu=mi, scatter=scatter, flags=flags)
        if arith == "exact-atomic":      # p + fl(delta) rounds once more than the store path: equal to 1 ulp
            np.testing.assert_allclose(got.P, P, rtol=3e-7, atol=1e-9)
            np.testing.assert_allclose(got.Q, Q, rtol=3e-7, atol=1e-9)
        else:
            assert np.array_equal(got.P, P) and np.array_equal(got.Q, Q)
    if arith != "exact-atomic":
        ring = mf.MatrixFactorizationSGD.factorize(u, i, r, n, n, k, 0.02, 0.03, 3, SEED, mode=capi.MODE_DSGD, n_gpus=4,
                                                   stripes_per_gpu=2, scatter=scatter, flags=flags | capi.FLAG_VIRTUAL_RING)
        assert np.array_equal(ring.P, P) and np.array_equal(ring.Q, Q)
    # the two arithmetics are the same rule: a few ulp apart per update
    Pe, Qe = orc.factorize(u, i, r, n, n, k, 0.02, 0.03, 3, SEED, orc.ORDER_SEQ)
    np.testing.assert_allclose(P, Pe, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(Q, Qe, rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("arith", ["fast", "exact", "fast-heavy", "exact-heavy", "fast-atomic-p"])
@pytest.mark.parametrize("k", [8, 32, 64, 100, 128, 200, 256, 512])
def test_hot_item_kernel_exact_sequential_runs(k, arith):
    """Run path: with one run per item (hot_chunk >= run length) the kernel applies an item's ratings strictly in
    bucket order with q_i in registers -- must equal the oracle bit for bit when users are pairwise distinct.
    Geometries of the run kernel: 8 lanes (k = 8 with idle lanes, 32: 4 runs side by side per warp), 16 lanes (64),
    32 lanes x 1..4 chunks (100 with a partial chunk, 128, 200, 256, 512). Cold records (distinct items) ride along."""
    n_hot, per_hot, n_cold = 5, 3000, 5003
    n = n_hot * per_hot + n_cold
    rng = np.random.default_rng(7)
    items = np.concatenate([np.repeat(np.arange(n_hot), per_hot), n_hot + np.arange(n_cold)]).astype(np.int32)
    perm = rng.permutation(n)
    i = items[perm]
    u = rng.permutation(n).astype(np.int32)
    r = (1 + 4 * rng.random(n)).astype(np.float32)
    ni = n_hot + n_cold
    # "-heavy": every user's records carry the heavy mark, "-atomic-p": MFSGD_SCATTER_ATOMIC_P -- both make the run kernel add
    # p_u's increment in memory (red.global.add) instead of storing the new row; the cold kernel does the same for marked records.
    heavy = arith.endswith("-heavy") or arith.endswith("-atomic-p")
    exact = arith.startswith("exact")
    cfg = mf.make_config(n, ni, k, 0.01, 0.03, seed=SEED, mode=capi.MODE_HOGWILD, stripes_per_gpu=1, rounds=1,
                         hot_chunk=4096, flags=capi.FLAG_NO_SHUFFLE | (capi.FLAG_EXACT_ARITH if exact else 0),
                         p_atomic_threshold=1e-9 if arith.endswith("-heavy") else -1.0,
                         scatter=capi.SCATTER_ATOMIC_P if arith.endswith("-atomic-p") else capi.SCATTER_STORE)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(u, i, r)
        assert eng.layout_info().n_hot_items == n_hot
        assert eng.layout_info().n_heavy_users == (n if arith.endswith("-heavy") else 0)
        mu_, _, _, _ = eng.records(with_marks=True)
        assert np.all((mu_ < 0) == arith.endswith("-heavy")) and np.array_equal(mu_ & 0x7fffffff, eng.records()[0])
        ou, oi, orr, off = eng.records()                  # the order the kernels will see (no reshuffle)
        eng.init_factors()
        eng.train(3)
        P, Q = eng.get_factors()
    Po, Qo = orc.init_factors(n, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    if arith == "fast-atomic-p":      # the cold kernel's atomic variants add exact-rule increments (a few ulp from the FMA twin)
        hot = oi < n_hot
        with orc.tree_lanes(orc.run_lanes(k)):
            orc.train(ou[hot].copy(), oi[hot].copy(), orr[hot].copy(), Po, Qo, 0.01, 0.03, 0, 3, SEED, orc.ORDER_WARP_TREE_FMA_PDELTA, shuffled=False)
        assert np.array_equal(P[ou[hot]], Po[ou[hot]]) and np.array_equal(Q[:n_hot], Qo[:n_hot])
        return
    train_runs_then_cold(ou, oi, orr, n_hot, Po, Qo, k, 0.01, 0.03, 0, 3,
                         orc.ORDER_WARP_TREE if exact else orc.ORDER_WARP_TREE_FMA, heavy=heavy)
    assert np.array_equal(P, Po) and np.array_equal(Q, Qo)



def plan_runs_of(off, n_hot, hot_items, rounds, chunk, boost, mu=1, IB=1, member=0):
    """The engine's own run plan for a layout (host-only hook mfsgd_plan_runs) as an oracle RunPlan + per-visit unit ranges."""
    import ctypes as C
    cap = int(((off[mu * IB + 1:] - off[mu * IB:-1]) // chunk + rounds + 1).sum()) + 16
    st, ct, it, wt = np.zeros(cap, np.int64), np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float32)
    n = C.c_int64(cap)
    visits = np.zeros(mu * rounds * IB + 1, np.int32)
    hbl = np.array([0, n_hot], np.int32) if IB == 1 else None
    hit = np.ascontiguousarray(hot_items, np.int32)
    capi.check(capi.lib.mfsgd_plan_runs(capi.ptr(off), mu, n_hot, IB, capi.ptr(hbl), capi.ptr(hit), rounds, chunk, SEED, member, boost,
                                        capi.ptr(st), capi.ptr(ct), capi.ptr(it), capi.ptr(wt), C.byref(n), capi.ptr(visits)))
    m = n.value
    return orc.RunPlan(st[:m], ct[:m], it[:m], wt[:m], off, member=member), visits


@pytest.mark.parametrize("boost,marks", [(1.0, "off"), (1.25, "off"), (1.25, "all")])
@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("k,rounds", [(128, 1), (32, 1), (128, 2)])
def test_run_kernel_averaged_merge_matches_its_oracle_twin(k, rounds, shuffle, boost, marks):
    """The dominant branch of the bench workload: several runs of one item in one launch, merged with
    red.global.add of weight * (q_run - q_start). Oracle twin = oracle.cpp orc_train_runs_launch fed with the plan the
    engine itself uses (mfsgd_plan_runs): every run walks its records sequentially from the launch-start q_i, the merge adds
    the same weighted differences -- the order of the float adds is the only freedom left, hence <= 1e-5 relative.
    Users are pairwise distinct (P rows commute); every item has 8 equally long runs per launch, which the launch's
    grid (>= 2 waves: 3 CTAs at k = 128, 1 CTA of 32 sub-warps at k = 32) takes item-aligned wave by wave, so which runs
    are in flight together does not depend on timing."""
    n_hot, pieces, chunk, n_cold = 6, 8, 64, 1003
    per_hot = pieces * chunk * rounds
    n = n_hot * per_hot + n_cold
    rng = np.random.default_rng(100 + k + rounds)
    items = np.concatenate([np.repeat(np.arange(n_hot), per_hot), n_hot + np.arange(n_cold)]).astype(np.int32)
    i = items[rng.permutation(n)]
    u = rng.permutation(n).astype(np.int32)
    nu = n
    r = (1 + 4 * rng.random(n)).astype(np.float32)
    ni, lr, lam, epochs = n_hot + n_cold, 0.01, 0.03, 3
    # "off": no heavy-user marks (p_u stored); "all": every user marked heavy (red.global.add of p_u's increment, PDELTA rounding)
    thr = {"off": -1.0, "all": 1e-9}[marks]
    cfg = mf.make_config(nu, ni, k, lr, lam, seed=SEED, mode=capi.MODE_HOGWILD, stripes_per_gpu=1, rounds=rounds, hot_chunk=chunk,
                         merge_boost=boost, flags=0 if shuffle else capi.FLAG_NO_SHUFFLE, p_atomic_threshold=thr)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(u, i, r)
        assert eng.layout_info().n_hot_items == n_hot
        ou, oi, orr, off = eng.records(with_marks=True)
        assert eng.layout_info().n_heavy_users == {"off": 0, "all": n}[marks]
        assert np.all((ou < 0) == (marks == "all"))
        eng.init_factors()
        eng.train(epochs)
        P, Q = eng.get_factors()
    plan, visits = plan_runs_of(off, n_hot, np.arange(n_hot), rounds, chunk, boost)
    assert len(plan.start) == n_hot * pieces * rounds and np.all(plan.count == chunk)
    assert np.allclose(plan.weight, min(1.0, boost / pieces))
    per_warp = 32 // orc.run_lanes(k)
    Po, Qo = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    cold = slice(int(off[0]), int(off[1]))
    for e in range(epochs):
        for rnd in range(rounds):                      # one sub-stripe: visit = round
            lo, hi = int(visits[rnd]), int(visits[rnd + 1])
            grid = max(1, -(-(hi - lo) // (16 * per_warp)))
            with orc.tree_lanes(orc.run_lanes(k)):
                orc.train_runs_launch(ou, orr, plan, lo, hi, Po, Qo, lr, lam, orc.ORDER_WARP_TREE_FMA, grid * 8 * per_warp, per_warp,
                                      virt=shuffle, seed=SEED, epoch=e)
        # the cold block rides along: distinct users and items, any order
        orc.train((ou[cold] & 0x7fffffff).copy(), oi[cold].copy(), orr[cold].copy(), Po, Qo, lr, lam, e, e + 1, SEED, orc.ORDER_WARP_TREE_FMA, shuffled=False)
    ou = ou & 0x7fffffff
    np.testing.assert_allclose(Q, Qo, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(P, Po, rtol=1e-5, atol=1e-7)
    # and the merge really was an average: a plain sequential walk of each item's bucket ends somewhere else
    Ps, Qs = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    train_runs_then_cold(ou, oi, orr, n_hot, Ps, Qs, k, lr, lam, 0, epochs, orc.ORDER_WARP_TREE_FMA)
    assert np.abs(Q[:n_hot] - Qs[:n_hot]).max() > 1e-3


def test_hot_item_path_can_be_disabled(midsize):
    m = midsize
    cfg = mf.make_config(m["nu"], m["ni"], m["k"], m["lr"], m["lam"], seed=SEED, hot_share=-1.0)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(*m["train"])
        assert eng.layout_info().n_hot_items == 0
    cfg = mf.make_config(m["nu"], m["ni"], m["k"], m["lr"], m["lam"], seed=SEED)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(*m["train"])
        assert eng.layout_info().n_hot_items > 500


def hogwild_run(m, **cfg_kw):
    cfg = mf.make_config(m.nu, m.ni, m.k, m.lr, m.lam, seed=SEED, mode=capi.MODE_HOGWILD, **cfg_kw)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(*m.train)
        eng.load_heldout(*m.held)
        eng.init_factors()
        eng.set_eval_every_epoch(True)
        stats = eng.train(m.epochs)
        got = eng.rmse(*m.held)
    assert abs(stats[-1].heldout_rmse - got) < 1e-9
    assert stats[0].heldout_rmse > stats[-1].heldout_rmse
    return got, [s.heldout_rmse for s in stats]


@pytest.mark.parametrize("flags", [0, capi.FLAG_MATERIALIZE_SHUFFLE])
@pytest.mark.parametrize("variant", ["default", "signal"])
def test_hogwild_rmse_parity(midsize, midsize_signal, variant, flags):
    """Held-out RMSE within 0.5 % of the sequential oracle at equal epochs, BOTH sides (median of three runs), on the noise-dominant
    set of the throughput workloads and on the signal-dominant one (where a wrong merge weight or lost updates cost whole per
    cents), on the layout the engine plans for the set.
    Default: the update kernels read every bucket through its per-epoch permutation (virtual reshuffle);
    MFSGD_FLAG_MATERIALIZE_SHUFFLE runs the reshuffle kernel instead."""
    m = midsize_signal if variant == "signal" else midsize
    _, curve = median_run(lambda: hogwild_run(m, flags=flags))
    assert_curve_parity(curve, m.curve)


@pytest.mark.parametrize("flags", [0, capi.FLAG_MATERIALIZE_SHUFFLE])
@pytest.mark.parametrize("variant", ["default", "signal"])
def test_hogwild_rmse_parity_forced_four_substripes_within_three_quarters_of_a_percent(midsize, midsize_signal, variant, flags):
    """The same set forced into 4 P sub-stripes of 3 450 users (a layout the planner would not choose for 13.8 K users: every launch
    then walks a quarter of the users with 862 sub-warps). Measured over 8 runs on the signal-dominant set: +0.06 ... +0.45 %
    (virtual reshuffle) and +0.07 ... +0.62 % (materialised), mean +0.2 / +0.3 % -- a configuration-made lag on top of the run-to-run
    spread; it is held to 0.75 % (median of three) and the name says so. The noise-dominant set stays inside 0.2 %."""
    m = midsize_signal if variant == "signal" else midsize
    got, curve = median_run(lambda: hogwild_run(m, stripes_per_gpu=4, flags=flags))
    assert abs(got / m.oracle_rmse - 1.0) <= 0.0075, (got, m.oracle_rmse)
    assert curve[-1] < curve[2] < curve[0]                            # and it keeps converging


@pytest.mark.parametrize("G,mu,mi", [(2, 1, 1), (4, 2, 2), (8, 1, 1)])
@pytest.mark.parametrize("variant", ["default", "signal"])
def test_dsgd_virtual_ring_rmse_parity(midsize, midsize_signal, variant, G, mu, mi):
    """The DSGD scheduler with G ring members placed on one GPU (streams instead of devices)."""
    m = midsize_signal if variant == "signal" else midsize
    cfg = mf.make_config(m.nu, m.ni, m.k, m.lr, m.lam, seed=SEED, mode=capi.MODE_DSGD, n_gpus=G,
                         stripes_per_gpu=mu, shards_per_gpu=mi, flags=capi.FLAG_VIRTUAL_RING | capi.FLAG_TIME_KERNELS)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(*m.train)
        ub, ib = eng.bounds()
        eng.init_factors()
        P0, Q0 = eng.get_factors()
        assert np.array_equal(P0, orc.init_factors(m.nu, m.k, SEED, 0))      # stripes reassemble the whole
        assert np.array_equal(Q0, orc.init_factors(m.ni, m.k, SEED, 1))
        stats = eng.train(m.epochs)
        assert all(s.updates == len(m.train[2]) for s in stats)
        rounds = eng.layout_info().rounds
        assert all(s.update_launches <= 2 * G * mu * rounds and s.update_kernel_ms > 0 for s in stats)   # cold + hot per visit
        got = eng.rmse(*m.held)
        P, Q = eng.get_factors()
    assert abs(got - orc.rmse(P, Q, *m.held)) / got < 1e-6                         # factors came home intact
    # the engine's strata are the oracle's (rating-count-balanced bounds)
    assert np.array_equal(ub[::mu], orc.balanced_bounds(m.train[0], m.nu, G))
    assert np.array_equal(ib[::mi], orc.balanced_bounds(m.train[1], m.ni, G))
    assert_ring_rmse_parity(got, m.oracle_rmse, m.dsgd_oracle_rmse(ub[::mu], ib[::mi]))


def test_set_get_factors_roundtrip_and_resume(midsize):
    m = midsize
    rng = np.random.default_rng(0)
    P = rng.random((m["nu"], m["k"]), dtype=np.float32)
    Q = rng.random((m["ni"], m["k"]), dtype=np.float32)
    cfg = mf.make_config(m["nu"], m["ni"], m["k"], m["lr"], m["lam"], seed=SEED, mode=capi.MODE_DSGD, n_gpus=4,
                         stripes_per_gpu=2, flags=capi.FLAG_VIRTUAL_RING)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(*m["train"])
        eng.set_factors(P, Q)
        P2, Q2 = eng.get_factors()
        assert np.array_equal(P, P2) and np.array_equal(Q, Q2)
        with pytest.raises(ValueError):
            eng.set_factors(P[:, :8], Q)


def test_state_errors():
    with mf.Engine(mf.make_config(10, 10, 8, 0.1, 0.1)) as eng:
        for fn in (eng.init_factors, lambda: eng.train(1), eng.get_factors, eng.layout_info):
            with pytest.raises(mf.MfsgdError) as ei:
                fn()
            assert ei.value.code == capi.E_STATE
        eng.load_ratings(np.zeros(3, np.int32), np.zeros(3, np.int32), np.ones(3, np.float32))
        with pytest.raises(mf.MfsgdError) as ei:
            eng.train(1)
        assert ei.value.code == capi.E_STATE


def test_c_harness_gpu_one_shot():
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.check_output([os.path.join(root, "tests", "c", "abi_harness"), capi.LIB_PATH, "gpu"], text=True)
    assert "GPU one-shot factorize" in out


def test_cpp_host_mirror_runs_every_entry_point():
    """host/factorize_demo: the C++ twin of the Java host through factorize, factorizeMixed, factorizeModel, factorizeEarlyStop,
    rmse and rmseModel (stand-in :109, :439, :305, :350, :169, :389) over dlopen/dlsym -- exit 0 only if every one trains."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([os.path.join(root, "host", "factorize_demo"), capi.LIB_PATH], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "early stopping:" in out.stdout and "extended model:" in out.stdout and "binary16 rows of P:" in out.stdout
