"""The run planner (csrc/run_plan.hpp through the host-only hook mfsgd_plan_runs): which records of which run bucket every
launch of the run kernel walks. Pure host logic -- runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

from matrixfactorizationsgd.java_b200 import _capi as capi

MIN_RUN = 16


def plan(sizes_cold, sizes_hot, mu, H, IB, hot_block_lo, hot_items, rounds, chunk, seed=7, member=0, boost=1.0):
    sizes = np.concatenate([np.asarray(sizes_cold, np.int64).ravel(), np.asarray(sizes_hot, np.int64).ravel()])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    cap = int(np.sum((np.asarray(sizes_hot) + chunk - 1) // chunk + rounds)) + 16
    start = np.zeros(cap, np.int64); count = np.zeros(cap, np.int32); item = np.zeros(cap, np.int32)
    weight = np.zeros(cap, np.float32)
    n = C.c_int64(cap)
    visits = np.zeros(mu * rounds * IB + 1, np.int32)
    hbl = np.asarray(hot_block_lo, np.int32); hit = np.asarray(hot_items, np.int32)
    capi.check(capi.lib.mfsgd_plan_runs(capi.ptr(off), mu, H, IB, capi.ptr(hbl), capi.ptr(hit), rounds, chunk, seed, member, boost,
                                        capi.ptr(start), capi.ptr(count), capi.ptr(item), capi.ptr(weight), C.byref(n), capi.ptr(visits)))
    m = n.value
    return off, start[:m], count[:m], item[:m], weight[:m], visits


@pytest.mark.parametrize("mu,IB,rounds,chunk", [(1, 1, 1, 256), (4, 1, 4, 256), (2, 3, 4, 64), (3, 2, 8, 96), (1, 2, 2, 4096)])
def test_every_record_of_every_run_bucket_is_walked_once_per_epoch(mu, IB, rounds, chunk):
    rng = np.random.default_rng(mu * 100 + IB * 10 + rounds)
    H = 37
    hot_items = np.sort(rng.choice(5000, H, replace=False)).astype(np.int32)
    cut = np.sort(rng.choice(np.arange(1, H), IB - 1, replace=False)) if IB > 1 else np.array([], int)
    hot_block_lo = np.concatenate([[0], cut, [H]]).astype(np.int32)
    sizes_cold = rng.integers(0, 500, (mu, IB))
    sizes_hot = rng.choice([0, 1, 15, 16, 31, 32, 63, 64, 65, 100, 255, 256, 257, 1000, 5000], (mu, H))
    off, start, count, item, weight, visits = plan(sizes_cold, sizes_hot, mu, H, IB, hot_block_lo, hot_items, rounds, chunk)
    assert visits[0] == 0 and visits[-1] == len(start) and np.all(np.diff(visits) >= 0)
    assert np.all(count >= 1) and np.all(count <= chunk)
    hot_base = mu * IB
    covered = np.zeros(off[-1], np.int32)
    for s, c in zip(start, count):
        covered[s:s + c] += 1
    assert np.all(covered[:off[hot_base]] == 0)                      # cold blocks are not the run kernel's
    assert np.all(covered[off[hot_base]:] == 1)                      # every run-bucket record exactly once per epoch
    item_block = np.searchsorted(hot_block_lo, np.arange(H), side="right") - 1
    for sa in range(mu):
        rounds_of_bucket = {}
        for rnd in range(rounds):
            for ib in range(IB):
                v = (sa * rounds + rnd) * IB + ib
                lo, hi = visits[v], visits[v + 1]
                assert np.all(np.diff(count[lo:hi]) <= 0)                # longest first
                for j in range(lo, hi):
                    hx = int(np.searchsorted(hot_items, item[j]))
                    assert hot_items[hx] == item[j] and item_block[hx] == ib
                    blk = hot_base + sa * H + hx
                    assert off[blk] <= start[j] and start[j] + count[j] <= off[blk + 1]     # inside its own bucket
                    rounds_of_bucket.setdefault(hx, {}).setdefault(rnd, []).append(j)
        for hx, per_round in rounds_of_bucket.items():
            bn = int(sizes_hot[sa, hx])
            assert len(per_round) == min(rounds, max(1, bn // (2 * MIN_RUN)))      # small buckets: one pass
            for rnd, js in per_round.items():
                n = int(count[js].sum())
                assert len(js) == -(-n // chunk)                                      # ceil(n / chunk) equal runs
                assert np.allclose(weight[js], 1.0 / len(js)) and count[js].max() - count[js].min() <= 1


def test_small_buckets_spread_evenly_over_the_rounds():
    mu, IB, rounds, H = 1, 1, 4, 4000
    hot_items = np.arange(H, dtype=np.int32)
    off, start, count, item, weight, visits = plan(np.zeros((1, 1)), np.full((1, H), 40), mu, H, IB, [0, H], hot_items, rounds, 256)
    per_round = np.diff(visits)
    assert per_round.sum() == H and per_round.min() > 0.8 * H / rounds and per_round.max() < 1.2 * H / rounds


def _plan_restated(off, mu, H, IB, hot_block_lo, hot_items, rounds, chunk, seed, member, boost):
    """The plan as run_plan.hpp's header comment states it, in plain Python (one thread, Python integers): slices of every
    bucket hashed over the rounds, cut into equal pieces, every visit's runs longest first (stable)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from np_restatement import hash64
    hot_base = mu * IB
    n_blocks = mu * (IB + H)
    visits = {}
    for sa in range(mu):
        for ib in range(IB):
            for hx in range(hot_block_lo[ib], hot_block_lo[ib + 1]):
                blk = hot_base + sa * H + hx
                bn = int(off[blk + 1] - off[blk])
                if bn <= 0:
                    continue
                spread = min(rounds, max(1, bn // (2 * MIN_RUN)))
                first = int(hash64(seed, 11, (sa << 32) | int(hot_items[hx]))) % rounds
                for sl in range(spread):
                    rnd = (first + sl * rounds // spread) % rounds
                    lo, hi = int(off[blk]) + bn * sl // spread, int(off[blk]) + bn * (sl + 1) // spread
                    n = hi - lo
                    if n <= 0:
                        continue
                    pieces = -(-n // chunk)
                    w = 1.0 if pieces <= 1 else min(1.0, float(np.float32(np.float32(max(boost, 1.0)) / np.float32(pieces))))
                    for pc in range(pieces):
                        st = lo + n * pc // pieces
                        visits.setdefault((sa * rounds + rnd) * IB + ib, []).append((st, lo + n * (pc + 1) // pieces - st, int(hot_items[hx]), w))
    out, first_of = [], [0]
    for v in range(mu * rounds * IB):
        runs = sorted(visits.get(v, []), key=lambda r: -r[1])          # Python's sort is stable
        out += runs
        first_of.append(len(out))
    return out, first_of


@pytest.mark.parametrize("threads", ["1", "3", "8", ""])
def test_threaded_plan_is_the_serial_plan_bit_for_bit(threads, monkeypatch):
    """Large plans are cut by several host threads (two-pass counting sort); whatever their number, the runs, their order inside
    every visit and the visit offsets are those of the one-thread restatement. 20 000 buckets: above the planner's own threshold
    for going parallel, so the empty setting exercises its automatic choice."""
    if threads:
        monkeypatch.setenv("MFSGD_PLAN_THREADS", threads)
    else:
        monkeypatch.delenv("MFSGD_PLAN_THREADS", raising=False)
    rng = np.random.default_rng(5)
    mu, IB, rounds, chunk, H = 4, 3, 4, 96, 5000
    hot_items = np.sort(rng.choice(60_000, H, replace=False)).astype(np.int32)
    hot_block_lo = np.concatenate([[0], np.sort(rng.choice(np.arange(1, H), IB - 1, replace=False)), [H]]).astype(np.int32)
    sizes_cold = rng.integers(0, 50, (mu, IB))
    sizes_hot = rng.choice([0, 1, 16, 31, 32, 64, 65, 95, 96, 97, 200, 500, 1000, 4000], (mu, H))
    off, start, count, item, weight, visits = plan(sizes_cold, sizes_hot, mu, H, IB, hot_block_lo, hot_items, rounds, chunk, seed=99, member=2, boost=1.25)
    want, first_of = _plan_restated(off, mu, H, IB, hot_block_lo, hot_items, rounds, chunk, 99, 2, 1.25)
    assert list(visits) == first_of and len(start) == len(want)
    assert np.array_equal(start, np.array([r[0] for r in want], np.int64))
    assert np.array_equal(count, np.array([r[1] for r in want], np.int32))
    assert np.array_equal(item, np.array([r[2] for r in want], np.int32))
    assert np.array_equal(weight, np.array([r[3] for r in want], np.float32))


def test_random_plans_match_the_restatement(monkeypatch):
    """60 random shapes (sub-stripes, item blocks, rounds, run lengths from 1 to 4096, empty and huge buckets, every merge
    boost) under random thread counts, each compared with the plain-Python restatement."""
    rng = np.random.default_rng(2)
    for _ in range(60):
        mu, IB = int(rng.integers(1, 6)), int(rng.integers(1, 5))
        rounds, chunk = int(rng.choice([1, 2, 3, 4, 8, 16])), int(rng.choice([1, 2, 16, 31, 32, 96, 256, 1024, 4096]))
        H = int(rng.integers(max(IB, 2), 60))
        hot_items = np.sort(rng.choice(100_000, H, replace=False)).astype(np.int32)
        cuts = np.sort(rng.choice(np.arange(1, H), IB - 1, replace=False)) if IB > 1 else np.array([], int)
        hot_block_lo = np.concatenate([[0], cuts, [H]]).astype(np.int32)
        sizes_hot = rng.choice([0, 1, 2, 15, 16, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1000, 5000], (mu, H))
        if chunk < 16:
            sizes_hot = np.minimum(sizes_hot, 200)
        boost, seed, member = float(rng.choice([1.0, 1.25, 1.9])), int(rng.integers(0, 2 ** 40)), int(rng.integers(0, 8))
        monkeypatch.setenv("MFSGD_PLAN_THREADS", str(int(rng.choice([1, 2, 5, 8]))))
        off, start, count, item, weight, visits = plan(rng.integers(0, 20, (mu, IB)), sizes_hot, mu, H, IB, hot_block_lo, hot_items, rounds,
                                                       chunk, seed=seed, member=member, boost=boost)
        want, first_of = _plan_restated(off, mu, H, IB, hot_block_lo, hot_items, rounds, chunk, seed, member, boost)
        assert list(visits) == first_of and len(start) == len(want), (mu, IB, rounds, chunk, H)
        assert np.array_equal(start, np.array([r[0] for r in want], np.int64)) and np.array_equal(count, np.array([r[1] for r in want], np.int32))
        assert np.array_equal(item, np.array([r[2] for r in want], np.int32)) and np.array_equal(weight, np.array([r[3] for r in want], np.float32))


def test_plan_arguments_are_checked():
    n = C.c_int64(0)
    v = np.zeros(2, np.int32)
    off = np.zeros(2, np.int64)
    hbl = np.zeros(2, np.int32)
    assert capi.lib.mfsgd_plan_runs(None, 1, 0, 1, capi.ptr(hbl), None, 1, 256, 0, 0, 1.0, None, None, None, None, C.byref(n), capi.ptr(v)) == capi.E_INVALID_ARG
    assert capi.lib.mfsgd_plan_runs(capi.ptr(off), 1, 0, 1, capi.ptr(hbl), None, 0, 256, 0, 0, 1.0, None, None, None, None, C.byref(n), capi.ptr(v)) == capi.E_INVALID_ARG
    assert capi.lib.mfsgd_plan_runs(capi.ptr(off), 1, 0, 1, capi.ptr(hbl), None, 1, 256, 0, 0, 1.0, None, None, None, None, C.byref(n), capi.ptr(v)) == capi.OK
    assert n.value == 0


# ------------------------------------------------------------------------------------------------
# automatic layout (plan_blocking / plan_rounds / plan_run_length) for the BASELINE.json shapes
# ------------------------------------------------------------------------------------------------
import matrixfactorizationsgd.java_b200 as mf

L2 = 132_644_864          # cudaDeviceProp::l2CacheSize of the pool's B200s (profiles/l2_peak.json: 132.6 MB)


def layout(name, G=1, world=1, resident_ctas=592, **kw):
    w = mf.WORKLOADS[name]
    mode = capi.MODE_DSGD if G > 1 else capi.MODE_HOGWILD
    cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, mode=mode, n_gpus=G, world_size=world, rank=0,
                         nccl_id=bytes(128) if world > 1 else None, **kw)
    n_train = int(w.n_ratings * 0.9)
    out = [C.c_int32(0) for _ in range(4)]
    capi.check(capi.lib.mfsgd_plan_layout(C.byref(cfg), L2, n_train // G, w.n_users // G, n_train // G, resident_ctas,
                                          *[C.byref(x) for x in out]))
    return tuple(x.value for x in out)      # (stripes, shards, rounds, run length)


def test_layout_of_the_baseline_shapes():
    # Netflix-shaped on 1 GPU: 246 MB of P in 4 L2-resident sub-stripes, 4 interleaved rounds, runs as long as a
    # sub-warp's share of the launch (round 2: long runs, few merges per item)
    assert layout("netflix") == (4, 1, 4, 960) and layout("netflix", resident_ctas=444) == (4, 1, 4, 1024)
    # one process per GPU: the rotation is pipelined over 2 item sub-shards, runs shorten with the launches
    # (two stream lanes overlap the sub-shards' launches, so a run may be as long as a sub-warp's share of two launches -- unless
    # that is still under the floor of 256: launches that small get runs of one share, down to 128; measured on 8 B200s)
    assert layout("netflix", G=2, world=2) == (2, 2, 4, 480)
    assert layout("netflix", G=4, world=4) == (1, 2, 2, 480)
    assert layout("netflix", G=8, world=8) == (1, 2, 1, 128) and layout("netflix", G=8, world=8, resident_ctas=444) == (1, 2, 1, 320)
    # a single process driving 8 devices (peer copies, no pipelining): one shard group per member
    assert layout("netflix", G=8, world=1) == (1, 1, 1, 256)
    assert layout("ml20m") == (2, 1, 4, 384)
    # 90 K ratings: one sub-stripe, but still 4 launches per epoch of >= 16 K records each; shortest runs (256: small
    # launches leave sub-warps idle rather than cut their items into pieces)
    assert layout("ml100k") == (1, 1, 4, 256)
    # the large shapes on their own configuration (8 GPUs) and squeezed onto one
    assert layout("yahoo", G=8, world=8) == (2, 2, 2, 416) and layout("powerlaw", G=8, world=8) == (7, 2, 1, 352)
    assert layout("yahoo") == (70, 1, 4, 384) and layout("powerlaw") == (193, 1, 4, 256)


def test_layout_overrides_and_modes():
    assert layout("netflix", stripes_per_gpu=7, shards_per_gpu=3, rounds=11, hot_chunk=100) == (7, 3, 11, 100)
    w = mf.WORKLOADS["ml100k"]
    cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, mode=capi.MODE_DETERMINISTIC)
    out = [C.c_int32(0) for _ in range(4)]
    capi.check(capi.lib.mfsgd_plan_layout(C.byref(cfg), L2, 90_000, w.n_users, 0, 592, *[C.byref(x) for x in out]))
    assert (out[0].value, out[1].value, out[2].value) == (1, 1, 1)          # parity mode: one block, one pass
    bad = mf.make_config(10, 10, 6, 0.1, 0.1)
    assert capi.lib.mfsgd_plan_layout(C.byref(bad), L2, 1, 1, 1, 1, *[C.byref(x) for x in out]) == capi.E_INVALID_ARG
