"""world_size-2 gloo tests (CPU) of the host-side ring plumbing: unique-id broadcast, RMSE reduction,
factor assembly, and the rotation schedule's DSGD invariants. The data path itself (NCCL inside
libmfsgd.so) needs GPUs and is covered by tests/test_gpu_multi.py."""
import os

import numpy as np
import pytest

from matrixfactorizationsgd.java_b200 import ring


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_rotation_schedule_invariants(G):
    sched = ring.ring_schedule(G)
    assert ring.check_schedule(sched)
    for s in range(G):
        for g in range(G):
            # member g receives in sub-epoch s+1 what member g+1 held in sub-epoch s
            assert sched[(s + 1) % G][g] == sched[s][(g + 1) % G] or s + 1 == G
    assert sched[0] == list(range(G))                       # every epoch starts with Q shards at home


def test_schedule_checker_rejects_collisions():
    assert not ring.check_schedule([[0, 0], [1, 1]])
    assert not ring.check_schedule([[0, 1], [0, 1]])


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nid = ring.broadcast_unique_id(dist, rank, make_id=lambda: bytes(range(128)))
        assert nid == bytes(range(128))
        # partial sums -> total RMSE: rank r contributes sse = r + 1 over n = 10 records
        total = ring.reduce_rmse(dist, float(rank + 1), 10)
        assert abs(total - np.sqrt(sum(range(1, world + 1)) / (10.0 * world))) < 1e-12
        assert ring.reduce_rmse(dist, 0.0, 0) == 0.0
        # each rank owns a row stripe; zero elsewhere -> the sum is the whole matrix
        P = np.zeros((4 * world, 3), dtype=np.float32)
        Q = np.zeros((2 * world, 3), dtype=np.float32)
        P[4 * rank:4 * rank + 4] = rank + 1
        Q[2 * rank:2 * rank + 2] = 10 * (rank + 1)
        Pa, Qa = ring.assemble_factors(dist, P, Q)
        for r in range(world):
            assert np.all(Pa[4 * r:4 * r + 4] == r + 1) and np.all(Qa[2 * r:2 * r + 2] == 10 * (r + 1))
        open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_ring_plumbing_gloo(world, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, 29611, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))
