"""Multi-process DSGD ring check (one process per GPU; launched by torchrun from tests/test_gpu_multi.py):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tests/mp/ring_parity.py
Every rank loads the same triplets, trains, and rank 0 compares the assembled result with the CPU oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch.distributed as dist
    import matrixfactorizationsgd.java_b200 as mf
    from matrixfactorizationsgd.java_b200 import ring
    import pyoracle as orc
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seed = mf.SEED
    nu, ni, n, k, lr, lam, epochs = 13_800, 2_700, 2_000_000, 32, 0.005, 0.05, 8
    u, i, r, held = orc.generate(seed, 0, n, nu, ni)
    tr = (u[~held].copy(), i[~held].copy(), r[~held].copy())
    ho = (u[held].copy(), i[held].copy(), r[held].copy())
    out = {"world": world}
    # 1. conflict-free data: the ring must reproduce the oracle bit for bit
    m = 5003
    rng = np.random.default_rng(3)
    cu, ci = rng.permutation(m).astype(np.int32), rng.permutation(m).astype(np.int32)
    cr = (1 + 4 * rng.random(m)).astype(np.float32)
    eng = ring.create_rank_engine(dist, rank, world, local, n_users=m, n_items=m, k=128, lr=0.02, lambda_=0.03, seed=seed)
    eng.load_ratings(cu, ci, cr)
    eng.init_factors()
    eng.train(3)
    P, Q = ring.assemble_factors(dist, *eng.get_factors())
    eng.close()
    # 1b. the same through the sharded load: every rank passes only its own (uneven) slice of the triplets; the records travel
    # to their stripe owners over NCCL (mfsgd_load_ratings_sharded). Same layout bounds, same factors, bit for bit.
    cuts = [0] + [int(m * (j + 1) ** 2 / world ** 2) for j in range(world)]          # uneven on purpose; the last rank gets most
    sl = slice(cuts[rank], cuts[rank + 1])
    eng = ring.create_rank_engine(dist, rank, world, local, n_users=m, n_items=m, k=128, lr=0.02, lambda_=0.03, seed=seed)
    eng.load_ratings_sharded(cu[sl], ci[sl], cr[sl])
    info_sh = eng.layout_info()
    eng.init_factors()
    eng.train(3)
    Psh, Qsh = ring.assemble_factors(dist, *eng.get_factors())
    eng.close()
    if rank == 0:
        Po, Qo = orc.factorize(cu, ci, cr, m, m, 128, 0.02, 0.03, 3, seed, orc.ORDER_WARP_TREE_FMA)
        out["conflict_free_bit_exact"] = bool(np.array_equal(P, Po) and np.array_equal(Q, Qo))
        out["sharded_load_bit_exact"] = bool(np.array_equal(Psh, Po) and np.array_equal(Qsh, Qo))
        out["sharded_n_train_total"] = int(info_sh.n_train_total)
    # 2. convergence parity on the mid-size workload
    eng = ring.create_rank_engine(dist, rank, world, local, n_users=nu, n_items=ni, k=k, lr=lr, lambda_=lam, seed=seed)
    nt = len(tr[2])
    mine = slice(nt * rank // world, nt * (rank + 1) // world)
    eng.load_ratings_sharded(tr[0][mine], tr[1][mine], tr[2][mine])       # each rank uploads 1/world of the set
    eng.load_heldout(*ho)
    eng.init_factors()
    part = eng.partition()
    stats = eng.train(epochs)
    _, sse, cnt = eng.rmse_heldout()
    got = ring.reduce_rmse(dist, sse, cnt)
    P, Q = ring.assemble_factors(dist, *eng.get_factors())
    info = eng.layout_info()
    eng.close()
    parts = [None] * world
    dist.all_gather_object(parts, part)
    if rank == 0:
        Po, Qo = orc.factorize(*tr, nu, ni, k, lr, lam, epochs, seed)
        want = orc.rmse(Po, Qo, *ho)
        # the sequential rule in this ring's own block order (the schedule's share of any deviation)
        ub, ib = orc.balanced_bounds(tr[0], nu, world), orc.balanced_bounds(tr[1], ni, world)
        Pd, Qd = orc.init_factors(nu, k, seed, 0), orc.init_factors(ni, k, seed, 1)
        for e in range(epochs):
            o = orc.dsgd_order(tr[0], tr[1], ub, ib, seed, e)
            orc.train(tr[0][o], tr[1][o], tr[2][o], Pd, Qd, lr, lam, e, e + 1, seed, shuffled=False)
        out.update({"gpu_rmse": got, "oracle_rmse": want, "oracle_dsgd_order_rmse": orc.rmse(Pd, Qd, *ho), "rel": (got - want) / want,
                    "assembled_rmse": orc.rmse(P, Q, *ho), "partitions": parts, "n_train_total": int(info.n_train_total),
                    "n_train": nt, "epoch_ms": [s.epoch_ms for s in stats], "stripes_per_gpu": int(info.stripes_per_gpu),
                    "shards_per_gpu": int(info.shards_per_gpu), "rounds": int(info.rounds)})
        print("RING_PARITY " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
