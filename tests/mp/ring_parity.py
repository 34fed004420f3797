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
    if rank == 0:
        Po, Qo = orc.factorize(cu, ci, cr, m, m, 128, 0.02, 0.03, 3, seed, orc.ORDER_WARP_TREE_FMA)
        out["conflict_free_bit_exact"] = bool(np.array_equal(P, Po) and np.array_equal(Q, Qo))
    # 2. convergence parity on the mid-size workload
    eng = ring.create_rank_engine(dist, rank, world, local, n_users=nu, n_items=ni, k=k, lr=lr, lambda_=lam, seed=seed)
    eng.load_ratings(*tr)
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
        out.update({"gpu_rmse": got, "oracle_rmse": want, "rel": (got - want) / want,
                    "assembled_rmse": orc.rmse(P, Q, *ho), "partitions": parts, "n_train_total": int(info.n_train_total),
                    "epoch_ms": [s.epoch_ms for s in stats]})
        print("RING_PARITY " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
