"""Stress check of the ring's Q rotation (one process per GPU; launched by torchrun):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tests/mp/ring_stress.py [epochs] [reps]
Conflict-free data (pairwise distinct users and items), so every launch is a few microseconds long and the ring runs at the speed
of its hand-overs: any slice that is read before it has arrived, or overwritten before it has left, changes the bits of the result.
The assembled factors must equal the sequential oracle's bit for bit after every repetition. Prints `RING_STRESS {json}`."""
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
    epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seed, m, k = mf.SEED, 20011, 128
    rng = np.random.default_rng(11)
    cu, ci = rng.permutation(m).astype(np.int32), rng.permutation(m).astype(np.int32)
    cr = (1 + 4 * rng.random(m)).astype(np.float32)
    want = orc.factorize(cu, ci, cr, m, m, k, 0.02, 0.03, epochs, seed, orc.ORDER_WARP_TREE_FMA) if rank == 0 else None
    bad = []
    eng = ring.create_rank_engine(dist, rank, world, local, n_users=m, n_items=m, k=k, lr=0.02, lambda_=0.03, seed=seed)
    for rep in range(reps):
        eng.load_ratings(cu, ci, cr)           # a reload of the same shape keeps the ring window (and its sequence numbers)
        eng.init_factors()
        for _ in range(4):                     # several train calls per repetition: the pipeline drains and restarts
            eng.train(epochs // 4)
        P, Q = ring.assemble_factors(dist, *eng.get_factors())
        if rank == 0:
            dp, dq = int((P != want[0]).sum()), int((Q != want[1]).sum())
            if dp or dq:
                bad.append({"rep": rep, "p_values_off": dp, "q_values_off": dq})
    eng.close()
    if rank == 0:
        print("RING_STRESS " + json.dumps({"world": world, "epochs": epochs // 4 * 4, "reps": reps, "bit_exact": not bad, "mismatches": bad,
                                           "transport": os.environ.get("MFSGD_RING_TRANSPORT", "window"),
                                           "signal": os.environ.get("MFSGD_RING_SIGNAL", "write")}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
