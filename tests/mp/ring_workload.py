"""One BASELINE.json workload on a one-process-per-GPU DSGD ring (launched by torchrun from tests/test_gpu_multi.py or by hand):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P tests/mp/ring_workload.py yahoo [epochs]
Synthetic ratings are generated on every rank's device (no host arrays: the 2 B-record shape would need 24 GB of them), the
held-out RMSE is evaluated after every epoch and rank 0 prints one line `RING_WORKLOAD {json}` with the curve, the sequential
oracle's curve (tests/golden/oracle_rmse_<name>.json) and the DSGD-ordered one where the fixture has it."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    import matrixfactorizationsgd.java_b200 as mf
    from matrixfactorizationsgd.java_b200 import ring
    capi = mf.capi
    name = sys.argv[1]
    w = mf.WORKLOADS[name]
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else w.epochs
    kw = {}
    if os.environ.get("RW_SHAPE"):           # experiments: the same generator on another shape, "users,items,ratings" (no oracle curve then)
        nu_, ni_, nr_ = (int(x) for x in os.environ["RW_SHAPE"].split(","))
        w = w._replace(name=w.name + "@%dx%dx%d" % (nu_, ni_, nr_), n_users=nu_, n_items=ni_, n_ratings=nr_)
        name = "none"
    for env, key in (("RW_STRIPES", "stripes_per_gpu"), ("RW_ROUNDS", "rounds"), ("RW_SHARDS", "shards_per_gpu"), ("RW_HOT_CHUNK", "hot_chunk")):
        if os.environ.get(env):
            kw[key] = int(os.environ[env])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = ring.create_rank_engine(dist, rank, world, local, n_users=w.n_users, n_items=w.n_items, k=w.k, lr=w.lr, lambda_=w.lambda_,
                                  seed=mf.SEED, flags=capi.FLAG_TIME_KERNELS, **kw)
    t0 = time.time()
    nt, nh = eng.generate_synthetic(mf.synth_params_of(w))
    setup_s = time.time() - t0
    info = eng.layout_info()
    eng.init_factors()
    curve, epoch_ms = [], []
    for e in range(epochs):
        st = eng.train(1)
        epoch_ms.append(st[0].epoch_ms)
        _, sse, cnt = eng.rmse_heldout()
        curve.append(ring.reduce_rmse(dist, sse, cnt))
    import torch
    t = torch.tensor([float(nt), float(nh), max(epoch_ms[1:] or epoch_ms)], dtype=torch.float64)
    tot = t.clone()
    dist.all_reduce(tot)
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    eng.close()
    if rank == 0:
        ms = sorted(epoch_ms[1:] or epoch_ms)
        out = {"workload": w.name, "ring": world, "k": w.k, "epochs": epochs, "n_train": int(tot[0].item()), "n_heldout": int(tot[1].item()),
               "stripes_per_gpu": int(info.stripes_per_gpu), "shards_per_gpu": int(info.shards_per_gpu), "rounds": int(info.rounds),
               "run_length": int(info.run_length), "hot_items": int(info.n_hot_items), "heavy_users": int(info.n_heavy_users),
               "setup_s": setup_s, "epoch_ms_rank0": epoch_ms, "heldout_rmse_per_epoch": curve,
               "note": "epochs are trained one mfsgd_train call at a time (a held-out evaluation after each), so epoch_ms includes the ring's pipeline start-up"}
        med = ms[len(ms) // 2]
        out["gupdates_per_s_rank0_clock"] = int(tot[0].item()) / med / 1e6
        out["roofline_frac_measured_peak"] = int(tot[0].item()) / world / (med * 1e-3) * mf.bytes_per_update(w.k) / 6552.6e9
        ref = os.path.join(ROOT, "tests", "golden", "oracle_rmse_%s.json" % name)
        if os.path.exists(ref):
            fx = json.load(open(ref))
            want = fx.get("heldout_rmse_per_epoch", [])
            out["oracle_rmse_per_epoch"] = want
            out["n_train_oracle"] = fx.get("n_train")
            m = min(len(want), epochs)
            out["rel_to_shuffled_oracle"] = [curve[e] / want[e] - 1 for e in range(m)]
            key = "dsgd%d" % world
            if key in fx:
                out["oracle_dsgd_order_rmse_per_epoch"] = fx[key]["heldout_rmse_per_epoch"]
        print("RING_WORKLOAD " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
