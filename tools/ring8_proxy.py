"""One GPU replaying the per-member work of an 8-member ring on the Netflix-shaped workload (no NCCL):
60 000 users (1/8), all 17 800 items in 8 item shards, 12.5 M ratings; MFSGD_FLAG_SPLIT_SHARDS makes every
launch pair one (P stripe 31 MB, Q shard 1.1 MB, ~1.4 M records) block, the size of one sub-epoch at G = 8.
usage: python tools/ring8_proxy.py [hot_share] [hot_chunk] [epochs]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import matrixfactorizationsgd.java_b200 as mf
capi = mf.capi
hot_share = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
hot_chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 6
w = mf.WORKLOADS["netflix"]
nu, n = w.n_users // 8, w.n_ratings // 8
cfg = mf.make_config(nu, w.n_items, w.k, w.lr, w.lambda_, seed=mf.SEED, mode=capi.MODE_HOGWILD, stripes_per_gpu=1, shards_per_gpu=int(os.environ.get("PROXY_SHARDS", "8")),
                     rounds=1, hot_share=hot_share, hot_chunk=hot_chunk,
                     flags=capi.FLAG_TIME_KERNELS | capi.FLAG_SPLIT_SHARDS)
with mf.Engine(cfg) as eng:
    eng.generate_synthetic(mf.synth_params(n, mf.SEED, w.log2_alpha_user, w.c_user, w.log2_alpha_item, w.c_item))
    eng.init_factors()
    eng.train(2)
    st = eng.train(epochs)
    info = eng.layout_info()
ms = float(np.median([s.epoch_ms for s in st]))
print(json.dumps({"hot_share_equiv": hot_share or 3e-5, "hot_chunk": hot_chunk, "hot_items": info.n_hot_items, "epoch_ms": ms, "per_subepoch_us": 1e3 * ms / 8, "shards": int(os.environ.get("PROXY_SHARDS", "8")),
                  "cold_us": 1e3 * float(np.median([s.cold_ms for s in st])) / 8, "hot_us": 1e3 * float(np.median([s.hot_ms for s in st])) / 8,
                  "gupdates_s": st[0].updates / ms / 1e6, "launches": st[0].update_launches}))
