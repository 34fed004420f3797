"""One-process-per-GPU DSGD ring: torch.distributed is used for the bootstrap only (shipping the
ncclUniqueId, summing the RMSE partials, assembling P and Q on the host). The Q-shard rotation itself
runs inside libmfsgd.so (ncclSend/ncclRecv on the engine's stream); nothing here touches the data path.
"""
import numpy as np

from . import _capi as capi
from .engine import Engine, make_config, nccl_unique_id


def ring_schedule(n_gpus):
    """held[s][g] = Q shard group ring member g trains on in sub-epoch s (and who it swaps with).

    Member g starts an epoch holding group g; after every sub-epoch it sends its group to g-1 and
    receives from g+1, so held[s][g] = (g + s) % G. Mirrors rotate_q() in csrc/engine.cu."""
    return [[(g + s) % n_gpus for g in range(n_gpus)] for s in range(n_gpus)]


def check_schedule(schedule):
    """DSGD invariants: within a sub-epoch no two members hold the same group; over an epoch every
    member sees every group exactly once."""
    n = len(schedule)
    for row in schedule:
        if sorted(row) != list(range(n)):
            return False
    for g in range(n):
        if sorted(schedule[s][g] for s in range(n)) != list(range(n)):
            return False
    return True


def broadcast_unique_id(dist, rank, make_id=nccl_unique_id):
    """Rank 0 creates the ncclUniqueId, everyone receives the same 128 bytes."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.tensor(list(make_id()), dtype=torch.uint8)
    dist.broadcast(buf, src=0)
    return bytes(buf.tolist())


def create_rank_engine(dist, rank, world_size, device, **cfg_kw):
    """Engine for ring member `rank` of a world_size-process ring (one GPU each)."""
    nid = broadcast_unique_id(dist, rank)
    cfg = make_config(mode=capi.MODE_DSGD, n_gpus=world_size, world_size=world_size, rank=rank, device=device,
                      nccl_id=nid, **cfg_kw)
    return Engine(cfg)


def reduce_rmse(dist, sse, n):
    """Total RMSE from the per-rank partial sums returned by Engine.rmse_heldout()/rmse_train()."""
    import torch
    t = torch.tensor([float(sse), float(n)], dtype=torch.float64)
    dist.all_reduce(t)
    return float(np.sqrt(t[0].item() / t[1].item())) if t[1].item() > 0 else 0.0


def assemble_factors(dist, P_local, Q_local):
    """Every rank's get_factors() fills only its own rows (others zero): the sum is the whole matrix."""
    import torch
    P, Q = torch.from_numpy(P_local.copy()), torch.from_numpy(Q_local.copy())
    dist.all_reduce(P)
    dist.all_reduce(Q)
    return P.numpy(), Q.numpy()
