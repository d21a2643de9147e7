"""Multi-GPU layer: independent chains / restarts are sharded over ranks (one process per GPU);
the only exchange is one all_gather of each rank's best (sum of scores, chain id, site vector)
(the reference's own commented PSeq lines parallelise the same axis, fsx:430).

Backend: NCCL on GPUs (over NVLink / NVSwitch), gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def shard_chains(n_chains: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block of global chain ids for `rank`: (first id, count). Streams are keyed by the
    GLOBAL chain id, so the union of all ranks' results does not depend on world_size."""
    if n_chains < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad shard request")
    base, rem = divmod(n_chains, world_size)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def select_best(sums: np.ndarray, chain_ids: np.ndarray) -> int:
    """Index of the winner: largest sum (strict >), lowest global chain id on ties -- the order in
    which the reference's sequential restart loop would have met them (fs:450, fs:166)."""
    best = -1
    for i in range(len(sums)):
        if best < 0:
            best = i
            continue
        if sums[i] > sums[best] or (sums[i] == sums[best] and chain_ids[i] < chain_ids[best]):
            best = i
    return best


def allgather_best(local_sum: float, local_chain_id: int, local_sites: np.ndarray, local_scores: np.ndarray,
                   device: Optional[str] = None):
    """One all_gather of (sum, chain id, sites, scores) per rank; returns the global winner.

    Works without an initialised process group (single process) and with gloo (CPU) or nccl (GPU).
    Returns (sum, global chain id, sites int32[n], scores float64[n], owner rank).
    """
    import torch
    import torch.distributed as dist

    n = int(local_sites.shape[0])
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_sum), int(local_chain_id), np.asarray(local_sites, np.int32), np.asarray(local_scores, np.float64), 0
    world = dist.get_world_size()
    dev = torch.device(device) if device else (torch.device("cuda", torch.cuda.current_device())
                                                if dist.get_backend() == "nccl" else torch.device("cpu"))
    # one float64 message: [sum, chain id, scores[n], sites[n]] (ids and sites are exact in float64)
    msg = torch.empty(2 + 2 * n, dtype=torch.float64)
    msg[0] = float(local_sum)
    msg[1] = float(local_chain_id)
    msg[2:2 + n] = torch.from_numpy(np.ascontiguousarray(local_scores, np.float64))
    msg[2 + n:] = torch.from_numpy(np.ascontiguousarray(local_sites, np.int32).astype(np.float64))
    msg = msg.to(dev)
    out = torch.empty(world * (2 + 2 * n), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, msg)
    out = out.cpu().numpy().reshape(world, 2 + 2 * n)
    sums = out[:, 0]
    ids = out[:, 1].astype(np.int64)
    # a rank without chains reports id -1
    valid = [r for r in range(world) if ids[r] >= 0]
    w = valid[select_best(sums[valid], ids[valid])]
    return float(sums[w]), int(ids[w]), out[w, 2 + n:].astype(np.int32), out[w, 2:2 + n].copy(), w
