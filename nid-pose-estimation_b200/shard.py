"""Whole-problem sharding across the GPUs of one box.

One evaluation of one frame pair does not shard profitably (3.7 MB, tens of microseconds), and the
reference has no multi-GPU path at all (device 0 is hard-coded, g2o/g2o/core/computeH.cu:378). The
multi-GPU unit is therefore an independent *problem*: a frame pair of a tracking sequence (BASELINE config
4) or a perturbed pose of a cost-surface sweep (config 5). Every rank owns whole problems, runs them on
its own `nid_ctx`, and only the per-problem results travel: one all-gather of a few doubles per problem
(NCCL on the GPU box, gloo in the CPU tests). There is no data-path collective.

Problem p goes to rank `p % world` ("cyclic": neighbouring frames of a sequence land on different GPUs, so
a sequence whose later pairs are harder does not load one GPU), or to contiguous blocks ("block": lets a
rank reuse the reference frame of consecutive pairs).
"""
from __future__ import annotations

import numpy as np


def owned(n_problems: int, world: int, rank: int, scheme: str = "cyclic") -> np.ndarray:
    """Global indices of the problems rank `rank` owns, ascending."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    if n_problems < 0:
        raise ValueError("n_problems < 0")
    if scheme == "cyclic":
        return np.arange(rank, n_problems, world, dtype=np.int64)
    if scheme == "block":
        base, rem = divmod(n_problems, world)
        lo = rank * base + min(rank, rem)
        return np.arange(lo, lo + base + (1 if rank < rem else 0), dtype=np.int64)
    raise ValueError(f"unknown scheme {scheme!r}")


def max_owned(n_problems: int, world: int) -> int:
    return -(-n_problems // world)


def gather_results(local: np.ndarray, n_problems: int, world: int, rank: int, scheme: str = "cyclic",
                   device=None) -> np.ndarray:
    """All-gather per-problem result rows. `local[i]` belongs to problem `owned(...)[i]`; returns the
    [n_problems, width] table in global problem order on every rank. Ragged shards are padded to the
    longest one for the collective and the padding dropped afterwards."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if local.ndim == 1:
        local = local[:, None]
    mine = owned(n_problems, world, rank, scheme)
    if local.shape[0] != mine.size:
        raise ValueError(f"rank {rank} owns {mine.size} problems but holds {local.shape[0]} result rows")
    width = local.shape[1]
    out = np.full((n_problems, width), np.nan)
    if world == 1:
        out[mine] = local
        return out
    import torch
    import torch.distributed as dist
    cap = max_owned(n_problems, world)
    buf = torch.full((cap, width), float("nan"), dtype=torch.float64)
    buf[:mine.size] = torch.from_numpy(local)
    if device is not None:
        buf = buf.to(device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    for r in range(world):
        idx = owned(n_problems, world, r, scheme)
        out[idx] = parts[r][:idx.size].cpu().numpy()
    return out


def run_sharded(n_problems: int, world: int, rank: int, run_local, scheme: str = "cyclic", device=None) -> np.ndarray:
    """`run_local(indices) -> [len(indices), width]` on this rank's problems, then gather."""
    mine = owned(n_problems, world, rank, scheme)
    res = run_local(mine) if mine.size else np.zeros((0, 1))
    res = np.asarray(res, dtype=np.float64)
    if res.ndim == 1:
        res = res[:, None]
    if world > 1:
        # ranks with no problems still need the row width for the collective
        import torch
        import torch.distributed as dist
        w = torch.tensor([res.shape[1] if mine.size else 0], dtype=torch.int64)
        if device is not None:
            w = w.to(device)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        if not mine.size:
            res = np.zeros((0, int(w.item())))
    return gather_results(res, n_problems, world, rank, scheme, device)
