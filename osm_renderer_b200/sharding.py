"""Multi-GPU decomposition of a tile batch: tiles are independent, the dataset is replicated on every
GPU, tile i of the global request list goes to rank i mod N (the round-robin of the reference server's worker
dispatch, src/http_server.rs:105-108).  No collective touches the data path; torch.distributed is used only to
combine the timing / tile counters at the end of a run.
"""
from __future__ import annotations

import numpy as np


def shard_indices(n_items: int, rank: int, world: int) -> np.ndarray:
    """Indices of the global request list rendered by `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, n_items, world, dtype=np.int64)


def shard_batch(tiles, area_begin, areas, rank: int, world: int):
    """Slice a (tiles, area_begin, areas) batch down to the tiles of `rank` (styler order inside a tile is kept)."""
    idx = shard_indices(len(tiles), rank, world)
    area_begin = np.asarray(area_begin, dtype=np.int64)
    parts = [areas[area_begin[i] : area_begin[i + 1]] for i in idx]
    begins = np.zeros(len(idx) + 1, dtype=np.uint32)
    if len(idx):
        begins[1:] = np.cumsum([len(p) for p in parts])
    sub = np.concatenate(parts) if parts else areas[:0]
    return tiles[idx], begins, sub, idx


def weak_scaling_request_list(n_batch_tiles: int, world: int) -> np.ndarray:
    """Global request list of the weak-scaling benchmark: `world` interleaved copies of one batch, so the round-robin
    shard of every rank is exactly one full batch.  Entry = index into the batch."""
    return np.repeat(np.arange(n_batch_tiles, dtype=np.int64), world)


def reduce_job(dist, seconds: float, tiles: int, device=None):
    """(max seconds over ranks, total tiles over ranks).  `dist` is torch.distributed or None (single process)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(seconds), int(tiles)
    import torch

    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    n = torch.tensor([int(tiles)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item())
