"""One process per GPU: shard a batch of independent paths across ranks and gather per-path summaries.

Paths never read each other's data, so the data path has NO collective: every rank profiles
its own contiguous shard and keeps the trajectories resident on its GPU.  The only exchange is an
all_gather of the [B_local, 5] fp64 summary rows (n_out, total_length, t_end, max|v|, status) -- 40 B/path --
over NCCL (NVLink 5 / NVSwitch) on GPU boxes, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .packing import PackedPaths


def shard_bounds(B: int, world: int, weights: np.ndarray | None = None) -> List[Tuple[int, int]]:
    """Contiguous shards; with `weights` (e.g. chord-length sums ~ distance samples) the cut points balance
    the summed weight instead of the path count."""
    if weights is None:
        cuts = [(B * r) // world for r in range(world + 1)]
    else:
        w = np.asarray(weights, dtype=np.float64)
        c = np.concatenate([[0.0], np.cumsum(w)])
        targets = c[-1] * np.arange(world + 1) / world
        cuts = [int(np.searchsorted(c, t, side="left")) for t in targets]
        cuts[0], cuts[-1] = 0, B
        for r in range(1, world + 1):
            cuts[r] = max(cuts[r], cuts[r - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def chord_weights(p: PackedPaths) -> np.ndarray:
    """Sum of chord lengths per path: cheap proxy for the number of distance samples D."""
    d = np.diff(p.node_attr[:, :, 0:2], axis=1)
    valid = (np.arange(1, p.N_max)[None, :] < p.n_nodes[:, None])
    return (np.hypot(d[:, :, 0], d[:, :, 1]) * valid).sum(axis=1)


def local_shard(p: PackedPaths, rank: int, world: int, balance: bool = True) -> Tuple[PackedPaths, Tuple[int, int]]:
    lo, hi = shard_bounds(p.B, world, chord_weights(p) if balance else None)[rank]
    return p.slice(lo, hi), (lo, hi)


def gather_summaries(summary: torch.Tensor, counts: List[int] | None = None) -> torch.Tensor:
    """all_gather the per-path summary rows of every rank into [B_total, 5] (same order as the shards)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return summary
    world = dist.get_world_size()
    n_local = torch.tensor([summary.shape[0]], dtype=torch.int64, device=summary.device)
    if counts is None:
        ns = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(ns, n_local)
        counts = [int(n.item()) for n in ns]
    m = max(counts)
    pad = torch.zeros((m, summary.shape[1]), dtype=summary.dtype, device=summary.device)
    pad[: summary.shape[0]] = summary
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
