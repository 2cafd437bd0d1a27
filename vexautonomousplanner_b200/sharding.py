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


def chord_weights(p: PackedPaths, dd: float = 0.005, dt: float = 0.01) -> np.ndarray:
    """Work estimate per path in rows: distance samples D ~ (sum of chord lengths) / dd, plus the rows
    the time stage inserts for waits (int(wait / dt) per node / action point) and turn-in-place profiles (trapezoid of
    arc |turn| * w / 2: never longer than 2 V / A + arc / V).  For plain paths this is proportional to the chord sum."""
    d = np.diff(p.node_attr[:, :, 0:2], axis=1)
    idx = np.arange(p.N_max)[None, :]
    valid = idx[:, 1:] < p.n_nodes[:, None]
    w = np.nan_to_num((np.hypot(d[:, :, 0], d[:, :, 1]) * valid).sum(axis=1) / dd, nan=0.0, posinf=0.0)
    node_ok = idx < p.n_nodes[:, None]
    waits = (np.floor(np.clip(p.node_attr[:, :, 3], 0, None) / dt) * node_ok).sum(axis=1)
    ap_ok = np.arange(p.A_max)[None, :] < p.n_ap[:, None]
    waits = waits + (np.floor(np.clip(p.ap_attr[:, :, 1], 0, None) / dt) * ap_ok).sum(axis=1)
    V, A, tw = p.cons[:, 0:1], p.cons[:, 1:2], p.cons[:, 5:6]
    turn = np.abs(p.node_attr[:, :, 2]) * node_ok
    with np.errstate(divide="ignore", invalid="ignore"):
        rows = np.where(turn != 0, np.ceil((2 * V / A + (turn * (np.pi / 180.0) * tw / 2) / V) / dt) + 3, 0.0)
    rows = np.nan_to_num(rows, nan=0.0, posinf=0.0).sum(axis=1)
    return w + np.minimum(waits + rows, 1.0e7)


def local_shard(p: PackedPaths, rank: int, world: int, balance: bool = True) -> Tuple[PackedPaths, Tuple[int, int]]:
    lo, hi = shard_bounds(p.B, world, chord_weights(p) if balance else None)[rank]
    return p.slice(lo, hi), (lo, hi)


class SummaryGatherer:
    """The one exchange of the sharded job, kept off the critical path: all_gather_into_tensor of the [B_local, 5] summary
    rows into a PREALLOCATED [world, B_pad, 5] buffer on a side stream.  submit() only snapshots the rows into a ring slot
    on the producing stream (a 40 B / path device copy) and lets the side stream wait for that; the compute streams never
    wait for the collective.  wait() joins before the gathered rows are read.  Works with NCCL (GPU) and gloo (CPU tests:
    no streams, the calls are synchronous)."""

    def __init__(self, counts: List[int], device, dtype=torch.float64, slots: int = 4):
        self.counts, self.world = list(counts), len(counts)
        self.B_pad = max(self.counts) if self.counts else 0
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.slots = [torch.zeros((self.B_pad, 5), dtype=dtype, device=self.device) for _ in range(slots)]
        self.outs = [torch.empty((self.world, self.B_pad, 5), dtype=dtype, device=self.device) for _ in range(slots)]
        self.side = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.done = [None] * slots               # per slot: the gather that last read it has finished
        self.k = 0
        self.last = None

    def submit(self, summary: torch.Tensor) -> int:
        """Enqueue the gather of one step's local rows; returns the slot whose `outs[slot]` will hold every rank's rows."""
        k = self.k % len(self.slots)
        self.k += 1
        if self.cuda and self.done[k] is not None:
            torch.cuda.current_stream(self.device).wait_event(self.done[k])      # the slot's previous gather has read it
        self.slots[k][: summary.shape[0]].copy_(summary, non_blocking=True)
        if self.world == 1:
            self.outs[k][0].copy_(self.slots[k])
        elif self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.side):
                self.side.wait_event(ev)
                dist.all_gather_into_tensor(self.outs[k].view(-1, 5), self.slots[k])
                self.done[k] = torch.cuda.Event()
                self.done[k].record(self.side)
        else:
            dist.all_gather_into_tensor(self.outs[k].view(-1, 5), self.slots[k])
        self.last = k
        return k

    def wait(self) -> None:
        if self.cuda and self.side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.side)

    def rows(self, slot: int | None = None) -> torch.Tensor:
        """[B_total, 5] rows of every rank in shard order (after wait())."""
        k = self.last if slot is None else slot
        return torch.cat([self.outs[k][r, :c] for r, c in enumerate(self.counts)], dim=0)


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
