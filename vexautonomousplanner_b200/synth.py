"""Synthetic workloads of BASELINE.json's configs.

Node pixels ~ U[15,1985]^2 (the GUI clamp, gui/node.py:159-181), converted with gui/path.py:365-367;
paths with two consecutive nodes closer than 30 px are redrawn (the reference's frac-wrap node detector,
motion_profile_generator.py:124, assumes less than one node per distance step).
"""
from __future__ import annotations

import numpy as np

from .packing import PackedPaths, pack_arrays, px_to_ft

FACTORY = (4.0, 8.0, 8.0, 0.8, 16.0, 12.5 / 12)      # src/config.yaml + ConfigManager defaults
REPO_ROOT = (4.0, 12.0, 12.0, 0.8, 43.0, 11.5 / 12)  # config.yaml at the repository root
CFG1_PX = np.array([[300, 300], [700, 500], [1000, 1200], [1400, 900], [1700, 1500], [1200, 1700]], dtype=np.float64)


def random_pixels(rng: np.random.Generator, B: int, N: int, min_sep: float = 30.0) -> np.ndarray:
    px = rng.uniform(15, 1985, (B, N, 2))
    while True:
        d = np.linalg.norm(np.diff(px, axis=1), axis=2)
        bad = (d < min_sep).any(axis=1)
        if not bad.any():
            return px
        px[bad] = rng.uniform(15, 1985, (int(bad.sum()), N, 2))


def cfg1(constraints=REPO_ROOT) -> PackedPaths:
    """Single 6-node path on the 2000x2000 field."""
    return pack_arrays(px_to_ft(CFG1_PX)[None], constraints)


def random_paths(B: int, N: int, seed: int, constraints=FACTORY) -> PackedPaths:
    """cfg2 (B=4096, N=8, seed 0) and cfg3 (B=2**20, N=16, seed 1): plain random-node paths, no actions."""
    rng = np.random.default_rng(seed)
    return pack_arrays(None, constraints, points_px=random_pixels(rng, B, N))


def mixed_paths(B: int, N: int = 8, seed: int = 3) -> PackedPaths:
    """cfg5: turn / wait / reverse / stop node actions, action points, per-path constraints; the second half of
    the batch is the mirror image of the first (gui/path.py:596-600)."""
    rng = np.random.default_rng(seed)
    H = (B + 1) // 2
    px = random_pixels(rng, H, N)
    idx = np.arange(N)[None, :]
    interior = (idx >= 1) & (idx <= N - 2)
    not_last = idx <= N - 2
    turn = np.where(interior & (rng.random((H, N)) < 0.15),
                    rng.choice([30.0, -30.0, 45.0, -45.0, 90.0, -90.0, 135.0, -135.0], (H, N)), 0.0)
    stop = interior & (rng.random((H, N)) < 0.1)
    rev = not_last & (rng.random((H, N)) < 0.1)
    wait = np.where(not_last & (rng.random((H, N)) < 0.15), rng.choice([0.1, 0.25, 0.5], (H, N)), 0.0)
    mv = np.where(rng.random((H, N)) < 0.1, rng.uniform(1.5, 3.5, (H, N)), 0.0)
    ma = np.where(rng.random((H, N)) < 0.1, rng.uniform(3.0, 7.0, (H, N)), 0.0)
    A = 2
    n_ap = rng.integers(0, A + 1, H).astype(np.int32)
    ap_t = np.sort(rng.uniform(0.2, N - 1.2, (H, A)), axis=1)
    ap_stop = rng.random((H, A)) < 0.1
    ap_wait = np.where(rng.random((H, A)) < 0.15, rng.choice([0.1, 0.25, 0.5], (H, A)), 0.0)
    ap_mv = np.where(rng.random((H, A)) < 0.1, rng.uniform(1.5, 3.5, (H, A)), 0.0)
    ap_ma = np.where(rng.random((H, A)) < 0.1, rng.uniform(3.0, 7.0, (H, A)), 0.0)
    cons = np.zeros((H, 6))
    cons[:, 0] = rng.uniform(2.5, 5.5, H)
    cons[:, 1] = rng.uniform(5.0, 14.0, H)
    cons[:, 2] = cons[:, 1]
    cons[:, 3] = 0.8
    cons[:, 4] = 16.0
    cons[:, 5] = rng.uniform(9.0, 15.0, H) / 12
    turn[:, 0] = 0.0                      # never: turn at node 0, turn/reverse at the last node (reference crashes)
    first = pack_arrays(None, cons, reverse=rev, stop=stop, turn=turn, wait=wait, max_velocity=mv,
                        max_acceleration=ma, ap_t=ap_t, ap_stop=ap_stop, ap_wait=ap_wait, ap_max_velocity=ap_mv,
                        ap_max_acceleration=ap_ma, n_ap=n_ap, points_px=px)
    second = first.mirrored()           # pixel-space mirror, turn -> -turn, nothing else (gui/path.py:596-600)
    cat = [np.concatenate([a, b])[:B] for a, b in zip(
        (first.node_attr, first.node_flags, first.n_nodes, first.ap_attr, first.ap_flags, first.n_ap, first.cons,
         first.points_px),
        (second.node_attr, second.node_flags, second.n_nodes, second.ap_attr, second.ap_flags, second.n_ap, second.cons,
         second.points_px))]
    return PackedPaths(*[np.ascontiguousarray(a) for a in cat[:7]], points_px=np.ascontiguousarray(cat[7]))


def long_path(N: int = 801, seed: int = 2, constraints=FACTORY) -> PackedPaths:
    """cfg4: one very long path (about 10^6 distance samples at dd = 0.005)."""
    rng = np.random.default_rng(seed)
    return pack_arrays(px_to_ft(random_pixels(rng, 1, N)), constraints)
