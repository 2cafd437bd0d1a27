"""'Next' row f3 of SURVEY.md section 8: the interactive geometry queries of the GUI canvas, served by one batched
device evaluation each instead of 25*N + 501 scalar Python calls per mouse move.

  preview_polyline            gui/path.py:370-386   25*N points of the path in pixels
  find_closest_point_on_path  gui/path.py:658-727   coarse pass over 25*N + 1 parameters, fine pass over 501
"""
from __future__ import annotations

import math

import numpy as np

from .packing import PX_TO_FT


def preview_polyline(spline_manager, n_nodes: int) -> np.ndarray:
    """Pixel coordinates of np.linspace(0, N-1, 25*N) along the path (update_spline, gui/path.py:370-384)."""
    t = np.linspace(0, n_nodes - 1, 25 * n_nodes)
    pts = np.asarray(spline_manager.get_point_at_parameter(t), dtype=np.float64)
    return (pts / PX_TO_FT + 0.5) * 2000


def find_closest_point_on_path(spline_manager, point_px, n_nodes: int):
    """(closest point in pixels, parameter) with the reference's two-pass search and first-strict-minimum rule."""
    point = (np.asarray(point_px, dtype=np.float64) / 2000 - 0.5) * PX_TO_FT
    min_dist, closest_point, closest_percent, closest_parameter = float("inf"), None, 0.0, 0.0

    def scan(percents):
        nonlocal min_dist, closest_point, closest_percent, closest_parameter
        params = np.array([spline_manager.percent_to_parameter(p) for p in percents], dtype=np.float64)
        pts = np.asarray(spline_manager.get_point_at_parameter(params), dtype=np.float64)     # one device launch
        for pc, pa, pt in zip(percents, params, pts):
            dist = math.hypot(pt[0] - point[0], pt[1] - point[1])
            if dist < min_dist:
                min_dist, closest_point, closest_percent, closest_parameter = dist, pt, pc, pa

    num_steps = 25 * n_nodes
    scan([i / num_steps for i in range(num_steps + 1)])
    coarse_percent = closest_percent
    start_percent = max(0.0, coarse_percent - 0.02)
    end_percent = min(1.0, coarse_percent + 0.02)
    percent_step = (end_percent - start_percent) / 500
    scan([start_percent + (i * percent_step) for i in range(501)])
    closest_px = (closest_point / PX_TO_FT + 0.5) * 2000
    return closest_px, closest_parameter
