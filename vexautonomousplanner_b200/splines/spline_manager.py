"""Mirror of src/splines/spline_manager.py (QuinticHermiteSplineManager, PathLookupTable) on the CUDA engine.

  build_path                     -> vap_build_path   (spline_manager.py:42-172)
  get_*_at_parameter             -> vap_eval         (:204-275)
  build_lookup_table             -> vap_build_lut    (:426-475)
  precompute_path_properties     -> vap_build_props  (:477-548)
  distance_to_time, get_heading,
  get_curvature                  -> vap_query_tables (:291-346, :550-580)
Error behaviour follows the reference: False on bad input, ValueError on an unbuilt object, IndexError for a turn or
reverse action on the last node (F7 in SURVEY.md).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from ..packing import pack_paths
from ..runtime import C, Keep, _p, check, dev, get_engine, stream
from .quintic_hermite_spline import QuinticHermiteSpline

logger = logging.getLogger(__name__)


@dataclass
class PathLookupTable:
    """Cache for quick parameter lookups based on distance"""
    distances: np.ndarray
    parameters: np.ndarray
    total_length: float


class QuinticHermiteSplineManager:
    """Manages multiple QuinticHermiteSplines to create a complete path through nodes with specific constraints."""

    def __init__(self):
        self.splines: List[QuinticHermiteSpline] = []
        self.nodes: List = []
        self.action_points: List = []
        self.path_parameters: Dict = {}
        self.arc_length = 0.0
        self.lookup_table: Optional[PathLookupTable] = None
        self._precomputed_properties: Optional[Dict] = None
        self._points: Optional[np.ndarray] = None
        self._db = None          # DeviceBatch (B = 1, default constraints)
        self._geo = None
        self._tables = None

    # ------------------------------------------------------------------ S0
    def build_path(self, points: np.ndarray, nodes: List, action_points: List) -> bool:
        if len(points) != len(nodes) or len(points) < 2:
            return False
        self.splines = []
        self.nodes = nodes
        self.action_points = action_points
        pts = np.asarray(points, dtype=np.float64)
        self._points = pts
        eng = get_engine()
        packed = pack_paths([(pts, nodes, action_points)], [0.0] * 6)
        self._db = eng.upload(packed)
        g = eng.build_geometry(self._db, with_params=True)
        st = int(g.status.item())
        if st == -2:
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (len(pts), len(pts)))
        if st != 0:
            logger.error(f"Failed to fit spline for points {pts}")
            return False
        self._geo = g
        S = int(g.n_splines.item())
        n = len(pts)
        first = g.first_node[0, : S + 1].cpu().numpy()
        seg = g.seg[0].cpu().numpy()
        seglen = g.seglen[0].cpu().numpy()
        params = g.params[0].cpu().numpy()
        derivs = g.derivs[0].cpu().numpy()
        pc = 0
        for k in range(S):
            a, b = int(first[k]), int(first[k + 1])
            m = b - a + 1
            sp = QuinticHermiteSpline()
            tangents = []
            for nd in nodes[a : b + 1]:
                if nd.tangent is not None:
                    tangents.append([nd.tangent * nd.incoming_magnitude, nd.tangent * nd.outgoing_magnitude])
                else:
                    tangents.append([None, None])
            sp.set_all_tangents(tangents)
            sp._load(pts[a : b + 1], seg[a:b], seglen[a:b], params[pc : pc + m], derivs[pc : pc + m])
            # boundary tangents as attributes (spline_manager.py:79-81,131-132,151-158): row 2 / 3 of the last segment
            if k > 0:
                sp.starting_tangent = np.array(sp.segments[-1][2])
            if k < S - 1:
                sp.ending_tangent = np.array(sp.segments[-1][3])
            self.splines.append(sp)
            pc += m
        self.arc_length = None
        self.lookup_table = None
        self._tables = None
        # The reference clears only lookup_table here (spline_manager.py:170-171) and would answer get_heading /
        # get_curvature from the PREVIOUS path's property tables until rebuild_tables() runs; the device tables are
        # keyed to this path's geometry, so the stale cache is dropped and the next query rebuilds it.
        self._precomputed_properties = None
        return True

    def _need(self):
        if not self.splines:
            raise ValueError("No splines have been initialized")

    # ------------------------------------------------------------------ evaluation
    def _eval(self, which: int, t):
        self._need()
        eng = get_engine()
        g = self._geo
        tt = np.atleast_1d(np.asarray(t, dtype=np.float64))
        out = eng._empty((tt.size, 2))
        k = Keep()
        check(eng.lib.vap_eval(C.c_int64(tt.size), k(torch.zeros(tt.size, dtype=torch.int32, device=eng.device)),
                               k(tt), C.c_int(which), C.c_int(self._db.N_max), _p(g.seg), _p(g.first_node),
                               _p(g.param_end), _p(g.n_splines), _p(out), stream()), "vap_eval")
        res = out.cpu().numpy()
        return res[0] if np.ndim(t) == 0 else res

    def get_point_at_parameter(self, t: float) -> np.ndarray:
        return self._eval(0, t)

    def get_derivative_at_parameter(self, t: float) -> np.ndarray:
        return self._eval(1, t)

    def get_second_derivative_at_parameter(self, t: float) -> np.ndarray:
        return self._eval(2, t)

    def _get_heading(self, t: float) -> float:
        return np.float64(self._eval(3, t)[0])

    def _get_curvature(self, t: float) -> float:
        """spline_manager.py:375-418 (speed_squared < 1e-10 -> 0)."""
        d = self._eval(1, t)
        if d[0] ** 2 + d[1] ** 2 < 1e-10:
            return 0.0
        return np.float64(self._eval(3, t)[1])

    def _map_parameter_to_spline(self, t: float) -> Tuple[int, float]:
        self._need()
        cumulative = 0
        for i, spline in enumerate(self.splines):
            num_points = len(spline.control_points)
            if t <= cumulative + num_points - 1 or i == len(self.splines) - 1:
                return i, t - cumulative
            cumulative += num_points - 1
        raise ValueError("Failed to map parameter to spline segment")

    def percent_to_parameter(self, percent: float):
        self._need()
        t = 0 + len(self.nodes) * (percent)
        t = max(t, 0)
        t = min(t, len(self.nodes) - 1)
        return t

    def get_magnitudes_at_parameter(self, idx):
        spline_idx, local_t = self._map_parameter_to_spline(idx)
        if self.nodes[idx].tangent is not None:
            return [self.nodes[idx].incoming_magnitude, self.nodes[idx].outgoing_magnitude]
        sp = self.splines
        if spline_idx == 0 and local_t == 0:
            return [0, sp[spline_idx].get_magnitude(0)]
        elif spline_idx == len(sp) - 1 and local_t == sp[spline_idx].percent_to_parameter(100):
            return [sp[spline_idx].get_magnitude(-1), 0]
        elif local_t == 0:
            return [sp[spline_idx - 1].get_magnitude(-1), sp[spline_idx].get_magnitude(0)]
        elif local_t == sp[spline_idx].percent_to_parameter(100):
            return [sp[spline_idx].get_magnitude(-1), sp[spline_idx + 1].get_magnitude(0)]
        return [sp[spline_idx].get_magnitude(round(local_t) - 1), sp[spline_idx].get_magnitude(round(local_t))]

    # ------------------------------------------------------------------ tables
    def build_lookup_table(self, min_samples=1000, max_samples=20000, tolerance=1e-6) -> None:
        if not self.splines:
            raise ValueError("No splines initialized")
        eng = get_engine()
        t = eng.build_lut(self._db, self._geo, samples=min_samples)
        if self._tables is not None and self._tables.prop_k is not None:
            t.spn, t.P_cap, t.prop_k, t.prop_h = self._tables.spn, self._tables.P_cap, self._tables.prop_k, self._tables.prop_h
        self._tables = t
        Q = min_samples * len(self.splines)
        self.lookup_table = PathLookupTable(distances=t.lut_d[0, :Q].cpu().numpy(), parameters=t.lut_t[0, :Q].cpu().numpy(),
                                            total_length=np.float64(t.total_len.item()))

    def precompute_path_properties(self, samples_per_node: int = 1000) -> None:
        if not self.splines:
            raise ValueError("No splines initialized")
        eng = get_engine()
        if self._tables is None:
            self._tables = eng.build_lut(self._db, self._geo)
            # the reference only materialises the LUT on demand; keep lookup_table as it was
        eng.build_props(self._db, self._geo, self._tables, spn=samples_per_node)
        P = len(self.nodes) * samples_per_node
        self._precomputed_properties = {
            "parameters": np.linspace(0, len(self.nodes) - 1, P),
            "curvatures": self._tables.prop_k[0, :P].cpu().numpy(),
            "headings": self._tables.prop_h[0, :P].cpu().numpy(),
        }

    def _query(self, what: int, x: float) -> float:
        eng = get_engine()
        if self._tables is None or (what != 0 and self._tables.prop_k is None):
            self.precompute_path_properties()
        t = self._tables
        out = eng._empty((1,))
        z = torch.zeros(1, dtype=torch.int32, device=eng.device)
        k = Keep()
        check(eng.lib.vap_query_tables(C.c_int64(1), _p(z), k([x]), C.c_int(what), _p(self._db.n_nodes),
                                       _p(self._geo.n_splines), C.c_int(t.samples), C.c_int64(t.Q_cap), _p(t.lut_d),
                                       _p(t.lut_t), _p(t.total_len), C.c_int(t.spn), C.c_int64(t.P_cap), _p(t.prop_k),
                                       _p(t.prop_h), _p(out), stream()), "vap_query_tables")
        return np.float64(out.item())

    def distance_to_time(self, distance: float) -> float:
        if self.lookup_table is None:
            self.build_lookup_table()
        if distance <= 0:
            return 0
        if distance >= self.lookup_table.total_length:
            return len(self.nodes) - 1
        return self._query(0, distance)

    def get_total_arc_length(self) -> float:
        self._need()
        if self.lookup_table is None:
            self.build_lookup_table()
        return self.lookup_table.total_length

    def get_heading(self, t: float) -> float:
        if self._precomputed_properties is None:
            self.precompute_path_properties()
        return self._query(1, t)

    def get_curvature(self, t: float) -> float:
        if self._precomputed_properties is None:
            self.precompute_path_properties()
        return self._query(2, t)

    def validate_path_continuity(self) -> bool:
        pass

    def rebuild_tables(self):
        self.build_lookup_table()
        self.precompute_path_properties()
