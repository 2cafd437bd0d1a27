"""Host-side packing of nodes / action points / constraints into the layouts of include/vap.h.

The reference hands the hot path duck-typed GUI objects (gui/node.py:17-51, gui/action_point.py:16-41);
only their attribute names matter.  `pack_paths` turns lists of such objects into the packed arrays the
CUDA engine consumes.  The rotation cos/sin of turn nodes are evaluated here with numpy exactly as
spline_manager.py:105-113 does (they are index-critical libm values).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

NA, APA = 12, 4
F_REVERSE, F_STOP, F_TANGENT = 1, 2, 4
(A_X, A_Y, A_TURN, A_WAIT, A_MAXVEL, A_MAXACC, A_TX, A_TY, A_INMAG, A_OUTMAG, A_RCOS, A_RSIN) = range(12)
PX_TO_FT = 12.1090395251      # gui/path.py:365-367


def px_to_ft(px):
    """Pixel -> feet conversion of PathWidget.update_spline (gui/path.py:365-367)."""
    px = np.asarray(px, dtype=np.float64)
    return (px / 2000 - 0.5) * PX_TO_FT


def rotation_table(turn_deg: np.ndarray, reverse: np.ndarray):
    """cos/sin of radians(turn) (+ pi when the node is also a reverse node), per node.

    Evaluated per distinct (turn, reverse) value through numpy scalars, i.e. through the very same
    np.radians / np.cos / np.sin calls the reference makes, then broadcast back.
    """
    turn_deg = np.asarray(turn_deg, dtype=np.float64)
    reverse = np.asarray(reverse, dtype=bool)
    rc = np.ones(turn_deg.shape)
    rs = np.zeros(turn_deg.shape)
    nz = turn_deg != 0
    if nz.any():
        key = np.stack([turn_deg[nz], reverse[nz].astype(np.float64)], axis=-1)
        uniq, inv = np.unique(key, axis=0, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        c = np.empty(len(uniq))
        s = np.empty(len(uniq))
        for i, (t, r) in enumerate(uniq):
            ang = np.radians(t)
            if r:
                ang = ang + np.pi
            c[i] = np.cos(ang)
            s[i] = np.sin(ang)
        rc[nz] = c[inv]
        rs[nz] = s[inv]
    return rc, rs


@dataclass
class PackedPaths:
    """A batch of B paths in the packed host layout (numpy, C-contiguous)."""
    node_attr: np.ndarray       # [B, N_max, 12] f64
    node_flags: np.ndarray      # [B, N_max] i32
    n_nodes: np.ndarray         # [B] i32
    ap_attr: np.ndarray         # [B, A_max, 4] f64
    ap_flags: np.ndarray        # [B, A_max] i32
    n_ap: np.ndarray            # [B] i32
    cons: np.ndarray            # [B, 6] f64
    points_px: Optional[np.ndarray] = None   # [B, N_max, 2] GUI pixel coordinates, when the batch was packed from pixels

    @property
    def B(self):
        return self.node_attr.shape[0]

    @property
    def N_max(self):
        return self.node_attr.shape[1]

    @property
    def A_max(self):
        return self.ap_attr.shape[1]

    def max_splines(self) -> int:
        """1 + the largest number of interior split nodes (reverse or turn) in the batch."""
        n = self.n_nodes[:, None]
        idx = np.arange(self.N_max)[None, :]
        split = (((self.node_flags & F_REVERSE) != 0) | (self.node_attr[:, :, A_TURN] != 0)) & (idx >= 1) & (idx < n - 1)
        return int(1 + split.sum(axis=1).max()) if self.B else 1

    def slice(self, lo: int, hi: int) -> "PackedPaths":
        return PackedPaths(*(np.ascontiguousarray(a[lo:hi]) for a in
                             (self.node_attr, self.node_flags, self.n_nodes, self.ap_attr, self.ap_flags, self.n_ap,
                              self.cons)),
                           points_px=None if self.points_px is None else np.ascontiguousarray(self.points_px[lo:hi]))

    def mirrored(self) -> "PackedPaths":
        """PathWidget.mirror_nodes (gui/path.py:596-600), exactly: every node moves to pixel x -> 2000 - x and gets
        turn -> -turn; NOTHING else changes (user tangents keep their direction, as in the reference).  The GUI mirrors
        in PIXEL space and only then converts to feet (path.py:365-367), which is not the same double as negating the
        feet, so the batch must have been packed from pixels (pack_arrays(points_px=...))."""
        if self.points_px is None:
            raise ValueError("mirrored() needs the pixel coordinates: pack the batch with pack_arrays(points_px=...) "
                             "(the reference mirrors in pixel space, gui/path.py:596-600)")
        px, turn = mirror_nodes_px(self.points_px, self.node_attr[:, :, A_TURN])
        na = self.node_attr.copy()
        na[:, :, 0:2] = px_to_ft(px)
        na[:, :, A_TURN] = turn
        rc, rs = rotation_table(turn, (self.node_flags & F_REVERSE) != 0)
        na[:, :, A_RCOS], na[:, :, A_RSIN] = rc, rs
        return PackedPaths(na, self.node_flags.copy(), self.n_nodes.copy(), self.ap_attr.copy(), self.ap_flags.copy(),
                           self.n_ap.copy(), self.cons.copy(), points_px=px)


def mirror_nodes_px(points_px, turn):
    """The transform of PathWidget.mirror_nodes (gui/path.py:596-600) on arrays: setPos(2000 - x, y), turn = -turn."""
    px = np.array(points_px, dtype=np.float64, copy=True)
    px[..., 0] = 2000 - px[..., 0]
    return px, 0.0 - np.asarray(turn, dtype=np.float64)     # the GUI's turn is an int: -0 is 0, never -0.0


def constraints_row(c) -> np.ndarray:
    """Constraints dataclass (motion_profile_generator.py:14-21) or a 6-sequence -> row of cons."""
    if hasattr(c, "max_vel"):
        return np.array([c.max_vel, c.max_acc, c.max_dec, c.friction_coef, c.max_jerk, c.track_width], dtype=np.float64)
    return np.asarray(c, dtype=np.float64).reshape(6)


def pack_arrays(points_ft, cons, reverse=None, stop=None, turn=None, wait=None, max_velocity=None,
                max_acceleration=None, tangent=None, in_mag=None, out_mag=None, n_nodes=None,
                ap_t=None, ap_stop=None, ap_wait=None, ap_max_velocity=None, ap_max_acceleration=None,
                n_ap=None, points_px=None) -> PackedPaths:
    """Vectorised packing from plain arrays: points_ft[B,N,2] (or None with points_px[B,N,2]: GUI pixels, converted
    as gui/path.py:365-367 does and kept for mirrored()); optional per-node arrays [B,N]; tangent[B,N,2] with NaN rows
    meaning "not set"; action-point arrays [B,A]; cons [6] or [B,6]."""
    if points_px is not None:
        points_px = np.ascontiguousarray(points_px, dtype=np.float64)
        if points_ft is None:
            points_ft = px_to_ft(points_px)
    pts = np.asarray(points_ft, dtype=np.float64)
    B, N = pts.shape[0], pts.shape[1]
    na = np.zeros((B, N, NA))
    nf = np.zeros((B, N), dtype=np.int32)
    na[:, :, 0:2] = pts

    def opt(a, default=0.0):
        return np.full((B, N), default) if a is None else np.asarray(a, dtype=np.float64).reshape(B, N)

    na[:, :, A_TURN] = opt(turn)
    na[:, :, A_WAIT] = opt(wait)
    na[:, :, A_MAXVEL] = opt(max_velocity)
    na[:, :, A_MAXACC] = opt(max_acceleration)
    rev = np.zeros((B, N), dtype=bool) if reverse is None else np.asarray(reverse, dtype=bool).reshape(B, N)
    stp = np.zeros((B, N), dtype=bool) if stop is None else np.asarray(stop, dtype=bool).reshape(B, N)
    nf |= rev.astype(np.int32) * F_REVERSE
    nf |= stp.astype(np.int32) * F_STOP
    if tangent is not None:
        tg = np.asarray(tangent, dtype=np.float64).reshape(B, N, 2)
        has = ~np.isnan(tg).any(axis=-1)
        na[:, :, A_TX] = np.where(has, tg[:, :, 0], 0.0)
        na[:, :, A_TY] = np.where(has, tg[:, :, 1], 0.0)
        na[:, :, A_INMAG] = np.where(has, opt(in_mag), 0.0)
        na[:, :, A_OUTMAG] = np.where(has, opt(out_mag), 0.0)
        nf |= has.astype(np.int32) * F_TANGENT
    na[:, :, A_RCOS], na[:, :, A_RSIN] = rotation_table(na[:, :, A_TURN], rev)
    nn = np.full(B, N, dtype=np.int32) if n_nodes is None else np.asarray(n_nodes, dtype=np.int32).reshape(B)
    if ap_t is None:
        apa = np.zeros((B, 1, APA)); apf = np.zeros((B, 1), dtype=np.int32); nap = np.zeros(B, dtype=np.int32)
    else:
        t = np.asarray(ap_t, dtype=np.float64)
        A = t.shape[1]
        apa = np.zeros((B, max(A, 1), APA)); apf = np.zeros((B, max(A, 1)), dtype=np.int32)
        apa[:, :A, 0] = t
        if ap_wait is not None:
            apa[:, :A, 1] = ap_wait
        if ap_max_velocity is not None:
            apa[:, :A, 2] = ap_max_velocity
        if ap_max_acceleration is not None:
            apa[:, :A, 3] = ap_max_acceleration
        if ap_stop is not None:
            apf[:, :A] = np.asarray(ap_stop, dtype=bool).astype(np.int32) * F_STOP
        nap = np.full(B, A, dtype=np.int32) if n_ap is None else np.asarray(n_ap, dtype=np.int32).reshape(B)
    cons = np.asarray(cons, dtype=np.float64)
    cons = np.tile(cons.reshape(1, 6), (B, 1)) if cons.size == 6 else cons.reshape(B, 6)
    return PackedPaths(np.ascontiguousarray(na), nf, nn, np.ascontiguousarray(apa), apf, nap, np.ascontiguousarray(cons),
                       points_px=points_px)


def pack_paths(paths: Sequence, constraints) -> PackedPaths:
    """Pack a list of (points_ft[N,2], nodes, action_points) triples of duck-typed reference objects.

    Attributes read: node.is_reverse_node, turn, wait_time, stop, tangent, incoming_magnitude,
    outgoing_magnitude, max_velocity, max_acceleration (spline_manager.py:61-158,
    motion_profile_generator.py:101-137,432-460,533-544); action point .t, stop, wait_time, max_velocity,
    max_acceleration (:143-163,548-553).  `constraints` is one Constraints (or 6-sequence) or one per path.
    """
    B = len(paths)
    N = max(len(p[1]) for p in paths)
    A = max([len(p[2]) if p[2] is not None else 0 for p in paths] + [1])
    na = np.zeros((B, N, NA)); nf = np.zeros((B, N), dtype=np.int32); nn = np.zeros(B, dtype=np.int32)
    apa = np.zeros((B, A, APA)); apf = np.zeros((B, A), dtype=np.int32); nap = np.zeros(B, dtype=np.int32)
    for b, (pts, nodes, aps) in enumerate(paths):
        pts = np.asarray(pts, dtype=np.float64)
        n = len(nodes)
        nn[b] = n
        na[b, :n, 0:2] = pts[:n]
        for i, nd in enumerate(nodes):
            r = na[b, i]
            r[A_TURN] = float(nd.turn)
            r[A_WAIT] = float(nd.wait_time)
            r[A_MAXVEL] = float(nd.max_velocity)
            r[A_MAXACC] = float(nd.max_acceleration)
            f = (F_REVERSE if nd.is_reverse_node else 0) | (F_STOP if nd.stop else 0)
            if nd.tangent is not None:
                f |= F_TANGENT
                r[A_TX], r[A_TY] = float(nd.tangent[0]), float(nd.tangent[1])
                r[A_INMAG] = float(nd.incoming_magnitude)
                r[A_OUTMAG] = float(nd.outgoing_magnitude)
            nf[b, i] = f
        aps = aps or []
        nap[b] = len(aps)
        for k, ap in enumerate(aps):
            apa[b, k] = (float(ap.t), float(ap.wait_time), float(ap.max_velocity), float(ap.max_acceleration))
            apf[b, k] = F_STOP if ap.stop else 0
    na[:, :, A_RCOS], na[:, :, A_RSIN] = rotation_table(na[:, :, A_TURN], (nf & F_REVERSE) != 0)
    if isinstance(constraints, (list, tuple)) and len(constraints) == B and not np.isscalar(constraints[0]):
        cons = np.stack([constraints_row(c) for c in constraints])
    else:
        cons = np.tile(constraints_row(constraints)[None], (B, 1))
    return PackedPaths(na, nf, nn, apa, apf, nap, np.ascontiguousarray(cons))
