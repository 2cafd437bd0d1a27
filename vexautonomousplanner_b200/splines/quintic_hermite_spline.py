"""Mirror of src/splines/quintic_hermite_spline.py (class QuinticHermiteSpline) on the CUDA engine.

Same public names, arguments, return types and error behaviour as the reference class
(quintic_hermite_spline.py:11-748); every numeric result comes from libvap.so:
  fit                      -> vap_fit_splines     (:30-219, :719-736)
  get_point / derivative   -> vap_eval            (:221-251, :473-541, :288-416)
  get_heading / curvature  -> vap_eval which=3    (spline.py:48-80)
  get_arc_length           -> vap_gl mode 0       (:592-644, Gauss-Legendre with numpy's leggauss nodes)
  get_parameter_by_arc_..  -> vap_gl mode 1       (:661-717, bisection)
"""
from __future__ import annotations

import logging
from typing import Optional

import numpy as np
import torch

from ..runtime import C, Keep, _p, check, dev, get_engine, stream
from .spline import Spline

logger = logging.getLogger(__name__)


class QuinticHermiteSpline(Spline):
    """Quintic Hermite spline through control points with first / second derivative constraints."""

    def __init__(self):
        super().__init__()
        self.first_derivatives: Optional[np.ndarray] = None
        self.second_derivatives: Optional[np.ndarray] = None
        self.starting_tangent: Optional[np.ndarray] = None
        self.ending_tangent: Optional[np.ndarray] = None
        self.set_tangents = None
        self.control_points: Optional[np.ndarray] = None
        self.parameters: Optional[np.ndarray] = None
        self.segments = []
        self.segment_lengths = []
        self._dev = None          # (seg, first_node, param_end, n_splines) on the device
        self._dirty = True

    # ------------------------------------------------------------------ fitting
    def fit(self, x, y, first_derivatives=None, second_derivatives=None) -> bool:
        if len(x) != len(y):
            return False
        if len(x) < 2:
            return False
        try:
            n = len(x)
            self.control_points = np.column_stack((x, y)).astype(np.float64)
            if first_derivatives is not None:
                if len(first_derivatives) != n:
                    return False
                self.first_derivatives = first_derivatives
            if second_derivatives is not None:
                if len(second_derivatives) != n:
                    return False
                self.second_derivatives = second_derivatives
            given = self.first_derivatives is not None and self.second_derivatives is not None
            # the reference indexes set_tangents[i] unconditionally inside a log call (:99-101): None -> False
            if self.set_tangents is None:
                raise TypeError("'NoneType' object is not subscriptable")
            has = np.zeros((1, n), dtype=np.int32)
            tin = np.zeros((1, n, 2)); tout = np.zeros((1, n, 2))
            for i in range(n):
                st = self.set_tangents[i]
                if st is None:
                    raise TypeError("'NoneType' object is not subscriptable")
                if st[0] is not None:
                    has[0, i] |= 1; tin[0, i] = np.asarray(st[0], dtype=np.float64)
                if st[1] is not None:
                    has[0, i] |= 2; tout[0, i] = np.asarray(st[1], dtype=np.float64)
            bh = 0
            bnd = np.zeros((1, 2, 2))
            for bit_apply, bit_attr, k, tg in ((1, 4, 0, self.starting_tangent), (2, 8, 1, self.ending_tangent)):
                if tg is not None:
                    bh |= bit_attr
                    if isinstance(tg, np.ndarray) and tg.shape == (2,):
                        bh |= bit_apply
                        bnd[0, k] = tg
            deriv_in = None
            if given:
                fd = np.asarray(self.first_derivatives, dtype=np.float64).reshape(n, 2)
                sd = np.asarray(self.second_derivatives, dtype=np.float64).reshape(n, 2)
                deriv_in = dev(np.concatenate([fd, sd], axis=1)[None])
                bh |= 16
            eng = get_engine()
            d_seg = eng._empty((1, max(n - 1, 1), 6, 2)); d_len = eng._empty((1, n)); d_par = eng._empty((1, n))
            d_st = eng._empty((1,), torch.int32); d_scr = eng._empty((1, n, 5)); d_der = eng._empty((1, n, 4))
            k = Keep()
            check(eng.lib.vap_fit_splines(C.c_int64(1), C.c_int(n), k([n], torch.int32),
                                          k(self.control_points[None]), k(has, torch.int32), k(tin),
                                          k(tout), k([bh], torch.int32), k(bnd), _p(d_seg), _p(d_len),
                                          _p(d_par), _p(d_st), _p(d_scr), _p(deriv_in), _p(d_der), stream()),
                  "vap_fit_splines")
            if int(d_st.item()) != 0:
                return False
            self._load(self.control_points, d_seg[0, : n - 1].cpu().numpy(), d_len[0, : n - 1].cpu().numpy(),
                       d_par[0].cpu().numpy(), d_der[0].cpu().numpy())
            return True
        except Exception as e:      # the reference swallows every exception (:136-138)
            logger.error(f"Error during fitting: {str(e)}")
            return False

    def _load(self, control_points, seg, seglen, params, derivs):
        """Populate the public attributes from engine results (also used by the manager mirror)."""
        self.control_points = np.asarray(control_points, dtype=np.float64)
        self.x_points = self.control_points[:, 0].copy()
        self.y_points = self.control_points[:, 1].copy()
        self.segments = [np.array(s) for s in seg]
        self.segment_lengths = [np.float64(v) for v in seglen]
        self.parameters = np.asarray(params, dtype=np.float64)
        self.first_derivatives = np.asarray(derivs[:, 0:2], dtype=np.float64).copy()
        self.second_derivatives = np.asarray(derivs[:, 2:4], dtype=np.float64).copy()
        self._dirty = True

    def _device(self):
        if self._dirty or self._dev is None:
            n = len(self.segments) + 1
            seg = np.stack(self.segments)[None]
            self._dev = (dev(seg), dev([[0, n - 1] + [0] * (n - 1)], torch.int32),
                         dev([[self.parameters[-1]] + [0.0] * (n - 1)]), dev([1], torch.int32), n)
            self._dirty = False
        return self._dev

    def set_tangent(self, tangent: np.ndarray, index: int):
        if self.set_tangents is None:
            self.set_tangents = np.zeros_like(self.control_points, dtype=float)
        self.set_tangents[index] = tangent

    def set_all_tangents(self, tangents):
        self.set_tangents = tangents

    # ------------------------------------------------------------------ evaluation
    def _eval(self, which: int, t) -> np.ndarray:
        if not self.segments:
            raise ValueError("Spline has not been fitted yet")
        seg, fn, pe, ns, n = self._device()
        eng = get_engine()
        tt = np.atleast_1d(np.asarray(t, dtype=np.float64))
        out = eng._empty((tt.size, 2))
        k = Keep()
        check(eng.lib.vap_eval(C.c_int64(tt.size), k(torch.zeros(tt.size, dtype=torch.int32, device=eng.device)),
                               k(tt), C.c_int(which), C.c_int(n), _p(seg), _p(fn), _p(pe), _p(ns), _p(out),
                               stream()), "vap_eval")
        res = out.cpu().numpy()
        return res[0] if np.ndim(t) == 0 else res

    def get_point(self, t: float) -> np.ndarray:
        return self._eval(0, t)

    def get_derivative(self, t: float, debug: bool = False) -> np.ndarray:
        return self._eval(1, t)

    def get_second_derivative(self, t: float, debug: bool = False) -> np.ndarray:
        return self._eval(2, t)

    def _heading_curvature(self, t: float):
        r = self._eval(3, t)
        return np.float64(r[0]), np.float64(r[1])

    def get_magnitude(self, idx):
        return self.segment_lengths[idx]

    def percent_to_point(self, percent: float) -> np.ndarray:
        if not self.segments:
            raise ValueError("Spline has not been fitted yet")
        return self.get_point(self.parameters[0] + self.parameters[-1] * (percent / 100))

    def percent_to_parameter(self, percent: float) -> float:
        if not self.segments:
            raise ValueError("Spline has not been fitted yet")
        return self.parameters[0] + self.parameters[-1] * (percent / 100)

    def _normalize_parameter(self, t: float):
        """(local_t, segment_idx) as in quintic_hermite_spline.py:506-541 (host arithmetic on two scalars)."""
        if not self.parameters.size:
            raise ValueError("Spline has not been fitted yet")
        t_min, t_max = self.parameters[0], self.parameters[-1]
        t = max(t_min, min(t, t_max))
        idx = int((t - t_min) / 1.0)
        if idx == len(self.segments):
            idx = len(self.segments) - 1
        return (t - (t_min + idx * 1.0)) / 1.0, idx

    def set_starting_tangent(self, tangent: np.ndarray) -> bool:
        if not isinstance(tangent, np.ndarray) or tangent.shape != (2,):
            return False
        self.first_derivatives[0] = tangent
        if len(self.segments) > 0:
            self.segments[-1][2] = tangent        # the reference writes the LAST segment (:561)
        self.starting_tangent = tangent
        self._dirty = True
        return True

    def set_ending_tangent(self, tangent: np.ndarray) -> bool:
        if not isinstance(tangent, np.ndarray) or tangent.shape != (2,):
            return False
        self.first_derivatives[-1] = tangent
        if len(self.segments) > 0:
            self.segments[-1][3] = tangent
        self.ending_tangent = tangent
        self._dirty = True
        return True

    # ------------------------------------------------------------------ arc length
    def _gl(self, mode: int, a: float, b: float, num_points: int = 20, max_iter: int = 50) -> float:
        seg, fn, pe, ns, n = self._device()
        eng = get_engine()
        pts, wts = np.polynomial.legendre.leggauss(num_points)
        out = eng._empty((1,)); qst = eng._empty((1,), torch.int32)
        z = torch.zeros(1, dtype=torch.int32, device=eng.device)
        k = Keep()
        check(eng.lib.vap_gl(C.c_int64(1), _p(z), _p(z), k([a]), k([b]), C.c_int(mode), C.c_int(max_iter),
                             C.c_int(num_points), k(pts), k(wts), C.c_int(n), _p(seg), _p(fn), _p(pe),
                             _p(out), _p(qst), stream()), "vap_gl")
        return float(out.item())

    def get_arc_length(self, t_start: float, t_end: float, num_points: int = 20) -> float:
        if not self.segments:
            raise ValueError("Spline has not been fitted yet")
        if t_start >= t_end:
            raise ValueError("t_start must be less than t_end")
        t_min, t_max = self.parameters[0], self.parameters[-1]
        if t_start < t_min or t_end > t_max:
            raise ValueError(f"Parameters must be within range [{t_min}, {t_max}]")
        return float(self._gl(0, t_start, t_end, num_points))

    def get_total_arc_length(self) -> float:
        if not self.segments:
            raise ValueError("Spline has not been fitted yet")
        return self.get_arc_length(self.parameters[0], self.parameters[-1])

    def get_parameter_by_arc_length(self, arc_length: float, tolerance: float = 1e-6, max_iterations: int = 50) -> float:
        if not self.segments:
            raise ValueError("Spline has not been fitted yet")
        if arc_length < 0:
            raise ValueError("Arc length must be non-negative")
        total_length = self.get_total_arc_length()
        if arc_length > total_length:
            raise ValueError(f"Arc length {arc_length} exceeds total length {total_length}")
        if arc_length == 0:
            return self.parameters[0]
        if arc_length == total_length:
            return self.parameters[-1]
        return self._gl(1, arc_length, tolerance, 20, max_iterations)

    def get_end_parameter(self) -> float:
        if not self.parameters.size:
            raise ValueError("Spline has not been fitted yet")
        return self.parameters[-1]
