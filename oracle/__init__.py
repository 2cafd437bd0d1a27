"""ctypes front end of the CPU oracle (oracle/vap_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of vap_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package.

Parity pin: tests/test_oracle_golden.py checks every stage against tests/golden/*.npz, which were
produced by the unmodified reference (tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "vap_oracle.c")
_SO = os.path.join(_HERE, "libvap_oracle.so")

NA, APA = 12, 4
F_REVERSE, F_STOP, F_TANGENT = 1, 2, 4
ERR = {-1: "False", -2: "IndexError", -3: "ValueError", -4: "capacity"}

_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (no FMA contraction; explicit fma() only)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-ffp-contract=off", "-fno-fast-math",
               "-mfma", "-fopenmp", "-o", _SO, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.ora_set_sq_mode.argtypes = [C.c_int]
        L.ora_set_threads.argtypes = [C.c_int]
        L.ora_get_sq_mode.restype = C.c_int
        L.ora_build_path.argtypes = [C.c_int, _dp, _ip, _dp, _ip, _dp, _dp, _dp]
        L.ora_build_path.restype = C.c_int
        L.ora_eval_spline.argtypes = [C.c_int, C.c_double, _dp, C.c_int, C.c_double, _dp]
        L.ora_eval.argtypes = [C.c_int, _ip, _dp, _dp, C.c_int, C.c_double, _dp]
        for f in (L.ora_exact_heading, L.ora_exact_curvature):
            f.argtypes = [C.c_int, _ip, _dp, _dp, C.c_double]
            f.restype = C.c_double
        L.ora_build_lut.argtypes = [C.c_int, _ip, _dp, _dp, C.c_int, _dp, _dp]
        L.ora_build_lut.restype = C.c_double
        L.ora_build_props.argtypes = [C.c_int, C.c_int, _ip, _dp, _dp, C.c_int, _dp, _dp]
        L.ora_distance_to_time.argtypes = [C.c_long, _dp, _dp, C.c_double, C.c_int, C.c_double]
        L.ora_distance_to_time.restype = C.c_double
        L.ora_snap.argtypes = [C.c_long, C.c_int, _dp, C.c_double]
        L.ora_snap.restype = C.c_double
        L.ora_dist_sample.argtypes = [C.c_int, _dp, _ip, C.c_int, _dp, _ip, _dp, C.c_double, C.c_double,
                                      C.c_long, _dp, _dp, C.c_double, C.c_long, _dp, _dp, C.c_long,
                                      _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int), _lp, _ip, C.POINTER(C.c_int)]
        L.ora_dist_sample.restype = C.c_long
        L.ora_fwd_bwd.argtypes = [C.c_long, _dp, _dp, _dp, _dp, C.c_double, _dp, C.c_int, _lp, _ip,
                                  C.c_double, C.c_double, C.c_int]
        L.ora_trapezoid.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_long, _dp]
        L.ora_trapezoid.restype = C.c_long
        L.ora_motion_profile_angle.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                               C.c_long, _dp, _dp]
        L.ora_motion_profile_angle.restype = C.c_long
        L.ora_lerp_uniform.argtypes = [C.c_double, C.c_double, C.c_long, _dp]
        L.ora_lerp_uniform.restype = C.c_double
        L.ora_profile.argtypes = [C.c_int, _dp, _ip, C.c_int, _dp, _ip, _dp, C.c_double, C.c_double,
                                  C.c_int, _ip, _dp, _dp, C.c_long, _dp, _dp, C.c_double, C.c_long, _dp, _dp,
                                  C.c_long, _dp, C.c_long, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp,
                                  _lp, C.POINTER(C.c_int), _lp, C.POINTER(C.c_int)]
        L.ora_profile.restype = C.c_long
        L.ora_gl_arclen.argtypes = [C.c_int, C.c_double, _dp, C.c_double, C.c_double, C.c_int, _dp, _dp,
                                    C.POINTER(C.c_double)]
        L.ora_gl_arclen.restype = C.c_int
        L.ora_gl_inverse.argtypes = [C.c_int, C.c_double, _dp, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp,
                                     C.POINTER(C.c_double)]
        L.ora_gl_inverse.restype = C.c_int
        L.ora_full.argtypes = [C.c_int, _dp, _ip, C.c_int, _dp, _ip, _dp, C.c_double, C.c_double, C.c_long, C.c_long,
                               _dp, C.POINTER(C.c_long), _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp,
                               _lp, C.POINTER(C.c_int), _lp, C.POINTER(C.c_int), _dp]
        L.ora_full.restype = C.c_long
        L.ora_full_batch.argtypes = [C.c_long, C.c_int, _dp, _ip, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, _dp,
                                     C.c_double, C.c_double, C.c_long, C.c_long, _dp]
        L.ora_full_batch_ex.argtypes = [C.c_long, C.c_int, _dp, _ip, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, _dp,
                                        C.c_double, C.c_double, C.c_long, C.c_long, _dp, _lp, _dp]
        _lib = L
    return _lib


def set_sq_mode(mode: int) -> None:
    """0: x**2 == libm pow(x, 2.0) (reference-faithful). 1: x**2 == x*x (what the CUDA engine does)."""
    lib().ora_set_sq_mode(int(mode))


class OracleError(Exception):
    def __init__(self, code):
        super().__init__(ERR.get(int(code), str(code)))
        self.code = int(code)


def _ap_arrays(ap_attr, ap_flags):
    if ap_attr is None or len(ap_attr) == 0:
        return np.zeros((1, APA)), np.zeros(1, dtype=np.int32), 0
    a = np.ascontiguousarray(ap_attr, dtype=np.float64).reshape(-1, APA)
    return a, np.ascontiguousarray(ap_flags, dtype=np.int32), a.shape[0]


class Geometry:
    """Result of build_path: segment tables + spline split map."""

    def __init__(self, node_attr, node_flags):
        L = lib()
        self.na = np.ascontiguousarray(node_attr, dtype=np.float64).reshape(-1, NA)
        self.nf = np.ascontiguousarray(node_flags, dtype=np.int32)
        n = self.n = self.na.shape[0]
        self.seg = np.zeros((max(n - 1, 1), 6, 2))
        fn = np.zeros(n + 1, dtype=np.int32)
        pe = np.zeros(n)
        self.seglen = np.zeros(n)
        pc = np.zeros(2 * n)
        S = L.ora_build_path(n, self.na, self.nf, self.seg.reshape(-1), fn, pe, self.seglen, pc)
        if S < 0:
            raise OracleError(S)
        self.S = S
        self.first_node = np.ascontiguousarray(fn[: S + 1])
        self.param_end = np.ascontiguousarray(pe[:S])
        self.seglen = self.seglen[: n - 1]
        self.params_concat = pc[: n - 1 + S]
        self.seg = self.seg[: n - 1]
        self._segflat = np.ascontiguousarray(self.seg.reshape(-1))

    def eval(self, which, t):
        out = np.zeros(2)
        lib().ora_eval(self.S, self.first_node, self.param_end, self._segflat, which, float(t), out)
        return out

    def eval_spline(self, k, which, t):
        out = np.zeros(2)
        a, b = int(self.first_node[k]), int(self.first_node[k + 1])
        lib().ora_eval_spline(b - a, float(self.param_end[k]), np.ascontiguousarray(self._segflat[a * 12:]), which,
                              float(t), out)
        return out

    def exact_heading(self, t):
        return lib().ora_exact_heading(self.S, self.first_node, self.param_end, self._segflat, float(t))

    def exact_curvature(self, t):
        return lib().ora_exact_curvature(self.S, self.first_node, self.param_end, self._segflat, float(t))

    def build_lut(self, samples=1000):
        d = np.zeros(samples * self.S)
        t = np.zeros(samples * self.S)
        total = lib().ora_build_lut(self.S, self.first_node, self.param_end, self._segflat, samples, d, t)
        return d, t, total

    def build_props(self, spn=1000):
        k = np.zeros(spn * self.n)
        h = np.zeros(spn * self.n)
        lib().ora_build_props(self.n, self.S, self.first_node, self.param_end, self._segflat, spn, k, h)
        return k, h

    def gl_arclen(self, k, t0, t1, pts, wts):
        out = C.c_double()
        a, b = int(self.first_node[k]), int(self.first_node[k + 1])
        r = lib().ora_gl_arclen(b - a, float(self.param_end[k]), np.ascontiguousarray(self._segflat[a * 12:]),
                                float(t0), float(t1), len(pts), np.ascontiguousarray(pts), np.ascontiguousarray(wts),
                                C.byref(out))
        if r < 0:
            raise OracleError(r)
        return out.value

    def gl_inverse(self, k, s, pts, wts, tol=1e-6, max_iter=50):
        out = C.c_double()
        a, b = int(self.first_node[k]), int(self.first_node[k + 1])
        r = lib().ora_gl_inverse(b - a, float(self.param_end[k]), np.ascontiguousarray(self._segflat[a * 12:]),
                                 float(s), tol, max_iter, len(pts), np.ascontiguousarray(pts),
                                 np.ascontiguousarray(wts), C.byref(out))
        if r < 0:
            raise OracleError(r)
        return out.value


def distance_to_time(lut_d, lut_t, total, n, d):
    return lib().ora_distance_to_time(len(lut_d), lut_d, lut_t, float(total), int(n), float(d))


def snap(vals, n, t):
    return lib().ora_snap(len(vals), int(n), vals, float(t))


def dist_sample(geom: Geometry, ap_attr, ap_flags, cons, dd, lut_d, lut_t, total, kap, th, end_vel=0.01, cap=None):
    L = lib()
    apa, apf, A = _ap_arrays(ap_attr, ap_flags)
    if cap is None:
        cap = int(total / dd) + 16
    t = np.zeros(cap); k = np.zeros(cap); h = np.zeros(cap); v = np.zeros(cap)
    ma = np.zeros(geom.n + A + 3)
    bidx = np.zeros(geom.n + A + 3, dtype=np.int64)
    bval = np.zeros(geom.n + A + 3, dtype=np.int32)
    n_acc, n_b = C.c_int(), C.c_int()
    D = L.ora_dist_sample(geom.n, geom.na, geom.nf, A, apa, apf, np.ascontiguousarray(cons, dtype=np.float64),
                          float(dd), float(end_vel), len(lut_d), lut_d, lut_t, float(total), len(kap), kap, th,
                          cap, t, k, h, v, ma, C.byref(n_acc), bidx, bval, C.byref(n_b))
    if D < 0:
        raise OracleError(D)
    return dict(D=D, t=t[:D].copy(), kap=k[:D].copy(), th=h[:D].copy(), v0=v[:D].copy(),
                max_accels=ma[: n_acc.value].copy(), bidx=bidx[: n_b.value].copy(), bval=bval[: n_b.value].copy())


def fwd_bwd(kap, th, v0, cons, dd, max_accels, bidx, bval, start_vel=0.01, end_vel=0.01, forward_only=False):
    v = np.array(v0, dtype=np.float64, copy=True)
    lib().ora_fwd_bwd(len(v), np.ascontiguousarray(kap), np.ascontiguousarray(th), v,
                      np.ascontiguousarray(cons, dtype=np.float64), float(dd), np.ascontiguousarray(max_accels),
                      len(bidx), np.ascontiguousarray(bidx, dtype=np.int64), np.ascontiguousarray(bval, dtype=np.int32),
                      start_vel, end_vel, int(forward_only))
    return v


def trapezoid(V, A, dist, dt=0.01, cap=1 << 16):
    v = np.zeros(cap)
    K = lib().ora_trapezoid(V, A, dist, dt, cap, v)
    if K < 0:
        raise OracleError(K)
    return v[:K].copy()


def motion_profile_angle(angle, V, A, w, dt=0.01, cap=1 << 16):
    h = np.zeros(cap); o = np.zeros(cap)
    K = lib().ora_motion_profile_angle(float(angle), V, A, w, dt, cap, h, o)
    if K < 0:
        raise OracleError(K)
    return h[:K].copy(), o[:K].copy()


def lerp_uniform(x, dd, ys):
    return lib().ora_lerp_uniform(float(x), float(dd), len(ys), np.ascontiguousarray(ys))


def profile(geom: Geometry, ap_attr, ap_flags, cons, dt, dd, lut_d, lut_t, total, kap, th, vel, cap=None):
    L = lib()
    apa, apf, A = _ap_arrays(ap_attr, ap_flags)
    if cap is None:
        cap = 8 * len(vel) + 4096
    o = [np.zeros(cap) for _ in range(8)]
    nmap = np.zeros(geom.n + 2, dtype=np.int64)
    amap = np.zeros(A + 2, dtype=np.int64)
    n_nm, n_am = C.c_int(), C.c_int()
    T = L.ora_profile(geom.n, geom.na, geom.nf, A, apa, apf, np.ascontiguousarray(cons, dtype=np.float64), dt, dd,
                      geom.S, geom.first_node, geom.param_end, geom._segflat, len(lut_d), lut_d, lut_t, float(total),
                      len(kap), kap, th, len(vel), np.ascontiguousarray(vel), cap, *o, nmap, C.byref(n_nm), amap,
                      C.byref(n_am))
    if T < 0:
        raise OracleError(T)
    names = ["times", "positions", "linear_vels", "accelerations", "headings", "angular_vels", "x", "y"]
    res = {k: a[:T].copy() for k, a in zip(names, o)}
    res["T"] = T
    res["nodes_map"] = nmap[: n_nm.value].copy()
    res["actions_map"] = amap[: n_am.value].copy()
    return res


def full(node_attr, node_flags, ap_attr, ap_flags, cons, dt=0.01, dd=0.005, cap_d=None, cap_t=None):
    """build_path + generate_motion_profile for one path.  Returns a dict or raises OracleError."""
    L = lib()
    na = np.ascontiguousarray(node_attr, dtype=np.float64).reshape(-1, NA)
    nf = np.ascontiguousarray(node_flags, dtype=np.int32)
    n = na.shape[0]
    apa, apf, A = _ap_arrays(ap_attr, ap_flags)
    if cap_d is None:
        chord = float(np.sum(np.hypot(np.diff(na[:, 0]), np.diff(na[:, 1]))))
        cap_d = int(4 * chord / dd) + 4096
    if cap_t is None:
        cap_t = 2 * cap_d
    vel = np.zeros(cap_d)
    o = [np.zeros(cap_t) for _ in range(8)]
    nmap = np.zeros(n + 2, dtype=np.int64)
    amap = np.zeros(A + 2, dtype=np.int64)
    n_nm, n_am, D = C.c_int(), C.c_int(), C.c_long()
    summ = np.zeros(5)
    T = L.ora_full(n, na, nf, A, apa, apf, np.ascontiguousarray(cons, dtype=np.float64), dt, dd, cap_d, cap_t, vel,
                   C.byref(D), *o, nmap, C.byref(n_nm), amap, C.byref(n_am), summ)
    if T < 0:
        raise OracleError(T)
    names = ["times", "positions", "linear_vels", "accelerations", "headings", "angular_vels", "x", "y"]
    res = {k: a[:T].copy() for k, a in zip(names, o)}
    res.update(T=T, D=D.value, vel=vel[: D.value].copy(), nodes_map=nmap[: n_nm.value].copy(),
               actions_map=amap[: n_am.value].copy(), summary=summ)
    return res


def full_batch(node_attr, node_flags, cons, n_ap=None, ap_attr=None, ap_flags=None, dt=0.01, dd=0.005,
               cap_d=60000, cap_t=40000, threads=None):
    """B same-length paths, OpenMP over paths; returns summaries[B,5] = T, L, t_end, max|v|, status."""
    L = lib()
    na = np.ascontiguousarray(node_attr, dtype=np.float64)
    B, n = na.shape[0], na.shape[1]
    nf = np.ascontiguousarray(node_flags, dtype=np.int32)
    cons = np.ascontiguousarray(cons, dtype=np.float64).reshape(B, 6)
    if threads:
        L.ora_set_threads(int(threads))
    summ = np.zeros((B, 5))
    if ap_attr is not None:
        apa = np.ascontiguousarray(ap_attr, dtype=np.float64)
        apf = np.ascontiguousarray(ap_flags, dtype=np.int32)
        nap = np.ascontiguousarray(n_ap, dtype=np.int32)
        Amax = apa.shape[1]
        L.ora_full_batch(B, n, na.reshape(-1), nf.reshape(-1), Amax, nap.ctypes.data, apa.ctypes.data, apf.ctypes.data,
                         cons.reshape(-1), dt, dd, cap_d, cap_t, summ.reshape(-1))
    else:
        L.ora_full_batch(B, n, na.reshape(-1), nf.reshape(-1), 0, None, None, None, cons.reshape(-1), dt, dd, cap_d,
                         cap_t, summ.reshape(-1))
    return summ


def full_batch_ex(node_attr, node_flags, cons, n_ap=None, ap_attr=None, ap_flags=None, dt=0.01, dd=0.005,
                  cap_d=60000, cap_t=40000, threads=None):
    """full_batch that also returns the integer outputs: (summaries[B,5], ints[B,K], checks[B,4]) with
    ints = D, len(nodes_map), len(actions_map), nodes_map (n+1 slots, -1 padded), actions_map (Amax slots);
    checks = sums of the velocity row, of times, of x, of y."""
    L = lib()
    na = np.ascontiguousarray(node_attr, dtype=np.float64)
    B, n = na.shape[0], na.shape[1]
    nf = np.ascontiguousarray(node_flags, dtype=np.int32)
    cons = np.ascontiguousarray(cons, dtype=np.float64).reshape(B, 6)
    if threads:
        L.ora_set_threads(int(threads))
    summ = np.zeros((B, 5))
    if ap_attr is not None:
        apa = np.ascontiguousarray(ap_attr, dtype=np.float64)
        apf = np.ascontiguousarray(ap_flags, dtype=np.int32)
        nap = np.ascontiguousarray(n_ap, dtype=np.int32)
        Amax = apa.shape[1]
        ptrs = (nap.ctypes.data, apa.ctypes.data, apf.ctypes.data)
    else:
        Amax, ptrs = 0, (None, None, None)
    ints = np.zeros((B, 3 + (n + 1) + Amax), dtype=np.int64)
    checks = np.zeros((B, 4))
    L.ora_full_batch_ex(B, n, na.reshape(-1), nf.reshape(-1), Amax, *ptrs, cons.reshape(-1), dt, dd, cap_d, cap_t,
                        summ.reshape(-1), ints.reshape(-1), checks.reshape(-1))
    return summ, ints, checks
