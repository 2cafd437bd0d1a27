#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference headless.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference's hot path (src/splines, src/motion_profiling_v2) imports as-is;
`splines.spline_manager` only needs `gui.node.Node` / `gui.action_point.ActionPoint`
as type annotations (spline_manager.py:8-9), so two attribute-only stand-ins are
registered in sys.modules before the import.  Nothing
from the reference is copied: this script only *calls* it and stores inputs,
intermediates and outputs.

Each fixture `case_<name>.npz` holds
  inputs : points_ft[N,2], node attribute arrays, action-point arrays, constraints[6],
           rot_cos/rot_sin[N] (the numpy cos/sin the reference itself evaluated at
           split nodes, spline_manager.py:105-113 -- index-critical libm values)
  S0     : seg[G,6,2] (all splines' segment tables concatenated in path order),
           spline_first_node[S+1], spline_param_end[S], spline_seglen[G]
  S1     : lut_d[Q], lut_t[Q], total_length
  S2     : prop_k[P], prop_h[P]  (parameters are analytic: linspace(0,N-1,1000N))
  S3     : t[D], kap[D], th[D]   (reference distance_to_time / get_curvature / get_heading at
           d_i accumulated the way forward_backward_pass does), max_accels, boundary_map (from the log)
  S4/S5  : vel[D]   (forward_backward_pass return value)
  S6     : times, positions, linear_vels, accelerations, headings, angular_vels, coords[T,2],
           nodes_map, actions_map
  API    : point / derivative / second-derivative / heading / curvature queries on the manager,
           GL arc length + inverse on each spline, motion_profile_angle, trapezoid profile.
"""
from __future__ import annotations

import logging
import os
import re
import sys
import types
import warnings

import numpy as np

REF_SRC = "/root/reference/src"
OUT_DIR = os.path.dirname(os.path.abspath(__file__))
PX2FT = 12.1090395251  # path.py:365-367


def _install_stubs():
    gui = types.ModuleType("gui")
    gui.__path__ = []
    node_mod = types.ModuleType("gui.node")
    ap_mod = types.ModuleType("gui.action_point")

    class Node:  # attribute names/defaults from gui/node.py:17-51
        def __init__(self, **kw):
            self.is_reverse_node = False
            self.turn = 0
            self.wait_time = 0
            self.stop = False
            self.tangent = None
            self.incoming_magnitude = None
            self.outgoing_magnitude = None
            self.max_velocity = 0
            self.max_acceleration = 0
            for k, v in kw.items():
                setattr(self, k, v)

    class ActionPoint:  # gui/action_point.py:16-41
        def __init__(self, t, **kw):
            self.t = t
            self.stop = False
            self.wait_time = 0
            self.max_velocity = 0
            self.max_acceleration = 0
            for k, v in kw.items():
                setattr(self, k, v)

    node_mod.Node = Node
    ap_mod.ActionPoint = ActionPoint
    sys.modules["gui"] = gui
    sys.modules["gui.node"] = node_mod
    sys.modules["gui.action_point"] = ap_mod
    return Node, ActionPoint


Node, ActionPoint = _install_stubs()
sys.path.insert(0, REF_SRC)
from splines.spline_manager import QuinticHermiteSplineManager  # noqa: E402
import motion_profiling_v2.motion_profile_generator as mpg  # noqa: E402
import motion_profiling_v2.one_dim_mp_generator as odm  # noqa: E402

logging.getLogger("splines.quintic_hermite_spline").setLevel(logging.ERROR)
logging.getLogger("splines.spline_manager").setLevel(logging.ERROR)
warnings.filterwarnings("ignore", category=RuntimeWarning)


class _Grab(logging.Handler):
    def __init__(self):
        super().__init__(level=logging.INFO)
        self.max_accels = None
        self.boundary_map = None

    def emit(self, record):
        msg = record.getMessage()
        if msg.startswith("Max Accels: "):
            self.max_accels = eval(msg[len("Max Accels: "):], {"np": np, "__builtins__": {}})
        elif msg.startswith("Boundary Map: "):
            self.boundary_map = eval(msg[len("Boundary Map: "):], {"np": np, "__builtins__": {}})


_grab = _Grab()
_lg = logging.getLogger("motion_profiling_v2.motion_profile_generator")
_lg.setLevel(logging.INFO)
_lg.addHandler(_grab)
_lg.propagate = False

FACTORY = dict(max_vel=4.0, max_acc=8.0, max_jerk=16.0, track_width=12.5 / 12)   # src/config.yaml
REPOROOT = dict(max_vel=4.0, max_acc=12.0, max_jerk=43.0, track_width=11.5 / 12)  # config.yaml


def px_to_ft(px):
    px = np.asarray(px, dtype=float)
    return (px / 2000 - 0.5) * PX2FT


def run_case(name, px, node_kw=None, aps=None, cons=FACTORY, dt=0.01, dd=0.005, n_api=40, seed=0):
    pts = px_to_ft(px)
    n = len(pts)
    node_kw = node_kw or {}
    nodes = []
    for i in range(n):
        kw = dict(node_kw.get(i, {}))
        if "tangent" in kw and kw["tangent"] is not None:
            kw["tangent"] = np.asarray(kw["tangent"], dtype=float)
        nodes.append(Node(**kw))
    aplist = [ActionPoint(**a) for a in (aps or [])]

    sm = QuinticHermiteSplineManager()
    ok = sm.build_path(pts, nodes, aplist)
    assert ok, name

    c = mpg.Constraints(max_vel=cons["max_vel"], max_acc=cons["max_acc"], max_dec=cons["max_acc"],
                        friction_coef=0.8, max_jerk=cons["max_jerk"], track_width=cons["track_width"])
    out = {}
    out["points_ft"] = pts
    out["points_px"] = np.asarray(px, dtype=float)      # the GUI works in pixels (mirror_nodes, gui/path.py:596-600)
    out["n_reverse"] = np.array([bool(nd.is_reverse_node) for nd in nodes])
    out["n_turn"] = np.array([float(nd.turn) for nd in nodes])
    out["n_wait"] = np.array([float(nd.wait_time) for nd in nodes])
    out["n_stop"] = np.array([bool(nd.stop) for nd in nodes])
    out["n_maxvel"] = np.array([float(nd.max_velocity) for nd in nodes])
    out["n_maxacc"] = np.array([float(nd.max_acceleration) for nd in nodes])
    out["n_has_tangent"] = np.array([nd.tangent is not None for nd in nodes])
    out["n_tangent"] = np.array([nd.tangent if nd.tangent is not None else [0.0, 0.0] for nd in nodes], dtype=float)
    out["n_inmag"] = np.array([float(nd.incoming_magnitude or 0.0) for nd in nodes])
    out["n_outmag"] = np.array([float(nd.outgoing_magnitude or 0.0) for nd in nodes])
    # the libm values the reference itself evaluates at turn nodes (spline_manager.py:105-113)
    rc = np.ones(n)
    rs = np.zeros(n)
    for i, nd in enumerate(nodes):
        if nd.turn != 0:
            ang = np.radians(nd.turn)
            if nd.is_reverse_node:
                ang = ang + np.pi
            rc[i] = np.cos(ang)
            rs[i] = np.sin(ang)
    out["rot_cos"], out["rot_sin"] = rc, rs
    out["ap_t"] = np.array([float(a.t) for a in aplist])
    out["ap_stop"] = np.array([bool(a.stop) for a in aplist])
    out["ap_wait"] = np.array([float(a.wait_time) for a in aplist])
    out["ap_maxvel"] = np.array([float(a.max_velocity) for a in aplist])
    out["ap_maxacc"] = np.array([float(a.max_acceleration) for a in aplist])
    out["constraints"] = np.array([c.max_vel, c.max_acc, c.max_dec, c.friction_coef, c.max_jerk, c.track_width])
    out["dt_dd"] = np.array([dt, dd])

    # ---- S0
    segs, first, pend, seglen, params = [], [0], [], [], []
    for sp in sm.splines:
        segs.extend(np.asarray(s, dtype=float) for s in sp.segments)
        first.append(first[-1] + len(sp.control_points) - 1)
        pend.append(float(sp.parameters[-1]))
        seglen.extend(float(x) for x in sp.segment_lengths)
        params.extend(float(x) for x in sp.parameters)
    out["seg"] = np.array(segs)
    out["spline_first_node"] = np.array(first, dtype=np.int64)
    out["spline_param_end"] = np.array(pend)
    out["spline_seglen"] = np.array(seglen)
    out["spline_params_concat"] = np.array(params)

    # ---- API queries (before tables are built; they do not depend on them)
    rng = np.random.default_rng(seed + 1000)
    tq = np.concatenate([rng.uniform(-0.2, n - 0.8, n_api), np.arange(n, dtype=float), [0.5, n - 1.5]])
    out["api_t"] = tq
    out["api_point"] = np.array([sm.get_point_at_parameter(t) for t in tq])
    out["api_d1"] = np.array([sm.get_derivative_at_parameter(t) for t in tq])
    out["api_d2"] = np.array([sm.get_second_derivative_at_parameter(t) for t in tq])
    out["api_heading_exact"] = np.array([sm._get_heading(t) for t in tq])
    out["api_curv_exact"] = np.array([sm._get_curvature(t) for t in tq])
    # GL arc length + inverse per spline (quintic_hermite_spline.py:592-717)
    gl_sp, gl_t0, gl_t1, gl_len, inv_sp, inv_s, inv_t, tot = [], [], [], [], [], [], [], []
    for k, sp in enumerate(sm.splines):
        pe = float(sp.parameters[-1])
        total = sp.get_total_arc_length()
        tot.append(total)
        for _ in range(6):
            a, b = np.sort(rng.uniform(0, pe, 2))
            if a >= b:
                continue
            gl_sp.append(k); gl_t0.append(a); gl_t1.append(b)
            gl_len.append(sp.get_arc_length(a, b))
        for frac in (0.0, 0.123, 0.5, 0.987, 1.0):
            s = total * frac
            inv_sp.append(k); inv_s.append(s)
            inv_t.append(float(sp.get_parameter_by_arc_length(s)))
    out["gl_spline"] = np.array(gl_sp, dtype=np.int64)
    out["gl_t0"], out["gl_t1"], out["gl_len"] = map(np.array, (gl_t0, gl_t1, gl_len))
    out["gl_total"] = np.array(tot)
    out["inv_spline"] = np.array(inv_sp, dtype=np.int64)
    out["inv_s"], out["inv_t"] = np.array(inv_s), np.array(inv_t)

    # ---- S1/S2
    sm.rebuild_tables()
    out["lut_d"] = np.array(sm.lookup_table.distances)
    out["lut_t"] = np.array(sm.lookup_table.parameters)
    out["total_length"] = np.array(float(sm.lookup_table.total_length))
    out["prop_k"] = np.array(sm._precomputed_properties["curvatures"])
    out["prop_h"] = np.array(sm._precomputed_properties["headings"])
    out["api_heading_snap"] = np.array([sm.get_heading(t) for t in tq])
    out["api_curv_snap"] = np.array([sm.get_curvature(t) for t in tq])
    dq = rng.uniform(-0.1, float(out["total_length"]) + 0.1, n_api)
    out["api_dist"] = dq
    out["api_dist_t"] = np.array([float(sm.distance_to_time(d)) for d in dq])

    # ---- S3 (same d accumulation as motion_profile_generator.py:112-122)
    L = sm.get_total_arc_length()
    d = 0
    ts = []
    while d < L:
        ts.append(float(sm.distance_to_time(d)))
        d += dd
    ts.append(float(sm.distance_to_time(L)))
    out["t"] = np.array(ts)
    out["kap"] = np.array([sm.get_curvature(t) for t in ts])
    out["th"] = np.array([sm.get_heading(t) for t in ts])

    # ---- S4/S5
    _grab.max_accels = _grab.boundary_map = None
    vel = mpg.forward_backward_pass(sm, c, dd)
    out["vel"] = np.array(vel, dtype=float)
    out["max_accels"] = np.array(_grab.max_accels, dtype=float)
    bm = _grab.boundary_map
    out["boundary_idx"] = np.array(sorted(bm.keys()), dtype=np.int64)
    out["boundary_val"] = np.array([bm[k] for k in sorted(bm.keys())], dtype=np.int64)
    assert len(vel) == len(ts)

    # ---- S6
    res = mpg.generate_motion_profile(sm, c, dt, dd)
    times, positions, lin, acc, head, ang, nodes_map, actions_map, coords = res
    out["times"] = np.array(times, dtype=float)
    out["positions"] = np.array(positions, dtype=float)
    out["linear_vels"] = np.array(lin, dtype=float)
    out["accelerations"] = np.array(acc, dtype=float)
    out["headings"] = np.array(head, dtype=float)
    out["angular_vels"] = np.array(ang, dtype=float)
    out["coords"] = np.array(coords, dtype=float).reshape(-1, 2)
    out["nodes_map"] = np.array(list(nodes_map) + [len(times)], dtype=np.int64)  # path.py:342
    out["actions_map"] = np.array(actions_map, dtype=np.int64)
    assert c.max_acc == cons["max_acc"] and c.max_dec == cons["max_acc"]

    path = os.path.join(OUT_DIR, f"case_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name:28s} N={n} S={len(sm.splines)} L={L:.4f} D={len(ts)} T={len(times)} "
          f"nodes_map={out['nodes_map'].tolist()} actions_map={out['actions_map'].tolist()} "
          f"-> {os.path.getsize(path) / 1024:.0f} KB")
    return out


CFG1_PX = [[300, 300], [700, 500], [1000, 1200], [1400, 900], [1700, 1500], [1200, 1700]]


def random_px(rng, n):
    while True:
        px = rng.uniform(15, 1985, (n, 2))
        if np.all(np.linalg.norm(np.diff(px, axis=0), axis=1) >= 30):
            return px


def misc_api():
    """motion_profile_angle / trapezoid profile / lerp known-answer vectors."""
    out = {}
    cases = []
    for cons in (FACTORY, REPOROOT, dict(max_vel=2.7, max_acc=5.3, max_jerk=1.0, track_width=0.9)):
        for deg in (30, -45, 90, -135, 170, 5, -1, 180):
            cases.append((cons, deg))
    heads, omegas, meta = [], [], []
    for cons, deg in cases:
        c = mpg.Constraints(cons["max_vel"], cons["max_acc"], cons["max_acc"], 0.8, cons["max_jerk"], cons["track_width"])
        h, w = mpg.motion_profile_angle(np.radians(deg), c, 0.01)
        meta.append([cons["max_vel"], cons["max_acc"], cons["track_width"], float(deg), float(np.radians(deg)), len(h)])
        heads.extend(float(x) for x in h)
        omegas.extend(float(x) for x in w)
    out["angle_meta"] = np.array(meta)
    out["angle_headings"] = np.array(heads)
    out["angle_omegas"] = np.array(omegas)
    trap_meta, trap_v = [], []
    for (v, a, dist) in [(4.0, 8.0, 1.0), (4.0, 8.0, 0.1), (3.0, 12.0, 5.0), (1.0, 1.0, 0.5), (2.5, 7.0, 0.8929)]:
        p = odm.generate_trapezoidal_profile(v, a, dist, 0.01)
        trap_meta.append([v, a, dist, len(p)])
        trap_v.extend(float(x) for x in p)
    out["trap_meta"] = np.array(trap_meta)
    out["trap_v"] = np.array(trap_v)
    xs = tuple(i * 0.005 for i in range(400))
    rng = np.random.default_rng(5)
    ys = tuple(float(x) for x in rng.uniform(0, 4, 400))
    q = np.concatenate([rng.uniform(-0.1, 2.1, 200), np.array(xs[:20]), [xs[-1], xs[-2]]])
    out["lerp_ys"] = np.array(ys)
    out["lerp_q"] = q
    out["lerp_out"] = np.array([float(mpg.lerp(x, xs, ys)) for x in q])
    lv = rng.uniform(-3, 3, 50); av = rng.uniform(-4, 4, 50)
    l, r = mpg.get_wheel_trajectory(list(lv), list(av), 12.5 / 12)
    out["wheel_lin"], out["wheel_ang"], out["wheel_l"], out["wheel_r"] = lv, av, np.array(l), np.array(r)
    np.savez_compressed(os.path.join(OUT_DIR, "misc_api.npz"), **out)
    print("misc_api written")


def main():
    only = set(sys.argv[1:])

    def want(nm):
        return not only or nm in only

    if want("cfg1_factory"):
        run_case("cfg1_factory", CFG1_PX, cons=FACTORY)
    if want("cfg1_reporoot"):
        run_case("cfg1_reporoot", CFG1_PX, cons=REPOROOT)
    if want("turn30"):
        run_case("turn30", CFG1_PX, {2: dict(turn=30)}, cons=REPOROOT)
    if want("turn_m60_wait"):
        run_case("turn_m60_wait", CFG1_PX, {3: dict(turn=-60, wait_time=0.25)}, cons=FACTORY)
    if want("turn135_rev"):
        run_case("turn135_rev", CFG1_PX, {2: dict(turn=135, is_reverse_node=True)}, cons=FACTORY)
    if want("reverse_mid"):
        run_case("reverse_mid", CFG1_PX, {3: dict(is_reverse_node=True)}, cons=REPOROOT)
    if want("node0_wait_rev"):
        run_case("node0_wait_rev", CFG1_PX, {0: dict(is_reverse_node=True, wait_time=0.37), 4: dict(is_reverse_node=True)},
                 cons=FACTORY)
    if want("stops_overrides"):
        run_case("stops_overrides", CFG1_PX,
                 {1: dict(stop=True), 2: dict(max_velocity=2.0, max_acceleration=5.0), 3: dict(max_velocity=3.1),
                  4: dict(stop=True, wait_time=0.1, max_acceleration=10.0), 0: dict(max_velocity=3.0, max_acceleration=6.0)},
                 cons=FACTORY)
    if want("actions"):
        run_case("actions", CFG1_PX, {2: dict(turn=-170)},
                 aps=[dict(t=0.6, stop=True, wait_time=0.3, max_acceleration=5.5), dict(t=2.4, max_velocity=2.2),
                      dict(t=3.7, wait_time=0.12, max_velocity=0, max_acceleration=0)], cons=REPOROOT)
    if want("tangents"):
        run_case("tangents", CFG1_PX,
                 {1: dict(tangent=[0.8, 0.6], incoming_magnitude=1.5, outgoing_magnitude=2.5),
                  3: dict(tangent=[0.0, 1.0], incoming_magnitude=2.0, outgoing_magnitude=1.0, turn=45),
                  4: dict(tangent=[-0.6, 0.8], incoming_magnitude=1.2, outgoing_magnitude=0.7, is_reverse_node=True)},
                 cons=FACTORY)
    if want("multi_split"):
        run_case("multi_split", CFG1_PX,
                 {1: dict(turn=90), 2: dict(is_reverse_node=True), 3: dict(turn=-45, is_reverse_node=True, stop=True),
                  4: dict(turn=30, wait_time=0.2)}, cons=FACTORY)
    if want("two_nodes"):
        run_case("two_nodes", [[400, 400], [1500, 1300]], cons=FACTORY)
    if want("three_nodes_turn"):
        run_case("three_nodes_turn", [[400, 400], [1500, 1300], [600, 1700]], {1: dict(turn=90)}, cons=REPOROOT)
    rng = np.random.default_rng(0)
    for k in range(3):
        px = random_px(rng, 8)
        if want(f"rand8_{k}"):
            run_case(f"rand8_{k}", px, cons=FACTORY if k % 2 == 0 else REPOROOT, seed=k)
    rng = np.random.default_rng(1)
    px = random_px(rng, 16)
    if want("rand16_0"):
        run_case("rand16_0", px, cons=FACTORY)
    # cfg5-style: mixed actions + per-path constraints
    rng = np.random.default_rng(3)
    for k in range(2):
        px = random_px(rng, 8)
        kw = {}
        for i in range(8):
            d = {}
            if 1 <= i <= 6 and rng.random() < 0.3:
                d["turn"] = int(rng.choice([30, -30, 45, -45, 90, -90, 135, -135]))
            if 1 <= i <= 6 and rng.random() < 0.2:
                d["stop"] = True
            if i <= 6 and rng.random() < 0.2:
                d["is_reverse_node"] = True
            if i <= 6 and rng.random() < 0.25:
                d["wait_time"] = float(rng.choice([0.1, 0.25, 0.5]))
            if rng.random() < 0.2:
                d["max_velocity"] = float(rng.uniform(1.5, 3.5))
            if rng.random() < 0.2:
                d["max_acceleration"] = float(rng.uniform(3, 7))
            if d:
                kw[i] = d
        aps = sorted([dict(t=float(rng.uniform(0.2, 5.8)), stop=bool(rng.random() < 0.3),
                           wait_time=float(rng.choice([0, 0.1, 0.25])),
                           max_velocity=float(rng.choice([0, 2.0, 3.0])),
                           max_acceleration=float(rng.choice([0, 4.0, 6.0]))) for _ in range(2)], key=lambda a: a["t"])
        cons = dict(max_vel=float(rng.uniform(2.5, 5.5)), max_acc=float(rng.uniform(5, 14)), max_jerk=16.0,
                    track_width=float(rng.uniform(9, 15) / 12))
        if want(f"mixed8_{k}"):
            run_case(f"mixed8_{k}", px, kw, aps=aps, cons=cons, seed=10 + k)
        # mirrored twin (path.py:596-600)
        pxm = px.copy(); pxm[:, 0] = 2000 - pxm[:, 0]
        kwm = {i: dict(d, **({"turn": -d["turn"]} if "turn" in d else {})) for i, d in kw.items()}
        if want(f"mixed8_{k}_mirror"):
            run_case(f"mixed8_{k}_mirror", pxm, kwm, aps=aps, cons=cons, seed=10 + k)
    if want("misc_api"):
        misc_api()


if __name__ == "__main__":
    main()
