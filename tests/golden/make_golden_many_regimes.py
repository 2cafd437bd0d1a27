#!/usr/bin/env python
"""Reference pins for paths with MORE THAN 256 acceleration regimes (one per node): run the UNMODIFIED reference (see
make_golden.py for the headless recipe) on two 300-node paths whose only real max_acceleration overrides sit behind the
256th node, and store packed inputs, integer outputs, summary values and strided samples in many_regimes_reference.npz
(same record layout as fuzz_reference.npz)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (installs the gui stubs, imports the reference)

OUT = os.path.join(mg.OUT_DIR, "many_regimes_reference.npz")
MAXN, MAXA, STRIDE = 300, 1, 37


def path_px(case):
    """270 closely spaced nodes on a gentle wave, then 30 widely spaced ones on a nearly straight line (where the
    acceleration limit, not the curvature, shapes the profile)."""
    x, y, px = 100.0, 1000.0, []
    for i in range(MAXN):
        px.append((x, y + (25.0 if case == 0 else 18.0) * np.sin(0.35 * i)))
        x += 32.0 if i < 269 else 260.0
    return np.array(px)


def main():
    rec = {k: [] for k in ("n", "node_attr", "node_flags", "n_ap", "ap_attr", "ap_flags", "cons", "D", "T", "L", "t_end",
                           "vmax", "nodes_map", "n_nm", "actions_map", "n_am", "samples", "vel_samples", "status")}
    for case in range(2):
        n = MAXN
        pts = mg.px_to_ft(path_px(case))
        cons = dict(mg.FACTORY)
        nodes = []
        for i in range(n):
            kw = {}
            if case == 0:
                kw["max_acceleration"] = cons["max_acc"]            # a regime per node, all equal to the path's own ...
                if i == 280:
                    kw["max_acceleration"] = 3.0                    # ... except one, behind the 256th
                if i == 279:
                    kw["stop"] = True
            elif i == 285:
                kw["max_acceleration"] = 5.0
            elif i == 284:
                kw["stop"] = True                                   # so that the robot accelerates under the override
            nodes.append(mg.Node(**kw))
        aps = [] if case == 0 else [mg.ActionPoint(290.5, stop=True, wait_time=0.0, max_velocity=0.0, max_acceleration=6.0)]
        A = len(aps)
        sm = mg.QuinticHermiteSplineManager()
        assert sm.build_path(pts, nodes, aps)
        c = mg.mpg.Constraints(cons["max_vel"], cons["max_acc"], cons["max_acc"], 0.8, cons["max_jerk"], cons["track_width"])
        vel = mg.mpg.forward_backward_pass(sm, c, 0.005)
        res = mg.mpg.generate_motion_profile(sm, c, 0.01, 0.005)
        times, positions, lin, acc, head, ang, nodes_map, actions_map, coords = res
        T = len(times)
        coords = np.array(coords, dtype=float).reshape(-1, 2)
        streams = np.stack([np.array(times, dtype=float), np.array(positions, dtype=float), np.array(lin, dtype=float),
                            np.array(acc, dtype=float), np.array(head, dtype=float), np.array(ang, dtype=float),
                            coords[:, 0], coords[:, 1]])
        na = np.zeros((MAXN, 12)); nf = np.zeros(MAXN, dtype=np.int32)
        for i, nd in enumerate(nodes):
            na[i, 0:2] = pts[i]
            na[i, 2], na[i, 3], na[i, 4], na[i, 5] = nd.turn, nd.wait_time, nd.max_velocity, nd.max_acceleration
            nf[i] |= (1 if nd.is_reverse_node else 0) | (2 if nd.stop else 0)
            na[i, 10], na[i, 11] = 1.0, 0.0
        apa = np.zeros((MAXA, 4)); apf = np.zeros(MAXA, dtype=np.int32)
        for k, a in enumerate(aps):
            apa[k] = (a.t, a.wait_time, a.max_velocity, a.max_acceleration); apf[k] = 2 if a.stop else 0
        nm = np.zeros(MAXN + 1, dtype=np.int64); am = np.zeros(MAXA, dtype=np.int64)
        full_nm = list(nodes_map) + [T]
        nm[: len(full_nm)] = full_nm; am[: len(actions_map)] = actions_map
        smp = np.full((8, 80), np.nan); idx = np.arange(0, T, STRIDE)[:80]
        smp[:, : len(idx)] = streams[:, idx]
        vs = np.full(80, np.nan); vidx = np.arange(0, len(vel), 211)[:80]; vs[: len(vidx)] = np.array(vel, dtype=float)[vidx]
        # the tail of the profile (where the overrides act) as well: the last 80 strided rows and velocities
        tidx = np.arange(T - 1, -1, -STRIDE)[:80]
        tsm = np.full((8, 80), np.nan); tsm[:, : len(tidx)] = streams[:, tidx]
        tv = np.full(80, np.nan); tvi = np.arange(len(vel) - 1, -1, -211)[:80]; tv[: len(tvi)] = np.array(vel, dtype=float)[tvi]
        rec.setdefault("tail_samples", []).append(tsm)
        rec.setdefault("tail_vel_samples", []).append(tv)
        for k, v in (("n", n), ("node_attr", na), ("node_flags", nf), ("n_ap", A), ("ap_attr", apa), ("ap_flags", apf),
                     ("cons", [c.max_vel, c.max_acc, c.max_dec, 0.8, c.max_jerk, c.track_width]), ("D", len(vel)), ("T", T),
                     ("L", float(sm.get_total_arc_length())), ("t_end", float(times[-1])),
                     ("vmax", float(np.max(np.abs(lin)))), ("nodes_map", nm), ("n_nm", len(full_nm)), ("actions_map", am),
                     ("n_am", len(actions_map)), ("samples", smp), ("vel_samples", vs), ("status", 0)):
            rec[k].append(v)
        print(f"case {case}: n={n} A={A} D={len(vel)} T={T} n_nodes_map={len(full_nm)} actions_map={list(actions_map)}", flush=True)
    np.savez_compressed(OUT, stride=STRIDE, **{k: np.array(v) for k, v in rec.items()})
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
