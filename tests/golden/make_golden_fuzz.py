#!/usr/bin/env python
"""More reference pins, summary level: run the UNMODIFIED reference (see make_golden.py for the headless recipe) on
random mixed paths -- 3..12 nodes, turns, reverse, stops, waits, node / action-point overrides, user tangents, sorted and
UNSORTED action points, per-path constraints -- and store the packed inputs with integer outputs, summary values and
strided samples of every output stream (small: a few KB per path) in fuzz_reference.npz."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (installs the gui stubs, imports the reference)

OUT = os.path.join(mg.OUT_DIR, "fuzz_reference.npz")
N_CASES = int(sys.argv[1]) if len(sys.argv) > 1 else 48
MAXN, MAXA, STRIDE = 12, 3, 37


def main():
    rng = np.random.default_rng(2024)
    rec = {k: [] for k in ("n", "node_attr", "node_flags", "n_ap", "ap_attr", "ap_flags", "cons", "D", "T", "L", "t_end",
                           "vmax", "nodes_map", "n_nm", "actions_map", "n_am", "samples", "vel_samples", "status")}
    for case in range(N_CASES):
        n = int(rng.integers(3, MAXN + 1))
        px = mg.random_px(rng, n)
        pts = mg.px_to_ft(px)
        nodes = []
        for i in range(n):
            kw = {}
            if 1 <= i <= n - 2 and rng.random() < 0.2:
                kw["turn"] = int(rng.choice([30, -30, 45, -45, 90, -90, 135, -135, 170, -10]))
            if 1 <= i <= n - 2 and rng.random() < 0.15:
                kw["stop"] = True
            if i <= n - 2 and rng.random() < 0.15:
                kw["is_reverse_node"] = True
            if i <= n - 2 and rng.random() < 0.2:
                kw["wait_time"] = float(rng.choice([0.1, 0.25, 0.5, 0.005]))
            if rng.random() < 0.15:
                kw["max_velocity"] = float(rng.uniform(1.5, 3.5))
            if rng.random() < 0.15:
                kw["max_acceleration"] = float(rng.uniform(3, 7))
            if rng.random() < 0.15:
                a = rng.uniform(0, 2 * np.pi)
                kw["tangent"] = np.array([np.cos(a), np.sin(a)])
                kw["incoming_magnitude"] = float(rng.uniform(0.5, 3.0))
                kw["outgoing_magnitude"] = float(rng.uniform(0.5, 3.0))
            nodes.append(mg.Node(**kw))
        if case % 6 == 5:
            nodes[0].turn = 0
        A = int(rng.integers(0, MAXA + 1))
        ts = rng.uniform(0.2, n - 1.2, A)
        if case % 3 != 2:
            ts = np.sort(ts)                       # every third case keeps the list unsorted (later points may be blocked)
        aps = [mg.ActionPoint(float(t), stop=bool(rng.random() < 0.3), wait_time=float(rng.choice([0, 0.1, 0.25])),
                              max_velocity=float(rng.choice([0, 2.0, 3.0])), max_acceleration=float(rng.choice([0, 4.0, 6.0])))
               for t in ts]
        cons = dict(max_vel=float(rng.uniform(2.5, 5.5)), max_acc=float(rng.uniform(5, 14)), max_jerk=16.0,
                    track_width=float(rng.uniform(9, 15) / 12))
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
        # packed inputs in the layout of include/vap.h
        na = np.zeros((MAXN, 12)); nf = np.zeros(MAXN, dtype=np.int32)
        for i, nd in enumerate(nodes):
            na[i, 0:2] = pts[i]
            na[i, 2], na[i, 3], na[i, 4], na[i, 5] = nd.turn, nd.wait_time, nd.max_velocity, nd.max_acceleration
            if nd.tangent is not None:
                na[i, 6:8] = nd.tangent; na[i, 8] = nd.incoming_magnitude; na[i, 9] = nd.outgoing_magnitude
                nf[i] |= 4
            nf[i] |= (1 if nd.is_reverse_node else 0) | (2 if nd.stop else 0)
            na[i, 10], na[i, 11] = 1.0, 0.0
            if nd.turn != 0:
                angle = np.radians(nd.turn) + (np.pi if nd.is_reverse_node else 0)
                na[i, 10], na[i, 11] = np.cos(angle), np.sin(angle)
        apa = np.zeros((MAXA, 4)); apf = np.zeros(MAXA, dtype=np.int32)
        for k, a in enumerate(aps):
            apa[k] = (a.t, a.wait_time, a.max_velocity, a.max_acceleration); apf[k] = 2 if a.stop else 0
        nm = np.zeros(MAXN + 1, dtype=np.int64); am = np.zeros(MAXA, dtype=np.int64)
        full_nm = list(nodes_map) + [T]
        nm[: len(full_nm)] = full_nm; am[: len(actions_map)] = actions_map
        smp = np.full((8, 80), np.nan); idx = np.arange(0, T, STRIDE)[:80]
        smp[:, : len(idx)] = streams[:, idx]
        vs = np.full(80, np.nan); vidx = np.arange(0, len(vel), 211)[:80]; vs[: len(vidx)] = np.array(vel, dtype=float)[vidx]
        for k, v in (("n", n), ("node_attr", na), ("node_flags", nf), ("n_ap", A), ("ap_attr", apa), ("ap_flags", apf),
                     ("cons", [c.max_vel, c.max_acc, c.max_dec, 0.8, c.max_jerk, c.track_width]), ("D", len(vel)), ("T", T),
                     ("L", float(sm.get_total_arc_length())), ("t_end", float(times[-1])),
                     ("vmax", float(np.max(np.abs(lin)))), ("nodes_map", nm), ("n_nm", len(full_nm)), ("actions_map", am),
                     ("n_am", len(actions_map)), ("samples", smp), ("vel_samples", vs), ("status", 0)):
            rec[k].append(v)
        print(f"case {case:3d}: n={n} A={A} D={len(vel)} T={T} nodes_map={full_nm} actions_map={list(actions_map)}", flush=True)
    np.savez_compressed(OUT, stride=STRIDE, **{k: np.array(v) for k, v in rec.items()})
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KB")


if __name__ == "__main__":
    main()
