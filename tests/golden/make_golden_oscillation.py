#!/usr/bin/env python
"""Reference pins for positions that move BACKWARDS: with max_dec > 0.2 / dt a step near a stop has delta_pos < 0
(motion_profile_generator.py:578-581), the position oscillates across a node boundary, every forward crossing counts as a
node transition (:527-530) and once node_idx runs past the last node `spline_manager.nodes[node_idx]` raises IndexError.
Runs the UNMODIFIED reference (see make_golden.py for the headless recipe) on small paths with stop nodes and large max_dec
and stores packed inputs + outcome (status 0 with T / nodes_map / strided samples, or -2 for the IndexError) in
oscillation_reference.npz.  Candidates are screened with the C oracle first (fast); the reference decides."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import make_golden as mg  # noqa: E402  (installs the gui stubs, imports the reference)
import oracle  # noqa: E402

OUT = os.path.join(mg.OUT_DIR, "oscillation_reference.npz")
MAXN, MAXA, STRIDE = 6, 2, 29


def pack(pts, nodes, aps):
    na = np.zeros((MAXN, 12)); nf = np.zeros(MAXN, dtype=np.int32)
    for i, nd in enumerate(nodes):
        na[i, 0:2] = pts[i]
        na[i, 2], na[i, 3], na[i, 4], na[i, 5] = nd.turn, nd.wait_time, nd.max_velocity, nd.max_acceleration
        nf[i] |= (1 if nd.is_reverse_node else 0) | (2 if nd.stop else 0)
        na[i, 10], na[i, 11] = 1.0, 0.0
        if nd.turn != 0:
            angle = np.radians(nd.turn) + (np.pi if nd.is_reverse_node else 0)
            na[i, 10], na[i, 11] = np.cos(angle), np.sin(angle)
    apa = np.zeros((MAXA, 4)); apf = np.zeros(MAXA, dtype=np.int32)
    for k, a in enumerate(aps):
        apa[k] = (a.t, a.wait_time, a.max_velocity, a.max_acceleration); apf[k] = 2 if a.stop else 0
    return na, nf, apa, apf


def main():
    rng = np.random.default_rng(77)
    rec = {k: [] for k in ("n", "node_attr", "node_flags", "n_ap", "ap_attr", "ap_flags", "cons", "dt", "dd", "status", "T",
                           "t_end", "nodes_map", "n_nm", "actions_map", "n_am", "samples")}
    want = {(0, False): 4, (-2, False): 4, (0, True): 5, (-2, True): 3}     # (outcome, with turns / waits / action points)
    tries = 0
    while any(v > 0 for v in want.values()) and tries < 4000:
        tries += 1
        n = int(rng.integers(3, MAXN + 1))
        px = mg.random_px(rng, n)
        pts = mg.px_to_ft(px)
        nodes = [mg.Node() for _ in range(n)]
        rich = bool(tries % 2)           # every other candidate also has turns, waits and action points (some with a stop)
        for i in range(1, n - 1):
            if rng.random() < 0.7:
                nodes[i].stop = True
            if rich and rng.random() < 0.3:
                nodes[i].turn = int(rng.choice([45, -90, 135]))
            if rich and rng.random() < 0.3:
                nodes[i].wait_time = float(rng.choice([0.1, 0.25]))
        aps = []
        if rich:
            ts = np.sort(rng.uniform(0.2, n - 1.2, int(rng.integers(1, MAXA + 1))))
            aps = [mg.ActionPoint(float(t), stop=bool(rng.random() < 0.5), wait_time=float(rng.choice([0, 0.1])),
                                  max_velocity=0.0, max_acceleration=0.0) for t in ts]
        dt = float(rng.choice([0.01, 0.02, 0.04])); dd = float(rng.choice([0.005, 0.0025]))
        cons = [float(rng.uniform(0.5, 6.0)), float(10 ** rng.uniform(-0.5, 1.5)), float(10 ** rng.uniform(1.2, 2.3)), 0.8, 16.0,
                float(rng.uniform(9, 15) / 12)]
        na, nf, apa, apf = pack(pts, nodes, aps)
        A = len(aps)
        # screen with the oracle: keep paths whose position really moves backwards (non-monotone positions or an IndexError)
        try:
            r = oracle.full(na[:n], nf[:n], apa[:A] if A else None, apf[:A] if A else None, cons, dt=dt, dd=dd)
            st_o = 0
            if not np.any(np.diff(r["positions"]) < 0):
                continue
        except oracle.OracleError as e:
            st_o = int(e.code) if hasattr(e, "code") else -2
        if st_o not in (0, -2) or want.get((st_o, rich), 0) <= 0:
            continue
        sm = mg.QuinticHermiteSplineManager()
        assert sm.build_path(pts, nodes, aps)
        c = mg.mpg.Constraints(*cons)
        try:
            res = mg.mpg.generate_motion_profile(sm, c, dt, dd)
            status = 0
        except IndexError:
            res, status = None, -2
        if want.get((status, rich), 0) <= 0:
            continue
        want[(status, rich)] -= 1
        nm = np.zeros(MAXN + 2, dtype=np.int64); smp = np.full((8, 120), np.nan); T = 0; t_end = 0.0; n_nm = 0
        am = np.zeros(MAXA, dtype=np.int64); n_am = 0
        if status == 0:
            times, positions, lin, acc, head, ang, nodes_map, actions_map, coords = res
            T = len(times); t_end = float(times[-1])
            coords = np.array(coords, dtype=float).reshape(-1, 2)
            streams = np.stack([np.array(x, dtype=float) for x in (times, positions, lin, acc, head, ang)] + [coords[:, 0], coords[:, 1]])
            full_nm = list(nodes_map) + [T]
            n_nm = len(full_nm); nm[:n_nm] = full_nm
            n_am = len(actions_map); am[:n_am] = actions_map
            idx = np.arange(0, T, STRIDE)[:120]
            smp[:, : len(idx)] = streams[:, idx]
        for k, v in (("n", n), ("node_attr", na), ("node_flags", nf), ("n_ap", A), ("ap_attr", apa), ("ap_flags", apf), ("cons", cons),
                     ("dt", dt), ("dd", dd), ("status", status), ("T", T), ("t_end", t_end), ("nodes_map", nm), ("n_nm", n_nm),
                     ("actions_map", am), ("n_am", n_am), ("samples", smp)):
            rec[k].append(v)
        print(f"try {tries}: n={n} A={A} rich={rich} dt={dt} max_dec={cons[2]:.1f} oracle={st_o} reference={status} T={T} "
              f"nodes_map={nm[:n_nm].tolist()} actions_map={am[:n_am].tolist()}", flush=True)
    np.savez_compressed(OUT, stride=STRIDE, **{k: np.array(v) for k, v in rec.items()})
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KB; still wanted:", want)


if __name__ == "__main__":
    main()
