#!/usr/bin/env python
"""Direct fuzz of the C oracle against the UNMODIFIED reference (needs /root/reference, so it runs in the build container
only; see make_golden.py for the headless recipe): random paths with turns, reverse, stops, waits, overrides, user tangents and (sorted or
unsorted) action points, constraints far from the factory values (max_vel 0.3 ... 14, max_acc 0.2 ... 40, max_dec up to 10^argv[3]), three dt
and three dd values.  Compares status (incl. the reference's IndexError / ValueError), T, nodes_map, actions_map exactly and
the streams within the north-star tolerances.  usage: fuzz_oracle_vs_reference.py [seed] [cases] [log10 of the largest
max_dec].  Last runs: seeds 11, 12, 21, 31 and 41 (the last with user tangents and unsorted action points), 24 + 40 + 60 + 80 + 80 (+ 60 more with seed 51)
cases, max_dec up to 18 / 160 / 160 / 20 / 20: 0 mismatches (two long, slowly
accelerating paths of seed 21 deviate by 2.3e-10 ft in position and x after 4000 rows: the <= 4 ulp of the tables, amplified)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import make_golden as mg
import oracle
oracle.set_sq_mode(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 11)
NC = int(sys.argv[2]) if len(sys.argv) > 2 else 24
MAXN, MAXA = 8, 3
bad = 0
t0 = time.time()
for case in range(NC):
    n = int(rng.integers(3, MAXN + 1))
    px = mg.random_px(rng, n); pts = mg.px_to_ft(px)
    nodes = []
    for i in range(n):
        kw = {}
        if 1 <= i <= n - 2 and rng.random() < 0.25: kw["turn"] = int(rng.choice([30, -45, 90, -135, 170]))
        if 1 <= i <= n - 2 and rng.random() < 0.2: kw["stop"] = True
        if i <= n - 2 and rng.random() < 0.15: kw["is_reverse_node"] = True
        if i <= n - 2 and rng.random() < 0.2: kw["wait_time"] = float(rng.choice([0.1, 0.25, 0.005]))
        if rng.random() < 0.15: kw["max_velocity"] = float(rng.uniform(0.5, 6))
        if rng.random() < 0.15: kw["max_acceleration"] = float(rng.uniform(0.5, 20))
        if rng.random() < 0.12:                      # user tangent with its own magnitudes
            a_ = rng.uniform(0, 2 * np.pi)
            kw["tangent"] = np.array([np.cos(a_), np.sin(a_)])
            kw["incoming_magnitude"] = float(rng.uniform(0.5, 3.0)); kw["outgoing_magnitude"] = float(rng.uniform(0.5, 3.0))
        nodes.append(mg.Node(**kw))
    if nodes[0].turn != 0 and rng.random() < 0.7: nodes[0].turn = 0
    A = int(rng.integers(0, MAXA + 1))
    ts = rng.uniform(0.2, n - 1.2, A)
    if case % 3 != 2: ts = np.sort(ts)              # every third case keeps its action points UNSORTED
    aps = [mg.ActionPoint(float(t), stop=bool(rng.random() < 0.3), wait_time=float(rng.choice([0, 0.1])),
                          max_velocity=float(rng.choice([0, 2.0])), max_acceleration=float(rng.choice([0, 4.0]))) for t in ts]
    cons = [float(rng.uniform(0.3, 14.0)), float(10 ** rng.uniform(-0.7, 1.6)), float(10 ** rng.uniform(-0.7, float(sys.argv[3]) if len(sys.argv) > 3 else 1.25)), 0.8, 16.0,
            float(rng.uniform(0.4, 2.5))]
    dt = float(rng.choice([0.01, 0.02, 0.005])); dd = float(rng.choice([0.005, 0.01, 0.0025]))
    na = np.zeros((n, 12)); nf = np.zeros(n, dtype=np.int32)
    for i, nd in enumerate(nodes):
        na[i, 0:2] = pts[i]; na[i, 2], na[i, 3], na[i, 4], na[i, 5] = nd.turn, nd.wait_time, nd.max_velocity, nd.max_acceleration
        nf[i] |= (1 if nd.is_reverse_node else 0) | (2 if nd.stop else 0)
        if nd.tangent is not None:
            na[i, 6:8] = nd.tangent; na[i, 8] = nd.incoming_magnitude; na[i, 9] = nd.outgoing_magnitude; nf[i] |= 4
        na[i, 10], na[i, 11] = 1.0, 0.0
        if nd.turn != 0:
            ang = np.radians(nd.turn) + (np.pi if nd.is_reverse_node else 0); na[i, 10], na[i, 11] = np.cos(ang), np.sin(ang)
    apa = np.zeros((max(A, 1), 4)); apf = np.zeros(max(A, 1), dtype=np.int32)
    for k, a in enumerate(aps): apa[k] = (a.t, a.wait_time, a.max_velocity, a.max_acceleration); apf[k] = 2 if a.stop else 0
    sm = mg.QuinticHermiteSplineManager()
    okb = sm.build_path(pts, nodes, aps)
    c = mg.mpg.Constraints(*cons)
    try:
        res = mg.mpg.generate_motion_profile(sm, c, dt, dd) if okb else None
        st_r = 0 if okb else -1
    except IndexError: res, st_r = None, -2
    except ValueError: res, st_r = None, -3
    except Exception as e: res, st_r = None, -99; print("  reference raised", type(e).__name__, e)
    try:
        r = oracle.full(na, nf, apa[:A] if A else None, apf[:A] if A else None, cons, dt=dt, dd=dd)
        st_o = 0
    except oracle.OracleError as e: r, st_o = None, e.code
    msg = ""
    if st_r != st_o: msg = f"STATUS ref {st_r} oracle {st_o}"
    elif st_r == 0:
        times, positions, lin, acc, head, ang, nodes_map, actions_map, coords = res
        T = len(times)
        if T != r["T"]: msg = f"T ref {T} oracle {r['T']}"
        elif list(nodes_map) + [T] != r["nodes_map"].tolist(): msg = f"nodes_map ref {list(nodes_map)+[T]} oracle {r['nodes_map'].tolist()}"
        elif list(actions_map) != r["actions_map"].tolist(): msg = f"actions_map ref {list(actions_map)} oracle {r['actions_map'].tolist()}"
        else:
            coords = np.array(coords, dtype=float).reshape(-1, 2)
            for nm_, a_, tol in (("times", times, 1e-6), ("positions", positions, 1e-9), ("linear_vels", lin, 1e-6), ("headings", head, 1e-9)):
                if not np.allclose(np.array(a_, dtype=float), r[nm_], rtol=tol, atol=1e-9): msg = f"stream {nm_} differs"; break
            if not msg and not np.allclose(coords[:, 0], r["x"], rtol=1e-9, atol=1e-9):   # north star: 1e-9 on positions
                dx = np.abs(coords[:, 0] - r["x"]); k = int(np.argmax(dx))
                dp = np.abs(np.array(positions, dtype=float) - r["positions"])
                msg = (f"x differs: max |dx| {dx.max():.3e} at row {k} of {T} (x = {coords[k, 0]!r} vs {r['x'][k]!r}), "
                       f"max |dpos| {dp.max():.3e}, |dpos| there {dp[k]:.3e}")
    bad += bool(msg)
    print(f"case {case}: n={n} A={A} dt={dt} dd={dd} cons={[round(x,2) for x in cons[:3]]} ref={st_r} {'OK' if not msg else 'MISMATCH ' + msg}", flush=True)
print("MISMATCHES", bad, "in", NC, "cases,", round(time.time() - t0), "s")
