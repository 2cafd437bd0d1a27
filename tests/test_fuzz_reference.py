"""48 more pins against the UNMODIFIED reference (tests/golden/make_golden_fuzz.py): random mixed paths with turns,
reverse, stops, waits, overrides, user tangents, sorted and unsorted action points, per-path constraints, 3..12 nodes.
Integer outputs must be exact; sampled values of every output stream within the north-star tolerances."""
import os

import numpy as np
import pytest

from golden_util import GOLDEN_DIR

FUZZ = os.path.join(GOLDEN_DIR, "fuzz_reference.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(FUZZ), reason="fuzz fixture not generated")


def _load():
    return dict(np.load(FUZZ))


def _check(f, i, D, T, L, nodes_map, actions_map, streams, vel):
    assert D == int(f["D"][i]), (i, D, int(f["D"][i]))
    assert T == int(f["T"][i]), (i, T, int(f["T"][i]))
    assert list(nodes_map) == f["nodes_map"][i][: int(f["n_nm"][i])].tolist(), i
    assert list(actions_map) == f["actions_map"][i][: int(f["n_am"][i])].tolist(), i
    np.testing.assert_allclose(L, f["L"][i], rtol=1e-13)
    stride = int(f["stride"])
    idx = np.arange(0, T, stride)[:80]
    want = f["samples"][i][:, : len(idx)]
    got = np.stack([s[idx] for s in streams])
    tol = {0: (1e-6, 1e-12), 1: (1e-9, 1e-10), 2: (1e-6, 1e-9), 3: (1e-6, 1e-6), 4: (1e-9, 1e-10), 5: (1e-6, 1e-9),
           6: (1e-9, 1e-10), 7: (1e-9, 1e-10)}
    for s in range(8):
        np.testing.assert_allclose(got[s], want[s], rtol=tol[s][0], atol=tol[s][1], err_msg=f"case {i} stream {s}")
    vidx = np.arange(0, D, 211)[:80]
    np.testing.assert_allclose(vel[vidx], f["vel_samples"][i][: len(vidx)], rtol=1e-6, atol=1e-12)


def test_oracle_against_reference_fuzz(oracle_mod):
    o = oracle_mod
    o.set_sq_mode(0)
    f = _load()
    for i in range(len(f["n"])):
        n, A = int(f["n"][i]), int(f["n_ap"][i])
        r = o.full(f["node_attr"][i][:n], f["node_flags"][i][:n], f["ap_attr"][i][:A], f["ap_flags"][i][:A], f["cons"][i])
        streams = [r[k] for k in ("times", "positions", "linear_vels", "accelerations", "headings", "angular_vels", "x", "y")]
        _check(f, i, r["D"], r["T"], r["summary"][1], r["nodes_map"], r["actions_map"], streams, r["vel"])
        np.testing.assert_allclose(r["summary"][2], f["t_end"][i], rtol=1e-6)
        np.testing.assert_allclose(r["summary"][3], f["vmax"][i], rtol=1e-6)


MANY = os.path.join(GOLDEN_DIR, "many_regimes_reference.npz")


def test_oracle_against_reference_many_regimes(oracle_mod):
    """Two 300-node paths (tests/golden/make_golden_many_regimes.py, the reference itself): one acceleration regime per
    node, the only real overrides behind the 256th node (a stop before one of them, an action-point override after the
    other).  Head AND tail samples of every stream: the overrides act at the end of the path."""
    o = oracle_mod
    o.set_sq_mode(0)
    f = dict(np.load(MANY))
    stride = int(f["stride"])
    tol = {0: (1e-6, 1e-12), 1: (1e-9, 1e-10), 2: (1e-6, 1e-9), 3: (1e-6, 1e-6), 4: (1e-9, 1e-10), 5: (1e-6, 1e-9),
           6: (1e-9, 1e-10), 7: (1e-9, 1e-10)}
    for i in range(len(f["n"])):
        n, A = int(f["n"][i]), int(f["n_ap"][i])
        assert n == 300
        r = o.full(f["node_attr"][i][:n], f["node_flags"][i][:n], f["ap_attr"][i][:A], f["ap_flags"][i][:A], f["cons"][i])
        streams = [r[k] for k in ("times", "positions", "linear_vels", "accelerations", "headings", "angular_vels", "x", "y")]
        _check(f, i, r["D"], r["T"], r["summary"][1], r["nodes_map"], r["actions_map"], streams, r["vel"])
        T, D = r["T"], r["D"]
        tidx = np.arange(T - 1, -1, -stride)[:80]
        for s_ in range(8):
            np.testing.assert_allclose(streams[s_][tidx], f["tail_samples"][i][s_][: len(tidx)], rtol=tol[s_][0],
                                       atol=tol[s_][1], err_msg=f"case {i} tail stream {s_}")
        tvi = np.arange(D - 1, -1, -211)[:80]
        np.testing.assert_allclose(r["vel"][tvi], f["tail_vel_samples"][i][: len(tvi)], rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(r["summary"][2], f["t_end"][i], rtol=1e-6)
        np.testing.assert_allclose(r["summary"][3], f["vmax"][i], rtol=1e-6)


@pytest.mark.gpu
def test_engine_against_reference_fuzz():
    import torch
    from vexautonomousplanner_b200.engine import Engine
    from vexautonomousplanner_b200.packing import PackedPaths
    f = _load()
    B = len(f["n"])
    packed = PackedPaths(np.ascontiguousarray(f["node_attr"]), np.ascontiguousarray(f["node_flags"]).astype(np.int32),
                         f["n"].astype(np.int32), np.ascontiguousarray(f["ap_attr"]),
                         np.ascontiguousarray(f["ap_flags"]).astype(np.int32), f["n_ap"].astype(np.int32),
                         np.ascontiguousarray(f["cons"]))
    for impl in (dict(), dict(velocity_impl="serial", time_impl="serial")):
        eng = Engine("cuda:0", **impl)
        res = eng.profile(eng.upload(packed))
        torch.cuda.synchronize()
        assert (res.status == 0).all()
        for i in range(B):
            p = res.path(i)
            streams = [p[k] for k in ("times", "positions", "linear_vels", "accelerations", "headings", "angular_vels", "x", "y")]
            _check(f, i, int(res.n_samples[i]), len(p["times"]), float(res.summary[i, 1]), p["nodes_map"], p["actions_map"],
                   streams, p["vel"])


@pytest.mark.gpu
def test_engine_against_reference_many_regimes():
    """The same two 300-node paths through the engine: more regimes than a pre-pass CTA has threads, real overrides only
    behind the 256th; integer outputs exact, sampled streams within the north-star tolerances."""
    import torch
    from vexautonomousplanner_b200.engine import Engine
    from vexautonomousplanner_b200.packing import PackedPaths
    f = dict(np.load(MANY))
    packed = PackedPaths(np.ascontiguousarray(f["node_attr"]), np.ascontiguousarray(f["node_flags"]).astype(np.int32),
                         f["n"].astype(np.int32), np.ascontiguousarray(f["ap_attr"]),
                         np.ascontiguousarray(f["ap_flags"]).astype(np.int32), f["n_ap"].astype(np.int32),
                         np.ascontiguousarray(f["cons"]))
    eng = Engine("cuda:0")
    res = eng.profile(eng.upload(packed))
    torch.cuda.synchronize()
    assert (res.status == 0).all()
    for i in range(len(f["n"])):
        p = res.path(i)
        streams = [p[k] for k in ("times", "positions", "linear_vels", "accelerations", "headings", "angular_vels", "x", "y")]
        _check(f, i, int(res.n_samples[i]), len(p["times"]), float(res.summary[i, 1]), p["nodes_map"], p["actions_map"],
               streams, p["vel"])


OSC = os.path.join(GOLDEN_DIR, "oscillation_reference.npz")
_TOL = {0: (1e-6, 1e-12), 1: (1e-9, 1e-10), 2: (1e-6, 1e-9), 3: (1e-6, 1e-6), 4: (1e-9, 1e-10), 5: (1e-6, 1e-9),
        6: (1e-9, 1e-10), 7: (1e-9, 1e-10)}


def _check_osc(f, i, status, T, nodes_map, actions_map, streams):
    """One case of oscillation_reference.npz (tests/golden/make_golden_oscillation.py: the unmodified reference on paths
    whose position moves backwards near a stop, max_dec > 0.2 / dt; half of them with turns, waits and action points):
    either the reference raised IndexError (-2: more node transitions than nodes) or it returned, and then T, nodes_map --
    extra transitions included -- and actions_map are exact."""
    assert status == int(f["status"][i]), (i, status, int(f["status"][i]))
    if status != 0:
        return
    assert T == int(f["T"][i]), (i, T, int(f["T"][i]))
    assert list(nodes_map) == f["nodes_map"][i][: int(f["n_nm"][i])].tolist(), i
    assert list(actions_map) == f["actions_map"][i][: int(f["n_am"][i])].tolist(), i
    idx = np.arange(0, T, int(f["stride"]))[:120]
    want = f["samples"][i][:, : len(idx)]
    got = np.stack([s[idx] for s in streams])
    assert np.any(np.diff(streams[1]) < 0), "the case is supposed to have steps that move backwards"
    for s in range(8):
        np.testing.assert_allclose(got[s], want[s], rtol=_TOL[s][0], atol=_TOL[s][1], err_msg=f"case {i} stream {s}")


def test_oracle_against_reference_oscillation(oracle_mod):
    o = oracle_mod
    o.set_sq_mode(0)
    f = dict(np.load(OSC))
    assert sorted(set(f["status"].tolist())) == [-2, 0] and len(f["n"]) == 16
    for i in range(len(f["n"])):
        n, A = int(f["n"][i]), int(f["n_ap"][i])
        try:
            r = o.full(f["node_attr"][i][:n], f["node_flags"][i][:n], f["ap_attr"][i][:A] if A else None,
                       f["ap_flags"][i][:A] if A else None, f["cons"][i], dt=float(f["dt"][i]), dd=float(f["dd"][i]))
        except o.OracleError as e:
            _check_osc(f, i, e.code, 0, [], [], None)
            continue
        streams = [r[k] for k in ("times", "positions", "linear_vels", "accelerations", "headings", "angular_vels", "x", "y")]
        _check_osc(f, i, 0, r["T"], r["nodes_map"], r["actions_map"], streams)
        np.testing.assert_allclose(r["summary"][2], f["t_end"][i], rtol=1e-6)


@pytest.mark.gpu
def test_engine_against_reference_oscillation():
    """The same cases through the engine, fast and reference-shaped kernels: IndexError cases come back as status -2, the
    others with the reference's T, nodes_map (the extra node transitions of an oscillating position included) and
    actions_map."""
    import torch
    from vexautonomousplanner_b200.engine import Engine
    from vexautonomousplanner_b200.packing import PackedPaths
    f = dict(np.load(OSC))
    B = len(f["n"])
    for dt, dd in sorted(set(zip(f["dt"].tolist(), f["dd"].tolist()))):
        sel = [i for i in range(B) if f["dt"][i] == dt and f["dd"][i] == dd]
        packed = PackedPaths(np.ascontiguousarray(f["node_attr"][sel]), np.ascontiguousarray(f["node_flags"][sel]).astype(np.int32),
                             f["n"][sel].astype(np.int32), np.ascontiguousarray(f["ap_attr"][sel]),
                             np.ascontiguousarray(f["ap_flags"][sel]).astype(np.int32), f["n_ap"][sel].astype(np.int32),
                             np.ascontiguousarray(f["cons"][sel]))
        for impl in (dict(), dict(velocity_impl="serial", time_impl="serial")):
            eng = Engine("cuda:0", dt=dt, dd=dd, **impl)
            res = eng.profile(eng.upload(packed))
            torch.cuda.synchronize()
            for j, i in enumerate(sel):
                st = int(res.status[j])
                if st != 0:
                    _check_osc(f, i, st, 0, [], [], None)
                    continue
                p = res.path(j)
                streams = [p[k] for k in ("times", "positions", "linear_vels", "accelerations", "headings", "angular_vels", "x", "y")]
                _check_osc(f, i, 0, len(p["times"]), p["nodes_map"], p["actions_map"], streams)
