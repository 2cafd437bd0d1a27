"""'Next' rows f1-f3 of SURVEY.md section 8: export rows / text, node JSON codec, interactive queries."""
import json
import math

import numpy as np
import pytest

from golden_util import load_case


def test_json_codec_roundtrip_cpu():
    from vexautonomousplanner_b200 import export as ex
    pts = np.array([[300.0, 300.0], [700.0, 500.0], [1000.0, 1200.0]])
    nodes = [ex.RouteNode() for _ in range(3)]
    nodes[0].is_start_node = True; nodes[2].is_end_node = True
    nodes[1].turn = 45; nodes[1].wait_time = 0.25; nodes[1].is_reverse_node = True
    nodes[1].tangent = np.array([0.6, 0.8]); nodes[1].incoming_magnitude = 1.5; nodes[1].outgoing_magnitude = 2.0
    nodes[1].action_values = [3, 0]
    ap = ex.RouteActionPoint(0.7); ap.stop = True; ap.wait_time = 0.1; ap.action_values = [1]
    s = ex.nodes_to_json(pts, nodes, [ap], [[500.0, 400.0]])
    data = json.loads(s)
    assert data[0][0][0] == ((300.0 / 2000) - 0.5) * 145.308474301
    assert data[0][1][4:8] == [1, 0, 45, 0.25] and data[0][1][8] == [0.6, 0.8] and data[0][1][11:] == [3, 0]
    assert "," in s and " " not in s                       # separators=(",", ":")
    p2, n2, a2, apx = ex.load_nodes(s)
    np.testing.assert_allclose(p2, pts, rtol=0, atol=1e-9)
    assert [n.turn for n in n2] == [0, 45, 0] and n2[1].is_reverse_node and n2[0].is_start_node and n2[2].is_end_node
    assert n2[1].tangent.tolist() == [0.6, 0.8] and n2[1].action_values == [3, 0]
    assert a2[0].t == 0.7 and a2[0].stop == 1 and a2[0].action_values == [1]   # stop stays the stored int (path.py:636)
    np.testing.assert_allclose(apx, [[500.0, 400.0]], atol=1e-9)
    ft = ex.px_points_to_ft(pts)
    assert ft[0, 0] == (300.0 / 2000 - 0.5) * 12.1090395251
    # legacy single-list files (gui/path.py:604-609)
    p3, n3, a3, _ = ex.load_nodes(json.dumps(data[0]))
    assert len(n3) == 3 and a3 == []


def test_splice_and_format_cpu():
    from vexautonomousplanner_b200 import export as ex
    traj = [[0, 0, 1.5, -2.0, 0.25, 12.0, 0.0], [0, np.float64(0.01), np.float64(1.25), 3.0, 0.5, 6.0, -0.125]]
    data = ex.splice_action_rows(traj, [0, 1, 2], [[7], [8], [9]], [1], [[5, 5]])
    # nodes rows go to map[i] + i = 0, 2, 4; the action row then to actions_map[0] + 0 = 1 (gui_manager.py:307-310)
    assert data[0] == [1, 7] and data[1] == [1, 5, 5] and data[3] == [1, 8] and data[-1] == [1, 9] and len(data) == 6
    txt = ex.format_rows(data)
    assert txt.splitlines()[0] == "1 7 " and txt.endswith("\n")
    assert "0 0.01 1.25 3.0 0.5 6.0 -0.125 " in txt


def test_routes_header_writer_cpu(tmp_path):
    """f1, legacy routes.h export: identical files to the reference's own fill_template (gui_manager.py:442-507) on the
    committed cases (tests/golden/gen_routes_header.py ran the reference method itself): new file, second route appended
    before #endif, route replaced in place, empty file, file without #endif (entry dropped, as the reference does),
    indented declaration."""
    import os
    from vexautonomousplanner_b200.export import update_routes_header, write_routes_header, routes_header_entry
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "routes_header.json")))
    assert len(cases) == 6
    for c in cases:
        before = None if c["before"] is None else c["before"].splitlines(keepends=True)
        assert "".join(update_routes_header(before, c["name"], c["rows"])) == c["after"], c["tag"]
        path = tmp_path / (c["tag"] + ".h")
        if c["before"] is not None:
            path.write_text(c["before"])
        write_routes_header(str(path), c["name"], c["rows"])
        assert path.read_text() == c["after"], c["tag"]
    assert routes_header_entry("r", [[1, 2]]) == "std::vector<std::vector<double>> r = {{1, 2}};\n"


@pytest.mark.gpu
def test_export_rows_on_device():
    import torch
    from vexautonomousplanner_b200 import export as ex, synth
    from vexautonomousplanner_b200.engine import Engine
    eng = Engine("cuda:0")
    packed = synth.mixed_paths(40, 8, seed=61)
    res = eng.profile(eng.upload(packed))
    rows, offsets = ex.export_rows(eng, res)
    torch.cuda.synchronize()
    rows, offsets = rows.cpu().numpy(), offsets.cpu().numpy()
    assert offsets[-1] == int(res.n_out.sum()) == rows.shape[0]
    for b in (0, 17, 39):
        p = res.path(b)
        want = np.stack([np.zeros_like(p["times"]), p["times"], p["x"] * 12, p["y"] * -12, p["headings"],
                         p["linear_vels"] * 12, p["angular_vels"]], axis=1)
        got = rows[offsets[b]:offsets[b + 1]]
        assert got.shape == want.shape and np.array_equal(got.view(np.int64), want.view(np.int64))
    # text: the device formatter must reproduce the reference's own expression f"{v} " on the same numbers
    db = eng.upload(packed)
    dev_text = ex.export_text_device(eng, res, db)
    kinds = ex.row_kinds(eng, res, db).cpu().numpy()
    n_int = 0
    for b in (0, 17, 39):
        p = res.path(b)
        txt = ex.trajectory_text(eng, res, b, [[i] for i in range(8)], [[9], [9]], device_text=dev_text)
        kb = kinds[b]
        n_int += int(((kb[: len(p["times"])] & ex.RK_INSERTED) != 0).sum())
        # the expression of save_nodes_to_file (gui_manager.py:284-295) on the reference's element types: int 0 where
        # the reference's lists hold ints (0 * 12 is the int 0), numpy floats elsewhere
        traj = [[0, (0 if kb[k] & ex.RK_TIME_INT else np.float64(p["times"][k])),
                 np.float64(p["x"][k] * 12), np.float64(p["y"][k] * -12), np.float64(p["headings"][k]),
                 (0 * 12 if kb[k] & ex.RK_INSERTED else np.float64(p["linear_vels"][k] * 12)),
                 (0 if kb[k] & ex.RK_OMEGA_INT else np.float64(p["angular_vels"][k]))] for k in range(len(p["times"]))]
        want = ex.format_rows(ex.splice_action_rows(traj, p["nodes_map"], [[i] for i in range(8)], p["actions_map"],
                                                    [[9], [9]]))
        assert txt == want, b
        assert txt.splitlines()[0] == "1 0 "
    assert n_int > 0          # the sampled paths do contain inserted rows


@pytest.mark.gpu
def test_device_repr_matches_python():
    """The device formatter is repr(float): random bit patterns, typical trajectory magnitudes and the edge cases."""
    import random
    import struct
    import torch
    from vexautonomousplanner_b200 import export as ex
    from vexautonomousplanner_b200.engine import Engine
    eng = Engine("cuda:0")
    random.seed(11)
    vals = [0.0, -0.0, 1.0, -1.5, 0.1, 1e16, 1e15, 9999999999999998.0, 1e-4, 1e-5, 123456789012345678.0, 5e-324,
            2.2250738585072014e-308, 1.7976931348623157e308, float("inf"), float("-inf"), 0.3, 2 / 3, 1e22, 1e23,
            0.30000000000000004, 9007199254740993.0, 5e-5, 0.001]
    while len(vals) < 300000:
        x = struct.unpack("<d", struct.pack("<Q", random.getrandbits(64)))[0]
        if x == x:
            vals.append(x)
    rng = np.random.default_rng(2)
    vals += list(rng.uniform(-80, 80, 200000)) + list(rng.uniform(-1e-3, 1e-3, 50000)) + [i * 0.01 for i in range(4000)]
    got = ex.format_doubles_device(eng, torch.tensor(vals, dtype=torch.float64, device="cuda:0"))
    bad = [(g, repr(float(v))) for g, v in zip(got, vals) if g != repr(float(v))]
    assert not bad, bad[:5]
    assert ex.format_doubles_device(eng, torch.tensor([float("nan")], dtype=torch.float64, device="cuda:0")) == ["nan"]


@pytest.mark.gpu
def test_interactive_queries():
    from test_gpu_dropin import objects_from_case
    from vexautonomousplanner_b200 import queries
    from vexautonomousplanner_b200.splines.spline_manager import QuinticHermiteSplineManager
    g = load_case("cfg1_factory")
    nodes, aps = objects_from_case(g)
    sm = QuinticHermiteSplineManager()
    assert sm.build_path(g["points_ft"], nodes, aps)
    n = g["n"]
    poly = queries.preview_polyline(sm, n)
    assert poly.shape == (25 * n, 2)
    want = np.array([sm.get_point_at_parameter(t) for t in np.linspace(0, n - 1, 25 * n)])
    assert np.array_equal(poly, (want / 12.1090395251 + 0.5) * 2000)
    np.testing.assert_allclose(poly[0], [300, 300], atol=1e-9)
    # closest point: the batched search must reproduce the reference's scalar two-pass search (gui/path.py:658-727)
    point_px = np.array([905.0, 1010.0])
    got_px, got_t = queries.find_closest_point_on_path(sm, point_px, n)
    point = (point_px / 2000 - 0.5) * 12.1090395251
    best, best_pc, best_t, best_pt = float("inf"), 0.0, 0.0, None
    for i in range(25 * n + 1):
        pc = i / (25 * n); t = sm.percent_to_parameter(pc); pt = sm.get_point_at_parameter(t)
        d = math.hypot(pt[0] - point[0], pt[1] - point[1])
        if d < best:
            best, best_pc, best_t, best_pt = d, pc, t, pt
    s0, e0 = max(0.0, best_pc - 0.02), min(1.0, best_pc + 0.02)
    for i in range(501):
        pc = s0 + i * ((e0 - s0) / 500); t = sm.percent_to_parameter(pc); pt = sm.get_point_at_parameter(t)
        d = math.hypot(pt[0] - point[0], pt[1] - point[1])
        if d < best:
            best, best_t, best_pt = d, t, pt
    assert got_t == best_t
    assert np.array_equal(got_px, (best_pt / 12.1090395251 + 0.5) * 2000)


def _gui_ref():
    import os
    from golden_util import GOLDEN_DIR
    return json.load(open(os.path.join(GOLDEN_DIR, "gui_reference.json")))


def _route_objects(state):
    """RouteNode / RouteActionPoint lists from a fixture snapshot (attribute-for-attribute what the GUI objects held)."""
    from vexautonomousplanner_b200 import export as ex
    nodes = []
    for d in state["nodes"]:
        n = ex.RouteNode()
        n.px = d["px"]; n.is_start_node, n.is_end_node = d["start"], d["end"]
        n.is_reverse_node, n.stop, n.turn, n.wait_time = d["reverse"], d["stop"], d["turn"], d["wait"]
        n.tangent = None if d["tangent"] is None else np.array(d["tangent"])
        n.incoming_magnitude, n.outgoing_magnitude, n.action_values = d["in_mag"], d["out_mag"], d["actions"]
        nodes.append(n)
    aps = []
    for d in state["action_points"]:
        a = ex.RouteActionPoint(d["t"]); a.px = d["px"]; a.stop, a.wait_time, a.action_values = d["stop"], d["wait"], d["actions"]
        aps.append(a)
    return nodes, aps


def test_json_codec_against_the_reference_gui_cpu():
    """f2 pinned to the reference: tests/golden/make_golden_gui.py ran the unmodified PathWidget.load_nodes /
    convert_point (gui/path.py:590-644) and AutonomousPlannerGUIManager.convert_nodes (gui_manager.py:388-427) headless.
    nodes_to_json must emit the reference's string byte for byte; load_nodes must land every node on the same pixel
    doubles, attributes, list order and build_path point order; a second convert must reproduce the reference's too."""
    from vexautonomousplanner_b200 import export as ex
    ref = _gui_ref()
    for name, r in ref.items():
        if "built" not in r:
            continue
        for stage in ("built", "loaded", "mirrored"):
            st = r[stage]
            nodes, aps = _route_objects(st)
            # convert_nodes ran BEFORE the path update moved the action points: the fixture holds their pixels at that moment
            s = ex.nodes_to_json([n.px for n in nodes], nodes, aps, st["ap_px_at_convert"])
            assert s == st["json"], (name, stage)                      # the reference's string, byte for byte
            assert ex.nodes_to_json([n.px for n in nodes], nodes, aps, st["ap_px_at_convert"], as_list=True) == st["as_list"]
        pts, nodes, aps, apx = ex.load_nodes(r["built"]["json"])
        L = r["loaded"]
        assert [n.px for n in nodes] == [d["px"] for d in L["nodes"]], name               # bit-equal pixel doubles
        assert pts.tolist() == L["ordered_px"], name
        for n, d in zip(nodes, L["nodes"]):
            assert (n.is_start_node, n.is_end_node, n.is_reverse_node, n.stop) == (d["start"], d["end"], d["reverse"], d["stop"])
            assert (n.turn, n.wait_time, n.incoming_magnitude, n.outgoing_magnitude) == (d["turn"], d["wait"], d["in_mag"], d["out_mag"])
            assert (None if n.tangent is None else n.tangent.tolist()) == d["tangent"]
            assert list(n.action_values) == d["actions"]
        assert [(a.t, a.stop, a.wait_time, list(a.action_values)) for a in aps] == \
               [(d["t"], d["stop"], d["wait"], d["actions"]) for d in L["action_points"]]
    # legacy single-list file (gui/path.py:604-609)
    pts, nodes, aps, _ = ex.load_nodes(ref["legacy_single_list"]["json"])
    assert [n.px for n in nodes] == [d["px"] for d in ref["legacy_single_list"]["nodes"]] and aps == []


@pytest.mark.gpu
def test_interactive_queries_against_the_reference_gui():
    """f3 pinned to the reference: the 25*N preview polyline returned by PathWidget.update_spline (gui/path.py:356-390)
    and the two-pass find_closest_point_on_path (gui/path.py:658-727), both run headless on the unmodified reference by
    tests/golden/make_golden_gui.py, for four routes (plain, turn / reverse / wait with action points, user tangents,
    8 random nodes), each also after load_nodes and after mirror_nodes, five mouse positions each."""
    from vexautonomousplanner_b200 import export as ex, queries
    from vexautonomousplanner_b200.splines.spline_manager import QuinticHermiteSplineManager
    ref = _gui_ref()
    n_q = 0
    for name, r in ref.items():
        if "built" not in r:
            continue
        for stage in ("built", "loaded", "mirrored"):
            st = r[stage]
            nodes, aps = _route_objects(st)
            sm = QuinticHermiteSplineManager()
            assert sm.build_path(ex.px_points_to_ft(st["ordered_px"]), nodes, aps)
            n = len(nodes)
            poly = queries.preview_polyline(sm, n)
            want = np.array(st["polyline"])
            assert poly.shape == want.shape == (25 * n, 2)
            np.testing.assert_allclose(poly, want, rtol=1e-12, atol=1e-9)          # pixels; 1e-9 px = 6e-12 ft
            for c in st["closest"]:
                got_px, got_t = queries.find_closest_point_on_path(sm, np.array(c["query"]), n)
                assert got_t == c["parameter"], (name, stage, c["query"])          # same sample wins: an index-like result
                np.testing.assert_allclose(got_px, c["px"], rtol=1e-12, atol=1e-9)
                n_q += 1
            # the action points land where _execute_update_path puts them (path.py:416-420)
            if aps:
                at = np.array([sm.get_point_at_parameter(a.t) for a in aps])
                np.testing.assert_allclose((at / 12.1090395251 + 0.5) * 2000, st["ap_px_after_update"], rtol=1e-12, atol=1e-9)
    assert n_q == 75


@pytest.mark.gpu
def test_trajectory_txt_against_the_reference_save_path():
    """f1 pinned to the reference's own save path: tests/golden/make_golden_gui.py ran the unmodified
    save_nodes_to_file + fill_txt_file (gui_manager.py:220-316) headless on two routes with turns, waits (also on node 0),
    reversal, stop and action points.  The engine's .txt body must have the same lines in the same places, print the
    SAME TOKEN wherever the reference printed an integer (row markers, action values, and the int 0 the reference's
    lists hold on inserted rows: "0", never "0.0"), and agree on every float within the north-star tolerances."""
    import gzip
    import os
    from golden_util import GOLDEN_DIR
    from vexautonomousplanner_b200 import export as ex
    from vexautonomousplanner_b200.engine import Engine
    from vexautonomousplanner_b200.packing import pack_paths
    from vexautonomousplanner_b200.synth import FACTORY
    ref = _gui_ref()
    eng = Engine("cuda:0")
    for name in ("turn_reverse_wait", "node0_wait"):
        st = ref[name]["built"]
        nodes, aps = _route_objects(st)
        want = gzip.open(os.path.join(GOLDEN_DIR, f"save_txt_{name}.txt.gz"), "rt").read()
        assert want.count("\n") == st["save_txt_lines"]
        packed = pack_paths([(ex.px_points_to_ft(st["ordered_px"]), nodes, aps)], FACTORY)
        db = eng.upload(packed)
        res = eng.profile(db)
        assert int(res.status[0]) == 0
        got = ex.trajectory_text(eng, res, 0, [n.action_values for n in nodes], [a.action_values for a in aps], db=db)
        gl, wl = got.splitlines(), want.splitlines()
        assert len(gl) == len(wl), name
        n_int_tokens = 0
        for r, (g, w) in enumerate(zip(gl, wl)):
            gt, wt = g.split(" "), w.split(" ")
            assert len(gt) == len(wt), (name, r, g, w)
            for c, (a, b) in enumerate(zip(gt, wt)):
                is_int = b.lstrip("-").isdigit()
                if is_int or b == "":
                    assert a == b, (name, r, c, g, w)              # ints print identically ("0", not "0.0")
                    n_int_tokens += is_int
                else:
                    assert not a.lstrip("-").isdigit(), (name, r, c, g, w)
                    tol = dict(rtol=1e-9, atol=1e-9) if c in (2, 3, 4) else dict(rtol=1e-6, atol=1e-6)
                    np.testing.assert_allclose(float(a), float(b), err_msg=f"{name} row {r} col {c}", **tol)
        assert n_int_tokens > len(wl)                               # column 0 plus the inserted rows' ints
    # the Python mirror returns the same element types as the reference (ints on inserted rows)
    from vexautonomousplanner_b200.motion_profiling_v2.motion_profile_generator import Constraints, generate_motion_profile
    from vexautonomousplanner_b200.splines.spline_manager import QuinticHermiteSplineManager
    st = ref["node0_wait"]["built"]
    nodes, aps = _route_objects(st)
    sm = QuinticHermiteSplineManager()
    assert sm.build_path(ex.px_points_to_ft(st["ordered_px"]), nodes, aps)
    out = generate_motion_profile(sm, Constraints(4.0, 8.0, 8.0, 0.8, 16.0, 12.5 / 12))
    times, positions, lin, acc, head, ang = out[:6]
    assert isinstance(times[0], np.float64) and times[0] == 0.0          # node 0 waits: current_time is a float by then
    assert lin[0] == 0 and isinstance(lin[0], int) and isinstance(ang[0], int) and isinstance(positions[0], int)
    assert isinstance(lin[40], np.float64) and isinstance(acc[0], int) and isinstance(head[0], np.float64)
