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
    wait0 = eng.upload(packed).node_attr[:, 0, 3]
    dev_text = ex.export_text_device(eng, res, wait0)
    for b in (0, 17, 39):
        p = res.path(b)
        txt = ex.trajectory_text(eng, res, b, [[i] for i in range(8)], [[9], [9]], device_text=dev_text)
        traj = [[0, (0 if (k == 0 and int(packed.node_attr[b, 0, 3] / 0.01) == 0) else np.float64(p["times"][k])),
                 np.float64(p["x"][k] * 12), np.float64(p["y"][k] * -12), np.float64(p["headings"][k]),
                 np.float64(p["linear_vels"][k] * 12), np.float64(p["angular_vels"][k])] for k in range(len(p["times"]))]
        want = ex.format_rows(ex.splice_action_rows(traj, p["nodes_map"], [[i] for i in range(8)], p["actions_map"],
                                                    [[9], [9]]))
        assert txt == want, b
        assert txt.splitlines()[0] == "1 0 "


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
