"""CPU-side tests (no GPU): the C ABI exports, host packing, synthetic workloads, sharding + gloo gather."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    """libvap.so must load without a GPU and export exactly the functions include/vap.h declares."""
    import vexautonomousplanner_b200 as vap
    from vexautonomousplanner_b200 import _lib
    so = vap.build()
    header = open(os.path.join(ROOT, "include", "vap.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(vap_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    lib = ctypes.CDLL(so)
    for name in declared:
        assert hasattr(lib, name), name
    L = vap.lib()
    assert L.vap_version() == 100
    assert L.vap_last_error() is not None
    nm = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (vap_\w+)", nm))
    assert declared <= exported


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vexautonomousplanner_b200 import VapError
    from vexautonomousplanner_b200.engine import Engine
    with pytest.raises(VapError):
        Engine("cuda:0")


def test_sass_is_sm100a_without_fma_contraction():
    """The shared library carries sm_100a code, and the index-critical kernels contain explicit DFMA only where
    fma() is written (build_path's norm / rotation), i.e. -fmad=false is in effect."""
    import vexautonomousplanner_b200 as vap
    so = vap.build()
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    full = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    lines, keep = [], False
    for ln in full.splitlines():
        if "Function :" in ln:
            keep = "k_build_lut" in ln
        if keep:
            lines.append(ln)
    sass = "\n".join(lines)
    assert "DMUL" in sass and "DADD" in sass
    # the only DFMA in the LUT kernel belong to the inlined sqrt / division sequences, never to a*b+c of the basis sums:
    # there must be at least as many DMUL+DADD as basis terms (6 polynomials x 3 terms + 12 accumulations)
    assert sass.count("DMUL") >= 20 and sass.count("DADD") >= 20


class _Node:
    def __init__(self, **kw):
        self.is_reverse_node = False; self.turn = 0; self.wait_time = 0; self.stop = False; self.tangent = None
        self.incoming_magnitude = None; self.outgoing_magnitude = None; self.max_velocity = 0; self.max_acceleration = 0
        self.__dict__.update(kw)


class _AP:
    def __init__(self, t, **kw):
        self.t = t; self.stop = False; self.wait_time = 0; self.max_velocity = 0; self.max_acceleration = 0
        self.__dict__.update(kw)


def test_pack_paths_layout_and_rotation_table():
    from vexautonomousplanner_b200 import pack_arrays, pack_paths, px_to_ft
    from vexautonomousplanner_b200.packing import F_REVERSE, F_STOP, F_TANGENT
    pts = px_to_ft([[300, 300], [700, 500], [1000, 1200], [1400, 900]])
    assert pts[0, 0] == (300 / 2000 - 0.5) * 12.1090395251
    nodes = [_Node(), _Node(turn=45, is_reverse_node=True, stop=True), _Node(tangent=np.array([0.6, 0.8]),
             incoming_magnitude=1.5, outgoing_magnitude=2.0, wait_time=0.2, max_velocity=2.5), _Node(max_acceleration=5.0)]
    aps = [_AP(1.5, stop=True, wait_time=0.1, max_velocity=2.0)]
    p = pack_paths([(pts, nodes, aps), (pts[:3], nodes[:3], [])], [4.0, 8.0, 8.0, 0.8, 16.0, 1.0])
    assert p.node_attr.shape == (2, 4, 12) and p.n_nodes.tolist() == [4, 3] and p.n_ap.tolist() == [1, 0]
    assert p.node_flags[0].tolist() == [0, F_REVERSE | F_STOP, F_TANGENT, 0]
    ang = np.radians(45) + np.pi
    assert p.node_attr[0, 1, 10] == np.cos(ang) and p.node_attr[0, 1, 11] == np.sin(ang)
    assert p.node_attr[0, 0, 10] == 1.0 and p.node_attr[0, 0, 11] == 0.0
    assert p.node_attr[0, 2, 6:10].tolist() == [0.6, 0.8, 1.5, 2.0]
    assert p.ap_attr[0, 0].tolist() == [1.5, 0.1, 2.0, 0.0] and p.ap_flags[0, 0] == F_STOP
    assert p.max_splines() == 2
    q = pack_arrays(pts[None], [4.0, 8.0, 8.0, 0.8, 16.0, 1.0], reverse=[[0, 1, 0, 0]], stop=[[0, 1, 0, 0]],
                    turn=[[0, 45, 0, 0]], wait=[[0, 0, 0.2, 0]], max_velocity=[[0, 0, 2.5, 0]],
                    max_acceleration=[[0, 0, 0, 5.0]], tangent=[[[np.nan, np.nan]] * 2 + [[0.6, 0.8]] + [[np.nan, np.nan]]],
                    in_mag=[[0, 0, 1.5, 0]], out_mag=[[0, 0, 2.0, 0]], ap_t=[[1.5]], ap_stop=[[True]], ap_wait=[[0.1]],
                    ap_max_velocity=[[2.0]])
    assert np.array_equal(q.node_attr[0], p.node_attr[0]) and np.array_equal(q.node_flags[0], p.node_flags[0])
    assert np.array_equal(q.ap_attr[0], p.ap_attr[0]) and np.array_equal(q.ap_flags[0], p.ap_flags[0])
    with pytest.raises(ValueError):
        q.mirrored()                  # packed from feet: the reference mirrors in pixel space (gui/path.py:596-600)


def test_mirrored_is_the_reference_transform():
    """f4: PackedPaths.mirrored() against the reference's own mirror_nodes (gui/path.py:596-600).  (a) the golden pairs
    mixed8_k / mixed8_k_mirror (inputs of the latter were built as 2000 - x_px, -turn and run through the unmodified
    reference): mirroring the packed first case must give the packed twin BIT FOR BIT, tangents untouched;
    (b) gui_reference.json: node pixels / turns after the reference's PathWidget.mirror_nodes ran headless."""
    import json
    import os
    from golden_util import GOLDEN_DIR, load_case
    from vexautonomousplanner_b200.packing import mirror_nodes_px, pack_arrays

    def pack(c):
        tg = np.where(c["n_has_tangent"][:, None], c["n_tangent"], np.nan)
        return pack_arrays(None, c["constraints"], reverse=c["n_reverse"][None], stop=c["n_stop"][None],
                           turn=c["n_turn"][None], wait=c["n_wait"][None], max_velocity=c["n_maxvel"][None],
                           max_acceleration=c["n_maxacc"][None], tangent=tg[None], in_mag=c["n_inmag"][None],
                           out_mag=c["n_outmag"][None], ap_t=c["ap_t"][None] if len(c["ap_t"]) else None,
                           ap_stop=c["ap_stop"][None] if len(c["ap_t"]) else None,
                           ap_wait=c["ap_wait"][None] if len(c["ap_t"]) else None,
                           ap_max_velocity=c["ap_maxvel"][None] if len(c["ap_t"]) else None,
                           ap_max_acceleration=c["ap_maxacc"][None] if len(c["ap_t"]) else None,
                           points_px=c["points_px"][None])

    for k in (0, 1):
        a, b = load_case(f"mixed8_{k}"), load_case(f"mixed8_{k}_mirror")
        pa = pack(a)
        assert np.array_equal(pa.node_attr[0], a["node_attr"]) and np.array_equal(pa.node_flags[0], a["node_flags"])
        m = pa.mirrored()
        assert np.array_equal(m.points_px[0], b["points_px"])
        assert m.node_attr[0].tobytes() == b["node_attr"].tobytes()          # feet, turn, rotation table, tangents: all bits
        assert np.array_equal(m.node_flags[0], b["node_flags"]) and np.array_equal(m.ap_attr[0], b["ap_attr"])
        np.testing.assert_allclose(m.mirrored().points_px, pa.points_px, rtol=0, atol=1e-9)   # an involution up to rounding
    # a node with a user tangent keeps it (the reference's mirror does not touch tangents)
    t = load_case("tangents")
    pt = pack(t)
    assert pt.node_flags[0].tolist() == t["node_flags"].tolist() and (pt.node_flags[0] & 4).any()
    mt = pt.mirrored()
    assert np.array_equal(mt.node_attr[0, :, 6:10], pt.node_attr[0, :, 6:10])
    # the reference's own mirror_nodes, run headless (tests/golden/make_golden_gui.py)
    ref = json.load(open(os.path.join(GOLDEN_DIR, "gui_reference.json")))
    for name, r in ref.items():
        if "mirrored" not in r:
            continue
        px0 = np.array([n["px"] for n in r["loaded"]["nodes"]]); tu0 = np.array([n["turn"] for n in r["loaded"]["nodes"]], dtype=float)
        px1, tu1 = mirror_nodes_px(px0, tu0)
        assert px1.tolist() == [n["px"] for n in r["mirrored"]["nodes"]], name
        assert tu1.tolist() == [float(n["turn"]) for n in r["mirrored"]["nodes"]], name
        assert [n["tangent"] for n in r["mirrored"]["nodes"]] == [n["tangent"] for n in r["loaded"]["nodes"]]


def test_synthetic_workloads_follow_the_spec():
    from vexautonomousplanner_b200 import synth
    p = synth.random_paths(64, 8, seed=0)
    px = (p.node_attr[:, :, 0:2] / 12.1090395251 + 0.5) * 2000
    assert px.min() >= 15 - 1e-9 and px.max() <= 1985 + 1e-9
    assert np.linalg.norm(np.diff(px, axis=1), axis=2).min() >= 30 - 1e-9
    assert np.array_equal(p.cons[0], np.array(synth.FACTORY))
    m = synth.mixed_paths(64, 8, seed=3)
    assert (m.node_attr[:, 0, 2] == 0).all() and (m.node_attr[:, -1, 2] == 0).all()       # no turn at first / last node
    assert ((m.node_flags[:, -1] & 1) == 0).all()                                            # no reverse at the last node
    # second half = mirror in PIXEL space (x_px -> 2000 - x_px, gui/path.py:596-600), so feet agree to rounding only
    assert np.allclose(m.node_attr[32:, :, 0], -m.node_attr[:32, :, 0], rtol=0, atol=1e-12)
    assert np.array_equal(m.node_attr[32:, :, 2], -m.node_attr[:32, :, 2])
    assert (m.node_attr[:, :, 2] != 0).any() and (m.node_attr[:, :, 3] > 0).any() and (m.n_ap > 0).any()
    assert synth.cfg1().node_attr.shape == (1, 6, 12)


def test_shard_bounds_cover_and_balance():
    from vexautonomousplanner_b200 import synth
    from vexautonomousplanner_b200.sharding import chord_weights, local_shard, shard_bounds
    for B, W in ((10, 3), (4096, 8), (5, 8), (1, 2)):
        b = shard_bounds(B, W)
        assert b[0][0] == 0 and b[-1][1] == B and all(b[i][1] == b[i + 1][0] for i in range(W - 1))
    p = synth.random_paths(512, 8, seed=4)
    w = chord_weights(p)
    b = shard_bounds(512, 4, w)
    assert b[0][0] == 0 and b[-1][1] == 512
    loads = [w[lo:hi].sum() for lo, hi in b]
    assert max(loads) / min(loads) < 1.05
    sub, (lo, hi) = local_shard(p, 1, 4)
    assert sub.B == hi - lo and np.array_equal(sub.node_attr, p.node_attr[lo:hi])
    # plain paths: the weight is the chord sum in samples; mixed paths (cfg5) add the rows of waits and turn profiles
    d = np.diff(p.node_attr[:, :, 0:2], axis=1)
    np.testing.assert_allclose(w, np.hypot(d[:, :, 0], d[:, :, 1]).sum(axis=1) / 0.005, rtol=1e-12)
    m = synth.mixed_paths(256, 8, seed=3)
    wm = chord_weights(m)
    dm = np.diff(m.node_attr[:, :, 0:2], axis=1)
    base = np.hypot(dm[:, :, 0], dm[:, :, 1]).sum(axis=1) / 0.005
    has_insert = ((m.node_attr[:, :, 2] != 0) | (m.node_attr[:, :, 3] > 0)).any(axis=1) | (m.ap_attr[:, :, 1] > 0).any(axis=1)
    assert np.all(wm[~has_insert & (m.n_ap == 0)] == base[~has_insert & (m.n_ap == 0)])
    assert np.all(wm[(m.node_attr[:, :, 3] > 0).any(axis=1)] > base[(m.node_attr[:, :, 3] > 0).any(axis=1)] + 9)
    assert np.all(wm[(m.node_attr[:, :, 2] != 0).any(axis=1)] > base[(m.node_attr[:, :, 2] != 0).any(axis=1)] + 50)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from vexautonomousplanner_b200 import synth
from vexautonomousplanner_b200.sharding import SummaryGatherer, gather_summaries, local_shard, shard_bounds, chord_weights
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
packed = synth.random_paths(37, 6, seed=2)
sub, (lo, hi) = local_shard(packed, rank, world)
# stand-in for the per-path summary rows a rank would produce (n_out, L, t_end, max|v|, status)
rows = torch.tensor(np.stack([np.arange(lo, hi), sub.node_attr[:, 0, 0], sub.node_attr[:, 1, 1],
                              sub.node_attr[:, 2, 0], np.zeros(hi - lo)], axis=1))
allrows = gather_summaries(rows)
assert allrows.shape == (37, 5), allrows.shape
assert torch.equal(allrows[:, 0], torch.arange(37, dtype=torch.float64))
assert torch.equal(allrows[:, 1], torch.tensor(packed.node_attr[:, 0, 0]))
# the preallocated gatherer the bench uses (ragged shard sizes, several steps through the slot ring)
counts = [b - a for a, b in shard_bounds(packed.B, world, chord_weights(packed))]
sg = SummaryGatherer(counts, "cpu")
for step in range(6):
    slot = sg.submit(rows + step)
    sg.wait()
    got = sg.rows(slot)
    assert got.shape == (37, 5) and torch.equal(got[:, 0], torch.arange(37, dtype=torch.float64) + step), step
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_gloo_world2_shard_and_gather(tmp_path):
    """The N > 1 path on CPU: two processes shard one batch and all_gather the summary rows over gloo."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT, "port": port})
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o
