#!/usr/bin/env python
"""Golden vectors for the "next" rows f2 / f3 / f4, produced by CALLING the unmodified
reference GUI code headless.  Build container only (needs /root/reference):

    python tests/golden/make_golden_gui.py        # writes tests/golden/gui_reference.json

What is driven, and how:
  * PyQt6 / matplotlib / qdarktheme are absent here, so a meta-path finder fabricates stand-in modules whose
    every attribute is an inert class; only QPointF (x / y / setX / setY) and the position bookkeeping of
    QGraphicsItem (setPos / pos / x / y) carry behaviour, because those are the only Qt facilities the methods
    below really use.  Nothing of the reference is copied: its own classes and methods run.
  * gui.path.PathWidget (real __init__), .add_node / .add_action_point / .load_nodes / .convert_point /
    .mirror_nodes / ._execute_update_path / .update_spline / .find_closest_point_on_path (path.py:356-425,
    439-478, 590-644, 658-727) and gui.gui_manager.AutonomousPlannerGUIManager.convert_nodes
    (gui_manager.py:388-427), called unbound on a stand-in main window whose central_widget is the PathWidget.

Per route the fixture stores: the node JSON string convert_nodes emits, what load_nodes makes of it (pixel
positions, attributes, the point order _execute_update_path hands to build_path), the 25*N preview polyline
update_spline returns, closest-point answers for a set of mouse positions, and the same again after
mirror_nodes.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import json
import logging
import os
import sys
import types

import numpy as np

REF_SRC = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gui_reference.json")


# ---------------------------------------------------------------------------------------------- inert Qt stand-ins
class _Inert:
    """Instantiable, callable, iterable-empty, attribute-complete nothing."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()

    def __bool__(self):
        return False

    def __iter__(self):
        return iter(())


class _InertMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert()


class QPointF:
    def __init__(self, x=0.0, y=0.0):
        self._x, self._y = float(x), float(y)          # Qt's qreal: coordinates are always doubles

    def x(self):
        return self._x

    def y(self):
        return self._y

    def setX(self, v):
        self._x = float(v)

    def setY(self, v):
        self._y = float(v)


class QGraphicsItem(metaclass=_InertMeta):
    """Position bookkeeping only (setPos accepts a QPointF or two numbers, like Qt)."""

    def __init__(self, *a, **k):
        self._pos = QPointF(0.0, 0.0)

    def setPos(self, *a):
        self._pos = QPointF(a[0].x(), a[0].y()) if len(a) == 1 else QPointF(a[0], a[1])

    def pos(self):
        return QPointF(self._pos.x(), self._pos.y())

    def x(self):
        return self._pos.x()

    def y(self):
        return self._pos.y()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert()


def _inert_class(name):
    return _InertMeta(name, (_Inert,), {})


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        if name == "QPointF":
            return QPointF
        if name == "QGraphicsItem":
            return QGraphicsItem
        cls = _inert_class(name)
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("PyQt6", "matplotlib", "qdarktheme")

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _StubFinder())
sys.path.insert(0, REF_SRC)
logging.disable(logging.CRITICAL)

from gui import path as ref_path  # noqa: E402
from gui import gui_manager as ref_gm  # noqa: E402


class _Config:
    """utilities.config_manager.ConfigManager stand-in: the three values the driven code reads."""

    def __init__(self, actions):
        self.actions = actions

    def get_value(self, section, key):
        return {"actions": self.actions, "width": 12.5, "length": 12.5}.get(key, 0)


def new_widget(n_actions=2):
    parent = _Inert()
    w = ref_path.PathWidget(_Config([f"a{i}" for i in range(n_actions)]), parent=parent, image_path=None)
    return w


def node_state(w):
    return [dict(px=[n.x(), n.y()], start=bool(n.is_start_node), end=bool(n.is_end_node),
                 reverse=bool(n.is_reverse_node), stop=bool(n.stop), turn=n.turn, wait=n.wait_time,
                 tangent=None if n.tangent is None else [float(n.tangent[0]), float(n.tangent[1])],
                 in_mag=n.incoming_magnitude, out_mag=n.outgoing_magnitude, actions=list(n.action_values))
            for n in w.nodes]


def ap_state(w):
    return [dict(px=[a.x(), a.y()], t=a.t, stop=a.stop, wait=a.wait_time, actions=list(a.action_values))
            for a in w.action_points]


def ordered_points(w):
    """The point list _execute_update_path builds (path.py:407-413): start node, the others, end node."""
    pts = [w.start_node.pos()]
    for nd in w.nodes:
        if nd.is_end_node or nd.is_start_node:
            continue
        pts.append(nd.pos())
    pts.append(w.end_node.pos())
    return pts


def snapshot(w, queries):
    gm = types.SimpleNamespace(central_widget=w)
    ap_px_at_convert = [[a.x(), a.y()] for a in w.action_points]     # convert_nodes reads the CURRENT pixel positions
    js = ref_gm.AutonomousPlannerGUIManager.convert_nodes(gm)
    as_list = ref_gm.AutonomousPlannerGUIManager.convert_nodes(gm, as_list=True)
    w._execute_update_path()                      # builds the spline manager + QPainterPath stand-in
    pts = ordered_points(w)
    poly = w.update_spline(pts, w.nodes, w.action_points)
    closest = []
    for q in queries:
        pt, par = w.find_closest_point_on_path(types.SimpleNamespace(isEmpty=lambda: False, length=lambda: 1.0),
                                               QPointF(q[0], q[1]))
        closest.append(dict(query=list(q), px=[float(pt.x()), float(pt.y())], parameter=float(par)))
    return dict(json=js, as_list=as_list, nodes=node_state(w), action_points=ap_state(w),
                ordered_px=[[p.x(), p.y()] for p in pts], polyline=np.asarray(poly).tolist(), closest=closest,
                ap_px_at_convert=ap_px_at_convert, ap_px_after_update=[[a.x(), a.y()] for a in w.action_points])


def build_route(spec):
    """Create the route through the GUI's own entry points (add_node, attribute assignment as the context menus do)."""
    w = new_widget(spec.get("n_actions", 2))
    for i, nd in enumerate(spec["nodes"]):
        n = w.add_node(QPointF(nd["px"][0], nd["px"][1]))
        for k, v in nd.items():
            if k == "px":
                continue
            if k == "tangent":
                n.set_tangent(np.array(v))
            elif k == "actions":
                n.set_action_values(list(v))
            else:
                setattr(n, k, v)
        if i == 0:
            w.start_node = n; n.is_start_node = True
        if i == len(spec["nodes"]) - 1:
            w.end_node = n; n.is_end_node = True
    for ap in spec.get("action_points", []):
        a = w.add_action_point(QPointF(0.0, 0.0), ap["t"])
        a.stop = ap.get("stop", False); a.wait_time = ap.get("wait", 0)
        if "actions" in ap:
            a.set_action_values(list(ap["actions"]))
    return w


ROUTES = {
    "cfg1": dict(nodes=[dict(px=[300, 300]), dict(px=[700, 500]), dict(px=[1000, 1200]), dict(px=[1400, 900]),
                        dict(px=[1700, 1500]), dict(px=[1200, 1700])]),
    "turn_reverse_wait": dict(nodes=[dict(px=[250.5, 410.25]), dict(px=[640.0, 880.0], turn=45, wait_time=0.25),
                                     dict(px=[1210.75, 640.5], is_reverse_node=True, stop=True, actions=[3, 0]),
                                     dict(px=[1500.0, 1333.0], turn=-90), dict(px=[900.0, 1700.0])],
                              action_points=[dict(t=0.7, stop=True, wait=0.1, actions=[1, 0]), dict(t=2.4)]),
    "tangents": dict(nodes=[dict(px=[400, 1500]), dict(px=[800, 1100], tangent=[0.6, 0.8], incoming_magnitude=1.5,
                                                       outgoing_magnitude=2.0),
                            dict(px=[1300, 1250], tangent=[1.0, -0.25], incoming_magnitude=0.75, outgoing_magnitude=1.25),
                            dict(px=[1650, 600])], n_actions=0),
    "rand8": None,       # filled below
    "node0_wait": dict(nodes=[dict(px=[300, 300], wait_time=0.37), dict(px=[700, 500]), dict(px=[1000, 1200], wait_time=0.2),
                              dict(px=[1400, 900], turn=30), dict(px=[1700, 1500]), dict(px=[1200, 1700])],
                       action_points=[dict(t=1.5, wait=0.05)]),
}
SAVE_TXT = ("turn_reverse_wait", "node0_wait")     # routes whose trajectory .txt is written by the reference's own save path
rng = np.random.default_rng(5)
ROUTES["rand8"] = dict(nodes=[dict(px=[float(x), float(y)]) for x, y in rng.uniform(100, 1900, (8, 2))])
QUERIES = [[905.0, 1010.0], [310.0, 290.0], [1999.0, 1.0], [1205.5, 1690.25], [20.0, 1800.0]]


def save_txt(w):
    """AutonomousPlannerGUIManager.save_nodes_to_file (gui_manager.py:232-316) + fill_txt_file (:220-230), unbound, on a
    stand-in main window: returns the trajectory .txt body the reference writes (factory constraints, src/config.yaml)."""
    import tempfile
    cls = ref_gm.AutonomousPlannerGUIManager
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "hdr")); os.makedirs(os.path.join(d, "routes"))
        gm = types.SimpleNamespace(central_widget=w, current_working_file="route", routes_header_path=os.path.join(d, "hdr"),
                                   routes_folder_path=os.path.join(d, "routes"), max_velocity=4.0, max_acceleration=8.0,
                                   max_jerk=16.0, track_width=12.5 / 12)
        gm.convert_nodes = lambda as_list=False: cls.convert_nodes(gm, as_list)
        gm.fill_txt_file = lambda nd: cls.fill_txt_file(gm, nd)
        cls.save_nodes_to_file(gm)
        with open(os.path.join(d, "hdr", "route.txt")) as f:
            return f.read()


def main():
    import gzip
    out = {}
    for name, spec in ROUTES.items():
        w = build_route(spec)
        first = snapshot(w, QUERIES)
        if name in SAVE_TXT:
            txt = save_txt(w)
            with gzip.GzipFile(os.path.join(os.path.dirname(OUT), f"save_txt_{name}.txt.gz"), "wb", mtime=0) as f:
                f.write(txt.encode())
            first["save_txt_lines"] = txt.count("\n")
        # f2: what load_nodes makes of the string convert_nodes wrote (fresh widget, the reference's own parser)
        w2 = new_widget(spec.get("n_actions", 2))
        w2.load_nodes(first["json"])
        loaded = snapshot(w2, QUERIES)
        # f4: mirror_nodes on the loaded route
        w2.mirror_nodes()
        mirrored = snapshot(w2, QUERIES)
        out[name] = dict(built=first, loaded=loaded, mirrored=mirrored)
    # legacy single-list files (path.py:604-609)
    w3 = new_widget(2)
    w3.load_nodes(json.dumps(out["cfg1"]["built"]["as_list"][0]))
    out["legacy_single_list"] = dict(json=json.dumps(out["cfg1"]["built"]["as_list"][0]), nodes=node_state(w3))
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, {k: (len(v["built"]["polyline"]) if "built" in v else len(v["nodes"])) for k, v in out.items()})


if __name__ == "__main__":
    main()
