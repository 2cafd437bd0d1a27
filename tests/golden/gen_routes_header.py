"""Golden vectors for the legacy routes.h writer: runs the reference's OWN fill_template (gui_manager.py:442-507), lifted
out of its Qt class with ast (PyQt6 is not importable here), on a scratch directory.  Needs /root/reference; run in the
build container:  python tests/golden/gen_routes_header.py  ->  tests/golden/routes_header.json"""
import ast
import json
import logging
import os
import tempfile
import types

REF = "/root/reference/src/gui/gui_manager.py"
ROWS_A = [[0, 0.0, 12.5, -3.25, 0.1, 1e-05, 0.0], [1, 0, 90, 0.5], [0, 0.01, 12.625, -3.5, 0.125, 24.0, -1.5e+16],
          [7, 8]]
ROWS_B = [[0, 0.0, 1.0, 2.0, 3.0, 4.0, 5.0], [1, 1, 0, 0], [0, 0.025, 1.5, 2.5, 3.5, 4.5, 0.1 + 0.2]]


def reference_fill_template():
    tree = ast.parse(open(REF).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "fill_template":
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"os": os, "logger": logging.getLogger("ref")}
            exec(compile(mod, REF, "exec"), ns)
            return ns["fill_template"]
    raise RuntimeError("fill_template not found")


def main():
    fill = reference_fill_template()
    cases = []
    with tempfile.TemporaryDirectory() as d:
        def run(tag, before, name, rows):
            path = os.path.join(d, tag + ".h")
            if before is not None:
                open(path, "w").write(before)
            fill(types.SimpleNamespace(current_working_file=name, routes_header_path=path), rows)
            cases.append(dict(tag=tag, before=before, name=name, rows=rows, after=open(path).read()))
            return cases[-1]["after"]

        first = run("missing", None, "sawp", ROWS_A)
        second = run("append", first, "elims", ROWS_B)
        run("replace", second, "sawp", ROWS_B)
        run("empty", "", "skills", ROWS_A)
        run("no_endif", "#include <vector>\n", "sawp", ROWS_B)
        run("indented_decl", "#ifndef ROUTES_H\n  std::vector<std::vector<double>> sawp = {{1, 2}};\n#endif\n", "sawp", ROWS_A)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "routes_header.json")
    json.dump(cases, open(out, "w"), indent=1)
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
