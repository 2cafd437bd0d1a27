"""Bounds diagnostics run (stand-in for compute-sanitizer memcheck, which is closed on this GPU pool).

Build the engine with -DVAP_BOUNDS_CHECK (every hand-computed index of the TMA rings, the chunk-interleaved pass arrays, the
sampling tiles and the time loop's ring is tested against the extent of its array before the access), run smoke()-like
batches that exercise every layout case, and print violations / checks executed per site.

    VAP_LIB_PATH=$PWD/vexautonomousplanner_b200/libvap_bounds.so VAP_NVCC_EXTRA=-DVAP_BOUNDS_CHECK \
        python -c "import vexautonomousplanner_b200 as v; v.build(force=True)"
    VAP_LIB_PATH=$PWD/vexautonomousplanner_b200/libvap_bounds.so python profiles/tools/bounds_check.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from vexautonomousplanner_b200 import _lib, synth  # noqa: E402
from vexautonomousplanner_b200.engine import Engine  # noqa: E402

SITES = {0: "ring: override-limit source alignment", 1: "ring: record source inside the path's rows, 16-byte aligned",
         2: "ring: shared-memory destination inside the CTA's ring", 3: "ring: forward-velocity source inside the path's row",
         4: "forward pass: extent of a chunk's loads / stores", 5: "backward pass: extent of a chunk's loads / stores",
         6: "forward sweep: shared-memory read inside the stage", 7: "backward sweep: shared-memory read inside the stage",
         8: "sampling tile: store index, sample index", 9: "sampling: tile read index, record offset",
         10: "time loop: TMA block inside the velocity row / ring slot",
         11: "time loop: the step's five shared-memory reads inside the thread's ring slice"}


def main():
    L = _lib.lib()
    buf = (C.c_uint64 * 32)()
    if L.vap_diag_read(buf, 1) != 0:
        print("this libvap.so has no bounds diagnostics:", L.vap_last_error().decode())
        return 2
    cases = [("cfg1 single path", synth.cfg1(), {}), ("8 x 6-node", synth.random_paths(8, 6, 7), {}),
             ("512 x 8-node (cfg2 slice)", synth.random_paths(512, 8, 0), {}),
             ("256 mixed (cfg5: turns, waits, reversals, overrides)", synth.mixed_paths(256, 8, 3), {}),
             ("64 x 16-node (cfg3 slice)", synth.random_paths(64, 16, 1), {}),
             ("3-node short paths, 8 chunks", synth.random_paths(32, 3, 5), dict(chunks=8)),
             ("64 chunks", synth.random_paths(64, 8, 2), dict(chunks=64)),
             ("one 121-node path, 256 chunks", synth.long_path(121, 2), dict(chunks=256))]
    bad = 0
    for name, packed, kw in cases:
        eng = Engine("cuda:0", **kw)
        res = eng.profile(eng.upload(packed))
        torch.cuda.synchronize()
        ok = bool((res.status == 0).all().item())
        assert L.vap_diag_read(buf, 1) == 0
        v = np.array(list(buf), dtype=np.uint64)
        viol, checks = v[:16], v[16:]
        bad += int(viol.sum())
        print(f"{name:55s} status ok={ok}  violations={int(viol.sum())}  checks executed={int(checks.sum())}")
        for s in range(12):
            if checks[s] or viol[s]:
                print(f"    site {s:2d} {SITES[s]:62s} checked {int(checks[s]):12d}  violations {int(viol[s])}")
    print("TOTAL violations:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
