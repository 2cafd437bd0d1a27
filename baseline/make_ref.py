#!/usr/bin/env python
"""Stage the UNMODIFIED reference hot path beside the repo so that it can be timed on the GPU box's host cores.

    python baseline/make_ref.py          # build container only: needs /root/reference

Copies /root/reference/src/{splines,motion_profiling_v2} (the two packages on the hot path, pure Python) to
baseline/_ref/src/.  baseline/_ref/ is git-ignored (reference sources never enter this repository's history) but not
gpurun-ignored, so it travels to the GPU box, where /root/reference does not exist.  The reference is not an installable
package (no setup.py / pyproject; `pip install /root/reference` has nothing to build), hence the plain copy.
bench.py times it as cpu_baseline.reference_python: build_path + generate_motion_profile (incl. rebuild_tables,
motion_profile_generator.py:402) per path, multiprocessing over all host cores.
"""
import os
import shutil
import sys

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "src")


def main() -> int:
    if not os.path.isdir(REF):
        print("no /root/reference here: nothing to stage (the GPU box uses the copy that travelled with the snapshot)")
        return 0
    for pkg in ("splines", "motion_profiling_v2"):
        dst = os.path.join(DST, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, pkg), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    print("staged", DST, sorted(os.listdir(DST)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
