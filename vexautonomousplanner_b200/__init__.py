"""vexautonomousplanner_b200 -- B200-native batched spline -> motion-profile engine.

Drop-in for the hot path of RohitMovva/VexAutonomousPlanner (src/splines + src/motion_profiling_v2):
  - `Engine` / `pack_paths` / `pack_arrays`: the batched API over libvap.so (sm_100a CUDA, include/vap.h)
  - `splines`, `motion_profiling_v2`: mirrors of the reference's Python call surface (B = 1 wrappers)
  - `sharding`: one-process-per-GPU batch sharding + NCCL gather of per-path summaries
There is no CPU fallback; importing works without a GPU, using the engine does not.
"""
from ._lib import VapError, build, lib  # noqa: F401
from .packing import PackedPaths, pack_arrays, pack_paths, px_to_ft  # noqa: F401

__all__ = ["VapError", "build", "lib", "PackedPaths", "pack_arrays", "pack_paths", "px_to_ft", "Engine"]


def __getattr__(name):
    if name in ("Engine", "DeviceBatch", "ProfileResult"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
