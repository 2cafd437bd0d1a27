"""Mirror of the reference package src/motion_profiling_v2 (same module and function names)."""
from . import motion_profile_generator, one_dim_mp_generator  # noqa: F401
