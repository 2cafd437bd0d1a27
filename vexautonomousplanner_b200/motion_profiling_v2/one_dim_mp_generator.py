"""Mirror of src/motion_profiling_v2/one_dim_mp_generator.py."""
import numpy as np
import torch

from ..runtime import C, _p, check, dev, get_engine, stream


def _turn_profile(x, max_velocity, max_acceleration, track_width, dt, mode):
    eng = get_engine()
    q = dev([[float(x), float(max_velocity), float(max_acceleration), float(track_width), float(dt)]])
    cap = 4096
    while True:
        a = eng._empty((1, cap)); b = eng._empty((1, cap)); cnt = eng._empty((1,), torch.int32)
        check(eng.lib.vap_turn_profile(C.c_int64(1), _p(q), C.c_int(mode), C.c_int64(cap), _p(a), _p(b), _p(cnt), stream()),
              "vap_turn_profile")
        K = int(cnt.item())
        if K <= cap:
            return a[0, :K].cpu().numpy(), b[0, :K].cpu().numpy()
        cap = K


def generate_trapezoidal_profile(max_velocity, max_acceleration, total_distance, time_step=0.01):
    """Trapezoidal / triangular velocity samples (one_dim_mp_generator.py:4-69); returns the velocity ndarray."""
    v, _ = _turn_profile(total_distance, max_velocity, max_acceleration, 1.0, time_step, mode=1)
    return v
