"""ctypes binding of libcc4_b200.so (the C-ABI declared in include/cc4_b200.h)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CC4_B200_LIB selects another build of the same C-ABI (tuning variants built by tools/build_variants.py)
LIB_PATH = os.environ.get("CC4_B200_LIB") or os.path.join(HERE, "libcc4_b200.so")

EXPORTS = [
    "cc4_create", "cc4_destroy", "cc4_set_episode_base", "cc4_set_red_agent", "cc4_state_bytes", "cc4_set_tape", "cc4_reset", "cc4_reset_host", "cc4_step", "cc4_step_host",
    "cc4_rollout", "cc4_expand_obs_bits", "cc4_host_write_bandwidth", "cc4_error_flags", "cc4_export_canon", "cc4_launch_count", "cc4_phase_clocks", "cc4_action_clocks", "cc4_last_error",
]

_lib = None


class NativeError(RuntimeError):
    pass


def load():
    """Load the extension; fails loudly when it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc -gencode arch=compute_100a,code=sm_100a); this engine has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.cc4_create.argtypes = [i32, i32, u64, u64, i32, i32, C.POINTER(vp)]
    L.cc4_destroy.argtypes = [vp]
    L.cc4_set_episode_base.argtypes = [vp, u64, u64]
    L.cc4_set_red_agent.argtypes = [vp, i32]
    L.cc4_state_bytes.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.cc4_set_tape.argtypes = [vp, vp, vp, C.c_uint32]
    L.cc4_reset.argtypes = [vp, vp, vp, vp]
    L.cc4_reset_host.argtypes = [vp, vp, vp, vp]
    L.cc4_step.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp]
    L.cc4_step_host.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp]
    L.cc4_rollout.argtypes = [vp, i32, u64, vp, i32, vp]
    L.cc4_error_flags.argtypes = [vp, vp]
    L.cc4_export_canon.argtypes = [vp, i32, vp, i32, C.POINTER(i32)]
    L.cc4_launch_count.argtypes = [vp, C.POINTER(u64)]
    L.cc4_expand_obs_bits.argtypes = [vp, vp, C.c_size_t, i32]
    L.cc4_host_write_bandwidth.argtypes = [C.c_size_t, i32, i32, i32, C.POINTER(C.c_double)]
    L.cc4_phase_clocks.argtypes = [vp, vp, i32]
    L.cc4_action_clocks.argtypes = [vp, vp]
    L.cc4_last_error.restype = C.c_char_p
    for name in EXPORTS:
        if name != "cc4_last_error":
            getattr(L, name).restype = i32
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        raise NativeError(f"{what} failed ({rc}): {load().cc4_last_error().decode()}")
