"""ctypes binding of the C oracle (TEST INFRASTRUCTURE — only tests/, smoke() and bench.py's CPU legs import this)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libcc4_oracle.so")
N_ACTORS = 91
DRAW_DTYPE = np.dtype([("val", "<u8"), ("n", "<u2"), ("site", "u1"), ("kind", "u1"), ("pad", "<u4")])


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "cc4_oracle.c")
    if force or not os.path.exists(LIB_PATH) or (os.path.exists(src) and os.path.getmtime(LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.cc4o_create.restype = C.c_void_p
        L.cc4o_create.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int]
        L.cc4o_destroy.argtypes = [C.c_void_p]
        L.cc4o_set_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.cc4o_reset.argtypes = [C.c_void_p]
        L.cc4o_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cc4o_observe.argtypes = [C.c_void_p, C.c_void_p]
        L.cc4o_action_mask.argtypes = [C.c_void_p, C.c_void_p]
        L.cc4o_export.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.cc4o_error_flags.argtypes = [C.c_void_p]
        L.cc4o_last_error.restype = C.c_char_p
        L.cc4o_last_error.argtypes = [C.c_void_p]
        L.cc4o_reward_breakdown.argtypes = [C.c_void_p, C.c_void_p]
        L.cc4o_run_episodes.restype = C.c_long
        L.cc4o_run_episodes.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def tape_from_fixture(fx, n_ticks: int | None = None):
    """(records, seg_off) in the C layout from a golden fixture's t_* arrays."""
    t_step = np.asarray(fx["t_step"]).astype(np.int64)
    t_actor = np.asarray(fx["t_actor"]).astype(np.int64)
    n_ticks = int(len(fx["actions"])) if n_ticks is None else n_ticks
    seg = np.where(t_step < 0, 0, 1 + t_step * N_ACTORS + t_actor)
    assert not np.any((t_step >= 0) & (t_actor == 255)), "draw outside any agent context during a tick"
    order = np.argsort(seg, kind="stable")
    recs = np.zeros(len(order), dtype=DRAW_DTYPE)
    recs["val"] = np.asarray(fx["t_val"])[order]
    recs["n"] = np.asarray(fx["t_n"])[order]
    recs["site"] = np.asarray(fx["t_site"])[order]
    recs["kind"] = np.asarray(fx["t_kind"])[order]
    n_segs = 1 + n_ticks * N_ACTORS
    counts = np.bincount(seg, minlength=n_segs)[:n_segs]
    seg_off = np.zeros(n_segs + 1, dtype=np.uint32)
    seg_off[1:] = np.cumsum(counts)
    return recs, seg_off


class OracleEnv:
    def __init__(self, seed=0, episode=0, steps=500, tape=None, red="fsm"):
        self.L = lib()
        self.h = self.L.cc4o_create(seed, episode, steps, 1 if tape is not None else 0)
        if red != "fsm":        # DiscoveryFSRed (FSMRedVariants.py:72-122) / RandomSelectRedAgent (RandomSelectRedAgent.py:8-139)
            self.L.cc4o_set_red_agent.argtypes = [C.c_void_p, C.c_int]
            self.L.cc4o_set_red_agent(self.h, {"discovery": 1, "random_select": 2}[red])
        self._tape = tape
        if tape is not None:
            recs, seg_off = tape
            self.L.cc4o_set_tape(self.h, recs.ctypes.data, seg_off.ctypes.data, len(seg_off) - 1)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.cc4o_destroy(self.h)
            self.h = None

    def reset(self):
        rc = self.L.cc4o_reset(self.h)
        return rc

    def step(self, actions, msgs=None):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        m = np.ascontiguousarray(msgs if msgs is not None else np.zeros(5), dtype=np.uint8)
        obs = np.zeros((5, 210), dtype=np.uint8)
        rew = np.zeros(1, dtype=np.float32)
        done = np.zeros(1, dtype=np.uint8)
        rc = self.L.cc4o_step(self.h, a.ctypes.data, m.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data)
        return rc, obs, float(rew[0]), int(done[0])

    def observe(self):
        obs = np.zeros((5, 210), dtype=np.uint8)
        self.L.cc4o_observe(self.h, obs.ctypes.data)
        return obs

    def mask(self):
        m = np.zeros((5, 242), dtype=np.uint8)
        self.L.cc4o_action_mask(self.h, m.ctypes.data)
        return m

    def export(self):
        buf = np.zeros(32768, dtype=np.int32)
        n = self.L.cc4o_export(self.h, buf.ctypes.data, len(buf))
        return buf[:n].copy()

    def error(self):
        return self.L.cc4o_error_flags(self.h), self.L.cc4o_last_error(self.h).decode()

    def reward_breakdown(self):
        """int32 [3 mission phases][3 kinds: LWF, ASF, RIA][9 subnets, creation order]: BlueRewardMachine events since the reset."""
        out = np.zeros((3, 3, 9), dtype=np.int32)
        self.L.cc4o_reward_breakdown(self.h, out.ctypes.data)
        return out


def run_episodes(seed, first_episode, n_eps, steps=500):
    ret = C.c_double(0)
    ticks = lib().cc4o_run_episodes(seed, first_episode, n_eps, steps, C.byref(ret))
    return ticks, ret.value
