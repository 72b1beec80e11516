"""Batched CC4 environment: N concurrent episodes stepped by the CUDA engine through the C-ABI.

Mirrors the call surface of the reference's wrapped environment
(CybORG/Agents/Wrappers/BlueFlatWrapper.py:61 reset, :80 step) with a leading batch dimension:
    reset() -> obs uint8 [N,5,210], mask bool [N,5,242]
    step(actions int32 [N,5], messages uint8 [N,5]) -> obs, reward float32 [N], done bool [N]
Agents 0-3 use the first 92 observation entries / 82 actions (BlueFlatWrapper.py:128-170,
BlueFixedActionWrapper.py:233-309); agent 4 uses 210 / 242.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native

N_BLUE, OBS_LEN, MASK_LEN = 5, 210, 242
N_ACTIONS = (82, 82, 82, 82, 242)
OBS_SIZES = (92, 92, 92, 92, 210)
SLEEP_INDEX = (49, 49, 49, 49, 145)
STEP_OBS, STEP_REWARD, STEP_ALL, STEP_OBS_COMPACT, STEP_OBS_BITS = 1, 2, 3, 4, 8
OBS_BITS_WORDS = 20        # CC4_STEP_OBS_BITS: the compact row as 578 bits in 19 words + the mission phase in word 19
OBS_COMPACT_LEN = 578      # 92 + 92 + 92 + 92 + 210: the five unpadded vectors back to back (CC4_STEP_OBS_COMPACT)
OBS_COMPACT_OFF = (0, 92, 184, 276, 368, 578)


def split_compact(obs):
    """[N, 578] compact observations -> the five per-agent views [N, 92] x 4 + [N, 210] (BlueFlatWrapper.py:128-170)."""
    return [obs[:, OBS_COMPACT_OFF[b]:OBS_COMPACT_OFF[b + 1]] for b in range(N_BLUE)]


def unpack_bits(rows: torch.Tensor) -> torch.Tensor:
    """uint32 / int32 [N, 20] bit rows (CC4_STEP_OBS_BITS) -> uint8 [N, 578] compact observations, on the rows' device."""
    w = rows.to(torch.int64) & 0xFFFFFFFF
    sh = torch.arange(32, device=rows.device, dtype=torch.int64)
    out = ((w[:, :OBS_BITS_WORDS - 1, None] >> sh) & 1).to(torch.uint8).reshape(rows.shape[0], -1)[:, :OBS_COMPACT_LEN].contiguous()
    out[:, list(OBS_COMPACT_OFF[:N_BLUE])] = w[:, OBS_BITS_WORDS - 1, None].to(torch.uint8)
    return out


DRAW_DTYPE = np.dtype([("val", "<u8"), ("n", "<u2"), ("site", "u1"), ("kind", "u1"), ("pad", "<u4")])


class BatchedCC4:
    """N independent CC4 episodes on one GPU.  Episode i draws from the RNG stream (seed, first_episode + i)."""

    RED_AGENTS = {"finite_state": 0, "fsm": 0, "FiniteStateRedAgent": 0, "discovery": 1, "DiscoveryFSRed": 1,
                  "random_select": 2, "RandomSelectRedAgent": 2}

    def __init__(self, num_envs: int, device: int | str | torch.device = 0, seed: int = 0, steps: int = 500,
                 first_episode: int = 0, rng: str = "philox", red_agent: str = "finite_state"):
        if not torch.cuda.is_available():
            raise _native.NativeError("BatchedCC4 needs a CUDA device (there is no CPU fallback)")
        self._L = _native.load()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.num_envs, self.steps, self.seed, self.first_episode = int(num_envs), int(steps), int(seed), int(first_episode)
        self.rng = rng
        h = C.c_void_p()
        _native.check(self._L.cc4_create(self.num_envs, self.device.index or 0, self.seed & (2**64 - 1), self.first_episode,
                                         self.steps, 1 if rng == "tape" else 0, C.byref(h)), "cc4_create")
        self._h = h
        self.red_agent = red_agent
        _native.check(self._L.cc4_set_red_agent(self._h, self.RED_AGENTS[red_agent]), "cc4_set_red_agent")
        n = self.num_envs
        self.obs = torch.zeros((n, N_BLUE, OBS_LEN), dtype=torch.uint8, device=self.device)
        self.mask = torch.zeros((n, N_BLUE, MASK_LEN), dtype=torch.uint8, device=self.device)
        self.reward = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.done = torch.zeros(n, dtype=torch.uint8, device=self.device)
        self._tape = None
        self._pinned = None
        self._pinned_reset = None

    # -- life cycle --------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.cc4_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- API ---------------------------------------------------------------------------------------
    def state_bytes(self) -> tuple[int, int, int]:
        """(hot, cold, static) bytes per episode; only the hot record is streamed every tick."""
        d, k, s = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _native.check(self._L.cc4_state_bytes(self._h, C.byref(d), C.byref(k), C.byref(s)), "cc4_state_bytes")
        return d.value, k.value, s.value

    def set_tape(self, recs: np.ndarray, seg_off: np.ndarray):
        """recs: DRAW_DTYPE [R]; seg_off uint32 [N, n_segs + 1] absolute offsets (tape mode)."""
        assert recs.dtype == DRAW_DTYPE and seg_off.dtype == np.uint32 and seg_off.shape[0] == self.num_envs
        recs_t = torch.from_numpy(recs.view(np.uint8).reshape(-1).copy()).to(self.device)
        seg_t = torch.from_numpy(seg_off.view(np.int32).copy()).to(self.device)      # same bits as uint32
        self._tape = (recs_t, seg_t)
        _native.check(self._L.cc4_set_tape(self._h, recs_t.data_ptr(), seg_t.data_ptr(), seg_off.shape[1] - 1), "cc4_set_tape")

    def set_episode_base(self, seed: int | None = None, first_episode: int | None = None):
        """Re-key the RNG streams used by the next reset(): episode i becomes (seed, first_episode + i)."""
        if seed is not None:
            self.seed = int(seed)
        if first_episode is not None:
            self.first_episode = int(first_episode)
        _native.check(self._L.cc4_set_episode_base(self._h, self.seed & (2**64 - 1), self.first_episode), "cc4_set_episode_base")

    def _check_inputs(self, actions, messages, on_device: bool):
        """Every entry point hands raw pointers to the C-ABI: a wrong dtype, device, shape or a non-contiguous view would be
        reinterpreted silently (an int64 tensor reads as int32 pairs), so all of them check first.  Returns the pointers."""
        n = self.num_envs

        def ptr(x, dtype_t, dtype_np, what):
            if isinstance(x, torch.Tensor):
                ok = x.dtype == dtype_t and x.is_cuda == on_device and tuple(x.shape) == (n, N_BLUE) and x.is_contiguous()
                if ok and on_device:
                    ok = x.device == self.device
                if not ok:
                    raise TypeError(f"{what}: expected a contiguous {dtype_t} tensor [{n}, {N_BLUE}] on {'%s' % self.device if on_device else 'the host'}, "
                                    f"got {x.dtype} {tuple(x.shape)} on {x.device}{'' if x.is_contiguous() else ' (not contiguous)'}")
                return x.data_ptr()
            if on_device or not isinstance(x, np.ndarray):
                raise TypeError(f"{what}: expected a {'CUDA tensor' if on_device else 'host tensor or numpy array'}, got {type(x).__name__}")
            if x.dtype != dtype_np or x.shape != (n, N_BLUE) or not x.flags["C_CONTIGUOUS"]:
                raise TypeError(f"{what}: expected a C-contiguous {np.dtype(dtype_np).name} array [{n}, {N_BLUE}], got {x.dtype} {x.shape}")
            return x.ctypes.data

        return ptr(actions, torch.int32, np.int32, "actions"), (0 if messages is None else ptr(messages, torch.uint8, np.uint8, "messages"))

    def reset(self):
        _native.check(self._L.cc4_reset(self._h, self.obs.data_ptr(), self.mask.data_ptr(), self._stream()), "cc4_reset")
        return self.obs, self.mask.view(torch.bool)       # the kernel writes 0 / 1 bytes: a view, not a conversion kernel over 79 MB

    def reset_host(self):
        """reset() with HOST outputs (pinned): obs uint8 [N,5,210], mask uint8 [N,5,242]."""
        n = self.num_envs
        if self._pinned_reset is None:
            self._pinned_reset = dict(obs=torch.zeros((n, N_BLUE, OBS_LEN), dtype=torch.uint8).pin_memory(),
                                      mask=torch.zeros((n, N_BLUE, MASK_LEN), dtype=torch.uint8).pin_memory())
        p = self._pinned_reset
        _native.check(self._L.cc4_reset_host(self._h, p["obs"].data_ptr(), p["mask"].data_ptr(), self._stream()), "cc4_reset_host")
        return p["obs"], p["mask"]

    def step(self, actions: torch.Tensor, messages: torch.Tensor | None = None, flags: int = STEP_ALL):
        assert not (flags & (STEP_OBS_COMPACT | STEP_OBS_BITS)), "compact / bit observations: use step_compact() / step_bits() / step_host(compact=True)"
        ap, mp = self._check_inputs(actions, messages, True)
        _native.check(self._L.cc4_step(self._h, ap, mp, self.obs.data_ptr(), self.reward.data_ptr(),
                                       self.done.data_ptr(), flags, self._stream()), "cc4_step")
        return self.obs, self.reward, self.done

    def step_host(self, actions: np.ndarray | torch.Tensor, messages=None, flags: int = STEP_ALL, compact: bool = False, bits: bool = False):
        """The same tick with HOST buffers (pinned torch tensors or numpy arrays): H2D + kernel + D2H + sync inside.
        compact=True returns the observations as uint8 [N, 578] (the five unpadded vectors back to back, see
        split_compact) instead of [N, 5, 210]: 45 % fewer bytes over PCIe.  bits=True returns them as bit rows int32
        [N, 20] (CC4_STEP_OBS_BITS, see unpack_bits): 80 B per episode, nothing to expand on the host."""
        n = self.num_envs
        a_ptr, m_ptr = self._check_inputs(actions, messages, False)
        if self._pinned is None:
            self._pinned = dict(obs=torch.zeros((n, N_BLUE, OBS_LEN), dtype=torch.uint8).pin_memory(),
                                reward=torch.zeros(n, dtype=torch.float32).pin_memory(),
                                done=torch.zeros(n, dtype=torch.uint8).pin_memory())
        p = self._pinned
        if bits:
            if "obs_bits" not in p:
                p["obs_bits"] = p["obs"].view(-1)[:n * OBS_BITS_WORDS * 4].view(torch.int32).view(n, OBS_BITS_WORDS)     # same pinned storage
            _native.check(self._L.cc4_step_host(self._h, a_ptr, m_ptr, p["obs_bits"].data_ptr(), p["reward"].data_ptr(), p["done"].data_ptr(),
                                                flags | STEP_OBS_BITS, self._stream()), "cc4_step_host")
            return p["obs_bits"], p["reward"], p["done"]
        if compact:
            flags |= STEP_OBS_COMPACT
            if "obs_compact" not in p:
                p["obs_compact"] = p["obs"].view(-1)[:n * OBS_COMPACT_LEN].view(n, OBS_COMPACT_LEN)     # same pinned storage
                p["ptrs"] = (p["obs_compact"].data_ptr(), p["reward"].data_ptr(), p["done"].data_ptr())
            o_ptr, r_ptr, d_ptr = p["ptrs"]
            rc = self._L.cc4_step_host(self._h, a_ptr, m_ptr, o_ptr, r_ptr, d_ptr, flags, self._stream())
            if rc:
                _native.check(rc, "cc4_step_host")
            return p["obs_compact"], p["reward"], p["done"]
        _native.check(self._L.cc4_step_host(self._h, a_ptr, m_ptr, p["obs"].data_ptr(), p["reward"].data_ptr(),
                                            p["done"].data_ptr(), flags, self._stream()), "cc4_step_host")
        return p["obs"], p["reward"], p["done"]

    def step_compact(self, actions: torch.Tensor, messages: torch.Tensor | None = None):
        """step() with compact device observations uint8 [N, 578] (see split_compact)."""
        if getattr(self, "_obs_compact", None) is None:
            self._obs_compact = torch.zeros((self.num_envs, OBS_COMPACT_LEN), dtype=torch.uint8, device=self.device)
        ap, mp = self._check_inputs(actions, messages, True)
        _native.check(self._L.cc4_step(self._h, ap, mp, self._obs_compact.data_ptr(), self.reward.data_ptr(),
                                       self.done.data_ptr(), STEP_ALL | STEP_OBS_COMPACT, self._stream()), "cc4_step")
        return self._obs_compact, self.reward, self.done

    def step_state_only(self, actions: torch.Tensor, messages: torch.Tensor | None = None):
        """The state transition alone (BASELINE config 2: observations and reward off); `done` is still written."""
        ap, mp = self._check_inputs(actions, messages, True)
        _native.check(self._L.cc4_step(self._h, ap, mp, 0, 0, self.done.data_ptr(), 0, self._stream()), "cc4_step")
        return self.done

    def capture_steps(self, actions_list, mode: str = "compact", accumulate: torch.Tensor | None = None):
        """Record len(actions_list) consecutive ticks (one k_step launch each, reading actions_list[i]) into ONE CUDA graph
        and return it; graph.replay() then costs one launch from the host for all of them, which matters when a tick is
        short (a sub-wave batch such as 8 192 episodes per GPU: ~0.07 ms of kernel against the per-launch cost of a Python
        call).  The outputs of the last captured tick are in the usual buffers; `accumulate` (float64 [N]) collects the
        reward of every tick.  Use an even number of ticks: successive ticks sweep the batch in alternating directions."""
        step = {"compact": self.step_compact, "padded": self.step, "bits": self.step_bits, "state": self.step_state_only}[mode]
        step(actions_list[0])                      # allocates the lazily created output buffers outside the capture
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for a in actions_list:
                step(a)
                if accumulate is not None:
                    accumulate.add_(self.reward)
        return g

    def step_bits(self, actions: torch.Tensor, messages: torch.Tensor | None = None):
        """step() with the observations as bit rows int32 [N, 20] on the device (CC4_STEP_OBS_BITS, 80 B per episode; see
        unpack_bits): for device-side policies that take bits, and what step_host moves over PCIe."""
        if getattr(self, "_obs_bits", None) is None:
            self._obs_bits = torch.zeros((self.num_envs, OBS_BITS_WORDS), dtype=torch.int32, device=self.device)
        ap, mp = self._check_inputs(actions, messages, True)
        _native.check(self._L.cc4_step(self._h, ap, mp, self._obs_bits.data_ptr(), self.reward.data_ptr(),
                                       self.done.data_ptr(), STEP_ALL | STEP_OBS_BITS, self._stream()), "cc4_step")
        return self._obs_bits, self.reward, self.done

    def rollout(self, n_ticks: int, policy_seed: int = 0, reward_sum: torch.Tensor | None = None, flags: int = STEP_REWARD):
        rp = 0 if reward_sum is None else reward_sum.data_ptr()
        _native.check(self._L.cc4_rollout(self._h, int(n_ticks), int(policy_seed), rp, flags, self._stream()), "cc4_rollout")

    def error_flags(self) -> np.ndarray:
        out = np.zeros(self.num_envs, dtype=np.uint32)
        _native.check(self._L.cc4_error_flags(self._h, out.ctypes.data), "cc4_error_flags")
        return out

    def export_canon(self, env: int) -> np.ndarray:
        buf = np.zeros(32768, dtype=np.int32)
        n = C.c_int(0)
        _native.check(self._L.cc4_export_canon(self._h, int(env), buf.ctypes.data, len(buf), C.byref(n)), "cc4_export_canon")
        return buf[:n.value].copy()

    def launch_count(self) -> int:
        v = C.c_uint64(0)
        _native.check(self._L.cc4_launch_count(self._h, C.byref(v)), "cc4_launch_count")
        return v.value
