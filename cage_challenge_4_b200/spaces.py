"""Minimal stand-ins for the two gymnasium spaces the reference wrappers return (used when gymnasium is absent)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium.spaces import Discrete, MultiDiscrete, MultiBinary  # noqa: F401
except Exception:  # noqa: BLE001
    class Discrete:
        def __init__(self, n: int):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.int64

        def contains(self, x) -> bool:
            return 0 <= int(x) < self.n

        def sample(self) -> int:
            return int(np.random.randint(self.n))

        def __eq__(self, other):
            return isinstance(other, Discrete) and other.n == self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class MultiDiscrete:
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            self.shape = self.nvec.shape
            self.dtype = np.int64

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all((0 <= x) & (x < self.nvec)))

        def sample(self):
            return (np.random.random(self.shape) * self.nvec).astype(np.int64)

        def __len__(self):
            return len(self.nvec)

        def __eq__(self, other):
            return isinstance(other, MultiDiscrete) and np.array_equal(other.nvec, self.nvec)

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"

    class MultiBinary:
        def __init__(self, n: int):
            self.n = int(n)
            self.shape = (self.n,)
            self.dtype = np.int8

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all((x == 0) | (x == 1)))

        def sample(self):
            return np.random.randint(0, 2, self.n).astype(np.int8)

        def __repr__(self):
            return f"MultiBinary({self.n})"
