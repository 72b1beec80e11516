import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = shape
        self.dtype = dtype
        self._np_random = None

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


class Discrete(Space):
    def __init__(self, n, start=0):
        super().__init__((), np.int64)
        self.n = int(n)
        self.start = int(start)

    def contains(self, x):
        try:
            return self.start <= int(x) < self.start + self.n
        except (TypeError, ValueError):
            return False

    def sample(self):
        rng = self._np_random or np.random.default_rng()
        return int(self.start + rng.integers(self.n))

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n and self.start == other.start

    def __hash__(self):
        return hash((self.n, self.start))

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiBinary(Space):
    def __init__(self, n):
        super().__init__((int(n),), np.int8)
        self.n = int(n)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == (self.n,) and bool(np.all((x == 0) | (x == 1)))

    def sample(self):
        rng = self._np_random or np.random.default_rng()
        return rng.integers(0, 2, size=self.n).astype(np.int8)

    def __repr__(self):
        return f"MultiBinary({self.n})"


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64):
        self.nvec = np.asarray(nvec, dtype=dtype)
        super().__init__(self.nvec.shape, dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.nvec.shape and bool(np.all((x >= 0) & (x < self.nvec)))

    def sample(self):
        rng = self._np_random or np.random.default_rng()
        return (rng.random(self.nvec.shape) * self.nvec).astype(self.dtype)

    def __len__(self):
        return len(self.nvec)

    def __eq__(self, other):
        return isinstance(other, MultiDiscrete) and np.array_equal(self.nvec, other.nvec)

    def __hash__(self):
        return hash(self.nvec.tobytes())

    def __repr__(self):
        return f"MultiDiscrete({self.nvec.tolist()})"


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        super().__init__(shape, dtype)
        self.low, self.high = low, high
