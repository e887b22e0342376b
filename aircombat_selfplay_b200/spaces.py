"""Observation / action space descriptors.

The reference declares its spaces with ``gymnasium.spaces`` (reference envs/JSBSim/tasks/task_base.py:2).  When gymnasium
is importable those classes are used unchanged, so algorithms that ``isinstance``-check them keep working; otherwise the
minimal stand-ins below expose the attributes the reference's runners and policies read (``shape``, ``n``, ``nvec``,
``low``/``high``, ``spaces``, ``sample``) -- the reference's algorithms dispatch on ``space.__class__.__name__``
(reference algorithms/utils/act.py), which these names match.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the installation
    from gymnasium.spaces import Box, Discrete, MultiDiscrete, Tuple  # noqa: F401
    HAVE_GYMNASIUM = True
except Exception:  # gymnasium is not installed in the build image
    HAVE_GYMNASIUM = False

    class _Space:
        def __init__(self, shape, dtype):
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return [seed]

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            shape = tuple(shape) if shape is not None else np.shape(low)
            super().__init__(shape, dtype)
            self.low = np.full(shape, low, dtype=dtype)
            self.high = np.full(shape, high, dtype=dtype)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Discrete(_Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = int(n)

        def sample(self):
            return int(self._rng.integers(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class MultiDiscrete(_Space):
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            super().__init__(self.nvec.shape, np.int64)

        def sample(self):
            return (self._rng.random(self.nvec.shape) * self.nvec).astype(np.int64)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"

    class Tuple(_Space):
        def __init__(self, spaces):
            self.spaces = tuple(spaces)
            super().__init__((), object)

        def sample(self):
            return tuple(s.sample() for s in self.spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def __len__(self):
            return len(self.spaces)

        def __iter__(self):
            return iter(self.spaces)

        def __repr__(self):
            return f"Tuple({', '.join(map(repr, self.spaces))})"


def flat_action_dim(space) -> int:
    """Number of scalars one agent's action row carries (how the reference's runners lay actions out in numpy)."""
    name = space.__class__.__name__
    if name == "Discrete":
        return 1
    if name == "MultiDiscrete":
        return int(len(space.nvec))
    if name == "Tuple":
        return sum(flat_action_dim(s) for s in space.spaces)
    if name == "Box":
        return int(np.prod(space.shape))
    raise NotImplementedError(name)
