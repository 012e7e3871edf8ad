"""Observation/action space containers.

The reference builds ``gymnasium.spaces`` objects (wfcrl/mdp.py:108-153, wfcrl/multiagent_env.py:70-88).  gymnasium is an
optional dependency here: when it is importable its classes are used unchanged, otherwise the minimal stand-ins below
provide the attributes this package (and typical RL code) reads: ``low``/``high``/``shape``/``dtype``, ``sample``,
``contains``, dict-style access.  Semantics follow gymnasium 0.29.1: ``Box`` defaults to float32 and scalar bounds with no
shape give shape ``(1,)`` (examples/demo.ipynb, last cell).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - exercised only where gymnasium is installed
    from gymnasium.spaces import Box, Dict, MultiDiscrete  # noqa: F401

    HAVE_GYMNASIUM = True
except Exception:  # gymnasium absent: minimal stand-ins
    HAVE_GYMNASIUM = False

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low) if np.ndim(low) > 0 else (np.shape(high) if np.ndim(high) > 0 else (1,))
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            lo = self.low.flat[0] if np.all(self.low == self.low.flat[0]) else self.low
            hi = self.high.flat[0] if np.all(self.high == self.high.flat[0]) else self.high
            return f"Box({lo}, {hi}, {self.shape}, {self.dtype})"

    class MultiDiscrete:
        def __init__(self, nvec, dtype=np.int64, seed=None):
            self.nvec = np.asarray(nvec, dtype=dtype)
            self.shape = self.nvec.shape
            self.dtype = np.dtype(dtype)
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return (self._rng.random(self.shape) * self.nvec).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

        def __getitem__(self, i):
            return _Discrete(int(self.nvec[i]))

        def __repr__(self):
            return f"MultiDiscrete({self.nvec})"

    class _Discrete:
        def __init__(self, n):
            self.n = n
            self.shape = ()
            self.dtype = np.dtype(np.int64)

        def sample(self):
            return int(np.random.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class Dict(OrderedDict):
        """Ordered mapping name -> space (``spaces.Dict``)."""

        def __init__(self, spaces=None):
            super().__init__(spaces or {})
            self.spaces = self

        def sample(self):
            return OrderedDict((k, s.sample()) for k, s in self.items())

        def contains(self, x):
            return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.items())

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {v}" for k, v in self.items()) + ")"
