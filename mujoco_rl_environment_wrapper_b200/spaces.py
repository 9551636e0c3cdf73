"""Minimal Box space (gymnasium.spaces.Box is used when gymnasium is importable).
The reference builds `Box(low=np.array(low), high=np.array(high))` (MuJoCo_Gym/mujoco_rl.py:191-192,211-212)."""
import numpy as np

try:  # pragma: no cover - gymnasium is not part of this image
    from gymnasium.spaces import Box  # type: ignore
except Exception:
    class Box:
        def __init__(self, low, high, dtype=np.float32, seed=None):
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape
            self.dtype = np.dtype(dtype)
            self._rng = np.random.default_rng(seed)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e6)
            hi = np.where(np.isfinite(self.high), self.high, 1e6)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"
