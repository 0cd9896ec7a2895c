"""gym.spaces when gym is installed, otherwise the two space types the USV path uses (Box, Dict)."""
try:  # pragma: no cover
    from gym.spaces import Box, Dict  # type: ignore
except Exception:
    import numpy as np

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape if shape is not None else np.shape(low)).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.low.shape).copy()
            self.shape = self.low.shape
            self.dtype = np.dtype(dtype)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Dict:
        def __init__(self, spaces):
            self.spaces = dict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def items(self):
            return self.spaces.items()

        def keys(self):
            return self.spaces.keys()

        def __repr__(self):
            return f"Dict({self.spaces})"
