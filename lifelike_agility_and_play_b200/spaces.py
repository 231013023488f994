"""Minimal stand-in for ``gym.spaces`` (Box / Dict / Tuple / Discrete) used when ``gym`` is not installed.

The reference declares its spaces with ``gym.spaces`` (primitive_level_env.py:117-124,
create_pybullet_envs.py:9-10); TLeague only reads ``.spaces``, ``.shape`` and ``.dtype`` from them.  If
``gym`` is importable the real classes are used, so a TLeague process sees genuine gym spaces."""
from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - depends on the host
    from gym import spaces as _gym_spaces
    Box, Dict, Tuple, Discrete = _gym_spaces.Box, _gym_spaces.Dict, _gym_spaces.Tuple, _gym_spaces.Discrete
    HAVE_GYM = True
except Exception:
    HAVE_GYM = False

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.dtype = np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

        def __repr__(self):
            return "Box(%s, %s, %s)" % (self.low.min(), self.high.max(), self.shape)

    class Discrete:
        def __init__(self, n):
            self.n, self.shape, self.dtype = int(n), (), np.dtype(np.int64)

    class Dict:
        def __init__(self, spaces):
            self.spaces = OrderedDict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

    class Tuple:
        def __init__(self, spaces):
            self.spaces = tuple(spaces)

        def __getitem__(self, i):
            return self.spaces[i]

        def __len__(self):
            return len(self.spaces)
