"""Gym spaces for the drop-in surface.

If the real ``gymnasium`` is importable it is used (so SB3 sees genuine spaces); otherwise a small local
implementation of the five space types the reference's surface needs is used.  Either way ``Dict`` built from a plain
dict orders its keys like gymnasium 0.29.1 does (sorted) -- that order defines the MARLon attacker's flattened
``MultiDiscrete`` layout (reference attack_wrapper.py:206-227), see ``config.KIND_ORDERS``.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Mapping, Optional, Sequence

import numpy as np

try:  # pragma: no cover - depends on the installation
    import gymnasium as _gym

    if str(getattr(_gym, "__version__", "")).endswith("+standin"):
        raise ImportError("test stand-in, not the real package")
    from gymnasium.spaces import Dict as _GDict
    from gymnasium.spaces import Discrete, MultiBinary, MultiDiscrete, Space, Tuple

    HAVE_GYMNASIUM = True

    class Dict(_GDict):
        def __init__(self, spaces=None, seed=None, **kw):
            if isinstance(spaces, Mapping) and not isinstance(spaces, OrderedDict):
                spaces = OrderedDict(sorted(spaces.items()))
            super().__init__(spaces, seed=seed, **kw)

except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Space:
        def __init__(self, shape=None, dtype=None, seed=None):
            self._shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self._rng = np.random.default_rng(seed)

        @property
        def shape(self):
            return self._shape

        @property
        def np_random(self):
            return self._rng

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return [seed]

        def sample(self, mask=None):
            raise NotImplementedError

        def contains(self, x) -> bool:
            raise NotImplementedError

        def __contains__(self, x):
            return self.contains(x)

    class Discrete(Space):
        def __init__(self, n, seed=None, start=0):
            self.n, self.start = int(n), int(start)
            super().__init__((), np.int64, seed)

        def sample(self, mask=None):
            return int(self.start + self._rng.integers(self.n))

        def contains(self, x):
            try:
                return self.start <= int(x) < self.start + self.n
            except Exception:  # noqa: BLE001
                return False

        def __eq__(self, o):
            return isinstance(o, Discrete) and (self.n, self.start) == (o.n, o.start)

        def __repr__(self):
            return f"Discrete({self.n})"

    class MultiDiscrete(Space):
        def __init__(self, nvec, dtype=np.int64, seed=None):
            self.nvec = np.array(nvec, dtype=dtype, copy=True)
            super().__init__(self.nvec.shape, dtype, seed)

        def sample(self, mask=None):
            return (self._rng.random(self.nvec.shape) * self.nvec).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(((x >= 0) & (x < self.nvec)).all())

        def __eq__(self, o):
            return isinstance(o, MultiDiscrete) and np.array_equal(self.nvec, o.nvec)

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"

    class MultiBinary(Space):
        def __init__(self, n, seed=None):
            if isinstance(n, (Sequence, np.ndarray)):
                self.n = tuple(int(i) for i in n)
                shape = self.n
            else:
                self.n = int(n)
                shape = (self.n,)
            super().__init__(shape, np.int8, seed)

        def sample(self, mask=None):
            return self._rng.integers(0, 2, size=self.shape, dtype=np.int8)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(((x == 0) | (x == 1)).all())

        def __eq__(self, o):
            return isinstance(o, MultiBinary) and self.n == o.n

        def __repr__(self):
            return f"MultiBinary({self.n})"

    class Tuple(Space):
        def __init__(self, spaces, seed=None):
            self.spaces = tuple(spaces)
            super().__init__(None, None, seed)

        def sample(self, mask=None):
            return tuple(s.sample() for s in self.spaces)

        def contains(self, x):
            return isinstance(x, (tuple, list)) and len(x) == len(self.spaces)

        def __len__(self):
            return len(self.spaces)

        def __getitem__(self, i):
            return self.spaces[i]

    class Dict(Space):
        def __init__(self, spaces=None, seed=None, **kw):
            if isinstance(spaces, Mapping) and not isinstance(spaces, OrderedDict):
                spaces = OrderedDict(sorted(spaces.items()))
            elif spaces is None:
                spaces = OrderedDict()
            else:
                spaces = OrderedDict(spaces)
            spaces.update(kw)
            self.spaces = spaces
            super().__init__(None, None, seed)

        def sample(self, mask=None):
            return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

        def contains(self, x):
            return isinstance(x, dict) and set(x.keys()) == set(self.spaces.keys())

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def __len__(self):
            return len(self.spaces)


class DummySpace(Space):
    """Placeholder for the non-numeric observation entries (reference cyberbattle_env.py:148-158)."""

    def __init__(self, sample: object):
        self._sample = sample

    def contains(self, x: object) -> bool:
        return True

    def sample(self, mask=None) -> object:
        return self._sample


class DiscriminatedUnion(Dict):
    """Exactly one key present per sample (reference _env/discriminatedunion.py:17-98)."""

    def __init__(self, spaces: Optional[Mapping[str, Any]] = None, seed=None, **kw):
        super().__init__(spaces, seed=None, **kw)
        self.union_np_random = np.random.default_rng(seed if isinstance(seed, int) else None)

    def sample(self, mask=None):
        keys = list(self.spaces.keys())
        k = keys[int(self.union_np_random.integers(0, len(keys)))]
        return OrderedDict([(k, self.spaces[k].sample())])

    def contains(self, x) -> bool:
        return isinstance(x, dict) and len(x) == 1 and next(iter(x)) in self.spaces

    @classmethod
    def is_of_kind(cls, key: str, sample_n: Mapping[str, object]) -> bool:
        return key in sample_n.keys()

    @classmethod
    def kind(cls, sample_n: Mapping[str, object]) -> str:
        keys = sample_n.keys()
        assert len(keys) == 1
        return list(keys)[0]

    def __repr__(self) -> str:
        return self.__class__.__name__ + "(" + ", ".join(f"{k}:{s}" for k, s in self.spaces.items()) + ")"
