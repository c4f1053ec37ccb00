"""Spaces of the gymnasium stand-in (test infrastructure; see package docstring)."""
from collections import OrderedDict
from collections.abc import Mapping, Sequence

import numpy as np

from ..utils import seeding
from . import space as _space_mod
from .space import Space


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n = int(n)
        self.start = int(start)
        super().__init__((), np.int64, seed)

    def sample(self, mask=None):
        return int(self.start + self.np_random.integers(self.n))

    def contains(self, x):
        try:
            xi = int(x)
        except Exception:
            return False
        return self.start <= xi < self.start + self.n

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n and self.start == other.start

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None, start=None):
        self.nvec = np.array(nvec, dtype=dtype, copy=True)
        assert (self.nvec > 0).all(), "nvec (counts) have to be positive"
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self, mask=None):
        return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool((x >= 0).all() and (x < self.nvec).all())

    def __eq__(self, other):
        return isinstance(other, MultiDiscrete) and np.array_equal(self.nvec, other.nvec)

    def __repr__(self):
        return f"MultiDiscrete({self.nvec})"


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
        return self.np_random.integers(low=0, high=2, size=self.shape, dtype=self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(((x == 0) | (x == 1)).all())

    def __eq__(self, other):
        return isinstance(other, MultiBinary) and self.n == other.n

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

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)

    def __eq__(self, other):
        return isinstance(other, Tuple) and self.spaces == other.spaces


class Dict(Space):
    """Key order follows gymnasium 0.29.1: a plain ``dict`` is SORTED by key."""

    def __init__(self, spaces=None, seed=None, **spaces_kwargs):
        if isinstance(spaces, Mapping) and not isinstance(spaces, OrderedDict):
            try:
                spaces = OrderedDict(sorted(spaces.items()))
            except TypeError:
                spaces = OrderedDict(spaces.items())
        elif isinstance(spaces, Sequence):
            spaces = OrderedDict(spaces)
        elif spaces is None:
            spaces = OrderedDict()
        for k, v in spaces_kwargs.items():
            spaces[k] = v
        self.spaces = spaces
        super().__init__(None, None, seed if not isinstance(seed, dict) else None)

    def sample(self, mask=None):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def contains(self, x):
        return isinstance(x, dict) and len(x) == len(self.spaces)

    def __getitem__(self, key):
        return self.spaces[key]

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def __len__(self):
        return len(self.spaces)

    def __eq__(self, other):
        return isinstance(other, Dict) and self.spaces == other.spaces

    def to_jsonable(self, sample_n):
        return {k: [s[k] for s in sample_n] for k in self.spaces}

    def from_jsonable(self, sample_n):
        n = len(next(iter(sample_n.values())))
        return [OrderedDict((k, sample_n[k][i]) for k in self.spaces) for i in range(n)]


__all__ = ["Space", "Discrete", "MultiDiscrete", "MultiBinary", "Tuple", "Dict", "seeding"]
