"""Minimal stand-in for gymnasium 0.29.1 -- TEST INFRASTRUCTURE ONLY.

The reference pins ``gymnasium==0.29.1`` (``/root/reference/requirements.txt``)
but the package is not installable here (no network, not in the wheelhouse).
This stand-in implements just the surface the reference's hot path touches so
the *unmodified* reference can be imported in this container to generate the
golden tapes under ``tests/golden/`` (see ``oracle/gen_golden.py``).

Behaviour that matters for parity and is reproduced deliberately:

* ``spaces.Dict`` built from a plain ``dict`` SORTS its keys (gymnasium 0.29.1
  ``spaces/dict.py``: ``OrderedDict(sorted(spaces.items()))``).  The MARLon
  attacker wrapper derives its flattened ``MultiDiscrete`` layout from
  ``cyber_env.action_space.spaces.items()`` (attack_wrapper.py:216-225), so the
  action kinds come out as 0=connect, 1=local_vulnerability,
  2=remote_vulnerability.
* ``utils.seeding.np_random`` returns a PCG64 ``numpy.random.Generator``.

Nothing in the product package imports this module.
"""
from . import spaces, utils, envs, error  # noqa: F401
from .core import Env, Wrapper, ObservationWrapper, ActionWrapper  # noqa: F401
from .spaces.space import Space  # noqa: F401
from .envs.registration import make, register, registry  # noqa: F401

__version__ = "0.29.1+standin"
