"""Import the UNMODIFIED reference (``/root/reference``) in this container.

TEST INFRASTRUCTURE ONLY -- used by ``oracle/gen_golden.py`` and by the
``reference``-marked tests that re-validate the oracle against the live
reference when ``/root/reference`` is present.  Nothing here runs on the GPU
box (``/root/reference`` does not exist there) and nothing in ``marlon_b200``
imports it.

Recipe:
  1. ``sys.path`` <- reference source roots + ``oracle/refshim`` (stand-ins for
     the pinned-but-absent ``gymnasium==0.29.1`` and ``boolean.py==4.0``).
  2. cosmetic imports (plotly, IPython, matplotlib, progressbar, ...) become
     ``MagicMock`` modules; ``cyberbattle.agents`` (torch DQL agents, out of
     scope) is pre-seeded as a mock so ``import cyberbattle`` stays light.
  3. ``numpy.can_cast`` accepts Python ints again (reference pins numpy 1.26;
     ``EnvironmentBounds.of_identifiers`` calls ``np.can_cast(int, np.int32)``,
     cyberbattle_env.py:206-214, which raises under NEP 50).
Nothing under ``/root/reference`` is modified.
"""
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("CBX_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")
_loaded = False


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "CyberBattleSim", "cyberbattle"))


def load():
    """Make ``cyberbattle`` and ``marlon`` (reference) importable. Idempotent."""
    global _loaded
    if _loaded:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for p in (_SHIM, os.path.join(REFERENCE_ROOT, "src", "CyberBattleSim"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)

    for name in [
        "plotly", "plotly.graph_objects", "plotly.subplots", "plotly.missing_ipywidgets",
        "plotly.utils", "plotly.express", "IPython", "IPython.display", "IPython.core",
        "IPython.core.display", "matplotlib", "matplotlib.pyplot", "progressbar",
        "asciichartpy", "cyberbattle.agents",
    ]:
        if name not in sys.modules:
            m = mock.MagicMock(name=name)
            m.__path__ = []  # looks like a package
            m.__spec__ = None
            sys.modules[name] = m

    import numpy as np

    if not getattr(np.can_cast, "_cbx_patched", False):
        _orig = np.can_cast

        def can_cast(from_, to, casting="safe"):
            if isinstance(from_, (int, np.integer)) and not isinstance(from_, bool):
                info = np.iinfo(to)
                return info.min <= int(from_) <= info.max
            return _orig(from_, to, casting)

        can_cast._cbx_patched = True
        np.can_cast = can_cast

    import cyberbattle  # noqa: F401  (registers the gym ids)

    _loaded = True


def make(env_id, **kwargs):
    load()
    import gymnasium as gym

    return gym.make(env_id, **kwargs)
