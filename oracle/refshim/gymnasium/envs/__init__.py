"""gymnasium.envs stand-in (test infrastructure)."""
from . import registration  # noqa: F401
