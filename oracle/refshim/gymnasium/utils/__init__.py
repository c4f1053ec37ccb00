"""gymnasium.utils stand-in (test infrastructure)."""
from . import seeding  # noqa: F401
