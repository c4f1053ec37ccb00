"""gymnasium.envs.registration stand-in: registry, EnvSpec, make (test infrastructure)."""
import importlib
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

registry: Dict[str, "EnvSpec"] = {}


@dataclass
class EnvSpec:
    id: str
    entry_point: Any = None
    reward_threshold: Optional[float] = None
    nondeterministic: bool = False
    max_episode_steps: Optional[int] = None
    order_enforce: bool = True
    autoreset: bool = False
    disable_env_checker: bool = False
    apply_api_compatibility: bool = False
    kwargs: dict = field(default_factory=dict)


def register(id, entry_point=None, **kwargs):
    registry[id] = EnvSpec(id, entry_point=entry_point, **kwargs)


def make(id, **kwargs):
    spec = registry[id]
    kwargs.pop("disable_env_checker", None)
    merged = dict(spec.kwargs)
    merged.update(kwargs)
    entry = spec.entry_point
    if isinstance(entry, str):
        mod_name, attr = entry.split(":")
        entry = getattr(importlib.import_module(mod_name), attr)
    env = entry(**merged)
    env.spec = spec
    return env
