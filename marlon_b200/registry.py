"""Gym ids and their default kwargs (reference ``cyberbattle/__init__.py:31-71``)."""
from __future__ import annotations

from typing import Any, Dict, Tuple

from . import model, scenarios
from .config import AttackerGoal, DefenderGoal

ENV_SPECS: Dict[str, Dict[str, Any]] = {
    "CyberBattleToyCtf-v0": dict(defender_agent=None, attacker_goal=AttackerGoal(own_atleast=6),
                                 defender_goal=DefenderGoal(eviction=True)),
    "CyberBattleChain-v0": dict(size=4, defender_agent=None, attacker_goal=AttackerGoal(own_atleast_percent=1.0),
                                defender_goal=DefenderGoal(eviction=True), winning_reward=5000.0, losing_reward=0.0),
    # cyberbattle_random.py:10-14 fixes maximum_discoverable_credentials_per_action=32 and takes no kwargs
    "CyberBattleRandom-v0": dict(maximum_discoverable_credentials_per_action=32),
}
SCENARIO_KWARGS = {"CyberBattleChain-v0": ("size",), "CyberBattleRandom-v0": ("seed",)}


def resolve(env_id: str, **kwargs) -> Tuple[model.Environment, Dict[str, Any]]:
    """-> (scenario environment, CyberBattleEnv ctor kwargs) the way ``gym.make(env_id, **kwargs)`` merges them."""
    if env_id not in ENV_SPECS:
        raise KeyError(f"No registered env with id: {env_id}")
    merged = dict(ENV_SPECS[env_id])
    merged.update(kwargs)
    scn_kw = {k: merged.pop(k) for k in SCENARIO_KWARGS.get(env_id, ()) if k in merged}
    return scenarios.make_environment(env_id, **scn_kw), merged
