"""Batch configuration: the reference's constructor kwargs -> ``cbx_config``.

Names and defaults follow ``CyberBattleEnv.__init__`` (reference
``_env/cyberbattle_env.py:470-485``), ``AttackerEnvWrapper.__init__``
(``attack_wrapper.py:34-42``) and ``DefenderEnvWrapper.__init__``
(``defend_wrapper.py:34-45``).
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

from . import _abi


class AttackerGoal(NamedTuple):
    """cyberbattle_env.py:227-241"""
    reward: float = 0.0
    low_availability: float = 1.0
    own_atleast: int = 0
    own_atleast_percent: float = 1.0


class DefenderGoal(NamedTuple):
    """cyberbattle_env.py:244-248"""
    eviction: bool


class DefenderConstraint(NamedTuple):
    """cyberbattle_env.py:251-254"""
    maintain_sla: float


class DefenderAgent:
    """_env/defender.py:19-25 -- marker base class; built-in defenders run inside the step kernel."""


class ScanAndReimageCompromisedMachines(DefenderAgent):
    """_env/defender.py:27-55: every ``scan_frequency`` steps scan ``scan_capacity`` random nodes, detect an installed
    agent with ``probability`` and re-image the node (fused into the step kernel; draws come from a counter-based
    Philox stream or from a recorded tape)."""

    def __init__(self, probability: float, scan_capacity: int, scan_frequency: int):
        self.probability = float(probability)
        self.scan_capacity = int(scan_capacity)
        self.scan_frequency = int(scan_frequency)


# gymnasium 0.29.1 `spaces.Dict` sorts the keys of a plain dict, so the reference's flattened attacker action is
# [3, connect(4), local_vulnerability(2), remote_vulnerability(3)] (attack_wrapper.py:206-227). "insertion" gives the
# order of the dict literal in cyberbattle_env.py:540-559 for stacks whose Dict space keeps insertion order.
KIND_ORDERS = {
    "gymnasium029": (_abi.KIND_CONNECT, _abi.KIND_LOCAL, _abi.KIND_REMOTE),
    "insertion": (_abi.KIND_LOCAL, _abi.KIND_REMOTE, _abi.KIND_CONNECT),
}


def make_config(
    mode: int = _abi.MODE_CYBERBATTLE,
    *,
    maximum_total_credentials: int = 1000,
    maximum_node_count: int = 100,
    maximum_discoverable_credentials_per_action: int = 5,
    defender_agent: Optional[DefenderAgent] = None,
    attacker_goal: Optional[AttackerGoal] = AttackerGoal(own_atleast_percent=1.0),
    defender_goal: DefenderGoal = DefenderGoal(eviction=True),
    defender_constraint: DefenderConstraint = DefenderConstraint(maintain_sla=0.0),
    winning_reward: float = 5000.0,
    losing_reward: float = 0.0,
    throws_on_invalid_actions: bool = True,
    seed: int = 0,
    env_index_base: int = 0,
    # AttackerEnvWrapper
    attacker_max_timesteps: int = 2000,
    attacker_invalid_action_reward_modifier: float = -1.0,
    action_kind_order: str | Sequence[int] = "gymnasium029",
    # DefenderEnvWrapper
    defender_enabled: bool = False,
    defender_max_timesteps: int = 100,
    defender_invalid_action_reward: float = 0.0,
    defender_reset_on_constraint_broken: bool = True,
    defender_loss_reward: float = -5000.0,
    defender_sla_worsening_penalty_scale: float = 200.0,
    defender_binding: str = "reference_stale",
    # batched-only knobs
    auto_reset: bool = True,
    mask_mode: int = _abi.MASK_DENSE,
    emit_terminal_obs: bool = False,
) -> _abi.Config:
    c = _abi.Config()
    c.abi_version = _abi.ABI_VERSION
    c.mode = int(mode)
    c.maximum_node_count = int(maximum_node_count)
    c.maximum_total_credentials = int(maximum_total_credentials)
    c.maximum_discoverable_credentials_per_action = int(maximum_discoverable_credentials_per_action or maximum_total_credentials)
    if c.maximum_total_credentials <= 0 or c.maximum_node_count <= 0:
        raise AssertionError("maximum_total_credentials and maximum_node_count must be positive")  # cyberbattle_env.py:208-209
    c.throws_on_invalid_actions = int(bool(throws_on_invalid_actions))
    if attacker_goal:
        c.has_attacker_goal = 1
        c.goal_reward = float(attacker_goal.reward)
        c.goal_low_availability = float(attacker_goal.low_availability)
        c.goal_own_atleast = int(attacker_goal.own_atleast)
        c.goal_own_atleast_percent = float(attacker_goal.own_atleast_percent)
    else:
        c.has_attacker_goal = 0
    c.defender_goal_eviction = int(bool(defender_goal.eviction))
    c.maintain_sla = float(defender_constraint.maintain_sla)
    c.winning_reward = float(winning_reward)
    c.losing_reward = float(losing_reward)
    if defender_agent is None:
        c.builtin_defender = _abi.BUILTIN_NONE
        c.scan_capacity, c.scan_frequency, c.scan_probability = 0, 1, 0.0
    elif type(defender_agent).__name__ == "ScanAndReimageCompromisedMachines":
        c.builtin_defender = _abi.BUILTIN_SCAN_AND_REIMAGE
        c.scan_probability = float(defender_agent.probability)
        c.scan_capacity = int(defender_agent.scan_capacity)
        c.scan_frequency = int(defender_agent.scan_frequency)
        if c.scan_frequency <= 0 or c.scan_capacity < 0:
            raise ValueError("scan_frequency must be positive and scan_capacity non-negative")
    else:
        raise NotImplementedError(f"built-in defender {type(defender_agent).__name__} is not on the batched path "
                                  "(only ScanAndReimageCompromisedMachines; SURVEY.md section 2 row 4)")
    c.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    c.env_index_base = int(env_index_base)
    order = KIND_ORDERS[action_kind_order] if isinstance(action_kind_order, str) else tuple(int(k) for k in action_kind_order)
    if sorted(order) != [0, 1, 2]:
        raise ValueError("action_kind_order must be a permutation of the three action kinds")
    for i, k in enumerate(order):
        c.kind_of_index[i] = k
    c.att_max_timesteps = int(attacker_max_timesteps)
    c.att_invalid_action_reward_modifier = float(attacker_invalid_action_reward_modifier)
    c.def_enabled = int(bool(defender_enabled))
    c.def_max_timesteps = int(defender_max_timesteps)
    c.def_reset_on_constraint_broken = int(bool(defender_reset_on_constraint_broken))
    c.auto_reset = int(bool(auto_reset))
    c.def_invalid_action_reward = float(defender_invalid_action_reward)
    c.def_loss_reward = float(defender_loss_reward)
    c.def_sla_worsening_penalty_scale = float(defender_sla_worsening_penalty_scale)
    c.mask_mode = int(mask_mode)
    c.emit_terminal_obs = int(bool(emit_terminal_obs))
    if defender_binding not in ("reference_stale", "live"):
        raise ValueError("defender_binding: 'reference_stale' (the reference as executed, SURVEY.md B.1) or 'live'")
    c.def_binding = _abi.DEF_BINDING_LIVE if defender_binding == "live" else _abi.DEF_BINDING_STALE
    return c


def attacker_action_layout(cfg: _abi.Config):
    """-> (nvec builder inputs) column slices of the MARLon MultiDiscrete action: {kind: (start, end)}."""
    out, col = {}, 1
    for i in range(3):
        k = cfg.kind_of_index[i]
        out[k] = (col, col + _abi.KIND_WIDTH[k])
        col += _abi.KIND_WIDTH[k]
    return out
