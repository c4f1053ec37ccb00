"""Single-env drop-ins for MARLon's wrappers over the fused step kernel.

``AttackerEnvWrapper`` (reference attack_wrapper.py:20-553), ``DefenderEnvWrapper`` (defend_wrapper.py:25-549),
``MaskedDiscreteAttackerWrapper`` (action_masking.py:30-165) and ``EnvironmentEventSource``
(environment_event_source.py:16-38) with the reference's constructor signatures, spaces, return shapes, counters and
reset protocol.  Both wrappers of one ``CyberBattleEnv`` share a 1-env MARLon-mode batch: ``attacker.step`` runs the
attacker half of the pair step, ``defender.step`` the defender half (``cbx_batch_step_ex``); ``reset`` maps to
``cbx_batch_reset_ex`` with the same ``who``.  The reset_request / notify_reset protocol lives in the kernel's state.
These objects are the legacy one-env view; rollouts at scale use ``MultiAgentUniversalEnv``.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import _abi, config, spaces
from .cyberbattle_env import CyberBattleEnv
from .universe import ATTACKER_OBS_KEYS, attacker_spaces, defender_spaces

WHO_ATT, WHO_DEF = 1, 2


class IEnvironmentObserver:
    def on_reset(self, last_reward):
        raise NotImplementedError


class IRewardStore:
    @property
    def episode_rewards(self) -> List[float]:
        raise NotImplementedError


class EnvironmentEventSource:
    """environment_event_source.py:16-38"""

    def __init__(self):
        self.observers: List[IEnvironmentObserver] = []

    def add_observer(self, observer: IEnvironmentObserver):
        self.observers.append(observer)

    def notify_reset(self, last_reward):
        for o in self.observers:
            o.on_reset(last_reward)


class _Binding:
    """The 1-env MARLon batch shared by the two wrappers of a CyberBattleEnv (created lazily, on first use)."""

    def __init__(self, cyber_env: CyberBattleEnv):
        self.cyber_env = cyber_env
        self.att_kwargs: Dict[str, Any] = {}
        self.def_kwargs: Optional[Dict[str, Any]] = None
        self.batch = None

    def get(self):
        if self.batch is None:
            from .batch import Batch

            a, d = self.att_kwargs, self.def_kwargs or {}
            cfg = config.make_config(
                _abi.MODE_MARLON, auto_reset=False,
                attacker_max_timesteps=a.get("max_timesteps", 2000),
                attacker_invalid_action_reward_modifier=a.get("invalid_action_reward_modifier", -1),
                defender_enabled=self.def_kwargs is not None, defender_max_timesteps=d.get("max_timesteps", 100),
                defender_invalid_action_reward=d.get("invalid_action_reward", 0),
                defender_reset_on_constraint_broken=d.get("reset_on_constraint_broken", True),
                defender_loss_reward=d.get("loss_reward", -5000.0),
                defender_sla_worsening_penalty_scale=d.get("sla_worsening_penalty_scale", 200.0),
                **self.cyber_env.env_kwargs)
            self.cfg = cfg
            self.batch = Batch(self.cyber_env.compiled, cfg, 1, device=self.cyber_env.device)
            self.cyber_env._marlon_batch = self.batch
        return self.batch

    def invalidate(self):
        if self.batch is not None:
            self.batch.close()
            self.batch = None
            self.cyber_env._marlon_batch = None


def _binding(cyber_env: CyberBattleEnv) -> _Binding:
    b = getattr(cyber_env, "_marlon_binding", None)
    if b is None:
        b = _Binding(cyber_env)
        cyber_env._marlon_binding = b
    return b


class AttackerEnvWrapper(IRewardStore, IEnvironmentObserver):
    int32_spaces = ATTACKER_OBS_KEYS
    _log = logging.getLogger("cyberbattle.attacker")
    metadata = {"render_modes": []}
    spec = None
    render_mode = None

    def __init__(self, cyber_env: CyberBattleEnv, event_source: Optional[EnvironmentEventSource] = None, max_timesteps=2000,
                 invalid_action_reward_modifier=-1, invalid_action_reward_multiplier=1, loss_reward=-5000,
                 log_episode_end: bool = False, episode_log_prefix: str = ""):
        self.cyber_env = cyber_env
        self._base_env = cyber_env.unwrapped
        self.bounds = self._base_env.bounds
        self.max_timesteps = max_timesteps
        self.invalid_action_reward_modifier = invalid_action_reward_modifier
        self.invalid_action_reward_multiplier = invalid_action_reward_multiplier  # stored, never used (attack_wrapper.py:51)
        self.loss_reward = loss_reward
        self._bind = _binding(self._base_env)
        self._bind.invalidate()
        self._bind.att_kwargs = dict(max_timesteps=max_timesteps, invalid_action_reward_modifier=invalid_action_reward_modifier)
        cfg = config.make_config(_abi.MODE_MARLON, **self._base_env.env_kwargs)
        self.observation_space, self.action_space = attacker_spaces(cfg, self._base_env.compiled)
        lay = config.attacker_action_layout(cfg)
        self.action_subspaces = {i: (_abi.KIND_NAMES[cfg.kind_of_index[i]],) + lay[cfg.kind_of_index[i]] for i in range(3)}
        self.node_count = int(self.bounds.maximum_node_count)
        self.timesteps = None
        self.cyber_rewards: List[float] = []
        self.rewards: List[float] = []
        self.valid_action_count = self.invalid_action_count = 0
        self.last_valid_action_count = self.last_invalid_action_count = 0
        self.last_action = None
        self.last_is_invalid = False
        self.last_cyber_reward = self.last_reward = 0.0
        self.last_terminated = self.last_truncated = False
        self.last_outcome: Optional[str] = None
        self.last_attempted_action_valid: Optional[bool] = None
        self._winning_reward = self._base_env._CyberBattleEnv__WINNING_REWARD
        self._losing_reward = self._base_env._CyberBattleEnv__LOSING_REWARD
        self._last_transformed_observation = None
        self._last_info: Dict[str, Any] = {}
        self._log_episode_end, self._episode_log_prefix = bool(log_episode_end), str(episode_log_prefix)
        self._step_log_enabled, self._step_log_prefix = False, ""
        self.event_source = event_source or EnvironmentEventSource()
        self.event_source.add_observer(self)

    def configure_step_logging(self, *, enabled: bool, prefix: str = "") -> None:
        self._step_log_enabled, self._step_log_prefix = bool(enabled), str(prefix)

    def configure_episode_end_logging(self, *, enabled: bool, prefix: str = "") -> None:
        self._log_episode_end, self._episode_log_prefix = bool(enabled), str(prefix)

    @property
    def unwrapped(self):
        return self

    @property
    def reset_request(self) -> bool:
        return bool(self._bind.get().export_state(0, 1)[0][6])

    def _observation(self) -> Dict[str, Any]:
        b = self._bind.get()
        sc = b.numpy("scalars")[0]
        obs = {k: int(sc[i]) for i, k in enumerate(ATTACKER_OBS_KEYS)}
        for k in ("leaked_credentials", "credential_cache_matrix", "discovered_nodes_properties", "nodes_privilegelevel",
                  "local_vulnerability", "remote_vulnerability", "connect"):
            obs[k] = b.numpy(k)[0].copy()
        return obs

    def step(self, action) -> Tuple[Dict[str, Any], float, bool, bool, Dict[str, Any]]:
        """attack_wrapper.py:255-398"""
        if self.timesteps is None:
            self.reset()
        b = self._bind.get()
        self.last_action = np.array(action, copy=True)
        act = np.asarray(action, dtype=np.int32).reshape(1, -1)
        b.step(act, None, who=WHO_ATT)
        info_raw = b.numpy("att_info")[0]
        intercepted = bool(info_raw[5])
        err = int(info_raw[3])
        if err == _abi.E_STEP_AFTER_DONE:
            raise RuntimeError("new episode must be started with env.reset()")
        if err in (_abi.E_SOURCE_NOT_OWNED, _abi.E_TARGET_NOT_DISCOVERED, _abi.E_CREDENTIAL_NOT_GATHERED):
            raise ValueError({1: "Agent does not owned the source node", 2: "Agent has not discovered the target node",
                              3: "Agent has not discovered credential"}[err])
        reward = float(b.numpy("att_reward")[0])
        terminated = bool(b.numpy("att_terminated")[0])
        truncated = bool(b.numpy("att_truncated")[0])
        cyber_reward = float(info_raw[0:1].view(np.float32)[0])
        self.last_attempted_action_valid = not intercepted
        self.last_is_invalid = intercepted
        if intercepted:
            self.invalid_action_count += 1
            obs = self._last_transformed_observation
            info = dict(self._last_info)
            info["invalid_action"] = True
            info["cyber_step_executed"] = False
        else:
            self.valid_action_count += 1
            obs = self._observation()
            self._last_transformed_observation = obs
            info = {"description": "CyberBattle simulation", "duration_in_ms": 0.0, "step_count": int(info_raw[4]),
                    "network_availability": float(b.numpy("network_availability")[0]),
                    "credential_cache": self._base_env.credential_cache}
            self._last_info = dict(info)
        self.cyber_rewards.append(cyber_reward)
        self.last_cyber_reward = cyber_reward
        self.last_outcome = None
        if terminated:
            if cyber_reward == float(self._winning_reward):
                self.last_outcome = "attacker_win"
            elif cyber_reward == float(self._losing_reward):
                self.last_outcome = "attacker_loss"
            else:
                self.last_outcome = "terminated"
        self.timesteps += 1
        if truncated and self.timesteps >= self.max_timesteps and self.last_outcome is None:
            self.last_outcome = "timeout"
        self.rewards.append(reward)
        self.last_reward, self.last_terminated, self.last_truncated = reward, terminated, truncated
        return obs, reward, terminated, truncated, info

    def reset(self, *, seed=None, options=None):
        """attack_wrapper.py:400-468"""
        b = self._bind.get()
        self.last_valid_action_count, self.last_invalid_action_count = self.valid_action_count, self.invalid_action_count
        b.reset(who=WHO_ATT)  # notifies the defender's wrapper (reset_request) inside the kernel when no reset is pending
        self.valid_action_count = self.invalid_action_count = 0
        self.timesteps = 0
        self.cyber_rewards, self.rewards = [], []
        self.last_action = None
        self.last_is_invalid = False
        self.last_cyber_reward = self.last_reward = 0.0
        self.last_terminated = self.last_truncated = False
        self.last_outcome = None
        self.last_attempted_action_valid = None
        self._last_transformed_observation = self._observation()
        info = {"description": "CyberBattle simulation", "duration_in_ms": 0, "step_count": 0,
                "network_availability": float(b.numpy("network_availability")[0]), "credential_cache": []}
        self._last_info = dict(info)
        return self._last_transformed_observation, info

    def on_reset(self, last_rewards):
        self._bind.get().notify_reset(WHO_ATT, 0.0)

    def transform_observation(self, observation):
        return observation  # observations already come out of the encoder in the normalised form

    @property
    def episode_rewards(self) -> List[float]:
        return self.cyber_rewards

    def close(self) -> None:
        self._bind.invalidate()

    def render(self, mode: str = "human") -> None:
        raise NotImplementedError("rendering is out of scope of the batched step path")


class DefenderEnvWrapper(IEnvironmentObserver):
    firewall_rule_list = ["RDP", "SSH", "HTTPS", "HTTP", "su", "sudo"]
    _log = logging.getLogger("cyberbattle.defender")
    metadata = {"render_modes": []}
    spec = None
    render_mode = None

    def __init__(self, cyber_env: CyberBattleEnv, attacker_reward_store: IRewardStore,
                 event_source: Optional[EnvironmentEventSource] = None, defender: bool = False, max_timesteps=100,
                 invalid_action_reward=0, reset_on_constraint_broken=True, loss_reward: float = -5000.0,
                 sla_worsening_penalty_scale: float = 200.0, log_episode_end: bool = False, episode_log_prefix: str = ""):
        assert defender is not None, "Attempting to use the defender environment without a defender present."
        self.cyber_env = cyber_env
        self._base_env = cyber_env.unwrapped
        self.bounds = self._base_env.bounds
        self.attacker_reward_store = attacker_reward_store
        self.max_timesteps = max_timesteps
        self.invalid_action_penalty = invalid_action_reward
        self.reset_on_constraint_broken = reset_on_constraint_broken
        self.loss_reward = loss_reward
        self.sla_worsening_penalty_scale = sla_worsening_penalty_scale
        self._defender_constraint = self._base_env._CyberBattleEnv__defender_constraint
        self._bind = _binding(self._base_env)
        self._bind.invalidate()
        self._bind.def_kwargs = dict(max_timesteps=max_timesteps, invalid_action_reward=invalid_action_reward,
                                     reset_on_constraint_broken=reset_on_constraint_broken, loss_reward=loss_reward,
                                     sla_worsening_penalty_scale=sla_worsening_penalty_scale)
        self.observation_space, self.action_space = defender_spaces(self._base_env.compiled)
        self.num_services = self._base_env.compiled.n_services
        self.timesteps = 0
        self.rewards: List[float] = []
        self.valid_action_count = self.invalid_action_count = 0
        self.last_valid_action_count = self.last_invalid_action_count = 0
        self.last_action = self.last_action_valid = None
        self.last_reward = 0.0
        self.last_terminated = self.last_truncated = False
        self.last_outcome: Optional[str] = None
        self.event_source = event_source or EnvironmentEventSource()
        self.event_source.add_observer(self)

    @property
    def unwrapped(self):
        return self

    def _state(self):
        return self._bind.get().export_state(0, 1)[0]

    @property
    def reset_request(self) -> bool:
        return bool(self._state()[7])

    @property
    def network_availability(self) -> float:
        n = self._base_env.compiled.n_nodes
        return (n - int(self._state()[14])) / n

    last_availability = network_availability

    @property
    def last_sla_breached(self) -> bool:
        return self.network_availability < float(self._defender_constraint.maintain_sla)

    def observe(self) -> Dict[str, np.ndarray]:
        b = self._bind.get()
        return {"infected_nodes": b.numpy("def_infected_nodes")[0].astype(np.int64),
                "incoming_firewall_status": b.numpy("def_incoming_firewall")[0].astype(np.int64),
                "outgoing_firewall_status": b.numpy("def_outgoing_firewall")[0].astype(np.int64),
                "services_status": b.numpy("def_services_status")[0].astype(np.int64)}

    def step(self, action):
        """defend_wrapper.py:197-327 (+ LearningDefender.executeAction on the stale copy)"""
        b = self._bind.get()
        if action is None or (hasattr(action, "__len__") and len(action) == 0):
            act = np.full((1, 12), -1, dtype=np.int32)
            self.last_action = np.array([], dtype=int)
        else:
            act = np.asarray(action, dtype=np.int32).reshape(1, 12)
            self.last_action = np.array(action, copy=True)
        before = self._state()
        b.step(None, act, who=WHO_DEF)
        after = self._state()
        self.last_action_valid = bool(after[11] > before[11])
        if self.last_action_valid:
            self.valid_action_count += 1
        else:
            self.invalid_action_count += 1
        reward = float(b.numpy("def_reward")[0])
        terminated, truncated = bool(b.numpy("def_terminated")[0]), bool(b.numpy("def_truncated")[0])
        self.timesteps += 1
        self.rewards.append(reward)
        self.last_reward, self.last_terminated, self.last_truncated = reward, terminated, truncated
        if terminated:
            self.last_outcome = "defender_win" if reward == float(self._base_env._CyberBattleEnv__WINNING_REWARD) else "sla_breached"
        elif truncated and self.timesteps >= self.max_timesteps:
            self.last_outcome = "timeout"
        return self.observe(), reward, terminated, truncated, {}

    def is_defender_action_valid(self, action) -> bool:
        """defend_wrapper.py:329-412 (live environment)"""
        st = self._base_env._state()
        comp = self._base_env.compiled
        running = st["countdown"] == 0
        nodes = [self._base_env.environment.get_node(k) for k in comp.node_ids]
        a = [int(x) for x in action]
        if a[0] == 0:
            return bool(running[a[1]] and nodes[a[1]].reimagable)
        if a[0] == 1:
            rules = nodes[a[2]].firewall.incoming if a[4] else nodes[a[2]].firewall.outgoing
            return bool(running[a[2]] and self.firewall_rule_list[a[3]] in [r.port for r in rules])
        if a[0] == 2:
            return bool(running[a[5]])
        if a[0] == 3:
            return bool(running[a[8]] and a[9] < len(nodes[a[8]].services))
        if a[0] == 4:
            return bool(running[a[10]] and a[11] < len(nodes[a[10]].services))
        return False

    def reset(self, *, seed=None, options=None):
        """defend_wrapper.py:414-477"""
        b = self._bind.get()
        b.reset(who=WHO_DEF)
        self.rewards, self.timesteps = [], 0
        self.last_valid_action_count, self.last_invalid_action_count = self.valid_action_count, self.invalid_action_count
        self.valid_action_count = self.invalid_action_count = 0
        self.last_action = self.last_action_valid = None
        self.last_reward = 0.0
        self.last_terminated = self.last_truncated = False
        self.last_outcome = None
        return self.observe(), {"description": "CyberBattle simulation", "duration_in_ms": 0, "step_count": 0,
                                "network_availability": 1.0, "credential_cache": []}

    def on_reset(self, last_reward):
        self._bind.get().notify_reset(WHO_DEF, float(last_reward))

    def set_reset_request(self, reset_request):
        if reset_request:
            self._bind.get().notify_reset(WHO_DEF, 0.0)

    def defender_constraints_broken(self):
        return self.network_availability < self._defender_constraint.maintain_sla

    def close(self) -> None:
        self._bind.invalidate()


class MaskedDiscreteAttackerWrapper:
    """action_masking.py:30-165: one Discrete(N*N*P*C + N*L + N*N*R) action space, order [connect, local, remote]."""

    def __init__(self, env: AttackerEnvWrapper):
        self.env = env
        obs_spaces = env.observation_space.spaces
        for k in ("local_vulnerability", "remote_vulnerability", "connect"):
            if k not in obs_spaces:
                raise KeyError("Observation must include top-level 'local_vulnerability', 'remote_vulnerability', and 'connect' masks")
        n_l = tuple(int(x) for x in obs_spaces["local_vulnerability"].n)
        n_r = tuple(int(x) for x in obs_spaces["remote_vulnerability"].n)
        n_c = tuple(int(x) for x in obs_spaces["connect"].n)
        self._n, self._l, self._r, self._p, self._c = n_l[0], n_l[1], n_r[2], n_c[2], n_c[3]
        self.connect_size = self._n * self._n * self._p * self._c
        self.local_size = self._n * self._l
        self.remote_size = self._n * self._n * self._r
        self.total = self.connect_size + self.local_size + self.remote_size
        self.action_space = spaces.Discrete(self.total)
        self.observation_space = env.observation_space
        self._kind_to_index = {kind: idx for idx, (kind, _, _) in env.action_subspaces.items()}
        self._kind_to_slice = {kind: (a, b) for _, (kind, a, b) in env.action_subspaces.items()}

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env

    def action_masks(self) -> np.ndarray:
        obs = getattr(self.env, "_last_transformed_observation", None)
        if obs is None:
            return np.ones((self.total,), dtype=np.bool_)
        return np.concatenate([np.asarray(obs["connect"], dtype=np.int8).reshape(-1),
                               np.asarray(obs["local_vulnerability"], dtype=np.int8).reshape(-1),
                               np.asarray(obs["remote_vulnerability"], dtype=np.int8).reshape(-1)]).astype(np.bool_)

    def _decode(self, action: int):
        a = int(action)
        if a < 0 or a >= self.total:
            raise ValueError(f"Invalid discrete action: {a}")
        if a < self.connect_size:
            a, cred = divmod(a, self._c)
            a, port = divmod(a, self._p)
            src, tgt = divmod(a, self._n)
            return "connect", (src, tgt, port, cred)
        if a < self.connect_size + self.local_size:
            src, vuln = divmod(a - self.connect_size, self._l)
            return "local_vulnerability", (src, vuln)
        a, vuln = divmod(a - self.connect_size - self.local_size, self._r)
        src, tgt = divmod(a, self._n)
        return "remote_vulnerability", (src, tgt, vuln)

    def _encode_for_inner_env(self, kind: str, coords) -> np.ndarray:
        nvec = np.asarray(self.env.action_space.nvec, dtype=np.int64)
        enc = np.zeros((len(nvec),), dtype=np.int64)
        enc[0] = self._kind_to_index[kind]
        a, b = self._kind_to_slice[kind]
        enc[a:b] = np.asarray(coords, dtype=np.int64)
        return enc

    def step(self, action):
        if isinstance(action, np.ndarray):
            action = int(action.item())
        kind, coords = self._decode(int(action))
        return self.env.step(self._encode_for_inner_env(kind, coords))

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)
