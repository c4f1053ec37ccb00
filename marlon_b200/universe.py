"""Batched multi-agent environment: the drop-in for the hot path behind ``MultiAgentUniverse``.

``MultiAgentUniversalEnv`` owns ``n_envs`` attacker-vs-defender environments in HBM and advances all of them with
one fused kernel launch per step pair (reference loop: ``marl_algorithm.collect_rollouts``, marl_algorithm.py:43-49:
attacker ``perform_step`` then defender ``perform_step``, each followed by SB3's ``DummyVecEnv`` auto-reset).
Constructor kwargs are those of ``MultiAgentUniverse.build`` (multiagent_universe.py:78-95) that concern the
environment; the agent builders / SB3 training loop stay outside (they are the caller).

``MultiAgentUniverse.build`` is kept with the reference signature: it wires single-env ``AttackerEnvWrapper`` /
``DefenderEnvWrapper`` objects (``wrappers.py``) and hands them to the caller's ``AgentBuilder``s.

Multi-GPU: env instances shard by contiguous index range, one process per GPU (``MultiAgentUniversalEnv.sharded``),
with no data-path collective; ``episode_statistics(reduce=True)`` all-reduces the 16-slot statistics vector over
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) -- the only collective of the path.
"""
from __future__ import annotations

import logging
import os
from typing import Any, Dict, Optional, Tuple

import numpy as np

from . import _abi, config, registry, scenario, spaces
from .config import DefenderConstraint

ATTACKER_OBS_KEYS = ["newly_discovered_nodes_count", "lateral_move", "customer_data_found", "probe_result", "escalation",
                     "credential_cache_length", "discovered_node_count"]


def attacker_spaces(cfg: _abi.Config, comp) -> Tuple[spaces.Dict, spaces.MultiDiscrete]:
    """AttackerEnvWrapper observation / action spaces (attack_wrapper.py:164-227)."""
    ident = comp.identifiers
    N, C, LEAK = cfg.maximum_node_count, cfg.maximum_total_credentials, cfg.maximum_discoverable_credentials_per_action
    P, L, R, props = len(ident.ports), len(ident.local_vulnerabilities), len(ident.remote_vulnerabilities), len(ident.properties)
    obs = spaces.Dict({
        "newly_discovered_nodes_count": spaces.Discrete(1 + N), "lateral_move": spaces.Discrete(2),
        "customer_data_found": spaces.Discrete(2), "probe_result": spaces.Discrete(3), "escalation": spaces.Discrete(4),
        "credential_cache_length": spaces.Discrete(C), "discovered_node_count": spaces.Discrete(N),
        "leaked_credentials": spaces.MultiDiscrete(np.tile(np.array([2, C, N, P], dtype=np.int32), LEAK)),
        "credential_cache_matrix": spaces.MultiDiscrete(np.tile(np.array([N, P], dtype=np.int32), C)),
        "discovered_nodes_properties": spaces.MultiDiscrete(np.full(N * props, 3, dtype=np.int32)),
        "nodes_privilegelevel": spaces.MultiDiscrete([4] * N),
        "local_vulnerability": spaces.MultiBinary(np.array([N, L])),
        "remote_vulnerability": spaces.MultiBinary(np.array([N, N, R])),
        "connect": spaces.MultiBinary(np.array([N, N, P, C], dtype=np.int32)),
    })
    dims = {_abi.KIND_LOCAL: [N, L], _abi.KIND_REMOTE: [N, N, R], _abi.KIND_CONNECT: [N, N, P, C]}
    nvec = [3]
    for i in range(3):
        nvec += dims[cfg.kind_of_index[i]]
    return obs, spaces.MultiDiscrete(nvec)


def defender_spaces(comp, n: Optional[int] = None, n_services: Optional[int] = None) -> Tuple[spaces.Dict, spaces.MultiDiscrete]:
    """DefenderEnvWrapper observation / action spaces (defend_wrapper.py:162-195); `n` / `n_services` override the scenario's
    own counts with the padded ones of a multi-scenario batch."""
    n = comp.n_nodes if n is None else int(n)
    n_services = comp.n_services if n_services is None else int(n_services)
    obs = spaces.Dict({"infected_nodes": spaces.MultiBinary(n), "incoming_firewall_status": spaces.MultiBinary(6 * n),
                       "outgoing_firewall_status": spaces.MultiBinary(6 * n), "services_status": spaces.MultiBinary(n_services)})
    return obs, spaces.MultiDiscrete([5, n, n, 6, 2, n, 6, 2, n, 3, n, 3])


class MultiAgentUniversalEnv:
    """n_envs attacker(+defender) environments stepped in lock-step on one GPU."""

    def __init__(self, env_id: str = "CyberBattleToyCtf-v0", n_envs: int = 1, *, device: int = 0, defender: bool = True,
                 attacker_invalid_action_reward_modifier: float = -1.0, attacker_invalid_action_reward_multiplier: float = 1.0,
                 defender_invalid_action_reward_modifier: float = -1, max_timesteps: int = 2000,
                 maximum_node_count: Optional[int] = None, maximum_total_credentials: Optional[int] = None,
                 maximum_discoverable_credentials_per_action: Optional[int] = None, throws_on_invalid_actions: bool = False,
                 attacker_loss_reward: float = -5000.0, defender_loss_reward: float = -5000.0, defender_maintain_sla: float = 0.60,
                 defender_reset_on_constraint_broken: bool = True, defender_binding: str = "reference_stale",
                 action_kind_order="gymnasium029", mask_mode: str = "dense", emit_terminal_obs: bool = False, seed: int = 0,
                 env_index_base: int = 0, scenario_seeds=None, **env_kwargs):
        """`scenario_seeds` (with ``env_id="CyberBattleRandom-v0"``): one generated network per seed, all in one batch
        (configs[4]); `n_envs` is then the env count per network (an int, or one count per seed; every count but the last a
        multiple of 32).  Node / credential / service counts differ between networks: the observation spaces are those of the
        bounds and of the largest network, smaller ones are zero-padded."""
        if defender_binding not in ("reference_stale", "live"):
            raise ValueError("defender_binding: 'reference_stale' (the reference as executed: the LearningDefender acts on a dead copy "
                             "of the environment, SURVEY.md B.1; default) or 'live' (its binding refreshed at every reset)")
        if defender_binding == "live" and scenario_seeds is not None:
            raise NotImplementedError("the live defender binding is single-scenario")
        del attacker_invalid_action_reward_multiplier, attacker_loss_reward  # stored but unused by the reference (attack_wrapper.py:51-52)
        kw = dict(env_kwargs)
        for k, v in (("maximum_node_count", maximum_node_count), ("maximum_total_credentials", maximum_total_credentials),
                     ("maximum_discoverable_credentials_per_action", maximum_discoverable_credentials_per_action)):
            if v is not None:
                kw[k] = int(v)
        kw["throws_on_invalid_actions"] = bool(throws_on_invalid_actions)
        if defender:  # multiagent_universe.py:158-165
            kw.setdefault("defender_constraint", DefenderConstraint(maintain_sla=defender_maintain_sla))
            kw.setdefault("losing_reward", defender_loss_reward)
        if scenario_seeds is not None:
            seeds = [int(x) for x in scenario_seeds]
            resolved = [registry.resolve(env_id, seed=sd, **kw) for sd in seeds]
            envs, merged = [r[0] for r in resolved], resolved[0][1]
            comps = [scenario.compile_scenario(e) for e in envs]
            counts = [int(n_envs)] * len(seeds) if np.isscalar(n_envs) else [int(x) for x in n_envs]
            n_envs = sum(counts)
            env = envs[0]
        else:
            env, merged = registry.resolve(env_id, **kw)
            comps, counts = None, None
        merged.pop("observation_padding", None)
        self.env_id, self.environment = env_id, env
        self.compiled = comps[0] if comps else scenario.compile_scenario(env)
        self.cfg = config.make_config(
            _abi.MODE_MARLON, attacker_max_timesteps=max_timesteps,
            attacker_invalid_action_reward_modifier=attacker_invalid_action_reward_modifier,
            action_kind_order=action_kind_order, defender_enabled=defender, defender_max_timesteps=max_timesteps,
            defender_invalid_action_reward=defender_invalid_action_reward_modifier,
            defender_reset_on_constraint_broken=defender_reset_on_constraint_broken, defender_loss_reward=defender_loss_reward,
            defender_binding=defender_binding, auto_reset=True, mask_mode=_abi.MASK_DENSE if mask_mode == "dense" else _abi.MASK_FACTORED,
            emit_terminal_obs=emit_terminal_obs, seed=seed, env_index_base=env_index_base, **merged)
        from .batch import Batch

        self.n_envs, self.device, self.has_defender = int(n_envs), device, bool(defender)
        self.batch = Batch(comps, self.cfg, counts, device=device) if comps else Batch(self.compiled, self.cfg, self.n_envs, device=device)
        self.attacker_observation_space, self.attacker_action_space = attacker_spaces(self.cfg, self.compiled)
        self.defender_observation_space, self.defender_action_space = defender_spaces(
            self.compiled, n=self.batch.views.n_nodes, n_services=self.batch.views.n_services)
        self.identifiers = env.identifiers
        self.logger = logging.getLogger("marlon")

    # ---- sharding ------------------------------------------------------------------------------------------------
    @classmethod
    def sharded(cls, env_id: str, total_envs: int, **kwargs) -> "MultiAgentUniversalEnv":
        """This rank's contiguous slice of `total_envs` (RANK / WORLD_SIZE / LOCAL_RANK from the environment)."""
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        lo, hi = shard_range(total_envs, rank, world)
        kwargs.setdefault("device", int(os.environ.get("LOCAL_RANK", "0")))
        return cls(env_id, hi - lo, env_index_base=lo, **kwargs)

    # ---- stepping -----------------------------------------------------------------------------------------------------
    def reset(self):
        self.batch.reset()
        return self.attacker_observation(), (self.defender_observation() if self.has_defender else None)

    def step(self, attacker_actions, defender_actions=None):
        """One attacker+defender step pair for every env; everything returned is a device tensor view."""
        self.batch.step(attacker_actions, defender_actions)
        return self.results()

    def step_attacker(self, attacker_actions):
        self.batch.step(attacker_actions, None, who=self.batch.WHO_ATTACKER)

    def step_defender(self, defender_actions):
        self.batch.step(None, defender_actions, who=self.batch.WHO_DEFENDER)

    def results(self) -> Dict[str, Any]:
        t = self.batch.tensors
        out = {"attacker_observation": self.attacker_observation(), "attacker_reward": t["att_reward"],
               "attacker_terminated": t["att_terminated"], "attacker_truncated": t["att_truncated"], "attacker_info": t["att_info"]}
        if self.has_defender:
            out.update({"defender_observation": self.defender_observation(), "defender_reward": t["def_reward"],
                        "defender_terminated": t["def_terminated"], "defender_truncated": t["def_truncated"]})
        return out

    def attacker_observation(self, terminal: bool = False) -> Dict[str, Any]:
        """AttackerEnvWrapper.transform_observation's dict (attack_wrapper.py:474-522), batched: [n_envs, ...] tensors."""
        t, p = self.batch.tensors, ("term_" if terminal else "")
        sc = t[p + "scalars"]
        obs = {k: sc[:, i] for i, k in enumerate(ATTACKER_OBS_KEYS)}
        obs.update({"leaked_credentials": t[p + "leaked_credentials"], "credential_cache_matrix": t[p + "credential_cache_matrix"],
                    "discovered_nodes_properties": t[p + "discovered_nodes_properties"],
                    "nodes_privilegelevel": t[p + "nodes_privilegelevel"]})
        if self.cfg.mask_mode == _abi.MASK_DENSE:
            obs.update({"local_vulnerability": t[p + "local_vulnerability"], "remote_vulnerability": t[p + "remote_vulnerability"],
                        "connect": t[p + "connect"]})
        elif not terminal:
            obs["owned_bits"] = t["owned_bits"]  # factored masks: owned bitset + the two counts above
        return obs

    def defender_observation(self, terminal: bool = False) -> Dict[str, Any]:
        t = self.batch.tensors
        return {"infected_nodes": t["term_def_infected_nodes" if terminal else "def_infected_nodes"],
                "incoming_firewall_status": t["def_incoming_firewall"], "outgoing_firewall_status": t["def_outgoing_firewall"],
                "services_status": t["def_services_status"]}

    def action_masks(self):
        """MaskedDiscreteAttackerWrapper.action_masks (action_masking.py:90-105), batched: bool [n_envs, N*N*P*C + N*L + N*N*R]
        in the order connect, local, remote."""
        import torch

        t, n = self.batch.tensors, self.n_envs
        if self.cfg.mask_mode != _abi.MASK_DENSE:
            raise RuntimeError("dense masks are not materialised in factored mode")
        return torch.cat([t["connect"].reshape(n, -1), t["local_vulnerability"].reshape(n, -1),
                          t["remote_vulnerability"].reshape(n, -1)], dim=1).to(torch.bool)

    def sample_actions(self, seed: int = 0):
        return self.batch.sample_actions(seed)

    # ---- statistics ----------------------------------------------------------------------------------------------------------
    def episode_statistics(self, reduce: bool = True, reset: bool = False) -> Dict[str, float]:
        """Per-rollout episode statistics (what SB3's ep_info_buffer / MARLon's EvalutionStats report, evaluation_stats.py:7-58).
        With reduce=True and an initialised torch.distributed group the 16 partial sums are all-reduced (SUM)."""
        import torch

        v = self.batch.stats_tensor.clone()
        if reduce:
            v = all_reduce_stats(v)
        if reset:
            self.batch.stats_reset()
        return summarize_stats(v.cpu().numpy())

    def vec_env(self, role: str = "attacker", observations: str = "torch", terminal_observations: str = "all"):
        """The SB3 ``VecEnv`` adapter of one agent (``vec_env.BatchedVecEnv``).  One adapter per (role, settings) is kept: it
        carries the Monitor-style episode return / length accumulators, which a fresh object on every access would lose."""
        from .vec_env import BatchedVecEnv

        key = (role, observations, terminal_observations)
        cache = self.__dict__.setdefault("_vec_envs", {})
        if key not in cache:
            cache[key] = BatchedVecEnv(self, role, observations=observations, terminal_observations=terminal_observations)
        return cache[key]

    @property
    def attacker_vec_env(self):
        return self.vec_env("attacker")

    @property
    def defender_vec_env(self):
        return self.vec_env("defender")

    def close(self):
        self.batch.close()


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous env-index range of `rank`: sizes differ by at most one, earlier ranks take the extras."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_stats(v):
    """SUM all-reduce of the statistics vector when torch.distributed is initialised (NCCL for CUDA tensors, gloo for CPU)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return v


def summarize_stats(s: np.ndarray) -> Dict[str, float]:
    d = {k: float(x) for k, x in zip(_abi.STAT_NAMES, s)}
    ep = max(d["episodes"], 1.0)

    def mean_std(sx, sxx, n):
        m = sx / n
        return m, float(np.sqrt(max(sxx / n - m * m, 0.0)))

    d["attacker_return_mean"], d["attacker_return_std"] = mean_std(d["att_return"], d["att_return_sq"], ep)
    d["episode_length_mean"], d["episode_length_std"] = mean_std(d["ep_len"], d["ep_len_sq"], ep)
    d["defender_return_mean"], d["defender_return_std"] = mean_std(d["def_return"], d["def_return_sq"], ep)
    return d


class AgentBuilder:
    """multiagent_universe.py:50-69"""

    def build(self, wrapper, logger):
        raise NotImplementedError


class MultiAgentUniverse:
    """Same factory signature as the reference (multiagent_universe.py:78-206).  The environment side is this package's;
    the agents are whatever the caller's builders return (SB3 models in the reference)."""

    @classmethod
    def build(cls, attacker_builder: AgentBuilder, attacker_invalid_action_reward_modifier: float = -1.0,
              attacker_invalid_action_reward_multiplier: float = 1.0, defender_builder: Optional[AgentBuilder] = None,
              defender_invalid_action_reward_modifier=-1, env_id: str = "CyberBattleToyCtf-v0", max_timesteps: int = 2000,
              maximum_node_count: Optional[int] = None, maximum_total_credentials: Optional[int] = None,
              maximum_discoverable_credentials_per_action: Optional[int] = None, observation_padding: Optional[bool] = None,
              throws_on_invalid_actions: Optional[bool] = False, attacker_action_masking: bool = False,
              attacker_loss_reward: float = -5000.0, defender_loss_reward: float = -5000.0, defender_maintain_sla: float = 0.60,
              defender_reset_on_constraint_broken: bool = True, device: int = 0):
        from . import cyberbattle_env as cbe
        from .wrappers import AttackerEnvWrapper, DefenderEnvWrapper, EnvironmentEventSource, MaskedDiscreteAttackerWrapper

        logger = logging.Logger("marlon", level=os.environ.get("LOGLEVEL", "INFO").upper())
        logger.addHandler(logging.StreamHandler())
        env_kwargs: Dict[str, Any] = {"device": device}
        for k, v in (("maximum_node_count", maximum_node_count), ("maximum_total_credentials", maximum_total_credentials),
                     ("maximum_discoverable_credentials_per_action", maximum_discoverable_credentials_per_action)):
            if v is not None:
                env_kwargs[k] = int(v)
        if observation_padding is not None:
            env_kwargs["observation_padding"] = bool(observation_padding)
        if throws_on_invalid_actions is not None:
            env_kwargs["throws_on_invalid_actions"] = bool(throws_on_invalid_actions)
        if defender_builder:
            cyber_env = cbe.make(env_id, defender_constraint=DefenderConstraint(maintain_sla=defender_maintain_sla),
                                 losing_reward=defender_loss_reward, **env_kwargs)
        else:
            cyber_env = cbe.make(env_id, **env_kwargs)
        event_source = EnvironmentEventSource()
        attacker_wrapper = AttackerEnvWrapper(cyber_env=cyber_env, event_source=event_source, max_timesteps=max_timesteps,
                                              invalid_action_reward_modifier=attacker_invalid_action_reward_modifier,
                                              invalid_action_reward_multiplier=attacker_invalid_action_reward_multiplier,
                                              loss_reward=attacker_loss_reward, log_episode_end=True)
        if attacker_action_masking:
            attacker_wrapper = MaskedDiscreteAttackerWrapper(attacker_wrapper)
        defender_wrapper = None
        if defender_builder:
            defender_wrapper = DefenderEnvWrapper(cyber_env=cyber_env, event_source=event_source, attacker_reward_store=attacker_wrapper,
                                                  max_timesteps=max_timesteps, invalid_action_reward=defender_invalid_action_reward_modifier,
                                                  defender=True, reset_on_constraint_broken=defender_reset_on_constraint_broken,
                                                  loss_reward=defender_loss_reward, log_episode_end=True)
        attacker_agent = attacker_builder.build(attacker_wrapper, logger)
        defender_agent = defender_builder.build(defender_wrapper, logger) if defender_builder else None
        return cls(attacker_agent=attacker_agent, defender_agent=defender_agent, max_timesteps=max_timesteps, logger=logger)

    def __init__(self, attacker_agent, defender_agent, max_timesteps: int, logger: logging.Logger):
        self.attacker_agent, self.defender_agent = attacker_agent, defender_agent
        self.max_timesteps, self.logger = max_timesteps, logger

    def run_episode(self, max_steps: Optional[int] = None):
        """marl_algorithm.run_episode (marl_algorithm.py:144-251) for agents exposing `.env`, `.wrapper`, `.predict`."""
        max_steps = self.max_timesteps if max_steps is None else max_steps
        a, d = self.attacker_agent, self.defender_agent
        obs1 = a.env.reset()
        obs1 = obs1[0] if isinstance(obs1, tuple) else obs1
        obs2 = None
        if d:
            d.wrapper.on_reset(0)
            obs2 = d.env.reset()
            obs2 = obs2[0] if isinstance(obs2, tuple) else obs2
        ar, dr = [], []
        for _ in range(max_steps):
            obs1, r1, t1, tr1, _ = a.env.step(a.predict(observation=obs1))
            ar.append(r1)
            done = t1 or tr1
            if d:
                obs2, r2, t2, tr2, _ = d.env.step(d.predict(observation=obs2))
                dr.append(r2)
                done = done or t2 or tr2
            if done:
                break
        return ar, dr
