"""stable-baselines3 ``VecEnv`` adapter over ``MultiAgentUniversalEnv``.

The reference only ever composes SB3's own ``DummyVecEnv`` / ``VecMonitor`` around one wrapper
(train_marl_multi.py:181-183, baseline_marlon_agent.py:34,47); this adapter presents the same contract for ``n_envs``
batched environments: ``reset()``, ``step_async(actions)`` / ``step_wait() -> (obs, rewards, dones, infos)`` with
auto-reset, ``infos[i]["terminal_observation"]``, ``infos[i]["TimeLimit.truncated"]`` and the Monitor / VecMonitor
``infos[i]["episode"] = {"r", "l", "t"}`` entry that ``ep_info_buffer`` consumers read (baseline_marlon_agent.py:179-188).

Two adapters share one universe: ``universe.attacker_vec_env`` steps the attacker half of the pair step and
``universe.defender_vec_env`` the defender half, in the reference's order (marl_algorithm.py:43-49).

``observations="numpy"`` (default, SB3 contract) copies the stacked observations to host arrays each step;
``observations="torch"`` returns the device tensors themselves (zero-copy views) for policies that live on the GPU.
"""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

_EMPTY: Dict[str, Any] = {}


class BatchedVecEnv:
    def __init__(self, universe, role: str = "attacker", observations: str = "numpy"):
        if role not in ("attacker", "defender"):
            raise ValueError("role must be 'attacker' or 'defender'")
        if role == "defender" and not universe.has_defender:
            raise ValueError("this universe was built without a defender")
        self.universe, self.role, self.observations = universe, role, observations
        self.num_envs = universe.n_envs
        self.observation_space = universe.attacker_observation_space if role == "attacker" else universe.defender_observation_space
        self.action_space = universe.attacker_action_space if role == "attacker" else universe.defender_action_space
        self.metadata = {"render_modes": []}
        self.render_mode = None
        self._actions = None
        self._t0 = time.time()
        n = self.num_envs
        self._ep_ret = np.zeros(n, dtype=np.float64)
        self._ep_len = np.zeros(n, dtype=np.int64)
        if not universe.cfg.emit_terminal_obs:
            self._warned = False

    # ---- helpers -------------------------------------------------------------------------------------------------------
    def _obs(self, terminal: bool = False):
        u = self.universe
        d = u.attacker_observation(terminal) if self.role == "attacker" else u.defender_observation(terminal)
        if self.observations == "torch":
            return d
        return {k: v.cpu().numpy() for k, v in d.items()}

    def _who(self) -> int:
        return 1 if self.role == "attacker" else 2

    # ---- VecEnv API -----------------------------------------------------------------------------------------------------
    def reset(self):
        self.universe.batch.reset(who=self._who())
        self._ep_ret[:] = 0
        self._ep_len[:] = 0
        return self._obs()

    def step_async(self, actions) -> None:
        self._actions = actions

    def step_wait(self):
        u, b = self.universe, self.universe.batch
        if self.role == "attacker":
            u.step_attacker(self._actions)
            rew, term, trunc = b.numpy("att_reward"), b.numpy("att_terminated"), b.numpy("att_truncated")
        else:
            u.step_defender(self._actions)
            rew, term, trunc = b.numpy("def_reward"), b.numpy("def_terminated"), b.numpy("def_truncated")
        term, trunc = term.astype(bool), trunc.astype(bool)
        dones = term | trunc
        self._ep_ret += rew
        self._ep_len += 1
        infos: List[Dict[str, Any]] = [_EMPTY] * self.num_envs
        if dones.any():
            idx = np.nonzero(dones)[0]
            term_obs = self._obs(terminal=True) if u.cfg.emit_terminal_obs else None
            now = round(time.time() - self._t0, 6)
            for i in idx:
                info: Dict[str, Any] = {"TimeLimit.truncated": bool(trunc[i] and not term[i]),
                                        "episode": {"r": round(float(self._ep_ret[i]), 6), "l": int(self._ep_len[i]), "t": now}}
                if term_obs is not None:
                    info["terminal_observation"] = {k: v[i] for k, v in term_obs.items()}
                infos[i] = info
            self._ep_ret[idx] = 0
            self._ep_len[idx] = 0
        return self._obs(), rew.copy(), dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        pass

    def seed(self, seed: Optional[int] = None) -> Sequence[Optional[int]]:
        return [seed] * self.num_envs

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        n = self.num_envs if indices is None else len(self._indices(indices))
        return [getattr(self, attr_name, getattr(self.universe, attr_name, None))] * n

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        setattr(self, attr_name, value)

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> List[Any]:
        if method_name == "action_masks":
            m = self.universe.action_masks().cpu().numpy()
            return [m[i] for i in self._indices(indices)]
        raise AttributeError(f"env_method {method_name!r} is not available on the batched environment")

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False] * len(self._indices(indices))

    def _indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        if isinstance(indices, int):
            return [indices]
        return list(indices)

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode: Optional[str] = None):
        return None

    @property
    def unwrapped(self):
        return self
