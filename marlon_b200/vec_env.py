"""stable-baselines3 ``VecEnv`` adapter over ``MultiAgentUniversalEnv``.

The reference only ever composes SB3's own ``DummyVecEnv`` / ``VecMonitor`` around one wrapper
(train_marl_multi.py:181-183, baseline_marlon_agent.py:34,47); this adapter presents the same contract for ``n_envs``
batched environments: ``reset()``, ``step_async(actions)`` / ``step_wait() -> (obs, rewards, dones, infos)`` with
auto-reset, ``infos[i]["terminal_observation"]``, ``infos[i]["TimeLimit.truncated"]`` and the Monitor / VecMonitor
``infos[i]["episode"] = {"r", "l", "t"}`` entry that ``ep_info_buffer`` consumers read (baseline_marlon_agent.py:179-188).

When stable-baselines3 is importable ``BatchedVecEnv`` IS a ``stable_baselines3.common.vec_env.VecEnv`` (SB3's
``BaseAlgorithm._wrap_env`` tests ``isinstance(env, VecEnv)`` and would otherwise wrap the object in a ``DummyVecEnv``);
without SB3 the same class stands on its own.

Two adapters share one universe: ``universe.attacker_vec_env`` steps the attacker half of the pair step and
``universe.defender_vec_env`` the defender half, in the reference's order (marl_algorithm.py:43-49).

Per step the host sees ONE device-to-host copy: rewards and done flags of both agents (12 bytes per env,
``cbx_batch_fetch_host``) -- plus, with ``observations="numpy"`` (the SB3 contract for host-side policies), the stacked
observation arrays in the same call.  ``observations="torch"`` (default) returns the device tensors themselves (zero-copy
views) for policies that live on the GPU.
"""
from __future__ import annotations

import time
import warnings
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _abi

try:  # pragma: no cover - depends on the installation
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv as _SB3VecEnv

    HAVE_SB3 = True
except Exception:  # noqa: BLE001
    _SB3VecEnv = object
    HAVE_SB3 = False

_EMPTY: Dict[str, Any] = {}
_ATT_SCALARS = ["newly_discovered_nodes_count", "lateral_move", "customer_data_found", "probe_result", "escalation",
                "credential_cache_length", "discovered_node_count"]
_ATT_FIELDS = {"leaked_credentials": "leaked_credentials", "credential_cache_matrix": "credential_cache_matrix",
               "discovered_nodes_properties": "discovered_nodes_properties", "nodes_privilegelevel": "nodes_privilegelevel",
               "local_vulnerability": "local_vulnerability", "remote_vulnerability": "remote_vulnerability", "connect": "connect",
               "owned_bits": "owned_bits"}
_DEF_FIELDS = {"infected_nodes": "def_infected_nodes", "incoming_firewall_status": "def_incoming_firewall",
               "outgoing_firewall_status": "def_outgoing_firewall", "services_status": "def_services_status"}


class _LazyRow(dict):
    """``infos[i]["terminal_observation"]``: row `j` of the gathered terminal observations, materialised on first use.
    A dict to every consumer (SB3's ``obs_to_tensor`` deep-copies it and iterates ``items()``); building thousands of 14-key
    dicts per step eagerly would cost more host time than the whole batch's step."""

    __slots__ = ("_src", "_j")

    def __init__(self, src, j):
        super().__init__()
        self._src, self._j = src, j

    def _fill(self):
        src = self._src
        if src is not None:
            self._src = None
            dict.update(self, {k: v[self._j] for k, v in src.items()})
        return self

    def __getitem__(self, k):
        return dict.__getitem__(self._fill(), k)

    def __iter__(self):
        return dict.__iter__(self._fill())

    def __len__(self):
        return dict.__len__(self._fill())

    def __contains__(self, k):
        return dict.__contains__(self._fill(), k)

    def keys(self):
        return dict.keys(self._fill())

    def values(self):
        return dict.values(self._fill())

    def items(self):
        return dict.items(self._fill())

    def get(self, k, default=None):
        return dict.get(self._fill(), k, default)

    def copy(self):
        return dict(self._fill())

    def __eq__(self, other):
        return dict.__eq__(self._fill(), other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return dict.__repr__(self._fill())

    def __deepcopy__(self, memo):
        import copy

        return {k: copy.deepcopy(v, memo) for k, v in self._fill().items()}

    def __reduce__(self):
        return (dict, (dict(self._fill()),))


class BatchedVecEnv(_SB3VecEnv):
    def __init__(self, universe, role: str = "attacker", observations: str = "torch", terminal_observations: str = "all"):
        """`observations`: "torch" (device tensor views) or "numpy" (host arrays, one packed copy per step).
        `terminal_observations`: "all" (SB3 contract: every finished env's info carries it), "truncated" (only where SB3's
        on-policy bootstrap reads it: ``TimeLimit.truncated`` episodes) or "none"."""
        if role not in ("attacker", "defender"):
            raise ValueError("role must be 'attacker' or 'defender'")
        if role == "defender" and not universe.has_defender:
            raise ValueError("this universe was built without a defender")
        if observations not in ("torch", "numpy") or terminal_observations not in ("all", "truncated", "none"):
            raise ValueError("observations: 'torch' | 'numpy'; terminal_observations: 'all' | 'truncated' | 'none'")
        self.universe, self.role, self.observations = universe, role, observations
        obs_space = universe.attacker_observation_space if role == "attacker" else universe.defender_observation_space
        act_space = universe.attacker_action_space if role == "attacker" else universe.defender_action_space
        if HAVE_SB3:
            _SB3VecEnv.__init__(self, universe.n_envs, obs_space, act_space)
        else:
            self.num_envs, self.observation_space, self.action_space = universe.n_envs, obs_space, act_space
            self.render_mode = None
            self.reset_infos: List[Dict[str, Any]] = [{} for _ in range(universe.n_envs)]
        self.metadata = {"render_modes": []}
        self._actions = None
        self._t0 = time.time()
        n = self.num_envs
        self._ep_ret = np.zeros(n, dtype=np.float64)
        self._ep_len = np.zeros(n, dtype=np.int64)
        if terminal_observations != "none" and not universe.cfg.emit_terminal_obs:
            warnings.warn("the universe was built with emit_terminal_obs=False: infos of finished envs will carry no "
                          "'terminal_observation' (build it with emit_terminal_obs=True for the full SB3 contract)", stacklevel=2)
            terminal_observations = "none"
        self.terminal_observations = terminal_observations
        dense = universe.cfg.mask_mode == _abi.MASK_DENSE
        if role == "attacker":
            self._fields = _abi.F_RESULTS | (_abi.F_OBS_DENSE if dense else _abi.F_OBS_FACTORED) & ~(0xF << 9)
        else:
            self._fields = _abi.F_RESULTS | (0xF << 9)

    # ---- helpers -------------------------------------------------------------------------------------------------------
    def _obs_from_host(self, h: Dict[str, np.ndarray]):
        if self.role == "attacker":
            sc = h["scalars"]
            d = {k: sc[:, i] for i, k in enumerate(_ATT_SCALARS)}
            d.update({k: h[v] for k, v in _ATT_FIELDS.items() if v in h})
            if "connect" in d:
                d.pop("owned_bits", None)
            return d
        return {k: h[v] for k, v in _DEF_FIELDS.items()}

    def _obs(self, terminal: bool = False, rows=None):
        """Stacked observation: device tensors ("torch") or host arrays ("numpy"); `rows` (index tensor) gathers on the device
        first so that only those envs cross PCIe."""
        u = self.universe
        d = u.attacker_observation(terminal) if self.role == "attacker" else u.defender_observation(terminal)
        if rows is not None:
            d = {k: v[rows] for k, v in d.items()}
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
        if self.observations == "numpy":
            return self._obs_from_host(self.universe.batch.fetch_host(self._fields & ~_abi.F_RESULTS))
        return self._obs()

    def step_async(self, actions) -> None:
        self._actions = actions

    def step_wait(self):
        u, b = self.universe, self.universe.batch
        if self.role == "attacker":
            u.step_attacker(self._actions)
        else:
            u.step_defender(self._actions)
        # the step's only device-to-host traffic: one call, one synchronisation
        h = b.fetch_host(self._fields if self.observations == "numpy" else _abi.F_RESULTS)
        p = "att_" if self.role == "attacker" else "def_"
        rew = h[p + "reward"].copy()
        term, trunc = h[p + "terminated"].astype(bool), h[p + "truncated"].astype(bool)
        dones = term | trunc
        self._ep_ret += rew
        self._ep_len += 1
        infos: List[Dict[str, Any]] = [_EMPTY] * self.num_envs
        idx = np.flatnonzero(dones)
        if idx.size:
            now = round(time.time() - self._t0, 6)
            tl = (trunc & ~term)[idx].tolist()
            rs, ls = np.round(self._ep_ret[idx], 6).tolist(), self._ep_len[idx].tolist()
            made = [{"TimeLimit.truncated": t, "episode": {"r": r, "l": ln, "t": now}} for t, r, ln in zip(tl, rs, ls)]
            if self.terminal_observations != "none":
                want = idx if self.terminal_observations == "all" else idx[np.asarray(tl, dtype=bool)]
                if want.size:
                    import torch

                    rows = self._obs(terminal=True, rows=torch.as_tensor(want, device=b.torch_device))
                    # rows are sliced when somebody reads them (_LazyRow): thousands of envs finish per step at bench sizes and
                    # one index call per env and key would cost more host time than the step itself
                    pos = {int(e): j for j, e in enumerate(want.tolist())}
                    for info, e in zip(made, idx.tolist()):
                        j = pos.get(e)
                        if j is not None:
                            info["terminal_observation"] = _LazyRow(rows, j)
            for info, e in zip(made, idx.tolist()):
                infos[e] = info
            self._ep_ret[idx] = 0
            self._ep_len[idx] = 0
        obs = self._obs_from_host(h) if self.observations == "numpy" else self._obs()
        return obs, rew, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        pass

    def seed(self, seed: Optional[int] = None) -> Sequence[Optional[int]]:
        return [seed] * self.num_envs

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        n = len(self._indices(indices))
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
