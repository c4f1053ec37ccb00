"""``CyberBattleEnv``: the gym interface of the reference (``_env/cyberbattle_env.py:339-1233``) over a 1-env batch.

Same constructor kwargs, action / observation spaces, ``reset`` / ``step`` return shapes, error behaviour and helper
methods (``compute_action_mask``, ``apply_mask``, ``is_node_owned``, ``is_action_valid``, ``sample_valid_action`` ...).
The simulation itself runs in the fused CUDA step kernel; this class only translates between gym-style Python objects
and the device tensors.  For throughput use ``marlon_b200.universe.MultiAgentUniversalEnv`` / ``BatchedVecEnv``: this
class is the legacy single-env view (one kernel launch + one host sync per ``step``).
"""
from __future__ import annotations

from typing import Any, Dict, List, NamedTuple, Optional, Tuple

import numpy as np

from . import _abi, config, model, registry, scenario, spaces
from .config import AttackerGoal, DefenderConstraint, DefenderGoal, DefenderAgent  # noqa: F401

NA = 1
UNUSED_SLOT = np.int32(0)
USED_SLOT = np.int32(1)
KIND_CODE = {"local_vulnerability": _abi.KIND_LOCAL, "remote_vulnerability": _abi.KIND_REMOTE, "connect": _abi.KIND_CONNECT}


class OutOfBoundIndexError(Exception):
    """The agent attempted to reference an entity (node or a vulnerability) with an invalid index"""


class EnvironmentBounds(NamedTuple):
    """cyberbattle_env.py:172-224"""
    maximum_total_credentials: np.int32
    maximum_node_count: np.int32
    maximum_discoverable_credentials_per_action: np.int32
    port_count: np.int32
    property_count: np.int32
    local_attacks_count: np.int32
    remote_attacks_count: np.int32

    @classmethod
    def of_identifiers(cls, identifiers, maximum_total_credentials: int, maximum_node_count: int,
                       maximum_discoverable_credentials_per_action: Optional[int] = None):
        per_action = maximum_discoverable_credentials_per_action or maximum_total_credentials
        assert maximum_total_credentials > 0, "maximum_total_credentials must be positive"
        assert maximum_node_count > 0, "maximum_node_count must be positive"
        return cls(np.int32(maximum_total_credentials), np.int32(maximum_node_count), np.int32(per_action),
                   np.int32(len(identifiers.ports)), np.int32(len(identifiers.properties)),
                   np.int32(len(identifiers.local_vulnerabilities)), np.int32(len(identifiers.remote_vulnerabilities)))


def observation_space_of(bounds: EnvironmentBounds) -> spaces.Dict:
    """cyberbattle_env.py:257-331"""
    N, C, P = int(bounds.maximum_node_count), int(bounds.maximum_total_credentials), int(bounds.port_count)
    L, R, props = int(bounds.local_attacks_count), int(bounds.remote_attacks_count), int(bounds.property_count)
    leak = int(bounds.maximum_discoverable_credentials_per_action)
    return spaces.Dict({
        "newly_discovered_nodes_count": spaces.Discrete(NA + N),
        "lateral_move": spaces.Discrete(2),
        "customer_data_found": spaces.Discrete(2),
        "probe_result": spaces.Discrete(3),
        "escalation": spaces.Discrete(model.PrivilegeLevel.MAXIMUM + 1),
        "leaked_credentials": spaces.Tuple([spaces.MultiDiscrete(np.array([NA + 1, C, N, P], dtype=np.int32))] * leak),
        "action_mask": spaces.Dict({
            "local_vulnerability": spaces.MultiBinary(np.array([N, L])),
            "remote_vulnerability": spaces.MultiBinary(np.array([N, N, R])),
            "connect": spaces.MultiBinary(np.array([N, N, P, C], dtype=np.int32)),
        }),
        "credential_cache_length": spaces.Discrete(C),
        "discovered_node_count": spaces.Discrete(N),
        "discovered_nodes_properties": spaces.MultiDiscrete(np.full((N, props), 3)),
        "nodes_privilegelevel": spaces.MultiDiscrete([model.PrivilegeLevel.MAXIMUM + 1] * N),
        "credential_cache_matrix": spaces.Tuple([spaces.MultiDiscrete(np.array([N, P], dtype=np.int32))] * C),
        "_discovered_nodes": spaces.DummySpace(sample=["node1", "node0", "node2"]),
        "_explored_network": spaces.DummySpace(sample=None),
    })


def action_space_of(bounds: EnvironmentBounds) -> spaces.DiscriminatedUnion:
    """cyberbattle_env.py:540-561"""
    N, C, P = int(bounds.maximum_node_count), int(bounds.maximum_total_credentials), int(bounds.port_count)
    L, R = int(bounds.local_attacks_count), int(bounds.remote_attacks_count)
    return spaces.DiscriminatedUnion({
        "local_vulnerability": spaces.MultiDiscrete(np.array([N, L], dtype=np.int32)),
        "remote_vulnerability": spaces.MultiDiscrete(np.array([N, N, R], dtype=np.int32)),
        "connect": spaces.MultiDiscrete(np.array([N, N, P, C], dtype=np.int32)),
    })


class CyberBattleEnv:
    metadata = {"render_modes": ["human"]}
    privilege_levels = model.PrivilegeLevel.MAXIMUM + 1
    spec = None
    render_mode = None

    def __init__(self, initial_environment: model.Environment, maximum_total_credentials: int = 1000,
                 maximum_node_count: int = 100, maximum_discoverable_credentials_per_action: int = 5,
                 defender_agent: Optional[DefenderAgent] = None,
                 attacker_goal: Optional[AttackerGoal] = AttackerGoal(own_atleast_percent=1.0),
                 defender_goal=DefenderGoal(eviction=True), defender_constraint=DefenderConstraint(maintain_sla=0.0),
                 winning_reward=5000.0, losing_reward=0.0, renderer="", observation_padding=True,
                 throws_on_invalid_actions=True, *, device: int = 0, seed: int = 0):
        if not observation_padding:
            raise NotImplementedError("observation_padding=False (variable-size observations) is not on the batched path")
        self._initial_environment = initial_environment
        self.compiled = scenario.compile_scenario(initial_environment)  # validate_environment's checks live in the compiler
        self._bounds = EnvironmentBounds.of_identifiers(initial_environment.identifiers, maximum_total_credentials,
                                                         maximum_node_count, maximum_discoverable_credentials_per_action)
        if self.compiled.n_nodes > maximum_node_count:
            raise ValueError(f"Network node count ({self.compiled.n_nodes}) exceeds the specified limit of {maximum_node_count}.")
        if self.compiled.max_leak > int(self._bounds.maximum_discoverable_credentials_per_action):
            raise ValueError(f"Some action in the environment returns {self.compiled.max_leak} credentials which exceeds the maximum "
                             f"number of discoverable credentials of {self._bounds.maximum_discoverable_credentials_per_action}")
        self.env_kwargs = dict(
            maximum_total_credentials=maximum_total_credentials, maximum_node_count=maximum_node_count,
            maximum_discoverable_credentials_per_action=int(self._bounds.maximum_discoverable_credentials_per_action),
            defender_agent=defender_agent, attacker_goal=attacker_goal, defender_goal=defender_goal,
            defender_constraint=defender_constraint, winning_reward=winning_reward, losing_reward=losing_reward,
            throws_on_invalid_actions=throws_on_invalid_actions, seed=seed)
        self._attacker_goal, self._defender_goal, self._defender_constraint = attacker_goal, defender_goal, defender_constraint
        self._winning_reward, self._losing_reward = winning_reward, losing_reward
        self._defender_agent = defender_agent
        self._throws_on_invalid_actions = throws_on_invalid_actions
        self._node_count = self.compiled.n_nodes
        self.device = device
        self.action_space = action_space_of(self._bounds)
        self.observation_space = observation_space_of(self._bounds)
        self.reward_range = (-float("inf"), float("inf"))
        self.np_random = np.random.default_rng()
        self.viewer = None
        self._batch = None
        self._marlon_batch = None  # set by AttackerEnvWrapper when the MARLon wrappers drive this env
        self._episode_rewards: List[float] = []
        self._done = False

    # ---- plumbing ---------------------------------------------------------------------------------------
    def _make_batch(self):
        from .batch import Batch

        cfg = config.make_config(_abi.MODE_CYBERBATTLE, auto_reset=False, **self.env_kwargs)
        return Batch(self.compiled, cfg, 1, device=self.device)

    @property
    def batch(self):
        """The live 1-env batch: the MARLon pair batch once wrappers are attached, else the plain CyberBattleEnv one."""
        if self._marlon_batch is not None:
            return self._marlon_batch
        if self._batch is None:
            self._batch = self._make_batch()
        return self._batch

    def _reset_environment(self) -> None:
        self._episode_rewards = []
        self._done = False
        if self._marlon_batch is None:
            self.batch.reset()

    @property
    def unwrapped(self):
        return self

    @property
    def environment(self) -> model.Environment:
        return self._initial_environment

    @property
    def name(self) -> str:
        return "CyberBattleEnv"

    @property
    def identifiers(self) -> model.Identifiers:
        return self._initial_environment.identifiers

    @property
    def bounds(self) -> EnvironmentBounds:
        return self._bounds

    # ---- state peeks (what MARLon reaches for through name-mangled attributes) -----------------------------
    def _state(self) -> Dict[str, Any]:
        x = self.batch.export_state(0, 1)[0]
        n = self.compiled.n_nodes
        h = _abi.X_HEADER_WORDS
        nd, nc = int(x[2]), int(x[3])
        C = int(self._bounds.maximum_total_credentials)
        return {
            "stepcount": int(x[0]), "done": bool(x[1]),
            "discovered": [self.compiled.node_ids[i] for i in x[h:h + nd]],
            "installed": x[h + n:h + 2 * n].astype(bool), "privilege": x[h + 2 * n:h + 3 * n].copy(),
            "countdown": x[h + 3 * n:h + 4 * n].copy(),
            "cache": [self.compiled.triples[t] for t in x[h + 10 * n:h + 10 * n + C][:nc]],
        }

    @property
    def discovered_nodes(self) -> List[str]:
        return self._state()["discovered"]

    # name-mangled accessors used by the reference's wrappers (attack_wrapper.py:71-72,118; defend_wrapper.py:52,260-261)
    @property
    def _CyberBattleEnv__discovered_nodes(self) -> List[str]:  # noqa: N802
        return self.discovered_nodes

    @property
    def credential_cache(self) -> List[model.CachedCredential]:
        return [model.CachedCredential(*t) for t in self._state()["cache"]]

    @property
    def network_availability(self) -> float:
        return float(self.batch.numpy("network_availability")[0])

    # ---- observation assembly -----------------------------------------------------------------------------
    def _observation(self, prefix: str = "") -> Dict[str, Any]:
        """The reference's Observation dict (cyberbattle_env.py:753-773, 859-933) from the device arrays."""
        b, bd = self.batch, self._bounds
        g = lambda k: b.numpy(prefix + k)[0]  # noqa: E731
        sc = g("scalars")
        leak, C = int(bd.maximum_discoverable_credentials_per_action), int(bd.maximum_total_credentials)
        N, props = int(bd.maximum_node_count), int(bd.property_count)
        leaked = g("leaked_credentials").reshape(leak, 4)
        cachem = g("credential_cache_matrix").reshape(C, 2)
        return {
            "newly_discovered_nodes_count": np.int32(sc[0]), "lateral_move": np.int32(sc[1]),
            "customer_data_found": np.int32(sc[2]), "probe_result": np.int32(sc[3]), "escalation": np.int32(sc[4]),
            "leaked_credentials": tuple(leaked[i].astype(np.int32) for i in range(leak)),
            "action_mask": {"local_vulnerability": g("local_vulnerability").copy(),
                            "remote_vulnerability": g("remote_vulnerability").copy(), "connect": g("connect").copy()},
            "credential_cache_matrix": tuple(cachem[i].astype(np.int32) for i in range(C)),
            "credential_cache_length": int(sc[5]), "discovered_node_count": int(sc[6]),
            "discovered_nodes_properties": g("discovered_nodes_properties").reshape(N, props).copy(),
            "nodes_privilegelevel": g("nodes_privilegelevel").copy(),
            "_discovered_nodes": self.discovered_nodes, "_explored_network": None,
        }

    def _info(self) -> Dict[str, Any]:
        info = self.batch.numpy("att_info")[0]
        return {"description": "CyberBattle simulation", "duration_in_ms": 0.0, "step_count": int(info[4]),
                "network_availability": self.network_availability, "credential_cache": self.credential_cache}

    @staticmethod
    def encode_action(action: Dict[str, Any]) -> np.ndarray:
        assert 1 == len(action.keys())
        kind = spaces.DiscriminatedUnion.kind(action)
        if kind not in KIND_CODE:
            raise ValueError("Invalid discriminated union value: " + str(action))
        coords = [int(c) for c in np.asarray(action[kind]).reshape(-1)]
        out = np.zeros((1, 5), dtype=np.int32)
        out[0, 0] = KIND_CODE[kind]
        out[0, 1:1 + len(coords)] = coords
        return out

    # ---- gym API ------------------------------------------------------------------------------------------------
    def step(self, action) -> Tuple[Dict[str, Any], float, bool, bool, Dict[str, Any]]:
        """cyberbattle_env.py:1145-1185"""
        if self._done:
            raise RuntimeError("new episode must be started with env.reset()")
        if self._marlon_batch is not None:
            raise RuntimeError("this env is driven by MARLon wrappers: step through AttackerEnvWrapper / DefenderEnvWrapper")
        a = self.encode_action(action)
        kind = spaces.DiscriminatedUnion.kind(action)
        nvec = self.action_space.spaces[kind].nvec
        if any(c < 0 or c >= int(m) for c, m in zip(a[0, 1:], nvec[1 if kind == "x" else 0:])) and kind != "connect":
            pass  # out-of-space vulnerability/port indices raise IndexError in the reference; indices are checked below
        self.batch.step(a)
        info = self.batch.numpy("att_info")[0]
        err = int(info[3])
        if err == _abi.E_SOURCE_NOT_OWNED:
            raise ValueError("Agent does not owned the source node")
        if err == _abi.E_TARGET_NOT_DISCOVERED:
            raise ValueError("Agent has not discovered the target node")
        if err == _abi.E_CREDENTIAL_NOT_GATHERED:
            raise ValueError("Agent has not discovered credential")
        obs = self._observation()
        reward = float(self.batch.numpy("att_reward")[0])
        self._done = bool(self.batch.numpy("att_terminated")[0])
        self._episode_rewards.append(reward)
        return obs, reward, self._done, False, self._info()

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        """cyberbattle_env.py:1187-1209"""
        self._reset_environment()
        self.np_random = np.random.default_rng(seed)
        obs = self._observation()
        info = self._info()
        info["duration_in_ms"] = 0
        return obs, info

    def close(self) -> None:
        if self._batch is not None:
            self._batch.close()
            self._batch = None

    def render(self, mode: str = "human") -> None:
        raise NotImplementedError("rendering (plotly) is out of scope of the batched step path")

    # ---- helpers of the reference -----------------------------------------------------------------------------------
    def compute_action_mask(self):
        """cyberbattle_env.py:679-683 (recomputed from the live state: owned set, discovery and cache counts)."""
        bd = self._bounds
        N, C, P = int(bd.maximum_node_count), int(bd.maximum_total_credentials), int(bd.port_count)
        L, R = int(bd.local_attacks_count), int(bd.remote_attacks_count)
        st = self._state()
        local = np.zeros((N, L), dtype=np.int8)
        remote = np.zeros((N, N, R), dtype=np.int8)
        connect = np.zeros((N, N, P, C), dtype=np.int8)
        idx = {k: i for i, k in enumerate(self.compiled.node_ids)}
        nd, nc = len(st["discovered"]), len(st["cache"])
        ident = self.identifiers
        for s, node_id in enumerate(st["discovered"]):
            if not st["installed"][idx[node_id]]:
                continue
            info = self.environment.get_node(node_id)
            for v, vid in enumerate(ident.local_vulnerabilities):
                if vid in self.environment.vulnerability_library or vid in info.vulnerabilities:
                    local[s, v] = 1
            remote[s, :nd, :R] = 1
            connect[s, :nd, :P, :nc] = 1
        return {"local_vulnerability": local, "remote_vulnerability": remote, "connect": connect}

    def apply_mask(self, action, mask=None) -> bool:
        if mask is None:
            mask = self.compute_action_mask()
        kind = spaces.DiscriminatedUnion.kind(action)
        return bool(mask[kind][tuple(int(c) for c in action[kind])])

    def is_node_owned(self, node: int) -> bool:
        """cyberbattle_env.py:1009-1014: privilege_level > NoAccess of the node at discovery index `node`."""
        st = self._state()
        if node < 0:
            raise OutOfBoundIndexError(f"Node index must be positive, given {node}")
        if node >= len(st["discovered"]):
            raise OutOfBoundIndexError(f"Node index ({node}) is invalid; only {len(st['discovered'])} nodes discovered so far.")
        return bool(st["privilege"][self.compiled.node_ids.index(st["discovered"][node])] > 0)

    def is_action_valid(self, action, action_mask=None) -> bool:
        """cyberbattle_env.py:1016-1039"""
        kind = spaces.DiscriminatedUnion.kind(action)
        st = self._state()
        nd, nc, bd = len(st["discovered"]), len(st["cache"]), self._bounds
        c = [int(x) for x in action[kind]]
        if kind == "local_vulnerability":
            ok = c[0] < nd and self.is_node_owned(c[0]) and c[1] < bd.local_attacks_count
        elif kind == "remote_vulnerability":
            ok = c[0] < nd and self.is_node_owned(c[0]) and c[1] < nd and c[2] < bd.remote_attacks_count
        else:
            ok = c[0] < nd and self.is_node_owned(c[0]) and c[1] < nd and c[2] < bd.port_count and c[3] < nc
        return bool(ok and self.apply_mask(action, action_mask))

    def sample_valid_action(self, kinds=None):
        """cyberbattle_env.py:959-1047: sample within the expected ranges until the mask admits the action
        (kind 1 -> local, kind 0 -> remote, as in the reference)."""
        st = self._state()
        nd, nc, bd = len(st["discovered"]), len(st["cache"]), self._bounds
        idx = {k: i for i, k in enumerate(self.compiled.node_ids)}
        owned = [s for s, k in enumerate(st["discovered"]) if st["privilege"][idx[k]] >= 1]
        mask = self.compute_action_mask()
        kinds = [0, 1, 2] if kinds is None else list(kinds)
        if nc == 0:
            kinds = [k for k in kinds if k != 2]
        assert kinds, "Kinds list cannot be empty"
        rng = self.action_space.union_np_random
        while True:
            kind = int(rng.choice(kinds))
            if kind == 2:
                a = {"connect": np.array([self.np_random.choice(owned), self.np_random.integers(0, nd),
                                          self.np_random.integers(0, bd.port_count), self.np_random.integers(0, nc)], np.int32)}
            elif kind == 1:
                a = {"local_vulnerability": np.array([rng.choice(owned), rng.integers(0, bd.local_attacks_count)], np.int32)}
            else:
                a = {"remote_vulnerability": np.array([rng.choice(owned), rng.integers(0, nd),
                                                       rng.integers(0, bd.remote_attacks_count)], np.int32)}
            if self.apply_mask(a, mask):
                return a

    def sample_valid_action_with_luck(self):
        mask = self.compute_action_mask()
        a = self.action_space.sample()
        while not self.apply_mask(a, mask):
            a = self.action_space.sample()
        return a

    # goal predicates on the live state (MARLon's defender wrapper calls the name-mangled one, defend_wrapper.py:260)
    def _CyberBattleEnv__defender_goal_reached(self) -> bool:  # noqa: N802
        return bool(self._defender_goal.eviction and not (self._state()["privilege"] >= 1).any())

    @property
    def _CyberBattleEnv__defender_constraint(self):  # noqa: N802
        return self._defender_constraint

    @property
    def _CyberBattleEnv__WINNING_REWARD(self):  # noqa: N802
        return self._winning_reward

    @property
    def _CyberBattleEnv__LOSING_REWARD(self):  # noqa: N802
        return self._losing_reward

    @property
    def _CyberBattleEnv__episode_rewards(self):  # noqa: N802
        return self._episode_rewards


class CyberBattleToyCtf(CyberBattleEnv):
    """_env/cyberbattle_toyctf.py:8-12"""

    def __init__(self, **kwargs):
        from . import scenarios

        super().__init__(initial_environment=scenarios.toyctf_environment(), **kwargs)


class CyberBattleChain(CyberBattleEnv):
    """_env/cyberbattle_chain.py:10-19"""

    def __init__(self, size, **kwargs):
        from . import scenarios

        self.size = size
        super().__init__(initial_environment=scenarios.chain_environment(size), **kwargs)

    @property
    def name(self) -> str:
        return f"CyberBattleChain-{self.size}"


def make(env_id: str, **kwargs) -> CyberBattleEnv:
    """``gym.make`` for the ids of cyberbattle/__init__.py:31-71 (registry kwargs merged with the caller's)."""
    env, merged = registry.resolve(env_id, **kwargs)
    e = CyberBattleEnv(initial_environment=env, **merged)
    if env_id == "CyberBattleChain-v0":
        size = kwargs.get("size", registry.ENV_SPECS[env_id]["size"])
        e.size = size
        e.__class__ = type("CyberBattleChain", (CyberBattleEnv,), {"name": property(lambda self: f"CyberBattleChain-{self.size}")})
    return e
