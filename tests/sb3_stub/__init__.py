"""A stand-in for the few stable-baselines3 names the reference's MARLon modules import -- TEST INFRASTRUCTURE ONLY.

stable-baselines3 (pinned 2.3.2 by the reference) is not installable here (no network, not in the wheelhouse), so the
reference's ``multiagent_universe`` / ``marl_algorithm`` / ``random_marlon_agent`` modules -- the CALLERS of the hot path --
cannot be imported as they are.  ``install()`` registers minimal modules under the ``stable_baselines3`` names so that those
files import unmodified and their loops (``collect_rollouts``, ``run_episode``, ``RandomMarlonAgent.perform_step``) can drive
either the reference's own wrappers or this package's.  The shapes follow SB3's documented interfaces:

* ``common.type_aliases.GymEnv`` -- a typing alias;
* ``common.callbacks.BaseCallback`` -- the abstract callback base (only subclassed by the reference);
* ``common.monitor.Monitor`` -- a transparent env wrapper that adds ``info["episode"] = {"r", "l", "t"}`` at episode end;
* ``common.vec_env.base_vec_env.VecEnv`` -- the abstract vectorised-env base: ``__init__(num_envs, observation_space,
  action_space)``, abstract ``reset / step_async / step_wait / close / get_attr / set_attr / env_method / env_is_wrapped``,
  concrete ``step``.

Nothing in the product package imports this module; ``marlon_b200.vec_env`` subclasses whatever
``stable_baselines3.common.vec_env.base_vec_env.VecEnv`` is importable, which in the tests that call ``install()`` is this one.
"""
import abc
import sys
import time
import types
from typing import Any


class BaseCallback(abc.ABC):
    def __init__(self, verbose: int = 0):
        self.verbose, self.n_calls, self.num_timesteps, self.locals, self.globals = verbose, 0, 0, {}, {}

    def update_locals(self, locals_):
        self.locals.update(locals_)

    def on_step(self) -> bool:
        self.n_calls += 1
        return self._on_step()

    @abc.abstractmethod
    def _on_step(self) -> bool:
        ...


class Monitor:
    """Transparent wrapper: forwards everything, records episode return / length (SB3 ``Monitor`` contract)."""

    def __init__(self, env, filename=None, allow_early_resets=True, **_):
        self.env = env
        self.t_start = time.time()
        self.rewards, self.episode_returns, self.episode_lengths = [], [], []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)

    def reset(self, **kwargs):
        self.rewards = []
        return self.env.reset(**kwargs)

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self.rewards.append(float(reward))
        if terminated or truncated:
            info = dict(info)
            info["episode"] = {"r": round(sum(self.rewards), 6), "l": len(self.rewards), "t": round(time.time() - self.t_start, 6)}
            self.episode_returns.append(info["episode"]["r"])
            self.episode_lengths.append(info["episode"]["l"])
        return obs, reward, terminated, truncated, info


class VecEnv(abc.ABC):
    def __init__(self, num_envs: int, observation_space, action_space):
        self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space
        self.reset_infos = [{} for _ in range(num_envs)]
        self._seeds = [None for _ in range(num_envs)]
        self._options = [{} for _ in range(num_envs)]
        self.render_mode = None

    @abc.abstractmethod
    def reset(self):
        ...

    @abc.abstractmethod
    def step_async(self, actions) -> None:
        ...

    @abc.abstractmethod
    def step_wait(self):
        ...

    @abc.abstractmethod
    def close(self) -> None:
        ...

    @abc.abstractmethod
    def get_attr(self, attr_name, indices=None):
        ...

    @abc.abstractmethod
    def set_attr(self, attr_name, value, indices=None) -> None:
        ...

    @abc.abstractmethod
    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        ...

    @abc.abstractmethod
    def env_is_wrapped(self, wrapper_class, indices=None):
        ...

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()


def install():
    """Register the stand-in under the stable_baselines3 module names (idempotent; a real installation wins)."""
    try:
        import stable_baselines3  # noqa: F401

        if not getattr(stable_baselines3, "__cbx_stub__", False):
            return False
    except ImportError:
        pass

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    root = mod("stable_baselines3", __cbx_stub__=True, __version__="2.3.2+standin")
    common = mod("stable_baselines3.common")
    root.common = common
    common.type_aliases = mod("stable_baselines3.common.type_aliases", GymEnv=Any)
    common.callbacks = mod("stable_baselines3.common.callbacks", BaseCallback=BaseCallback)
    common.monitor = mod("stable_baselines3.common.monitor", Monitor=Monitor)
    base = mod("stable_baselines3.common.vec_env.base_vec_env", VecEnv=VecEnv)
    common.vec_env = mod("stable_baselines3.common.vec_env", VecEnv=VecEnv, base_vec_env=base)
    return True
