"""Env / Wrapper base classes of the gymnasium stand-in (test infrastructure)."""
from typing import Generic, TypeVar

ObsType = TypeVar("ObsType")
ActType = TypeVar("ActType")


class Env(Generic[ObsType, ActType]):
    metadata = {"render_modes": []}
    render_mode = None
    spec = None
    action_space = None
    observation_space = None
    np_random = None

    def step(self, action):
        raise NotImplementedError

    def reset(self, *, seed=None, options=None):
        raise NotImplementedError

    def render(self):
        raise NotImplementedError

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def close(self):
        return self.env.close()

    @property
    def unwrapped(self):
        return self.env.unwrapped


class ObservationWrapper(Wrapper):
    pass


class ActionWrapper(Wrapper):
    pass
