"""Device-resident rollout collection.

The reference collects experience with ``marl_algorithm.collect_rollouts`` (marl_algorithm.py:17-54): per step each agent's
``perform_step`` (baseline_marlon_agent.py:100-167) moves the observation to the policy's device, runs the policy, copies
the actions back to numpy, steps its one environment and appends to an SB3 ``RolloutBuffer``; ``on_rollout_end``
(:276-284) bootstraps the last value and calls ``compute_returns_and_advantage``.  At 1e8 env-steps/s the host round
trips would be the whole cost, so here nothing leaves the GPU: observations are the batch's own tensors, the policies are
callables on device tensors, actions / rewards / episode starts / values / log-probabilities land in ``[n_steps, n_envs]``
device buffers and the advantages come from one kernel (``cbx_gae``).

Like MARLon's ``perform_step`` -- and unlike stock SB3 -- no value bootstrap is added for time-limit truncations.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Dict, Optional, Tuple

from . import _lib

# policy(observation: dict of device tensors, action_masks or None) -> (actions int32 [n, A], values float32 [n], log_probs float32 [n])
Policy = Callable[[Dict[str, Any], Optional[Any]], Tuple[Any, Any, Any]]


class DeviceRolloutBuffer:
    """What SB3's ``RolloutBuffer`` holds for one agent, as device tensors; observations are stored through `obs_keys`
    (default: none -- a 65 536-env rollout of dense masks does not fit anywhere; store the features your policy needs)."""

    def __init__(self, n_steps: int, n_envs: int, action_width: int, device, gamma: float = 0.99, gae_lambda: float = 0.95,
                 obs_spec: Optional[Dict[str, Tuple[Tuple[int, ...], Any]]] = None):
        import torch

        self._torch = torch
        self.n_steps, self.n_envs, self.gamma, self.gae_lambda = int(n_steps), int(n_envs), float(gamma), float(gae_lambda)
        self.device = device
        f32 = dict(dtype=torch.float32, device=device)
        self.actions = torch.zeros((n_steps, n_envs, action_width), dtype=torch.int32, device=device)
        self.rewards = torch.zeros((n_steps, n_envs), **f32)
        self.values = torch.zeros((n_steps, n_envs), **f32)
        self.log_probs = torch.zeros((n_steps, n_envs), **f32)
        self.episode_starts = torch.zeros((n_steps, n_envs), dtype=torch.uint8, device=device)
        self.advantages = torch.zeros((n_steps, n_envs), **f32)
        self.returns = torch.zeros((n_steps, n_envs), **f32)
        self.observations = {k: torch.zeros((n_steps, n_envs) + tuple(shape), dtype=dt, device=device)
                             for k, (shape, dt) in (obs_spec or {}).items()}
        self.pos, self.full = 0, False

    def reset(self):
        self.pos, self.full = 0, False

    def store_observation(self, obs: Dict[str, Any]):
        """Keep the observation the next action is chosen on (call BEFORE the step: the batch's tensors are overwritten by it)."""
        for k, buf in self.observations.items():
            buf[self.pos].copy_(obs[k].reshape(buf[self.pos].shape))

    def add(self, obs: Optional[Dict[str, Any]], actions, rewards, episode_starts, values, log_probs):
        t = self.pos
        if obs is not None:
            self.store_observation(obs)
        self.actions[t].copy_(actions.reshape(self.actions[t].shape))
        self.rewards[t].copy_(rewards)
        self.episode_starts[t].copy_(episode_starts)
        self.values[t].copy_(values.reshape(-1))
        self.log_probs[t].copy_(log_probs.reshape(-1))
        self.pos += 1
        self.full = self.pos == self.n_steps

    def compute_returns_and_advantage(self, last_values, dones):
        """SB3 ``RolloutBuffer.compute_returns_and_advantage`` (GAE) on the device: one launch of ``cbx_gae``."""
        torch = self._torch
        L = _lib.load()
        lv = last_values.reshape(-1).to(torch.float32).contiguous()
        ld = dones.reshape(-1).to(torch.uint8).contiguous()
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        with torch.cuda.device(self.device):
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(L.cbx_gae(p(self.rewards), p(self.values), p(self.episode_starts), p(lv), p(ld), self.gamma, self.gae_lambda,
                                 self.n_steps, self.n_envs, p(self.advantages), p(self.returns), stream))
        self._keep = (lv, ld)

    def get(self, batch_size: Optional[int] = None, generator=None):
        """SB3 ``RolloutBuffer.get``: shuffled minibatches of ``RolloutBufferSamples`` over the whole buffer (device tensors; the
        flat sample index is SB3's env-major ``swap_and_flatten`` order, drawn with ``torch.randperm`` on the device)."""
        from .ppo import RolloutBufferSamples

        torch = self._torch
        assert self.full, "rollout buffer not full"
        T, n = self.n_steps, self.n_envs
        total = T * n
        perm = torch.randperm(total, device=self.device, generator=generator)
        batch_size = total if batch_size is None else int(batch_size)
        for start in range(0, total, batch_size):
            idx = perm[start:start + batch_size]
            e, t = idx // T, idx % T  # swap_and_flatten: sample i = (env i // n_steps, step i % n_steps)
            obs = {k: v[t, e] for k, v in self.observations.items()}
            yield RolloutBufferSamples(obs, self.actions[t, e], self.values[t, e], self.log_probs[t, e], self.advantages[t, e],
                                       self.returns[t, e])


def collect_rollouts(universe, attacker_policy: Policy, attacker_buffer: DeviceRolloutBuffer,
                     defender_policy: Optional[Policy] = None, defender_buffer: Optional[DeviceRolloutBuffer] = None,
                     attacker_action_masks: bool = False) -> bool:
    """``marl_algorithm.collect_rollouts`` (marl_algorithm.py:17-54) over a ``MultiAgentUniversalEnv``: attacker step then
    defender step per iteration, both agents' buffers filled, GAE at the end -- every tensor stays on the GPU.

    The universe must have been reset; its current observations are the agents' ``_last_obs``."""
    import torch

    n = universe.n_envs
    dev = universe.batch.torch_device
    has_def = universe.has_defender and defender_policy is not None
    for attr in ("_att_starts", "_def_starts"):  # _last_episode_starts: true right after reset()
        if not hasattr(universe, attr):
            setattr(universe, attr, torch.ones(n, dtype=torch.uint8, device=dev))
    attacker_buffer.reset()
    if has_def:
        defender_buffer.reset()
    t = universe.batch.tensors
    n_steps = min(attacker_buffer.n_steps, defender_buffer.n_steps) if has_def else attacker_buffer.n_steps
    with torch.no_grad():
        for _ in range(n_steps):
            aobs = universe.attacker_observation()
            masks = universe.action_masks() if attacker_action_masks else None
            a_act, a_val, a_lp = attacker_policy(aobs, masks)
            a_act = a_act.to(torch.int32)
            attacker_buffer.store_observation(aobs)  # before the step overwrites the batch's tensors
            if has_def:
                dobs = universe.defender_observation()
                defender_buffer.store_observation(dobs)
                d_act, d_val, d_lp = defender_policy(dobs, None)
                d_act = d_act.to(torch.int32)
                universe.batch.step(a_act, d_act)  # one launch: the attacker's move, then the defender's
            else:
                universe.batch.step(a_act, None, who=universe.batch.WHO_ATTACKER if universe.has_defender else 3)
            a_done = t["att_terminated"] | t["att_truncated"]
            attacker_buffer.add(None, a_act, t["att_reward"], universe._att_starts, a_val, a_lp)
            universe._att_starts = a_done.clone()
            if has_def:
                d_done = t["def_terminated"] | t["def_truncated"]
                defender_buffer.add(None, d_act, t["def_reward"], universe._def_starts, d_val, d_lp)
                universe._def_starts = d_done.clone()
        # on_rollout_end (baseline_marlon_agent.py:276-284): bootstrap from the value of the observation after the last step
        _, a_last, _ = attacker_policy(universe.attacker_observation(), universe.action_masks() if attacker_action_masks else None)
        attacker_buffer.compute_returns_and_advantage(a_last, universe._att_starts)
        if has_def:
            _, d_last, _ = defender_policy(universe.defender_observation(), None)
            defender_buffer.compute_returns_and_advantage(d_last, universe._def_starts)
    return True
