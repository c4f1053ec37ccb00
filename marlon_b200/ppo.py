"""PPO over the device-resident rollout (second half).

The reference trains with stable-baselines3's PPO behind ``BaselineMarlonAgent.train`` (``baseline_marlon_agent.py:208-209`` ->
``PPO.train``): minibatches of ``RolloutBufferSamples`` from the rollout buffer, clipped surrogate objective, value loss,
entropy bonus, gradient clipping.  Here the rollout already lives in HBM (``rollout.DeviceRolloutBuffer``), so the update
runs over it in place: ``DeviceRolloutBuffer.get(batch_size)`` yields the same named tuple of samples (device tensors,
SB3's env-major flattening) and ``ppo_update`` is SB3's ``PPO.train`` loop restated -- same loss, same defaults
(``clip_range=0.2, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, n_epochs=10, batch_size=64``, advantage normalisation per
minibatch).  ``MultiDiscretePolicy`` is a small actor-critic for the two MARLon ``MultiDiscrete`` action spaces (one
categorical head per component, as SB3's ``MultiCategoricalDistribution``) over the integer observation fields.

PyTorch here is plumbing for a consumer of the hot path, not the hot path: the environment step, the observation encoding and
the advantage computation are the CUDA library's (``cbx_batch_step``, ``cbx_gae``).
"""
from __future__ import annotations

from typing import Any, Dict, List, NamedTuple, Optional, Sequence

import torch
from torch import nn


class RolloutBufferSamples(NamedTuple):
    """stable_baselines3.common.type_aliases.RolloutBufferSamples / DictRolloutBufferSamples (observations: tensor or dict)."""
    observations: Any
    actions: torch.Tensor
    old_values: torch.Tensor
    old_log_prob: torch.Tensor
    advantages: torch.Tensor
    returns: torch.Tensor


ATTACKER_FEATURES = ["newly_discovered_nodes_count", "lateral_move", "customer_data_found", "probe_result", "escalation",
                     "credential_cache_length", "discovered_node_count", "leaked_credentials", "credential_cache_matrix",
                     "discovered_nodes_properties", "nodes_privilegelevel"]
DEFENDER_FEATURES = ["infected_nodes", "incoming_firewall_status", "outgoing_firewall_status", "services_status"]


def flatten_observation(obs: Dict[str, torch.Tensor], keys: Sequence[str]) -> torch.Tensor:
    """[n, F] float32 features from the integer observation fields named in `keys` (the dense masks are left out: a
    policy reads them as action masks, not as 11 268 input features)."""
    cols = [obs[k].reshape(obs[k].shape[0], -1).to(torch.float32) for k in keys]
    return torch.cat(cols, dim=1)


class MultiDiscretePolicy(nn.Module):
    """Actor-critic with one categorical head per component of a ``MultiDiscrete`` action space.  All heads are evaluated
    together: the logits are scattered into a padded [n, components, max choices] tensor (-inf padding), so sampling
    (Gumbel-max), log-probabilities and entropies are a handful of kernels whatever the number of components -- the same
    distribution as one ``torch.distributions.Categorical`` per head, without its per-head launches and argument checks."""

    def __init__(self, n_features: int, nvec: Sequence[int], feature_keys: Sequence[str], hidden: int = 64):
        super().__init__()
        self.nvec = [int(x) for x in nvec]
        self.feature_keys = list(feature_keys)
        self.body = nn.Sequential(nn.Linear(n_features, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh())
        self.pi = nn.Linear(hidden, sum(self.nvec))
        self.vf = nn.Linear(hidden, 1)
        A, K = len(self.nvec), max(self.nvec)
        index = torch.zeros(A, K, dtype=torch.long)   # column of the flat logits that feeds padded slot (a, k)
        valid = torch.zeros(A, K, dtype=torch.bool)
        off = 0
        for a, n in enumerate(self.nvec):
            index[a, :n] = torch.arange(off, off + n)
            valid[a, :n] = True
            off += n
        self.register_buffer("_index", index.reshape(-1), persistent=False)
        self.register_buffer("_valid", valid, persistent=False)
        # one head much larger than the rest (generated networks: 192 credentials next to 3 action kinds) would make the padded
        # tensor several times the logits: such spaces are evaluated head by head instead
        self._padded = A * K <= 2 * sum(self.nvec)

    @classmethod
    def for_space(cls, observation: Dict[str, torch.Tensor], nvec, feature_keys, hidden: int = 64) -> "MultiDiscretePolicy":
        n_features = flatten_observation({k: observation[k][:1] for k in feature_keys}, feature_keys).shape[1]
        return cls(n_features, nvec, feature_keys, hidden)

    def _log_probs(self, obs):
        """-> (log-probabilities [n, A, K] with -inf at the padding, values [n])"""
        x = obs if torch.is_tensor(obs) else flatten_observation(obs, self.feature_keys)
        h = self.body(x)
        A, K = self._valid.shape
        logits = self.pi(h).index_select(1, self._index).reshape(-1, A, K).masked_fill(~self._valid, float("-inf"))
        return torch.log_softmax(logits, dim=2), self.vf(h).squeeze(1)

    def _per_head(self, obs, actions=None):
        """The same distribution head by head (no padding): -> (actions [n, A], values, log_prob, entropy)."""
        x = obs if torch.is_tensor(obs) else flatten_observation(obs, self.feature_keys)
        h = self.body(x)
        logits, values = self.pi(h), self.vf(h).squeeze(1)
        picked, lps, ents, off = [], 0.0, 0.0, 0
        for a, n in enumerate(self.nvec):
            lp = torch.log_softmax(logits[:, off:off + n], dim=1)
            off += n
            if actions is None:
                u = torch.rand_like(lp).clamp_(1e-20, 1.0)
                act = torch.argmax(lp - torch.log(-torch.log(u)), dim=1)
            else:
                act = actions[:, a].long()
            picked.append(act)
            lps = lps + lp.gather(1, act.unsqueeze(1)).squeeze(1)
            ents = ents - (torch.exp(lp) * lp).sum(1)
        return torch.stack(picked, dim=1), values, lps, ents

    def forward(self, obs, action_masks=None):
        """-> (actions int32 [n, A], values [n], log_probs [n]); the call signature ``rollout.collect_rollouts`` expects."""
        if not self._padded:
            actions, values, lp, _ = self._per_head(obs)
            return actions.to(torch.int32), values, lp
        logp, values = self._log_probs(obs)
        u = torch.rand_like(logp).clamp_(1e-20, 1.0)
        actions = torch.argmax(logp - torch.log(-torch.log(u)), dim=2)  # Gumbel-max: a sample of each head's categorical
        return actions.to(torch.int32), values, logp.gather(2, actions.unsqueeze(2)).squeeze(2).sum(1)

    def evaluate_actions(self, obs, actions):
        """-> (values, log_prob, entropy) of `actions` under the current policy (SB3 ``ActorCriticPolicy.evaluate_actions``)."""
        if not self._padded:
            _, values, lp, ent = self._per_head(obs, actions)
            return values, lp, ent
        logp, values = self._log_probs(obs)
        chosen = logp.gather(2, actions.long().unsqueeze(2)).squeeze(2).sum(1)
        safe = logp.masked_fill(~self._valid, 0.0)  # not where(valid, p * logp, 0): 0 * -inf poisons the gradient
        entropy = -(torch.exp(safe) * safe).masked_fill(~self._valid, 0.0).sum((1, 2))
        return values, chosen, entropy


def ppo_loss(policy, s: RolloutBufferSamples, clip_range: float = 0.2, ent_coef: float = 0.0, vf_coef: float = 0.5,
             normalize_advantage: bool = True):
    """The loss of one minibatch as stable-baselines3's ``PPO.train`` computes it (no value clipping: ``clip_range_vf=None``)."""
    values, log_prob, entropy = policy.evaluate_actions(s.observations, s.actions)
    adv = s.advantages
    if normalize_advantage and adv.numel() > 1:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    ratio = torch.exp(log_prob - s.old_log_prob)
    policy_loss = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - clip_range, 1 + clip_range)).mean()
    value_loss = torch.nn.functional.mse_loss(s.returns, values)
    entropy_loss = -entropy.mean()
    loss = policy_loss + ent_coef * entropy_loss + vf_coef * value_loss
    return loss, dict(policy_loss=policy_loss.detach(), value_loss=value_loss.detach(), entropy_loss=entropy_loss.detach(),
                      approx_kl=((ratio - 1) - (log_prob - s.old_log_prob)).mean().detach())


def ppo_update(policy, optimizer, buffer, n_epochs: int = 10, batch_size: int = 64, clip_range: float = 0.2, ent_coef: float = 0.0,
               vf_coef: float = 0.5, max_grad_norm: float = 0.5, generator: Optional[torch.Generator] = None) -> List[Dict[str, float]]:
    """``PPO.train`` over a full ``DeviceRolloutBuffer``: `n_epochs` passes of shuffled minibatches, one optimiser step each."""
    logs = []
    policy.train()
    for _ in range(n_epochs):
        for s in buffer.get(batch_size, generator=generator):
            loss, log = ppo_loss(policy, s, clip_range, ent_coef, vf_coef)
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(policy.parameters(), max_grad_norm)
            optimizer.step()
            log["loss"] = loss.detach()
            logs.append(log)
    return [{k: float(v) for k, v in log.items()} for log in logs]  # one host synchronisation, after the last step
