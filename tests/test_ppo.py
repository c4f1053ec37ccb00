"""PPO over the device rollout (marlon_b200/ppo.py, rollout.DeviceRolloutBuffer.get): the minibatch loss against a plain
restatement of stable-baselines3's ``PPO.train`` formulas, the sample iterator's coverage and SB3's flattening order (CPU
tensors), and -- on the GPU -- a full collect_rollouts + ppo_update round for both agents with everything resident in HBM."""
import numpy as np
import pytest
import torch

from marlon_b200 import ppo
from marlon_b200.rollout import DeviceRolloutBuffer


def test_ppo_loss_is_sb3s():
    torch.manual_seed(0)
    n, F, nvec = 96, 19, [3, 4, 5]
    pol = ppo.MultiDiscretePolicy(F, nvec, ["x"])
    obs = {"x": torch.randn(n, F)}
    a, v, lp = pol(obs)
    assert a.shape == (n, 3) and a.dtype == torch.int32 and all(int(a[:, k].max()) < nvec[k] for k in range(3))
    v2, lp2, ent = pol.evaluate_actions(obs, a)
    assert torch.allclose(lp, lp2, atol=1e-6) and torch.allclose(v, v2)
    s = ppo.RolloutBufferSamples(obs, a, v.detach(), lp.detach() + 0.1 * torch.randn(n), torch.randn(n), torch.randn(n))
    loss, log = ppo.ppo_loss(pol, s, clip_range=0.2, ent_coef=0.01, vf_coef=0.5)
    adv = (s.advantages - s.advantages.mean()) / (s.advantages.std() + 1e-8)  # PPO.train: normalize_advantage
    ratio = torch.exp(lp2 - s.old_log_prob)
    want = (-torch.min(adv * ratio, adv * ratio.clamp(0.8, 1.2)).mean() + 0.01 * (-ent.mean()) + 0.5 * ((s.returns - v2) ** 2).mean())
    assert torch.allclose(loss, want, atol=1e-6)
    assert set(log) == {"policy_loss", "value_loss", "entropy_loss", "approx_kl"}
    loss.backward()  # heads of different sizes are padded with -inf logits: the padding must not poison the gradient
    assert all(torch.isfinite(p.grad).all() for p in pol.parameters())
    # the batched heads against one torch.distributions.Categorical per head
    logits = torch.split(pol.pi(pol.body(obs["x"])), nvec, dim=1)
    dists = [torch.distributions.Categorical(logits=lg) for lg in logits]
    assert torch.allclose(ent, sum(d.entropy() for d in dists), atol=1e-5)
    assert torch.allclose(lp2, sum(d.log_prob(a[:, k].long()) for k, d in enumerate(dists)), atol=1e-5)


def test_rollout_buffer_get_covers_every_sample_once_in_sb3_order():
    T, n = 5, 7
    buf = DeviceRolloutBuffer(T, n, 2, torch.device("cpu"), obs_spec={"o": ((3,), torch.float32)})
    for t in range(T):
        base = torch.arange(n, dtype=torch.float32) * 100 + t  # value encodes (env, step)
        buf.add({"o": base[:, None].repeat(1, 3)}, torch.stack([torch.arange(n), torch.full((n,), t)], 1).to(torch.int32), base,
                torch.zeros(n, dtype=torch.uint8), base + 0.5, -base)
    buf.advantages.copy_(buf.rewards)
    buf.returns.copy_(buf.rewards * 2)
    seen = []
    g = torch.Generator().manual_seed(1)
    for s in buf.get(batch_size=8, generator=g):
        assert s.actions.shape[1] == 2 and s.observations["o"].shape[1:] == (3,)
        env, step = s.actions[:, 0].float(), s.actions[:, 1].float()
        assert torch.equal(s.advantages, env * 100 + step) and torch.equal(s.old_values, env * 100 + step + 0.5)
        assert torch.equal(s.old_log_prob, -(env * 100 + step)) and torch.equal(s.returns, 2 * (env * 100 + step))
        assert torch.equal(s.observations["o"][:, 0], env * 100 + step)
        seen += [(int(e), int(t)) for e, t in zip(env, step)]
    assert sorted(seen) == [(e, t) for e in range(n) for t in range(T)]
    whole = next(iter(buf.get(None, generator=torch.Generator().manual_seed(2))))
    assert whole.actions.shape[0] == T * n


def test_skewed_action_space_is_evaluated_head_by_head_with_the_same_distribution():
    """One head much larger than the others (generated networks: 192 credentials next to 3 action kinds): the padded batch of
    heads would be several times the logits, so the policy evaluates head by head -- same log-probabilities and entropies."""
    torch.manual_seed(0)
    pol = ppo.MultiDiscretePolicy(50, [3, 72, 5, 72, 72, 9, 72, 72, 7, 192], ["x"])
    assert not pol._padded
    x = torch.randn(64, 50)
    actions, values, lp = pol(x)
    assert actions.dtype == torch.int32 and all(int(actions[:, a].max()) < n for a, n in enumerate(pol.nvec))
    v1, lp1, ent1 = pol.evaluate_actions(x, actions)
    assert torch.allclose(lp, lp1, atol=1e-5) and torch.allclose(values, v1)
    pol._padded = True  # the padded evaluation of the same heads
    v2, lp2, ent2 = pol.evaluate_actions(x, actions)
    assert torch.allclose(lp1, lp2, atol=1e-5) and torch.allclose(ent1, ent2, atol=1e-4) and torch.allclose(v1, v2)


@pytest.mark.gpu
def test_collect_rollouts_and_ppo_update_on_device():
    from marlon_b200.rollout import collect_rollouts
    from marlon_b200.universe import MultiAgentUniversalEnv

    torch.manual_seed(3)
    n, T = 512, 16
    u = MultiAgentUniversalEnv("CyberBattleToyCtf-v0", n, maximum_node_count=12, maximum_total_credentials=10, max_timesteps=40)
    aobs, dobs = u.reset()
    dev = u.batch.torch_device
    apol = ppo.MultiDiscretePolicy.for_space(aobs, u.attacker_action_space.nvec, ppo.ATTACKER_FEATURES).to(dev)
    dpol = ppo.MultiDiscretePolicy.for_space(dobs, u.defender_action_space.nvec, ppo.DEFENDER_FEATURES).to(dev)
    a_spec = {k: (tuple(aobs[k].shape[1:]), aobs[k].dtype) for k in ppo.ATTACKER_FEATURES}
    d_spec = {k: (tuple(dobs[k].shape[1:]), dobs[k].dtype) for k in ppo.DEFENDER_FEATURES}
    ab = DeviceRolloutBuffer(T, n, 10, dev, obs_spec=a_spec)
    db = DeviceRolloutBuffer(T, n, 12, dev, obs_spec=d_spec)
    assert collect_rollouts(u, apol, ab, dpol, db)
    assert ab.full and db.full and u.episode_statistics(reduce=False)["env_steps"] == n * T
    for pol, buf in ((apol, ab), (dpol, db)):
        opt = torch.optim.Adam(pol.parameters(), lr=3e-4)
        before = [p.detach().clone() for p in pol.parameters()]
        # the stored log-probabilities are the policy's own: ratio 1, approx_kl 0 on the first minibatch
        first = next(iter(buf.get(256)))
        _, lp, _ = pol.evaluate_actions(first.observations, first.actions)
        assert torch.allclose(lp, first.old_log_prob, atol=1e-4)
        logs = ppo.ppo_update(pol, opt, buf, n_epochs=2, batch_size=1024)
        assert len(logs) == 2 * (n * T // 1024) and all(np.isfinite(list(l.values())).all() for l in logs)
        assert abs(logs[0]["approx_kl"]) < 1e-5
        assert any(not torch.equal(b, p.detach()) for b, p in zip(before, pol.parameters()))
    u.close()
