"""Acceptance harness with the REAL callers: the reference's own, unmodified rollout / episode loops
(``marlon/baseline_models/multiagent/marl_algorithm.py:17-54`` ``collect_rollouts``, ``:144-251`` ``run_episode``), its own
``MultiAgentUniverse.build`` factory (``multiagent_universe.py:78-206``) and its own ``RandomMarlonAgent``
(``random_marlon_agent.py:31-140``) are pointed first at the reference's wrappers, then at this package's wrappers, with the
same action stream; every reward and done flag must agree.  That is the "drop-in" claim, executed.

Runs only where ``/root/reference`` exists (dev container; no GPU there): the package's host-side Python is exercised for real,
the step semantics come from the CPU oracle through ``tests/oracle_backed.OracleBackedBatch`` (that the CUDA library
computes the same arrays is the job of the ``-m gpu`` parity tests).  stable-baselines3 is not installable here, so the three
SB3 names the reference's modules import come from ``tests/sb3_stub``.
"""
import logging

import numpy as np
import pytest

pytestmark = pytest.mark.reference

ENV_KW = dict(env_id="CyberBattleToyCtf-v0", max_timesteps=45, maximum_node_count=12, maximum_total_credentials=10,
              maximum_discoverable_credentials_per_action=5)


@pytest.fixture(scope="module")
def ref():
    """The reference's caller modules, imported unmodified."""
    import ref_loader
    import sb3_stub

    ref_loader.load()
    sb3_stub.install()
    logging.disable(logging.CRITICAL)
    from marlon.baseline_models.multiagent import marl_algorithm, multiagent_universe, random_marlon_agent

    yield dict(marl=marl_algorithm, universe=multiagent_universe, random_agent=random_marlon_agent)
    logging.disable(logging.NOTSET)


@pytest.fixture
def oracle_batches(monkeypatch):
    """marlon_b200's host code on the CPU oracle (no GPU in the container the reference lives in)."""
    import marlon_b200.batch
    from oracle_backed import OracleBackedBatch

    monkeypatch.setattr(marlon_b200.batch, "Batch", OracleBackedBatch)


class _SharedSampler:
    """Both stacks draw their random actions from identical streams (the two gym ``MultiDiscrete.sample`` implementations
    differ; what is under test is the environment, not the sampler)."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def __call__(self, agent):
        nvec = np.asarray(agent.env.action_space.nvec)
        return (self.rng.random(len(nvec)) * nvec).astype(np.int64)


def _build(universe_cls, builder_cls, sampler_a, sampler_d, monkeypatch, agent_cls, n_rollout_steps):
    monkeypatch.setattr(agent_cls, "_sample_action", lambda self: (sampler_a if self.role == "attacker" else sampler_d)(self), raising=False)
    u = universe_cls.build(attacker_builder=builder_cls(n_rollout_steps=n_rollout_steps),
                           defender_builder=builder_cls(n_rollout_steps=n_rollout_steps), **ENV_KW)
    u.attacker_agent.role, u.defender_agent.role = "attacker", "defender"
    return u


def _trace(agent):
    """Record what the agent's env returns, step by step (reward, done) and reset by reset."""
    log = []
    env = agent.env
    step, reset = env.step, env.reset

    def traced_step(action):
        out = step(action)
        log.append(("step", float(out[1]), bool(out[2]), bool(out[3])))
        return out

    def traced_reset(**kw):
        log.append(("reset",))
        return reset(**kw)

    env.step, env.reset = traced_step, traced_reset
    return log


def test_reference_collect_rollouts_and_run_episode_drive_both_stacks(ref, oracle_batches, monkeypatch):
    import marlon_b200.universe as ours

    marl, RandomAgent = ref["marl"], ref["random_agent"].RandomMarlonAgent
    Builder = ref["random_agent"].RandomAgentBuilder
    results = {}
    for name, universe_cls in (("reference", ref["universe"].MultiAgentUniverse), ("marlon_b200", ours.MultiAgentUniverse)):
        u = _build(universe_cls, Builder, _SharedSampler(11), _SharedSampler(12), monkeypatch, RandomAgent, n_rollout_steps=400)
        a, d = u.attacker_agent, u.defender_agent
        assert type(a) is RandomAgent and type(a.env).__name__ == "Monitor"  # the reference's agent class around ... whose wrapper?
        assert type(a.wrapper).__module__.startswith("marlon." if name == "reference" else "marlon_b200.")
        la, ld = _trace(a), _trace(d)
        # SB3's learn() resets every env before the first rollout: attacker first, then defender (marl_algorithm.learn)
        a.env.reset()
        d.env.reset()
        a.n_eval_episodes = d.n_eval_episodes = 10 ** 9  # RandomMarlonAgent stops a rollout by episode count otherwise
        assert marl.collect_rollouts(a, d) is True      # the reference's loop, 400 attacker+defender step pairs
        ep_a, ep_d, _ = marl.run_episode(a, d, max_steps=60)  # then the reference's evaluation loop
        results[name] = dict(att=la, dfn=ld, ep_a=[float(x) for x in ep_a], ep_d=[float(x) for x in ep_d],
                             att_episodes=list(a.env.episode_returns), def_episodes=list(d.env.episode_returns))
    r, o = results["reference"], results["marlon_b200"]
    assert len(r["att"]) > 400 and sum(1 for x in r["att"] if x[0] == "reset") > 5  # episodes ended and restarted on the way
    assert o["att"] == r["att"]
    assert o["dfn"] == r["dfn"]
    assert o["ep_a"] == r["ep_a"] and o["ep_d"] == r["ep_d"]
    assert o["att_episodes"] == r["att_episodes"] and o["def_episodes"] == r["def_episodes"]


def test_batched_vec_env_is_an_sb3_vec_env_and_serves_the_reference_agent_loop(ref, oracle_batches):
    """With stable-baselines3 importable (here: the stand-in) ``BatchedVecEnv`` IS a ``VecEnv`` -- what SB3's ``_wrap_env`` tests
    -- and a ``perform_step``-shaped loop (baseline_marlon_agent.py:100-167: ``env.step(actions)`` -> 4-tuple, ``_update_info_buffer``
    reading ``info["episode"]``, ``_last_episode_starts = dones``) over the adapter reproduces, env by env, what the reference's
    wrappers produce under SB3's DummyVecEnv protocol for the same actions."""
    import importlib

    import marlon_b200.vec_env as vec_env_mod
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv

    vec_env_mod = importlib.reload(vec_env_mod)  # pick up the (stand-in) SB3 base class
    from marlon_b200.universe import MultiAgentUniversalEnv

    n, steps = 6, 260
    u = MultiAgentUniversalEnv("CyberBattleToyCtf-v0", n, maximum_node_count=12, maximum_total_credentials=10,
                               maximum_discoverable_credentials_per_action=5, max_timesteps=45, emit_terminal_obs=True)
    av = vec_env_mod.BatchedVecEnv(u, "attacker", observations="numpy")
    dv = vec_env_mod.BatchedVecEnv(u, "defender", observations="numpy")
    assert vec_env_mod.HAVE_SB3 and isinstance(av, VecEnv) and isinstance(dv, VecEnv)
    assert av.num_envs == n and len(av.reset_infos) == n
    # the reference stack, one env at a time under the DummyVecEnv protocol
    ref_u = [ref["universe"].MultiAgentUniverse.build(attacker_builder=ref["random_agent"].RandomAgentBuilder(),
                                                      defender_builder=ref["random_agent"].RandomAgentBuilder(), **ENV_KW)
             for _ in range(n)]
    for ru in ref_u:
        ru.attacker_agent.env.reset()
        ru.defender_agent.env.reset()
    aobs, dobs = av.reset(), dv.reset()
    rng = np.random.default_rng(5)
    a_nvec, d_nvec = np.asarray(av.action_space.nvec), np.asarray(dv.action_space.nvec)
    ep_info_buffer = []
    last_episode_starts = np.ones(n, dtype=bool)
    for s in range(steps):
        a_act = (rng.random((n, len(a_nvec))) * a_nvec).astype(np.int64)
        d_act = (rng.random((n, len(d_nvec))) * d_nvec).astype(np.int64)
        aobs, ar, adone, ainfos = av.step(a_act)
        dobs, dr, ddone, dinfos = dv.step(d_act)
        for info in ainfos:  # BaseAlgorithm._update_info_buffer
            if info.get("episode") is not None:
                ep_info_buffer.append(info["episode"])
        last_episode_starts = adone
        for i, ru in enumerate(ref_u):
            a, d = ru.attacker_agent, ru.defender_agent
            o, r, term, trunc, info = a.env.step(a_act[i])
            assert ar[i] == pytest.approx(r, rel=1e-6) and adone[i] == bool(term or trunc), (s, i)
            if term or trunc:
                assert ainfos[i]["TimeLimit.truncated"] == bool(trunc and not term)
                assert ainfos[i]["episode"]["l"] == info["episode"]["l"] and ainfos[i]["episode"]["r"] == pytest.approx(info["episode"]["r"])
                assert np.array_equal(ainfos[i]["terminal_observation"]["discovered_nodes_properties"], o["discovered_nodes_properties"])
                assert np.array_equal(ainfos[i]["terminal_observation"]["connect"], o["connect"])
                o, _ = a.env.reset()
            assert np.array_equal(aobs["discovered_nodes_properties"][i], o["discovered_nodes_properties"]), (s, i)
            assert np.array_equal(aobs["connect"][i], o["connect"]) and aobs["discovered_node_count"][i] == o["discovered_node_count"]
            o2, r2, term2, trunc2, _ = d.env.step(d_act[i])
            assert dr[i] == pytest.approx(r2, rel=1e-6) and ddone[i] == bool(term2 or trunc2), (s, i)
            if term2 or trunc2:
                o2, _ = d.env.reset()
            assert np.array_equal(dobs["infected_nodes"][i], o2["infected_nodes"]), (s, i)
    assert len(ep_info_buffer) > 10 and last_episode_starts.shape == (n,)
    u.close()
