"""GPU tests of the reference-shaped Python surface: CyberBattleEnv, the MARLon wrappers (driven exactly like the
golden-tape generator drove the reference's), the batched universe / VecEnv adapter, half steps and sharding."""
import numpy as np
import pytest

import helpers
from marlon_b200 import _abi, config, scenario, scenarios

pytestmark = pytest.mark.gpu

SCALAR_KEYS = ["newly_discovered_nodes_count", "lateral_move", "customer_data_found", "probe_result", "escalation",
               "credential_cache_length", "discovered_node_count"]


def _check_cyber_obs(obs, z, s, t=0):
    assert [int(obs[k]) for k in SCALAR_KEYS] == z["scalars"][s, t, :7].tolist()
    assert np.array_equal(np.concatenate(obs["leaked_credentials"]), z["leaked"][s, t])
    assert np.array_equal(np.concatenate(obs["credential_cache_matrix"]), z["cachem"][s, t])
    assert np.array_equal(obs["discovered_nodes_properties"].reshape(-1), z["props"][s, t].astype(np.int32))
    assert np.array_equal(obs["nodes_privilegelevel"], z["priv"][s, t].astype(np.int32))
    am = obs["action_mask"]
    assert np.array_equal(am["local_vulnerability"], z["local"][s, t])
    assert helpers.crc_rows(am["remote_vulnerability"][None])[0] == z["remote_crc"][s, t]
    assert helpers.crc_rows(am["connect"][None])[0] == z["connect_crc"][s, t]


def test_cyberbattle_env_replays_the_reference_chain10_fixture():
    """cyberbattle_env_test.py:43-114 through the gym-shaped CyberBattleEnv: same observations, rewards, done flag,
    RuntimeError on the step after done."""
    from marlon_b200 import cyberbattle_env as cbe

    meta, z = helpers.load_tape("raw_chain10_fixture")
    env = cbe.make("CyberBattleChain-v0", size=10, maximum_node_count=12, maximum_total_credentials=12,
                   attacker_goal=cbe.AttackerGoal(own_atleast_percent=1.0))
    assert env.name == "CyberBattleChain-10" and env.bounds.maximum_node_count == 12
    obs, info = env.reset()
    assert obs["discovered_node_count"] == 1 and info["step_count"] == 0
    kinds = {0: "local_vulnerability", 1: "remote_vulnerability", 2: "connect"}
    width = {0: 2, 1: 3, 2: 4}
    for s in range(meta["steps"]):
        a = z["action"][s, 0]
        action = {kinds[int(a[0])]: a[1:1 + width[int(a[0])]]}
        assert env.is_action_valid(action)
        obs, reward, done, truncated, info = env.step(action)
        assert reward == z["reward"][s, 0] and done == bool(z["terminated"][s, 0]) and truncated is False
        assert info["step_count"] == z["stepcount"][s, 0]
        _check_cyber_obs(obs, z, s)
    assert done and reward == 5000.0
    with pytest.raises(RuntimeError, match=r"new episode must be started with env\.reset\(\)"):
        env.step({"connect": np.array([10, 5, 2, 4])})
    env.reset()
    a = env.sample_valid_action()
    assert env.apply_mask(a)
    env.close()


def test_cyberbattle_env_invalid_actions():
    from marlon_b200 import cyberbattle_env as cbe

    env = cbe.make("CyberBattleToyCtf-v0", maximum_node_count=12, maximum_total_credentials=10)  # throws_on_invalid_actions=True
    env.reset()
    obs, r, done, _, _ = env.step({"local_vulnerability": np.array([5, 0])})  # OutOfBoundIndexError swallowed: blank observation
    assert r == 0.0 and not done and (obs["discovered_nodes_properties"] == 2).all() and not obs["action_mask"]["connect"].any()
    assert obs["discovered_node_count"] == 1
    with pytest.raises(cbe.OutOfBoundIndexError):
        env.is_node_owned(7)
    env.step({"local_vulnerability": np.array([0, 2])})  # SearchEdgeHistory: discovers Website
    with pytest.raises(ValueError, match="does not owned"):
        env.step({"local_vulnerability": np.array([1, 0])})  # Website is discovered but not owned
    env.close()


@pytest.mark.parametrize("name", ["marlon_toyctf_short", "marlon_chain10_valid"])
def test_wrappers_replay_reference_tape(name):
    """AttackerEnvWrapper / DefenderEnvWrapper objects driven with the DummyVecEnv protocol the tape was recorded with."""
    from marlon_b200 import cyberbattle_env as cbe
    from marlon_b200.wrappers import AttackerEnvWrapper, DefenderEnvWrapper, EnvironmentEventSource

    meta, z = helpers.load_tape(name)
    kw = helpers._decode_kwargs(meta["env_kwargs"])
    env = cbe.make(meta["env_id"], **kw)
    es = EnvironmentEventSource()
    att = AttackerEnvWrapper(env, es, **meta["att_kwargs"])
    dfn = DefenderEnvWrapper(env, att, es, defender=True, **meta["def_kwargs"])
    assert att.action_space.nvec.tolist() == meta["att_nvec"] and dfn.action_space.nvec.tolist() == meta["def_nvec"]
    aobs, _ = att.reset()
    dobs, _ = dfn.reset()
    t = 0  # tape 0 of the file
    for s in range(min(meta["steps"], 300)):
        aobs, ar, aterm, atrunc, ainfo = att.step(z["att_action"][s, t])
        assert abs(ar - z["att_reward"][s, t]) <= 1e-6 * max(1, abs(z["att_reward"][s, t])), (s, ar)
        assert (aterm, atrunc) == (bool(z["att_terminated"][s, t]), bool(z["att_truncated"][s, t])), s
        assert bool(ainfo.get("invalid_action", False)) == bool(z["intercepted"][s, t])
        if aterm or atrunc:
            aobs, _ = att.reset()
        assert [aobs[k] for k in SCALAR_KEYS] == z["scalars"][s, t, :7].tolist(), s
        assert np.array_equal(aobs["discovered_nodes_properties"], z["props"][s, t].astype(np.int32))
        assert np.array_equal(aobs["local_vulnerability"], z["local"][s, t])
        assert helpers.crc_rows(aobs["connect"][None])[0] == z["connect_crc"][s, t]
        d_act = z["def_action"][s, t]
        dobs, dr, dterm, dtrunc, _ = dfn.step([] if d_act[0] < 0 else d_act)
        assert abs(dr - z["def_reward"][s, t]) <= 1e-6 * max(1, abs(z["def_reward"][s, t])), (s, dr)
        assert (dterm, dtrunc) == (bool(z["def_terminated"][s, t]), bool(z["def_truncated"][s, t])), s
        if dterm or dtrunc:
            dobs, _ = dfn.reset()
        assert np.array_equal(dobs["infected_nodes"], z["infected"][s, t])
        assert np.array_equal(dobs["incoming_firewall_status"], z["fw_in"][s, t])
        assert att.valid_action_count == z["digest"][s, t, 9] and att.invalid_action_count == z["digest"][s, t, 10]
        assert att.reset_request == bool(z["digest"][s, t, 6]) and dfn.reset_request == bool(z["digest"][s, t, 7])
    att.close()


def test_masked_discrete_wrapper_replays_reference_tape():
    """MaskedDiscreteAttackerWrapper (action_masking.py:30-165) against a tape recorded through the REFERENCE's wrapper: the
    CONTENT of action_masks() every step (concat order connect, local, remote: CRC + count of valid actions) and the
    Discrete -> MultiDiscrete decoding of every action that was played."""
    import zlib

    from marlon_b200 import cyberbattle_env as cbe
    from marlon_b200.wrappers import AttackerEnvWrapper, DefenderEnvWrapper, EnvironmentEventSource, MaskedDiscreteAttackerWrapper

    meta, z = helpers.load_tape("marlon_toyctf_masked")
    env = cbe.make(meta["env_id"], **helpers._decode_kwargs(meta["env_kwargs"]))
    es = EnvironmentEventSource()
    att = AttackerEnvWrapper(env, es, **meta["att_kwargs"])
    dfn = DefenderEnvWrapper(env, att, es, defender=True, **meta["def_kwargs"])
    mw = MaskedDiscreteAttackerWrapper(att)
    assert mw.action_space.n == 12 * 12 * 7 * 10 + 12 * 3 + 12 * 12 * 8
    mw.reset()
    dfn.reset()
    t = 0
    for s in range(meta["steps"]):
        mask = mw.action_masks()
        assert mask.dtype == np.bool_ and mask.shape == (11268,)
        assert int(mask.sum()) == int(z["mask_count"][s, t]), s
        assert (zlib.crc32(mask.astype(np.int8).tobytes()) & 0xFFFFFFFF) == int(z["mask_crc"][s, t]), s
        a = int(z["att_discrete"][s, t])
        assert np.array_equal(mw._encode_for_inner_env(*mw._decode(a)), z["att_action"][s, t]), s
        aobs, ar, aterm, atrunc, _ = mw.step(np.array(a))
        assert abs(ar - z["att_reward"][s, t]) <= 1e-6 * max(1, abs(z["att_reward"][s, t])), s
        assert (aterm, atrunc) == (bool(z["att_terminated"][s, t]), bool(z["att_truncated"][s, t])), s
        if aterm or atrunc:
            mw.reset()
        d_act = z["def_action"][s, t]
        _, _, dterm, dtrunc, _ = dfn.step([] if d_act[0] < 0 else d_act)
        if dterm or dtrunc:
            dfn.reset()
    att.close()


def test_batched_action_masks_match_reference_tape():
    """The batched face of the same wrapper (universe.action_masks(): bool [n, 11268] built on the device from the dense masks)
    against the reference-recorded mask CRCs, all tapes of the file side by side in one batch."""
    import zlib

    from marlon_b200.batch import Batch

    meta, z = helpers.load_tape("marlon_toyctf_masked")
    comp, cfg = helpers.config_from_meta(meta)
    n = meta["n_tapes"]
    b = Batch(comp, cfg, n)
    b.reset()
    import torch

    for s in range(meta["steps"]):
        t = b.tensors
        m = torch.cat([t["connect"].reshape(n, -1), t["local_vulnerability"].reshape(n, -1), t["remote_vulnerability"].reshape(n, -1)],
                      dim=1).to(torch.bool).cpu().numpy()
        for e in range(n):
            assert int(m[e].sum()) == int(z["mask_count"][s, e]), (s, e)
            assert (zlib.crc32(m[e].astype(np.int8).tobytes()) & 0xFFFFFFFF) == int(z["mask_crc"][s, e]), (s, e)
        b.step(z["att_action"][s], z["def_action"][s])
    b.close()


def test_half_steps_and_notify_match_oracle():
    from marlon_b200.batch import Batch
    from oracle import OracleBatch

    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = config.make_config(_abi.MODE_MARLON, maximum_node_count=12, maximum_total_credentials=10, throws_on_invalid_actions=False,
                             attacker_goal=config.AttackerGoal(own_atleast=6), defender_constraint=config.DefenderConstraint(0.6),
                             losing_reward=-5000.0, defender_enabled=True, attacker_max_timesteps=40, defender_max_timesteps=30,
                             auto_reset=False)
    n = 300
    b, o = Batch(comp, cfg, n), OracleBatch(comp, cfg, n)
    rng = np.random.default_rng(5)
    b.reset(); o.reset()
    for s in range(120):
        att, dfn = b.sample_actions(seed=9)
        att, dfn = att.cpu().numpy(), dfn.cpu().numpy()
        b.step(att, None, who=1); o.step(att, None, who=1)
        adone = (o.arrays["att_terminated"] | o.arrays["att_truncated"]).astype(bool)
        if adone.any():
            b.reset(mask=adone, who=1); o.reset(mask=adone.astype(np.uint8), who=1)
        b.step(None, dfn, who=2); o.step(None, dfn, who=2)
        ddone = (o.arrays["def_terminated"] | o.arrays["def_truncated"]).astype(bool)
        if ddone.any():
            b.reset(mask=ddone, who=2); o.reset(mask=ddone.astype(np.uint8), who=2)
        if s % 17 == 3:  # an outside notify_reset, as marl_algorithm.run_episode does for the defender
            m = rng.random(n) < 0.1
            b.notify_reset(2, 3.0, mask=m); o.notify_reset(2, 3.0, mask=m)
        for k, t in b.tensors.items():
            if not k.startswith("term_"):
                assert np.array_equal(t.cpu().numpy(), o.arrays[k]), (s, k)
        assert np.array_equal(b.export_state(), o.export_state()), s
    b.close()


def test_universe_and_vec_env_contract():
    from marlon_b200.universe import MultiAgentUniversalEnv
    from oracle import OracleBatch

    n = 96
    u = MultiAgentUniversalEnv("CyberBattleToyCtf-v0", n, maximum_node_count=12, maximum_total_credentials=10,
                               max_timesteps=25, emit_terminal_obs=True)
    o = OracleBatch(u.compiled, u.cfg, n)
    o.reset()
    assert u.attacker_vec_env is u.attacker_vec_env  # one adapter per role: the episode accumulators live on it
    av, dv = u.vec_env("attacker", observations="numpy"), u.vec_env("defender", observations="numpy")
    assert av.num_envs == n and av.action_space.nvec.tolist() == [3, 12, 12, 7, 10, 12, 3, 12, 12, 8]
    aobs, dobs = av.reset(), dv.reset()
    assert aobs["connect"].shape == (n, 12, 12, 7, 10) and aobs["connect"].dtype == np.int8 and dobs["infected_nodes"].shape == (n, 10)
    episodes = 0
    for s in range(60):
        att, dfn = u.sample_actions(seed=3)
        att, dfn = att.cpu().numpy(), dfn.cpu().numpy()
        aobs, ar, adone, ainfos = av.step(att)
        dobs, dr, ddone, dinfos = dv.step(dfn)
        o.step(att, dfn)
        assert np.array_equal(ar, o.arrays["att_reward"]) and np.array_equal(dr, o.arrays["def_reward"])
        assert np.array_equal(adone, (o.arrays["att_terminated"] | o.arrays["att_truncated"]).astype(bool))
        assert np.array_equal(aobs["connect"], o.arrays["connect"]) and np.array_equal(dobs["infected_nodes"], o.arrays["def_infected_nodes"])
        for i in np.nonzero(adone)[0]:
            info = ainfos[i]
            episodes += 1
            assert set(info) >= {"TimeLimit.truncated", "episode", "terminal_observation"}
            assert info["TimeLimit.truncated"] == bool(o.arrays["att_truncated"][i] and not o.arrays["att_terminated"][i])
            assert info["episode"]["l"] >= 1 and info["terminal_observation"]["connect"].shape == (12, 12, 7, 10)
            assert np.array_equal(info["terminal_observation"]["discovered_nodes_properties"], o.arrays["term_discovered_nodes_properties"][i])
            assert np.array_equal(info["terminal_observation"]["connect"], o.arrays["term_connect"][i])
        for i in np.nonzero(ddone)[0]:
            assert np.array_equal(dinfos[i]["terminal_observation"]["infected_nodes"], o.arrays["term_def_infected_nodes"][i])
        assert all(ainfos[i] == {} for i in np.nonzero(~adone)[0])
        # the device-tensor adapter of the same universe sees the same observation without any copy
        tobs = u.attacker_observation()
        assert tobs["connect"].data_ptr() == u.batch.tensors["connect"].data_ptr()
        assert np.array_equal(tobs["discovered_node_count"].cpu().numpy(), aobs["discovered_node_count"])
    assert episodes > 0
    masks = av.env_method("action_masks")
    assert len(masks) == n and masks[0].shape == (11268,) and masks[0].dtype == np.bool_
    st = u.episode_statistics(reduce=True)
    assert st["env_steps"] == n * 60 and st["episodes"] == episodes
    u.close()


def test_sharding_does_not_change_the_random_draws():
    """env_index_base: envs [128, 256) of a 256-env run == a separate 128-env batch with base 128 (Philox keyed by global index)."""
    from marlon_b200.batch import Batch

    comp = scenario.compile_scenario(scenarios.toyctf_environment())

    def cfg(base):
        return config.make_config(_abi.MODE_CYBERBATTLE, maximum_node_count=12, maximum_total_credentials=10, throws_on_invalid_actions=False,
                                  attacker_goal=config.AttackerGoal(own_atleast=6), seed=77, env_index_base=base,
                                  defender_agent=config.ScanAndReimageCompromisedMachines(0.9, 3, 2),
                                  defender_constraint=config.DefenderConstraint(0.5), auto_reset=True)

    whole, part = Batch(comp, cfg(0), 256), Batch(comp, cfg(128), 128)
    whole.reset(); part.reset()
    for s in range(80):
        att, _ = whole.sample_actions(seed=1)
        whole.step(att)
        part.step(att[128:].contiguous())
    assert np.array_equal(whole.export_state(128, 256), part.export_state())
    assert np.array_equal(whole.numpy("connect")[128:], part.numpy("connect"))
    assert whole.export_state()[:, 13].sum() > 0  # some nodes are being re-imaged: the defender did act
    whole.close(); part.close()


def test_universe_over_generated_random_networks():
    """configs[4] through the Python face: one batch over several generated CyberBattleRandom networks (factored masks)."""
    from marlon_b200.universe import MultiAgentUniversalEnv

    u = MultiAgentUniversalEnv("CyberBattleRandom-v0", 64, scenario_seeds=[1, 2, 3], maximum_node_count=72, maximum_total_credentials=104,
                               max_timesteps=50, mask_mode="factored")
    assert u.n_envs == 192 and len(u.batch.scenarios) == 3
    n_max = max(c.n_nodes for c in u.batch.scenarios)
    assert u.defender_action_space.nvec.tolist() == [5, n_max, n_max, 6, 2, n_max, 6, 2, n_max, 3, n_max, 3]
    assert u.attacker_action_space.nvec.tolist()[0] == 3
    aobs, dobs = u.reset()
    assert aobs["discovered_nodes_properties"].shape == (192, 72) and dobs["infected_nodes"].shape == (192, n_max)
    assert "owned_bits" in aobs and "connect" not in aobs
    assert (aobs["discovered_node_count"] == 1).all()  # one breach node per generated network (generate_network.py:205-219)
    for s in range(80):
        att, dfn = u.sample_actions(seed=9)
        r = u.step(att, dfn)
    st = u.episode_statistics(reduce=False)
    assert st["env_steps"] == 192 * 80 and st["episodes"] >= 192 and st["att_invalid"] == 0
    assert float(r["attacker_reward"].abs().sum()) >= 0
    u.close()


def _gae_numpy(rewards, values, starts, last_values, last_dones, gamma, lam):
    """stable-baselines3 RolloutBuffer.compute_returns_and_advantage, restated (float32 like SB3's numpy buffers)."""
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    last = np.zeros(rewards.shape[1], dtype=np.float32)
    for t in reversed(range(T)):
        if t == T - 1:
            nnt, nv = 1.0 - last_dones.astype(np.float32), last_values
        else:
            nnt, nv = 1.0 - starts[t + 1].astype(np.float32), values[t + 1]
        delta = rewards[t] + np.float32(gamma) * nv * nnt - values[t]
        last = delta + np.float32(gamma) * np.float32(lam) * nnt * last
        adv[t] = last
    return adv, adv + values


def test_device_rollout_and_gae():
    """collect_rollouts with everything on the device; GAE kernel against the SB3 formula."""
    import torch

    from marlon_b200.rollout import DeviceRolloutBuffer, collect_rollouts
    from marlon_b200.universe import MultiAgentUniversalEnv

    n, T = 300, 24
    u = MultiAgentUniversalEnv("CyberBattleToyCtf-v0", n, maximum_node_count=12, maximum_total_credentials=10, max_timesteps=9)
    u.reset()
    dev = u.batch.torch_device
    calls = {"k": 0}

    def attacker_policy(obs, masks):
        calls["k"] += 1
        att, _ = u.sample_actions(seed=calls["k"])
        value = obs["discovered_node_count"].to(torch.float32) * 0.5 + obs["credential_cache_length"].to(torch.float32)
        return att, value, torch.full((n,), -1.25, device=dev)

    def defender_policy(obs, masks):
        _, dfn = u.sample_actions(seed=1000 + calls["k"])
        return dfn, obs["infected_nodes"].to(torch.float32).sum(dim=1), torch.full((n,), -2.5, device=dev)

    ab = DeviceRolloutBuffer(T, n, 10, dev, gamma=0.99, gae_lambda=0.95, obs_spec={"discovered_node_count": ((), torch.int32)})
    db = DeviceRolloutBuffer(T, n, 12, dev, gamma=0.9, gae_lambda=0.8)
    assert collect_rollouts(u, attacker_policy, ab, defender_policy, db)
    assert ab.full and db.full
    for buf, gamma, lam, last_starts in ((ab, 0.99, 0.95, u._att_starts), (db, 0.9, 0.8, u._def_starts)):
        r, v, s = buf.rewards.cpu().numpy(), buf.values.cpu().numpy(), buf.episode_starts.cpu().numpy()
        assert s[0].all()  # _last_episode_starts right after reset
        adv, ret = buf.advantages.cpu().numpy(), buf.returns.cpu().numpy()
        # the kernel's bootstrap inputs are recomputed here the way collect_rollouts produced them
        if buf is ab:
            o = u.attacker_observation()
            lv = (o["discovered_node_count"].to(torch.float32) * 0.5 + o["credential_cache_length"].to(torch.float32)).cpu().numpy()
        else:
            lv = u.defender_observation()["infected_nodes"].to(torch.float32).sum(dim=1).cpu().numpy()
        want_adv, want_ret = _gae_numpy(r, v, s, lv, last_starts.cpu().numpy(), gamma, lam)
        assert np.allclose(adv, want_adv, rtol=1e-5, atol=1e-3) and np.allclose(ret, want_ret, rtol=1e-5, atol=1e-3)
    # episode starts after step 0 are the done flags of the step before: as many as episodes finished (minus those of the last step)
    st = u.episode_statistics(reduce=False)
    assert int(ab.episode_starts[1:].sum().item()) + int(u._att_starts.sum().item()) == int(st["episodes"])
    assert st["env_steps"] == n * T
    assert (ab.observations["discovered_node_count"][0] == 1).all()  # the observation the first action was chosen on
    assert ab.actions.shape == (T, n, 10) and db.actions.shape == (T, n, 12)
    u.close()
