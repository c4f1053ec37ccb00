"""Host-side logic that needs no GPU: spaces / action layouts, masked-discrete decoding, sharding, statistics and the
world_size-2 all-reduce of the episode-statistics vector over gloo (the only collective of the path)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from marlon_b200 import _abi, config, registry, scenario, scenarios, spaces, universe


def _toy_cfg(**kw):
    return config.make_config(_abi.MODE_MARLON, maximum_node_count=12, maximum_total_credentials=10, **kw)


def test_attacker_action_layout_follows_gymnasium_029_key_order():
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    obs, act = universe.attacker_spaces(_toy_cfg(), comp)
    # spaces.Dict sorts a plain dict (gymnasium 0.29.1): connect, local_vulnerability, remote_vulnerability
    assert act.nvec.tolist() == [3, 12, 12, 7, 10, 12, 3, 12, 12, 8]
    _, act2 = universe.attacker_spaces(_toy_cfg(action_kind_order="insertion"), comp)
    assert act2.nvec.tolist() == [3, 12, 3, 12, 12, 8, 12, 12, 7, 10]  # SURVEY.md 8a (insertion order)
    assert obs.spaces["connect"].shape == (12, 12, 7, 10) and obs.spaces["leaked_credentials"].nvec.tolist() == [2, 10, 12, 7] * 5
    assert list(spaces.Dict({"b": spaces.Discrete(2), "a": spaces.Discrete(2)}).spaces) == ["a", "b"]


def test_defender_spaces():
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    obs, act = universe.defender_spaces(comp)
    assert act.nvec.tolist() == [5, 10, 10, 6, 2, 10, 6, 2, 10, 3, 10, 3]  # defend_wrapper.py:174-195
    assert [obs.spaces[k].n for k in ("infected_nodes", "incoming_firewall_status", "outgoing_firewall_status", "services_status")] == [10, 60, 60, 13]


def test_masked_discrete_layout_and_decode():
    from marlon_b200.wrappers import MaskedDiscreteAttackerWrapper

    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = _toy_cfg()
    obs, act = universe.attacker_spaces(cfg, comp)
    lay = config.attacker_action_layout(cfg)

    class Inner:
        observation_space, action_space = obs, act
        action_subspaces = {i: (_abi.KIND_NAMES[cfg.kind_of_index[i]],) + lay[cfg.kind_of_index[i]] for i in range(3)}
        _last_transformed_observation = None

        def step(self, a):
            return a

    m = MaskedDiscreteAttackerWrapper(Inner())
    assert m.action_space.n == 10080 + 36 + 1152 == 11268
    assert m._decode(0) == ("connect", (0, 0, 0, 0))
    assert m._decode(10079) == ("connect", (11, 11, 6, 9))
    assert m._decode(10080 + 3 * 5 + 2) == ("local_vulnerability", (5, 2))
    assert m._decode(10080 + 36 + (7 * 12 + 4) * 8 + 3) == ("remote_vulnerability", (7, 4, 3))
    enc = m.step(10080 + 3 * 5 + 2)
    assert enc[0] == 1 and enc[5:7].tolist() == [5, 2]
    assert m.action_masks().all() and m.action_masks().shape == (11268,)
    with pytest.raises(ValueError):
        m._decode(11268)


def test_shard_range_is_a_partition():
    for total, world in [(65536, 8), (1_000_000, 8), (10, 3), (7, 8)]:
        ranges = [universe.shard_range(total, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1


def test_summarize_stats():
    s = np.zeros(_abi.STAT_COUNT)
    s[_abi.STAT_EPISODES], s[_abi.STAT_ATT_RETURN], s[_abi.STAT_ATT_RETURN_SQ] = 4, 20.0, 120.0
    s[_abi.STAT_EP_LEN], s[_abi.STAT_EP_LEN_SQ] = 40, 500
    d = universe.summarize_stats(s)
    assert d["attacker_return_mean"] == 5.0 and abs(d["attacker_return_std"] - np.sqrt(5.0)) < 1e-12
    assert d["episode_length_mean"] == 10.0 and abs(d["episode_length_std"] - 5.0) < 1e-12


def test_registry_merges_kwargs_like_gym_make():
    env, kw = registry.resolve("CyberBattleChain-v0", size=10, maximum_node_count=12)
    assert len(list(env.network.nodes)) == 12 and kw["maximum_node_count"] == 12 and kw["winning_reward"] == 5000.0
    _, kw = registry.resolve("CyberBattleToyCtf-v0")
    assert kw["attacker_goal"].own_atleast == 6 and kw["attacker_goal"].own_atleast_percent == 1.0  # __init__.py:38
    with pytest.raises(KeyError):
        registry.resolve("NoSuchEnv-v0")


def test_make_config_rejects_unsupported_builtin_defender():
    class ExternalRandomEvents(config.DefenderAgent):
        pass

    with pytest.raises(NotImplementedError):
        config.make_config(defender_agent=ExternalRandomEvents())
    c = config.make_config(defender_agent=config.ScanAndReimageCompromisedMachines(0.6, 2, 5), env_index_base=1 << 33)
    assert (c.builtin_defender, c.scan_capacity, c.scan_frequency, c.env_index_base) == (1, 2, 5, 1 << 33)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from marlon_b200 import _abi, universe
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
lo, hi = universe.shard_range(1001, rank, 2)
v = torch.zeros(_abi.STAT_COUNT, dtype=torch.float64)
v[_abi.STAT_EPISODES] = hi - lo          # pretend every env of the shard finished one episode of return = env index
idx = torch.arange(lo, hi, dtype=torch.float64)
v[_abi.STAT_ATT_RETURN] = idx.sum(); v[_abi.STAT_ATT_RETURN_SQ] = (idx * idx).sum(); v[_abi.STAT_ENV_STEPS] = 7 * (hi - lo)
out = universe.summarize_stats(universe.all_reduce_stats(v).numpy())
ref = np.arange(1001, dtype=np.float64)
assert out["episodes"] == 1001 and out["env_steps"] == 7007
assert abs(out["attacker_return_mean"] - ref.mean()) < 1e-9 and abs(out["attacker_return_std"] - ref.std()) < 1e-6
dist.destroy_process_group()
print("ok", rank)
'''


def test_stats_all_reduce_world_size_2_gloo(tmp_path):
    """N>1 path on CPU: two ranks shard 1001 envs, all-reduce the statistics vector over gloo, both see global stats."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), root, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_encode_action():
    from marlon_b200.cyberbattle_env import CyberBattleEnv

    a = CyberBattleEnv.encode_action({"connect": np.array([1, 2, 3, 4])})
    assert a.tolist() == [[2, 1, 2, 3, 4]]
    assert CyberBattleEnv.encode_action({"local_vulnerability": np.array([5, 1])}).tolist() == [[0, 5, 1, 0, 0]]
    with pytest.raises(AssertionError):
        CyberBattleEnv.encode_action({"connect": [0, 0, 0, 0], "local_vulnerability": [0, 0]})


def test_wrappers_construct_without_a_gpu_and_expose_reference_spaces():
    """Constructors build spaces only; the device batch is created on first reset/step (which needs a GPU)."""
    from marlon_b200 import cyberbattle_env as cbe
    from marlon_b200.wrappers import AttackerEnvWrapper, DefenderEnvWrapper, EnvironmentEventSource, MaskedDiscreteAttackerWrapper

    env = cbe.make("CyberBattleToyCtf-v0", maximum_node_count=12, maximum_total_credentials=10, throws_on_invalid_actions=False)
    assert list(env.action_space.spaces) == ["connect", "local_vulnerability", "remote_vulnerability"]
    assert env.action_space.spaces["connect"].nvec.tolist() == [12, 12, 7, 10]
    assert env.bounds.property_count == 10 and env.identifiers.ports[0] == "GIT" and env.name == "CyberBattleEnv"
    es = EnvironmentEventSource()
    att = AttackerEnvWrapper(env, es)
    dfn = DefenderEnvWrapper(env, att, es, defender=True)
    assert len(es.observers) == 2
    assert att.action_subspaces == {0: ("connect", 1, 5), 1: ("local_vulnerability", 5, 7), 2: ("remote_vulnerability", 7, 10)}
    assert att.max_timesteps == 2000 and dfn.max_timesteps == 100 and dfn.num_services == 13
    assert MaskedDiscreteAttackerWrapper(att).action_space.n == 11268
    assert cbe.make("CyberBattleChain-v0", size=10, maximum_node_count=12, maximum_total_credentials=12).name == "CyberBattleChain-10"
    with pytest.raises(ValueError, match="exceeds the specified limit"):
        cbe.make("CyberBattleChain-v0", size=100)  # 102 nodes > default maximum_node_count=100 (cyberbattle_env.py:417-418)


def test_narrow_actions_checks_the_range_and_fills_caller_buffers():
    """int64 policy output -> int16 elements for the host-buffer step (cbx_batch_step_host_i16)."""
    import numpy as np
    import pytest

    from marlon_b200.batch import narrow_actions

    a = np.array([[2, 11, 9, 6, 9, 0, 2, 0, 0, 7], [-1, 0, 0, 0, 0, 0, 0, 0, 0, 999]], dtype=np.int64)
    n = narrow_actions(a)
    assert n.dtype == np.int16 and n.flags["C_CONTIGUOUS"] and np.array_equal(n, a)
    out = np.full(a.shape, 7, dtype=np.int16)
    assert narrow_actions(a[:, ::1], out=out) is out and np.array_equal(out, a)
    with pytest.raises(OverflowError):
        narrow_actions(np.array([[40000]], dtype=np.int64))
    with pytest.raises(OverflowError):
        narrow_actions(np.array([[-40000]], dtype=np.int32))
    with pytest.raises(TypeError):
        narrow_actions(np.array([[1.0]]))
    with pytest.raises(ValueError):
        narrow_actions(a, out=np.zeros((2, 9), dtype=np.int16))
    assert narrow_actions(np.zeros((0, 10), dtype=np.int64)).shape == (0, 10)
