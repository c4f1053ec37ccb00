"""The oracle's restatement of ``CyberBattleEnv.sample_valid_action`` (ENV:959-1047; rejection sampling over whole proposals,
quirk B.9: kind 1 builds a LOCAL action, kind 0 a REMOTE one) against the live reference: same frequencies of every action, on
states with and without cached credentials.  The CUDA sampler is compared bit for bit with the oracle's on the GPU
(tests/test_gpu_parity.py::test_device_sampler_equals_oracle_sampler), which closes the chain reference -> oracle -> device."""
import collections

import numpy as np
import pytest

from marlon_b200 import _abi, config, scenario, scenarios

pytestmark = pytest.mark.reference

DRAWS = 40000


def _chi2_ok(ref_counts, orc_counts):
    """Two-sample chi-square over the union of categories; cells pooled below 5 expected counts."""
    keys = sorted(set(ref_counts) | set(orc_counts))
    a = np.array([ref_counts.get(k, 0) for k in keys], dtype=np.float64)
    b = np.array([orc_counts.get(k, 0) for k in keys], dtype=np.float64)
    tot = a + b
    small = tot < 10
    if small.any():
        a = np.append(a[~small], a[small].sum())
        b = np.append(b[~small], b[small].sum())
        tot = a + b
    k1, k2 = np.sqrt(b.sum() / a.sum()), np.sqrt(a.sum() / b.sum())
    chi2 = float((((k1 * a - k2 * b) ** 2) / np.maximum(tot, 1)).sum())
    dof = len(tot) - 1
    # mean dof, variance 2 dof: six sigma keeps the test deterministic-in-practice yet sharp (a swapped kind, a different
    # owned-node set or per-kind instead of per-proposal rejection moves chi2 by thousands)
    return chi2, dof, chi2 < dof + 6.0 * np.sqrt(2.0 * dof) + 10.0


@pytest.mark.parametrize("prefix_steps", [0, 25])
def test_oracle_sampler_has_the_reference_distribution(prefix_steps):
    import ref_loader
    from oracle import OracleBatch

    ref_loader.load()
    env = ref_loader.make("CyberBattleToyCtf-v0", maximum_node_count=12, maximum_total_credentials=10, throws_on_invalid_actions=False)
    env.reset(seed=3)
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    cfg = config.make_config(_abi.MODE_CYBERBATTLE, maximum_node_count=12, maximum_total_credentials=10, throws_on_invalid_actions=False,
                             attacker_goal=config.AttackerGoal(own_atleast=6), auto_reset=False)
    o = OracleBatch(comp, cfg, 1)
    o.reset()
    code = {"local_vulnerability": 0, "remote_vulnerability": 1, "connect": 2}
    cache = env._CyberBattleEnv__credential_cache
    steps = 0
    # bring both to the same non-trivial state: at least `prefix_steps` steps and, then, cached credentials and two owned nodes
    while prefix_steps and (steps < prefix_steps or len(cache) < 2 or len(env._CyberBattleEnv__get__owned_nodes_indices()) < 2):
        steps += 1
        assert steps < 3000
        a = env.sample_valid_action(kinds=[0, 1, 2])
        kind = next(iter(a))
        enc = np.zeros((1, 5), dtype=np.int32)
        enc[0, 0] = code[kind]
        enc[0, 1:1 + len(a[kind])] = a[kind]
        env.step(a)
        o.step(enc)
    if prefix_steps:
        cache = env._CyberBattleEnv__credential_cache
        assert len(cache) >= 2 and o.export_state()[0, _abi.X_NAMES.index("n_cached")] == len(cache)
    ref_counts = collections.Counter()
    for _ in range(DRAWS):
        a = env.sample_valid_action(kinds=[0, 1, 2])
        kind = next(iter(a))
        ref_counts[(code[kind],) + tuple(int(x) for x in a[kind])] += 1
    orc_counts = collections.Counter()
    for k in range(DRAWS):
        att, _ = o.sample_actions(seed=99, step=k)
        kind = int(att[0, 0])
        orc_counts[(kind,) + tuple(int(x) for x in att[0, 1:1 + _abi.KIND_WIDTH[kind]])] += 1
    chi2, dof, ok = _chi2_ok(ref_counts, orc_counts)
    assert ok, (chi2, dof)
    # the marginal over kinds on its own (what quirk B.9 and whole-proposal rejection decide)
    rk = np.array([sum(v for k, v in ref_counts.items() if k[0] == q) for q in range(3)]) / DRAWS
    ok_ = np.array([sum(v for k, v in orc_counts.items() if k[0] == q) for q in range(3)]) / DRAWS
    assert np.abs(rk - ok_).max() < 0.012, (rk, ok_)
    if prefix_steps == 0:  # ToyCtf at reset: one owned node with one of three local vulnerabilities, no credentials
        assert abs(rk[0] - 0.25) < 0.012 and rk[2] == 0 and ok_[2] == 0  # (1/2 * 1/3) / (1/6 + 1/2)
