"""The factored action masks carry the same information as the dense ones: dense masks rebuilt from
(owned bits, discovered-node count, cached-credential count, discovery order, static presence table) equal the dense masks
of a batch that played the same actions.  CPU: on the oracle; GPU: on the CUDA library (Chain-10, and the factored kernel
against the dense one)."""
import numpy as np
import pytest

from marlon_b200 import _abi, config, masks, scenario, scenarios


def _cfgs(**kw):
    # The defender never ends an episode on its own here: its auto-reset is applied AFTER the step's observations were taken
    # (DummyVecEnv order), so the state digest used below for the discovery order would run ahead of the attacker's
    # observation.  Attacker-side endings (time limit) rewrite the observation together with the state.
    base = dict(maximum_node_count=12, maximum_total_credentials=12, throws_on_invalid_actions=False,
                defender_constraint=config.DefenderConstraint(0.0), losing_reward=-5000.0, defender_enabled=True,
                defender_max_timesteps=10 ** 6, defender_reset_on_constraint_broken=False, attacker_max_timesteps=150)
    base.update(kw)
    return (config.make_config(_abi.MODE_MARLON, mask_mode=_abi.MASK_DENSE, **base),
            config.make_config(_abi.MODE_MARLON, mask_mode=_abi.MASK_FACTORED, **base))


def _valid_actions(arr_dense, rng, n, cfg, comp):
    """attacker actions drawn from the dense masks (so both success and repeat branches occur), uniform defender actions"""
    lay = config.attacker_action_layout(cfg)
    att = np.zeros((n, 10), dtype=np.int32)
    for e in range(n):
        kind = int(rng.integers(0, 3))
        m = {0: arr_dense["local_vulnerability"], 1: arr_dense["remote_vulnerability"], 2: arr_dense["connect"]}[kind][e]
        idx = np.argwhere(m)
        if len(idx) == 0:
            kind, idx = 0, np.argwhere(arr_dense["local_vulnerability"][e])
        pick = idx[rng.integers(0, len(idx))]
        a0, a1 = lay[kind]
        att[e, 0] = [k for k in range(3) if cfg.kind_of_index[k] == kind][0]
        att[e, a0:a0 + len(pick)] = pick
    nn = comp.n_nodes
    dfn = (rng.random((n, 12)) * np.array([5, nn, nn, 6, 2, nn, 6, 2, nn, 3, nn, 3])).astype(np.int32)
    return att, dfn


def _check(dense_get, fact_get, fact_export, comp, cfg_f, step):
    N, C = cfg_f.maximum_node_count, cfg_f.maximum_total_credentials
    ident = comp.identifiers
    sc = fact_get("scalars")
    order = fact_export()[:, _abi.X_HEADER_WORDS: _abi.X_HEADER_WORDS + comp.n_nodes]
    loc, rem, con = masks.dense_masks_from_factored(fact_get("owned_bits"), sc[:, 6], sc[:, 5], order, masks.local_presence_table(comp),
                                                    N, len(ident.remote_vulnerabilities), len(ident.ports), C)
    assert np.array_equal(loc, dense_get("local_vulnerability")), step
    assert np.array_equal(rem, dense_get("remote_vulnerability")), step
    assert np.array_equal(con, dense_get("connect")), step
    for k in ("scalars", "leaked_credentials", "credential_cache_matrix", "discovered_nodes_properties", "nodes_privilegelevel", "att_reward"):
        assert np.array_equal(fact_get(k), dense_get(k)), (step, k)


def test_dense_masks_are_a_function_of_the_factored_observation_oracle():
    from oracle import OracleBatch

    comp = scenario.compile_scenario(scenarios.chain_environment(10))
    cfg_d, cfg_f = _cfgs()
    n = 48
    od, of = OracleBatch(comp, cfg_d, n), OracleBatch(comp, cfg_f, n)
    od.reset(); of.reset()
    rng = np.random.default_rng(3)
    progressed, seen_connect = 0, False
    _check(lambda k: od.arrays[k], lambda k: of.arrays[k], of.export_state, comp, cfg_f, -1)
    for s in range(200):
        att, dfn = _valid_actions(od.arrays, rng, n, cfg_d, comp)
        od.step(att, dfn); of.step(att, dfn)
        _check(lambda k: od.arrays[k], lambda k: of.arrays[k], of.export_state, comp, cfg_f, s)
        progressed = max(progressed, int(od.arrays["scalars"][:, 6].max()))
        seen_connect = seen_connect or bool(od.arrays["connect"].any())
    progressed = max(progressed, int(od.arrays["scalars"][:, 6].max()))
    assert progressed >= 3 and seen_connect  # the game actually progressed: nodes discovered, credentials usable


@pytest.mark.gpu
def test_dense_masks_are_a_function_of_the_factored_observation_cuda():
    from marlon_b200.batch import Batch

    comp = scenario.compile_scenario(scenarios.chain_environment(10))
    cfg_d, cfg_f = _cfgs()
    n = 200
    bd, bf = Batch(comp, cfg_d, n), Batch(comp, cfg_f, n)
    bd.reset(); bf.reset()
    rng = np.random.default_rng(4)
    for s in range(60):
        att, dfn = _valid_actions({k: bd.numpy(k) for k in ("local_vulnerability", "remote_vulnerability", "connect")}, rng, n, cfg_d, comp)
        bd.step(att, dfn); bf.step(att, dfn)
        if s % 5 == 0 or s == 59:
            _check(bd.numpy, bf.numpy, bf.export_state, comp, cfg_f, s)
    assert bd.kernel_info()["name"] != "" and bd.numpy("connect").any()
    bd.close(); bf.close()
