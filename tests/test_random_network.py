"""CyberBattleRandom networks (configs[4]): the seeded generator reproduces the reference's generator, and the gym id works.

The fingerprints were recorded by ``oracle/gen_golden.py --random`` from the reference's own
``generate_random_traffic_network`` + ``cyberbattle_model_from_traffic_graph`` (generate_network.py:22-263) with
``new_environment``'s parameters and ``random.seed(seed)``."""
import json
import os

import pytest

import helpers
from marlon_b200 import random_network, registry, scenario


def _fps():
    return json.load(open(os.path.join(helpers.GOLDEN, "scenario_fingerprints.json")))


@pytest.mark.parametrize("seed", range(12))
def test_generated_network_tables_match_reference_generator(seed):
    comp = scenario.compile_scenario(random_network.random_environment(seed))
    assert comp.fingerprint() == _fps()[f"CyberBattleRandom-v0:seed={seed}"]


def test_generated_network_shape():
    env = random_network.random_environment(3)
    comp = scenario.compile_scenario(env)
    ident = comp.identifiers
    # P=3 ports, L=3 local, R=1 remote, one property; 65 nodes (50 clients + 15 servers with traffic)
    assert (len(ident.ports), len(ident.local_vulnerabilities), len(ident.remote_vulnerabilities), len(ident.properties)) == (3, 3, 1, 1)
    assert comp.n_nodes == 65
    entries = [k for k, v in env.nodes() if v.agent_installed]
    assert len(entries) == 1 and env.get_node(entries[0]).properties == ["breach_node"] and not env.get_node(entries[0]).reimagable
    # services alias the password lists, which kept growing after the services were created (generate_network.py:222-241)
    for _, info in env.nodes():
        for svc in info.services:
            assert svc.name in ("SMB", "RDP") and len(svc.allowedCredentials) >= 1
    assert random_network.random_environment(3).network.nodes._n.keys() == env.network.nodes._n.keys()  # deterministic
    assert scenario.compile_scenario(random_network.random_environment(4)).fingerprint() != comp.fingerprint()


def test_registry_resolves_cyberbattle_random():
    env, kw = registry.resolve("CyberBattleRandom-v0", seed=5, maximum_node_count=72)
    assert kw["maximum_discoverable_credentials_per_action"] == 32 and kw["maximum_node_count"] == 72  # cyberbattle_random.py:14
    assert scenario.compile_scenario(env).fingerprint() == _fps()["CyberBattleRandom-v0:seed=5"]


@pytest.mark.reference
def test_live_reference_generator_agrees():
    import random

    import numpy as np
    import ref_loader

    ref_loader.load()
    from cyberbattle.simulation import generate_network as gen
    from cyberbattle.simulation import model as ref_model

    seed = 13
    random.seed(seed)
    traffic = gen.generate_random_traffic_network(seed=seed, n_clients=50, n_servers={"SMB": 15, "HTTP": 15, "RDP": 15},
                                                  alpha=np.array([(1, 1), (0.2, 0.5)], dtype=float),
                                                  beta=np.array([(1000, 10), (10, 100)], dtype=float))
    net = gen.cyberbattle_model_from_traffic_graph(traffic, cached_rdp_password_probability=0.8, cached_smb_password_probability=0.7,
                                                   cached_accessed_network_shares_probability=0.8,
                                                   cached_password_has_changed_probability=0.01,
                                                   probability_two_nodes_use_same_password_to_access_given_resource=0.9)
    ref = ref_model.Environment(network=net, vulnerability_library={}, identifiers=gen.ENV_IDENTIFIERS)
    assert scenario.compile_scenario(ref).blob.tolist() == scenario.compile_scenario(random_network.random_environment(seed)).blob.tolist()
