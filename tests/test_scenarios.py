"""Scenario tables: this package's scenario definitions compile to the same bytes as the reference's own objects."""
import json
import os

import pytest

import helpers
from marlon_b200 import scenario, scenarios


def _fps():
    return json.load(open(os.path.join(helpers.GOLDEN, "scenario_fingerprints.json")))


def test_toyctf_tables_match_reference_fingerprint():
    assert scenario.compile_scenario(scenarios.toyctf_environment()).fingerprint() == _fps()["CyberBattleToyCtf-v0"]


@pytest.mark.parametrize("size", [4, 10, 100])
def test_chain_tables_match_reference_fingerprint(size):
    assert scenario.compile_scenario(scenarios.chain_environment(size)).fingerprint() == _fps()[f"CyberBattleChain-v0:size={size}"]


def test_toyctf_dimensions():
    c = scenario.compile_scenario(scenarios.toyctf_environment())
    ident = c.identifiers
    # P=7 ports, L=3, R=8, props=10; 10 nodes; 13 services
    assert (len(ident.ports), len(ident.local_vulnerabilities), len(ident.remote_vulnerabilities), len(ident.properties)) == (7, 3, 8, 10)
    assert c.n_nodes == 10 and c.n_services == 13 and len(c.triples) == 5
    assert c.node_ids[0] == "Website" and c.node_ids[-1] == "client"
    # Website.incoming and Website[user=monitor].outgoing are one list object
    assert [("Website", "incoming"), ("Website[user=monitor]", "outgoing")] in c.alias_groups.values()


def test_chain_node_order():
    c = scenario.compile_scenario(scenarios.chain_environment(10))
    assert c.node_ids[:4] == ["start", "11_LinuxNode", "1_LinuxNode", "2_WindowsNode"]
    with pytest.raises(ValueError):
        scenarios.chain_environment(3)


def test_precondition_parser():
    from marlon_b200.model import Precondition

    # actions_test.py:410-422 of the reference
    p = Precondition("Windows&Win10&(~(privilege_2|privilege_3))")
    assert p.evaluate(["Windows", "Win10"]) and not p.evaluate(["Windows", "Win10", "privilege_2"])
    assert not p.evaluate(["Windows"]) and Precondition("true").evaluate([]) and not Precondition("false").evaluate([])
    assert Precondition("SasUrlInCommit&GitHub").evaluate(["GitHub", "SasUrlInCommit"])


@pytest.mark.reference
def test_tables_match_live_reference_objects():
    import ref_loader

    ref_loader.load()
    from cyberbattle.samples.chainpattern import chainpattern
    from cyberbattle.samples.toyctf import toy_ctf

    assert scenario.compile_scenario(toy_ctf.new_environment()).blob.tolist() == scenario.compile_scenario(scenarios.toyctf_environment()).blob.tolist()
    for size in (4, 10):
        assert scenario.compile_scenario(chainpattern.new_environment(size)).blob.tolist() == scenario.compile_scenario(scenarios.chain_environment(size)).blob.tolist()
