"""Scenario sources: ToyCtf, Chain(size) -- the static data the step kernel plays on.

Scenario *content* (node names, services, vulnerability ids, costs, credentials)
has to be identical to the reference's for trajectories to be bit-exact; it is
restated here in a compact tabular form and pinned against the reference's own
objects by ``tests/test_scenarios.py`` (table fingerprints in
``tests/golden/scenario_fingerprints.json``).  Reference sources:
``cyberbattle/samples/toyctf/toy_ctf.py:22-200`` and
``cyberbattle/samples/chainpattern/chainpattern.py:56-243``.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

from . import model as m

ALLOW, BLOCK = m.RulePermission.ALLOW, m.RulePermission.BLOCK
LOCAL, REMOTE = m.VulnerabilityType.LOCAL, m.VulnerabilityType.REMOTE


def _rules(*spec: Tuple[str, m.RulePermission]) -> List[m.FirewallRule]:
    return [m.FirewallRule(p, perm) for p, perm in spec]


def _std_rules() -> List[m.FirewallRule]:
    return _rules(("RDP", ALLOW), ("SSH", ALLOW), ("HTTPS", ALLOW), ("HTTP", ALLOW))


def _svc(*spec) -> List[m.ListeningService]:
    out = []
    for s in spec:
        name, creds = (s, []) if isinstance(s, str) else (s[0], list(s[1:]))
        out.append(m.ListeningService(name, allowedCredentials=creds))
    return out


def _leak_creds(*triples: Tuple[str, str, str]) -> m.LeakedCredentials:
    return m.LeakedCredentials([m.CachedCredential(*t) for t in triples])


def _vuln(kind, outcome, cost=1.0, precondition="true") -> m.VulnerabilityInfo:
    return m.VulnerabilityInfo(type=kind, outcome=outcome, cost=cost, precondition=m.Precondition(precondition))


# ------------------------------------------------------------------------------------------------
# ToyCtf
# ------------------------------------------------------------------------------------------------
def toyctf_nodes() -> Dict[str, m.NodeInfo]:
    """Ten nodes, order as in the reference.  The `Website` incoming rule list and the
    `Website[user=monitor]` outgoing rule list are ONE list object."""
    shared = _std_rules()
    mon = "Website[user=monitor]"
    arm, arm_mon = "AzureResourceManager", "AzureResourceManager[user=monitor]"
    return {
        "Website": m.NodeInfo(
            services=_svc("HTTPS", ("SSH", "ReusedMySqlCred-web")),
            firewall=m.FirewallConfiguration(incoming=shared, outgoing=shared + _rules(("su", ALLOW), ("sudo", ALLOW))),
            value=100, properties=["MySql", "Ubuntu", "nginx/1.10.3"],
            vulnerabilities={
                "ScanPageContent": _vuln(REMOTE, m.LeakedNodesId(["GitHubProject"])),
                "ScanPageSource": _vuln(REMOTE, m.LeakedNodesId(["Website.Directory"])),
                "CredScanBashHistory": _vuln(LOCAL, _leak_creds((mon, "SSH", "monitorBashCreds"))),
            }),
        "Website.Directory": m.NodeInfo(
            services=_svc("HTTPS"), value=50,
            properties=["Ubuntu", "nginx/1.10.3", "CTFFLAG:Readme.txt-Discover secret data"],
            vulnerabilities={
                "NavigateWebDirectoryFurther": _vuln(REMOTE, _leak_creds(("Website", "MySQL", "ReusedMySqlCred-web"))),
                "NavigateWebDirectory": _vuln(REMOTE, m.LeakedNodesId(["Sharepoint"])),
            }),
        mon: m.NodeInfo(
            services=_svc(("SSH",), ("SSH-key", "unkownkey"), ("su", "monitorBashCreds")),
            value=100, properties=["MySql", "Ubuntu", "nginx/1.10.3"],
            firewall=m.FirewallConfiguration(
                outgoing=shared,
                incoming=_rules(("SSH", BLOCK), ("sudo", BLOCK), ("su", ALLOW)) + shared),
            vulnerabilities={
                "CredScan-HomeDirectory": _vuln(LOCAL, _leak_creds((arm_mon, "HTTPS", "azuread_user_credentials"))),
            }),
        "GitHubProject": m.NodeInfo(
            services=_svc("GIT"), value=10, properties=["GitHub", "SasUrlInCommit"],
            vulnerabilities={
                "CredScanGitHistory": _vuln(REMOTE, _leak_creds(("AzureStorage", "HTTPS", "SASTOKEN1")),
                                            precondition="SasUrlInCommit&GitHub"),
            }),
        "AzureStorage": m.NodeInfo(
            services=_svc(("HTTPS", "SASTOKEN1")), value=50, properties=["CTFFLAG:LeakedCustomerData"],
            vulnerabilities={"AccessDataWithSASToken": _vuln(REMOTE, m.CustomerData())}),
        "Sharepoint": m.NodeInfo(
            services=_svc("HTTPS"), value=100, properties=["SharepointLeakingPassword"],
            firewall=m.FirewallConfiguration(incoming=_rules(("SSH", ALLOW), ("HTTP", ALLOW), ("HTTPS", ALLOW)), outgoing=[]),
            vulnerabilities={
                "ScanSharepointParentDirectory": _vuln(REMOTE, _leak_creds((arm, "HTTPS", "ADPrincipalCreds"))),
            }),
        arm: m.NodeInfo(
            services=_svc(("HTTPS", "ADPrincipalCreds", "azuread_user_credentials")), value=50,
            properties=["CTFFLAG:LeakedCustomerData2"],
            vulnerabilities={"ListAzureResources": _vuln(REMOTE, m.LeakedNodesId(["AzureVM"]))}),
        arm_mon: m.NodeInfo(services=_svc(("HTTPS", "azuread_user_credentials")), value=50, properties=[]),
        "AzureVM": m.NodeInfo(
            services=_svc("PING", "SSH"), value=100, properties=["CTFFLAG:VMPRIVATEINFO"],
            firewall=m.FirewallConfiguration(incoming=_rules(("SSH", BLOCK)), outgoing=[])),
        "client": m.NodeInfo(
            services=[], value=0, agent_installed=True, reimagable=False,
            vulnerabilities={"SearchEdgeHistory": _vuln(LOCAL, m.LeakedNodesId(["Website"]))}),
    }


def toyctf_environment() -> m.Environment:
    nodes = toyctf_nodes()
    return m.Environment(network=m.create_network(nodes), vulnerability_library={},
                         identifiers=m.infer_constants_from_nodes(nodes.items(), {}))


# ------------------------------------------------------------------------------------------------
# Chain
# ------------------------------------------------------------------------------------------------
CHAIN_IDENTIFIERS = m.Identifiers(
    properties=["Windows", "Linux", "ApacheWebSite", "IIS_2019", "IIS_2020_patched", "MySql", "Ubuntu",
                "nginx/1.10.3", "SMB_vuln", "SMB_vuln_patched", "SQLServer", "Win10", "Win10Patched", "FLAG:Linux"],
    ports=["HTTPS", "GIT", "SSH", "RDP", "PING", "MySQL", "SSH-key", "su"],
    local_vulnerabilities=["ScanBashHistory", "ScanExplorerRecentFiles", "SudoAttempt", "CrackKeepPassX", "CrackKeepPass"],
    remote_vulnerabilities=["ProbeLinux", "ProbeWindows"],
)


def _linux(i: int) -> str:
    return f"{i}_LinuxNode"


def _windows(i: int) -> str:
    return f"{i}_WindowsNode"


def _ssh_pw(i: int) -> str:
    return f"LinuxPassword!{i}"


def _rdp_pw(i: int) -> str:
    return f"WindowsPassword!{i}"


def chain_nodes(size: int) -> Dict[str, m.NodeInfo]:
    """start -> (Linux_i -> Windows_{i+1})* -> Linux_{size+1}[flag]; node order: start, flag node, then links."""
    if size % 2 == 1:
        raise ValueError(f"Chain size must be even: {size}")
    linux_rules = _std_rules()  # every Linux link shares this one list for incoming AND outgoing
    trap = m.ExploitFailed
    nodes: Dict[str, m.NodeInfo] = {
        "start": m.NodeInfo(
            services=[], value=0, agent_installed=True, reimagable=False,
            vulnerabilities={"ScanExplorerRecentFiles": _vuln(LOCAL, _leak_creds((_linux(1), "SSH", _ssh_pw(1))))}),
        _linux(size + 1): m.NodeInfo(
            services=_svc("HTTPS", ("SSH", _ssh_pw(size + 1))), value=1000,
            properties=["MySql", "Ubuntu", "nginx/1.10.3", "FLAG:Linux"], vulnerabilities={}),
    }
    for i in range(1, size, 2):
        nodes[_linux(i)] = m.NodeInfo(
            services=_svc("HTTPS", ("SSH", _ssh_pw(i))),
            firewall=m.FirewallConfiguration(incoming=linux_rules, outgoing=linux_rules),
            value=100, properties=["MySql", "Ubuntu", "nginx/1.10.3"],
            vulnerabilities={
                "ProbeLinux": _vuln(REMOTE, m.ProbeSucceeded(["Ubuntu"]), cost=5.0),
                "ProbeWindows": _vuln(REMOTE, m.ProbeFailed(), cost=5.0),
                "ScanBashHistory": _vuln(LOCAL, m.LeakedNodesId([_windows(i + 1)])),
                "ScanExplorerRecentFiles": _vuln(LOCAL, trap(), cost=10.0),
                "SudoAttempt": _vuln(LOCAL, trap(), cost=100.0),
                "CrackKeepPassX": _vuln(LOCAL, _leak_creds((_windows(i + 1), "RDP", _rdp_pw(i + 1)))),
            })
        nodes[_windows(i + 1)] = m.NodeInfo(
            services=_svc("HTTPS", ("RDP", _rdp_pw(i + 1))), value=100,
            properties=["Windows", "Win10", "Win10Patched"],
            vulnerabilities={
                "ProbeLinux": _vuln(REMOTE, m.ProbeFailed()),
                "ProbeWindows": _vuln(REMOTE, m.ProbeSucceeded(["Windows"])),
                "ScanBashHistory": _vuln(LOCAL, trap(), cost=100.0),
                "ScanExplorerRecentFiles": _vuln(LOCAL, m.LeakedNodesId([_linux(i + 2)])),
                "SudoAttempt": _vuln(LOCAL, trap(), cost=100.0),
                "CrackKeepPassX": _vuln(LOCAL, trap(), cost=100.0),
                "CrackKeepPass": _vuln(LOCAL, _leak_creds((_linux(i + 2), "SSH", _ssh_pw(i + 2)))),
            })
    return nodes


def chain_environment(size: int) -> m.Environment:
    return m.Environment(network=m.create_network(chain_nodes(size)), vulnerability_library={}, identifiers=CHAIN_IDENTIFIERS)


# ------------------------------------------------------------------------------------------------
# registry (gym ids of cyberbattle/__init__.py:31-71)
# ------------------------------------------------------------------------------------------------
def make_environment(env_id: str, **kwargs) -> m.Environment:
    if env_id == "CyberBattleToyCtf-v0":
        return toyctf_environment()
    if env_id == "CyberBattleChain-v0":
        return chain_environment(int(kwargs.get("size", 4)))
    if env_id == "CyberBattleRandom-v0":
        from .random_network import random_environment

        return random_environment(seed=kwargs.get("seed", 0))
    raise KeyError(f"unknown environment id {env_id!r}")
