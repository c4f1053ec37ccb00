"""Seeded generator of ``CyberBattleRandom`` networks (configs[4]).

The reference builds these networks in two stages
(``src/CyberBattleSim/cyberbattle/simulation/generate_network.py``):

1. ``generate_random_traffic_network`` (:22-77): for each protocol a two-block stochastic block model (clients, servers of
   that protocol) whose edge probabilities come from a Beta draw, scaled per protocol and clipped; the union of the three
   edge sets, labelled with the protocols seen on each edge, is the traffic graph.
2. ``cyberbattle_model_from_traffic_graph`` (:80-263): a random entry node; on every node vulnerabilities that leak the
   node's traffic neighbours (cached RDP / SMB credentials, recent network shares, traceroute) with per-target coin flips;
   passwords are created, rotated or shared by further coin flips.

``new_environment`` (:266-294), which ``CyberBattleRandom-v0`` calls, passes ``seed=None`` and so cannot be reproduced.
Here every draw comes from generators seeded by the caller: the Beta draw from ``numpy.random.RandomState(seed)``, the
block model from ``networkx.stochastic_block_model(..., seed=seed)`` (a library call, like the reference's) and every coin
flip from one ``random.Random(seed)`` consumed in the reference's order -- so ``random_environment(seed)`` equals what the
reference's two functions return for ``seed`` after ``random.seed(seed)`` (``tests/test_random_network.py`` pins the
compiled tables of a few seeds recorded from the reference; ``oracle/gen_golden.py --random`` wrote them).

Quirk kept on purpose (the compiled tables depend on it): a node only gets a listening service for the (node, port) pairs
that already had a password when the ENTRY node's vulnerabilities were generated (:222-228 runs before :237-241); the
password lists of those services keep growing afterwards because the service holds the same list object.
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import model as m

IDENTIFIERS = m.Identifiers(
    properties=["breach_node"],
    ports=["SMB", "HTTP", "RDP"],
    local_vulnerabilities=["ScanWindowsCredentialManagerForRDP", "ScanWindowsExplorerRecentFiles",
                           "ScanWindowsCredentialManagerForSMB"],
    remote_vulnerabilities=["Traceroute"],
)

PROTOCOLS = ("SMB", "HTTP", "RDP")  # iteration order of the reference's n_servers dict
PROTOCOL_SCALE = {"SMB": 3.0, "HTTP": 1.0, "RDP": 4.0}


def traffic_edges(seed: int, n_clients: int, n_servers: Dict[str, int], alpha: np.ndarray, beta: np.ndarray,
                  tolerance: float = 1e-3) -> Dict[Tuple[int, int], set]:
    """Stage 1: ordered map edge -> protocols seen on it (first-seen order, which later fixes the node order)."""
    import networkx as nx

    labels: Dict[Tuple[int, int], set] = {}
    tol = np.float32(tolerance)
    for proto in n_servers:
        probs = np.random.RandomState(seed).beta(a=alpha, b=beta, size=(2, 2)) * PROTOCOL_SCALE.get(proto, 1.0)
        probs = np.clip(probs, a_min=tol, a_max=np.float32(1.0 - tol))
        block = nx.stochastic_block_model(sizes=[n_clients, n_servers[proto]], p=probs, directed=True, seed=seed)
        for edge in block.edges:
            labels.setdefault(edge, set()).add(proto)
    return labels


class _Passwords:
    """Password bookkeeping of stage 2: which passwords open (node, port), in creation order."""

    def __init__(self, rng: random.Random, p_changed: float, p_shared: float):
        self.rng, self.p_changed, self.p_shared = rng, p_changed, p_shared
        self.count = 0
        self.valid: Dict[Tuple[str, str], List[str]] = {}

    def fresh(self) -> str:
        self.count += 1
        return f"unique_pwd{self.count}"

    def fresh_valid(self, node: str, port: str) -> str:
        pwd = self.fresh()
        self.valid.setdefault((node, port), []).append(pwd)
        return pwd

    def cached(self, node: str, port: str) -> str:
        """The password a client has cached for (node, port): rotated since (invalid), shared with others, or its own."""
        if self.rng.random() < self.p_changed:
            return self.fresh()
        if self.rng.random() < self.p_shared and (node, port) in self.valid:
            return self.rng.choice(self.valid[node, port])
        return self.fresh_valid(node, port)


def network_from_traffic(labels: Dict[Tuple[int, int], set], rng: random.Random, *, p_cached_smb=0.75, p_cached_rdp=0.8,
                         p_cached_shares=0.6, p_password_changed=0.1, p_traceroute=0.5, p_shared_password=0.8) -> Dict[str, m.NodeInfo]:
    """Stage 2: node id -> NodeInfo, in the reference's node order."""
    order: List[str] = []
    seen = set()
    out_edges: Dict[str, List[Tuple[str, set]]] = {}
    for (u, v), protos in labels.items():
        su, sv = str(u), str(v)
        for x in (su, sv):
            if x not in seen:
                seen.add(x)
                order.append(x)
        out_edges.setdefault(su, []).append((sv, protos))
    # the reference walks graph.edges(): grouped by source node in node order, targets in insertion order per source
    def targets(node: str, proto: str) -> List[str]:
        return [t for t, ps in out_edges.get(node, []) if proto in ps]

    pw = _Passwords(rng, p_password_changed, p_shared_password)

    def leak_vulnerabilities(node: str) -> m.VulnerabilityLibrary:
        lib: m.VulnerabilityLibrary = {}
        rdp, smb = targets(node, "RDP"), targets(node, "SMB")
        if rdp:
            creds = [m.CachedCredential(node=t, port="RDP", credential=pw.cached(t, "RDP")) for t in rdp if rng.random() < p_cached_rdp]
            lib["ScanWindowsCredentialManagerForRDP"] = m.VulnerabilityInfo(
                description="Look for RDP credentials in the Windows Credential Manager", type=m.VulnerabilityType.LOCAL,
                outcome=m.LeakedCredentials(credentials=creds), reward_string="Discovered creds in the Windows Credential Manager",
                cost=2.0)
        if smb:
            shares = [t for t in smb if rng.random() < p_cached_shares]
            lib["ScanWindowsExplorerRecentFiles"] = m.VulnerabilityInfo(
                description="Look for network shares in the Windows Explorer Recent files", type=m.VulnerabilityType.LOCAL,
                outcome=m.LeakedNodesId(shares), reward_string="Windows Explorer Recent Files revealed network shares", cost=1.0)
            creds = [m.CachedCredential(node=t, port="SMB", credential=pw.cached(t, "SMB")) for t in smb if rng.random() < p_cached_smb]
            lib["ScanWindowsCredentialManagerForSMB"] = m.VulnerabilityInfo(
                description="Look for network credentials in the Windows Credential Manager", type=m.VulnerabilityType.LOCAL,
                outcome=m.LeakedCredentials(credentials=creds), reward_string="Discovered SMB creds in the Windows Credential Manager",
                cost=2.0)
        if smb and rdp:  # `smb_neighbors or rdp_neighbors` is the SMB list whenever this branch is taken
            hops = [t for t in smb if rng.random() < p_traceroute]
            lib["Traceroute"] = m.VulnerabilityInfo(
                description="Attempt to discvover network nodes using Traceroute", type=m.VulnerabilityType.REMOTE,
                outcome=m.LeakedNodesId(hops), reward_string="Discovered new network nodes via traceroute", cost=5.0)
        return lib

    def firewall() -> m.FirewallConfiguration:
        rules = lambda: [m.FirewallRule("RDP", m.RulePermission.ALLOW), m.FirewallRule("SMB", m.RulePermission.ALLOW)]  # noqa: E731
        return m.FirewallConfiguration(rules(), rules())

    entry = order[rng.randrange(len(order))]
    nodes: Dict[str, m.NodeInfo] = {k: None for k in order}  # type: ignore[misc]
    nodes[entry] = m.NodeInfo(services=[], value=0, properties=["breach_node"], vulnerabilities=leak_vulnerabilities(entry),
                              agent_installed=True, firewall=firewall(), reimagable=False)
    for node in order:  # services first (they alias the password lists), then the node values
        if node == entry:
            continue
        services = [m.ListeningService(name=port, allowedCredentials=pw.valid[(t, port)]) for (t, port) in list(pw.valid) if t == node]
        nodes[node] = m.NodeInfo(services=services, value=rng.randint(0, 100), agent_installed=False, firewall=firewall())
    for node in order:
        if node != entry:
            nodes[node].vulnerabilities = leak_vulnerabilities(node)
    return nodes


def random_environment(seed: Optional[int] = 0, n_servers_per_protocol: int = 15, n_clients: int = 50) -> m.Environment:
    """``new_environment(n_servers_per_protocol)`` (generate_network.py:266-294) with every draw seeded by `seed`."""
    if seed is None:
        seed = random.SystemRandom().randrange(2 ** 31)
    labels = traffic_edges(int(seed), n_clients, {p: n_servers_per_protocol for p in PROTOCOLS},
                           alpha=np.array([(1, 1), (0.2, 0.5)], dtype=float), beta=np.array([(1000, 10), (10, 100)], dtype=float))
    nodes = network_from_traffic(labels, random.Random(int(seed)), p_cached_rdp=0.8, p_cached_smb=0.7, p_cached_shares=0.8,
                                 p_password_changed=0.01, p_shared_password=0.9)
    return m.Environment(network=m.create_network(nodes), vulnerability_library={}, identifiers=IDENTIFIERS)
