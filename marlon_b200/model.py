"""Scenario data model (host side).

Mirrors the names and field meanings of the reference data model
(``src/CyberBattleSim/cyberbattle/simulation/model.py:63-345`` in the
reference) so scenario definitions written for CyberBattleSim read the same
here.  These objects are only ever *inputs to the scenario compiler*
(``marlon_b200/scenario.py``): the simulation itself runs on the compiled
tables in HBM, never on these Python objects.

Differences from the reference, on purpose:
  * no networkx: ``Environment.network`` is an insertion-ordered node mapping
    exposing just ``.nodes`` (order = node index, reference: dict insertion
    order of the scenario);
  * preconditions are parsed by a tiny recursive-descent parser
    (``Precondition.evaluate``) instead of boolean.py;
  * no YAML / plotting / random labelling (out of scope).
"""
from __future__ import annotations

import enum
import re
from dataclasses import dataclass, field
from typing import Dict, Iterable, Iterator, List, NamedTuple, Optional, Sequence, Tuple

NodeID = str
PortName = str
CredentialID = str
VulnerabilityID = str
PropertyName = str


class VulnerabilityType(enum.Enum):
    LOCAL = 1
    REMOTE = 2


class PrivilegeLevel(enum.IntEnum):
    NoAccess = 0
    LocalUser = 1
    Admin = 2
    System = 3
    MAXIMUM = 3


class RulePermission(enum.Enum):
    ALLOW = 0
    BLOCK = 1


class MachineStatus(enum.Enum):
    Stopped = 0
    Running = 1
    Imaging = 2


# ---- vulnerability outcomes (reference model.py:118-195) ---------------------------------
class VulnerabilityOutcome:
    pass


class LateralMove(VulnerabilityOutcome):
    pass


class CustomerData(VulnerabilityOutcome):
    pass


class ProbeFailed(VulnerabilityOutcome):
    pass


class ExploitFailed(VulnerabilityOutcome):
    pass


class PrivilegeEscalation(VulnerabilityOutcome):
    def __init__(self, level: PrivilegeLevel):
        self.level = PrivilegeLevel(level)

    @property
    def tag(self) -> str:
        return f"privilege_{self.level}"


class AdminEscalation(PrivilegeEscalation):
    def __init__(self):
        super().__init__(PrivilegeLevel.Admin)


class SystemEscalation(PrivilegeEscalation):
    def __init__(self):
        super().__init__(PrivilegeLevel.System)


class ProbeSucceeded(VulnerabilityOutcome):
    def __init__(self, discovered_properties: Sequence[PropertyName]):
        self.discovered_properties = list(discovered_properties)


class CachedCredential(NamedTuple):
    node: NodeID
    port: PortName
    credential: CredentialID


class LeakedCredentials(VulnerabilityOutcome):
    def __init__(self, credentials: Sequence[CachedCredential]):
        self.credentials = list(credentials)


class LeakedNodesId(VulnerabilityOutcome):
    def __init__(self, nodes: Sequence[NodeID]):
        self.nodes = list(nodes)


# ---- preconditions -------------------------------------------------------------------------
_TOK = re.compile(r"\s*(?:(?P<op>[&|~!()])|(?P<sym>[^\s&|~!()]+))")


class Precondition:
    """Boolean expression over node property names: ``&``, ``|``, ``~``/``!``,
    parentheses, ``true``/``false``.  Reference: model.py:208-223 (boolean.py
    syntax); evaluated the way ``AgentActions._check_prerequisites`` does
    (actions.py:158-171): a symbol is true iff it names a property of the node.
    """

    def __init__(self, expression: str = "true"):
        self.expression = str(expression)
        self._rpn = self._parse(self.expression)

    @staticmethod
    def _parse(text):
        toks = []
        pos = 0
        text = text.strip()
        while pos < len(text):
            m = _TOK.match(text, pos)
            if m is None:
                raise ValueError(f"bad precondition {text!r}")
            pos = m.end()
            toks.append(m.group("op") or ("$" + m.group("sym")))
        out, i = [], 0

        def p_or():
            nonlocal i
            p_and()
            while i < len(toks) and toks[i] == "|":
                i += 1
                p_and()
                out.append("|")

        def p_and():
            nonlocal i
            p_not()
            while i < len(toks) and toks[i] == "&":
                i += 1
                p_not()
                out.append("&")

        def p_not():
            nonlocal i
            if i < len(toks) and toks[i] in "~!":
                i += 1
                p_not()
                out.append("~")
                return
            if i >= len(toks):
                raise ValueError(f"bad precondition {text!r}")
            t = toks[i]
            i += 1
            if t == "(":
                p_or()
                if i >= len(toks) or toks[i] != ")":
                    raise ValueError(f"missing ) in {text!r}")
                i += 1
            elif t.startswith("$"):
                out.append(t)
            else:
                raise ValueError(f"bad precondition {text!r}")

        p_or()
        if i != len(toks):
            raise ValueError(f"trailing tokens in precondition {text!r}")
        return out

    def symbols(self) -> List[str]:
        return [t[1:] for t in self._rpn if t.startswith("$") and t[1:].lower() not in ("true", "false", "1", "0")]

    def evaluate(self, properties: Iterable[str]) -> bool:
        props = set(properties)
        st: List[bool] = []
        for t in self._rpn:
            if t == "&":
                b, a = st.pop(), st.pop()
                st.append(a and b)
            elif t == "|":
                b, a = st.pop(), st.pop()
                st.append(a or b)
            elif t == "~":
                st.append(not st.pop())
            else:
                name = t[1:]
                low = name.lower()
                if low in ("true", "1"):
                    st.append(True)
                elif low in ("false", "0"):
                    st.append(False)
                else:
                    st.append(name in props)
        assert len(st) == 1
        return st[0]


@dataclass
class VulnerabilityInfo:
    description: str = ""
    type: VulnerabilityType = VulnerabilityType.LOCAL
    outcome: VulnerabilityOutcome = field(default_factory=ExploitFailed)
    precondition: Precondition = field(default_factory=Precondition)
    cost: float = 1.0
    reward_string: str = ""
    URL: str = ""


VulnerabilityLibrary = Dict[VulnerabilityID, VulnerabilityInfo]


@dataclass
class ListeningService:
    name: PortName
    allowedCredentials: List[CredentialID] = field(default_factory=list)
    running: bool = True
    sla_weight: float = 1.0


@dataclass(frozen=True)
class FirewallRule:
    port: PortName
    permission: RulePermission
    reason: str = ""


def _default_rules() -> List[FirewallRule]:
    return [FirewallRule(p, RulePermission.ALLOW) for p in ("RDP", "SSH", "HTTPS", "HTTP")]


@dataclass
class FirewallConfiguration:
    outgoing: List[FirewallRule] = field(default_factory=_default_rules)
    incoming: List[FirewallRule] = field(default_factory=_default_rules)


@dataclass
class NodeInfo:
    services: List[ListeningService] = field(default_factory=list)
    vulnerabilities: VulnerabilityLibrary = field(default_factory=dict)
    value: int = 0
    properties: List[PropertyName] = field(default_factory=list)
    firewall: FirewallConfiguration = field(default_factory=FirewallConfiguration)
    agent_installed: bool = False
    privilege_level: PrivilegeLevel = PrivilegeLevel.NoAccess
    reimagable: bool = True
    owned_string: str = ""
    sla_weight: float = 1.0


class Identifiers(NamedTuple):
    properties: List[PropertyName] = []
    ports: List[PortName] = ["Null"]
    local_vulnerabilities: List[VulnerabilityID] = []
    remote_vulnerabilities: List[VulnerabilityID] = []


class _NodeView:
    """``network.nodes`` look-alike: ordered, ``nodes[id]["data"]`` -> NodeInfo."""

    def __init__(self, nodes: Dict[NodeID, NodeInfo]):
        self._n = nodes

    def __iter__(self):
        return iter(self._n)

    def __len__(self):
        return len(self._n)

    def __contains__(self, k):
        return k in self._n

    def __getitem__(self, k):
        return {"data": self._n[k]}

    def items(self):
        return ((k, {"data": v}) for k, v in self._n.items())


class Network:
    def __init__(self, nodes: Dict[NodeID, NodeInfo]):
        self.nodes = _NodeView(dict(nodes))

    def has_node(self, k) -> bool:
        return k in self.nodes


def create_network(nodes: Dict[NodeID, NodeInfo]) -> Network:
    return Network(nodes)


@dataclass
class Environment:
    network: Network
    vulnerability_library: VulnerabilityLibrary
    identifiers: Identifiers

    def nodes(self) -> Iterator[Tuple[NodeID, NodeInfo]]:
        for k, v in self.network.nodes.items():
            yield k, v["data"]

    def get_node(self, node_id: NodeID) -> NodeInfo:
        return self.network.nodes[node_id]["data"]


# ---- identifier inference (reference model.py:413-459) --------------------------------------
def infer_constants_from_nodes(nodes: Iterable[Tuple[NodeID, NodeInfo]], vulnerabilities: VulnerabilityLibrary) -> Identifiers:
    nodes = list(nodes)
    props, ports, local, remote = set(), set(), set(), set()

    def scan(vulns: VulnerabilityLibrary):
        for vid, v in vulns.items():
            (local if v.type == VulnerabilityType.LOCAL else remote).add(vid)
            if isinstance(v.outcome, LeakedCredentials):
                ports.update(c.port for c in v.outcome.credentials)

    scan(vulnerabilities)
    for _, info in nodes:
        props.update(info.properties)
        ports.update(s.name for s in info.services)
        scan(info.vulnerabilities)
    return Identifiers(sorted(props), sorted(ports), sorted(local), sorted(remote))
