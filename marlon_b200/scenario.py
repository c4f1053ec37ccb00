"""Scenario compiler: ``model.Environment`` -> dense static tables (the wire format of include/cbx.h).

Run once on the host per scenario; the resulting ``uint32`` blob is what
``cbx_scenario_create`` uploads and what every CTA of the step kernel stages
into shared memory.  The compiler is duck-typed on the attribute names of the
reference data model (``cyberbattle/simulation/model.py:63-345``), so the same
code compiles this package's own scenario definitions and -- in the test
container only -- the reference's live ``model.Environment`` objects, which is
how the scenario definitions in ``marlon_b200/scenarios.py`` are pinned
(``tests/test_scenarios.py``).

Table semantics and the reference lines they fold in:
  * node order = ``network.nodes`` order (defend_wrapper.py:340, _env/defender.py:45);
  * ``FW_OUT``/``FW_IN``: first rule whose port matches decides, no rule = blocked
    (actions.py:504-515), evaluated per attacker port name;
  * ``LISTEN``: ``port in [s.name for s in services]`` (actions.py:573);
  * ``AUTH``: secrets accepted by a running service of that name (actions.py:608-621);
  * vulnerability lookup: library first, then the node's own dict (actions.py:339-353);
  * precondition: symbol true iff it names a node property (actions.py:158-171), tabulated
    over the 16 combinations of dynamic ``privilege_N`` tags (actions.py:378);
  * defender observation bits: any rule naming RDP/SSH/HTTPS/HTTP/su/sudo
    (defend_wrapper.py:31,506-517).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Optional, Any, Dict, List, Tuple

import numpy as np

SCN_MAGIC = 0x31584243
SCN_VERSION = 3
H_WORDS = 24
(H_MAGIC, H_VERSION, H_TOTAL_WORDS, H_N_NODES, H_N_PORTS, H_N_PROPS, H_N_LOCAL, H_N_REMOTE, H_N_SECRETS,
 H_N_TRIPLES, H_N_SERVICES, H_MAX_LEAK, H_FLAGS, H_OFF_NODE, H_OFF_AUTH, H_OFF_VULN, H_OFF_PAYLOAD,
 H_N_PAYLOAD, H_OFF_TRIPLE) = range(19)
NODE_WORDS = 8
VULN_WORDS = 4
(OUT_EXPLOIT_FAILED, OUT_LEAKED_CREDENTIALS, OUT_LEAKED_NODES, OUT_LATERAL_MOVE, OUT_CUSTOMER_DATA,
 OUT_PROBE_SUCCEEDED, OUT_PROBE_FAILED, OUT_ESCALATION) = range(8)

DEFENDER_FIREWALL_RULE_LIST = ["RDP", "SSH", "HTTPS", "HTTP", "su", "sudo"]  # defend_wrapper.py:31
PRIVILEGE_TAGS = [f"privilege_{i}" for i in range(4)]  # str(PrivilegeLevel(i)) of an IntEnum is its value


class UnsupportedScenario(ValueError):
    """The scenario is valid for the reference but not expressible in the static tables."""


def _outcome_kind(outcome) -> int:
    names = {c.__name__ for c in type(outcome).__mro__}
    if "LeakedCredentials" in names:
        return OUT_LEAKED_CREDENTIALS
    if "LeakedNodesId" in names:
        return OUT_LEAKED_NODES
    if "LateralMove" in names:
        return OUT_LATERAL_MOVE
    if "CustomerData" in names:
        return OUT_CUSTOMER_DATA
    if "ProbeSucceeded" in names:
        return OUT_PROBE_SUCCEEDED
    if "ProbeFailed" in names:
        return OUT_PROBE_FAILED
    if "PrivilegeEscalation" in names:
        return OUT_ESCALATION
    return OUT_EXPLOIT_FAILED  # ExploitFailed and any unknown outcome: no special handling anywhere


def _eval_precondition(precondition, properties) -> bool:
    if hasattr(precondition, "evaluate"):
        return bool(precondition.evaluate(properties))
    expr = precondition.expression
    if isinstance(expr, str):
        from .model import Precondition

        return Precondition(expr).evaluate(properties)
    # boolean.py-style expression object (reference objects in the test container)
    import importlib

    algebra = importlib.import_module(type(expr).__module__).BooleanAlgebra()
    t, f = algebra.parse("true"), algebra.parse("false")
    mapping = {s: (t if str(s) in properties else f) for s in expr.get_symbols()}
    return expr.subs(mapping).simplify() == t


def _is_type(v, name: str) -> bool:
    return getattr(v.type, "name", str(v.type)) == name


def _passes(rules, port: str) -> bool:
    for r in rules:
        if r.port == port:
            return getattr(r.permission, "name", str(r.permission)) == "ALLOW"
    return False


@dataclass
class CompiledScenario:
    """The blob plus the host-side name tables needed to talk about it."""

    blob: np.ndarray  # uint32
    node_ids: List[str]
    identifiers: Any
    secrets: List[str]
    triples: List[Tuple[str, str, str]]
    n_services: int
    services_per_node: List[int]
    max_leak: int
    alias_groups: Dict[str, List[Tuple[str, str]]]
    fw_ext: Optional[np.ndarray] = None  # firewall extension tables for the `live` defender binding (FWX_* layout below)

    @property
    def n_nodes(self) -> int:
        return len(self.node_ids)

    def header(self, idx: int) -> int:
        return int(self.blob[idx])

    def tobytes(self) -> bytes:
        return self.blob.astype("<u4").tobytes()

    def fingerprint(self) -> str:
        import hashlib

        return hashlib.sha256(self.tobytes()).hexdigest()

    def fw_fingerprint(self) -> str:
        """Of the firewall extension tables (alias groups, rule lists as the `live` binding sees them)."""
        import hashlib

        return hashlib.sha256(np.asarray(self.fw_ext, dtype="<u4").tobytes()).hexdigest()


def compile_scenario(env) -> CompiledScenario:
    ident = env.identifiers
    ports, props = list(ident.ports), list(ident.properties)
    local_ids, remote_ids = list(ident.local_vulnerabilities), list(ident.remote_vulnerabilities)
    node_ids = list(env.network.nodes)
    infos = [env.network.nodes[k]["data"] for k in node_ids]
    n, P, L, R = len(node_ids), len(ports), len(local_ids), len(remote_ids)
    if not (ports and props and local_ids and remote_ids):
        raise ValueError("identifiers must define ports, properties, local and remote vulnerabilities")  # cyberbattle_env.py:411-414
    if n > 255:
        raise UnsupportedScenario(f"{n} nodes > 255")
    if P > 32:
        raise UnsupportedScenario(f"{P} ports > 32")
    if len(props) > 64:
        raise UnsupportedScenario(f"{len(props)} properties > 64")
    node_index = {k: i for i, k in enumerate(node_ids)}
    library = dict(env.vulnerability_library)

    def lookup(info, vid):
        if vid in library:  # global library wins (actions.py:339-345)
            return library[vid]
        return info.vulnerabilities.get(vid)

    # ---- intern secrets and (node, port, credential) triples, in deterministic scenario order
    secrets: Dict[str, int] = {}
    triples: Dict[Tuple[str, str, str], int] = {}

    def intern_secret(s):
        if s not in secrets:
            secrets[s] = len(secrets)
        return secrets[s]

    all_vulns = []
    for info in infos:
        for vid in local_ids + remote_ids:
            v = lookup(info, vid)
            if v is not None:
                all_vulns.append(v)
    for v in all_vulns:
        if _outcome_kind(v.outcome) == OUT_LEAKED_CREDENTIALS:
            for c in v.outcome.credentials:
                if c.node not in node_index:
                    raise ValueError(f"leaked credential references unknown node {c.node!r}")
                if c.port not in ports:
                    raise ValueError(f"The network has references to undefined port names: {{{c.port!r}}}")  # cyberbattle_env.py:437-440
                intern_secret(c.credential)
                key = (c.node, c.port, c.credential)
                if key not in triples:
                    triples[key] = len(triples)
    for info in infos:
        for s in info.services:
            for cred in s.allowedCredentials:
                intern_secret(cred)
    n_secrets = max(1, len(secrets))
    Ws = (n_secrets + 31) // 32

    # ---- node records, auth table
    node_tab = np.zeros((n, NODE_WORDS), dtype=np.uint32)
    auth = np.zeros((n, P, Ws), dtype=np.uint32)
    services_per_node = []
    svc_off = 0
    for i, info in enumerate(infos):
        if len(info.services) > 255:
            raise UnsupportedScenario("more than 255 services on a node")
        node_tab[i, 0] = np.int32(int(info.value)).view(np.uint32)
        flags = (1 if info.reimagable else 0) | ((1 if info.agent_installed else 0) << 1) \
            | ((int(info.privilege_level) & 3) << 2) | (len(info.services) << 8)
        node_tab[i, 1] = flags
        pb = 0
        for p in info.properties:
            if p in PRIVILEGE_TAGS:
                raise UnsupportedScenario("static privilege_N property")
            if p not in props:
                raise ValueError(f"The network has references to undefined property names: {{{p!r}}}")  # cyberbattle_env.py:442-445
            pb |= 1 << props.index(p)
        node_tab[i, 2] = pb & 0xFFFFFFFF
        node_tab[i, 3] = pb >> 32
        fo = fi = li = 0
        for pi, pname in enumerate(ports):
            if _passes(info.firewall.outgoing, pname):
                fo |= 1 << pi
            if _passes(info.firewall.incoming, pname):
                fi |= 1 << pi
            if pname in [s.name for s in info.services]:
                li |= 1 << pi
        node_tab[i, 4], node_tab[i, 5], node_tab[i, 6] = fo, fi, li
        dob = 0
        for ri, rname in enumerate(DEFENDER_FIREWALL_RULE_LIST):
            if any(r.port == rname for r in info.firewall.incoming):
                dob |= 1 << ri
            if any(r.port == rname for r in info.firewall.outgoing):
                dob |= 1 << (8 + ri)
        node_tab[i, 7] = dob | (svc_off << 16)
        if svc_off > 0xFFFF:
            raise UnsupportedScenario("too many services")
        for s in info.services:
            if s.name not in ports:
                raise ValueError(f"The network has references to undefined port names: {{{s.name!r}}}")
            if s.running:
                for cred in s.allowedCredentials:
                    sid = secrets[cred]
                    auth[i, ports.index(s.name), sid // 32] |= np.uint32(1 << (sid % 32))
        services_per_node.append(len(info.services))
        svc_off += len(info.services)

    # ---- vulnerabilities
    vuln_tab = np.zeros((n, L + R, VULN_WORDS), dtype=np.uint32)
    payload: List[int] = []
    has_escalation = False
    max_leak = 0
    for i, info in enumerate(infos):
        for vi, vid in enumerate(local_ids + remote_ids):
            v = lookup(info, vid)
            if v is None:
                continue
            expected = "LOCAL" if vi < L else "REMOTE"
            if not _is_type(v, expected):
                # reference raises ValueError out of step() (actions.py:357-358); ids are per-type in every in-scope scenario
                raise UnsupportedScenario(f"vulnerability id {vid!r} is used with both LOCAL and REMOTE types")
            kind = _outcome_kind(v.outcome)
            truth = 0
            for tags in range(16):
                plist = list(info.properties) + [PRIVILEGE_TAGS[b] for b in range(4) if tags >> b & 1]
                if _eval_precondition(v.precondition, plist):
                    truth |= 1 << tags
            level = 0
            off, cnt = len(payload), 0
            if kind == OUT_LEAKED_CREDENTIALS:
                for c in v.outcome.credentials:
                    payload.append(triples[(c.node, c.port, c.credential)])
                cnt = len(v.outcome.credentials)
                max_leak = max(max_leak, cnt)
            elif kind == OUT_LEAKED_NODES:
                for nid in v.outcome.nodes:
                    if nid not in node_index:
                        raise ValueError(f"leaked node id {nid!r} is not in the network")
                    payload.append(node_index[nid])
                cnt = len(v.outcome.nodes)
            elif kind == OUT_PROBE_SUCCEEDED:
                pb = 0
                for p in v.outcome.discovered_properties:
                    if p not in info.properties:  # reference asserts (actions.py:387-388)
                        raise ValueError(f"Discovered property {p} must belong to the set of properties associated with the node.")
                    if p not in PRIVILEGE_TAGS:
                        pb |= 1 << props.index(p)
                payload += [pb & 0xFFFFFFFF, pb >> 32]
                cnt = 2
            elif kind == OUT_ESCALATION:
                level = int(v.outcome.level) & 3
                has_escalation = True
            vuln_tab[i, vi, 0] = 1 | (kind << 1) | (level << 4) | (truth << 8)
            vuln_tab[i, vi, 1] = np.float32(float(v.cost)).view(np.uint32)
            vuln_tab[i, vi, 2] = off
            vuln_tab[i, vi, 3] = cnt

    triple_tab = np.zeros((max(1, len(triples)), 3), dtype=np.uint32)
    for (nid, pname, cred), t in triples.items():
        triple_tab[t] = (node_index[nid], ports.index(pname), secrets[cred])

    # ---- alias groups of firewall rule lists: recorded for the `live` defender binding
    seen: Dict[int, str] = {}
    alias_groups: Dict[str, List[Tuple[str, str]]] = {}
    for k, info in zip(node_ids, infos):
        for direction in ("incoming", "outgoing"):
            lst = getattr(info.firewall, direction)
            g = seen.setdefault(id(lst), f"g{len(seen)}")
            alias_groups.setdefault(g, []).append((k, direction))

    fw_ext = _firewall_extension(node_ids, infos, ports, alias_groups)

    # ---- assemble
    sections = [node_tab.ravel(), auth.ravel(), vuln_tab.ravel(),
                np.asarray(payload if payload else [0], dtype=np.uint32), triple_tab.ravel()]
    header = np.zeros(H_WORDS, dtype=np.uint32)
    off = H_WORDS
    offs = []
    for s in sections:
        offs.append(off)
        off += int(s.size)
    header[H_MAGIC], header[H_VERSION], header[H_TOTAL_WORDS] = SCN_MAGIC, SCN_VERSION, off
    header[H_N_NODES], header[H_N_PORTS], header[H_N_PROPS] = n, P, len(props)
    header[H_N_LOCAL], header[H_N_REMOTE] = L, R
    header[H_N_SECRETS], header[H_N_TRIPLES] = n_secrets, len(triples)
    header[H_N_SERVICES], header[H_MAX_LEAK] = svc_off, max_leak
    header[H_FLAGS] = 1 if has_escalation else 0
    header[H_OFF_NODE], header[H_OFF_AUTH], header[H_OFF_VULN], header[H_OFF_PAYLOAD], header[H_OFF_TRIPLE] = offs
    header[H_N_PAYLOAD] = len(payload)
    blob = np.concatenate([header] + sections).astype(np.uint32)
    # pad to a multiple of 4 words so the blob can be bulk-copied (cp.async.bulk needs 16-byte granules)
    if blob.size % 4:
        blob = np.concatenate([blob, np.zeros(4 - blob.size % 4, dtype=np.uint32)])
        blob[H_TOTAL_WORDS] = blob.size
    inv_secrets = [None] * len(secrets)
    for s, i in secrets.items():
        inv_secrets[i] = s
    inv_triples = [None] * len(triples)
    for t, i in triples.items():
        inv_triples[i] = t
    return CompiledScenario(blob=blob, node_ids=node_ids, identifiers=ident, secrets=inv_secrets, triples=inv_triples,
                            n_services=svc_off, services_per_node=services_per_node, max_leak=max_leak,
                            alias_groups=alias_groups, fw_ext=fw_ext)


# ---- firewall extension tables (`live` defender binding) -------------------------------------------------
# Under the live binding the LearningDefender edits the firewall rule LISTS of the environment the attacker plays in
# (marlon/defender_agents/defender.py:50-69), and several (node, direction) pairs share one list object (toy_ctf.py:14-19,
# chainpattern.py:48-53; deepcopy keeps the sharing), so the unit of state is the list = alias group.  What the path ever asks of
# a list is, per port NAME: is there a rule for it (defend_wrapper.py:360-372, :506-517) and does the first one ALLOW
# (actions.py:504-515); block removes every rule of a name, allow appends an ALLOW rule only where none exists -- so two bits
# per (group, name) carry the whole list.  Names: the defender's six first (its actions index them), then the attacker's ports.
# Layout (uint32 words): [0] magic "CBXF"  [1] names  [2] groups  [3] 0
#   [4, 4 + P)           name index of every attacker port
#   [.., + n)            per node: incoming group | outgoing group << 16
#   [.., + 2 * groups)   per group: names present, names whose first rule allows      (the initial per-env state)
FWX_MAGIC = 0x46584243


def _firewall_extension(node_ids, infos, ports, alias_groups) -> np.ndarray:
    names = list(DEFENDER_FIREWALL_RULE_LIST) + [p for p in ports if p not in DEFENDER_FIREWALL_RULE_LIST]
    if len(names) > 32:
        raise UnsupportedScenario("more than 32 firewall port names")
    group_ids = list(alias_groups)
    group_of = {}
    for g, members in alias_groups.items():
        for node, direction in members:
            group_of[(node, direction)] = group_ids.index(g)
    if len(group_ids) > 0xFFFF:
        raise UnsupportedScenario("too many firewall rule lists")
    node_words = [group_of[(k, "incoming")] | (group_of[(k, "outgoing")] << 16) for k in node_ids]
    init = []
    by_node = dict(zip(node_ids, infos))
    for g in group_ids:
        node, direction = alias_groups[g][0]
        rules = getattr(by_node[node].firewall, direction)
        present = allow = 0
        for i, nm in enumerate(names):
            if any(r.port == nm for r in rules):
                present |= 1 << i
                if _passes(rules, nm):
                    allow |= 1 << i
        init += [present, allow]
    words = [FWX_MAGIC, len(names), len(group_ids), 0] + [names.index(p) for p in ports] + node_words + init
    while len(words) % 4:
        words.append(0)
    return np.asarray(words, dtype=np.uint32)

