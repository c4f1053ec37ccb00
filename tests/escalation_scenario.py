"""A small network whose vulnerabilities ESCALATE privileges -- none of the reference's registered scenarios has such an
outcome, so the `privilege_N` tag path (actions.py:369-378: append the tag to the node's properties; preconditions then see
it, actions.py:158-171) had no reference-recorded tape.  TEST INFRASTRUCTURE.

The scenario is written once, against whatever ``model`` module it is given: ``oracle/gen_golden.py --round2`` builds it from
the REFERENCE's ``cyberbattle.simulation.model`` and records tapes on the unmodified reference env; the tests build it from
``marlon_b200.model`` and replay those tapes (``helpers.config_from_meta`` checks that both compile to the same table bytes).
Shapes follow the reference's own escalation fixtures (actions_test.py:24-75: UAC bypass -> AdminEscalation guarded by
``~(privilege_2|privilege_3)``, a credential dump that needs one of the two tags) and add what they do not reach through the
gym action space: an escalation with precondition ``true`` (repeat -> -1), a REMOTE escalation that takes an unowned node
straight to System, library precedence over a node's own entry of the same id, and re-imagable nodes for the
ScanAndReimage tape (tags survive a re-image, the privilege level does not).
"""
ENV_ID = "test:escalation"


def build(model):
    m = model
    admin, system = m.AdminEscalation().tag, m.SystemEscalation().tag  # "privilege_2", "privilege_3"
    LOCAL, REMOTE = m.VulnerabilityType.LOCAL, m.VulnerabilityType.REMOTE
    library = {
        "UACBypass": m.VulnerabilityInfo(
            description="UAC bypass", type=LOCAL, precondition=m.Precondition(f"Windows&Win10&(~({admin}|{system}))"),
            outcome=m.AdminEscalation(), cost=1.0),
        "GetSystem": m.VulnerabilityInfo(
            description="token theft, works anywhere, any number of times", type=LOCAL, outcome=m.SystemEscalation(), cost=2.0),
    }
    nodes = {
        "start": m.NodeInfo(
            services=[], value=0, properties=["Windows", "Win10"], agent_installed=True, reimagable=False,
            vulnerabilities=dict(
                ListNeighbors=m.VulnerabilityInfo(description="reveal other nodes", type=LOCAL,
                                                  outcome=m.LeakedNodesId(nodes=["ws", "srv", "dc"])),
                DumpCreds=m.VulnerabilityInfo(
                    description="needs an elevated token", type=LOCAL, precondition=m.Precondition(f"Windows&({admin}|{system})"),
                    outcome=m.LeakedCredentials([m.CachedCredential("srv", "SSH", "root_pw"), m.CachedCredential("dc", "RDP", "da_pw")]),
                    cost=1.0),
                # same id as a library entry: the library's wins (actions.py:339-345)
                GetSystem=m.VulnerabilityInfo(description="shadowed by the library", type=LOCAL, outcome=m.CustomerData(), cost=9.0),
            )),
        "ws": m.NodeInfo(
            services=[m.ListeningService("RDP", allowedCredentials=["da_pw"])], value=40,
            properties=["Windows", "Win10", "PortRDPOpen"], reimagable=True,
            vulnerabilities=dict(
                RDPBF=m.VulnerabilityInfo(description="RDP brute force", type=REMOTE, precondition=m.Precondition("Windows&PortRDPOpen"),
                                          outcome=m.LateralMove(), cost=1.0),
            )),
        "srv": m.NodeInfo(
            services=[m.ListeningService("SSH", allowedCredentials=["root_pw"])], value=80, properties=["Linux", "PortSSHOpen", "PortSQLOpen"],
            reimagable=True,
            vulnerabilities=dict(
                KernelExploit=m.VulnerabilityInfo(description="remote root", type=REMOTE, precondition=m.Precondition("Linux"),
                                                  outcome=m.SystemEscalation(), cost=5.0),
                SudoCheck=m.VulnerabilityInfo(description="only as root", type=LOCAL, precondition=m.Precondition(f"Linux&{system}"),
                                              outcome=m.ProbeSucceeded(["PortSQLOpen"]), cost=1.0),
            )),
        "dc": m.NodeInfo(
            services=[m.ListeningService("RDP", allowedCredentials=["da_pw"])], value=1000, properties=["Windows", "PortRDPOpen"],
            reimagable=False,
            vulnerabilities=dict(
                RDPBF=m.VulnerabilityInfo(description="RDP brute force (patched here)", type=REMOTE,
                                          precondition=m.Precondition("Windows&Win10&PortRDPOpen"), outcome=m.LateralMove(), cost=1.0),
                Mimikatz=m.VulnerabilityInfo(description="needs SYSTEM", type=LOCAL, precondition=m.Precondition(f"Windows&{system}"),
                                             outcome=m.CustomerData(), cost=1.0),
            )),
    }
    identifiers = m.Identifiers(
        properties=["Windows", "Win10", "Linux", "PortRDPOpen", "PortSSHOpen", "PortSQLOpen"],
        ports=["RDP", "SSH"],
        local_vulnerabilities=["ListNeighbors", "UACBypass", "GetSystem", "DumpCreds", "SudoCheck", "Mimikatz"],
        remote_vulnerabilities=["RDPBF", "KernelExploit"])
    return m.Environment(network=m.create_network(nodes), vulnerability_library=library, identifiers=identifiers)
