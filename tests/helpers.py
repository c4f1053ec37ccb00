"""Shared test machinery: load golden tapes, build configs from their headers, replay them through any
batch object (the C oracle's ``OracleBatch`` or the CUDA ``Batch``) and compare everything bit for bit."""
import glob
import json
import os
import zlib

import numpy as np

from marlon_b200 import _abi, config, registry, scenario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# rewards are sums of small integers and x.0 costs: exact in fp32; BASELINE.json allows 1e-6 relative
REWARD_RTOL = 1e-6


def golden_tapes(kind=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        name = os.path.basename(p)[:-4]
        if name.startswith("kat_"):
            continue
        if kind and not name.startswith(kind):
            continue
        out.append(name)
    return out


class _Tape(dict):
    """The arrays of one .npz, decompressed once (NpzFile inflates a member again at every ``z[key]``)."""

    @property
    def files(self):
        return list(self)


def load_tape(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as npz:
        z = _Tape((k, npz[k]) for k in npz.files)
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, z


def _decode_kwargs(kw):
    out = {}
    for k, v in kw.items():
        if isinstance(v, dict) and "__nt__" in v:
            cls = getattr(config, v["__nt__"])
            out[k] = cls(**{a: b for a, b in v.items() if a != "__nt__"})
        elif isinstance(v, dict) and v.get("__scan__"):
            out[k] = config.ScanAndReimageCompromisedMachines(v["probability"], v["scan_capacity"], v["scan_frequency"])
        else:
            out[k] = v
    return out


def config_from_meta(meta, **overrides):
    """-> (CompiledScenario, cbx_config) for a tape header."""
    if meta["env_id"] == "test:escalation":  # not a registered id: the test-only escalation network (tests/escalation_scenario.py)
        import escalation_scenario
        from marlon_b200 import model

        env, env_kw = escalation_scenario.build(model), _decode_kwargs(meta["env_kwargs"])
    else:
        env, env_kw = registry.resolve(meta["env_id"], **_decode_kwargs(meta["env_kwargs"]))
    comp = scenario.compile_scenario(env)
    assert comp.fingerprint() == meta["fingerprint"], "scenario tables differ from the ones the tape was recorded on"
    env_kw.pop("observation_padding", None)
    if meta["kind"] == "raw":
        cfg = config.make_config(_abi.MODE_CYBERBATTLE, auto_reset=meta["auto_reset"], emit_terminal_obs=True, **env_kw, **overrides)
    else:
        a, d = meta["att_kwargs"], meta["def_kwargs"]
        cfg = config.make_config(
            _abi.MODE_MARLON, auto_reset=True, emit_terminal_obs=True,
            attacker_max_timesteps=a.get("max_timesteps", 2000),
            attacker_invalid_action_reward_modifier=a.get("invalid_action_reward_modifier", -1),
            action_kind_order=tuple(meta["kind_of_index"]),
            defender_enabled=meta["with_defender"],
            defender_max_timesteps=d.get("max_timesteps", 100),
            defender_invalid_action_reward=d.get("invalid_action_reward", 0),
            defender_reset_on_constraint_broken=d.get("reset_on_constraint_broken", True),
            defender_loss_reward=d.get("loss_reward", -5000.0),
            defender_sla_worsening_penalty_scale=d.get("sla_worsening_penalty_scale", 200.0),
            defender_binding=meta.get("defender_binding", "reference_stale"),
            **env_kw, **overrides)
        if meta.get("fw_fingerprint"):
            assert comp.fw_fingerprint() == meta["fw_fingerprint"], "firewall rule-list groups differ from the ones the tape was recorded on"
    return comp, cfg


def crc_rows(a):
    """crc32 of each env's slice of a [n, ...] int8 array."""
    a = np.ascontiguousarray(a)
    return np.array([zlib.crc32(a[i].tobytes()) & 0xFFFFFFFF for i in range(a.shape[0])], dtype=np.uint32)


def _eq(name, step, got, want):
    got, want = np.asarray(got), np.asarray(want)
    if got.shape != want.shape or not np.array_equal(got, want):
        bad = np.argwhere(got != want)[:5] if got.shape == want.shape else "shape"
        raise AssertionError(f"step {step}: {name} differs (first mismatches at {bad if isinstance(bad, str) else bad.tolist()})\n"
                             f" got  {got.reshape(got.shape[0], -1)[:2] if got.ndim > 1 else got[:8]}\n"
                             f" want {want.reshape(want.shape[0], -1)[:2] if want.ndim > 1 else want[:8]}")


def _close(name, step, got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    tol = REWARD_RTOL * np.maximum(1.0, np.abs(want))
    if not (np.abs(got - want) <= tol).all():
        i = int(np.argmax(np.abs(got - want) - tol))
        raise AssertionError(f"step {step}: {name}[{i}] = {got[i]!r}, reference {want[i]!r}")


def check_attacker_obs(step, get, z, prefix="", rows=None):
    """Compare the attacker observation arrays of a batch with a tape row. `rows`: envs to check (default all)."""
    sel = (lambda a: a) if rows is None else (lambda a: a[rows])
    t = "term_" if prefix else ""
    _eq(prefix + "scalars", step, sel(get(t + "scalars")), sel(z[prefix + "scalars"][step]))
    _eq(prefix + "leaked_credentials", step, sel(get(t + "leaked_credentials")), sel(z[prefix + "leaked"][step]))
    _eq(prefix + "credential_cache_matrix", step, sel(get(t + "credential_cache_matrix")), sel(z[prefix + "cachem"][step]))
    _eq(prefix + "discovered_nodes_properties", step, sel(get(t + "discovered_nodes_properties")),
        sel(z[prefix + "props"][step].astype(np.int32)))
    _eq(prefix + "nodes_privilegelevel", step, sel(get(t + "nodes_privilegelevel")), sel(z[prefix + "priv"][step].astype(np.int32)))
    _eq(prefix + "local_vulnerability", step, sel(get(t + "local_vulnerability")), sel(z[prefix + "local"][step]))
    _eq(prefix + "remote_vulnerability(crc)", step, sel(crc_rows(get(t + "remote_vulnerability"))), sel(z[prefix + "remote_crc"][step]))
    _eq(prefix + "connect(crc)", step, sel(crc_rows(get(t + "connect"))), sel(z[prefix + "connect_crc"][step]))
    if not prefix:
        _eq("owned_bits", step, sel(get("owned_bits")), sel(z["owned_bits"][step]))


def replay(meta, z, batch, get, export, max_steps=None, check_every=1):
    """Feed a tape's inputs to `batch`; compare outputs and state with what the reference produced.

    batch: object with reset() and step(att, def, scan_u, detect_u); get(name) -> numpy array of a view;
    export() -> canonical state [n, W]."""
    raw = meta["kind"] == "raw"
    steps = meta["steps"] if max_steps is None else min(max_steps, meta["steps"])
    batch.reset()
    has_scan = "scan_u" in z.files
    for s in range(steps):
        if raw:
            att = z["action"][s]
            batch.step(att, None, z["scan_u"][s] if has_scan else None, z["detect_u"][s] if has_scan else None)
        else:
            batch.step(z["att_action"][s], z["def_action"][s] if meta["with_defender"] else None, None, None)
        if s % check_every and s != steps - 1:
            continue
        info = get("att_info")
        if raw:
            _close("reward", s, get("att_reward"), z["reward"][s])
            _close("raw reward", s, info[:, 1].copy().view(np.float32), z["raw"][s])
            _eq("outcome", s, info[:, 2], z["outcome"][s])
            _eq("terminated", s, get("att_terminated"), z["terminated"][s])
            _eq("stepcount", s, info[:, 4], z["stepcount"][s])
            _eq("network_availability", s, get("network_availability"), z["availability"][s])
            done = z["terminated"][s].astype(bool)
        else:
            _close("att_reward", s, get("att_reward"), z["att_reward"][s])
            _close("cyber_reward", s, info[:, 0].copy().view(np.float32), z["cyber_reward"][s])
            _close("raw reward", s, info[:, 1].copy().view(np.float32), z["raw"][s])
            _eq("outcome", s, info[:, 2], z["outcome"][s])
            _eq("intercepted", s, info[:, 5], z["intercepted"][s])
            _eq("network_availability", s, get("network_availability"), z["availability"][s])
            _eq("att_terminated", s, get("att_terminated"), z["att_terminated"][s])
            _eq("att_truncated", s, get("att_truncated"), z["att_truncated"][s])
            done = (z["att_terminated"][s] | z["att_truncated"][s]).astype(bool)
            if meta["with_defender"]:
                _close("def_reward", s, get("def_reward"), z["def_reward"][s])
                _eq("def_terminated", s, get("def_terminated"), z["def_terminated"][s])
                _eq("def_truncated", s, get("def_truncated"), z["def_truncated"][s])
                _eq("infected_nodes", s, get("def_infected_nodes"), z["infected"][s])
                _eq("incoming_firewall_status", s, get("def_incoming_firewall"), z["fw_in"][s])
                _eq("outgoing_firewall_status", s, get("def_outgoing_firewall"), z["fw_out"][s])
                _eq("services_status", s, get("def_services_status"), z["services"][s])
                ddone = (z["def_terminated"][s] | z["def_truncated"][s]).astype(bool)
                if ddone.any():
                    _eq("terminal infected_nodes", s, get("term_def_infected_nodes")[ddone], z["term_infected"][s][ddone])
        check_attacker_obs(s, get, z)
        if done.any() and (not raw or meta["auto_reset"]):
            check_attacker_obs(s, get, z, prefix="term_", rows=np.nonzero(done)[0])
        _eq("state digest", s, export(), z["digest"][s])
    return steps
