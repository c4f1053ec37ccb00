"""Record golden tapes from the UNMODIFIED reference.  TEST INFRASTRUCTURE; runs only where /root/reference exists.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz and scenario_fingerprints.json

Each tape is one ``.npz``: a JSON header (``meta``), the inputs that were fed to the reference (actions, the
built-in defender's random draws) and everything the reference produced at every step (observations in the
AttackerEnvWrapper-normalised form of SURVEY.md section E, rewards, flags, and a canonical state digest in the
``CBX_X_*`` layout of include/cbx.h).  ``tests/test_oracle_golden.py`` replays the inputs through the C oracle
(and ``tests/test_gpu_parity.py`` through the CUDA library) and demands bit-equality.

Because stable-baselines3 is not installable here, the SB3 ``DummyVecEnv`` protocol is emulated exactly as
documented: ``done = terminated or truncated``; on done the observation is kept as the
terminal observation and replaced by ``env.reset()``'s.  The two-agent lock-step is MARLon's ``collect_rollouts``
(marl_algorithm.py:43-49): attacker ``perform_step`` then defender ``perform_step``.
"""
import json
import math
import os
import random
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_loader  # noqa: E402

ref_loader.load()

import numpy.random  # noqa: E402
from cyberbattle._env import cyberbattle_env as ref_env  # noqa: E402
from cyberbattle._env import defender as ref_defender  # noqa: E402
from cyberbattle.simulation import actions as ref_actions  # noqa: E402
from cyberbattle.simulation import model as ref_model  # noqa: E402
from marlon.baseline_models.env_wrappers.attack_wrapper import AttackerEnvWrapper  # noqa: E402
from marlon.baseline_models.env_wrappers.defend_wrapper import DefenderEnvWrapper  # noqa: E402
from marlon.baseline_models.env_wrappers.environment_event_source import EnvironmentEventSource  # noqa: E402

from marlon_b200 import _abi, scenario  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# ---- instrumentation (wrappers around reference callables; no reference file is modified) ---------------
LAST = {"raw": 0.0, "outcome": None}


def _wrap_actuator(name):
    orig = getattr(ref_actions.AgentActions, name)

    def wrapped(self, *a, **k):
        res = orig(self, *a, **k)
        LAST["raw"], LAST["outcome"] = float(res.reward), res.outcome
        return res

    setattr(ref_actions.AgentActions, name, wrapped)


for _n in ("exploit_local_vulnerability", "exploit_remote_vulnerability", "connect_to_remote_machine"):
    _wrap_actuator(_n)


class DrawTape:
    """Feeds / records the draws of ScanAndReimage: random.random() inside random.choices and numpy.random.random()."""

    def __init__(self, rng, cap):
        self.rng, self.cap = rng, max(cap, 1)
        self.begin()

    def begin(self):
        self.scan = [math.nan] * self.cap
        self.detect = [math.nan] * self.cap
        self.k_scan = 0
        self.slot = -1

    def py_random(self):
        u = float(self.rng.random())
        self.scan[self.k_scan] = u
        self.k_scan += 1
        return u

    def np_random(self):
        u = float(self.rng.random())
        self.detect[self.slot] = u
        return u


class TapedScanAndReimage(ref_defender.ScanAndReimageCompromisedMachines):
    """Same step() body as the reference class (inherited, unmodified); only the two RNG *sources* are redirected
    for the duration of the call so the consumed draws can be recorded per slot."""

    tape: DrawTape = None

    def step(self, environment, actions, t):
        tape = self.tape
        tape.begin()
        orig_choices, orig_np = ref_defender.random.choices, ref_defender.numpy.random.random

        class SlotList(list):
            # the reference draws numpy.random.random() only for Running+installed nodes while iterating the
            # scanned nodes in order; iterating this list tells the tape which slot a detection draw belongs to
            def __iter__(inner):
                for i, x in enumerate(list.__iter__(inner)):
                    tape.slot = i
                    yield x

        def choices(population, k=1):  # random.choices without weights: population[floor(random() * n)]
            n = len(population)
            return SlotList(population[math.floor(tape.py_random() * n)] for _ in range(k))

        ref_defender.random.choices = choices
        ref_defender.numpy.random.random = tape.np_random
        try:
            super().step(environment, actions, t)
        finally:
            ref_defender.random.choices = orig_choices
            ref_defender.numpy.random.random = orig_np


# ---- canonical digest of the reference's objects (CBX_X_* layout) ------------------------------------------
def digest(env, comp, C, att=None, dfn=None):
    n = comp.n_nodes
    node_ids = comp.node_ids
    idx = {k: i for i, k in enumerate(node_ids)}
    ident = env.identifiers
    vids = [(v, True) for v in ident.local_vulnerabilities] + [(v, False) for v in ident.remote_vulnerabilities]
    Ws = (max(1, len(comp.secrets)) + 31) // 32
    x = np.zeros(_abi.X_HEADER_WORDS + 10 * n + C + Ws, dtype=np.int32)
    disc = env._CyberBattleEnv__discovered_nodes
    cache = env._CyberBattleEnv__credential_cache
    act = env._actuator
    live_prog = env._defender_actuator.node_reimaging_progress
    x[0] = env._CyberBattleEnv__stepcount
    x[1] = int(env._CyberBattleEnv__done)
    x[2], x[3] = len(disc), len(cache)
    if att is not None:
        x[4] = att.timesteps or 0
        x[6] = int(att.reset_request)
        x[9], x[10] = att.valid_action_count, att.invalid_action_count
    sh_prog = {}
    if dfn is not None:
        x[5] = dfn.timesteps
        x[7] = int(dfn.reset_request)
        x[8] = int(dfn._has_breached_sla)
        x[11], x[12] = dfn.valid_action_count, dfn.invalid_action_count
        sh_prog = dfn.defender._actuator.node_reimaging_progress
        x[14] = int(round((1.0 - dfn._actuator.network_availability) * n))
        x[15] = int(round((1.0 - dfn._prev_network_availability) * n))
    infos = [env.environment.get_node(k) for k in node_ids]
    x[13] = sum(1 for i in infos if i.status != ref_model.MachineStatus.Running)
    p = _abi.X_HEADER_WORDS
    x[p:p + n] = -1
    for k, nid in enumerate(disc):
        x[p + k] = idx[nid]
    p += n
    x[p:p + n] = [int(i.agent_installed) for i in infos]
    p += n
    x[p:p + n] = [int(i.privilege_level) for i in infos]
    p += n
    x[p:p + n] = [live_prog[k] + 1 if k in live_prog else 0 for k in node_ids]
    p += n
    x[p:p + n] = [sh_prog[k] + 1 if k in sh_prog else 0 for k in node_ids]
    p += n
    tr = act._discovered_nodes
    x[p:p + n] = [int(k in tr and tr[k].last_owned_at is not None) for k in node_ids]
    p += n
    for j, k in enumerate(node_ids):
        bits = 0
        if k in tr:
            for q in tr[k].discovered_properties:
                bits |= 1 << q
        x[p + j] = np.uint32(bits & 0xFFFFFFFF).view(np.int32)
        x[p + n + j] = np.uint32(bits >> 32).view(np.int32)
    p += 2 * n
    for j, k in enumerate(node_ids):
        bits = 0
        if k in tr:
            for vi, key in enumerate(vids):
                if key in tr[k].last_attack:
                    bits |= 1 << (2 * vi)
                    lr = infos[j].last_reimaging
                    if lr is None or tr[k].last_attack[key] >= lr:
                        bits |= 2 << (2 * vi)
        x[p + j] = np.uint32(bits).view(np.int32)
    p += n
    for j, info in enumerate(infos):
        x[p + j] = sum(1 << b for b in range(4) if f"privilege_{b}" in info.properties)
    p += n
    x[p:p + C] = -1
    for k, c in enumerate(cache):
        x[p + k] = comp.triples.index((c.node, c.port, c.credential))
    p += C
    for s in act._gathered_credentials:
        sid = comp.secrets.index(s)
        x[p + sid // 32] |= np.int32(1 << (sid % 32))
    return x


RES_CODE = {"NoneType": 0, "ExploitFailed": 1, "LeakedCredentials": 2, "LeakedNodesId": 3, "LateralMove": 4,
            "CustomerData": 5, "ProbeSucceeded": 6, "ProbeFailed": 7, "PrivilegeEscalation": 8, "AdminEscalation": 8,
            "SystemEscalation": 8}

SCALAR_KEYS = ["newly_discovered_nodes_count", "lateral_move", "customer_data_found", "probe_result", "escalation",
               "credential_cache_length", "discovered_node_count"]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


class Recorder:
    def __init__(self):
        self.rows = {}

    def add(self, **kw):
        for k, v in kw.items():
            self.rows.setdefault(k, []).append(np.array(v))

    def arrays(self):
        return {k: np.stack(v) for k, v in self.rows.items()}


def normalise_obs(obs, N, nprops, C, LEAK):
    """AttackerEnvWrapper.transform_observation's form (attack_wrapper.py:474-522) from a raw CyberBattleEnv obs
    or an already transformed one."""
    if "action_mask" in obs:
        am = obs["action_mask"]
        local, remote, connect = am["local_vulnerability"], am["remote_vulnerability"], am["connect"]
    else:
        local, remote, connect = obs["local_vulnerability"], obs["remote_vulnerability"], obs["connect"]
    leaked = obs["leaked_credentials"]
    if isinstance(leaked, tuple):
        leaked = np.concatenate([np.asarray(x, dtype=np.int32) for x in leaked])
    cm = obs["credential_cache_matrix"]
    if isinstance(cm, tuple):
        cm = np.concatenate([np.asarray(x, dtype=np.int32) for x in cm])
    props = np.asarray(obs["discovered_nodes_properties"], dtype=np.int32).reshape(-1)
    scal = np.array([int(obs[k]) for k in SCALAR_KEYS] + [int(bool((props == 2).all()))], dtype=np.int32)
    owned = 0
    loc = np.asarray(local, dtype=np.int8)
    rem = np.asarray(remote, dtype=np.int8)
    con = np.asarray(connect, dtype=np.int8)
    for s in range(N):
        if rem[s].any():
            owned |= 1 << s
    ow = np.array([(owned >> (32 * w)) & 0xFFFFFFFF for w in range((N + 31) // 32)], dtype=np.uint32)
    return dict(scalars=scal, leaked=np.asarray(leaked, dtype=np.int32).reshape(-1),
                cachem=np.asarray(cm, dtype=np.int32).reshape(-1), props=props.astype(np.int8),
                priv=np.asarray(obs["nodes_privilegelevel"], dtype=np.int8).reshape(-1), local=loc, owned_bits=ow,
                local_crc=np.uint32(crc(loc)), remote_crc=np.uint32(crc(rem)), connect_crc=np.uint32(crc(con)))


def save(name, meta, rec: Recorder):
    arrs = rec.arrays()
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrs)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(next(iter(arrs.values())))} steps")


def env_kwargs_meta(kw):
    out = {}
    for k, v in kw.items():
        if hasattr(v, "_asdict"):
            out[k] = {"__nt__": type(v).__name__, **v._asdict()}
        elif isinstance(v, ref_defender.ScanAndReimageCompromisedMachines):
            out[k] = {"__scan__": True, "probability": v.probability, "scan_capacity": v.scan_capacity,
                      "scan_frequency": v.scan_frequency}
        else:
            out[k] = v
    return out


# ---- raw CyberBattleEnv tapes ---------------------------------------------------------------------------
def sample_raw_action(env, rng, p_valid):
    """50/50 mix of a valid action and a uniform one over the gym nvec."""
    b = env.bounds
    if rng.random() < p_valid:
        a = env.sample_valid_action(kinds=[0, 1, 2])
    else:
        kind = ["local_vulnerability", "remote_vulnerability", "connect"][int(rng.integers(3))]
        nvec = env.action_space.spaces[kind].nvec
        a = {kind: np.array([int(rng.integers(int(m))) for m in nvec], dtype=np.int32)}
    kind = next(iter(a))
    code = {"local_vulnerability": 0, "remote_vulnerability": 1, "connect": 2}[kind]
    out = np.zeros(5, dtype=np.int32)
    out[0] = code
    out[1:1 + len(a[kind])] = np.asarray(a[kind], dtype=np.int32)
    return a, out


def record_raw(name, env_id, env_kwargs, n_tapes, steps, seed, p_valid=0.5, scripted=None, auto_reset=True, factory=None,
               meta_kwargs=None):
    """One file, `n_tapes` independent envs x `steps` steps; arrays are [steps, n_tapes, ...]."""
    scan = env_kwargs.get("defender_agent")
    cap = scan.scan_capacity if scan else 0
    rec = Recorder()
    envs, tapes, rngs = [], [], []
    for t in range(n_tapes):
        kw = dict(env_kwargs)
        rng = np.random.default_rng(seed + t)
        if scan:
            d = TapedScanAndReimage(scan.probability, scan.scan_capacity, scan.scan_frequency)
            d.tape = DrawTape(rng, cap)
            kw["defender_agent"] = d
            tapes.append(d.tape)
        env = factory(**kw) if factory else ref_loader.make(env_id, **kw)
        env.reset(seed=seed + t)
        envs.append(env)
        rngs.append(rng)
    comp = scenario.compile_scenario(envs[0]._CyberBattleEnv__initial_environment)
    b = envs[0].bounds
    N, C, LEAK = int(b.maximum_node_count), int(b.maximum_total_credentials), int(b.maximum_discoverable_credentials_per_action)
    nprops = int(b.property_count)
    for step in range(steps):
        row = {}
        for t, env in enumerate(envs):
            if scripted is not None:
                a = scripted[step]
                kind = next(iter(a))
                enc = np.zeros(5, dtype=np.int32)
                enc[0] = {"local_vulnerability": 0, "remote_vulnerability": 1, "connect": 2}[kind]
                enc[1:1 + len(a[kind])] = a[kind]
            else:
                a, enc = sample_raw_action(env, rngs[t], p_valid)
            if tapes:
                tapes[t].begin()
            LAST["raw"], LAST["outcome"] = 0.0, "unset"
            obs, reward, done, _trunc, info = env.step(a)
            o = normalise_obs(obs, N, nprops, C, LEAK)
            # connect with a credential index outside the cache never reaches the actuator (cyberbattle_env.py:736-737)
            if LAST["outcome"] == "unset":
                raw, code = (0.0, 9) if o["scalars"][7] == 1 else (-1.0, 0)
            else:
                raw, code = LAST["raw"], RES_CODE[type(LAST["outcome"]).__name__]
            r = dict(action=enc, reward=np.float64(reward), raw=np.float64(raw), outcome=np.int32(code),
                     terminated=np.uint8(done), availability=np.float64(info["network_availability"]),
                     stepcount=np.int32(info["step_count"]))
            if tapes:
                r["scan_u"] = np.array(tapes[t].scan[:cap], dtype=np.float64)
                r["detect_u"] = np.array(tapes[t].detect[:cap], dtype=np.float64)
            r["pre_reset_digest"] = digest(env, comp, C)
            if done and auto_reset:
                for k, v in o.items():
                    r["term_" + k] = v
                obs, _ = env.reset()
                o = normalise_obs(obs, N, nprops, C, LEAK)
            else:
                for k, v in o.items():
                    r["term_" + k] = np.zeros_like(v)
            r.update(o)
            r["digest"] = digest(env, comp, C)
            for k, v in r.items():
                row.setdefault(k, []).append(v)
        rec.add(**{k: np.stack(v) for k, v in row.items()})
    meta = dict(kind="raw", env_id=env_id, env_kwargs=env_kwargs_meta(dict(env_kwargs, **(meta_kwargs or {}))), n_tapes=n_tapes, steps=steps,
                auto_reset=auto_reset, N=N, C=C, LEAK=LEAK, fingerprint=comp.fingerprint(), node_ids=comp.node_ids,
                scan_capacity=cap)
    save(name, meta, rec)


# ---- MARLon attacker+defender tapes ------------------------------------------------------------------------
def record_marlon(name, env_id, env_kwargs, att_kwargs, def_kwargs, n_tapes, steps, seed, with_defender=True,
                  p_att_valid=0.0, p_def_empty=0.0, factory=None, meta_kwargs=None, masked=False, live=False):
    """`masked`: the attacker is driven through the reference's MaskedDiscreteAttackerWrapper (action_masking.py:30-165): per step
    its action_masks() (CRC + count) and the Discrete action fed to its step() are recorded next to the MultiDiscrete encoding.
    `live`: the LIVE defender binding -- after every CyberBattleEnv.reset() the DefenderEnvWrapper's and the LearningDefender's
    cached actuator / environment (defend_wrapper.py:51, defender.py:29-30) are pointed at the objects that reset just created,
    so the defender acts on the environment the attacker plays in.  No reference file is modified: the instance's `reset` is
    wrapped."""
    rec = Recorder()
    units = []
    for t in range(n_tapes):
        env = factory(**env_kwargs) if factory else ref_loader.make(env_id, **env_kwargs)
        es = EnvironmentEventSource()
        att = AttackerEnvWrapper(env, es, **att_kwargs)
        dfn = DefenderEnvWrapper(env, att, es, defender=True, **def_kwargs) if with_defender else None
        if live:
            def _rebinding_reset(*a, _orig=env.reset, _env=env, _dfn=dfn, **k):
                out = _orig(*a, **k)
                base = _env.unwrapped
                _dfn._actuator = base._defender_actuator
                _dfn.defender._actuator = base._defender_actuator
                _dfn.defender._environment = base.environment
                return out

            env.reset = _rebinding_reset
        # SB3 learn() resets every env once before the first step: attacker first, then defender
        aobs, _ = att.reset()
        dobs = dfn.reset()[0] if dfn else None
        mw = None
        if masked:
            from marlon.baseline_models.env_wrappers.action_masking import MaskedDiscreteAttackerWrapper

            mw = MaskedDiscreteAttackerWrapper(att)
        units.append(dict(env=env, att=att, dfn=dfn, aobs=aobs, dobs=dobs, rng=np.random.default_rng(seed + t), mw=mw))
    env0 = units[0]["env"]
    comp = scenario.compile_scenario(env0._CyberBattleEnv__initial_environment)
    b = env0.bounds
    N, C, LEAK = int(b.maximum_node_count), int(b.maximum_total_credentials), int(b.maximum_discoverable_credentials_per_action)
    nprops = int(b.property_count)
    subspaces = units[0]["att"].action_subspaces
    kind_code = {"local_vulnerability": 0, "remote_vulnerability": 1, "connect": 2}
    kind_of_index = [kind_code[subspaces[i][0]] for i in range(3)]
    att_nvec = [int(x) for x in units[0]["att"].action_space.nvec]
    def_nvec = [int(x) for x in units[0]["dfn"].action_space.nvec] if with_defender else []
    for step in range(steps):
        row = {}
        for u in units:
            env, att, dfn, rng = u["env"], u["att"], u["dfn"], u["rng"]
            # ---- attacker.perform_step
            extra = {}
            if masked:
                mw = u["mw"]
                mask = mw.action_masks()
                valid = np.flatnonzero(mask)
                if len(valid) and rng.random() < p_att_valid:
                    a_disc = int(valid[int(rng.integers(len(valid)))])
                else:
                    a_disc = int(rng.integers(int(mw.action_space.n)))
                a_act = mw._encode_for_inner_env(*mw._decode(a_disc)).astype(np.int64)
                extra = dict(att_discrete=np.int64(a_disc), mask_crc=np.uint32(crc(mask.astype(np.int8))), mask_count=np.int32(mask.sum()))
            elif rng.random() < p_att_valid:
                va = env.sample_valid_action(kinds=[0, 1, 2])
                kind = next(iter(va))
                a_act = np.zeros(len(att_nvec), dtype=np.int64)
                ki = [i for i in range(3) if subspaces[i][0] == kind][0]
                a_act[0] = ki
                a_act[subspaces[ki][1]:subspaces[ki][2]] = va[kind]
            else:
                a_act = np.array([int(rng.integers(m)) for m in att_nvec], dtype=np.int64)
            LAST["raw"], LAST["outcome"] = 0.0, "unset"
            if masked:
                aobs, ar, aterm, atrunc, ainfo = u["mw"].step(np.array(extra["att_discrete"]))
            else:
                aobs, ar, aterm, atrunc, ainfo = att.step(a_act)
            intercepted = bool(ainfo.get("invalid_action", False))
            if LAST["outcome"] == "unset":
                raw, code = (0.0, 0) if intercepted else (-1.0, 0)
            else:
                raw, code = LAST["raw"], RES_CODE[type(LAST["outcome"]).__name__]
            o = normalise_obs(aobs, N, nprops, C, LEAK)
            if o["scalars"][7] == 1 and not intercepted:
                code = 9
            r = dict(att_action=a_act.astype(np.int32), att_reward=np.float64(ar), att_terminated=np.uint8(aterm),
                     att_truncated=np.uint8(atrunc), cyber_reward=np.float64(att.last_cyber_reward), raw=np.float64(raw),
                     outcome=np.int32(code), intercepted=np.uint8(intercepted),
                     availability=np.float64(env._defender_actuator.network_availability))
            r.update(extra)
            if aterm or atrunc:
                for k, v in o.items():
                    r["term_" + k] = v
                aobs, _ = att.reset()
                o = normalise_obs(aobs, N, nprops, C, LEAK)
            else:
                for k, v in o.items():
                    r["term_" + k] = np.zeros_like(v)
            r.update(o)
            r["mid_digest"] = digest(env, comp, C, att, dfn)
            # ---- defender.perform_step
            if dfn is not None:
                if rng.random() < p_def_empty:
                    d_act = np.full(len(def_nvec), -1, dtype=np.int64)
                    dobs, dr, dterm, dtrunc, _ = dfn.step([])
                else:
                    d_act = np.array([int(rng.integers(m)) for m in def_nvec], dtype=np.int64)
                    dobs, dr, dterm, dtrunc, _ = dfn.step(d_act)
                r.update(def_action=d_act.astype(np.int32), def_reward=np.float64(dr), def_terminated=np.uint8(dterm),
                         def_truncated=np.uint8(dtrunc))
                r["term_infected"] = np.asarray(dobs["infected_nodes"], dtype=np.int8)
                if dterm or dtrunc:
                    dobs, _ = dfn.reset()
                r["infected"] = np.asarray(dobs["infected_nodes"], dtype=np.int8)
                r["fw_in"] = np.asarray(dobs["incoming_firewall_status"], dtype=np.int8)
                r["fw_out"] = np.asarray(dobs["outgoing_firewall_status"], dtype=np.int8)
                r["services"] = np.asarray(dobs["services_status"], dtype=np.int8)
            r["digest"] = digest(env, comp, C, att, dfn)
            for k, v in r.items():
                row.setdefault(k, []).append(v)
        rec.add(**{k: np.stack(v) for k, v in row.items()})
    meta = dict(kind="marlon", env_id=env_id, env_kwargs=env_kwargs_meta(dict(env_kwargs, **(meta_kwargs or {}))), att_kwargs=att_kwargs,
                def_kwargs=def_kwargs, with_defender=with_defender, n_tapes=n_tapes, steps=steps, N=N, C=C, LEAK=LEAK,
                fingerprint=comp.fingerprint(), node_ids=comp.node_ids, kind_of_index=kind_of_index,
                att_nvec=att_nvec, def_nvec=def_nvec, defender_binding="live" if live else "reference_stale",
                fw_fingerprint=comp.fw_fingerprint())
    save(name, meta, rec)


# ---- the reference's own fixtures -------------------------------------------------------------------------
def chain10_fixture_actions():
    """The action list of cyberbattle_env_test.py:43-98, parsed from the reference file (not copied)."""
    import ast

    path = os.path.join(ref_loader.REFERENCE_ROOT, "src/CyberBattleSim/cyberbattle/_env/cyberbattle_env_test.py")
    tree = ast.parse(open(path).read())
    for fn in tree.body:
        if isinstance(fn, ast.FunctionDef) and fn.name == "test_step_after_done":
            assign = fn.body[0]
            out = []
            for d in assign.value.elts:
                key = d.keys[0].value
                vals = [c.value for c in d.values[0].args[0].elts]
                out.append({key: np.array(vals, dtype=np.int32)})
            return out
    raise RuntimeError("fixture not found")


def toyctf_kat():
    """commandcontrol_test.py:14-73 replayed through the reference's AgentActions on node/vuln NAMES; records the
    per-call (kind, indices, reward) so the oracle's L1 entry points can be pinned on 389.0."""
    from cyberbattle.samples.toyctf import toy_ctf as ctf

    env = ref_model.Environment(network=ref_model.create_network(ctf.nodes), vulnerability_library={}, identifiers=ctf.ENV_IDENTIFIERS)
    import copy

    env = copy.deepcopy(env)
    comp = scenario.compile_scenario(env)
    act = ref_actions.AgentActions(env, throws_on_invalid_actions=True)
    ident = env.identifiers
    ni = {k: i for i, k in enumerate(comp.node_ids)}
    calls = []
    total = 0.0

    def local(node, vuln):
        nonlocal total
        r = act.exploit_local_vulnerability(node, vuln)
        calls.append([0, ni[node], ident.local_vulnerabilities.index(vuln), 0, 0, r.reward])
        total += r.reward
        return r.outcome

    def remote(src, tgt, vuln):
        nonlocal total
        r = act.exploit_remote_vulnerability(src, tgt, vuln)
        calls.append([1, ni[src], ni[tgt], ident.remote_vulnerabilities.index(vuln), 0, r.reward])
        total += r.reward
        return r.outcome

    def connect(src, tgt, port, cred):
        nonlocal total
        r = act.connect_to_remote_machine(src, tgt, port, cred)
        pidx = ident.ports.index(port) if port in ident.ports else -1
        calls.append([2, ni[src], ni[tgt], pidx, comp.secrets.index(cred), r.reward])
        total += r.reward
        return r.outcome

    local("client", "SearchEdgeHistory")
    remote("client", "Website", "ScanPageContent")
    sas = remote("client", "GitHubProject", "CredScanGitHistory").credentials[0].credential
    connect("client", "AzureStorage", "HTTPS", sas)
    remote("client", "Website", "ScanPageSource")
    mysql = remote("client", "Website.Directory", "NavigateWebDirectoryFurther").credentials[0].credential
    remote("client", "Website.Directory", "NavigateWebDirectory")
    ad = remote("client", "Sharepoint", "ScanSharepointParentDirectory").credentials[0].credential
    connect("client", "AzureResourceManager", "HTTPS", ad)
    remote("client", "AzureResourceManager", "ListAzureResources")
    connect("client", "AzureVM", "SSH", mysql)
    connect("client", "Website", "SSH", mysql)
    mon = local("Website", "CredScanBashHistory").credentials[0].credential
    # ("sudo" is not an attacker port: commandcontrol_test.py:54 is blocked by the outgoing rule lookup by NAME;
    #  the tables only know identifiers.ports, so that call is recorded with port -1 and skipped by the replay)
    connect("Website", "Website[user=monitor]", "sudo", mon)
    connect("client", "Website[user=monitor]", "SSH", mon)
    connect("Website", "Website[user=monitor]", "su", mon)
    aad = local("Website[user=monitor]", "CredScan-HomeDirectory").credentials[0].credential
    connect("client", "AzureResourceManager", "HTTPS", aad)
    assert total == 389.0, total
    np.savez_compressed(os.path.join(GOLDEN, "kat_toyctf_commandcontrol.npz"), calls=np.array(calls, dtype=np.float64),
                        total=np.float64(total),
                        meta=np.frombuffer(json.dumps(dict(fingerprint=comp.fingerprint())).encode(), dtype=np.uint8))
    print("wrote kat_toyctf_commandcontrol.npz: total", total)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    random.seed(0)
    numpy.random.seed(0)
    AG, DG, DC = ref_env.AttackerGoal, ref_env.DefenderGoal, ref_env.DefenderConstraint
    Scan = ref_defender.ScanAndReimageCompromisedMachines

    # scenario fingerprints of the reference's own scenario objects
    from cyberbattle.samples.chainpattern import chainpattern
    from cyberbattle.samples.toyctf import toy_ctf

    fps = {"CyberBattleToyCtf-v0": scenario.compile_scenario(toy_ctf.new_environment()).fingerprint()}
    for size in (4, 10, 100):
        fps[f"CyberBattleChain-v0:size={size}"] = scenario.compile_scenario(chainpattern.new_environment(size)).fingerprint()
    json.dump(fps, open(os.path.join(GOLDEN, "scenario_fingerprints.json"), "w"), indent=1, sort_keys=True)

    toyctf_kat()

    # (1) the reference's Chain-10 fixture: 56 scripted actions, done=True r=5000 at the end (default bounds are
    #     100 x 1000; the tape uses N=12, C=12 so the dense masks stay small -- indices are unaffected)
    fixture = chain10_fixture_actions()
    c10 = dict(size=10, maximum_node_count=12, maximum_total_credentials=12, attacker_goal=AG(own_atleast_percent=1.0))
    record_raw("raw_chain10_fixture", "CyberBattleChain-v0", c10, 1, 56, seed=1, scripted=fixture[:56], auto_reset=False)

    # (2) config 2: Chain-10 attacker-only, mixed valid / uniform actions
    c10b = dict(size=10, maximum_node_count=12, maximum_total_credentials=12, throws_on_invalid_actions=False)
    record_raw("raw_chain10_mixed", "CyberBattleChain-v0", c10b, 24, 400, seed=1000, p_valid=0.5)
    record_raw("raw_chain10_valid", "CyberBattleChain-v0", c10b, 8, 1500, seed=2000, p_valid=1.0)

    # (3) config 3: ToyCtf + ScanAndReimage(0.6, 2, 5), SLA 0.80 (notebook_withdefender.py:57-63 parameters)
    t3 = dict(maximum_node_count=12, maximum_total_credentials=10, throws_on_invalid_actions=False,
              defender_agent=Scan(probability=0.6, scan_capacity=2, scan_frequency=5),
              defender_constraint=DC(maintain_sla=0.80))
    record_raw("raw_toyctf_scan", "CyberBattleToyCtf-v0", t3, 24, 400, seed=3000, p_valid=0.8)
    c3 = dict(size=10, maximum_node_count=12, maximum_total_credentials=12, throws_on_invalid_actions=False,
              defender_agent=Scan(probability=0.6, scan_capacity=2, scan_frequency=5),
              defender_constraint=DC(maintain_sla=0.80))
    record_raw("raw_chain10_scan", "CyberBattleChain-v0", c3, 16, 600, seed=4000, p_valid=0.9)
    # aggressive scanner so that reimaging / eviction / SLA endings all occur
    c3b = dict(size=4, maximum_node_count=6, maximum_total_credentials=6, throws_on_invalid_actions=False,
               defender_agent=Scan(probability=0.9, scan_capacity=3, scan_frequency=2),
               defender_constraint=DC(maintain_sla=0.50))
    record_raw("raw_chain4_scan_aggressive", "CyberBattleChain-v0", c3b, 16, 600, seed=5000, p_valid=0.9)

    # (4) config 1: ToyCtf MARLon attacker+defender, uniform random MultiDiscrete actions (MultiAgentUniverse.build
    #     defaults: multiagent_universe.py:78-95,159-198)
    t1 = dict(maximum_node_count=12, maximum_total_credentials=10, maximum_discoverable_credentials_per_action=5,
              throws_on_invalid_actions=False, defender_constraint=DC(maintain_sla=0.60), losing_reward=-5000.0)
    akw = dict(max_timesteps=2000, invalid_action_reward_modifier=-1.0, invalid_action_reward_multiplier=1.0, loss_reward=-5000.0)
    dkw = dict(max_timesteps=2000, invalid_action_reward=-1, reset_on_constraint_broken=True, loss_reward=-5000.0)
    record_marlon("marlon_toyctf_uniform", "CyberBattleToyCtf-v0", t1, akw, dkw, 2, 3000, seed=12345)
    record_marlon("marlon_toyctf_valid", "CyberBattleToyCtf-v0", t1, akw, dkw, 8, 600, seed=22345, p_att_valid=0.7, p_def_empty=0.5)
    # short episodes so time-limit truncations and the cross-agent reset protocol are exercised many times
    akw2 = dict(akw, max_timesteps=37, invalid_action_reward_modifier=0.0)
    dkw2 = dict(dkw, max_timesteps=23, reset_on_constraint_broken=False, invalid_action_reward=0)
    record_marlon("marlon_toyctf_short", "CyberBattleToyCtf-v0", t1, akw2, dkw2, 8, 500, seed=32345, p_att_valid=0.6, p_def_empty=0.3)
    c1 = dict(size=10, maximum_node_count=12, maximum_total_credentials=12, throws_on_invalid_actions=False,
              defender_constraint=DC(maintain_sla=0.60), losing_reward=-5000.0)
    record_marlon("marlon_chain10_valid", "CyberBattleChain-v0", c1, akw, dkw, 8, 600, seed=42345, p_att_valid=0.9, p_def_empty=0.8)
    record_marlon("marlon_chain10_attacker_only", "CyberBattleChain-v0", c1, dict(akw, max_timesteps=150), {}, 8, 600,
                  seed=52345, with_defender=False, p_att_valid=0.9)


# ---- CyberBattleRandom: the reference's generator functions, seeded (configs[4]) ---------------------------------
def ref_random_environment(seed):
    """What ``generate_network.new_environment(15)`` builds (generate_network.py:266-294), with the seed it leaves out:
    the reference's own two functions, its own parameter values, ``random.seed(seed)`` for the global stream they draw from."""
    from cyberbattle.simulation import generate_network as gen
    from cyberbattle.simulation import model as ref_model

    random.seed(seed)
    traffic = gen.generate_random_traffic_network(
        seed=seed, n_clients=50, n_servers={"SMB": 15, "HTTP": 15, "RDP": 15},
        alpha=np.array([(1, 1), (0.2, 0.5)], dtype=float), beta=np.array([(1000, 10), (10, 100)], dtype=float))
    network = gen.cyberbattle_model_from_traffic_graph(
        traffic, cached_rdp_password_probability=0.8, cached_smb_password_probability=0.7,
        cached_accessed_network_shares_probability=0.8, cached_password_has_changed_probability=0.01,
        probability_two_nodes_use_same_password_to_access_given_resource=0.9)
    return ref_model.Environment(network=network, vulnerability_library=dict([]), identifiers=gen.ENV_IDENTIFIERS)


def main_random():
    """Only the CyberBattleRandom fixtures (the other tapes are left alone): table fingerprints of seeds 0..11 and tapes on
    two of the networks."""
    fp_path = os.path.join(GOLDEN, "scenario_fingerprints.json")
    fps = json.load(open(fp_path))
    for seed in range(12):
        comp = scenario.compile_scenario(ref_random_environment(seed))
        fps[f"CyberBattleRandom-v0:seed={seed}"] = comp.fingerprint()
        print("seed", seed, "nodes", comp.n_nodes, "credentials", len(comp.triples), "services", comp.n_services)
    json.dump(fps, open(fp_path, "w"), indent=1, sort_keys=True)
    AG, DC = ref_env.AttackerGoal, ref_env.DefenderConstraint
    Scan = ref_defender.ScanAndReimageCompromisedMachines

    def factory(seed):
        return lambda **kw: ref_env.CyberBattleEnv(initial_environment=ref_random_environment(seed), **kw)

    # CyberBattleRandom fixes maximum_discoverable_credentials_per_action=32 (cyberbattle_random.py:14); 65 nodes,
    # <= 100 credentials: bounds (72, 104) keep the dense masks of a tape small
    r3 = dict(maximum_node_count=72, maximum_total_credentials=104, maximum_discoverable_credentials_per_action=32,
              throws_on_invalid_actions=False, attacker_goal=AG(own_atleast_percent=1.0))
    record_raw("raw_random3_valid", "CyberBattleRandom-v0", r3, 6, 400, seed=6000, p_valid=0.9, factory=factory(3), meta_kwargs=dict(seed=3))
    r0 = dict(r3, defender_agent=Scan(probability=0.6, scan_capacity=2, scan_frequency=5), defender_constraint=DC(maintain_sla=0.80))
    record_raw("raw_random0_scan", "CyberBattleRandom-v0", r0, 6, 400, seed=7000, p_valid=0.9, factory=factory(0), meta_kwargs=dict(seed=0))
    akw = dict(max_timesteps=300, invalid_action_reward_modifier=-1.0, invalid_action_reward_multiplier=1.0, loss_reward=-5000.0)
    dkw = dict(max_timesteps=300, invalid_action_reward=-1, reset_on_constraint_broken=True, loss_reward=-5000.0)
    m3 = dict(maximum_node_count=72, maximum_total_credentials=104, maximum_discoverable_credentials_per_action=32,
              throws_on_invalid_actions=False, defender_constraint=DC(maintain_sla=0.60), losing_reward=-5000.0)
    record_marlon("marlon_random3_valid", "CyberBattleRandom-v0", m3, akw, dkw, 4, 400, seed=62345, p_att_valid=0.9, p_def_empty=0.5,
                  factory=factory(3), meta_kwargs=dict(seed=3))


def main_round2(only_masked=False):
    """Round-2 tapes (the other tapes are left alone): a scenario with PRIVILEGE ESCALATION outcomes and dynamic privilege_N
    tags (tests/escalation_scenario.py, built here from the reference's model classes), and Chain-100 -- the config-4 scenario --
    with the MARLon pair and with the built-in ScanAndReimage defender."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import escalation_scenario

    AG, DC = ref_env.AttackerGoal, ref_env.DefenderConstraint
    Scan = ref_defender.ScanAndReimageCompromisedMachines

    def esc_factory(**kw):
        return ref_env.CyberBattleEnv(initial_environment=escalation_scenario.build(ref_model), **kw)

    e1 = dict(maximum_node_count=6, maximum_total_credentials=4, throws_on_invalid_actions=False, attacker_goal=AG(own_atleast_percent=1.0))
    if only_masked:
        t1 = dict(maximum_node_count=12, maximum_total_credentials=10, maximum_discoverable_credentials_per_action=5,
                  throws_on_invalid_actions=False, defender_constraint=DC(maintain_sla=0.60), losing_reward=-5000.0)
        akw_m = dict(max_timesteps=80, invalid_action_reward_modifier=-1.0, invalid_action_reward_multiplier=1.0, loss_reward=-5000.0)
        dkw_m = dict(max_timesteps=80, invalid_action_reward=-1, reset_on_constraint_broken=True, loss_reward=-5000.0)
        record_marlon("marlon_toyctf_masked", "CyberBattleToyCtf-v0", t1, akw_m, dkw_m, 4, 400, seed=8300, p_att_valid=0.9, p_def_empty=0.5,
                      masked=True)
        return
    record_raw("raw_escalation_valid", escalation_scenario.ENV_ID, e1, 16, 300, seed=8000, p_valid=0.85, factory=esc_factory)
    e2 = dict(e1, defender_agent=Scan(probability=0.7, scan_capacity=2, scan_frequency=3), defender_constraint=DC(maintain_sla=0.5))
    record_raw("raw_escalation_scan", escalation_scenario.ENV_ID, e2, 16, 400, seed=8100, p_valid=0.9, factory=esc_factory)
    akw = dict(max_timesteps=60, invalid_action_reward_modifier=-1.0, invalid_action_reward_multiplier=1.0, loss_reward=-5000.0)
    dkw = dict(max_timesteps=45, invalid_action_reward=-1, reset_on_constraint_broken=True, loss_reward=-5000.0)
    e3 = dict(maximum_node_count=6, maximum_total_credentials=4, throws_on_invalid_actions=False,
              defender_constraint=DC(maintain_sla=0.60), losing_reward=-5000.0)
    record_marlon("marlon_escalation_valid", escalation_scenario.ENV_ID, e3, akw, dkw, 8, 400, seed=8200, p_att_valid=0.8, p_def_empty=0.4,
                  factory=esc_factory)

    # the attacker behind the reference's MaskedDiscreteAttackerWrapper: mask content and Discrete -> MultiDiscrete decoding
    t1 = dict(maximum_node_count=12, maximum_total_credentials=10, maximum_discoverable_credentials_per_action=5,
              throws_on_invalid_actions=False, defender_constraint=DC(maintain_sla=0.60), losing_reward=-5000.0)
    akw_m = dict(max_timesteps=80, invalid_action_reward_modifier=-1.0, invalid_action_reward_multiplier=1.0, loss_reward=-5000.0)
    dkw_m = dict(max_timesteps=80, invalid_action_reward=-1, reset_on_constraint_broken=True, loss_reward=-5000.0)
    record_marlon("marlon_toyctf_masked", "CyberBattleToyCtf-v0", t1, akw_m, dkw_m, 4, 400, seed=8300, p_att_valid=0.9, p_def_empty=0.5,
                  masked=True)

    # Chain-100 (chainpattern.py:198-243): bounds must be raised to 102 nodes / 102 credentials
    c100 = dict(size=100, maximum_node_count=102, maximum_total_credentials=102, throws_on_invalid_actions=False,
                defender_constraint=DC(maintain_sla=0.60), losing_reward=-5000.0)
    akw4 = dict(max_timesteps=2000, invalid_action_reward_modifier=-1.0, invalid_action_reward_multiplier=1.0, loss_reward=-5000.0)
    dkw4 = dict(max_timesteps=2000, invalid_action_reward=-1, reset_on_constraint_broken=True, loss_reward=-5000.0)
    record_marlon("marlon_chain100_valid", "CyberBattleChain-v0", c100, akw4, dkw4, 2, 300, seed=9000, p_att_valid=0.95, p_def_empty=0.5)
    c100s = dict(size=100, maximum_node_count=102, maximum_total_credentials=102, throws_on_invalid_actions=False,
                 defender_agent=Scan(probability=0.6, scan_capacity=2, scan_frequency=5), defender_constraint=DC(maintain_sla=0.80))
    record_raw("raw_chain100_scan", "CyberBattleChain-v0", c100s, 2, 300, seed=9100, p_valid=0.95)


def main_live():
    """Tapes of the LIVE defender binding: the reference's own DefenderEnvWrapper / LearningDefender
    classes with their binding refreshed at every reset, so re-imaging, block_traffic and allow_traffic act on the environment
    the attacker plays in -- rule lists shared between nodes included (toy_ctf.py:14-19)."""
    DC = ref_env.DefenderConstraint
    t1 = dict(maximum_node_count=12, maximum_total_credentials=10, maximum_discoverable_credentials_per_action=5,
              throws_on_invalid_actions=False, defender_constraint=DC(maintain_sla=0.60), losing_reward=-5000.0)
    akw = dict(max_timesteps=90, invalid_action_reward_modifier=-1.0, invalid_action_reward_multiplier=1.0, loss_reward=-5000.0)
    dkw = dict(max_timesteps=70, invalid_action_reward=-1, reset_on_constraint_broken=True, loss_reward=-5000.0)
    record_marlon("marlon_toyctf_live", "CyberBattleToyCtf-v0", t1, akw, dkw, 8, 700, seed=9500, p_att_valid=0.85, p_def_empty=0.35,
                  live=True)
    # SLA breaches do not end the episode here, so long stretches with several nodes re-imaging and many firewall edits occur
    dkw2 = dict(dkw, max_timesteps=200, reset_on_constraint_broken=False)
    record_marlon("marlon_toyctf_live_long", "CyberBattleToyCtf-v0", t1, dict(akw, max_timesteps=200), dkw2, 4, 700, seed=9600,
                  p_att_valid=0.9, p_def_empty=0.2, live=True)
    c1 = dict(size=10, maximum_node_count=12, maximum_total_credentials=12, throws_on_invalid_actions=False,
              defender_constraint=DC(maintain_sla=0.60), losing_reward=-5000.0)
    record_marlon("marlon_chain10_live", "CyberBattleChain-v0", c1, akw, dkw, 6, 600, seed=9700, p_att_valid=0.9, p_def_empty=0.5,
                  live=True)


if __name__ == "__main__":
    if "--random" in sys.argv:
        main_random()
    elif "--round2" in sys.argv:
        main_round2()
    elif "--masked" in sys.argv:
        main_round2(only_masked=True)
    elif "--live" in sys.argv:
        main_live()
    else:
        main()
