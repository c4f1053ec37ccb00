#!/usr/bin/env python
"""bench.py -- attacker+defender env-steps/sec of the batched step, against the HBM roofline.

    python bench.py --gpus 1 --steps 200 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU arm (oracle port of the reference step on the host cores)

Workload (BASELINE.json configs[2]): CyberBattleToyCtf-v0 with MARLon bounds (N=12, C=10), the
MARLon attacker+defender pair step (AttackerEnvWrapper.step + DefenderEnvWrapper.step, reference_stale binding, SB3
auto-reset), dense int8 action masks, 65536 envs per GPU; random VALID attacker actions and uniform defender actions
(recorded once with the on-device sampler, then replayed: the env dynamics are deterministic so the replay is valid).
One "step" = one launch of the fused step kernel over all envs of the rank.  Env instances shard across ranks with
no data-path collective ("weak" scaling: fixed envs per GPU); the only collective is the per-rollout all-reduce of
the 16-slot episode-statistics vector.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


WORKLOADS = {
    "toyctf": "CyberBattleToyCtf-v0 (N=12,C=10) MARLon attacker+defender pair step (AttackerEnvWrapper + DefenderEnvWrapper, "
              "reference_stale defender, SB3 auto-reset)",
    "chain100": "CyberBattleChain-v0 size=100 (N=102,C=102) MARLon attacker+defender pair step (config 4: the 1M-env sharded case; "
                "factored masks -- a dense connect mask would be 8.5 MB per env)",
    "toyctf_scan": "CyberBattleToyCtf-v0 (N=12,C=10) CyberBattleEnv step with the built-in ScanAndReimageCompromisedMachines(0.6, 2, 5) "
                   "defender (configs[2] to the letter: notebook_withdefender.py:57-63 parameters, SLA 0.80, Philox detection draws, "
                   "auto-reset as under the SB3 VecEnv adapter)",
    "chain100_scan": "CyberBattleChain-v0 size=100 (N=102,C=102) CyberBattleEnv step with the built-in ScanAndReimageCompromisedMachines(0.6, 2, 5) "
                     "defender (config 4's other variant: SLA 0.80, Philox detection draws, auto-reset; factored masks)",
    "toyctf_live": "CyberBattleToyCtf-v0 (N=12,C=10) MARLon attacker+defender pair step with the LIVE LearningDefender binding (SURVEY 8f row 4: "
                   "the defender re-images / blocks / allows on the environment the attacker plays in; firewall rule lists are per-env state)",
    "random16": "CyberBattleRandom-v0, 16 generated 65-node networks (seeds 0-15) side by side in one batch (config 5: padded layout, "
                "N=72, C=192, 32 leak slots), MARLon attacker+defender pair step",
}


def workload_config(mask_mode=0, workload="toyctf"):
    from marlon_b200 import _abi, config, scenario, scenarios

    if workload == "toyctf_scan":
        comp = scenario.compile_scenario(scenarios.toyctf_environment())
        cfg = config.make_config(
            _abi.MODE_CYBERBATTLE, maximum_node_count=12, maximum_total_credentials=10, maximum_discoverable_credentials_per_action=5,
            throws_on_invalid_actions=False, attacker_goal=config.AttackerGoal(own_atleast=6),
            defender_agent=config.ScanAndReimageCompromisedMachines(0.6, 2, 5), defender_constraint=config.DefenderConstraint(0.80),
            auto_reset=True, mask_mode=mask_mode, seed=2026)
        return comp, cfg
    if workload == "chain100_scan":
        comp = scenario.compile_scenario(scenarios.chain_environment(100))
        cfg = config.make_config(
            _abi.MODE_CYBERBATTLE, maximum_node_count=102, maximum_total_credentials=102, maximum_discoverable_credentials_per_action=5,
            throws_on_invalid_actions=False, attacker_goal=config.AttackerGoal(own_atleast_percent=1.0),
            defender_agent=config.ScanAndReimageCompromisedMachines(0.6, 2, 5), defender_constraint=config.DefenderConstraint(0.80),
            auto_reset=True, mask_mode=1, seed=2026)
        return comp, cfg
    if workload == "random16":
        from marlon_b200 import random_network

        comps = [scenario.compile_scenario(random_network.random_environment(sd)) for sd in range(16)]
        cfg = config.make_config(
            _abi.MODE_MARLON, maximum_node_count=72, maximum_total_credentials=192, maximum_discoverable_credentials_per_action=32,
            throws_on_invalid_actions=False, attacker_goal=config.AttackerGoal(own_atleast_percent=1.0),
            defender_constraint=config.DefenderConstraint(0.60), losing_reward=-5000.0,
            attacker_max_timesteps=2000, attacker_invalid_action_reward_modifier=-1.0,
            defender_enabled=True, defender_max_timesteps=2000, defender_invalid_action_reward=-1,
            defender_reset_on_constraint_broken=True, defender_loss_reward=-5000.0, mask_mode=1)
        return comps, cfg
    if workload == "chain100":
        comp = scenario.compile_scenario(scenarios.chain_environment(100))
        cfg = config.make_config(
            _abi.MODE_MARLON, maximum_node_count=102, maximum_total_credentials=102, maximum_discoverable_credentials_per_action=5,
            throws_on_invalid_actions=False, attacker_goal=config.AttackerGoal(own_atleast_percent=1.0),
            defender_constraint=config.DefenderConstraint(0.60), losing_reward=-5000.0,
            attacker_max_timesteps=2000, attacker_invalid_action_reward_modifier=-1.0,
            defender_enabled=True, defender_max_timesteps=2000, defender_invalid_action_reward=-1,
            defender_reset_on_constraint_broken=True, defender_loss_reward=-5000.0, mask_mode=1)
        return comp, cfg
    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    # MultiAgentUniverse.build defaults (multiagent_universe.py:78-95) with ppo/train_marl.py:12-14 bounds
    cfg = config.make_config(
        _abi.MODE_MARLON, maximum_node_count=12, maximum_total_credentials=10, maximum_discoverable_credentials_per_action=5,
        throws_on_invalid_actions=False, attacker_goal=config.AttackerGoal(own_atleast=6),
        defender_constraint=config.DefenderConstraint(0.60), losing_reward=-5000.0,
        attacker_max_timesteps=2000, attacker_invalid_action_reward_modifier=-1.0,
        defender_enabled=True, defender_max_timesteps=2000, defender_invalid_action_reward=-1,
        defender_reset_on_constraint_broken=True, defender_loss_reward=-5000.0, mask_mode=mask_mode,
        **({"defender_binding": "live"} if workload == "toyctf_live" else {}))
    return comp, cfg


def algorithmic_bytes_per_env_step(comp, cfg, S_words):
    """reference dtypes (int8 masks, int32 other observation fields), each state byte read once and
    written once, actions read, rewards/flags written, no re-reads."""
    n, N, C, LEAK = comp.n_nodes, cfg.maximum_node_count, cfg.maximum_total_credentials, cfg.maximum_discoverable_credentials_per_action
    ident = comp.identifiers
    P, L, R, props = len(ident.ports), len(ident.local_vulnerabilities), len(ident.remote_vulnerabilities), len(ident.properties)
    att_small = 4 * (7 + 4 * LEAK + 2 * C + N * props + N)
    masks = (N * L + N * N * R + N * N * P * C) if cfg.mask_mode == 0 else 4 * ((N + 31) // 32)
    defender = (n + 12 * n + comp.n_services) if cfg.def_enabled else 0
    state = 2 * 4 * S_words
    io = (40 if cfg.mode == 1 else 20) + (48 if cfg.def_enabled else 0) + 4 + 4 + 4  # actions (MARLon 10 / CyberBattleEnv 5 words), rewards, flags
    return dict(attacker_obs=att_small, masks=masks, defender_obs=defender, state_rw=state, actions_rewards_flags=io,
                total=att_small + masks + defender + state + io)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        """Start sampling (call well before the timed region: nvidia-smi needs ~0.1-0.3 s before its first row)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=3.0):
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        t0, t1 = getattr(self, "t_begin", 0.0), getattr(self, "t_end", float("inf"))
        inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.012]  # a row describes the ~10 ms before it was printed
        window = "timed region"
        if not inside:  # region shorter than the sampling period: the rows right around it
            before = [r for t, r in self.rows if t < t0][-1:]
            after = [r for t, r in self.rows if t > t1][:2]
            inside, window = before + after, "rows adjacent to the timed region (region shorter than the 10 ms sampling period)"
        sm, mx, reasons = [], [], set()
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference step on the host cores (one process per core, numpy valid-action sampler)
# ------------------------------------------------------------------------------------------------------------------
def numpy_valid_actions(arrays, cfg, rng, n_nodes):
    """Uniform valid attacker actions + uniform defender actions from the observation arrays (host side)."""
    from marlon_b200 import _abi, config

    sc = arrays["scalars"]
    n = sc.shape[0]
    nd, nc = sc[:, 6], sc[:, 5]
    owned = arrays["owned_bits"][:, 0]
    N, L = arrays["local_vulnerability"].shape[1:]
    R, P = arrays["remote_vulnerability"].shape[3], arrays["connect"].shape[3]
    lay = config.attacker_action_layout(cfg)
    idx_of_kind = {cfg.kind_of_index[i]: i for i in range(3)}
    att = np.zeros((n, 10), dtype=np.int32)
    kind = (rng.random(n) * np.where(nc > 0, 3, 2)).astype(np.int64)
    bits = ((owned[:, None] >> np.arange(N)[None, :]) & 1).astype(bool)
    pick = rng.random((n, N)) * bits
    src = pick.argmax(1)
    loc = arrays["local_vulnerability"][np.arange(n), src]  # [n, L]
    lv = (rng.random((n, L)) * (loc > 0)).argmax(1)
    kind = np.where((kind == 0) & (loc.sum(1) == 0), 1, kind)
    tgt = (rng.random(n) * np.maximum(nd, 1)).astype(np.int64)
    for k in (0, 1, 2):
        m = kind == k
        a0 = lay[k][0]
        att[m, 0] = idx_of_kind[k]
        att[m, a0] = src[m]
        if k == _abi.KIND_LOCAL:
            att[m, a0 + 1] = lv[m]
        elif k == _abi.KIND_REMOTE:
            att[m, a0 + 1] = tgt[m]
            att[m, a0 + 2] = (rng.random(m.sum()) * R).astype(np.int64)
        else:
            att[m, a0 + 1] = tgt[m]
            att[m, a0 + 2] = (rng.random(m.sum()) * P).astype(np.int64)
            att[m, a0 + 3] = (rng.random(m.sum()) * np.maximum(nc[m], 1)).astype(np.int64)
    nvec = np.array([5, n_nodes, n_nodes, 6, 2, n_nodes, 6, 2, n_nodes, 3, n_nodes, 3])
    dfn = (rng.random((n, 12)) * nvec).astype(np.int32)
    return att, dfn


def _cpu_worker(args):
    seed, n_envs, seconds = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import OracleBatch

    comp, cfg = workload_config()
    o = OracleBatch(comp, cfg, n_envs)
    o.reset()
    rng = np.random.default_rng(seed)
    steps, spent = 0, 0.0
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        att, dfn = numpy_valid_actions(o.arrays, cfg, rng, comp.n_nodes)
        t0 = time.perf_counter()
        o.step(att, dfn)
        spent += time.perf_counter() - t0
        steps += 1
    return steps * n_envs, spent


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(seconds=12.0, n_envs=256, cores=None):
    """Oracle (C port of the reference step) on every host core at once; only the step call is timed."""
    import multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle

    oracle.build()
    cores = cores or len(os.sched_getaffinity(0))
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(1000 + i, n_envs, seconds) for i in range(cores)])
    value = sum(s / t for s, t in res if t > 0)
    return {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "cpu_model": _cpu_model(), "per_core": value / cores,
            "sample": f"{cores} processes x {n_envs} envs, ToyCtf(12,10) MARLon attacker+defender pair step, random valid actions, "
                      f"{seconds:.0f} s wall each, only oracle step() timed ({sum(s for s, _ in res)} env-steps)"}


def workload_label(workload, factored):
    """The `config.workload` string: one function for both arms, so that the driver's same_config comparison holds."""
    return WORKLOADS[workload] + ", " + ("factored" if factored else "dense int8") + " action masks, random valid actions"


def python_reference_baseline():
    """The north star's CPU baseline -- the UNMODIFIED Python reference under multiprocessing.Pool on every core -- cannot run
    on the GPU box (/root/reference does not travel): it was timed on the dev container by oracle/time_python_reference.py
    and is carried from profiles/r02_python_reference_cpu.json, labelled as such."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_python_reference_cpu.json")))
        d["note"] = "measured on the dev container (no GPU), not on this box: " + d.get("note", "")
        return d
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    per = max(2.0, min(20.0, 1.5 * (args.steps + args.warmup) / 10.0))
    cb = cpu_baseline(seconds=per)
    line = {
        "impl": "reference", "metric": "attacker+defender env-steps/sec", "value": cb["value"], "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_label("toyctf", False),
                   "note": "reference arm = CPU oracle port of the reference step (the reference is pure Python and does not travel); "
                           "one process per host core, bounded sample"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    pr = python_reference_baseline()
    if pr:
        line["cpu_baseline"]["python_reference"] = pr
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def pin_to_gpu_numa(local):
    """Run this rank on the CPUs next to its GPU (NVML's affinity mask for the device): the page-locked action / result
    buffers are then allocated on that NUMA node and the per-step launch + sync loop does not cross sockets."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {i * 64 + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"pinned_cpus": len(cpus), "first_cpu": min(cpus) if cpus else None}
    except Exception as e:  # noqa: BLE001
        return {"pinned_cpus": 0, "error": str(e)[:80]}


def record_tape(comp, cfg, counts, n, steps, seed, local, dev):
    """Synthetic action tape: valid attacker actions for the evolving state (sampler kernel) + uniform defender actions,
    recorded once on a scratch batch -- the env dynamics are deterministic, so replaying it on a fresh batch stays valid."""
    import torch

    from marlon_b200 import _abi
    from marlon_b200.batch import Batch

    rec = Batch(comp, cfg, counts, device=local)
    rec.reset()
    aw = rec.att_width  # 10: MARLon MultiDiscrete attacker action; 5: CyberBattleEnv [kind, 4 coordinates]
    has_def = bool(cfg.mode == _abi.MODE_MARLON and cfg.def_enabled)
    tape_a = torch.empty((steps, n, aw), dtype=torch.int32, device=dev)
    tape_d = torch.empty((steps, n, 12), dtype=torch.int32, device=dev) if has_def else [None] * steps
    for s in range(steps):
        rec.sample_actions(seed=seed, attacker_out=tape_a[s], defender_out=tape_d[s] if has_def else None)
        rec.step(tape_a[s], tape_d[s])
    torch.cuda.synchronize()
    rec.close()
    return tape_a, tape_d, aw, has_def


def in_place_kernel_bytes(comp, cfg, ab):
    """Algorithmic bytes per env-step of the IN-PLACE kernel (cbx_wide_kernel), from the SURVEY 8(d) figure `ab`:
      * state: only the words a step has to touch -- the wrappers' words, the stale copy's countdowns and the header read +
        written; discovery order / inverse map, installed bits, privilege levels, property bitsets and the credential cache
        read once to build the observation.  The rest of the per-env state (attacked bits, live countdowns, gathered / cached
        bitsets, ever-owned, not-running) is touched only by the few envs whose action needs it and is NOT counted;
      * defender observation: only infected_nodes -- the firewall / service rows cannot change under the reference's stale
        defender binding and are written when an env is created or reset, not by a step.
    Both corrections LOWER the numerator; the SURVEY figure is kept under `survey_8d_total`."""
    n = comp.n_nodes
    ident = comp.identifiers
    props = len(ident.properties)
    Wn, PW = (n + 31) // 32, (props + 31) // 32
    ntr = len(comp.triples)
    rw = 13 + (n + 3) // 4 + 2                     # wrapper words + shadow countdowns + stepcount / episode sum
    ro = Wn + (n + 15) // 16 + 2 * ((n + 3) // 4) + n * PW + (ntr + 2) // 2
    out = dict(ab)
    out["survey_8d_total"] = ab["total"]
    out["state_rw"] = 4 * (2 * rw + ro)
    if cfg.def_enabled:
        out["defender_obs"] = n
    out["total"] = sum(out[k] for k in ("attacker_obs", "masks", "defender_obs", "state_rw", "actions_rewards_flags"))
    out["note"] = ("in-place kernel: state = wrapper/header/stale-copy words read+written + observation sources read once; defender "
                   "observation = infected_nodes only (static firewall / service rows are written at reset, not per step)")
    return out


def measure_device(args, workload, n_req, K, W, world, rank, local, dev, clocks=None):
    """W untimed + K timed device-resident steps of `workload` with n_req envs on this rank, then the per-rollout statistics
    all-reduce INSIDE the timed region.  -> dict (rank 0 fills the derived numbers)."""
    import torch
    import torch.distributed as dist

    from marlon_b200 import _abi
    from marlon_b200.batch import Batch

    comp, cfg = workload_config(mask_mode=1 if args.factored else 0, workload=workload)
    factored = bool(cfg.mask_mode)
    multi = isinstance(comp, list)
    counts = [n_req // len(comp)] * len(comp) if multi else n_req
    n = sum(counts) if multi else n_req
    tape_a, tape_d, aw, has_def = record_tape(comp, cfg, counts, n, W + K, args.seed + rank, local, dev)
    b = Batch(comp, cfg, counts, device=local)
    b.reset()
    for s in range(W):
        b.step(tape_a[s], tape_d[s])
    warm = b.stats_tensor.clone()  # the collective's path once before the clock starts (allocator, NCCL channel set-up)
    if world > 1:
        dist.all_reduce(warm, op=dist.ReduceOp.SUM)
    b.stats_reset()
    b.enable_timing(True)
    launches0 = b.launch_count
    torch.cuda.synchronize()
    if clocks is not None:
        clocks.wait_first()
    if world > 1:
        dist.barrier()
    ev0, evc, ev1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    if clocks is not None:
        clocks.mark_begin()
    ev0.record()
    for s in range(W, W + K):
        b.step(tape_a[s], tape_d[s])
    b.enable_timing(False)  # closes the kernel-time bracket here (an event on the stream, no synchronisation)
    evc.record()
    # the only collective of the path: one SUM all-reduce of the 16-slot episode-statistics vector per rollout
    stats = b.stats_tensor.clone()
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    ev1.record()
    torch.cuda.synchronize()
    if clocks is not None:
        clocks.mark_end()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1), evc.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms, coll_ms = (float(x) for x in ms.tolist())
    launches = b.launch_count - launches0
    kernel_ms, kernel_n = b.step_kernel_ms()
    kinfo = b.kernel_info()
    b.close()
    out = {"comp": comp, "cfg": cfg, "counts": counts, "n": n, "tape_a": tape_a, "tape_d": tape_d, "aw": aw, "has_def": has_def,
           "factored": factored, "multi": multi, "total_ms": total_ms, "collective_ms": coll_ms, "launches": launches,
           "kernel_ms": kernel_ms, "kernel_n": kernel_n, "kinfo": kinfo, "stats": stats.cpu().numpy()}
    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        packed_state_words = int(os.environ.get("CBX_STATE_WORDS", "0")) or None
        if multi:  # padded layout: the largest scenario's dimensions, as the library lays the batch out
            import copy

            big = copy.copy(max(comp, key=lambda c: c.n_nodes))
            big.n_services = max(c.n_services for c in comp)
            S_packed = packed_state_words or max(_packed_state_words(c, cfg) for c in comp)
            ab = algorithmic_bytes_per_env_step(big, cfg, S_packed)
            ref_comp = big
        else:
            S_packed = packed_state_words or _packed_state_words(comp, cfg)
            ab = algorithmic_bytes_per_env_step(comp, cfg, S_packed)
            ref_comp = comp
        if kinfo["name"] == "cbx_wide_kernel":  # in-place kernel: only what a step has to move counts
            ab = in_place_kernel_bytes(ref_comp, cfg, ab)
        achieved = ab["total"] * n / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
        total_envs = n * world
        out.update({
            "value": total_envs * K / (total_ms * 1e-3), "ms_per_step": total_ms / K, "total_envs": total_envs, "ab": ab,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": ncu_traffic(kinfo["name"], workload, n, factored),
                         "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                         "kernel": kinfo["name"], "kernel_launch": kinfo, "kernel_ms": kernel_ms, "kernel_launches_timed": kernel_n}})
    return out


class HostTape:
    """The action tape in page-locked HOST memory, CH steps at a time (bounded pinned memory); refills are outside the timed calls."""

    def __init__(self, m, dtype, rows=None, CH=64):
        import torch

        self.m, self.dtype, self.CH = m, dtype, CH
        self.rows = rows if rows is not None else slice(None)
        n = m["tape_a"][0][self.rows].shape[0]
        self.h_a = torch.empty((CH, n, m["aw"]), dtype=dtype, pin_memory=True)
        self.h_d = torch.empty((CH, n, 12), dtype=dtype, pin_memory=True) if m["has_def"] else None
        self.a_np = self.h_a.numpy()
        self.d_np = self.h_d.numpy() if m["has_def"] else None
        self.base = -1

    def get(self, s):
        """-> (attacker actions, defender actions or None) of tape step s as numpy views of the pinned chunk."""
        import torch

        c0 = (s // self.CH) * self.CH
        if c0 != self.base:
            c1 = min(len(self.m["tape_a"]), c0 + self.CH)
            self.h_a[: c1 - c0].copy_(self.m["tape_a"][c0:c1, self.rows].to(self.dtype))
            if self.h_d is not None:
                self.h_d[: c1 - c0].copy_(self.m["tape_d"][c0:c1, self.rows].to(self.dtype))
            torch.cuda.synchronize()
            self.base = c0
        return self.a_np[s - c0], (self.d_np[s - c0] if self.d_np is not None else None)


def _max_over_ranks(secs, world, dev):
    import torch
    import torch.distributed as dist

    t = torch.tensor([secs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def e2e_results_only(m, K, W, dtype, world, local, dev):
    """K host-buffer steps (cbx_batch_step_host[_i16]) of a fresh batch; W untimed calls of the SAME entry point first, all
    allocations done before (host_prepare).  -> seconds, max over ranks."""
    import torch
    import torch.distributed as dist

    from marlon_b200.batch import Batch

    tape = HostTape(m, dtype)
    b2 = Batch(m["comp"], m["cfg"], m["counts"], device=local)
    b2.reset()
    b2.host_prepare()
    for s in range(W):
        b2.step_host(*tape.get(s))
    tape.get(W)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    secs, out = 0.0, None
    s = W
    while s < W + K:
        tape.get(s)  # refill (untimed) when s enters a new chunk
        e = min(W + K, (s // tape.CH + 1) * tape.CH)
        t0 = time.perf_counter()
        for q in range(s, e):
            out = b2.step_host(*tape.get(q))
        secs += time.perf_counter() - t0
        s = e
    n = m["n"]
    assert out["att_reward"].shape[0] == n
    # the host-side results must be the device's (the kernel wrote both)
    assert (out["att_reward"] == b2.numpy("att_reward")).all() and (out["def_truncated"] == b2.numpy("def_truncated")).all()
    b2.close()
    return _max_over_ranks(secs, world, dev)


def e2e_results_two_halves(m, K, W, world, local, dev):
    """The results-only host-buffer step as a double-buffered host loop drives it: the envs are two half batches on two
    streams; the loop waits for half A's results of step s-1 (they are in host memory then: the point where a policy would
    read them and write A's next actions), submits A's step s, and does the same for B -- so the GPU works on one half while
    the host handles the other.  Every step of a half is submitted only after that half's previous results have reached
    the host.  -> seconds for K steps of all envs, max over ranks (None when the batch does not split)."""
    import copy

    import torch
    import torch.distributed as dist

    from marlon_b200.batch import Batch

    n = m["n"]
    if m["multi"] or n % 128:
        return None
    h = n // 2
    parts = []
    for q in range(2):
        cfg = copy.copy(m["cfg"])
        cfg.env_index_base = m["cfg"].env_index_base + q * h
        bq = Batch(m["comp"], cfg, h, device=local)
        bq.reset()
        bq.host_prepare()
        parts.append((bq, HostTape(m, torch.int16, rows=slice(q * h, (q + 1) * h)), torch.cuda.Stream(device=dev)))
    outs = [None, None]

    def one(s):
        for q, (bq, tape, st) in enumerate(parts):
            a, d = tape.get(s)
            st.synchronize()  # this half's previous results are on the host
            with torch.cuda.stream(st):
                outs[q] = bq.step_host(a, d, sync=False)

    for s in range(W):
        one(s)
    for _, tape, st in parts:
        st.synchronize()
        tape.get(W)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    secs, s = 0.0, W
    CH = parts[0][1].CH
    while s < W + K:
        for _, tape, st in parts:
            st.synchronize()
            tape.get(s)
        e = min(W + K, (s // CH + 1) * CH)
        t0 = time.perf_counter()
        for q in range(s, e):
            one(q)
        for _, _, st in parts:
            st.synchronize()
        secs += time.perf_counter() - t0
        s = e
    for (bq, _, _), out in zip(parts, outs):
        assert (out["att_reward"] == bq.numpy("att_reward")).all() and (out["def_truncated"] == bq.numpy("def_truncated")).all()
        bq.close()
    return _max_over_ranks(secs, world, dev)


def e2e_obs_factored(m, K, W, world, local, dev, halves):
    """Host-buffer steps that also bring the observation a host-side policy needs to page-locked host memory every step:
    scalars, leaked credentials, credential cache, property matrix, privilege levels, owned bits (= the factored action masks,
    SURVEY.md A.4), the defender's observation, rewards and done flags (cbx_batch_fetch_host).  halves = 1: step, fetch,
    synchronise.  halves = 2: the envs are two half batches on two streams, so one half's copies overlap the other half's
    step (what a host loop that acts on half A while half B steps would see).  -> (seconds max over ranks, D2H bytes/step)."""
    import torch
    import torch.distributed as dist

    from marlon_b200 import _abi
    from marlon_b200.batch import Batch

    n = m["n"]
    if m["multi"] or n % (64 * halves):
        return None, 0
    h = n // halves
    import copy

    parts = []
    for q in range(halves):
        cfg = copy.copy(m["cfg"])
        cfg.env_index_base = m["cfg"].env_index_base + q * h
        bq = Batch(m["comp"], cfg, h, device=local)
        bq.reset()
        bq.host_prepare()
        parts.append((bq, HostTape(m, torch.int16, rows=slice(q * h, (q + 1) * h)), torch.cuda.Stream(device=dev)))
    fields = _abi.F_OBS_FACTORED | _abi.F_RESULTS
    d2h = 0

    def one(s):
        nonlocal d2h
        got = []
        for bq, tape, st in parts:
            a, d = tape.get(s)
            with torch.cuda.stream(st):
                bq.step_host(a, d, sync=False)
                got.append(bq.fetch_host(fields, sync=False))
        for _, _, st in parts:
            st.synchronize()
        d2h = sum(sum(v.nbytes for k, v in g.items()) for g in got)
        return got

    for s in range(W):
        one(s)
    for _, tape, _ in parts:
        tape.get(W)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    secs, s, got = 0.0, W, None
    CH = parts[0][1].CH
    while s < W + K:
        for _, tape, _ in parts:
            tape.get(s)
        e = min(W + K, (s // CH + 1) * CH)
        t0 = time.perf_counter()
        for q in range(s, e):
            got = one(q)
        secs += time.perf_counter() - t0
        s = e
    # what reached the host is what the device holds
    for (bq, _, _), g in zip(parts, got):
        assert (g["discovered_nodes_properties"] == bq.numpy("discovered_nodes_properties")).all()
        assert (g["att_reward"] == bq.numpy("att_reward")).all()
        bq.close()
    return _max_over_ranks(secs, world, dev), d2h


def e2e_vecenv(m, K, W, world, local, dev, observations):
    """The SB3 VecEnv adapter: attacker_vec_env.step then defender_vec_env.step per env-step pair (marl_algorithm.py:43-49) on
    host int64 action arrays (what a policy's .cpu().numpy() hands over), infos built.  observations="torch": the policy is on
    the GPU, only rewards / done flags cross PCIe.  "numpy": the full stacked dense observation goes to the host every step."""
    import torch
    import torch.distributed as dist

    from marlon_b200.universe import MultiAgentUniversalEnv

    n = m["n"]
    u = MultiAgentUniversalEnv("CyberBattleToyCtf-v0", n, device=local, maximum_node_count=12, maximum_total_credentials=10,
                               maximum_discoverable_credentials_per_action=5, max_timesteps=2000, emit_terminal_obs=True)
    av = u.vec_env("attacker", observations=observations, terminal_observations="truncated")
    dv = u.vec_env("defender", observations=observations, terminal_observations="truncated")
    av.reset()
    dv.reset()
    ta = m["tape_a"][: W + K].cpu().numpy().astype(np.int64)
    td = m["tape_d"][: W + K].cpu().numpy().astype(np.int64)
    for s in range(W):
        av.step(ta[s])
        dv.step(td[s])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(W, W + K):
        obs, r, done, infos = av.step(ta[s])
        dobs, dr, ddone, dinfos = dv.step(td[s])
    secs = time.perf_counter() - t0
    assert r.shape[0] == n and len(infos) == n
    u.close()
    return _max_over_ranks(secs, world, dev)


def device_rollout(m, world, local, dev, T=32, rounds=3, workload="toyctf"):
    """rollout collection with everything resident in HBM -- observations are the batch's own tensors, two small
    MLP actor-critics (marlon_b200.ppo.MultiDiscretePolicy) pick both agents' MultiDiscrete actions on the device, one fused
    attacker+defender launch per step, rewards / values / log-probabilities into [T, n] device buffers, GAE by cbx_gae.
    -> seconds for rounds x T steps (CUDA events, max over ranks)."""
    import torch
    import torch.distributed as dist

    from marlon_b200 import ppo
    from marlon_b200.rollout import DeviceRolloutBuffer, collect_rollouts
    from marlon_b200.universe import MultiAgentUniversalEnv

    n = m["n"]
    torch.manual_seed(7)
    if workload == "random16":  # configs[4]: 16 generated networks side by side, factored masks (workload_config("random16"))
        u = MultiAgentUniversalEnv("CyberBattleRandom-v0", n // 16, device=local, scenario_seeds=range(16), maximum_node_count=72,
                                   maximum_total_credentials=192, maximum_discoverable_credentials_per_action=32,
                                   max_timesteps=2000, mask_mode="factored")
    else:
        u = MultiAgentUniversalEnv("CyberBattleToyCtf-v0", n, device=local, maximum_node_count=12, maximum_total_credentials=10,
                                   maximum_discoverable_credentials_per_action=5, max_timesteps=2000)
    aobs, dobs = u.reset()
    apol = ppo.MultiDiscretePolicy.for_space(aobs, u.attacker_action_space.nvec, ppo.ATTACKER_FEATURES).to(dev)
    dpol = ppo.MultiDiscretePolicy.for_space(dobs, u.defender_action_space.nvec, ppo.DEFENDER_FEATURES).to(dev)
    ab, db = DeviceRolloutBuffer(T, n, 10, dev), DeviceRolloutBuffer(T, n, 12, dev)
    collect_rollouts(u, apol, ab, dpol, db)  # warm-up: allocator, cuBLAS handles
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        collect_rollouts(u, apol, ab, dpol, db)
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3
    u.close()
    return _max_over_ranks(secs, world, dev), T * rounds


def run_ours(args):
    import torch
    import torch.distributed as dist

    from marlon_b200 import _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pin = pin_to_gpu_numa(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    K, W = args.steps, args.warmup
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    m = measure_device(args, args.workload, args.envs_per_gpu, K, W, world, rank, local, dev, clocks)
    clk = clocks.stop() if clocks else None
    n, aw, has_def = m["n"], m["aw"], m["has_def"]

    # ---- e2e: the same step through the public API with HOST action buffers and host-side results ----
    e2e = None
    if not args.no_e2e and has_def and not m["multi"]:
        Ke = K
        e2e_s = e2e_results_only(m, Ke, W, torch.int16, world, local, dev)
        e2e32_s = e2e_results_only(m, Ke, W, torch.int32, world, local, dev)
        e2e2h_s = e2e_results_two_halves(m, Ke, W, world, local, dev)
        Ko = min(K, 200)
        obs2_s, d2h_obs = e2e_obs_factored(m, Ko, W, world, local, dev, halves=2)
        obs1_s, _ = e2e_obs_factored(m, Ko, W, world, local, dev, halves=1)
        vt_s = vn_s = roll_s = None
        Kv, Kn, Kr = min(K, 100), min(K, 5), 0
        if args.workload == "toyctf" and not m["factored"]:
            roll_s, Kr = device_rollout(m, world, local, dev)
            vt_s = e2e_vecenv(m, Kv, W, world, local, dev, "torch")
            if world == 1:
                vn_s = e2e_vecenv(m, Kn, min(W, 3), world, local, dev, "numpy")
        total_envs = n * world
        act_b = n * (aw + 12)
        e2e = {"value": total_envs * Ke / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": act_b * 2, "d2h_bytes_per_step": n * 12,
               "which": "results_only_int16",
               "variants": {
                   "results_only_int16": {"value": total_envs * Ke / e2e_s, "steps": Ke, "h2d_bytes_per_step": act_b * 2, "d2h_bytes_per_step": n * 12},
                   "results_only_int32": {"value": total_envs * Ke / e2e32_s, "steps": Ke, "h2d_bytes_per_step": act_b * 4, "d2h_bytes_per_step": n * 12},
                   "results_only_int16_two_halves": None if e2e2h_s is None else {
                       "value": total_envs * Ke / e2e2h_s, "steps": Ke, "h2d_bytes_per_step": act_b * 2, "d2h_bytes_per_step": n * 12},
                   "obs_factored_two_halves": None if obs2_s is None else {
                       "value": total_envs * Ko / obs2_s, "steps": Ko, "h2d_bytes_per_step": act_b * 2, "d2h_bytes_per_step": d2h_obs},
                   "obs_factored_sequential": None if obs1_s is None else {
                       "value": total_envs * Ko / obs1_s, "steps": Ko, "h2d_bytes_per_step": act_b * 2, "d2h_bytes_per_step": d2h_obs},
                   "device_rollout_mlp_policies": None if roll_s is None else {
                       "value": total_envs * Kr / roll_s, "steps": Kr, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   "vecenv_torch_obs": None if vt_s is None else {
                       "value": total_envs * Kv / vt_s, "steps": Kv, "h2d_bytes_per_step": act_b * 4, "d2h_bytes_per_step": n * 24},
                   "vecenv_numpy_dense_obs": None if vn_s is None else {
                       "value": total_envs * Kn / vn_s, "steps": Kn, "h2d_bytes_per_step": act_b * 4,
                       "d2h_bytes_per_step": n * (m["ab"]["attacker_obs"] + m["ab"]["masks"] + m["ab"]["defender_obs"] + 24) if rank == 0 else None}},
               "note": "value = results_only_int16: cbx_batch_step_host_i16 per step with HOST action buffers (int16 elements) in "
                       "page-locked memory -- the kernel reads each tile's actions over PCIe in place (TMA bulk loads from the mapped "
                       "host buffers) and writes rewards + done flags straight into a rotating page-locked result block, stream sync; "
                       "observations stay in HBM as torch tensors (the consumer is a GPU-resident policy); results_only_int16_two_halves is the "
                       "same work as a double-buffered host loop submits it: two half batches on two streams, a half's next step "
                       "goes out only after its previous results have reached the host, so the GPU steps one half while the host "
                       "handles the other (the single-batch figure pays a full launch + pipeline fill + sync per step).  For a HOST-side policy "
                       "the observation has to cross PCIe as well: obs_factored_* bring every small field + the factored masks + the "
                       "defender observation to pinned host memory each step (cbx_batch_fetch_host; two_halves overlaps one half "
                       "batch's copies with the other's step), device_rollout_mlp_policies is rollout.collect_rollouts with two small MLP "
                       "actor-critics choosing both agents' actions on the device and GAE by cbx_gae (nothing crosses PCIe), vecenv_* go through the SB3 VecEnv adapter (attacker step + defender "
                       "step, int64 numpy actions, infos built; numpy_dense_obs copies the full dense observation, 12.4 KB per env, "
                       "every step).  All variants: allocations and W warm-up calls of the same entry point before the clock starts."}

    # ---- config 4 as BASELINE.json names it, when the job has more than one GPU: Chain-100, 1 048 576 envs in total ----
    cfg4 = None
    if world > 1 and not args.no_config4:
        K4 = min(K, 50)
        per = (1048576 // world) // 32 * 32
        m4 = measure_device(args, "chain100", per, K4, W, world, rank, local, dev)
        if rank == 0:
            cfg4 = {"workload": workload_label("chain100", True), "envs_per_gpu": m4["n"], "total_envs": m4["total_envs"], "steps": K4,
                    "value": m4["value"], "unit": "env-steps/s", "ms_per_step": m4["ms_per_step"], "collective_ms": m4["collective_ms"],
                    "roofline": m4["roofline"], "algorithmic_bytes_per_env_step": m4["ab"], "gpu_launches": m4["launches"]}

    # ---- config 5 as BASELINE.json names it (generated networks, PPO rollout across the GPUs): the step kernel on 16 networks
    # side by side, then rollout collection with MLP policies on the device (rollout.collect_rollouts + cbx_gae) ----
    cfg5 = None
    if (world > 1 and not args.no_config4) or args.config5:
        K5 = min(K, 50)
        m5 = measure_device(args, "random16", 131072, K5, W, world, rank, local, dev)
        roll5_s, Kr5 = device_rollout(m5, world, local, dev, T=16, rounds=2, workload="random16")
        if rank == 0:
            cfg5 = {"workload": workload_label("random16", True), "envs_per_gpu": m5["n"], "total_envs": m5["total_envs"], "steps": K5,
                    "value": m5["value"], "unit": "env-steps/s", "ms_per_step": m5["ms_per_step"], "collective_ms": m5["collective_ms"],
                    "roofline": m5["roofline"], "algorithmic_bytes_per_env_step": m5["ab"], "gpu_launches": m5["launches"],
                    "device_rollout_mlp_policies": {"value": m5["total_envs"] * Kr5 / roll5_s, "unit": "env-steps/s", "steps": Kr5,
                                                    "note": "rollout.collect_rollouts: two MLP actor-critics choose both agents' actions on "
                                                            "the device, one fused launch per step, GAE by cbx_gae; nothing crosses PCIe"}}

    if rank == 0:
        ab = m["ab"]
        line = {
            "metric": "attacker+defender env-steps/sec", "value": m["value"], "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_label(args.workload, m["factored"]),
                       "envs_per_gpu": n, "total_envs": m["total_envs"], "parallelism": f"env-sharded x{world}, no data-path collective",
                       "l2": "per-step output (%.0f MB/GPU) larger than the 126 MB L2; no flush needed" % (ab["total"] * n / 1e6),
                       "algorithmic_bytes_per_env_step": ab, "numa": pin},
            "roofline": m["roofline"],
            "collective_ms": m["collective_ms"],
            "collective": "one SUM all-reduce of the 16-slot fp64 episode-statistics vector per K-step rollout, inside the timed region"
                          + ("" if world > 1 else " (single rank: the clone only, nothing to reduce)"),
            "gpu_launches": m["launches"],
            "clocks": clk,
            "episode_stats": {k: float(v) for k, v in zip(_abi.STAT_NAMES, m["stats"])},
        }
        if e2e:
            line["e2e"] = e2e
        if cfg4:
            line["config4"] = cfg4
        if cfg5:
            line["config5"] = cfg5
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(seconds=args.cpu_seconds)
            pr = python_reference_baseline()
            if pr:
                line["cpu_baseline"]["python_reference"] = pr
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(kernel, workload, envs, factored):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel from the committed `ncu --set full` captures
    (profiles/ncu_traffic.json, written by scripts/ncu_summary.py: one entry per kernel + workload); None when no capture
    matches this kernel, workload and batch size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        for c in (t if isinstance(t, list) else [t]):
            if (c.get("kernel") == kernel and c.get("workload", "toyctf") == workload and int(c.get("envs", -1)) == int(envs)
                    and bool(c.get("factored", False)) == bool(factored)):
                return float(c["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def _packed_state_words(comp, cfg):
    """Words of the packed per-env state (mirrors compute_layout in csrc/cbx_api.cu)."""
    n = comp.n_nodes
    ident = comp.identifiers
    L, R, props = len(ident.local_vulnerabilities), len(ident.remote_vulnerabilities), len(ident.properties)
    Wn, PW, AW = (n + 31) // 32, (props + 31) // 32, (2 * (L + R) + 31) // 32
    nsec, ntr = max(1, len(comp.secrets)), len(comp.triples)
    has_tags = int(comp.blob[12]) & 1
    fw = 2 * int(comp.fw_ext[2]) if getattr(cfg, "def_binding", 0) and getattr(comp, "fw_ext", None) is not None else 0  # live: 2 words per rule-list group
    return (13 + (n + 3) // 4 + 2 + 3 * Wn + (n + 15) // 16 + (((n + 7) // 8) if has_tags else 0) + 3 * ((n + 3) // 4) + fw
            + n * PW + n * AW + (nsec + 31) // 32 + max(1, (ntr + 31) // 32) + (ntr + 2) // 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=65536)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--factored", action="store_true", help="factored masks instead of dense int8 masks")
    ap.add_argument("--workload", default="toyctf", choices=sorted(WORKLOADS), help="toyctf = the headline configuration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (kernel experiments)")
    ap.add_argument("--no-config4", action="store_true", help="multi-GPU runs: skip the Chain-100 / 1M-env and generated-network blocks")
    ap.add_argument("--config5", action="store_true", help="add the generated-network block (config 5) on a single GPU too")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
