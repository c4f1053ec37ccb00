"""Time the UNMODIFIED Python reference on this machine's CPU cores -- the north star's CPU baseline.

    python oracle/time_python_reference.py [--seconds 12] [--out profiles/r02_python_reference_cpu.json]

TEST / MEASUREMENT INFRASTRUCTURE; runs only where /root/reference exists (the dev container: the reference is pure Python
with pinned dependencies that are absent from the image, so it cannot travel to the GPU box -- bench.py carries the JSON this
script writes as ``cpu_baseline.python_reference``, labelled "measured on the dev container").

What runs, per worker process (``multiprocessing.Pool``, one per core; SB3's SubprocVecEnv is not installable here, the
fallback BASELINE.md section 3 names): the reference's ``CyberBattleToyCtf-v0`` env with MARLon's bounds (12, 10) behind the
reference's ``AttackerEnvWrapper`` + ``DefenderEnvWrapper`` (``MultiAgentUniverse.build`` defaults,
multiagent_universe.py:78-95,159-198), stepped as ``marl_algorithm.collect_rollouts`` does (attacker step, defender step,
DummyVecEnv-style reset on done; marl_algorithm.py:43-49) with random VALID attacker actions (``env.sample_valid_action``)
and uniform defender actions.  3 warm-up iterations, then `seconds` of wall time; one env-step = one attacker+defender pair.
A second figure times the bare ``CyberBattleEnv.step`` with the built-in ScanAndReimage defender (configs[2] to the letter).
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _worker(args):
    seed, seconds, mode = args
    import logging

    logging.disable(logging.CRITICAL)
    import numpy as np

    import ref_loader

    ref_loader.load()
    from cyberbattle._env import cyberbattle_env as ref_env
    from cyberbattle._env import defender as ref_defender
    from marlon.baseline_models.env_wrappers.attack_wrapper import AttackerEnvWrapper
    from marlon.baseline_models.env_wrappers.defend_wrapper import DefenderEnvWrapper
    from marlon.baseline_models.env_wrappers.environment_event_source import EnvironmentEventSource

    rng = np.random.default_rng(seed)
    bounds = dict(maximum_node_count=12, maximum_total_credentials=10, maximum_discoverable_credentials_per_action=5,
                  throws_on_invalid_actions=False)
    if mode == "pair":
        env = ref_loader.make("CyberBattleToyCtf-v0", defender_constraint=ref_env.DefenderConstraint(maintain_sla=0.60),
                              losing_reward=-5000.0, **bounds)
        es = EnvironmentEventSource()
        att = AttackerEnvWrapper(env, es, max_timesteps=2000, invalid_action_reward_modifier=-1.0, invalid_action_reward_multiplier=1.0,
                                 loss_reward=-5000.0)
        dfn = DefenderEnvWrapper(env, att, es, defender=True, max_timesteps=2000, invalid_action_reward=-1, reset_on_constraint_broken=True,
                                 loss_reward=-5000.0)
        att.reset()
        dfn.reset()
        sub = att.action_subspaces
        a_nvec, d_nvec = [int(x) for x in att.action_space.nvec], [int(x) for x in dfn.action_space.nvec]

        def one():
            va = env.sample_valid_action(kinds=[0, 1, 2])
            kind = next(iter(va))
            a = np.zeros(len(a_nvec), dtype=np.int64)
            ki = [i for i in range(3) if sub[i][0] == kind][0]
            a[0] = ki
            a[sub[ki][1]:sub[ki][2]] = va[kind]
            _, _, term, trunc, _ = att.step(a)
            if term or trunc:
                att.reset()
            d = np.array([int(rng.integers(m)) for m in d_nvec], dtype=np.int64)
            _, _, term, trunc, _ = dfn.step(d)
            if term or trunc:
                dfn.reset()
    else:
        env = ref_loader.make("CyberBattleToyCtf-v0",
                              defender_agent=ref_defender.ScanAndReimageCompromisedMachines(probability=0.6, scan_capacity=2, scan_frequency=5),
                              defender_constraint=ref_env.DefenderConstraint(maintain_sla=0.80), **bounds)
        env.reset()

        def one():
            _, _, done, _, _ = env.step(env.sample_valid_action(kinds=[0, 1, 2]))
            if done:
                env.reset()

    for _ in range(3):
        one()
    steps, t0 = 0, time.perf_counter()
    t_end = t0 + seconds
    while time.perf_counter() < t_end:
        one()
        steps += 1
    return steps, time.perf_counter() - t0


def measure(mode, cores, seconds):
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_worker, [(7000 + i, seconds, mode) for i in range(cores)])
    return sum(s / t for s, t in res), sum(s for s, _ in res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=12.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_python_reference_cpu.json"))
    args = ap.parse_args()
    cores = len(os.sched_getaffinity(0))
    pair, pair_steps = measure("pair", cores, args.seconds)
    scan, scan_steps = measure("scan", cores, args.seconds)
    import numpy
    import platform

    out = {
        "value": pair, "unit": "env-steps/s", "cores": cores, "per_core": pair / cores, "kind": "reference", "cpu_model": _cpu_model(),
        "sample": f"{cores} processes (multiprocessing.Pool, one per core), ToyCtf(12,10) AttackerEnvWrapper + DefenderEnvWrapper pair step, "
                  f"random valid attacker actions, {args.seconds:.0f} s wall each after 3 warm-up iterations ({pair_steps} env-steps)",
        "toyctf_scan_and_reimage": {"value": scan, "per_core": scan / cores,
                                    "sample": f"bare CyberBattleEnv.step + ScanAndReimage(0.6, 2, 5), SLA 0.80 ({scan_steps} env-steps)"},
        "note": f"unmodified reference from /root/reference, Python {platform.python_version()}, numpy {numpy.__version__} "
                "(reference pins 1.26.4), gymnasium / boolean.py stand-ins of oracle/refshim (SURVEY.md section D); "
                "script: oracle/time_python_reference.py",
    }
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
