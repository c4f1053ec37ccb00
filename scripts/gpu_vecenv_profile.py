"""cProfile of the SB3 VecEnv adapter at bench size (where does a step pair's host time go?)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from marlon_b200.universe import MultiAgentUniversalEnv
n = int(os.environ.get("ENVS", 65536))
u = MultiAgentUniversalEnv("CyberBattleToyCtf-v0", n, maximum_node_count=12, maximum_total_credentials=10,
                           maximum_discoverable_credentials_per_action=5, max_timesteps=2000, emit_terminal_obs=True)
mode = os.environ.get("OBS", "torch")
av, dv = u.vec_env("attacker", observations=mode, terminal_observations="truncated"), u.vec_env("defender", observations=mode, terminal_observations="truncated")
av.reset(); dv.reset()
acts = []
for s in range(30):
    a, d = u.sample_actions(seed=s)
    acts.append((a.cpu().numpy().astype(np.int64), d.cpu().numpy().astype(np.int64)))
    av.step(acts[-1][0]); dv.step(acts[-1][1])
av.reset(); dv.reset()
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for s in range(30):
    av.step(acts[s][0]); dv.step(acts[s][1])
pr.disable()
dt = time.perf_counter() - t0
print(f"{mode}: {dt / 30 * 1e3:.2f} ms per step pair, {n * 30 / dt:.4g} env-steps/s")
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
