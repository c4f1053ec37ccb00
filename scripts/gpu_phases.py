"""Per-phase cycle breakdown of the step kernel on the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from marlon_b200.batch import Batch
comp, cfg = bench.workload_config(workload=os.environ.get("WORKLOAD", "toyctf"))
n = int(os.environ.get("ENVS", 65536))
if isinstance(comp, list):  # multi-scenario workload: envs per scenario
    n = [n // len(comp)] * len(comp)
b = Batch(comp, cfg, n); b.reset()
acts = []
for s in range(25):
    a, d = b.sample_actions(seed=1); acts.append((a.clone(), d.clone())); b.step(a, d)
b.close()
b = Batch(comp, cfg, n); b.reset()
for s in range(5): b.step(*acts[s])
b.phase_cycles(True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(5, 25): b.step(*acts[s])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
pc = b.phase_cycles(False)
tot = max(sum(pc.values()), 1)
print(f"kernel+launch {ms:.4f} ms/step; grid CTAs x tiles; phase share of CTA time:")
for k, v in pc.items():
    if v: print(f"  {k:22s} {v/tot:6.3f}  ({v/20/1e6:8.2f} Mcycles/step summed over CTAs)")
