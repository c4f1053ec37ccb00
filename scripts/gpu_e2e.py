"""Where the end-to-end (host-buffer) step spends its time: kernel under zero-copy, copies, host overhead."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from marlon_b200.batch import Batch
comp, cfg = bench.workload_config()
n = 65536
b = Batch(comp, cfg, n); b.reset()
K = 60
ta = torch.empty((K, n, 10), dtype=torch.int32, device="cuda"); td = torch.empty((K, n, 12), dtype=torch.int32, device="cuda")
for s in range(K):
    b.sample_actions(seed=5, attacker_out=ta[s], defender_out=td[s]); b.step(ta[s], td[s])
ha = torch.empty((K, n, 10), dtype=torch.int32, pin_memory=True); hd = torch.empty((K, n, 12), dtype=torch.int32, pin_memory=True)
ha.copy_(ta); hd.copy_(td); torch.cuda.synchronize()
han, hdn = ha.numpy(), hd.numpy()
b.close()
for dt in (torch.int32, torch.int16):
    ha2, hd2 = torch.empty((K, n, 10), dtype=dt, pin_memory=True), torch.empty((K, n, 12), dtype=dt, pin_memory=True)
    ha2.copy_(ha.to(dt)); hd2.copy_(hd.to(dt))
    a_n, d_n = ha2.numpy(), hd2.numpy()
    b = Batch(comp, cfg, n); b.reset()
    for s in range(10): b.step_host(a_n[s], d_n[s])
    t0 = time.perf_counter()
    for s in range(10, K): b.step_host(a_n[s], d_n[s])
    wall0 = (time.perf_counter() - t0) / (K - 10)
    b.close()
    b = Batch(comp, cfg, n); b.reset()
    for s in range(10): b.step_host(a_n[s], d_n[s])
    b.enable_timing(True)
    t0 = time.perf_counter()
    for s in range(10, K): b.step_host(a_n[s], d_n[s])
    wall = (time.perf_counter() - t0) / (K - 10)
    kms, kn = b.step_kernel_ms()
    print(f"step_host {dt}: wall {wall0*1e3:.4f} ms/step untimed, {wall*1e3:.4f} with events, kernel {kms:.4f} ms ({kn} launches), rest {wall*1e3-kms:.4f} ms")
    b.close()
# D2H alone
out = torch.empty(n * 12, dtype=torch.uint8, pin_memory=True)
src = torch.empty(n * 12, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): out.copy_(src, non_blocking=True); torch.cuda.synchronize()
print(f"D2H 0.79 MB + sync: {(time.perf_counter()-t0)/50*1e3:.4f} ms")
