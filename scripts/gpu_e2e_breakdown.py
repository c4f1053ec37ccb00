"""Where an end-to-end (host-buffer) step's time goes: the kernel's own duration when every launch is synchronised, the
host-side call, and the same loop without Python (python -X importtime is not the question: the ctypes call is)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from marlon_b200.batch import Batch

class A: pass
args = A(); args.seed = 2026; args.factored = False
dev = torch.device("cuda", 0)
m = bench.measure_device(args, "toyctf", 65536, 20, 5, 1, 0, 0, dev)
for dtype in (torch.int16,):
    tape = bench.HostTape(m, dtype)
    b = Batch(m["comp"], m["cfg"], m["counts"], device=0); b.reset(); b.host_prepare()
    for s in range(10): b.step_host(*tape.get(s))
    K = 54
    b.enable_timing(True)
    t0 = time.perf_counter()
    for s in range(10, 10 + K): b.step_host(*tape.get(s))
    t1 = time.perf_counter()
    b.enable_timing(False)
    kms, kn = b.step_kernel_ms()
    print(f"{dtype}: wall {1e3*(t1-t0)/K:.4f} ms/step   kernel (events, per launch incl. gaps when bracketed) {kms:.4f} ms over {kn} launches   info {b.kernel_info()['overlapped_launches']}")
    # python overhead alone: the same calls with tape.get only
    t0 = time.perf_counter()
    for s in range(10, 10 + K): tape.get(s)
    t1 = time.perf_counter()
    print(f"   tape.get alone {1e6*(t1-t0)/K:.2f} us/step")
    # nosync: launches back to back, one sync at the end (what pipelining the host loop would buy)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(10, 10 + K): b.step_host(*tape.get(s), sync=False)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"   sync=False wall {1e3*(t1-t0)/K:.4f} ms/step")
    b.close()
