"""Per-warp timeline of one launch of the warp-per-tile kernel (experiment build libcbx_trace.so: globaltimer stamps at the phase
boundaries of every tile, warps of CTAs 0, 37, 74, 111).  CBX_LIB=marlon_b200/libcbx_trace.so python scripts/gpu_wide_trace.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from marlon_b200.batch import Batch
w = os.environ.get("WORKLOAD", "chain100")
comp, cfg = bench.workload_config(workload=w)
n = int(os.environ.get("ENVS", 131072))
if isinstance(comp, list): n = [n // len(comp)] * len(comp)
b = Batch(comp, cfg, n); b.reset()
acts = []
for s in range(12):
    a, d = b.sample_actions(seed=1); acts.append((a.clone(), None if d is None else d.clone())); b.step(a, d)
b.phase_cycles(True)
b.step(*acts[5]); torch.cuda.synchronize()
out = (C.c_uint64 * 8192)()
b._L.cbx_batch_debug_read.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
assert b._L.cbx_batch_debug_read(b._h, out, 8192) == 0
t = np.frombuffer(out, dtype=np.uint64)[:4 * 14 * 4 * 8].reshape(4, 14, 4, 8).astype(np.int64)
nw = b.kernel_info()["threads"] // 32
t0 = t[t > 0].min()
names = ["start", "acts", "attack", "term", "defend", "fields", "defobs"]
print(f"{w}: {nw} warps per CTA; times in us from the first stamp; per tile: start | +actions | +attacker | +defender,desc | +fields | +defender obs")
for c in range(4):
    for wp in range(nw):
        row = []
        for u in range(4):
            s = t[c, wp, u]
            if s[0] == 0: continue
            rel = lambda k: (s[k] - t0) / 1e3 if s[k] else float("nan")
            row.append("[%6.1f | %5.1f %5.1f %5.1f %5.1f %5.1f]" % (rel(0), (s[1]-s[0])/1e3, (s[2]-s[1])/1e3, (s[4]-s[2])/1e3 if s[4] else float('nan'), (s[5]-s[4])/1e3 if s[5] and s[4] else float('nan'), (s[6]-s[5])/1e3 if s[6] and s[5] else float('nan')))
        print(f"cta {37*c:3d} warp {wp:2d}: " + "  ".join(row))
ends = t[:, :, :, 6]
print("last stamp at %.1f us" % ((ends.max() - t0) / 1e3))
