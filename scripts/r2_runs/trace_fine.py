import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, bench
from marlon_b200.batch import Batch
w = os.environ.get("WORKLOAD", "chain100")
comp, cfg = bench.workload_config(workload=w)
n = int(os.environ.get("ENVS", 131072))
if isinstance(comp, list): n = [n // len(comp)] * len(comp)
b = Batch(comp, cfg, n); b.reset()
acts = []
for s in range(40):
    a, d = b.sample_actions(seed=1); acts.append((a.clone(), None if d is None else d.clone())); b.step(a, d)
a, d = b.sample_actions(seed=1)
b.phase_cycles(True)
b.step(a, d); torch.cuda.synchronize()
out = (C.c_uint64 * 8192)()
b._L.cbx_batch_debug_read.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
assert b._L.cbx_batch_debug_read(b._h, out, 8192) == 0
t = np.frombuffer(out, dtype=np.uint64)[:4 * 14 * 4 * 16].reshape(4, 14, 4, 16).astype(np.int64)
print(w, "per tile (us): logic->fields start | scalars+leaked | cachem gather | cachem emit | props gather | props emit+rest | ncmax ndmax")
for c in range(4):
    for wp in range(0, 14, 3):
        for u in range(2):
            s = t[c, wp, u]
            if s[0] == 0 or s[5] == 0: continue
            d = lambda a, b: (s[a] - s[b]) / 1e3
            print(f"cta {37*c:3d} warp {wp:2d} tile {u}: fields {d(5,4):5.1f} = {d(7,4):5.1f} + {d(8,7):5.1f} + {d(9,8):5.1f} + {d(10,9):5.1f} + {d(5,10):5.1f}   ncmax {s[12]} ndmax {s[13]}")
