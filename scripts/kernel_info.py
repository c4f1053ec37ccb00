"""Print which kernel and tile order a batch of the headline workload uses at a few sizes (helper for gpurun checks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from marlon_b200.batch import Batch
comp, cfg = bench.workload_config()
for n in (65536, 262144):
    b = Batch(comp, cfg, n); b.reset()
    k = b.kernel_info()
    print(n, k["name"], k["tile_order"], k["ctas"], k["threads"])
    b.close()
