"""Pure-write vs copy bandwidth on this GPU (is a write-only stream limited below the read+write copy figure?)."""
import torch

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

for mb in (256, 831, 2048, 8192):
    nbytes = mb * 1000 * 1000
    a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    t_fill = timeit(lambda: a.fill_(1))
    t_zero = timeit(lambda: a.zero_())
    t_copy = timeit(lambda: a.copy_(b))
    a32 = a.view(torch.int32)
    t_fill32 = timeit(lambda: a32.fill_(7))
    t_read = timeit(lambda: a32.sum())
    print(f"{mb:5d} MB: fill_u8 {nbytes/t_fill/1e6:8.1f} GB/s  zero {nbytes/t_zero/1e6:8.1f} GB/s  fill_i32 {nbytes/t_fill32/1e6:8.1f} GB/s  "
          f"copy(r+w) {2*nbytes/t_copy/1e6:8.1f} GB/s  read(sum) {nbytes/t_read/1e6:8.1f} GB/s")
