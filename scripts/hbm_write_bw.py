"""Pure-write, pure-read and copy bandwidth of this GPU's HBM (torch kernels, CUDA events): the yardstick for the
write-dominated dense build.  usage: python scripts/hbm_write_bw.py [GB]"""
import sys

import torch

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 1.5
n = int(gb * 1e9 / 4)
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_fill = timed(lambda: x.fill_(-2.3))
t_zero = timed(lambda: x.zero_())
t_sum = timed(lambda: x.sum())
t_copy = timed(lambda: y.copy_(x))
print(f"buffer {gb:.2f} GB: fill {gb / t_fill * 1e3:.0f} GB/s ({t_fill:.3f} ms), memset {gb / t_zero * 1e3:.0f} GB/s, "
      f"read(sum) {gb / t_sum * 1e3:.0f} GB/s, copy {2 * gb / t_copy * 1e3:.0f} GB/s (read+write)")
