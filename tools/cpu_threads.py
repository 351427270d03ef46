"""Host-core sanity check for the CPU baseline: one 3x3 conv of the up_3 shape (77 GFLOP) at several thread counts."""
import os, time, torch
import torch.nn.functional as F
print("os.cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "torch threads", torch.get_num_threads())
x = torch.randn(1, 128, 512, 512); w = torch.randn(128, 128, 3, 3)
for t in (8, 16, 32, 64, 128):
    torch.set_num_threads(t)
    F.conv2d(x, w, padding=1)
    t0 = time.perf_counter()
    for _ in range(3):
        F.conv2d(x, w, padding=1)
    dt = (time.perf_counter() - t0) / 3
    print("threads %3d: %.3f s  %.1f GFLOP/s" % (t, dt, 77.3 / dt), flush=True)
