"""Host enqueue time vs device time of one pipeline step (is the launch path host-bound?), and e2e group-size sweep."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import yond_public_b200 as Y  # noqa: E402
from yond_public_b200 import synth  # noqa: E402

sd = synth.bench_state_dict(bench.ARCH, seed=0)
drv = Y.YOND_SIDD(bench.ARCH, bench.PIPE_FRAME, state_dict=sd)
frames = bench.synth_frames(8, seed=1)
host_in = torch.from_numpy(frames.reshape(8, 1, bench.FRAME_H, bench.FRAME_W)).pin_memory()
host_out = torch.empty((8, bench.FRAME_H, bench.FRAME_W)).pin_memory()
dev_in = host_in.cuda()
for n in (1, 2, 4, 8):
    x = dev_in[:n]
    for _ in range(2):
        drv.iter_denoise_dev(x, dict(bench.P0))
    torch.cuda.synchronize()
    l0 = Y._lib.launch_count()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = drv.iter_denoise_dev(x, dict(bench.P0))
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{n} frames: host enqueue {1e3 * (t1 - t0):.2f} ms, device {e0.elapsed_time(e1):.2f} ms, launches {Y._lib.launch_count() - l0}", flush=True)
for group in (1, 2, 4, [1, 2, 2, 2, 1], [1, 3, 3, 1], 8):
    for _ in range(2):
        drv.iter_denoise_host(host_in, host_out, dict(bench.P0), group=group)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        drv.iter_denoise_host(host_in, host_out, dict(bench.P0), group=group)
    torch.cuda.synchronize()
    print(f"e2e group {group}: {1e3 * (time.perf_counter() - t0) / 5:.2f} ms per 8 frames", flush=True)
# raw copy rates
torch.cuda.synchronize()
for name, fn in (("H2D", lambda: dev_in.copy_(host_in, non_blocking=True)), ("D2H", lambda: host_out.copy_(dev_in[:, 0], non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {host_in.numel() * 4 / dt / 1e9:.1f} GB/s ({1e3 * dt:.2f} ms per 390 MB)")
