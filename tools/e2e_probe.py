"""Where the end-to-end (host buffers in/out) step of the 12 MP workload spends its time: device-resident step in groups of
4 and 8 frames, copies alone, and the streamed host path for several group sizes."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import yond_public_b200 as Y  # noqa: E402
from yond_public_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
sd = synth.bench_state_dict(bench.ARCH, seed=0)
N = bench.N_FRAMES
frames = bench.synth_frames(N, seed=2024)
host_in = torch.from_numpy(frames.reshape(N, 1, bench.FRAME_H, bench.FRAME_W)).pin_memory()
host_outs = [torch.empty((N, bench.FRAME_H, bench.FRAME_W), dtype=torch.float32).pin_memory() for _ in range(2)]
dev_in = host_in.to(dev)
drv = Y.YOND_SIDD(bench.ARCH, bench.PIPE_FRAME, state_dict=sd, device=dev)


def timed(fn, reps=4):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print(f"device-resident, 8 frames at once: {timed(lambda: drv.iter_denoise_batch(dev_in, dict(bench.P0))):.2f} ms")
print(f"device-resident, 2 x 4 frames:     {timed(lambda: [drv.iter_denoise_dev(dev_in[a:a + 4], dict(bench.P0)) for a in (0, 4)]):.2f} ms")
dbuf = torch.empty_like(dev_in)
dout = torch.empty((N, bench.FRAME_H, bench.FRAME_W), device=dev)
print(f"H2D 8 frames: {timed(lambda: dbuf.copy_(host_in, non_blocking=True)):.2f} ms;  D2H 8 frames: {timed(lambda: host_outs[0].copy_(dout, non_blocking=True)):.2f} ms")
s2 = torch.cuda.Stream()
def both():
    dbuf.copy_(host_in, non_blocking=True)
    with torch.cuda.stream(s2):
        host_outs[0].copy_(dout, non_blocking=True)
print(f"H2D + D2H concurrently: {timed(both):.2f} ms")
for grp in (2, 4, 8):
    jobs = []
    seq = [0]
    def step():
        seq[0] += 1
        jobs.append(drv.iter_denoise_host(host_in, host_outs[seq[0] % 2], dict(bench.P0), group=grp, wait=False))
        while len(jobs) > 1:
            jobs.pop(0).result()
    t = timed(step, 5)
    while jobs:
        jobs.pop(0).result()
    print(f"host path streamed, groups of {grp}: {t:.2f} ms/step")
t = timed(lambda: drv.iter_denoise_host(host_in, host_outs[0], dict(bench.P0), group=4, wait=True), 4)
print(f"host path, one batch at a time (wait=True), groups of 4: {t:.2f} ms/step")
# host time to SUBMIT one streamed step (everything is asynchronous: this is pure host work)
jobs = []
torch.cuda.synchronize()
ts = []
for i in range(6):
    t0 = time.perf_counter()
    jobs.append(drv.iter_denoise_host(host_in, host_outs[i % 2], dict(bench.P0), group=8, wait=False))
    ts.append((time.perf_counter() - t0) * 1e3)
    while len(jobs) > 1:
        jobs.pop(0).result()
while jobs:
    jobs.pop(0).result()
print("host submit ms per step:", [round(t, 1) for t in ts])
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
j = drv.iter_denoise_host(host_in, host_outs[0], dict(bench.P0), group=8, wait=False)
pr.disable()
j.result()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
