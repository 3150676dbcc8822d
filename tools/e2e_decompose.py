"""Decomposes the end-to-end overhead of the streamed host path: pure asynchronous compute, + D2H only, + H2D only, both."""
import os, sys, time
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
host_out = torch.empty((N, bench.FRAME_H, bench.FRAME_W), dtype=torch.float32).pin_memory()
dev_in = host_in.to(dev)
stage = [torch.empty_like(dev_in) for _ in range(2)]
drv = Y.YOND_SIDD(bench.ARCH, bench.PIPE_FRAME, state_dict=sd, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
cur = torch.cuda.current_stream()


def run(h2d, d2h, steps=6):
    def one(k):
        src = dev_in
        if h2d:
            with torch.cuda.stream(s_in):
                stage[k % 2].copy_(host_in, non_blocking=True)
                ev = torch.cuda.Event(); ev.record(s_in)
            cur.wait_event(ev)
            src = stage[k % 2]
        res = drv.iter_denoise_dev(src, dict(bench.P0))
        if d2h:
            done = torch.cuda.Event(); done.record(cur)
            s_out.wait_event(done)
            with torch.cuda.stream(s_out):
                host_out.copy_(res["final"], non_blocking=True)
            res["final"].record_stream(s_out)
    for k in range(2):
        one(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        one(k)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


for h2d, d2h in ((0, 0), (0, 1), (1, 0), (1, 1)):
    print(f"h2d={h2d} d2h={d2h}: {run(h2d, d2h):.2f} ms/step")
