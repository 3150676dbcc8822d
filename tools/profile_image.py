"""Profiling driver: warms up, then runs ONE SIDD-shaped image (32 blocks of 256x256) through iter_denoise_device
between cudaProfilerStart/Stop (use with `ncu --profile-from-start off`).  `--batch N` profiles N blocks through the
network only (conv-stack capture)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import yond_public_b200 as Y  # noqa: E402
from yond_public_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--net-only", type=int, default=0, help="profile only the network forward on this many 128x128 packed blocks")
ap.add_argument("--arch", default="gru")
ap.add_argument("--step", type=int, default=0, help="profile one batched pipeline step over this many images")
ap.add_argument("--frames", type=int, default=0, help="profile one batched pipeline step over this many 12 MP frames (bench.py's headline workload)")
ap.add_argument("--time", type=int, default=0, help="with --net-only: time this many forwards with CUDA events instead of profiling")
ap.add_argument("--frame", default=None, help="HxW packed frame for --net-only, e.g. 1536x2016")
args = ap.parse_args()
arch = bench.ARCH if args.arch == "gru" else {"name": "UNetSeeInDark", "in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": True, "norm": True}
sd = synth.random_init_state_dict(arch, seed=0)
drv = Y.YOND_SIDD(arch, bench.PIPE, state_dict=sd)
if args.net_only:
    B = args.net_only
    H, W = (128, 128) if not args.frame else tuple(int(v) for v in args.frame.split("x"))
    z = torch.rand((B, H, W, 4), device="cuda")
    ub = z.amax(dim=(1, 2, 3)).contiguous()
    t = torch.full((B,), 0.04, device="cuda")
    for _ in range(3):
        drv.net.forward_nhwc(z, ub, t if "guided" in arch else None)
    torch.cuda.synchronize()
    if args.time:
        drv.net.enable_profile(True) if hasattr(drv.net, "enable_profile") else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.time):
            drv.net.forward_nhwc(z, ub, t if "guided" in arch else None)
        e1.record()
        torch.cuda.synchronize()
        print(f"net forward {B} x {H}x{W}: {e0.elapsed_time(e1) / args.time:.3f} ms")
        sys.exit(0)
    torch.cuda.cudart().cudaProfilerStart()
    drv.net.forward_nhwc(z, ub, t if "guided" in arch else None)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
elif args.frames:
    sd = synth.bench_state_dict(arch, seed=0)
    drv = Y.YOND_SIDD(arch, bench.PIPE_FRAME, state_dict=sd)
    fr = bench.synth_frames(args.frames, seed=2024)
    dev_in = torch.from_numpy(fr.reshape(args.frames, 1, bench.FRAME_H, bench.FRAME_W)).cuda()
    for i in range(2):
        drv.iter_denoise_batch(dev_in, dict(bench.P0))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    drv.iter_denoise_batch(dev_in, dict(bench.P0))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
elif args.step:
    imgs, _ = bench.synth_images(args.step)
    dev_in = torch.from_numpy(imgs).cuda()
    for i in range(2):
        drv.iter_denoise_batch(dev_in, dict(bench.P0))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    drv.iter_denoise_batch(dev_in, dict(bench.P0))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    import time
    tm = {}
    for _ in range(3):
        tm.pop("_t", None)
        drv.iter_denoise_batch(dev_in, dict(bench.P0), timings=tm)
    print("stage ms/step:", {k: round(v / 3 * 1e3, 2) for k, v in tm.items() if k != "_t"})
    for chunk in (32, 64, 128, 256):
        drv.engine.chunk = chunk
        drv.iter_denoise_batch(dev_in, dict(bench.P0))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            drv.iter_denoise_batch(dev_in, dict(bench.P0))
        torch.cuda.synchronize()
        print(f"chunk {chunk}: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms/step")
else:
    imgs, _ = bench.synth_images(2)
    dev_in = torch.from_numpy(imgs).cuda()
    for i in range(3):
        drv.iter_denoise_device(dev_in[i % 2], dict(bench.P0))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    drv.iter_denoise_device(dev_in[0], dict(bench.P0))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("profiled ok")
