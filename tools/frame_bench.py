"""Times BASELINE.json configs[2]-style work on one GPU: a full-resolution Bayer frame (default 4032x3024) through
estimate + VST + network + inverse (`YOND_SIDD.iter_denoise_device`, device-resident), whole-frame and halo-tiled.
    python tools/frame_bench.py [H W] [reps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import yond_public_b200 as Y  # noqa: E402
from yond_public_b200 import synth  # noqa: E402

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3024, 4032)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rng = np.random.default_rng(3)
noisy = synth.noisy(rng, synth.clean_smooth(rng, H, W), 3.0, 5.0)
pipe = dict(bench.PIPE, full_dn=True, iter="once")
for name, arch in (("GuidedResUnet", bench.ARCH), ("UNetSeeInDark", {"name": "UNetSeeInDark", "in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": True, "norm": True})):
    drv = Y.YOND_SIDD(arch, pipe, state_dict=synth.random_init_state_dict(arch, seed=0))
    x = torch.from_numpy(noisy).cuda()
    p = {"wp": 1023, "bl": 64, "ratio": 1, "gain": 1, "sigma": 0, "scale": 959.0}

    def whole():
        return drv.iter_denoise_device(x, dict(p))

    def tiled():
        reg = np.asarray(drv.iter_denoise_device(x, dict(p))["regs"][0]) if False else REG
        return drv.engine.vst_denoise_tiled(x, REG[0] * 959, np.sqrt(max(REG[1], 0)) * 959, 959.0, core=512)

    REG = np.asarray(whole()["regs"][0])
    for label, fn in (("whole frame (estimate + VST + net + inverse)", whole), ("halo-tiled denoise only (core 512, halo 128)", tiled)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / reps * 1e3
        print(f"{name:14s} {H}x{W} {label}: {ms:7.2f} ms/frame = {H * W / 1e6 / (ms / 1e3):7.1f} MP/s")
