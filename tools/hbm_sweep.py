"""BASELINE configs[4], second half: isolated HBM-kernel sweep — pack, unpack, uint16 ingest, the estimator (maps + fit), the fused
VST front end (pack + LUT bias + VST + normalise + pad) and the fused back end (de-normalise + inverse VST + unpack) — over
1 ... 256 MP of Bayer pixels, device-resident, CUDA events, algorithmic bytes per Bayer pixel of SURVEY 8(d).
Prints one table; `sweep(sizes)` returns the rows (bench.py embeds a short version as `kernel_sweep`)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yond_public_b200 as Y  # noqa: E402
from yond_public_b200 import _lib, isp, nlf  # noqa: E402
from yond_public_b200._lib import check, ptr, stream_ptr  # noqa: E402
from yond_public_b200.pipeline import VstParams, YondEngine  # noqa: E402

SHAPES = {1: (1, 1024, 1024), 4: (1, 2048, 2048), 12: (1, 3024, 4032), 49: (4, 3024, 4032), 98: (8, 3024, 4032), 256: (21, 3024, 4032)}
BYTES = {"pack": 8, "unpack": 8, "ingest_u16": 6, "estimate_self (maps + fit)": 16 + 4 * 3 + 8 + 12, "vst_fwd (fused front)": 8, "vst_inv (fused back)": 8,
         "render_srgb (float64 gamma)": 7, "block_metrics (PSNR + SSIM, float64)": 8}


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def sweep(sizes=(1, 4, 12, 49, 98, 256), peak_gbs=6543.1):
    lib = _lib.load()
    eng = YondEngine(None, {"name": "UNetSeeInDark"}, biaslut=Y.BiasLUT())
    est = nlf._estimator()
    rows = []
    for mp in sizes:
        B, H, W = SHAPES[mp]
        npx = B * H * W
        g = torch.Generator(device="cuda").manual_seed(mp)
        x = torch.rand((B, H, W), device="cuda", generator=g) * 0.6
        reps = 3 if mp >= 49 else 10
        h, w = H // 2, W // 2
        res = {}
        packed = torch.empty((B, h, w, 4), device="cuda")
        res["pack"] = timeit(lambda: check(lib.yond_pack(ptr(x), ptr(packed), B, H, W, stream_ptr())), reps)
        back = torch.empty_like(x)
        res["unpack"] = timeit(lambda: check(lib.yond_unpack(ptr(packed), ptr(back), B, h, w, stream_ptr())), reps)
        raw = (x * 959 + 64).to(torch.int16)
        res["ingest_u16"] = timeit(lambda: check(lib.yond_ingest_mosaic(ptr(raw), ptr(back), npx, 64.0, 1023.0, 1.0, 0, stream_ptr())), reps)
        x4 = x.reshape(B, 1, H, W)
        res["estimate_self (maps + fit)"] = timeit(lambda: est.estimate_dev(x4, None, 29), reps)
        regs = torch.tensor([[4e-3, 1e-5]] * B, device="cuda", dtype=torch.float64)
        ch = eng.chain_params(regs, None, B, 1, 959, 959.0, 959, 1, "pre", "exact")
        pl, pr, pt, pb = isp.get_p2d((B, 4, h, w), base=32)
        hp, wp = h + pt + pb, w + pl + pr
        z = torch.empty((B, hp, wp, 4), device="cuda")
        ub = torch.empty((B,), device="cuda")
        res["vst_fwd (fused front)"] = timeit(lambda: check(lib.yond_vst_fwd(ptr(x), ptr(z), ptr(ub), B, H, W, pl, pr, pt, pb, ptr(ch["params"]),
                                                                              ptr(ch["rows"]), ptr(ch["xnodes"]), ch["stride"], stream_ptr())), reps)
        out = torch.empty((B, H, W), device="cuda")
        res["vst_inv (fused back)"] = timeit(lambda: check(lib.yond_vst_inv_place(ptr(z), ptr(out), B, H, W, pl, pr, pt, pb, ptr(ch["params"]), 1, 1, 0,
                                                                                 None, 1, None, stream_ptr())), reps)
        if mp <= 98:  # evaluation-side kernels: float64 arithmetic bounds them, the HBM fraction is reported for scale
            bgr = torch.empty((B, H, W, 3), device="cuda", dtype=torch.uint8)
            gains = (C.c_double * 3)(0.5, 1.0, 0.6)
            ccm = (C.c_double * 9)(1.6, -0.5, -0.1, -0.2, 1.5, -0.3, 0.0, -0.6, 1.6)
            res["render_srgb (float64 gamma)"] = timeit(lambda: check(lib.yond_render_srgb(ptr(x), ptr(bgr), B, H, W, 0, 0, gains, ccm, stream_ptr())), reps)
            res["block_metrics (PSNR + SSIM, float64)"] = timeit(lambda: Y.block_metrics(x, out, 1), reps)
            del bgr
        for k, ms in res.items():
            gbs = BYTES[k] * npx / ms / 1e6
            rows.append({"kernel": k, "MP": round(npx / 1e6, 1), "ms": round(ms, 4), "algorithmic_B_per_px": BYTES[k], "GBps": round(gbs, 1), "frac": round(gbs / peak_gbs, 3)})
        del x, packed, back, raw, z, out
        torch.cuda.empty_cache()
    return rows


if __name__ == "__main__":
    rows = sweep()
    names = list(BYTES)
    sizes = sorted({r["MP"] for r in rows})
    print(f"{'kernel (B/px algorithmic)':40s}" + "".join(f"{s:>10.0f} MP" for s in sizes) + "   [fraction of 6543 GB/s]")
    for k in names:
        print(f"{k + ' (' + str(BYTES[k]) + ')':40s}" + "".join(f"{next((r['frac'] for r in rows if r['kernel'] == k and r['MP'] == s), float('nan')):>13.3f}" for s in sizes))
    print(f"{'-- time (ms)':40s}")
    for k in names:
        print(f"{k:40s}" + "".join(f"{next((r['ms'] for r in rows if r['kernel'] == k and r['MP'] == s), float('nan')):>13.3f}" for s in sizes))
