"""Bring-up helper for the GPU box: runs each kernel family against the oracle in its own subprocess (a trapped
kernel poisons its CUDA context, so stages must not share a process) and prints one compact line per check.

    python tools/gpu_stage_check.py            # all stages
    python tools/gpu_stage_check.py conv 0     # one stage in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _torch_conv_ref(mode, x, w, b):
    import torch.nn.functional as F
    if mode == 0:
        return F.conv2d(x, w, b, padding=1)
    if mode == 1:
        return F.conv2d(x, w, b)
    if mode == 2:
        return F.conv2d(x, w, b, stride=2, padding=1)
    return F.conv_transpose2d(x, w, b, stride=2)


def run_conv_case(mode, impl, B, H, W, cin0, cin1, cout, act=0, use_res=False, use_scale=False, dual=False, seed=0):
    import ctypes as C

    import numpy as np
    import torch

    from yond_public_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    cin = cin0 + cin1
    x = torch.randn(B, cin, H, W, generator=g).to(torch.bfloat16)
    if mode == 3:
        w = (torch.randn(cin, cout, 2, 2, generator=g) * 0.1)
    else:
        k = 1 if mode == 1 else 3
        w = torch.randn(cout, cin, k, k, generator=g) * (0.5 / (cin * k * k) ** 0.5)
    wq = w.to(torch.bfloat16).float()
    b = torch.randn(cout, generator=g) * 0.1
    ref = _torch_conv_ref(mode, x.float(), wq, b)
    Ho, Wo = ref.shape[-2:]
    scale = shift = None
    if use_scale:
        scale = torch.randn(B, cout, generator=g) * 0.5 + 1.0
        shift = torch.randn(B, cout, generator=g) * 0.1
        ref = ref * scale[:, :, None, None] + shift[:, :, None, None]
    if act == 1:
        ref = torch.nn.functional.leaky_relu(ref, 0.2)
    elif act == 2:
        ref = torch.nn.functional.silu(ref)
    res = None
    if use_res:
        res = torch.randn(B, cout, Ho, Wo, generator=g).to(torch.bfloat16)
        ref = ref + res.float()
    dev = torch.device("cuda")
    xn = x.permute(0, 2, 3, 1).contiguous()
    s0 = xn[..., :cin0].contiguous().to(dev)
    s1 = xn[..., cin0:].contiguous().to(dev) if cin1 else None
    out0 = torch.zeros(B, Ho, Wo, cout, dtype=torch.bfloat16, device=dev)
    out1 = torch.zeros_like(out0) if dual else None
    resd = res.permute(0, 2, 3, 1).contiguous().to(dev) if use_res else None
    wc = np.ascontiguousarray(w.numpy())
    p = _lib.ptr
    bd = b.to(dev)  # keep device operands referenced until the (synchronous) call returns
    scd = scale.to(dev).contiguous() if use_scale else None
    shd = shift.to(dev).contiguous() if use_scale else None
    rc = lib.yond_conv2d(mode, impl, B, H, W, cin0, cin1, p(s0), p(s1), cout, wc.ctypes.data_as(C.c_void_p),
                         p(bd), p(scd), p(shd), act, 0.2, p(resd), p(out0), p(out1), _lib.stream_ptr())
    if rc:
        return f"rc={rc} {lib.yond_last_error().decode()}"
    got = out0.float().cpu().permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    tol = 2.0 ** -7 * max(1.0, ref.abs().max().item())  # bf16 output rounding is 2^-9 of the value; fp32 accumulation-order noise on top
    msg = f"max_err={err:.4g} (|ref|max={ref.abs().max().item():.3g})"
    if dual:
        e1 = (out1.float().cpu().permute(0, 3, 1, 2) - torch.nn.functional.silu(ref)).abs().max().item()
        msg += f" dual_err={e1:.4g}"
        err = max(err, e1)
    return ("OK " if err < tol else "BAD ") + msg


CONV_CASES = [
    # mode, B, H, W, cin0, cin1, cout, act, res, scale, dual
    (1, 2, 16, 16, 64, 0, 64, 0, False, False, False),     # 1x1, SW128, plain tiles
    (1, 2, 16, 16, 32, 0, 32, 0, False, False, False),     # 1x1, SW64
    (0, 1, 16, 16, 64, 0, 64, 0, False, False, False),     # 3x3 slab SW128
    (0, 1, 16, 16, 32, 0, 32, 1, False, False, False),     # 3x3 slab SW64
    (0, 3, 8, 8, 64, 0, 128, 2, True, True, True),         # NB=2 tiles (C,W,B,H map), full epilogue
    (0, 2, 24, 40, 32, 32, 32, 1, False, False, False),    # concat, ragged tiles
    (0, 1, 16, 16, 256, 256, 256, 1, False, False, False), # streamed weights, NT=256
    (0, 2, 8, 8, 512, 0, 512, 1, False, False, False),     # two n tiles
    (2, 2, 16, 16, 32, 0, 64, 0, False, False, True),      # stride 2
    (2, 1, 32, 48, 64, 0, 128, 0, False, False, True),
    (3, 2, 8, 8, 64, 0, 32, 0, False, False, False),       # convT
    (3, 1, 8, 8, 512, 0, 256, 0, False, False, False),     # convT, N=1024
    (1, 1, 16, 16, 64, 64, 64, 0, False, False, True),     # shortcut 1x1 on concat
    (0, 4, 128, 128, 32, 0, 32, 1, False, False, False),   # many tiles (persistent loop, phases)
    (0, 1, 4, 2, 512, 0, 512, 1, False, False, False),     # tiny maps (golden net fixture level 4)
    (0, 4, 16, 16, 128, 0, 128, 2, True, True, False),     # T=2 sub-tiles along the batch, streamed weights
    (0, 2, 64, 64, 64, 0, 64, 1, False, False, True),      # resident weights, T=2 along H
    (0, 1, 40, 24, 128, 0, 128, 0, True, False, False),    # ragged super-tiles along H
    (0, 3, 16, 16, 256, 0, 256, 1, False, False, False),   # odd batch with T=2 along the batch, two n tiles
    (0, 1, 96, 126, 32, 32, 32, 2, False, True, False),    # full-frame-like ragged width, concat, T=4
    (0, 1, 16, 24, 128, 0, 128, 1, False, False, False),   # CTA pair with an odd number of M tiles (the peer's last tile is empty)
    (0, 5, 8, 8, 512, 0, 512, 2, True, True, False),       # CTA pair, two n tiles, NB=2 image pairs, odd batch, full epilogue
    (0, 2, 24, 40, 32, 0, 32, 2, True, True, True),        # pixel-pair formulation (32 -> 32): FiLM + SiLU + residual + dual store
    (0, 3, 96, 126, 32, 0, 32, 2, False, True, False),     # pixel pairs, ragged width (63 pairs), T=4, scale only
    (0, 1, 16, 15, 32, 0, 32, 1, True, False, False),      # odd width: falls back to the N=32 path
]


def stage_conv(i):
    c = CONV_CASES[i]
    mode, B, H, W, c0, c1, co, act, res, sc, dual = c
    r_ref = run_conv_case(mode, 1, B, H, W, c0, c1, co, act, res, sc, dual)
    r_tc = run_conv_case(mode, 0, B, H, W, c0, c1, co, act, res, sc, dual)
    print(f"conv[{i}] {c}: ref {r_ref} | tc {r_tc}")


def stage_isp():
    import numpy as np
    import torch

    import yond_public_b200 as Y
    from oracle import yond_oracle as O
    rng = np.random.default_rng(0)
    bay = rng.standard_normal((12, 16)).astype(np.float32)
    print("pack exact:", np.array_equal(Y.bayer2rggb(bay), O.bayer2rggb(bay)),
          "unpack exact:", np.array_equal(Y.rggb2bayer(O.bayer2rggb(bay)), bay))
    bay = rng.standard_normal((3, 6, 10)).astype(np.float32)
    print("pack scalar path exact:", np.array_equal(Y.bayer2rggb(bay), np.stack([O.bayer2rggb(b) for b in bay])))
    x = rng.uniform(-30, 960, 5000).astype(np.float32)
    z = Y.VST(x, 5.1, gain=3.7)
    zr = O.VST(x, np.float64(5.1), gain=np.float64(3.7))
    print("VST max rel err:", float(np.max(np.abs(z - zr) / np.maximum(zr, 1e-3))))
    lut = Y.BiasLUT()
    olut = O.BiasLUT(lut.bias_lut)
    xs = rng.uniform(0, 960, 4000).astype(np.float32)
    for K, s in [(3.7, 5.1), (0.31, 1.9), (21.0, 30.5), (1.0, 0.0)]:
        d = np.abs(lut.get_lut(xs, K, s) - olut.get_lut(xs, K=np.float64(K), sigGs=np.float64(s))).max()
        print(f"BiasLUT K={K} s={s}: max abs err {d:.3g}")
    torch.cuda.synchronize()


def stage_net(key):
    import numpy as np
    import torch

    import yond_public_b200 as Y
    from oracle import yond_oracle as O
    archs = {"unet": {"name": "UNetSeeInDark", "in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": True, "norm": True},
             "gru": {"name": "GuidedResUnet", "guided": True, "in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": True, "norm": True},
             "snr": {"name": "SNRnet", "guided": True, "in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": True, "norm": True}}
    arch = archs[key]
    for scale in (None, 4.0):
        sd = O.init_state_dict(arch, seed=5, weight_scale=scale)
        net = Y.build_net(arch)
        net.load_state_dict(sd)
        g = torch.Generator().manual_seed(1)
        x = torch.rand(2, 4, 64, 96, generator=g)
        x[1] *= 0.6
        t = torch.tensor(0.043) if "guided" in arch else None
        with torch.no_grad():
            ref = O.net_forward(arch, sd, x, t)
            refb = O.net_forward(arch, sd, x, t, bf16=True)
        for impl in (1, 0):
            net.conv_impl = impl
            y = (net(x.cuda(), t.cuda()) if t is not None else net(x.cuda())).cpu()
            print(f"net {key} wscale={scale} impl={impl}: max|y-ref_fp32|={float((y - ref).abs().max()):.3g} "
                  f"max|y-ref_bf16emu|={float((y - refb).abs().max()):.3g} |ref|max={float(ref.abs().max()):.3g} "
                  f"|ref-x|max={float((ref - x).abs().max()):.3g}")


def stage_nlf():
    import numpy as np

    import yond_public_b200 as Y
    from oracle import yond_oracle as O
    r2 = np.random.default_rng(77)
    clean = O.synth_clean(r2, 256, 384)
    noisy = O.synth_noisy(r2, clean, 6.0, 9.0)
    rggb = O.bayer2rggb(noisy)
    print("blur19 max err:", float(np.abs(Y.blur(rggb, 19) - O.blur(rggb, 19)).max()),
          "std29 max err:", float(np.abs(Y.stdfilt(rggb, 29) - O.stdfilt(rggb, 29)).max()))
    reg = Y.SimpleNLF(noisy, k=29, setting={"mode": "self"})
    ref = O.SimpleNLF(noisy, k=29, setting={"mode": "self"})
    print("SimpleNLF self:", reg, ref, "rel err", np.abs(reg - ref) / np.abs(ref))


def main():
    if len(sys.argv) > 1:
        st = sys.argv[1]
        if st == "conv":
            stage_conv(int(sys.argv[2]))
        elif st == "isp":
            stage_isp()
        elif st == "net":
            stage_net(sys.argv[2])
        elif st == "nlf":
            stage_nlf()
        return
    stages = ([] if os.environ.get("STAGE_ONLY_CONV3") else [["isp"]]) + [["conv", str(i)] for i in (range(len(CONV_CASES)) if not os.environ.get("STAGE_ONLY_CONV3") else [i for i, c in enumerate(CONV_CASES) if c[0] == 0])] + ([] if os.environ.get("STAGE_ONLY_CONV3") else [["net", k] for k in ("unet", "gru", "snr")] + [["nlf"]])
    for st in stages:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + st, capture_output=True, text=True, timeout=300)
            out = (r.stdout + ("\nSTDERR: " + r.stderr[-1500:] if r.returncode else "")).strip()
        except subprocess.TimeoutExpired:
            out = "TIMEOUT"
        print(f"=== {' '.join(st)} ({time.time() - t0:.1f}s)\n{out}", flush=True)


if __name__ == "__main__":
    main()
