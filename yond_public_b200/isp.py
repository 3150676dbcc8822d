"""Function-level surface of the YOND hot path, same names / argument meaning as the reference
(utils/isp_ops.py, utils/isp_algos.py, utils/utils.py, YOND_SIDD.py) — backed by libyond_b200 kernels.

Inputs may be NumPy arrays (host, like the reference) or CUDA torch tensors; the result comes back in the
same kind.  Host inputs are staged through pinned memory.  There is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
DEFAULT_LUT = os.path.join(_DATA, "bias_lut_2d_f32.npz")


def _dev():
    if not torch.cuda.is_available():
        raise _lib.YondError("yond_public_b200 needs a CUDA device (B200); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def to_dev(a, dtype=torch.float32):
    """numpy / tensor -> contiguous CUDA tensor; returns (tensor, was_numpy)."""
    dev = _dev()  # raises without CUDA: there is no CPU path
    if isinstance(a, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32 if dtype == torch.float32 else None))
        return t.pin_memory().to(dev, non_blocking=True), True
    if not torch.is_tensor(a):
        a = torch.as_tensor(np.asarray(a, dtype=np.float32))
        return a.to(dev), True
    return a.to(device=dev, dtype=dtype).contiguous(), False


def _back(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t


# ------------------------------------------------------------------ A1 / A2  utils/isp_ops.py:57-71
def bayer2rggb(bayer):
    """(H,W) -> (H/2,W/2,4); also (B,H,W) -> (B,H/2,W/2,4) (the reference's batched `bayer2rggbs`)."""
    x, np_in = to_dev(bayer)
    batched = x.dim() == 3
    xb = x if batched else x[None]
    B, H, W = xb.shape
    out = torch.empty((B, H // 2, W // 2, 4), device=x.device, dtype=torch.float32)
    check(_lib.load().yond_pack(ptr(xb), ptr(out), B, H, W, stream_ptr()))
    return _back(out if batched else out[0], np_in)


def rggb2bayer(rggb):
    x, np_in = to_dev(rggb)
    batched = x.dim() == 4
    xb = x if batched else x[None]
    B, h, w, c = xb.shape
    assert c == 4
    out = torch.empty((B, 2 * h, 2 * w), device=x.device, dtype=torch.float32)
    check(_lib.load().yond_unpack(ptr(xb), ptr(out), B, h, w, stream_ptr()))
    return _back(out if batched else out[0], np_in)


bayer2rggbs = bayer2rggb


_ROT_K = {((1, 2), (2, 3)): 0, ((2, 1), (3, 2)): 3, ((2, 3), (1, 2)): 1, ((3, 2), (2, 1)): 2}


def rot_bayer(image, bayer_pattern, rev=False):
    """utils/sidd_utils.py:198-213: quarter-turn rotation that brings the CFA to its canonical phase (np.rot90 over the last
    two axes; `rev=True` undoes it).  (H,W) or (B,H,W) float32, NumPy or CUDA tensor."""
    key = tuple(tuple(int(v) for v in row) for row in bayer_pattern)
    if key not in _ROT_K:
        raise ValueError(f"unknown Bayer pattern {bayer_pattern}")
    k = _ROT_K[key]
    if rev:
        k = (4 - k) % 4
    x, np_in = to_dev(image)
    batched = x.dim() == 3
    xb = x if batched else x[None]
    B, H, W = xb.shape
    out = torch.empty((B, W, H) if k % 2 else (B, H, W), device=x.device, dtype=torch.float32)
    check(_lib.load().yond_rot90(ptr(xb), ptr(out), B, H, W, k, stream_ptr()))
    return _back(out if batched else out[0], np_in)


# ------------------------------------------------------------------ 8(f)-1  data_process/process.py:40-64
def pack_raw_bayer(raw, wp=1023, clip=True, raw_pattern=None, black_level_per_channel=None, interleaved=False):
    """RAW ingest.  `raw`: a rawpy-like object (`.raw_image_visible`, `.raw_pattern`, `.black_level_per_channel`) exactly
    as the reference takes it, or a uint16 mosaic (H,W) / batch (B,H,W) (NumPy or CUDA tensor) with `raw_pattern` and
    `black_level_per_channel` given.  Returns (4,H/2,W/2) float32 planes in R, G1, B, G2 order like the reference
    ((B,4,h,w) for a batch); `interleaved=True` returns (h,w,4), the layout the rest of the path consumes."""
    if raw_pattern is None:
        img, raw_pattern, black_level_per_channel = raw.raw_image_visible, raw.raw_pattern, raw.black_level_per_channel
    else:
        img = raw
    dev = _dev()  # raises without CUDA: there is no CPU path
    np_in = not torch.is_tensor(img)
    if np_in:
        a = np.ascontiguousarray(np.asarray(img))
        assert a.dtype == np.uint16, "the sensor mosaic must be uint16"
        x = torch.from_numpy(a.view(np.int16)).to(dev)  # same bits; torch's uint16 support is partial
    else:
        assert img.dtype in (torch.uint16, torch.int16), "the sensor mosaic must be a 16-bit integer tensor"
        x = img.to(dev).contiguous()
    batched = x.dim() == 3
    xb = x if batched else x[None]
    B, H, W = xb.shape
    pat = np.asarray(raw_pattern)
    pos = (C.c_int * 4)(*[int(2 * np.where(pat == c)[0][0] + np.where(pat == c)[1][0]) for c in range(4)])
    black = (C.c_float * 4)(*[float(np.float32(b)) for b in black_level_per_channel])
    h, w = H // 2, W // 2
    out = torch.empty((B, h, w, 4) if interleaved else (B, 4, h, w), device=dev, dtype=torch.float32)
    check(_lib.load().yond_pack_raw(ptr(xb), ptr(out), B, H, W, pos, black, float(np.float32(wp)), int(bool(clip)), int(bool(interleaved)),
                                    stream_ptr()))
    return _back(out if batched else out[0], np_in)
rggb2bayers = rggb2bayer


# ------------------------------------------------------------------ A3 / A4  utils/isp_algos.py:5-33
def VST(x, sigma, mu=0, gain=1.0):
    if np.isscalar(x) or (isinstance(x, np.ndarray) and x.ndim == 0):
        # scalars (lower = VST(0), upper = VST(scale), YOND_SIDD.py:264-265) are host arithmetic in float64
        fz = max(gain * float(x) + (3 / 8) * gain ** 2 + sigma ** 2 - gain * mu, 0.0)
        return np.float64(2 / gain * fz ** 0.5)
    assert mu == 0, "the YOND path only uses mu = 0"
    t, np_in = to_dev(x)
    out = torch.empty_like(t)
    check(_lib.load().yond_vst(ptr(t), ptr(out), t.numel(), float(sigma), float(gain), stream_ptr()))
    return _back(out, np_in)


def inverse_VST(z, sigma, gain=1, exact=False):
    t, np_in = to_dev(z)
    out = torch.empty_like(t)
    check(_lib.load().yond_inverse_vst(ptr(t), ptr(out), t.numel(), float(sigma), float(gain), int(bool(exact)), stream_ptr()))
    return _back(out, np_in)


# ------------------------------------------------------------------ A5  utils/isp_algos.py:162-231
def lut_grids():
    """x-grid (electrons): 128 linear nodes on [0,2^-4) + 1793 log-spaced nodes 2^-4..2^10; sigma-grid: 200 nodes
    [0,1) + 901 nodes [1,10]  (isp_algos.py:168-177)."""
    sp = 128
    x_lut = np.concatenate((np.linspace(0, 2 ** -4, sp, endpoint=False),
                            np.exp(np.linspace(np.log(2 ** (-4)), np.log(2 ** 10), 14 * sp + 1))))
    sg_lut = np.concatenate((np.linspace(0, 1, 200, endpoint=False), np.linspace(1, 10, 901)))
    return x_lut, sg_lut


def sigma_pos(sg_lut, sg):
    """Fractional sigma index — BiasLUT.pos_interp (isp_algos.py:179-186), host float64."""
    data = np.concatenate(([-np.inf], sg_lut))
    idx = int(np.clip(np.searchsorted(data, sg), 0, len(data) - 1))
    w = data[idx] - sg
    diff = data[idx] - data[idx - 1]
    return idx - w / diff - 1


class BiasLUT:
    """Bilinear lookup of the VST bias table, on device.  `lut_path`: .npy (the authors' file layout,
    (1921,1101) [x,sigma]) or the .npz stand-in shipped in yond_public_b200/data (key 'bias_lut')."""

    def __init__(self, lut_path="checkpoints/bias_lut_2d.npy"):
        if not os.path.exists(lut_path) and lut_path == "checkpoints/bias_lut_2d.npy":
            lut_path = DEFAULT_LUT
        arr = np.load(lut_path)
        table = arr["bias_lut"] if hasattr(arr, "files") else arr
        self.x_lut, self.sg_lut = lut_grids()
        assert table.shape == (len(self.x_lut), len(self.sg_lut)), f"bias LUT must be (1921,1101), got {table.shape}"
        self.bias_lut = np.ascontiguousarray(table, dtype=np.float32)
        self._dev_table = None
        self._dev_nodes = None

    def device_table(self):
        if self._dev_table is None:
            self._dev_table = torch.from_numpy(self.bias_lut).to(_dev())
            self._dev_nodes = torch.from_numpy(self.x_lut.astype(np.float32)).to(_dev())
        return self._dev_table, self._dev_nodes

    def in_range(self, K, sigGs):
        return sigma_pos(self.sg_lut, sigGs / K) <= len(self.sg_lut) - 1

    def sigma_row(self, K, sigGs, out=None):
        """(1921,) device row for this frame's sigma (isp_algos.py:225)."""
        table, _ = self.device_table()
        pos = float(sigma_pos(self.sg_lut, float(sigGs) / float(K)))
        if out is None:
            out = torch.empty(len(self.x_lut), device=table.device, dtype=torch.float32)
        check(_lib.load().yond_lut_row(ptr(table), len(self.x_lut), len(self.sg_lut), pos, ptr(out), stream_ptr()))
        return out

    def get_lut(self, x, K=1, sigGs=2, func=False):
        assert not func, "func=True (scipy interp1d object) is host-only in the reference; not part of the device path"
        if not self.in_range(K, sigGs):
            raise _lib.YondError("sigma/K outside the BiasLUT range (>= 10 e-): use the fallback table (get_bias)")
        t, np_in = to_dev(x)
        row = self.sigma_row(K, sigGs)
        _, nodes = self.device_table()
        out = torch.empty_like(t)
        check(_lib.load().yond_lut_apply(ptr(t), ptr(out), t.numel(), ptr(row), ptr(nodes), len(self.x_lut), float(K),
                                         float(sigGs), stream_ptr()))
        return _back(out, np_in)


# ------------------------------------------------------------------ A7  utils/isp_algos.py:234-242
def _as_batch4(t):
    """(h,w,C) with C = 4*n -> (n,h,w,4) batch: the SIDD_256 channel stack is 32 independent 4-channel images."""
    h, w, c = t.shape
    assert c % 4 == 0
    if c == 4:
        return t[None].contiguous()
    return t.reshape(h, w, c // 4, 4).permute(2, 0, 1, 3).contiguous()


def _from_batch4(b, c):
    n, h, w, _ = b.shape
    if c == 4:
        return b[0]
    return b.permute(1, 2, 0, 3).reshape(h, w, c).contiguous()


def blur(img, k):
    """cv2.blur(img,(k,k)) for float32 HWC images: normalised box, BORDER_REFLECT_101, float64 sums."""
    t, np_in = to_dev(img)
    b = _as_batch4(t)
    B, h, w, _ = b.shape
    lib = _lib.load()
    work = torch.empty(lib.yond_nlf_work_bytes(B, h, w, 4), device=t.device, dtype=torch.uint8)
    out = torch.empty_like(b)
    check(lib.yond_box_blur(ptr(b), ptr(out), B, h, w, 4, int(k), 0, ptr(work), stream_ptr()))
    return _back(_from_batch4(out, t.shape[2]), np_in)


def stdfilt(img, k=5):
    t, np_in = to_dev(img)
    b = _as_batch4(t)
    B, h, w, _ = b.shape
    lib = _lib.load()
    work = torch.empty(lib.yond_nlf_work_bytes(B, h, w, 4), device=t.device, dtype=torch.uint8)
    var = torch.empty_like(b)
    mean = torch.empty_like(b)
    lap = torch.empty_like(b)
    # collab maps with both inputs = img give lap = std_k(img)
    check(lib.yond_nlf_maps(ptr(b), ptr(b), ptr(var), ptr(mean), ptr(lap), B, h, w, 4, int(k), 1, ptr(work), stream_ptr()))
    return _back(_from_batch4(lap, t.shape[2]), np_in)


# ------------------------------------------------------------------ A13  utils/utils.py:246-252
def get_p2d(shape, base=16):
    xb, xc, xh, xw = shape
    yh, yw = ((xh - 1) // base + 1) * base, ((xw - 1) // base + 1) * base
    diffY, diffX = yh - xh, yw - xw
    return (diffX // 2, diffX - diffX // 2, diffY // 2, diffY - diffY // 2)


# ------------------------------------------------------------------ A6  fallback bias table (host-side generator)
def get_bias_table(img_max, sigGs, K, pho_min=1, close_form=True):
    """Node positions / values of the reference's fallback table `get_bias` (isp_algos.py:98-140): numeric
    Poisson (*) Gaussian expectation on a piecewise grid, Foi's closed form above 50*sqrt(K).  The table
    (a few hundred nodes) is host arithmetic (SciPy), exactly like the reference; its per-pixel application
    (`interp1d`) runs on the device inside yond_vst_fwd.  Device generation is a SURVEY §8(f) 'next' row."""
    from scipy.signal import convolve
    from scipy.stats import norm, poisson
    # dtype flow of the reference call sites (YOND_SIDD.py:256,395,452): the bound is a float32 scalar, K and sigma float64
    img_max, sigGs, K = np.float32(img_max), np.float64(sigGs), np.float64(K)
    lb, ub = 0, np.ceil(img_max) + 1
    if ub < 50:
        lams = np.linspace(lb, ub, int((ub - lb) / 0.1) + 2)
    elif ub < 500:
        lams = np.concatenate((np.linspace(lb, 50, int((50 - lb) / 0.1) + 1), np.linspace(50, ub, int(ub - 50) + 2)))
    else:
        lams = np.concatenate((np.linspace(lb, 50, int((50 - lb) / 0.1) + 1), np.linspace(50, 500, 451),
                               np.linspace(500, ub, int(ub - 500) // 10 + 2)))
    bias = np.zeros(len(lams), np.float32)
    pho = int(np.maximum(int(K ** 0.5), pho_min))
    sg = sigGs / K
    if close_form:
        th = 50 * K if K < 1 else 50 * K ** 0.5
        hi = lams > th
        y = lams[hi] / K
        yh = y + 3 / 8 + sg ** 2
        bias[hi] = 2 * yh ** 0.5 * (-1 / 8 * (y + sg ** 2) / yh ** 2 + 1 / 16 * y / yh ** 3
                                    - 5 / 128 * (y + 3 * (y + sg ** 2) ** 2) / yh ** 4)
    else:
        th = lams.max() + 1

    def vst(v):
        return 2 / K * np.maximum(K * v + (3 / 8) * K ** 2 + sigGs ** 2, 0) ** 0.5
    for i, lam in enumerate(lams[lams <= th]):
        r = int(lam * (1 / K) * 2 + sigGs * 2 + lam + 10)
        xs = np.linspace(-r, r, 2 * pho * r + 1)
        if sigGs > 0:
            pdf = convolve(poisson.pmf(xs, lam / K), norm.pdf(xs, loc=0, scale=sg), mode="same")
        else:
            pdf = poisson.pmf(xs, lam / K)
        pdf[pdf < 0] = 0
        pdf = pdf / (pdf.sum() / pho)
        bias[i] = np.sum(pdf * vst(K * xs) / pho) - vst(lam)
    return lams, bias
