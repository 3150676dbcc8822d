"""Function-level surface of the YOND hot path, same names / argument meaning as the reference
(utils/isp_ops.py, utils/isp_algos.py, utils/utils.py, YOND_SIDD.py) — backed by libyond_b200 kernels.

Inputs may be NumPy arrays (host, like the reference) or CUDA torch tensors; the result comes back in the
same kind.  Host inputs are staged through pinned memory.  There is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

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


_STAGE = threading.local()


def _pinned_stage(nbytes):
    """Grow-only pinned host buffer of this thread: NumPy inputs go host -> pinned -> device without a cudaHostAlloc per call."""
    buf = getattr(_STAGE, "buf", None)
    if buf is None or buf.numel() < nbytes:
        buf = _STAGE.buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8).pin_memory()
    return buf


def pinned_view(name, like):
    """A view shaped / typed like the tensor `like` into a grow-only pinned buffer of this thread named `name`."""
    n = like.numel() * like.element_size()
    buf = getattr(_STAGE, name, None)
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(int(n), 1 << 20), dtype=torch.uint8).pin_memory()
        setattr(_STAGE, name, buf)
    return buf[:n].view(like.dtype).view(like.shape)


_COPY_POOL = None


def _host_copy(dst: np.ndarray, src: np.ndarray):
    """dst[...] = src for large host arrays with a few threads (NumPy releases the GIL in copyto): the per-call NumPy surface of
    the reference moves ~150 MB of host memory per 12 MP frame, and one core copies at only ~6-8 GB/s."""
    n = dst.size * dst.itemsize
    if n < (8 << 20) or not (dst.flags.c_contiguous and src.flags.c_contiguous):
        np.copyto(dst, src)
        return
    global _COPY_POOL
    if _COPY_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(max_workers=4, thread_name_prefix="yond-copy")
    d, s_ = dst.reshape(-1), src.reshape(-1)
    step = -(-d.size // 4)
    list(_COPY_POOL.map(lambda i: np.copyto(d[i:i + step], s_[i:i + step]), range(0, d.size, step)))


def to_dev(a, dtype=torch.float32):
    """numpy / tensor -> contiguous CUDA tensor; returns (tensor, was_numpy)."""
    dev = _dev()  # raises without CUDA: there is no CPU path
    if isinstance(a, np.ndarray):
        src = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32 if dtype == torch.float32 else None))
        n = src.numel() * src.element_size()
        stage = _pinned_stage(n)[:n].view(src.dtype).view(src.shape)
        torch.cuda.current_stream(dev).synchronize()  # the previous upload out of this buffer has left it
        _host_copy(stage.numpy(), src.numpy())
        return stage.to(dev, non_blocking=True), True
    if not torch.is_tensor(a):
        a = torch.as_tensor(np.asarray(a, dtype=np.float32))
        return a.to(dev), True
    return a.to(device=dev, dtype=dtype).contiguous(), False


def to_host(tensors):
    """CUDA tensors -> fresh NumPy arrays through ONE pinned staging region and one synchronisation (a plain .cpu() stages
    every tensor through the driver's bounce buffers at a fraction of the PCIe rate)."""
    if not tensors:
        return []
    tensors = [t.contiguous() for t in tensors]
    sizes = [t.numel() * t.element_size() for t in tensors]
    offs = np.concatenate([[0], np.cumsum([(n + 255) // 256 * 256 for n in sizes])]).astype(np.int64)
    stage = getattr(_STAGE, "down", None)
    if stage is None or stage.numel() < int(offs[-1]):
        stage = _STAGE.down = torch.empty(max(int(offs[-1]), 1 << 20), dtype=torch.uint8).pin_memory()
    views = []
    for t, o, n in zip(tensors, offs[:-1], sizes):
        v = stage[int(o):int(o) + n].view(t.dtype).view(t.shape)
        v.copy_(t, non_blocking=True)
        views.append(v)
    torch.cuda.current_stream(tensors[0].device).synchronize()
    outs = []
    for v in views:
        o = np.empty(tuple(v.shape), dtype=v.numpy().dtype)
        _host_copy(o, v.numpy())
        outs.append(o)
    return outs


def _back(t, was_numpy):
    return to_host([t])[0] if was_numpy else t


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


def normalize_raw(raw, bl, wp, ratio=1, clip=False):
    """The `data['lr']` of the 14-bit dataset drivers (data_process/yond_datasets.py:955-961, :1053-1056):
    (raw.astype(float32) - bl) * ratio / (wp - bl) on the uint16 mosaic (any shape; NumPy or a 16-bit CUDA tensor), float32 result
    on the device for tensors / as NumPy for NumPy input.  The low-light gain `ratio` is applied here, which is why the driver's
    p['scale'] is (wp - bl) / ratio."""
    dev = _dev()
    np_in = not torch.is_tensor(raw)
    if np_in:
        a = np.ascontiguousarray(np.asarray(raw))
        assert a.dtype == np.uint16, "the sensor mosaic must be uint16"
        x = torch.from_numpy(a.view(np.int16)).to(dev)
    else:
        assert raw.dtype in (torch.uint16, torch.int16), "the sensor mosaic must be a 16-bit integer tensor"
        x = raw.to(dev).contiguous()
    out = torch.empty(x.shape, device=dev, dtype=torch.float32)
    check(_lib.load().yond_ingest_mosaic(ptr(x), ptr(out), x.numel(), float(bl), float(wp), float(ratio), int(bool(clip)), stream_ptr()))
    return _back(out, np_in)


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
    (1921,1101) [x,sigma]) or the .npz stand-in shipped in yond_public_b200/data (key 'bias_lut').

    The reference enables the LUT only when checkpoints/bias_lut_2d.npy exists (YOND_SIDD.py:171); the authors' file is
    not distributed, so `BiasLUT()` without that file loads the stand-in — a float32 regeneration with the reference's own
    get_bias_points(x_lut, 1, sg, pho_min=100, close_form=True) (tests/golden/make_bias_lut.py) — and says so in
    `self.source`.  Pass `standin=False` to get the reference behaviour (FileNotFoundError -> the driver runs LUT-less)."""

    def __init__(self, lut_path="checkpoints/bias_lut_2d.npy", standin=True):
        self.source = lut_path
        if not os.path.exists(lut_path) and lut_path == "checkpoints/bias_lut_2d.npy":
            if not standin:
                raise FileNotFoundError(lut_path)
            lut_path = DEFAULT_LUT
            self.source = f"stand-in table {DEFAULT_LUT} (checkpoints/bias_lut_2d.npy not found)"
        arr = np.load(lut_path)
        table = arr["bias_lut"] if hasattr(arr, "files") else arr
        self.x_lut, self.sg_lut = lut_grids()
        assert table.shape == (len(self.x_lut), len(self.sg_lut)), f"bias LUT must be (1921,1101), got {table.shape}"
        self.bias_lut = np.ascontiguousarray(table, dtype=np.float32)
        self._dev = None
        self._lock = threading.Lock()

    def device_arrays(self):
        """(table (1921,1101) f32, x nodes (1921) f32, sigma nodes (1101) f64) on the current device; uploaded once, under a
        lock and synchronised, so concurrent host lanes never see a partially uploaded table."""
        if self._dev is None:
            with self._lock:
                if self._dev is None:
                    dev = _dev()
                    arrs = (torch.from_numpy(self.bias_lut).to(dev), torch.from_numpy(self.x_lut.astype(np.float32)).to(dev),
                            torch.from_numpy(np.ascontiguousarray(self.sg_lut, np.float64)).to(dev))
                    torch.cuda.synchronize(dev)
                    self._dev = arrs
        return self._dev

    def device_table(self):
        t, x, _ = self.device_arrays()
        return t, x

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
            # sigma/K beyond the table: the reference falls back to get_bias (more than 1000 points: the interpolated
            # table up to x.max()) or get_bias_points (exact per point, pho_min = 100) — isp_algos.py:204-212
            t, np_in = to_dev(x)
            if t.numel() > 1000:
                nodes, vals = get_bias_table(float(t.max()), sigGs, K, device=True)
                out = torch.empty_like(t)
                n = nodes.numel()
                # piecewise-linear evaluation on the device: the table as a one-row fallback table of the fused front end
                check(_lib.load().yond_table_apply(ptr(t), ptr(out), t.numel(), ptr(vals), ptr(nodes), n, stream_ptr()))
                return _back(out, np_in)
            return _back(get_bias_points(t.reshape(-1).double(), K, sigGs, pho_min=100).reshape(t.shape).float(), np_in)
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


# ------------------------------------------------------------------ A6  fallback bias table (device generator)
def bias_table_nodes(img_max):
    """Node positions of get_bias (isp_algos.py:101-108) in the reference's own dtype flow: `ub` is a float32 scalar, so
    the pieces ending in `ub` are float32 linspaces.  Host index math only (the values come from the device)."""
    img_max = np.float32(img_max)
    lb, ub = 0, np.ceil(img_max) + 1
    if ub < 50:
        return np.linspace(lb, ub, int((ub - lb) / 0.1) + 2)
    if ub < 500:
        return np.concatenate((np.linspace(lb, 50, int((50 - lb) / 0.1) + 1), np.linspace(50, ub, int(ub - 50) + 2)))
    return np.concatenate((np.linspace(lb, 50, int((50 - lb) / 0.1) + 1), np.linspace(50, 500, 451),
                           np.linspace(500, ub, int(ub - 500) // 10 + 2)))


def get_bias_table(img_max, sigGs, K, pho_min=1, close_form=True, device=False):
    """Node positions / values of the reference's fallback table `get_bias` (isp_algos.py:98-140): numeric
    Poisson (*) Gaussian expectation on a piecewise grid, Foi's closed form above 50*sqrt(K) — generated ON THE DEVICE
    (yond_bias_table).  Returns (nodes float64, values float32) NumPy arrays like `interp1d(...).x/.y`, or
    with device=True the (nodes, values) float32 CUDA tensors the kernels consume."""
    assert pho_min == 1 and close_form, "the YOND path calls get_bias with its defaults (pho_min=1, close_form=True)"
    lib = _lib.load()
    dev = _dev()
    bound = float(np.float32(img_max))
    n = int(lib.yond_bias_table_nodes(bound))
    nodes = torch.empty(n, device=dev, dtype=torch.float32)
    vals = torch.empty(n, device=dev, dtype=torch.float32)
    work = torch.empty(int(lib.yond_chain_work_bytes(1)), device=dev, dtype=torch.uint8)
    check(lib.yond_bias_table(float(K), float(sigGs), bound, ptr(nodes), ptr(vals), n, None, ptr(work), stream_ptr()))
    if device:
        return nodes, vals
    lams = bias_table_nodes(img_max)
    assert len(lams) == n
    return lams, vals.cpu().numpy()


def get_bias_points(lams, K, sigGs, pho_min=100, close_form=True):
    """get_bias_points (isp_algos.py:142-160) on the device: the bias at explicit points (ascending, like the reference
    assumes), float64.  NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out."""
    assert close_form, "only close_form=True is used on the YOND path"
    lib = _lib.load()
    dev = _dev()
    np_in = not torch.is_tensor(lams)
    t = (torch.from_numpy(np.ascontiguousarray(lams, np.float64)) if np_in else lams).to(device=dev, dtype=torch.float64).contiguous()
    out = torch.empty_like(t)
    work = torch.empty(int(lib.yond_chain_work_bytes(1)), device=dev, dtype=torch.uint8)
    check(lib.yond_bias_points(ptr(t), t.numel(), float(K), float(sigGs), int(pho_min), ptr(out), ptr(work), stream_ptr()))
    return out.cpu().numpy() if np_in else out
