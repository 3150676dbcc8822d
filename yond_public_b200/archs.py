"""Arch plugins with the reference's names, constructor (`Cls(arch_dict)` from the yml `arch:` block,
YOND_SIDD.py:177), call convention (`net(x)` / `net(x, t)`, :283-288) and state_dict layout
(archs/Unet.py:4-104, :288-378, :380-470) — executed by libyond_b200's tcgen05 conv stack.

The modules hold ordinary torch parameters in the reference's registration order, so `initialize_weights`
(archs/__init__.py:10-17), `load_state_dict` and `load_weights` (utils/utils.py:160-209) work unchanged; the
parameters are never used for compute by torch.  On the first forward (and whenever a parameter changes) the
weights are handed to the C library, which repacks them once into its tensor-core layout.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr, stream_ptr

ARCH_IDS = {"UNetSeeInDark": 0, "GuidedResUnet": 1, "SNRnet": 2, "ResUnet2": 3, "SelfResUNet": 4, "GuidedSelfUnet": 5}


def conv1x1(in_nc, out_nc):
    return nn.Conv2d(in_nc, out_nc, kernel_size=1, stride=1)


class conv3x3(nn.Module):
    """Stride-2 'pool' of GuidedResUnet/SNRnet.  The reference registers a ReLU as a child of the Conv2d, so it
    never runs (archs/modules.py:117-125); only `conv.weight/bias` exist in the state_dict."""

    def __init__(self, in_nc, out_nc, stride=2):
        super().__init__()
        self.conv = nn.Conv2d(in_nc, out_nc, kernel_size=3, padding=1, stride=stride)


class GuidedResidualBlock(nn.Module):  # parameters only — archs/modules.py:163-183
    def __init__(self, in_c, out_c):
        super().__init__()
        self.conv1 = nn.Conv2d(out_c, out_c, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(out_c, out_c, 3, 1, 1, bias=True)
        self.gamma = nn.Sequential(conv1x1(1, out_c), nn.SiLU(), conv1x1(out_c, out_c))
        self.beta = nn.Sequential(nn.SiLU(), conv1x1(out_c, out_c))
        self.short_cut = nn.Sequential(conv1x1(in_c, out_c)) if in_c != out_c else nn.Sequential(OrderedDict([]))


class SNR_Block(nn.Module):  # parameters only — archs/modules.py:198-218
    def __init__(self, in_c, out_c):
        super().__init__()
        self.conv1 = nn.Conv2d(out_c, out_c, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(out_c, out_c, 3, 1, 1, bias=True)
        self.sfm1 = nn.Sequential(conv1x1(1, out_c), nn.SiLU(), conv1x1(out_c, out_c))
        self.sfm2 = nn.Sequential(conv1x1(1, out_c), nn.SiLU(), conv1x1(out_c, out_c))
        self.short_cut = nn.Sequential(conv1x1(in_c, out_c)) if in_c != out_c else nn.Sequential(OrderedDict([]))


class _B200Net(nn.Module):
    """Shared machinery: C handle, weight sync, workspace, forward through yond_net_forward[_nchw]."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.nframes = args.get("nframes", 1)
        self.cf = 0
        self.res = args["res"]
        self.norm = args["norm"] if "norm" in args else False
        assert self.nframes == 1 and args["in_nc"] == 4 and args["out_nc"] == 4, \
            "the B200 plugin builds the packed-Bayer configuration of the shipped yml files (in_nc=out_nc=4, nframes=1)"
        self._handle = None
        self._synced = None
        self._ws = None
        self.conv_impl = 0  # 0 = tcgen05 kernels; 1 = CUDA-core cross-check (tests only)

    # -- C handle ------------------------------------------------------------------------------
    def _get_handle(self):
        if self._handle is None:
            lib = _lib.load()
            h = C.c_void_p()
            check(lib.yond_net_create(ARCH_IDS[type(self).__name__], 4, 4, int(self.args["nf"]), int(bool(self.res)),
                                      int(bool(self.norm)), C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().yond_net_destroy(self._handle)
        except Exception:
            pass

    def expected_state(self):
        """(key, shape) list the C library expects — must equal this module's state_dict."""
        lib, h = _lib.load(), self._get_handle()
        out = []
        shp = (C.c_int64 * 4)()
        for i in range(lib.yond_net_num_keys(h)):
            nd = lib.yond_net_key_shape(h, i, shp)
            out.append((lib.yond_net_key(h, i).decode(), tuple(int(shp[d]) for d in range(nd))))
        return out

    def _sync_weights(self):
        sd = self.state_dict()
        version = tuple((k, v._version, v.data_ptr()) for k, v in sd.items())
        if version == self._synced:
            return
        lib, h = _lib.load(), self._get_handle()
        for k, v in sd.items():
            a = np.ascontiguousarray(v.detach().float().cpu().numpy())
            shape = (C.c_int64 * a.ndim)(*a.shape)
            check(lib.yond_net_set_tensor(h, k.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim))
        buf = C.create_string_buffer(4096)
        if lib.yond_net_missing(h, buf, 4096) != 0:
            raise _lib.YondError(f"state_dict lacks tensors the network needs: {buf.value.decode()}")
        self._synced = version

    def invalidate_weights(self):
        """Call after modifying parameters through `.data` in place (which torch's version counters do not see)."""
        self._synced = None

    def load_state_dict(self, *a, **kw):
        self._synced = None
        return super().load_state_dict(*a, **kw)

    def _apply(self, fn, *a, **kw):
        self._synced = None
        return super()._apply(fn, *a, **kw)

    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(int(nbytes) + 1024, device=device, dtype=torch.uint8)
        off = (-self._ws.data_ptr()) % 1024
        return self._ws[off:]

    def flops(self, B, H, W):
        return float(_lib.load().yond_net_flops(self._get_handle(), B, H, W))

    def set_profile(self, enable=True):
        check(_lib.load().yond_net_profile(self._get_handle(), int(enable)))

    def read_profile(self, reset=True):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int()
        check(_lib.load().yond_net_profile_read(self._get_handle(), C.byref(ms), C.byref(fl), C.byref(n), int(reset)))
        return dict(conv_ms=ms.value, conv_flops=fl.value, launches=n.value)

    # -- forward -------------------------------------------------------------------------------
    def _t_vector(self, t, B, device):
        t = torch.as_tensor(t, dtype=torch.float32, device=device).reshape(-1)
        return (t.expand(B) if t.numel() == 1 else t).contiguous()

    def forward_nhwc(self, z, ub, t=None, out=None):
        """Fused-pipeline entry: z (B,H,W,4) f32 CUDA in [0,1], ub (B) per-sample max, t (B) or None."""
        lib, h = _lib.load(), self._get_handle()
        self._sync_weights()
        check(lib.yond_net_set_conv_impl(h, int(self.conv_impl)))
        B, H, W, _ = z.shape
        ws = self._workspace(lib.yond_net_workspace_bytes(h, B, H, W), z.device)
        y = torch.empty_like(z) if out is None else out
        tv = self._t_vector(t, B, z.device) if t is not None else None
        check(lib.yond_net_forward(h, ptr(z), ptr(ub), ptr(tv), ptr(y), B, H, W, ptr(ws), ws.numel(), stream_ptr()))
        return y

    def _forward_nchw(self, x, t=None):
        if not x.is_cuda:
            raise _lib.YondError("the B200 arch plugins run on CUDA tensors only (no CPU path): move the net and input to 'cuda'")
        lib, h = _lib.load(), self._get_handle()
        self._sync_weights()
        check(lib.yond_net_set_conv_impl(h, int(self.conv_impl)))
        x = x.float().contiguous()
        B, Cc, H, W = x.shape
        assert Cc == 4
        ws = self._workspace(lib.yond_net_workspace_bytes(h, B, H, W), x.device)
        y = torch.empty_like(x)
        tv = self._t_vector(t, B, x.device) if t is not None else None
        check(lib.yond_net_forward_nchw(h, ptr(x), ptr(tv), ptr(y), B, H, W, ptr(ws), ws.numel(), stream_ptr()))
        return y


class UNetSeeInDark(_B200Net):
    """archs/Unet.py:4-104."""

    def __init__(self, args=None):
        super().__init__(args)
        nf, in_nc, out_nc = args["nf"], args["in_nc"], args["out_nc"]
        chans = [nf, nf * 2, nf * 4, nf * 8, nf * 16]
        prev = in_nc * self.nframes
        for i, c in enumerate(chans, start=1):
            setattr(self, f"conv{i}_1", nn.Conv2d(prev, c, kernel_size=3, stride=1, padding=1))
            setattr(self, f"conv{i}_2", nn.Conv2d(c, c, kernel_size=3, stride=1, padding=1))
            if i < 5:
                setattr(self, f"pool{i}", nn.MaxPool2d(kernel_size=2))
            prev = c
        for i, c in zip(range(6, 10), chans[-2::-1]):
            setattr(self, f"upv{i}", nn.ConvTranspose2d(c * 2, c, 2, stride=2))
            setattr(self, f"conv{i}_1", nn.Conv2d(c * 2, c, kernel_size=3, stride=1, padding=1))
            setattr(self, f"conv{i}_2", nn.Conv2d(c, c, kernel_size=3, stride=1, padding=1))
        self.conv10_1 = nn.Conv2d(nf, out_nc, kernel_size=1, stride=1)
        self.relu = nn.LeakyReLU(0.2, inplace=True)

    def forward(self, x):
        return self._forward_nchw(x)


class _GuidedBase(_B200Net):
    _block = None

    def __init__(self, args=None):
        super().__init__(args)
        nf, in_nc, out_nc = args["nf"], args["in_nc"], args["out_nc"]
        Block = self._block
        self.conv_in = nn.Conv2d(in_nc * self.nframes, nf, kernel_size=3, stride=1, padding=1)
        c = nf
        for i in range(1, 5):
            setattr(self, f"conv{i}", Block(c, c))
            setattr(self, f"pool{i}", conv3x3(c, c * 2))
            c *= 2
        self.conv5 = Block(c, c)
        for i in range(6, 10):
            setattr(self, f"upv{i}", nn.ConvTranspose2d(c, c // 2, 2, stride=2))
            setattr(self, f"conv{i}", Block(c, c // 2))
            c //= 2
        self.conv10 = nn.Conv2d(nf, out_nc, kernel_size=1, stride=1)
        self.lrelu = nn.LeakyReLU(inplace=True)

    def forward(self, x, t):
        return self._forward_nchw(x, t)


class GuidedResUnet(_GuidedBase):
    """archs/Unet.py:380-470 ('GRU' in the yml names)."""
    _block = GuidedResidualBlock


class SNRnet(_GuidedBase):
    """archs/Unet.py:288-378."""
    _block = SNR_Block


class ResBlock(GuidedResidualBlock):
    """archs/modules.py:235-265 — registers the same conv1 / conv2 / gamma / beta / short_cut modules as the guided block
    (is_activate=False: SiLU); its forward never touches gamma / beta."""


class ResUnet2(_GuidedBase):
    """archs/Unet.py:197-286: GuidedResUnet's graph without the noise-level conditioning; LeakyReLU(0.2) after conv_in;
    forward(x, noise_map=None) ignores the second argument."""
    _block = ResBlock

    def forward(self, x, noise_map=None):
        return self._forward_nchw(x)


class LR(nn.Module):  # parameters only — archs/comp.py:709-722
    def __init__(self, in_size, out_size, ksize=3, slope=0.1):
        super().__init__()
        self.block = nn.Sequential(nn.Conv2d(in_size, out_size, kernel_size=ksize, padding=ksize // 2, bias=True),
                                   nn.LeakyReLU(slope, inplace=False))


class Res(nn.Module):  # parameters only — archs/comp.py:830-850 (RUP, :804-828, registers the same modules)
    def __init__(self, in_size, out_size, slope=0.1, ksize=3):
        super().__init__()
        self.conv_1 = LR(out_size, out_size, ksize=ksize, slope=slope)
        self.conv_2 = LR(out_size, out_size, ksize=ksize, slope=slope)
        self.short_cut = nn.Sequential(conv1x1(in_size, out_size)) if in_size != out_size else nn.Sequential(OrderedDict([]))


class GLR(nn.Module):  # parameters only — archs/comp.py:912-934 (conv, z*tk + tb, LeakyReLU)
    def __init__(self, in_size, out_size, ksize=3, slope=0.1):
        super().__init__()
        self.block = nn.Conv2d(in_size, out_size, kernel_size=ksize, padding=ksize // 2, bias=True)
        self.act = nn.LeakyReLU(slope, inplace=False)
        self.gamma = nn.Sequential(conv1x1(1, out_size), nn.SiLU(), conv1x1(out_size, out_size))
        self.beta = nn.Sequential(nn.SiLU(), conv1x1(out_size, out_size))


class GRes(nn.Module):  # parameters only — archs/comp.py:936-954 (GUP, :956-983, registers the same modules)
    def __init__(self, in_size, out_size, slope=0.1, ksize=3):
        super().__init__()
        self.conv_1 = LR(out_size, out_size, ksize=ksize)
        self.conv_2 = GLR(out_size, out_size, ksize=ksize)
        self.short_cut = nn.Sequential(conv1x1(in_size, out_size)) if in_size != out_size else nn.Sequential(OrderedDict([]))


class GuidedSelfUnet(_B200Net):
    """archs/comp.py:852-910: SelfResUNet's graph with noise-level conditioning (GRes head / last, single GLR down
    levels, GUP up levels); called as net(x, t).  `res` must be False: the reference's res branch adds the 2nf-channel features to the
    4-channel output and cannot run."""

    def __init__(self, args):
        args = dict(args)
        args.setdefault("res", False)
        assert not args["res"], "GuidedSelfUnet: res=True cannot run in the reference (archs/comp.py:904-905)"
        super().__init__(args)
        nf = args["nf"] if "nf" in args else 32
        assert args.get("depth", 5) == 5 and args.get("slope", 0.1) == 0.1, "the B200 plugin builds the class defaults (depth 5, slope 0.1)"
        in_nc, out_nc, depth = args["in_nc"], args["out_nc"], 5
        self.depth = depth
        self.head = GRes(in_nc, nf)
        self.down_path = nn.ModuleList([GLR(nf, nf, 3) for _ in range(depth)])
        self.up_path = nn.ModuleList([GRes((nf * 2 if i == 0 else nf * 3) if i != depth - 1 else nf * 2 + in_nc, nf * 2) for i in range(depth)])
        self.last = GRes(2 * nf, 2 * nf, ksize=1)
        self.out = conv1x1(2 * nf, out_nc)

    def forward(self, x, t):
        return self._forward_nchw(x, t)


class SelfResUNet(_B200Net):
    """archs/comp.py:745-802: constant-width residual U-Net — Res(4, nf) head, five max-pool + Res(nf, nf) levels,
    five nearest-neighbour up levels RUP(., 2 nf) that concatenate the pooled features (the network input at the last one),
    Res(2 nf, 2 nf, ksize=1), 1x1 output.  depth = 5, slope = 0.1 (the class defaults); H and W multiples of 32."""

    def __init__(self, args):
        args = dict(args)
        args.setdefault("res", False)
        super().__init__(args)
        nf = args["nf"] if "nf" in args else 32
        assert args.get("depth", 5) == 5 and args.get("slope", 0.1) == 0.1, "the B200 plugin builds the class defaults (depth 5, slope 0.1)"
        in_nc, out_nc, depth = args["in_nc"], args["out_nc"], 5
        self.depth = depth
        self.head = Res(in_nc, nf)
        self.down_path = nn.ModuleList([Res(nf, nf) for _ in range(depth)])
        self.up_path = nn.ModuleList([Res((nf * 2 if i == 0 else nf * 3) if i != depth - 1 else nf * 2 + in_nc, nf * 2) for i in range(depth)])
        self.last = Res(2 * nf, 2 * nf, ksize=1)
        self.out = conv1x1(2 * nf, out_nc)

    def forward(self, x):
        return self._forward_nchw(x)


def initialize_weights(net):
    """archs/__init__.py:10-17 — N(0,0.02) for conv weight+bias and ConvT weight (ConvT bias: torch default)."""
    for m in net.modules():
        if isinstance(m, nn.Conv2d):
            m.weight.data.normal_(0.0, 0.02)
            if m.bias is not None:
                m.bias.data.normal_(0.0, 0.02)
        if isinstance(m, nn.ConvTranspose2d):
            m.weight.data.normal_(0.0, 0.02)
    if hasattr(net, "invalidate_weights"):
        net.invalidate_weights()


def load_weights(model, pretrained_dict, multi_gpu=False, by_name=False):
    """utils/utils.py:160-209: `tsm_shift` -> `tsm_buffer` key remap, optional by-name filtering (unknown keys and
    shape mismatches are dropped), then update-and-load over the model's own state_dict."""
    target = model.module if multi_gpu else model
    model_dict = target.state_dict()
    pretrained_dict = dict(pretrained_dict)
    for k in [k for k in pretrained_dict if "tsm_shift" in k]:
        pretrained_dict[k.replace("tsm_shift", "tsm_buffer")] = pretrained_dict[k]
    if by_name:
        pretrained_dict = {k: v for k, v in pretrained_dict.items()
                           if k in model_dict and model_dict[k].shape == v.shape}
    model_dict.update(pretrained_dict)
    target.load_state_dict(model_dict)
    return model
