"""Image-quality metrics of the SIDD driver on the device: the reference evaluates raw PSNR and a
MATLAB-style SSIM per 256x256 block of every denoised mosaic in CPU worker threads (YOND_SIDD.py:635-697); here one call
measures every block of a batch of mosaics on the GPU.  Same names as the reference: `compare_psnr` (its alias of
skimage.metrics.peak_signal_noise_ratio), `ssim`, `calculate_ssim`; `block_metrics` is the batched form the driver uses.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .isp import to_dev


def gaussian_kernel(ksize=11, sigma=1.5):
    """cv2.getGaussianKernel(ksize, sigma) for sigma > 0: exp(-(i-c)^2 / (2 sigma^2)) normalised to sum 1, float64."""
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-0.5 / (sigma * sigma) * x * x)
    return k * (1.0 / k.sum())


_WIN = np.ascontiguousarray(gaussian_kernel(11, 1.5))


def block_metrics(a, b, nblk=1, data_range=1.0, ssim_scale=255.0, psnr=True, ssim=True):
    """a, b: (nimg, H, nblk*Wb) (or (H, nblk*Wb)) mosaics.  Returns device tensors (psnr, ssim), each (nimg, nblk) float64 (None
    when not requested): raw PSNR and the SSIM of (a*255, b*255) of every block, like YOND_SIDD.py:649-652."""
    ta, _ = to_dev(a)
    tb, _ = to_dev(b)
    if ta.dim() == 2:
        ta, tb = ta[None], tb[None]
    ta, tb = ta.contiguous(), tb.contiguous()
    assert ta.shape == tb.shape and ta.dtype == torch.float32 and tb.dtype == torch.float32
    nimg, H, Wm = ta.shape
    out_p = torch.empty((nimg, nblk), device=ta.device, dtype=torch.float64) if psnr else None
    out_s = torch.empty((nimg, nblk), device=ta.device, dtype=torch.float64) if ssim else None
    check(_lib.load().yond_block_metrics(ptr(ta), ptr(tb), nimg, H, Wm, int(nblk), float(data_range), float(ssim_scale),
                                         _WIN.ctypes.data_as(C.POINTER(C.c_double)), ptr(out_p), ptr(out_s), stream_ptr()))
    return out_p, out_s


def compare_psnr(image_true, image_test, data_range=1):
    """skimage.metrics.peak_signal_noise_ratio for float32 images (YOND_SIDD.py:651)."""
    p, _ = block_metrics(image_true, image_test, 1, data_range=data_range, ssim=False)
    return float(p.cpu()[0, 0])


def ssim(prediction, target):
    """YOND_SIDD.py:679-697 on one 2-D image pair already in [0, 255]."""
    _, s = block_metrics(np.asarray(prediction, np.float32), np.asarray(target, np.float32), 1, ssim_scale=1.0, psnr=False)
    return float(s.cpu()[0, 0])


def calculate_ssim(target, ref):
    """YOND_SIDD.py:700-721: 2-D images, or (H, W, 3) / (H, W, 1) channel means."""
    img1, img2 = np.asarray(target), np.asarray(ref)
    if img1.shape != img2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if img1.ndim == 2:
        return ssim(img1, img2)
    if img1.ndim == 3:
        if img1.shape[2] == 3:
            return float(np.mean([ssim(img1[:, :, i], img2[:, :, i]) for i in range(3)]))
        if img1.shape[2] == 1:
            return ssim(np.squeeze(img1), np.squeeze(img2))
    raise ValueError("Wrong input image dimensions.")


def sidd_image_metrics(outputs, hr_raw, nblk=32):
    """The per-image numbers of multiprocess_plot (YOND_SIDD.py:643-656): mean over the mosaic's blocks of raw PSNR and SSIM; an
    output whose maximum is <= 0 scores (-1, -1).  outputs / hr_raw: (H, nblk*Wb) mosaics (NumPy or CUDA)."""
    t, _ = to_dev(outputs)
    if float(t.max()) <= 0:
        return -1.0, -1.0
    p, s = block_metrics(t, hr_raw, nblk)
    return float(p.mean().cpu()), float(s.mean().cpu())
