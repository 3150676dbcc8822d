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


def _is_u8(a):
    return (isinstance(a, np.ndarray) and a.dtype == np.uint8) or (torch.is_tensor(a) and a.dtype == torch.uint8)


def block_metrics_rgb8(a, b, nblk=1, psnr=True, ssim=True):
    """a, b: (nimg, H, nblk*Wb, 3) (or (H, nblk*Wb, 3)) uint8 pictures.  Returns device tensors (psnr, ssim), each (nimg, nblk)
    float64: PSNR at data_range 255 and the channel-mean SSIM of every block, like YOND_SIDD.py:661-665."""
    ta, _ = to_dev(a, dtype=torch.uint8)
    tb, _ = to_dev(b, dtype=torch.uint8)
    if ta.dim() == 3:
        ta, tb = ta[None], tb[None]
    ta, tb = ta.contiguous(), tb.contiguous()
    assert ta.shape == tb.shape and ta.dtype == torch.uint8 and tb.dtype == torch.uint8 and ta.shape[-1] == 3
    nimg, H, Wm, _ = ta.shape
    out_p = torch.empty((nimg, nblk), device=ta.device, dtype=torch.float64) if psnr else None
    out_s = torch.empty((nimg, nblk), device=ta.device, dtype=torch.float64) if ssim else None
    check(_lib.load().yond_block_metrics_rgb8(ptr(ta), ptr(tb), nimg, H, Wm, int(nblk), _WIN.ctypes.data_as(C.POINTER(C.c_double)),
                                              ptr(out_p), ptr(out_s), stream_ptr()))
    return out_p, out_s


def compare_psnr(image_true, image_test, data_range=1):
    """skimage.metrics.peak_signal_noise_ratio: float32 images (YOND_SIDD.py:651) or uint8 (H,W,3) pictures (:663)."""
    if _is_u8(image_true) and _is_u8(image_test):
        if data_range != 255:
            raise ValueError("uint8 pictures are measured at data_range=255")
        p, _ = block_metrics_rgb8(image_true, image_test, 1, ssim=False)
    else:
        p, _ = block_metrics(image_true, image_test, 1, data_range=data_range, ssim=False)
    return float(p.cpu()[0, 0])


def ssim(prediction, target):
    """YOND_SIDD.py:679-697 on one 2-D image pair already in [0, 255]."""
    _, s = block_metrics(np.asarray(prediction, np.float32), np.asarray(target, np.float32), 1, ssim_scale=1.0, psnr=False)
    return float(s.cpu()[0, 0])


def calculate_ssim(target, ref):
    """YOND_SIDD.py:700-721: 2-D images, or (H, W, 3) / (H, W, 1) channel means."""
    if _is_u8(target) and _is_u8(ref) and target.ndim == 3 and target.shape[2] == 3:
        if tuple(target.shape) != tuple(ref.shape):
            raise ValueError("Input images must have the same dimensions.")
        _, s = block_metrics_rgb8(target, ref, 1, psnr=False)
        return float(s.cpu()[0, 0])
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


def sidd_rgb_metrics(img_dn, img_hr, nblk=32):
    """YOND_SIDD.py:660-665: mean over the picture's blocks (split along W) of the sRGB PSNR and SSIM.  (H, nblk*Wb, 3) uint8."""
    p, s = block_metrics_rgb8(img_dn, img_hr, nblk)
    return float(p.mean().cpu()), float(s.mean().cpu())


def sidd_eval_image(outputs, hr_raw, meta=None, nblk=32, rgb=True):
    """The scoring half of multiprocess_plot (YOND_SIDD.py:635-668) for one image, all on the device: for every iteration's
    mosaic in `outputs` ((n_iter, H, nblk*Wb) or a list) the raw PSNR / SSIM against `hr_raw`, and — with `meta` (keys 'bayer_2by2',
    'wb', 'cst2') and rgb=True, the reference's save_plot branch — the sRGB PSNR / SSIM of the rendered pictures.  An iteration whose
    output has no positive value scores -1 like the reference and is not rendered.  Returns {'psnr': [...], 'ssim': [...],
    'psnr_rgb': [...], 'ssim_rgb': [...], 'img_hr': CUDA uint8 picture or None, 'img_dn': [pictures]}."""
    from .render import process_sidd_image
    hr, _ = to_dev(hr_raw)
    res = {"psnr": [], "ssim": [], "psnr_rgb": [], "ssim_rgb": [], "img_hr": None, "img_dn": []}
    do_rgb = bool(rgb and meta is not None)
    if do_rgb:
        res["img_hr"] = process_sidd_image(hr, meta["bayer_2by2"], meta["wb"], meta["cst2"])
    for out in outputs:
        t, _ = to_dev(out)
        if float(t.max()) <= 0:
            res["psnr"].append(-1.0)
            res["ssim"].append(-1.0)
            res["img_dn"].append(None)
            continue
        p, s = block_metrics(t, hr, nblk)
        res["psnr"].append(float(p.mean().cpu()))
        res["ssim"].append(float(s.mean().cpu()))
        if do_rgb:
            img = process_sidd_image(t, meta["bayer_2by2"], meta["wb"], meta["cst2"])
            pr, sr = block_metrics_rgb8(img, res["img_hr"], nblk)
            res["psnr_rgb"].append(float(pr.mean().cpu()))
            res["ssim_rgb"].append(float(sr.mean().cpu()))
            res["img_dn"].append(img)
    return res
