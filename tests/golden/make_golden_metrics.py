"""Golden vectors for the SIDD driver's metrics from the UNMODIFIED reference: `calculate_ssim` / `ssim`
(YOND_SIDD.py:679-721) are called as they are; `compare_psnr` is scikit-image's peak_signal_noise_ratio, which is not installed
in the build container, so its value comes from the oracle's restatement and is stored for regression only.

    python tests/golden/make_golden_metrics.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import yond_oracle as O  # noqa: E402
from oracle.ref_harness import load_reference  # noqa: E402


def main():
    ref = load_reference()
    import cv2
    rng = np.random.default_rng(31)
    nblk, H, Wb = 4, 96, 80
    clean = np.concatenate([O.synth_clean(rng, H, Wb) for _ in range(nblk)], axis=-1)
    noisy = O.synth_noisy(rng, clean, 4.0, 6.0)
    dn = (0.7 * clean + 0.3 * noisy).astype(np.float32)  # a stand-in "denoised" mosaic
    dn_ = np.array(np.split(dn, nblk, axis=-1))
    hr_ = np.array(np.split(clean, nblk, axis=-1))
    ssim_blocks = np.array([ref.Y.calculate_ssim(d * 255, h * 255) for d, h in zip(dn_, hr_)], np.float64)
    ssim_noisy = np.float64(ref.Y.calculate_ssim(noisy[:, :Wb] * 255, clean[:, :Wb] * 255))
    rgb = rng.uniform(0, 255, (40, 48, 3)).astype(np.float32)
    rgb2 = np.clip(rgb + rng.normal(0, 9, rgb.shape), 0, 255).astype(np.float32)
    ssim_rgb = np.float64(ref.Y.calculate_ssim(rgb, rgb2))
    psnr_blocks = np.array([O.compare_psnr(d, h, data_range=1) for d, h in zip(dn_, hr_)], np.float64)
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), seed=31, nblk=nblk, H=H, Wb=Wb, clean=clean, noisy=noisy, dn=dn,
                        ssim_blocks=ssim_blocks, ssim_noisy=ssim_noisy, rgb=rgb, rgb2=rgb2, ssim_rgb=ssim_rgb,
                        psnr_blocks=psnr_blocks, gauss11=cv2.getGaussianKernel(11, 1.5).reshape(-1))
    print("ssim", ssim_blocks, ssim_noisy, ssim_rgb, "psnr", psnr_blocks)


if __name__ == "__main__":
    main()
