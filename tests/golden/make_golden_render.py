"""Golden vectors for the SIDD driver's sRGB render from the UNMODIFIED reference: `process_sidd_image`
(utils/sidd_utils.py:156-180) is called as it is, with the OpenCV build of this container (opencv-python 4.13.0) behind its
`cv2.COLOR_BayerBG2RGB_EA` demosaic; the raw demosaic of two uint16 mosaics by cv2 itself is stored as well.

    python tests/golden/make_golden_render.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import yond_oracle as O  # noqa: E402
from oracle.ref_harness import load_reference  # noqa: E402

PATTERNS = [[[1, 2], [2, 3]], [[2, 1], [3, 2]], [[2, 3], [1, 2]], [[3, 2], [2, 1]]]


def main():
    ref = load_reference()
    import cv2
    from utils.sidd_utils import process_sidd_image  # the reference's own
    rng = np.random.default_rng(47)
    out = {"cv2_version": cv2.__version__, "seed": 47}
    # (1) the whole render, one case per CFA phase; the last two images leave [0,1] / hold flat and saturated areas
    H, W = 72, 104
    for i, pat in enumerate(PATTERNS):
        clean = O.synth_clean(rng, H, W)
        img = O.synth_noisy(rng, clean, 2.0 + i, 3.0, clip=False).astype(np.float32)
        if i == 2:
            img[10:30, 20:60] = 0.25          # flat: every gradient comparison ties
            img[40:60, 50:90] = 1.5           # saturated
        if i == 3:
            img = (img * 1.8 - 0.1).astype(np.float32)
        wb = np.array([[rng.uniform(1.6, 2.4), 1.0, rng.uniform(1.4, 2.0)]])
        cst = np.array([[0.8, 0.25, -0.05], [-0.3, 1.1, 0.2], [0.02, -0.2, 0.9]]) + rng.normal(0, 0.03, (3, 3))
        out[f"img{i}"], out[f"pat{i}"], out[f"wb{i}"], out[f"cst{i}"] = img, np.array(pat), wb, cst
        out[f"srgb{i}"] = process_sidd_image(img.copy(), pat, wb, cst)
        if i == 0:  # the sRGB numbers of multiprocess_plot (YOND_SIDD.py:660-665): picture of the noisy mosaic vs picture of the clean one
            out["srgb_clean0"] = process_sidd_image(clean.astype(np.float32), pat, wb, cst)
            out["ssim_u8"] = np.float64(ref.Y.calculate_ssim(out["srgb0"], out["srgb_clean0"]))
            halves = [np.split(out[k], 2, axis=-2) for k in ("srgb0", "srgb_clean0")]
            out["ssim_u8_blocks"] = np.array([ref.Y.calculate_ssim(a, b) for a, b in zip(*halves)], np.float64)
    # (2) cv2's demosaic alone: full 14-bit range, and a 2-bit range where ties dominate; odd-ish sizes
    for j, (h, w, hi) in enumerate(((34, 50, 16384), (26, 38, 4))):
        b = rng.integers(0, hi, size=(h, w), dtype=np.uint16)
        out[f"bayer{j}"] = b
        out[f"ea{j}"] = cv2.cvtColor(b, cv2.COLOR_BayerBG2RGB_EA)
    np.savez_compressed(os.path.join(HERE, "render.npz"), **out)
    for i in range(4):
        mine = O.process_sidd_image(out[f"img{i}"], PATTERNS[i], out[f"wb{i}"], out[f"cst{i}"])
        print("pattern", PATTERNS[i], "oracle == reference:", np.array_equal(mine, out[f"srgb{i}"]), "mean", out[f"srgb{i}"].mean())
    for j in range(2):
        print("ea", j, np.array_equal(O.demosaic_ea_u16(out[f"bayer{j}"]), out[f"ea{j}"]))
    print("bytes", os.path.getsize(os.path.join(HERE, "render.npz")))


if __name__ == "__main__":
    main()
