"""Generates the stand-in `bias_lut_2d` table with the REFERENCE's own code (build container only).

The authors' checkpoints/bias_lut_2d.npy is not distributed (/root/reference/.MISSING_LARGE_BLOBS).
The table consumed by `BiasLUT` (utils/isp_algos.py:162-177) has shape (1921 x-nodes, 1101 sigma-nodes),
row-major [x, sigma], in electrons (K = 1).  Each column is produced by the reference's
`get_bias_points(x_lut, K=1, sigGs=sg, pho_min=100, close_form=True)` (utils/isp_algos.py:142-160).

Output: yond_public_b200/data/bias_lut_2d_f32.npz (float32, compressed) — consumed by BOTH the oracle and
the CUDA path, so parity does not depend on its content; it only has to be a faithful stand-in.

Usage (≈10 min on 8 cores):  python tests/golden/make_bias_lut.py [nproc]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def _grids():
    sp = 128
    x_lut = np.concatenate((np.linspace(0, 2 ** -4, sp, endpoint=False),
                            np.exp(np.linspace(np.log(2 ** (-4)), np.log(2 ** 10), 14 * sp + 1))))
    sg_lut = np.concatenate((np.linspace(0, 1, 200, endpoint=False), np.linspace(1, 10, 901)))
    return x_lut, sg_lut


def _column(j):
    from oracle.ref_harness import load_reference
    ref = load_reference()
    x_lut, sg_lut = _grids()
    col = ref.utils.get_bias_points(x_lut.copy(), 1.0, float(sg_lut[j]), pho_min=100, close_form=True)
    return j, col


def main():
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else max(1, (os.cpu_count() or 2) - 1)
    x_lut, sg_lut = _grids()
    out = np.zeros((len(x_lut), len(sg_lut)), np.float64)
    t0 = time.time()
    from multiprocessing import Pool
    with Pool(nproc) as pool:
        for n, (j, col) in enumerate(pool.imap_unordered(_column, range(len(sg_lut)), chunksize=4)):
            out[:, j] = col
            if n % 50 == 0:
                print(f"{n}/{len(sg_lut)} columns, {time.time() - t0:.0f}s", flush=True)
    dst = os.path.join(ROOT, "yond_public_b200", "data")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "bias_lut_2d_f32.npz"), bias_lut=out.astype(np.float32))
    print("done", out.shape, f"{time.time() - t0:.0f}s", "nan:", int(np.isnan(out).sum()))


if __name__ == "__main__":
    main()
