"""Golden vectors for BASELINE configs[3] (ELD / LRID-like 14-bit frames, noclip, low-light gain) — runs the UNMODIFIED
reference (/root/reference) in the build container, like make_golden.py.

The reference's `full_dn` branch (YOND_SIDD.py:387-389, :456-458) denoises the whole frame in one VST_Denoiser call; the
14-bit datasets normalise as (raw - bl) * ratio / (wp - bl) without clipping (data_process/yond_datasets.py:955-961,
1053-1056) and the driver's `p` carries wp = 16383, bl = 512, scale = (wp - bl) / ratio.  The shipped driver
hard-codes SIDD_256 = True in round 2 (:431), which needs the packed width to be a multiple of 32: the frame here is
160 x 2048 (packed 80 x 1024 -> 32 strips of 32 columns).

    python tests/golden/make_golden_c4.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import yond_oracle as O  # noqa: E402
from oracle.ref_harness import load_reference, make_driver  # noqa: E402

ARCH = {"name": "GuidedResUnet", "guided": True, "in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": True, "norm": True}
PIPE = {"data_type": "ELD", "full_est": True, "est_type": "simple+full", "k": 29, "full_dn": True, "vst_type": "exact",
        "bias_corr": "pre", "denoiser_type": "gru32n", "iter": "iter", "max_iter": 1, "clip": False}
LUT_PATH = os.path.join(ROOT, "yond_public_b200", "data", "bias_lut_2d_f32.npz")
H, W = 160, 2048


def frame_for(ratio, seed):
    """Synthetic low-light 14-bit observation in the dataset's normalisation: true signal in DN = clean * 15871 / ratio."""
    rng = np.random.default_rng(seed)
    clean = O.synth_clean_smooth(rng, H, W)
    K_true, S_true = 2.2, 3.1  # DN per electron, read noise in DN (Sony-like)
    return O.synth_noisy(rng, clean, K_true * ratio, S_true * ratio, scale=16383.0 - 512.0, clip=False)


def main():
    ref = load_reference()
    import scipy  # noqa: F401
    blut = ref.utils.BiasLUT(lut_path=_lut_npy())
    out = {}
    for ratio in (1, 100):
        noisy = frame_for(ratio, 4000 + ratio)
        for wname, sd in (("smooth", O.smoother_state_dict(ARCH)), ("mix", _mix())):
            drv = make_driver(ref, ARCH, PIPE, biaslut=blut, seed=5)
            drv.net.load_state_dict(sd)
            p = {"wp": 16383, "bl": 512, "ratio": ratio, "gain": 1, "sigma": 0, "scale": (16383 - 512) / ratio, "cfa": [[1, 2], [2, 3]]}
            res = drv.IterDenoise({"lr_path_full": None, "lr": noisy[None].copy(), "meta": None, "name": "a_b_XX_00100_x"},
                                  {"p": p, "img_id": 0})
            tag = f"r{ratio}_{wname}"
            out[f"{tag}_nrounds"] = len(res["raw_dns"])
            out[f"{tag}_regs"] = np.array([np.asarray(r, np.float64) for r in res["regs"]])
            for i, dn in enumerate(res["raw_dns"]):
                out[f"{tag}_dn{i}_sub"] = dn[::4, ::8].astype(np.float32)
                out[f"{tag}_dn{i}_mean"] = np.float64(dn.astype(np.float64).mean())
            print(tag, out[f"{tag}_nrounds"], out[f"{tag}_regs"])
    np.savez_compressed(os.path.join(HERE, "c4_frame.npz"), H=H, W=W, **out)
    print("wrote c4_frame")


def _mix():
    rnd = O.init_state_dict(ARCH, seed=0)
    sm = O.smoother_state_dict(ARCH)
    return {k: sm[k] + 0.25 * rnd[k] for k in rnd}


def _lut_npy():
    """The reference's BiasLUT takes a .npy path: write the stand-in table next to its scratch directory."""
    import tempfile
    path = os.path.join(tempfile.mkdtemp(prefix="yond_lut_"), "bias_lut_2d.npy")
    np.save(path, np.load(LUT_PATH)["bias_lut"])
    return path


if __name__ == "__main__":
    main()
