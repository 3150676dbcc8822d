"""Golden vectors for SURVEY 8(f)-1 (RAW ingest): runs the UNMODIFIED reference `pack_raw_bayer`
(/root/reference/data_process/process.py:40-64) on stand-in rawpy objects, and `rot_bayer` (utils/sidd_utils.py:198-213), and stores inputs + outputs in
tests/golden/pack_raw.npz.

    python tests/golden/make_golden_raw.py        # needs /root/reference
"""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_harness import load_reference  # noqa: E402

load_reference()
P = importlib.import_module("data_process.process")

rng = np.random.default_rng(11)
cases = {}
patterns = {"rggb": [[0, 1], [3, 2]], "bggr": [[2, 3], [1, 0]], "grbg": [[1, 0], [2, 3]], "gbrg": [[3, 2], [0, 1]]}
specs = [("rggb", 10, 1023, [64, 64, 64, 64], True), ("bggr", 14, 16383, [512, 512, 512, 512], False),
         ("grbg", 12, 4095, [255, 257, 256, 254], True), ("gbrg", 14, 15360, [2047, 2048, 2049, 2050], False)]
for i, (name, bits, wp, black, clip) in enumerate(specs):
    H, W = 24 + 2 * i, 36 + 4 * i
    img = rng.integers(0, 2 ** bits, size=(H, W), dtype=np.uint16)
    img[0, :4] = [0, 1, 2 ** bits - 1, black[0]]  # extremes: below black, saturated, exactly black
    raw = types.SimpleNamespace(raw_image_visible=img, raw_pattern=np.array(patterns[name], dtype=np.uint8),
                                black_level_per_channel=list(black))
    out = P.pack_raw_bayer(raw, wp=wp, clip=clip)
    assert out.dtype == np.float32 and out.shape == (4, H // 2, W // 2)
    cases.update({f"img{i}": img, f"pattern{i}": np.array(patterns[name], dtype=np.int32), f"black{i}": np.array(black, dtype=np.int32),
                  f"wp{i}": np.int32(wp), f"clip{i}": np.int32(clip), f"out{i}": out})
# rot_bayer (utils/sidd_utils.py:198-213) through the reference: every pattern, forward and reverse, 2-D and batched
S = importlib.import_module("utils.sidd_utils")
rot_in = rng.random((2, 18, 26)).astype(np.float32)
cases["rot_in"] = rot_in
for pi, pat in enumerate([[[1, 2], [2, 3]], [[2, 1], [3, 2]], [[2, 3], [1, 2]], [[3, 2], [2, 1]]]):
    cases[f"rot_pat{pi}"] = np.array(pat, dtype=np.int32)
    for rev in (0, 1):
        cases[f"rot_out{pi}_{rev}"] = np.ascontiguousarray(S.rot_bayer(rot_in, pat, rev=bool(rev)))
        cases[f"rot2d_out{pi}_{rev}"] = np.ascontiguousarray(S.rot_bayer(rot_in[0], pat, rev=bool(rev)))
cases["n"] = np.int32(len(specs))
np.savez_compressed(os.path.join(HERE, "pack_raw.npz"), **cases)
print("wrote pack_raw.npz:", len(specs), "cases")
