"""Times yond_pack_raw on 24 MP 14-bit frames: algorithmic bytes = 2 B/px in + 4 B/px out."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yond_public_b200 as Y  # noqa: E402

B, H, W = 8, 4000, 6000
rng = np.random.default_rng(0)
raw = torch.from_numpy(rng.integers(0, 2 ** 14, size=(B, H, W), dtype=np.uint16).view(np.int16)).cuda()
pat, black = [[0, 1], [3, 2]], [512, 512, 512, 512]
for inter in (False, True):
    for _ in range(3):
        Y.pack_raw_bayer(raw, wp=16383, clip=True, raw_pattern=pat, black_level_per_channel=black, interleaved=inter)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        Y.pack_raw_bayer(raw, wp=16383, clip=True, raw_pattern=pat, black_level_per_channel=black, interleaved=inter)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = B * H * W * 6 / 1e9
    print(f"pack_raw {'interleaved' if inter else 'planes'}: {ms:.3f} ms for {B} x 24 MP = {gb / ms * 1e3:.0f} GB/s algorithmic "
          f"({B * H * W / 1e6 / ms * 1e3 / 1e3:.1f} GPixel/s)")
