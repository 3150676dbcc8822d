"""One self + one collab estimate on N 12 MP frames between cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yond_public_b200 as Y  # noqa: E402,F401
from yond_public_b200 import nlf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, W = 3024, 4032
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand((n, 1, H, W), device="cuda", generator=g) * 0.5
y = (x.reshape(n, H, W) * 0.9 + 0.01).contiguous()
est = nlf._estimator()
seg = torch.zeros(n, device="cuda")
for _ in range(2):
    est.estimate_dev(x, None, 29, seg_max=seg)
    est.estimate_dev(x, y, 29, y_mosaic=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
est.estimate_dev(x, None, 29, seg_max=seg)
est.estimate_dev(x, y, 29, y_mosaic=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled ok")
