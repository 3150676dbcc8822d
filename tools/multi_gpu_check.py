"""torchrun --nproc-per-node N tools/multi_gpu_check.py — multi-GPU paths on real GPUs (NCCL):
(1) one 12 MP frame (C3) and one 24 MP 14-bit frame (C4) with the network stage band-sharded across the ranks == the
single-GPU whole-frame result; (2) image-parallel gather."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import yond_public_b200 as Y  # noqa: E402
from yond_public_b200 import parallel, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
drv = Y.YOND_SIDD(bench.ARCH, bench.PIPE_FRAME, state_dict=synth.bench_state_dict(bench.ARCH, seed=0), device=dev)
for name, (H, W), p in (("C3 12 MP", (3024, 4032), bench.P0), ("C4 24 MP 14-bit x100", (4000, 6000), bench.P_C4)):
    rng = np.random.default_rng(7)  # same frame on every rank
    frame = torch.from_numpy(bench.synth_frame(rng, H, W, p)).to(dev)
    res = parallel.denoise_frame_sharded(drv, frame, dict(p))
    ref = drv.iter_denoise_dev(frame.reshape(1, 1, H, W), dict(p))
    d1 = float((res["dn1"] - ref["dn1"]).abs().max())
    d2 = float((res["final"] - ref["final"]).abs().max())
    r = (res["regs2"] - ref["regs2"]).abs().max().item()
    if rank == 0:
        print(f"{name}: band-sharded x{world} vs single GPU: round-1 max abs diff {d1:.3g}, final {d2:.3g}, regs2 diff {r:.3g}, ok {res['ok'].tolist()}")
    assert d1 < 1e-5 and d2 < 1e-4
units = torch.arange(10, dtype=torch.float32, device=dev).reshape(10, 1, 1).expand(10, 4, 4).contiguous()
got = parallel.run_sharded(units, lambda u: u * 2, dst=0)
if rank == 0:
    assert torch.equal(got, units * 2)
    print("image-parallel gather ok")
dist.destroy_process_group()
