"""torchrun --nproc-per-node 2 tools/multi_gpu_check.py — multi-GPU paths on real GPUs (NCCL):
(1) one 12 MP frame tile-sharded across the ranks == the single-GPU result; (2) image-parallel gather."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yond_public_b200 as Y  # noqa: E402
from yond_public_b200 import synth  # noqa: E402
from yond_public_b200 import parallel  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
arch = {"name": "GuidedResUnet", "guided": True, "in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": True, "norm": True}
net = Y.build_net(arch, dev)
net.load_state_dict(synth.random_init_state_dict(arch, seed=5))
eng = Y.YondEngine(net, arch, Y.BiasLUT())
rng = np.random.default_rng(7)  # same frame on every rank
frame = torch.from_numpy(synth.noisy(rng, synth.clean_smooth(rng, 3024, 4032), 3.0, 5.0)).to(dev)
out = parallel.denoise_frame_tile_sharded(eng, frame, 3.1, 5.2, 959.0, core=512)
if rank == 0:
    ref = eng.vst_denoise_tiled(frame, 3.1, 5.2, 959.0, core=512)
    print("tile-sharded == single-GPU tiled: max abs diff", float((out - ref).abs().max()))
    assert torch.equal(out, ref)
units = torch.arange(10, dtype=torch.float32, device=dev).reshape(10, 1, 1).expand(10, 4, 4).contiguous()
got = parallel.run_sharded(units, lambda u: u * 2, dst=0)
if rank == 0:
    assert torch.equal(got, units * 2)
    print("image-parallel gather ok")
dist.destroy_process_group()
