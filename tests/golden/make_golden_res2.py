"""Golden vectors for the ResUnet2 plugin (archs/Unet.py:197-286) — runs the UNMODIFIED reference
(/root/reference) in the build container, like make_golden.py: state-dict identity under the random-init recipe and one forward.

    python tests/golden/make_golden_res2.py
"""
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import yond_oracle as O  # noqa: E402
from oracle.ref_harness import load_reference, make_driver  # noqa: E402

ARCH = {"name": "ResUnet2", "in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": True, "norm": True}
ARCH_GSELF = {"name": "GuidedSelfUnet", "guided": True, "in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": False, "norm": True}  # :852-910
ARCH_SELF = {"name": "SelfResUNet", "in_nc": 4, "out_nc": 4, "nf": 32, "nframes": 1, "res": True, "norm": True}  # archs/comp.py:745-802
PIPE = {"full_est": True, "est_type": "simple+full", "k": 29, "full_dn": False, "vst_type": "exact", "bias_corr": "pre",
        "iter": "iter", "max_iter": 1}


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def main():
    ref = load_reference()
    for arch, fname in ((ARCH, "net_res2"), (ARCH_SELF, "net_selfres"), (ARCH_GSELF, "net_gself")):
        one(ref, arch, fname)


def one(ref, ARCH, fname):
    drv = make_driver(ref, ARCH, PIPE, seed=5)
    sd_ref = drv.net.state_dict()
    sd = O.init_state_dict(ARCH, seed=5)
    assert list(sd_ref.keys()) == list(sd.keys()), "state-dict keys / order differ"
    for k in sd:
        assert torch.equal(sd_ref[k], sd[k]), k
    rng = np.random.default_rng(2024)
    xin = torch.from_numpy(rng.uniform(0, 1, (2, 4, 64, 32)).astype(np.float32))
    xin[1] *= 0.6
    with torch.no_grad():
        y = drv.net(xin, torch.tensor(0.043, dtype=torch.float32)) if "guided" in ARCH else drv.net(xin)
    np.savez_compressed(os.path.join(HERE, fname + ".npz"), x=xin.numpy(), y=y.numpy(), nparams=sum(v.numel() for v in sd.values()),
                        keys=np.array(list(sd.keys())), shapes=np.array([str(tuple(v.shape)) for v in sd.values()]),
                        sd_crc=np.array([crc(v.numpy()) for v in sd.values()], np.uint32))
    print("wrote", fname, float(y.abs().max()))


if __name__ == "__main__":
    main()
