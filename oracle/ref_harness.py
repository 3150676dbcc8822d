"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference (/root/reference) under stub modules.

Only the golden-vector generators under tests/golden/ (run in the build container, where
/root/reference exists) may use this file.  Nothing shipped, benchmarked or run on the GPU
box imports it: /root/reference does not exist there.

The reference imports a dozen packages the hot path never touches (matplotlib, skimage,
rawpy, h5py, kornia, bm3d ...).  They are absent from this image, so empty stand-in modules
are installed in sys.modules before the import.
"""
import importlib.machinery
import importlib.util
import os
import sys
import tempfile
import types

REF_ROOT = os.environ.get("YOND_REFERENCE_ROOT", "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_LOADED = {}


def load_reference():
    """Returns a namespace with the reference modules: .utils, .archs, .Y (YOND_SIDD.py as a module)."""
    if _LOADED:
        return types.SimpleNamespace(**_LOADED)
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    _stub("skimage")
    _stub("skimage.metrics", peak_signal_noise_ratio=None, structural_similarity=None)
    for n in ["exifread", "rawpy", "rawpy.enhance", "h5py", "lpips", "torchsummary", "kornia", "kornia.filters"]:
        _stub(n)
    _stub("bm3d", bm3d=None)
    _stub("natsort", natsort=None)
    sys.path.insert(0, REF_ROOT)
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="yond_ref_")
    os.chdir(scratch)  # the driver script creates ./logs etc. relative to cwd
    try:
        import utils as ref_utils  # noqa
        import archs as ref_archs  # noqa
        spec = importlib.util.spec_from_file_location("YOND_SIDD_ref", os.path.join(REF_ROOT, "YOND_SIDD.py"))
        Y = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(Y)
    finally:
        os.chdir(cwd)
    Y.dataload = lambda p: None  # YOND_SIDD.py:339 calls dataload(None) before its None check
    _LOADED.update(utils=ref_utils, archs=ref_archs, Y=Y, scratch=scratch)
    return types.SimpleNamespace(**_LOADED)


def make_driver(ref, arch, pipe, biaslut=None, seed=0, weight_scale=None):
    """Builds a YOND_SIDD object without argparse / dataset loading."""
    import torch
    Y = ref.Y
    o = object.__new__(Y.YOND_SIDD)
    o.device = torch.device("cpu")
    o.arch = dict(arch)
    o.pipe = dict(pipe)
    o.biaslut = biaslut
    o.args = {}
    o.est_args = {}
    o.logfile = None
    o.dst = {"root_dir": ""}
    torch.manual_seed(seed)
    o.net = getattr(Y, arch["name"])(arch)
    Y.initialize_weights(o.net)
    if weight_scale is not None:
        with torch.no_grad():
            for p in o.net.parameters():
                p.mul_(weight_scale)
    o.net.eval()
    return o
