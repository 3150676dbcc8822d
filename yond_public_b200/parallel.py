"""Image / tile parallelism across the GPUs of one box (one process per GPU, torch.distributed).

Frames (and the halo tiles of one frame) are independent units: ranks take contiguous shares, run the whole path
locally, and the only exchange is the FINAL GATHER of the denoised output (NCCL over NVLink on GPUs, gloo in the CPU
tests).  There is no data-path collective.  A SINGLE frame is split into row bands for the network stage
(BandShardedForward): one all-gather of the network output per round.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int):
    """Pins this process (and therefore its first-touch host allocations: the pinned staging buffers of the host path) to the CPUs of
    the NUMA node the GPU hangs off.  With one process per GPU on a two-socket box, unpinned ranks put their pinned memory wherever
    they happen to run and half of the H2D / D2H traffic crosses the socket interconnect.  Returns the node, or None when the
    topology is not visible (containers without /sys NUMA information): nothing is changed then."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs a 4-digit one
            bdf = bdf[4:]
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_range(n_units: int, rank: int, world: int):
    """Contiguous, balanced share of `n_units` for `rank`: the first (n_units % world) ranks get one extra unit."""
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_units: int, world: int):
    return [shard_range(n_units, r, world)[1] - shard_range(n_units, r, world)[0] for r in range(world)]


def gather_units(local: torch.Tensor, n_units: int, dst: int = 0):
    """Gathers per-rank outputs (first dim = this rank's units, possibly ragged across ranks) onto `dst` in unit order.
    Returns the (n_units, ...) tensor on `dst`, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(n_units, world)
    assert local.shape[0] == sizes[rank]
    pad = max(sizes)
    buf = local
    if local.shape[0] < pad:  # equal-size exchange; the padding rows are dropped on arrival
        buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[:local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf.contiguous(), outs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0)


def run_sharded(units: torch.Tensor, fn, dst: int = 0):
    """units: (n, ...) available on every rank; each rank applies `fn` to its share; `dst` gets all results."""
    if not dist.is_initialized():
        return fn(units)
    rank, world = dist.get_rank(), dist.get_world_size()
    a, b = shard_range(units.shape[0], rank, world)
    return gather_units(fn(units[a:b]), units.shape[0], dst)


def band_range(hp: int, rank: int, world: int, align: int = 16):
    """Rows [r0, r1) of a padded frame of `hp` rows owned by `rank`: equal bands of a multiple of `align` rows (the network
    has four stride-2 stages: tile origins on multiples of 16); trailing ranks may get a short or empty band.  Returns
    (r0, r1, band) with band = the common band height."""
    band = -(-hp // (world * align)) * align
    r0 = min(hp, rank * band)
    return r0, min(hp, r0 + band), band


class BandShardedForward:
    """Network forward of ONE padded frame split across the ranks into row bands (tile-sharded single frame, BASELINE
    configs[2]/[3]).  A rank forwards its band plus a halo of `halo` rows that is cut at the frame border — the halo covers
    the receptive field (107 / 123 packed pixels), full-width bands need no left / right halo and the slice is a
    contiguous view of the NHWC frame, so there is no tile copy at all — and the bands are exchanged with ONE all-gather
    (NCCL over NVLink; gloo in the CPU test).  Everything else of the pipeline (estimate, VST, inverse) is cheap and HBM
    bound and is replicated on every rank instead of exchanged, so every rank ends with the whole result."""

    def __init__(self, forward, rank=None, world=None, halo=128, group=None):
        self.forward = forward
        live = dist.is_initialized()
        self.rank = (dist.get_rank(group) if live else 0) if rank is None else rank
        self.world = (dist.get_world_size(group) if live else 1) if world is None else world
        self.halo, self.group = halo, group
        self._buf = {}

    def __call__(self, z, ub, t=None, out=None):
        B, hp, wp, c = z.shape
        assert B == 1, "band sharding splits one frame; batches of frames are image-parallel"
        r0, r1, band = band_range(hp, self.rank, self.world)
        key = (band, wp, c, z.device, z.dtype)
        if key not in self._buf:
            self._buf = {key: (torch.zeros((band, wp, c), device=z.device, dtype=z.dtype),
                               torch.empty((self.world * band, wp, c), device=z.device, dtype=z.dtype))}
        mine, full = self._buf[key]
        if r1 > r0:
            ty0, ty1 = max(0, r0 - self.halo), min(hp, r1 + self.halo)
            yt = self.forward(z[:, ty0:ty1], ub, t)
            mine[:r1 - r0].copy_(yt[0, r0 - ty0:r1 - ty0])
        if self.world > 1:
            dist.all_gather_into_tensor(full.view(-1), mine.view(-1), group=self.group)
        else:
            full = mine
        y = torch.empty_like(z) if out is None else out
        y[0].copy_(full[:hp])
        return y


def denoise_frame_sharded(drv, frame, p):
    """IterDenoise of ONE full-resolution frame (H,W) with the network stage band-sharded across the ranks of the default
    process group; returns the device result dict of YOND_SIDD.iter_denoise_dev on every rank."""
    eng = drv.engine
    prev = eng.forward
    eng.forward = BandShardedForward(drv.net.forward_nhwc)
    try:
        return drv.iter_denoise_dev(frame.reshape(1, 1, *frame.shape[-2:]), p)
    finally:
        eng.forward = prev
