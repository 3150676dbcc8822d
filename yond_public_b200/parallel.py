"""Image / tile parallelism across the GPUs of one box (one process per GPU, torch.distributed).

Frames (and the halo tiles of one frame) are independent units: ranks take contiguous shares, run the whole path
locally, and the only exchange is the FINAL GATHER of the denoised output (NCCL over NVLink on GPUs, gloo in the CPU
tests).  There is no data-path collective.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


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


def gather_disjoint(partial: torch.Tensor, dst: int = 0):
    """Final gather for tile-sharded frames: every rank holds the full-size padded output with ITS tile cores filled
    and zeros elsewhere; the supports are disjoint, so a SUM reduction onto `dst` assembles the frame exactly
    (x + 0 is exact in floating point).  One collective of frame size (48.8 MB for a 12 MP frame)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return partial
    dist.reduce(partial, dst=dst, op=dist.ReduceOp.SUM)
    return partial if dist.get_rank() == dst else None


def denoise_frame_tile_sharded(engine, frame, gain, sigma, scale, bias_corr="pre", vst_type="exact", clip01=True,
                               core=512, dst=0):
    """One full-resolution frame across the ranks: every rank runs the cheap HBM-bound pre-stage (pack + bias + VST +
    normalise + pad + global max) on the whole frame — replicated rather than exchanged — denoises its
    share of the halo tiles, and the cores are gathered on `dst`, which applies the inverse VST."""
    import numpy as np  # noqa: F401
    from . import isp
    from ._lib import check, ptr, stream_ptr
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    H, W = frame.shape
    h, w = H // 2, W // 2
    pl, pr, pt, pb = isp.get_p2d((1, 4, h, w), base=32)
    ntiles = len(engine.tile_grid(h + pt + pb, w + pl + pr, core))
    a, b = shard_range(ntiles, rank, world)
    y, (params, p2d) = engine.vst_denoise_tiled(frame, gain, sigma, scale, bias_corr, vst_type, clip01, core=core,
                                                 tiles=range(a, b), return_padded=True)
    y = gather_disjoint(y, dst)
    if y is None:
        return None
    out = torch.empty_like(frame)
    check(engine.lib.yond_vst_inv(ptr(y), ptr(out), 1, H, W, *p2d, ptr(params), int(clip01), stream_ptr()))
    return out
