"""Noise-level-function estimators (SimpleNLF / SelfNLF / CollabNLF, YOND_SIDD.py:62-124) on device.

Device work: box statistics (yond_nlf_maps), exact order statistics (yond_order_stats), score3 bin occupancy
(yond_score3_bins), masked regression sums (yond_masked_sums).  Host work (a few dozen scalars, float64,
exactly like the reference): np.percentile's linear interpolation, the score argmin, the reference's guards
and the 2x2 least-squares solve.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .isp import _as_batch4, _dev, bayer2rggb, to_dev


def _percentiles_from_order_stats(lo_vals, hi_vals, gamma):
    """np.percentile(method='linear') given the two bracketing order statistics (float32) and the fractional
    part gamma (float64): NumPy's _lerp — the difference is taken in float32, the lerp in float64."""
    a = lo_vals.astype(np.float32)
    b = hi_vals.astype(np.float32)
    diff = (b - a)  # float32, like np.subtract on the float32 array
    out = a + diff * gamma
    hi_side = gamma >= 0.5
    out[hi_side] = (b - diff * (1 - gamma))[hi_side]
    return np.asarray(out, dtype=np.float64)


class NlfEstimator:
    """Reusable device scratch for the estimator."""

    def __init__(self):
        self.lib = _lib.load()
        self._sel_work = None
        self._small = None

    def _scratch(self, dev):
        if self._sel_work is None or self._sel_work.device != dev:
            self._sel_work = torch.empty(self.lib.yond_select_work_bytes(64), device=dev, dtype=torch.uint8)
            self._small = torch.empty(8192, device=dev, dtype=torch.uint8)
        return self._sel_work, self._small

    # -- maps -----------------------------------------------------------------------------------
    def maps(self, lr_rggb, hr_rggb=None, k=29):
        """lr_rggb / hr_rggb: (B,h,w,4) CUDA float32.  Returns var, mean, lap (same shape)."""
        B, h, w, _ = lr_rggb.shape
        work = torch.empty(self.lib.yond_nlf_work_bytes(B, h, w, 4), device=lr_rggb.device, dtype=torch.uint8)
        var, mean, lap = torch.empty_like(lr_rggb), torch.empty_like(lr_rggb), torch.empty_like(lr_rggb)
        mode = 0 if hr_rggb is None else 1
        check(self.lib.yond_nlf_maps(ptr(lr_rggb), ptr(hr_rggb), ptr(var), ptr(mean), ptr(lap), B, h, w, 4, int(k), mode,
                                     ptr(work), stream_ptr()))
        return var, mean, lap

    # -- percentiles ----------------------------------------------------------------------------
    def percentiles(self, data, quants):
        """np.percentile(data.reshape(-1), quants, method='linear') with exact device order statistics."""
        n = data.numel()
        q = np.asarray(quants, np.float64)
        qq = np.true_divide(q, 100)
        vi = (n - 1) * qq  # NumPy's virtual index for method='linear'
        lo = np.floor(vi).astype(np.int64)
        hi = np.minimum(lo + 1, n - 1)
        gamma = vi - lo
        ranks = np.concatenate([lo, hi]).astype(np.uint64)
        sel_work, _ = self._scratch(data.device)
        ranks_dev = torch.from_numpy(ranks.view(np.int64)).to(data.device)
        out = torch.empty(len(ranks), device=data.device, dtype=torch.float32)
        check(self.lib.yond_order_stats(ptr(data), n, ptr(ranks_dev), len(ranks), ptr(out), ptr(sel_work), stream_ptr()))
        vals = out.cpu().numpy()
        return _percentiles_from_order_stats(vals[:len(q)], vals[len(q):], gamma)

    # -- get_threshold(mode='score3'), YOND_SIDD.py:22-49 ----------------------------------------
    def threshold_score3(self, lap, mean, step=5):
        quants = np.linspace(step, 100, 100 // step, endpoint=True)
        ths = self.percentiles(lap, quants)
        _, small = self._scratch(lap.device)
        ths_dev = torch.from_numpy(ths).to(lap.device)
        npk = torch.empty(len(ths), device=lap.device, dtype=torch.int32)
        check(self.lib.yond_score3_bins(ptr(lap), ptr(mean), lap.numel(), ptr(ths_dev), len(ths), ptr(npk), ptr(small),
                                        stream_ptr()))
        npeaks = npk.cpu().numpy().astype(np.float64)
        score = ths / (quants * npeaks)
        i = int(np.argmin(score[1:]) + 1)  # start_pos = 1 skips the 5 % quantile (:46-47)
        return ths[i], quants[i], dict(ths=ths, npeaks=npeaks, score=score)

    # -- masked fit, YOND_SIDD.py:77-86 + utils/isp_algos.py:345-365 ------------------------------
    def masked_fit(self, var, mean, lap, th):
        sums_dev = torch.empty(12, device=lap.device, dtype=torch.float64)

        def sums(thr):
            check(self.lib.yond_masked_sums(ptr(lap), ptr(mean), ptr(var), lap.numel(), float(thr), ptr(sums_dev), stream_ptr()))
            return sums_dev.cpu().numpy()
        s = sums(th)
        if s[0] == 0:  # "no flat area": fall back to the 25th percentile (:79-84)
            th_backup = float(self.percentiles(lap, [25.0])[0])
            if th != th_backup:
                th = th_backup
                s = sums(th)
        use = s[6:12] if s[6] > 0.01 * s[0] else s[0:6]  # polyfit keeps 1e-4 < x < 0.8 when that is > 1 % (:348-350)
        N, Sx, Sy, Sxx, Sxy = use[0], use[1], use[2], use[3], use[4]
        det = N * Sxx - Sx * Sx
        b1 = (N * Sxy - Sx * Sy) / det
        b2 = (Sy - b1 * Sx) / N
        return np.array([b1, b2], np.float64), th

    # -- SelfNLF / CollabNLF ----------------------------------------------------------------------
    def estimate(self, lr_rggb, hr_rggb=None, k=29, details=False):
        var, mean, lap = self.maps(lr_rggb, hr_rggb, k)
        th, pct, info = self.threshold_score3(lap, mean, step=5)
        reg, th = self.masked_fit(var, mean, lap, th)
        if details:
            return reg, dict(th=th, pct=pct, **info)
        return reg


_EST = None


def _estimator():
    global _EST
    if _EST is None:
        _EST = NlfEstimator()
    return _EST


def _rggb_batch(raw, sidd_256):
    """Bayer (H,W) [or (B,H,W)] -> packed (B,h,w,4) device batch.  SIDD_256: the mosaic's 32 blocks become 32
    separate images, which is what stacking them on the channel axis means for a box filter (YOND_SIDD.py:65,91-93)."""
    t, _ = to_dev(raw)
    if t.dim() == 2:
        t = t[None]
    if sidd_256:
        B, H, W = t.shape
        assert B == 1 and W % 32 == 0
        t = t.reshape(H, 32, W // 32).permute(1, 0, 2).contiguous()
    return bayer2rggb(t)


def SelfNLF(lr_rggb, k=29, kwargs=None):
    kwargs = kwargs or {}
    t, _ = to_dev(lr_rggb)
    b = _as_batch4(t) if t.dim() == 3 else t
    if kwargs.get("SIDD_256"):
        B, h, w, _ = b.shape
        b = b.reshape(h, 32, w // 32, 4).permute(1, 0, 2, 3).contiguous()
    return _estimator().estimate(b, None, k)


def CollabNLF(lr_rggb, hr_rggb, k=29, kwargs=None):
    kwargs = kwargs or {}
    tl, _ = to_dev(lr_rggb)
    th_, _ = to_dev(hr_rggb)
    bl = _as_batch4(tl) if tl.dim() == 3 else tl
    bh = _as_batch4(th_) if th_.dim() == 3 else th_
    if kwargs.get("SIDD_256"):
        B, h, w, _ = bl.shape
        bl = bl.reshape(h, 32, w // 32, 4).permute(1, 0, 2, 3).contiguous()
        bh = bh.reshape(h, 32, w // 32, 4).permute(1, 0, 2, 3).contiguous()
    return _estimator().estimate(bl, bh, k)


def SimpleNLF(lr_raw, hr_raw=None, k=29, setting=None):
    """YOND_SIDD.py:117-124 — pack + dispatch.  Returns array-like (beta1, beta2), float64."""
    setting = setting or {"mode": "self"}
    sidd = bool(setting.get("SIDD_256", False))
    lr = _rggb_batch(lr_raw, sidd)
    if setting["mode"] == "self":
        return _estimator().estimate(lr, None, k)
    if setting["mode"] == "collab":
        return _estimator().estimate(lr, _rggb_batch(hr_raw, sidd), k)
    raise NotImplementedError(setting["mode"])
