"""Device-side `EvaluationMetrics`: the torch half of the reference class
/root/reference/synth_sod/src/synth_sod/model_training/metrics.py:213-421 - MAE, MaxF / AvgF over the 255-threshold
precision-recall sweep, and the S-measure - with the reductions done by the CUDA library in two passes over the image
instead of the reference's ~770 small torch kernels per image.

Same surface: `EvaluationMetrics(device, sm_only=False)`, `step(pred, mask)`, `compute_metrics()`, `reset()`.
Differences, on purpose: `step` does not binarise the caller's `mask` in place (the reference does, metrics.py:269-270);
the E-measure (metrics.py:14-137: its cumulative-histogram form) is evaluated from two 256-bin histograms the same
device pass produces instead of numpy on a host copy; the weighted F-measure (scipy distance transform + 7x7 Gaussian on
the CPU in the reference, metrics.py:140-210) runs on the device too (`s3od_metrics_weighted_f`: exact separable Euclidean
feature transform with scipy's tie-breaking, double-accumulated 7x7 filter, weighted sums), so `compute_metrics()` returns MAE,
MaxF, AvgF, Sm, Em and wF like the reference (or Sm alone with `sm_only=True`).  The ground truth of the E-measure is `mask >= 0.5`, which for masks in
[0, 1] is what the reference's in-place binarisation followed by `gt > 0` amounts to.  Values agree with the reference to float32 rounding (sums are accumulated in
double here, in float32 there).  There is no CPU fallback.
"""
import ctypes
from typing import Dict, List

import numpy as np
import torch

from .engine import _check, _stream_ptr, load_library


class _Stats(ctypes.Structure):
    _fields_ = [("abs_err", ctypes.c_double), ("sum_p", ctypes.c_double), ("sum_y", ctypes.c_double),
                ("fg_p", ctypes.c_double), ("fg_p2", ctypes.c_double), ("bg_q", ctypes.c_double), ("bg_q2", ctypes.c_double),
                ("n_fg", ctypes.c_uint64), ("sum_mx", ctypes.c_uint64), ("sum_my", ctypes.c_uint64),
                ("hist_cnt", ctypes.c_uint64 * 256), ("hist_y", ctypes.c_double * 256),
                ("em_all", ctypes.c_uint64 * 256), ("em_fg", ctypes.c_uint64 * 256)]


class _WfmSums(ctypes.Structure):
    _fields_ = [("fg_ew", ctypes.c_double), ("bg_ew", ctypes.c_double), ("n_fg", ctypes.c_uint64)]


class _Region(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double * 4) for n in ("sp", "sm", "spp", "smm", "spm")]


def _round_half_even(v: float) -> int:
    return int(np.round(v))                                           # torch.round semantics (metrics.py:376-377)


class EvaluationMetrics:
    def __init__(self, device=None, sm_only: bool = False):
        self.device = torch.device(device if device else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("s3od_b200.metrics.EvaluationMetrics runs on CUDA devices only; there is no CPU fallback")
        self.sm_only = sm_only
        self.metrics: Dict[str, List[float]] = {"mae": [], "max_f": [], "avg_f": [], "s_score": []}
        self.changeable_ems: List[np.ndarray] = []                     # EMeasure.metrics['changeable_ems'] (metrics.py:18-21)
        self.weighted_fms: List[float] = []                            # WeightedFMeasure.metrics['weighted_fms'] (metrics.py:143-145)
        self._lib = load_library()
        assert self._lib.s3od_metrics_stats_bytes() == ctypes.sizeof(_Stats)
        assert self._lib.s3od_metrics_region_bytes() == ctypes.sizeof(_Region)
        self._th = torch.linspace(0, 1 - 1e-10, 255).to(self.device)  # metrics.py:319 (float32, same generator call)
        self._d_stats = torch.empty(ctypes.sizeof(_Stats), dtype=torch.uint8, device=self.device)
        self._d_region = torch.empty(ctypes.sizeof(_Region), dtype=torch.uint8, device=self.device)
        self._lib.s3od_metrics_weighted_f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                      ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        self._lib.s3od_metrics_weighted_f_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
        self._lib.s3od_metrics_weighted_f_workspace_bytes.restype = ctypes.c_size_t
        self._d_wfm = torch.empty(ctypes.sizeof(_WfmSums), dtype=torch.uint8, device=self.device)
        self._wfm_ws = None

    # ---- device passes ---------------------------------------------------------------------------------------------
    def _stats(self, pred: torch.Tensor, mask: torch.Tensor) -> _Stats:
        h, w = pred.shape
        with torch.cuda.device(self.device):
            _check(self._lib, self._lib.s3od_metrics_stats(pred.data_ptr(), mask.data_ptr(), h, w, self._th.data_ptr(),
                                                          self._d_stats.data_ptr(), self._d_stats.numel(), _stream_ptr(self.device)),
                   "s3od_metrics_stats")
        return _Stats.from_buffer_copy(self._d_stats.cpu().numpy().tobytes())

    def _region(self, pred: torch.Tensor, mask: torch.Tensor, X: int, Y: int) -> _Region:
        h, w = pred.shape
        with torch.cuda.device(self.device):
            _check(self._lib, self._lib.s3od_metrics_region(pred.data_ptr(), mask.data_ptr(), h, w, X, Y, self._d_region.data_ptr(),
                                                           self._d_region.numel(), _stream_ptr(self.device)), "s3od_metrics_region")
        return _Region.from_buffer_copy(self._d_region.cpu().numpy().tobytes())

    def _weighted_f(self, pred: torch.Tensor, mask: torch.Tensor, beta: float = 1.0) -> float:
        """WeightedFMeasure.step (metrics.py:146-190) with the per-pixel work on the device."""
        h, w = pred.shape
        need = self._lib.s3od_metrics_weighted_f_workspace_bytes(h, w)
        if self._wfm_ws is None or self._wfm_ws.numel() < need:
            self._wfm_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _check(self._lib, self._lib.s3od_metrics_weighted_f(pred.data_ptr(), mask.data_ptr(), h, w, self._wfm_ws.data_ptr(), self._wfm_ws.numel(),
                                                               self._d_wfm.data_ptr(), _stream_ptr(self.device)), "s3od_metrics_weighted_f")
        sm = _WfmSums.from_buffer_copy(self._d_wfm.cpu().numpy().tobytes())
        if sm.n_fg == 0:                                               # metrics.py:148-150
            return 0.0
        eps = float(np.spacing(1))
        tpw, fpw = sm.n_fg - sm.fg_ew, sm.bg_ew
        r = 1.0 - sm.fg_ew / sm.n_fg
        p = tpw / (tpw + fpw + eps)
        return float((1 + beta) * r * p / (r + beta * p + eps))

    # ---- formulas of the reference on the reduced quantities ---------------------------------------------------------
    @staticmethod
    def _object(s: float, s2: float, n: int) -> float:                # metrics.py:338-344 (mean, unbiased std)
        if n == 0:
            return float("nan")
        x = s / n
        var = (s2 - n * x * x) / (n - 1) if n > 1 else float("nan")
        sigma = float(np.sqrt(max(var, 0.0))) if var == var else float("nan")
        return 2.0 * x / (x * x + 1.0 + sigma + 1e-20)

    @staticmethod
    def _ssim(n: int, sp: float, sm: float, spp: float, smm: float, spm: float) -> float:       # metrics.py:405-421
        if n == 0:
            return float("nan")
        x, y = sp / n, sm / n
        den = n - 1 + 1e-20
        sx, sy, sxy = (spp - n * x * x) / den, (smm - n * y * y) / den, (spm - n * x * y) / den
        alpha = 4 * x * y * sxy
        beta = (x * x + y * y) * (sx + sy)
        if alpha != 0:
            return alpha / (beta + 1e-20)
        return 1.0 if beta == 0 else 0.0

    def _s_measure(self, st: _Stats, pred, mask, h: int, w: int) -> float:      # metrics.py:259-273
        n = h * w
        y_mean = st.sum_y / n
        if y_mean == 0:
            return 1.0 - st.sum_p / n
        if y_mean == 1:
            return st.sum_p / n
        n_fg, n_bg = int(st.n_fg), n - int(st.n_fg)
        u = n_fg / n
        s_obj = u * self._object(st.fg_p, st.fg_p2, n_fg) + (1 - u) * self._object(st.bg_q, st.bg_q2, n_bg)
        if n_fg == 0:                                                  # mask.sum() == 0 after binarisation (metrics.py:361-367)
            X, Y = round(w / 2), round(h / 2)
        else:
            X, Y = _round_half_even(st.sum_mx / n_fg), _round_half_even(st.sum_my / n_fg)
        rg = self._region(pred, mask, X, Y)
        area = h * w
        w1, w2, w3 = X * Y / area, (w - X) * Y / area, X * (h - Y) / area
        w4 = 1 - w1 - w2 - w3
        sizes = (Y * X, Y * (w - X), (h - Y) * X, (h - Y) * (w - X))
        q = [self._ssim(sizes[k], rg.sp[k], rg.sm[k], rg.spp[k], rg.smm[k], rg.spm[k]) for k in range(4)]
        Q = 0.5 * s_obj + 0.5 * (w1 * q[0] + w2 * q[1] + w3 * q[2] + w4 * q[3])
        return max(Q, 0.0) if Q == Q else Q

    @staticmethod
    def _changeable_em(st: _Stats, gt_size: int) -> np.ndarray:
        """EMeasure.cal_em_with_cumsumhistogram + generate_parts_numel_combinations (metrics.py:80-132) on the device
        histograms; same float64 numpy arithmetic, the counts are exact integers."""
        eps = np.spacing(1)
        all_hist = np.array(st.em_all, dtype=np.int64)
        fg_fg_hist = np.array(st.em_fg, dtype=np.int64)
        fg_bg_hist = all_hist - fg_fg_hist
        gt_fg_numel = int(fg_fg_hist.sum())
        fg_fg = np.cumsum(np.flip(fg_fg_hist), axis=0)
        fg_bg = np.cumsum(np.flip(fg_bg_hist), axis=0)
        fg__ = fg_fg + fg_bg
        bg__ = gt_size - fg__
        if gt_fg_numel == 0:
            enhanced = bg__
        elif gt_fg_numel == gt_size:
            enhanced = fg__
        else:
            bg_fg = gt_fg_numel - fg_fg
            bg_bg = bg__ - bg_fg
            parts = [fg_fg, fg_bg, bg_fg, bg_bg]
            mean_pred, mean_gt = fg__ / gt_size, gt_fg_numel / gt_size
            dp_fg, dp_bg, dg_fg, dg_bg = 1 - mean_pred, 0 - mean_pred, 1 - mean_gt, 0 - mean_gt
            combos = [(dp_fg, dg_fg), (dp_fg, dg_bg), (dp_bg, dg_fg), (dp_bg, dg_bg)]
            res = np.empty((4, 256), dtype=np.float64)
            for i, (numel, (a, b)) in enumerate(zip(parts, combos)):
                align = 2 * (a * b) / (a ** 2 + b ** 2 + eps)
                res[i] = (align + 1) ** 2 / 4 * numel
            enhanced = res.sum(axis=0)
        return enhanced / (gt_size - 1 + eps)

    def step(self, pred: torch.Tensor, mask: torch.Tensor) -> None:
        pred = pred.to(self.device, torch.float32).contiguous()
        mask = mask.to(self.device, torch.float32).contiguous()
        if pred.ndim != 2 or pred.shape != mask.shape:
            pred, mask = pred.reshape(pred.shape[-2:]), mask.reshape(mask.shape[-2:])
        h, w = pred.shape
        st = self._stats(pred, mask)
        s_score = float(self._s_measure(st, pred, mask, h, w))
        if self.sm_only:
            self.metrics["s_score"].append(s_score)
            return
        n = h * w
        cnt = np.array(st.hist_cnt, dtype=np.float64)
        ysum = np.array(st.hist_y, dtype=np.float64)
        # threshold i selects the pixels with bin >= i + 1 (suffix sums);  metrics.py:323-326 in float32 like the reference
        sel = torch.from_numpy(np.cumsum(cnt[::-1])[::-1][1:].copy()).float()
        tp = torch.from_numpy(np.cumsum(ysum[::-1])[::-1][1:].copy()).float()
        total_y = torch.tensor(st.sum_y).float()
        prec, recall = tp / (sel + 1e-20), tp / (total_y + 1e-20)
        f_score = (1 + 0.3) * prec * recall / (0.3 * prec + recall)  # metrics.py:250-252
        f_score[f_score != f_score] = 0
        self.changeable_ems.append(self._changeable_em(st, n))
        self.weighted_fms.append(self._weighted_f(pred, mask))
        self.metrics["mae"].append(st.abs_err / n)
        self.metrics["max_f"].append(f_score.max().item())
        self.metrics["avg_f"].append(f_score.mean().item())
        self.metrics["s_score"].append(s_score)

    def compute_metrics(self) -> dict:
        if self.sm_only:
            return {"Sm": np.mean(self.metrics["s_score"])}
        return {"MAE": np.mean(self.metrics["mae"]), "MaxF": np.mean(self.metrics["max_f"]), "AvgF": np.mean(self.metrics["avg_f"]),
                "Sm": np.mean(self.metrics["s_score"]),
                "Em": np.mean(np.array(self.changeable_ems, dtype=np.float64), axis=0).mean(),          # metrics.py:134-137
                "wF": np.mean(np.array(self.weighted_fms))}                                               # metrics.py:206-210

    def reset(self) -> None:
        for v in self.metrics.values():
            v.clear()
        self.changeable_ems.clear()
        self.weighted_fms.clear()
