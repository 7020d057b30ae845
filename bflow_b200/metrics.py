"""Flow read-out metrics on the device — scope row (f2): the step immediately after ``RAFTSpline.forward``.

Function-level mirrors of the reference's ``utils/metrics.py`` with the same names, argument meaning and return values
(``modules/raft_spline.py:230-238,270-298`` are the callers): ``epe_masked`` (:196-213), ``ae_masked`` (:259-296),
``n_pixel_error_masked`` (:161-193), ``epe_masked_multi`` (:216-240), ``ae_masked_multi`` (:242-257) and
``predictions_from_lin_assumption`` (:298-305).  On CUDA tensors every metric is ONE pass of ``bflow_flow_metrics`` over the
(N, C, *) prediction / target pair; the torchmetrics ``Metric`` classes around them are control plane and are not mirrored.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import check


def flow_metric_sums(source: torch.Tensor, target: torch.Tensor, valid_mask: Optional[torch.Tensor] = None,
                     n_pixels: Sequence[float] = (), source_scale: float = 1.0) -> torch.Tensor:
    """Raw sums of one pass (float64, 8 entries): [sum EPE, valid count, sum angular error (rad), N-pixel-error counts ...]
    with ``source_scale * source`` as the prediction."""
    assert source.is_cuda and target.is_cuda, 'bflow_b200.metrics runs on CUDA tensors'
    assert source.ndim > 2 and source.shape == target.shape
    src, tgt = source.float().contiguous(), target.float().contiguous()
    N, Cc = src.shape[:2]
    HW = src.numel() // (N * Cc)
    v = None
    if valid_mask is not None:
        assert valid_mask.shape[0] == target.shape[0]
        assert valid_mask.ndim == target.ndim - 1
        assert valid_mask.dtype == torch.bool
        assert valid_mask.numel() == N * HW
        v = valid_mask.to(torch.uint8).contiguous()
    th = (C.c_float * 4)(*[float(x) for x in n_pixels][:4])
    assert len(n_pixels) <= 4
    with torch.cuda.device(src.device):
        out = torch.zeros(8, device=src.device, dtype=torch.float64)
        check(_lib.lib().bflow_flow_metrics(src.data_ptr(), tgt.data_ptr(), v.data_ptr() if v is not None else None, N, Cc, HW, float(source_scale),
                                            th, len(n_pixels), out.data_ptr(), torch.cuda.current_stream(src.device).cuda_stream), 'flow_metrics')
    return out


def epe_masked(source: torch.Tensor, target: torch.Tensor, valid_mask: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    s = flow_metric_sums(source, target, valid_mask)
    if valid_mask is not None and float(s[1]) == 0:
        return None                                     # metrics.py:210-211
    return (s[0] / s[1]).float()


def ae_masked(source: torch.Tensor, target: torch.Tensor, valid_mask: Optional[torch.Tensor] = None, degrees: bool = True) -> torch.Tensor:
    s = flow_metric_sums(source, target, valid_mask)
    ae = s[2] / s[1]                                    # 0/0 = nan, as the reference's sum()/sum() over an empty mask
    return (ae / math.pi * 180 if degrees else ae).float()


def n_pixel_error_masked(source: torch.Tensor, target: torch.Tensor, valid_mask: Optional[torch.Tensor], n_pixels: float) -> torch.Tensor:
    s = flow_metric_sums(source, target, valid_mask, n_pixels=(n_pixels,))
    if valid_mask is not None:
        assert float(s[1]) > 0                          # metrics.py:173-174
    return (s[3] / s[1] * 100).float()


def epe_masked_multi(source_lst: List[torch.Tensor], target_lst: List[torch.Tensor],
                     valid_mask_lst: Optional[List[torch.Tensor]] = None) -> Optional[torch.Tensor]:
    num_preds = len(source_lst)
    assert num_preds > 0
    assert len(target_lst) == num_preds, len(target_lst)
    if valid_mask_lst is not None:
        assert len(valid_mask_lst) == num_preds, len(valid_mask_lst)
    else:
        valid_mask_lst = [None] * num_preds
    epe_sum, denominator = 0, 0
    for source, target, valid_mask in zip(source_lst, target_lst, valid_mask_lst):
        epe = epe_masked(source, target, valid_mask)
        if epe is not None:
            epe_sum = epe_sum + epe
            denominator += 1
    if denominator == 0:
        return None
    return epe_sum / denominator


def ae_masked_multi(source_lst: List[torch.Tensor], target_lst: List[torch.Tensor], valid_mask_lst: Optional[List[torch.Tensor]] = None,
                    degrees: bool = True) -> torch.Tensor:
    num_preds = len(source_lst)
    assert num_preds > 0
    assert len(target_lst) == num_preds, len(target_lst)
    if valid_mask_lst is not None:
        assert len(valid_mask_lst) == num_preds, len(valid_mask_lst)
    else:
        valid_mask_lst = [None] * num_preds
    ae_sum = 0
    for source, target, valid_mask in zip(source_lst, target_lst, valid_mask_lst):
        ae_sum = ae_sum + ae_masked(source, target, valid_mask, degrees)
    return ae_sum / num_preds


def predictions_from_lin_assumption(source: torch.Tensor, target_timestamps: List[float]) -> List['ScaledPrediction']:
    """metrics.py:298-305: the prediction at time t under the linearity assumption is t * (final flow).  The products are not
    materialised: every entry is a view (tensor, scale) that the metric functions of this module consume directly."""
    assert max(target_timestamps) <= 1
    assert 0 <= min(target_timestamps)
    return [ScaledPrediction(source, float(ts)) for ts in target_timestamps]


class ScaledPrediction:
    """``scale * tensor`` evaluated inside the metric kernel (``source_scale``); ``.tensor()`` materialises it."""

    def __init__(self, base: torch.Tensor, scale: float):
        self.base, self.scale = base, scale
        self.shape, self.ndim = base.shape, base.ndim

    def tensor(self) -> torch.Tensor:
        return self.base * self.scale


def lin_assumption_metrics(final_flow: torch.Tensor, target_timestamps: List[float], target_lst: List[torch.Tensor],
                           valid_mask_lst: Optional[List[torch.Tensor]] = None, degrees: bool = True):
    """(epe_multi_lin, ae_multi_lin) of modules/raft_spline.py:290-297 in len(target_timestamps) kernel passes, no t * flow tensors."""
    assert len(target_timestamps) == len(target_lst)
    masks = valid_mask_lst if valid_mask_lst is not None else [None] * len(target_lst)
    epe_sum, epe_n, ae_sum = 0, 0, 0
    for ts, tgt, m in zip(target_timestamps, target_lst, masks):
        assert 0 <= ts <= 1
        s = flow_metric_sums(final_flow, tgt, m, source_scale=float(ts))
        if m is None or float(s[1]) > 0:
            epe_sum = epe_sum + (s[0] / s[1]).float()
            epe_n += 1
        ae = s[2] / s[1]
        ae_sum = ae_sum + (ae / math.pi * 180 if degrees else ae).float()
    return (epe_sum / epe_n if epe_n else None), ae_sum / len(target_lst)
