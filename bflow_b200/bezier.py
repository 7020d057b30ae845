"""Return type of :meth:`bflow_b200.RAFTSpline.forward`.

Mirrors the public surface of the reference's ``BezierCurves`` (models/raft_spline/bezier.py:17-216)
that its callers use: ``get_flow_from_reference``, ``get_params``, ``detach``,
``cpu``, ``requires_grad``, ``batch_size/height/width/degree/dim``, ``delta_update_params``,
``create_upsampled`` and construction from a raw ``(B, 2*deg, H, W)`` tensor.  The curve is in
Bernstein form with P0 == 0; channel ``d*deg + (i-1)`` holds coordinate d of control point P_i
(bezier.py:134-135).  Evaluation on CUDA tensors runs the ``bflow_bezier_eval`` kernel.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Union

import numpy as np
import torch as th


def bernstein_coeffs(timestamps: Sequence[float], degree: int) -> np.ndarray:
    """C(n,i) (1-t)^(n-i) t^i for i = 1..n as float64 (bezier.py:141-163)."""
    ts = np.asarray(timestamps, dtype='float64').reshape(-1)
    assert ts.size > 0 and ts.min() >= 0 and ts.max() <= 1
    i = np.arange(1, degree + 1, dtype='float64')
    binom = np.array([math.comb(degree, int(k)) for k in i], dtype='float64')
    return binom[None, :] * (1.0 - ts[:, None]) ** (degree - i[None, :]) * ts[:, None] ** i[None, :]


class BezierCurves:
    CTRL_DIM: int = 2

    def __init__(self, bezier_params: th.Tensor, ready_event=None):
        assert bezier_params.ndim == 4
        assert bezier_params.shape[1] % 2 == 0
        self._raw = bezier_params
        # results of a pipelined forward (RAFTSpline.forward(non_blocking=True)) are in flight until this event completes
        self._ready_event = ready_event
        self.batch, channels, self.ht, self.wd = bezier_params.shape
        self.n_ctrl_pts = channels // self.CTRL_DIM + 1

    # ---- constructors ---------------------------------------------------------------------
    @classmethod
    def create_from_specification(cls, batch_size: int, n_ctrl_pts: int, height: int, width: int,
                                  device: th.device) -> 'BezierCurves':
        assert batch_size > 0 and n_ctrl_pts > 1 and height > 0 and width > 0
        return cls(th.zeros(batch_size, cls.CTRL_DIM * (n_ctrl_pts - 1), height, width, device=device))

    @classmethod
    def from_2view(cls, flow_tensor: th.Tensor) -> 'BezierCurves':
        assert flow_tensor.shape[1] == cls.CTRL_DIM
        return cls(flow_tensor)

    @classmethod
    def create_from_voxel_grid(cls, voxel_grid: th.Tensor, downsample_factor: int = 8,
                               bezier_degree: int = 2) -> 'BezierCurves':
        assert isinstance(downsample_factor, int) and downsample_factor >= 1
        batch, _, ht, wd = voxel_grid.shape
        assert ht % 8 == 0 and wd % 8 == 0
        return cls.create_from_specification(batch, bezier_degree + 1, ht // downsample_factor,
                                             wd // downsample_factor, voxel_grid.device)

    # ---- tensor plumbing -------------------------------------------------------------------
    @property
    def _params(self) -> th.Tensor:
        if self._ready_event is not None:
            self._ready_event.synchronize()
            self._ready_event = None
        return self._raw

    @_params.setter
    def _params(self, value: th.Tensor) -> None:
        self._raw = value
        self._ready_event = None

    @property
    def device(self):
        return self._raw.device

    @property
    def dtype(self):
        return self._raw.dtype

    @property
    def requires_grad(self):
        return self._raw.requires_grad

    @property
    def batch_size(self):
        return self._raw.shape[0]

    @property
    def degree(self):
        return self.n_ctrl_pts - 1

    @property
    def dim(self):
        return self._raw.shape[1]

    @property
    def height(self):
        return self._raw.shape[-2]

    @property
    def width(self):
        return self._raw.shape[-1]

    def get_params(self) -> th.Tensor:
        return self._params

    def detach(self, clone: bool = False, cpu: bool = False) -> 'BezierCurves':
        p = self._params.detach()
        if cpu:
            return BezierCurves(p.cpu())
        return BezierCurves(p.clone() if clone else p)

    def detach_(self, cpu: bool = False) -> None:
        self._params = self._params.detach()
        if cpu:
            self._params = self._params.cpu()

    def cpu(self) -> 'BezierCurves':
        return BezierCurves(self._params.cpu())

    def cpu_(self) -> None:
        self._params = self._params.cpu()

    def delta_update_params(self, delta_bezier: th.Tensor) -> None:
        assert delta_bezier.shape == self._params.shape
        self._params = self._params + delta_bezier

    def create_upsampled(self, mask: th.Tensor) -> 'BezierCurves':
        """Convex 8x upsampling of every control-point channel (bezier.py:81-84 → utils.py:33-48)."""
        from . import ops
        return BezierCurves(ops.cvx_upsample(self._params, mask))

    # ---- evaluation ------------------------------------------------------------------------
    def _eval(self, ts: np.ndarray) -> th.Tensor:
        coef = bernstein_coeffs(ts, self.degree)
        if self._params.is_cuda:
            from . import ops
            return ops.bezier_eval(self._params, coef)
        c = th.from_numpy(coef).float()
        p = self._params.view(self.batch, 2, self.degree, self.ht, self.wd)
        return (p[None] * c[:, None, None, :, None, None]).sum(dim=3)

    def get_flow_from_reference(self, time: Union[float, int, List[float], np.ndarray]) -> th.Tensor:
        """bezier.py:188-216: scalar → (B,2,H,W); list/array of T timestamps → (T,B,2,H,W)."""
        scalar = isinstance(time, (int, float))
        if scalar:
            assert 0.0 <= time <= 1.0
            if time == 1:
                return self._params.view(self.batch, 2, self.degree, self.ht, self.wd)[:, :, -1]
            if time == 0:
                return th.zeros((self.batch, 2, self.ht, self.wd), dtype=self.dtype, device=self.device)
            ts = np.array([time], dtype='float64')
        elif isinstance(time, list):
            ts = np.asarray(time, dtype='float64')
        else:
            assert isinstance(time, np.ndarray) and time.dtype == 'float64'
            ts = time
        flows = self._eval(ts)
        return flows[0] if scalar else flows
