"""Event representation on the GPU — scope row (f1): the step immediately before ``RAFTSpline.forward``.

Mirrors ``data/utils/representations.py`` of the reference (``VoxelGrid.convert`` :64-111, ``norm_voxel_grid`` :9-18) with the same
constructor and call signature, but on CUDA tensors: the reference builds every voxel grid on the CPU inside DataLoader workers.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from ._lib import check


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def norm_voxel_grid(voxel_grid: torch.Tensor) -> torch.Tensor:
    """In place: standardise the non-zero voxels (mean / unbiased std over them)."""
    assert voxel_grid.is_cuda and voxel_grid.dtype == torch.float32 and voxel_grid.is_contiguous()
    with torch.cuda.device(voxel_grid.device):
        stats = torch.empty(3, device=voxel_grid.device, dtype=torch.float64)
        check(_lib.lib().bflow_voxel_norm(voxel_grid.data_ptr(), voxel_grid.numel(), stats.data_ptr(), _stream(voxel_grid)), 'voxel_norm')
    return voxel_grid


class VoxelGrid:
    def __init__(self, channels: int, height: int, width: int, check_bounds: bool = True):
        """``check_bounds``: raise IndexError (one device->host read) when integer coordinates fall outside the grid, like the
        reference; False skips the read -- such events are then silently dropped, never written out of bounds."""
        assert channels > 1 and height > 1 and width > 1
        self.nb_channels, self.height, self.width = channels, height, width
        self.check_bounds = check_bounds

    def _get_dt(self, t0_center: int, t1_center: int):
        assert t1_center > t0_center
        return (t1_center - t0_center) / (self.nb_channels - 1)

    def get_extended_time_window(self, t0_center: int, t1_center: int):
        dt = self._get_dt(t0_center, t1_center)
        return math.floor(t0_center - dt), math.ceil(t1_center + dt)

    def convert(self, x: torch.Tensor, y: torch.Tensor, pol: torch.Tensor, time: torch.Tensor, t0_center: Optional[int] = None,
                t1_center: Optional[int] = None) -> torch.Tensor:
        assert x.is_cuda and x.device == y.device == pol.device == time.device, 'bflow_b200.events runs on CUDA tensors'
        assert type(t0_center) == type(t1_center)
        assert x.shape == y.shape == pol.shape == time.shape and x.ndim == 1
        assert not torch.is_floating_point(time) and not torch.is_complex(time)
        is_int_xy = not torch.is_floating_point(x)
        if is_int_xy:
            assert not torch.is_floating_point(y)
            xs, ys = x.long().contiguous(), y.long().contiguous()
        else:
            xs, ys = x.float().contiguous(), y.float().contiguous()
        t = time.long().contiguous()
        p = pol.to(torch.uint8).contiguous()
        t0 = int(t0_center) if t0_center is not None else int(t[0])
        t1 = int(t1_center) if t1_center is not None else int(t[-1])
        with torch.cuda.device(x.device):
            out = torch.zeros(self.nb_channels, self.height, self.width, device=x.device, dtype=torch.float32)
            oob = torch.zeros(1, device=x.device, dtype=torch.int32) if is_int_xy else None
            check(_lib.lib().bflow_voxelize(xs.data_ptr(), ys.data_ptr(), int(not is_int_xy), p.data_ptr(), t.data_ptr(), t.numel(), t0, t1,
                                            self.nb_channels, self.height, self.width, out.data_ptr(), oob.data_ptr() if oob is not None else None,
                                            _stream(x)), 'voxelize')
            if oob is not None and self.check_bounds:
                n_bad = int(oob.item())
                if n_bad:       # the reference's put_ raises for an index outside the grid (representations.py:96-99); never written here
                    raise IndexError(f'VoxelGrid.convert: {n_bad} events with integer coordinates outside [0, {self.width}) x [0, {self.height})')
        return out


def epe_sum_count(flow: torch.Tensor, target: torch.Tensor, valid: Optional[torch.Tensor] = None):
    """Scope row (f2): masked end-point error on the device (utils/metrics.py:196-213) → (sum f64, count i64) tensors."""
    assert flow.is_cuda and flow.shape == target.shape and flow.ndim > 2
    f, t = flow.float().contiguous(), target.float().contiguous()
    N, Cc = f.shape[:2]
    HW = f.numel() // (N * Cc)
    v = None
    if valid is not None:
        assert valid.shape[0] == N and valid.numel() == N * HW
        v = valid.to(torch.uint8).contiguous()
    with torch.cuda.device(f.device):
        out = torch.zeros(2, device=f.device, dtype=torch.float64)
        check(_lib.lib().bflow_epe_masked(f.data_ptr(), t.data_ptr(), v.data_ptr() if v is not None else None, N, Cc, HW, out.data_ptr(), _stream(f)),
              'epe_masked')
    return out[0], out[1].to(torch.int64)
