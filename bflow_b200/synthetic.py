"""Seeded synthetic inputs of the shapes the reference's datasets deliver.

``sparse_norm`` mimics ``norm_voxel_grid`` (data/utils/representations.py:9-18): ~10 % non-zero
voxels, non-zeros standardised to zero mean / unit std.  Everything is drawn on the CPU from a
``torch.Generator`` so the same tensors are reproduced on the build container and the GPU box.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import config as _cfg


def voxel_grid(cfg: dict, batch: int, height: int, width: int, seed: int = 1234, kind: str = 'sparse_norm',
               pinned: bool = False) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    c = _cfg.input_channels(cfg)
    x = torch.randn(batch, c, height, width, generator=g)
    if kind == 'sparse_norm':
        keep = torch.rand(batch, c, height, width, generator=g) < 0.10
        x = x * keep
        nz = x != 0
        v = x[nz]
        x[nz] = (v - v.mean()) / v.std()
    elif kind != 'randn':
        raise ValueError(kind)
    return x.pin_memory() if pinned else x


def images(batch: int, height: int, width: int, seed: int = 4321, pinned: bool = False) -> List[torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    out = [torch.randint(0, 256, (batch, 3, height, width), generator=g).float() for _ in range(2)]
    return [t.pin_memory() for t in out] if pinned else out


def inputs(cfg: dict, batch: int, height: int, width: int, seed: int = 1234, kind: str = 'sparse_norm',
           pinned: bool = False) -> Tuple[Optional[torch.Tensor], Optional[List[torch.Tensor]]]:
    vg = voxel_grid(cfg, batch, height, width, seed, kind, pinned) if cfg['use_events'] else None
    im = images(batch, height, width, seed + 1, pinned) if cfg['use_boundary_images'] else None
    return vg, im


def lookup_case(batch: int, h: int, w: int, dim: int = 256, targets: int = 1, seed: int = 7):
    """Config #5 (correlation-lookup microbench): random feature maps and perturbed coordinates."""
    g = torch.Generator().manual_seed(seed)
    f1 = torch.randn(batch, dim, h, w, generator=g)
    f2 = torch.randn(targets, batch, dim, h, w, generator=g)
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    grid = torch.stack([xs, ys], 0).float()
    coords = grid[None, None] + 8 * torch.randn(targets, batch, 2, h, w, generator=g)
    return f1, f2, coords
