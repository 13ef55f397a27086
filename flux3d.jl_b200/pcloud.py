"""PointCloud — mirror of the reference struct (src/rep/pcloud.jl:25-57).

The reference stores ``points::Array{Float32,3}`` of Julia shape (D, N, B) (column-major).  The same
bytes read row-major are a C / torch array of shape (B, N, D), which is the shape used here: a torch
tensor ``points[b, n, :]`` is point n of batch element b.  A 2-D (N, D) input becomes B = 1
(pcloud.jl:30-43); any element type is converted to Float32 (pcloud.jl:45-51)."""
from __future__ import annotations

import numpy as np
import torch


def as_f32_tensor(x, device=None) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if x.dtype != torch.float32:
        x = x.to(torch.float32)
    if device is not None and x.device != torch.device(device):
        x = x.to(device, non_blocking=True)
    return x.contiguous()


class PointCloud:
    """Batched point cloud.  ``points``: (B, N, 3) float32 (== Julia (3, N, B)); optional ``normals``
    of the same shape."""

    def __init__(self, points, normals=None, device=None):
        p = as_f32_tensor(points, device)
        if p.dim() == 2:
            p = p.unsqueeze(0)
        if p.dim() != 3:
            raise ValueError("points must be (N, D) or (B, N, D)")
        self.points = p
        if normals is not None:
            n = as_f32_tensor(normals, device)
            if n.dim() == 2:
                n = n.unsqueeze(0)
            if n.shape != p.shape:
                raise ValueError("normals must have the same shape as points")  # pcloud.jl:36-39
            self.normals = n
        else:
            self.normals = None

    def __len__(self):
        return self.points.shape[0]

    def __getitem__(self, i):
        return self.points[i]

    def npoints(self):
        return self.points.shape[1]

    def to(self, device):
        return PointCloud(self.points.to(device), None if self.normals is None else self.normals.to(device))

    def cuda(self):
        return self.to("cuda")

    def __repr__(self):
        return (f"PointCloud{{Float32}} Structure:\n    Batch size: {self.points.shape[0]}\n"
                f"    Points: {self.points.shape[1]}\n    Dimension: {self.points.shape[2]}\n"
                f"    Storage: {self.points.device}")
