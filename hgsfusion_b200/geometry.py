"""Host-side grid arithmetic, mirroring the reference expression for expression.

  grid_size      pcdet/datasets/processor/data_processor.py:135-136
  centre offsets pcdet/models/backbones_3d/vfe/pillar_vfe.py:76-81
"""
from __future__ import annotations

import numpy as np

from . import _lib


def grid_size(point_cloud_range, voxel_size) -> np.ndarray:
    """round((range[3:6] - range[0:3]) / VOXEL_SIZE) as int64, as DataProcessor computes it."""
    r = np.asarray(point_cloud_range)
    g = (r[3:6] - r[0:3]) / np.array(voxel_size)
    return np.round(g).astype(np.int64)


def make_geometry(point_cloud_range, voxel_size, grid=None) -> _lib.Geometry:
    """`point_cloud_range` is used with whatever scalar type the caller holds (OpenPCDet passes an
    np.float32 array), so `voxel/2 + range_min` rounds exactly as it does inside the reference."""
    r = point_cloud_range
    grid = grid_size(np.asarray(r, dtype=np.float32), voxel_size) if grid is None else grid
    vx, vy, vz = voxel_size[0], voxel_size[1], voxel_size[2]
    x_off = vx / 2 + r[0]
    y_off = vy / 2 + r[1]
    z_off = vz / 2 + r[2]
    g = _lib.Geometry()
    g.pc_range[:] = [float(np.float32(v)) for v in r]
    g.voxel_size[:] = [float(np.float32(vx)), float(np.float32(vy)), float(np.float32(vz))]
    g.grid[:] = [int(grid[0]), int(grid[1]), int(grid[2])]
    g.centre_off[:] = [float(np.float32(x_off)), float(np.float32(y_off)), float(np.float32(z_off))]
    return g
