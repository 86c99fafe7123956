"""Device-side mirror of the reference's `DataProcessor.transform_points_to_voxels` hook
(pcdet/datasets/processor/data_processor.py:133-183): the same config keys, the same data_dict keys in and out, but the
voxelization runs on the GPU through the native library instead of spconv on a DataLoader worker.

    proc = TransformPointsToVoxels(config, point_cloud_range, num_point_features, mode='test')
    data_dict = proc(data_dict)        # data_dict['points']: CUDA float32 [N, F] of ONE frame (as the reference's hook sees it)

Config keys (as in the YAML): VOXEL_SIZE, MAX_POINTS_PER_VOXEL, MAX_NUMBER_OF_VOXELS {'train','test'}, DOUBLE_FLIP (optional);
plus SPCONV_VERSION (optional, 1 or 2: whose overflow semantics to reproduce, see ops.PillarPath).  `data_dict['use_lead_xyz']`
False drops the xyz columns of `voxels` (data_processor.py:158-159).  Outputs: voxels [M,P,F(-3)], voxel_coords [M,3] int32
(z, y, x), voxel_num_points [M] int32 -- or, with DOUBLE_FLIP, lists of four (original, yflip, xflip, xyflip; :161-178).
"""
from __future__ import annotations

import numpy as np
import torch

from .ops import PillarPath


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class TransformPointsToVoxels:
    def __init__(self, config, point_cloud_range, num_point_features: int, mode: str = "test"):
        self.config = config
        self.voxel_size = list(_get(config, "VOXEL_SIZE"))
        self.point_cloud_range = np.asarray(point_cloud_range, dtype=np.float32)
        mv = _get(config, "MAX_NUMBER_OF_VOXELS")
        self.max_voxels = int(mv[mode] if isinstance(mv, dict) else mv)
        self.max_points = int(_get(config, "MAX_POINTS_PER_VOXEL"))
        self.double_flip = bool(_get(config, "DOUBLE_FLIP", False))
        # grid_size / voxel_size as the reference's hook publishes them on first call (:135-138)
        self.grid_size = np.round((self.point_cloud_range[3:6] - self.point_cloud_range[0:3]) / np.array(self.voxel_size)).astype(np.int64)
        self.path = PillarPath(self.point_cloud_range, self.voxel_size, self.max_points, self.max_voxels, int(num_point_features),
                               spconv_version=int(_get(config, "SPCONV_VERSION", 2)))

    def _one(self, res, use_lead_xyz):
        out = res.trim()
        voxels = out["voxels"] if use_lead_xyz else out["voxels"][..., 3:]        # remove xyz in voxels (:158-159)
        return voxels, out["voxel_coords"][:, 1:], out["voxel_num_points"]

    def __call__(self, data_dict):
        points = data_dict["points"]
        if not (torch.is_tensor(points) and points.is_cuda):
            raise ValueError("data_dict['points'] must be a CUDA tensor (hgsfusion_b200 has no CPU path)")
        use_lead_xyz = bool(data_dict.get("use_lead_xyz", True))
        offs = torch.tensor([0, points.shape[0]], dtype=torch.int32, device=points.device)
        if self.double_flip:
            res = self.path.pillarize_double_flip(points, 1, xyz_col=0, batch_col=-1, frame_offsets=offs)
            vs, cs, ns = zip(*[self._one(r, use_lead_xyz) for r in res])
            data_dict["voxels"], data_dict["voxel_coords"], data_dict["voxel_num_points"] = list(vs), list(cs), list(ns)
        else:
            v, c, n = self._one(self.path.pillarize(points, 1, xyz_col=0, batch_col=-1, frame_offsets=offs), use_lead_xyz)
            data_dict["voxels"], data_dict["voxel_coords"], data_dict["voxel_num_points"] = v, c, n
        return data_dict
