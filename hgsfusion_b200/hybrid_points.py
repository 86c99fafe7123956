"""Hybrid point assembly on the device: host-side mirror of what the reference's datasets do per sample in numpy
(SURVEY.md section 8(f) rank 3), through hgsf_assemble_hybrid_points.  torch is device memory and the current stream only.

Reference (file:line under the HGSFusion repo):
  VODDataset.__getitem__ points block      pcdet/datasets/kitti/vod_dataset.py:498-529   (USE_VIRTUAL_POINTS, NO_DUP, FOV_POINTS_ONLY)
  TJ4DDataset.__getitem__ points block     pcdet/datasets/kitti/tj4d_dataset.py:588-618
  DataProcessor.mask_points_and_boxes_outside_range -> mask_points_by_range      data_processor.py:79-93, common_utils.py:78-81
  Calibration.lidar_to_rect / rect_to_img  pcdet/utils/calibration_kitti.py:68-88
The result is the collated `points [sum N, 1+F]` tensor (column 0 = frame index, dataset.py:237-244) that
batch_dict['points'] holds after load_data_to_gpu, plus the frame offsets -- what hgsf_points_to_bev, hgsf_split_encode and
the PillarNet reader take, so the frames never visit the host between the files and the canvas.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


def calib_row(V2C, R0, P2, image_shape) -> np.ndarray:
    """One frame's [26] float32 calibration row: the [4,3] lidar->rect matrix exactly as Calibration.lidar_to_rect forms
    it (np.dot(V2C.T, R0.T) on the float32 matrices, calibration_kitti.py:74), P2 [3,4], image height, image width."""
    V2C, R0, P2 = (np.asarray(a, dtype=np.float32) for a in (V2C, R0, P2))
    if V2C.shape != (3, 4) or R0.shape != (3, 3) or P2.shape != (3, 4):
        raise ValueError("V2C [3,4], R0 [3,3], P2 [3,4] expected")
    m = np.dot(V2C.T, R0.T)
    return np.concatenate([m.reshape(-1), P2.reshape(-1), np.asarray(image_shape[:2], dtype=np.float32)]).astype(np.float32)


@dataclass
class HybridPoints:
    """points [capacity, 1+F] (rows [0, frame_offsets[-1]) are valid), frame_offsets int32 [B+1], both on the device."""
    points: torch.Tensor
    frame_offsets: torch.Tensor

    def trim(self) -> torch.Tensor:
        """[N, 1+F]: one host sync for the row count (the reference's contract needs the exact shape)."""
        return self.points[: int(self.frame_offsets[-1].item())]


def _offsets(counts_or_offsets, B, device, name):
    t = torch.as_tensor(counts_or_offsets)
    if t.numel() != B + 1:
        raise ValueError(f"{name} must hold batch_size + 1 offsets")
    return t.to(device=device, dtype=torch.int32).contiguous()


def assemble_hybrid_points(real, real_offsets, gt_real=None, gt_offsets=None, virt=None, virt_offsets=None, *, batch_size: int,
                           calib=None, point_cloud_range=None, no_dup: bool = False, dup_threshold: float = 0.001,
                           dataset: str = "vod", out: HybridPoints | None = None, workspace=None) -> HybridPoints:
    """real [sum Nr, Fr], gt_real / virt [.., Fr+8] float32 CUDA tensors, frames concatenated; *_offsets [B+1] (host lists
    or tensors).  gt_real = virt = None: USE_VIRTUAL_POINTS False.  calib = [B,26] rows of `calib_row` (None:
    FOV_POINTS_ONLY False); point_cloud_range = the 6 Python floats of POINT_CLOUD_RANGE (None: no range mask)."""
    lib = _lib.load()
    if not real.is_cuda or real.dtype != torch.float32 or real.dim() != 2:
        raise ValueError("real must be a float32 CUDA tensor [n, Fr] (hgsfusion_b200 has no CPU path)")
    dev, B = real.device, int(batch_size)
    real = real.contiguous()
    Fr = int(real.shape[1])
    hybrid = gt_real is not None or virt is not None
    host_off = [torch.as_tensor(o).cpu().to(torch.int64) for o in (real_offsets,) + ((gt_offsets, virt_offsets) if hybrid else ())]
    ro = _offsets(real_offsets, B, dev, "real_offsets")
    s = _lib.HybridInputs()
    keep = [real, ro]
    s.real, s.real_offsets, s.real_features, s.batch_size = real.data_ptr(), ro.data_ptr(), Fr, B
    n = int(host_off[0][-1])
    W = 0
    if hybrid:
        if gt_real is None or virt is None or gt_offsets is None or virt_offsets is None:
            raise ValueError("gt_real, virt and their offsets come together")
        W = Fr + 8
        for name, t in (("gt_real", gt_real), ("virt", virt)):
            if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or int(t.shape[1]) != W:
                raise ValueError(f"{name} must be a float32 CUDA tensor [n, {W}]")
        if dataset == "tj4d":
            ng, nv = host_off[1][1:] - host_off[1][:-1], host_off[2][1:] - host_off[2][:-1]
            if bool(((ng == 0) & (nv > 0)).any()):
                # tj4d_dataset.py:603-605 assigns the [Nr, 8] sweep into Nr + Nv rows: numpy raises a broadcast error
                raise ValueError("TJ4D frame with virtual points but no mask points: the reference raises here")
        gt_real, virt = gt_real.contiguous(), virt.contiguous()
        go, vo = _offsets(gt_offsets, B, dev, "gt_offsets"), _offsets(virt_offsets, B, dev, "virt_offsets")
        keep += [gt_real, virt, go, vo]
        s.gt_real, s.virt, s.gt_offsets, s.virt_offsets = gt_real.data_ptr(), virt.data_ptr(), go.data_ptr(), vo.data_ptr()
        n += int(host_off[1][-1]) + int(host_off[2][-1])
    s.n_candidates, s.hybrid_features, s.no_dup, s.dup_threshold = n, W, int(bool(no_dup)), float(dup_threshold)
    cal_ptr = None
    if calib is not None:
        cal = torch.as_tensor(np.asarray(calib, dtype=np.float32) if not torch.is_tensor(calib) else calib)
        cal = cal.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(cal.shape) != (B, 26):
            raise ValueError("calib must be [batch_size, 26] (see calib_row)")
        keep.append(cal)
        cal_ptr = C.c_void_p(cal.data_ptr())
    rng = None
    if point_cloud_range is not None:
        # the reference compares against DatasetTemplate.point_cloud_range = np.array(..., dtype=np.float32) (dataset.py:26):
        # the limits are the float32 roundings of the YAML values, widened exactly to double (float32(51.2) > 51.2)
        r = [float(np.float32(v)) for v in point_cloud_range]
        rng = (C.c_double * 4)(r[0], r[1], r[3], r[4])
    need = C.c_size_t(0)
    _lib.check(lib.hgsf_hybrid_workspace_size(n, C.byref(need)), "hgsf_hybrid_workspace_size")
    if workspace is None or workspace.numel() < need.value + 256:
        workspace = torch.empty(need.value + 256, dtype=torch.uint8, device=dev)
    ws_ptr = (workspace.data_ptr() + 255) // 256 * 256
    F = W + 2 if hybrid else Fr
    if out is None:
        out = HybridPoints(points=torch.empty((max(n, 1), 1 + F), dtype=torch.float32, device=dev),
                           frame_offsets=torch.empty(B + 1, dtype=torch.int32, device=dev))
    elif out.points.shape[0] < n or out.points.shape[1] != 1 + F or out.frame_offsets.numel() != B + 1:
        raise ValueError("out: capacity / width mismatch")
    st = lib.hgsf_assemble_hybrid_points(C.byref(s), cal_ptr, rng, C.c_void_p(ws_ptr), need.value, C.c_void_p(out.points.data_ptr()),
                                         C.c_void_p(out.frame_offsets.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(st, "hgsf_assemble_hybrid_points")
    del keep
    return out
