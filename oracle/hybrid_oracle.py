"""TEST INFRASTRUCTURE -- numpy restatement of the reference's hybrid point assembly, the
step in front of the pillar path: raw radar sweep + real points inside instance masks + virtual (RHGM) points -> one
[N, F+2] array with two flag columns, FOV filter, range mask.  Only tests/ may import this.

PINNED: tests/golden/hybrid_*.npz hold the outputs of the reference's own statements, executed unchanged from
/root/reference by tests/golden/make_hybrid_golden.py; tests/test_hybrid_oracle.py checks this file against them bit for bit.

Reference (file:line under the HGSFusion repo):
  assembly      pcdet/datasets/kitti/vod_dataset.py:498-522, pcdet/datasets/kitti/tj4d_dataset.py:588-610
  NO_DUP        vod_dataset.py:13-19 (calc_dist), :511-514
  FOV filter    vod_dataset.py:181-197,525-528; pcdet/utils/calibration_kitti.py:68-88
  range mask    pcdet/utils/common_utils.py:78-81 through data_processor.py:83-85
  float32 cast  pcdet/models/__init__.py:23-36 (load_data_to_gpu: .float())
All arithmetic is float64, as in the reference (np.ones / the sweep normalisation make the array float64).
"""
from __future__ import annotations

import numpy as np


def lidar_to_rect_matrix(V2C: np.ndarray, R0: np.ndarray) -> np.ndarray:
    """The [4,3] matrix of Calibration.lidar_to_rect (calibration_kitti.py:74): V2C^T . R0^T, a float32 product."""
    return np.dot(np.asarray(V2C, dtype=np.float32).T, np.asarray(R0, dtype=np.float32).T)


def assemble_frame(real, gt_real, virt, *, use_virtual=True, no_dup=False, fov=None, pc_range=None, dataset="vod"):
    """real [Nr,Fr], gt_real [Ng,Fr+8], virt [Nv,Fr+8] float32 -> points [N', Fr+10] float32 (or [N', Fr] without virtual points).

    fov = None or (lidar_to_rect [4,3] float32, P2 [3,4] float32, (img_h, img_w)); pc_range = None or 6 Python floats."""
    real = np.asarray(real, dtype=np.float32).astype(np.float64)
    Fr = real.shape[1]
    if use_virtual:
        gt = np.asarray(gt_real, dtype=np.float32).astype(np.float64)
        vt = np.asarray(virt, dtype=np.float32).astype(np.float64)
        W = Fr + 8
        if len(gt) == 0:
            # vod_dataset.py:507-509: only the sweep survives, every other column stays 1
            if dataset == "tj4d" and len(vt) > 0:
                raise ValueError("tj4d_dataset.py:603-605 broadcasts [Nr,8] into Nr+Nv rows: the reference raises here")
            pts = np.ones((len(real), W + 2))
            pts[:, :Fr] = real
        else:
            if no_dup and len(real) > 0:
                d = gt[:, None, :3] - real[None, :, :3]                       # calc_dist: [Ng, Nr, 3]
                d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
                real = real[np.abs(d2.min(0)) > 0.001]
            pts = np.ones((len(real) + len(gt) + len(vt), W + 2))
            pts[:len(real), :Fr] = real                                      # flags (1, 1): raw radar
            pts[len(real):, :W] = np.concatenate([gt, vt])
            pts[len(real):, W] = 0.0                                          # (0, 0): real points inside a mask
            pts[len(real):, W + 1] = 0.0
            # (0, 1): virtual points.  vod_dataset.py:521 writes points[-Nv:, -1] = 1; with Nv == 0 that slice is the
            # WHOLE array, so a frame with masks but no virtual points gets 1 in the last column of every row
            pts[(len(pts) - len(vt)) if len(vt) > 0 else 0:, W + 1] = 1.0
    else:
        pts = real
    if fov is not None:
        M, P2, (img_h, img_w) = fov
        M = np.asarray(M, dtype=np.float32).astype(np.float64)
        P2 = np.asarray(P2, dtype=np.float32).astype(np.float64)
        x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
        rect = [x * M[0, j] + y * M[1, j] + z * M[2, j] + M[3, j] for j in range(3)]
        hom = [rect[0] * P2[j, 0] + rect[1] * P2[j, 1] + rect[2] * P2[j, 2] + P2[j, 3] for j in range(3)]
        with np.errstate(divide="ignore", invalid="ignore"):
            u, v = hom[0] / rect[2], hom[1] / rect[2]
        depth = hom[2] - P2[2, 3]
        pts = pts[(u >= 0) & (u < img_w) & (v >= 0) & (v < img_h) & (depth >= 0)]
    if pc_range is not None:
        r = [float(np.float32(v)) for v in pc_range]      # limit_range is a float32 array in the reference (dataset.py:26)
        pts = pts[(pts[:, 0] >= r[0]) & (pts[:, 0] <= r[3]) & (pts[:, 1] >= r[1]) & (pts[:, 1] <= r[4])]
    return pts.astype(np.float32)


def assemble_frame_from_fixture(data, b: int) -> np.ndarray:
    fov = None
    if int(data["fov"]):
        fov = (lidar_to_rect_matrix(data[f"V2C_{b}"], data[f"R0_{b}"]), data[f"P2_{b}"], tuple(int(v) for v in data["image_shape"]))
    dataset = "tj4d" if int(data["Fr"]) == 8 else "vod"
    return assemble_frame(data[f"real{b}"], data[f"gt{b}"], data[f"virt{b}"], use_virtual=bool(int(data["use_virtual"])),
                          no_dup=bool(int(data["no_dup"])), fov=fov, pc_range=list(data["pc_range"]), dataset=dataset)


def assemble_batch(frames, **kw) -> tuple[np.ndarray, np.ndarray]:
    """frames = [(real, gt, virt, fov)] -> collated points [sum N', 1+F] (column 0 = frame index, dataset.py:237-244), offsets."""
    rows, offs = [], [0]
    for b, (real, gt, virt, fov) in enumerate(frames):
        p = assemble_frame(real, gt, virt, fov=fov, **kw)
        rows.append(np.concatenate([np.full((len(p), 1), b, np.float32), p], 1))
        offs.append(offs[-1] + len(p))
    return np.concatenate(rows, 0), np.asarray(offs, dtype=np.int32)
