"""Seeded synthetic radar frames and PFN weights.

Pure numpy; shared by tests, bench.py and the golden-vector generator so that every
arm sees identical inputs.  Point layouts follow the reference's loaders:
VoD 7 features (x, y, z, rcs, v_r, v_r_comp, time), TJ4D 8 features
(x, y, z, V_r, Range, Power, Alpha, Beta); `batch_points` prepends the batch-index
column the way DatasetTemplate.collate_batch does (pcdet/datasets/dataset.py:237-244).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# BASELINE.json configs (geometry; see SURVEY.md §8 preamble and §8d)
CONFIGS = {
    "vod":    dict(pc_range=[0, -25.6, -3, 51.2, 25.6, 2], voxel_size=[0.16, 0.16, 5], F=7),
    "tj4d":   dict(pc_range=[0, -39.68, -4, 69.12, 39.68, 2], voxel_size=[0.16, 0.16, 6], F=8),
    "stress": dict(pc_range=[0, -25.6, -3, 51.2, 25.6, 2], voxel_size=[0.1, 0.1, 5], F=7),
}


def make_frame(n: int, pc_range, F: int, seed: int, mode: str = "clustered",
               oob_fraction: float = 0.0) -> np.ndarray:
    """One frame [n, F] fp32.

    mode "uniform":   x, y, z ~ U(range).
    mode "clustered": 10 % uniform "raw radar" followed by 90 % in 40 Gaussian blobs
                      (sigma 0.6 m in x, y; z uniform), blob by blob -- the RHGM virtual points
                      arrive per instance mask (hybrid_pts/hybrid_radar_pts_vod.py:32,152).
                      Points are clipped by the inclusive range mask (common_utils.py:78-81 semantics).
    oob_fraction:     share of points pushed outside the range (exercises the range test).
    Point order is generation order: first-seen pillar order is non-trivial.
    """
    rng = np.random.default_rng(seed)
    r = np.asarray(pc_range, dtype=np.float64)
    lo, hi = r[:3], r[3:]
    if mode == "uniform":
        xyz = rng.uniform(lo, hi, size=(n, 3))
    elif mode == "clustered":
        n_raw = n // 10
        raw = rng.uniform(lo, hi, size=(n_raw, 3))
        K = 40
        centres = rng.uniform(lo[:2], hi[:2], size=(K, 2))
        sizes = np.full(K, (n - n_raw) // K)
        sizes[: (n - n_raw) - sizes.sum()] += 1
        parts = [raw]
        for k in range(K):
            xy = centres[k] + rng.normal(0.0, 0.6, size=(sizes[k], 2))
            z = rng.uniform(lo[2], hi[2], size=(sizes[k], 1))
            parts.append(np.concatenate([xy, z], axis=1))
        xyz = np.concatenate(parts, axis=0)
        xyz[:, :2] = np.clip(xyz[:, :2], lo[:2], hi[:2])
    else:
        raise ValueError(mode)
    if oob_fraction > 0:
        k = int(n * oob_fraction)
        idx = rng.choice(n, size=k, replace=False)
        xyz[idx] += rng.choice([-1.0, 1.0], size=(k, 3)) * (hi - lo) * rng.uniform(0.0, 1.2, size=(k, 3))
    feats = rng.normal(0.0, 1.0, size=(n, F - 3))
    return np.concatenate([xyz, feats], axis=1).astype(np.float32)


def batch_points(frames) -> tuple[np.ndarray, np.ndarray]:
    """collate_batch for the 'points' key: [sum n, 1+F] with the batch index in column 0,
    plus frame_offsets [B+1] int32."""
    rows, offs = [], [0]
    for b, f in enumerate(frames):
        rows.append(np.concatenate([np.full((f.shape[0], 1), b, dtype=np.float32), f], axis=1))
        offs.append(offs[-1] + f.shape[0])
    F1 = rows[0].shape[1] if rows else 1
    pts = np.concatenate(rows, axis=0) if rows else np.zeros((0, F1), np.float32)
    return np.ascontiguousarray(pts, dtype=np.float32), np.asarray(offs, dtype=np.int32)


def make_batch(config: str, B: int, n: int, mode: str = "clustered", seed0: int = 0, oob_fraction: float = 0.0):
    cfg = CONFIGS[config]
    frames = [make_frame(n, cfg["pc_range"], cfg["F"], seed0 + b, mode, oob_fraction) for b in range(B)]
    return batch_points(frames)


@dataclass
class PfnWeights:
    """Random-init single-layer PFN with randomised BatchNorm statistics so that the
    padding constant relu(beta - mean*gamma/sqrt(var+eps)) is non-zero."""
    weight: np.ndarray
    gamma: np.ndarray
    beta: np.ndarray
    running_mean: np.ndarray
    running_var: np.ndarray
    eps: float = 1e-3
    bias: np.ndarray | None = field(default=None)


def make_pfn(Cin: int, C: int = 64, seed: int = 0) -> PfnWeights:
    rng = np.random.default_rng(1000 + seed)
    bound = 1.0 / np.sqrt(Cin)          # nn.Linear default init range
    return PfnWeights(
        weight=rng.uniform(-bound, bound, size=(C, Cin)).astype(np.float32),
        gamma=rng.uniform(0.5, 1.5, size=C).astype(np.float32),
        beta=rng.normal(0.0, 0.5, size=C).astype(np.float32),
        running_mean=rng.normal(0.0, 1.0, size=C).astype(np.float32),
        running_var=rng.uniform(0.5, 2.0, size=C).astype(np.float32),
    )
