"""ctypes front end of oracle/liboracle.so plus a slow pure-Python twin.

TEST INFRASTRUCTURE ONLY (see oracle/pillar_oracle.c header).  The C file is the
oracle; `voxelize_py` below is an independent loop-level restatement used to
cross-check the C on small cases.

Reference citations (under /root/reference):
  voxelizer      pcdet/datasets/processor/data_processor.py:37-43,55-60 (spconv call site)
  PillarVFE      pcdet/models/backbones_3d/vfe/pillar_vfe.py:52-123
  scatter        pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:5-41
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "pillar_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return so


class _VfeCfg(C.Structure):
    _fields_ = [("F", C.c_int32), ("P", C.c_int32), ("C", C.c_int32),
                ("use_absolute_xyz", C.c_int32), ("with_distance", C.c_int32), ("use_norm", C.c_int32),
                ("vx", C.c_float), ("vy", C.c_float), ("vz", C.c_float),
                ("x_off", C.c_float), ("y_off", C.c_float), ("z_off", C.c_float),
                ("eps", C.c_float)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_voxelize.restype = C.c_int32
        _LIB.orc_voxelize_mode.restype = C.c_int32
        _LIB.orc_voxelize_batch.restype = C.c_int64
        _LIB.orc_points_to_bev.restype = C.c_int64
        _LIB.orc_mask_points_by_range.restype = C.c_int64
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


@dataclass
class Geometry:
    """Voxel grid of one config.  `pc_range` fp32[6], `voxel_size` python floats (as in the YAML)."""
    pc_range: np.ndarray
    voxel_size: tuple

    def __post_init__(self):
        self.pc_range = np.asarray(self.pc_range, dtype=np.float32)
        self.voxel_size = tuple(float(v) for v in self.voxel_size)

    @property
    def grid(self) -> np.ndarray:
        g = np.zeros(3, dtype=np.int32)
        vs = np.asarray(self.voxel_size, dtype=np.float64)
        lib().orc_grid_size(_p(self.pc_range), _p(vs), _p(g))
        return g

    @property
    def vsize_f32(self) -> np.ndarray:
        return np.asarray(self.voxel_size, dtype=np.float32)

    def centre_offsets(self):
        """pillar_vfe.py:79-81 evaluated the way the reference evaluates it: python float / 2 + range[j]
        where range[j] is whatever scalar type the caller's point_cloud_range holds (np.float32 here)."""
        return tuple(float(np.float32(self.voxel_size[j] / 2 + self.pc_range[j])) for j in range(3))


def set_num_threads(n: int):
    lib().orc_set_num_threads(int(n))


def mask_points_by_range(points, pc_range, xcol=0):
    pts = _f32(points)
    keep = np.zeros(pts.shape[0], dtype=np.uint8)
    lib().orc_mask_points_by_range(_p(pts), C.c_int64(pts.shape[0]), C.c_int(pts.shape[1]), C.c_int(xcol),
                                   _p(np.asarray(pc_range, dtype=np.float32)), _p(keep))
    return keep.astype(bool)


def voxelize(points, geom: Geometry, P: int, max_voxels: int, F: int | None = None, xcol: int = 0,
             return_point_pillar: bool = False, spconv1_break: bool = False):
    """One frame.  points [n, stride] fp32.  Returns voxels [M,P,F], coords [M,3] int32 (z,y,x), num [M] int32.
    spconv1_break: spconv 1.x overflow semantics (the loop stops at the first point that would open pillar max_voxels + 1)."""
    pts = _f32(points)
    n, stride = pts.shape
    F = stride - xcol if F is None else F
    grid = geom.grid
    lookup = np.full(int(grid[0]) * int(grid[1]) * int(grid[2]), -1, dtype=np.int32)
    cap = max(1, min(n, max_voxels))
    voxels = np.zeros((cap, P, F), dtype=np.float32)
    coords = np.zeros((cap, 3), dtype=np.int32)
    num = np.zeros(cap, dtype=np.int32)
    pp = np.zeros(max(n, 1), dtype=np.int32) if return_point_pillar else None
    m = lib().orc_voxelize_mode(_p(pts), C.c_int64(n), C.c_int(stride), C.c_int(xcol), C.c_int(F),
                                _p(geom.pc_range), _p(geom.vsize_f32), _p(grid), C.c_int(P), C.c_int(max_voxels),
                                _p(lookup), _p(voxels), _p(coords), _p(num), _p(pp), C.c_int(int(spconv1_break)))
    assert (lookup == -1).all()
    out = (voxels[:m].copy(), coords[:m].copy(), num[:m].copy())
    return out + (pp[:n],) if return_point_pillar else out


def voxelize_py(points, geom: Geometry, P: int, max_voxels: int, F: int | None = None, xcol: int = 0, spconv1_break: bool = False):
    """Independent pure-Python twin of `voxelize` (slow; small cases only)."""
    pts = _f32(points)
    n, stride = pts.shape
    F = stride - xcol if F is None else F
    grid = [int(g) for g in geom.grid]
    rng, vs = geom.pc_range, geom.vsize_f32
    table = {}
    voxels, coords, num = [], [], []
    for i in range(n):
        c = []
        for j in range(3):
            q = np.float32(np.float32(pts[i, xcol + j] - rng[j]) / vs[j])
            f = np.floor(q)
            if not (f >= 0 and f < grid[j]):
                c = None
                break
            c.append(int(f))
        if c is None:
            continue
        key = (c[2], c[1], c[0])
        v = table.get(key, -1)
        if v == -1:
            if len(coords) >= max_voxels:
                if spconv1_break:
                    break
                continue
            v = len(coords)
            table[key] = v
            coords.append(key)
            num.append(0)
            voxels.append(np.zeros((P, F), dtype=np.float32))
        if num[v] < P:
            voxels[v][num[v]] = pts[i, xcol:xcol + F]
            num[v] += 1
    M = len(coords)
    return (np.stack(voxels) if M else np.zeros((0, P, F), np.float32),
            np.asarray(coords, dtype=np.int32).reshape(M, 3), np.asarray(num, dtype=np.int32))


@dataclass
class PfnParams:
    """Weights of the single last-layer PFN (pillar_vfe.py:8-27)."""
    weight: np.ndarray                 # [C, Cin]  pfn_layers.0.linear.weight
    gamma: np.ndarray | None = None    # norm.weight
    beta: np.ndarray | None = None     # norm.bias
    running_mean: np.ndarray | None = None
    running_var: np.ndarray | None = None
    bias: np.ndarray | None = None     # linear.bias when USE_NORM is False
    eps: float = 1e-3


def _cfg(geom: Geometry, F, P, Cout, use_absolute_xyz, with_distance, use_norm, eps):
    xo, yo, zo = geom.centre_offsets()
    vs = geom.vsize_f32
    return _VfeCfg(F, P, Cout, int(use_absolute_xyz), int(with_distance), int(use_norm),
                   float(vs[0]), float(vs[1]), float(vs[2]), xo, yo, zo, float(eps))


def pillar_vfe(voxels, coords_bzyx, num_points, geom: Geometry, pfn: PfnParams,
               use_absolute_xyz=True, with_distance=False, invstd_override=None):
    vox = _f32(voxels)
    M, P, F = vox.shape
    W = _f32(pfn.weight)
    Cout = W.shape[0]
    use_norm = pfn.gamma is not None
    cfg = _cfg(geom, F, P, Cout, use_absolute_xyz, with_distance, use_norm, pfn.eps)
    out = np.zeros((M, Cout), dtype=np.float32)
    opt = lambda a: None if a is None else _f32(a)
    keep = [opt(pfn.bias), opt(pfn.gamma), opt(pfn.beta), opt(pfn.running_mean), opt(pfn.running_var)]
    co, nu = _f32(coords_bzyx), _f32(num_points)
    inv = opt(invstd_override)
    lib().orc_pillar_vfe(C.byref(cfg), C.c_int64(M), _p(vox), _p(co), _p(nu), _p(W), *[_p(k) for k in keep],
                         _p(inv), _p(out))
    return out


def pointpillar_scatter(pillar_features, coords_bzyx, B, C_out, ny, nx):
    pf, co = _f32(pillar_features), _f32(coords_bzyx)
    canvas = np.empty((B, C_out, ny, nx), dtype=np.float32)
    lib().orc_pointpillar_scatter(C.c_int64(pf.shape[0]), C.c_int(C_out), C.c_int(B), C.c_int(ny), C.c_int(nx),
                                  _p(pf), _p(co), _p(canvas))
    return canvas


def points_to_bev(points, frame_offsets, geom: Geometry, pfn: PfnParams, P: int, max_voxels: int,
                  F: int | None = None, xcol: int = 0, use_absolute_xyz=True, with_distance=False,
                  want_canvas=True, want_voxels=True):
    """Whole path on a batch.  points [n_total, stride] fp32, frame_offsets [B+1].
    Returns dict(voxels, voxel_coords [M,4] int32 (b,z,y,x), voxel_num_points [M] int32,
                 pillar_features [M,C], spatial_features [B,C,ny,nx], frame_pillars [B])."""
    pts = _f32(points)
    stride = pts.shape[1]
    F = stride - xcol if F is None else F
    fo = np.ascontiguousarray(frame_offsets, dtype=np.int64)
    B = fo.shape[0] - 1
    grid = geom.grid
    W = _f32(pfn.weight)
    Cout = W.shape[0]
    use_norm = pfn.gamma is not None
    cfg = _cfg(geom, F, P, Cout, use_absolute_xyz, with_distance, use_norm, pfn.eps)
    cap = int(sum(min(int(fo[b + 1] - fo[b]), max_voxels) for b in range(B)))
    cap = max(cap, 1)
    voxels = np.empty((cap, P, F), dtype=np.float32)
    coords = np.zeros((cap, 4), dtype=np.int32)
    num = np.zeros(cap, dtype=np.int32)
    feat = np.zeros((cap, Cout), dtype=np.float32)
    canvas = np.empty((B, Cout, int(grid[1]), int(grid[0])), dtype=np.float32) if want_canvas else None
    fp = np.zeros(max(B, 1), dtype=np.int32)
    opt = lambda a: None if a is None else _f32(a)
    keep = [opt(pfn.bias), opt(pfn.gamma), opt(pfn.beta), opt(pfn.running_mean), opt(pfn.running_var)]
    M = lib().orc_points_to_bev(_p(pts), _p(fo), C.c_int(B), C.c_int(stride), C.c_int(xcol), C.byref(cfg),
                                _p(geom.pc_range), _p(geom.vsize_f32), _p(grid), C.c_int(max_voxels),
                                _p(W), *[_p(k) for k in keep],
                                _p(voxels), _p(coords), _p(num), _p(feat), _p(canvas), _p(fp))
    return dict(voxels=voxels[:M] if want_voxels else None, voxel_coords=coords[:M], voxel_num_points=num[:M],
                pillar_features=feat[:M], spatial_features=canvas, frame_pillars=fp[:B], num_pillars=int(M))


# ---- stacked PFN (numpy restatement; test infrastructure) -------------------------------------------------------------------
def _fma_rows(x, w):
    """x [..., K] fp32, w [C, K] fp32 -> [..., C]: acc = fmaf(x[k], w[c][k], acc) for k ascending, as the kernels evaluate nn.Linear
    (pillar_vfe.py:37).  The product of two fp32 values is exact in float64; the sum is rounded to float64 and then to fp32 (a
    double rounding that differs from a true fma in rare last-bit cases: the stacked comparison carries a 1e-5 tolerance anyway)."""
    acc = np.zeros(x.shape[:-1] + (w.shape[0],), dtype=np.float32)
    for k in range(x.shape[-1]):
        acc = (acc.astype(np.float64) + x[..., k:k + 1].astype(np.float64) * w[:, k].astype(np.float64)).astype(np.float32)
    return acc


def _bn_relu(x, p: PfnParams):
    if p.gamma is None:
        y = x + _f32(p.bias)
    else:
        inv = (np.float32(1.0) / np.sqrt(_f32(p.running_var) + np.float32(p.eps))).astype(np.float32)
        y = ((x - _f32(p.running_mean)) * inv) * _f32(p.gamma) + _f32(p.beta)      # four roundings, as BatchNorm1d eval does (:39)
    return np.maximum(y, np.float32(0)).astype(np.float32)


def pillar_vfe_layers(voxels, coords_bzyx, num_points, geom: Geometry, layers, use_absolute_xyz=True, with_distance=False):
    """PillarVFE.forward (pillar_vfe.py:94-123) with a STACK of PFN layers (`layers`: list of PfnParams; every layer but the last
    concatenates each slot's features with the pillar's max, PFNLayer.forward :29-49).  Plain numpy, float32 throughout."""
    vox = _f32(voxels)
    M, P, F = vox.shape
    num = _f32(num_points)
    co = _f32(coords_bzyx)
    xo, yo, zo = (np.float32(v) for v in geom.centre_offsets())
    vs = geom.vsize_f32
    # torch CPU sum(dim=1): four interleaved partial sums over the leading 4*floor(P/4) slots, the tail into partial 0
    P4 = (P // 4) * 4
    acc = [np.zeros((M, 3), np.float32) for _ in range(4)]
    for s in range(P):
        q = (s % 4) if s < P4 else 0
        acc[q] = acc[q] + vox[:, s, :3]
    mean = (((acc[0] + acc[1]) + acc[2]) + acc[3]) / num[:, None]
    f_cluster = vox[:, :, :3] - mean[:, None, :]
    centre = np.stack([co[:, 3] * vs[0] + xo, co[:, 2] * vs[1] + yo, co[:, 1] * vs[2] + zo], axis=1).astype(np.float32)
    f_center = vox[:, :, :3] - centre[:, None, :]
    feats = [vox if use_absolute_xyz else vox[:, :, 3:], f_cluster, f_center]
    if with_distance:
        x, y, z = vox[:, :, 0], vox[:, :, 1], vox[:, :, 2]
        feats.append(np.sqrt((z.astype(np.float64) * z + (y.astype(np.float64) * y + (x * x).astype(np.float32)).astype(np.float32))
                             .astype(np.float32))[:, :, None].astype(np.float32))
    x = np.concatenate(feats, axis=2).astype(np.float32)
    mask = (num.astype(np.int32)[:, None] > np.arange(P)[None, :]).astype(np.float32)
    x = x * mask[:, :, None]
    for li, p in enumerate(layers):
        h = _bn_relu(_fma_rows(x, _f32(p.weight)), p)
        hmax = h.max(axis=1, keepdims=True)
        if li == len(layers) - 1:
            return hmax[:, 0, :]
        x = np.concatenate([h, np.repeat(hmax, P, axis=1)], axis=2)
