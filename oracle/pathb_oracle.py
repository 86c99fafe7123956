"""Numpy restatement of the reference's Path B ops (TEST INFRASTRUCTURE; see oracle/pillar_oracle.c header).

  gen_indice_pairs + flatten_indices   pcdet/ops/pillar_ops/pillar_utils.py:84-132, group_utils.py:12-31,
                                       src/pillar_ops_gpu.cu:13-117, src/group_ops_gpu.cu:9-24
  gather_feature (+grad)               group_ops_gpu.cu:42-70
  scatter_max (+arg, +grad)            scatter_ops_gpu.cu:13-58

Pinned on the GPU box against the reference's OWN kernels compiled from /root/reference
(oracle/_ref/libref_pillar_ops.so, tests/test_gpu_pillarnet.py); the restatement here is the CPU-side checker.
"""
import numpy as np


def trunc_div(x: np.ndarray, s: float) -> np.ndarray:
    """int(x / s) as the CUDA kernel evaluates it: fp32 divide, cvt.rzi.s32.f32 (truncate, saturate, NaN -> 0)."""
    q = (x.astype(np.float32) / np.float32(s)).astype(np.float32)
    q = np.where(np.isnan(q), np.float32(0), q)
    return np.clip(np.trunc(q.astype(np.float64)), -2 ** 31, 2 ** 31 - 1).astype(np.int64).astype(np.int32)


def gen_indice_pairs_flat(xyz, cnt, bev_size, H, W):
    xyz = np.asarray(xyz, dtype=np.float32)
    cnt = np.asarray(cnt, dtype=np.int64)
    N, B = xyz.shape[0], cnt.shape[0]
    incl = np.cumsum(cnt)
    # frame of point p: first b with p < incl[b] among b < B-1, else B-1 (pillar_ops_gpu.cu:22-27)
    bid = np.full(N, B - 1, dtype=np.int64)
    if B > 1 and N:
        first = np.searchsorted(incl[:-1], np.arange(N), side="right")
        bid = np.minimum(first, B - 1)
    xid, yid = trunc_div(xyz[:, 0], bev_size), trunc_div(xyz[:, 1], bev_size)
    ok = ~((xid < 0) | (xid >= W) | (yid < 0) | (yid >= H))
    mask = np.zeros((B, H, W), dtype=bool)
    mask[bid[ok], yid[ok], xid[ok]] = True
    loc = np.cumsum(mask.reshape(-1)).astype(np.int32)
    M = int(loc[-1]) if loc.size else 0
    bev = (loc.reshape(B, H, W) * mask - 1).astype(np.int32)
    pillars = np.argwhere(mask).astype(np.int32)                 # raster order (b, y, x)
    pairs = np.full(N, -1, dtype=np.int32)
    pairs[ok] = bev[bid[ok], yid[ok], xid[ok]]
    valid = pairs > -1
    point_idx = np.flatnonzero(valid).astype(np.int32)
    return dict(pillars=pillars, pillar_bev_indices=bev, indice_pairs=pairs.reshape(N, 1),
                point_set_indices=point_idx, pillar_set_indices=pairs[valid], M=M, L=int(valid.sum()))


def gather_feature(features, idx):
    return np.asarray(features)[np.asarray(idx)]


def gather_feature_grad(idx, grad_out, N):
    g = np.zeros((N, grad_out.shape[1]), dtype=np.float64)
    np.add.at(g, np.asarray(idx), grad_out.astype(np.float64))
    return g.astype(np.float32)


def scatter_max(src, index, M):
    """src [C,L], index [L] -> out [C,M] = max(0, max src)."""
    src = np.asarray(src, dtype=np.float32)
    out = np.zeros((src.shape[0], M), dtype=np.float32)
    np.maximum.at(out, (np.arange(src.shape[0])[:, None], np.asarray(index)[None, :]), np.where(np.isnan(src), -np.inf, src))
    return out


def check_arg(arg, src, index, out):
    """arg is valid iff every non-negative entry points at an element of the right pillar within 1e-5 of the max,
    and it is -1 exactly where no element qualifies."""
    Cc, L = src.shape
    M = out.shape[1]
    ok = True
    qualifies = np.abs(src - out[:, index]) < 1e-5                  # [C, L]
    has = np.zeros((Cc, M), dtype=bool)
    np.logical_or.at(has, (np.arange(Cc)[:, None], np.asarray(index)[None, :]), qualifies)
    ok &= np.array_equal(arg >= 0, has)
    c, m = np.nonzero(arg >= 0)
    a = arg[c, m]
    ok &= bool(np.all(a // L == c)) and bool(np.all(index[a % L] == m)) and bool(np.all(qualifies[c, a % L]))
    return ok
