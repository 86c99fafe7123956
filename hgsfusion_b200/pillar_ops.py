"""Path B -- the PillarNet reader ops, mirroring the reference's `pcdet/ops/pillar_ops` Python surface on top of the
native library (no `pillar_cuda`, no per-call host syncs inside the index generation).

  reference                                              here
  pillar_utils.bev_spatial_shape              :16-19     bev_spatial_shape
  pillar_utils.gen_indice_pairs (GenIndicePairs) :84-132 + group_utils.flatten_indices :12-31   gen_indice_pairs_flat
  group_utils.gather_feature  (GatherFeature)  :34-58    gather_feature      (autograd, native backward)
  scatter_utils.scatter_max   (ScatterMaxFunction) :7-45 scatter_max         (autograd, native backward)
  pillar_utils.PillarQueryAndGroup              :22-54   PillarQueryAndGroup
  pillar_modules.PillarMaxPooling               :10-82   PillarMaxPooling    (same parameter names: shared_mlps.0.weight, shared_mlps.1.*)
  vfe/pillarnet.py:51-58 (per-frame split) + pillarnet_modules/dynamic_pillar_encoder.py:55-118   split_encode
  dynamic_pillar_encoder.DynamicPillarFeatureNet :9-121  DynamicPillarFeatureNet (forward takes the collated points or the reference's list)

`PillarMaxPooling.forward` returns `(pillar_features [M,C], pillars [M,3] (b,y,x), spatial_shape (Ny,Nx), batch_size)`
and wraps them in `spconv.SparseConvTensor` only when spconv is importable (pillar_modules.py:82).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib


def _s(t=None):
    """The current stream of tensor `t`'s device (of the current device without a tensor)."""
    return C.c_void_p(torch.cuda.current_stream(None if t is None else t.device).cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _on_device(fn):
    """Run `fn` with the device of its first CUDA tensor argument current: the native library launches on the current device
    and `_s()` then returns that device's current stream (the reference's ops have no device guard at all, cuda_utils.h)."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapped


def _need_cuda(t, name, dtype):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (hgsfusion_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")     # the reference's CHECK_INPUT, cuda_utils.h:19-21


def bev_spatial_shape(point_cloud_range, pillar_size):
    H = round((point_cloud_range[3] - point_cloud_range[0]) / pillar_size)
    W = round((point_cloud_range[4] - point_cloud_range[1]) / pillar_size)
    return int(W), int(H)          # the reference's quirk: (W, H) is then unpacked as `H, W = spatial_shape` (:102)


@_on_device
@torch.no_grad()
def gen_indice_pairs_flat(xyz, xyz_batch_cnt, pillar_size, spatial_shape, sync=True):
    """xyz [N,3] relative coords, xyz_batch_cnt [B] int32 -> dict(pillars [M,3], pillar_bev_indices [B,H,W],
    indice_pairs [N,1], point_set_indices [L], pillar_set_indices [L]).  One native launch; `sync=True` reads
    (M, L) once to slice the outputs as the reference does, `sync=False` leaves them capacity-sized with `counts`."""
    _need_cuda(xyz, "xyz", torch.float32)
    _need_cuda(xyz_batch_cnt, "xyz_batch_cnt", torch.int32)
    assert xyz.shape[1] == 3
    lib = _lib.load()
    N, B = int(xyz.shape[0]), int(xyz_batch_cnt.numel())
    H, W = spatial_shape
    dev = xyz.device
    cap = max(1, min(N, B * H * W))
    bev = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    pillars = torch.empty((cap, 3), dtype=torch.int32, device=dev)
    pairs = torch.empty((N, 1), dtype=torch.int32, device=dev)
    first = torch.empty((max(N, 1),), dtype=torch.int32, device=dev)
    second = torch.empty((max(N, 1),), dtype=torch.int32, device=dev)
    counts = torch.empty((2,), dtype=torch.int32, device=dev)
    need = C.c_size_t(0)
    _lib.check(lib.hgsf_pillarnet_workspace_size(N, C.byref(need)), "hgsf_pillarnet_workspace_size")
    ws = torch.empty(need.value + 256, dtype=torch.uint8, device=dev)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    st = lib.hgsf_pillarnet_indices(float(pillar_size), _p(xyz), _p(xyz_batch_cnt), N, B, H, W, _p(bev), _p(pillars), _p(pairs),
                                    _p(first), _p(second), _p(counts), C.c_void_p(ws_ptr), need.value, _s())
    _lib.check(st, "hgsf_pillarnet_indices")
    out = dict(pillar_bev_indices=bev, indice_pairs=pairs, counts=counts)
    if sync:
        M, L = (int(v) for v in counts.tolist())
        out.update(pillars=pillars[:M], point_set_indices=first[:L], pillar_set_indices=second[:L])
    else:
        out.update(pillars=pillars, point_set_indices=first, pillar_set_indices=second)
    return out


class GatherFeature(Function):
    @staticmethod
    @_on_device
    def forward(ctx, features: torch.Tensor, set_indices: torch.Tensor):
        _need_cuda(features, "features", torch.float32)
        _need_cuda(set_indices, "set_indices", torch.int32)
        out = features.new_empty((set_indices.shape[0], features.shape[1]))
        _lib.check(_lib.load().hgsf_gather_feature(_p(set_indices), _p(features), int(set_indices.shape[0]),
                                                  int(features.shape[1]), _p(out), _s()), "hgsf_gather_feature")
        ctx.for_backwards = (features.shape[0], features.shape[1], set_indices)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, grad_out):
        N, Cc, set_indices = ctx.for_backwards
        grad_features = grad_out.new_zeros((N, Cc))
        g = grad_out.contiguous()
        _lib.check(_lib.load().hgsf_gather_feature_grad(_p(set_indices), _p(g), int(set_indices.shape[0]), int(Cc),
                                                       _p(grad_features), _s()), "hgsf_gather_feature_grad")
        return grad_features, None


gather_feature = GatherFeature.apply


class ScatterMaxFunction(Function):
    @staticmethod
    @_on_device
    def forward(ctx, src: torch.Tensor, index: torch.Tensor, M: int):
        """src (C, L), index (L,) -> out (C, M)"""
        _need_cuda(src, "src", torch.float32)
        _need_cuda(index, "index", torch.int32)
        Cc, L = src.size()
        arg = torch.empty((Cc, M), dtype=torch.int32, device=src.device)
        out = src.new_empty((Cc, M))
        _lib.check(_lib.load().hgsf_scatter_max(_p(index), _p(src), int(Cc), int(L), int(M), _p(arg), _p(out), _s()),
                   "hgsf_scatter_max")
        ctx.for_backwards = (Cc, L, arg)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, grad_out):
        Cc, L, arg = ctx.for_backwards
        grad_src = grad_out.new_zeros((Cc, L))
        g = grad_out.contiguous()
        _lib.check(_lib.load().hgsf_scatter_max_grad(_p(arg), _p(g), int(Cc), int(arg.shape[1]), _p(grad_src), _s()),
                   "hgsf_scatter_max_grad")
        return grad_src, None, None


scatter_max = ScatterMaxFunction.apply


class PillarQueryAndGroup(nn.Module):
    def __init__(self, pillar_size, point_cloud_range):
        super().__init__()
        self.pillar_size = pillar_size
        self.spatial_shape = bev_spatial_shape(point_cloud_range, pillar_size)
        self.z_center = (point_cloud_range[5] + point_cloud_range[2]) / 2
        self.point_cloud_range = point_cloud_range

    def forward(self, xyz, xyz_batch_cnt, point_features):
        """xyz (N,3) relative, xyz_batch_cnt (B,), point_features (N,C) ->
        pillars (M,3) [b y x], pillar_set_indices (L,), group_features (L, C+6)   (pillar_utils.py:31-54)"""
        r = gen_indice_pairs_flat(xyz, xyz_batch_cnt, self.pillar_size, self.spatial_shape)
        pillars, point_set_indices, pillar_set_indices = r["pillars"], r["point_set_indices"], r["pillar_set_indices"]
        self.last_pillar_bev_indices = r["pillar_bev_indices"]      # the cell table, for the pillar-list consumer (subm_neighbors)
        # pillar centres as the reference builds them (pillar_utils.py:117-121): z centre in absolute coordinates
        pillar_centers = torch.zeros([pillars.shape[0], 3], dtype=torch.float32, device=xyz.device)
        pillar_centers[:, 0] = (pillars[:, 2] + 0.5) * self.pillar_size
        pillar_centers[:, 1] = (pillars[:, 1] + 0.5) * self.pillar_size
        pillar_centers[:, 2] = self.z_center
        group_point_features = gather_feature(point_features, point_set_indices)
        group_point_xyz = gather_feature(xyz, point_set_indices)
        group_pillar_centers = gather_feature(pillar_centers, pillar_set_indices)
        group_pillar_centers = group_point_xyz - group_pillar_centers
        group_features = torch.cat([group_point_features.detach(), group_point_xyz.detach(),
                                    group_pillar_centers.detach()], dim=1)
        return pillars, pillar_set_indices, group_features


class PillarMaxPooling(nn.Module):
    """pillar_modules.py:10-82.  After a forward, `pillar_bev_indices` holds the [B,H,W] cell table (pillar id per cell, -1
    none) of that call -- what PillarEncoderConv1 / PillarEncoderConv2 need next to the returned features and indices."""

    @property
    def pillar_bev_indices(self):
        return getattr(self.groups, "last_pillar_bev_indices", None)

    def __init__(self, mlps: List[int], pillar_size: float, point_cloud_range: List[float]):
        super().__init__()
        self.bev_width, self.bev_height = bev_spatial_shape(point_cloud_range, pillar_size)
        self.groups = PillarQueryAndGroup(pillar_size, point_cloud_range)
        shared_mlp = []
        for k in range(len(mlps) - 1):
            shared_mlp.extend([nn.Linear(mlps[k], mlps[k + 1], bias=False),
                               nn.BatchNorm1d(mlps[k + 1], eps=1e-3, momentum=0.01), nn.ReLU()])
        self.shared_mlps = nn.Sequential(*shared_mlp)       # the reference's own cuBLAS MLP (pillar_modules.py:19-26)

    def _fused_ok(self, pt_feature) -> bool:
        """Eval without gradients and the single 32-channel layer every shipped config uses (hgsfusion_vod.yaml:107-108):
        the whole reader is one native launch after the index generation."""
        if self.training or len(self.shared_mlps) != 3:
            return False
        lin = self.shared_mlps[0]
        if lin.out_features != 32 or lin.in_features > 40 or lin.in_features != pt_feature.shape[1] + 6:
            return False
        return not (torch.is_grad_enabled() and any(p.requires_grad for p in self.shared_mlps.parameters()))

    @torch.no_grad()
    @_on_device
    def _forward_fused(self, xyz, xyz_batch_cnt, pt_feature):
        from .ops import PfnWeights
        _need_cuda(pt_feature, "pt_feature", torch.float32)
        g = self.groups
        r = gen_indice_pairs_flat(xyz, xyz_batch_cnt, g.pillar_size, g.spatial_shape)
        pillars, point_idx, pillar_idx = r["pillars"], r["point_set_indices"], r["pillar_set_indices"]
        g.last_pillar_bev_indices = r["pillar_bev_indices"]         # the cell table, for the pillar-list consumer (subm_neighbors)
        lin, bn = self.shared_mlps[0], self.shared_mlps[1]
        pfn = PfnWeights(weight=lin.weight.detach(), bn_weight=bn.weight.detach(), bn_bias=bn.bias.detach(),
                         running_mean=bn.running_mean, running_var=bn.running_var, eps=bn.eps)
        pf = pfn.to_struct()
        M, L = int(pillars.shape[0]), int(point_idx.shape[0])
        out = torch.empty((M, 32), dtype=torch.float32, device=xyz.device)
        st = _lib.load().hgsf_pillarnet_reader(_p(xyz), _p(pt_feature), int(pt_feature.shape[1]), _p(point_idx), _p(pillar_idx), L,
                                               _p(pillars), M, float(g.pillar_size), float(g.z_center), C.byref(pf), _p(out), _s())
        _lib.check(st, "hgsf_pillarnet_reader")
        return pillars, out

    def forward(self, xyz, xyz_batch_cnt, pt_feature):
        B = xyz_batch_cnt.shape[0]
        if self._fused_ok(pt_feature):
            pillar_indices, pillar_features = self._forward_fused(xyz, xyz_batch_cnt, pt_feature)
        else:
            pillar_indices, pillar_set_indices, group_features = self.groups(xyz, xyz_batch_cnt, pt_feature)
            group_features = self.shared_mlps(group_features)
            group_features = group_features.transpose(1, 0).contiguous()
            pillar_features = scatter_max(group_features, pillar_set_indices, pillar_indices.shape[0])
            pillar_features = pillar_features.transpose(1, 0)
        try:
            try:
                import spconv.pytorch as spconv
            except ImportError:
                import spconv
            # (bev_width, bev_height) as the reference passes it (pillar_modules.py:82): bev_spatial_shape returns (W, H) =
            # (Ny, Nx), so this is (Ny, Nx) and matches the (b, y, x) indices -- on TJ4D (496, 432)
            return spconv.SparseConvTensor(pillar_features, pillar_indices, (self.bev_width, self.bev_height), B)
        except ImportError:
            return pillar_features, pillar_indices, (self.bev_width, self.bev_height), B


@_on_device
def sparse_to_dense(features, indices, spatial_shape, batch_size: int):
    """`SparseConvTensor(features [M,C], indices [M,3] int32 (b, y, x), spatial_shape (Ny, Nx), batch_size).dense()` ->
    [B, C, Ny, Nx], as the PillarNet branch calls it on its backbone outputs (pillarnet_modules/lss_fpn.py:111-113,
    rpn.py:243-247) -- one pass over the dense tensor through hgsf_sparse_to_dense (zeros included).  Forward only."""
    _need_cuda(features, "features", torch.float32)
    _need_cuda(indices, "indices", torch.int32)
    if features.dim() != 2 or indices.dim() != 2 or indices.shape[1] != 3 or indices.shape[0] != features.shape[0]:
        raise ValueError("features [M,C] and indices [M,3] expected")
    lib = _lib.load()
    M, Cc = int(features.shape[0]), int(features.shape[1])
    ny, nx = int(spatial_shape[0]), int(spatial_shape[1])
    need = C.c_size_t(0)
    _lib.check(lib.hgsf_sparse_to_dense_workspace_size(int(batch_size), ny, nx, C.byref(need)), "hgsf_sparse_to_dense_workspace_size")
    ws = torch.empty(need.value + 256, dtype=torch.uint8, device=features.device)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    dense = torch.empty((int(batch_size), Cc, ny, nx), dtype=torch.float32, device=features.device)
    st = lib.hgsf_sparse_to_dense(_p(features), _p(indices), M, Cc, int(batch_size), ny, nx, C.c_void_p(ws_ptr), need.value,
                                  _p(dense), _s())
    _lib.check(st, "hgsf_sparse_to_dense")
    return dense


_ENCODINGS = {"split": 0, "mixed": 1, "direct": 2}
_SPLIT_N = {"vod": 12, "tj4d": 13}          # dynamic_pillar_encoder.py:72-76


@_on_device
@torch.no_grad()
def split_encode(points, pc_range, num_input, virtual=True, encoding_type="split", dataset="vod", batch_size=None):
    """Collated `points [L, 1+Fin]` (frame index in column 0) -> `(xyz [L',3], xyz_batch_cnt [B] int32, pt_features
    [L', Fout])`, what `PillarNet.forward` (vfe/pillarnet.py:51-58) and `DynamicPillarFeatureNet.forward`
    (dynamic_pillar_encoder.py:55-118) hand to the reader.  One native launch and one 8-byte read (rows kept, order
    flag) instead of `.max().item()` plus a boolean-mask pass per frame; `batch_size=None` reads it from column 0 like the
    reference, `batch_dict['batch_size']` avoids that sync."""
    _need_cuda(points, "points", torch.float32)
    lib = _lib.load()
    L, Fin = int(points.shape[0]), int(points.shape[1]) - 1
    if batch_size is None:
        batch_size = int(points[:, 0].max().item()) + 1                  # pillarnet.py:52
    if not virtual:
        mode, Fout, n = 1, Fin, 0
    elif encoding_type == "split":
        if dataset not in _SPLIT_N:
            raise NotImplementedError(dataset)                            # :77-78
        mode, Fout, n = 0, int(num_input), _SPLIT_N[dataset]
    elif encoding_type == "mixed":
        mode, Fout, n = 1, Fin, 0
    elif encoding_type == "direct":
        mode, Fout, n = 2, Fin - 2, 0
    else:
        raise NotImplementedError(encoding_type)                          # :97-98
    dev = points.device
    xyz = torch.empty((L, 3), dtype=torch.float32, device=dev)
    feat = torch.empty((L, Fout), dtype=torch.float32, device=dev)
    cnt = torch.empty((batch_size,), dtype=torch.int32, device=dev)
    info = torch.empty((2,), dtype=torch.int32, device=dev)
    pc_min = (C.c_float * 3)(*[float(v) for v in pc_range[:3]])

    def run(order):
        _lib.check(lib.hgsf_split_encode(_p(points), L, Fin, Fout, n, batch_size, mode, pc_min, _p(order), _p(xyz), _p(feat),
                                         _p(cnt), _p(info), _s()), "hgsf_split_encode")
        return [int(v) for v in info.tolist()]

    flags, kept = run(None)
    if flags & 1:
        # rows not grouped by frame (never the case for a collated batch): a stable order by frame, dropped rows last
        b = points[:, 0]
        key = torch.where((b >= 0) & (b < batch_size) & (b == b.trunc()), b, torch.full_like(b, float(batch_size)))
        order = torch.sort(key, stable=True).indices.to(torch.int32)
        flags, kept = run(order)
        assert not (flags & 1)
    return xyz[:kept], cnt, feat[:kept]


class DynamicPillarFeatureNet(nn.Module):
    """Drop-in for pillarnet_modules/dynamic_pillar_encoder.py:9-121 (same constructor, same `pfn_layers.*` parameter
    names).  `forward(example)` accepts the reference's `dict(points=[per-frame tensors])` or, without the Python split,
    `dict(points=collated [L,1+Fin], batch_size=B)`."""

    def __init__(self, num_input_features=2, num_filters=(32,), pillar_size=0.1, virtual=False,
                 pc_range=(0, -40, -3, 70.4, 40, 1), encoding_type="split", dataset="vod", **kwargs):
        super().__init__()
        self.pc_range = pc_range
        assert len(num_filters) > 0
        self.num_input = num_input_features
        self.pfn_layers = PillarMaxPooling(mlps=[6 + num_input_features] + list(num_filters), pillar_size=pillar_size,
                                           point_cloud_range=pc_range)
        self.virtual = virtual
        self.encoding_type = encoding_type
        self.dataset = dataset

    def forward(self, example, **kwargs):
        points = example.pop("points")
        batch_size = example.get("batch_size")
        if isinstance(points, (list, tuple)):
            batch_size = len(points)
            points = torch.cat([torch.cat([torch.full((p.shape[0], 1), float(i), dtype=p.dtype, device=p.device), p], dim=1)
                                for i, p in enumerate(points)], dim=0).contiguous()
        xyz, cnt, feat = split_encode(points, self.pc_range, self.num_input, self.virtual, self.encoding_type, self.dataset,
                                      batch_size)
        return self.pfn_layers(xyz, cnt, feat)


# ---------------------------------------------------------------------------------------------------------------
# The pillar-list consumer: SpMiddlePillarEncoder18.conv1 on the reader's pillar list
# ---------------------------------------------------------------------------------------------------------------

@_on_device
@torch.no_grad()
def subm_neighbors(pillar_bev_indices, pillars, num_rows_dev=None):
    """The rule book of one `indice_key` for 3x3 submanifold convolutions: [M, 9] int32, entry ky*3+kx = pillar id at
    (y+ky-1, x+kx-1) or -1.  pillar_bev_indices [B,H,W] and pillars [M,3] (b,y,x) come from gen_indice_pairs_flat /
    PillarMaxPooling.  num_rows_dev: optional device int32 row count (capacity-sized `pillars`, no host sync)."""
    _need_cuda(pillar_bev_indices, "pillar_bev_indices", torch.int32)
    _need_cuda(pillars, "pillars", torch.int32)
    B, H, W = (int(v) for v in pillar_bev_indices.shape)
    M = int(pillars.shape[0])
    nbr = torch.empty((M, 9), dtype=torch.int32, device=pillars.device)
    st = _lib.load().hgsf_subm_neighbors(_p(pillar_bev_indices), _p(pillars), M, _p(num_rows_dev), B, H, W, _p(nbr), _s())
    _lib.check(st, "hgsf_subm_neighbors")
    return nbr


@_on_device
@torch.no_grad()
def subm_conv3x3(features, neighbors, weight, bias=None, bn=None, residual=None, relu=False, weight_layout="KRSC",
                 num_rows_dev=None, out=None):
    """act(BN(SubMConv2d_3x3(features)) + residual) on the pillar list (hgsf_subm_conv3x3).  weight [Cout,3,3,Cin]
    ("KRSC", spconv 2.x) or [3,3,Cin,Cout] ("RSCK", spconv 1.x); bn = an eval-mode nn.BatchNorm1d or None."""
    _need_cuda(features, "features", torch.float32)
    _need_cuda(neighbors, "neighbors", torch.int32)
    _need_cuda(weight, "weight", torch.float32)
    if weight_layout not in ("KRSC", "RSCK") or weight.dim() != 4:
        raise ValueError("weight must be [Cout,3,3,Cin] (KRSC) or [3,3,Cin,Cout] (RSCK)")
    Cout, Cin = (int(weight.shape[0]), int(weight.shape[3])) if weight_layout == "KRSC" else (int(weight.shape[3]), int(weight.shape[2]))
    M = int(neighbors.shape[0])          # output rows; a stride-2 rule book (sparse_conv_s2_indices) has its own row count
    if features.dim() != 2 or features.shape[1] != Cin or neighbors.dim() != 2 or neighbors.shape[1] != 9:
        raise ValueError("features [M_in,Cin] and neighbors [M_out,9] expected")
    if bn is not None and bn.training:
        raise NotImplementedError("the pillar-list consumer is an inference path: BatchNorm must be in eval mode")
    if residual is not None:
        _need_cuda(residual, "residual", torch.float32)
    if out is None:
        out = torch.empty((M, Cout), dtype=torch.float32, device=features.device)
    cv = _lib.SubmConv()
    cv.weight = weight.data_ptr(); cv.weight_layout = 0 if weight_layout == "KRSC" else 1
    cv.bias = bias.data_ptr() if bias is not None else None
    if bn is not None:
        cv.bn_weight, cv.bn_bias = bn.weight.data_ptr(), bn.bias.data_ptr()
        cv.bn_mean, cv.bn_var, cv.bn_eps = bn.running_mean.data_ptr(), bn.running_var.data_ptr(), float(bn.eps)
    cv.in_channels, cv.out_channels, cv.relu = Cin, Cout, int(bool(relu))
    st = _lib.load().hgsf_subm_conv3x3(C.byref(cv), _p(features), _p(neighbors), M, _p(num_rows_dev), _p(residual), _p(out), _s())
    _lib.check(st, "hgsf_subm_conv3x3")
    return out


class SubMConv2d(nn.Module):
    """Parameter holder with spconv.SubMConv2d's names and spconv 2.x's weight layout [Cout, 3, 3, Cin] (kernel 3,
    stride 1, padding 1: what conv2D3x3 builds, pcnres18.py:82-95), so a reference checkpoint's `*.weight` / `*.bias` load."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, 3, 3, in_channels))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        if bias:
            bound = 1.0 / (9 * in_channels) ** 0.5
            nn.init.uniform_(self.bias, -bound, bound)
        self._rsck = None

    def rsck(self):
        """The weight as [3, 3, Cin, Cout], the kernel's native layout (a straight copy into shared memory; the KRSC layout
        is transposed on the fly by every CTA, ~8 % slower at config 2).  Cached until the parameter changes."""
        key = (self.weight.data_ptr(), self.weight._version)
        if self._rsck is None or self._rsck[0] != key:
            self._rsck = (key, self.weight.detach().permute(1, 2, 3, 0).contiguous())
        return self._rsck[1]


def _conv_bn(inp, out):
    return nn.Sequential(SubMConv2d(inp, out, bias=True), nn.BatchNorm1d(out, eps=1e-3, momentum=0.01))


class Sparse2DBasicBlockV(nn.Module):
    """pcnres18.py:108-151 with the same sub-module names (conv0.0 / conv0.1 ...); forward(features, neighbors) ->
    features.  Three fused launches: conv0+BN+ReLU, conv1+BN+ReLU, conv2+BN+identity+ReLU."""

    def __init__(self, inplanes, planes):
        super().__init__()
        self.conv0, self.conv1, self.conv2 = _conv_bn(inplanes, planes), _conv_bn(planes, planes), _conv_bn(planes, planes)

    def forward(self, features, neighbors, num_rows_dev=None):
        f = lambda seq, x, res: subm_conv3x3(x, neighbors, seq[0].rsck(), seq[0].bias, seq[1], residual=res, relu=True,
                                             weight_layout="RSCK", num_rows_dev=num_rows_dev)
        identity = f(self.conv0, features, None)
        out = f(self.conv1, identity, None)
        return f(self.conv2, out, identity)


class Sparse2DBasicBlock(nn.Module):
    """pcnres18.py:154-188; two fused launches."""

    def __init__(self, inplanes, planes):
        super().__init__()
        self.conv1, self.conv2 = _conv_bn(planes, planes), _conv_bn(planes, planes)

    def forward(self, features, neighbors, num_rows_dev=None):
        f = lambda seq, x, res: subm_conv3x3(x, neighbors, seq[0].rsck(), seq[0].bias, seq[1], residual=res, relu=True,
                                             weight_layout="RSCK", num_rows_dev=num_rows_dev)
        return f(self.conv2, f(self.conv1, features, None), features)


class PillarEncoderConv1(nn.Module):
    """`SpMiddlePillarEncoder18.conv1` (pcnres18.py:212-215: Sparse2DBasicBlockV(32,32) + Sparse2DBasicBlock(32,32), indice
    key "res1") on the pillar list: one rule-book launch + five fused convolution launches, no dense canvas.  The
    sub-modules are named `0` and `1` like the reference's SparseSequential, so `conv1.*` checkpoint keys load."""

    def __init__(self, planes=32):
        super().__init__()
        self.add_module("0", Sparse2DBasicBlockV(planes, planes))
        self.add_module("1", Sparse2DBasicBlock(planes, planes))

    def forward(self, pillar_features, pillars, pillar_bev_indices, num_rows_dev=None):
        nbr = subm_neighbors(pillar_bev_indices, pillars, num_rows_dev)
        x = getattr(self, "0")(pillar_features, nbr, num_rows_dev)
        return getattr(self, "1")(x, nbr, num_rows_dev)


@_on_device
@torch.no_grad()
def sparse_conv_s2_indices(pillar_bev_indices, pillars, num_rows_dev=None, sync=True):
    """Indices of `SparseConv2d(kernel 3, stride 2, padding 1)` (the first layer of conv2 / conv3 / conv4, pcnres18.py:217-221)
    from the input's cell table and pillar list: dict(pillars [Mo,3] raster order, pillar_bev_indices [B,Ho,Wo], neighbors [Mo,9]
    (input pillar ids), counts).  `sync=True` reads Mo once to slice; `sync=False` leaves the outputs capacity-sized."""
    _need_cuda(pillar_bev_indices, "pillar_bev_indices", torch.int32)
    _need_cuda(pillars, "pillars", torch.int32)
    B, H, W = (int(v) for v in pillar_bev_indices.shape)
    M = int(pillars.shape[0])
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dev = pillars.device
    cap = min(4 * M, B * Ho * Wo)
    lib = _lib.load()
    need = C.c_size_t(0)
    _lib.check(lib.hgsf_sparse_conv_s2_workspace_size(M, B, C.byref(need)), "hgsf_sparse_conv_s2_workspace_size")
    ws = torch.empty(need.value + 256, dtype=torch.uint8, device=dev)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    out_bev = torch.empty((B, Ho, Wo), dtype=torch.int32, device=dev)
    out_pillars = torch.empty((cap, 3), dtype=torch.int32, device=dev)
    nbr = torch.empty((cap, 9), dtype=torch.int32, device=dev)
    counts = torch.empty((2,), dtype=torch.int32, device=dev)
    st = lib.hgsf_sparse_conv_s2_indices(_p(pillar_bev_indices), _p(pillars), M, _p(num_rows_dev), B, H, W, _p(out_bev), _p(out_pillars),
                                         _p(counts), _p(nbr), cap, C.c_void_p(ws_ptr), need.value, _s())
    _lib.check(st, "hgsf_sparse_conv_s2_indices")
    if sync:
        Mo = int(counts[0].item())
        out_pillars, nbr = out_pillars[:Mo], nbr[:Mo]
    return dict(pillars=out_pillars, pillar_bev_indices=out_bev, neighbors=nbr, counts=counts)


class SparseConv2d(SubMConv2d):
    """Parameter holder for spconv.SparseConv2d(in, out, 3, 2, padding=1, bias=False) (pcnres18.py:218-220), weight [Cout,3,3,Cin]."""

    def __init__(self, in_channels, out_channels, bias=False):
        super().__init__(in_channels, out_channels, bias=bias)


class PillarEncoderConv2(nn.Module):
    """`SpMiddlePillarEncoder18.conv2` (pcnres18.py:217-225): SparseConv2d(32, 64, 3, 2, padding 1) + BatchNorm1d + ReLU, then two
    Sparse2DBasicBlock(64, 64) on the down-sampled active set (indice key "res2").  Sub-modules `0` .. `4` like the reference's
    SparseSequential (`2` is the parameter-free ReLU).  forward -> (features [Mo,64], pillars [Mo,3], pillar_bev_indices [B,Ho,Wo])."""

    def __init__(self, inplanes=32, planes=64):
        super().__init__()
        self.add_module("0", SparseConv2d(inplanes, planes))
        self.add_module("1", nn.BatchNorm1d(planes, eps=1e-3, momentum=0.01))
        self.add_module("2", nn.ReLU())
        self.add_module("3", Sparse2DBasicBlock(planes, planes))
        self.add_module("4", Sparse2DBasicBlock(planes, planes))

    def forward(self, features, pillars, pillar_bev_indices):
        r = sparse_conv_s2_indices(pillar_bev_indices, pillars)
        conv, bn = getattr(self, "0"), getattr(self, "1")
        x = subm_conv3x3(features, r["neighbors"], conv.rsck(), conv.bias, bn, relu=True, weight_layout="RSCK")
        nbr = subm_neighbors(r["pillar_bev_indices"], r["pillars"])
        x = getattr(self, "3")(x, nbr)
        x = getattr(self, "4")(x, nbr)
        return x, r["pillars"], r["pillar_bev_indices"]


class PillarEncoderConv3(PillarEncoderConv2):
    """`SpMiddlePillarEncoder18.conv3` (pcnres18.py:227-235): SparseConv2d(64, 128, 3, 2, padding 1) + BatchNorm1d + ReLU + two
    Sparse2DBasicBlock(128, 128) (indice key "res3"); same structure and sub-module names as conv2."""

    def __init__(self, inplanes=64, planes=128):
        super().__init__(inplanes, planes)


class PillarEncoderConv4(PillarEncoderConv2):
    """`SpMiddlePillarEncoder18.conv4` (pcnres18.py:237-245): 128 -> 256, stride 2, + two Sparse2DBasicBlock(256, 256) ("res4")."""

    def __init__(self, inplanes=128, planes=256):
        super().__init__(inplanes, planes)


class SpMiddlePillarEncoder18(nn.Module):
    """The whole sparse encoder behind the reader (pcnres18.py:200-285) on pillar lists: conv1 .. conv4 with the reference's
    attribute names, so that a reference checkpoint's `conv1.*` .. `conv4.*` keys load.  forward(pillar_features, pillars,
    pillar_bev_indices) -> [(features, pillars [M,3] (b,y,x), pillar_bev_indices [B,H,W])] for x_conv1 .. x_conv4 filtered by
    `out_indices` as the reference does (:278-281); `sparse_to_dense` turns an entry into the [B,C,H,W] tensor `.dense()` gives."""

    def __init__(self, in_planes=32, out_indices=(1, 2, 3)):
        super().__init__()
        self.out_indices = list(out_indices)
        self.conv1 = PillarEncoderConv1(in_planes)
        self.conv2 = PillarEncoderConv2(in_planes, 64)
        self.conv3 = PillarEncoderConv3(64, 128)
        self.conv4 = PillarEncoderConv4(128, 256)
        self.backbone_channels = {'x_conv1': 32, 'x_conv2': 64, 'x_conv3': 128, 'x_conv4': 256}
        self.backbone_strides = {'x_conv1': 1, 'x_conv2': 2, 'x_conv3': 4, 'x_conv4': 8}

    def forward(self, pillar_features, pillars, pillar_bev_indices):
        x1 = (self.conv1(pillar_features, pillars, pillar_bev_indices), pillars, pillar_bev_indices)
        x2 = self.conv2(*x1)
        x3 = self.conv3(*x2)
        x4 = self.conv4(*x3)
        outs = [x1, x2, x3, x4]
        return [outs[i] for i in self.out_indices]
