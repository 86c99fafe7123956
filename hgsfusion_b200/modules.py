"""Drop-in modules for OpenPCDet / HGSFusion's registries, backed by libhgsfusion_b200.so.

They keep the reference's plugin API (constructor keywords, `forward(batch_dict) -> batch_dict`,
`get_output_feature_dim`, `num_bev_features`) and parameter names, so reference checkpoints load:

  PillarVFE            <-> pcdet/models/backbones_3d/vfe/pillar_vfe.py:52-123   (VFE.NAME: PillarVFE)
  Radar7PillarVFE      <-> pcdet/models/backbones_3d/vfe/pillar_vfe.py:125-271  (VFE.NAME: Radar7PillarVFE)
  PointPillarScatter   <-> pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:5-41
  FusedPillarVFE       points -> pillars -> PillarVFE -> canvas in one native call; used with
                       DATA_PROCESSOR `transform_points_to_voxels_placeholder`
                       (pcdet/datasets/processor/data_processor.py:107-115) so that no CPU voxelizer runs
  PillarScatterPassthrough   MAP_TO_BEV plugin for FusedPillarVFE: spatial_features already exists

Training works through the same modules: in train mode the PFN's BatchNorm1d runs on batch statistics (zero-padded
rows included, running statistics updated) and `linear.weight` / `norm.weight` / `norm.bias` receive gradients from
native backward kernels (csrc/train_ops.cu) -- what torch autograd computes over PFNLayer.forward
(pillar_vfe.py:29-49) and PointPillarScatter.forward (pointpillar_scatter.py:33-35) in the reference.
There is no PyTorch fallback: without the CUDA library these modules raise at construction.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .ops import PfnWeights, PillarPath


class _PillarVFEFunction(torch.autograd.Function):
    """PFNLayer.forward (pillar_vfe.py:29-49) on the contract tensors, differentiable w.r.t. the layer's parameters.
    The inputs (voxels, coords, counts) are data: the reference's voxels tensor does not require grad either."""

    @staticmethod
    def forward(ctx, path, voxels, coords, num, weight, gamma, beta, bias, running_mean, running_var, eps, momentum,
                batch_stats, use_absolute_xyz, with_distance):
        kw = dict(eps=eps, use_absolute_xyz=use_absolute_xyz, with_distance=with_distance)
        w = weight.detach().contiguous()
        stats = None
        if gamma is not None:
            mean, var = running_mean, running_var
            if batch_stats:
                probe = PfnWeights(weight=w, bn_weight=gamma.detach(), bn_bias=beta.detach(), running_mean=running_mean,
                                   running_var=running_var, **kw)
                mean, var, stats = path.pillar_vfe_batch_stats(voxels, coords, num, probe, momentum, running_mean, running_var)
            pfn = PfnWeights(weight=w, bn_weight=gamma.detach(), bn_bias=beta.detach(), running_mean=mean, running_var=var, **kw)
        else:
            pfn = PfnWeights(weight=w, bias=bias.detach(), **kw)
        out = path.pillar_vfe(voxels, coords, num, pfn)
        ctx.path, ctx.pfn, ctx.stats = path, pfn, stats
        ctx.save_for_backward(voxels, coords, num)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        voxels, coords, num = ctx.saved_tensors
        dW, dg, db = ctx.path.pillar_vfe_backward(voxels, coords, num, ctx.pfn, grad_out.contiguous(), ctx.stats)
        bn = ctx.pfn.bn_weight is not None
        return (None, None, None, None, dW, dg if bn else None, db if bn else None, None if bn else db,
                None, None, None, None, None, None, None)


class _ScatterFunction(torch.autograd.Function):
    """PointPillarScatter.forward (pointpillar_scatter.py:14-41); backward = gather at the pillar cells."""

    @staticmethod
    def forward(ctx, path, pillar_features, coords, batch_size):
        ctx.path, ctx.batch_size = path, batch_size
        ctx.save_for_backward(coords)
        return path.pointpillar_scatter(pillar_features, coords, batch_size)

    @staticmethod
    def backward(ctx, grad_canvas):
        (coords,) = ctx.saved_tensors
        return None, ctx.path.pointpillar_scatter_backward(grad_canvas.contiguous(), coords, ctx.batch_size), None, None


class _FusedTrainFunction(torch.autograd.Function):
    """Train-mode FusedPillarVFE on the batch statistics: points -> (pillar_features, spatial_features) in three launches,
    backward in three, no host round trip in either (hgsf_points_to_bev_train / _backward).  Differentiable w.r.t.
    linear.weight, norm.weight, norm.bias; the points are data."""

    @staticmethod
    def forward(ctx, path, points, batch_size, weight, gamma, beta, running_mean, running_var, eps, momentum,
                use_absolute_xyz, with_distance):
        kw = dict(eps=eps, use_absolute_xyz=use_absolute_xyz, with_distance=with_distance)
        w, g, b = weight.detach().contiguous(), gamma.detach(), beta.detach()
        probe = PfnWeights(weight=w, bn_weight=g, bn_bias=b, running_mean=running_mean, running_var=running_var, **kw)
        res, mean, var, stats = path.points_to_bev_train(points, batch_size, probe, momentum, running_mean, running_var)
        ctx.path, ctx.batch_size = path, batch_size
        ctx.pfn = PfnWeights(weight=w, bn_weight=g, bn_bias=b, running_mean=mean, running_var=var, **kw)
        # only what the backward reads, and through save_for_backward: keeping `res` itself on ctx would tie the differentiable
        # outputs to their own grad_fn (a reference cycle: every step's 1 GB of outputs would wait for the cyclic GC)
        ctx.save_for_backward(res.voxels, res.voxel_coords, res.voxel_num_points, res.num_pillars, stats)
        ctx.set_materialize_grads(False)             # an unused output's cotangent stays None (no capacity-sized zero fill)
        for t in (res.voxel_coords, res.voxel_num_points, res.num_pillars, res.voxels):
            ctx.mark_non_differentiable(t)
        return res.pillar_features, res.spatial_features, res.voxel_coords, res.voxel_num_points, res.num_pillars, res.voxels

    @staticmethod
    def backward(ctx, g_feats, g_canvas, *unused):
        if g_feats is None and g_canvas is None:
            return (None,) * 12
        voxels, coords, num, counts, stats = ctx.saved_tensors
        res = SimpleNamespace(voxels=voxels, voxel_coords=coords, voxel_num_points=num, num_pillars=counts)
        dW, dg, db = ctx.path.points_to_bev_train_backward(res, ctx.pfn, stats, ctx.batch_size, g_canvas, g_feats)
        return None, None, None, dW, dg, db, None, None, None, None, None, None


def _pfn_forward(path, layer, voxels, coords, num, use_absolute_xyz, with_distance, training, weight=None):
    """Eval without gradients: the plain native forward.  Otherwise the autograd function (batch statistics in train mode)."""
    weight = layer.linear.weight if weight is None else weight
    params = [weight] + ([layer.norm.weight, layer.norm.bias] if layer.use_norm else [layer.linear.bias])
    needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    batch_stats = bool(training and layer.use_norm)
    if not needs_grad and not batch_stats:
        pfn = layer.weights(use_absolute_xyz, with_distance)
        pfn.weight = weight.detach()
        return path.pillar_vfe(voxels, coords, num, pfn)
    if batch_stats and layer.norm.num_batches_tracked is not None:
        layer.norm.num_batches_tracked += 1
    if layer.use_norm:
        n = layer.norm
        return _PillarVFEFunction.apply(path, voxels, coords, num, weight, n.weight, n.bias, None, n.running_mean, n.running_var,
                                        n.eps, n.momentum, batch_stats, use_absolute_xyz, with_distance)
    return _PillarVFEFunction.apply(path, voxels, coords, num, weight, None, None, layer.linear.bias, None, None, 1e-3, 0.0,
                                    False, use_absolute_xyz, with_distance)


class _PFNLayerParams(nn.Module):
    """Holds the parameters of the reference's PFNLayer under the same names (pillar_vfe.py:8-27):
    `linear.weight` (+ `linear.bias` when USE_NORM is False) and `norm.*` (BatchNorm1d eps 1e-3, momentum 0.01)."""

    def __init__(self, in_channels, out_channels, use_norm=True, last_layer=True):
        super().__init__()
        self.last_vfe = last_layer
        self.use_norm = use_norm
        if not self.last_vfe:
            out_channels = out_channels // 2
        if self.use_norm:
            self.linear = nn.Linear(in_channels, out_channels, bias=False)
            self.norm = nn.BatchNorm1d(out_channels, eps=1e-3, momentum=0.01)
        else:
            self.linear = nn.Linear(in_channels, out_channels, bias=True)
        self.part = 50000

    def weights(self, use_absolute_xyz: bool, with_distance: bool) -> PfnWeights:
        if self.use_norm:
            return PfnWeights(weight=self.linear.weight.detach(), bn_weight=self.norm.weight.detach(),
                              bn_bias=self.norm.bias.detach(), running_mean=self.norm.running_mean,
                              running_var=self.norm.running_var, eps=self.norm.eps,
                              use_absolute_xyz=use_absolute_xyz, with_distance=with_distance)
        return PfnWeights(weight=self.linear.weight.detach(), bias=self.linear.bias.detach(),
                          use_absolute_xyz=use_absolute_xyz, with_distance=with_distance)


class _VFEBase(nn.Module):
    def __init__(self, model_cfg, num_point_features, voxel_size, point_cloud_range, **kwargs):
        super().__init__()
        _lib.load()                                    # fail loudly here if the CUDA library is missing
        self.model_cfg = model_cfg
        self.use_norm = self.model_cfg.USE_NORM
        self.with_distance = self.model_cfg.WITH_DISTANCE
        self.use_absolute_xyz = self.model_cfg.USE_ABSLOTE_XYZ      # (sic) the reference's key, pillar_vfe.py:58
        self.num_raw_features = int(num_point_features)
        num_point_features += 6 if self.use_absolute_xyz else 3
        if self.with_distance:
            num_point_features += 1
        self.num_filters = list(self.model_cfg.NUM_FILTERS)
        assert len(self.num_filters) > 0
        if len(self.num_filters) > 2:
            raise NotImplementedError("PFN stacks of more than two layers (NUM_FILTERS with three entries) are not built")
        # the reference's loop (pillar_vfe.py:63-74): every layer but the last halves its out_channels and concatenates the max
        widths = [num_point_features] + self.num_filters
        self.pfn_layers = nn.ModuleList([_PFNLayerParams(widths[i], widths[i + 1], self.use_norm, last_layer=(i >= len(widths) - 2))
                                         for i in range(len(widths) - 1)])
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = point_cloud_range
        self.voxel_x, self.voxel_y, self.voxel_z = self.voxel_size
        # kept for parity with the reference attributes (pillar_vfe.py:79-81); the native geometry struct
        # evaluates the same expressions (geometry.make_geometry)
        self.x_offset = self.voxel_x / 2 + point_cloud_range[0]
        self.y_offset = self.voxel_y / 2 + point_cloud_range[1]
        self.z_offset = self.voxel_z / 2 + point_cloud_range[2]

    def get_output_feature_dim(self):
        return self.num_filters[-1]

    def _pfn(self) -> PfnWeights:
        return self.pfn_layers[0].weights(self.use_absolute_xyz, self.with_distance)

    @property
    def stacked(self) -> bool:
        return len(self.pfn_layers) == 2

    @property
    def fused_ok(self) -> bool:
        """The one-pass points -> canvas kernel covers the single-layer 64-channel PFN (every shipped config); other widths
        and the stacked PFN run as pillarize -> PFN -> scatter, three native launches on the same kernels' siblings."""
        return (not self.stacked) and self.num_filters[-1] == 64

    def _pfn_eval(self, path, voxels, coords, num):
        """Eval-mode PFN stack on the contract tensors (one native launch)."""
        if self.stacked:
            if self.training or (torch.is_grad_enabled() and any(p.requires_grad for p in self.pfn_layers.parameters())):
                raise NotImplementedError("training through a stacked (two-layer) PFN is not built: use eval() / no_grad()")
            return path.pillar_vfe_stacked(voxels, coords, num, self.pfn_layers[0].weights(self.use_absolute_xyz, self.with_distance),
                                           self.pfn_layers[1].weights(True, False))
        return _pfn_forward(path, self.pfn_layers[0], voxels, coords, num, self.use_absolute_xyz, self.with_distance,
                            self.training)


class PillarVFE(_VFEBase):
    """batch_dict contract mode: voxels, voxel_num_points, voxel_coords in; pillar_features out."""

    def __init__(self, model_cfg, num_point_features, voxel_size, point_cloud_range, grid_size=None, **kwargs):
        super().__init__(model_cfg, num_point_features, voxel_size, point_cloud_range)
        self.path = PillarPath(point_cloud_range, self.voxel_size, max_points_per_voxel=1, max_voxels=1,
                               num_point_features=self.num_raw_features, grid_size=grid_size)

    def forward(self, batch_dict, **kwargs):
        voxels, num, coords = batch_dict['voxels'], batch_dict['voxel_num_points'], batch_dict['voxel_coords']
        features = self._pfn_eval(self.path, voxels, coords, num)
        batch_dict['pillar_features'] = features.view(-1, 1, features.shape[-1]).squeeze()   # pillar_vfe.py:121
        return batch_dict


class Radar7PillarVFE(nn.Module):
    """The VoD radar variant (pillar_vfe.py:125-271): the raw features entering the PFN are chosen by the flags
    USE_XYZ / USE_RCS / USE_VR / USE_VR_COMP / USE_TIME (ascending column order x y z rcs v_r v_r_comp time), and with
    USE_ELEVATION False z is zeroed IN PLACE in batch_dict['voxels'] before anything else, exactly as the reference does.

    It runs on the same native kernel as PillarVFE: the Linear weight [C, Cin_selected] is expanded to the full
    [C, 7 + 6] layout with zero columns for the unselected features.  The kernel's dot product is a sequential fmaf in
    column order, and fmaf(x, 0, acc) == acc exactly for finite x, so the result is bit-identical to evaluating the
    selected columns only (checked against the reference's own outputs, tests/golden/radar7_*.npz).
    USE_DISTANCE must be False: the reference itself fails with it (it appends the range without widening the Linear)."""

    AVAILABLE = ['x', 'y', 'z', 'rcs', 'v_r', 'v_r_comp', 'time']

    def __init__(self, model_cfg, num_point_features, voxel_size, point_cloud_range, grid_size=None, **kwargs):
        super().__init__()
        _lib.load()
        self.model_cfg = model_cfg
        self.use_norm = self.model_cfg.USE_NORM
        self.use_xyz = self.model_cfg.USE_XYZ
        self.with_distance = self.model_cfg.USE_DISTANCE
        params = ["USE_RCS", "USE_VR", "USE_VR_COMP", "USE_TIME", "USE_ELEVATION"]
        if not all(hasattr(self.model_cfg, a) for a in params):
            raise Exception("config does not have the right parameters, please use a radar config")
        if self.with_distance:
            raise NotImplementedError("Radar7PillarVFE with USE_DISTANCE: the reference's own forward fails "
                                      "(Linear width does not count the range feature)")
        self.use_elevation = self.model_cfg.USE_ELEVATION
        sel = []
        if self.use_xyz:
            sel += [0, 1, 2]
        for flag, name in (("USE_RCS", 'rcs'), ("USE_VR", 'v_r'), ("USE_VR_COMP", 'v_r_comp'), ("USE_TIME", 'time')):
            if getattr(self.model_cfg, flag):
                sel.append(self.AVAILABLE.index(name))
        self.selected_indexes = torch.LongTensor(sel)
        self.z_ind = 2
        self.num_filters = list(self.model_cfg.NUM_FILTERS)
        if len(self.num_filters) != 1:
            raise NotImplementedError("single-layer PFN only (NUM_FILTERS: [64])")
        self.pfn_layers = nn.ModuleList([_PFNLayerParams(len(sel) + 6, self.num_filters[0], self.use_norm, True)])
        self.voxel_size = [float(v) for v in voxel_size]
        self.path = PillarPath(point_cloud_range, self.voxel_size, max_points_per_voxel=1, max_voxels=1,
                               num_point_features=7, grid_size=grid_size)

    def get_output_feature_dim(self):
        return self.num_filters[-1]

    def _expanded_weight(self) -> torch.Tensor:
        """[C, len(sel)+6] -> [C, 13] with zero columns; built with differentiable torch ops so that the gradient of the
        expanded weight flows back to the selected columns of the parameter."""
        w = self.pfn_layers[0].linear.weight
        k = len(self.selected_indexes)
        cols = torch.cat([self.selected_indexes.to(w.device), torch.arange(7, 13, device=w.device)])
        return torch.zeros((w.shape[0], 13), dtype=w.dtype, device=w.device).index_copy(1, cols, w[:, :k + 6])

    def forward(self, batch_dict, **kwargs):
        voxels, num, coords = batch_dict['voxels'], batch_dict['voxel_num_points'], batch_dict['voxel_coords']
        if not self.use_elevation:
            voxels[:, :, self.z_ind] = 0          # in place on the caller's tensor, like pillar_vfe.py:232-233
        features = _pfn_forward(self.path, self.pfn_layers[0], voxels, coords, num, True, False, self.training,
                                weight=self._expanded_weight())
        batch_dict['pillar_features'] = features.view(-1, 1, features.shape[-1]).squeeze()
        return batch_dict


class PointPillarScatter(nn.Module):
    """pillar_features + voxel_coords -> spatial_features [B, C, ny, nx]."""

    def __init__(self, model_cfg, grid_size, point_cloud_range=None, voxel_size=None, **kwargs):
        super().__init__()
        _lib.load()
        self.model_cfg = model_cfg
        self.num_bev_features = self.model_cfg.NUM_BEV_FEATURES
        self.nx, self.ny, self.nz = (int(v) for v in grid_size)
        assert self.nz == 1
        # the scatter needs only the grid; range / voxel size are placeholders for the geometry struct
        self.path = PillarPath(np.zeros(6, np.float32) if point_cloud_range is None else point_cloud_range,
                               [1.0, 1.0, 1.0] if voxel_size is None else voxel_size, 1, 1, 4,
                               grid_size=[self.nx, self.ny, self.nz])
        self.batch_size_from_coords = True     # the reference's rule (pointpillar_scatter.py:21); one host sync

    def forward(self, batch_dict, **kwargs):
        pillar_features, coords = batch_dict['pillar_features'], batch_dict['voxel_coords']
        if self.batch_size_from_coords or 'batch_size' not in batch_dict:
            batch_size = coords[:, 0].max().int().item() + 1
        else:
            batch_size = int(batch_dict['batch_size'])
        if pillar_features.dim() == 1:                 # the reference's squeeze() dropped M == 1
            pillar_features = pillar_features.view(1, -1)
        if torch.is_grad_enabled() and pillar_features.requires_grad:
            batch_dict['spatial_features'] = _ScatterFunction.apply(self.path, pillar_features, coords, batch_size)
        else:
            batch_dict['spatial_features'] = self.path.pointpillar_scatter(pillar_features, coords, batch_size)
        return batch_dict


class FusedPillarVFE(_VFEBase):
    """VFE plugin that starts from batch_dict['points'] ([sum N, 1+F], column 0 = batch index,
    pcdet/datasets/dataset.py:237-244) and produces everything the detector reads downstream:
    voxel_coords, voxel_num_points, pillar_features and spatial_features (and voxels on request).

    model_cfg keys beyond PillarVFE's: MAX_POINTS_PER_VOXEL, MAX_NUMBER_OF_VOXELS (int or {'train','test'}),
    optional RETURN_VOXELS (default False), TRIM (default True: slice outputs to [M, ...] like the
    reference, one host sync; False keeps capacity-sized tensors and batch_dict['num_pillars'] on device)."""

    def __init__(self, model_cfg, num_point_features, voxel_size, point_cloud_range, grid_size=None, **kwargs):
        super().__init__(model_cfg, num_point_features, voxel_size, point_cloud_range)
        mv = self.model_cfg.MAX_NUMBER_OF_VOXELS
        self._max_voxels = mv if isinstance(mv, dict) else {'train': int(mv), 'test': int(mv)}
        self.max_points = int(self.model_cfg.MAX_POINTS_PER_VOXEL)
        self.return_voxels = bool(getattr(self.model_cfg, 'RETURN_VOXELS', False))
        self.trim = bool(getattr(self.model_cfg, 'TRIM', True))
        self._grid = grid_size
        self._paths = {}

    def _path(self) -> PillarPath:
        mode = 'train' if self.training else 'test'
        if mode not in self._paths:
            self._paths[mode] = PillarPath(self.point_cloud_range, self.voxel_size, self.max_points,
                                           int(self._max_voxels[mode]), self.num_raw_features, grid_size=self._grid)
        return self._paths[mode]

    def _forward_train(self, points, batch_size):
        """Train mode / parameters that need gradients: pillarize natively, then the differentiable PFN and scatter
        (the same native kernels PillarVFE and PointPillarScatter use); one host sync for the pillar count."""
        path = self._path()
        out = path.pillarize(points, batch_size, xyz_col=1, batch_col=0, want_voxels=True).trim()
        out.pop('num_pillars')
        feats = self._pfn_eval(path, out['voxels'], out['voxel_coords'], out['voxel_num_points'])
        out['pillar_features'] = feats
        if torch.is_grad_enabled() and feats.requires_grad:
            out['spatial_features'] = _ScatterFunction.apply(path, feats, out['voxel_coords'], batch_size)
        else:
            out['spatial_features'] = path.pointpillar_scatter(feats, out['voxel_coords'], batch_size)
        if not self.return_voxels:
            out.pop('voxels')
        return out

    def _forward_train_fused(self, points, batch_size):
        """Train mode with BatchNorm inside the fused kernel's domain: three launches from points to the canvas, the batch
        statistics computed on the way, nothing read back unless TRIM asks for reference-shaped outputs."""
        path, layer = self._path(), self.pfn_layers[0]
        n = layer.norm
        if n.num_batches_tracked is not None:
            n.num_batches_tracked += 1
        feats, canvas, coords, num, counts, voxels = _FusedTrainFunction.apply(
            path, points, batch_size, layer.linear.weight, n.weight, n.bias, n.running_mean, n.running_var, n.eps,
            n.momentum, self.use_absolute_xyz, self.with_distance)
        out = dict(voxel_coords=coords, voxel_num_points=num, pillar_features=feats, spatial_features=canvas)
        if self.return_voxels:
            out['voxels'] = voxels
        if self.trim:
            M = int(counts[0])                       # the one host sync, after everything has been queued
            for k in ('voxel_coords', 'voxel_num_points', 'pillar_features', 'voxels'):
                if k in out:
                    out[k] = out[k][:M]
        else:
            out['num_pillars'] = counts
        return out

    def forward(self, batch_dict, **kwargs):
        points = batch_dict['points']
        batch_size = int(batch_dict['batch_size'])
        layer = self.pfn_layers[0]
        if (self.fused_ok and self.training and layer.use_norm and layer.norm.momentum is not None and
                layer.norm.track_running_stats and points.is_cuda and self._path().fused_train_supported()):
            batch_dict.update(self._forward_train_fused(points, batch_size))
            return batch_dict
        if (not self.fused_ok or (self.training and layer.use_norm) or
                (torch.is_grad_enabled() and any(p.requires_grad for p in layer.parameters()))):
            # train mode / gradients, or a PFN outside the fused kernel set (stacked, or not 64 channels): the composed native path
            batch_dict.update(self._forward_train(points, batch_size))
            return batch_dict
        res = self._path().points_to_bev(points, batch_size, self._pfn(), xyz_col=1, batch_col=0,
                                         want_voxels=self.return_voxels)
        if self.trim:
            out = res.trim()
            out.pop('num_pillars')
        else:
            out = dict(voxel_coords=res.voxel_coords, voxel_num_points=res.voxel_num_points,
                       pillar_features=res.pillar_features, spatial_features=res.spatial_features,
                       num_pillars=res.num_pillars)
            if res.voxels is not None:
                out['voxels'] = res.voxels
        batch_dict.update(out)
        return batch_dict


class PillarScatterPassthrough(nn.Module):
    """MAP_TO_BEV plugin to pair with FusedPillarVFE: the canvas is already in batch_dict."""

    def __init__(self, model_cfg, grid_size, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = self.model_cfg.NUM_BEV_FEATURES
        self.nx, self.ny, self.nz = (int(v) for v in grid_size)
        assert self.nz == 1

    def forward(self, batch_dict, **kwargs):
        if 'spatial_features' not in batch_dict:
            raise KeyError("PillarScatterPassthrough expects FusedPillarVFE upstream (no 'spatial_features')")
        return batch_dict


def register(vfe_all: dict | None = None, map_to_bev_all: dict | None = None, override: bool = False):
    """Adds the modules to the reference's name -> class registries
    (pcdet/models/backbones_3d/vfe/__init__.py:13-27, pcdet/models/backbones_2d/map_to_bev/__init__.py:7-14).
    With override=True the stock 'PillarVFE' / 'PointPillarScatter' names resolve to the native modules."""
    if vfe_all is not None:
        vfe_all['FusedPillarVFE'] = FusedPillarVFE
        vfe_all['PillarVFEB200'] = PillarVFE
        vfe_all['Radar7PillarVFEB200'] = Radar7PillarVFE
        if override:
            vfe_all['PillarVFE'] = PillarVFE
            vfe_all['Radar7PillarVFE'] = Radar7PillarVFE
    if map_to_bev_all is not None:
        map_to_bev_all['PillarScatterPassthrough'] = PillarScatterPassthrough
        map_to_bev_all['PointPillarScatterB200'] = PointPillarScatter
        if override:
            map_to_bev_all['PointPillarScatter'] = PointPillarScatter
