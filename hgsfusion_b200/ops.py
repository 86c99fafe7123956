"""Tensor-level front end of the C ABI (include/hgsfusion_b200.h).

torch is plumbing here: device memory, the current stream, nothing else.  Every function passes
raw pointers to libhgsfusion_b200.so; none of them computes anything in PyTorch, and none falls
back to the CPU.

Reference interfaces mirrored (file:line under the HGSFusion repo):
  pillarize            DataProcessor.transform_points_to_voxels   datasets/processor/data_processor.py:133-183
  pillar_vfe           PillarVFE.forward                          models/backbones_3d/vfe/pillar_vfe.py:94-123
  pointpillar_scatter  PointPillarScatter.forward                 models/backbones_2d/map_to_bev/pointpillar_scatter.py:14-41
  points_to_bev        the three above fused
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from .geometry import make_geometry


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device=None):
    """The current stream OF THE TENSORS' DEVICE (not of whatever device happens to be current)."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (hgsfusion_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


@dataclass
class PfnWeights:
    """Device tensors of the single PFN layer, named as in the reference state_dict
    (pfn_layers.0.linear.weight, pfn_layers.0.norm.{weight,bias,running_mean,running_var})."""
    weight: torch.Tensor
    bn_weight: torch.Tensor | None = None
    bn_bias: torch.Tensor | None = None
    running_mean: torch.Tensor | None = None
    running_var: torch.Tensor | None = None
    bias: torch.Tensor | None = None
    eps: float = 1e-3
    use_absolute_xyz: bool = True
    with_distance: bool = False

    def to_struct(self) -> _lib.Pfn:
        w = _f32c(self.weight, "weight")
        s = _lib.Pfn()
        self._keep = [w]
        s.weight = w.data_ptr()
        for field, t in (("bias", self.bias), ("bn_weight", self.bn_weight), ("bn_bias", self.bn_bias),
                         ("bn_mean", self.running_mean), ("bn_var", self.running_var)):
            if t is not None:
                t = _f32c(t, field)
                self._keep.append(t)
                setattr(s, field, t.data_ptr())
        s.bn_eps = float(self.eps)
        s.out_channels, s.in_channels = int(w.shape[0]), int(w.shape[1])
        s.use_absolute_xyz, s.with_distance = int(self.use_absolute_xyz), int(self.with_distance)
        return s


@dataclass
class PillarResult:
    """Outputs of the pillar path, capacity-sized on the device.  `num_pillars` is int32 [1+B]
    (total, then per frame).  `trim()` reads the total (one host sync, as the reference's
    contract needs [M, ...] tensors) and returns the batch_dict entries."""
    voxel_coords: torch.Tensor
    voxel_num_points: torch.Tensor
    num_pillars: torch.Tensor
    voxels: torch.Tensor | None = None
    pillar_features: torch.Tensor | None = None
    spatial_features: torch.Tensor | None = None

    def trim(self) -> dict:
        m = int(self.num_pillars[0].item())
        out = dict(voxel_coords=self.voxel_coords[:m], voxel_num_points=self.voxel_num_points[:m], num_pillars=m)
        if self.voxels is not None:
            out["voxels"] = self.voxels[:m]
        if self.pillar_features is not None:
            out["pillar_features"] = self.pillar_features[:m]
        if self.spatial_features is not None:
            out["spatial_features"] = self.spatial_features
        return out


class PillarPath:
    """A configured pillar path: geometry + limits + (optionally) PFN weights.

    Holds the workspace so that steady-state calls allocate nothing; a call enqueues one memset and
    four kernels on the current stream and returns without synchronising."""

    def __init__(self, point_cloud_range, voxel_size, max_points_per_voxel: int, max_voxels: int,
                 num_point_features: int, grid_size=None, spconv_version: int = 2):
        """spconv_version: whose overflow semantics to reproduce when a frame holds more than max_voxels pillars -- 2
        (Point2VoxelCPU3d: later NEW pillars are refused) or 1 (VoxelGenerator: the loop stops there, dropping every later
        point).  VoxelGeneratorWrapper picks whichever spconv is installed (data_processor.py:16-26)."""
        if spconv_version not in (1, 2):
            raise ValueError("spconv_version must be 1 or 2")
        self.flags = _lib.POINTS_SPCONV1_BREAK if spconv_version == 1 else 0
        self.lib = _lib.load()
        self.geom = make_geometry(point_cloud_range, voxel_size, grid_size)
        self.nx, self.ny, self.nz = (int(v) for v in self.geom.grid)
        self.P, self.max_voxels, self.F = int(max_points_per_voxel), int(max_voxels), int(num_point_features)
        self._ws = None
        self.last_launches = 0

    # -- helpers ---------------------------------------------------------------------------------
    def capacity(self, n_points: int, batch_size: int) -> int:
        cap = self.lib.hgsf_pillar_capacity(C.byref(self.geom), n_points, batch_size, self.max_voxels)
        if cap < 0:
            raise _lib.HgsfError(_lib.ERR_INVALID_ARG, "hgsf_pillar_capacity")
        return int(cap)

    def _workspace(self, n_points: int, batch_size: int, device) -> torch.Tensor:
        need = C.c_size_t(0)
        _lib.check(self.lib.hgsf_workspace_size(C.byref(self.geom), n_points, batch_size, self.F, C.byref(need)),
                   "hgsf_workspace_size")
        if self._ws is None or self._ws.numel() < need.value or self._ws.device != device:
            self._ws = torch.empty(need.value + 256, dtype=torch.uint8, device=device)
        return self._ws

    def _points_struct(self, points, batch_size, xyz_col, batch_col, frame_offsets, flip=(False, False)):
        pts = _f32c(points, "points")
        if pts.dim() != 2:
            raise ValueError("points must be [n, stride]")
        s = _lib.Points()
        s.data = pts.data_ptr()
        s.n, s.stride = int(pts.shape[0]), int(pts.shape[1])
        s.xyz_col, s.num_features = int(xyz_col), self.F
        s.batch_col = int(batch_col) if frame_offsets is None else -1
        keep = [pts]
        if frame_offsets is not None:
            fo = frame_offsets.to(device=pts.device, dtype=torch.int32).contiguous()
            if fo.numel() != batch_size + 1:
                raise ValueError("frame_offsets must have batch_size + 1 entries")
            s.frame_offsets = fo.data_ptr()
            keep.append(fo)
        s.batch_size = int(batch_size)
        s.flags = self.flags | (_lib.POINTS_FLIP_X if flip[0] else 0) | (_lib.POINTS_FLIP_Y if flip[1] else 0)
        return s, keep

    def _run(self, points, batch_size, pfn, xyz_col, batch_col, frame_offsets, want_voxels, want_features,
             want_canvas, out: PillarResult | None, flip=(False, False), train=None):
        ps, keep = self._points_struct(points, batch_size, xyz_col, batch_col, frame_offsets, flip)
        dev = points.device
        cap = max(self.capacity(ps.n, batch_size), 1)
        ws = self._workspace(ps.n, batch_size, dev)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        ws_bytes = ws.numel() - (ws_ptr - ws.data_ptr())
        if out is None:
            C_out = int(pfn.weight.shape[0]) if pfn is not None else 0
            out = PillarResult(
                voxel_coords=torch.empty((cap, 4), dtype=torch.int32, device=dev),
                voxel_num_points=torch.empty((cap,), dtype=torch.int32, device=dev),
                num_pillars=torch.empty((1 + batch_size,), dtype=torch.int32, device=dev),
                voxels=torch.empty((cap, self.P, self.F), dtype=torch.float32, device=dev) if want_voxels else None,
                pillar_features=torch.empty((cap, C_out), dtype=torch.float32, device=dev)
                if (pfn is not None and want_features) else None,
                spatial_features=torch.empty((batch_size, C_out * self.nz, self.ny, self.nx), dtype=torch.float32,
                                             device=dev) if (pfn is not None and want_canvas) else None)
        o = _lib.PillarOutputs()
        o.voxel_coords, o.voxel_num_points = out.voxel_coords.data_ptr(), out.voxel_num_points.data_ptr()
        o.num_pillars = out.num_pillars.data_ptr()
        o.voxels = out.voxels.data_ptr() if out.voxels is not None else None
        o.pillar_features = out.pillar_features.data_ptr() if out.pillar_features is not None else None
        o.spatial_features = out.spatial_features.data_ptr() if out.spatial_features is not None else None
        o.pillar_capacity = int(out.voxel_coords.shape[0])
        with torch.cuda.device(dev):           # the library launches on the current device: make it the tensors' device
            if pfn is None:
                st = self.lib.hgsf_pillarize(C.byref(self.geom), C.byref(ps), self.P, self.max_voxels,
                                             C.c_void_p(ws_ptr), ws_bytes, C.byref(o), _stream(dev))
                _lib.check(st, "hgsf_pillarize")
            elif train is not None:
                pf = pfn.to_struct()
                momentum, run_mean, run_var, mean, var, stats = train
                st = self.lib.hgsf_points_to_bev_train(C.byref(self.geom), C.byref(ps), C.byref(pf), self.P, self.max_voxels,
                                                       C.c_void_p(ws_ptr), ws_bytes, C.byref(o), float(momentum), _ptr(run_mean),
                                                       _ptr(run_var), _ptr(mean), _ptr(var), _ptr(stats), _stream(dev))
                _lib.check(st, "hgsf_points_to_bev_train")
            else:
                pf = pfn.to_struct()
                st = self.lib.hgsf_points_to_bev(C.byref(self.geom), C.byref(ps), C.byref(pf), self.P, self.max_voxels,
                                                 C.c_void_p(ws_ptr), ws_bytes, C.byref(o), _stream(dev))
                _lib.check(st, "hgsf_points_to_bev")
            self.last_launches = int(self.lib.hgsf_last_launch_count())
        del keep
        return out

    # -- public ----------------------------------------------------------------------------------
    def pillarize(self, points, batch_size, xyz_col=1, batch_col=0, frame_offsets=None, want_voxels=True,
                  out: PillarResult | None = None, flip=(False, False)) -> PillarResult:
        """points [n, stride] -> voxels, voxel_coords, voxel_num_points (transform_points_to_voxels + collate).
        flip = (flip_x, flip_y): voxelize the mirrored cloud (DataProcessor.double_flip) without copying the points."""
        return self._run(points, batch_size, None, xyz_col, batch_col, frame_offsets, want_voxels, False, False, out, flip)

    def pillarize_double_flip(self, points, batch_size, xyz_col=1, batch_col=0, frame_offsets=None):
        """DOUBLE_FLIP test-time augmentation (data_processor.py:161-178): the cloud and its y-, x- and xy-mirrored copies,
        voxelized one after the other -> four PillarResults in the reference's order [original, yflip, xflip, xyflip]."""
        return [self.pillarize(points, batch_size, xyz_col, batch_col, frame_offsets, True, None, flip)
                for flip in ((False, False), (False, True), (True, False), (True, True))]

    def points_to_bev(self, points, batch_size, pfn: PfnWeights, xyz_col=1, batch_col=0, frame_offsets=None,
                      want_voxels=False, want_features=True, want_canvas=True,
                      out: PillarResult | None = None) -> PillarResult:
        """points [n, stride] -> voxel_coords, voxel_num_points, pillar_features, spatial_features, one pass."""
        return self._run(points, batch_size, pfn, xyz_col, batch_col, frame_offsets, want_voxels, want_features,
                         want_canvas, out)

    def fused_train_supported(self) -> bool:
        """hgsf_points_to_bev_train's domain (besides BatchNorm and 64 channels, which the module checks)."""
        return self.P <= 32 and self.nz == 1 and self.nx % 4 == 0 and not (self.flags & _lib.POINTS_SPCONV1_BREAK)

    def points_to_bev_train(self, points, batch_size, pfn: PfnWeights, momentum: float, running_mean=None, running_var=None,
                            xyz_col=1, batch_col=0, frame_offsets=None):
        """Train-mode forward from points in three launches and no host round trip: BatchNorm1d on the statistics of this
        batch (pillar_vfe.py:38-40), running statistics updated in place.  Returns (PillarResult with voxels -- all at
        capacity rows, the count stays in num_pillars on the device --, batch_mean, batch_var, stats for the backward)."""
        dev = points.device
        Cc = int(pfn.weight.shape[0])
        mean = torch.empty(Cc, dtype=torch.float32, device=dev)
        var = torch.empty(Cc, dtype=torch.float32, device=dev)
        stats = torch.empty(int(self.lib.hgsf_train_stats_doubles(Cc, int(pfn.weight.shape[1]))), dtype=torch.float64, device=dev)
        res = self._run(points, batch_size, pfn, xyz_col, batch_col, frame_offsets, True, True, True, None,
                        train=(momentum, running_mean, running_var, mean, var, stats))
        return res, mean, var, stats

    def points_to_bev_train_backward(self, res: PillarResult, pfn: PfnWeights, stats, batch_size, grad_spatial_features=None,
                                     grad_pillar_features=None):
        """Gradients (linear.weight, norm.weight, norm.bias) of points_to_bev_train; `pfn` carries the batch statistics the
        forward returned as bn_mean / bn_var, `stats` its statistics buffer.  Nothing is read back: the pillar count comes from res.num_pillars."""
        vox = res.voxels
        dev = vox.device
        cap, P, F = (int(v) for v in vox.shape)
        Cc, cin = int(pfn.weight.shape[0]), int(pfn.weight.shape[1])
        gc = _f32c(grad_spatial_features, "grad_spatial_features") if grad_spatial_features is not None else None
        gp = _f32c(grad_pillar_features, "grad_pillar_features") if grad_pillar_features is not None else None
        if gp is not None and tuple(gp.shape) != (cap, Cc):
            raise ValueError("grad_pillar_features must be [capacity, C]")
        rows = torch.empty((cap, Cc), dtype=torch.float32, device=dev)
        scratch = torch.empty(int(self.lib.hgsf_train_scratch_doubles(Cc, cin)), dtype=torch.float64, device=dev)
        dW = torch.empty((Cc, cin), dtype=torch.float32, device=dev)
        dg = torch.empty(Cc, dtype=torch.float32, device=dev)
        db = torch.empty(Cc, dtype=torch.float32, device=dev)
        pf = pfn.to_struct()
        with torch.cuda.device(dev):
            st = self.lib.hgsf_points_to_bev_train_backward(
                C.byref(self.geom), C.byref(pf), _ptr(vox), _ptr(res.voxel_coords), _ptr(res.voxel_num_points), cap,
                _ptr(res.num_pillars), P, F, int(batch_size), _ptr(gc), _ptr(gp), _ptr(rows), _ptr(stats), _ptr(scratch),
                _ptr(dW), _ptr(dg), _ptr(db), _stream(dev))
        _lib.check(st, "hgsf_points_to_bev_train_backward")
        self.last_launches = int(self.lib.hgsf_last_launch_count())
        return dW, dg, db

    def pillar_vfe(self, voxels, voxel_coords, voxel_num_points, pfn: PfnWeights) -> torch.Tensor:
        """batch_dict contract: voxels [M,P,F], voxel_coords [M,4], voxel_num_points [M] -> pillar_features [M,C].
        coords / counts may be float32 (as load_data_to_gpu delivers them) or int32."""
        vox = _f32c(voxels, "voxels")
        M, P, F = (int(v) for v in vox.shape)
        co, cf = _coords(voxel_coords, "voxel_coords")
        nu, nf = _coords(voxel_num_points, "voxel_num_points")
        pf = pfn.to_struct()
        out = torch.empty((M, int(pfn.weight.shape[0])), dtype=torch.float32, device=vox.device)
        with torch.cuda.device(vox.device):
            st = self.lib.hgsf_pillar_vfe(C.byref(self.geom), C.byref(pf), _ptr(vox), _ptr(co), _ptr(nu), cf, nf, M, P, F,
                                      _ptr(out), _stream(vox.device))
        _lib.check(st, "hgsf_pillar_vfe")
        self.last_launches = int(self.lib.hgsf_last_launch_count())
        return out

    def pillar_vfe_stacked(self, voxels, voxel_coords, voxel_num_points, pfn0: PfnWeights, pfn1: PfnWeights) -> torch.Tensor:
        """The same for a two-layer (stacked) PFN: pfn0 = Linear(Cin -> H) (+BN), pfn1 = Linear(2H -> C1) (+BN); eval mode."""
        vox = _f32c(voxels, "voxels")
        M, P, F = (int(v) for v in vox.shape)
        co, cf = _coords(voxel_coords, "voxel_coords")
        nu, nf = _coords(voxel_num_points, "voxel_num_points")
        p0, p1 = pfn0.to_struct(), pfn1.to_struct()
        out = torch.empty((M, int(pfn1.weight.shape[0])), dtype=torch.float32, device=vox.device)
        with torch.cuda.device(vox.device):
            st = self.lib.hgsf_pillar_vfe_stacked(C.byref(self.geom), C.byref(p0), C.byref(p1), _ptr(vox), _ptr(co), _ptr(nu), cf, nf,
                                                  M, P, F, _ptr(out), _stream(vox.device))
        _lib.check(st, "hgsf_pillar_vfe_stacked")
        self.last_launches = int(self.lib.hgsf_last_launch_count())
        return out

    # -- training (include/hgsfusion_b200.h "Training through the path") ---------------------------------------
    def _contract_args(self, voxels, voxel_coords, voxel_num_points):
        vox = _f32c(voxels, "voxels")
        M, P, F = (int(v) for v in vox.shape)
        co, cf = _coords(voxel_coords, "voxel_coords")
        nu, nf = _coords(voxel_num_points, "voxel_num_points")
        return vox, co, cf, nu, nf, M, P, F

    def pillar_vfe_batch_stats(self, voxels, voxel_coords, voxel_num_points, pfn: PfnWeights, momentum: float,
                               running_mean=None, running_var=None):
        """BatchNorm1d train-mode statistics of PFNLayer.forward (pillar_vfe.py:38-40) over all M*P rows: returns
        (batch_mean, batch_var (biased), stats) and updates running_mean / running_var in place like torch does."""
        vox, co, cf, nu, nf, M, P, F = self._contract_args(voxels, voxel_coords, voxel_num_points)
        pf = pfn.to_struct()
        Cc, cin = int(pfn.weight.shape[0]), int(pfn.weight.shape[1])
        stats = torch.empty(int(self.lib.hgsf_train_stats_doubles(Cc, cin)), dtype=torch.float64, device=vox.device)
        mean = torch.empty(Cc, dtype=torch.float32, device=vox.device)
        var = torch.empty(Cc, dtype=torch.float32, device=vox.device)
        with torch.cuda.device(vox.device):
            st = self.lib.hgsf_pillar_vfe_batch_stats(C.byref(self.geom), C.byref(pf), _ptr(vox), _ptr(co), _ptr(nu), cf, nf, M, P, F,
                                                  float(momentum), _ptr(running_mean), _ptr(running_var), _ptr(mean), _ptr(var),
                                                  _ptr(stats), _stream(vox.device))
        _lib.check(st, "hgsf_pillar_vfe_batch_stats")
        self.last_launches = int(self.lib.hgsf_last_launch_count())
        return mean, var, stats

    def pillar_vfe_backward(self, voxels, voxel_coords, voxel_num_points, pfn: PfnWeights, grad_out, stats=None):
        """grad of pillar_features [M,C] -> (grad_weight [C,Cin], grad_gamma [C] or None, grad_beta_or_bias [C]).
        `pfn` must carry the statistics the forward normalised with; `stats` = what pillar_vfe_batch_stats returned
        (train mode) or None (running statistics, constants)."""
        vox, co, cf, nu, nf, M, P, F = self._contract_args(voxels, voxel_coords, voxel_num_points)
        pf = pfn.to_struct()
        Cc, cin = int(pfn.weight.shape[0]), int(pfn.weight.shape[1])
        g = _f32c(grad_out, "grad_pillar_features").view(M, Cc)
        scratch = torch.empty(int(self.lib.hgsf_train_scratch_doubles(Cc, cin)), dtype=torch.float64, device=vox.device)
        dW = torch.empty((Cc, cin), dtype=torch.float32, device=vox.device)
        dg = torch.empty(Cc, dtype=torch.float32, device=vox.device) if pfn.bn_weight is not None else None
        db = torch.empty(Cc, dtype=torch.float32, device=vox.device)
        with torch.cuda.device(vox.device):
            st = self.lib.hgsf_pillar_vfe_backward(C.byref(self.geom), C.byref(pf), _ptr(vox), _ptr(co), _ptr(nu), cf, nf, M, P, F,
                                               _ptr(g), _ptr(stats), _ptr(scratch), _ptr(dW), _ptr(dg), _ptr(db), _stream(vox.device))
        _lib.check(st, "hgsf_pillar_vfe_backward")
        self.last_launches = int(self.lib.hgsf_last_launch_count())
        return dW, dg, db

    def pointpillar_scatter_backward(self, grad_spatial_features, voxel_coords, batch_size: int) -> torch.Tensor:
        gc = _f32c(grad_spatial_features, "grad_spatial_features")
        co, cf = _coords(voxel_coords, "voxel_coords")
        M, Cc = int(co.shape[0]), int(gc.shape[1])
        out = torch.empty((M, Cc), dtype=torch.float32, device=gc.device)
        with torch.cuda.device(gc.device):
            st = self.lib.hgsf_pointpillar_scatter_backward(C.byref(self.geom), _ptr(gc), _ptr(co), cf, M, Cc, int(batch_size),
                                                        _ptr(out), _stream(gc.device))
        _lib.check(st, "hgsf_pointpillar_scatter_backward")
        self.last_launches = int(self.lib.hgsf_last_launch_count())
        return out

    def pointpillar_scatter(self, pillar_features, voxel_coords, batch_size: int) -> torch.Tensor:
        """pillar_features [M,C] + voxel_coords [M,4] -> spatial_features [B, C, ny, nx]."""
        pf = _f32c(pillar_features, "pillar_features")
        co, cf = _coords(voxel_coords, "voxel_coords")
        M, Cc = int(pf.shape[0]), int(pf.shape[1])
        need = C.c_size_t(0)
        _lib.check(self.lib.hgsf_scatter_workspace_size(C.byref(self.geom), batch_size, C.byref(need)),
                   "hgsf_scatter_workspace_size")
        ws = torch.empty(need.value + 256, dtype=torch.uint8, device=pf.device)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        canvas = torch.empty((batch_size, Cc * self.nz, self.ny, self.nx), dtype=torch.float32, device=pf.device)
        with torch.cuda.device(pf.device):
            st = self.lib.hgsf_pointpillar_scatter(C.byref(self.geom), _ptr(pf), _ptr(co), cf, M, Cc, batch_size,
                                               C.c_void_p(ws_ptr), need.value, _ptr(canvas), _stream(pf.device))
        _lib.check(st, "hgsf_pointpillar_scatter")
        self.last_launches = int(self.lib.hgsf_last_launch_count())
        return canvas


def _coords(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor")
    if t.dtype == torch.float32:
        return t.contiguous(), 1
    if t.dtype == torch.int32:
        return t.contiguous(), 0
    if t.dtype in (torch.int64, torch.int16, torch.uint8):
        return t.to(torch.int32).contiguous(), 0
    raise TypeError(f"{name}: unsupported dtype {t.dtype}")
