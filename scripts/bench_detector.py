"""BASELINE.json configs[4] ("full HGSFusion-VoD forward ... with the new pillar path dropped in"), reduced as SURVEY.md
section 8(d) config 5 prescribes: the full detector cannot be built here or on the GPU box (spconv, mmcv, kornia, three source files
missing from the reference repository, DeepLab weights), so this measures a REDUCED detector with random-init weights and
synthetic radar points -- the radar branch only:

    points -> [pillar path] -> spatial_features [B,64,320,320] -> BEV backbone (3 stages 3/5/5 convs, 64/128/256 channels,
              strides 2/2/2, three 128-channel up-sampling branches; the layout of pcdet/models/backbones_2d/base_bev_backbone.py:6-60)
           -> 1x1 conv heads (class / box / direction for 6 anchors per cell; anchor_head_single.py)

with the pillar path in two arms on the SAME B200, the SAME batch and the SAME backbone module:

  eager : what the reference executes -- voxels / voxel_coords / voxel_num_points produced on the CPU
          (transform_points_to_voxels; here the oracle's C voxelizer, one frame per host thread like DataLoader workers),
          copied host -> device as load_data_to_gpu does, then PillarVFE.forward (pillar_vfe.py:94-123) and
          PointPillarScatter.forward (pointpillar_scatter.py:14-41) as the same sequence of torch ops, eager, cuBLAS / native kernels.
  ours  : points host -> device, hgsf_points_to_bev (k_front + k_emit).

Reported: milliseconds of each stage (CUDA events), detector frames/s of both arms with the CPU voxelization of the eager arm
(a) excluded -- as if perfectly hidden by DataLoader workers -- and (b) included, plus max |difference| of the two canvases.
The torch restatement is test infrastructure (a GPU cross-check of the kernels, 1e-5), not a product path.
"""
from __future__ import annotations

import json
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as Fn

from hgsfusion_b200 import synthetic
from hgsfusion_b200.ops import PillarPath
from util import device_pfn


# ---- eager restatement of the reference modules (torch ops in the reference's order) --------------------------------------
def eager_pillar_vfe(voxels, num_points, coords, W, bn, vsize, offs):
    """voxels [M,P,F], num_points [M], coords [M,4] (b,z,y,x) float -> [M,C]."""
    M, P, _ = voxels.shape
    mean = voxels[:, :, :3].sum(dim=1, keepdim=True) / num_points.type_as(voxels).view(-1, 1, 1)
    f_cluster = voxels[:, :, :3] - mean
    f_center = torch.zeros_like(voxels[:, :, :3])
    f_center[:, :, 0] = voxels[:, :, 0] - (coords[:, 3].to(voxels.dtype).unsqueeze(1) * vsize[0] + offs[0])
    f_center[:, :, 1] = voxels[:, :, 1] - (coords[:, 2].to(voxels.dtype).unsqueeze(1) * vsize[1] + offs[1])
    f_center[:, :, 2] = voxels[:, :, 2] - (coords[:, 1].to(voxels.dtype).unsqueeze(1) * vsize[2] + offs[2])
    feats = torch.cat([voxels, f_cluster, f_center], dim=-1)
    mask = (torch.arange(P, device=voxels.device).view(1, -1) < num_points.view(-1, 1).int()).unsqueeze(-1).type_as(voxels)
    feats = feats * mask
    x = Fn.linear(feats, W)
    x = Fn.batch_norm(x.permute(0, 2, 1), bn["mean"], bn["var"], bn["weight"], bn["bias"], False, 0.01, 1e-3).permute(0, 2, 1)
    x = Fn.relu(x)
    return torch.max(x, dim=1, keepdim=True)[0].squeeze(1)


def eager_scatter(pillar_features, coords, B, ny, nx):
    C = pillar_features.shape[1]
    out = []
    for b in range(B):
        canvas = torch.zeros(C, ny * nx, dtype=pillar_features.dtype, device=pillar_features.device)
        m = coords[:, 0] == b
        c = coords[m]
        idx = (c[:, 1] + c[:, 2] * nx + c[:, 3]).long()
        canvas[:, idx] = pillar_features[m].t()
        out.append(canvas)
    return torch.stack(out, 0).view(B, C, ny, nx)


# ---- the reduced detector's dense part (random init; identical for both arms) ----------------------------------------------
class BevBackbone(nn.Module):
    def __init__(self, cin=64, nums=(3, 5, 5), strides=(2, 2, 2), filters=(64, 128, 256), ups=(1, 2, 4), upf=(128, 128, 128)):
        super().__init__()
        self.blocks, self.deblocks = nn.ModuleList(), nn.ModuleList()
        cins = [cin, *filters[:-1]]
        for i in range(len(nums)):
            layers = [nn.Conv2d(cins[i], filters[i], 3, stride=strides[i], padding=1, bias=False),
                      nn.BatchNorm2d(filters[i], eps=1e-3, momentum=0.01), nn.ReLU()]
            for _ in range(nums[i]):
                layers += [nn.Conv2d(filters[i], filters[i], 3, padding=1, bias=False),
                           nn.BatchNorm2d(filters[i], eps=1e-3, momentum=0.01), nn.ReLU()]
            self.blocks.append(nn.Sequential(*layers))
            self.deblocks.append(nn.Sequential(nn.ConvTranspose2d(filters[i], upf[i], ups[i], stride=ups[i], bias=False),
                                               nn.BatchNorm2d(upf[i], eps=1e-3, momentum=0.01), nn.ReLU()))
        c = sum(upf)
        self.cls, self.box, self.dir = nn.Conv2d(c, 6 * 3, 1), nn.Conv2d(c, 6 * 7, 1), nn.Conv2d(c, 6 * 2, 1)

    def forward(self, x):
        ups = []
        for blk, de in zip(self.blocks, self.deblocks):
            x = blk(x)
            ups.append(de(x))
        f = torch.cat(ups, 1)
        return self.cls(f), self.box(f), self.dir(f)


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    from oracle import oracle                      # the CPU voxelizer of the eager arm (= the reference's DataLoader work)
    B, n, P, maxv, mode = 16, 30000, int(os.environ.get("DET_P", "32")), 40000, os.environ.get("DET_MODE", "clustered")
    cfg = synthetic.CONFIGS["vod"]
    F = cfg["F"]
    dev = torch.device("cuda:0")
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    pts, offs = synthetic.make_batch("vod", B, n, mode, seed0=0)
    w = synthetic.make_pfn(F + 6, 64, 0)
    pfn = device_pfn(w, dev)
    path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], P, maxv, F)
    nx, ny = path.nx, path.ny

    # ---- eager arm inputs: CPU voxelization per frame (threads = host cores), collate, pinned ----
    from concurrent.futures import ThreadPoolExecutor
    threads = os.cpu_count() or 1

    def vox_frame(b):
        v, c, k = oracle.voxelize(pts[offs[b]:offs[b + 1]], geom, P, maxv, F=F, xcol=1)
        return v, np.concatenate([np.full((len(c), 1), b, np.int32), c], 1), k

    def cpu_voxelize():
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(vox_frame, range(B)))
        return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts]))

    cpu_voxelize()
    t0 = time.perf_counter()
    for _ in range(3):
        voxels, coords, nump = cpu_voxelize()
    cpu_vox_ms = (time.perf_counter() - t0) / 3 * 1e3
    hv, hc, hn = (torch.from_numpy(a).pin_memory() for a in (voxels, coords.astype(np.float32), nump.astype(np.float32)))
    hp = torch.from_numpy(pts).pin_memory()
    W = pfn.weight
    bn = dict(mean=pfn.running_mean, var=pfn.running_var, weight=pfn.bn_weight, bias=pfn.bn_bias)
    vs = [float(np.float32(v)) for v in cfg["voxel_size"]]
    of = [float(np.float32(v / 2 + r)) for v, r in zip(cfg["voxel_size"], cfg["pc_range"][:3])]

    net = BevBackbone().to(dev).eval()
    state = {}

    def eager_front():
        dv, dc, dn = hv.to(dev, non_blocking=True), hc.to(dev, non_blocking=True), hn.to(dev, non_blocking=True)   # load_data_to_gpu: float32
        pf = eager_pillar_vfe(dv, dn, dc, W, bn, vs, of)
        state["eager"] = eager_scatter(pf, dc, B, ny, nx)

    res = path.points_to_bev(hp.to(dev), B, pfn)

    def ours_front():
        dp = hp.to(dev, non_blocking=True)
        path.points_to_bev(dp, B, pfn, out=res)
        state["ours"] = res.spatial_features

    with torch.no_grad():
        ms_eager = timed(eager_front, 10)
        ms_ours = timed(ours_front, 50)
        diff = (state["eager"] - state["ours"]).abs().max().item()
        ref_max = state["eager"].abs().max().item()
        ms_net = timed(lambda: net(state["ours"]), 10)
        ms_det_eager = timed(lambda: (eager_front(), net(state["eager"])), 10)
        ms_det_ours = timed(lambda: (ours_front(), net(state["ours"])), 10)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms_net_bf16 = timed(lambda: net(state["ours"]), 10)
            ms_det_eager_bf16 = timed(lambda: (eager_front(), net(state["eager"])), 10)
            ms_det_ours_bf16 = timed(lambda: (ours_front(), net(state["ours"])), 10)
    fps = lambda ms: B / ms * 1e3
    out = dict(
        what="reduced detector (radar branch: pillar path -> BEV backbone 3/5/5 -> 1x1 heads), random init, synthetic points; "
             "the full HGSFusion forward is blocked (SURVEY 8c)",
        workload=f"vod_{mode}_b{B}_n{n}_P{P}", pillars=int(len(nump)), host_cores=threads,
        stage_ms=dict(cpu_voxelize_eager_arm=cpu_vox_ms, eager_h2d_vfe_scatter=ms_eager, ours_h2d_points_to_bev=ms_ours,
                      backbone_heads_fp32=ms_net, backbone_heads_bf16_autocast=ms_net_bf16),
        h2d_bytes=dict(eager=int(hv.numel() * 4 + hc.numel() * 4 + hn.numel() * 4), ours=int(hp.numel() * 4)),
        detector_frames_per_s=dict(
            fp32=dict(eager_cpu_voxelization_hidden=fps(ms_det_eager), eager_cpu_voxelization_serial=fps(ms_det_eager + cpu_vox_ms),
                      ours=fps(ms_det_ours)),
            bf16_autocast_backbone=dict(eager_cpu_voxelization_hidden=fps(ms_det_eager_bf16),
                                        eager_cpu_voxelization_serial=fps(ms_det_eager_bf16 + cpu_vox_ms), ours=fps(ms_det_ours_bf16))),
        pillar_path_speedup_vs_eager_gpu=ms_eager / ms_ours,
        canvas_max_abs_diff_vs_eager=diff, canvas_max_abs=ref_max)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
