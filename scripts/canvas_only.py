"""Times the path with no points (k_canvas writes an all-zero canvas) and at the bench shapes:
whole step, host issue time per call, k_canvas in-stream duration."""
import os, sys, time, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from hgsfusion_b200 import synthetic, _lib
from hgsfusion_b200.ops import PillarPath
from util import device_pfn
cfg = synthetic.CONFIGS["vod"]
dev = torch.device("cuda:0")
B = 16
lib = _lib.load()
path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], 32, 40000, 7)
pf = device_pfn(synthetic.make_pfn(13, 64), dev)
cases = [("n=0", torch.zeros((0, 8), device=dev))]
for mode in ("uniform", "clustered"):
    pts, _ = synthetic.make_batch("vod", B, 30000, mode)
    cases.append((mode, torch.from_numpy(pts).to(dev)))
for name, pts in cases:
    res = path.points_to_bev(pts, B, pf)
    for _ in range(5): path.points_to_bev(pts, B, pf, out=res)
    torch.cuda.synchronize()
    lib.hgsf_emit_timing_begin(50)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(50): path.points_to_bev(pts, B, pf, out=res)
    e1.record()
    t_issue = (time.perf_counter() - t0) / 50
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    buf = (C.c_float * 64)(); n = lib.hgsf_emit_timing_collect(buf, 64)
    kc = sum(buf[:n]) / max(n, 1)
    lib.hgsf_emit_timing_begin(0)
    print(f"{name:10s} step {ms*1e3:7.1f} us | host issue {t_issue*1e6:6.1f} us/call | k_canvas {kc*1e3:6.1f} us = {res.spatial_features.numel()*4/kc/1e6:.0f} GB/s canvas write")
