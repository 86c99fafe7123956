"""one 32 -> 32 submanifold convolution on the bench workload's pillar list: ms and fp32 TFLOP/s (HGSF_SUBM_VARIANT with an
experiment build selects the thread tile)"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np, torch
from hgsfusion_b200 import synthetic, pillar_ops as po
dev = torch.device("cuda:0")
for mode in ("clustered", "uniform"):
    pts = synthetic.make_batch("vod", 16, 30000, mode, seed0=0)[0]
    xyz = torch.from_numpy(np.ascontiguousarray(pts[:, 1:4] - np.asarray(synthetic.CONFIGS["vod"]["pc_range"][:3], dtype=np.float32))).to(dev)
    cnt = torch.from_numpy(np.bincount(pts[:, 0].astype(np.int64), minlength=16).astype(np.int32)).to(dev)
    r = po.gen_indice_pairs_flat(xyz, cnt, 0.16, (320, 320))
    nbr = po.subm_neighbors(r["pillar_bev_indices"], r["pillars"])
    M = nbr.shape[0]
    torch.manual_seed(0)
    CH = int(os.environ.get("SUBM_CH", "32"))
    if CH == 64:      # the conv2 stage's active set
        r2 = po.sparse_conv_s2_indices(r["pillar_bev_indices"], r["pillars"])
        nbr = po.subm_neighbors(r2["pillar_bev_indices"], r2["pillars"]); M = nbr.shape[0]
    f = torch.rand((M, CH), device=dev); w = torch.randn((CH, 3, 3, CH), device=dev) * 0.1
    ref = None
    out = torch.empty((M, CH), device=dev)
    for layout, wt in (("KRSC", w), ("RSCK", w.permute(1, 2, 3, 0).contiguous())):
        for _ in range(5): po.subm_conv3x3(f, nbr, wt, relu=True, out=out, weight_layout=layout)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100): po.subm_conv3x3(f, nbr, wt, relu=True, out=out, weight_layout=layout)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 100
        print(f"variant {os.environ.get('HGSF_SUBM_VARIANT', '0')} {layout} {mode} M={M}: {ms:.4f} ms  {2 * 9 * CH * CH * M / ms / 1e9:.1f} TFLOP/s  checksum {out.double().sum().item():.6f}")
