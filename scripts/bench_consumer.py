"""SURVEY 8(f) rank 4 measured: SpMiddlePillarEncoder18.conv1 (five 3x3 submanifold convolutions 32 -> 32 with BatchNorm,
residuals and ReLU, pcnres18.py:212-215) on the reader's pillar list, against the dense formulation a maintainer without spconv
would run on the same B200 -- SparseConvTensor.dense() + five cuDNN conv2d (+ BN + ReLU + active-set mask) in eager torch,
fp32 with TF32 off and on.  spconv itself is not in the image, so the reference's own sparse kernels cannot be timed.
usage: python scripts/bench_consumer.py [clustered|uniform] [B] [n]  -> one JSON line"""
import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
import torch.nn.functional as F
from hgsfusion_b200 import synthetic, pillar_ops as po

mode = sys.argv[1] if len(sys.argv) > 1 else "clustered"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30000
dev = torch.device("cuda:0")
cfg = synthetic.CONFIGS["vod"]
pts = synthetic.make_batch("vod", B, n, mode, seed0=0)[0]                      # [sum N, 1 + F], col 0 = frame
xyz = torch.from_numpy(np.ascontiguousarray(pts[:, 1:4] - np.asarray(cfg["pc_range"][:3], dtype=np.float32))).to(dev)
cnt = torch.from_numpy(np.bincount(pts[:, 0].astype(np.int64), minlength=B).astype(np.int32)).to(dev)
H = W = 320
r = po.gen_indice_pairs_flat(xyz, cnt, 0.16, (H, W))
pillars, bev = r["pillars"], r["pillar_bev_indices"]
M = int(pillars.shape[0])
torch.manual_seed(0)
feats = torch.rand((M, 32), device=dev)
enc = po.PillarEncoderConv1(32).to(dev).eval()

def timed(fn, iters=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

with torch.no_grad():
    ours = enc(feats, pillars, bev)
    t_ours = timed(lambda: enc(feats, pillars, bev))
    nbr = po.subm_neighbors(bev, pillars)
    t_nbr = timed(lambda: po.subm_neighbors(bev, pillars))
    c0 = getattr(enc, "0").conv0
    t_one = timed(lambda: po.subm_conv3x3(feats, nbr, c0[0].weight, c0[0].bias, c0[1], relu=True))

    # dense formulation of the same stage
    idx = pillars.long()
    mask = torch.zeros((B, 1, H, W), device=dev); mask[idx[:, 0], 0, idx[:, 1], idx[:, 2]] = 1
    def fold(seq):
        w = seq[0].weight.permute(0, 3, 1, 2).contiguous()
        bn = seq[1]
        s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        return w, seq[0].bias, s.view(1, -1, 1, 1), (bn.bias - bn.running_mean * s).view(1, -1, 1, 1)
    b0, b1 = getattr(enc, "0"), getattr(enc, "1")
    P = {k: fold(v) for k, v in dict(a=b0.conv0, b=b0.conv1, c=b0.conv2, d=b1.conv1, e=b1.conv2).items()}
    def cb(x, p, res=None):
        y = F.conv2d(x, p[0], p[1], padding=1) * p[2] + p[3]
        if res is not None: y = y + res
        return torch.relu(y) * mask                                        # submanifold: inactive cells stay empty
    def dense_stage():
        x = po.sparse_to_dense(feats, pillars, (H, W), B)
        i = cb(x, P["a"]); x = cb(cb(i, P["b"]), P["c"], i)
        return cb(cb(x, P["d"]), P["e"], x)
    res = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        d = dense_stage()
        got = d[idx[:, 0], :, idx[:, 1], idx[:, 2]]
        res[tf32] = (timed(dense_stage, iters=20, warm=3), (got - ours).abs().max().item())
    # ---- conv2: SparseConv2d(32, 64, 3, 2, 1) + BN + ReLU + 2 x Sparse2DBasicBlock(64) on the down-sampled active set ----
    enc2 = po.PillarEncoderConv2(32, 64).to(dev).eval()
    x2, p2, bev2 = enc2(ours, pillars, bev)
    M2 = int(p2.shape[0])
    t_ours2 = timed(lambda: enc2(ours, pillars, bev))
    t_idx2 = timed(lambda: po.sparse_conv_s2_indices(bev, pillars, sync=False))
    idx2 = p2.long()
    mask2 = torch.zeros((B, 1, H // 2, W // 2), device=dev); mask2[idx2[:, 0], 0, idx2[:, 1], idx2[:, 2]] = 1
    w0 = getattr(enc2, "0").weight.permute(0, 3, 1, 2).contiguous()
    bn0 = getattr(enc2, "1")
    s0 = (bn0.weight / torch.sqrt(bn0.running_var + bn0.eps)).view(1, -1, 1, 1)
    o0 = (bn0.bias.view(1, -1, 1, 1) - bn0.running_mean.view(1, -1, 1, 1) * s0)
    Q = {k: fold(v) for k, v in dict(a=getattr(enc2, "3").conv1, b=getattr(enc2, "3").conv2, c=getattr(enc2, "4").conv1, d=getattr(enc2, "4").conv2).items()}
    def cb2(x, p, res=None):
        y = F.conv2d(x, p[0], p[1], padding=1) * p[2] + p[3]
        if res is not None: y = y + res
        return torch.relu(y) * mask2
    def dense_stage2():
        x = po.sparse_to_dense(ours, pillars, (H, W), B)
        x = torch.relu(F.conv2d(x, w0, None, stride=2, padding=1) * s0 + o0) * mask2
        x = cb2(cb2(x, Q["a"]), Q["b"], x)
        return cb2(cb2(x, Q["c"]), Q["d"], x)
    res2 = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        d = dense_stage2()
        got = d[idx2[:, 0], :, idx2[:, 1], idx2[:, 2]]
        res2[tf32] = (timed(dense_stage2, iters=20, warm=3), (got - x2).abs().max().item())
fma = 9 * 32 * 32 * M
print(json.dumps({"what": "SpMiddlePillarEncoder18.conv1 on the pillar list vs the dense cuDNN formulation, same B200",
                  "workload": f"vod_{mode}_b{B}_n{n}", "pillars": M, "active_fraction": M / (B * H * W),
                  "ours_stage_ms": t_ours, "ours_rulebook_ms": t_nbr, "ours_one_conv_ms": t_one,
                  "ours_one_conv_tflops_fp32": 2 * fma / (t_one * 1e-3) / 1e12,
                  "dense_cudnn_fp32_stage_ms": res[False][0], "dense_cudnn_tf32_stage_ms": res[True][0],
                  "max_abs_diff_vs_dense_fp32": res[False][1], "max_abs_diff_vs_dense_tf32": res[True][1],
                  "speedup_vs_dense_fp32": res[False][0] / t_ours, "speedup_vs_dense_tf32": res[True][0] / t_ours,
                  "conv2": {"what": "SpMiddlePillarEncoder18.conv2 (stride-2 32->64 + BN + ReLU + 2 residual blocks at 64 channels)",
                            "out_pillars": M2, "ours_stage_ms": t_ours2, "ours_indices_ms": t_idx2,
                            "dense_cudnn_fp32_stage_ms": res2[False][0], "dense_cudnn_tf32_stage_ms": res2[True][0],
                            "max_abs_diff_vs_dense_fp32": res2[False][1], "max_abs_diff_vs_dense_tf32": res2[True][1],
                            "speedup_vs_dense_fp32": res2[False][0] / t_ours2, "speedup_vs_dense_tf32": res2[True][0] / t_ours2}}))
