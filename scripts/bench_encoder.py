"""The whole sparse encoder behind the reader (SpMiddlePillarEncoder18 conv1 .. conv4, pcnres18.py:200-285) on pillar lists, stage by
stage, against the dense formulation (SparseConvTensor.dense() + cuDNN conv2d + BN + ReLU + active-set mask, eager torch, fp32 with
TF32 off / on) on the same B200.  usage: python scripts/bench_encoder.py [clustered|uniform] [B] [n]  -> one JSON line"""
import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np, torch
import torch.nn.functional as F
from hgsfusion_b200 import synthetic, pillar_ops as po

mode = sys.argv[1] if len(sys.argv) > 1 else "clustered"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30000
dev = torch.device("cuda:0")
cfg = synthetic.CONFIGS["vod"]
pts = synthetic.make_batch("vod", B, n, mode, seed0=0)[0]
xyz = torch.from_numpy(np.ascontiguousarray(pts[:, 1:4] - np.asarray(cfg["pc_range"][:3], dtype=np.float32))).to(dev)
cnt = torch.from_numpy(np.bincount(pts[:, 0].astype(np.int64), minlength=B).astype(np.int32)).to(dev)
H = W = 320
r = po.gen_indice_pairs_flat(xyz, cnt, 0.16, (H, W))
pillars, bev = r["pillars"], r["pillar_bev_indices"]
torch.manual_seed(0)
feats = torch.rand((pillars.shape[0], 32), device=dev)
enc = po.SpMiddlePillarEncoder18(32, out_indices=(0, 1, 2, 3)).to(dev).eval()


def timed(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def fold(seq):
    conv, bn = seq
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return conv.weight.permute(0, 3, 1, 2).contiguous(), conv.bias, s.view(1, -1, 1, 1), (bn.bias - bn.running_mean * s).view(1, -1, 1, 1)


def cb(x, p, msk, res=None, stride=1):
    y = F.conv2d(x, p[0], p[1], stride=stride, padding=1) * p[2] + p[3]
    if res is not None: y = y + res
    return torch.relu(y) * msk


out = {"what": "SpMiddlePillarEncoder18 conv1..conv4 on pillar lists vs dense cuDNN, same B200", "workload": f"vod_{mode}_b{B}_n{n}", "stages": []}
with torch.no_grad():
    outs = enc(feats, pillars, bev)
    stage_in = [(feats, pillars, bev)] + outs[:3]
    stages = [enc.conv1, enc.conv2, enc.conv3, enc.conv4]
    for si, (stage, inp, res) in enumerate(zip(stages, stage_in, outs)):
        f_in, p_in, bev_in = inp
        f_out, p_out, bev_out = res
        t_ours = timed(lambda: stage(f_in, p_in, bev_in))
        Hi, Wi, Ho, Wo = bev_in.shape[1], bev_in.shape[2], bev_out.shape[1], bev_out.shape[2]
        io = p_out.long()
        msk = torch.zeros((B, 1, Ho, Wo), device=dev); msk[io[:, 0], 0, io[:, 1], io[:, 2]] = 1
        if si == 0:
            b0, b1 = getattr(stage, "0"), getattr(stage, "1")
            Pm = [fold(b0.conv0), fold(b0.conv1), fold(b0.conv2), fold(b1.conv1), fold(b1.conv2)]

            def dense():
                x = po.sparse_to_dense(f_in, p_in, (Hi, Wi), B)
                i = cb(x, Pm[0], msk); x = cb(cb(i, Pm[1], msk), Pm[2], msk, i)
                return cb(cb(x, Pm[3], msk), Pm[4], msk, x)
        else:
            Pm = [fold((getattr(stage, "0"), getattr(stage, "1")))] + [fold(getattr(getattr(stage, k), c)) for k in ("3", "4") for c in ("conv1", "conv2")]

            def dense():
                x = po.sparse_to_dense(f_in, p_in, (Hi, Wi), B)
                x = cb(x, Pm[0], msk, stride=2)
                x = cb(cb(x, Pm[1], msk), Pm[2], msk, x)
                return cb(cb(x, Pm[3], msk), Pm[4], msk, x)
        ent = dict(stage=f"conv{si + 1}", in_pillars=int(p_in.shape[0]), out_pillars=int(p_out.shape[0]), channels=int(f_out.shape[1]),
                   active_fraction_out=float(p_out.shape[0]) / (B * Ho * Wo), ours_ms=t_ours)
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            d = dense()
            got = d[io[:, 0], :, io[:, 1], io[:, 2]]
            key = "tf32" if tf32 else "fp32"
            ent[f"dense_cudnn_{key}_ms"] = timed(dense, iters=10, warm=2)
            ent[f"max_abs_diff_vs_dense_{key}"] = (got - f_out).abs().max().item()
        ent["speedup_vs_dense_fp32"] = ent["dense_cudnn_fp32_ms"] / t_ours
        ent["speedup_vs_dense_tf32"] = ent["dense_cudnn_tf32_ms"] / t_ours
        out["stages"].append(ent)
    out["ours_total_ms"] = timed(lambda: enc(feats, pillars, bev))
print(json.dumps(out))
