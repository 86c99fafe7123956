"""Path B (PillarNet reader) forward: the REFERENCE's own CUDA kernels (oracle/_ref/libref_pillar_ops.so, compiled unmodified
from pcdet/ops/pillar_ops/src) inside the reference's Python glue (pillar_utils.py:99-124, group_utils.py:20-29,
pillar_modules.py:74-82, torch for the MLP) against this repo's native reader, same GPU, same inputs.
Prints one JSON line; wall-clock per forward with a device synchronize on both sides (both sides contain host syncs)."""
import ctypes as C
import json
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import torch

from hgsfusion_b200 import pillar_ops as po

REF_SO = os.path.join(R, "oracle", "_ref", "libref_pillar_ops.so")
dev = torch.device("cuda:0")
B, n, H, W, s, Cf = 16, 30000, 320, 320, 0.16, 29
rng = np.random.default_rng(0)
N = B * n
xyz_np = np.stack([rng.uniform(0, W * s, N), rng.uniform(0, H * s, N), rng.uniform(0, 5, N)], axis=1).astype(np.float32)
k = int(0.9 * N)                                     # 90 % of the points in 40 blobs per frame, like the hybrid radar clouds
centres = rng.uniform(0, W * s, size=(B * 40, 2))
blob = rng.integers(0, 40, k) + 40 * (np.arange(k) * B // k)
xyz_np[:k, :2] = (centres[blob] + rng.normal(0, 0.6, size=(k, 2))).astype(np.float32)
order = np.argsort(np.concatenate([np.arange(k) * B // k, rng.integers(0, B, N - k)]), kind="stable")
xyz_np = xyz_np[order]
cnt_np = np.bincount(np.sort(np.concatenate([np.arange(k) * B // k, rng.integers(0, B, N - k)])), minlength=B).astype(np.int32)
xyz = torch.from_numpy(xyz_np).to(dev)
cnt = torch.from_numpy(cnt_np).to(dev)
pf = torch.randn((N, Cf), device=dev)
m = po.PillarMaxPooling([Cf + 6, 32], s, [0, -25.6, -3, 51.2, 25.6, 2]).to(dev).eval()
with torch.no_grad():
    m.shared_mlps[1].running_mean.normal_(); m.shared_mlps[1].running_var.uniform_(0.5, 2.0)
p = lambda t: C.c_void_p(t.data_ptr())


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


out = dict(workload=f"PillarNet reader forward, VoD 320x320, B={B}, N={n}/frame, Cf={Cf} -> 32 channels", unit="ms per forward")
with torch.no_grad():
    ours = lambda: m(xyz, cnt, pf)
    res = ours()
    feats = res[0] if isinstance(res, tuple) else res.features
    out["native_fused_ms"] = timed(ours)
    out["native_indices_ms"] = timed(lambda: po.gen_indice_pairs_flat(xyz, cnt, s, (H, W)))
    out["M"] = int(feats.shape[0])
    if os.path.exists(REF_SO):
        ref = C.CDLL(REF_SO)

        def ref_indices():
            mask = torch.zeros((B, H, W), dtype=torch.bool, device=dev)
            ref.ref_create_pillar_indices_stack(N, B, H, W, C.c_float(s), p(xyz), p(cnt), p(mask))
            location = torch.cumsum(mask.view(-1), 0).int()
            M = location[-1].item()
            bev = (location.view(B, H, W) * mask - 1).int().contiguous()
            pillars = torch.zeros((M, 3), dtype=torch.int32, device=dev)
            ref.ref_create_pillar_indices(B, H, W, p(bev), p(pillars))
            pairs = torch.full((N, 1), -1, dtype=torch.int32, device=dev)
            ref.ref_create_pillar_indice_pairs_stack(N, B, H, W, C.c_float(s), p(xyz), p(cnt), p(bev), p(pairs))
            valid = pairs.view(-1) > -1
            position = torch.cumsum(valid, 0).int()
            L = position[-1].item()
            position = (position * valid - 1).int().contiguous()
            first = torch.zeros(L, dtype=torch.int32, device=dev)
            second = torch.zeros(L, dtype=torch.int32, device=dev)
            ref.ref_flatten_indice_pairs(N, 1, p(pairs), p(position), p(first), p(second))
            return pillars, first, second, M, L

        def ref_gather(src, idx, L):
            o = torch.zeros((L, src.shape[1]), device=dev)
            ref.ref_gather_feature(L, src.shape[1], p(idx), p(src), p(o))
            return o

        def ref_forward():
            pillars, first, second, M, L = ref_indices()
            centers = torch.zeros([M, 3], dtype=torch.float32, device=dev)
            centers[:, 0] = (pillars[:, 2] + 0.5) * s
            centers[:, 1] = (pillars[:, 1] + 0.5) * s
            centers[:, 2] = -0.5
            gp, gx, gc = ref_gather(pf, first, L), ref_gather(xyz, first, L), ref_gather(centers, second, L)
            g = torch.cat([gp, gx, gx - gc], dim=1)
            h = m.shared_mlps(g).transpose(1, 0).contiguous()
            arg = torch.full((32, M), -1, dtype=torch.int32, device=dev)
            o = torch.zeros((32, M), device=dev)
            ref.ref_scatter_max(32, L, M, p(second), p(h), p(arg), p(o))
            return o.transpose(1, 0)

        exp = ref_forward()
        torch.cuda.synchronize()
        out["max_abs_diff_vs_reference"] = float((feats - exp).abs().max())
        out["reference_kernels_ms"] = timed(ref_forward)
        out["reference_indices_ms"] = timed(ref_indices)
        out["speedup_forward"] = out["reference_kernels_ms"] / out["native_fused_ms"]
        out["speedup_indices"] = out["reference_indices_ms"] / out["native_indices_ms"]
    else:
        out["reference_kernels_ms"] = None
print(json.dumps(out))
