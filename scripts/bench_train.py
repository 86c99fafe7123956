"""Train-step timing of FusedPillarVFE: the fused train path (hgsf_points_to_bev_train, 3 + 4 launches, no read-back) against the
contract-layout chain (pillarize -> trim -> batch statistics -> PFN -> scatter): python scripts/bench_train.py [config mode B n]"""
import json, os, sys
from types import SimpleNamespace
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
from hgsfusion_b200 import modules, synthetic

cfgname, mode, B, n = (sys.argv[1:5] + ["vod", "clustered", "16", "30000"][len(sys.argv) - 1:])[:4]
B, n = int(B), int(n)
cfg = synthetic.CONFIGS[cfgname]
dev = torch.device("cuda:0")
pts, _ = synthetic.make_batch(cfgname, B, n, mode)
d = torch.from_numpy(pts).to(dev)
mc = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64], MAX_POINTS_PER_VOXEL=32,
                     MAX_NUMBER_OF_VOXELS={'train': 40000, 'test': 40000}, TRIM=False)
m = modules.FusedPillarVFE(model_cfg=mc, num_point_features=cfg["F"], voxel_size=list(cfg["voxel_size"]),
                           point_cloud_range=np.array(cfg["pc_range"], dtype=np.float32)).to(dev).train()
Rc = None


def run(fused):
    """GPU time of the forward and of the backward with the launches already queued (a spin kernel in front keeps the GPU busy
    while the host enqueues), and the host time of the same calls with the GPU idle."""
    global Rc
    import time
    K = 20
    gf = gb = hf = hb = 0.0
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for i in range(-5, K):
        m.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        torch.cuda._sleep(6_000_000)                      # ~3 ms: the host runs ahead
        e0.record()
        out = m._forward_train_fused(d, B) if fused else m._forward_train(d, B)
        e1.record()
        torch.cuda.synchronize()
        if Rc is None:
            Rc = torch.randn_like(out['spatial_features'])
        torch.cuda._sleep(6_000_000)
        e2.record()
        torch.autograd.backward([out['spatial_features']], [Rc])
        e3.record()
        torch.cuda.synchronize()
        # host time, GPU idle
        m.zero_grad(set_to_none=True)
        t0 = time.perf_counter()
        out = m._forward_train_fused(d, B) if fused else m._forward_train(d, B)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        torch.autograd.backward([out['spatial_features']], [Rc])
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        if i >= 0:
            gf += e0.elapsed_time(e1); gb += e2.elapsed_time(e3); hf += (t1 - t0) * 1e3; hb += (t3 - t2) * 1e3
    return dict(forward_gpu_ms=round(gf / K, 4), backward_gpu_ms=round(gb / K, 4), forward_host_ms=round(hf / K, 4),
                backward_host_ms=round(hb / K, 4))


res = {"workload": f"{cfgname}_{mode}_b{B}_n{n}"}
for name, fused in (("fused_train", True), ("contract_chain", False)):
    res[name] = run(fused)
res["note"] = ("train-mode FusedPillarVFE, TRIM False; forward = points -> canvas on batch statistics (fused: 3 launches, no read-back; "
               "contract chain: pillarize -> trim (host sync) -> batch statistics -> PFN -> scatter); backward = canvas cotangent -> "
               "grads of linear.weight / norm.weight / norm.bias, no loss kernels; *_gpu_ms = device time with the launches queued behind a "
               "spin kernel (the contract chain's forward contains a host sync, so its figure includes that round trip); *_host_ms = time "
               "of the Python call with the GPU idle")
print(json.dumps(res))
