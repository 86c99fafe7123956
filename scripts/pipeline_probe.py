"""Two steps in flight: step i+1's k_front (latency-bound, cooperative) under step i's k_emit (two streams, two slots of
workspace + outputs).  Prints ms/step of the serial loop and of the two-slot loop on the bench workload."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np, torch
from hgsfusion_b200 import synthetic
from hgsfusion_b200.ops import PillarPath
from util import device_pfn
cfg = synthetic.CONFIGS["vod"]
B, n, ring = 16, 30000, 6
dev = torch.device("cuda:0")
w = synthetic.make_pfn(13, 64)
pf = device_pfn(w, dev)
for mode in ["clustered", "uniform"]:
    batches = [torch.from_numpy(synthetic.make_batch("vod", B, n, mode, seed0=r * B)[0]).to(dev) for r in range(ring)]
    paths = [PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], 32, 40000, 7) for _ in range(2)]
    res = [p.points_to_bev(batches[0], B, pf) for p in paths]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    def serial(steps):
        for i in range(steps):
            paths[0].points_to_bev(batches[i % ring], B, pf, out=res[0])
    def piped(steps):
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event(); fork.record(main)
        for s in streams: s.wait_event(fork)
        for i in range(steps):
            with torch.cuda.stream(streams[i & 1]):
                paths[i & 1].points_to_bev(batches[i % ring], B, pf, out=res[i & 1])
        for s in streams:
            j = torch.cuda.Event(); j.record(s); main.wait_event(j)
    for name, fn in (("serial", serial), ("two-slot", piped)):
        fn(10); torch.cuda.synchronize()
        e0.record(); fn(200); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 200
        print(f"{mode:10s} {name:9s} ms/step {ms:.4f}  frames/s {B / ms * 1e3:.0f}", flush=True)
