#!/usr/bin/env python
"""bench.py -- points -> BEV frames/s of the radar pillarization hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 200 --warmup 10          # our arm, one B200
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 20 --warmup 3    # the reference's own CPU modules (baseline/_ref/) on host cores

A "step" is one pass of the hot path (points -> pillars -> PillarVFE -> PointPillarScatter canvas) over one
batch of synthetic frames.  Default workload = BASELINE.json configs[1]: VoD hybrid-point density, 30 000 points
per frame, 7 features, 0.16 m pillars on 320x320, C = 64, batch 16 per GPU (weak scaling: every rank owns its own
16 frames, no data-path collective).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from hgsfusion_b200 import synthetic  # noqa: E402

METRIC = "points_to_bev_frames_per_sec"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="vod", choices=list(synthetic.CONFIGS))
    ap.add_argument("--mode", default="clustered", choices=["clustered", "uniform"])
    ap.add_argument("--batch", type=int, default=16, help="frames per GPU per step")
    ap.add_argument("--points", type=int, default=30000, help="points per frame")
    ap.add_argument("--max-points", type=int, default=32)
    ap.add_argument("--max-voxels", type=int, default=40000)
    ap.add_argument("--ring", type=int, default=10, help="distinct input batches cycled through (defeats L2 reuse)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(args):
    """The `config` object of BOTH arms (ours and --impl reference): everything in it follows from the arguments, so the two
    lines carry the same dictionary.  `l2` / `parallelism` describe the GPU arm's timed region."""
    cfg = synthetic.CONFIGS[args.config]
    rng, vs = cfg["pc_range"], cfg["voxel_size"]
    nx, ny = int(round((rng[3] - rng[0]) / vs[0])), int(round((rng[4] - rng[1]) / vs[1]))
    ring = max(1, args.ring)
    in_mb = ring * args.batch * args.points * (1 + cfg["F"]) * 4 / 1e6
    canvas_mb = args.batch * 64 * ny * nx * 4 / 1e6
    return dict(workload=f"{args.config}_{args.mode}_b{args.batch}_n{args.points}", dataset_shape=args.config,
                points_per_frame=args.points, point_features=cfg["F"], frames_per_gpu_per_step=args.batch,
                max_points_per_voxel=args.max_points, max_voxels=args.max_voxels, channels=64,
                pc_range=cfg["pc_range"], voxel_size=cfg["voxel_size"], point_distribution=args.mode, grid=[nx, ny, 1],
                l2=f"ring of {ring} distinct input batches ({in_mb:.0f} MB) and a {canvas_mb:.0f} MB canvas rewritten every step, "
                   f"both > 126 MB L2",
                parallelism=f"frames sharded by rank, {args.batch} per GPU, no collective")


def alg_bytes_per_frame(n, F, C, ny, nx, M):
    """points read + canvas write + pillar_features + voxel_coords + voxel_num_points."""
    return 4 * n * F + 4 * C * ny * nx + 4 * M * C + 16 * M + 4 * M


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is working (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, f"/tmp/hgsf_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        # keep the samples taken under load (upper half of the observed clocks)
        if sm:
            top = sorted(sm)[len(sm) // 2:]
            return {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons)}
        return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def _ref_frames(args, sample_frames):
    cfg = synthetic.CONFIGS[args.config]
    return [synthetic.make_frame(args.points, cfg["pc_range"], cfg["F"], seed, args.mode) for seed in range(sample_frames)]


def cpu_reference_arm(args, steps, warmup, threads=None, sample_frames=None):
    """The reference's CPU path for this workload: per-frame spconv-style voxelization (the C restatement of
    the spconv loop, one frame per thread as the DataLoader workers do), collate, then the REFERENCE'S OWN PillarVFE and
    PointPillarScatter (pillar_vfe.py:52-123, pointpillar_scatter.py:5-41, staged unmodified into baseline/_ref/ and imported by
    path) in eval mode on torch CPU with `threads` intra-op threads.  A step = `sample_frames` frames of the bench workload (a
    bounded sample: the full batch would take seconds per step).  Falls back to the all-C oracle port when baseline/_ref/ is absent."""
    import concurrent.futures as cf
    from types import SimpleNamespace

    from baseline import ref_modules
    from oracle import oracle
    cfg = synthetic.CONFIGS[args.config]
    F = cfg["F"]
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    threads = threads or (os.cpu_count() or 1)
    S = sample_frames or min(args.batch, 2)
    frames = _ref_frames(args, S)
    w = synthetic.make_pfn(F + 6, 64, 0)
    if not ref_modules.available():
        # all-C port (voxelize + VFE + scatter in oracle/pillar_oracle.c)
        oracle.set_num_threads(threads)
        pfn = oracle.PfnParams(w.weight, w.gamma, w.beta, w.running_mean, w.running_var)
        pts, offs = synthetic.batch_points(frames)
        run = lambda: oracle.points_to_bev(pts, offs, geom, pfn, args.max_points, args.max_voxels, F=F, xcol=1, want_voxels=False)
        kind, how = "port", f"oracle/pillar_oracle.c, {threads} pthreads (baseline/_ref/ absent)"
    else:
        import torch
        torch.set_num_threads(threads)
        PillarVFE, PointPillarScatter = ref_modules.load()
        model_cfg = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64])
        vfe = PillarVFE(model_cfg=model_cfg, num_point_features=F, voxel_size=cfg["voxel_size"],
                        point_cloud_range=np.asarray(cfg["pc_range"], dtype=np.float32)).eval()
        with torch.no_grad():
            vfe.pfn_layers[0].linear.weight.copy_(torch.from_numpy(w.weight))
            bn = vfe.pfn_layers[0].norm
            bn.weight.copy_(torch.from_numpy(w.gamma)); bn.bias.copy_(torch.from_numpy(w.beta))
            bn.running_mean.copy_(torch.from_numpy(w.running_mean)); bn.running_var.copy_(torch.from_numpy(w.running_var))
        scatter = PointPillarScatter(model_cfg=SimpleNamespace(NUM_BEV_FEATURES=64), grid_size=np.asarray(geom.grid)).eval()
        pool = cf.ThreadPoolExecutor(max_workers=min(threads, S))
        oracle.set_num_threads(1)

        def run():
            # transform_points_to_voxels per frame (data_processor.py:133-183), collate_batch (dataset.py:232-244),
            # load_data_to_gpu's .float() (models/__init__.py:23-36), then the two reference modules
            vox = list(pool.map(lambda f: oracle.voxelize(f, geom, args.max_points, args.max_voxels, F=F, xcol=0), frames))
            voxels = np.concatenate([v[0] for v in vox])
            coords = np.concatenate([np.concatenate([np.full((v[1].shape[0], 1), b, np.int32), v[1]], axis=1) for b, v in enumerate(vox)])
            num = np.concatenate([v[2] for v in vox])
            bd = dict(voxels=torch.from_numpy(voxels), voxel_coords=torch.from_numpy(coords).float(),
                      voxel_num_points=torch.from_numpy(num).float())
            with torch.no_grad():
                bd = scatter(vfe(bd))
            return bd["spatial_features"]
        kind = "_ref+port"
        how = (f"voxelizer = oracle/pillar_oracle.c (spconv restated; one frame per thread), PillarVFE + PointPillarScatter = the reference's "
               f"own modules from baseline/_ref/ on torch CPU, {threads} intra-op threads")
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    return dict(value=S * steps / dt, unit=UNIT, cores=threads, kind=kind,
                sample=f"{steps} steps x {S} frames of the bench workload ({dt:.1f} s); {how}"), dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # every step is a bounded sample (2 frames of the workload) so that K steps end within a few minutes on any host
    steps, warmup = max(1, args.steps), max(args.warmup, 3)
    base, dt = cpu_reference_arm(args, steps, warmup)
    # the same with one torch thread (BASELINE.md section 4 asks for both), on a shorter run
    one, _ = cpu_reference_arm(args, max(1, min(steps, 10)), 1, threads=1)
    base["single_thread"] = dict(value=one["value"], unit=UNIT, cores=1, sample=one["sample"])
    line = dict(metric=METRIC, value=base["value"], unit=UNIT, impl="reference", n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=1e3 * dt / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=workload(args), cpu_baseline=base,
                e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                note="reference arm = the reference's CPU implementation of the path on host cores: its own PillarVFE / PointPillarScatter "
                     "modules (staged from /root/reference into baseline/_ref/) behind the C restatement of spconv's voxelizer (spconv "
                     "itself is not installed anywhere); a step is a bounded sample of the workload, ms_per_step is per sample")
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local: int) -> dict:
    """Pins this rank's threads to the CPUs of its GPU's NUMA node BEFORE the pinned host buffers are allocated (first
    touch then places them on that node).  Returns what was found / done for the record."""
    info = dict(numa_node=None, cpus=None, bound=False)
    try:
        import torch
        props = torch.cuda.get_device_properties(local)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info["numa_node"] = node
        if node >= 0:
            cpulist = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
            cpus = set()
            for part in cpulist.split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            allowed = cpus & os.sched_getaffinity(0)
            if allowed:
                os.sched_setaffinity(0, allowed)
                info["cpus"], info["bound"] = cpulist, True
    except Exception as exc:  # topology not exposed in the container: leave the affinity alone
        info["error"] = type(exc).__name__
    return info


def gpu_reference_arm(args, path, dpts, B, w, dev, steps=10):
    """The honest GPU comparator: the reference's own PillarVFE + PointPillarScatter modules (baseline/_ref/) in eager torch
    on the SAME GPU, fed with the voxels / coords / counts tensors our pillarizer produced on the device (the reference gets
    them from spconv on the CPU and copies 20x more bytes over PCIe; neither is charged here).  CUDA events, frames/s."""
    import torch
    from types import SimpleNamespace

    from baseline import ref_modules
    if not ref_modules.available():
        return None
    cfg = synthetic.CONFIGS[args.config]
    F = cfg["F"]
    PillarVFE, PointPillarScatter = ref_modules.load()
    model_cfg = SimpleNamespace(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64])
    vfe = PillarVFE(model_cfg=model_cfg, num_point_features=F, voxel_size=cfg["voxel_size"],
                    point_cloud_range=np.asarray(cfg["pc_range"], dtype=np.float32)).to(dev).eval()
    with torch.no_grad():
        vfe.pfn_layers[0].linear.weight.copy_(torch.from_numpy(w.weight))
        bn = vfe.pfn_layers[0].norm
        bn.weight.copy_(torch.from_numpy(w.gamma)); bn.bias.copy_(torch.from_numpy(w.beta))
        bn.running_mean.copy_(torch.from_numpy(w.running_mean)); bn.running_var.copy_(torch.from_numpy(w.running_var))
    scatter = PointPillarScatter(model_cfg=SimpleNamespace(NUM_BEV_FEATURES=64), grid_size=np.asarray([path.nx, path.ny, path.nz])).eval()
    r = path.pillarize(dpts, B, want_voxels=True).trim()
    bd0 = dict(voxels=r["voxels"], voxel_coords=r["voxel_coords"].float(), voxel_num_points=r["voxel_num_points"].float())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for _ in range(3):
            scatter(vfe(dict(bd0)))
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            scatter(vfe(dict(bd0)))
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del vfe, scatter, bd0, r
    torch.cuda.empty_cache()
    return dict(value=B / ms * 1e3, unit=UNIT, ms_per_step=ms, steps=steps, kind="reference modules, eager torch, same GPU",
                note="the reference's own PillarVFE + PointPillarScatter (baseline/_ref/) on this GPU from device-resident voxels; "
                     "its CPU voxelization and the H2D copy of the padded voxels tensor are NOT included")


def main_ours(args):
    import torch
    import torch.distributed as dist

    from hgsfusion_b200 import _lib, sharding
    from hgsfusion_b200.ops import PfnWeights, PillarPath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hgsfusion_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)             # before any pinned allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    cfg = synthetic.CONFIGS[args.config]
    F, B, n = cfg["F"], args.batch, args.points
    path = PillarPath(np.asarray(cfg["pc_range"], dtype=np.float32), cfg["voxel_size"], args.max_points, args.max_voxels, F)
    nx, ny = path.nx, path.ny
    w = synthetic.make_pfn(F + 6, 64, 0)
    t = lambda a: torch.from_numpy(a).to(dev)
    pfn = PfnWeights(weight=t(w.weight), bn_weight=t(w.gamma), bn_bias=t(w.beta), running_mean=t(w.running_mean),
                     running_var=t(w.running_var))

    # a ring of distinct input batches: weak scaling -> every rank owns its own frames (seeds differ per rank)
    ring = max(1, args.ring)
    host = [synthetic.make_batch(args.config, B, n, args.mode, seed0=(rank * ring + r) * B)[0] for r in range(ring)]
    dpts = [torch.from_numpy(h).to(dev) for h in host]
    res = path.points_to_bev(dpts[0], B, pfn)                      # allocates outputs + workspace once
    torch.cuda.synchronize()
    M = int(res.num_pillars[0].item())
    launches_per_step = path.last_launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- device-resident throughput: W warm-up, K timed steps, CUDA events, max over ranks.  Nothing but the library's own
    #      launches is enqueued inside the timed region (no per-kernel events: they cost ~8 us per step and break the programmatic
    #      dependent launch of the second kernel) ----
    warm = max(args.warmup, 3)
    for i in range(warm):
        path.points_to_bev(dpts[i % ring], B, pfn, out=res)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        path.points_to_bev(dpts[i % ring], B, pfn, out=res)
    e1.record()
    barrier()
    ms_dev = sharding.reduce_max(e0.elapsed_time(e1), dev)

    # ---- the same K steps once more with a CUDA event pair around the dominant kernel (k_emit) on the launch stream: the
    #      roofline leg.  Its step time is reported too (instrumented_ms_per_step) ----
    lib.hgsf_emit_timing_begin(min(args.steps, 1024))
    barrier()
    e0.record()
    for i in range(args.steps):
        path.points_to_bev(dpts[i % ring], B, pfn, out=res)
    e1.record()
    barrier()
    ms_inst = e0.elapsed_time(e1) / args.steps
    buf = (C.c_float * 1024)()
    n_ev = lib.hgsf_emit_timing_collect(buf, 1024)
    emit_ms = statistics.fmean(buf[:n_ev]) if n_ev > 0 else None
    lib.hgsf_emit_timing_begin(0)

    # ---- end to end through the public API with HOST buffers: pinned H2D of the step's points, the native call, D2H of the
    #      step's result summary (pillar counts); the canvas stays on the device for the 2D backbone, as
    #      batch_dict['spatial_features'] does in the reference.  Copies ride a second stream, double buffered.
    #      Default layout = the frames' own rows [sum N, F] + frame_offsets (hgsf_points.frame_offsets): what a collate hook
    #      that does not add the batch column ships (12.5 % fewer bytes than the collated [sum N, 1+F]) ----
    copy_stream = torch.cuda.Stream(device=dev)
    pinned_nf = [torch.from_numpy(np.ascontiguousarray(h[:, 1:])).pin_memory() for h in host]
    pinned_col = [torch.from_numpy(h).pin_memory() for h in host]
    stage_nf = [torch.empty((pinned_nf[0].shape[0], F), dtype=torch.float32, device=dev) for _ in range(2)]
    stage_col = [torch.empty_like(dpts[0]) for _ in range(2)]
    offs_dev = torch.arange(0, (B + 1) * n, n, dtype=torch.int32, device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    counts_host = [torch.empty(1 + B, dtype=torch.int32).pin_memory() for _ in range(2)]
    main_stream = torch.cuda.current_stream()
    call_nf = lambda pts_dev: path.points_to_bev(pts_dev, B, pfn, xyz_col=0, frame_offsets=offs_dev, out=res)
    call_col = lambda pts_dev: path.points_to_bev(pts_dev, B, pfn, out=res)

    def e2e_loop(steps, srcs, stages, call, with_canvas=False, canvas_host=None, copy_only=False):
        for s_ in consumed:
            s_.record(main_stream)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[0])
            stages[0].copy_(srcs[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(steps):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < steps:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[nxt])
                    stages[nxt].copy_(srcs[(i + 1) % ring], non_blocking=True)
                    ready[nxt].record(copy_stream)
            main_stream.wait_event(ready[cur])
            if not copy_only:
                call(stages[cur])
            consumed[cur].record(main_stream)
            if not copy_only:
                counts_host[cur].copy_(res.num_pillars, non_blocking=True)
                if with_canvas:
                    canvas_host.copy_(res.spatial_features, non_blocking=True)

    def timed(steps, **kw):
        e2e_loop(3, **kw)
        barrier()
        e0.record()
        e2e_loop(steps, **kw)
        e1.record()
        barrier()
        return sharding.reduce_max(e0.elapsed_time(e1), dev)

    e2e_steps = max(10, min(args.steps, 200))
    ms_e2e = timed(e2e_steps, srcs=pinned_nf, stages=stage_nf, call=call_nf)
    assert int(counts_host[(e2e_steps - 1) & 1][0]) > 0
    ms_e2e_col = timed(e2e_steps, srcs=pinned_col, stages=stage_col, call=call_col)
    # copy-only: the same pinned buffers and stream choreography without the kernels -> the H2D rate this rank gets while every
    # other rank copies too (separates a host / PCIe ceiling from a software one)
    ms_copy = timed(e2e_steps, srcs=pinned_nf, stages=stage_nf, call=call_nf, copy_only=True)
    h2d = int(pinned_nf[0].numel() * 4)
    d2h = int((1 + B) * 4)
    h2d_gbs_rank = h2d * e2e_steps / (ms_e2e * 1e-3) / 1e9
    copy_gbs_rank = h2d * e2e_steps / (ms_copy * 1e-3) / 1e9
    # the same with the whole canvas also copied to the host every step (PCIe bound; for the record only)
    canvas_steps = 5
    canvas_host = torch.empty(res.spatial_features.shape, dtype=torch.float32).pin_memory()
    e2e_loop(1, srcs=pinned_nf, stages=stage_nf, call=call_nf, with_canvas=True, canvas_host=canvas_host)
    barrier()
    e0.record()
    e2e_loop(canvas_steps, srcs=pinned_nf, stages=stage_nf, call=call_nf, with_canvas=True, canvas_host=canvas_host)
    e1.record()
    barrier()
    ms_e2e_canvas = sharding.reduce_max(e0.elapsed_time(e1), dev)
    del canvas_host

    clocks = sampler.stop() if rank == 0 else None
    gpu_ref = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            gpu_ref = gpu_reference_arm(args, path, dpts[0], B, w, dev)
        except Exception as exc:  # the comparator must never take the bench line down
            gpu_ref = dict(unavailable=f"{type(exc).__name__}: {exc}"[:200])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frames = B * world
    value = frames * args.steps / (ms_dev * 1e-3)
    alg = alg_bytes_per_frame(n, F, 64, ny, nx, M / B) * B            # bytes one launch of the path moves, per GPU
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"
    dominant_kernel = ("k_emit (per canvas tile of 32 cells x 64 channels: order + decorate + PFN + max of the tile's pillars, pillar "
                       "rows, and the tile itself in one TMA tensor store, zero tiles included)")
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        ent = json.load(open(tp)).get(f"{args.config}_{args.mode}", {})
        traffic = ent.get("dominant_kernel_dram_bytes_per_launch")
        traffic_src = dict(kind="ncu capture committed under profiles/ (NOT measured in this run)", source=ent.get("source"),
                           commit=ent.get("commit"))
        try:
            # stale = the path's kernel sources changed since the capture (docs / scripts / other kernels moving HEAD do not count); no .git on
            # the GPU box: no check
            if ent.get("commit"):
                r = subprocess.run(["git", "diff", "--quiet", ent["commit"], "HEAD", "--"] +
                                   [f"hgsfusion_b200/csrc/{f}" for f in ("pillar_path.cu", "pillar_path.cuh", "common.cuh", "pfn.cuh")],
                                   cwd=ROOT, capture_output=True, text=True)
                if r.returncode == 1:
                    traffic_src["warning"] = f"kernel sources changed since the capture at {ent['commit']}"
        except OSError:
            pass
    step_ms = ms_dev / args.steps
    roof = dict(bound="hbm", kernel=dominant_kernel,
                achieved=(alg / (emit_ms * 1e-3) / 1e9) if emit_ms else None, peak=peak, unit="GB/s",
                frac=(alg / (emit_ms * 1e-3) / 1e9 / peak) if emit_ms else None, traffic=traffic, traffic_source=traffic_src,
                peak_source=peak_src, kernel_ms=emit_ms, kernel_launches_timed=n_ev,
                algorithmic_bytes_per_launch=alg, instrumented_ms_per_step=ms_inst,
                step_achieved=alg / (step_ms * 1e-3) / 1e9, step_frac=alg / (step_ms * 1e-3) / 1e9 / peak,
                note="achieved = SURVEY 8(d) algorithmic bytes of one batch / mean duration of the dominant kernel (CUDA event pairs on the "
                     "launch stream, in a second pass over the same K steps: instrumented_ms_per_step); step_* = the same bytes / the "
                     "whole-step time of the un-instrumented timed region (k_front + k_emit)")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=warm,
                ms_per_step=step_ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="ours",
                config=workload(args), pillars_per_frame=M / B,
                e2e=dict(value=frames * e2e_steps / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         steps=e2e_steps, h2d_gbs_per_rank=h2d_gbs_rank,
                         note="pinned host points [sum N, F] + frame_offsets -> H2D -> hgsf_points_to_bev -> D2H pillar counts; the canvas "
                              "stays on the device as spatial_features does in the reference",
                         copy_only=dict(h2d_gbs_per_rank=copy_gbs_rank, h2d_gbs_aggregate=copy_gbs_rank * world,
                                        frames_per_s_ceiling=frames * e2e_steps / (ms_copy * 1e-3),
                                        note="the same pinned buffers, streams and events without the kernels (slowest rank): the "
                                             "host-to-device ceiling of this box at this rank count"),
                         numa=numa,
                         collated_layout=dict(value=frames * e2e_steps / (ms_e2e_col * 1e-3), unit=UNIT,
                                              h2d_bytes_per_step=int(pinned_col[0].numel() * 4),
                                              note="points as the collated [sum N, 1+F] with the batch column"),
                         with_canvas_d2h=dict(value=frames * canvas_steps / (ms_e2e_canvas * 1e-3), unit=UNIT,
                                              d2h_bytes_per_step=int(res.spatial_features.numel() * 4) + d2h)),
                gpu_launches=launches_per_step * args.steps, gpu_launches_per_step=launches_per_step,
                clocks=clocks, roofline=roof)
    if gpu_ref is not None:
        if "value" in gpu_ref:
            gpu_ref["ours_over_reference"] = value / world / gpu_ref["value"]
        line["gpu_reference"] = gpu_ref
    if world == 1 and not args.no_cpu_baseline:
        base, _ = cpu_reference_arm(args, steps=12, warmup=2)
        line["cpu_baseline"] = base
    if world > 1:
        dist.destroy_process_group()
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    # Only the JSON line may reach stdout: libraries (NCCL's version banner, torchrun notices) write there too, so
    # file descriptor 1 is pointed at stderr for the whole run and the JSON goes to the saved descriptor.
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
