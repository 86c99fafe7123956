#!/usr/bin/env python
"""bench.py -- points -> BEV frames/s of the radar pillarization hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 200 --warmup 10          # our arm, one B200
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 5 --warmup 1     # the CPU path of the reference (oracle port) on host cores

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
    cfg = synthetic.CONFIGS[args.config]
    return dict(workload=f"{args.config}_{args.mode}_b{args.batch}_n{args.points}", dataset_shape=args.config,
                points_per_frame=args.points, point_features=cfg["F"], frames_per_gpu_per_step=args.batch,
                max_points_per_voxel=args.max_points, max_voxels=args.max_voxels, channels=64,
                pc_range=cfg["pc_range"], voxel_size=cfg["voxel_size"], point_distribution=args.mode)


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
def cpu_reference_arm(args, steps, warmup, threads=None):
    """The reference's CPU path for this workload: per-frame spconv-style voxelization (one frame per thread, as the
    DataLoader workers do), PillarVFE and PointPillarScatter -- the oracle port in C, all host threads."""
    from oracle import oracle
    cfg = synthetic.CONFIGS[args.config]
    F = cfg["F"]
    geom = oracle.Geometry(cfg["pc_range"], cfg["voxel_size"])
    threads = threads or (os.cpu_count() or 1)
    oracle.set_num_threads(threads)
    w = synthetic.make_pfn(F + 6, 64, 0)
    pfn = oracle.PfnParams(w.weight, w.gamma, w.beta, w.running_mean, w.running_var)
    pts, offs = synthetic.make_batch(args.config, args.batch, args.points, args.mode, seed0=0)
    run = lambda: oracle.points_to_bev(pts, offs, geom, pfn, args.max_points, args.max_voxels, F=F, xcol=1,
                                       want_voxels=False)
    for _ in range(max(warmup, 1)):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    oracle.set_num_threads(1)
    return dict(value=args.batch * steps / dt, unit=UNIT, cores=threads, kind="port",
                sample=f"{steps} steps x {args.batch} frames of the bench workload ({dt:.1f} s); oracle/pillar_oracle.c, "
                       f"{threads} pthreads (frames in parallel for voxelization, pillars in parallel for the PFN)"), dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: cap the step count so that the run ends within a few minutes on any host
    steps, warmup = max(1, min(args.steps, 20)), max(1, min(args.warmup, 2))
    base, dt = cpu_reference_arm(args, steps, warmup)
    line = dict(metric=METRIC, value=base["value"], unit=UNIT, impl="reference", n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=1e3 * dt / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=workload(args), cpu_baseline=base,
                e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                note="reference arm = the reference's CPU implementation of the path (oracle port; the reference itself is "
                     "Python/torch + spconv and cannot travel to the GPU box), host cores only")
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
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
    pinned = [torch.from_numpy(h).pin_memory() for h in host]
    dpts = [p.to(dev) for p in pinned]
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

    # ---- device-resident throughput: W warm-up, K timed steps, CUDA events, max over ranks ----
    for i in range(max(args.warmup, 3)):
        path.points_to_bev(dpts[i % ring], B, pfn, out=res)
    lib.hgsf_emit_timing_begin(min(args.steps, 1024))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        path.points_to_bev(dpts[i % ring], B, pfn, out=res)
    e1.record()
    barrier()
    ms_dev = sharding.reduce_max(e0.elapsed_time(e1), dev)
    buf = (C.c_float * 1024)()
    n_ev = lib.hgsf_emit_timing_collect(buf, 1024)
    emit_ms = statistics.fmean(buf[:n_ev]) if n_ev > 0 else None
    lib.hgsf_emit_timing_begin(0)

    # ---- end to end through the public API with HOST buffers: pinned H2D of the step's points, the native call,
    #      D2H of the step's result summary (pillar counts); the canvas stays on the device for the 2D backbone,
    #      as batch_dict['spatial_features'] does in the reference.  Copies ride a second stream, double buffered. ----
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty_like(dpts[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    counts_host = [torch.empty(1 + B, dtype=torch.int32).pin_memory() for _ in range(2)]
    main_stream = torch.cuda.current_stream()

    def e2e_loop(steps, with_canvas=False, canvas_host=None, srcs=None, stages=None, call=None):
        srcs = pinned if srcs is None else srcs
        stages = stage if stages is None else stages
        call = (lambda pts_dev: path.points_to_bev(pts_dev, B, pfn, out=res)) if call is None else call
        for s in consumed:
            s.record(main_stream)
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
            call(stages[cur])
            consumed[cur].record(main_stream)
            counts_host[cur].copy_(res.num_pillars, non_blocking=True)
            if with_canvas:
                canvas_host.copy_(res.spatial_features, non_blocking=True)

    e2e_steps = max(10, min(args.steps, 200))
    e2e_loop(3)
    barrier()
    e0.record()
    e2e_loop(e2e_steps)
    e1.record()
    barrier()
    ms_e2e = sharding.reduce_max(e0.elapsed_time(e1), dev)
    assert int(counts_host[(e2e_steps - 1) & 1][0]) > 0
    h2d = int(pinned[0].numel() * 4)
    d2h = int((1 + B) * 4)
    # the same call fed with the frames' own rows [sum N, F] plus frame offsets (hgsf_points.frame_offsets) instead of the
    # collated [sum N, 1+F] layout: what a collate hook that does not add the batch column would ship over PCIe
    pinned_nf = [torch.from_numpy(np.ascontiguousarray(h[:, 1:])).pin_memory() for h in host]
    stage_nf = [torch.empty((pinned_nf[0].shape[0], F), dtype=torch.float32, device=dev) for _ in range(2)]
    offs_dev = torch.arange(0, (B + 1) * n, n, dtype=torch.int32, device=dev)
    call_nf = lambda pts_dev: path.points_to_bev(pts_dev, B, pfn, xyz_col=0, frame_offsets=offs_dev, out=res)
    e2e_loop(3, srcs=pinned_nf, stages=stage_nf, call=call_nf)
    barrier()
    e0.record()
    e2e_loop(e2e_steps, srcs=pinned_nf, stages=stage_nf, call=call_nf)
    e1.record()
    barrier()
    ms_e2e_nf = sharding.reduce_max(e0.elapsed_time(e1), dev)
    assert int(counts_host[(e2e_steps - 1) & 1][0]) > 0
    # the same with the whole canvas also copied to the host every step (PCIe bound; for the record only)
    canvas_steps = 5
    canvas_host = torch.empty(res.spatial_features.shape, dtype=torch.float32).pin_memory()
    e2e_loop(1, True, canvas_host)
    barrier()
    e0.record()
    e2e_loop(canvas_steps, True, canvas_host)
    e1.record()
    barrier()
    ms_e2e_canvas = sharding.reduce_max(e0.elapsed_time(e1), dev)

    clocks = sampler.stop() if rank == 0 else None
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
    dominant_kernel = "k_emit (fused: order + decorate + PFN + max + pillar rows + canvas tiles via TMA, zeros included)"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{args.config}_{args.mode}", {}).get("dominant_kernel_dram_bytes_per_launch")
    roof = dict(bound="hbm", kernel=dominant_kernel,
                achieved=(alg / (emit_ms * 1e-3) / 1e9) if emit_ms else None, peak=peak, unit="GB/s",
                frac=(alg / (emit_ms * 1e-3) / 1e9 / peak) if emit_ms else None, traffic=traffic,
                peak_source=peak_src, kernel_ms=emit_ms, kernel_launches_timed=n_ev,
                algorithmic_bytes_per_launch=alg,
                step_achieved=alg / (ms_dev / args.steps * 1e-3) / 1e9, step_frac=alg / (ms_dev / args.steps * 1e-3) / 1e9 / peak,
                note="achieved = SURVEY 8(d) algorithmic bytes of one batch / mean duration of the dominant kernel (CUDA events on the launch "
                     "stream inside the timed region); step_* = the same bytes / whole-step time (k_front + k_emit); traffic = ncu dram read+write of that kernel per launch")
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="ours",
                config=dict(workload(args), pillars_per_frame=M / B, grid=[nx, ny, 1],
                            l2=f"ring of {ring} distinct input batches ({ring * h2d / 1e6:.0f} MB) and a "
                               f"{res.spatial_features.numel() * 4 / 1e6:.0f} MB canvas rewritten every step, both > 126 MB L2",
                            parallelism=f"frames sharded by rank, {B} per GPU, no collective"),
                e2e=dict(value=frames * e2e_steps / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         steps=e2e_steps,
                         note="pinned host points -> H2D -> hgsf_points_to_bev -> D2H pillar counts; canvas stays on device as "
                              "spatial_features does in the reference",
                         per_frame_rows_layout=dict(value=frames * e2e_steps / (ms_e2e_nf * 1e-3), unit=UNIT,
                                                    h2d_bytes_per_step=int(pinned_nf[0].numel() * 4),
                                                    note="points as [sum N, F] + frame_offsets (no batch column)"),
                         with_canvas_d2h=dict(value=frames * canvas_steps / (ms_e2e_canvas * 1e-3), unit=UNIT,
                                              d2h_bytes_per_step=int(res.spatial_features.numel() * 4) + d2h)),
                gpu_launches=launches_per_step * args.steps, gpu_launches_per_step=launches_per_step,
                clocks=clocks, roofline=roof)
    if world == 1 and not args.no_cpu_baseline:
        base, _ = cpu_reference_arm(args, steps=8, warmup=1)
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
