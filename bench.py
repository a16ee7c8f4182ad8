#!/usr/bin/env python
"""bench.py — frames/s and points/s of the 64x2048 cloud -> range -> ground -> cloud path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step is one pass of the hot path over one batch of F synthetic HDL-64E frames (64 x 2048, ~120 k points each;
BASELINE config C4: a pool of 64 distinct frames tiled to F = 4096) PER GPU: frames are independent, so ranks shard
frame-parallel with no data-path collective (weak scaling; torch.distributed is used for the barrier and the
max-over-ranks time only).

  value      device-resident throughput: inputs already in HBM, gs_process_batch_device on the context's stream,
             timed with CUDA events on that stream (max over ranks)
  e2e        the same metric through gs_process_batch_host with pinned HOST buffers (H2D of the points and D2H of the
             no-ground images + clouds inside the timed region)
  roofline   algorithmic bytes (16 N + 2 P + 16 N_ng per frame) / device time, against MEASURED_PEAKS.json
  cpu_baseline / --impl reference
             the reference's CPU path (the oracle port: the reference itself needs ROS/PCL/OpenCV and cannot be
             built) on this box's host cores, bounded sample of the same workload
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from groundsegmentationrosnode_b200 import params as P, synth  # noqa: E402

METRIC = "frames/s, 64x2048 cloud->range->ground->cloud (points/s alongside)"
WORKLOAD = "C4: synthetic HDL-64E 64x2048 frames (~120k pts), pool of 64 distinct frames tiled to the batch"
N_DISTINCT = 64
# dram__bytes_read.sum + dram__bytes_write.sum of one k_frame launch / frames in it (ncu --set full, profiles/r1_fused_ncu.md)
TRAFFIC_BYTES_PER_FRAME = 901476608 / 296  # ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch over 296 frames (profiles/r1_final_ncu_raw_selected.txt)


def rank_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def make_workload(n_frames, n_distinct=N_DISTINCT):
    p = P.preset(P.GS_PRESET_HDL64E_2048)
    n_distinct = min(n_distinct, n_frames)
    pool, offs = synth.make_pool(p, n_distinct)
    return p, pool, offs


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def cpu_reference_run(p, pool, offs, seconds_budget, threads):
    """Times the oracle port of the reference's CPU path (Alfa_GS::process_pointcloud hw==0) on a bounded sample."""
    from oracle import oracle as O  # the one place bench.py may execute oracle/ (cpu_baseline / --impl reference)

    n_distinct = len(offs) - 1
    # calibrate on one frame, then size the sample
    t0 = time.perf_counter()
    O.process_batch(p, pool[: offs[1]], offs[:2], n_threads=1, want_outputs=False)
    per_frame = max(time.perf_counter() - t0, 1e-4)
    n_frames = int(max(threads, min(4096, seconds_budget / per_frame * threads)))
    pts, offsets = synth.tile_pool(pool, offs, n_frames)
    out = O.process_batch(p, pts, offsets, n_threads=threads, want_outputs=False)
    n_points = int(offsets[-1])
    return {"frames": n_frames, "points": n_points, "seconds": out["seconds"], "fps": n_frames / out["seconds"],
            "pps": n_points / out["seconds"], "stage_ms_per_frame": (out["stage_ms"] / n_frames).tolist(), "distinct": n_distinct}


def run_reference(args):
    rank, _, world = rank_info()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    p, pool, offs = make_workload(64)
    per_step = max(2.0, 60.0 / max(args.steps + args.warmup, 1))
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_run(p, pool, offs, per_step, cores)
        if i >= args.warmup:
            vals.append(r)
    frames = sum(v["frames"] for v in vals)
    secs = sum(v["seconds"] for v in vals)
    pts = sum(v["points"] for v in vals)
    fps = frames / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "points_per_s": pts / secs,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(len(vals), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32/u16 (reference arithmetic)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "frames_per_step": vals[0]["frames"], "shape": "64x2048"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d frames per step on %d host threads (independent frame streams); oracle port of "
                                   "Alfa_GS::process_pointcloud hw==0 — the reference needs ROS/PCL/OpenCV and cannot be compiled" % (vals[0]["frames"], cores)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=4096, help="frames per GPU per step (BASELINE config C4: 4096)")
    ap.add_argument("--e2e-frames", type=int, default=1024, help="frames per GPU per end-to-end step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--staged", action="store_true", help="force the per-stage kernels instead of the fused kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from groundsegmentationrosnode_b200 import capi

    rank, local_rank, world = rank_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    p, pool, offs = make_workload(args.frames)
    H, W = p.shape
    Pn = H * W
    F = args.frames
    pts_np, offsets = synth.tile_pool(pool, offs, F)
    n_points = int(offsets[-1])

    ctx = capi.Context(p, device=local_rank)
    if args.staged:
        ctx.set_option(capi.GS_OPT_FORCE_STAGED, 1)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    d_pts = torch.from_numpy(pts_np).cuda()
    d_res = torch.empty((F, H, W), dtype=torch.int16, device="cuda")
    d_xyz = torch.empty((F, Pn, 4), dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    def step():
        ctx.process_batch_device(d_pts.data_ptr(), offsets, F, 0, 0, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr())

    for _ in range(args.warmup):
        step()
    ctx.synchronize()
    launches0 = ctx.stats().kernel_launches
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.finish()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.stats().kernel_launches - launches0
    n_ng = int(d_cnt.sum().item())

    # ---- end to end through the host-buffer C-ABI call (pinned buffers, copies inside the timed region)
    Fe = min(args.e2e_frames, F)
    e_offsets = offsets[: Fe + 1].copy()
    e_points = int(e_offsets[-1])
    h_pts = torch.from_numpy(pts_np[:e_points]).pin_memory()
    h_res = torch.empty((Fe, H, W), dtype=torch.int16).pin_memory()
    cap = int(d_cnt[:Fe].sum().item()) + 1024
    h_xyz = torch.empty((cap, 4), dtype=torch.float32).pin_memory()
    h_out_offs = np.zeros(Fe + 1, np.int64)

    def e2e_step():
        ctx.process_batch_host(h_pts.numpy(), e_offsets, h_res.numpy().view(np.uint16), h_xyz.numpy(), h_out_offs)

    for _ in range(2):
        e2e_step()
    barrier()
    torch.cuda.synchronize()
    e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e_sec = time.perf_counter() - t0
    e_out_records = int(h_out_offs[-1])

    # ---- max over ranks
    t = torch.tensor([ms_total, e_sec], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(F), float(n_points), float(n_ng), float(launches), float(Fe), float(e_points)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, e_sec = float(t[0]), float(t[1])
    F_all, pts_all, ng_all, launches_all, Fe_all, e_pts_all = (float(x) for x in tot)

    if rank == 0:
        sec = ms_total / 1e3
        fps = F_all * args.steps / sec
        pps = pts_all * args.steps / sec
        peak, peak_src = measured_peak()
        bytes_per_step_gpu = 16.0 * n_points + 2.0 * Pn * F + 16.0 * n_ng  # this rank's algorithmic bytes per step
        step_s = sec / args.steps
        achieved = bytes_per_step_gpu / step_s / 1e9
        launches_per_step = launches / max(args.steps, 1)
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "points_per_s": pps, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64/f32/u16 (reference arithmetic: double projection and back-projection, float angles, u16 images)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": F, "shape": "64x2048", "points_per_frame": n_points / F,
                       "l2": "inputs (%.1f GB per step per GPU) are larger than L2; no flush needed" % (16.0 * n_points / 1e9),
                       "path": "staged kernels" if args.staged else "default (fused kernel when the configuration fits, else staged)",
                       "sharding": "frame-parallel, no collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": TRAFFIC_BYTES_PER_FRAME * F if (TRAFFIC_BYTES_PER_FRAME and not args.staged) else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_frame": bytes_per_step_gpu / F,
                         "algorithmic_bytes_per_launch": bytes_per_step_gpu,
                         "kernel": ("k_frame (fused persistent cluster kernel): the whole step is one launch" if launches_per_step <= 1.5
                                    else "whole step (%g staged launches per step); per-kernel shares in profiles/" % launches_per_step),
                         "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per frame x frames per launch (profiles/)"},
            "e2e": {"value": Fe_all * e_steps / e_sec, "unit": "frames/s", "points_per_s": e_pts_all * e_steps / e_sec,
                    "h2d_bytes_per_step": int(16 * e_points), "d2h_bytes_per_step": int(2 * Pn * Fe + 16 * e_out_records + 4 * Fe),
                    "frames_per_step": Fe, "api": "gs_process_batch_host (pinned host buffers, chunked copy/compute overlap)"},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            one = cpu_reference_run(p, pool, offs, 6.0, 1)
            allc = cpu_reference_run(p, pool, offs, 12.0, cores) if cores > 1 else one
            line["cpu_baseline"] = {
                "value": allc["fps"], "unit": "frames/s", "cores": cores, "kind": "port",
                "sample": "%d frames on %d host threads (independent frame streams); 1-thread (the reference's execution model): "
                          "%.2f frames/s on %d frames" % (allc["frames"], cores, one["fps"], one["frames"]),
                "one_core_value": one["fps"], "points_per_s": allc["pps"], "stage_ms_per_frame_one_core": one["stage_ms_per_frame"],
            }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
