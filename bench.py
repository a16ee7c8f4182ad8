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
  parity     K frames of the timed run's outputs compared with the CPU oracle after the timed region
  cpu_baseline / --impl reference
             the reference's OWN sources (src/alfa_gs.cpp + src/alfa_node.cpp + Labeler/label.cpp compiled verbatim into
             oracle/_ref against stand-in ROS / PCL / OpenCV headers) driven frame by frame the way ROS drives
             the node, one node per host core, on a bounded sample of the same workload
  --latency  one frame at a time through gs_process_frame (the reference's real execution model, alfa_node.cpp:298-319)
  --config   C4 (default, the headline: 64x2048 batch) | C3 (128x2048 with the Labeler CCL labels as an output)
  --native   in-process multi-GPU mode: the native driver (gs_multi) over --gpus N devices, end to end from host buffers
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
N_DISTINCT = 64
CONFIGS = {
    # name: (preset, workload string, frames per GPU per step, labels as an output)
    "C4": (P.GS_PRESET_HDL64E_2048, "C4: synthetic HDL-64E 64x2048 frames (~120k pts), pool of 64 distinct frames tiled to the batch", 4096, False),
    "C3": (P.GS_PRESET_OS1_128, "C3: synthetic OS1-128 128x2048 frames (~260k pts), full pipeline including the Labeler CCL (component labels out), "
                                "pool of 64 distinct frames tiled to the batch", 1024, True),
}
LABEL_BETA_DEG = 10.0  # the Labeler's threshold argument when it runs as an all-components CCL (BASELINE config C3 / C5)


def config_dict(name, frames):
    """The `config` object of the JSON line: identical in the GPU arm and the reference arm."""
    preset, workload, _, labels = CONFIGS[name]
    H, W = P.preset(preset).shape
    return {"workload": workload, "shape": "%dx%d" % (H, W), "frames_per_gpu_per_step": frames, "labels": bool(labels),
            "sharding": "frame-parallel, no collective"}


def ncu_traffic(shape):
    """dram__bytes_read.sum + dram__bytes_write.sum per frame from the committed ncu --set full capture (profiles/r2_ncu_traffic.json),
    or None when there is no capture for this shape.  A static figure: the bench does not run under a profiler."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")) as f:
            t = json.load(f)
        return t.get(shape)
    except Exception:
        return None


def bind_near_gpu(index):
    """Pin this process to the CPUs next to its GPU (NUMA node of the PCI device) before any pinned buffer is allocated: the
    host copies of 8 ranks otherwise share one node's memory.  Returns a note for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        node = int(open(dev + "/numa_node").read())
        cpus = open(dev + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"numa_node": node, "cpus": cpus, "bound": bool(ids)}
    except Exception as e:  # no NVML / sysfs in this container: leave the affinity alone
        return {"numa_node": None, "cpus": None, "bound": False, "why": str(e)[:80]}


def rank_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def make_workload(n_frames, n_distinct=N_DISTINCT, preset=P.GS_PRESET_HDL64E_2048):
    p = P.preset(preset)
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


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs: the reference itself (oracle/_ref, compiled verbatim) and the oracle port — the one place bench.py executes oracle/
# ---------------------------------------------------------------------------------------------------------------------
_W = {}


def _ref_worker(args):
    """One reference node (Alfa_GS + AlfaNode + Labeler, verbatim) on this process: frames through AlfaNode::cloud_cb."""
    wid, n_workers, n_frames, seconds = args
    from oracle import oracle as O

    p, pool, offs = _W["p"], _W["pool"], _W["offs"]
    node = O.RefAlfaGS(p)
    nd = len(offs) - 1
    frames = [np.ascontiguousarray(pool[offs[i]:offs[i + 1]]) for i in range(nd)]
    node.process_frame(frames[wid % nd], want_outputs=False)  # first-touch of the node's buffers
    done, pts = 0, 0
    t0 = time.perf_counter()
    k = wid
    while done < n_frames:
        f = frames[k % nd]
        if node.process_frame(f, want_outputs=False) < 0:
            raise RuntimeError("the reference node published nothing")
        done += 1
        pts += f.shape[0]
        k += n_workers
        if seconds and time.perf_counter() - t0 > seconds:
            break
    return done, pts, time.perf_counter() - t0


def reference_run(p, pool, offs, seconds, workers):
    """The reference's own code on `workers` host processes for about `seconds` of wall clock (independent frame streams:
    the reference itself is one ros::spin thread per node, alfa_node.cpp:387)."""
    import multiprocessing as mp

    _W.update(p=p, pool=pool, offs=offs)
    t0 = time.perf_counter()
    if workers == 1:
        res = [_ref_worker((0, 1, 1 << 30, seconds))]
    else:
        with mp.get_context("fork").Pool(workers) as pl:
            res = pl.map(_ref_worker, [(w, workers, 1 << 30, seconds) for w in range(workers)])
    wall = time.perf_counter() - t0
    frames = sum(r[0] for r in res)
    pts = sum(r[1] for r in res)
    busy = max(r[2] for r in res)  # the workers run side by side: the slowest one bounds the sample
    return {"frames": frames, "points": pts, "seconds": busy, "fps": frames / busy, "pps": pts / busy, "wall": wall}


def port_run(p, pool, offs, seconds_budget, threads):
    """The oracle port (oracle/gs_oracle.cpp: restated, tighter loops, no discarded moving average) on a bounded sample."""
    from oracle import oracle as O

    t0 = time.perf_counter()
    O.process_batch(p, pool[: offs[1]], offs[:2], n_threads=1, want_outputs=False)
    per_frame = max(time.perf_counter() - t0, 1e-4)
    n_frames = int(max(threads, min(4096, seconds_budget / per_frame * threads)))
    pts, offsets = synth.tile_pool(pool, offs, n_frames)
    out = O.process_batch(p, pts, offsets, n_threads=threads, want_outputs=False)
    return {"frames": n_frames, "points": int(offsets[-1]), "seconds": out["seconds"], "fps": n_frames / out["seconds"],
            "pps": int(offsets[-1]) / out["seconds"], "stage_ms_per_frame": (out["stage_ms"] / n_frames).tolist()}


def have_reference():
    try:
        from oracle import oracle as O

        return O.ref() is not None
    except Exception:
        return False


def cpu_baseline_block(p, pool, offs, cores, all_seconds=8.0, one_seconds=3.0):
    block = {}
    port_one = port_run(p, pool, offs, 2.0, 1)
    port_all = port_run(p, pool, offs, 4.0, cores) if cores > 1 else port_one
    if have_reference():
        one = reference_run(p, pool, offs, one_seconds, 1)
        allc = reference_run(p, pool, offs, all_seconds, cores) if cores > 1 else one
        block.update({
            "value": allc["fps"], "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": "%d frames in %.1f s on %d host processes, one reference node each (src/alfa_gs.cpp + src/alfa_node.cpp + "
                      "Labeler/label.cpp compiled verbatim against stand-in ROS/PCL/OpenCV headers, PointCloud2 -> cloud_cb -> "
                      "process_pointcloud -> publish); one process: %.2f frames/s on %d frames"
                      % (allc["frames"], allc["seconds"], cores, one["fps"], one["frames"]),
            "one_core_value": one["fps"], "points_per_s": allc["pps"],
        })
    else:
        block.update({"value": port_all["fps"], "unit": "frames/s", "cores": cores, "kind": "port",
                      "sample": "%d frames on %d host threads: oracle port (oracle/_ref did not travel)" % (port_all["frames"], cores),
                      "one_core_value": port_one["fps"], "points_per_s": port_all["pps"]})
    # the restated oracle beside it: the same algorithm in tight loops on raw arrays, without the moving average the reference
    # computes and discards — what a tuned CPU build of the reference could reach
    block["port_value"] = port_all["fps"]
    block["port_one_core_value"] = port_one["fps"]
    block["port_stage_ms_per_frame_one_core"] = port_one["stage_ms_per_frame"]
    return block


def run_reference(args):
    rank, _, world = rank_info()
    if rank != 0:
        return 0
    name = args.config
    preset, _, frames_default, _ = CONFIGS[name]
    F = args.frames or frames_default
    cores = len(os.sched_getaffinity(0)) or 1
    p, pool, offs = make_workload(64, preset=preset)
    per_step = max(2.0, min(20.0, 90.0 / max(args.steps + args.warmup, 1)))
    ref = have_reference()
    vals = []
    for i in range(args.warmup + args.steps):
        r = reference_run(p, pool, offs, per_step, cores) if ref else port_run(p, pool, offs, per_step, cores)
        if i >= args.warmup:
            vals.append(r)
    frames = sum(v["frames"] for v in vals)
    secs = sum(v["seconds"] for v in vals)
    pts = sum(v["points"] for v in vals)
    fps = frames / secs
    kind = "reference" if ref else "port"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "points_per_s": pts / secs,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(len(vals), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32/u16 (reference arithmetic)",
        "data": "synthetic", "config": config_dict(name, F),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": "each step: about %.0f s of %d host processes, one reference node each, frames of the same pool (%d frames per step "
                                   "on average); %s" % (per_step, cores, frames // max(len(vals), 1),
                                                        "the reference's own sources compiled verbatim (oracle/_ref)" if ref else "oracle port (oracle/_ref did not travel)")},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def parity_block(p, pool, offs, offsets, d_res, d_xyz, d_cnt, d_lab, k_frames, beta):
    """K frames of the timed run's outputs against the CPU oracle (parity gates travel with every number)."""
    import platform

    from oracle import oracle as O

    nd = len(offs) - 1
    F = len(offsets) - 1
    picks = sorted(set(int(x) for x in np.linspace(0, F - 1, k_frames)))
    bad_px = bad_xyz = bad_cnt = bad_lab = near = 0
    for f in picks:
        src = pool[offs[f % nd]:offs[f % nd + 1]]
        want = O.process_frame(p, src)
        _, ne, _ = O.create_range_image(p, src, with_counts=True)
        near += int(ne)
        res = d_res[f].cpu().numpy().view(np.uint16)
        bad_px += int(np.count_nonzero(res != want["no_ground"]))
        n = int(d_cnt[f].item())
        bad_cnt += int(n != want["xyz"].shape[0])
        if n == want["xyz"].shape[0]:
            bad_xyz += int(np.count_nonzero(np.any(d_xyz[f, :n].cpu().numpy() != want["xyz"], axis=1)))
        if d_lab is not None:
            wl, _ = O.label_components(p, want["no_ground"], want["angle"], beta)
            bad_lab += int(np.count_nonzero(d_lab[f].cpu().numpy() != wl))
    out = {"frames_checked": len(picks), "mismatches": bad_px + bad_xyz + bad_cnt + bad_lab, "no_ground_pixels_differing": bad_px,
           "cloud_points_differing": bad_xyz, "cloud_sizes_differing": bad_cnt, "near_edge_points_in_checked_frames": near,
           "bar": "no-ground image and cloud bit-exact (north star: masks bit-exact except counted near-edge pixels, XYZ within 1e-6 relative)",
           "checker": "CPU oracle (oracle/gs_oracle.cpp, pinned to the reference's own sources by tests/test_reference_verbatim.py)",
           "libm": "%s %s" % platform.libc_ver()}
    if d_lab is not None:
        out["component_labels_differing"] = bad_lab
    return out


def pcie_probe(torch, barrier, in_per_frame, out_per_frame, n_bytes=256 << 20):
    """Copy-only ceiling of the e2e figure: this rank moves pinned host -> device and device -> pinned host bytes in the
    proportion of the workload (one frame = in_per_frame bytes in, out_per_frame bytes out), all ranks at the same time.
    Returns (frames/s this rank could move, H2D GB/s, D2H GB/s) plus the H2D-alone rate."""
    n_out = max(1 << 20, int(n_bytes * out_per_frame / max(in_per_frame, 1.0)))
    h_in = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    reps = 12
    for both in (False, True):
        torch.cuda.synchronize()
        barrier()  # buffers exist on every rank: the copies of all ranks overlap in time
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
            if both:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if both:
            out["h2d_gbs"] = reps * n_bytes / dt / 1e9
            out["d2h_gbs"] = reps * n_out / dt / 1e9
            out["frames_per_s"] = reps * n_bytes / in_per_frame / dt
        else:
            out["h2d_alone_gbs"] = reps * n_bytes / dt / 1e9
    return out


def run_latency(args):
    """One frame per call, the way the reference works (subscriber queue 1, one callback per cloud): pinned host buffers in,
    no-ground image + cloud out, through gs_process_frame."""
    import torch

    from groundsegmentationrosnode_b200 import capi

    _, local_rank, _ = rank_info()
    torch.cuda.set_device(local_rank)
    rows = {}
    for name, preset in (("C1 16x1800", P.GS_PRESET_VLP16), ("C2 64x2048", P.GS_PRESET_HDL64E_2048), ("C3 128x2048", P.GS_PRESET_OS1_128)):
        p = P.preset(preset)
        H, W = p.shape
        pool, offs = synth.make_pool(p, 8)
        frames = [torch.from_numpy(pool[offs[i]:offs[i + 1]].copy()).pin_memory() for i in range(8)]
        h_res = torch.empty((H, W), dtype=torch.int16).pin_memory()
        h_xyz = torch.empty((H * W, 4), dtype=torch.float32).pin_memory()
        p_res = np.empty((H, W), np.uint16)  # pageable outputs: the copying path of gs_process_frame
        p_xyz = np.empty((H * W, 4), np.float32)
        row = {}
        with capi.Context(p, device=local_rank) as ctx:
            def pinned(i):
                return ctx.process_frame_into(frames[i % 8].numpy(), h_res.numpy().view(np.uint16), h_xyz.numpy())

            def copying(i):
                return ctx.process_frame_into(frames[i % 8].numpy(), p_res, p_xyz)

            # four variants measured in alternating blocks (the GPU's clocks ramp with load: variants measured one after the other
            # would not see the same clocks); the first block of every variant is warm-up
            variants = [("p50_us", pinned, 0), ("p50_us_outputs_copied_back", copying, 0),
                        ("p50_us_batch_plan_cluster", pinned, 1), ("p50_us_batch_plan_cluster_outputs_copied_back", copying, 1)]
            samples = {k: [] for k, _, _ in variants}
            block = max(25, args.latency_frames // 6)
            for rnd in range(7):
                for key, fn, no_lone in variants:
                    ctx.set_option(capi.GS_OPT_NO_LONE_PLAN, no_lone)
                    for i in range(block):
                        t0 = time.perf_counter()
                        fn(i)
                        dt = time.perf_counter() - t0
                        if rnd:
                            samples[key].append(dt * 1e6)
            ts = np.array(samples["p50_us"])
            row.update({"p50_us": float(np.percentile(ts, 50)), "p99_us": float(np.percentile(ts, 99)), "mean_us": float(ts.mean()),
                        "frames": len(ts), "points_per_frame": float(np.mean([f.shape[0] for f in frames])),
                        "path": "page-locked output buffers: the kernel stores image, cloud and count straight into host memory; "
                                "one frame spread over the widest cluster that fits"})
            # where the latency comes from: the same call with the outputs copied back after the kernel, and on the batch plan's cluster
            for key, _, _ in variants[1:]:
                row[key] = float(np.percentile(np.array(samples[key]), 50))
        if not args.no_cpu_baseline:
            cpu = reference_run(p, pool, offs, 2.0, 1) if have_reference() else port_run(p, pool, offs, 2.0, 1)
            row["cpu_one_thread_us"] = 1e6 / cpu["fps"]
            row["cpu_kind"] = "reference" if have_reference() else "port"
        rows[name] = row
    print(json.dumps({"metric": "single-frame latency, pinned host cloud in -> no-ground image + cloud out (gs_process_frame)", "unit": "us",
                      "higher_is_better": False, "latency": rows, "data": "synthetic"}), flush=True)
    return 0


def run_native(args):
    """In-process multi-GPU mode: the native driver (gs_multi: one host thread + context per GPU, NUMA-local pinned buffers) over
    --gpus N devices of this box, end to end from host buffers.  No torchrun, no torch.distributed."""
    import torch

    from groundsegmentationrosnode_b200 import capi

    G = min(args.gpus, torch.cuda.device_count())
    preset, _, frames_default, _ = CONFIGS[args.config]
    Fg = min(args.e2e_frames, args.frames or frames_default)
    F = Fg * G
    p, pool, offs = make_workload(64, preset=preset)
    H, W = p.shape
    pts, offsets = synth.tile_pool(pool, offs, F)
    with capi.MultiGpu(p, list(range(G))) as mg:
        h_pts = mg.host_array(pts.shape, np.float32)
        h_pts[:] = pts
        h_res = mg.host_array((F, H, W), np.uint16)
        h_xyz = mg.host_array((F * (H * W // 2 + 4096), 4), np.float32)
        for _ in range(max(1, args.warmup // 2)):
            _, _, first, count = mg.process_batch_host(h_pts, offsets, h_res, h_xyz)
        steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(steps):
            _, _, first, count = mg.process_batch_host(h_pts, offsets, h_res, h_xyz)
        dt = (time.perf_counter() - t0) / steps
        workers = mg.workers()
    print(json.dumps({
        "metric": METRIC, "mode": "native multi-GPU driver, one process (gs_multi_process_batch_host)", "n_gpus": G, "unit": "frames/s",
        "higher_is_better": True, "scaling": "weak", "data": "synthetic", "config": config_dict(args.config, Fg),
        "e2e": {"value": F / dt, "unit": "frames/s", "h2d_bytes_per_step": int(pts.nbytes), "d2h_bytes_per_step": int(h_res.nbytes + int(count.sum()) * 16),
                "frames_per_step": F, "ms_per_step": 1e3 * dt},
        "workers": workers}), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C4", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default: the configuration's, C4: 4096)")
    ap.add_argument("--e2e-frames", type=int, default=1024, help="frames per call of the end-to-end leg (a step is frames / this many calls)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--latency", action="store_true", help="single-frame latency of gs_process_frame instead of batch throughput")
    ap.add_argument("--latency-frames", type=int, default=300)
    ap.add_argument("--staged", action="store_true", help="force the per-stage kernels instead of the fused kernel")
    ap.add_argument("--native", action="store_true", help="in-process multi-GPU mode: the native driver over --gpus N devices, end to end (no torchrun)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.latency:
        return run_latency(args)
    if args.native:
        return run_native(args)

    rank, local_rank, world = rank_info()
    affinity = bind_near_gpu(local_rank)  # before torch allocates pinned memory

    import torch
    import torch.distributed as dist

    from groundsegmentationrosnode_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    preset, _, frames_default, want_labels = CONFIGS[args.config]
    F = args.frames or frames_default
    p, pool, offs = make_workload(F, preset=preset)
    H, W = p.shape
    Pn = H * W
    pts_np, offsets = synth.tile_pool(pool, offs, F)
    n_points = int(offsets[-1])
    beta = float(np.radians(LABEL_BETA_DEG))

    ctx = capi.Context(p, device=local_rank)
    if args.staged:
        ctx.set_option(capi.GS_OPT_FORCE_STAGED, 1)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    d_pts = torch.from_numpy(pts_np).cuda()
    d_res = torch.empty((F, H, W), dtype=torch.int16, device="cuda")
    d_xyz = torch.empty((F, Pn, 4), dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
    d_lab = torch.empty((F, H, W), dtype=torch.int32, device="cuda") if want_labels else None
    torch.cuda.synchronize()

    def step():
        if want_labels:
            ctx.process_batch_device_labels(d_pts.data_ptr(), offsets, F, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr(), d_lab.data_ptr(), beta)
        else:
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

    # ---- end to end through the host-buffer C-ABI call (pinned buffers, copies inside the timed region): a step covers the same F frames,
    # in calls of --e2e-frames frames each (the pinned staging stays at one call's size)
    Fe = min(args.e2e_frames, F)
    calls = max(1, F // Fe)
    e_offsets = offsets[: Fe + 1].copy()
    e_points = int(e_offsets[-1])
    h_pts = torch.from_numpy(pts_np[:e_points]).pin_memory()
    h_res = torch.empty((Fe, H, W), dtype=torch.int16).pin_memory()
    cap = int(d_cnt[:Fe].sum().item()) + 1024
    h_xyz = torch.empty((cap, 4), dtype=torch.float32).pin_memory()
    h_out_offs = np.zeros(Fe + 1, np.int64)

    def e2e_step():
        for _ in range(calls):
            ctx.process_batch_host(h_pts.numpy(), e_offsets, h_res.numpy().view(np.uint16), h_xyz.numpy(), h_out_offs)

    ctx.process_batch_host(h_pts.numpy(), e_offsets, h_res.numpy().view(np.uint16), h_xyz.numpy(), h_out_offs)
    barrier()
    torch.cuda.synchronize()
    e_steps = max(2, min(args.steps, 4))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e_sec = time.perf_counter() - t0
    e_out_records = int(h_out_offs[-1])
    barrier()
    # every rank at the same time: the copy-only ceiling the ranks share
    pcie = pcie_probe(torch, barrier, 16.0 * e_points / Fe, (2.0 * Pn * Fe + 16.0 * e_out_records + 4.0 * Fe) / Fe)
    barrier()

    # ---- max over ranks
    t = torch.tensor([ms_total, e_sec], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(F), float(n_points), float(n_ng), float(launches), float(Fe * calls), float(e_points * calls)], dtype=torch.float64, device="cuda")
    pc = torch.tensor([pcie["h2d_alone_gbs"], pcie["h2d_gbs"], pcie["d2h_gbs"], pcie["frames_per_s"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(pc, op=dist.ReduceOp.SUM)
    ms_total, e_sec = float(t[0]), float(t[1])
    F_all, pts_all, ng_all, launches_all, Fe_all, e_pts_all = (float(x) for x in tot)

    if rank == 0:
        sec = ms_total / 1e3
        fps = F_all * args.steps / sec
        pps = pts_all * args.steps / sec
        peak, peak_src = measured_peak()
        label_bytes = 4.0 * Pn * F if want_labels else 0.0
        bytes_per_step_gpu = 16.0 * n_points + 2.0 * Pn * F + 16.0 * n_ng + label_bytes  # this rank's algorithmic bytes per step
        step_s = sec / args.steps
        achieved = bytes_per_step_gpu / step_s / 1e9
        launches_per_step = launches / max(args.steps, 1)
        traffic = None if args.staged else ncu_traffic("%dx%d" % (H, W))
        e_bytes_in = 16.0 * e_points * calls
        e_bytes_out = (2.0 * Pn * Fe + 16.0 * e_out_records + 4.0 * Fe) * calls
        e_fps = Fe_all * e_steps / e_sec
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "points_per_s": pps, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64/f32/u16 (reference arithmetic: double projection and back-projection, float angles, u16 images)",
            "data": "synthetic",
            "config": config_dict(args.config, F),
            "run": {"points_per_frame": n_points / F,
                    "l2": "inputs (%.1f GB per step per GPU) are larger than L2; no flush needed" % (16.0 * n_points / 1e9),
                    "path": "staged kernels" if args.staged else "fused kernel (thread-block clusters + software clusters on the SMs they strand)"
                            + (", then the batched CCL kernels for the labels" if want_labels else ""),
                    "host_affinity": affinity},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["bytes_per_frame"] * F if traffic else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_frame": bytes_per_step_gpu / F,
                         "algorithmic_bytes_per_launch": bytes_per_step_gpu,
                         "kernel": ("k_frame (fused persistent kernel: %g launches per step — the cluster instantiation and its software-cluster twin share one frame queue)" % launches_per_step
                                    if launches_per_step <= 2.5 and not want_labels
                                    else "whole step (%g launches per step); per-kernel shares in profiles/" % launches_per_step),
                         "traffic_source": (traffic or {}).get("source", "no ncu capture for this shape")},
            "e2e": {"value": e_fps, "unit": "frames/s", "points_per_s": e_pts_all * e_steps / e_sec,
                    "h2d_bytes_per_step": int(e_bytes_in), "d2h_bytes_per_step": int(e_bytes_out),
                    "frames_per_step": Fe * calls, "calls_per_step": calls,
                    "api": "gs_process_batch_host (pinned host buffers, waves of frames, copy-in / compute / copy-out on three streams)",
                    "pcie": {"ceiling_frames_per_s": float(pc[3]), "frac_of_ceiling": e_fps / max(float(pc[3]), 1e-9),
                             "h2d_gbs_all_ranks": float(pc[1]), "d2h_gbs_all_ranks": float(pc[2]), "h2d_alone_gbs_all_ranks": float(pc[0]),
                             "note": "copy-only ceiling measured in this run: every rank moves pinned host -> device and device -> pinned host bytes in the "
                                     "proportion of one frame (input records in, image + cloud out), all ranks at once, no kernel; "
                                     "profiles/r2_e2e_probe_8gpu.txt has the same at 1 / 2 / 4 / 8 GPUs from one process"}},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
        }
        if not args.no_parity:
            line["parity"] = parity_block(p, pool, offs, offsets, d_res, d_xyz, d_cnt, d_lab, 8, beta)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_block(p, pool, offs, len(os.sched_getaffinity(0)) or 1)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
