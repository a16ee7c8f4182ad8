#!/usr/bin/env python
"""Randomised parity run: the CUDA path (through the C ABI) against the CPU oracle on random sensor configurations and clouds
for a fixed time, beyond what the test-suite's seeds cover.

    gpurun -- 'python tools/fuzz_parity.py 240 > gpurun_out/fuzz.txt'          (seconds; default 120)

Every round draws a sensor shape (16 / 32 / 64 / 128 beams, odd and even widths), a vertical field of view that may reach past
+-45 degrees (the fast projection leaves steeper points to the exact path), range gates, thresholds, both chain variants,
RepairGaps and smoothing on / off, both pixel policies, keep-ground mode, and a cloud that is ordered / shuffled, dense / sparse,
jittered up to half a cell or adversarial (points on cell edges, collisions, non-finite values).  The frame runs through the
fused kernel's production and checked instantiations (the latter also counts fast-path / exact-path disagreements) and the
per-stage kernels; range image, angle image, no-ground image and cloud must equal the oracle's bit for bit.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from groundsegmentationrosnode_b200 import capi, params as P, synth  # noqa: E402
from oracle import oracle  # noqa: E402  (the checker, test infrastructure)


def main():
    dry = "--dry" in sys.argv  # generator + oracle only (no GPU): checks that every drawn configuration is a legal one
    argv = [a for a in sys.argv[1:] if a != "--dry"]
    budget = float(argv[0]) if len(argv) > 0 else 120.0
    seed0 = int(argv[1]) if len(argv) > 1 else 20261019
    rng = np.random.default_rng(seed0)
    t_end = time.time() + budget
    rounds = frames = points = 0
    bad = []
    counters = dict(fallbacks=0, mismatches=0, near_edge=0, steep=0, fused=0, staged_only=0)
    while time.time() < t_end:
        rounds += 1
        H = int(rng.choice([16, 32, 64, 128]))
        W = int(rng.choice([130, 517, 900, 1000, 1024, 1500, 1800, 2000, 2048, 2500, 3000]))
        if H == 128:
            W = min(W, 2048)
        steep = rng.random() < 0.25
        lo = -float(rng.uniform(0.2, 1.2 if steep else 0.5))
        hi = float(rng.uniform(0.05, 1.2 if steep else 0.4))
        variant = int(rng.integers(0, 2))
        p = P.GsParams(n_beams=H, rpm=600, sample_freq=W * 10, min_angle=lo, max_angle=hi, min_range=float(rng.uniform(0.1, 1.0)),
                       max_range=float(rng.uniform(60, 250)), ground_angle_threshold=float(rng.uniform(0.02, 0.3)),
                       start_angle_threshold=float(rng.uniform(0.1, 1.0)), window_size=int(rng.choice([5, 7, 9, 11])),
                       erase_ground=int(rng.integers(0, 2)), variant=variant, pixel_policy=int(rng.integers(0, 2)),
                       col_policy=int(rng.integers(0, 2)), smoothing=int(rng.integers(0, 2)),
                       repair_gaps=int(rng.integers(0, 2)) if variant == P.GS_VARIANT_CLOUD2RANGE else 0)
        kind = rng.random()
        seed = int(rng.integers(1, 1 << 30))
        if kind < 0.2:
            pts = synth.make_adversarial_frame(p, seed, kind=str(rng.choice(["edges", "collisions", "nonfinite", "sparse"])))
        else:
            pts = synth.make_frame(p, seed, dropout=float(rng.choice([0.0, 0.05, 0.3, 0.8])), shuffle=bool(rng.integers(0, 2)),
                                   jitter=float(rng.choice([0.0, 0.3, 0.45, 0.5])), noise_sigma=float(rng.choice([0.0, 0.02, 0.2])))
        if rng.random() < 0.15:  # scale some points far out / close in, move some onto the axes
            pts = pts.copy()
            idx = rng.integers(0, len(pts), size=max(1, len(pts) // 50))
            pts[idx, :3] *= rng.choice([1e-3, 0.5, 3.0, 50.0])
            idx = rng.integers(0, len(pts), size=max(1, len(pts) // 200))
            pts[idx, int(rng.integers(0, 3))] = 0.0
        want = oracle.process_frame(p, pts)
        frames += 1
        points += len(pts)
        counters["steep"] += int(steep)
        if dry:
            continue
        with capi.Context(p) as ctx:
            for mode in ("checked", "production", "staged"):
                ctx.set_option(capi.GS_OPT_FORCE_STAGED, 1 if mode == "staged" else 0)
                ctx.set_option(capi.GS_OPT_COLLECT_STATS, 1 if mode == "checked" else 0)
                ctx.set_option(capi.GS_OPT_NO_LONE_PLAN, 1 if mode == "checked" else 0)
                ctx.reset_stats()
                got = ctx.process_frame(pts)
                st = ctx.stats()
                if mode == "checked":
                    counters["fallbacks"] += st.slow_path_points
                    counters["mismatches"] += st.fast_path_mismatches
                    counters["near_edge"] += st.near_edge_project + st.near_edge_angle
                    if st.kernel_launches == 1:
                        counters["fused"] += 1
                    else:
                        counters["staged_only"] += 1
                diff = [k for k in ("range", "angle", "no_ground", "xyz") if not np.array_equal(got[k], want[k])]
                if diff or (mode == "checked" and st.fast_path_mismatches):
                    bad.append((rounds, mode, diff, int(st.fast_path_mismatches), H, W, lo, hi, variant, int(p.pixel_policy), int(p.repair_gaps), seed, kind))
    print("fuzz_parity: %d frames (%d points) in %.0f s, seed %d" % (frames, points, budget, seed0))
    print("  configurations on the fused kernel: %d, on the per-stage kernels only: %d, with a field of view past 45 degrees: %d"
          % (counters["fused"], counters["staged_only"], counters["steep"]))
    print("  checked instantiation: %d points took the exact path, %d points / pixels within 1e-5 of a cell edge, fast-path / exact-path disagreements: %d"
          % (counters["fallbacks"], counters["near_edge"], counters["mismatches"]))
    print("  frames differing from the oracle (any of range / angle / no-ground image, cloud; production, checked and per-stage runs): %d" % len(bad))
    for b in bad[:20]:
        print("   ", b)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
