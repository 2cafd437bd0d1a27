#!/usr/bin/env python
"""bench.py -- profiled trajectories/s of the batched spline -> motion-profile engine on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--paths 4096] [--nodes 8]

A "step" is one pass of the whole hot path (build_path -> tables -> distance sampling -> forward/backward ->
time-domain resampling) over one batch of synthetic random-node paths.  At N = 1 the workload is
BASELINE.json configs[1]: 4096 random 8-node paths, factory constraints.  At N > 1 every rank profiles its
own 4096-path batch (weak scaling, no data-path collective; the only exchange is the NCCL all_gather of the
40-byte per-path summary rows).  Rank 0 prints ONE JSON line.

`--impl reference` times the CPU restatement of the reference (oracle/, a C port: the reference itself is
pure Python and is not present on the GPU box) on all host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "profiled trajectories/sec (8-node paths)"
UNIT = "trajectories/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(packed, seconds: float, threads: int):
    """Time the C oracle (OpenMP over paths) on a bounded sample; returns (paths/s, sample size, elapsed)."""
    import oracle
    oracle.build()
    oracle.set_sq_mode(0)
    n0 = min(packed.B, max(2 * threads, 16))
    t = time.perf_counter()
    oracle.full_batch(packed.node_attr[:n0], packed.node_flags[:n0], packed.cons[:n0], threads=threads)
    r0 = n0 / (time.perf_counter() - t)
    n = int(min(packed.B, max(n0, r0 * seconds)))
    reps = max(1, int(round(r0 * seconds / n)))          # bounded sample: about `seconds` of CPU work
    t = time.perf_counter()
    for _ in range(reps):
        s = oracle.full_batch(packed.node_attr[:n], packed.node_flags[:n], packed.cons[:n], threads=threads)
    el = time.perf_counter() - t
    assert (s[:, 4] == 0).all()
    return n * reps / el, f"{reps} x the first {n}", el


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from vexautonomousplanner_b200 import synth
    threads = os.cpu_count() or 1
    packed = synth.random_paths(args.paths, args.nodes, seed=0)
    import oracle
    oracle.build()
    sample = min(packed.B, args.ref_sample)
    sub = packed.slice(0, sample)
    for _ in range(args.warmup):
        oracle.full_batch(sub.node_attr, sub.node_flags, sub.cons, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.full_batch(sub.node_attr, sub.node_flags, sub.cons, threads=threads)
    el = time.perf_counter() - t0
    value = sample * args.steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.paths} random {args.nodes}-node paths per GPU, factory constraints, dt=0.01 dd=0.005 "
                               "(BASELINE.json configs[1])",
                   "paths_per_gpu": args.paths, "nodes": args.nodes, "sample": f"first {sample} paths per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} paths x {args.steps} steps, C restatement of the reference (oracle/), OpenMP"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torch.distributed as dist

    from vexautonomousplanner_b200 import synth
    from vexautonomousplanner_b200.engine import Engine
    from vexautonomousplanner_b200.sharding import gather_summaries

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N = args.paths, args.nodes
    packed = synth.random_paths(B, N, seed=rank)      # weak scaling: every rank its own cfg2 batch
    eng = Engine(dev, chunks=args.chunks)
    db = eng.upload(packed)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    graphed = eng.capture(db, tiles=args.tiles) if args.graph else None

    def step():
        res = graphed.run() if graphed else eng.profile(db, reuse_plan=True, tiles=args.tiles)
        if world > 1:
            gather_summaries(res.summary, counts=[B] * world)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        res = step()
    barrier()
    # ---------------- timed region: K steps, device time per step, L2 flushed between steps
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = eng.launches
    times = []
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        res = step()
        e.record()
        times.append((s, e))
    barrier()
    wall = time.perf_counter() - wall0
    launches = (eng.launches - l0) // args.steps
    dev_ms = sum(s.elapsed_time(e) for s, e in times)
    clocks = sampler.stop() if sampler else None
    tt = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms = float(tt.item())
    ms_per_step = dev_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    assert bool((res.status == 0).all().item())

    # ---------------- per-stage device times (separate eager pass, same command, CUDA events on the launch stream).
    # One untimed eager step first (it fills the allocator pools of the eager path); per stage the MEDIAN over the steps,
    # because these intervals also contain host-side launch gaps (a graph replay has none).
    eng.profile(db, reuse_plan=True)
    torch.cuda.synchronize()
    per_step = []
    for _ in range(max(args.steps, 3)):
        flush.zero_()
        eng.stage_events = []
        res = eng.profile(db, reuse_plan=True)
        torch.cuda.synchronize()
        acc = {}
        for name, s, e in eng.stage_events:
            acc[name] = acc.get(name, 0.0) + s.elapsed_time(e)
        per_step.append(acc)
    eng.stage_events = None
    stage_ms = {k: float(np.median([a[k] for a in per_step])) for k in per_step[0]}
    # algorithmic bytes per stage: 8 B per fp64 element, each array once per stage that must touch it
    Dsum = float(res.n_samples.double().sum().item())
    Tsum = float(res.n_out.double().sum().item())
    Ssum = float(B)            # random paths have one spline each
    Q, P = 1000.0 * Ssum, 1000.0 * N * B
    stage_bytes = {"S0_build_path": 8 * (2 * N + 9 * N) * B, "S1_lut": 8 * 2 * Q, "S2_props": 8 * 2 * P,
                   "S3_dist_sample": 8 * 5 * Dsum, "S45_fwd_bwd": 8 * 4 * Dsum, "S6_resample": 8 * (Dsum + 9 * Tsum)}
    peak, peak_src = measured_peak()
    stages = {k: {"ms": round(v, 4), "alg_GB": round(stage_bytes[k] / 1e9, 4),
                  "GBps": round(stage_bytes[k] / 1e9 / (v * 1e-3), 1)} for k, v in stage_ms.items()}
    dom = max(stage_ms, key=stage_ms.get)
    achieved = stage_bytes[dom] / 1e9 / (stage_ms[dom] * 1e-3)
    total_bytes = sum(stage_bytes.values())
    # measured DRAM traffic of the dominant stage from the committed ncu capture (profiles/, same workload), if present,
    # and what ncu measured per kernel (real DRAM bytes, not the algorithmic model) next to it
    traffic, ncu_kernels, prof_name = None, None, "profiles/r01b_kernel_table.json"
    try:
        with open(os.path.join(ROOT, prof_name)) as f:
            prof = json.load(f)
        if (B, N) == (4096, 8):
            traffic = prof["stages"][dom]["dram_bytes_per_step"]
            ncu_kernels = {k: {"ms": v["ms"], "dram_GBps": v["dram_GBps"], "dram_frac_of_peak": round(v["dram_GBps"] / peak, 3),
                               "fp64_pipe_pct": v["fp64_pipe_pct"]} for k, v in prof["kernels"].items() if v["ms"] >= 0.05}
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": prof_name + " (ncu dram__bytes_read+write, sum over the stage's kernels)",
                "whole_step": {"alg_GB": total_bytes / 1e9, "GBps": total_bytes / 1e9 / (ms_per_step * 1e-3),
                               "frac": total_bytes / 1e9 / (ms_per_step * 1e-3) / peak},
                "stages": stages, "ncu_kernels": ncu_kernels}

    # ---------------- end to end through the public API with HOST buffers (h2d + d2h inside the timed region):
    # numpy node tables on the host -> Engine -> every output stream of every path back in (pinned) host memory
    g2h = eng.capture(eng.upload(packed), tiles=args.e2e_tiles, to_host=True) if args.e2e_mode == "graph" else None
    hstate = None
    e2e_ms = []
    if args.e2e_mode == "stream":
        # K batches through the 3-deep pipeline of Engine.stream_to_host (fill and drain inside the timed region)
        for hres in eng.stream_to_host([packed] * 3, tiles=args.e2e_tiles):
            pass
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for hres in eng.stream_to_host([packed] * args.steps, tiles=args.e2e_tiles):
            pass
        e2e_ms = [(time.perf_counter() - t0) * 1e3 / args.steps]
    else:
        for it in range(3 + args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            if g2h is not None:
                hres = g2h.run_host(packed)
            else:
                hres = eng.profile_to_host(packed, tiles=args.e2e_tiles, state=hstate)
                hstate = hres.state
            el = (time.perf_counter() - t0) * 1e3
            if it >= 3:
                e2e_ms.append(el)
    assert int(hres.n_out.sum()) == int(res.n_out.sum().item())
    b0 = hres.path(B - 1)
    assert np.array_equal(b0["x"], res.path(B - 1)["x"])
    tt = torch.tensor([sum(e2e_ms) / len(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_step_ms = float(tt.item())
    h2d = sum(a.nbytes for a in (packed.node_attr, packed.node_flags, packed.n_nodes, packed.ap_attr, packed.ap_flags,
                                 packed.n_ap, packed.cons))
    e2e = {"value": world * B / (e2e_step_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": hres.bytes_per_step(), "ms_per_step": e2e_step_ms,
           "how": (f"host numpy node tables -> pinned -> device, {args.e2e_tiles} tile(s), dense result "
                   "rows " + ("written to pinned host memory by the pack kernels inside one CUDA graph"
                              if g2h is not None else "packed on the device and moved by the copy engine tile by tile")
                   + (f"; {args.steps} batches streamed through a 3-deep pipeline (batch n+1 computes and batch n+2 is enqueued "
                      "while batch n crosses PCIe), wall clock of the whole stream incl. fill and drain / batches"
                      if args.e2e_mode == "stream" else ", one batch at a time, wall clock incl. the final sync"))}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, n, el = cpu_port_rate(packed, args.cpu_seconds, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n} of the {B} paths, {el:.1f} s, C restatement of the reference (oracle/) with OpenMP; "
                         "the Python reference itself measured 0.37 paths/s/core on this workload (BASELINE.md)"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{B} random {N}-node paths per GPU, factory constraints, dt=0.01 dd=0.005 "
                                   "(BASELINE.json configs[1])",
                       "paths_per_gpu": B, "nodes": N, "l2": "flushed between timed steps (256 MB write)",
                       "tiles": args.tiles, "cuda_graph": bool(args.graph),
                       "mean_D": Dsum / B, "mean_T": Tsum / B, "wall_s": wall},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--paths", type=int, default=4096)
    ap.add_argument("--nodes", type=int, default=8)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-sample", type=int, default=1024)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--chunks", type=int, default=32, help="speculative chunks per path in the velocity passes")
    ap.add_argument("--tiles", type=int, default=1, help="row tiles of the batch, one CUDA stream each")
    ap.add_argument("--e2e-mode", default="stream", choices=["stream", "copy", "graph"])
    ap.add_argument("--e2e-tiles", type=int, default=1, help="tiles of the end-to-end (host in / host out) run")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the step as a CUDA graph (default), 0: eager launches")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
