#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 render core.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c4|c2]

Workload (BASELINE.json): C4 — the reference's scene literal (main.rs:810-1083) at 3840x2160, thin-lens
DoF + scatter tracer (main.rs:1129-1167), 256 epochs, depth 5, focus 3.0, blur 0.04, seed 0.  One "step"
is the whole 256-epoch frame.  With N GPUs the epochs are split contiguously over the ranks (strong
scaling: total work fixed) and the float4 {sum.rgb,count} accumulation buffers are summed with one NCCL
all-reduce inside the timed region.  `value` = primary samples ("rays" in the reference's own print,
main.rs:1169) per second over all ranks, in Mrays/s, with inputs resident in HBM.  `e2e` is the same
through the host-buffer C-ABI entry point (pinned host accumulators, H2D + D2H inside the timed region).

--impl reference times the reference's CPU algorithm (the oracle port: the Rust crate cannot be built in
this image) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, width, height, depth, epochs, tracer)
    "c4": ("C4 fixture scene 3840x2160 DoF+scatter 256 epochs depth 5 (epoch-sharded)", 3840, 2160, 5, 256, "distributed"),
    "c1": ("C1 fixture scene 1280x960 Whitted depth 5, the reference's own frame (row-sharded)", 1280, 960, 5, 1, "whitted"),
    "c2": ("C2 fixture scene 1920x1080 Whitted depth 8 (row-sharded)", 1920, 1080, 8, 1, "whitted"),
    "c3": ("C3 fixture scene 3840x2160 Whitted depth 5 (row-sharded)", 3840, 2160, 5, 1, "whitted"),
}
FLOP_TRI, FLOP_SPH = 36.0, 28.0   # algorithmic flop per ray x triangle / ray x sphere pair


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if "Active" in r[5 + k] and "Not" not in r[5 + k]:
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            sm.sort()
            # median of the samples under load (upper half of the clock distribution excludes idle samples)
            load = [v for v in sm if v >= 0.5 * max(sm)]
            out["sm_mhz"] = load[len(load) // 2]
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def shard(total: int, rank: int, world: int):
    """Contiguous split [g*T/G, (g+1)*T/G)."""
    import __graft_entry__ as ge
    ge.load_package()
    from b200rt.sharding import shard_range
    return shard_range(total, rank, world)


# --------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU algorithm (oracle port) on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob      # bench.py's reference arm is one of the places allowed to run the oracle
    b = ob.b
    desc, W, H, depth, epochs, tracer = WORKLOADS[args.workload]
    world = b.World.fixture()
    cam = b.fixture_camera()
    cores = ob.max_threads()
    # bounded sample: a band of rows through the middle of the frame x a few epochs, sized from a probe
    rows, ep = 32, 1
    params = b.default_params(width=W, height=H, depth=depth, seed=0, row_begin=H // 2 - rows // 2, row_count=rows)

    def one(rows_, ep_):
        p = b.copy_params(params, row_begin=H // 2 - rows_ // 2, row_count=rows_)
        t0 = time.perf_counter()
        if tracer == "distributed":
            ob.render_distributed(world.scene(), cam, p, 0, ep_)
        else:
            ob.render_whitted(world.scene(), cam, p)
        return time.perf_counter() - t0

    t_probe = one(rows, ep)
    target_s = 8.0   # per step; (warmup + steps) * target stays within a few minutes
    scale = max(1.0, target_s / max(t_probe, 1e-3))
    if tracer == "distributed":
        ep = int(min(8, max(1, round(scale ** 0.5))))
        rows = int(min(H, max(32, round(rows * scale / ep))))
    else:
        rows = int(min(H, max(32, round(rows * scale))))
    for _ in range(args.warmup):
        one(rows, ep)
    ts = [one(rows, ep) for _ in range(args.steps)]
    samples = rows * W * ep
    t = sum(ts) / len(ts)
    v = samples / t / 1e6
    sample = f"rows [{H // 2 - rows // 2},{H // 2 - rows // 2 + rows}) x {ep} epoch(s) of {desc}: {samples} samples/step"
    line = {
        "impl": "reference", "metric": "Mrays/s (primary samples/s, the reference's own 'rays/s', main.rs:1169)",
        "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": desc, "width": W, "height": H, "depth": depth, "epochs": epochs,
                                        "sample": sample},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C++ restatement of the reference CPU algorithm (the Rust crate cannot be built here: no rustc/cargo)",
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    b = ge.load_package()

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world_size,
                                device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    desc, W, H, depth, epochs, tracer = WORKLOADS[args.workload]
    if args.width: W = args.width
    if args.height: H = args.height
    if args.epochs: epochs = args.epochs
    reduced = bool(args.width or args.height or args.epochs)

    ctx = b.Context(local_rank)          # raises without a GPU: there is no CPU fallback
    scene_world = b.World.fixture()
    ctx.upload_scene(scene_world)
    cam = b.fixture_camera()
    tracer_id = b.TRACER_WAVEFRONT if args.tracer == "wavefront" else b.TRACER_MEGAKERNEL
    params = b.default_params(width=W, height=H, depth=depth, seed=0, tracer=tracer_id)
    # every launch, copy, collective and timing event of the bench goes on this one stream
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    if tracer == "distributed":
        e0, en = shard(epochs, rank, world_size)
        d_accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
        h_accum = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
        samples_total = W * H * epochs

        def step_device():
            d_accum.zero_()
            ctx.render_distributed_device(cam, params, e0, en, d_accum.data_ptr(), stream)
            if world_size > 1:
                dist.all_reduce(d_accum, op=dist.ReduceOp.SUM)

        def step_e2e():
            h_accum.zero_()
            if world_size == 1:
                # host-buffer C-ABI entry: pinned accumulators travel H2D, are added to, and travel back
                ctx.render_distributed(cam, params, e0, en, h_accum.numpy())
            else:
                # N > 1: the ranks' accumulators are summed on the devices, so the caller does the copies around the
                # device-pointer entry of the same C ABI: pinned H2D, render, all-reduce, D2H of the reduced image
                d_accum.copy_(h_accum, non_blocking=True)
                ctx.render_distributed_device(cam, params, e0, en, d_accum.data_ptr(), stream)
                dist.all_reduce(d_accum, op=dist.ReduceOp.SUM)
                h_accum.copy_(d_accum, non_blocking=True)
                torch.cuda.synchronize()
            return float(h_accum[H // 2, W // 2, 3])

        h2d = d2h = W * H * 16
        launches_per_step = 1
    else:
        r0, rn = shard(H, rank, world_size)
        p_rank = b.copy_params(params, row_begin=r0, row_count=rn)
        d_rgb = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
        h_rgb = torch.zeros((H, W, 3), dtype=torch.float32).pin_memory()
        samples_total = W * H

        def step_device():
            ctx.render_whitted_device(cam, p_rank, d_rgb.data_ptr(), 0, stream)
            if world_size > 1:
                dist.all_reduce(d_rgb, op=dist.ReduceOp.SUM)   # disjoint rows: sum == gather

        def step_e2e():
            ctx.render_whitted(cam, p_rank, out_rgb=h_rgb.numpy(), want_prim_id=False)
            return float(h_rgb[r0, W // 2, 0])

        h2d, d2h = 0, rn * W * 12
        launches_per_step = 1

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    ctx.reset_stats()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler: sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = ev0.elapsed_time(ev1)
    st = ctx.stats()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = samples_total / (ms_per_step * 1e-3) / 1e6

    # per-launch duration of the dominant kernel, measured live with CUDA events on the launching stream.
    #  * wavefront: the dominant kernel is wf_cast_kernel (World::cast for every ray of a round); the library brackets
    #    each of its launches with events (b200rt_set_kernel_timing) during one extra step;
    #  * megakernel / Whitted: one trace_kernel launch per step.
    barrier()
    cast_launches = 0
    logic_ms = None
    wavefront = tracer == "distributed" and args.tracer == "wavefront"
    if wavefront:
        ctx.set_kernel_timing(True)
        d_accum.zero_()
        ctx.render_distributed_device(cam, params, e0, en, d_accum.data_ptr(), stream)
        barrier()
        kst = ctx.stats()
        ctx.set_kernel_timing(False)
        k_ms_total = kst["cast_kernel_ms"]
        cast_launches = kst["cast_kernel_launches"]
        logic_ms = kst["logic_kernel_ms"]
        k_ms = k_ms_total / max(cast_launches, 1)
        launches_per_step = kst["kernel_launches"]
        step_kernel_ms = kst["kernel_ms"]
    else:
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, c in kev:
            if tracer == "distributed":
                d_accum.zero_()
                a.record()
                ctx.render_distributed_device(cam, params, e0, en, d_accum.data_ptr(), stream)
                c.record()
            else:
                a.record()
                ctx.render_whitted_device(cam, p_rank, d_rgb.data_ptr(), 0, stream)
                c.record()
        barrier()
        kernel_ms = [a.elapsed_time(c) for a, c in kev]
        k_ms = sum(kernel_ms) / len(kernel_ms)
        k_ms_total = k_ms
        cast_launches = 1
        step_kernel_ms = k_ms

    # ---- end-to-end through the host-buffer C ABI --------------------------------------------------------------
    e2e_steps = max(1, min(args.steps, 3))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = samples_total / (e2e_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (this rank's launch) -----------------------------------------------------
    peaks, peak_src = measured_peaks()
    info = ctx.device_info()
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    peak_tflops = info["sm_count"] * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    tri_pairs = st["tri_pair_tests"] / args.steps
    sph_pairs = st["sph_pair_tests"] / args.steps
    flops = tri_pairs * FLOP_TRI + sph_pairs * FLOP_SPH          # algorithmic flop of one step on this rank
    achieved = flops / (k_ms_total * 1e-3) / 1e12                # ... over the device time of the dominant kernel's launches
    live_peak, live_mhz = ctx.measure_fp32_peak()
    # scenes of one 64-triangle tile run the rays-in-lanes cast kernel (rt_wavefront.cu), larger ones the warp-transposed one
    kname = (("wf_cast_rl_kernel" if scene_world.scene().n_triangles <= 64 else "wf_cast_kernel") if wavefront
             else f"trace_kernel<{tracer}>")
    # DRAM bytes per launch of the dominant kernel from the committed ncu capture (same workload only)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if wavefront and not (args.width or args.height) and world_size == 1 and kname in tj:
            traffic = tj[kname]["dram_bytes_per_launch"]
    except Exception:
        traffic = None
    n_l = max(cast_launches, 1)
    roofline = {
        "bound": "fp32", "kernel": kname, "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": achieved / peak_tflops, "traffic": traffic,
        "peak_source": f"SMs({info['sm_count']}) x 128 lanes x 2 flop x sm_max_mhz({sm_max_mhz:.0f}, {peak_src} MEASURED_PEAKS.json)",
        "ffma_loop_tflops_live": live_peak, "frac_of_live_ffma_loop": achieved / live_peak if live_peak else None,
        "launches_per_step": cast_launches, "avg_launch_ms": k_ms,
        "algorithmic_flop_per_launch": flops / n_l,
        "pair_tests_per_launch": {"tri": tri_pairs / n_l, "sph": sph_pairs / n_l}, "kernel_ms": k_ms_total,
        "casts_per_launch": st["casts"] / args.steps / n_l,
        "share_of_step": k_ms_total / step_kernel_ms if step_kernel_ms else None,
        "other_kernels_ms": logic_ms,
        "whole_step_frac": flops / (step_kernel_ms * 1e-3) / 1e12 / peak_tflops if step_kernel_ms else None,
    }

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample (rank 0, N=1 only) ------------------
    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_binding as ob
        cores = ob.max_threads()
        rows, ep = 32, 1
        p = b.copy_params(params, row_begin=H // 2 - rows // 2, row_count=rows)
        t0 = time.perf_counter()
        (ob.render_distributed(scene_world.scene(), cam, p, 0, ep) if tracer == "distributed"
         else ob.render_whitted(scene_world.scene(), cam, p))
        t_probe = time.perf_counter() - t0
        scale = max(1.0, 15.0 / max(t_probe, 1e-3))
        rows = int(min(H, max(32, round(rows * scale))))
        p = b.copy_params(params, row_begin=H // 2 - rows // 2, row_count=rows)
        t0 = time.perf_counter()
        (ob.render_distributed(scene_world.scene(), cam, p, 0, ep) if tracer == "distributed"
         else ob.render_whitted(scene_world.scene(), cam, p))
        t_cpu = time.perf_counter() - t0
        n = rows * W * ep
        cpu = {"value": n / t_cpu / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
               "sample": f"rows [{H // 2 - rows // 2},{H // 2 - rows // 2 + rows}) x {ep} epoch of the same workload: {n} samples in {t_cpu:.1f} s"}

    if rank == 0:
        line = {
            "metric": "Mrays/s (primary samples/s, the reference's own 'rays/s', main.rs:1169)",
            "value": value, "unit": "Mrays/s", "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc + (" [REDUCED SIZE: dev run]" if reduced else ""), "width": W, "height": H,
                       "depth": depth, "epochs": epochs, "sharding": ("epochs" if tracer == "distributed" else "rows"),
                       "cast_mode": "two_phase", "tracer": args.tracer if tracer == "distributed" else "megakernel", "l2": "path state + ray buffers (%.1f GB per 16-epoch batch) and the accumulation buffer (%.1f MB) exceed L2; scene records stay in registers / L1" % (W * H * 16 * 440 / 1e9, W * H * 16 / 1e6)},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms},
            "gpu_launches": int(launches_per_step) * args.steps, "clocks": clocks,
            "casts_per_s_in_dominant_kernel": st["casts"] / args.steps / (k_ms_total * 1e-3),
            "pair_tests_per_s_in_dominant_kernel": (tri_pairs + sph_pairs) / (k_ms_total * 1e-3),
        }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--width", type=int, default=0, help="dev only: override (marks the line REDUCED)")
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--epochs", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tracer", default="wavefront", choices=["wavefront", "megakernel"],
                    help="GPU schedule of the stochastic tracer (same samples, same bits)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
