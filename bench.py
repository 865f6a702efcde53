#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 render core.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c4|c1|c2|c3|c5]

Workload (BASELINE.json): C4 — the reference's scene literal (main.rs:810-1083) at 3840x2160, thin-lens
DoF + scatter tracer (main.rs:1129-1167), 256 epochs, depth 5, focus 3.0, blur 0.04, seed 0.  One "step"
is the whole 256-epoch frame.  Every rank is one member of a device group of the C ABI (b200rt_group_*,
include/b200rt.h): the library splits the epochs contiguously over the ranks (strong scaling: total work
fixed) and sums the float4 {sum.rgb,count} accumulation buffers with one ncclReduce to rank 0 inside the
timed region.  `value` = pixel samples per second over all ranks, in Mrays/s, with the scene resident in
HBM and the frame left on rank 0's GPU.  `e2e` is the same through the host-buffer entry points (scene,
camera, parameters in; the reduced frame out to a pinned host buffer on rank 0).  c1/c2/c3 are the Whitted
frames (rows sharded), c5 the 100 352-triangle mesh with 10 M one-sample photons (rows sharded).

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
    "c5": ("C5 fixture scene + 100 352-triangle OBJ height field, 4000x2500 = 10 M one-sample photons, depth 5 (row-sharded)", 4000, 2500, 5, 1, "distributed"),
}
FLOP_TRI, FLOP_SPH = 36.0, 28.0   # algorithmic flop per ray x triangle / ray x sphere pair
# "rays" = pixel samples GENERATED per second (width x height x epochs / time).  The reference's own print (main.rs:1169)
# counts the samples it ACCEPTED (non-normal ones are dropped, main.rs:1157-1160: ~20 % on this scene); that figure is
# reported beside it as accepted_samples_per_s.  Both arms use the same convention.
METRIC = "Mrays/s (pixel samples generated per second; accepted samples, the reference's own 'rays/s' of main.rs:1169, beside it)"


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
    """The reference's CPU algorithm (oracle port) on all host cores; rank 0 only.  The scene comes from
    tests/golden/fixture_scene.npz: this arm loads oracle/liboracle.so and nothing of the product."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob      # bench.py's reference arm is one of the places allowed to run the oracle
    desc, W, H, depth, epochs, tracer = WORKLOADS[args.workload]
    fx = ob.GoldenFixture()
    scene, cam = fx.scene, fx.camera
    if args.workload == "c5":
        print(json.dumps({"impl": "reference", "unavailable": "c5 needs the 100 352-triangle OBJ mesh through the product's importer; its CPU sample is bench.py's cpu_baseline leg"}), flush=True)
        return 0
    # torchrun exports OMP_NUM_THREADS=1: the thread count is passed explicitly (every core this process may use)
    cores = ob.host_threads()
    # bounded sample: a band of rows through the middle of the frame x a few epochs, sized from a probe
    rows, ep = 32, 1

    def one(rows_, ep_):
        p = fx.params(width=W, height=H, depth=depth, seed=0, row_begin=H // 2 - rows_ // 2, row_count=rows_)
        t0 = time.perf_counter()
        if tracer == "distributed":
            _, cnt = ob.render_distributed(scene, cam, p, 0, ep_, n_threads=cores)
        else:
            _, _, cnt = ob.render_whitted(scene, cam, p, n_threads=cores)
        return time.perf_counter() - t0, cnt["samples"]

    t_probe, _ = one(rows, ep)
    target_s = 8.0   # per step; (warmup + steps) * target stays within a few minutes
    scale = max(1.0, target_s / max(t_probe, 1e-3))
    if tracer == "distributed":
        ep = int(min(8, max(1, round(scale ** 0.5))))
        rows = int(min(H, max(32, round(rows * scale / ep))))
    else:
        rows = int(min(H, max(32, round(rows * scale))))
    for _ in range(args.warmup):
        one(rows, ep)
    runs = [one(rows, ep) for _ in range(args.steps)]
    samples = rows * W * ep
    t = sum(r[0] for r in runs) / len(runs)
    accepted = runs[-1][1]
    v = samples / t / 1e6
    sample = f"rows [{H // 2 - rows // 2},{H // 2 - rows // 2 + rows}) x {ep} epoch(s) of {desc}: {samples} samples/step"
    line = {
        "impl": "reference", "metric": METRIC,
        "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": desc, "width": W, "height": H, "depth": depth, "epochs": epochs,
                                        "sample": sample},
        "accepted_samples_per_s": accepted / t / 1e6 if tracer == "distributed" else v,
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
    root = rank == 0

    desc, W, H, depth, epochs, tracer = WORKLOADS[args.workload]
    if args.width: W = args.width
    if args.height: H = args.height
    if args.epochs: epochs = args.epochs
    reduced = bool(args.width or args.height or args.epochs)
    by_rows = args.workload == "c5"     # one epoch: the frame is split by rows; C4 splits its 256 epochs

    # ---- the scene ------------------------------------------------------------------------------------------------
    if args.workload == "c5":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from scene_util import fixture_plus_mesh
        tmp = tempfile.mkdtemp(prefix="b200rt_c5_")
        scene_world, _ = fixture_plus_mesh(b, tmp, 225)      # fixture scene + the 100 352-triangle OBJ height field
    else:
        scene_world = b.World.fixture()
    sc = scene_world.scene()
    scene_bytes = (sc.n_triangles * ctypes.sizeof(b.Triangle) + sc.n_spheres * ctypes.sizeof(b.Sphere) +
                   sc.n_materials * ctypes.sizeof(b.Material) + sc.n_lights * ctypes.sizeof(b.Light))
    cam = b.fixture_camera()
    tracer_id = b.TRACER_WAVEFRONT if args.tracer == "wavefront" else b.TRACER_MEGAKERNEL
    cast_id = b.CAST_BVH if args.cast == "bvh" else b.CAST_TWO_PHASE
    params = b.default_params(width=W, height=H, depth=depth, seed=0, tracer=tracer_id, cast_mode=cast_id)

    # ---- the device group: one rank per process, NCCL inside the library (include/b200rt.h, "device groups") -------------
    # The id of the NCCL communicator is made on rank 0 and handed to the other ranks (here: a torch broadcast).
    uid = None
    if world_size > 1:
        t = torch.zeros(b.GROUP_ID_BYTES, dtype=torch.uint8, device=dev)
        if root:
            t.copy_(torch.frombuffer(bytearray(b.Group.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        uid = bytes(t.cpu().numpy().tobytes())
    group = b.Group.rank(local_rank, rank, world_size, uid)     # raises without a GPU: there is no CPU fallback
    group.upload_scene(scene_world)
    ctx = group.member(0)

    if tracer == "distributed":
        d_accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev) if root else None
        h_accum = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory() if root else None
        samples_total = W * H * epochs

        def step_device():
            group.render_distributed_device(cam, params, 0, epochs, d_accum.data_ptr() if root else 0, by_rows=by_rows)
            return group.last_render_ms()

        def step_e2e():
            # the call a user makes: scene + camera + parameters in (host), the reduced frame out (pinned host buffer on
            # rank 0).  The accumulators of a fresh frame start at zero ON the devices: nothing else travels H2D.
            group.upload_scene(scene_world)
            group.render_distributed(cam, params, 0, epochs, h_accum.numpy() if root else None, by_rows=by_rows)
            return float(h_accum[H // 2, W // 2, 3]) if root else 0.0

        h2d, d2h = scene_bytes + ctypes.sizeof(b.Camera) + ctypes.sizeof(b.Params), W * H * 16
    else:
        d_rgb = torch.zeros((H, W, 3), dtype=torch.float32, device=dev) if root else None
        h_rgb = torch.zeros((H, W, 3), dtype=torch.float32).pin_memory() if root else None
        samples_total = W * H

        def step_device():
            group.render_whitted_device(cam, params, d_rgb.data_ptr() if root else 0)
            return group.last_render_ms()

        def step_e2e():
            group.upload_scene(scene_world)
            group.render_whitted(cam, params, out_rgb=h_rgb.numpy() if root else None, want_prim_id=False)
            return float(h_rgb[H // 2, W // 2, 0]) if root else 0.0

        h2d, d2h = scene_bytes + ctypes.sizeof(b.Camera) + ctypes.sizeof(b.Params), W * H * 12

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing: CUDA events on the launching stream around render + collective, inside the library ------
    for _ in range(args.warmup):
        step_device()
    barrier()
    ctx.reset_stats()
    sampler = ClockSampler(local_rank) if root else None
    if sampler: sampler.start()
    barrier()
    t0 = time.perf_counter()
    ms_total = 0.0
    for _ in range(args.steps):
        ms_total += step_device()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler else None
    st = ctx.stats()
    ms_total = max_over_ranks(ms_total)
    wall_ms = max_over_ranks(wall_ms)
    ms_per_step = ms_total / args.steps
    value = samples_total / (ms_per_step * 1e-3) / 1e6
    accepted = torch.tensor([float(st["samples"]) / args.steps], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(accepted, op=dist.ReduceOp.SUM)
    accepted = float(accepted.item())

    # per-launch duration of the dominant kernel, measured live with CUDA events on the launching stream.
    #  * wavefront: the dominant kernel is the cast kernel (World::cast for every ray of a round); the library brackets
    #    each of its launches with events (b200rt_set_kernel_timing) during one extra step;
    #  * megakernel / Whitted: one trace_kernel launch per step.
    barrier()
    logic_ms = None
    wavefront = tracer == "distributed" and args.tracer == "wavefront"
    if wavefront:
        ctx.set_kernel_timing(True)
        step_device()
        barrier()
        kst = ctx.stats()
        ctx.set_kernel_timing(False)
        k_ms_total = kst["cast_kernel_ms"]
        cast_launches = kst["cast_kernel_launches"]
        logic_ms = kst["logic_kernel_ms"]
        primary_ms = kst.get("primary_kernel_ms", 0.0)
        k_ms = k_ms_total / max(cast_launches, 1)
        launches_per_step = kst["kernel_launches"]
        step_kernel_ms = kst["kernel_ms"]
    else:
        step_device()
        barrier()
        kst = ctx.stats()
        k_ms = k_ms_total = step_kernel_ms = kst["kernel_ms"]
        cast_launches = 1
        launches_per_step = 1
        primary_ms = 0.0

    # ---- end-to-end through the host-buffer C ABI: wall clock around the call a user makes -----------------------------
    e2e_steps = max(1, min(args.steps, 3))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    e2e_value = samples_total / (e2e_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (this rank's launches) ---------------------------------------------------------
    peaks, peak_src = measured_peaks()
    info = ctx.device_info()
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    peak_tflops = info["sm_count"] * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    tri_pairs = st["tri_pair_tests"] / args.steps
    sph_pairs = st["sph_pair_tests"] / args.steps
    step_flops = tri_pairs * FLOP_TRI + sph_pairs * FLOP_SPH     # algorithmic flop of one step on this rank (every cast)
    casts_step = st["casts"] / args.steps
    primary_casts = 0.0
    if wavefront and primary_ms > 0.0:
        # round 0 runs in another kernel (wf_cast_rl[_tiled]_primary_kernel: camera-ray generation + the primary cast, one
        # per pixel sample of this rank): its casts and its time are not part of the dominant kernel's figures
        primary_casts = float(st["samples_generated"]) / args.steps if "samples_generated" in st else float(
            (shard(H, rank, world_size)[1] * W * epochs) if by_rows else (W * H * shard(epochs, rank, world_size)[1]))
        tri_pairs -= primary_casts * sc.n_triangles
        sph_pairs -= primary_casts * sc.n_spheres
        casts_step -= primary_casts
    flops = tri_pairs * FLOP_TRI + sph_pairs * FLOP_SPH          # algorithmic flop of the dominant kernel's launches
    achieved = flops / (k_ms_total * 1e-3) / 1e12                # ... over the device time of the dominant kernel's launches
    live_peak, live_mhz = ctx.measure_fp32_peak()
    # scenes of one 64-triangle tile run the rays-in-lanes cast kernel (rt_wavefront.cu), larger ones its tiled (TMA) form
    kname = (("wf_cast_rl_kernel" if sc.n_triangles <= 64 else "wf_cast_rl_tiled_kernel") if wavefront
             else f"trace_kernel<{tracer}>")
    # DRAM bytes per launch of the dominant kernel: NOT measured by this run - the figure of the committed ncu capture of
    # the same command line (profiles/r2_traffic.json), reported with its source
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if wavefront and not reduced and world_size == 1 and kname in tj:
            traffic = tj[kname]["dram_bytes_per_launch"]
            traffic_src = "profiles/r2_traffic.json: " + tj[kname].get("source", "ncu capture")
    except Exception:
        traffic = None
    n_l = max(cast_launches, 1)
    roofline = {
        "bound": "fp32", "kernel": kname, "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": achieved / peak_tflops, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": f"SMs({info['sm_count']}) x 128 lanes x 2 flop x sm_max_mhz({sm_max_mhz:.0f}, {peak_src} MEASURED_PEAKS.json)",
        "ffma_loop_tflops_live": live_peak, "frac_of_live_ffma_loop": achieved / live_peak if live_peak else None,
        "launches_per_step": cast_launches, "avg_launch_ms": k_ms,
        "algorithmic_flop_per_launch": flops / n_l,
        "pair_tests_per_launch": {"tri": tri_pairs / n_l, "sph": sph_pairs / n_l}, "kernel_ms": k_ms_total,
        "casts_per_launch": casts_step / n_l,
        "share_of_step": k_ms_total / step_kernel_ms if step_kernel_ms else None,
        "other_kernels_ms": logic_ms,
        "primary_cast_kernel_ms": primary_ms, "primary_casts_per_step": primary_casts,
        "whole_step_frac": step_flops / (step_kernel_ms * 1e-3) / 1e12 / peak_tflops if step_kernel_ms else None,
    }

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample (rank 0, N=1 only) ------------------
    cpu = None
    if root and world_size == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_binding as ob
        cores = ob.host_threads()
        rows, ep = (1, 1) if args.workload == "c5" else (32, 1)    # (a C5 cast tests 100 420 primitives)

        def cpu_run(rows_):
            p = b.copy_params(params, row_begin=H // 2 - rows_ // 2, row_count=rows_)
            t0 = time.perf_counter()
            (ob.render_distributed(sc, cam, p, 0, ep, n_threads=cores) if tracer == "distributed"
             else ob.render_whitted(sc, cam, p, n_threads=cores))
            return time.perf_counter() - t0

        t_probe = cpu_run(rows)
        rows = int(min(H, max(rows, round(rows * max(1.0, 15.0 / max(t_probe, 1e-3))))))
        t_cpu = cpu_run(rows)
        n = rows * W * ep
        cpu = {"value": n / t_cpu / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
               "sample": f"rows [{H // 2 - rows // 2},{H // 2 - rows // 2 + rows}) x {ep} epoch of the same workload: {n} samples in {t_cpu:.1f} s"}

    if root:
        line = {
            "metric": METRIC,
            "value": value, "unit": "Mrays/s", "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc + (" [REDUCED SIZE: dev run]" if reduced else ""), "width": W, "height": H,
                       "depth": depth, "epochs": epochs, "sharding": ("rows" if (by_rows or tracer != "distributed") else "epochs"),
                       "collective": "NCCL inside libb200rt.so (b200rt_group_*): " + ("ncclSend/ncclRecv gather of row bands to rank 0" if (by_rows or tracer != "distributed") else "one ncclReduce(sum) of the accumulators to rank 0"),
                       "cast_mode": "two_phase" if args.cast != "bvh" else "bvh (SURVEY 8f N1: the same hits through the acceleration structure - NOT the brute-force walk the roofline figure is defined on; roofline.frac of this line counts the pairs the reference would test)",
                       "tracer": args.tracer if tracer == "distributed" else "megakernel",
                       "timing": "per step: CUDA events on the launching stream around accumulator clear + render + collective (b200rt_group_last_render_ms), max over ranks",
                       "l2": "path state + ray buffers (%.1f GB per 16-epoch batch) and the accumulation buffer (%.1f MB) exceed L2; scene records stay in the constant bank / shared memory" % (W * H * 16 * 440 / 1e9, W * H * 16 / 1e6)},
            "accepted_samples_per_s": accepted / (ms_per_step * 1e-3) / 1e6,
            "wall_ms_per_step": wall_ms / args.steps,
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms,
                    "what": "wall clock around b200rt_group_upload_scene + b200rt_group_render_* with host buffers: scene / camera / parameters H2D, render, collective, D2H of the frame on rank 0"},
            "gpu_launches": int(launches_per_step) * args.steps, "clocks": clocks,
            "casts_per_s_in_dominant_kernel": casts_step / (k_ms_total * 1e-3),
            "pair_tests_per_s_in_dominant_kernel": (tri_pairs + sph_pairs) / (k_ms_total * 1e-3),
        }
        print(json.dumps(line), flush=True)
    group.close()
    if world_size > 1:
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------------------------
def run_k2(args):
    """K2: World::cast on its own (b200rt_intersect_device / b200rt_intersect): the primary camera rays of the 3840x2160
    frame of the fixture scene (8.3 M rays, the population the tracers produce), resident AoS rays and hits of the public
    API.  One step = one cast of every ray; value = casts per second (M rays/s).  With N GPUs the rays are split
    contiguously over the ranks (no collective: the path has no exchange step)."""
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
        dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    W, H = 3840, 2160
    ctx = b.Context(local_rank)
    world = b.World.fixture()
    ctx.upload_scene(world)
    sc = world.scene()
    cam = b.fixture_camera()
    # Camera::shoot (main.rs:84-99) for every pixel, in numpy (test / bench infrastructure: the rays are the INPUT)
    toward = np.array(cam.toward, dtype=np.float64); toward /= np.linalg.norm(toward)
    right = np.cross(toward, np.array(cam.up, dtype=np.float64)); right /= np.linalg.norm(right)
    up2 = np.cross(right, toward); up2 /= np.linalg.norm(up2)
    t = np.tan(cam.fovy / 2)
    ys, xs = np.mgrid[0:H, 0:W]
    d = ((xs - W / 2) / H)[..., None] * (t * right) + ((H / 2 - ys) / H)[..., None] * (t * up2) + toward
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    rays = np.zeros(W * H, dtype=b.RAY_DTYPE)
    rays["origin"] = (np.array(cam.center) + toward * cam.near).astype(np.float32)
    rays["direction"] = d.reshape(-1, 3).astype(np.float32)
    rays["exclude_prim"] = -1
    r0, rn = shard(W * H, rank, world_size)
    rays = rays[r0:r0 + rn]
    n = len(rays)
    h_rays = torch.from_numpy(rays.view(np.uint8).copy()).pin_memory()
    d_rays = h_rays.to(dev)
    d_hits = torch.empty(n * b.HIT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    h_hits = np.zeros(n, dtype=b.HIT_DTYPE)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # L2: 8.3 M rays x (36 B in + 48 B out) = 697 MB per step exceed the 126 MB L2
    for _ in range(max(args.warmup, 3)):
        ctx.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), b.CAST_TWO_PHASE, stream)
    barrier()
    ctx.reset_stats()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler: sampler.start()
    steps = max(args.steps, 20)            # a step is 0.7 ms: enough of them for the clock sampler to see the load
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        ctx.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), b.CAST_TWO_PHASE, stream)
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_per_step = max_over_ranks(ev0.elapsed_time(ev1) / steps)
    st = ctx.stats()
    total_rays = W * H
    value = total_rays / (ms_per_step * 1e-3) / 1e6
    # end to end: host rays in, host hits out
    ctx.intersect(rays)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        h_hits = ctx.intersect(rays)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / 3)
    peaks, peak_src = measured_peaks()
    info = ctx.device_info()
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    peak_tflops = info["sm_count"] * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    flops = n * (sc.n_triangles * FLOP_TRI + sc.n_spheres * FLOP_SPH)
    achieved = flops / (ms_per_step * 1e-3) / 1e12
    if rank == 0:
        line = {
            "metric": "Mrays/s (World::cast calls per second, the intersection kernel on its own)", "value": value, "unit": "Mrays/s",
            "n_gpus": world_size, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "K2: World::cast of the 8 294 400 primary camera rays of the fixture scene at 3840x2160 (64 triangles + 4 spheres per cast)",
                       "sharding": "rays", "l2": "697 MB of rays + hits per step exceed L2"},
            "roofline": {"bound": "fp32", "kernel": "intersect_rl_kernel", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops, "traffic": None,
                         "peak_source": f"SMs({info['sm_count']}) x 128 lanes x 2 flop x sm_max_mhz({sm_max_mhz:.0f}, {peak_src} MEASURED_PEAKS.json)",
                         "exact_tests_per_cast": st["exact_confirms"] / max(st["casts"], 1), "hit_fraction": float((h_hits["prim_id"] >= 0).mean())},
            "cpu_baseline": None,
            "e2e": {"value": total_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": n * 36, "d2h_bytes_per_step": n * 48,
                    "ms_per_step": e2e_ms},
            "gpu_launches": steps, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world_size > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["k2"])
    ap.add_argument("--width", type=int, default=0, help="dev only: override (marks the line REDUCED)")
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--epochs", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cast", default="two_phase", choices=["two_phase", "bvh"],
                    help="World::cast: the reference's brute-force walk (two-phase cast; the headline) or the acceleration structure (same hits)")
    ap.add_argument("--tracer", default="wavefront", choices=["wavefront", "megakernel"],
                    help="GPU schedule of the stochastic tracer (same samples, same bits)")
    args = ap.parse_args()
    if args.workload == "k2":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "k2 is the GPU kernel's own micro-workload; the CPU arm is timed on c1..c5"}), flush=True)
            return 0
        return run_k2(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
