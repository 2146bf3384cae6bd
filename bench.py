#!/usr/bin/env python
"""bench.py — throughput of the path-tracing hot path (Scene::update) on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg4] [--partition rows|samples|samples-strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload: BASELINE config 4, the north-star target (random_spheres 3840x2160, 4096 spp, depth 50), at every N.
One "step" = one `Scene::update` of the workload (one frame of `samples` spp over the whole image).
  value   : Mrays/s, whole job, image resident in HBM (pt_render_device on torch's current stream, CUDA events).
  e2e     : same metric through the reference-facing call with a HOST buffer — ONE `Scene::update` mirror -> pt_render
            call per step whatever N is (N > 1: a multi-device scene, the library fans out one host thread per GPU),
            previous frame uploaded and result downloaded every step (frame_num >= 1).  Steps longer than 2 s: one e2e step.
  roofline: the megakernel's ALGORITHMIC flop — 16 per (ray, sphere) test x rays x spheres, the FP32 formulation of SURVEY §8d —
            against the FP32 FMA peak, as in round 1.  Since round 2 the two dot products of every test run on the tensor
            path (mma.sync f16 split operands, pt_sweep_mma.cuh), so the fraction is a comparison of builds, not a pipe
            utilisation: the note says how much HMMA work that is; HBM stays idle.
  cpu_baseline: the CPU oracle (restated reference, list mode = the reference's live path) on the host cores,
            bounded sample, rank 0 at N=1 only.
N > 1 (one process per GPU): default `--partition rows` is STRONG scaling of one frame by interleaved row tiles with no
collective in the data path (the image partitions by rows, north_star); `--partition samples` is weak scaling — every
rank renders the same image with its own frame seed (the reference's progressive-frame semantics, scene.rs:86-87,99-101)
and one NCCL reduce forms the equal-weight mean.
  extras  : (N=1) one short run each of the other BASELINE configs so that every config has a driver-visible figure.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (preset, width, height, spp, depth)   — BASELINE.json configs
    "cfg1": ("random_spheres", 200, 100, 100, 50),
    "cfg2": ("random_spheres", 1200, 800, 1024, 50),
    "cfg3": ("two_perlin_spheres", 1920, 1080, 1024, 50),
    "cfg4": ("random_spheres", 3840, 2160, 4096, 50),
    "cfg5": ("stress100k", 1920, 1080, 512, 50),
}
FLOP_PER_TEST = 16  # one (ray, sphere) test in the SoA form


def workload_name(key, spp):
    preset, w, h, s, d = WORKLOADS[key]
    return "%s %dx%d %dspp depth%d (%s)" % (preset, w, h, spp, d, key)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner under the driver's
    NCCL_DEBUG setting, which this script leaves alone): from here on file descriptor 1 points at stderr, and only emit()
    writes to the real stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (json.dumps(line) + "\n").encode())


def run_reference(args):
    """--impl reference: the reference's CPU path (restated: oracle, list mode) on the host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    preset, w, h, spp_full, depth = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    spp = args.ref_spp
    scene = orc.Scene(preset, w, h)
    times, rays_total = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, rays = scene.update(spp, depth, frame_num=0, mode=orc.HIT_LIST, nthreads=cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            rays_total += rays
    total = sum(times)
    mrays = rays_total / 1e6 / total
    sample = "%s at %d spp of %d (full resolution, list mode = live reference path, %d threads)" % (workload_name(args.workload, spp_full), spp, spp_full, cores)
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong" if args.partition != "samples" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, spp_full), "sample": sample},
        "samples_per_s": w * h * spp * args.steps / total,
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Rust and cannot be built in this image (no rustc/cargo): this is the C++ restatement in oracle/",
    }
    emit(line)
    return 0


def cpu_baseline(workload, budget_s=12.0):
    """Oracle timed on the host cores for a bounded sample of the same workload (reported, not a target)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    preset, w, h, spp_full, depth = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    scene = orc.Scene(preset, w, h)
    out = {}
    for key, mode in (("list", orc.HIT_LIST), ("soa_avx2", orc.HIT_SOA_AVX2 if orc.lib().orc_has_avx2() else orc.HIT_SOA_SCALAR)):
        t0 = time.perf_counter()
        _, rays = scene.update(1, depth, mode=mode, nthreads=cores)  # calibration pass: 1 spp
        dt = time.perf_counter() - t0
        spp = int(max(1, min(spp_full, (budget_s / 2) / max(dt, 1e-3))))
        t0 = time.perf_counter()
        _, rays = scene.update(spp, depth, mode=mode, nthreads=cores)
        dt = time.perf_counter() - t0
        out[key] = (rays / 1e6 / dt, spp, dt)
    v, spp, dt = out["list"]
    return {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": "%s at %d spp of %d, full resolution, %.1f s, list mode (the reference's live HitableList path)" % (workload_name(workload, spp_full), spp, spp_full, dt),
            "soa_avx2_mrays_s": out["soa_avx2"][0],
            "soa_avx2_sample": "%d spp, %.1f s (spheres_soa.rs hit_avx2: bench-only in the reference)" % (out["soa_avx2"][1], out["soa_avx2"][2])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS),
                    help="default cfg4 = the north-star target (random_spheres 3840x2160, 4096 spp, depth 50) at every N")
    ap.add_argument("--partition", default="rows", choices=["rows", "samples", "samples-strong"],
                    help="N>1: rows = ONE frame split by interleaved row tiles, no collective (strong scaling, the default); samples = every "
                         "rank renders the workload's full spp with its own frame seed + one NCCL reduce (weak); samples-strong = the "
                         "workload's spp split into N frame seeds of spp/N + one reduce (strong)")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (a reduced run is labelled as such)")
    ap.add_argument("--ref-spp", type=int, default=1, help="--impl reference: spp of the bounded CPU sample per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of the other BASELINE configs (N=1 only)")
    ap.add_argument("--fast", action="store_true", help="the e2e leg runs one step without its own warm-up (automatic when a step exceeds 2 s)")
    ap.add_argument("--resident-kernel", type=int, default=0, help="PtOptions.resident_kernel (0 = the library's default)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import pathtrace_rs_b200 as pt
    from pathtrace_rs_b200 import parallel

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        # host-side rendezvous for the e2e leg: an NCCL barrier parks a spinning kernel on every waiting rank's GPU, and rank 0
        # drives ALL GPUs through one library call there — the waiting ranks must leave their GPUs idle
        cpu_group = dist.new_group(backend="gloo")

    preset_name, w, h, spp, depth = WORKLOADS[args.workload]
    reduced = args.spp > 0 and args.spp != spp
    if args.spp > 0:
        spp = args.spp
    full_spp = spp
    samples_strong = world > 1 and args.partition == "samples-strong"
    if samples_strong:
        if spp % world:
            raise SystemExit("--partition samples-strong needs spp divisible by the number of GPUs")
        spp //= world  # each rank renders spp/N samples of every pixel with its own frame seed; one reduce forms the mean
    params = pt.Params(w, h, spp, depth)
    options = pt.PtOptions(resident_kernel=args.resident_kernel) if args.resident_kernel else None
    preset = pt.Preset(preset_name, params).create_scene(local_rank, options)
    n_spheres = len(preset)
    info = pt.device_info(local_rank)

    d_rgb = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
    d_rays = torch.zeros(1, dtype=torch.int64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()
    rows_mode = world > 1 and args.partition == "rows"
    part = parallel.partition_for(rank, world) if rows_mode else None

    def step_device(frame):
        """one Scene::update, image resident in HBM; returns nothing (async on torch's stream)"""
        if world > 1 and not rows_mode:
            d_rgb.zero_()                      # sample slices: every rank renders its frame seed into an empty buffer ...
        preset.update_device(params, frame, d_rgb.data_ptr(), d_rays.data_ptr(), stream.cuda_stream, part)
        if world > 1 and not rows_mode:
            parallel.combine_sample_slices(d_rgb, frame, world, dist)  # ... and ONE NCCL reduce forms the equal-weight mean

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_run(step_fn, n_warm, n_steps):
        for i in range(n_warm):
            step_fn(i)
            flush.fill_(float(i))
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        rays = 0
        t_wall = time.perf_counter()
        for i in range(n_steps):
            evs[i][0].record(stream)
            r = step_fn(i)
            evs[i][1].record(stream)
            if r is None:
                evs[i][1].synchronize()
                r = int(d_rays.item())
            rays += r
            flush.fill_(float(i))  # L2 flush between timed iterations (outside the event pair)
        barrier()
        wall = time.perf_counter() - t_wall
        ms = sum(a.elapsed_time(b) for a, b in evs)
        return ms, rays, wall

    # ---- device-resident throughput (value) -------------------------------------------------------------------
    frame_of = (lambda i: rank) if (world > 1 and not rows_mode) else (lambda i: 0)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, rays, wall = timed_run(lambda i: step_device(frame_of(i)), args.warmup, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    step_ms = ms / max(1, args.steps)
    long_steps = step_ms > 2000.0 or args.fast

    # kernel time for the roofline.  With row tiles (or one GPU) a step IS one megakernel launch between the two events, so
    # the value leg's own events are the measurement; only the sample-slice modes (zero fill + NCCL reduce inside the step)
    # time the bare launch separately.
    if world > 1 and not rows_mode:
        k_steps = 1 if long_steps else max(1, min(args.steps, 3))
        k_ms, k_rays, _ = timed_run(lambda i: preset.update_device(params, frame_of(i), d_rgb.data_ptr(), d_rays.data_ptr(), stream.cuda_stream, part),
                                    0 if long_steps else 1, k_steps)
    else:
        k_steps, k_ms, k_rays = args.steps, ms, rays

    # ---- end to end through the reference-facing call, host buffers ----------------------------------------------
    # ONE call per step — `scene.update(&params, &camera, frame_num, &mut buffer)` (src/offline.rs:29) -> pt_render — whatever
    # the number of GPUs: at N > 1 rank 0 owns a multi-device scene (pt_scene_create_multi over all N devices) and the library
    # fans the call out to one host thread per GPU, each copying its rows from / into the caller's buffer; the other ranks
    # wait at the barrier.  frame_num >= 1, so the previous accumulation is uploaded and the result downloaded every step.
    # The buffer is registered with pt_host_register (the contract asks for pinned memory; a Rust caller's pageable Vec is
    # measured beside it in `extras` at cfg2's size).
    e_steps = 1 if long_steps else args.steps
    host_img = np.zeros((h, w, 3), np.float32)
    e_ms = e_rays = e_wall = 0.0
    multi = None
    if rank == 0:
        L = pt.libptgpu()
        import ctypes as C
        pt.ffi.check(L.pt_host_register(host_img.ctypes.data_as(C.c_void_p), host_img.nbytes))
        if world > 1:
            multi = pt.Preset(preset_name, pt.Params(w, h, full_spp, depth)).create_scene(list(range(world)), options)
        e2e_scene, e2e_params = (multi, pt.Params(w, h, full_spp, depth)) if world > 1 else (preset, params)
    def cpu_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    barrier()
    cpu_barrier()
    if rank == 0:
        if not long_steps:
            e2e_scene.update(e2e_params, frame_num=1, buffer=host_img)
        t0 = time.perf_counter()
        for i in range(e_steps):
            _, r = e2e_scene.update(e2e_params, frame_num=1 + i, buffer=host_img)
            e_rays += r
        e_wall = time.perf_counter() - t0
        per_gpu_kernel_ms = [st.kernel_ms for st in e2e_scene.device_stats()]
        pt.ffi.check(L.pt_host_unregister(host_img.ctypes.data_as(C.c_void_p)))
    cpu_barrier()
    h2d_b = w * h * 12
    d2h_b = w * h * 12 + 8 * world

    # ---- reduce over ranks: time = max, work = sum ------------------------------------------------------------------
    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ms_max, rays_sum = allmax(ms), allsum(rays)
    k_ms_max, k_rays_sum = allmax(k_ms), allsum(k_rays)
    ms_min = -allmax(-ms)
    samples_per_step = w * h * spp * (world if (world > 1 and not rows_mode) else 1)

    if rank == 0:
        value = rays_sum / 1e6 / (ms_max * 1e-3)
        peak_nominal = info.fp32_fma_peak_flops * world
        achieved = k_rays_sum * FLOP_PER_TEST * n_spheres / (k_ms_max * 1e-3)
        try:
            peak_probe = pt.probe_fp32_peak(local_rank)
        except Exception:
            peak_probe = None
        traffic, traffic_note = None, "no ncu capture of this workload at this size is committed (profiles/ncu_traffic.json)"
        try:
            table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            t = table.get(args.workload) if not reduced else table.get("%s@%d" % (args.workload, full_spp))
            if t is None and not reduced:
                near = sorted(k for k in table if k.startswith(args.workload + "@"))
                if near:
                    traffic_note = "no capture of the full-size launch; at reduced spp (%s): %s" % (near[0], table[near[0]]["note"])
            if t:
                traffic = (t["dram_read_bytes"] + t["dram_write_bytes"]) * (1 if rows_mode or world == 1 else world)
                traffic_note = "dram bytes read+written per launch, ncu --set full: %s; %s" % (t["source"], t["note"])
        except (OSError, ValueError, KeyError):
            pass
        st = preset.stats()
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            peaks = {}
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.partition in ("rows", "samples-strong") else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, full_spp) + (" [REDUCED spp]" if reduced else ""),
                       "n_spheres": n_spheres, "partition": ("rows: interleaved 4-row tiles, no collective" if rows_mode else
                                                             (("samples-strong: %d spp per GPU x %d frame seeds + one NCCL reduce" % (spp, world)) if samples_strong else
                                                              ("samples: one frame seed per GPU + one NCCL reduce" if world > 1 else "single GPU"))),
                       "l2": "flushed between timed iterations (256 MB fill); the scene is a %d KB pre-filter image %s" % (max(1, n_spheres * (32 if st.resident in (2, 3) else 16) // 1024), "resident in shared memory" if st.resident else "streamed from L2 in TMA tiles"),
                       "prefilter": ("tensor path: the two dot products of every (ray, sphere) test as mma.sync.m16n8k16 f16 split-operand MMAs (HMMA.16816.F32), "
                                     "A'^2 + B' and the sign test in packed FP32; the exact f32 test of the flagged spheres decides every hit" if st.resident in (2, 3) else
                                     "packed FP32 (FFMA2), 7 instructions per 2 tests; the exact f32 test of the flagged spheres decides every hit"),
                       "kernel": "%d CTAs x %d threads, %d B shared memory per CTA" % (st.grid_ctas, st.cta_threads, st.smem_bytes),
                       "timing": "CUDA events per step on the launching stream, max over ranks (slowest rank %.1f ms, fastest %.1f ms per step)" % (ms_max / args.steps, ms_min / args.steps)},
            "samples_per_s": samples_per_step * args.steps / (ms_max * 1e-3),
            "rays_per_sample": rays_sum / (samples_per_step * args.steps),
            "e2e": {"value": e_rays / 1e6 / e_wall, "unit": "Mrays/s", "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
                    "ms_per_step": 1e3 * e_wall / e_steps, "steps": e_steps,
                    "path": ("Scene::update mirror -> ONE pt_render call on a %d-device scene (pt_scene_create_multi: host thread per GPU, interleaved row tiles, "
                             "each GPU copies its rows from/into the caller's buffer)" % world) if world > 1 else
                            "Scene::update mirror -> pt_render (host buffer up and down every step, frame_num >= 1)",
                    "host_buffer": "numpy array registered with pt_host_register (pinned)",
                    "per_gpu_kernel_ms": per_gpu_kernel_ms},
            "gpu_launches": args.steps * world,
            "roofline": {"bound": "fp32_fma", "achieved": achieved / 1e12, "peak": peak_nominal / 1e12, "unit": "TFLOP/s",
                         "frac": achieved / peak_nominal, "traffic": traffic, "traffic_unit": "bytes", "traffic_note": traffic_note,
                         "peak_source": "sm_count*128*2*max SM clock (%d SMs, %d MHz); MEASURED_PEAKS.json has no FP32 figure — "
                                        "a pure-FFMA probe kernel measured %s TFLOP/s on this GPU in this run"
                                        % (info.sm_count, info.sm_clock_khz // 1000, ("%.1f" % (peak_probe / 1e12)) if peak_probe else "n/a"),
                         "algorithmic": "16 flop x %d spheres x %d rays per launch (brute force, every ray tests every sphere)" % (n_spheres, int(k_rays_sum / k_steps / world)),
                         "kernel_ms": k_ms_max / k_steps,
                         "tensor_work": ({"executed": achieved / FLOP_PER_TEST * 64.0 / 1e12 * ((n_spheres + 15) // 16 * 16) / max(1, n_spheres), "unit": "TFLOP/s",
                                          "what": "HMMA.16816 work of the pre-filter: 2 dot products x 16-deep f16 products x 2 flop per (ray, sphere) test, padded sphere count",
                                          "mma_sync_ceiling": 554.0 * world, "mma_sync_ceiling_source": "tools/probe_mma.cu on B200: 8.6 clk per m16n8k16 per SM sub-partition (profiles/probe_mma_r2a.txt)",
                                          "dense_bf16_peak": (peaks.get("bf16_tflops") or 0.0) * world, "dense_bf16_peak_source": "MEASURED_PEAKS.json bf16_tflops (tcgen05 path, cuBLAS)"}
                                         if st.resident in (2, 3) else None),
                         "note": ("algorithmic flop (the FP32 formulation's 16 per test) against the FP32 FMA peak, as in round 1, so the two builds compare; "
                                  "in this build 12 of the 16 (the two 3-term dot products) execute on the tensor pipe as 2 x 16-deep f16 products = 64 flop per test: %.0f TFLOP/s of HMMA work "
                                  "(mma.sync ceiling measured on B200: 550 TFLOP/s with a register accumulator, tools/probe_mma.cu); the kernel is bound by instruction issue around the MMAs, "
                                  "not by HBM: algorithmic HBM bytes are 12-24 B/pixel/launch (accumulation buffer)" % (achieved / FLOP_PER_TEST * 64.0 / 1e12)) if st.resident in (2, 3) else
                                 "bound by the FP32 FMA pipe, not by HBM: algorithmic HBM bytes are 12-24 B/pixel/launch (accumulation buffer)"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_extras:
            try:
                line["extras"] = extras(pt, np, torch, local_rank, options)
            except Exception as e:
                line["extras"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args.workload)
            except Exception as e:  # the baseline is a reported extra; never lose the GPU line over it
                line["cpu_baseline"] = {"error": repr(e)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def extras(pt, np, torch, device, options):
    """Short driver-visible figures for the other BASELINE configs (one warm-up + one timed `pt_render` each, kernel time from
    the library's own CUDA events): cfg1 and cfg2 and cfg3 at FULL size, cfg5 at full resolution but 8 spp (labelled)."""
    out = {}
    info = pt.device_info(device)
    for key, spp_override in (("cfg1", 0), ("cfg2", 0), ("cfg3", 0), ("cfg5", 8)):
        preset_name, w, h, spp, depth = WORKLOADS[key]
        full = spp
        if spp_override:
            spp = spp_override
        params = pt.Params(w, h, spp, depth)
        pr = pt.Preset(preset_name, params).create_scene(device, options)
        buf = np.zeros((h, w, 3), np.float32)
        pr.update(params, frame_num=0, buffer=buf)
        t0 = time.perf_counter()
        _, rays = pr.update(params, frame_num=1, buffer=buf)
        wall = time.perf_counter() - t0
        st = pr.stats()
        n = len(pr)
        rec = {"workload": workload_name(key, full) + ("" if spp == full else " [REDUCED to %d spp]" % spp), "kernel_ms": st.kernel_ms,
               "mrays_s": rays / 1e3 / st.kernel_ms, "samples_per_s": w * h * spp / (st.kernel_ms * 1e-3),
               "e2e_mrays_s_pageable_host_buffer": rays / 1e6 / wall, "rays_per_sample": rays / (w * h * spp),
               "fp32_frac": rays * FLOP_PER_TEST * n / (st.kernel_ms * 1e-3) / info.fp32_fma_peak_flops, "n_spheres": n,
               "lane_efficiency_of_the_sweep": rays / 32.0 / max(1, st.warp_sweeps),
               "kernel": {0: "streamed, packed-FP32 pre-filter", 1: "resident, packed-FP32 pre-filter", 2: "resident, tensor-path pre-filter",
                          3: "streamed, tensor-path pre-filter"}.get(int(st.resident), "?")}
        if key == "cfg3":
            # every scatter evaluates the Noise texture once (both spheres are noise-textured Lambertians, presets.rs:271-315) and a
            # sample ends by exactly one miss or one depth-limit hit, so texture lookups = rays - samples; turb = 7 octaves of noise
            turb = rays - w * h * spp
            rec["noise_evals_per_s"] = 7.0 * turb / (st.kernel_ms * 1e-3)
            rec["turb_evals_per_s"] = turb / (st.kernel_ms * 1e-3)
            rec["note"] = ("N = 2 spheres: the FP32-FMA roofline of the sweep does not bound this config (SURVEY §8d); it is bound by instruction issue in "
                           "Perlin turb (7 octaves x 8 lattice corners of gathers + Hermite/trilinear arithmetic, perlin.rs:54-111)")
        out[key] = rec
        del pr
    return out


if __name__ == "__main__":
    sys.exit(main())
