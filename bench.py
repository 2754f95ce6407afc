#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native path-tracing core.

Workload (BASELINE.json configs[1]): cornell-dielectric 1280x720, plain path tracing + NEE/MIS, power heuristic,
maxDepth 30, 16 spp per frame (the glass shell is the documented stand-in of
scenes/make_standins.py because the reference's shell.obj is a missing blob).  One "step" = one frame = 16 spp over the
whole image.  Metric: Mrays/s = (closest-hit extend rays + MIS probe rays + shadow rays) per second, whole job.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun, one rank per GPU: ranks render disjoint frames (frame_seed = tea(step*N + rank, seed)),
the accumulation image is all-reduced over NCCL inside the timed region ("weak" scaling: per-GPU work is fixed).
--impl reference times the CPU restatement of the reference's tracer (oracle/, all host threads) on a bounded sample.
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
WIDTH, HEIGHT, SPP = 1280, 720, int(os.environ.get("B200PT_BENCH_SPP", "16"))   # the override is a tuning aid (tail analysis), not a bench mode
SCENE = os.path.join(ROOT, "scenes", "cornell-dielectric", "cornell-dielectric.xml")
SEED = 0xC0FFEE
BYTES_PER_RAY = 152          # algorithmic wavefront-state bytes per extend/shadow ray
WORKLOAD = "cornell-dielectric 1280x720 NEE+MIS (power heuristic), maxDepth 30, 16 spp per step, stand-in shell.obj"
METRIC, UNIT = "Mrays/s (extend+shadow) cornell-dielectric 1280x720 NEE+MIS", "Mrays/s"


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def push_constants(P, frame_seed, previous_frames, spp=SPP):
    return P.default_push_constants(randomUInt=frame_seed, previousFrames=previous_frames, samplesPerPixel=spp, enableNEE=1,
                                    enableMIS=1, usePowerHeuristic=1, numNEE=1, maxDepth=30, maxFollowDiscrete=3)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_run(P, width, height, spp, frames, threads):
    """The oracle (CPU port of the reference's shader megakernel) on a bounded sample of the workload: a centred
    width x height crop of the 1280x720 frame is not available through the camera model, so the sample is the full
    view rendered at width x height (same scene, camera, push constants), `frames` frames of `spp` spp."""
    O = _load("b200pt_oracle", os.path.join(ROOT, "oracle", "oracle.py"))
    scene = P.Scene(SCENE)
    view, proj = scene.camera_matrices(WIDTH / HEIGHT)
    o = O.TracerOracle(width, height, 0, accel=True)
    o.set_scene(scene.desc)
    o.set_camera(view, proj, P.mat4_inverse(view), P.mat4_inverse(proj))
    times, rays = [], []
    for f in range(frames):
        o.reset_counters()
        pc = push_constants(P, P.tea(f, SEED), 0, spp)
        t0 = time.perf_counter()
        o.render_region(pc, threads=threads)
        times.append(time.perf_counter() - t0)
        c = o.counters()
        rays.append(c["extend_rays"] + c["shadow_rays"])
    return times, rays


EM_SPLITS, EM_PER_REGION = 8, 57600     # BASELINE configs[0]: 2^8 regions x (1280*720*16 / 256) records = the full sample buffer
EM_BYTES_PER_SAMPLE_ITER = 16           # direction + weight per sample per EM iteration


def em_batches(P, aabbs):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import guiding_data                      # synthetic sample generator (numpy only; not part of oracle/)
    return [guiding_data.make_batch(aabbs, EM_PER_REGION, seed, invalid_fraction=0.0) for seed in (0x5EED0001, 0x5EED0002)]


def em_cpu_reference(P, batches, threads):
    """The reference's own lightpmm + guiding headers (oracle/_ref, compiled from /root/reference in the build container)
    driven by the restated PathGuiding::update: first update = fit, second = updateFit; all regions on `threads` cores."""
    O = _load("b200pt_oracle", os.path.join(ROOT, "oracle", "oracle.py"))
    if not O.ref_guiding_available():
        return None
    scene = P.Scene(SCENE)
    g = O.GuidingRef(EM_SPLITS, [float(x) for x in scene.desc.scene_min[:]], [float(x) for x in scene.desc.scene_max[:]], P.default_guiding_params())
    times = []
    for b in batches:
        t0 = time.perf_counter()
        g.update(b, threads=threads)
        times.append(time.perf_counter() - t0)
    n = sum(len(b) for b in batches)
    return {"value": n / sum(times), "unit": "samples/s", "cores": threads, "kind": "reference",
            "sample": "full workload: 2 updates (fit, updateFit) of %d records each, lightpmm SSE build, region loop on %d threads" % (len(batches[0]), threads),
            "seconds": times, "em_sample_iterations": g.em_sample_iterations()}


def em_gpu(P, r, batches, torch, device):
    """Guiding update on the GPU: `value` with the records resident in HBM, `e2e` from pinned host memory through
    b200pt_guiding_update_host (H2D copy inside the timed region).  CUDA events on the library's stream."""
    gp = P.default_guiding_params()
    dev = [torch.from_numpy(b.view("u1").reshape(len(b), 40)).to(device) for b in batches]
    pinned = [torch.from_numpy(b.view("u1").reshape(len(b), 40)).pin_memory() for b in batches]
    res = {}
    for mode in ("warmup", "device", "host"):
        r.guiding_reset(gp)
        r.stats_reset()
        ms = []
        for i in range(len(batches)):
            r.timer_start()
            if mode == "host":
                P._check(P.lib().b200pt_guiding_update_host(r._h, gp, pinned[i].data_ptr(), len(batches[i])))
            else:
                r.guiding_update_device(dev[i].data_ptr(), len(batches[i]), gp)
            ms.append(r.timer_stop())
        st = r.stats()
        res[mode] = {"ms": ms, "ms_sort": st.ms_guiding_sort, "ms_fit": st.ms_guiding_fit, "sample_iters": int(st.guiding_em_sample_iterations),
                     "samples": int(st.guiding_samples), "launches": int(st.launches_guiding)}
    return res


def run_reference(args, rank, world):
    """Reference arm: the reference's tracer has no CPU implementation (GLSL + RT cores), so this times the CPU port
    (oracle/) with every host thread, each step a bounded sample of the workload."""
    if rank != 0:
        return
    P = _load("b200pt_binding", os.path.join(ROOT, "rtx-pathtracer_b200", "b200pt.py"))
    threads = os.cpu_count() or 1
    w, h, spp = 320, 180, 4
    times, rays = cpu_oracle_run(P, w, h, spp, args.warmup + args.steps, threads)
    t = sum(times[args.warmup:])
    r = sum(rays[args.warmup:])
    value = r / t / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "spp_per_step": SPP},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%dx%d view of the same scene/camera, %d spp per step, oracle/tracer_oracle.cpp with its own BVH" % (w, h, spp)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if world == 1 and not args.no_em:
        # second headline: the reference's CPU guiding fit (its own lightpmm code), all host threads
        O = _load("b200pt_oracle", os.path.join(ROOT, "oracle", "oracle.py"))
        if O.ref_guiding_available():
            scene = P.Scene(SCENE)
            g = O.GuidingRef(EM_SPLITS, [float(x) for x in scene.desc.scene_min[:]], [float(x) for x in scene.desc.scene_max[:]], P.default_guiding_params())
            em = em_cpu_reference(P, em_batches(P, g.aabbs()), threads)
            line["em"] = {"metric": "guiding EM samples/s, 256 regions x 57600 records, fit + updateFit", "value": em["value"], "unit": "samples/s", "cpu_baseline": em}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-em", action="store_true", help="skip the guiding-EM leg (second headline metric)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    P = _load("b200pt_binding", os.path.join(ROOT, "rtx-pathtracer_b200", "b200pt.py"))
    S = _load("b200pt_sharding", os.path.join(ROOT, "rtx-pathtracer_b200", "sharding.py"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene = P.Scene(SCENE)
    view, proj = scene.camera_matrices(WIDTH / HEIGHT)
    r = P.Renderer(WIDTH, HEIGHT, 0, 0, device=local_rank)
    r.set_scene(scene)
    r.set_camera(view, proj)
    accum = torch.zeros((HEIGHT, WIDTH, 4), dtype=torch.float32, device="cuda:%d" % local_rank)
    host_img = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory()

    def barrier():
        r.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        # device-resident step: frame i of this rank; running mean over this rank's frames
        r.render_frame(push_constants(P, S.frame_seed(i, rank, world, SEED), i))

    for i in range(args.warmup):
        step(i)
    # one more untimed pass through the call the timed region makes, so that the library's one-time work for a batch
    # of K frames (second half of the per-pixel sums, the pool of CUDA events) is not inside the timed region
    r.set_stage_timing(2)            # CUDA events around the trace kernel only (every stage kernel: ~10 % slower frames)
    r.render_frames([push_constants(P, S.frame_seed(i, rank, world, SEED), i) for i in range(args.steps)])
    if world > 1:   # warm the collective
        r.read_image_device(P.IMAGE_OUTPUT, accum.data_ptr())
        S.combine_images(accum, args.warmup)

    # ---- timed region 1: `value` (inputs resident, K frames + the image reduction) --------------------------------
    barrier()
    r.stats_reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    r.timer_start()                  # CUDA event on the library's stream (the one every kernel is launched on)
    # the K frames go to the library in ONE call (b200pt_render_frames): same images as K calls of render_frame, but a
    # pixel that has finished frame i starts frame i + 1 without waiting for the frame's slowest pixels
    r.render_frames([push_constants(P, S.frame_seed(i, rank, world, SEED), i) for i in range(args.steps)])
    r.read_image_device(P.IMAGE_OUTPUT, accum.data_ptr())
    dev_ms = r.timer_stop()
    final = S.combine_images(accum, args.steps)      # N > 1: one NCCL all-reduce of the image (+ a scalar); N = 1: no-op
    barrier()
    wall = time.perf_counter() - t0
    # single GPU: device time between the two events; multi GPU: the collective runs on torch's stream, so the region
    # is closed by the barrier and timed by the host clock around it (max over ranks below)
    elapsed = dev_ms * 1e-3 if world == 1 else wall
    clocks = sampler.summary()
    st = r.stats()
    rays = int(st.extend_rays) + int(st.shadow_rays)
    stats = {k: getattr(st, k) for k in ("extend_rays", "shadow_rays", "iterations", "kernel_launches", "launches_extend", "launches_shadow",
                                         "launches_shade", "ms_extend", "ms_shadow", "ms_shade", "ms_total")}

    # ---- timed region 2: `e2e` through the C ABI with HOST buffers (camera + push constants in, image out) ---------
    r.set_stage_timing(0)
    barrier()
    r.stats_reset()
    t1 = time.perf_counter()
    r.timer_start()
    for i in range(args.steps):
        r.set_camera(view, proj)                                  # host -> device: 2 x mat4 (the reference's UBO update)
        r.render_frame(push_constants(P, S.frame_seed(i, rank, world, SEED), i))   # 192 B of push constants
        P._check(P.lib().b200pt_read_image(r._h, P.IMAGE_OUTPUT, host_img.data_ptr()))   # device -> pinned host, 16 B/px
    e2e_dev_ms = r.timer_stop()
    barrier()
    e2e_elapsed = e2e_dev_ms * 1e-3 if world == 1 else time.perf_counter() - t1
    st2 = r.stats()
    e2e_rays = int(st2.extend_rays) + int(st2.shadow_rays)

    # ---- untimed: two frames with CUDA events around EVERY stage kernel -> the trace / shade split of a frame ---------
    r.set_stage_timing(1)
    r.stats_reset()
    for i in range(2):
        step(i)
    st3 = r.stats()
    split = {"trace": st3.ms_extend, "shade": st3.ms_shade, "frame_total": st3.ms_total, "frames": 2,
             "note": "separate pass of 2 single frames with events around every kernel (slows a frame by ~10 %); shares, not absolutes"}
    r.set_stage_timing(0)

    # max over ranks, totals over ranks
    if world > 1:
        tt = torch.tensor([elapsed, e2e_elapsed], dtype=torch.float64, device="cuda:%d" % local_rank)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed, e2e_elapsed = float(tt[0]), float(tt[1])
        rr = torch.tensor([rays, e2e_rays, int(st.kernel_launches)], dtype=torch.float64, device="cuda:%d" % local_rank)
        dist.all_reduce(rr)
        total_rays, total_e2e_rays, launches = float(rr[0]), float(rr[1]), int(rr[2])
    else:
        total_rays, total_e2e_rays, launches = float(rays), float(e2e_rays), int(st.kernel_launches)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # dominant kernel = k_trace (persistent BVH8 traversal of the path, probe and shadow queues): algorithmic bytes
        # per launch = rays per launch x 152 B; duration = CUDA events around every launch on the launching stream
        ext_launches = max(1, stats["launches_extend"])
        ext_ms = stats["ms_extend"] / ext_launches
        ext_bytes = (stats["extend_rays"] + stats["shadow_rays"]) / ext_launches * BYTES_PER_RAY
        achieved = ext_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        # DRAM traffic of one k_trace launch from the committed ncu --set full capture (profiles/), if present
        traffic = None
        try:
            rd = wr = None
            for l in open(os.path.join(ROOT, "profiles", "r01e_ncu_trace.txt")):
                f = l.split()
                if l.startswith("dram__bytes_read.sum") and rd is None:
                    rd = float(f[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[1]]
                if l.startswith("dram__bytes_write.sum") and wr is None:
                    wr = float(f[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[1]]
            if rd is not None and wr is not None:
                traffic = rd + wr
        except Exception:
            pass
        value = total_rays / elapsed / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "spp_per_step": SPP, "frames_per_rank": args.steps, "parallelism": "spp-sharded x%d" % world,
                       "frames": "value: the K frames of a rank in one b200pt_render_frames call (pixels walk from frame to frame; images identical to K single calls, tests/test_frame_batch_gpu.py); e2e: one b200pt_render_frame + image read-back per step",
                       "l2": "no flush: the wavefront queues touched per iteration (~230 MB at 921600 paths) exceed the 126 MB L2"},
            "spp_per_s": SPP * args.steps * world / elapsed,
            "e2e": {"value": total_e2e_rays / e2e_elapsed / 1e6, "unit": UNIT, "h2d_bytes_per_step": 128 + 192, "d2h_bytes_per_step": WIDTH * HEIGHT * 16},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_trace (persistent BVH8 traversal: closest hit + any hit)", "launches": stats["launches_extend"], "avg_launch_us": 1e3 * ext_ms,
                         "trace_Mrays_per_s": (stats["extend_rays"] + stats["shadow_rays"]) / max(stats["ms_extend"], 1e-9) / 1e3, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "algorithmic_bytes_per_launch": ext_bytes, "peak_source": peak_src,
                         "note": "152 algorithmic bytes/ray of wavefront state (SURVEY 8d); the kernel is latency/issue bound, see profiles/"},
            "stage_ms": {"trace": stats["ms_extend"], "frame_total": stats["ms_total"], "device": dev_ms, "wall": 1e3 * wall},
            "stage_split": split,
            "rays": {"extend": stats["extend_rays"], "shadow": stats["shadow_rays"], "iterations": stats["iterations"]},
        }
        if world == 1 and not args.no_em:
            # ---- second headline metric: guiding EM samples/s (BASELINE configs[0]) ------------------------------------
            rg = P.Renderer(64, 64, 0, EM_SPLITS, device=local_rank)
            rg.set_scene(scene)
            batches = em_batches(P, rg.guiding_aabbs())
            g = em_gpu(P, rg, batches, torch, "cuda:%d" % local_rank)
            n = sum(len(b) for b in batches)
            dev_s, host_s = sum(g["device"]["ms"]) * 1e-3, sum(g["host"]["ms"]) * 1e-3
            fit_s = g["device"]["ms_fit"] * 1e-3
            em = {"metric": "guiding EM samples/s: PathGuiding::update (sort + preFit + fit|updateFit + merge/split + statistics + pack), "
                            "256 regions x 57600 records, 2 updates",
                  "value": n / dev_s, "unit": "samples/s", "ms_per_update": g["device"]["ms"], "ms_sort_total": g["device"]["ms_sort"], "ms_fit_total": g["device"]["ms_fit"],
                  "em_sample_iterations": g["device"]["sample_iters"], "gpu_launches": g["device"]["launches"],
                  "e2e": {"value": n / host_s, "unit": "samples/s", "h2d_bytes_per_step": len(batches[0]) * 40, "d2h_bytes_per_step": 0, "ms_per_update": g["host"]["ms"]},
                  "roofline": {"bound": "hbm", "kernel": "k_guiding_update (one block per region)", "achieved": g["device"]["sample_iters"] * EM_BYTES_PER_SAMPLE_ITER / fit_s / 1e9,
                               "peak": hbm_peak, "unit": "GB/s", "frac": g["device"]["sample_iters"] * EM_BYTES_PER_SAMPLE_ITER / fit_s / 1e9 / hbm_peak, "traffic": None,
                               "sample_iterations_per_s": g["device"]["sample_iters"] / fit_s,
                               "note": "16 B per sample per EM iteration of fit/updateFit (masked post-split fits and the statistics passes are extra work "
                                       "not counted here); the kernel is bounded by per-region serial depth, not HBM — see DESIGN.md"}}
            if not args.no_cpu_baseline:
                cpu = em_cpu_reference(P, batches, os.cpu_count() or 1)
                if cpu:
                    em["cpu_baseline"] = cpu
                    em["speedup_vs_cpu_all_cores"] = em["value"] / cpu["value"]
            line["em"] = em
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            w, h, spp = 640, 360, 8          # ~10 s of CPU work on 16 threads
            times, crays = cpu_oracle_run(P, w, h, spp, 2, threads)
            line["cpu_baseline"] = {"value": crays[-1] / times[-1] / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%dx%d view of the same scene/camera, %d spp, 1 warm-up + 1 timed frame of oracle/tracer_oracle.cpp (own BVH)" % (w, h, spp)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
