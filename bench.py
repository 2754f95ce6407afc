#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native path-tracing core.

Workload (BASELINE.json configs[1]): cornell-dielectric 1280x720, plain path tracing + NEE/MIS, power heuristic,
maxDepth 30, 16 spp per frame (the glass shell is the documented stand-in of
scenes/make_standins.py because the reference's shell.obj is a missing blob).  One "step" = one frame = 16 spp over the
whole image.  Metric: Mrays/s = (closest-hit extend rays + MIS probe rays + shadow rays) per second, whole job.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun, one rank per GPU: ranks render disjoint frames (frame_seed = tea(step*N + rank, seed)) and
the images are combined by the library's own b200pt_reduce_image (NCCL all-reduce on the library's stream) inside the
timed region ("weak" scaling: per-GPU work is fixed).  Every N is timed the same way: CUDA events on the stream all the
kernels and the collective run on, max over ranks.  torch.distributed only carries the communicator id and the final
max / sum of the per-rank numbers.  Extra objects of the JSON line: `strong` (1024 spp split over the N GPUs, time to the
reduced image), `config5` (BASELINE configs[4]: 4K guiding training with the region-sharded refit, per-phase times),
`em` (BASELINE configs[0], N = 1 only).
--impl reference times the reference's own shader source compiled as C++ (oracle/_ref/libshader_ref.so, one process per host
core; without that library the CPU restatement oracle/tracer_oracle.cpp) on the SAME view, resolution and spp, restricted to a
bounded sample of the image rows (every 8th block of 10 rows).
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


CPU_SAMPLE_ROWS = [(y0, y0 + 10) for y0 in range(0, HEIGHT, 80)]       # every 8th block of 10 rows: 90 of the 720 rows
CPU_SAMPLE = "same scene, camera, 1280x720 view and %d spp per step; rows %s of the image (1/8 of the pixels, spread over the frame)" % (
    SPP, ", ".join("%d-%d" % (a, b - 1) for a, b in CPU_SAMPLE_ROWS))


# what "port" means here: the reference has no CPU renderer (its path tracer is GLSL on RT cores); the port is oracle/tracer_oracle.cpp,
# whose frames are bit-equal to the reference's own shader source compiled as C++ and run on the CPU (tests/test_shader_ref.py)
PORT_PIN = "frames bit-equal to the reference's shader source compiled as C++ (oracle/shader_ref.cpp, tests/test_shader_ref.py)"


def cpu_oracle_run(P, frames, threads):
    """The oracle (CPU port of the reference's shader megakernel, oracle/tracer_oracle.cpp, own BVH) on a bounded sample
    of the workload: the configuration is the bench's (view, 1280x720, SPP spp per step, same push constants), the
    sample is a fixed subset of the image rows."""
    O = _load("b200pt_oracle", os.path.join(ROOT, "oracle", "oracle.py"))
    scene = P.Scene(SCENE)
    view, proj = scene.camera_matrices(WIDTH / HEIGHT)
    o = O.TracerOracle(WIDTH, HEIGHT, 0, accel=True)
    o.set_scene(scene.desc)
    o.set_camera(view, proj, P.mat4_inverse(view), P.mat4_inverse(proj))
    times, rays = [], []
    for f in range(frames):
        o.reset_counters()
        pc = push_constants(P, P.tea(f, SEED), 0, SPP)
        t0 = time.perf_counter()
        for y0, y1 in CPU_SAMPLE_ROWS:
            o.render_region(pc, 0, y0, WIDTH, y1, threads=threads)
        times.append(time.perf_counter() - t0)
        c = o.counters()
        rays.append(c["extend_rays"] + c["shadow_rays"])
    return times, rays


SHADER_REF = os.path.join(ROOT, "oracle", "_ref", "libshader_ref.so")
REFERENCE_KIND = ("the reference's own shader source (raytrace.rgen + hit / miss / intersection shaders) compiled as C++ and run on the CPU, one process per "
                  "host core (oracle/shader_ref.cpp); ray / primitive intersection, texture filtering and the elementary functions — what the reference "
                  "leaves to the driver — come from oracle/tracer_oracle.cpp")
_ref_worker = {}


def _ref_worker_init():
    """one per process: the compiled reference shaders keep their descriptor sets in globals, like a shader invocation does"""
    import ctypes as C
    import numpy as np
    P = _load("b200pt_binding", os.path.join(ROOT, "rtx-pathtracer_b200", "b200pt.py"))
    O = _load("b200pt_oracle", os.path.join(ROOT, "oracle", "oracle.py"))
    scene = P.Scene(SCENE)
    view, proj = scene.camera_matrices(WIDTH / HEIGHT)
    o = O.TracerOracle(8, 8, 0, accel=True)        # only its traversal and its sampler are used (callbacks): no frame-sized buffers needed
    o.set_scene(scene.desc)
    S = C.CDLL(SHADER_REF)
    S.shader_ref_rays_traced.restype = C.c_uint64
    S.shader_ref_init(WIDTH, HEIGHT, 0)
    S.shader_ref_set_scene(C.byref(scene.desc))
    f = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(C.POINTER(C.c_float))
    S.shader_ref_set_camera(f(view), f(proj), f(P.mat4_inverse(view)), f(P.mat4_inverse(proj)))
    L = O.lib()
    S.shader_ref_set_callbacks(o._h, C.cast(L.oracle_trace_one, C.c_void_p), C.cast(L.oracle_texture, C.c_void_p))
    _ref_worker.update(P=P, S=S, scene=scene, oracle=o, C=C)


def _ref_worker_rows(job):
    frame, rows = job
    P, S, C = _ref_worker["P"], _ref_worker["S"], _ref_worker["C"]
    pc = push_constants(P, P.tea(frame, SEED), 0, SPP)
    before = S.shader_ref_rays_traced()
    for y in rows:
        assert S.shader_ref_render(C.byref(pc), 0, y, WIDTH, y + 1) == 0
    return S.shader_ref_rays_traced() - before


def cpu_reference_shaders_run(frames, procs):
    """The reference's shader pipeline compiled as C++ (oracle/_ref/libshader_ref.so) on the bench's bounded sample: the rows of
    CPU_SAMPLE_ROWS dealt round-robin to `procs` processes; wall time per frame around the whole pool."""
    import multiprocessing as mp
    rows = [y for y0, y1 in CPU_SAMPLE_ROWS for y in range(y0, y1)]
    shares = [rows[k::procs] for k in range(procs)]
    times, rays = [], []
    with mp.get_context("fork").Pool(procs, initializer=_ref_worker_init) as pool:
        pool.map(_ref_worker_rows, [(0, [])] * procs)          # every worker is up before the clock starts
        for f in range(frames):
            t0 = time.perf_counter()
            counts = pool.map(_ref_worker_rows, [(f, sh) for sh in shares], chunksize=1)
            times.append(time.perf_counter() - t0)
            rays.append(sum(counts))
    return times, rays


def cpu_baseline_reference_shaders():
    """cpu_baseline of the GPU arm's line: the reference arm (1 warm-up + 2 timed steps of the bounded sample) in a process of its own —
    its workers are forked, and nothing is forked from a process that holds a CUDA context.  None when the library is absent or the run fails."""
    if not os.path.exists(SHADER_REF):
        return None
    try:
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1", "--no-em"],
                             capture_output=True, text=True, timeout=240, env=env)
        cpu = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        return cpu if cpu.get("kind") == "reference" and cpu.get("value", 0) > 0 else None
    except Exception:
        return None


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
    """Reference arm: the reference's tracer has no CPU implementation (GLSL + RT cores).  What can run on the host is its shader source
    compiled as C++ (oracle/_ref/libshader_ref.so, built from /root/reference in the build container): that is what is timed here, one
    process per host core, each step a bounded sample of the workload.  Without that library: the CPU port (oracle/tracer_oracle.cpp)."""
    if rank != 0:
        return
    P = _load("b200pt_binding", os.path.join(ROOT, "rtx-pathtracer_b200", "b200pt.py"))
    threads = os.cpu_count() or 1
    # the reference's own code when it compiled here (oracle/_ref travels with the repository), else the port
    use_shaders = os.path.exists(SHADER_REF)
    if use_shaders:
        times, rays = cpu_reference_shaders_run(args.warmup + args.steps, threads)
    else:
        times, rays = cpu_oracle_run(P, args.warmup + args.steps, threads)
    t = sum(times[args.warmup:])
    r = sum(rays[args.warmup:])
    value = r / t / 1e6
    baseline = ({"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": CPU_SAMPLE, "what": REFERENCE_KIND} if use_shaders else
                {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": CPU_SAMPLE, "port_checked_against": PORT_PIN})
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "spp_per_step": SPP},
            "cpu_baseline": baseline,
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


def config5_run(P, torch, dist, rank, world, local_rank, width=3840, height=2160, spp=2, guided_frames=2):
    """BASELINE configs[4]: 3840x2160, GUIDING_SPLITS = 8, six guiding optimisation frames through the frame driver (each
    followed by PathGuiding::update — across ranks the region-sharded refit b200pt_guiding_update_all_ranks), then guided
    frames.  Stand-in scene sponzaXML (fireplace_room.obj is a missing blob of the reference checkout).  Device times, max
    over ranks; rays and samples summed over ranks."""
    scene_name = "sponzaXML"
    scene = P.Scene(os.path.join(ROOT, "scenes", scene_name, scene_name + ".xml"))
    view, proj = scene.camera_matrices(width / height)
    r = P.Renderer(width, height, 0, 8, device=local_rank)
    r.set_scene(scene)
    r.set_camera(view, proj)
    dev = "cuda:%d" % local_rank
    if world > 1:
        ids = [P.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        r.comm_init(ids[0], rank, world)
    app = P.App(r, accumulate=True, samplesPerPixel=spp, enableNEE=1, enableMIS=1, updateGuiding=1, useParallaxCompensation=1)
    out = dict(workload="%s (stand-in for fireplace) %dx%d, GUIDING_SPLITS 8 (256 regions), %d spp per frame, spp-sharded x%d" % (scene_name, width, height, spp, world),
               n_gpus=world, phases=[])

    def red(vals, op):
        if world == 1:
            return vals
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return t.tolist()

    step = [0]

    def phase(label, n, reduce_image=False):
        r.stats_reset()
        r.synchronize()
        if world > 1:
            dist.barrier()
        r.timer_start()
        for _ in range(n):
            app.draw_frame(P.tea(step[0] * world + rank, SEED))
            step[0] += 1
        if reduce_image and world > 1:
            r.reduce_image(P.IMAGE_OUTPUT, n)
        total = r.timer_stop()
        st = r.stats()
        M = dist.ReduceOp.MAX if world > 1 else None
        S_ = dist.ReduceOp.SUM if world > 1 else None
        total, render, sort, xchg, fit, gather = red([total, st.ms_total, st.ms_guiding_sort, st.ms_guiding_exchange, st.ms_guiding_fit, st.ms_guiding_gather], M)
        rays, fitted, recv = red([float(st.extend_rays + st.shadow_rays), float(st.guiding_samples), float(st.guiding_bytes_received)], S_)
        ph = dict(phase=label, frames_per_rank=n, ms_total=round(total, 2), ms_render=round(render, 2), ms_guiding_sort=round(sort, 2), ms_guiding_exchange=round(xchg, 2),
                  ms_guiding_fit=round(fit, 2), ms_guiding_gather=round(gather, 2), Mrays_per_s=round(rays / max(total, 1e-9) / 1e3, 1),
                  spp_per_s=round(spp * n * world / max(total, 1e-9) * 1e3, 2))
        if st.guiding_samples_all_ranks:
            ph.update(guiding_samples_all_ranks=int(st.guiding_samples_all_ranks), guiding_samples_fitted_this_rank=int(st.guiding_samples),
                      guiding_samples_fitted_sum=int(fitted), exchange_bytes_all_ranks=int(recv),
                      guiding_samples_per_s=round(st.guiding_samples_all_ranks / max(sort + xchg + fit + gather, 1e-9) * 1e3, 0))
        out["phases"].append(ph)

    phase("training: 7 frames, 6 refits (updateGuiding)", app.state.numGuidingOptimizations + 1)
    app.settings.useGuiding = 1
    app.settings.guidingProb = 0.5
    app.input_changed()
    phase("guided render (useGuiding, guidingProb 0.5, parallax compensation)", guided_frames, reduce_image=True)
    vm = r.guiding_get_vmms()
    if world > 1:
        v = torch.from_numpy(vm.view("u1").copy()).to(dev)
        parts = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(parts, v)
        out["mixtures_identical_on_all_ranks"] = all(bool(torch.equal(parts[0], p)) for p in parts)
        out["exchange_mode"] = {0: "none", 1: "ncclSend/ncclRecv", 2: "CUDA-IPC peer reads over NVLink"}[r.comm_exchange_mode()]
    img = r.read_image()[..., :3]
    out.update(image_mean=float(img.mean()), finite=bool((img == img).all()), mean_components=float(vm["usedDistributions"].mean()))
    if world > 1:
        r.comm_destroy()
    del r
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-em", action="store_true", help="skip the guiding-EM leg (second headline metric)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (1024 spp split over the GPUs)")
    ap.add_argument("--no-config5", action="store_true", help="skip the 4K guiding-training leg (BASELINE configs[4])")
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
    dev = "cuda:%d" % local_rank
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene = P.Scene(SCENE)
    view, proj = scene.camera_matrices(WIDTH / HEIGHT)
    r = P.Renderer(WIDTH, HEIGHT, 0, 0, device=local_rank)
    r.set_scene(scene)
    r.set_camera(view, proj)
    host_img = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory()
    if world > 1:      # the product's own communicator (NCCL inside libb200pt.so); torch only hands the id around
        ids = [P.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        r.comm_init(ids[0], rank, world)

    def barrier():
        r.synchronize()
        if world > 1:
            dist.barrier()

    def over_ranks(vals, op):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return t.tolist()

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
        r.reduce_image(P.IMAGE_OUTPUT, args.steps)

    # ---- timed region 1: `value` (inputs resident, K frames + the image reduction) --------------------------------
    # Timed on the device for every N: CUDA events on the library's stream — the stream of every kernel AND of the NCCL
    # all-reduce inside b200pt_reduce_image — bracketed by a barrier on both sides; max over ranks.
    barrier()
    r.stats_reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    r.timer_start()
    # the K frames go to the library in ONE call (b200pt_render_frames): same images as K calls of render_frame, but a
    # pixel that has finished frame i starts frame i + 1 without waiting for the frame's slowest pixels
    r.render_frames([push_constants(P, S.frame_seed(i, rank, world, SEED), i) for i in range(args.steps)])
    if world > 1:
        r.reduce_image(P.IMAGE_OUTPUT, args.steps)       # mean over all frames of all ranks, on every rank
    dev_ms = r.timer_stop()
    barrier()
    wall = time.perf_counter() - t0
    elapsed = dev_ms * 1e-3
    clocks = sampler.summary()
    st = r.stats()
    rays = int(st.extend_rays) + int(st.shadow_rays)
    stats = {k: getattr(st, k) for k in ("extend_rays", "shadow_rays", "iterations", "kernel_launches", "launches_extend", "launches_shadow",
                                         "launches_shade", "ms_extend", "ms_shadow", "ms_shade", "ms_total")}

    # ---- timed region 2: `e2e` through the C ABI with HOST buffers (camera + push constants in, image out) ---------
    r.set_stage_timing(0)
    barrier()
    r.stats_reset()
    r.timer_start()
    for i in range(args.steps):
        r.set_camera(view, proj)                                  # host -> device: 2 x mat4 (the reference's UBO update)
        r.render_frame(push_constants(P, S.frame_seed(i, rank, world, SEED), i))   # 192 B of push constants
        P._check(P.lib().b200pt_read_image(r._h, P.IMAGE_OUTPUT, host_img.data_ptr()))   # device -> pinned host, 16 B/px
    if world > 1:
        r.reduce_image(P.IMAGE_OUTPUT, args.steps)
        P._check(P.lib().b200pt_read_image(r._h, P.IMAGE_OUTPUT, host_img.data_ptr()))
    e2e_elapsed = r.timer_stop() * 1e-3
    barrier()
    st2 = r.stats()
    e2e_rays = int(st2.extend_rays) + int(st2.shadow_rays)

    # ---- strong scaling: 1024 spp in total (64 frames of 16 spp) split over the N GPUs, time to the reduced image ------
    strong = None
    if not args.no_strong:
        total_frames = 1024 // SPP
        mine = [f for f in range(total_frames) if f % world == rank]
        pcs = [push_constants(P, P.tea(f, SEED), i) for i, f in enumerate(mine)]
        barrier()
        r.stats_reset()
        r.timer_start()
        r.render_frames(pcs)
        if world > 1:
            r.reduce_image(P.IMAGE_OUTPUT, len(mine))
        s_ms = r.timer_stop()
        barrier()
        st4 = r.stats()
        s_elapsed = over_ranks([s_ms], dist.ReduceOp.MAX if world > 1 else None)[0] * 1e-3
        s_rays = over_ranks([float(st4.extend_rays + st4.shadow_rays)], dist.ReduceOp.SUM if world > 1 else None)[0]
        strong = {"workload": "1024 spp of the same view in total: %d frames of %d spp dealt round-robin to the %d GPU(s), image all-reduced" % (total_frames, SPP, world),
                  "scaling": "strong", "seconds_to_image": s_elapsed, "spp_per_s": 1024 / s_elapsed, "Mrays_per_s": s_rays / s_elapsed / 1e6, "frames_this_rank": len(mine)}

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
    MAX = dist.ReduceOp.MAX if world > 1 else None
    SUM = dist.ReduceOp.SUM if world > 1 else None
    elapsed, e2e_elapsed = over_ranks([elapsed, e2e_elapsed], MAX)
    total_rays, total_e2e_rays, launches = over_ranks([rays, e2e_rays, int(st.kernel_launches)], SUM)
    launches = int(launches)
    if world > 1:
        r.comm_destroy()

    c5 = None
    if not args.no_config5:
        del r
        torch.cuda.empty_cache()
        c5 = config5_run(P, torch, dist, rank, world, local_rank)
        r = None

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
        # the kernel's own traffic: an extend ray reads 32 B and writes a 16 B hit; a shadow ray reads 32 + 16 B and adds 16 B
        trace_only_bytes = (stats["extend_rays"] * 48.0 + stats["shadow_rays"] * 64.0) / ext_launches
        # DRAM traffic of one k_trace launch from the committed ncu --set full capture (profiles/), if present
        traffic, traffic_src = None, None
        for name in ("r02b_ncu_trace.txt", "r02_ncu_trace.txt", "r01e_ncu_trace.txt"):
            try:
                rd = wr = None
                for l in open(os.path.join(ROOT, "profiles", name)):
                    f = l.split()
                    if l.startswith("dram__bytes_read.sum") and rd is None:
                        rd = float(f[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[1]]
                    if l.startswith("dram__bytes_write.sum") and wr is None:
                        wr = float(f[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[1]]
                if rd is not None and wr is not None:
                    traffic, traffic_src = rd + wr, "profiles/" + name
                    break
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
                       "multi_gpu": "b200pt_comm_init + b200pt_reduce_image (NCCL all-reduce on the library's stream) inside the timed region; torch.distributed only as launcher",
                       "timing": "CUDA events on the library's stream for every N, barrier on both sides, max over ranks",
                       "l2": "no flush: the wavefront queues touched per iteration (~230 MB at 921600 paths) exceed the 126 MB L2"},
            "spp_per_s": SPP * args.steps * world / elapsed,
            "e2e": {"value": total_e2e_rays / e2e_elapsed / 1e6, "unit": UNIT, "h2d_bytes_per_step": 128 + 192, "d2h_bytes_per_step": WIDTH * HEIGHT * 16},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_trace (persistent BVH8 traversal: closest hit + any hit)", "launches": stats["launches_extend"], "avg_launch_us": 1e3 * ext_ms,
                         "trace_Mrays_per_s": (stats["extend_rays"] + stats["shadow_rays"]) / max(stats["ms_extend"], 1e-9) / 1e3, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": ext_bytes, "peak_source": peak_src,
                         "bytes_per_ray": {"whole_vertex": BYTES_PER_RAY, "meaning": "152 B = half of the 304 B of wavefront state one path vertex with one light sample moves through ALL kernels "
                                           "(SURVEY 8d) — most of it is k_shade's traffic; k_trace alone reads a 32 B ray (+16 B contribution for shadow rays) and writes 16 B"},
                         "trace_only": {"algorithmic_bytes_per_launch": trace_only_bytes, "achieved": trace_only_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0,
                                        "frac": (trace_only_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0) / hbm_peak},
                         "whole_step": {"achieved": total_rays / world / elapsed * BYTES_PER_RAY / 1e9, "frac": total_rays / world / elapsed * BYTES_PER_RAY / 1e9 / hbm_peak,
                                        "meaning": "rays per second of the whole timed step (per GPU) x 152 B: the figure the 152 B were defined for"},
                         "note": "the kernel is latency / issue bound, not HBM bound: ncu counters in profiles/"},
            "stage_ms": {"trace": stats["ms_extend"], "frame_total": stats["ms_total"], "device": dev_ms, "wall": 1e3 * wall},
            "stage_split": split,
            "rays": {"extend": stats["extend_rays"], "shadow": stats["shadow_rays"], "iterations": stats["iterations"]},
        }
        if strong:
            line["strong"] = strong
        if c5:
            line["config5"] = c5
        if world == 1 and not args.no_em:
            # ---- second headline metric: guiding EM samples/s (BASELINE configs[0]) ------------------------------------
            rg = P.Renderer(64, 64, 0, EM_SPLITS, device=local_rank)
            rg.set_scene(scene)
            batches = em_batches(P, rg.guiding_aabbs())
            n = sum(len(b) for b in batches)
            per_order = {}
            for order, name in ((P.GUIDING_ORDER_REORDERED, "reordered"), (P.GUIDING_ORDER_STRICT, "strict")):
                rg.guiding_set_order(order)
                g = em_gpu(P, rg, batches, torch, dev)
                dev_s, host_s, fit_s = sum(g["device"]["ms"]) * 1e-3, sum(g["host"]["ms"]) * 1e-3, g["device"]["ms_fit"] * 1e-3
                per_order[name] = {"value": n / dev_s, "unit": "samples/s", "ms_per_update": g["device"]["ms"], "ms_sort_total": g["device"]["ms_sort"], "ms_fit_total": g["device"]["ms_fit"],
                                   "em_sample_iterations": g["device"]["sample_iters"], "gpu_launches": g["device"]["launches"], "e2e_value": n / host_s, "e2e_ms_per_update": g["host"]["ms"],
                                   "sample_iterations_per_s": g["device"]["sample_iters"] / fit_s}
            d = per_order["reordered"]
            em = {"metric": "guiding EM samples/s: PathGuiding::update (sort + preFit + fit|updateFit + merge/split + statistics + pack), "
                            "256 regions x 57600 records, 2 updates",
                  "value": d["value"], "unit": "samples/s", "ms_per_update": d["ms_per_update"], "ms_sort_total": d["ms_sort_total"], "ms_fit_total": d["ms_fit_total"],
                  "em_sample_iterations": d["em_sample_iterations"], "gpu_launches": d["gpu_launches"],
                  "summation_order": {"default": "reordered (block-parallel sums, work shared between all blocks): TOL_REORDERED of tests/test_guiding_cpu.py",
                                      "strict": "the reference's sequential float sums on the device: mixtures bit-identical to lightpmm (tests/test_guiding_gpu.py)",
                                      "reordered": per_order["reordered"], "strict_values": per_order["strict"]},
                  "e2e": {"value": d["e2e_value"], "unit": "samples/s", "h2d_bytes_per_step": len(batches[0]) * 40, "d2h_bytes_per_step": 0, "ms_per_update": d["e2e_ms_per_update"]},
                  "roofline": {"bound": "hbm", "kernel": "k_guiding_update_shared (persistent, chunked passes shared between all blocks)",
                               "achieved": d["em_sample_iterations"] * EM_BYTES_PER_SAMPLE_ITER / (d["ms_fit_total"] * 1e-3) / 1e9,
                               "peak": hbm_peak, "unit": "GB/s", "frac": d["em_sample_iterations"] * EM_BYTES_PER_SAMPLE_ITER / (d["ms_fit_total"] * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                               "sample_iterations_per_s": d["sample_iterations_per_s"],
                               "note": "16 B per sample per EM iteration of fit/updateFit (masked post-split fits and the statistics passes are extra work "
                                       "not counted here); ~30 flop/B: the kernel is FP32-issue bound (ncu: profiles/r02_ncu_guiding.txt), not HBM bound"}}
            if not args.no_cpu_baseline:
                cpu = em_cpu_reference(P, batches, os.cpu_count() or 1)
                if cpu:
                    em["cpu_baseline"] = cpu
                    em["speedup_vs_cpu_all_cores"] = em["value"] / cpu["value"]
                    em["strict_speedup_vs_cpu_all_cores"] = per_order["strict"]["value"] / cpu["value"]
            line["em"] = em
        if not args.no_cpu_baseline:
            cpu = cpu_baseline_reference_shaders() if world == 1 else None
            if cpu is None:      # the compiled reference shaders did not travel (or N > 1): the port, threads inside this process
                threads = os.cpu_count() or 1
                times, crays = cpu_oracle_run(P, 3, threads)       # 1 warm-up + 2 timed steps of the bounded sample
                cpu = {"value": sum(crays[1:]) / sum(times[1:]) / 1e6, "unit": UNIT, "cores": threads, "kind": "port", "sample": CPU_SAMPLE, "port_checked_against": PORT_PIN}
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
