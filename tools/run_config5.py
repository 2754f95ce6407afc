"""BASELINE config 5 (stand-in scene: sponzaXML — fireplace_room.obj is missing from the reference checkout): 3840x2160,
GUIDING_SPLITS=8, 6 guiding optimisation frames (updateGuiding) through the frame driver, then guided frames
(useGuiding, guidingProb 0.5, parallax compensation).  Prints one JSON line (rank 0).
  1 GPU :  python tools/run_config5.py [W H frames spp]
  N GPUs:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/run_config5.py [W H frames spp]
With N ranks the run is spp-sharded: rank g renders frame f with seed tea(f*N+g, seed), every training
frame ends with the region-sharded refit b200pt_guiding_update_all_ranks (compacted records exchanged over NVLink, each
region fitted once), and the image is combined with b200pt_reduce_image.  Times are device times, max over ranks."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers  # noqa: E402

P = helpers.pt()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
W, H = (int(args[0]), int(args[1])) if len(args) > 1 else (3840, 2160)
FRAMES = int(args[2]) if len(args) > 2 else 4
SPP = int(args[3]) if len(args) > 3 else 2
SCENE = "sponzaXML"
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scene = P.Scene(helpers.scene_path(SCENE))
view, proj = scene.camera_matrices(W / H)
r = P.Renderer(W, H, 0, 8, device=local)
r.set_scene(scene)
r.set_camera(view, proj)
if world > 1:
    ids = [P.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    r.comm_init(ids[0], rank, world)
app = P.App(r, accumulate=True, samplesPerPixel=SPP, enableNEE=1, enableMIS=1, updateGuiding=1, useParallaxCompensation=1)
out = dict(scene=SCENE + " (stand-in for fireplace)", width=W, height=H, spp_per_frame=SPP, regions=r.guiding_region_count(), n_gpus=world, phases=[])
step = 0


def maxed(vals):
    """max over ranks of a list of floats"""
    if world == 1:
        return vals
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def summed(vals):
    if world == 1:
        return vals
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def phase(label, n):
    global step
    r.stats_reset()
    if world > 1:
        dist.barrier()
    t = time.time()
    for _ in range(n):
        app.draw_frame(P.tea(step * world + rank, 0xC0FFEE))
        step += 1
    wall = (time.time() - t) * 1e3
    st = r.stats()
    render, sort, xchg, fit, gather, wall = maxed([st.ms_total, st.ms_guiding_sort, st.ms_guiding_exchange, st.ms_guiding_fit, st.ms_guiding_gather, wall])
    rays, owned, recv = summed([float(st.extend_rays + st.shadow_rays), float(st.guiding_samples), float(st.guiding_bytes_received)])
    out["phases"].append(dict(phase=label, frames_per_rank=n, wall_ms=round(wall, 1), render_ms=round(render, 1), guiding_sort_ms=round(sort, 2),
                              guiding_exchange_ms=round(xchg, 2), guiding_fit_ms=round(fit, 2), guiding_gather_ms=round(gather, 2),
                              guiding_samples_all_ranks=int(st.guiding_samples_all_ranks), guiding_samples_fitted_sum=int(owned),
                              guiding_samples_fitted_this_rank=int(st.guiding_samples), exchange_bytes_all_ranks=int(recv),
                              Mrays_per_s=round(rays / max(render, 1e-6) / 1e3, 1)))


phase("training (updateGuiding, 6 refits)", app.state.numGuidingOptimizations + 1)
assert app.settings.updateGuiding == 0
app.settings.useGuiding = 1
app.settings.guidingProb = 0.5
app.input_changed()
phase("guided render", FRAMES)
vm = r.guiding_get_vmms()
if world > 1:
    t0 = time.time()
    r.reduce_image(P.IMAGE_OUTPUT, FRAMES)
    out["reduce_image_ms"] = round(maxed([(time.time() - t0) * 1e3])[0], 2)
    v = torch.from_numpy(vm.view("u1").copy()).cuda()
    parts = [torch.empty_like(v) for _ in range(world)]
    dist.all_gather(parts, v)
    out["mixtures_identical_on_all_ranks"] = all(bool(torch.equal(parts[0], p)) for p in parts)
    out["exchange_mode"] = {0: "none", 1: "ncclSend/ncclRecv", 2: "CUDA-IPC peer reads over NVLink"}[r.comm_exchange_mode()]
img = r.read_image()[..., :3]
out.update(image_mean=float(img.mean()), finite=bool((img == img).all()), mean_components=float(vm["usedDistributions"].mean()))
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    r.comm_destroy()
    dist.destroy_process_group()
