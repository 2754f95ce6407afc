"""Run under `gpurun --gpus 2` (or more) with torchrun: the native NCCL entry points of the C ABI against the
torch.distributed path (sharding.py) — same reduced image, same gathered records, identical refit on every rank.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_native_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
P = helpers.pt()
S = helpers._load("b200pt_sharding", os.path.join(helpers.PKG_DIR, "sharding.py"))
W, H = 320, 180
scene = P.Scene(helpers.scene_path("cornell-dielectric"))
view, proj = scene.camera_matrices(W / H)
r = P.Renderer(W, H, 0, 3, device=torch.cuda.current_device())
r.set_scene(scene)
r.set_camera(view, proj)
ids = [P.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
r.comm_init(ids[0], rank, world)
steps = 3 + (1 if rank == 0 else 0)        # ragged frame counts
for s in range(steps):
    r.render_frame(P.default_push_constants(randomUInt=S.frame_seed(s, rank, world, 0xC0FFEE), previousFrames=s, samplesPerPixel=2, enableMIS=1, updateGuiding=1))
local = torch.from_numpy(r.read_image()).cuda()
ref = S.combine_images(local, steps).cpu().numpy()
r.reduce_image(P.IMAGE_OUTPUT, steps)
mine = r.read_image()
err = float(np.abs(mine - ref).max() / max(np.abs(ref).max(), 1e-9))
# sample gather + identical refit
n_native = r.allgather_samples()
r2 = P.Renderer(W, H, 0, 3, device=torch.cuda.current_device())
r2.set_scene(scene)
r2.guiding_put_samples(r.guiding_get_samples())
n_torch = S.guiding_update_all_ranks(r2)
r.guiding_update_all_ranks()
same_fit = bool(np.array_equal(r.guiding_get_vmms().view(np.uint8), r2.guiding_get_vmms().view(np.uint8)))
vm = torch.from_numpy(r.guiding_get_vmms().view(np.uint8).copy()).cuda()
allvm = [torch.empty_like(vm) for _ in range(world)]
dist.all_gather(allvm, vm)
same_ranks = all(bool(torch.equal(allvm[0], v)) for v in allvm)
print("rank %d/%d: reduce_image max rel err vs torch %.2e | gathered %d (torch %d) | refit == torch path: %s | mixtures identical on all ranks: %s"
      % (rank, world, err, n_native, n_torch, same_fit, same_ranks), flush=True)
assert err < 1e-6 and n_native == n_torch and same_fit and same_ranks
dist.barrier()
r.comm_destroy()
dist.destroy_process_group()
