#!/usr/bin/env python3
"""torchrun --nproc-per-node 2 tools/multi_gpu_guiding_check.py  (under gpurun --gpus 2)
Two ranks render different training frames, all-gather their samples over NCCL and refit; checks that both ranks end
with bit-identical mixtures and that those equal a single-GPU refit on the concatenated records."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P = helpers.pt()
S = helpers._load("b200pt_sharding", os.path.join(helpers.PKG_DIR, "sharding.py"))
W, H = 160, 90
scene = P.Scene(helpers.scene_path("cornell-dielectric"))
view, proj = scene.camera_matrices(W / H)
r = P.Renderer(W, H, 0, 5, device=local); r.set_scene(scene); r.set_camera(view, proj)
for step in range(2):
    pc = P.default_push_constants(randomUInt=S.frame_seed(step, rank, world, 0xC0FFEE), previousFrames=0, samplesPerPixel=4, enableMIS=1, updateGuiding=1, useGuiding=int(step > 0))
    r.render_frame(pc)
    mine = r.guiding_get_samples()
    n = S.guiding_update_all_ranks(r)
    vm = torch.from_numpy(r.guiding_get_vmms().view(np.uint8).copy()).cuda()
    parts = [torch.empty_like(vm) for _ in range(world)]
    dist.all_gather(parts, vm)
    same = all(torch.equal(parts[0], p) for p in parts)
    # single-GPU reference: gather the raw records on rank 0 and refit there in a fresh context driven identically
    allmine = [None] * world
    dist.all_gather_object(allmine, mine)
    if rank == 0:
        if step == 0:
            r1 = P.Renderer(W, H, 0, 5, device=local); r1.set_scene(scene)
        r1.guiding_update_host(np.concatenate(allmine))
        ok = np.array_equal(r1.guiding_get_vmms().view(np.uint8), r.guiding_get_vmms().view(np.uint8))
        print("step %d: %d records gathered, ranks identical: %s, equals single-GPU refit of the concatenation: %s" % (step, n, same, ok), flush=True)
        assert same and ok
dist.barrier()
dist.destroy_process_group()
