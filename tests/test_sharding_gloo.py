"""World-size-2 checks of the multi-GPU host logic on CPU (gloo): frame-seed sharding, the image reduction and the
variable-size guiding-sample all-gather (rtx-pathtracer_b200/sharding.py).  The per-rank "renderer" here is the CPU
oracle at a tiny resolution — test scaffolding only; on the GPU box the same functions wrap the CUDA renderer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers

W, H, STEPS, SEED = 24, 16, 3, 0xC0FFEE


def _sharding():
    return helpers._load("b200pt_sharding", os.path.join(helpers.PKG_DIR, "sharding.py"))


def _render_frames(frame_indices):
    """running mean over the given global frame indices, rendered by the oracle (1 spp, NEE + MIS)"""
    P = helpers.pt()
    scene, _, o = helpers.make_pair("veachMIS", W, H, gpu=False)
    for k, f in enumerate(frame_indices):
        pc = P.default_push_constants(randomUInt=P.tea(f, SEED), previousFrames=k, samplesPerPixel=1, enableMIS=1)
        o.render_region(pc, threads=1)
    return o.image().astype(np.float32).copy()


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = _sharding()
    steps = STEPS + (1 if rank == 0 else 0)                      # ragged: rank 0 renders one more frame
    idx = [s * world + rank for s in range(steps)]
    assert [S.frame_seed(s, rank, world, SEED) for s in range(steps)] == [S.tea(i, SEED) for i in idx]
    local = torch.from_numpy(_render_frames(idx))
    combined = S.combine_images(local, steps)
    # variable-size sample gather: rank r contributes 5 + 3 r records tagged with its rank
    rec = torch.zeros((5 + 3 * rank, 40), dtype=torch.uint8)
    rec[:, 36] = rank + 1
    rec[:, 0] = torch.arange(rec.shape[0], dtype=torch.uint8)
    gathered = S.allgather_samples(rec)
    out.put((rank, combined.numpy(), gathered.numpy(), idx))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_image_reduction_and_sample_gather():
    helpers.ensure_built()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, img0, g0, idx0), (_, img1, g1, idx1) = results
    assert sorted(idx0 + idx1) == list(range(2 * STEPS)) + [2 * STEPS]           # disjoint frames, nothing skipped
    assert np.array_equal(img0, img1) and np.array_equal(g0, g1)                # every rank ends with the same data
    # the reduced image equals the mean over all frames computed in one process
    ref = np.mean([_render_frames([f]) for f in sorted(idx0 + idx1)], axis=0)
    assert np.allclose(img0[..., :3], ref[..., :3], rtol=1e-5, atol=1e-6)
    # gathered records: rank order, sizes 5 and 8, contents intact
    assert g0.shape == (13, 40) and list(g0[:, 36]) == [1] * 5 + [2] * 8 and list(g0[:5, 0]) == list(range(5)) and list(g0[5:, 0]) == list(range(8))


def test_single_process_passthrough():
    S = _sharding()
    x = torch.ones((2, 2, 4))
    assert S.combine_images(x, 3) is x
    r = torch.zeros((4, 40), dtype=torch.uint8)
    assert S.allgather_samples(r) is r
    assert S.frame_seed(2, 1, 4, 7) == helpers.pt().tea(9, 7)


def test_app_frame_seed_schedule_is_common_during_ic_preparation():
    """Prepare frames and the ADRRS estimate frame: same seed on every rank (identical caches without communication);
    afterwards the frames are sharded by global frame index.  Pure host logic."""
    P, S = helpers.pt(), _sharding()
    world = 4
    seeds = []
    for rank in range(world):
        app = P.App(accumulate=True, useADRRS=1, samplesPerPixel=2)
        app.state.irradianceCachePrepareFrames = 3
        mine = []
        for step in range(7):
            sd = S.app_frame_seed(app.state, step, rank, world, SEED)
            pc = app.begin_frame(sd)
            mine.append((sd, pc.isIrradiancePrepareFrame, pc.storeEstimate, pc.useADRRS))
            app.end_frame()
        seeds.append(mine)
    for step in range(4):                    # 3 prepare frames + the estimate frame
        assert len({seeds[r][step][0] for r in range(world)}) == 1
    assert [seeds[0][s][1] for s in range(7)] == [1, 1, 1, 0, 0, 0, 0] and seeds[0][3][2] == 1
    for step in range(5, 7):                 # (step 4 re-uses the estimate frame's seed: reference behaviour, see test_app_driver)
        assert len({seeds[r][step][0] for r in range(world)}) == world
        assert all(seeds[r][step][0] == S.frame_seed(step, r, world, SEED) and seeds[r][step][3] == 1 for r in range(world))
