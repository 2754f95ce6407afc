#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/multi_gpu_guiding_check.py [regions_log2 per_region rounds]   (under gpurun --gpus N)
Region-sharded guiding refit (b200pt_guiding_update_all_ranks*): every rank holds different records (ragged region
counts, INVALID slots, one region only some ranks see, one region nobody sees); after each of three updates
  * the mixtures (full fit state + packed VMM_Thetas) are bit-identical on all ranks, and
  * they equal, bit for bit, a single-GPU update on the concatenation of the ranks' buffers in rank order.
Run once per exchange mode: B200PT_EXCHANGE unset (CUDA-IPC peer reads) and B200PT_EXCHANGE=nccl (send/recv)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import guiding_data  # noqa: E402
import helpers  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P = helpers.pt()
SPLITS = int(sys.argv[1]) if len(sys.argv) > 1 else 5
PER_REGION = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
ROUNDS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
scene = P.Scene(helpers.scene_path("cornell-dielectric"))
r = P.Renderer(64, 64, 0, SPLITS, device=local)
r.set_scene(scene)
ids = [P.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
r.comm_init(ids[0], rank, world)
ref = P.Renderer(64, 64, 0, SPLITS, device=local)          # single-GPU reference on the concatenation
ref.set_scene(scene)
aabbs = r.guiding_aabbs()
R = len(aabbs)
ok_all = True
for rnd in range(ROUNDS):
    rng = np.random.default_rng(100 * rnd + 7)
    counts = rng.integers(PER_REGION // 3, PER_REGION, size=(world, R))
    counts[:, rnd % R] = 0                                  # a region nobody has samples for
    counts[1:, (rnd + 3) % R] = 0                           # a region only rank 0 sees
    counts[:, (rnd + 5) % R] = 2                            # too few samples to fit (< 2K)
    batches = [guiding_data.make_batch(aabbs, list(counts[s]), 5000 + 31 * rnd + s) for s in range(world)]
    n = max(len(b) for b in batches)                        # equal buffer size on every rank: pad with INVALID records
    padded = []
    for b in batches:
        pad = np.zeros(n - len(b), dtype=guiding_data.DD)
        pad["flags"] = guiding_data.INVALID
        padded.append(np.concatenate([b, pad]))
    mine = torch.from_numpy(padded[rank].view(np.uint8).reshape(-1, 40).copy()).cuda()
    torch.cuda.synchronize()                                # the library runs on its own stream
    r.stats_reset()
    r.guiding_update_all_ranks_device(mine.data_ptr(), n)
    st = r.stats()
    ref.guiding_update_host(np.concatenate(padded))
    vm, vr = r.guiding_get_vmms().view(np.uint8), ref.guiding_get_vmms().view(np.uint8)
    same_ref = bool(np.array_equal(vm, vr))
    for g in range(R):
        a, b = r.guiding_state(g), ref.guiding_state(g)
        for k in a:
            same_ref = same_ref and bool(np.array_equal(np.asarray(a[k]), np.asarray(b[k]), equal_nan=True))
    t = torch.from_numpy(vm.copy()).cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    same_ranks = all(bool(torch.equal(parts[0], p)) for p in parts)
    tot = int(counts.sum())
    print("round %d rank %d/%d mode %d: owned %d of %d samples in %d regions, %.2f MB from peers | sort %.3f exchange %.3f fit %.3f gather %.3f ms | == single-GPU refit: %s | identical on all ranks: %s"
          % (rnd, rank, world, r.comm_exchange_mode(), st.guiding_samples, st.guiding_samples_all_ranks, st.guiding_regions_fit, st.guiding_bytes_received / 1e6,
             st.ms_guiding_sort, st.ms_guiding_exchange, st.ms_guiding_fit, st.ms_guiding_gather, same_ref, same_ranks), flush=True)
    ok_all = ok_all and same_ref and same_ranks and st.guiding_samples_all_ranks == tot
flag = torch.tensor([1.0 if ok_all else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
r.comm_destroy()
dist.destroy_process_group()
if flag.item() != 1.0:
    sys.exit("multi_gpu_guiding_check FAILED")
if rank == 0:
    print("multi_gpu_guiding_check OK (%d ranks)" % world)
