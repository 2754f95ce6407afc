"""BASELINE config 5 (stand-in scene: sponzaXML — fireplace_room.obj is missing from the reference checkout): 3840x2160,
GUIDING_SPLITS=8, 6 guiding optimisation frames (updateGuiding) through the frame driver, then guided frames
(useGuiding, guidingProb 0.5, parallax compensation).  Prints one JSON line (rank 0); the work is bench.config5_run.
  1 GPU :  python tools/run_config5.py [W H frames spp]
  N GPUs:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/run_config5.py [W H frames spp]
With N ranks the run is spp-sharded: rank g renders frame f with seed tea(f*N+g, seed), every training
frame ends with the region-sharded refit b200pt_guiding_update_all_ranks (compacted records exchanged over NVLink, each
region fitted once), and the image is combined with b200pt_reduce_image.  Times are device times, max over ranks."""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
W, H = (int(args[0]), int(args[1])) if len(args) > 1 else (3840, 2160)
FRAMES = int(args[2]) if len(args) > 2 else 4
SPP = int(args[3]) if len(args) > 3 else 2
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
if world > 1:
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P = bench._load("b200pt_binding", os.path.join(bench.ROOT, "rtx-pathtracer_b200", "b200pt.py"))
out = bench.config5_run(P, torch, dist, rank, world, local, W, H, SPP, FRAMES)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
