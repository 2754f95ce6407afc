"""Multi-GPU plumbing of the path tracer: one process per GPU, `torch.distributed` (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  SURVEY.md §8(e): the image shards by SAMPLE INDEX — every rank renders disjoint frames of the
full image with its own frame-seed stream, scene + BVH are replicated, and there is no data-path collective inside a
frame.  Two exchange steps exist:
  * image:   one all-reduce (sum) of the per-rank RGBA32F running means, weighted by the frames each rank rendered;
  * guiding: an all-gather of the ranks' DirectionalData buffers before a refit, after which every rank runs the same
             deterministic `b200pt_guiding_update_device` on the same records (so no VMM broadcast is needed).
The reference is single-GPU (no collective call sites); this module is our addition and has no counterpart to mirror.
It is the torch.distributed VARIANT of the two exchange steps, kept as the checker of the native path and for the gloo
tests; the product path is inside the library: b200pt_reduce_image, and b200pt_guiding_update_all_ranks (region-sharded
refit: only compacted valid records travel, every region is fitted once) — bench.py and the frame driver use those."""
import torch
import torch.distributed as dist

RECORD_BYTES = 40      # sizeof(DirectionalData), shaders/guiding.glsl:99-113


def tea(val0, val1):
    """shaders/random.glsl:13-27"""
    v0, v1, s0 = val0 & 0xFFFFFFFF, val1 & 0xFFFFFFFF, 0
    for _ in range(16):
        s0 = (s0 + 0x9E3779B9) & 0xFFFFFFFF
        v0 = (v0 + ((((v1 << 4) & 0xFFFFFFFF) + 0xA341316C) ^ (v1 + s0) ^ ((v1 >> 5) + 0xC8013EA4))) & 0xFFFFFFFF
        v1 = (v1 + ((((v0 << 4) & 0xFFFFFFFF) + 0xAD90777D) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7E95761E))) & 0xFFFFFFFF
    return v0


def frame_seed(step, rank, world, seed):
    """Frame `step` of rank `rank`: global frame index step*world + rank, so N ranks x K steps consume exactly the
    seeds a single GPU would use for N*K frames."""
    return tea(step * world + rank, seed)


def app_frame_seed(app_state, step, rank, world, seed):
    """Seed of the next frame when the frames are drawn through the frame driver (b200pt_app_*).  Irradiance-cache
    prepare frames and the ADRRS estimate frame get the SAME seed on every rank, so all ranks build identical caches and
    estimate images without any communication; every other frame is sharded like frame_seed().
    `app_state` is the driver's state BEFORE begin_frame (the condition mirrors src/RayTracingApp.cpp:1130-1150)."""
    st = app_state
    uses_cache = bool(st.settings.useIrradianceCache or st.settings.useADRRS)
    prepare = uses_cache and st.currentPrepareFrames < st.irradianceCachePrepareFrames
    estimate = uses_cache and bool(st.activateADRRSAfterPrepareFrames) and st.currentPrepareFrames == st.irradianceCachePrepareFrames
    if prepare or estimate:
        return tea(step, seed ^ 0x1C1C1C1C)
    return frame_seed(step, rank, world, seed)


def global_frame_indices(steps, rank, world):
    return [s * world + rank for s in range(steps)]


def combine_images(local_mean, frames_local, group=None):
    """All ranks: running-mean image of this rank's `frames_local` frames (H, W, 4 float32, on the device the backend
    expects) -> mean over ALL frames of all ranks.  In place on a scaled copy; one all-reduce of the image + one of a
    scalar."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_mean
    acc = local_mean * float(frames_local)
    n = torch.tensor([float(frames_local)], dtype=torch.float64, device=local_mean.device)
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return acc / float(n.item())


def allgather_samples(records, group=None):
    """records: (n, 40) uint8 tensor of this rank's DirectionalData (n may differ per rank; INVALID records may be
    included — the device sort drops them).  Returns the concatenation over ranks in rank order, identical on every
    rank, so that each rank's refit sees the same records in the same order."""
    assert records.dtype == torch.uint8 and records.dim() == 2 and records.shape[1] == RECORD_BYTES
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return records
    world = dist.get_world_size(group)
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=records.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes)
    padded = torch.zeros((cap, RECORD_BYTES), dtype=torch.uint8, device=records.device)
    padded[:records.shape[0]] = records
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def guiding_update_all_ranks(renderer, params=None, group=None):
    """Training-frame exchange: every rank contributes the DirectionalData its frame recorded (binding 18, W*H*16
    records incl. INVALID slots), the buffers are all-gathered over NCCL, and every rank refits on the identical
    concatenation — the mixtures stay bit-identical across ranks without a broadcast."""
    n = renderer.guiding_sample_capacity()
    device = torch.device("cuda", torch.cuda.current_device())
    local = torch.empty((n, RECORD_BYTES), dtype=torch.uint8, device=device)
    renderer.guiding_get_samples_device(local.data_ptr(), n)
    allrec = allgather_samples(local, group)
    # the library launches on its own non-blocking stream: everything torch queued for `allrec` (NCCL all-gather, the
    # concatenation) has to be complete before the sort kernels read it
    torch.cuda.current_stream().synchronize()
    renderer.guiding_update_device(allrec.data_ptr(), allrec.shape[0], params)
    return allrec.shape[0]
