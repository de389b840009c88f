"""Multi-GPU plumbing for the hot path: one process per GPU, independent samples per rank
(batch split of the denoising batch), NO data-path collective inside a denoise step, and ONE collective at the end
of the schedule: the all-gather of the decoded uint8 frames over NVLink (NCCL; gloo on CPU for the tests).

The reference has no distributed layer at all on this path; nothing here mirrors reference code.
"""
import os

import torch
import torch.distributed as dist


def init(backend=None, device=None):
    """Initialise the default process group from torchrun's env (RANK / WORLD_SIZE / MASTER_*). No-op for 1 process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return int(os.environ.get("RANK", "0")), world
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl" and device is not None:
        kw["device_id"] = torch.device(device)
    dist.init_process_group(backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def shard(n_items, rank, world):
    """Contiguous split of n_items samples over `world` ranks (first n_items % world ranks get one more)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def make_cfg_pairs():
    """CFG-pair split: ranks (2k, 2k+1) form one pair working on sample k.  Every rank must call this (new_group is
    collective).  Returns (pair_group, cfg_rank, sample_index, n_pairs)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world % 2:
        raise ValueError("CFG-pair split needs an even number of ranks")
    mine = None
    for k in range(world // 2):
        g = dist.new_group([2 * k, 2 * k + 1])
        if rank // 2 == k:
            mine = g
    return mine, rank % 2, rank // 2, world // 2


def allgather_frames(frames, group=None):
    """frames: uint8 tensor [n_local, 3, F, H, W] (same F,H,W on every rank; n_local may differ by one).
    Returns the frames of ALL samples in global sample order on every rank -- the single collective of the schedule."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return frames
    world = dist.get_world_size(group)
    counts = torch.zeros(world, dtype=torch.int64, device=frames.device)
    counts[dist.get_rank(group)] = frames.shape[0]
    dist.all_reduce(counts, group=group)
    nmax = int(counts.max())
    pad = frames
    if frames.shape[0] < nmax:
        pad = torch.cat([frames, frames.new_zeros((nmax - frames.shape[0],) + tuple(frames.shape[1:]))], 0)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)          # ncclAllGather on GPUs, gloo in the CPU tests
    return torch.cat([out[r][: int(counts[r])] for r in range(world)], 0)


def generate_batch(make_denoiser, contexts, context_null, latent_shape, seeds, device=None, fused=None):
    """Batch-split generation: rank r denoises + decodes samples shard(len(seeds), r, world) with its own denoiser
    (weights replicated; both Wan2.2 experts fit one 180 GB GPU), then all ranks all-gather the uint8 frames.
    `make_denoiser()` returns a wan2gp_b200.pipeline.WanDenoiser with a VAE attached.

    fused (default: on for NCCL with one sample per rank): the decoded frames stay fp32 on the device and the quantisation to uint8 is
    fused with the all-gather (`FusedFrameGather`: every rank stores its bytes straight into all peers' buffers over NVLink) --
    no CPU round trip, no separate ncclAllGather.  Otherwise frames_to_u8 + allgather_frames (NCCL / gloo)."""
    rank, world = init(device=device)
    if fused is None:
        fused = world > 1 and dist.is_initialized() and dist.get_backend() == "nccl" and len(seeds) == world
    mine = shard(len(seeds), rank, world)
    den = make_denoiser()
    dev = device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
    frame_shape = (3, 4 * (latent_shape[1] - 1) + 1, 8 * latent_shape[2], 8 * latent_shape[3])
    outs, aborted = [], False
    for i in mine:
        res = den.generate(contexts[i], context_null, latent_shape, seed=seeds[i], device_frames=fused)
        if res is None:                              # this rank was interrupted: stop working, but still enter the collectives below
            aborted = True
            break
        outs.append(res["x"])
    # `_interrupt` is per process: agree on the abort collectively so that no rank is left waiting in the all-gather
    if world > 1:
        flag = torch.tensor([1.0 if aborted else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        aborted = bool(flag.item() > 0)
    if aborted:
        return None
    if fused and world > 1 and len(seeds) == world:
        # one sample per rank, frames still fp32 on the device: quantise + all-gather in ONE kernel over NVLink peer memory
        fr = outs[0].contiguous()
        key = (fr.numel(), str(dev))
        fg = _FUSED.get(key)
        if fg is None:
            fg = _FUSED[key] = FusedFrameGather(fr.numel(), dev)
        return fg.gather(fr).view(world, *frame_shape).clone()
    local = torch.stack(outs, 0).to(dev) if outs else torch.empty((0,) + frame_shape, dtype=torch.uint8, device=dev)
    return allgather_frames(local)


_FUSED = {}          # (bytes per rank, device) -> FusedFrameGather (symmetric buffers are expensive to rendezvous: keep them)


class FusedFrameGather:
    """Fused `frames -> uint8 -> all-gather` over NVLink peer memory (csrc/vae_ops.cu::frames_to_u8_allgather_kernel).

    One symmetric buffer [world, n] uint8 per GPU (torch.distributed._symmetric_memory: CUDA VMM allocations mapped into
    every peer); each rank's quantisation kernel stores its bytes into slot `rank` of ALL buffers, then a signal-pad barrier
    makes them visible.  Replaces frames_to_u8 + ncclAllGather (allgather_frames) when all ranks share one NVSwitch box."""

    def __init__(self, n_bytes_per_rank, device, group=None):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n = int(n_bytes_per_rank)
        self.buf = symm_mem.empty(self.world * self.n, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        ptrs = list(self.hdl.buffer_ptrs)
        self._ptrs = (ctypes.c_uint64 * len(ptrs))(*ptrs)

    def gather(self, frames_fp32):
        """frames_fp32: this rank's fp32 frames (numel == n).  Returns uint8 [world, n] on this GPU (all ranks' frames)."""
        import ctypes

        from . import _lib
        assert frames_fp32.numel() == self.n and frames_fp32.is_contiguous() and frames_fp32.dtype == torch.float32
        self.hdl.barrier(channel=0)                 # every rank has finished reading the previous contents
        _lib.call("b200_frames_to_u8_allgather", frames_fp32.data_ptr(), ctypes.cast(self._ptrs, ctypes.c_void_p), self.world,
                  self.rank, self.n, torch.cuda.current_stream().cuda_stream)
        self.hdl.barrier(channel=1)                 # all peers' stores have landed
        return self.buf.view(self.world, self.n)
