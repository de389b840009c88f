"""2+ GPU check (torchrun): fused quantise+all-gather over peer memory == frames_to_u8 + ncclAllGather, and its timing."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wan2gp_b200 import _lib  # noqa: E402
from wan2gp_b200 import dist as wd  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n = 3 * 81 * 720 * 1280
g = torch.Generator(device=dev).manual_seed(rank)
frames = torch.randn(n, device=dev, generator=g) * 0.7
fg = wd.FusedFrameGather(n, dev)
out = fg.gather(frames).clone()
# reference path: local quantise kernel + NCCL all-gather
u8 = torch.empty(n, device=dev, dtype=torch.uint8)
_lib.call("b200_frames_to_u8", frames.data_ptr(), u8.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
ref = [torch.empty_like(u8) for _ in range(world)]
dist.all_gather(ref, u8)
ok = all(torch.equal(out[r], ref[r]) for r in range(world))


def timeit(fn, it=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it


def nccl_path():
    _lib.call("b200_frames_to_u8", frames.data_ptr(), u8.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    dist.all_gather(ref, u8)


t_f, t_n = timeit(lambda: fg.gather(frames)), timeit(nccl_path)
tt = torch.tensor([t_f, t_n], device=dev)
dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"fused_gather_check world={world} equal={ok} fused_ms={float(tt[0]):.3f} u8+nccl_allgather_ms={float(tt[1]):.3f} bytes_per_rank={n}")
dist.destroy_process_group()
