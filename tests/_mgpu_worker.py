"""torchrun worker of tests/test_multigpu_gpu.py: batch-split generation of a reduced Wan config on N GPUs through
wan2gp_b200.dist.generate_batch, once with the fused quantise + all-gather over NVLink peer memory (the default on NCCL) and once
with frames_to_u8 + ncclAllGather; every rank must hold identical uint8 frames of ALL samples, bit-equal between the two paths."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from wan2gp_b200 import dist as wd, synth
    from wan2gp_b200.pipeline import WanDenoiser
    from wan2gp_b200.wan import WanModel, WanVAE
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rank, world = wd.init(device=dev)
    cfg = synth.WAN_CONFIGS["tiny"]
    sd = synth.make_wan_state_dict(cfg, 0)
    vsd = synth.make_vae_state_dict(synth.VAE_CFG_TINY, 0)

    def make():
        m = WanModel(**cfg, device=dev)
        m.load_state_dict(sd)
        return WanDenoiser(m, vae=WanVAE(device=dev, state_dict=vsd, cfg=synth.VAE_CFG_TINY), num_steps=2, shift=5.0, guide_scale=4.0, device=dev)
    g = torch.Generator().manual_seed(0)
    ctxs = [torch.randn(1, cfg["text_len"], cfg["text_dim"], generator=g) for _ in range(world)]
    null = torch.zeros(1, cfg["text_len"], cfg["text_dim"])
    seeds = list(range(100, 100 + world))
    fused = wd.generate_batch(make, ctxs, null, (16, 3, 8, 12), seeds, device=dev, fused=True)
    plain = wd.generate_batch(make, ctxs, null, (16, 3, 8, 12), seeds, device=dev, fused=False)
    assert fused.shape == plain.shape == (world, 3, 9, 64, 96) and fused.dtype == torch.uint8, (fused.shape, plain.shape)
    assert torch.equal(fused, plain.to(fused.device)), "fused gather differs from frames_to_u8 + ncclAllGather"
    assert len({int(fused[i].sum()) for i in range(world)}) == world          # the samples differ (different seeds / prompts)
    # every rank holds the same gathered tensor
    chk = torch.tensor([float(fused.float().sum())], device=dev, dtype=torch.float64)
    lo, hi = chk.clone(), chk.clone()
    torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
    torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
    assert float(lo) == float(hi)
    # CFG-pair split on GPUs (BASELINE configs[2]): rank 2k runs the cond branch, rank 2k+1 the uncond branch, one NCCL all-gather of the
    # prediction (+ abort flag) per step; the latents of both ranks stay identical and equal to the single-rank joint pass
    if world % 2 == 0:
        grp, cfg_rank, sample, _ = wd.make_cfg_pairs()
        m = WanModel(**cfg, device=dev)
        m.load_state_dict(sd)
        gg = torch.Generator().manual_seed(7 + sample)
        lat0 = torch.randn(1, 16, 3, 8, 12, generator=gg).to(dev)
        c, cn = ctxs[sample].to(dev), null.to(dev)
        for solver in ("euler", "unipc"):
            split = WanDenoiser(m, num_steps=3, shift=5.0, guide_scale=4.0, device=dev, cfg_group=grp, cfg_rank=cfg_rank, sample_solver=solver, cfg_star_switch=True)
            joint = WanDenoiser(m, num_steps=3, shift=5.0, guide_scale=4.0, device=dev, sample_solver=solver, cfg_star_switch=True)
            a, b = lat0.clone(), lat0.clone()
            for i in range(3):
                assert split.step(a, i, c, cn) is a and joint.step(b, i, c, cn) is b
            assert torch.equal(a, b), f"cfg split ({solver}) differs from the joint pass: {float((a - b).abs().max())}"
    if rank == 0:
        print("MGPU_OK", world)
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
