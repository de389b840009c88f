"""BASELINE configs[4]: decode-only sweep of the three VAEs -- latent T in {9, 33, 65, 129} at 720p / 1080p -- on 1 or N GPUs (one clip
per GPU under torchrun: every rank decodes its own clip, value = total frames/s, max over ranks).  One JSON line per case ->
stdout and gpurun_out/vae_sweep.jsonl.

    python tools/vae_sweep.py [--decoders wan,hyvae15,hyvae10] [--T 9,33,65,129] [--res 720p,1080p] [--max-seconds 60]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/vae_sweep.py ...

Mode per case: the whole clip in one pass while its activations fit (the B200-native default); beyond that the Wan decoder runs the
streamed (time-sliced, bit-identical) decode and the Hunyuan decoders the reference's own tiling (enable_tiling(): temporal tiles of
16(+1) latent frames, 256 px spatial tiles with cross-faded seams -- what the reference pipelines always use)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wan2gp_b200 import _lib, synth  # noqa: E402

RES = {"720p": (720, 1280), "1080p": (1080, 1920)}


def build(name, dev):
    if name == "wan":
        from wan2gp_b200.wan import WanVAE
        vae = WanVAE(device=dev, state_dict=synth.make_vae_state_dict(seed=0))
        return vae, 16, 8
    from wan2gp_b200.hyvideo import AutoencoderKLCausal3D, AutoencoderKLConv3D, HYVAE10Decoder, HYVAEDecoder
    if name == "hyvae15":
        cfg = synth.HYVAE_CONFIGS["hyvae15"]
        vae = AutoencoderKLConv3D(device=dev)                        # upstream defaults = the production config (tiling parameters included)
        vae.decoder = HYVAEDecoder(cfg, dev)
        vae.decoder.load_state_dict(synth.make_hyvae_state_dict(cfg, 0, device=dev))
        return vae, cfg["z_channels"], cfg["ffactor_spatial"]
    cfg = synth.HYVAE10_CONFIGS["hyvae10"]
    vae = AutoencoderKLCausal3D(device=dev)
    vae.decoder = HYVAE10Decoder(cfg, dev)
    vae.decoder.load_state_dict(synth.make_hyvae10_state_dict(cfg, 0, device=dev))
    return vae, cfg["latent_channels"], 8


def decode(name, vae, z, mode):
    if name == "wan":
        if mode == "whole":
            return vae.model.decode_frames(z[0], vae.mean, vae.std)
        return vae.model.decode_frames_streamed(z[0], vae.mean, vae.std, chunk=mode[1])
    if mode == "whole":
        vae.disable_tiling()
    else:
        vae.enable_tiling()
    return vae.decode(z, return_dict=False)[0][0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--decoders", default="wan,hyvae15,hyvae10")
    ap.add_argument("--T", default="9,33,65,129")
    ap.add_argument("--res", default="720p,1080p")
    ap.add_argument("--max-seconds", type=float, default=45.0, help="skip cases whose predicted decode time exceeds this")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    logf = open(os.path.join(ROOT, "gpurun_out", f"vae_sweep_n{world}.jsonl"), "a") if rank == 0 else None
    rate = {}                                  # decoder -> measured seconds per (output frame x megapixel), to predict the next case
    for name in args.decoders.split(","):
        vae, zc, fs = build(name, dev)
        for res in args.res.split(","):
            Hp, Wp = RES[res]
            h, w = Hp // fs, Wp // fs
            for T in [int(t) for t in args.T.split(",")]:
                frames = 4 * (T - 1) + 1
                mpix = frames * h * fs * w * fs / 1e6
                rec = {"decoder": name, "latent": [zc, T, h, w], "frames": frames, "resolution": [h * fs, w * fs], "n_gpus": world}
                if name in rate and rate[name] * mpix > args.max_seconds:
                    rec["skipped"] = f"predicted {rate[name] * mpix:.0f} s > --max-seconds"
                else:
                    # whole clip while ~5 full-resolution bf16 activations (top level: 96 / 128 channels) fit in ~150 GB
                    top_c = 96 if name == "wan" else 128
                    live = {"wan": 5.0, "hyvae15": 3.0, "hyvae10": 4.5}[name]      # measured peaks: 106 / 72.5 / 114 GB at 129 frames x 720p
                    whole_bytes = live * frames * h * fs * w * fs * top_c * 2
                    if whole_bytes < 150e9:
                        mode = "whole"
                    elif name == "wan":
                        per_lat = 5.0 * 4 * h * fs * w * fs * top_c * 2
                        mode = ("streamed", max(3, int(60e9 // per_lat)))
                    else:
                        mode = "reference tiling"
                    rec["mode"] = mode if isinstance(mode, str) else f"streamed, {mode[1]} latent frames per slice"
                    try:
                        z = torch.randn(1, zc, T, h, w, generator=torch.Generator().manual_seed(rank)).to(dev)
                        torch.cuda.reset_peak_memory_stats()
                        if mpix < 400:                      # warm-up only where it is cheap
                            decode(name, vae, z, mode)
                        if dist is not None:
                            dist.barrier()
                        torch.cuda.synchronize()
                        l0 = _lib.launch_count()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        out = decode(name, vae, z, mode)
                        e1.record()
                        torch.cuda.synchronize()
                        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
                        if dist is not None:
                            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                        ms = float(ms[0])
                        rec.update({"ms_per_clip": ms, "frames_per_sec": world * frames / (ms / 1e3), "gpu_launches": _lib.launch_count() - l0,
                                    "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "finite": bool(torch.isfinite(out[:, :2]).all()),
                                    "out_shape": list(out.shape)})
                        rate[name] = ms / 1e3 / mpix
                        del out, z
                    except Exception as e:                  # noqa: BLE001
                        rec["error"] = repr(e)[:300]
                    torch.cuda.empty_cache()
                if rank == 0:
                    print(json.dumps(rec), flush=True)
                    logf.write(json.dumps(rec) + "\n"); logf.flush()
        del vae
        torch.cuda.empty_cache()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
