"""Time the un-tiled Hunyuan VAE decoders at the 720p x 129f clip size (rows H5/H6) -- prints one JSON line per decoder.
usage: python tools/hyvae_bench.py [hyvae10|hyvae15 ...] [--small]"""
import json
import sys
import time

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from wan2gp_b200 import _lib, synth                                    # noqa: E402
from wan2gp_b200.hyvideo import HYVAE10Decoder, HYVAE10Encoder, HYVAEDecoder, HYVAEEncoder            # noqa: E402


def run(name, small):
    dev = torch.device("cuda:0")
    if name == "hyvae10":
        cfg, dec = synth.HYVAE10_CONFIGS["hyvae10"], None
        zshape = (16, 9, 45, 80) if small else (16, 33, 90, 160)
        dec = HYVAE10Decoder(cfg, dev)
        dec.load_state_dict(synth.make_hyvae10_state_dict(cfg, 0, device=dev))
    else:
        cfg = synth.HYVAE_CONFIGS["hyvae15"]
        zshape = (32, 9, 22, 40) if small else (32, 33, 45, 80)
        dec = HYVAEDecoder(cfg, dev)
        dec.load_state_dict(synth.make_hyvae_state_dict(cfg, 0, device=dev))
    z = torch.randn(1, *zshape, generator=torch.Generator().manual_seed(0)).to(dev)
    res = {"decoder": name, "latent": list(zshape)}
    try:
        torch.cuda.reset_peak_memory_stats()
        out = dec(z)
        torch.cuda.synchronize()
        res["frames"], res["resolution"] = out.shape[2], list(out.shape[3:])
        res["finite"] = bool(torch.isfinite(out).all())
        res["absmean"] = float(out.abs().mean())
        del out
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        out = dec(z)
        e1.record()
        torch.cuda.synchronize()
        res["ms_per_clip"] = e0.elapsed_time(e1)
        res["wall_s"] = time.time() - t0
        res["frames_per_sec"] = out.shape[2] / (res["ms_per_clip"] / 1e3)
        res["gpu_launches"] = _lib.launch_count() - l0
        res["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    except Exception as e:                                               # noqa: BLE001
        res["error"] = repr(e)[:300]
    print(json.dumps(res), flush=True)
    del dec
    torch.cuda.empty_cache()


def run_encode(name, small):
    """Un-tiled encode of a 720p x 129f clip (or --small): one JSON line."""
    dev = torch.device("cuda:0")
    if name == "hyvae10":
        cfg = synth.HYVAE10_CONFIGS["hyvae10"]
        enc = HYVAE10Encoder(cfg, dev)
        enc.load_state_dict(synth.make_hyvae10_state_dict(cfg, 0, device=dev, encoder=True))
    else:
        cfg = synth.HYVAE_CONFIGS["hyvae15"]
        enc = HYVAEEncoder(cfg, dev)
        enc.load_state_dict(synth.make_hyvae_state_dict(cfg, 0, device=dev, encoder=True))
    shape = (1, 3, 33, 368, 640) if small else (1, 3, 129, 720, 1280)
    x = (torch.rand(*shape, generator=torch.Generator().manual_seed(0)) * 2 - 1).to(dev)
    res = {"encoder": name, "video": list(shape[1:])}
    try:
        torch.cuda.reset_peak_memory_stats()
        out = enc(x)
        torch.cuda.synchronize()
        res["moments"], res["finite"] = list(out.shape[1:]), bool(torch.isfinite(out).all())
        del out
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = enc(x)
        e1.record()
        torch.cuda.synchronize()
        res["ms_per_clip"] = e0.elapsed_time(e1)
        res["frames_per_sec"] = shape[2] / (res["ms_per_clip"] / 1e3)
        res["gpu_launches"] = _lib.launch_count() - l0
        res["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    except Exception as e:                                               # noqa: BLE001
        res["error"] = repr(e)[:300]
    print(json.dumps(res), flush=True)
    del enc
    torch.cuda.empty_cache()


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["hyvae10", "hyvae15"]
    for n in names:
        (run_encode if "--encode" in sys.argv else run)(n, "--small" in sys.argv)
