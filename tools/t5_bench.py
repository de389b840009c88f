"""Time the umT5-XXL text encoder (wan2gp_b200/wan/t5.py) on one prompt of 512 token ids with random weights: one JSON line.
The reference runs this encoder on the CPU (any2video.py:125 builds it with device='cpu'); it sits in front of the denoise path."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wan2gp_b200 import _lib, synth                    # noqa: E402
from wan2gp_b200.wan.t5 import umt5_xxl_encoder        # noqa: E402

dev = torch.device("cuda:0")
cfg = synth.T5_CONFIGS["umt5_xxl"]
g = torch.Generator(device=dev).manual_seed(0)
sd = {}
for name, shape in synth.t5_param_shapes(cfg).items():
    std = 1.0 if name == "token_embedding.weight" else 0.5 if "pos_embedding" in name else 0.1 if "norm" in name else shape[-1] ** -0.5
    t = torch.randn(shape, device=dev, generator=g, dtype=torch.float32 if len(shape) == 1 or "pos_embedding" in name else torch.bfloat16) * std
    sd[name] = t + (1.0 if "norm" in name else 0.0)
enc = umt5_xxl_encoder(dev)
enc.load_state_dict(sd)
del sd
ids = torch.randint(1, cfg["vocab_size"], (512,), device=dev)
for n_valid in (512, 60):
    enc.encode_one(ids, n_valid)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    reps = 5
    for _ in range(reps):
        out = enc.encode_one(ids, n_valid)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 24 * 2.0 * 512 * (4096 * 3 * 4096 + 4096 * 4096 + 3 * 4096 * 10240) + 24 * 4.0 * 512 * 512 * 4096
    print(json.dumps({"encoder": "umt5_xxl", "tokens": 512, "n_valid": n_valid, "ms_per_prompt": ms, "wall_ms": (time.perf_counter() - t0) * 1e3 / reps,
                      "tflops": flops / ms / 1e9, "weights_gb": 24 * 2 * (4 * 4096 * 4096 + 3 * 4096 * 10240) / 1e9,
                      "weight_stream_gbs": 24 * 2 * (4 * 4096 * 4096 + 3 * 4096 * 10240) / ms / 1e6,
                      "gpu_launches": (_lib.launch_count() - l0) // reps, "finite": bool(torch.isfinite(out).all())}))
