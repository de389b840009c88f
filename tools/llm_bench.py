"""Time the Qwen2.5-VL-7B language tower (wan2gp_b200/hyvideo/llm.py) on one prompt with random weights: one JSON line per prompt length.
The reference runs this model through transformers inside its TextEncoder (text_encoder_1_5.py:470-476); it sits in front of the denoise path."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wan2gp_b200 import _lib, synth                              # noqa: E402
from wan2gp_b200.hyvideo.llm import LlamaLikeTextModel           # noqa: E402

dev = torch.device("cuda:0")
cfg = dict(synth.LLM_CONFIGS["qwen25_vl_7b"], vocab_size=32768)  # the embedding table is only gathered from: a smaller one saves 0.4 GB of random numbers
g = torch.Generator(device=dev).manual_seed(0)
sd = {}
for name, shape in synth.llm_param_shapes(cfg).items():
    one_d = len(shape) == 1
    std = 0.1 if one_d else (1.0 if name == "embed_tokens.weight" else shape[-1] ** -0.5)
    t = torch.randn(shape, device=dev, generator=g, dtype=torch.float32 if one_d else torch.bfloat16) * std
    sd[name] = t + (1.0 if name.endswith("norm.weight") or name.endswith("layernorm.weight") else 0.0)
m = LlamaLikeTextModel.from_state_dict(sd, cfg["num_heads"], cfg["num_kv_heads"], cfg["rms_eps"], cfg["rope_theta"], device=dev)
del sd
D, Fi, H, Hk, NL = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_heads"], cfg["num_kv_heads"], cfg["num_layers"]
wbytes = NL * 2 * (D * (H + 2 * Hk) * 128 + H * 128 * D + 3 * D * Fi)
for n in (620, 128):                                             # text_len 512 + the template's ~108 tokens; a short prompt
    ids = torch.randint(1, cfg["vocab_size"], (n,), device=dev)
    m.hidden_states_one(ids)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    reps = 3
    for _ in range(reps):
        hs = m.hidden_states_one(ids)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = NL * (2.0 * n * (D * (H + 2 * Hk) * 128 + H * 128 * D + 3 * D * Fi) + 4.0 * n * n / 2 * H * 128)
    print(json.dumps({"encoder": "qwen2.5-vl-7b language tower (28 layers)", "tokens": n, "ms_per_prompt": ms, "wall_ms": (time.perf_counter() - t0) * 1e3 / reps,
                      "tflops": flops / ms / 1e9, "weights_gb": wbytes / 1e9, "weight_stream_gbs": wbytes / ms / 1e6,
                      "gpu_launches": (_lib.launch_count() - l0) // reps, "finite": bool(torch.isfinite(hs[-3]).all())}))
