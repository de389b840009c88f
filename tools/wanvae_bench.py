"""Time the Wan VAE decode (row V1-V7) at the 720p x 81f clip size without loading a DiT: one JSON line."""
import json
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from wan2gp_b200 import _lib, synth            # noqa: E402
from wan2gp_b200.wan import WanVAE              # noqa: E402

dev = torch.device("cuda:0")
T, H, W = (21, 90, 160) if "--small" not in sys.argv else (5, 45, 80)
vae = WanVAE(device=dev, state_dict=synth.make_vae_state_dict(seed=0, encoder=True))
z = torch.randn(16, T, H, W, generator=torch.Generator().manual_seed(0)).to(dev)
vae.model.decode_frames(z, vae.mean, vae.std)
torch.cuda.synchronize()
l0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 3
for _ in range(reps):
    fr = vae.model.decode_frames(z, vae.mean, vae.std)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nfr = 4 * (T - 1) + 1
print(json.dumps({"decoder": "wan_vae", "latent": [16, T, H, W], "frames": nfr, "ms_per_clip": ms, "frames_per_sec": nfr / (ms / 1e3),
                  "gpu_launches": (_lib.launch_count() - l0) // reps, "finite": bool(torch.isfinite(fr).all())}))

if "--encode" in sys.argv:
    vid = (torch.rand(3, nfr, 8 * H, 8 * W, generator=torch.Generator().manual_seed(1)) * 2 - 1).to(dev)
    vae.encode([vid], tile_size=0)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0.record()
    mu = vae.encode([vid], tile_size=0)[0]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"encoder": "wan_vae", "video": [3, nfr, 8 * H, 8 * W], "latent": list(mu.shape), "ms_per_clip": ms,
                      "frames_per_sec": nfr / (ms / 1e3), "gpu_launches": _lib.launch_count() - l0, "finite": bool(torch.isfinite(mu).all())}))
