"""Attention kernel variants A/B: parity (sampled rows vs fp64) + CUDA-event timing at the Wan 720p / 480p self-attention shapes.
One subprocess per variant (the kernel choice is a process-wide env switch), each under a timeout so a hang cannot eat the lease.
Writes gpurun_out/attn_ab.json.   Usage: python tools/attn_ab.py [variant ...]   (default: v1 v100 v104 v103 v102; p0/p1/p2/p4 = the CTA-pair kernel family)"""
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NAMES = {"p0": "single-CTA (r01 kernel)", "p1": "pair, packed-fp32 softmax", "p2": "pair, scalar softmax", "p4": "pair, packed + 1/4 exp2 on FMA pipe",
         "v1": "single-CTA r01 kernel (scalar softmax, all exp2 on MUFU)", "v100": "single-CTA, packed softmax", "v104": "single-CTA, packed + 1/4 of the exp2 pairs on the FMA pipe",
         "v103": "single-CTA, packed + 1/3 on the FMA pipe", "v102": "single-CTA, packed + 1/2 on the FMA pipe",
         "v500": "one Q tile per CTA, double-buffered scores, K/V multicast over a 2-CTA cluster, 8 softmax warps (16x256b)",
         "v503": "same + 1/3 of the exp2 pairs on the FMA pipe", "v504": "same + 1/4", "v502": "same + 1/2",
         "v600": "one Q tile per CTA, three score buffers, two softmax warpgroups alternating over the K/V tiles, fixed reference maximum",
         "v603": "same + 1/3 of the exp2 pairs on the FMA pipe", "v604": "same + 1/4",
         "v613": "three score buffers, no per-tile max pass, 1/3 poly, both P halves published after one wait", "v614": "same, 1/4 poly", "v612": "same, 1/2 poly",
         "v901": "ABLATION of v103: exponentials replaced by a move", "v902": "ABLATION: half of each S row read from TMEM",
         "v903": "ABLATION: no softmax (MMA / smem / barrier ceiling)", "v904": "ABLATION: S read from TMEM, nothing computed or stored",
         "v905": "ABLATION: full arithmetic, P never stored"}


def leg():
    import torch
    from wan2gp_b200 import ops
    bf16 = torch.bfloat16
    out = {"variant": os.environ.get("ATT_LEG"), "parity": [], "timing": []}
    g = torch.Generator(device="cuda").manual_seed(0)
    # v9xx = limiter ablations (csrc/attn_sm100.cuh ABL): wrong output by construction, timing only
    ablation = out["variant"].startswith("v9")
    out["ablation"] = ablation
    # boost > 1: keys beyond the first 200 are scaled up so that later scores outrun the first tile's row maximum by far more than 2^60
    # (the repeat pass of attn6_sm100.cuh; the lazy rescale of the other kernels)
    for Lq, Lk, H, boost in ([] if ablation else [(1024, 1024, 2, 1), (2000, 1333, 3, 1), (1100, 512, 2, 1), (9000, 9000, 2, 1), (100, 100, 1, 1),
                                                   (300, 256, 1, 1), (129, 384, 2, 1), (700, 1500, 2, 60), (700, 1500, 2, -132), (640, 900, 1, -133)]):
        D = H * 128
        q, k, v = (torch.randn(n, D, device="cuda", generator=g).to(bf16) for n in (Lq, Lk, Lk))
        if boost > 1:
            k[200:] *= boost
        elif boost < 0:                  # ONE key (at a position whose exponential runs on the FMA pipe, resp. on MUFU) far above everything
            k[-boost] *= 80
        o = ops.attention(q, k, v, H)
        torch.cuda.synchronize()
        qh, kh, vh = (t.double().reshape(-1, H, 128).permute(1, 0, 2) for t in (q, k, v))
        ref = (torch.softmax(qh @ kh.transpose(1, 2) / math.sqrt(128), -1) @ vh).permute(1, 0, 2).reshape(Lq, D)
        rel = float((o.double() - ref).norm() / ref.norm())
        out["parity"].append({"Lq": Lq, "Lk": Lk, "H": H, "boost": boost, "rel_l2": rel, "ok": rel < 4e-3})
        print(out["parity"][-1], flush=True)
    if all(p["ok"] for p in out["parity"]):
        for L, H in ([(75600, 40)] if ablation else [(75600, 40), (32760, 40)]):
            D = H * 128
            qkv = torch.randn(L, 3 * D, device="cuda").to(bf16)
            o = torch.empty(L, D, device="cuda", dtype=bf16)
            fn = lambda: ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], H, out=o)
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(6 if ablation else 4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = sorted(ts)[len(ts) // 2]
            out["timing"].append({"L": L, "H": H, "ms": ms, "tflops": 4.0 * L * L * D / ms / 1e9, "all_ms": ts})
            print(out["timing"][-1], flush=True)
            del qkv, o
    print("LEG " + json.dumps(out))


if __name__ == "__main__":
    if os.environ.get("ATT_LEG"):
        leg()
        sys.exit(0)
    res = {}
    for i, var in enumerate(sys.argv[1:] or ["v1", "v100", "v104", "v103", "v102"]):
        # "pN": B200_ATT_PAIR=N (pair kernel family); "vN": single-CTA kernel, B200_ATT_VARIANT=N
        env = dict(os.environ, ATT_LEG=var, B200_ATT_PAIR=var[1:] if var[0] == "p" else "0", B200_ATT_VARIANT=var[1:] if var[0] == "v" else "1")
        key = var if var not in res else f"{var}#{i}"              # a variant listed twice (drift check) keeps both results
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=240)
            line = [l for l in r.stdout.splitlines() if l.startswith("LEG ")]
            res[key] = dict(json.loads(line[-1][4:]), name=NAMES.get(var, var)) if line else {"name": NAMES.get(var, var), "error": (r.stdout + r.stderr)[-1500:]}
        except subprocess.TimeoutExpired:
            res[key] = {"name": NAMES.get(var, var), "error": "timeout (hang)"}
        print(key, json.dumps(res[key])[:600], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", os.environ.get("ATT_AB_OUT", "attn_ab.json")), "w"), indent=1)
