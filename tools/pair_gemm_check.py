"""CTA-pair GEMM (csrc/gemm2_sm100.cuh) vs the single-CTA kernel: parity on a few shapes, then A/B timings at the four
Wan2.2-14B linear-layer shapes.  The kernel choice is a process-wide switch (B200_GEMM_PAIR), so the A/B legs are subprocesses.
Usage: python tools/pair_gemm_check.py [check|time]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def leg():
    import torch
    from wan2gp_b200 import ops
    bf16, f32 = torch.bfloat16, torch.float32
    res = {"pair": os.environ.get("B200_GEMM_PAIR", "1"), "parity": [], "timing": []}
    g = torch.Generator(device="cuda").manual_seed(0)

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm())
    for M, N, K in [(512, 256, 64), (512, 256, 512), (1000, 768, 1536), (4096, 1536, 1536), (777, 512, 200), (2048, 1024, 4096), (75600, 256, 128)]:
        a = torch.randn(M, K, device="cuda", generator=g).to(bf16)
        b = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(bf16)
        bias, gate = torch.randn(N, device="cuda", generator=g), torch.randn(N, device="cuda", generator=g)
        ref = a.double() @ b.double().t() + bias.double()
        e_bf = rel(ops.gemm(a, b, bias=bias), ref)
        e_ge = rel(ops.gemm(a, b, bias=bias, act=1), torch.nn.functional.gelu(ref.float(), approximate="tanh"))
        x0 = torch.randn(M, N, device="cuda", generator=g)
        x = x0.clone()
        ops.gemm(a, b, out=x, bias=bias, gate=gate, accumulate=True)
        e_acc = rel(x, x0.double() + ref * gate.double())
        r = torch.randn(M, N, device="cuda", generator=g).to(bf16)
        e_res = rel(ops.gemm(a, b, bias=bias, residual=r), ref + r.double())
        res["parity"].append({"M": M, "N": N, "K": K, "bf16": e_bf, "gelu": e_ge, "acc_fp32": e_acc, "residual": e_res})
        print(res["parity"][-1], flush=True)
    if "time" in sys.argv:
        L, D, F = 75600, 5120, 13824

        def timeit(fn, iters=5, warm=2):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return sorted(ts)[len(ts) // 2]
        for name, N, K, kw in [("qkv", 3 * D, D, {}), ("o-proj(+gate,+=x)", D, D, {"acc": True}), ("ffn.0(+GELU)", F, D, {"act": 1}), ("ffn.2(+gate,+=x)", D, F, {"acc": True})]:
            A = torch.randn(L, K, device="cuda").to(bf16)
            w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(bf16)
            bias = torch.randn(N, device="cuda")
            if kw.get("acc"):
                x = torch.zeros(L, N, device="cuda", dtype=f32); gate = torch.randn(N, device="cuda")
                fn = lambda: ops.gemm(A, w, out=x, bias=bias, gate=gate, accumulate=True)
            else:
                out = torch.empty(L, N, device="cuda", dtype=bf16)
                fn = lambda: ops.gemm(A, w, out=out, bias=bias, act=kw.get("act", 0))
            ms = timeit(fn)
            res["timing"].append({"gemm": name, "M": L, "N": N, "K": K, "ms": ms, "tflops": 2.0 * L * N * K / ms / 1e9})
            print(res["timing"][-1], flush=True)
            del A, w
    print("LEG " + json.dumps(res))


if __name__ == "__main__":
    if os.environ.get("PAIR_LEG"):
        leg()
        sys.exit(0)
    out = {}
    for pair in ("1", "0"):
        env = dict(os.environ, PAIR_LEG="1", B200_GEMM_PAIR=pair)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, capture_output=True, text=True, timeout=420)
            line = [l for l in r.stdout.splitlines() if l.startswith("LEG ")]
            out["pair" if pair == "1" else "single"] = json.loads(line[-1][4:]) if line else {"error": (r.stdout + r.stderr)[-1500:]}
        except subprocess.TimeoutExpired as e:
            out["pair" if pair == "1" else "single"] = {"error": "timeout (hang?)", "tail": str(e.stdout)[-800:] if e.stdout else ""}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "pair_gemm.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
