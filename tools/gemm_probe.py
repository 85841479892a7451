"""Probe the UMMA shared-memory descriptor parameters on a real B200.

Runs each (orientation, LBO, SBO, K-step) hypothesis in its own subprocess (a bad descriptor
can fault the context) and reports which ones reproduce torch's fp32 matmul.  Used once to
confirm the MN-major descriptor fields in csrc/gemm_sm100.cu; kept for regression triage.

    python tools/gemm_probe.py            # sweep
    python tools/gemm_probe.py one a_mn b_mn M N K bn   # single case, env overrides apply
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(a_mn, b_mn, M, N, K, bn):
    import torch
    from cellcomm_b200 import ops
    g = torch.Generator().manual_seed(0)
    a = torch.randn((K, M) if a_mn else (M, K), generator=g).to(torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), generator=g).to(torch.bfloat16)
    da = ops.alloc2d(*a.shape); da.copy_(a)
    db = ops.alloc2d(*b.shape); db.copy_(b)
    out = ops.alloc2d(M, N, dtype=torch.float32)
    ops.gemm(M, N, [da], [db], [K], a_mn, b_mn, out32=out, bn=bn, use_ws=False)
    torch.cuda.synchronize()
    A = a.float().t() if a_mn else a.float()
    B = b.float() if b_mn else b.float().t()
    ref = A @ B
    err = (out.cpu() - ref).abs()
    ok = (err <= 2e-3 * K ** 0.5 + 1e-2 * ref.abs()).float().mean().item()
    print(json.dumps({"max_err": err.max().item(), "frac_ok": ok,
                      "row_ok": (err.max(1).values < 0.05 * K ** 0.5).float().mean().item(),
                      "col_ok": (err.max(0).values < 0.05 * K ** 0.5).float().mean().item()}))


def sweep():
    cases = []
    # K-major both (dgrad orientation)
    for lbo in (16, 0):
        cases.append(("KK", 0, 0, {"CC_GEMM_K_LBO": lbo}))
    # MN-major hypotheses for B (fwd) and A+B (wgrad)
    for (lbo, sbo, kstep) in ((8192, 1024, 2048), (1024, 8192, 2048), (8192, 1024, 256),
                              (128, 1024, 2048), (8192, 128, 2048)):
        env = {"CC_GEMM_MN_LBO": lbo, "CC_GEMM_MN_SBO": sbo, "CC_GEMM_MN_KSTEP": kstep}
        cases.append(("K-MN", 0, 1, env))
        cases.append(("MN-MN", 1, 1, env))
    results = []
    for name, a_mn, b_mn, envo in cases:
        for (M, N, K, bn) in ((128, 128, 64, 128), (128, 256, 256, 256), (200, 300, 500, 256)):
            env = dict(os.environ)
            env.update({k: str(v) for k, v in envo.items()})
            cmd = [sys.executable, os.path.abspath(__file__), "one", str(a_mn), str(b_mn), str(M),
                   str(N), str(K), str(bn)]
            try:
                res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=180)
                line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
                out = json.loads(line) if line.startswith("{") else {
                    "error": (res.stderr or res.stdout)[-400:]}
            except subprocess.TimeoutExpired:
                out = {"error": "timeout"}
            rec = {"case": name, "env": envo, "shape": [M, N, K, bn], **out}
            print(json.dumps(rec), flush=True)
            results.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_probe.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one(*[int(v) for v in sys.argv[2:8]])
    else:
        sweep()
