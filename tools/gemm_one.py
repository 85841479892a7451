"""Run one GEMM shape a few times (profiling target): python tools/gemm_one.py M N K a_mn b_mn out[f32|bf16] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cellcomm_b200 import ops  # noqa: E402

M, N, K, a_mn, b_mn = (int(v) for v in sys.argv[1:6])
out = sys.argv[6] if len(sys.argv) > 6 else "f32"
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 3
a = ops.alloc2d(K, M) if a_mn else ops.alloc2d(M, K)
b = ops.alloc2d(K, N) if b_mn else ops.alloc2d(N, K)
a.normal_()
b.normal_()
o = ops.alloc2d(M, N, dtype=torch.float32 if out == "f32" else torch.bfloat16)
kw = {"out32": o} if out == "f32" else {"out16": o}
for _ in range(iters):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.gemm(M, N, [a], [b], [K], a_mn, b_mn, use_ws=False, **kw)
    e.record()
    torch.cuda.synchronize()
    print(f"{s.elapsed_time(e):.4f} ms  {2.0 * M * N * K / s.elapsed_time(e) / 1e9:.1f} TFLOP/s")
