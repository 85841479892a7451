"""Profiling target: the fused-RMSprop wgrad GEMM on one layer shape.
    python tools/wgrad_fused_one.py K N batch [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cellcomm_b200 import ops  # noqa: E402

K, N, B = (int(v) for v in sys.argv[1:4])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ld = ops.pad_ld(N)
mk = lambda dt=torch.float32: torch.zeros(K, ld, dtype=dt, device="cuda")[:, :N]
p32, ms, mom, p16 = mk(), mk(), mk(), mk(torch.bfloat16)
x = ops.alloc2d(B, K); x.normal_()
dz = ops.alloc2d(B, N); dz.normal_(std=1e-3)
rms = (p32, p16, ms, mom, 0.0075, 0.85, 0.1, 1e-7)
for _ in range(iters):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.dense_wgrad(x, dz, None, rms=rms)
    e.record()
    torch.cuda.synchronize()
    t = s.elapsed_time(e)
    print(f"{t:.4f} ms  {26.0 * K * N / t / 1e6:.0f} GB/s  {2.0 * B * K * N / t / 1e9:.0f} TFLOP/s")
