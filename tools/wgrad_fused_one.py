"""Profiling target: the fused-RMSprop wgrad GEMM on one layer shape.
    python tools/wgrad_fused_one.py K N batch[,batch...] [iters] [--blocked] [--row0=R]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cellcomm_b200 import ops  # noqa: E402

blocked = "--blocked" in sys.argv
row0_arg = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--row0=")), 0)
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
K, N = int(argv[0]), int(argv[1])
batches = [int(v) for v in argv[2].split(",")]
iters = int(argv[3]) if len(argv) > 3 else 3
ld = ops.pad_ld(N)
K32 = (K + row0_arg + 31) // 32 * 32      # the layer: row0 rows of another segment first
mk = lambda dt=torch.float32: torch.zeros(K32, ld, dtype=dt, device="cuda")[:K, :N]
p16 = mk(torch.bfloat16)
if blocked:     # the layer's flat blocked state arrays (cc_gemm_desc.rms_blocked)
    flat = lambda: torch.zeros(K32 * ld, device="cuda")
    rms, row0 = (flat(), p16, flat(), flat(), 0.0075, 0.85, 0.1, 1e-7), row0_arg
else:
    rms, row0 = (mk(), p16, mk(), mk(), 0.0075, 0.85, 0.1, 1e-7), None
for B in batches:
    x = ops.alloc2d(B, K); x.normal_()
    dz = ops.alloc2d(B, N); dz.normal_(std=1e-3)
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.dense_wgrad(x, dz, None, rms=rms, rms_row0=row0)
        e.record()
        torch.cuda.synchronize()
        t = s.elapsed_time(e)
        print(f"batch {B}: {t:.4f} ms  {26.0 * K * N / t / 1e6:.0f} GB/s  {2.0 * B * K * N / t / 1e9:.0f} TFLOP/s")
