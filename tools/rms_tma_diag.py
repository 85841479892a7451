"""Diagnostic for the TMA-state fused RMSprop epilogue: which padding elements does the bf16
TMA store touch when the tensor's inner extent is not a multiple of 8 elements?
    CC_GEMM_RMS_P16_TMA=1 python tools/rms_tma_diag.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cellcomm_b200 import ops  # noqa: E402

for K, N, B in ((64, 300, 40), (64, 2049, 40), (64, 1300, 40), (64, 296, 40), (64, 290, 40)):
    ld = ops.pad_ld(N) + 64
    x = ops.alloc2d(B, K); x.normal_()
    dz = ops.alloc2d(B, N); dz.normal_(std=1e-2)
    mk = lambda dt: torch.full((K, ld), 7.0, dtype=dt, device="cuda")
    p32, ms, mom, p16 = mk(torch.float32), mk(torch.float32), mk(torch.float32), mk(torch.bfloat16)
    ops.dense_wgrad(x, dz, None, rms=(p32[:, :N], p16[:, :N], ms[:, :N], mom[:, :N], 0.0075, 0.85, 0.1, 1e-7))
    torch.cuda.synchronize()
    for name, t in (("p32", p32), ("ms", ms), ("mom", mom), ("p16", p16)):
        pad = t[:, N:].float()
        touched = (pad != 7.0)
        cols = sorted(set((torch.nonzero(touched)[:, 1] + N).tolist()))
        rows = sorted(set(torch.nonzero(touched)[:, 0].tolist()))
        vals = sorted(set(pad[touched].tolist()))[:5]
        print(f"K={K} N={N} ld={ld} {name}: touched padding cols {cols[:12]}{'...' if len(cols) > 12 else ''} "
              f"rows {len(rows)} values {vals}")
