"""Bandwidth of the fused-RMSprop wgrad epilogue in isolation: batch 64 makes the MMA main
loop negligible, so the kernel time is the epilogue's 26 B/element stream."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cellcomm_b200 import ops  # noqa: E402
from tools.gemm_bench import timeit  # noqa: E402

K, N = 6738, 33694
for B in (64, 2048):
    x = ops.alloc2d(B, K); x.normal_()
    dz = ops.alloc2d(B, N); dz.normal_(std=1e-3)
    ld = ops.pad_ld(N)
    mk = lambda dt=torch.float32: torch.zeros(K, ld, dtype=dt, device="cuda")[:, :N]
    p32, ms, mom, p16, dw = mk(), mk(), mk(), mk(torch.bfloat16), mk()
    rms = (p32, p16, ms, mom, 0.0075, 0.85, 0.1, 1e-7)
    for name, fn, bytes_ in (
            ("wgrad->fp32 grad", lambda: ops.dense_wgrad(x, dz, dw), 4.0 * K * N),
            ("wgrad+fused rmsprop", lambda: ops.dense_wgrad(x, dz, None, rms=rms), 26.0 * K * N),
            ("standalone rmsprop sweep", lambda: ops.rmsprop_step(p32, p16, dw, ms, mom, 0.0075, 0.85, 0.1, 1e-7), 30.0 * K * N)):
        t = timeit(fn)
        print(json.dumps({"batch": B, "kernel": name, "ms": t, "GB/s": bytes_ / t / 1e6}), flush=True)
