"""Fused wgrad + RMSprop register epilogue: row-major optimiser state vs the blocked (32 x 32
blocks of 4 KB) state layout (CC_GEMM_RMS_BLOCKED), timing only.
    python tools/rms_blocked_bench.py [out.jsonl]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cellcomm_b200 import ops  # noqa: E402
from tools.gemm_bench import timeit  # noqa: E402

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None
for (K, N) in ((6738, 33694), (33694, 10108), (3369, 6738), (33694, 3369)):
    ld = ops.pad_ld(N)
    K32 = (K + 31) // 32 * 32
    mk = lambda dt=torch.float32: torch.zeros(K32, ld, dtype=dt, device="cuda")[:K, :N]
    p32, ms, mom, p16 = mk(), mk(), mk(), mk(torch.bfloat16)
    rms = (p32, p16, ms, mom, 0.0075, 0.85, 0.1, 1e-7)
    flat = lambda: torch.zeros(K32 * ld, device="cuda")
    rms_b = (flat(), p16, flat(), flat(), 0.0075, 0.85, 0.1, 1e-7)
    for B in (128, 2048):
        x = ops.alloc2d(B, K); x.normal_()
        dz = ops.alloc2d(B, N); dz.normal_(std=1e-3)
        for tma, blocked in ((0, 0), (0, 1), (-1, 0)):
            os.environ["CC_GEMM_RMS_TMA"] = str(tma)
            ops.reload_env()
            if blocked:
                t = timeit(lambda: ops.dense_wgrad(x, dz, None, rms=rms_b, rms_row0=0))
            else:
                t = timeit(lambda: ops.dense_wgrad(x, dz, None, rms=rms))
            rec = {"K": K, "N": N, "batch": B, "tma_state": tma, "blocked": blocked, "ms": t,
                   "GB/s": 26.0 * K * N / t / 1e6}
            print(json.dumps(rec), flush=True)
            if out:
                out.write(json.dumps(rec) + "\n")
        del x, dz
    del p32, ms, mom, p16, rms, rms_b
    torch.cuda.empty_cache()
