"""Bandwidth of the fused-RMSprop wgrad epilogue: batch 64 makes the MMA main loop negligible,
so the kernel time is the epilogue's 26 B/element stream; batch 2048 is the bench's regime.
Sweeps the tile raster / cache-hint / prefetch knobs of gemm_sm100.cu (read per launch).

    python tools/rms_epilogue_bench.py [out.jsonl]
"""
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cellcomm_b200 import ops  # noqa: E402
from tools.gemm_bench import timeit  # noqa: E402

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None


def emit(rec):
    line = json.dumps(rec)
    print(line, flush=True)
    if out:
        out.write(line + "\n")
        out.flush()


for (K, N) in ((6738, 33694), (33694, 10108)):
    ld = ops.pad_ld(N)
    mk = lambda dt=torch.float32: torch.zeros(K, ld, dtype=dt, device="cuda")[:, :N]
    p32, ms, mom, p16, dw = mk(), mk(), mk(), mk(torch.bfloat16), mk()
    rms = (p32, p16, ms, mom, 0.0075, 0.85, 0.1, 1e-7)
    for B in (64, 2048):
        x = ops.alloc2d(B, K); x.normal_()
        dz = ops.alloc2d(B, N); dz.normal_(std=1e-3)
        t = timeit(lambda: ops.dense_wgrad(x, dz, dw))
        emit({"K": K, "N": N, "batch": B, "kernel": "wgrad->fp32 grad", "ms": t,
              "GB/s": 4.0 * K * N / t / 1e6, "TFLOP/s": 2.0 * B * K * N / t / 1e9})
        t2 = timeit(lambda: ops.rmsprop_step(p32, p16, dw, ms, mom, 0.0075, 0.85, 0.1, 1e-7))
        emit({"K": K, "N": N, "batch": B, "kernel": "standalone rmsprop sweep", "ms": t2,
              "GB/s": 30.0 * K * N / t2 / 1e6, "unfused_total_ms": t + t2})
        for nfast, cs, pf in itertools.product((0, 1), (0, 1), (0, 1)):
            os.environ["CC_GEMM_RMS_NFAST"] = str(nfast)
            os.environ["CC_GEMM_RMS_CS"] = str(cs)
            os.environ["CC_GEMM_RMS_PREFETCH"] = str(pf)
            ops.reload_env()
            tf = timeit(lambda: ops.dense_wgrad(x, dz, None, rms=rms))
            emit({"K": K, "N": N, "batch": B, "kernel": "wgrad+fused rmsprop", "nfast": nfast,
                  "cs": cs, "prefetch": pf, "ms": tf, "GB/s": 26.0 * K * N / tf / 1e6,
                  "vs_unfused": (t + t2) / tf})
        del x, dz
    del p32, ms, mom, p16, dw, rms
    torch.cuda.empty_cache()
