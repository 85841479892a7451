"""Blocked-state fused epilogue: cluster on/off x raster, at the big layer shapes.
    python tools/rms_blocked_sweep.py [out.jsonl]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cellcomm_b200 import ops  # noqa: E402
from tools.gemm_bench import timeit  # noqa: E402

out = open(sys.argv[1], "w") if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else None
for (K, N) in ((6738, 33694), (33694, 10108), (33694, 3369)):
    ld = ops.pad_ld(N)
    K32 = (K + 31) // 32 * 32
    p16 = torch.zeros(K32, ld, dtype=torch.bfloat16, device="cuda")[:K, :N]
    flat = lambda: torch.zeros(K32 * ld, device="cuda")
    rms_b = (flat(), p16, flat(), flat(), 0.0075, 0.85, 0.1, 1e-7)
    for B in ((512, 1024, 2048, 4096) if "--keep" in sys.argv else (128, 2048)):
        x = ops.alloc2d(B, K); x.normal_()
        dz = ops.alloc2d(B, N); dz.normal_(std=1e-3)
        for cluster, nfast, keep in ((-1, -1, 1), (-1, -1, 0), (-1, 0, 1), (-1, 1, 1)) if "--keep" in sys.argv \
                else [(c, n, 1) for c in (1, 0) for n in (-1, 0, 1)]:
            if True:
                os.environ["CC_GEMM_RMS_KEEP_OPERANDS"] = str(keep)
                os.environ["CC_GEMM_RMS_CLUSTER"] = str(cluster)
                os.environ["CC_GEMM_RMS_NFAST"] = str(nfast)
                ops.reload_env()
                t = timeit(lambda: ops.dense_wgrad(x, dz, None, rms=rms_b, rms_row0=0))
                rec = {"K": K, "N": N, "batch": B, "cluster": cluster, "nfast": nfast, "keep_operands": keep, "ms": t,
                       "GB/s": 26.0 * K * N / t / 1e6}
                print(json.dumps(rec), flush=True)
                if out:
                    out.write(json.dumps(rec) + "\n")
        del x, dz
    del p16, rms_b
    torch.cuda.empty_cache()
