"""Fused wgrad + RMSprop epilogue: register path (CC_GEMM_RMS_TMA=0) vs TMA-state path, with the
bf16 copy by TMA store or by row stores, both rasters, at the two big layer shapes.
    python tools/rms_tma_bench.py [out.jsonl]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cellcomm_b200 import ops  # noqa: E402
from tools.gemm_bench import timeit  # noqa: E402

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None
for (K, N) in ((6738, 33694), (33694, 10108), (3369, 6738)):
    ld = ops.pad_ld(N)
    mk = lambda dt=torch.float32: torch.zeros(K, ld, dtype=dt, device="cuda")[:, :N]
    p32, ms, mom, p16 = mk(), mk(), mk(), mk(torch.bfloat16)
    rms = (p32, p16, ms, mom, 0.0075, 0.85, 0.1, 1e-7)
    for B in (128, 2048):
        x = ops.alloc2d(B, K); x.normal_()
        dz = ops.alloc2d(B, N); dz.normal_(std=1e-3)
        for tma, pair, inter, nfast, l2 in ((0, 0, 0, -1, 0), (0, 0, 0, -1, 1), (1, 0, 1, -1, 0),
                                            (1, 0, 1, -1, 1), (1, 0, 0, -1, 1)):
            os.environ["CC_GEMM_RMS_L2_256"] = str(l2)
            os.environ["CC_GEMM_RMS_TMA"] = str(tma)
            os.environ["CC_GEMM_RMS_PAIR"] = str(pair)
            os.environ["CC_GEMM_RMS_INTERLEAVE"] = str(inter)
            os.environ["CC_GEMM_RMS_NFAST"] = str(nfast)
            ops.reload_env()
            t = timeit(lambda: ops.dense_wgrad(x, dz, None, rms=rms))
            rec = {"K": K, "N": N, "batch": B, "tma_state": tma, "pair": pair, "interleave": inter,
                   "nfast": nfast, "l2_256": l2, "ms": t, "GB/s": 26.0 * K * N / t / 1e6}
            print(json.dumps(rec), flush=True)
            if out:
                out.write(json.dumps(rec) + "\n")
        del x, dz
    del p32, ms, mom, p16, rms
    torch.cuda.empty_cache()
