"""Does operand major-ness (K-major vs MN-major UMMA descriptors) or the output type change
the tcgen05 GEMM's throughput?  wgrad-shaped problem: M = 33694 (features), N = 10108, K = batch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cellcomm_b200 import ops  # noqa: E402
from tools.gemm_bench import timeit  # noqa: E402


def main():
    M, N, K = 33694, 10108, 2048
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            a = ops.alloc2d(K, M) if a_mn else ops.alloc2d(M, K)
            b = ops.alloc2d(K, N) if b_mn else ops.alloc2d(N, K)
            a.normal_()
            b.normal_()
            for out in ("f32", "bf16"):
                o = ops.alloc2d(M, N, dtype=torch.float32 if out == "f32" else torch.bfloat16)
                kw = {"out32": o} if out == "f32" else {"out16": o}
                ms = timeit(lambda: ops.gemm(M, N, [a], [b], [K], a_mn, b_mn, use_ws=False, **kw),
                            flush=flush)
                print(json.dumps({"a_mn": a_mn, "b_mn": b_mn, "out": out, "ms": ms,
                                  "tflops": 2.0 * M * N * K / ms / 1e9}), flush=True)
                del o
            del a, b
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
