"""Micro-benchmark of the hot kernels on one B200: the dominant Dense GEMMs of the Continuous
BiGAN in their three orientations, the RMSprop sweep and the batch gather.
CUDA-event timing, warm-up, inputs larger than L2 or an L2 flush between iterations.

    python tools/gemm_bench.py [batch] [out.json]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cellcomm_b200 import ops  # noqa: E402


def timeit(fn, iters=5, warmup=2, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "gemm_bench.json")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    res = []
    G = 33694
    layers = [("E1/fwd", G, 3369), ("Dx1/fwd", G, 10108), ("G6/fwd", 6738, G), ("G5/fwd", 3369, 6738),
              ("Dx3/fwd", 3369, 1684)]
    for name, K, N in layers:
        x = ops.alloc2d(B, K)
        x.normal_()
        w = ops.alloc2d(K, N)
        w.normal_(std=0.01)
        bias = torch.zeros(N, device="cuda")
        y = ops.alloc2d(B, N)
        dz = ops.alloc2d(B, N)
        dz.normal_()
        dx = ops.alloc2d(B, K)
        dw = ops.alloc2d(K, N, dtype=torch.float32)
        flops = 2.0 * B * K * N
        for kind, fn, extra_bytes in (
                ("fwd", lambda: ops.dense_fwd([x], w, [0], bias, 1, out16=y), 0),
                ("dgrad", lambda: ops.dense_dgrad([dz], [w], dx), 0),
                ("wgrad", lambda: ops.dense_wgrad(x, dz, dw), 0)):
            ms = timeit(fn, flush=flush)
            bytes_ = 2.0 * (B * K + K * N + B * N) + (2.0 * K * N if kind == "wgrad" else 0)
            rec = {"layer": name.split("/")[0], "kind": kind, "M": B, "K": K, "N": N, "ms": ms,
                   "tflops": flops / ms / 1e9, "gbs_min": bytes_ / ms / 1e6}
            print(json.dumps(rec), flush=True)
            res.append(rec)
        del x, w, y, dz, dx, dw
        torch.cuda.empty_cache()
    # RMSprop sweep over D's parameters (494.7 M): 28 B + 2 B per element
    n = 494_700_000 // 4 * 4
    p32 = torch.zeros(n, device="cuda")
    g = torch.randn(n, device="cuda")
    ms_, mom = torch.zeros_like(p32), torch.zeros_like(p32)
    p16 = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    t = timeit(lambda: ops.rmsprop_step(p32, p16, g, ms_, mom, 0.0075, 0.85, 0.1, 1e-7))
    rec = {"kernel": "rmsprop", "elems": n, "ms": t, "gbs": n * 30.0 / t / 1e6}
    print(json.dumps(rec), flush=True)
    res.append(rec)
    del p32, g, ms_, mom, p16
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
