"""Bandwidth of the peer-memory optimiser kernel (csrc/peer_optimizer.cu).

    python tools/peer_bench.py                       one GPU: world 1, and two emulated ranks
    torchrun --nproc-per-node N tools/peer_bench.py  N GPUs over NVLink (symmetric memory)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cellcomm_b200 import ops  # noqa: E402

HP = (0.0075, 0.85, 0.1, 1e-7)
n = 1 << 27          # 128 Mi elements per bucket (512 MB fp32)


def time_it(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


if "RANK" not in os.environ:
    dev = "cuda"
    p32, ms, mom = (torch.zeros(n, device=dev) for _ in range(3))
    p16 = torch.zeros(n, dtype=torch.bfloat16, device=dev)
    g = torch.randn(n, device=dev)
    t = time_it(lambda: ops.rmsprop_step(p32, p16, g, ms, mom, *HP))
    print(json.dumps({"kernel": "rmsprop_kernel (flat sweep)", "ms": t, "GB/s": 30.0 * n / t / 1e6}))
    flags = torch.full((16,), 2 ** 30, dtype=torch.int32, device=dev)
    t = time_it(lambda: ops.peer_rmsprop(1, 0, [g.data_ptr()], [p16.data_ptr()], p32, ms, mom, 0, n,
                                         True, *HP, flags.data_ptr(), 1))
    print(json.dumps({"kernel": "peer_rmsprop world=1", "ms": t, "GB/s": 30.0 * n / t / 1e6}))
    g2 = torch.randn(n, device=dev)
    q16 = torch.zeros(n, dtype=torch.bfloat16, device=dev)
    t = time_it(lambda: ops.peer_rmsprop(2, 0, [g.data_ptr(), g2.data_ptr()],
                                         [p16.data_ptr(), q16.data_ptr()], p32, ms, mom, 0, n // 2,
                                         True, *HP, flags.data_ptr(), 1))
    print(json.dumps({"kernel": "peer_rmsprop world=2 emulated on one GPU, half range", "ms": t,
                      "GB/s": 36.0 * (n // 2) / t / 1e6}))
else:
    import torch.distributed as dist
    from cellcomm_b200 import engine as eng
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    td = eng.TorchDist()
    g, g_ptrs, h1 = td.symmetric_zeros(n, torch.float32, dev)
    p16, w_ptrs, h2 = td.symmetric_zeros(n, torch.bfloat16, dev)
    flags, f_ptrs, h3 = td.symmetric_zeros(2 * world, torch.int32, dev)
    g.normal_()
    p32, ms, mom = (torch.zeros(n, device=dev) for _ in range(3))
    shard = n // world
    epoch = [0]

    def step():
        epoch[0] += 1
        ops.peer_signal([f_ptrs[t] + 4 * rank for t in range(world)], epoch[0])
        ops.peer_rmsprop(world, rank, g_ptrs, w_ptrs, p32, ms, mom, rank * shard, shard, True, *HP,
                         f_ptrs[rank], epoch[0])
        ops.peer_signal([f_ptrs[t] + 4 * (world + rank) for t in range(world)], epoch[0])
        ops.peer_wait(f_ptrs[rank] + 4 * world, world, epoch[0])

    dist.barrier()
    t = time_it(step)
    nv_in = 4.0 * shard * (world - 1)
    if rank == 0:
        print(json.dumps({"kernel": f"peer_rmsprop world={world} over NVLink", "elements": n,
                          "ms": t, "nvlink_in_GB/s": nv_in / t / 1e6,
                          "hbm_GB/s": (30.0 + 4 * (world - 1)) * shard / t / 1e6}), flush=True)
    # NCCL equivalent: reduce-scatter fp32 + flat sweep on the shard + all-gather bf16
    gs = torch.empty(shard, device=dev)
    sl = slice(rank * shard, (rank + 1) * shard)

    def nccl_step():
        dist.reduce_scatter_tensor(gs, g)
        ops.rmsprop_step(p32[sl], p16[sl], gs, ms[sl], mom[sl], *HP)
        dist.all_gather_into_tensor(p16, p16[sl])

    t2 = time_it(nccl_step)
    if rank == 0:
        print(json.dumps({"kernel": f"NCCL reduce-scatter + sweep + all-gather world={world}",
                          "ms": t2}), flush=True)
    dist.destroy_process_group()
