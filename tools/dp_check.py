"""torchrun -n N tools/dp_check.py : data-parallel sanity on real GPUs over NCCL.
Every rank trains on its own rows; after a few steps all ranks must hold bit-identical
weights (gradients / BN statistics were all-reduced), and the losses must be finite."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cellcomm_b200 import engine as eng, ops  # noqa: E402

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, G = 256, 3000
e = eng.BiGanEngine("cont", 3, G, max_batch=B, device="cuda", seed=0, dist=eng.TorchDist())
g = torch.Generator().manual_seed(100 + rank)
for step in range(3):
    x = ((torch.rand(B, G, generator=g) < 0.06).float() *
         (torch.poisson(torch.full((B, G), 1.2), generator=g) + 1))
    x16 = ops.alloc2d(B, G)
    x16.copy_(x)
    e.set_latents(torch.rand(B, 3, generator=g), torch.rand(B, 3, generator=g), B)
    losses = [float(v) for v in e.train_step(x16)]
    assert all(np.isfinite(losses)), losses
# the same step as ONE CUDA graph per rank (gather + priors + eight sub-steps + peer-memory
# gradient exchange / optimiser / all-reduces with device-resident epochs), replayed twice
graph_mode = "eager only"
if e.peer_graphable():
    from cellcomm_b200.cell_type_training import CellMatrix
    dense = ((torch.rand(4 * B, G, generator=g) < 0.06).float() *
             (torch.poisson(torch.full((4 * B, G), 1.2), generator=g) + 1)).numpy().astype(np.float64)
    csr = CellMatrix.from_dense(dense).device_csr("cuda")
    gs = e.capture_step(csr, G, B, latents="device")
    for step in range(2):
        idx = torch.from_numpy(np.random.RandomState(rank * 10 + step).permutation(4 * B)[:B])
        losses = [float(v) for v in gs.replay(idx)]
        assert all(np.isfinite(losses)), losses
    graph_mode = f"+ 2 CUDA-graph replays ({gs.launches_per_replay} kernels each)"
e.join()
torch.cuda.synchronize()
for name, n in e.nets.items():
    c16 = n.p16.clone()
    r16 = c16.clone()
    dist.broadcast(r16, src=0)
    assert torch.equal(c16, r16), f"rank {rank}: {name} bf16 compute copies differ from rank 0"
    n.gather_master()          # sharded optimiser: fp32 master is current only on its owner
    mine = n.p32.clone()
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(mine, ref), f"rank {rank}: {name} weights differ from rank 0"
    bn = torch.cat([L["moving_mean"] for L in n.layers if L["kind"] == "bn"] or [torch.zeros(1, device="cuda")])
    refbn = bn.clone()
    dist.broadcast(refbn, src=0)
    assert torch.equal(bn, refbn), f"rank {rank}: {name} BN stats differ"
if rank == 0:
    pr = e.G.peer
    mode = "NCCL reduce-scatter / all-gather" if pr is None else (
        "peer-memory push + fused optimiser, all-gather by " +
        ("NVLS multicast stores" if pr["p16_mc"] else "P2P stores"))
    print(f"dp_check ok: world={world} losses={losses} [{mode}] [{graph_mode}]")
dist.destroy_process_group()
