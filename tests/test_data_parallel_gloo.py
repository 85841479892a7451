"""Data-parallel host logic with world_size 2 over gloo on the CPU (kernels emulated): every
rank holds the full weights and half of each batch; BN batch statistics, parameter gradients
and the loss buffer are sum-all-reduced, so the result must equal the single-process oracle on
the whole batch (reference semantics: one global batch)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, variant, Z, G, B, out, bucket_elems=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if bucket_elems:     # read when cellcomm_b200.engine is imported
        os.environ["CELLCOMM_B200_BUCKET_ELEMS"] = str(bucket_elems)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ops_emulator
    from cellcomm_b200 import engine as eng
    from oracle import bigan_oracle as O
    import test_parity_gpu as P
    eng.ops = ops_emulator
    torch.set_num_threads(2)
    orc = O.OracleBiGan(variant, Z, G, seed=0, dtype=torch.float64)
    e = eng.BiGanEngine(variant, Z, G, max_batch=B // world, device="cpu", seed=0,
                        dist=eng.TorchDist())
    if bucket_elems:     # the big layers must really be cut into several row pieces / buckets
        assert max(len(v) for v in e.D.pieces.values()) >= 2 and len(e.D.buckets) >= 3
    for n in ("G", "E", "D"):
        e.nets[n].set_weights([w.numpy() for w in orc.get_weights(n)])
    x, z, r = P._inputs(variant, Z, G, B, 11)
    masks = O.make_masks(variant, Z, G, B, 3)
    lo, hi = rank * B // world, (rank + 1) * B // world
    local_masks = {s: {n: [m[lo:hi] for m in ms] for n, ms in d.items()} for s, d in masks.items()}
    x16 = ops_emulator.alloc2d(hi - lo, G)
    x16.copy_(x[lo:hi])
    e.set_latents(z[lo:hi], r[lo:hi], hi - lo)
    got = [float(v) for v in e.train_step(x16, local_masks)]
    weights = {n: e.nets[n].get_weights() for n in ("G", "E", "D")}
    if rank == 0:
        ref = orc.trainings_step(x, z, r, masks)
        ok = all(abs(a - b) <= 1e-4 * abs(b) + 1e-6 for a, b in zip(got, ref))
        worst = 0.0
        for n in ("G", "E", "D"):
            for a, b in zip(weights[n], orc.get_weights(n)):
                if b.numel():
                    worst = max(worst, float(np.abs(a - b.numpy()).max()))
        torch.save({"ok": ok, "got": got, "ref": list(ref), "worst_weight_err": worst}, out)
    # both ranks must end with identical weights
    flat = torch.cat([torch.as_tensor(w).flatten() for n in ("G", "E", "D") for w in weights[n]])
    other = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(other, flat)
    assert torch.equal(other[0], other[1])
    dist.destroy_process_group()


@pytest.mark.parametrize("variant,Z,G,B", [("cont", 3, 150, 12), ("classify", 4, 80, 8)])
def test_two_ranks_equal_one_global_batch(tmp_path, variant, Z, G, B):
    out = str(tmp_path / "res.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, variant, Z, G, B, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["ok"], res
    assert res["worst_weight_err"] <= 2e-4, res


def test_two_ranks_bucketed_row_chunks(tmp_path):
    """Same check with tiny gradient buckets: big kernels are cut along their rows, every
    bucket is reduce-scattered / updated / all-gathered on its own."""
    out = str(tmp_path / "res.pt")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, "cont", 3, 150, 12, out, 2048), nprocs=2, join=True)
    res = torch.load(out)
    assert res["ok"], res
    assert res["worst_weight_err"] <= 2e-4, res
