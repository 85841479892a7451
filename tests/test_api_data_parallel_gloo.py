"""The PRODUCT entry points under data parallel, world_size 2 over gloo on the CPU (kernels
emulated by tests/ops_emulator.py): `CellTraining.run` on every rank with rank-0-only
interceptors (DbRecorder on the in-memory Mongo, EncodingFiles, Checkpoints), ONE global
`permutation(N)[:B]` split contiguously over the ranks, global priors, the row-sharded
encode-all-cells pass gathered to rank 0, and checkpoint gathers of the sharded optimiser
state -- compared with the same run in a single process (reference call sites
src/__main__.py:44-66, src/cell_type_training.py:40-50, src/intercepts/db_recorder.py:82-108).
"""
import os
import pickle
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")


def _write_source(tmp, N, G, seed):
    rng = np.random.default_rng(seed)
    dense = (rng.random((N, G)) < 0.3) * rng.integers(1, 9, (N, G))
    dense[np.arange(N), rng.integers(0, G, N)] += 1          # every barcode has an entry
    dense[rng.integers(0, N, G), np.arange(G)] += 1          # every gene occurs
    src = {k: os.path.join(tmp, f"s_{k}.{'mtx' if k == 'matrix' else 'tsv'}")
           for k in ("matrix", "barcodes", "genes")}
    with open(src["matrix"], "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n%\n0 0 0\n")
        for b in range(N):
            for g in np.flatnonzero(dense[b]):
                f.write(f"{g + 1} {b + 1} {int(dense[b, g])}\n")
    with open(src["barcodes"], "w") as f:
        f.writelines(f"BC{b:04d}-1\n" for b in range(N))
    with open(src["genes"], "w") as f:
        f.writelines(f"ENS{g:05d}\tsym{g}\n" for g in range(G))
    return src


def _run(rank, world, port, tmp, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CELLCOMM_B200_DEVICE="cpu",
                      RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import ops_emulator
    from cellcomm_b200 import engine as eng, intercepts
    from cellcomm_b200.cell_type_training import CellTraining, load_matrix
    from cellcomm_b200.intercepts import db_recorder as dbr
    from cellcomm_b200.intercepts.fake_mongo import MongoClient as FakeMongo
    eng.ops = ops_emulator
    src = {k: os.path.join(tmp, f"s_{k}.{'mtx' if k == 'matrix' else 'tsv'}")
           for k in ("matrix", "barcodes", "genes")}
    data = load_matrix(src["matrix"])
    np.random.seed(100 + rank)                 # run() must replace this with rank 0's state
    trainer = CellTraining(data, batch_size=8, encoding_size=3, batches_per_iteration=2)
    net = trainer.network
    assert net._world() == world
    # identical initial weights: a deterministic re-initialisation on every rank
    for name, n in net._engine.nets.items():
        gen = torch.Generator().manual_seed(7 + len(name))
        n.set_weights([(torch.rand(w.shape, generator=gen) - 0.5).numpy() * 0.2 if i % 2 == 0
                       else w for i, w in enumerate(n.get_weights())])
    net._engine.rng_seed = 4321
    seen = []
    icpt = None
    if rank == 0:
        np.random.seed(5)
        net._prior_rng = np.random.default_rng(9)
        log_dir = os.path.join(tmp, f"logs_w{world}")
        rec = dbr.DbRecorder("run", src, client_factory=FakeMongo)
        rec.setup()
        ck = intercepts.Checkpoints(log_dir)
        icpt = intercepts.combined_interceptors((
            lambda it, l: seen.append((it, [float(v) for v in l])),
            rec.create_interceptor(trainer),
            intercepts.EncodingFiles(log_dir).create_interceptor(trainer),
            ck.create_interceptor(trainer)))
    # dropout off in this comparison: the Philox streams are keyed by rank, so masks differ
    # between a 1-rank and a 2-rank run by design (priors and batches do not)
    for n in net._engine.nets.values():
        for node in n.g.nodes:
            if node["kind"] == "dropout":
                node["rate"] = 0.0
    trainer.run(2, icpt)
    if rank == 0:
        docs = FakeMongo(dbr.MONGO_URL)[dbr.MONGO_DB][dbr.ITERATIONS_COLLECTION].find({"eid": "run"})
        with open(os.path.join(log_dir, "encodings", "1.enc"), "rb") as f:
            enc_file = pickle.load(f)
        ckpt = dict(np.load(os.path.join(log_dir, "checkpoint.npz")))
        torch.save({"seen": seen, "xs": [d["xs"] for d in docs], "its": [d["it"] for d in docs],
                    "enc": enc_file, "ckpt": ckpt}, out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_two_ranks_run_equals_one_process(tmp_path):
    tmp = str(tmp_path)
    _write_source(tmp, N=21, G=40, seed=1)            # 21 rows: ragged shards (11 + 10)
    one, two = str(tmp_path / "one.pt"), str(tmp_path / "two.pt")
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_run, args=(1, port, tmp, one), nprocs=1, join=True)
    mp.spawn(_run, args=(2, port + 1, tmp, two), nprocs=2, join=True)
    a, b = torch.load(one, weights_only=False), torch.load(two, weights_only=False)
    assert a["its"] == b["its"] == [0, 1]
    for (ia, la), (ib, lb) in zip(a["seen"], b["seen"]):
        assert ia == ib
        assert np.allclose(la, lb, rtol=2e-4, atol=1e-6), (la, lb)
    assert np.allclose(a["xs"], b["xs"], rtol=0, atol=2e-2)         # x255 encodings
    assert a["enc"].shape == b["enc"].shape == (21, 3) and a["enc"].dtype == np.float32
    assert np.allclose(a["enc"], b["enc"], atol=1e-4)
    # checkpoint written by rank 0 of the sharded run holds EVERY rank's optimiser state
    assert int(b["ckpt"]["meta/iteration"]) == 1
    keys = [k for k in a["ckpt"] if "/rms" in k or "/mom" in k or "/w" in k]
    assert keys and set(a["ckpt"]) == set(b["ckpt"])
    for k in keys:
        assert np.allclose(a["ckpt"][k], b["ckpt"][k], rtol=1e-3, atol=2e-5), k
    rms = [k for k in keys if "/rms" in k and b["ckpt"][k].size > 64]
    assert all(np.count_nonzero(b["ckpt"][k]) > 0.9 * b["ckpt"][k].size for k in rms)
