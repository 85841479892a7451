"""Worker of tests/test_data_parallel_gpu.py (one process per GPU under torchrun, or a single
process for the 1-GPU baseline of the API comparison).  Not collected by pytest.

    torchrun --nproc-per-node N tests/dp_gpu_worker.py engine <out.json>
    [torchrun ...] python tests/dp_gpu_worker.py api <source-dir> <out.pt>
"""
import json
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def _init():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world


def _same_on_all_ranks(t, what):
    ref = t.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(t, ref), f"rank {dist.get_rank()}: {what} differs from rank 0"


def engine_mode(out_path):
    """N ranks x B/N rows == the oracle's step on the global batch (explicit masks), then two
    captured-graph steps; all ranks must end with bit-identical state."""
    from cellcomm_b200 import engine as eng, ops
    from oracle import bigan_oracle as O
    import test_parity_gpu as P
    rank, world = _init()
    variant, Z, G, B = "cont", 3, 3000, 256
    orc = O.OracleBiGan(variant, Z, G, seed=0, dtype=torch.float32)
    e = eng.BiGanEngine(variant, Z, G, max_batch=B // world, device="cuda", seed=0,
                        dist=eng.TorchDist())
    for n in ("G", "E", "D"):
        e.nets[n].set_weights([w.numpy() for w in orc.get_weights(n)])
    x, z, r = P._inputs(variant, Z, G, B, 11)
    masks = O.make_masks(variant, Z, G, B, 3)
    lo, hi = rank * B // world, (rank + 1) * B // world
    lm = {s: {n: [m[lo:hi].cuda() for m in ms] for n, ms in d.items()} for s, d in masks.items()}
    x16 = ops.alloc2d(hi - lo, G)
    x16.copy_(x[lo:hi])
    e.set_latents(z[lo:hi], r[lo:hi], hi - lo)
    before = {n: [w.clone() for w in orc.get_weights(n)] for n in ("G", "E", "D")}
    got = [float(v) for v in e.train_step(x16, lm)]
    e.join()
    torch.cuda.synchronize()
    weights = {n: e.nets[n].get_weights() for n in ("G", "E", "D")}      # collective
    res = {"world": world, "losses": got}
    if rank == 0:
        ref = orc.trainings_step(x, z, r, masks)
        res["oracle_losses"] = list(ref)
        res["update_cosine"] = {}
        for n in ("G", "E", "D"):
            du_r = np.concatenate([(a - b).numpy().ravel() for a, b in zip(orc.get_weights(n), before[n])])
            du_g = np.concatenate([(a - b.numpy()).ravel() for a, b in zip(weights[n], before[n])])
            res["update_cosine"][n] = P._cos(du_g, du_r)
    # graph path: gather + priors + eight sub-steps + peer-memory exchange in ONE graph per rank
    res["graph"] = "eager only"
    if e.peer_graphable():
        from cellcomm_b200.cell_type_training import CellMatrix
        g = torch.Generator().manual_seed(5)
        dense = ((torch.rand(4 * B, G, generator=g) < 0.06).float() *
                 (torch.poisson(torch.full((4 * B, G), 1.2), generator=g) + 1)).numpy().astype(np.float64)
        csr = CellMatrix.from_dense(dense).device_csr("cuda")
        gs = e.capture_step(csr, G, B // world, latents="device")
        for step in range(2):
            idx = np.random.RandomState(step).permutation(4 * B)[:B][lo:hi]
            losses = [float(v) for v in gs.replay(torch.from_numpy(idx))]
            assert all(np.isfinite(losses)), losses
        res["graph"] = f"2 CUDA-graph replays, {gs.launches_per_replay} kernels each"
        res["graph_losses"] = losses
    e.join()
    torch.cuda.synchronize()
    if world > 1:
        for name, n in e.nets.items():
            _same_on_all_ranks(n.p16.clone(), f"{name} bf16 compute copy")
            if n.p16lo is not None:
                _same_on_all_ranks(n.p16lo.clone(), f"{name} low-order bf16 terms")
            n.gather_master()
            n.gather_slots()
            _same_on_all_ranks(n.p32.clone(), f"{name} fp32 master weights")
            _same_on_all_ranks(n.ms.clone(), f"{name} rms slots")
            bn = torch.cat([L["moving_mean"] for L in n.layers if L["kind"] == "bn"] +
                           [L["moving_var"] for L in n.layers if L["kind"] == "bn"])
            _same_on_all_ranks(bn, f"{name} BN moving statistics")
    if rank == 0:
        pr = e.G.peer
        res["exchange"] = "NCCL reduce-scatter / all-gather" if pr is None else (
            "peer-memory push + fused optimiser, all-gather by " +
            ("NVLS multicast stores" if pr["p16_mc"] else "P2P stores"))
        with open(out_path, "w") as f:
            json.dump(res, f)
        print("dp_check ok:", json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def api_mode(src_dir, out_path):
    """`CellTraining.run` with rank-0 interceptors (recorder on the in-memory Mongo, .enc files,
    checkpoints); dropout rates set to 0 so that 1-rank and N-rank runs are comparable (the
    Philox dropout streams are keyed by rank)."""
    from cellcomm_b200 import intercepts
    from cellcomm_b200.cell_type_training import CellTraining, load_matrix
    from cellcomm_b200.intercepts import db_recorder as dbr
    from cellcomm_b200.intercepts.fake_mongo import MongoClient as FakeMongo
    rank, world = _init()
    src = {k: os.path.join(src_dir, f"s_{k}.{'mtx' if k == 'matrix' else 'tsv'}")
           for k in ("matrix", "barcodes", "genes")}
    data = load_matrix(src["matrix"])
    np.random.seed(100 + rank)                       # run() replaces this with rank 0's state
    trainer = CellTraining(data, batch_size=64, encoding_size=3, batches_per_iteration=2)
    net = trainer.network
    assert net._world() == world
    for name, n in net._engine.nets.items():
        gen = torch.Generator().manual_seed(7 + len(name) + ord(name))
        ws = n.get_weights()
        n.set_weights([((torch.rand(w.shape, generator=gen) - 0.5) * 0.1).numpy()
                       if (w.ndim == 2) else w for w in ws])
        for node in n.g.nodes:
            if node["kind"] == "dropout":
                node["rate"] = 0.0
    seen, icpt = [], None
    if rank == 0:
        np.random.seed(5)
        net._prior_rng = np.random.default_rng(9)
        log_dir = os.path.join(src_dir, f"logs_w{world}")
        FakeMongo(dbr.MONGO_URL).drop_database(dbr.MONGO_DB)
        rec = dbr.DbRecorder("run", src, client_factory=FakeMongo)
        rec.setup()
        icpt = intercepts.combined_interceptors((
            lambda it, l: seen.append((it, [float(v) for v in l])),
            rec.create_interceptor(trainer),
            intercepts.EncodingFiles(log_dir).create_interceptor(trainer),
            intercepts.Checkpoints(log_dir).create_interceptor(trainer)))
    trainer.run(2, icpt)
    if rank == 0:
        docs = FakeMongo(dbr.MONGO_URL)[dbr.MONGO_DB][dbr.ITERATIONS_COLLECTION].find({"eid": "run"})
        with open(os.path.join(log_dir, "encodings", "1.enc"), "rb") as f:
            enc_file = pickle.load(f)
        ckpt = dict(np.load(os.path.join(log_dir, "checkpoint.npz")))
        torch.save({"seen": seen, "xs": [d["xs"] for d in docs], "its": [d["it"] for d in docs],
                    "enc": enc_file, "rms_nonzero": {k: float(np.count_nonzero(v)) / max(v.size, 1)
                                                     for k, v in ckpt.items() if "/rms" in k and v.size > 4096},
                    "graphs": len(net._engine._graphs)}, out_path)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    if sys.argv[1] == "engine":
        engine_mode(sys.argv[2])
    else:
        api_mode(sys.argv[2], sys.argv[3])
