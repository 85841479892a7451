"""The PUBLIC Python API on the GPU against the oracle (no emulator, no mocks on the compute
path): `load_matrix` -> `CellTraining.run` -> `trainings_step(CellBatch)` (the CUDA-graph
path) -> `encoding_prediction(CellMatrix)` (cc_encode_stream) -> `DbRecorder.intercept` /
`EncodingFiles` / `Checkpoints`, plus `evaluate_discriminator_accuracy` and checkpoint ->
resume.  Reference call sites: src/__main__.py:44-66, src/cell_type_training.py:37-50,
src/bigan_classify.py:126-155, src/bigan_basic.py:29-64, src/intercepts/db_recorder.py:82-108.

The oracle is fed exactly what the API path consumed: the batches (`np.random` global state),
the priors (the network's private generator, seeded here) and the dropout masks (regenerated
from the engine's Philox streams with cc_dropout_mask).

Tolerances: losses rel 1e-2 for a step from identical weights, 3e-2 for the sum over an
iteration of two free-running steps; encodings per-cell cosine >= 0.999 and |diff| <= 2e-2;
document bookkeeping (ids, names, iteration numbers, duplicate groups) exact; checkpoint ->
resume continues the uninterrupted run up to reduction-order noise (the column reductions and
loss sums use float atomics, so two runs of the same step differ in the last bit).
"""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import bigan_oracle as O

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
SOURCES = {"matrix": os.path.join(GOLDEN, "example_matrix.mtx"),
           "barcodes": os.path.join(GOLDEN, "example_barcodes.tsv"),
           "genes": os.path.join(GOLDEN, "example_genes.tsv")}
NETS_OF = {1: ("G", "D"), 2: ("E", "G"), 3: ("E", "D"), 4: ("G", "E"), 6: ("D",), 8: ("D",)}


def _write_source(tmp, N, G, seed):
    """A 10x-shaped source (matrix.mtx in barcode-major order + barcodes / genes TSVs)."""
    rng = np.random.default_rng(seed)
    dense = (rng.random((N, G)) < 0.06) * rng.geometric(0.45, (N, G))
    dense = dense * np.where(rng.random((N, G)) < 0.01, 50, 1)
    dense[np.arange(N), rng.integers(0, G, N)] += 1
    dense[rng.integers(0, N, G), np.arange(G)] += 1
    src = {k: os.path.join(tmp, f"s_{k}.{'mtx' if k == 'matrix' else 'tsv'}")
           for k in ("matrix", "barcodes", "genes")}
    with open(src["matrix"], "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate integer general\n%\n{G} {N} {(dense > 0).sum()}\n")
        for b in range(N):
            for g in np.flatnonzero(dense[b]):
                f.write(f"{g + 1} {b + 1} {int(dense[b, g])}\n")
    with open(src["barcodes"], "w") as f:
        f.writelines(f"BC{b:05d}-1\n" for b in range(N))
    with open(src["genes"], "w") as f:
        f.writelines(f"ENS{g:06d}\tsym{g}\n" for g in range(G))
    return src, dense.astype(np.float32)


def _engine_masks(eng, step, B):
    """The keep-masks the engine's Philox streams produce in training step `step` (device RNG
    counter = step), in the oracle's {substep: {net: [masks in call order]}} layout."""
    from cellcomm_b200 import ops
    counter = torch.tensor([step], dtype=torch.int64, device="cuda")
    out = {}
    for s, nets in NETS_OF.items():
        out[s] = {}
        for name in nets:
            args = eng._drop_args(s, name, None)
            seed, _, base = args["rng"]
            ms = []
            for node in eng.nets[name].g.nodes:
                if node["kind"] != "dropout":
                    continue
                w = eng.nets[name].g.widths[node["out"]]
                m = torch.ones((B, max(w, 1)), dtype=torch.uint8, device="cuda")[:, :w]
                if w > 0:
                    m = torch.empty((B, w), dtype=torch.uint8, device="cuda")
                    ops.dropout_mask(m, node["rate"], seed=seed, counter=counter,
                                     stream_id=base + node["drop"])
                ms.append(m.cpu())
            out[s][name] = ms
    return out


def _sync_oracle(orc, eng):
    """oracle <- engine: weights, BN moving statistics, RMSprop slots"""
    for n in ("G", "E", "D"):
        orc.set_weights(n, eng.nets[n].get_weights())
        orc.set_slots(n, eng.nets[n].get_slots())


def _trainer(data, B, bpi, seed, variant_seed=0):
    from cellcomm_b200.cell_type_training import CellTraining
    np.random.seed(seed)
    trainer = CellTraining(data, batch_size=B, encoding_size=3, batches_per_iteration=bpi)
    net = trainer.network
    net._prior_rng = np.random.default_rng(seed + 1)
    net._engine.rng_seed = 777 + variant_seed
    net._engine._graphs.clear()
    return trainer, net


@pytest.mark.parametrize("graph", ["1", "0"])
def test_run_with_interceptors_matches_oracle(tmp_path, monkeypatch, graph):
    """graph = "1": trainings_step replays the captured CUDA graph (production default);
    "0": the eager launch sequence."""
    monkeypatch.setenv("CELLCOMM_B200_GRAPH", graph)
    from cellcomm_b200 import intercepts
    from cellcomm_b200.cell_type_training import load_matrix
    from cellcomm_b200.intercepts import db_recorder as dbr
    from cellcomm_b200.intercepts.fake_mongo import MongoClient as FakeMongo
    N, G, B, BPI, ITS = 333, 1500, 64, 2, 3
    src, dense = _write_source(str(tmp_path), N, G, seed=3)
    data = load_matrix(src["matrix"])
    assert data.shape == (N, G)
    trainer, net = _trainer(data, B, BPI, seed=11)
    eng = net._engine
    orc = O.OracleBiGan("cont", 3, G, seed=0)
    _sync_oracle(orc, eng)

    # what the API path will draw, replayed for the oracle: batches from numpy's global state,
    # priors from the private generator (encodings first, then noise: reference :130-131)
    st_np, st_prior = np.random.get_state(), net._prior_rng.bit_generator.state
    feed = []
    for step in range(ITS * BPI):
        pos = np.random.permutation(N)[:B]
        enc = net._prior_rng.random((B, 3), dtype=np.float32)
        noise = net._prior_rng.random((B, 3), dtype=np.float32)
        feed.append((pos, enc, noise))
    np.random.set_state(st_np)
    net._prior_rng.bit_generator.state = st_prior

    log_dir = str(tmp_path / "logs" / "run")
    client = FakeMongo(dbr.MONGO_URL)
    client.drop_database(dbr.MONGO_DB)
    rec = dbr.DbRecorder("api-gpu", src, client_factory=FakeMongo)
    rec.setup()
    enc_files = intercepts.EncodingFiles(log_dir)
    checked = []

    # every step is compared on its own: the oracle is re-synchronised with the engine (weights,
    # BN statistics, RMSprop slots) before each public trainings_step call, runs the same step
    # on the same batch / priors / masks, and the three returned losses are compared
    per_step, inner = [], net.trainings_step

    def checked_step(batch):
        step = len(per_step)
        pos, enc, noise = feed[step]
        assert np.array_equal(batch.positions, pos), "the trainer sampled a different batch"
        _sync_oracle(orc, eng)
        one = np.array(orc.trainings_step(torch.from_numpy(dense[pos]), enc, noise,
                                          _engine_masks(eng, step, B)))
        out = inner(batch)
        mine = np.array([float(v) for v in out])
        # very first step (zero slots): rel 1e-2.  Later steps start from identical state too,
        # but inside a step each side applies its own six RMSprop updates (sub-step 4's loss
        # sees the G / E that sub-steps 1-3 moved on each side): rel 5e-2
        rel, ab = (1e-2, 2e-3) if step == 0 else (5e-2, 5e-3)
        assert np.all(np.abs(mine - one) <= rel * np.abs(one) + ab), \
            f"step {step}: losses {mine} vs oracle {one}"
        per_step.append(out)
        return out

    net.trainings_step = checked_step

    def against_oracle(it, losses):
        got = [float(v) for v in losses]
        assert np.allclose(got, [sum(float(s[j]) for s in per_step[it * BPI:(it + 1) * BPI])
                                 for j in range(3)], rtol=1e-6), "run() must sum the step losses"
        # encode-all-cells from IDENTICAL weights: oracle <- engine, then both encode all cells
        _sync_oracle(orc, eng)
        checked.append((it, orc.encoding_prediction(torch.from_numpy(dense)).numpy()))

    trainer.run(ITS, intercepts.combined_interceptors((
        against_oracle, rec.create_interceptor(trainer), enc_files.create_interceptor(trainer),
        intercepts.Checkpoints(log_dir).create_interceptor(trainer))))
    assert [it for it, _ in checked] == [0, 1, 2]
    assert int(eng.rng_counter.item()) == ITS * BPI

    docs = client[dbr.MONGO_DB][dbr.ITERATIONS_COLLECTION].find({"eid": "api-gpu"}, {"_id": 0})
    assert [d["it"] for d in docs] == [0, 1, 2]
    for d, (it, ref) in zip(docs, checked):
        assert d["cids"] == list(range(1, N + 1)) and d["ns"][0] == "BC00000-1" and len(d["ns"]) == N
        got = np.stack([d["xs"], d["ys"], d["zs"]], 1) / 255.0
        cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
        assert cos.min() >= 0.999, f"iteration {it}: encoding cosine {cos.min()}"
        assert np.abs(got - ref).max() <= 2e-2
        assert d["ds"] == dbr.find_duplicate_ids(np.stack([d["xs"], d["ys"], d["zs"]], 1).astype(np.float32))
        with open(enc_files.path(it), "rb") as f:
            enc = pickle.load(f)
        assert enc.dtype == np.float32 and enc.shape == (N, 3)
        assert np.array_equal(np.multiply(enc, 255)[:, 0].tolist(), d["xs"])   # same pass, same floats
    run = client[dbr.MONGO_DB][dbr.ENCODINGS_COLLECTION].find_one({"_id": "api-gpu"})
    assert run["defit"] == 2 and run["showits"] == [0, 1, 2]
    assert os.path.exists(os.path.join(log_dir, "checkpoint.npz"))


def test_first_step_through_the_api_is_tight():
    """ONE trainings_step(CellBatch) from identical weights: the three returned losses at the
    single-step tolerance (rel 1e-2), for both the graph replay and the dense-array input."""
    from cellcomm_b200.cell_type_training import CellMatrix
    N, G, B = 200, 1200, 48
    rng = np.random.default_rng(0)
    dense = ((rng.random((N, G)) < 0.06) * (rng.poisson(1.2, (N, G)) + 1)).astype(np.float32)
    data = CellMatrix.from_dense(dense.astype(np.float64))
    for as_batch in (True, False):
        trainer, net = _trainer(data, B, 1, seed=5, variant_seed=int(as_batch))
        eng = net._engine
        orc = O.OracleBiGan("cont", 3, G, seed=0)
        _sync_oracle(orc, eng)
        st = net._prior_rng.bit_generator.state
        enc = net._prior_rng.random((B, 3), dtype=np.float32)
        noise = net._prior_rng.random((B, 3), dtype=np.float32)
        net._prior_rng.bit_generator.state = st
        batch = trainer.sample_cell_data(random_seed=3)
        masks = _engine_masks(eng, 0, B)
        ref = orc.trainings_step(torch.from_numpy(dense[batch.positions]), enc, noise, masks)
        got = net.trainings_step(batch if as_batch else dense[batch.positions])
        for a, r in zip(got, ref):
            assert abs(float(a) - r) <= 1e-2 * abs(r) + 2e-3, (as_batch, [float(v) for v in got], ref)


def test_losses_of_successive_steps_are_independent_values():
    """train_on_batch returns independent floats in the reference; the graph path must not
    hand out aliases of one device buffer (history.append(net.trainings_step(b)))."""
    from cellcomm_b200.cell_type_training import CellMatrix
    rng = np.random.default_rng(1)
    dense = ((rng.random((120, 700)) < 0.06) * (rng.poisson(1.2, (120, 700)) + 1))
    trainer, net = _trainer(CellMatrix.from_dense(dense.astype(np.float64)), 32, 1, seed=2)
    history = [net.trainings_step(trainer.sample_cell_data()) for _ in range(3)]
    torch.cuda.synchronize()
    firsts = [float(h[0]) for h in history]
    again = [float(h[0]) for h in history]
    assert firsts == again and len(set(firsts)) == 3, firsts


def test_evaluate_discriminator_accuracy_matches_oracle():
    from cellcomm_b200.cell_type_training import CellMatrix
    N, G, B = 150, 900, 96
    rng = np.random.default_rng(4)
    dense = ((rng.random((N, G)) < 0.06) * (rng.poisson(1.2, (N, G)) + 1)).astype(np.float32)
    trainer, net = _trainer(CellMatrix.from_dense(dense.astype(np.float64)), B, 1, seed=8)
    for _ in range(2):          # move D's outputs away from 0.5
        net.trainings_step(trainer.sample_cell_data())
    eng = net._engine
    orc = O.OracleBiGan("cont", 3, G, seed=0)
    _sync_oracle(orc, eng)
    batch = trainer.sample_cell_data(random_seed=9)
    st = net._prior_rng.bit_generator.state
    enc = net._prior_rng.random((B, 3), dtype=np.float32)
    noise = net._prior_rng.random((B, 3), dtype=np.float32)
    net._prior_rng.bit_generator.state = st
    x = torch.from_numpy(dense[batch.positions])
    # oracle probabilities: rows within 2e-2 of 0.5 may legitimately round either way
    p_fake = orc.discriminator_predict(enc, orc.generate_cells(enc, noise)).numpy().ravel()
    p_real = orc.discriminator_predict(orc.encoding_prediction(x), x).numpy().ravel()
    ref_tp, ref_tn = orc.evaluate_discriminator_accuracy(x, enc, noise)
    tp, tn = net.evaluate_discriminator_accuracy(batch)
    assert isinstance(tp, int) and isinstance(tn, int)
    assert abs(tp - ref_tp) <= int((np.abs(p_real - 0.5) < 2e-2).sum()), (tp, ref_tp)
    assert abs(tn - ref_tn) <= int((np.abs(p_fake - 0.5) < 2e-2).sum()), (tn, ref_tn)
    # and through the reference-shaped generic path (host round trip) the same counts
    from cellcomm_b200.bigan_basic import BasicBiGan
    net._prior_rng.bit_generator.state = st
    tp2, tn2 = BasicBiGan.evaluate_discriminator_accuracy(net, dense[batch.positions])
    assert abs(tp2 - tp) <= 2 and abs(tn2 - tn) <= 2


@pytest.mark.parametrize("G", [1000, 6000])
def test_checkpoint_resume_continues_the_run_on_the_gpu(tmp_path, G):
    """Fused optimiser + captured graphs: a run resumed from the Checkpoints interceptor's file
    continues with the same batches, priors and dropout streams; the results equal the
    uninterrupted run's up to the summation order of the atomics-based reductions.  At 6,000
    genes the wide kernels' fp32 state lives in the blocked device layout between steps
    (Net._state_layout; incl. two-segment kernels whose second segment starts off a block
    row): the checkpoint written mid-run, the resumed run and the final state all go through
    the layout conversions."""
    from cellcomm_b200 import intercepts
    from cellcomm_b200.cell_type_training import CellMatrix
    N, B = 180, 32
    rng = np.random.default_rng(6)
    dense = ((rng.random((N, G)) < 0.06) * (rng.poisson(1.2, (N, G)) + 1))
    data = CellMatrix.from_dense(dense.astype(np.float64))
    log_dir = str(tmp_path / "logs")
    ck = intercepts.Checkpoints(log_dir)

    a, net_a = _trainer(data, B, 2, seed=21)
    seen_a = []
    a.run(2, intercepts.combined_interceptors((
        lambda it, l: seen_a.append((it, [float(v) for v in l])), ck.create_interceptor(a))))
    # ... `a` carries on for one more iteration: the uninterrupted run
    os.rename(ck.path, ck.path + ".keep")
    a.run(1, lambda it, l: seen_a.append((it, [float(v) for v in l])), start_iteration=2)
    os.rename(ck.path + ".keep", ck.path)
    enc_a = net_a.encoding_prediction(data)

    from cellcomm_b200.cell_type_training import CellTraining
    np.random.seed(12345)                                  # everything below comes from the file
    b = CellTraining(data, batch_size=B, encoding_size=3, batches_per_iteration=2)
    start = ck.resume(b)
    assert start == 2
    seen_b = []
    b.run(1, lambda it, l: seen_b.append((it, [float(v) for v in l])), start_iteration=start)
    assert [it for it, _ in seen_b] == [2]
    assert np.allclose(seen_b[0][1], seen_a[2][1], rtol=1e-5, atol=1e-6), (seen_a, seen_b)
    if G == 6000:
        eng_a = net_a._engine
        assert eng_a.E.blockable and eng_a.D.blockable and eng_a.G.blockable
        assert eng_a.D.state_blocked and b.network._engine.D.state_blocked   # left so by the last step
    for n in ("G", "E", "D"):
        na, nb = net_a._engine.nets[n], b.network._engine.nets[n]
        na._rows()
        nb._rows()
        for ta, tb in ((na.p32, nb.p32), (na.ms, nb.ms), (na.mom, nb.mom)):
            assert float((ta - tb).abs().max()) <= 1e-5 * float(ta.abs().max()) + 1e-7
        assert float((na.p16.float() - nb.p16.float()).abs().max()) <= 2e-2 * float(na.p16.float().abs().max())
    assert np.allclose(enc_a, b.network.encoding_prediction(data), atol=1e-4)


def test_fixture_config_through_the_api():
    """BASELINE.json configs[0]: test/example_matrix.mtx (5 cells x 5 genes, Dense(0) layers,
    SURVEY.md D10) trained and recorded through the public API on the GPU."""
    from cellcomm_b200 import intercepts
    from cellcomm_b200.cell_type_training import CellTraining, load_matrix
    from cellcomm_b200.intercepts import db_recorder as dbr
    from cellcomm_b200.intercepts.fake_mongo import MongoClient as FakeMongo
    data = load_matrix(SOURCES["matrix"])
    assert data.shape == (5, 5)
    np.random.seed(0)
    trainer = CellTraining(data, batch_size=3, encoding_size=3, batches_per_iteration=4)
    FakeMongo(dbr.MONGO_URL).drop_database(dbr.MONGO_DB)
    rec = dbr.DbRecorder("fixture", SOURCES, client_factory=FakeMongo)
    rec.setup()
    seen = []
    trainer.run(2, intercepts.combined_interceptors((
        lambda it, l: seen.append([float(v) for v in l]), rec.create_interceptor(trainer))))
    assert len(seen) == 2 and np.all(np.isfinite(seen))
    docs = FakeMongo(dbr.MONGO_URL)[dbr.MONGO_DB][dbr.ITERATIONS_COLLECTION].find({"eid": "fixture"})
    assert [d["it"] for d in docs] == [0, 1] and docs[0]["cids"] == [1, 2, 3, 4, 5]
    assert docs[0]["ns"][0] == "AAACCTGGTGTCCTCT-1"
    # encodings against the oracle from identical weights
    eng = trainer.network._engine
    orc = O.OracleBiGan("cont", 3, 5, seed=0)
    _sync_oracle(orc, eng)
    ref = orc.encoding_prediction(torch.from_numpy(data.to_numpy(np.float32))).numpy()
    got = np.stack([docs[1]["xs"], docs[1]["ys"], docs[1]["zs"]], 1) / 255.0
    assert np.abs(got - ref).max() <= 2e-2


@pytest.mark.parametrize("variant,Z,G", [("cont", 3, 2500), ("classify", 10, 1700)])
def test_encode_stream_is_the_tile_loop_in_one_call(variant, Z, G):
    """cc_encode_stream (gather + encoder forward per tile, one C call) against the composed
    path (cc_gather_rows + Net.forward per tile): same kernels in the same order, so the
    encodings are BIT-IDENTICAL -- for both reference encoders, a row range that starts inside
    the matrix, tiles that do not divide the range, and a tile larger than the range."""
    from cellcomm_b200 import engine as eng, ops
    from cellcomm_b200.cell_type_training import CellMatrix
    N = 700
    rng = np.random.default_rng(7)
    dense = ((rng.random((N, G)) < 0.06) * (rng.poisson(1.2, (N, G)) + 1)).astype(np.float64)
    csr = CellMatrix.from_dense(dense).device_csr("cuda")
    e = eng.BiGanEngine(variant, Z, G, max_batch=64, device="cuda", seed=3)
    assert e.encode_plan(128) is not None
    for lo, hi, tile in ((0, N, 256), (37, 611, 128), (5, 90, 4096)):
        ref = torch.empty(hi - lo, Z, device="cuda")
        buf = ops.alloc2d(tile, G)
        for s in range(lo, hi, tile):
            m = min(tile, hi - s)
            ops.gather_rows(*csr, G, row_start=s, n_rows=m, out16=buf[:m])
            e.encode(buf[:m], out32=ref[s - lo:s - lo + m])
        got = torch.empty(hi - lo, Z, device="cuda")
        e.encode_stream(*csr, lo, hi, got, tile_rows=tile)
        torch.cuda.synchronize()
        assert torch.equal(got, ref), (variant, lo, hi, tile, float((got - ref).abs().max()))
    orc = O.OracleBiGan(variant, Z, G, seed=0)
    _sync_oracle(orc, e)
    want = orc.encoding_prediction(torch.from_numpy(dense[5:90].astype(np.float32))).numpy()
    assert np.abs(got.cpu().numpy() - want).max() <= 2e-2
