"""The reference's own unit tests, replayed against the drop-in Python surface
(test/bigans_basic_test.py, test/bigans_cc_test.py, test/cell_type_training_test.py,
test/db_recorder_test.py).  CPU-only: models are mocks or structure-only, or run on the ops
emulator (host logic); the Mongo server is the in-memory stand-in."""
import os
from datetime import datetime
from unittest.mock import MagicMock

import numpy as np
import pytest
import torch

import ops_emulator
from cellcomm_b200 import engine as eng
from cellcomm_b200.bigan_basic import BasicBiGan, components_changed
from cellcomm_b200.bigan_classify import ClassifyCellBiGan
from cellcomm_b200.bigan_cont import ContinuousCellBiGan
from cellcomm_b200.cell_type_training import CellTraining, load_matrix
from cellcomm_b200.models import RMSprop, activations, losses
from cellcomm_b200 import intercepts
from cellcomm_b200.intercepts import db_recorder as dbr
from cellcomm_b200.intercepts.data_sink import DataSink
from cellcomm_b200.intercepts.fake_mongo import MongoClient as FakeMongo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TEST_MATRIX_FILE = os.path.join(GOLDEN, "example_matrix.mtx")


@pytest.fixture(autouse=True)
def _cpu_engine(monkeypatch):
    """structure / wiring tests: engine buffers on the CPU over the ops emulator"""
    monkeypatch.setattr(eng, "ops", ops_emulator)
    monkeypatch.setenv("CELLCOMM_B200_DEVICE", "cpu")
    yield


# ------------------------------------------------------------------ test/bigans_basic_test.py
@pytest.fixture
def mock_bigan():
    gen, enc, dis = MagicMock(), MagicMock(), MagicMock()
    b = BasicBiGan(encoding_size=15, gene_size=1000,
                   generator_factory=MagicMock(return_value=gen),
                   encoder_factory=MagicMock(return_value=enc),
                   discriminator_factory=MagicMock(return_value=dis))
    return b, gen, enc, dis


def test_create_components(mock_bigan):
    b, gen, enc, dis = mock_bigan
    assert b._encoder is enc and b._generator is gen and b._discriminator is dis
    assert b.all_components == (gen, enc, dis) and b.encoding_size == 15


def test_encoding_prediction_delegates(mock_bigan):
    b, *_ = mock_bigan
    genes = [[5, 3, 1, 4], [1, 5, 13, 7]]
    pred = [[0.1, 0.3], [0.7, 0.3], [0.001, 0.99]]
    b._encoder.predict = MagicMock(return_value=pred)
    assert b.encoding_prediction(genes) == pred
    b._encoder.predict.assert_called_with(genes)


def test_cell_data_prediction_rounds_half_even(mock_bigan):
    b, *_ = mock_bigan
    encoding_in, random_in = [[1, 0], [0, 1]], [[0.2, 0.], [0.99, 0.99]]
    pred = [[0.3, 12.59939265, 2.4894546, 0.01], [0.9, 4.7007282, 0, 2.07244989]]
    b._generator.predict = MagicMock(return_value=pred)
    cells = b.generate_cells(encoding_in, random_in)
    assert np.array_equal(cells, [[0, 13, 2, 0], [1, 5, 0, 2]])
    b._generator.predict.assert_called_with((encoding_in, random_in))
    assert len(cells) == 2


def test_random_uniform_vector(mock_bigan):
    b, *_ = mock_bigan
    hv = b.random_uniform_vector(7)
    assert hv.shape == (7, 15) and hv.dtype == np.float32
    assert np.all(hv >= 0) and np.all(hv < 1)


def test_get_accuracy(mock_bigan):
    b, *_ = mock_bigan
    data_batch = [[1, 2], [3, 4], [5, 6]]
    b.generate_cells = MagicMock(return_value=[6, 7, 8])
    b.random_encoding_vector = MagicMock(return_value=[4, 5, 6])
    b.encoding_prediction = MagicMock(return_value=[1, 2, 3])
    b._discriminator.predict = discr = MagicMock(side_effect=[[0.45, 0.9, 0.55], [0.9, 0.9, 0.55]])
    assert b.evaluate_discriminator_accuracy(data_batch) == (3, 1)
    b.random_encoding_vector.assert_called_once_with(3)
    b.generate_cells.assert_called_once_with([4, 5, 6])
    discr.assert_any_call(([4, 5, 6], [6, 7, 8]), use_multiprocessing=True)
    discr.assert_any_call(([1, 2, 3], data_batch), use_multiprocessing=True)


def test_components_changed():
    a = [np.ones(3), np.zeros(2)]
    assert not components_changed(a, [np.ones(3), np.zeros(2)])
    assert components_changed(a, [np.ones(3), np.array([0.0, 1.0])])


# ------------------------------------------------------------------ test/bigans_cc_test.py
def test_classify_generator_train_model():
    bigan = ClassifyCellBiGan(encoding_size=15, gene_size=1000)
    g = bigan._generator
    assert g.layers[0].name == "gen_encoding_in" and g.layers[1].name == "gen_random_in"
    assert g.layers[0].input_shape == [(None, 15)] and g.layers[1].input_shape == [(None, 15)]
    assert g.output_shape == (None, 1000)
    assert g.layers[-1].activation is activations.relu
    t = bigan._train_gen_w_discr
    assert t._is_compiled and t.loss is losses.binary_crossentropy
    assert type(t.optimizer) is RMSprop
    assert t.layers[-1].layers[-1].activation is activations.sigmoid
    cfg = t.optimizer.get_config()
    assert (cfg["learning_rate"], cfg["rho"], cfg["momentum"]) == (0.0075, 0.85, 0.1)


def test_classify_encoder_and_discriminator_models():
    bigan = ClassifyCellBiGan(encoding_size=12, gene_size=100)
    e = bigan._encoder
    assert e.input_shape == (None, 100) and e.output_shape == (None, 12)
    assert e.layers[-1].activation is activations.softmax
    t = bigan._train_enc_w_discr
    assert t._is_compiled and t.loss is losses.binary_crossentropy
    assert t.layers[-1].layers[-1].activation is activations.sigmoid
    d = ClassifyCellBiGan(encoding_size=4, gene_size=5)._discriminator
    assert d._is_compiled and d.loss is losses.binary_crossentropy
    assert d.get_input_shape_at(0) == [(None, 4), (None, 5)]
    assert d.get_output_shape_at(0) == (None, 1)
    assert d.layers[-1].activation is activations.sigmoid and type(d.optimizer) is RMSprop


def test_classify_random_encoding_vector_golden():
    np.random.seed(21)
    hv = ClassifyCellBiGan(encoding_size=4, gene_size=1).random_encoding_vector(5)
    assert hv.shape == (5, 4)
    assert hv.tolist() == [[0, 1, 0, 0], [0, 0, 0, 1], [1, 0, 0, 0], [1, 0, 0, 0], [1, 0, 0, 0]]


def test_training_models_shapes():
    for cls, z, g in ((ClassifyCellBiGan, 4, 6), (ContinuousCellBiGan, 7, 11)):
        b = cls(encoding_size=z, gene_size=g)
        assert b._train_gen_w_discr.input_shape == [(None, z), (None, z)]
        assert b._train_gen_w_discr.output_shape == (None, 1)
        assert b._train_enc_w_discr.input_shape == (None, g)
        assert b._train_enc_w_discr.output_shape == (None, 1)
        assert b._train_gen_w_enc.loss is losses.mse and b._train_enc_w_gen.loss is losses.mse


def test_trainings_encoding_prediction_variants():
    b = ClassifyCellBiGan(encoding_size=2, gene_size=4)
    pred = [[0.1, 0.3], [0.7, 0.3], [0.001, 0.99]]
    b.encoding_prediction = MagicMock(return_value=pred)
    genes = [[5, 3, 1, 4], [1, 5, 13, 7]]
    assert b.trainings_encoding_prediction(genes).tolist() == [[0, 1], [1, 0], [0, 1]]
    b.encoding_prediction.assert_called_with(genes)
    c = ContinuousCellBiGan(encoding_size=2, gene_size=4)
    c._encoder.predict = MagicMock(return_value=pred)
    assert c.trainings_encoding_prediction(genes) == pred
    c._encoder.predict.assert_called_with(genes)


def test_continuous_structure_and_prior():
    b = ContinuousCellBiGan(encoding_size=13, gene_size=100)
    g = b._generator
    assert g.layers[0].name == "gen_encoding_in" and g.layers[1].name == "gen_random_in"
    assert g.output_shape == (None, 100) and g.layers[-1].activation is activations.relu
    assert b._encoder.layers[-1].activation is activations.sigmoid
    assert b._train_gen_w_discr.layers[-1].layers[-1].activation is activations.sigmoid
    c = ContinuousCellBiGan(encoding_size=20, gene_size=1)
    for _ in range(30):
        v = c.random_encoding_vector(3)
        assert v.shape == (3, 20) and np.all(v >= 0) and np.all(v < 1)


def test_predict_and_step_run_through_the_engine():
    """end-to-end host path on the emulator: sample -> gather -> step -> predict"""
    data = load_matrix(TEST_MATRIX_FILE)
    tr = CellTraining(data, 3, 3, batches_per_iteration=2)
    seen = []
    tr.run(2, lambda it, losses: seen.append((it, [float(v) for v in losses])))
    assert [s[0] for s in seen] == [0, 1] and all(np.isfinite(v) for s in seen for v in s[1])
    enc = tr.network.encoding_prediction(tr.data)
    assert enc.shape == (5, 3) and enc.dtype == np.float32
    cells = tr.network.generate_cells(tr.network.random_encoding_vector(4))
    assert cells.shape == (4, 5) and np.array_equal(cells, np.round(cells))
    tp, tn = tr.network.evaluate_discriminator_accuracy(tr.sample_cell_data())
    assert 0 <= tp <= 3 and 0 <= tn <= 3
    # a dense host batch (what the reference passes) takes the upload path
    g, e, d = tr.network.trainings_step(data.values[[2, 0, 1]])
    assert np.isfinite(float(g + e + d))
    w = tr.network._encoder.layers[1].get_weights()
    assert len(w) == 2 and w[0].shape[0] == 5
    tr.network.summary()
    tr.network.print_params_changes("after")


# ------------------------------------------------------------------ test/cell_type_training_test.py
def test_bigan_setup_and_run_bookkeeping():
    data = load_matrix(TEST_MATRIX_FILE)
    trainer = CellTraining(data, 3, 8, batches_per_iteration=4)
    b = trainer.network
    assert b.encoding_size == 8
    assert b._generator.get_input_shape_at(0) == [(None, 8), (None, 8)]
    assert b._generator.output_shape == (None, 5)
    trainer.sample_cell_data = sample = MagicMock(return_value=['some', 'data'])
    trainer.network.trainings_step = step = MagicMock(return_value=(0.7, 0.8, 0.8))
    trainer.run(6, None)
    assert sample.call_count == 24 and step.call_count == 24
    step.assert_called_with(['some', 'data'])
    icpt = MagicMock()
    trainer.network.trainings_step = MagicMock(return_value=(1, 2, 3))
    trainer.run(3, icpt)
    icpt.assert_called_with(2, (4, 8, 12))


# ------------------------------------------------------------------ intercepts
def test_interceptor_combinators(capsys):
    calls = []
    ic = intercepts.combined_interceptors([lambda it, l: calls.append(("a", it)),
                                           lambda it, l: calls.append(("b", it))])
    ic(3, (1, 2, 3))
    assert calls == [("a", 3), ("b", 3)]
    calls.clear()
    sk = intercepts.skip_iterations(3, lambda it, l: calls.append(it))
    for it in range(7):
        sk(it, None)
    assert calls == [2, 5]
    calls.clear()
    off = intercepts.offset_iterations(2, lambda it, l: calls.append(it))
    for it in range(4):
        off(it, None)
    assert calls == [2, 3]
    intercepts.print_losses("run-x")(7, (1.0, 2.5, eng.LossScalar(torch.tensor(0.25))))
    out = capsys.readouterr().out
    assert "run-x it:      7  TOT:  3.750  G-L:  1.000  E-L:  2.500  D-L:  0.250" in out


def test_data_sink_and_sink_intercepts(tmp_path):
    s = DataSink(log_dir=str(tmp_path), batch_size=2)
    s.add_graph_header("g", ["a", "b"])
    with pytest.raises(AssertionError, match="duplicate graph name: g"):
        s.add_graph_header("g", ["a"])
    with pytest.raises(AssertionError, match="unknown graph: nope"):
        s.add_data("nope", [1])
    with pytest.raises(AssertionError, match=r"expected 2 values, received: \[1\]"):
        s.add_data("g", [1])
    s.add_data("g", [1, 2])
    assert (tmp_path / "g.csv").read_text() == "a,b\n"          # batched
    s.add_data("g", [3, 4])
    assert (tmp_path / "g.csv").read_text() == "a,b\n1,2\n3,4\n"
    si = intercepts.SinkIntercepts(str(tmp_path))
    rec = si.save_losses()
    rec(0, (1.0, 2.0, 3.0))
    assert (tmp_path / "losses.csv").read_text() == \
        "iteration,total-loss,g-loss,e-loss,d-loss\n0,6.0,1.0,2.0,3.0\n"
    trainer = MagicMock()
    trainer.sample_cell_data.return_value = [1, 2, 3, 4]
    trainer.network.evaluate_discriminator_accuracy.return_value = (3, 1)
    si.save_accuracy(trainer)(5, None)
    assert (tmp_path / "accuracy.csv").read_text() == "iteration,pos-pct,neg-pct\n5,0.75,0.25\n"


# ------------------------------------------------------------------ test/db_recorder_test.py
TEST_DB = "cellcomm-test"
RUN = "test-run"
SOURCES = {"matrix": TEST_MATRIX_FILE,
           "barcodes": os.path.join(GOLDEN, "example_barcodes.tsv"),
           "genes": os.path.join(GOLDEN, "example_genes.tsv")}


@pytest.fixture
def recorder():
    FakeMongo(dbr.MONGO_URL).drop_database(TEST_DB)
    return dbr.DbRecorder(RUN, SOURCES, TEST_DB, client_factory=FakeMongo)


def _coll(name):
    return FakeMongo(dbr.MONGO_URL)[TEST_DB][name]


def test_check_files():
    bad = dict(SOURCES, barcodes=os.path.join(GOLDEN, "does.not.exist"))
    with pytest.raises(AssertionError) as cm:
        dbr.DbRecorder("fail-run-id", bad, client_factory=FakeMongo)
    assert str(cm.value) == f"File not found: {bad['barcodes']}"


def test_stores_encoding_run_and_rejects_duplicates(recorder):
    recorder.store_encoding_run()
    run = _coll(dbr.ENCODINGS_COLLECTION).find_one({"_id": RUN})
    assert run["_id"] == RUN and run["defit"] == 0 and run["showits"] == []
    assert (datetime.now() - run["date"]).total_seconds() < 1
    assert run["srcs"] == {"matrix": "example_matrix.mtx", "barcodes": "example_barcodes.tsv",
                           "genes": "example_genes.tsv"}
    with pytest.raises(AssertionError) as cm:
        recorder.store_encoding_run()
    assert str(cm.value) == f"Encoding run id already exists: {RUN}"


def test_stores_cells_and_genes(recorder):
    recorder.setup()
    cells = _coll(dbr.CELLS_COLLECTION).find({"sid": "example_barcodes.tsv"})
    assert len(cells) == 5
    assert (cells[0]["cid"], cells[0]["n"]) == (1, "AAACCTGGTGTCCTCT-1")
    assert cells[0]["g"] == [{"e": "ENSMUSG00000025902", "m": "Sox17", "v": 11},
                             {"e": "ENSMUSG00000102343", "m": "Gm37381", "v": 6},
                             {"e": "ENSMUSG00000089699", "m": "Gm1992", "v": 1},
                             {"e": "ENSMUSG00000109048", "m": "Rp1", "v": 1}]
    assert (cells[4]["cid"], cells[4]["n"]) == (5, "AAAGATGGTGATAAAC-1")
    assert cells[4]["g"] == [{"e": "ENSMUSG00000109048", "m": "Rp1", "v": 2}]
    genes = _coll(dbr.GENES_COLLECTION).find({"sid": "example_barcodes.tsv"}, {"_id": 0})
    assert len(genes) == 5
    assert genes[0] == {"sid": "example_barcodes.tsv", "e": "ENSMUSG00000089699", "m": "Gm1992",
                        "cids": [1, 2, 3]}
    assert genes[4] == {"sid": "example_barcodes.tsv", "e": "ENSMUSG00000051951", "m": "Xkr4",
                        "cids": [2, 3, 4]}
    assert recorder.cell_ids == [1, 2, 3, 4, 5]


def test_use_existing_cells(recorder):
    _coll(dbr.CELLS_COLLECTION).insert_many([{"sid": "example_barcodes.tsv", "n": "bla-bla"},
                                             {"sid": "example_barcodes.tsv", "n": "blu-blu"}])
    recorder.setup()
    assert recorder.barcodes == ["bla-bla", "blu-blu"] and recorder.cell_ids == [1, 2]


def test_ordering_assertions(recorder):
    with pytest.raises(AssertionError) as cm:
        recorder.load_barcodes()
    assert recorder.barcodes is None
    assert str(cm.value) == "Cannot load barcodes without encoding!"
    with pytest.raises(AssertionError) as cm:
        recorder.create_interceptor(None)
    assert str(cm.value) == "Cannot store iterations without barcodes!"


def _trainer(encs):
    t = MagicMock()
    t.network.encoding_prediction = MagicMock(return_value=encs)
    return t


ENCS = np.array([[0.5, 0.5, 0.0], [1.0, 0.2, 1.0], [0.5, 0.5, 0.5], [0.5, 0.5, 0.5],
                 [1.0, 0.2, 1.0]])


def test_intercept_stores_iteration(recorder):
    recorder.setup()
    icpt = recorder.create_interceptor(_trainer(ENCS))
    icpt(2009, {"what": "ever"})
    icpt(3009, {"what": "ever"})
    its = _coll(dbr.ITERATIONS_COLLECTION).find({"eid": RUN}, {"_id": 0})
    assert len(its) == 2
    assert its[0] == {
        "eid": RUN, "it": 2009, "cids": [1, 2, 3, 4, 5],
        "ns": ["AAACCTGGTGTCCTCT-1", "AAACGGGCAGGTCTCG-1", "AAACGGGTCCGCTGTT-1",
               "AAACGGGTCTGATTCT-1", "AAAGATGGTGATAAAC-1"],
        "xs": [127.5, 255, 127.5, 127.5, 255], "ys": [127.5, 51, 127.5, 127.5, 51],
        "zs": [0.0, 255, 127.5, 127.5, 255], "ds": [[3, 4], [2, 5]]}
    enc = _coll(dbr.ENCODINGS_COLLECTION).find_one({"_id": RUN})
    assert enc["defit"] == 3009 and enc["showits"] == [2009, 3009]


def test_intercept_shape_and_duplicate_assertions(recorder):
    recorder.setup()
    with pytest.raises(AssertionError) as cm:
        recorder.create_interceptor(_trainer(np.array([[], [], [], []])))(2, None)
    assert str(cm.value) == "encodings + barcodes have different length: 4 != 5"
    with pytest.raises(AssertionError) as cm:
        recorder.create_interceptor(_trainer(np.array([[1, 2]] * 5)))(3, None)
    assert str(cm.value) == "encodings vector length = 2, not in x, y, z format"
    icpt = recorder.create_interceptor(_trainer(ENCS))
    icpt(2009, None)
    with pytest.raises(AssertionError) as cm:
        icpt(2009, None)
    assert str(cm.value) == "duplicate iteration 2009"


def test_find_duplicate_ids_matches_reference_semantics():
    from oracle import loader_oracle as LO
    rng = np.random.default_rng(0)
    coords = rng.integers(0, 4, (300, 3)).astype(np.float32) * 63.75
    got = dbr.find_duplicate_ids(coords)
    ref = LO.find_duplicate_ids(coords)
    assert sorted(got) == sorted(ref)
    assert dbr.find_duplicate_ids(np.zeros((0, 3))) == []


def test_log_dir_guard(tmp_path, monkeypatch):
    from cellcomm_b200 import __main__ as entry
    d = tmp_path / "logs" / "x"
    entry.check_log_dir(str(d))
    with pytest.raises(AssertionError) as cm:
        entry.check_log_dir(str(d))
    assert str(cm.value) == f"duplicate run-id, log-dir: {d}"
    assert entry.RUN_ID == "test" and entry.DATA_SOURCES is entry.SOURCES[1]
    assert entry.SOURCES[1]["matrix"].endswith("GSE122930_TAC_4_weeks_repA+B_matrix.mtx")
