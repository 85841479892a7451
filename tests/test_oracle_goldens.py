"""Pin the oracle against every golden the reference's own tests hold for this path
(SURVEY.md §8c items 1-10).  The NN arithmetic itself has no golden in the reference
("parity unpinned", see oracle/bigan_oracle.py); it is pinned here against fp64 finite
differences and hand-derived cases instead."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import bigan_oracle as O
from oracle import loader_oracle as LO

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_loader_and_sampler_goldens():
    with open(os.path.join(GOLDEN, "loader_golden.json")) as f:
        gold = json.load(f)
    for name, file in (("fixture", "example_matrix.mtx"), ("synthetic_dups", "synthetic_dups.mtx")):
        dense, rows, cols = LO.load_matrix_numpy(os.path.join(GOLDEN, file))
        assert rows.tolist() == gold[name]["index"] and cols.tolist() == gold[name]["columns"]
        assert dense.tolist() == gold[name]["values"]
        df = LO.load_matrix_pandas(os.path.join(GOLDEN, file))
        assert df.values.tolist() == gold[name]["values"]
    # test/cell_type_training_test.py:14-21 and :36-41
    assert gold["fixture"]["values"] == [[0, 1, 6, 1, 11], [4, 1, 0, 1, 6], [1, 1, 1, 1, 6],
                                         [1, 0, 14, 0, 1], [0, 0, 0, 2, 0]]
    assert LO.sample_indices(5, 3, 0).tolist() == [2, 0, 1]
    rows = np.array(gold["fixture"]["index"])
    assert rows[LO.sample_indices(5, 3, 0)].tolist() == gold["fixture"]["sample_seed0_rows"]
    rows = np.array(gold["synthetic_dups"]["index"])
    assert rows[LO.sample_indices(len(rows), 8, 3)].tolist() == \
        gold["synthetic_dups"]["sample_seed3_rows"]


def test_round_half_even_golden():
    # test/bigans_basic_test.py:36-42
    pred = torch.tensor([[0.3, 12.59939265, 2.4894546, 0.01], [0.9, 4.7007282, 0, 2.07244989]])
    assert O.round_half_even(pred).tolist() == [[0, 13, 2, 0], [1, 5, 0, 2]]
    assert O.round_half_even(torch.tensor([0.5, 1.5, 2.5, -0.5])).tolist() == [0, 2, 2, -0.0]


def test_classify_prior_golden():
    # test/bigans_cc_test.py:57-63
    np.random.seed(21)
    hv = O.classify_random_encoding_vector(4, 5)
    assert hv.tolist() == [[0, 1, 0, 0], [0, 0, 0, 1], [1, 0, 0, 0], [1, 0, 0, 0], [1, 0, 0, 0]]


def test_onehot_argmax_golden():
    # test/bigans_cc_test.py:77-85
    p = torch.tensor([[0.1, 0.3], [0.7, 0.3], [0.001, 0.99]])
    assert O.to_categorical_argmax(p).tolist() == [[0, 1], [1, 0], [0, 1]]


def test_accuracy_bookkeeping_golden():
    # test/bigans_basic_test.py:57-67: D outputs [0.45,0.9,0.55] on fakes, [0.9,0.9,0.55] on reals
    fake = torch.tensor([0.45, 0.9, 0.55])
    real = torch.tensor([0.9, 0.9, 0.55])
    false_neg = int(torch.count_nonzero(torch.round(fake)))
    assert (int(torch.count_nonzero(torch.round(real))), 3 - false_neg) == (3, 1)


def test_encits_golden():
    # test/db_recorder_test.py:120-149
    encs = np.array([[0.5, 0.5, 0.0], [1.0, 0.2, 1.0], [0.5, 0.5, 0.5], [0.5, 0.5, 0.5],
                     [1.0, 0.2, 1.0]])
    coords = np.multiply(encs, 255)
    assert coords[:, 0].tolist() == [127.5, 255, 127.5, 127.5, 255]
    assert coords[:, 1].tolist() == [127.5, 51, 127.5, 127.5, 51]
    assert sorted(LO.find_duplicate_ids(coords)) == [[2, 5], [3, 4]]


def test_structural_shapes():
    # test/bigans_cc_test.py:16-55,89-121 and Appendix B parameter counts at G=33,694
    def count(spec):
        return sum(i * o + o if k == "dense" else 0 for k, *rest in spec
                   for i, o in [rest[:2] if k == "dense" else (0, 0)])
    Z, G = 3, 33694
    assert count(O.cont_encoder_spec(Z, G)) == 176_197_306 or abs(
        count(O.cont_encoder_spec(Z, G)) - 176.2e6) < 0.1e6
    assert abs(count(O.cont_generator_spec(Z, G)) - 250.7e6) < 0.1e6
    assert abs(count(O.discriminator_spec(Z, G)) - 494.7e6) < 0.1e6
    m = O.OracleBiGan("classify", 4, 6)
    x = torch.rand(5, 6)
    z = torch.tensor(O.classify_random_encoding_vector(4, 5))
    assert m.generator_predict(z, torch.rand(5, 4)).shape == (5, 6)
    enc = m.encoding_prediction(x)
    assert enc.shape == (5, 4) and torch.allclose(enc.sum(-1), torch.ones(5), atol=1e-6)
    assert m.discriminator_predict(z, x).shape == (5, 1)
    c = O.OracleBiGan("cont", 7, 11)
    e = c.encoding_prediction(torch.rand(3, 11))
    assert e.shape == (3, 7) and bool(((e > 0) & (e < 1)).all())
    # the 5-gene fixture yields zero-width Dense layers
    f = O.OracleBiGan("cont", 8, 5)
    assert [l["kernel"].shape for l in f.enc_layers if l["kind"] == "dense"][:2] == [
        torch.Size([5, 0]), torch.Size([5, 0])]


def test_gradients_against_finite_differences():
    """fp64 central differences on a tiny model pin the autograd graph of every sub-step."""
    torch.manual_seed(0)
    for variant, Z, G in (("cont", 2, 12), ("classify", 3, 12)):
        B = 5
        x = torch.poisson(torch.rand(B, G, dtype=torch.float64) * 3)
        z = torch.rand(B, Z, dtype=torch.float64)
        r = torch.rand(B, Z, dtype=torch.float64)
        masks = O.make_masks(variant, Z, G, B, 1)
        for k, net in ((1, "G"), (2, "G"), (3, "E"), (4, "E"), (6, "D"), (8, "D")):
            m = O.OracleBiGan(variant, Z, G, seed=3, dtype=torch.float64)
            if k in (6, 8):
                m.substep(5, x, z, r, masks)
                m.substep(7, x, z, r, masks)
            state = {n: m.get_weights(n) for n in "GED"}
            m.substep(k, x, z, r, masks)
            grads = m.last_grads[str(k)]
            params_idx = [i for i, l in enumerate(m.nets()[net])]
            # probe a few scalar entries of a few tensors
            rng = np.random.default_rng(k)
            flat_params = O.trainable_params(O.OracleBiGan(variant, Z, G, seed=3,
                                                           dtype=torch.float64).nets()[net])
            for t_i in rng.choice(len(flat_params), size=min(4, len(flat_params)), replace=False):
                if flat_params[t_i].numel() == 0:
                    continue
                e_i = int(rng.integers(flat_params[t_i].numel()))
                vals = []
                for sign in (+1, -1):
                    mm = O.OracleBiGan(variant, Z, G, seed=3, dtype=torch.float64)
                    for n in "GED":
                        mm.set_weights(n, state[n])
                    if k in (6, 8):
                        mm.gen_cells, mm.gen_enc = m.gen_cells, m.gen_enc
                    p = O.trainable_params(mm.nets()[net])[t_i]
                    p.view(-1)[e_i] += sign * 1e-6
                    vals.append(mm.substep(k, x, z, r, masks))
                fd = (vals[0] - vals[1]) / 2e-6
                an = float(grads[t_i].view(-1)[e_i])
                assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)) + 1e-7, (variant, k, t_i, fd, an)


def test_rmsprop_hand_case_and_bn_freeze_rules():
    # Keras RMSprop(momentum): first step from zero slots = lr*g/sqrt((1-rho) g^2 + eps)
    p = torch.tensor([1.0, -2.0])
    g = torch.tensor([0.5, 1e-5])
    slots = {}
    O.rmsprop_apply([p], [g], slots)
    exp = torch.tensor([1.0, -2.0]) - O.LR * g / torch.sqrt((1 - O.RHO) * g * g + O.EPSILON)
    assert torch.allclose(p, exp, atol=1e-7)
    ms, mom = slots[id(p)]
    O.rmsprop_apply([p], [g], slots)      # second step: momentum 0.1 carries over
    ms2 = O.RHO * (1 - O.RHO) * g * g + (1 - O.RHO) * g * g
    assert torch.allclose(ms, ms2)
    # frozen BN runs in inference mode and keeps its moving stats; the trained net's move
    m = O.OracleBiGan("cont", 3, 40, seed=1)
    x = torch.poisson(torch.rand(6, 40) * 2)
    z, r = torch.rand(6, 3), torch.rand(6, 3)
    d_mean = [l["moving_mean"].clone() for l in m.dis_layers if l["kind"] == "bn"]
    g_mean = [l["moving_mean"].clone() for l in m.gen_layers if l["kind"] == "bn"]
    m.substep(1, x, z, r)                  # trains G through a frozen D
    assert all(torch.equal(a, l["moving_mean"]) for a, l in
               zip(d_mean, [l for l in m.dis_layers if l["kind"] == "bn"]))
    assert any(not torch.equal(a, l["moving_mean"]) for a, l in
               zip(g_mean, [l for l in m.gen_layers if l["kind"] == "bn"]))
    # loss bookkeeping: g = 1+2, e = 3+4, d = mean(6, 8)   (src/bigan_classify.py:140-152)
    out = m.trainings_step(x, z, r)
    L = m.last_losses
    assert out[0] == pytest.approx(L["1"] + L["2"]) and out[1] == pytest.approx(L["3"] + L["4"])
    assert out[2] == pytest.approx((L["6"] + L["8"]) / 2)
