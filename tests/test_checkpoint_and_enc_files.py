"""SURVEY.md 8(f) "next" rows f3 / f4: the `.enc` encoding-file writer that the reference's
offline tools read (import_encodings.py:21-27, convert_encodings_to_mp4.py:27-30), and weight
/ optimiser checkpointing in Keras layout.  CPU-only: kernels emulated (tests/ops_emulator.py).
"""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import bigan_oracle as O

import ops_emulator
from cellcomm_b200 import engine as eng
from cellcomm_b200.intercepts import EncodingFiles
from cellcomm_b200.intercepts.encoding_files import load_encodings


@pytest.fixture(autouse=True)
def _emulated_ops(monkeypatch):
    monkeypatch.setattr(eng, "ops", ops_emulator)
    yield


def _inputs(Z, G, B, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.poisson(torch.rand(B, G, generator=g) * 3, generator=g)
    return x, torch.rand(B, Z, generator=g), torch.rand(B, Z, generator=g)


def _step(e, x, z, r, masks):
    x16 = ops_emulator.alloc2d(x.shape[0], x.shape[1])
    x16.copy_(x)
    e.set_latents(z, r, x.shape[0])
    return [float(v) for v in e.train_step(x16, masks)]


def test_checkpoint_resume_continues_bit_identically(tmp_path):
    Z, G, B = 3, 120, 10
    a = eng.BiGanEngine("cont", Z, G, max_batch=B, device="cpu", seed=5)
    masks = [O.make_masks("cont", Z, G, B, 40 + i) for i in range(3)]
    data = [_inputs(Z, G, B, 60 + i) for i in range(3)]
    _step(a, *data[0], masks[0])
    path = str(tmp_path / "ckpt.npz")
    a.save_checkpoint(path)
    ref = [_step(a, *data[i], masks[i]) for i in (1, 2)]

    b = eng.BiGanEngine("cont", Z, G, max_batch=B, device="cpu", seed=999)   # different init
    b.load_checkpoint(path)
    got = [_step(b, *data[i], masks[i]) for i in (1, 2)]
    assert got == ref                      # same weights, slots and BN statistics => same floats
    for n in ("G", "E", "D"):
        for wa, wb in zip(a.nets[n].get_weights(), b.nets[n].get_weights()):
            assert np.array_equal(wa, wb)
        for (ma, va), (mb, vb) in zip(a.nets[n].get_slots(), b.nets[n].get_slots()):
            assert np.array_equal(ma, mb) and np.array_equal(va, vb)
    assert b.rng_seed == a.rng_seed and int(b.rng_counter) == int(a.rng_counter)


def test_checkpoint_is_keras_layout(tmp_path):
    Z, G = 3, 100
    e = eng.BiGanEngine("cont", Z, G, max_batch=4, device="cpu", seed=1)
    path = str(tmp_path / "c.npz")
    e.save_checkpoint(path)
    with np.load(path) as f:
        # encoder: Dense(int(G*0.1)) on the cell -> kernel [in, out], bias  (bigan_cont.py:28-32)
        assert f["E/w0"].shape == (G, int(G * 0.1)) and f["E/w1"].shape == (int(G * 0.1),)
        # generator's first BatchNormalization: gamma, beta, moving_mean, moving_variance
        assert [f[f"G/w{i}"].shape for i in (6, 7, 8, 9)] == [(256,)] * 4
        assert np.all(f["G/w9"] == 1.0) and np.all(f["G/w8"] == 0.0)
        assert f["E/rms0"].shape == f["E/w0"].shape and f["E/mom0"].shape == f["E/w0"].shape
        assert str(f["meta/variant"]) == "cont" and int(f["meta/gene_size"]) == G


def test_checkpoint_rejects_other_architecture(tmp_path):
    e = eng.BiGanEngine("cont", 3, 100, max_batch=4, device="cpu", seed=1)
    path = str(tmp_path / "c.npz")
    e.save_checkpoint(path)
    other = eng.BiGanEngine("cont", 3, 120, max_batch=4, device="cpu", seed=1)
    with pytest.raises(ValueError, match="gene_size"):
        other.load_checkpoint(path)


class _Net:
    def __init__(self, enc):
        self.enc, self.calls = enc, 0

    def encoding_prediction(self, data):
        self.calls += 1
        return self.enc


class _Trainer:
    def __init__(self, enc):
        self.data = np.zeros((len(enc), 7))
        self.network = _Net(enc)


def test_enc_files_are_what_the_reference_tools_read(tmp_path):
    enc = np.random.default_rng(0).random((6, 3)).astype(np.float32)
    trainer = _Trainer(enc)
    log_dir = str(tmp_path / "logs" / "run1")
    intercept = EncodingFiles(log_dir).create_interceptor(trainer)
    intercept(0, (1.0, 2.0, 3.0))
    intercept(7, (1.0, 2.0, 3.0))
    assert sorted(os.listdir(os.path.join(log_dir, "encodings"))) == ["0.enc", "7.enc"]
    # import_encodings.py:24-27 verbatim
    with open(f"{log_dir}/encodings/7.enc", "rb") as f:
        encodings = pickle.load(f)
        coords = np.multiply(encodings, 255)
    assert encodings.dtype == np.float32 and encodings.shape == (6, 3)
    assert np.array_equal(coords, enc * np.float32(255))
    assert np.array_equal(load_encodings(log_dir, 0), enc)
    # import_encodings.py:68: iteration names recovered with rstrip('.enc')
    assert sorted(n.rstrip(".enc") for n in os.listdir(f"{log_dir}/encodings")) == ["0", "7"]
    assert trainer.network.calls == 2


def test_enc_files_length_assertion(tmp_path):
    trainer = _Trainer(np.zeros((4, 3), np.float32))
    trainer.data = np.zeros((5, 7))
    intercept = EncodingFiles(str(tmp_path)).create_interceptor(trainer)
    with pytest.raises(AssertionError, match="different length: 4 != 5"):
        intercept(0, None)


class _CkptNet:
    def __init__(self):
        self.saved, self.loaded = [], []

    def save_checkpoint(self, path, extra=None):
        self.saved.append(path)
        self.extra = extra
        open(path, "wb").write(b"x")

    def load_checkpoint(self, path):
        self.loaded.append(path)
        return dict(self.extra)


def test_checkpoint_interceptor_saves_every_n_and_resumes(tmp_path):
    from cellcomm_b200.intercepts import Checkpoints

    class T:
        network = _CkptNet()

    ck = Checkpoints(str(tmp_path / "logs" / "r"), every=2)
    assert ck.resume(T) == 0 and T.network.loaded == []
    intercept = ck.create_interceptor(T)
    for it in range(5):
        intercept(it, (0.0, 0.0, 0.0))
    assert len(T.network.saved) == 2 and T.network.saved[0].endswith("checkpoint.npz")   # it 1, 3
    # the newest checkpoint was written at iteration 3: the resumed run continues with 4
    assert ck.resume(T) == 4 and T.network.loaded == [ck.path]
