"""Parity with the oracle AT THE BASELINE SHAPE (BASELINE.json configs[1]: gene_size 33,694,
encoding_size 3): the shapes where the production code paths run — split-K forward / dgrad at
the reference's batch 128, 2-CTA clusters and the tile-width pick on every wide layer, the
fused wgrad + RMSprop epilogue on 227 M / 340 M-element kernels with both rasters (N-fast at
batch 128 and for Dx1 at batch 2048, M-fast for G6 at batch 2048), and the encode pass.

Protocol and tolerances are those of tests/test_parity_gpu.py (same constants): every update
is compared from IDENTICAL weights and RMSprop slots (src/bigan_classify.py:126-155, one
`train_on_batch` at a time), on the loss, every parameter gradient, the applied update
(w_after - w_before) and the BN moving statistics.  Comparisons of the 0.25-0.5 G-element
tensors run on the device in float64 (torch as plumbing).
"""
import numpy as np
import pytest
import torch

from oracle import bigan_oracle as O

import test_parity_gpu as P

pytestmark = pytest.mark.gpu

VARIANT, Z, G = "cont", 3, 33694
SLOT = {1: 0, 2: 1, 3: 2, 4: 3, 6: 4, 8: 5}


class _Pair:
    """One oracle + one engine at the BASELINE shape, shared by the tests of this module
    (building 921.6 M parameters twice takes longer than the comparisons)."""

    def __init__(self):
        from cellcomm_b200 import engine as eng
        self.orc = O.OracleBiGan(VARIANT, Z, G, seed=0, dtype=torch.float32)
        self.e = eng.BiGanEngine(VARIANT, Z, G, max_batch=128, device="cuda", seed=0)
        for n in ("G", "E", "D"):
            self.sync(n)

    def sync(self, name):
        """engine <- oracle: weights, BN moving statistics and RMSprop slots of one network"""
        net = self.e.nets[name]
        net.set_weights([w.numpy() for w in self.orc.get_weights(name)])
        net.set_slots([(a.numpy(), b.numpy()) for a, b in self.orc.get_slots(name)])


@pytest.fixture(scope="module")
def pair():
    p = _Pair()
    yield p
    del p
    torch.cuda.empty_cache()


def _log(rec):
    """one JSON line per comparison (pytest -s shows them; CELLCOMM_PARITY_LOG=path keeps them)"""
    import json
    import os
    line = json.dumps(rec)
    print(line)
    path = os.environ.get("CELLCOMM_PARITY_LOG")
    if path:
        with open(path, "a") as f:
            f.write(line + "\n")


def _dots(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float(a @ b), float(a @ a), float(b @ b)


def _cos3(d):
    ab, aa, bb = d
    if aa < 1e-60 or bb < 1e-60:
        return 1.0 if abs(aa - bb) < 1e-24 else 0.0
    return ab / (aa * bb) ** 0.5


def _trainable(net):
    out = []
    for L in net.layers:
        out += [(L["w32"], L["dw"]), (L["b32"], L["db"])] if L["kind"] == "dense" else \
            [(L["gamma"], L["dgamma"]), (L["beta"], L["dbeta"])]
    return out


def _check_update(p, k, x16, xo, zo, ro, masks, dmasks, B, *, resync=True):
    """One train_on_batch (sub-step k) on both sides from identical state; asserts loss,
    gradients, updates and BN moving statistics.  Leaves both sides re-synchronised."""
    from cellcomm_b200 import ops
    name = P.UPDATES[k]
    net = p.e.nets[name]
    before = [w.cuda() for w in O.trainable_params(p.orc.nets()[name])]
    ref_loss = p.orc.substep(k, xo, zo, ro, masks)
    ops.fill_f32(p.e.loss_buf, 0.0)
    p.e.substep(k, x16, dmasks)
    p.e.join()
    net._rows()         # the fused epilogue keeps the wide kernels' fp32 state blocked: back to
    torch.cuda.synchronize()   # rows before the layer views below are read
    got_loss = float(p.e.loss_buf[SLOT[k]])
    assert abs(got_loss - ref_loss) <= 1e-2 * abs(ref_loss) + 1e-3, \
        f"sub-step {k} (B={B}): loss {got_loss} vs oracle {ref_loss}"

    ref_g = p.orc.last_grads[str(k)]
    ref_w = O.trainable_params(p.orc.nets()[name])
    tens = _trainable(net)
    assert len(tens) == len(ref_g) == len(ref_w) == len(before)
    gd, ud = [], []
    for (w, dw), rg, rw, w0 in zip(tens, ref_g, ref_w, before):
        if rg.numel() == 0:
            gd.append((0.0, 0.0, 0.0))
            ud.append((0.0, 0.0, 0.0))
            continue
        gd.append(_dots(dw, rg.cuda()))
        ud.append(_dots(w - w0, rw.cuda() - w0))
    failures = []
    for what, dots, flat_min, tensor_min in (("gradient", gd, P.GRAD_FLAT, P.GRAD_TENSOR),
                                             ("update", ud, P.UPD_FLAT, P.UPD_TENSOR)):
        flat = _cos3(tuple(sum(d[i] for d in dots) for i in range(3)))
        total = sum(d[2] for d in dots) ** 0.5
        per = {i: _cos3(d) for i, d in enumerate(dots) if d[2] ** 0.5 >= 2e-2 * total}
        worst = min(per.values()) if per else 1.0
        _log({"substep": k, "net": name, "batch": B, "fused": bool(net.fuse_optimizer),
              "what": what, "flat_cosine": flat, "worst_tensor_cosine": worst,
              "loss": got_loss, "oracle_loss": ref_loss})
        if flat < flat_min:
            failures.append(f"sub-step {k} {name} (B={B}): flat {what} cosine {flat}")
        for i, c in per.items():
            if c < tensor_min:
                failures.append(f"sub-step {k} {name} (B={B}) tensor {i}: {what} cosine {c}")
    # BN moving statistics of the trained network (momentum 0.99 update with batch statistics)
    for L, ol in zip(net.layers, p.orc.nets()[name]):
        if L["kind"] == "bn":
            for key in ("moving_mean", "moving_var"):
                ref = ol[key].cuda()
                err = float((L[key] - ref).abs().max())
                assert err <= 1e-3 * float(ref.abs().max()) + 2e-4, \
                    f"sub-step {k} {name}: {key} off by {err}"
    if resync:
        p.sync(name)
    assert not failures, "; ".join(failures)


def _stage(p, B, seed):
    from cellcomm_b200 import ops
    x, z, r = P._inputs(VARIANT, Z, G, B, seed)
    masks = O.make_masks(VARIANT, Z, G, B, seed + 1)
    x16 = ops.alloc2d(B, G)
    x16.copy_(x)
    p.e.set_latents(z, r, B)
    return x16, p.orc.t(x), p.orc.t(z), p.orc.t(r), masks, P._dev_masks(masks)


@pytest.mark.parametrize("fused", [True, False])
def test_every_update_at_reference_batch_128(pair, fused):
    """B = 128 (src/__main__.py:44): all eight sub-steps in order, each update from identical
    weights.  fused = RMSprop in the wgrad epilogue (1-GPU default) / flat sweep (DP path)."""
    p, B = pair, 128
    p.e.set_fused_optimizer(fused, keep_grads=True)
    x16, xo, zo, ro, masks, dmasks = _stage(p, B, 11)
    for k in (1, 2, 3, 4, 5, 6, 7, 8):
        if k in P.UPDATES:
            _check_update(p, k, x16, xo, zo, ro, masks, dmasks, B)
            continue
        p.orc.substep(k, xo, zo, ro, masks)
        p.e.substep(k, x16, dmasks)
        torch.cuda.synchronize()
        if k == 5:   # generated cells: integer-valued, equal up to rounding flips of G's output
            got, ref = p.e.gen_cells[:B].float(), p.orc.gen_cells.cuda()
            flips = float((got != ref.to(torch.bfloat16).float()).float().mean())
            assert flips <= 2e-2, f"generated cells differ in {flips:.3%} of the entries"
            assert abs(float(got.sum()) - float(ref.sum())) <= 2e-2 * float(ref.sum()) + B
        else:        # trainings_encoding_prediction(batch): the encodings of the minibatch
            got, ref = p.e.gen_enc32[:B], p.orc.gen_enc.cuda()
            cos = torch.nn.functional.cosine_similarity(got.double(), ref.double(), dim=1)
            assert float(cos.min()) >= 0.999 and float((got - ref).abs().max()) <= 2e-2


def test_updates_at_saturating_batch_2048(pair):
    """B = 2048 (the bench's per-GPU batch): one G update (sub-step 2: G6 takes the M-fast
    raster of the fused epilogue), one E update (3) and one D update (8: Dx1 takes the N-fast
    raster), 2-CTA clusters and 192/256-wide tiles on every wide layer."""
    p, B = pair, 2048
    p.e.set_fused_optimizer(True, keep_grads=True)
    x16, xo, zo, ro, masks, dmasks = _stage(p, B, 23)
    p.orc.substep(7, xo, zo, ro, masks)         # D's real-pair input: E.predict(batch)
    p.e.substep(7, x16, dmasks)
    for k in (2, 3, 8):
        _check_update(p, k, x16, xo, zo, ro, masks, dmasks, B)


def test_free_running_step_at_batch_128(pair):
    """The whole trainings_step without re-synchronisation, production settings (fused
    optimiser, gradients not kept): six sub-step losses and the returned (g, e, d)."""
    p, B = pair, 128
    p.e.set_fused_optimizer(True, keep_grads=False)
    x16, xo, zo, ro, masks, dmasks = _stage(p, B, 31)
    ref = p.orc.trainings_step(xo, zo, ro, masks)
    got = p.e.train_step(x16, dmasks)
    torch.cuda.synchronize()
    six = [p.orc.last_losses[k] for k in ("1", "2", "3", "4", "6", "8")]
    for name, a, b in zip("123468", p.e.last_losses[:6].tolist(), six):
        assert abs(a - b) <= 1e-2 * abs(b) + 1e-3, f"sub-step {name} loss {a} vs oracle {b}"
    for a, b in zip(got, ref):
        assert abs(float(a) - b) <= 1e-2 * abs(b) + 2e-3
    for n in ("G", "E", "D"):
        p.sync(n)


def test_encode_pass_at_baseline_width(pair):
    """encoding_prediction on 1,000 cells x 33,694 genes (src/bigan_basic.py:29-30): the tile
    path the encode-all-cells interceptor uses (4096-row tiles -> here one 1000-row tile and
    a ragged 1000 = 512 + 488 split)."""
    from cellcomm_b200 import ops
    p, N = pair, 1000
    x, _, _ = P._inputs(VARIANT, Z, G, N, 41)
    ref = p.orc.encoding_prediction(x).cuda()
    x16 = ops.alloc2d(N, G)
    x16.copy_(x)
    out = torch.empty(N, Z, device="cuda")
    p.e.encode(x16, out32=out)
    out2 = torch.empty(N, Z, device="cuda")
    p.e.encode(x16[:512], out32=out2[:512])
    p.e.encode(x16[512:], out32=out2[512:])
    torch.cuda.synchronize()
    for got in (out, out2):
        cos = torch.nn.functional.cosine_similarity(got.double(), ref.double(), dim=1)
        assert float(cos.min()) >= 0.999, float(cos.min())
        assert float((got - ref).abs().max()) <= 2e-2
    assert float((out - out2).abs().max()) <= 1e-4     # tile boundaries: summation order only
