"""Host logic of cellcomm_b200.engine (graph wiring, freeze pattern, gradient routing, RMSprop
bookkeeping) against the oracle, with the kernels replaced by tests/ops_emulator.py.

CPU-only (`-m "not gpu"`).  The real kernels are checked on the GPU in test_parity_gpu.py.
"""
import numpy as np
import pytest
import torch

from oracle import bigan_oracle as O

import ops_emulator
from cellcomm_b200 import engine as eng


@pytest.fixture(autouse=True)
def _emulated_ops(monkeypatch):
    monkeypatch.setattr(eng, "ops", ops_emulator)
    yield


def _pair(variant, Z, G, B, seed=0, fused=True):
    orc = O.OracleBiGan(variant, Z, G, seed=seed, dtype=torch.float64)
    e = eng.BiGanEngine(variant, Z, G, max_batch=B, device="cpu", seed=seed)
    e.set_fused_optimizer(fused, keep_grads=True)
    for n in ("G", "E", "D"):
        e.nets[n].set_weights([w.numpy() for w in orc.get_weights(n)])
    return orc, e


def _inputs(variant, Z, G, B, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.poisson(torch.rand(B, G, generator=g) * 3, generator=g)
    if variant == "cont":
        z = torch.rand(B, Z, generator=g)
    else:
        z = torch.nn.functional.one_hot(torch.randint(0, Z, (B,), generator=g), Z).float()
    r = torch.rand(B, Z, generator=g)
    return x, z, r


def _close(a, b, rtol=2e-4, atol=2e-6, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    assert np.all(err <= tol), f"{what}: max err {err.max():.3e} (ref max {np.abs(b).max():.3e})"


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("variant,Z,G,B", [("cont", 3, 200, 16), ("classify", 4, 60, 12),
                                           ("cont", 8, 5, 3)])
def test_trainings_step_matches_oracle(variant, Z, G, B, fused):
    """fused: RMSprop applied inside the wgrad epilogue (single GPU); else one flat sweep"""
    orc, e = _pair(variant, Z, G, B, fused=fused)
    for step in range(2):
        x, z, r = _inputs(variant, Z, G, B, 100 + step)
        masks = O.make_masks(variant, Z, G, B, 7 + step)
        ref = orc.trainings_step(x, z, r, masks)
        e.set_latents(z, r, B)
        x16 = ops_emulator.alloc2d(B, G)
        x16.copy_(x)
        got = e.train_step(x16, masks)
        _close([float(v) for v in got], ref, rtol=1e-4, what=f"losses step {step}")
        six = e.last_losses[:6].tolist()
        _close(six, [orc.last_losses[k] for k in ("1", "2", "3", "4", "6", "8")], rtol=1e-4,
               what="per-substep losses")
        # gradients of the LAST update of each net (G: sub-step 2, E: 4, D: 8)
        for net, sub in (("G", "2"), ("E", "4"), ("D", "8")):
            n = e.nets[net]
            got_g = []
            for L in n.layers:
                got_g += [L["dw"], L["db"]] if L["kind"] == "dense" else [L["dgamma"], L["dbeta"]]
            for i, (a, b) in enumerate(zip(got_g, orc.last_grads[sub])):
                if b.numel() == 0:
                    assert a.numel() == 0
                    continue
                scale = float(b.abs().max()) + 1e-12
                _close(a.numpy() / scale, b.numpy() / scale, rtol=2e-3, atol=3e-4,
                       what=f"{net} grad {i} step {step}")
        # post-step weights, BN moving statistics
        for net in ("G", "E", "D"):
            for i, (a, b) in enumerate(zip(e.nets[net].get_weights(), orc.get_weights(net))):
                _close(a, b.numpy(), rtol=1e-3, atol=1e-4, what=f"{net} weight {i} step {step}")


def test_blocked_state_layout_is_invisible_outside_the_fused_epilogue(monkeypatch):
    """Fused-optimiser runs keep p32 / ms / mom of the wide kernels in the blocked layout
    (Net._state_layout).  With the narrow-layer threshold lowered so that this small model HAS
    blockable kernels (incl. two-segment ones whose second segment starts at a row that is not a
    multiple of 32): two training steps equal the row-major run exactly, get_weights /
    get_slots / snapshot+restore / set_weights see rows, the layers' row views are valid again
    after any of them, and the next step converts back."""
    monkeypatch.setattr(eng, "_NARROW_ELEMS", 4096)
    variant, Z, G, B = "cont", 3, 1400, 16

    def run(blocked):
        monkeypatch.setattr(eng, "_BLOCKED_STATE", blocked)
        orc, e = _pair(variant, Z, G, B, fused=True)
        assert bool(e.D.blockable) == blocked and bool(e.G.blockable) == blocked
        if blocked:
            segs = [e.D.layers[i]["in_widths"] for i in e.D.blockable]
            assert any(len(w) > 1 and w[0] % 32 for w in segs), segs
        outs = []
        for step in range(2):
            x, z, r = _inputs(variant, Z, G, B, 100 + step)
            masks = O.make_masks(variant, Z, G, B, 7 + step)
            e.set_latents(z, r, B)
            x16 = ops_emulator.alloc2d(B, G)
            x16.copy_(x)
            outs.append([float(v) for v in e.train_step(x16, masks)])
            assert e.D.state_blocked == blocked
            if step == 0:
                snap = e.snapshot_state()
                w_mid = {n: e.nets[n].get_weights() for n in ("G", "E", "D")}   # -> rows
                assert not e.D.state_blocked
                for L in e.D.layers:
                    if L["kind"] == "dense":     # row views valid: p16 is bf16(p32)
                        assert torch.equal(L["w16"].float(), L["w32"].to(L["w16"].dtype).float())
        return e, outs, w_mid, snap

    e0, o0, w0, _ = run(False)
    e1, o1, w1, snap = run(True)
    assert o0 == o1
    for n in ("G", "E", "D"):
        for a, b in zip(w0[n], w1[n]):
            assert np.array_equal(a, b)
        for a, b in zip(e0.nets[n].get_weights(), e1.nets[n].get_weights()):
            assert np.array_equal(a, b)
        for (a, c), (b, d) in zip(e0.nets[n].get_slots(), e1.nets[n].get_slots()):
            assert np.array_equal(a, b) and np.array_equal(c, d)
    # the snapshot was taken in the blocked layout: restoring it brings layout and data back
    e1.restore_state(snap)
    assert e1.D.state_blocked
    for n in ("G", "E", "D"):
        for a, b in zip(w1[n], e1.nets[n].get_weights()):
            assert np.array_equal(a, b)
    # padding rows / columns of the state stay zero through the conversions
    for i in e1.D.blockable:
        L = e1.D.layers[i]
        R = L["Kv"]
        assert all(v % 32 == 0 for v in L["seg_vrow"])
        reg = e1.D.ms[L["w_off"]:L["w_off"] + R * L["ld"]].view(R, L["ld"])
        assert float(reg[L["K"]:].abs().sum()) == 0.0 and float(reg[:, L["N"]:].abs().sum()) == 0.0


def test_blocked_layout_index_formula():
    """ops.state_rows_to_blocked implements the element order documented for
    cc_gemm_desc.rms_blocked (include/cellcomm_b200.h): element (r, c) of a [R, ld] array at
    ((r/32) * (ld/32) + c/32) * 1024 + ((c%32)/4) * 128 + (r%32) * 4 + c%4; the emulator's
    restatement agrees and the inverse restores the rows."""
    from cellcomm_b200 import ops as real_ops
    R, ld = 96, 192
    rows = torch.arange(R * ld, dtype=torch.float32).reshape(R, ld)
    flat = real_ops.state_rows_to_blocked(rows)
    assert torch.equal(flat, ops_emulator.state_rows_to_blocked(rows))
    r = torch.arange(R).view(-1, 1).expand(R, ld)
    c = torch.arange(ld).view(1, -1).expand(R, ld)
    idx = ((r // 32) * (ld // 32) + c // 32) * 1024 + ((c % 32) // 4) * 128 + (r % 32) * 4 + c % 4
    assert torch.equal(flat[idx.reshape(-1)], rows.reshape(-1))
    assert torch.equal(real_ops.state_blocked_to_rows(flat, R, ld), rows)


def test_predict_paths_match_oracle():
    variant, Z, G, B = "cont", 3, 120, 10
    orc, e = _pair(variant, Z, G, B)
    x, z, r = _inputs(variant, Z, G, B, 5)
    x16 = ops_emulator.alloc2d(B, G)
    x16.copy_(x)
    out32 = torch.zeros(B, Z)
    e.encode(x16, out32=out32)
    _close(out32.numpy(), orc.encoding_prediction(x).numpy(), what="encode")
    e.set_latents(z, r, B)
    g32 = torch.zeros(B, G)
    e.generate(B, out32=g32)
    _close(g32.numpy(), orc.generator_predict(z, r).numpy(), what="generate")
    p = torch.zeros(B, 1)
    e.discriminate(e.z32[:B], x16, p)
    _close(p.numpy(), orc.discriminator_predict(z, x).numpy(), what="discriminate")


def test_rng_dropout_mode_runs_and_advances_counter():
    e = eng.BiGanEngine("cont", 3, 64, max_batch=8, device="cpu", seed=1)
    x16 = ops_emulator.alloc2d(8, 64)
    x16.copy_(torch.rand(8, 64).round())
    e.draw_latents(8)
    g, ee, d = e.train_step(x16)
    assert all(np.isfinite(float(v)) for v in (g, ee, d))
    assert int(e.rng_counter.item()) == 1


def test_loss_scalar_behaves_like_a_float():
    a = eng.LossScalar(torch.tensor(1.5))
    acc = 0
    acc += a
    acc += a
    assert float(acc) == 3.0
    assert f"{acc:6.3f}" == " 3.000"
    assert float(sum((a, a, a))) == 4.5


@pytest.mark.parametrize("bucket_elems", [1 << 11, 1 << 14, 1 << 30])
def test_gradient_buckets_tile_the_kernel_region(monkeypatch, bucket_elems):
    """Data-parallel gradient buckets (engine.Net._build_buckets): the row pieces of every
    Dense kernel are disjoint, 64-element aligned (so 1/world shards of 2, 4 or 8 ranks stay
    16-byte aligned), cover every kernel element, respect Concatenate segment boundaries, and
    consecutive pieces group into contiguous buckets."""
    monkeypatch.setattr(eng, "_BUCKET_ELEMS", bucket_elems)
    e = eng.BiGanEngine("cont", 3, 700, max_batch=4, device="cpu", seed=0)
    for net in e.nets.values():
        covered = torch.zeros(net.small_off, dtype=torch.int32)
        for (layer, seg), pieces in net.pieces.items():
            L = net.layers[layer]
            seg_lo = sum(L["in_widths"][:seg])
            seg_hi = seg_lo + L["in_widths"][seg]
            assert pieces[0]["lo"] == seg_lo and pieces[-1]["hi"] == seg_hi
            for a, b in zip(pieces, pieces[1:]):
                assert a["hi"] == b["lo"] and (b["lo"] - seg_lo) % 64 == 0
            for pc in pieces:
                assert pc["start"] == L["w_off"] + pc["lo"] * L["ld"]
                assert pc["start"] % 64 == 0 and pc["end"] % 64 == 0
                assert pc["end"] >= L["w_off"] + pc["hi"] * L["ld"]
                covered[pc["start"]:pc["end"]] += 1
                bk = pc["bucket"]
                assert bk["start"] <= pc["start"] and pc["end"] <= bk["end"]
        assert int(covered.max()) <= 1
        for L in net.layers:
            if L["kind"] == "dense" and L["K"] * L["N"] > 0:
                assert bool((covered[L["w_off"]:L["w_off"] + L["K"] * L["ld"]] == 1).all())
        for a, b in zip(net.buckets, net.buckets[1:]):
            assert a["end"] <= b["start"]
        for bk in net.buckets:
            assert (bk["end"] - bk["start"]) % 64 == 0 and bk["pieces"] >= 1
        if bucket_elems == 1 << 11:
            assert max(len(p) for p in net.pieces.values()) >= 2     # big kernels were cut


@pytest.mark.parametrize("world", [2, 4, 8])
def test_routed_gradient_addresses_meet_the_optimiser_shards(monkeypatch, world):
    """Host-side addressing of the peer-memory data-parallel path, without a GPU: replay what
    the routed wgrad epilogue does with the (world, shard, off0, bases) tuples `Net._route`
    hands to cc_gemm (element rel = off0 + row*ld + col goes to bases[rel // shard] + 4*rel)
    on fake per-rank staging arrays, then check that the ranges `_launch_bucket` gives to
    cc_peer_rmsprop (this rank's 1/world slice of every bucket, read from staging slot q for
    every source rank q) hold exactly rank q's values for every kernel element."""
    monkeypatch.setattr(eng, "_BUCKET_ELEMS", 1 << 12)
    e = eng.BiGanEngine("cont", 3, 300, max_batch=4, device="cpu", seed=0)
    net = e.D
    n_flat = net.n_flat
    BASE = 1 << 44                                  # fake device address of rank r's staging
    stage = [np.full(world * n_flat, -1.0) for _ in range(world)]

    class FakeDist:
        world_size = world
        rank = 0

    kernel_elems = np.zeros(n_flat, dtype=bool)
    for me in range(world):
        FakeDist.rank = me
        net.dist = FakeDist
        net.peer = {"stage": [r * BASE for r in range(world)]}
        for pieces in net.pieces.values():
            for pc in pieces:
                W, shard, off0, bases = net._route(pc)
                L = net.layers[pc["layer"]]
                rows, ld, N = pc["hi"] - pc["lo"], L["ld"], L["N"]
                r_idx, c_idx = np.meshgrid(np.arange(rows), np.arange(N), indexing="ij")
                rel = off0 + r_idx * ld + c_idx
                owner = np.minimum(rel // shard, W - 1)
                for r in range(world):
                    sel = owner == r
                    addr = bases[r] + 4 * rel[sel]               # byte address on rank r
                    slot_elem = (addr - r * BASE) // 4            # element of rank r's staging
                    flat = pc["bucket"]["start"] + rel[sel]
                    assert np.all(slot_elem == me * n_flat + flat)
                    stage[r][slot_elem] = me * 1e9 + flat         # "rank me's gradient value"
                    kernel_elems[flat] = True
    # optimiser side: rank r updates [start + r*n, start + (r+1)*n) of every bucket
    covered = np.zeros(n_flat, dtype=bool)
    for bk in net.buckets:
        n = (bk["end"] - bk["start"]) // world
        assert n * world == bk["end"] - bk["start"] and n % 8 == 0
        for r in range(world):
            lo = bk["start"] + r * n
            offs = np.arange(lo, lo + n)
            live = kernel_elems[offs]                              # (padding is never written)
            for q in range(world):
                got = stage[r][q * n_flat + offs]
                assert np.all(got[live] == q * 1e9 + offs[live])
            covered[offs] = True
    assert np.all(covered[kernel_elems])
