"""Tail kernels (dropout, BN, losses, RMSprop, ...) through the C ABI vs plain PyTorch fp32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from cellcomm_b200 import ops
    return ops


def _dev(t, ops):
    out = ops.alloc2d(t.shape[0], t.shape[1], dtype=t.dtype)
    out.copy_(t)
    return out


def _r(rows, cols, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(rows, cols, generator=g) * scale).to(torch.bfloat16)


def close16(got, ref, rtol=2 ** -7, atol=1e-6):
    got, ref = got.float().cpu(), ref.float().cpu()
    bad = ((got - ref).abs() > atol + rtol * ref.abs()).sum().item()
    assert bad == 0, f"{bad} mismatches, max err {(got - ref).abs().max().item()}"


@pytest.mark.parametrize("rows,cols", [(128, 3369), (5, 3), (300, 103), (2048, 256)])
def test_colsum_and_bn(rows, cols):
    ops = _ops()
    x = _r(rows, cols, 1, 2.0)
    dx = _dev(x, ops)
    out = torch.empty(cols, device="cuda")
    ops.colsum(dx, out)
    ref = x.float().sum(0)
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-2 * rows ** 0.5)
    # BN train
    sums = torch.empty(2 * cols, device="cuda")
    ops.bn_stats(dx, sums)
    gamma = torch.rand(cols, device="cuda") + 0.5
    beta = torch.randn(cols, device="cuda")
    mm, mv = torch.zeros(cols, device="cuda"), torch.ones(cols, device="cuda")
    sm, sr = torch.empty(cols, device="cuda"), torch.empty(cols, device="cuda")
    y = ops.alloc2d(rows, cols)
    ops.bn_train_apply(dx, y, sums, rows, gamma, beta, 1e-3, 0.99, mm, mv, sm, sr)
    xf = x.float()
    mean, var = xf.mean(0), xf.var(0, unbiased=False)
    ref_y = (xf - mean) / torch.sqrt(var + 1e-3) * gamma.cpu() + beta.cpu()
    close16(y, ref_y, atol=2e-2)
    assert torch.allclose(mm.cpu(), 0.01 * mean, atol=1e-4)
    assert torch.allclose(mv.cpu(), 0.99 + 0.01 * var, rtol=1e-3, atol=1e-4)
    # BN backward
    dy = _r(rows, cols, 2)
    ddy = _dev(dy, ops)
    sums2 = torch.empty(2 * cols, device="cuda")
    ops.bn_bwd_stats(ddy, dx, sm, sr, sums2)
    dxo = ops.alloc2d(rows, cols)
    dg, db = torch.empty(cols, device="cuda"), torch.empty(cols, device="cuda")
    ops.bn_bwd_apply(ddy, dx, dxo, gamma, sm, sr, sums2, rows, dg, db)
    xt = xf.clone().requires_grad_(True)
    gt, bt = gamma.cpu().clone().requires_grad_(True), beta.cpu().clone().requires_grad_(True)
    yy = (xt - xt.mean(0)) / torch.sqrt(xt.var(0, unbiased=False) + 1e-3) * gt + bt
    yy.backward(dy.float())
    close16(dxo, xt.grad, rtol=3e-2, atol=3e-2)
    assert torch.allclose(db.cpu(), bt.grad, rtol=1e-3, atol=1e-2 * rows ** 0.5)
    assert torch.allclose(dg.cpu(), gt.grad, rtol=2e-2, atol=2e-2 * rows ** 0.5)
    # BN inference + its backward
    yi = ops.alloc2d(rows, cols)
    ops.bn_infer(dx, yi, gamma, beta, mm, mv, 1e-3)
    ref_i = (xf - mm.cpu()) / torch.sqrt(mv.cpu() + 1e-3) * gamma.cpu() + beta.cpu()
    close16(yi, ref_i, atol=2e-2)
    dxi = ops.alloc2d(rows, cols)
    ops.bn_infer_bwd(ddy, dxi, gamma, mv, 1e-3)
    close16(dxi, dy.float() * gamma.cpu() / torch.sqrt(mv.cpu() + 1e-3), atol=1e-3)


def test_dropout_explicit_mask_and_rng():
    ops = _ops()
    rows, cols = 130, 1001
    x = _r(rows, cols, 3)
    dx = _dev(x, ops)
    mask = (torch.rand(rows, cols) > 0.15).to(torch.uint8).cuda()
    out = ops.alloc2d(rows, cols)
    ops.dropout(dx, out, 0.15, mask=mask)
    close16(out, x.float() * mask.cpu().float() / 0.85)
    # RNG path: same call twice => same mask; matches cc_dropout_mask; keep rate ~ 1-rate
    counter = torch.tensor([7], dtype=torch.int64, device="cuda")
    o1, o2 = ops.alloc2d(rows, cols), ops.alloc2d(rows, cols)
    ops.dropout(dx, o1, 0.15, seed=123, counter=counter, stream_id=5)
    ops.dropout(dx, o2, 0.15, seed=123, counter=counter, stream_id=5)
    assert torch.equal(o1, o2)
    m = torch.empty(rows, cols, dtype=torch.uint8, device="cuda")
    ops.dropout_mask(m, 0.15, seed=123, counter=counter, stream_id=5)
    close16(o1, x.float() * m.cpu().float() / 0.85)
    keep = m.float().mean().item()
    assert abs(keep - 0.85) < 0.01
    ops.counter_add(counter, 1)
    m2 = torch.empty_like(m)
    ops.dropout_mask(m2, 0.15, seed=123, counter=counter, stream_id=5)
    assert not torch.equal(m, m2)
    assert counter.item() == 8


def test_uniform_range_and_mean():
    ops = _ops()
    u = ops.alloc2d(1000, 3, dtype=torch.float32)
    ops.uniform(out32=u, seed=9, stream_id=1)
    assert u.min().item() >= 0.0 and u.max().item() < 1.0
    assert abs(u.mean().item() - 0.5) < 0.03


def test_act_bwd_copy_cast_round():
    ops = _ops()
    rows, cols = 77, 515
    y = torch.sigmoid(_r(rows, cols, 4).float()).to(torch.bfloat16)
    dy = _r(rows, cols, 5)
    out = ops.alloc2d(rows, cols)
    ops.act_bwd(_dev(dy, ops), _dev(y, ops), out, ops.ACT_SIGMOID)
    close16(out, dy.float() * y.float() * (1 - y.float()))
    yr = torch.relu(_r(rows, cols, 6).float()).to(torch.bfloat16)
    ops.act_bwd(_dev(dy, ops), _dev(yr, ops), out, ops.ACT_RELU)
    close16(out, dy.float() * (yr.float() > 0).float())
    # copy into a column slice with accumulate
    dst = ops.alloc2d(rows, cols + 9)
    dst.fill_(1.0)
    ops.copy2d(_dev(dy, ops), dst[:, 9:], beta=1)
    close16(dst[:, 9:], dy.float() + 1.0, rtol=2 ** -7)
    assert torch.all(dst[:, :9] == 1.0)
    # casts
    f = torch.randn(rows, cols, device="cuda")
    c16 = ops.alloc2d(rows, cols)
    ops.cast_f32_to_bf16(f, c16)
    assert torch.equal(c16, f.to(torch.bfloat16))
    f2 = torch.empty(rows, cols, device="cuda")
    ops.cast_bf16_to_f32(c16, f2, 255.0)
    assert torch.equal(f2, c16.float() * 255.0)
    # round half to even: golden from test/bigans_basic_test.py:36-42
    g = torch.tensor([[0.3, 12.59939265, 2.4894546, 0.01], [0.9, 4.7007282, 0, 2.07244989],
                      [0.5, 1.5, 2.5, 3.5]]).to(torch.bfloat16)
    r32 = torch.empty(3, 4, device="cuda")
    ops.round_half_even(_dev(g, ops), out32=r32)
    assert r32.cpu().tolist() == [[0, 13, 2, 0], [1, 5, 0, 2], [0, 2, 2, 4]]


def test_softmax_argmax():
    ops = _ops()
    rows, cols = 300, 10
    x = _r(rows, cols, 7, 3.0)
    y16, y32 = ops.alloc2d(rows, cols), ops.alloc2d(rows, cols, dtype=torch.float32)
    ops.softmax_fwd(_dev(x, ops), y16, y32)
    ref = torch.softmax(x.float(), -1)
    assert torch.allclose(y32.cpu(), ref, rtol=1e-4, atol=1e-6)
    dy = _r(rows, cols, 8)
    dxo = ops.alloc2d(rows, cols)
    ops.softmax_bwd(_dev(dy, ops), y16, dxo)
    yf = y16.float().cpu()
    close16(dxo, yf * (dy.float() - (dy.float() * yf).sum(-1, keepdim=True)), atol=1e-3)
    # golden: test/bigans_cc_test.py:77-85
    p = torch.tensor([[0.1, 0.3], [0.7, 0.3], [0.001, 0.99]], device="cuda")
    oh = torch.empty(3, 2, device="cuda")
    ops.argmax_onehot(p, out32=oh)
    assert oh.cpu().tolist() == [[0, 1], [1, 0], [0, 1]]


def test_losses():
    ops = _ops()
    rows = 128
    logits = torch.randn(rows, 1, device="cuda") * 3
    for target in (0.95, 0.0):
        loss = torch.zeros(1, device="cuda")
        dz = ops.alloc2d(rows, 1)
        ops.bce_fwd_bwd(logits, target, rows, loss, dz, from_logits=True)
        lt = logits.cpu().double().requires_grad_(True)
        ref = torch.nn.functional.binary_cross_entropy_with_logits(
            lt, torch.full_like(lt, target))
        ref.backward()
        assert abs(loss.item() - ref.item()) < 1e-5 * max(1, abs(ref.item()))
        close16(dz, lt.grad.float(), atol=1e-6)
    # probabilities path
    p = torch.sigmoid(logits)
    loss = torch.zeros(1, device="cuda")
    ops.bce_fwd_bwd(p, 0.95, rows, loss, None, from_logits=False)
    ref = torch.nn.functional.binary_cross_entropy(p.cpu().double(),
                                                   torch.full((rows, 1), 0.95, dtype=torch.double))
    assert abs(loss.item() - ref.item()) < 1e-4
    # mse
    cols = 3370
    pred, tgt = _r(rows, cols, 9), _r(rows, cols, 10)
    loss = torch.zeros(1, device="cuda")
    dp = ops.alloc2d(rows, cols, dtype=torch.float32)
    ops.mse_fwd_bwd(_dev(pred, ops), rows, loss, target=_dev(tgt, ops), dpred=dp)
    d = pred.float() - tgt.float()
    assert abs(loss.item() - (d * d).mean().item()) < 1e-4 * (d * d).mean().item()
    close16(dp, 2 * d / (rows * cols), atol=1e-9)
    t32 = torch.rand(rows, 3, device="cuda")
    p3 = _r(rows, 3, 11)
    loss = torch.zeros(1, device="cuda")
    ops.mse_fwd_bwd(_dev(p3, ops), rows, loss, target=t32)
    assert abs(loss.item() - ((p3.float() - t32.cpu()) ** 2).mean().item()) < 1e-5


def test_rmsprop_matches_keras_formula():
    ops = _ops()
    rows, cols = 50, 103
    ld = ops.pad_ld(cols)
    p32 = torch.zeros(rows, ld, device="cuda")[:, :cols]
    p32.copy_(torch.randn(rows, cols))
    p16 = torch.zeros(rows, ld, dtype=torch.bfloat16, device="cuda")[:, :cols]
    g = torch.zeros(rows, ld, device="cuda")[:, :cols]
    ms = torch.zeros(rows, ld, device="cuda")[:, :cols]
    mom = torch.zeros(rows, ld, device="cuda")[:, :cols]
    w = p32.cpu().double().numpy().copy()
    s = np.zeros_like(w)
    m = np.zeros_like(w)
    lr, rho, mo, eps = 0.0075, 0.85, 0.1, 1e-7
    for it in range(3):
        gg = torch.randn(rows, cols) * 10 ** (-it * 2)
        g.copy_(gg)
        ops.rmsprop_step(p32, p16, g, ms, mom, lr, rho, mo, eps)
        gn = gg.double().numpy()
        s = rho * s + (1 - rho) * gn * gn
        m = mo * m + lr * gn / np.sqrt(s + eps)
        w = w - m
    assert np.allclose(p32.cpu().numpy(), w, rtol=1e-5, atol=1e-6)
    assert np.allclose(ms.cpu().numpy(), s, rtol=1e-5, atol=1e-12)
    assert np.allclose(mom.cpu().numpy(), m, rtol=1e-4, atol=1e-8)
    assert torch.equal(p16, p32.to(torch.bfloat16))
    # 1-D parameter (bias)
    b32, bg = torch.randn(77, device="cuda"), torch.randn(77, device="cuda")
    bms, bmom = torch.zeros(77, device="cuda"), torch.zeros(77, device="cuda")
    ref = b32.cpu() - lr * bg.cpu() / torch.sqrt(0.15 * bg.cpu() ** 2 + eps)
    ops.rmsprop_step(b32, None, bg, bms, bmom, lr, rho, mo, eps)
    assert torch.allclose(b32.cpu(), ref, rtol=1e-5, atol=1e-6)


def test_fp32_operands_bias_grad_and_split():
    """dtype-mask paths: fp32 activations / gradients, fp32 bias gradient, bf16 hi+lo split."""
    ops = _ops()
    rows, cols = 96, 333
    x32 = torch.randn(rows, cols, device="cuda") * 0.05 + 0.5
    # split: hi + lo reproduces x to ~16 bits
    hi, lo = ops.alloc2d(rows, cols), ops.alloc2d(rows, cols)
    ops.split_bf16(x32, hi, lo)
    assert torch.equal(hi, x32.to(torch.bfloat16))
    assert (hi.float() + lo.float() - x32).abs().max().item() <= 2 ** -16
    # fp32 -> fp32 dropout with explicit mask, fp32 -> bf16 BN
    mask = (torch.rand(rows, cols, device="cuda") > 0.1).to(torch.uint8)
    d32 = torch.empty(rows, cols, device="cuda")
    ops.dropout(x32, d32, 0.1, mask=mask)
    assert torch.allclose(d32, x32 * mask.float() / 0.9, rtol=1e-6, atol=1e-7)
    sums = torch.empty(2 * cols, device="cuda")
    ops.bn_stats(d32, sums)
    gamma, beta = torch.ones(cols, device="cuda"), torch.zeros(cols, device="cuda")
    mm, mv = torch.zeros(cols, device="cuda"), torch.ones(cols, device="cuda")
    sm, sr = torch.empty(cols, device="cuda"), torch.empty(cols, device="cuda")
    y = ops.alloc2d(rows, cols)
    ops.bn_train_apply(d32, y, sums, rows, gamma, beta, 1e-3, 0.99, mm, mv, sm, sr)
    ref = (d32 - d32.mean(0)) / torch.sqrt(d32.var(0, unbiased=False) + 1e-3)
    close16(y, ref.cpu(), atol=2e-2)
    # fp32 dy through BN backward into fp32 dx
    dy32 = torch.randn(rows, cols, device="cuda")
    s2 = torch.empty(2 * cols, device="cuda")
    ops.bn_bwd_stats(dy32, d32, sm, sr, s2)
    dx32 = torch.empty(rows, cols, device="cuda")
    ops.bn_bwd_apply(dy32, d32, dx32, gamma, sm, sr, s2, rows)
    xt = d32.clone().requires_grad_(True)
    yy = (xt - xt.mean(0)) / torch.sqrt(xt.var(0, unbiased=False) + 1e-3)
    yy.backward(dy32)
    assert torch.allclose(dx32, xt.grad, rtol=2e-3, atol=2e-3)
    # act_bwd fp32 dy, fp32 y -> bf16 dz ; bias_grad from the fp32 operands
    ysig = torch.sigmoid(torch.randn(rows, cols, device="cuda"))
    dz = ops.alloc2d(rows, cols)
    ops.act_bwd(dy32, ysig, dz, ops.ACT_SIGMOID)
    close16(dz, (dy32 * ysig * (1 - ysig)).cpu())
    db = torch.empty(cols, device="cuda")
    ops.bias_grad(dy32, ysig, ops.ACT_SIGMOID, db)
    assert torch.allclose(db, (dy32 * ysig * (1 - ysig)).sum(0), rtol=1e-4, atol=1e-4)
    ops.bias_grad(dy32, ysig, ops.ACT_NONE, db)
    assert torch.allclose(db, dy32.sum(0), rtol=1e-4, atol=1e-4)
    # one launch for "dz as hi + lo, plus the bias gradient" (accumulating into db), and the
    # frozen-layer form without a bias gradient: same hi / lo as act_bwd + split_bf16
    v = dy32 * ysig * (1 - ysig)
    want_hi, want_lo = ops.alloc2d(rows, cols), ops.alloc2d(rows, cols)
    ops.split_bf16(v, want_hi, want_lo)
    for out in (torch.full((cols,), 2.0, device="cuda"), None):
        zh, zl = ops.alloc2d(rows, cols), ops.alloc2d(rows, cols)
        ops.bias_grad(dy32, ysig, ops.ACT_SIGMOID, out, dz=zh, beta=1, dz_lo=zl)
        # (torch and the kernel multiply dy * y * (1 - y) in different orders: the fp32 products
        # differ in the last bit, which is exactly what the low-order term records)
        assert (zh.float() - want_hi.float()).abs().max().item() <= 2 ** -8 * v.abs().max().item()
        assert (zh.float() + zl.float() - v).abs().max().item() <= 2 ** -15 * v.abs().max().item()
        if out is not None:
            assert torch.allclose(out, 2.0 + v.sum(0), rtol=1e-4, atol=1e-4)
    # copy2d as cast + scale + accumulate across dtypes
    acc = torch.ones(rows, cols, device="cuda")
    ops.copy2d(hi, acc, beta=1, scale=2.0)
    assert torch.allclose(acc, 1 + 2 * hi.float())
    # dgrad into an fp32 gradient buffer with accumulate
    K, N = 200, 150
    dzz = ops.alloc2d(rows, N)
    dzz.normal_()
    w = ops.alloc2d(K, N)
    w.normal_()
    g32 = torch.ones(rows, K, device="cuda")
    ops.dense_dgrad([dzz], [w], g32, beta=1)
    assert torch.allclose(g32, 1 + dzz.float() @ w.float().t(), rtol=1e-3, atol=1e-2)
    # wgrad with hi/lo x segments and two dz terms
    xw = torch.randn(rows, K, device="cuda")
    xh, xl = ops.alloc2d(rows, K), ops.alloc2d(rows, K)
    ops.split_bf16(xw, xh, xl)
    dw = torch.empty(K, N, device="cuda")
    ops.dense_wgrad([xh, xl], dzz, dw)
    assert torch.allclose(dw, xw.t() @ dzz.float(), rtol=1e-3, atol=2e-3)


def test_peer_rmsprop_two_ranks_emulated_on_one_gpu():
    """csrc/peer_optimizer.cu with both "ranks" on one device: every rank sums both gradient
    buffers (P2P loads), updates its shard of the fp32 master / slots and stores the bf16
    result into BOTH weight buffers; the flags order signal -> kernel -> signal -> wait.
    Reference: Keras RMSprop(momentum) on the summed gradient, fp32."""
    import torch
    from cellcomm_b200 import ops
    n, W = 4096, 2
    lr, rho, mu, eps = 0.0075, 0.85, 0.1, 1e-7
    g = torch.Generator(device="cuda").manual_seed(3)
    grads = [torch.randn(n, device="cuda", generator=g) for _ in range(W)]
    p32 = [torch.randn(n, device="cuda", generator=g) for _ in range(W)]
    p32[1].copy_(p32[0])                                  # replicas start identical
    ms = [torch.rand(n, device="cuda", generator=g) for _ in range(W)]
    ms[1].copy_(ms[0])
    mom = [torch.randn(n, device="cuda", generator=g) * 0.01 for _ in range(W)]
    mom[1].copy_(mom[0])
    p16 = [torch.zeros(n, dtype=torch.bfloat16, device="cuda") for _ in range(W)]
    flags = [torch.zeros(2 * 2 * W, dtype=torch.int32, device="cuda") for _ in range(W)]
    w0, s0, m0 = p32[0].clone(), ms[0].clone(), mom[0].clone()

    def fptr(on, kind, bucket, src):
        return flags[on].data_ptr() + 4 * ((kind * 2 + bucket) * W + src)

    epoch, start, count = 1, 64, n - 128                  # a sub-range; the rest must not move
    half = count // W
    # one stream plays both ranks in protocol order (two spinning kernels on one device could
    # be serialised onto one hardware queue): ready signals, kernels, done signals, waits
    for r in range(W):
        ops.peer_signal([fptr(t, 0, 0, r) for t in range(W)], epoch)
    for r in range(W):
        ops.peer_rmsprop(W, r, [t.data_ptr() for t in grads], [t.data_ptr() for t in p16],
                         p32[r], ms[r], mom[r], start + r * half, half, True, lr, rho, mu, eps,
                         fptr(r, 0, 0, 0), epoch)
        ops.peer_signal([fptr(t, 1, 0, r) for t in range(W)], epoch)
    for r in range(W):
        ops.peer_wait(fptr(r, 1, 0, 0), W, epoch)
    torch.cuda.synchronize()
    gsum = grads[0] + grads[1]
    s_ref = rho * s0 + (1 - rho) * gsum * gsum
    m_ref = mu * m0 + lr * gsum / torch.sqrt(s_ref + eps)
    w_ref = w0 - m_ref
    for r in range(W):
        sl = slice(start + r * half, start + (r + 1) * half)
        assert torch.allclose(ms[r][sl], s_ref[sl], rtol=1e-5, atol=1e-7)
        assert torch.allclose(mom[r][sl], m_ref[sl], rtol=2e-4, atol=1e-7)
        assert torch.allclose(p32[r][sl], w_ref[sl], rtol=1e-5, atol=1e-6)
        other = slice(start + (1 - r) * half, start + (2 - r) * half)
        assert torch.equal(p32[r][other], w0[other])      # the peer's shard is not mine to touch
        assert torch.equal(p32[r][:start], w0[:start]) and torch.equal(p32[r][start + count:], w0[start + count:])
    # both weight buffers hold bf16(updated master) over the whole range, bit-identical
    full = torch.cat([p32[0][start:start + half], p32[1][start + half:start + count]])
    assert torch.equal(p16[0][start:start + count], full.to(torch.bfloat16))
    assert torch.equal(p16[0], p16[1])
    assert torch.count_nonzero(p16[0][:start]) == 0


def test_peer_rmsprop_replicated_tail():
    """broadcast = 0 (biases / BN parameters): every rank updates the whole range from the sum
    of all gradients and stores only its own bf16 copy."""
    import torch
    from cellcomm_b200 import ops
    n, W = 512, 2
    grads = [torch.full((n,), 0.5, device="cuda"), torch.full((n,), 1.5, device="cuda")]
    p32, ms, mom = torch.ones(n, device="cuda"), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    p16 = [torch.zeros(n, dtype=torch.bfloat16, device="cuda") for _ in range(W)]
    flags = torch.full((W,), 7, dtype=torch.int32, device="cuda")      # already signalled
    ops.peer_rmsprop(W, 1, [t.data_ptr() for t in grads], [t.data_ptr() for t in p16], p32, ms, mom,
                     0, n, False, 0.0075, 0.85, 0.1, 1e-7, flags.data_ptr(), 7)
    torch.cuda.synchronize()
    s = 0.15 * 4.0
    w = 1.0 - 0.0075 * 2.0 / (s + 1e-7) ** 0.5
    assert torch.allclose(p32, torch.full_like(p32, w), rtol=1e-5)
    assert torch.equal(p16[1], p32.to(torch.bfloat16))
    assert torch.count_nonzero(p16[0]) == 0               # rank 0's copy is rank 0's job
