"""TEST INFRASTRUCTURE — a torch-CPU stand-in for `cellcomm_b200.ops`.

It lets the `-m "not gpu"` suite drive the *host* logic of cellcomm_b200.engine (graph
wiring, freeze pattern, gradient routing, optimiser bookkeeping, data-parallel reductions)
against the oracle without a GPU.  It is never imported by the product; the product path
(cellcomm_b200.ops) has no fallback and fails loudly without the CUDA library.

Activations are kept in float32 here (COMPUTE_DTYPE) so host-logic errors show up at 1e-5
instead of hiding under bf16 rounding; kernel numerics are the GPU tests' job.
"""
import torch

COMPUTE_DTYPE = torch.float32
ACT_NONE, ACT_SIGMOID, ACT_RELU = 0, 1, 2
_launches = [0]


def launch_count():
    return _launches[0]


def pad_ld(cols, mult=64):
    return max(mult, (cols + mult - 1) // mult * mult)


def alloc2d(rows, cols, dtype=torch.bfloat16, device="cpu", zero=True):
    if dtype == torch.bfloat16:
        dtype = COMPUTE_DTYPE
    buf = torch.zeros((max(rows, 1), pad_ld(cols)), dtype=dtype, device="cpu")
    return buf[:rows, :cols]


def _act(v, act):
    if act == ACT_SIGMOID:
        return torch.sigmoid(v)
    if act == ACT_RELU:
        return torch.relu(v)
    return v


def _store(dst, v, beta=0):
    _launches[0] += 1
    if dst is None:
        return
    if beta:
        v = v + dst.to(v.dtype)
    dst.copy_(v.to(dst.dtype))


def dense_fwd(xs, w16, row_offsets, bias, act, out16=None, out32=None, hi_lo=None):
    ws = list(w16) if isinstance(w16, (list, tuple)) else [w16] * len(xs)
    acc = 0
    for x, w, ro in zip(xs, ws, row_offsets):
        acc = acc + x.double() @ w[ro:ro + x.shape[1]].double()
    v = _act(acc + bias.double(), act).float()
    _store(out16, v)
    if out32 is not None:
        out32.copy_(v)
    if hi_lo is not None:
        split_bf16(v, hi_lo[0], hi_lo[1])


def dense_dgrad(dzs, ws16, out16, *, dact_y=None, dact=0, alpha=1.0, beta=0):
    acc = 0
    for dz, w in zip(dzs, ws16):
        acc = acc + dz.double() @ w.double().t()
    acc = acc * alpha
    if dact_y is not None and dact:
        y = dact_y.double()
        acc = acc * (y * (1 - y) if dact == ACT_SIGMOID else (y > 0).double())
    _store(out16, acc.float(), beta)


def state_rows_to_blocked(rows):
    """cellcomm_b200.ops.state_rows_to_blocked (the documented index formula, restated)"""
    R, ld = rows.shape
    return rows.reshape(R // 32, 32, ld // 32, 8, 4).permute(0, 2, 3, 1, 4).reshape(-1)


def state_blocked_to_rows(flat, R, ld):
    return flat.reshape(R // 32, ld // 32, 8, 32, 4).permute(0, 3, 1, 2, 4).reshape(R, ld)


def dense_wgrad(x, dz, dw32, beta=0, rms=None, route=None, rms_row0=None, rms_lo=None):
    assert route is None, "routed outputs need peer memory (GPU only)"
    xs = list(x) if isinstance(x, (list, tuple)) else [x]
    dzs = list(dz) if isinstance(dz, (list, tuple)) else [dz] * len(xs)
    acc = 0
    for t, d in zip(xs, dzs):
        acc = acc + t.double().t() @ d.double()
    g = acc.float()
    if rms is not None and rms_row0 is not None:
        # blocked fp32 state: the layer's flat arrays; update rows [row0, row0 + K) in place
        p32, p16, ms, mom, lr, rho, momentum, eps = rms
        K, N = p16.shape
        ld = p16.stride(0)
        R = p32.numel() // ld
        assert R % 32 == 0 and R * ld == p32.numel() and rms_row0 + K <= R
        rows = [state_blocked_to_rows(t, R, ld).clone() for t in (p32, ms, mom)]
        sl = slice(rms_row0, rms_row0 + K)
        rmsprop_step(rows[0][sl, :N], p16, g, rows[1][sl, :N], rows[2][sl, :N], lr, rho, momentum, eps)
        for t, r in zip((p32, ms, mom), rows):
            t.copy_(state_rows_to_blocked(r))
        if rms_lo is not None:
            w = rows[0][sl, :N]
            _store(rms_lo, w - w.to(p16.dtype).float())
    elif rms is not None:
        p32, p16, ms, mom, lr, rho, momentum, eps = rms
        rmsprop_step(p32, p16, g, ms, mom, lr, rho, momentum, eps)
    _store(dw32, g, beta)


def bias_grad(dy, y, act, out32, dz=None, beta=0, dz_lo=None):
    yy = y.double()
    d = yy * (1 - yy) if act == ACT_SIGMOID else ((yy > 0).double() if act == ACT_RELU else 1.0)
    v = (dy.double() * d).float()
    _store(out32, v.double().sum(0).float(), beta)
    if dz_lo is not None:
        split_bf16(v, dz, dz_lo)
    elif dz is not None:
        _store(dz, v)


def split_bf16(x, hi, lo):
    h = x.float().to(torch.bfloat16).float()
    _store(hi, h)
    _store(lo, x.float() - h)


def bias_act(bias, act, rows, out16=None, out32=None):
    t = out16 if out16 is not None else out32
    v = _act(bias.float(), act).unsqueeze(0).expand(rows, t.shape[1])
    _store(out16, v)
    if out32 is not None:
        out32.copy_(v)


def gather_rows(rowptr, colidx, values, n_cols, *, row_idx=None, row_start=0, n_rows=None,
                out16=None, out32=None):
    if n_rows is None:
        n_rows = row_idx.shape[0]
    for i in range(n_rows):
        r = int(row_idx[i]) if row_idx is not None else row_start + i
        b, e = int(rowptr[r]), int(rowptr[r + 1])
        for out in (out16, out32):
            if out is not None:
                out[i].zero_()
                out[i, colidx[b:e].long()] = values[b:e].to(out.dtype)
    _launches[0] += 1


def colsum(x16, out32, beta=0):
    _store(out32, x16.double().sum(0).float(), beta)


def dropout(x16, out16, rate, *, mask=None, seed=0, counter=None, stream_id=0):
    if mask is None:
        g = torch.Generator().manual_seed((int(seed) * 1000003 + int(counter.item()) * 4099 +
                                          int(stream_id)) % (2 ** 62))
        mask = (torch.rand(x16.shape, generator=g) >= rate)
    _store(out16, x16.float() * mask.float() / (1.0 - rate))


def uniform(out32=None, out16=None, *, seed=0, counter=None, stream_id=0):
    t = out32 if out32 is not None else out16
    c = int(counter.item()) if counter is not None else 0
    g = torch.Generator().manual_seed((int(seed) * 1000003 + c * 4099 + int(stream_id)) % (2 ** 62))
    u = torch.rand(t.shape, generator=g)
    _store(out32, u)
    _store(out16, u)


def counter_add(counter, inc=1):
    counter += inc


def act_bwd(dy16, y16, dz16, act):
    y = y16.float()
    d = y * (1 - y) if act == ACT_SIGMOID else ((y > 0).float() if act == ACT_RELU else 1.0)
    if dz16.shape[1] == 0:
        return
    _store(dz16, dy16.float() * d)


def copy2d(src16, dst16, beta=0, scale=1.0):
    _store(dst16, src16.float() * scale, beta)


def cast_f32_to_bf16(src32, dst16):
    _store(dst16, src32)


def cast_bf16_to_f32(src16, dst32, scale=1.0):
    _store(dst32, src16.float() * scale)


def bn_stats(x16, sums):
    n = x16.shape[1]
    x = x16.double()
    sums[:n] = x.sum(0).float()
    sums[n:2 * n] = (x * x).sum(0).float()
    _launches[0] += 1


def bn_train_apply(x16, y16, sums, n_total, gamma, beta, eps, momentum, moving_mean, moving_var,
                   save_mean, save_rstd):
    n = x16.shape[1]
    mean = sums[:n].double() / n_total
    var = (sums[n:2 * n].double() / n_total - mean * mean).clamp_min(0)
    save_mean[:n] = mean.float()
    save_rstd[:n] = (1.0 / torch.sqrt(var + eps)).float()
    moving_mean.mul_(momentum).add_(mean.float() * (1 - momentum))
    moving_var.mul_(momentum).add_(var.float() * (1 - momentum))
    _store(y16, ((x16.double() - mean) / torch.sqrt(var + eps) * gamma.double() + beta.double()).float())


def bn_infer(x16, y16, gamma, beta, moving_mean, moving_var, eps):
    _store(y16, (x16.float() - moving_mean) / torch.sqrt(moving_var + eps) * gamma + beta)


def bn_bwd_stats(dy16, x16, save_mean, save_rstd, sums2):
    n = x16.shape[1]
    dy = dy16.double()
    xhat = (x16.double() - save_mean[:n].double()) * save_rstd[:n].double()
    sums2[:n] = dy.sum(0).float()
    sums2[n:2 * n] = (dy * xhat).sum(0).float()
    _launches[0] += 1


def bn_bwd_apply(dy16, x16, dx16, gamma, save_mean, save_rstd, sums2, n_total, dgamma=None,
                 dbeta=None):
    n = x16.shape[1]
    if dbeta is not None:
        dbeta.copy_(sums2[:n])
    if dgamma is not None:
        dgamma.copy_(sums2[n:2 * n])
    if dx16 is None:
        return
    rstd = save_rstd[:n].double()
    xhat = (x16.double() - save_mean[:n].double()) * rstd
    dx = gamma.double() * rstd * (dy16.double() - sums2[:n].double() / n_total
                                  - xhat * sums2[n:2 * n].double() / n_total)
    _store(dx16, dx.float())


def bn_infer_bwd(dy16, dx16, gamma, moving_var, eps):
    _store(dx16, dy16.float() * gamma / torch.sqrt(moving_var + eps))


def softmax_fwd(x16, y16=None, y32=None):
    v = torch.softmax(x16.float(), -1)
    _store(y16, v)
    if y32 is not None:
        y32.copy_(v)


def softmax_bwd(dy16, y16, dx16):
    y, dy = y16.float(), dy16.float()
    _store(dx16, y * (dy - (dy * y).sum(-1, keepdim=True)))


def bce_fwd_bwd(x32, target, n_total, loss_out, dz=None, from_logits=True):
    dz16 = dz
    x = x32.double()
    if from_logits:
        loss = torch.clamp(x, min=0) - x * target + torch.log1p(torch.exp(-x.abs()))
        p = torch.sigmoid(x)
    else:
        p = x
        q = x.clamp(1e-7, 1 - 1e-7)
        loss = -(target * torch.log(q) + (1 - target) * torch.log(1 - q))
    loss_out += (loss.sum() / n_total).float()
    if dz16 is not None:
        _store(dz16, ((p - target) / n_total).float())


def mse_fwd_bwd(pred16, n_total, loss_out, *, target, dpred=None):
    dpred16 = dpred
    d = pred16.double() - target.double()
    cols = pred16.shape[1]
    loss_out += ((d * d).sum() / (n_total * cols)).float()
    if dpred16 is not None:
        _store(dpred16, (2 * d / (n_total * cols)).float())


def round_half_even(x16, out16=None, out32=None):
    v = torch.round(x16.float())
    _store(out16, v)
    if out32 is not None:
        out32.copy_(v)


def argmax_onehot(p32, out16=None, out32=None):
    v = torch.nn.functional.one_hot(torch.argmax(p32, -1), p32.shape[1]).float()
    _store(out16, v)
    if out32 is not None:
        out32.copy_(v)


def rmsprop_step(p32, p16, g, ms, mom, lr, rho, momentum, eps, grad_scale=1.0):
    gg = g * grad_scale
    ms.mul_(rho).add_((1 - rho) * gg * gg)
    mom.mul_(momentum).add_(lr * gg / torch.sqrt(ms + eps))
    p32.sub_(mom)
    if p16 is not None:
        p16.copy_(p32.to(p16.dtype))
    _launches[0] += 1


def fill_f32(t, value):
    t.fill_(value)
    _launches[0] += 1
