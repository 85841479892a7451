"""Thin torch-tensor wrappers over the C ABI (include/cellcomm_b200.h).

PyTorch is plumbing here: it owns device memory and the current CUDA stream; every function
passes raw pointers + leading dimensions to libcellcomm_b200.so.  2-D tensors may be strided
views (stride(1) == 1), so column slices of padded buffers are passed without copies.

There is deliberately NO fallback: on a tensor that is not on a CUDA device, or when the
library is missing, these raise.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmDesc, check

ACT_NONE, ACT_SIGMOID, ACT_RELU = 0, 1, 2
ACT_IDS = {None: 0, "none": 0, "linear": 0, "sigmoid": 1, "relu": 2}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"cellcomm_b200.ops: {name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"cellcomm_b200.ops: {name} must be {dtype}, got {t.dtype}")
    if t.dim() == 2 and t.shape[1] > 1 and t.stride(1) != 1:
        raise ValueError(f"cellcomm_b200.ops: {name} must be row-major (stride(1)==1)")


def _p(t):
    return None if t is None else t.data_ptr()


def _ld(t):
    if t is None:
        return 0
    if t.dim() == 1:
        return t.shape[0]
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def pad_ld(cols, mult=64):
    """Leading dimension used for activations / weights: multiple of 64 elements (128 B rows)."""
    return max(mult, (cols + mult - 1) // mult * mult)


def alloc2d(rows, cols, dtype=torch.bfloat16, device="cuda", zero=True):
    """[rows, cols] view of a zero-initialised [rows, pad_ld(cols)] buffer."""
    ld = pad_ld(cols)
    buf = (torch.zeros if zero else torch.empty)((max(rows, 1), ld), dtype=dtype, device=device)
    return buf[:rows, :cols]


_WS = {}


def workspace(device):
    """Per-device split-K scratch (fp32)."""
    key = str(device)
    ws = _WS.get(key)
    if ws is None:
        ws = torch.empty(296 * 128 * 256 * 2, dtype=torch.float32, device=device)
        _WS[key] = ws
    return ws


def launch_count():
    return int(_lib.load().cc_launch_count())


# --------------------------------------------------------------------------- GEMM
def gemm(M, N, a_list, b_list, k_list, a_mn, b_mn, *, bias=None, act=0, dact_y=None, dact=0,
         alpha=1.0, out16=None, beta16=0, out32=None, beta32=0, splits=0, bn=0, use_ws=True):
    lib = _lib.load()
    d = GemmDesc()
    d.M, d.N = int(M), int(N)
    d.a_mn_major, d.b_mn_major = int(a_mn), int(b_mn)
    d.nseg = len(a_list)
    for s, (a, b, k) in enumerate(zip(a_list, b_list, k_list)):
        _req(a, torch.bfloat16, "A")
        _req(b, torch.bfloat16, "B")
        d.a[s], d.lda[s] = a.data_ptr(), _ld(a)
        d.b[s], d.ldb[s] = b.data_ptr(), _ld(b)
        d.k[s] = int(k)
    d.alpha = float(alpha)
    _req(bias, torch.float32, "bias")
    d.bias = _p(bias)
    d.act = int(act)
    _req(dact_y, torch.bfloat16, "dact_y")
    d.dact_y, d.ld_dact, d.dact = _p(dact_y), _ld(dact_y), int(dact)
    _req(out16, torch.bfloat16, "out16")
    _req(out32, torch.float32, "out32")
    d.out16, d.ld16, d.beta16 = _p(out16), _ld(out16), int(beta16)
    d.out32, d.ld32, d.beta32 = _p(out32), _ld(out32), int(beta32)
    if use_ws:
        ws = workspace(a_list[0].device)
        d.workspace, d.workspace_elems = ws.data_ptr(), ws.numel()
    d.force_splits, d.force_bn = int(splits), int(bn)
    check(lib.cc_gemm(C.byref(d), _stream()))


def dense_fwd(xs, w16, row_offsets, bias, act, out16=None, out32=None):
    """out = act(sum_s xs[s] @ w16[roff_s : roff_s + xs[s].shape[1], :] + bias).

    xs: list of [M, K_s] bf16 (Concatenate segments, consumed without materialising);
    w16: [K_total, N] bf16 Keras-layout kernel.  Dense forward, src/bigan_classify.py:10-75."""
    M, N = xs[0].shape[0], w16.shape[1]
    bs = [w16[ro:ro + x.shape[1]] for x, ro in zip(xs, row_offsets)]
    gemm(M, N, xs, bs, [x.shape[1] for x in xs], 0, 1, bias=bias, act=act, out16=out16, out32=out32)


def dense_dgrad(dzs, ws16, out16, *, dact_y=None, dact=0, alpha=1.0, beta=0):
    """out16[M,K] (+)= alpha * (sum_s dzs[s] @ ws16[s]^T) * act'(dact_y); ws16[s]: [K, N_s]."""
    M, K = out16.shape
    gemm(M, K, dzs, ws16, [dz.shape[1] for dz in dzs], 0, 0, dact_y=dact_y, dact=dact, alpha=alpha,
         out16=out16, beta16=beta)


def dense_wgrad(x, dz, dw32, beta=0):
    """dw32[K,N] (+)= x[M,K]^T @ dz[M,N]."""
    K, N = dw32.shape
    gemm(K, N, [x], [dz], [x.shape[0]], 1, 1, out32=dw32, beta32=beta, use_ws=False)


# --------------------------------------------------------------------------- data
def gather_rows(rowptr, colidx, values, n_cols, *, row_idx=None, row_start=0, n_rows=None,
                out16=None, out32=None):
    lib = _lib.load()
    _req(rowptr, torch.int64, "rowptr")
    _req(colidx, torch.int32, "colidx")
    _req(values, torch.float32, "values")
    _req(row_idx, torch.int64, "row_idx")
    _req(out16, torch.bfloat16, "out16")
    _req(out32, torch.float32, "out32")
    if n_rows is None:
        n_rows = row_idx.shape[0]
    check(lib.cc_gather_rows(rowptr.data_ptr(), colidx.data_ptr(), values.data_ptr(), _p(row_idx),
                             int(row_start), int(n_rows), int(n_cols), _p(out16), _ld(out16),
                             _p(out32), _ld(out32), _stream()))


# --------------------------------------------------------------------------- tail kernels
def colsum(x16, out32, beta=0):
    _req(x16, torch.bfloat16, "x")
    _req(out32, torch.float32, "out")
    check(_lib.load().cc_colsum(_p(x16), _ld(x16), x16.shape[0], x16.shape[1], _p(out32), int(beta),
                                _stream()))


def dropout(x16, out16, rate, *, mask=None, seed=0, counter=None, stream_id=0):
    _req(x16, torch.bfloat16, "x")
    _req(out16, torch.bfloat16, "out")
    _req(mask, torch.uint8, "mask")
    _req(counter, torch.int64, "counter")
    check(_lib.load().cc_dropout(_p(x16), _ld(x16), _p(out16), _ld(out16), x16.shape[0],
                                 x16.shape[1], float(rate), _p(mask), _ld(mask), int(seed),
                                 _p(counter), int(stream_id), _stream()))


def dropout_mask(mask, rate, *, seed=0, counter=None, stream_id=0):
    _req(mask, torch.uint8, "mask")
    check(_lib.load().cc_dropout_mask(_p(mask), _ld(mask), mask.shape[0], mask.shape[1],
                                      float(rate), int(seed), _p(counter), int(stream_id),
                                      _stream()))


def uniform(out32=None, out16=None, *, seed=0, counter=None, stream_id=0):
    t = out32 if out32 is not None else out16
    _req(out32, torch.float32, "out32")
    _req(out16, torch.bfloat16, "out16")
    if out32 is not None and out16 is not None and _ld(out32) != _ld(out16):
        raise ValueError("uniform: out32/out16 must share a leading dimension")
    check(_lib.load().cc_uniform(_p(out32), _p(out16), _ld(t), t.shape[0], t.shape[1], int(seed),
                                 _p(counter), int(stream_id), _stream()))


def counter_add(counter, inc=1):
    check(_lib.load().cc_counter_add(_p(counter), int(inc), _stream()))


def act_bwd(dy16, y16, dz16, act):
    check(_lib.load().cc_act_bwd(_p(dy16), _ld(dy16), _p(y16), _ld(y16), _p(dz16), _ld(dz16),
                                 dy16.shape[0], dy16.shape[1], int(act), _stream()))


def copy2d(src16, dst16, beta=0):
    _req(src16, torch.bfloat16, "src")
    _req(dst16, torch.bfloat16, "dst")
    check(_lib.load().cc_copy2d(_p(src16), _ld(src16), _p(dst16), _ld(dst16), src16.shape[0],
                                src16.shape[1], int(beta), _stream()))


def cast_f32_to_bf16(src32, dst16):
    _req(src32, torch.float32, "src")
    _req(dst16, torch.bfloat16, "dst")
    check(_lib.load().cc_cast_f32_to_bf16(_p(src32), _ld(src32), _p(dst16), _ld(dst16),
                                          src32.shape[0], src32.shape[1], _stream()))


def cast_bf16_to_f32(src16, dst32, scale=1.0):
    _req(src16, torch.bfloat16, "src")
    _req(dst32, torch.float32, "dst")
    check(_lib.load().cc_cast_bf16_to_f32(_p(src16), _ld(src16), _p(dst32), _ld(dst32),
                                          src16.shape[0], src16.shape[1], float(scale), _stream()))


def bn_stats(x16, sums):
    check(_lib.load().cc_bn_stats(_p(x16), _ld(x16), x16.shape[0], x16.shape[1], _p(sums),
                                  _stream()))


def bn_train_apply(x16, y16, sums, n_total, gamma, beta, eps, momentum, moving_mean, moving_var,
                   save_mean, save_rstd):
    check(_lib.load().cc_bn_train_apply(_p(x16), _ld(x16), _p(y16), _ld(y16), x16.shape[0],
                                        x16.shape[1], _p(sums), int(n_total), _p(gamma), _p(beta),
                                        float(eps), float(momentum), _p(moving_mean),
                                        _p(moving_var), _p(save_mean), _p(save_rstd), _stream()))


def bn_infer(x16, y16, gamma, beta, moving_mean, moving_var, eps):
    check(_lib.load().cc_bn_infer(_p(x16), _ld(x16), _p(y16), _ld(y16), x16.shape[0], x16.shape[1],
                                  _p(gamma), _p(beta), _p(moving_mean), _p(moving_var), float(eps),
                                  _stream()))


def bn_bwd_stats(dy16, x16, save_mean, save_rstd, sums2):
    check(_lib.load().cc_bn_bwd_stats(_p(dy16), _ld(dy16), _p(x16), _ld(x16), dy16.shape[0],
                                      dy16.shape[1], _p(save_mean), _p(save_rstd), _p(sums2),
                                      _stream()))


def bn_bwd_apply(dy16, x16, dx16, gamma, save_mean, save_rstd, sums2, n_total, dgamma=None,
                 dbeta=None):
    check(_lib.load().cc_bn_bwd_apply(_p(dy16), _ld(dy16), _p(x16), _ld(x16), _p(dx16), _ld(dx16),
                                      dy16.shape[0], dy16.shape[1], _p(gamma), _p(save_mean),
                                      _p(save_rstd), _p(sums2), int(n_total), _p(dgamma),
                                      _p(dbeta), _stream()))


def bn_infer_bwd(dy16, dx16, gamma, moving_var, eps):
    check(_lib.load().cc_bn_infer_bwd(_p(dy16), _ld(dy16), _p(dx16), _ld(dx16), dy16.shape[0],
                                      dy16.shape[1], _p(gamma), _p(moving_var), float(eps),
                                      _stream()))


def softmax_fwd(x16, y16=None, y32=None):
    check(_lib.load().cc_softmax_fwd(_p(x16), _ld(x16), _p(y16), _ld(y16), _p(y32), _ld(y32),
                                     x16.shape[0], x16.shape[1], _stream()))


def softmax_bwd(dy16, y16, dx16):
    check(_lib.load().cc_softmax_bwd(_p(dy16), _ld(dy16), _p(y16), _ld(y16), _p(dx16), _ld(dx16),
                                     dy16.shape[0], dy16.shape[1], _stream()))


def bce_fwd_bwd(x32, target, n_total, loss_out, dz16=None, from_logits=True):
    _req(x32, torch.float32, "x")
    _req(loss_out, torch.float32, "loss_out")
    check(_lib.load().cc_bce_fwd_bwd(_p(x32), _ld(x32), x32.shape[0], int(bool(from_logits)),
                                     float(target), int(n_total), _p(loss_out), _p(dz16),
                                     _ld(dz16), _stream()))


def mse_fwd_bwd(pred16, n_total, loss_out, *, target16=None, target32=None, dpred16=None):
    check(_lib.load().cc_mse_fwd_bwd(_p(pred16), _ld(pred16), _p(target16), _ld(target16),
                                     _p(target32), _ld(target32), pred16.shape[0], pred16.shape[1],
                                     int(n_total), _p(loss_out), _p(dpred16), _ld(dpred16),
                                     _stream()))


def round_half_even(x16, out16=None, out32=None):
    check(_lib.load().cc_round_half_even(_p(x16), _ld(x16), _p(out16), _ld(out16), _p(out32),
                                         _ld(out32), x16.shape[0], x16.shape[1], _stream()))


def argmax_onehot(p32, out16=None, out32=None):
    check(_lib.load().cc_argmax_onehot(_p(p32), _ld(p32), _p(out16), _ld(out16), _p(out32),
                                       _ld(out32), p32.shape[0], p32.shape[1], _stream()))


def rmsprop_step(p32, p16, g, ms, mom, lr, rho, momentum, eps, grad_scale=1.0):
    """In-place Keras RMSprop(momentum) on a [rows, cols] view; all fp32 views share p32's ld."""
    rows, cols = (p32.shape[0], p32.shape[1]) if p32.dim() == 2 else (1, p32.shape[0])
    ld = _ld(p32) if p32.dim() == 2 else cols
    for t in (g, ms, mom):
        if (t.dim() == 2 and _ld(t) != ld) or t.shape != p32.shape:
            raise ValueError("rmsprop_step: gradient/slots must share the parameter's layout")
    if p16 is not None and p16.dim() == 2 and _ld(p16) != ld:
        raise ValueError("rmsprop_step: p16 must share the parameter's leading dimension")
    check(_lib.load().cc_rmsprop_step(_p(p32), _p(p16), _p(g), _p(ms), _p(mom), rows, cols, ld,
                                      float(lr), float(rho), float(momentum), float(eps),
                                      float(grad_scale), _stream()))


def bias_act(bias, act, rows, out16=None, out32=None):
    t = out16 if out16 is not None else out32
    check(_lib.load().cc_bias_act(_p(bias), int(act), _p(out16), _ld(out16), _p(out32), _ld(out32),
                                  int(rows), t.shape[1], _stream()))


def fill_f32(t, value):
    check(_lib.load().cc_fill_f32(_p(t), float(value), t.numel(), _stream()))
