"""cc_gather_rows (csrc/gather.cu) against `dense[idx]` — the dense DataFrame row gather of
`DataFrame.sample(B)` (src/cell_type_training.py:37-38).  Indexing work: BIT-EXACT.

Covered: bf16 and fp32 outputs (together and alone), the shared-memory path and the global
zero-fill + scatter fallback (taken for a 16-byte-misaligned output and for rows wider than
shared memory), `row_idx` (sampled batch) and `row_start` (encode tile) addressing, empty
rows, duplicate indices, ragged widths, and the BASELINE width G = 33,694 with counts beyond
bf16's exact integer range.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _matrix(N, G, seed, density=0.06, empty_rows=()):
    rng = np.random.default_rng(seed)
    dense = (rng.random((N, G)) < density) * (rng.geometric(0.45, (N, G)))
    big = rng.random((N, G)) < 0.01
    dense = dense * np.where(big, 50, 1)
    for r in empty_rows:
        dense[r] = 0
    return dense.astype(np.float64)


def _csr(dense):
    from cellcomm_b200.cell_type_training import CellMatrix
    m = CellMatrix.from_dense(dense)
    return m, m.device_csr("cuda")


def _expect16(rows):
    return torch.from_numpy(rows.astype(np.float32)).to(torch.bfloat16)


@pytest.mark.parametrize("N,G", [(64, 5), (97, 333), (300, 1200), (50, 4097)])
@pytest.mark.parametrize("mode", ["idx", "start"])
def test_bf16_and_fp32_outputs_bit_exact(N, G, mode):
    from cellcomm_b200 import ops
    dense = _matrix(N, G, seed=N + G, empty_rows=(0, N // 2))
    _, csr = _csr(dense)
    if mode == "idx":
        idx = np.random.RandomState(1).permutation(N)[:max(1, N // 2)]
        idx[-1] = idx[0]                       # a duplicate row index is legal
        kw = {"row_idx": torch.from_numpy(idx).cuda()}
        want = dense[idx]
    else:
        s, n = N // 4, N // 2
        kw = {"row_start": s, "n_rows": n}
        want = dense[s:s + n]
    B = want.shape[0]
    out16 = ops.alloc2d(B, G)
    out32 = ops.alloc2d(B, G, dtype=torch.float32)
    out16.fill_(7.0)                           # stale contents must be overwritten, padding too
    out32.fill_(7.0)
    ops.gather_rows(*csr, G, out16=out16, out32=out32, **kw)
    torch.cuda.synchronize()
    assert torch.equal(out32.cpu(), torch.from_numpy(want.astype(np.float32)))
    assert torch.equal(out16.cpu(), _expect16(want))
    # the padding columns of the row (ld > G) are zero-filled: TMA reads them as K padding
    full16 = out16.as_strided((B, ops.pad_ld(G)), (ops.pad_ld(G), 1))
    assert float(full16[:, G:].abs().sum()) == 0.0
    # each output alone
    o16 = ops.alloc2d(B, G)
    ops.gather_rows(*csr, G, out16=o16, **kw)
    o32 = ops.alloc2d(B, G, dtype=torch.float32)
    ops.gather_rows(*csr, G, out32=o32, **kw)
    assert torch.equal(o16, out16) and torch.equal(o32, out32)


def test_fallback_path_misaligned_output():
    """out16 not 16-byte aligned -> zero-fill + scatter in global memory (no smem staging)."""
    from cellcomm_b200 import ops
    N, G = 80, 700
    dense = _matrix(N, G, seed=3)
    _, csr = _csr(dense)
    idx = np.random.RandomState(2).permutation(N)[:33]
    ld = ops.pad_ld(G) + 64
    buf = torch.full((33, ld), 3.0, dtype=torch.bfloat16, device="cuda")
    view = buf[:, 1:1 + G]                     # base pointer 2 bytes off a 16-byte boundary
    assert view.data_ptr() % 16 != 0
    ops.gather_rows(*csr, G, row_idx=torch.from_numpy(idx).cuda(), out16=view)
    torch.cuda.synchronize()
    assert torch.equal(view.cpu(), _expect16(dense[idx]))
    assert float((buf[:, 0].float() - 3.0).abs().sum()) == 0.0   # nothing written before the view


def test_fallback_path_rows_wider_than_shared_memory():
    from cellcomm_b200 import ops
    N, G = 6, 120_000                          # 240 KB per bf16 row > 227 KB of shared memory
    rng = np.random.default_rng(5)
    dense = np.zeros((N, G))
    for r in range(N):
        cols = rng.choice(G, 500, replace=False)
        dense[r, cols] = rng.integers(1, 300, 500)
    _, csr = _csr(dense)
    out16 = ops.alloc2d(N, G)
    ops.gather_rows(*csr, G, row_start=0, n_rows=N, out16=out16)
    torch.cuda.synchronize()
    assert torch.equal(out16.cpu(), _expect16(dense))


@pytest.mark.parametrize("B", [128, 2048])
def test_baseline_width_33694(B):
    """BASELINE.json configs[1] width; counts > 256 round to bf16 exactly like a host cast."""
    from cellcomm_b200 import ops
    from cellcomm_b200.cell_type_training import CellMatrix
    N, G = 3000, 33694
    rng = np.random.default_rng(20260101)
    nnz_row = np.clip(np.round(rng.lognormal(np.log(2000), 0.35, N)), 200, 8000).astype(np.int64)
    rowptr = np.zeros(N + 1, np.int64)
    np.cumsum(nnz_row, out=rowptr[1:])
    colidx = np.concatenate([np.sort(rng.choice(G, k, replace=False)) for k in nnz_row])
    vals = rng.geometric(0.45, rowptr[-1]).astype(np.float64)
    vals[rng.random(len(vals)) < 0.01] *= 50
    m = CellMatrix(rowptr, colidx.astype(np.int32), vals, np.arange(1, N + 1), np.arange(1, G + 1))
    csr = m.device_csr("cuda")
    idx = np.random.RandomState(0).permutation(N)[:B]
    out16 = ops.alloc2d(B, G)
    out32 = ops.alloc2d(B, G, dtype=torch.float32)
    ops.gather_rows(*csr, G, row_idx=torch.from_numpy(idx).cuda(), out16=out16, out32=out32)
    torch.cuda.synchronize()
    want = m.dense_rows(idx, np.float32)
    assert torch.equal(out32.cpu(), torch.from_numpy(want))
    assert torch.equal(out16.cpu(), torch.from_numpy(want).to(torch.bfloat16))
    # row_start addressing on the same matrix = a contiguous slice
    ops.gather_rows(*csr, G, row_start=17, n_rows=B if B < N - 17 else N - 17, out32=out32)
    torch.cuda.synchronize()
    n = B if B < N - 17 else N - 17
    assert torch.equal(out32[:n].cpu(), torch.from_numpy(m.dense_rows(np.arange(17, 17 + n), np.float32)))
