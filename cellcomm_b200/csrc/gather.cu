// CSR row gather -> dense padded minibatch (bf16 for the GEMMs, optional exact fp32).
//
// Replaces the dense DataFrame row gather of DataFrame.sample(B)
// (src/cell_type_training.py:37-38) and the DataFrame -> float32 feed of
// train_on_batch / predict (src/bigan_classify.py:144-155, src/bigan_basic.py:29-30).
// The reference applies NO normalisation: raw counts go in.
//
// One CTA per output row.  bf16 path: the row is assembled in shared memory (zero fill,
// scatter the row's non-zeros) and streamed out with 16-byte coalesced stores, so HBM sees
// one sequential write per row plus the CSR reads.  Rows too wide for shared memory, and the
// fp32 output, use zero-fill + scatter in global memory (same CTA, ordered by a barrier).
#include "common.cuh"

namespace cc {

constexpr int GATHER_THREADS = 256;

__global__ void __launch_bounds__(GATHER_THREADS)
gather_rows_kernel(const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                   const float* __restrict__ values, const long long* __restrict__ row_idx,
                   long long row_start, long long n_rows, long long n_cols,
                   bf16* __restrict__ out16, long long ld16, float* __restrict__ out32,
                   long long ld32, int use_smem) {
  extern __shared__ __align__(16) unsigned char smem[];
  bf16* srow = reinterpret_cast<bf16*>(smem);
  const int tid = threadIdx.x;
  for (long long i = blockIdx.x; i < n_rows; i += gridDim.x) {
    const long long r = row_idx ? row_idx[i] : row_start + i;
    const long long beg = rowptr[r], end = rowptr[r + 1];
    if (out16 != nullptr) {
      bf16* orow = out16 + i * ld16;
      if (use_smem) {
        // rows are 16-byte aligned on this path (host wrapper).  Only columns [0, n_cols) are
        // written, so `out16` may be a column slice of a wider buffer.
        const long long nvec = n_cols / 8, nz = (n_cols + 7) / 8;
        uint4* s4 = reinterpret_cast<uint4*>(srow);
        for (long long v = tid; v < nz; v += GATHER_THREADS) s4[v] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        for (long long e = beg + tid; e < end; e += GATHER_THREADS)
          srow[colidx[e]] = f2bf(values[e]);
        __syncthreads();
        uint4* o4 = reinterpret_cast<uint4*>(orow);
        for (long long v = tid; v < nvec; v += GATHER_THREADS) o4[v] = s4[v];
        for (long long c = nvec * 8 + tid; c < n_cols; c += GATHER_THREADS) orow[c] = srow[c];
        __syncthreads();
      } else {
        for (long long c = tid; c < n_cols; c += GATHER_THREADS) orow[c] = f2bf(0.f);
        __syncthreads();
        for (long long e = beg + tid; e < end; e += GATHER_THREADS)
          orow[colidx[e]] = f2bf(values[e]);
        __syncthreads();
      }
    }
    if (out32 != nullptr) {
      float* orow = out32 + i * ld32;
      for (long long c = tid; c < n_cols; c += GATHER_THREADS) orow[c] = 0.f;
      __syncthreads();
      for (long long e = beg + tid; e < end; e += GATHER_THREADS) orow[colidx[e]] = values[e];
      __syncthreads();
    }
  }
}

}  // namespace cc

extern "C" int cc_gather_rows(const int64_t* rowptr_dev, const int32_t* colidx_dev,
                              const float* values_dev, const int64_t* row_idx_dev,
                              int64_t row_start, int64_t n_rows, int64_t n_cols, void* out16,
                              int64_t ld16, float* out32, int64_t ld32, cc_stream_t stream) {
  using namespace cc;
  if (n_rows <= 0) return 0;
  CC_REQUIRE(out16 == nullptr || ld16 >= n_cols, "cc_gather_rows: ld16 %lld < cols %lld",
             (long long)ld16, (long long)n_cols);
  CC_REQUIRE(out32 == nullptr || ld32 >= n_cols, "cc_gather_rows: ld32 %lld < cols %lld",
             (long long)ld32, (long long)n_cols);
  static int sms = 0, max_smem = 0;
  if (sms == 0) {
    int dev = 0;
    CC_CHECK_CUDA(cudaGetDevice(&dev));
    CC_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CC_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CC_CHECK_CUDA(cudaFuncSetAttribute(gather_rows_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  }
  size_t smem = 0;
  int use_smem = 0;
  if (out16 != nullptr && ld16 % 8 == 0 && (((uintptr_t)out16) & 15) == 0 &&
      (size_t)(n_cols + 7) / 8 * 16 <= (size_t)max_smem) {
    use_smem = 1;
    smem = (size_t)(n_cols + 7) / 8 * 16;
  }
  // rows per CTA: grid-stride; cap the grid at a few waves so tiny rows do not over-launch
  long long grid = n_rows;
  const long long cap = (long long)sms * 16;
  if (grid > cap) grid = cap;
  gather_rows_kernel<<<(unsigned)grid, GATHER_THREADS, smem, (cudaStream_t)stream>>>(
      (const long long*)rowptr_dev, colidx_dev, values_dev, (const long long*)row_idx_dev,
      row_start, n_rows, n_cols, (bf16*)out16, ld16, out32, ld32, use_smem);
  CC_CHECK_LAUNCH();
  return 0;
}
