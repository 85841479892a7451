// Layer-level GEMM entry points: map a Keras Dense layer's forward / input-gradient /
// weight-gradient onto cc_gemm's operand-major flags, so that the row-major activations
// X[M,K], gradients dZ[M,N] and Keras-layout kernels W[K,N] are consumed as they lie
// (no transposed copies).
//
//   forward  Y  = act(X W + b)          A = X  (K-major)   B = W    (MN-major: N contiguous)
//   dgrad    dX = dZ W^T                A = dZ (K-major)   B = W    (K-major:  rows of W are dX cols)
//   wgrad    dW = X^T dZ                A = X  (MN-major)  B = dZ   (MN-major)
//
// Reference: layers.Dense inside Model.train_on_batch / predict,
// src/bigan_classify.py:10-75,144-155 and src/bigan_cont.py:7-41.
#include "common.cuh"

extern "C" int cc_dense_fwd(int32_t M, int32_t N, int32_t nseg, const void* const* x,
                            const int64_t* ldx, const int32_t* k, const void* const* w,
                            const int64_t* ldw, const float* bias, int32_t act, void* out16,
                            int64_t ld16, float* out32, int64_t ld32, float* workspace,
                            int64_t workspace_elems, cc_stream_t stream) {
  CC_REQUIRE(nseg >= 1 && nseg <= CC_GEMM_MAX_SEG, "cc_dense_fwd: nseg=%d", nseg);
  cc_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = M;
  d.N = N;
  d.a_mn_major = 0;
  d.b_mn_major = 1;
  d.nseg = nseg;
  for (int s = 0; s < nseg; ++s) {
    d.a[s] = x[s];
    d.lda[s] = ldx[s];
    d.b[s] = w[s];
    d.ldb[s] = ldw[s];
    d.k[s] = k[s];
  }
  d.alpha = 1.f;
  d.bias = bias;
  d.act = act;
  d.out16 = out16;
  d.ld16 = ld16;
  d.out32 = out32;
  d.ld32 = ld32;
  d.workspace = workspace;
  d.workspace_elems = workspace_elems;
  return cc_gemm(&d, stream);
}

extern "C" int cc_dense_dgrad(int32_t M, int32_t K, int32_t nseg, const void* const* dz,
                              const int64_t* lddz, const int32_t* n, const void* const* w,
                              const int64_t* ldw, const void* dact_y, int64_t ld_dact,
                              int32_t dact, float alpha, void* out16, int64_t ld16, int32_t beta16,
                              float* workspace, int64_t workspace_elems, cc_stream_t stream) {
  CC_REQUIRE(nseg >= 1 && nseg <= CC_GEMM_MAX_SEG, "cc_dense_dgrad: nseg=%d", nseg);
  cc_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = M;
  d.N = K;  // output columns = the layer's input features
  d.a_mn_major = 0;
  d.b_mn_major = 0;
  d.nseg = nseg;
  for (int s = 0; s < nseg; ++s) {
    d.a[s] = dz[s];
    d.lda[s] = lddz[s];
    d.b[s] = w[s];  // W[K rows, N_s cols]: row index = output column, contiguous = reduction
    d.ldb[s] = ldw[s];
    d.k[s] = n[s];
  }
  d.alpha = alpha;
  d.dact_y = dact_y;
  d.ld_dact = ld_dact;
  d.dact = dact;
  d.out16 = out16;
  d.ld16 = ld16;
  d.beta16 = beta16;
  d.workspace = workspace;
  d.workspace_elems = workspace_elems;
  return cc_gemm(&d, stream);
}

extern "C" int cc_dense_wgrad(int32_t M, int32_t K, int32_t N, const void* x, int64_t ldx,
                              const void* dz, int64_t lddz, float* dw, int64_t lddw, int32_t beta,
                              cc_stream_t stream) {
  cc_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = K;  // output rows = input features
  d.N = N;
  d.a_mn_major = 1;  // X stored [batch, K]: the MMA-M index (feature) is contiguous
  d.b_mn_major = 1;  // dZ stored [batch, N]
  d.nseg = 1;
  d.a[0] = x;
  d.lda[0] = ldx;
  d.b[0] = dz;
  d.ldb[0] = lddz;
  d.k[0] = M;  // reduction over the batch
  d.alpha = 1.f;
  d.out32 = dw;
  d.ld32 = lddw;
  d.beta32 = beta;
  return cc_gemm(&d, stream);
}
