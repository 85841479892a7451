// Microbenchmark: does the RUN LENGTH of the access pattern bound the fused RMSprop epilogue?
// Streams three fp32 state arrays (read + write) and a bf16 copy (write) over a [K, ld] matrix
// the way a tile epilogue does: each warp instruction group touches RUN bytes of ROWS rows that
// are ld*4 bytes apart; tiles are 128 rows x 256 columns, visited like the persistent kernel
// visits them (one CTA per SM, N-fast raster).  RUN = 128 is today's epilogue, RUN = 1024 is a
// whole tile row per warp.  Usage: rms_pattern K N
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int RUN_FLOATS>   // contiguous floats per row per warp step (32 = 128 B ... 256 = 1 KB)
__global__ void __launch_bounds__(256, 1)
pattern_kernel(float* __restrict__ w, float* __restrict__ s, float* __restrict__ m,
               __nv_bfloat16* __restrict__ h, long long ld, int K, int N, int tiles_m, int tiles_n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int LANES_PER_ROW = RUN_FLOATS / 4;          // lanes covering one row's run (float4 each)
  constexpr int ROWS_PER_INSTR = 32 / LANES_PER_ROW;     // rows one warp instruction touches
  const int q = warp & 3, half = warp >> 2;              // 8 warps: 4 row quarters x 2 column halves
  for (int t = blockIdx.x; t < tiles_m * tiles_n; t += gridDim.x) {
    const int m0 = (t / tiles_n) * 128 + q * 32, n0 = (t % tiles_n) * 256 + half * 128;
    // the warp's region: 32 rows x 128 columns, walked in column blocks of RUN_FLOATS
    for (int c = 0; c < 128; c += (RUN_FLOATS < 128 ? RUN_FLOATS : 128)) {
      float4 a[8], b[8], d[8];
      constexpr int COLS = RUN_FLOATS < 128 ? RUN_FLOATS : 128;
      constexpr int LPR = COLS / 4, RPI = 32 / LPR, NI = 32 / RPI;   // instructions per block
      const int col = n0 + c + (lane % LPR) * 4;
      for (int rb = 0; rb < 32; rb += 8 * RPI) {   // 8 instructions of loads in flight at a time
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = m0 + rb + i * RPI + lane / LPR;
        if (row < K && col + 3 < N) {
          const long long off = (long long)row * ld + col;
          a[i] = __ldcs(reinterpret_cast<const float4*>(w + off));
          b[i] = __ldcs(reinterpret_cast<const float4*>(s + off));
          d[i] = __ldcs(reinterpret_cast<const float4*>(m + off));
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = m0 + rb + i * RPI + lane / LPR;
        if (row < K && col + 3 < N) {
          const long long off = (long long)row * ld + col;
          float4 ss = make_float4(0.85f * b[i].x + 1e-9f, 0.85f * b[i].y + 1e-9f, 0.85f * b[i].z + 1e-9f, 0.85f * b[i].w + 1e-9f);
          float4 mm = make_float4(0.1f * d[i].x + ss.x, 0.1f * d[i].y + ss.y, 0.1f * d[i].z + ss.z, 0.1f * d[i].w + ss.w);
          float4 ww = make_float4(a[i].x - mm.x, a[i].y - mm.y, a[i].z - mm.z, a[i].w - mm.w);
          __stcs(reinterpret_cast<float4*>(s + off), ss);
          __stcs(reinterpret_cast<float4*>(m + off), mm);
          __stcs(reinterpret_cast<float4*>(w + off), ww);
          __nv_bfloat162 lo = __floats2bfloat162_rn(ww.x, ww.y), hi = __floats2bfloat162_rn(ww.z, ww.w);
          uint2 u;
          u.x = *reinterpret_cast<unsigned*>(&lo);
          u.y = *reinterpret_cast<unsigned*>(&hi);
          __stcs(reinterpret_cast<uint2*>(h + off), u);
        }
      }
      }
      (void)NI;
    }
  }
  (void)ROWS_PER_INSTR;
}

template <int RUN>
static float run(float* w, float* s, float* m, __nv_bfloat16* h, long long ld, int K, int N) {
  const int tm = (K + 127) / 128, tn = (N + 255) / 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) pattern_kernel<RUN><<<148, 256>>>(w, s, m, h, ld, K, N, tm, tn);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) pattern_kernel<RUN><<<148, 256>>>(w, s, m, h, ld, K, N, tm, tn);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 6738, N = argc > 2 ? atoi(argv[2]) : 33694;
  const long long ld = (N + 63) / 64 * 64, n = (long long)K * ld;
  float *w, *s, *m;
  __nv_bfloat16* h;
  cudaMalloc(&w, n * 4);
  cudaMalloc(&s, n * 4);
  cudaMalloc(&m, n * 4);
  cudaMalloc(&h, n * 2);
  cudaMemset(w, 0, n * 4);
  cudaMemset(s, 0, n * 4);
  cudaMemset(m, 0, n * 4);
  const double bytes = 26.0 * K * N;
  float t;
  t = run<32>(w, s, m, h, ld, K, N);
  printf("{\"K\": %d, \"N\": %d, \"run_bytes\": 128, \"ms\": %.4f, \"GB/s\": %.0f}\n", K, N, t, bytes / t / 1e6);
  t = run<64>(w, s, m, h, ld, K, N);
  printf("{\"K\": %d, \"N\": %d, \"run_bytes\": 256, \"ms\": %.4f, \"GB/s\": %.0f}\n", K, N, t, bytes / t / 1e6);
  t = run<128>(w, s, m, h, ld, K, N);
  printf("{\"K\": %d, \"N\": %d, \"run_bytes\": 512, \"ms\": %.4f, \"GB/s\": %.0f}\n", K, N, t, bytes / t / 1e6);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
