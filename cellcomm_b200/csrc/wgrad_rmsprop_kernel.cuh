// Weight-gradient GEMM with the Keras RMSprop(momentum) update applied in the epilogue, the
// optimiser state streamed by TMA (included by gemm_sm100.cu).
//
//   dW[K, N] = X[batch, K]^T dZ[batch, N]             (both operands MN-major, fp32 in TMEM)
//   ms  = rho*ms + (1-rho) dW^2 ;  mom = momentum*mom + lr*dW/sqrt(ms+eps) ;  w -= mom ;
//   w16 = bf16(w)                                      (optimizers.RMSprop, src/bigan_classify.py:88)
//
// The launch is HBM-bound: 26 B per parameter (read w, ms, mom; write w, ms, mom, w16) against
// 2*batch FLOPs.  The first version of this epilogue moved the state through registers with
// per-row address arithmetic (~1,500 warp instructions per 32x32 block, 57 % of the DRAM
// peak).  Here every byte of state is moved by the TMA engine:
//
//   warp 0      A / B producer (2-stage ring; a tile's operands come from L2)
//   warp 1      tcgen05.mma issuer, accumulator double-buffered in TMEM (2 x 256 columns)
//   warps 2..9  epilogue.  Warp (q = warp % 4, h) owns rows 32q..32q+31 and columns
//               128h..128h+127 of the tile, in four 32 x 32 blocks.  Per block: one lane issues
//               three 2-D TMA loads (w, ms, mom: 32 rows x 128 B each, 128-byte swizzle) into
//               the warp's private staging buffer; lane r reads the accumulator row r from TMEM
//               (tcgen05.ld 32x32b: no transpose needed -- a lane owns a row in TMEM and a row
//               in the swizzled staging buffer alike, bank-conflict free), updates the 32
//               elements in place in shared memory, and one lane issues four TMA stores
//               (w, ms, mom, bf16 w).  Edges need no code: TMA zero-fills loads and clips
//               stores at the tensor bounds, so padding is never written.
//
// All state traffic carries an L2 evict-first policy (it is touched once per update).
#pragma once

namespace cc {

struct RmsMaps {
  CUtensorMap a[MAX_SEG];
  CUtensorMap b[MAX_SEG];
  CUtensorMap p32, ms, mom, p16;
};

__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                                 int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, uint32_t src, int c0,
                                                  int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::
          "l"(map),
      "r"(src), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

constexpr int RMS_BN = 256;
constexpr int RMS_STAGES = 2;
constexpr int RMS_EPI_WARPS = 8;
constexpr uint32_t RMS_F32_BLOCK = 32 * 32 * 4;                   // one 32x32 fp32 block
constexpr uint32_t RMS_B16_BLOCK = 32 * 32 * 2;
constexpr uint32_t RMS_EPI_BUF = 3 * RMS_F32_BLOCK + RMS_B16_BLOCK;  // per epilogue warp
constexpr uint32_t RMS_STAGE_BYTES = BM * BK * 2 + RMS_BN * BK * 2;
constexpr size_t RMS_SMEM_BYTES = (size_t)RMS_STAGES * RMS_STAGE_BYTES +
                                  (size_t)RMS_EPI_WARPS * RMS_EPI_BUF + 8 * (2 * RMS_STAGES + 4 + RMS_EPI_WARPS) +
                                  16 + 1024;

// One warp's share of a tile: rows r0..r0+31 (lane = row = TMEM lane), columns
// n0 + [col_lo, col_hi) in 32-column blocks taken `col_step` apart (the two warps of a lane
// quarter either split the tile's columns in halves or interleave their blocks, so that the
// pair touches 256 contiguous bytes of every row at about the same time).  t_row: TMEM address
// of the warp's lane quarter at the accumulator's column 0.  buf_w: the warp's staging buffer
// (w | ms | mom | bf16 w).
__device__ __forceinline__ void rms_epilogue_tile(const EpiParams& e, const RmsMaps& maps,
                                                  const uint32_t t_row, const int r0, const int n0,
                                                  const int col_lo, const int col_hi,
                                                  const int col_step, const int lane,
                                                  const uint32_t buf_w, const uint32_t sbar,
                                                  uint32_t& sphase, const uint64_t pol,
                                                  const bool p16_direct) {
  if (r0 >= e.M) return;  // warp-uniform (the odd row tile of a pair may lie below the matrix)
  const uint32_t buf_s = buf_w + RMS_F32_BLOCK;
  const uint32_t buf_m = buf_s + RMS_F32_BLOCK;
  const uint32_t buf_h = buf_m + RMS_F32_BLOCK;
  const bool has16 = e.rms_p16 != nullptr;
  const float rho = e.rms_rho, omr = 1.f - e.rms_rho, mu = e.rms_momentum, lr = e.rms_lr,
              eps = e.rms_eps;
  // swizzled position of this lane's 16-byte pieces: 128-byte rows, piece j at (j ^ (lane & 7))
  const uint32_t row128 = (uint32_t)lane * 128u, sw128 = (uint32_t)(lane & 7);
  const uint32_t row64 = (uint32_t)lane * 64u, sw64 = (uint32_t)((lane >> 1) & 3);
#pragma unroll 1
  for (int c = col_lo; c < col_hi; c += col_step) {
    const int c0 = n0 + c;
    if (c0 >= e.N) break;  // warp-uniform
    if (elect_one()) {
      bulk_wait_read0();  // the previous block's stores have drained the staging buffer
      mbar_expect_tx(sbar, 3 * RMS_F32_BLOCK);
      tma_load_2d_hint(buf_w, &maps.p32, sbar, c0, r0, pol);
      tma_load_2d_hint(buf_s, &maps.ms, sbar, c0, r0, pol);
      tma_load_2d_hint(buf_m, &maps.mom, sbar, c0, r0, pol);
    }
    __syncwarp();
    uint32_t raw[32];
    tmem_ld32(t_row + (uint32_t)c, raw);
    tmem_ld_wait();
    if (e.out32 != nullptr && r0 + lane < e.M) {
      // parity tests also want the gradient itself (keep_grads): plain row stores
      float* o = e.out32 + (long long)(r0 + lane) * e.ld32 + c0;
      const int nc = min(32, e.N - c0);
      if (nc == 32 && (e.ld32 & 3) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(o + j) = make_uint4(raw[j], raw[j + 1], raw[j + 2], raw[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nc) o[j] = __uint_as_float(raw[j]);
      }
    }
    mbar_wait(sbar, sphase, 25);
    sphase ^= 1u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t off = row128 + ((((uint32_t)j) ^ sw128) << 4);
      const float4 w = lds128(buf_w + off), s = lds128(buf_s + off), m = lds128(buf_m + off);
      const float g0 = __uint_as_float(raw[4 * j]), g1 = __uint_as_float(raw[4 * j + 1]),
                  g2 = __uint_as_float(raw[4 * j + 2]), g3 = __uint_as_float(raw[4 * j + 3]);
      float4 ss, mm, ww;
      ss.x = fmaf(rho, s.x, omr * g0 * g0);
      ss.y = fmaf(rho, s.y, omr * g1 * g1);
      ss.z = fmaf(rho, s.z, omr * g2 * g2);
      ss.w = fmaf(rho, s.w, omr * g3 * g3);
      mm.x = fmaf(mu, m.x, lr * g0 * rsqrtf(ss.x + eps));
      mm.y = fmaf(mu, m.y, lr * g1 * rsqrtf(ss.y + eps));
      mm.z = fmaf(mu, m.z, lr * g2 * rsqrtf(ss.z + eps));
      mm.w = fmaf(mu, m.w, lr * g3 * rsqrtf(ss.w + eps));
      ww.x = w.x - mm.x;
      ww.y = w.y - mm.y;
      ww.z = w.z - mm.z;
      ww.w = w.w - mm.w;
      sts128(buf_s + off, ss);
      sts128(buf_m + off, mm);
      sts128(buf_w + off, ww);
      // bf16 copy: two float4 make one 16-byte piece; park the packed words in raw[] (the
      // gradient words 4j..4j+3 were consumed above, 2j and 2j+1 lie at or below them)
      __nv_bfloat162 lo = __floats2bfloat162_rn(ww.x, ww.y);
      __nv_bfloat162 hi = __floats2bfloat162_rn(ww.z, ww.w);
      raw[2 * j] = *reinterpret_cast<uint32_t*>(&lo);
      raw[2 * j + 1] = *reinterpret_cast<uint32_t*>(&hi);
    }
    if (has16 && !p16_direct) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        sts128u(buf_h + row64 + ((((uint32_t)t) ^ sw64) << 4),
                make_uint4(raw[4 * t], raw[4 * t + 1], raw[4 * t + 2], raw[4 * t + 3]));
    } else if (has16 && r0 + lane < e.M) {
      // each lane owns 64 contiguous bytes of its row: two full 32-byte sectors
      bf16* o = e.rms_p16 + (long long)(r0 + lane) * e.rms_ld + c0;
      if (c0 + 32 <= e.N) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
          __stcs(reinterpret_cast<uint4*>(o) + t,
                 make_uint4(raw[4 * t], raw[4 * t + 1], raw[4 * t + 2], raw[4 * t + 3]));
      } else {
        const int nc = e.N - c0;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          if (2 * t < nc) o[2 * t] = __ushort_as_bfloat16((unsigned short)(raw[t] & 0xFFFFu));
          if (2 * t + 1 < nc) o[2 * t + 1] = __ushort_as_bfloat16((unsigned short)(raw[t] >> 16));
        }
      }
    }
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA engine
    __syncwarp();
    if (elect_one()) {
      tma_store_2d_hint(&maps.ms, buf_s, c0, r0, pol);
      tma_store_2d_hint(&maps.mom, buf_m, c0, r0, pol);
      tma_store_2d_hint(&maps.p32, buf_w, c0, r0, pol);
      if (has16 && !p16_direct) tma_store_2d_hint(&maps.p16, buf_h, c0, r0, pol);
      bulk_commit();
    }
  }
}

template <bool N_FAST, int CLUSTER>
__global__ void __launch_bounds__(64 + 32 * RMS_EPI_WARPS, 1)
wgrad_rmsprop_tma_kernel(const __grid_constant__ RmsMaps maps, const __grid_constant__ GemmParams p,
                         const int tiles_m, const int tiles_n) {
  constexpr int BN = RMS_BN, STAGES = RMS_STAGES, EPI_WARPS = RMS_EPI_WARPS;
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t STAGE_BYTES = RMS_STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;          // 1024-byte aligned
  const uint32_t bar_base = epi_base + EPI_WARPS * RMS_EPI_BUF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  auto state_bar = [&](int w) { return bar_base + 8u * (2 * STAGES + 4 + w); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4 + EPI_WARPS);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkb = p.total_kblocks;
  const int rank = (CLUSTER == 2) ? (int)cluster_ctarank() : 0;
  const int tiles_mu = (tiles_m + CLUSTER - 1) / CLUSTER;
  const int num_units = tiles_mu * tiles_n;
  const int unit0 = (int)blockIdx.x / CLUSTER, unit_step = (int)gridDim.x / CLUSTER;
  auto unit_m0 = [&](int u) { return ((N_FAST ? u / tiles_n : u % tiles_mu) * CLUSTER + rank) * BM; };
  auto unit_n0 = [&](int u) { return (N_FAST ? u % tiles_n : u / tiles_mu) * BN; };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CLUSTER);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), EPI_WARPS);
    }
#pragma unroll
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(state_bar(w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CLUSTER == 2) cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_acc = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== operand producer (X^T and dZ tiles, from L2) =====================
    uint32_t it = 0;
    for (int u = unit0; u < num_units; u += unit_step) {
      const int m0 = unit_m0(u), n0 = unit_n0(u);
      int seg = 0, kb_in_seg = 0;
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u, 21);
        if (elect_one()) {
          const uint32_t a_dst = smem_base + s * STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_BYTES;
          const int k0 = kb_in_seg * BK;
          mbar_expect_tx(full_bar(s), STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < BM / 64; ++j)
            tma_load_2d(a_dst + j * (64 * BK * 2), &maps.a[seg], full_bar(s), m0 + 64 * j, k0);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) {
            if (CLUSTER == 2) {
              if ((j & 1) == rank)
                tma_load_2d_mc(b_dst + j * (64 * BK * 2), &maps.b[seg], full_bar(s), n0 + 64 * j, k0,
                               (uint16_t)3);
            } else {
              tma_load_2d(b_dst + j * (64 * BK * 2), &maps.b[seg], full_bar(s), n0 + 64 * j, k0);
            }
          }
        }
        __syncwarp();
        if (++kb_in_seg >= p.kblocks[seg] && seg < p.nseg - 1) {
          kb_in_seg = 0;
          ++seg;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t it = 0, tl = 0;
    for (int u = unit0; u < num_units; u += unit_step, ++tl) {
      const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
      mbar_wait(tempty_bar(acc), aph ^ 1u, 22);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_acc + acc * BN;
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        mbar_wait(full_bar(s), ph, 23);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a_src = smem_base + s * STAGE_BYTES;
          const uint32_t b_src = a_src + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = p.adesc_hi | (uint64_t)(((a_src + k * p.a_kstep) & 0x3FFFFu) >> 4);
            const uint64_t bdesc = p.bdesc_hi | (uint64_t)(((b_src + k * p.b_kstep) & 0x3FFFFu) >> 4);
            umma_bf16(d_tmem, adesc, bdesc, p.idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          if (CLUSTER == 2) umma_commit_mc(empty_bar(s), (uint16_t)3);
          else umma_commit(empty_bar(s));
          if (i == nkb - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: RMSprop on the accumulator, state moved by TMA ==========
    const EpiParams& e = p.epi;
    const int q = warp & 3;                 // TMEM lane quarter (hardware rule: warp id mod 4)
    const int ew = warp - 2;
    const int col_lo = (ew >> 2) * (BN / 2);
    const uint32_t buf_w = epi_base + ew * RMS_EPI_BUF;      // w, then ms, mom, bf16 w
    const uint32_t sbar = state_bar(ew);
    const uint64_t pol = l2_evict_first_policy();
    const bool interleave = (e.rms_cs & 4) != 0;   // the quarter's two warps alternate blocks
    uint32_t tl = 0, sphase = 0;
    for (int u = unit0; u < num_units; u += unit_step, ++tl) {
      const int m0 = unit_m0(u), n0 = unit_n0(u);
      const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
      mbar_wait(tfull_bar(acc), aph, 24);
      tcgen05_fence_after();
      rms_epilogue_tile(e, maps, tmem_acc + ((uint32_t)(q * 32) << 16) + acc * BN, m0 + q * 32, n0,
                        interleave ? (ew >> 2) * 32 : col_lo, interleave ? BN : col_lo + BN / 2,
                        interleave ? 64 : 32, lane, buf_w, sbar, sphase, pol, (e.rms_cs & 2) != 0);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
    __syncwarp();
    if (elect_one()) bulk_wait_read0();   // staging memory stays valid until the last store read it
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CLUSTER == 2) cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc),
                 "r"(TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair version (tcgen05 cta_group::2): the two SMs of a cluster compute ONE 256 x 256 tile.
// Each CTA stages its own 128 rows of X^T and only HALF of the dZ tile (the tensor cores read
// the other half from the peer's shared memory), so a k-block costs 32 KB of shared memory per
// CTA instead of 48 KB: four stages fit next to the eight epilogue staging buffers, against two
// in the kernel above -- at batch 2048 (32 k-blocks per tile) the operand ring of the one-CTA
// kernels is latency-bound and stretches the HBM-bound epilogue.  The leader CTA (cluster rank
// 0) issues every MMA; completion is multicast to both CTAs' barriers; each CTA's epilogue warps
// drain their own 128 accumulator rows from their own TMEM.
constexpr int RMSP_STAGES = 4;
constexpr uint32_t RMSP_STAGE_BYTES = BM * BK * 2 + (RMS_BN / 2) * BK * 2;
constexpr uint32_t RMSP_EPI_BUF = 3 * RMS_F32_BLOCK;   // the bf16 copy is stored from registers
constexpr size_t RMSP_SMEM_BYTES = (size_t)RMSP_STAGES * RMSP_STAGE_BYTES +
                                   (size_t)RMS_EPI_WARPS * RMSP_EPI_BUF +
                                   8 * (2 * RMSP_STAGES + 4 + RMS_EPI_WARPS) + 16 + 1024;

template <bool N_FAST>
__global__ void __launch_bounds__(64 + 32 * RMS_EPI_WARPS, 1)
wgrad_rmsprop_pair_kernel(const __grid_constant__ RmsMaps maps, const __grid_constant__ GemmParams p,
                          const int tiles_m, const int tiles_n) {
  constexpr int BN = RMS_BN, STAGES = RMSP_STAGES, EPI_WARPS = RMS_EPI_WARPS;
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t STAGE_BYTES = RMSP_STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = epi_base + EPI_WARPS * RMSP_EPI_BUF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };              // used in the leader
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };  // one per CTA
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };  // leader's
  auto state_bar = [&](int w) { return bar_base + 8u * (2 * STAGES + 4 + w); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4 + EPI_WARPS);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkb = p.total_kblocks;
  const int rank = (int)cluster_ctarank();
  const bool leader = rank == 0;
  const int tiles_mu = (tiles_m + 1) / 2;
  const int num_units = tiles_mu * tiles_n;
  const int unit0 = (int)blockIdx.x / 2, unit_step = (int)gridDim.x / 2;
  auto unit_m0 = [&](int u) { return ((N_FAST ? u / tiles_n : u % tiles_mu) * 2 + rank) * BM; };
  auto unit_n0 = [&](int u) { return (N_FAST ? u % tiles_n : u / tiles_mu) * BN; };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2);    // one arrival per CTA of the pair (+ the bytes of both)
      mbar_init(empty_bar(s), 1);   // the leader's commit, multicast to both CTAs
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * EPI_WARPS);   // the epilogue warps of both CTAs
    }
#pragma unroll
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(state_bar(w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {   // one warp of EACH CTA: the pair allocates the same columns in both TMEMs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_acc = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== operand producer: own X^T rows + own half of dZ ====================
    uint32_t it = 0;
    for (int u = unit0; u < num_units; u += unit_step) {
      const int m0 = unit_m0(u), n0 = unit_n0(u) + rank * (BN / 2);
      int seg = 0, kb_in_seg = 0;
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u, 31);
        if (elect_one()) {
          const uint32_t a_dst = smem_base + s * STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_BYTES;
          const uint32_t lbar = mapa_cluster(full_bar(s), 0);   // the leader's barrier
          const int k0 = kb_in_seg * BK;
          if (leader) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES);
          else mbar_arrive_cluster(lbar);
#pragma unroll
          for (int j = 0; j < BM / 64; ++j)
            tma_load_2d_pair(a_dst + j * (64 * BK * 2), &maps.a[seg], lbar, m0 + 64 * j, k0);
#pragma unroll
          for (int j = 0; j < BN / 128; ++j)
            tma_load_2d_pair(b_dst + j * (64 * BK * 2), &maps.b[seg], lbar, n0 + 64 * j, k0);
        }
        __syncwarp();
        if (++kb_in_seg >= p.kblocks[seg] && seg < p.nseg - 1) {
          kb_in_seg = 0;
          ++seg;
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA only): M = 256 across the pair ============
      const uint32_t idesc = (p.idesc & ~(0x1Fu << 24)) | ((uint32_t)((2 * BM) >> 4) << 24);
      uint32_t it = 0, tl = 0;
      for (int u = unit0; u < num_units; u += unit_step, ++tl) {
        const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
        mbar_wait(tempty_bar(acc), aph ^ 1u, 32);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_acc + acc * BN;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(full_bar(s), ph, 33);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t a_src = smem_base + s * STAGE_BYTES;
            const uint32_t b_src = a_src + A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adesc = p.adesc_hi | (uint64_t)(((a_src + k * p.a_kstep) & 0x3FFFFu) >> 4);
              const uint64_t bdesc = p.bdesc_hi | (uint64_t)(((b_src + k * p.b_kstep) & 0x3FFFFu) >> 4);
              umma_bf16_pair(d_tmem, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_pair(empty_bar(s), (uint16_t)3);   // both CTAs may refill the stage
            if (i == nkb - 1) umma_commit_pair(tfull_bar(acc), (uint16_t)3);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs: own 128 rows of the pair's tile) =============
    const EpiParams& e = p.epi;
    const int q = warp & 3;
    const int ew = warp - 2;
    const int col_lo = (ew >> 2) * (BN / 2);
    const uint32_t buf_w = epi_base + ew * RMSP_EPI_BUF;
    const uint32_t sbar = state_bar(ew);
    const uint64_t pol = l2_evict_first_policy();
    const bool interleave = (e.rms_cs & 4) != 0;
    uint32_t tl = 0, sphase = 0;
    for (int u = unit0; u < num_units; u += unit_step, ++tl) {
      const int m0 = unit_m0(u), n0 = unit_n0(u);
      const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
      mbar_wait(tfull_bar(acc), aph, 34);
      tcgen05_fence_after();
      rms_epilogue_tile(e, maps, tmem_acc + ((uint32_t)(q * 32) << 16) + acc * BN, m0 + q * 32, n0,
                        interleave ? (ew >> 2) * 32 : col_lo, interleave ? BN : col_lo + BN / 2,
                        interleave ? 64 : 32, lane, buf_w, sbar, sphase, pol, true);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_cluster(tempty_bar(acc), 0));
    }
    __syncwarp();
    if (elect_one()) bulk_wait_read0();
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc),
                 "r"(TMEM_COLS)
                 : "memory");
  }
}

}  // namespace cc
