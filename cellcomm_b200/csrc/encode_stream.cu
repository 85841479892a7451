// cc_encode_stream: the encode-all-cells pass as ONE call.
//
// Replaces `encoder.predict(all cells)` = BasicBiGan.encoding_prediction(trainer.data)
// (src/bigan_basic.py:29-30, called from src/intercepts/db_recorder.py:85): for every tile of
// `tile_rows` cells -- gather the CSR rows into a dense bf16 tile, run the encoder's layers in
// inference mode (Dense GEMMs with fused bias / activation, BatchNormalization on the moving
// statistics, softmax), write the fp32 encodings of the tile -- all enqueued on one stream
// from one C call, no host work between tiles.
//
// The encoder is described by a small program over numbered scratch slots (cc_encode_plan,
// include/cellcomm_b200.h): slot 0 is the gathered cell tile, slot -1 the output.  The caller
// (cellcomm_b200/engine.py: BiGanEngine.encode_stream) derives the program from the layer graph
// and the precision policy, so this pass runs exactly the kernels of Net.forward(bn_train =
// False, dropout = "off") and gives bit-identical encodings.
#include "common.cuh"

namespace cc {

int gemm_impl(const cc_gemm_desc* d, cudaStream_t st);

static inline int64_t pad64(int64_t n) { return n <= 64 ? 64 : (n + 63) / 64 * 64; }

struct Slot {
  char* ptr;
  int64_t ld;   // elements
  int fp32;
  int width;
};

}  // namespace cc

extern "C" int64_t cc_encode_scratch_bytes(const cc_encode_plan* plan) {
  if (plan == nullptr) return -1;
  int64_t total = 0;
  for (int s = 0; s < plan->n_slots; ++s) {
    const int64_t bytes = (int64_t)plan->tile_rows * cc::pad64(plan->slot_width[s]) *
                          (plan->slot_fp32[s] ? 4 : 2);
    total += (bytes + 255) / 256 * 256;
  }
  return total;
}

extern "C" int cc_encode_stream(const int64_t* rowptr_dev, const int32_t* colidx_dev,
                                const float* values_dev, int64_t row_begin, int64_t row_end,
                                const cc_encode_plan* plan, float* out32, int64_t ld_out,
                                cc_stream_t stream) {
  using namespace cc;
  CC_REQUIRE(plan != nullptr && out32 != nullptr, "cc_encode_stream: null plan / output");
  CC_REQUIRE(row_end >= row_begin, "cc_encode_stream: rows [%lld, %lld)", (long long)row_begin,
             (long long)row_end);
  CC_REQUIRE(plan->n_slots >= 1 && plan->n_slots <= CC_ENC_MAX_SLOTS && plan->n_ops >= 1 &&
                 plan->n_ops <= CC_ENC_MAX_OPS && plan->tile_rows >= 1,
             "cc_encode_stream: bad plan (%d slots, %d ops, tile %d)", plan->n_slots, plan->n_ops,
             plan->tile_rows);
  CC_REQUIRE(plan->slot_fp32[0] == 0 && plan->slot_width[0] == plan->n_cols,
             "cc_encode_stream: slot 0 must be the bf16 cell tile");
  CC_REQUIRE(plan->scratch != nullptr && (((uintptr_t)plan->scratch) & 255) == 0 &&
                 plan->scratch_bytes >= cc_encode_scratch_bytes(plan),
             "cc_encode_stream: scratch of %lld bytes needed (256-byte aligned)",
             (long long)cc_encode_scratch_bytes(plan));
  Slot slots[CC_ENC_MAX_SLOTS];
  {
    char* p = (char*)plan->scratch;
    for (int s = 0; s < plan->n_slots; ++s) {
      slots[s].ptr = p;
      slots[s].ld = pad64(plan->slot_width[s]);
      slots[s].fp32 = plan->slot_fp32[s];
      slots[s].width = plan->slot_width[s];
      const int64_t bytes = (int64_t)plan->tile_rows * slots[s].ld * (slots[s].fp32 ? 4 : 2);
      p += (bytes + 255) / 256 * 256;
    }
  }
  auto check_slot = [&](int s, bool allow_out) {
    return (s >= 0 && s < plan->n_slots) || (allow_out && s == -1);
  };

  for (int64_t r0 = row_begin; r0 < row_end; r0 += plan->tile_rows) {
    const int64_t rows = (row_end - r0 < plan->tile_rows) ? row_end - r0 : plan->tile_rows;
    float* out_tile = out32 + (r0 - row_begin) * ld_out;
    // the tile's padding columns were zeroed once by the caller (scratch is zero-initialised)
    int rc = cc_gather_rows(rowptr_dev, colidx_dev, values_dev, nullptr, r0, rows, plan->n_cols,
                            slots[0].ptr, slots[0].ld, nullptr, 0, stream);
    if (rc) return rc;
    for (int i = 0; i < plan->n_ops; ++i) {
      const cc_enc_op& op = plan->ops[i];
      CC_REQUIRE(check_slot(op.out, true), "cc_encode_stream: op %d writes slot %d", i, op.out);
      // destination: a scratch slot, or (slot -1) the fp32 output tile
      const bool to_out = op.out == -1;
      char* dptr = to_out ? (char*)out_tile : slots[op.out].ptr;
      const int64_t dld = to_out ? ld_out : slots[op.out].ld;
      const int dfp32 = to_out ? 1 : slots[op.out].fp32;
      switch (op.kind) {
        case CC_ENC_DENSE: {
          CC_REQUIRE(op.n_in >= 1 && op.n_in <= CC_GEMM_MAX_SEG, "cc_encode_stream: op %d has %d "
                     "segments", i, op.n_in);
          cc_gemm_desc d;
          memset(&d, 0, sizeof(d));
          d.M = (int32_t)rows;
          d.N = op.width;
          d.a_mn_major = 0;
          d.b_mn_major = 1;
          d.nseg = op.n_in;
          for (int s = 0; s < op.n_in; ++s) {
            CC_REQUIRE(check_slot(op.in[s], false) && !slots[op.in[s]].fp32,
                       "cc_encode_stream: op %d segment %d reads slot %d (must be bf16)", i, s,
                       op.in[s]);
            d.a[s] = slots[op.in[s]].ptr;
            d.lda[s] = slots[op.in[s]].ld;
            d.b[s] = (const char*)op.w16 + (int64_t)op.w_row[s] * op.ldw * 2;
            d.ldb[s] = op.ldw;
            d.k[s] = slots[op.in[s]].width;
          }
          d.alpha = 1.f;
          d.bias = op.bias;
          d.act = op.act;
          if (dfp32) {
            d.out32 = (float*)dptr;
            d.ld32 = dld;
          } else {
            d.out16 = dptr;
            d.ld16 = dld;
          }
          d.workspace = plan->workspace;
          d.workspace_elems = plan->workspace_elems;
          rc = gemm_impl(&d, (cudaStream_t)stream);
          break;
        }
        case CC_ENC_SPLIT: {   // fp32 slot -> bf16 hi (op.out) + lo (op.out2)
          CC_REQUIRE(check_slot(op.in[0], false) && slots[op.in[0]].fp32 && !to_out &&
                         check_slot(op.out2, false),
                     "cc_encode_stream: op %d: bad split", i);
          const Slot& x = slots[op.in[0]];
          rc = cc_split_bf16(x.ptr, x.ld, dptr, dld, slots[op.out2].ptr, slots[op.out2].ld, rows,
                             x.width, 1, stream);
          break;
        }
        case CC_ENC_COPY: {    // cast / copy between slots (or to the output tile)
          CC_REQUIRE(check_slot(op.in[0], false), "cc_encode_stream: op %d: bad copy", i);
          const Slot& x = slots[op.in[0]];
          rc = cc_copy2d(x.ptr, x.ld, dptr, dld, rows, x.width, 0, 1.f,
                         (x.fp32 ? 1 : 0) | (dfp32 ? 2 : 0), stream);
          break;
        }
        case CC_ENC_BN_INFER: {
          CC_REQUIRE(check_slot(op.in[0], false), "cc_encode_stream: op %d: bad bn", i);
          const Slot& x = slots[op.in[0]];
          rc = cc_bn_infer(x.ptr, x.ld, dptr, dld, rows, x.width, op.gamma, op.beta, op.mean,
                           op.var, op.eps, (x.fp32 ? 1 : 0) | (dfp32 ? 2 : 0), stream);
          break;
        }
        case CC_ENC_SOFTMAX: {  // op.out: activation slot; fp32 copy to the output tile
          CC_REQUIRE(check_slot(op.in[0], false) && !to_out, "cc_encode_stream: op %d: bad "
                     "softmax", i);
          const Slot& x = slots[op.in[0]];
          rc = cc_softmax_fwd(x.ptr, x.ld, dptr, dld, out_tile, ld_out, rows, x.width,
                              (x.fp32 ? 1 : 0) | (dfp32 ? 2 : 0), stream);
          break;
        }
        default:
          CC_REQUIRE(false, "cc_encode_stream: op %d has unknown kind %d", i, op.kind);
      }
      if (rc) return rc;
    }
  }
  return 0;
}
