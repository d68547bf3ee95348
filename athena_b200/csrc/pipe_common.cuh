// Pieces shared by the fused gather kernels (pipe_tc.cu: list gather in shared memory;
// pipe_tcg.cu: tensor-core gather): launch arguments, activation helpers and the epilogue
// of one 128-row accumulator tile.
#pragma once
#include "athena_internal.h"
#include "tc_common.cuh"

namespace athena {
namespace pipe {

using namespace tc;

template <int ACT>
__device__ __forceinline__ float act_fwd(float x) {
  if (ACT == ATHENA_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == ATHENA_ACT_LEAKY_RELU) return fmaxf(x * 0.01f, x);
  if (ACT == ATHENA_ACT_SIGMOID) return 1.f / (1.f + expf(-x));
  if (ACT == ATHENA_ACT_TANH) return tanhf(x);
  return x;
}
template <int ACT>
__device__ __forceinline__ float act_bwd(float y, float g) {
  if (ACT == ATHENA_ACT_RELU) return y > 0.f ? g : 0.f;
  if (ACT == ATHENA_ACT_LEAKY_RELU) return y > 0.f ? g : g * 0.01f;
  if (ACT == ATHENA_ACT_SIGMOID) return g * (y * (1.f - y));
  if (ACT == ATHENA_ACT_TANH) return g * (1.f - y * y);
  return g;
}

constexpr int EPI_ACT = 0;      // out = act(v)
constexpr int EPI_ACTGRAD = 1;  // out = v * act'(Hin)
constexpr int EPI_MSE = 2;      // out = d MSE / d pre-activation; the loss is reduced per CTA

struct GatherArgs {
  const int4* tiles;
  int num_tiles;
  const int32_t* row_ptr;  // CSR row pointers or CSC column pointers
  const uint8_t* col8;     // neighbour index relative to the tile's first row (Batch::col8/csc8)
  const float* rs;         // deg^-1/2 per vertex; nullptr -> unit coefficients
  const float* X;          // [V][F]
  const float* W;
  float* P;                // optional [V][F]
  float* out;              // [V][N]
  const float* aux;        // EPI_ACTGRAD: [V][N] saved activations (nullptr: act' == 1);
                           // EPI_MSE: [V][N] target
  int act;
  // relu / leaky_relu: the sign of the pre-activation as one bit per element, [V][N/32]
  // words.  The forward writes it (mask_out), the reverse sweep of the consuming step reads
  // it (mask_in) instead of the saved activations: 8 bytes per row instead of 256.
  uint32_t* mask_out;
  const uint32_t* mask_in;
  const uint4* abits;     // pipe_tcg: adjacency rows as bit masks (Batch::abits / atbits)
  long long num_rows;     // pipe_tcg: V (extent of the TMA tensor map of `aux`)
  int32_t* nonfinite;     // pipe_tcg forward: set to 1 when a feature value is NaN / Inf (such a
                          // value reaches every row of its tile through 0 * Inf in the dense
                          // adjacency product; the caller re-runs with the list gather)
  int aux_tma;            // pipe_tcg fwd+MSE: the target tile arrives by TMA tensor copies
  int store_tma;          // pipe_tcg: bit 0 / 1: `out` / `P` leave through TMA tensor stores
  int reverse;            // pipe_tcg: walk the tiles from the last to the first
  // L2 policies of the bulk copies (0 default, 1 evict_first, 2 evict_last):
  int hint_x, hint_aux, hint_out, hint_p;
  // EPI_MSE (mse_loss_type%compute for graph outputs, athena_loss.f90:416-427)
  const int32_t* vcount;  // [V] vertices of the vertex's graph (Batch::vcount)
  float* loss_part;       // [gridDim.x] sum over this CTA's rows of (p-e)^2 / (N * nv_s)
  int dbg;                // ATHENA_DEBUG_PIPE bitmask (experiments only): 1 no P store,
                          // 2 no output store, 4 no gather loop
  long long* trace;       // ATHENA_DEBUG_TRACE: [role 0..3][tile][8] clock64 stamps of CTA 0
};

// Epilogue of one 128-row tile for the warp that owns TMEM lanes 32q .. 32q+31.
// tcgen05.ld hands every thread one ROW of the accumulator.  All arithmetic happens in
// that layout; the second operand (saved activations / target) is this thread's row of a
// padded shared-memory tile (pitch 272 B: conflict-free row-per-thread LDS.128) that the
// thread itself prefetches with a 256-byte TMA bulk copy one tile ahead.  A row-per-thread
// global store would touch 32 different 128-byte lines per instruction, so the warp
// transposes the result through a private padded patch, 32 columns at a time:
// row-per-thread STS.128 (pitch 144 B), then LDS.128 / STG.128 with eight lanes per row
// segment, i.e. four full 128-byte lines per instruction.
//   EPI_ACT      out = act(v)
//   EPI_ACTGRAD  out = v .* act'(aux),  aux = saved activations (USE_AUX = false: out = v)
//   EPI_MSE      out = act'(p) .* (p - aux) * row_scale,  p = act(v), aux = target;
//                returns this row's sum (p - aux)^2 * row_scale
//                (the caller halves the total: athena_loss.f90:414-427)
constexpr int EPI_PITCH = 36;                 // floats per staged row (32 + 4 pad)
constexpr int EPI_PATCH = 32 * EPI_PITCH;
constexpr int AUX_PITCH = 68;                 // floats per operand row (64 + 4 pad)
// STACKED: the accumulator holds hi | lo partial products side by side (N columns each).
// [HB, HE): the 32-column halves this call handles (pipe_tcg.cu gives each half its own warp).
// AUX_SWZ: `aux_row` is the BASE of an operand tile that TMA wrote as two [128 rows x 32
// floats] boxes with the 128-byte swizzle (16-byte chunk c of row r at chunk c ^ (r % 8)): the
// row-per-thread reads are conflict-free without padding.
template <int ACT, int EPI, int N, bool STACKED = true, int HB = 0, int HE = N / 32,
          bool AUX_SWZ = false>
__device__ __forceinline__ float epilogue_tile(uint32_t tacc, int q, int lane, int nrows,
                                               float* __restrict__ out_tile,
                                               const float* aux_row /* shared memory */,
                                               float row_scale, float* patch,
                                               uint64_t* acc_empty, bool no_store,
                                               uint32_t* mask_out_row, bool use_mask,
                                               const uint32_t (&mask_in)[N / 32]) {
  float lsum = 0.f;
  float* srow = patch + lane * EPI_PITCH;
  auto aux4 = [&](int col) {
    if (AUX_SWZ) {
      const int r = q * 32 + lane;
      return *reinterpret_cast<const float4*>(
          reinterpret_cast<const uint8_t*>(aux_row) + (col >> 5) * (TILE_ROWS * 128) + r * 128 +
          ((((col & 31) >> 2) ^ (r & 7)) << 4));
    }
    return *reinterpret_cast<const float4*>(aux_row + col);
  };
#pragma unroll
  for (int half = HB; half < HE; ++half) {
    uint32_t mbits = 0;
#pragma unroll
    for (int cg = 0; cg < 2; ++cg) {
      float vh[16], vl[16];
      const int col0 = half * 32 + cg * 16;
      const uint32_t taddr = tacc + (static_cast<uint32_t>(q * 32) << 16) + col0;
      tmem_ld16(taddr, vh);
      if (STACKED) tmem_ld16(taddr + N, vl);
      if (half == HE - 1 && cg == 1) {
        tc_fence_before();
        mbar_arrive(acc_empty);  // the accumulator buffer may be overwritten now
      }
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = STACKED ? vh[i + k] + vl[i + k] : vh[i + k];
        if (EPI == EPI_ACT) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (ACT == ATHENA_ACT_RELU || ACT == ATHENA_ACT_LEAKY_RELU)
              mbits |= (o[k] > 0.f ? 1u : 0u) << (cg * 16 + i + k);
            o[k] = act_fwd<ACT>(o[k]);
          }
        } else if (EPI == EPI_ACTGRAD && use_mask) {
          // act'(H) from the sign bits: relu 1 / 0, leaky_relu 1 / 0.01
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool pos = (mask_in[half] >> (cg * 16 + i + k)) & 1u;
            o[k] = pos ? o[k] : (ACT == ATHENA_ACT_LEAKY_RELU ? o[k] * 0.01f : 0.f);
          }
        } else if (EPI == EPI_MSE) {
          const float4 t4 = aux4(col0 + i);
          const float t[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float pk = act_fwd<ACT>(o[k]);
            const float d = pk - t[k];
            if (row_scale != 0.f) lsum += d * d * row_scale;
            o[k] = act_bwd<ACT>(pk, d * row_scale);
          }
        } else if (ACT != ATHENA_ACT_NONE) {
          const float4 h4 = aux4(col0 + i);
          const float h[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = act_bwd<ACT>(h[k], o[k]);
        }
        *reinterpret_cast<float4*>(srow + cg * 16 + i) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    if (EPI == EPI_ACT && mask_out_row != nullptr &&
        (ACT == ATHENA_ACT_RELU || ACT == ATHENA_ACT_LEAKY_RELU))
      mask_out_row[half] = mbits;
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = it * 32 + lane;
      const int r = idx >> 3, c = idx & 7;
      const int trow = q * 32 + r;
      if (trow < nrows && !no_store)
        *reinterpret_cast<float4*>(out_tile + static_cast<size_t>(trow) * N + half * 32 + c * 4) =
            *reinterpret_cast<const float4*>(patch + r * EPI_PITCH + c * 4);
    }
    __syncwarp();
  }
  return lsum;
}

}  // namespace pipe

// pipe_tcg.cu: the same three fused steps with the gather on the tensor core
bool pipe_tcg_supported(Batch* b, int F, int N);
int launch_pipe_tcg(const pipe::GatherArgs& a, bool transb, int epi, int width = 64);

}  // namespace athena
