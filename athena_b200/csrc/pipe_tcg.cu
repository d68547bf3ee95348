// Tensor-core gather: the fused Kipf step for mini-batches of small graphs with BOTH
// products on the tcgen05 pipe.
//
//   out[v,:] = epi( ( sum_{w in row v} c_w * X[col[w],:] ) . op(W) )         (as pipe_tc.cu)
//
// A tile of whole graphs (<= 128 vertices) is block diagonal, so its propagate step is a
// dense [128 x 128] x [128 x 64] product with a 0/1 adjacency matrix:
//     P = D^-1/2 . ( A . ( D^-1/2 X ) )          (athena_diffstruc_extd_sub_kipf.f90:29-46)
// 0 and 1 are exact in tf32, so the adjacency needs no hi/lo split; it is expanded from 16
// bytes per row (Batch::abits) straight into TENSOR MEMORY and used as the A operand of
// tcgen05.mma (A from TMEM, B = hi/lo split feature tile in shared memory).  The
// accumulator P already has the layout of a TMEM A operand (lane = vertex, one column per
// feature), so the weight transform out = P . W reads it in place: the propagated tile never
// passes through shared memory.  Compared with the list gather of pipe_tc.cu (~7 random
// 256-byte shared-memory row reads per output row + hi/lo operand stores: ~6300 shared
// memory wavefronts per tile) a tile costs ~2900 wavefronts, which moves the kernel from
// the shared-memory bandwidth wall to the HBM stream.
//
// Roles (one persistent CTA per SM, 18 warps):
//   producer (1 lane)   TMA bulk copies: feature rows + deg^-1/2 into a shared-memory ring
//   split warps (4)     ring -> registers (scaled by deg_u^-1/2), hi/lo split, MN-major
//                       swizzled B operand (the layout k_pipe_tn uses)
//   build warps (4)     adjacency bits -> 1.0f / 0.0f -> tcgen05.st into TMEM (A operand)
//   MMA (1 lane)        G(j):  P  = ADJ . [Xhi ; Xlo]            32 x (128 x 64 x 8)
//                       T(j):  O  = P . Whi + P . Wlo + Plo . Whi 24 x (128 x 64 x 8)
//                       issued G(j+1) before T(j) so the pipe never waits for the fix warps
//   fix warps (4)       tcgen05.ld P, * deg_v^-1/2, hi -> P, lo -> Plo (tcgen05.st), and the
//                       coalesced global store of P (saved for dW) through a padded patch
//   epilogue warps (4)  the epilogue of pipe_tc.cu (activation / act' / fused MSE)
// TMEM (512 columns): ADJ 0..127 | P0,Plo0 128..255 | P1,Plo1 256..383 | O0 384..447 | O1 448..511
//
// Summation order: the tensor core adds the row's terms in column order with fp32
// (truncating) accumulation, not in the reference's entry order; with <= 128 terms the
// difference is ~1e-7 relative, inside the 1e-5 parity tolerance.
// Batches in which some (row, column) pair repeats (multi-edges) keep the list kernels.
#include <algorithm>
#include <cstdio>

#include "athena_internal.h"
#include "pipe_common.cuh"
#include "tc_common.cuh"

namespace athena {

using namespace tc;
using namespace pipe;

namespace {

// mbarrier wait that traps (instead of hanging the GPU) when a barrier never completes
__device__ __forceinline__ void mbar_wait_g(uint64_t* bar, uint32_t parity, int tag) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
    if (done) return;
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    if (now - t0 > 4000000000ll) {
      printf("k_pipe_tcg: barrier %d never completed (block %d thread %d parity %u)\n", tag,
             blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

template <int EPI>
struct TcgCfg {
  static constexpr int F = 64, N = 64;
  static constexpr int NS = EPI == EPI_MSE ? 1 : 2;             // ring stages (shared memory budget)
  static constexpr int SPLIT_WARP0 = 0;                          // warps 0..3
  static constexpr int BUILD_WARP0 = 4;                          // warps 4..7   (warp % 4 = TMEM lane quarter)
  static constexpr int FIX_WARP0 = 8;                            // warps 8..11
  static constexpr int EPI_WARP0 = 12;                           // warps 12..15
  static constexpr int PRODUCER_WARP = 16;
  static constexpr int MMA_WARP = 17;
  static constexpr int THREADS = 18 * 32;
  static constexpr int X_BYTES = TILE_ROWS * F * 4;              // 32 KB of raw feature rows
  static constexpr int RS_BYTES = (TILE_ROWS + 8) * 4;
  static constexpr int STAGE_BYTES = (X_BYTES + RS_BYTES + 127) / 128 * 128;
  static constexpr int OP_BLK = TILE_ROWS * 128;                 // [128 vertices x 32 features]
  static constexpr int OP_BYTES = (F / 32) * OP_BLK;             // hi (or lo) feature operand
  static constexpr int W_BLK = 2 * N * 128;                      // [hi(W') ; lo(W')] x 32 k
  static constexpr int OFF_BHI = 0;
  static constexpr int OFF_BLO = OP_BYTES;
  static constexpr int OFF_W = 2 * OP_BYTES;
  static constexpr int OFF_RING = OFF_W + (F / 32) * W_BLK;
  static constexpr int OFF_AUX = OFF_RING + NS * STAGE_BYTES;
  static constexpr int AUX_BYTES = EPI != EPI_ACT ? TILE_ROWS * AUX_PITCH * 4 : 0;
  static constexpr int OFF_BAR = OFF_AUX + AUX_BYTES;
  static constexpr int OFF_EPI = OFF_BAR + 256 + 512;
  static constexpr int OFF_FIX = OFF_EPI + 4 * EPI_PATCH * 4;
  static constexpr int FIX_BYTES = EPI != EPI_ACTGRAD ? 4 * EPI_PATCH * 4 : 0;  // P is stored forward only
  static constexpr int SMEM = 1024 + OFF_FIX + FIX_BYTES;
  static_assert(SMEM <= 232448, "shared memory budget");
  static constexpr uint32_t T_ADJ = 0, T_P = 128, T_O = 384;     // TMEM columns
};

template <bool TRANSB, int EPI>
__global__ void __launch_bounds__(TcgCfg<EPI>::THREADS, 1) k_pipe_tcg(GatherArgs a) {
  using Cfg = TcgCfg<EPI>;
  constexpr int F = Cfg::F, N = Cfg::N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sBhi = smem + Cfg::OFF_BHI;
  uint8_t* sBlo = smem + Cfg::OFF_BLO;
  uint8_t* sW = smem + Cfg::OFF_W;
  uint8_t* ring = smem + Cfg::OFF_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;                  // [NS]  TMA landed
  uint64_t* empty = bars + 2;             // [NS]  split warps hold the stage in registers
  uint64_t* ops_ready = bars + 4;         // feature operand (hi/lo) of tile j staged
  uint64_t* adj_ready = bars + 5;         // adjacency of tile j in TMEM
  uint64_t* g_done = bars + 6;            // G(j) finished: operand buffers + ADJ reusable
  uint64_t* p_full = bars + 7;            // [2] P accumulator of tile j complete
  uint64_t* p_fixed = bars + 9;           // [2] scaled hi/lo P back in TMEM
  uint64_t* o_full = bars + 11;           // [2]
  uint64_t* o_empty = bars + 13;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
  float* loss_red = reinterpret_cast<float*>(smem + Cfg::OFF_BAR + 256);  // [128], EPI_MSE
  float* sAux = reinterpret_cast<float*>(smem + Cfg::OFF_AUX);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int step = gridDim.x;

  if (warp == Cfg::MMA_WARP) tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    for (int s = 0; s < Cfg::NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 128);
    }
    mbar_init(ops_ready, 128);
    mbar_init(adj_ready, 128);
    mbar_init(g_done, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&p_full[b], 1);
      mbar_init(&p_fixed[b], 128);
      mbar_init(&o_full[b], 1);
      mbar_init(&o_empty[b], 128);
    }
    mbar_fence_init();
  }
  // stacked weight operand [hi(W') ; lo(W')], W' = op(W) as [N][F] K-major (as k_pipe_gather)
  for (int item = tid; item < N * (F / 4); item += Cfg::THREADS) {
    const int n = item / (F / 4), kc = item - n * (F / 4);
    float4 w;
    if (!TRANSB) {  // W row-major [F][N]: W'[n][k] = W[k][n]
      const float* src = a.W + (kc * 4) * N + n;
      w = make_float4(__ldg(src), __ldg(src + N), __ldg(src + 2 * N), __ldg(src + 3 * N));
    } else {        // W row-major [N][F] used as is
      w = __ldg(reinterpret_cast<const float4*>(a.W + n * F + kc * 4));
    }
    float4 hi, lo;
    split_tf32(w, hi, lo);
    uint8_t* blk = sW + (kc >> 3) * Cfg::W_BLK;
    *reinterpret_cast<float4*>(blk + sw128_off(n, kc & 7)) = hi;
    *reinterpret_cast<float4*>(blk + sw128_off(N + n, kc & 7)) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool has_coef = a.rs != nullptr;

  if (warp == Cfg::PRODUCER_WARP) {
    // ===================== producer: TMA bulk copies into the ring =====================
    if (lane == 0) {
      int j = 0;
      for (int t = blockIdx.x; t < a.num_tiles; t += step, ++j) {
        const int s = j % Cfg::NS;
        const uint32_t ph = (j / Cfg::NS) & 1;
        mbar_wait_g(&empty[s], ph ^ 1u, 0);
        const int4 ti = __ldg(a.tiles + t);
        const int r0 = ti.x, nrows = ti.y;
        const int ra = r0 & ~3, rcnt = (r0 + nrows - ra + 3) & ~3;
        uint8_t* st = ring + s * Cfg::STAGE_BYTES;
        const uint32_t xb = nrows * F * 4, rb = rcnt * 4;
        mbar_arrive_expect_tx(&full[s], xb + (has_coef ? rb : 0u));
        bulk_g2s(st, a.X + static_cast<size_t>(r0) * F, xb, &full[s]);
        if (has_coef) bulk_g2s(st + Cfg::X_BYTES, a.rs + ra, rb, &full[s]);
      }
    }
  } else if (warp == Cfg::MMA_WARP) {
    // ===================== MMA issuer ==================================================
    if (lane == 0) {
      const uint32_t bHi = smem_u32(sBhi), bLo = smem_u32(sBlo), wAddr = smem_u32(sW);
      constexpr uint32_t IDESC_G = make_idesc(128, F, false, true);   // B = features, MN-major
      constexpr uint32_t IDESC_T = make_idesc(128, N, false, false);  // B = W', K-major
      int my_tiles = 0;
      for (int t = blockIdx.x; t < a.num_tiles; t += step) ++my_tiles;
      for (int j = 0; j <= my_tiles; ++j) {
        if (j < my_tiles) {
          // ---- G(j): P[j&1] = ADJ . (Xhi + Xlo)
          mbar_wait_g(ops_ready, j & 1, 1);
          mbar_wait_g(adj_ready, j & 1, 2);
          tc_fence_after();
          const uint32_t tp = tmem + Cfg::T_P + (j & 1) * 128;
          // The tensor core adds into its fp32 accumulator with truncation: the small (lo)
          // terms go first, while the accumulator is small, so that only the hi terms (one
          // per neighbour block) contribute a full-size truncation step.
#pragma unroll
          for (int ks = 0; ks < TILE_ROWS / 8; ++ks) {
            const uint64_t dl = make_desc_mn32(bLo + ks * 1024, Cfg::OP_BLK, 512);
            umma_tf32_ts(tp, tmem + Cfg::T_ADJ + ks * 8, dl, IDESC_G, ks ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < TILE_ROWS / 8; ++ks) {
            const uint64_t dh = make_desc_mn32(bHi + ks * 1024, Cfg::OP_BLK, 512);
            umma_tf32_ts(tp, tmem + Cfg::T_ADJ + ks * 8, dh, IDESC_G, 1u);
          }
          umma_commit(g_done);
          umma_commit(&p_full[j & 1]);
        }
        if (j >= 1) {
          // ---- T(j-1): O = P . Whi + P . Wlo + Plo . Whi
          const int jj = j - 1, b = jj & 1;
          mbar_wait_g(&p_fixed[b], (jj >> 1) & 1, 3);
          mbar_wait_g(&o_empty[b], ((jj >> 1) & 1) ^ 1u, 4);
          tc_fence_after();
          const uint32_t tp = tmem + Cfg::T_P + b * 128, to = tmem + Cfg::T_O + b * N;
          // cross terms first (see G): P . Wlo, Plo . Whi, then P . Whi
#pragma unroll
          for (int k8 = 0; k8 < F / 8; ++k8) {
            const uint32_t wk = wAddr + (k8 >> 2) * Cfg::W_BLK + (k8 & 3) * 32;
            const uint64_t dwh = make_desc(wk, 16, 1024);
            const uint64_t dwl = make_desc(wk + N * 128, 16, 1024);
            umma_tf32_ts(to, tp + k8 * 8, dwl, IDESC_T, k8 ? 1u : 0u);
            umma_tf32_ts(to, tp + 64 + k8 * 8, dwh, IDESC_T, 1u);
          }
#pragma unroll
          for (int k8 = 0; k8 < F / 8; ++k8) {
            const uint32_t wk = wAddr + (k8 >> 2) * Cfg::W_BLK + (k8 & 3) * 32;
            umma_tf32_ts(to, tp + k8 * 8, make_desc(wk, 16, 1024), IDESC_T, 1u);
          }
          umma_commit(&o_full[b]);
        }
      }
    }
  } else if (warp >= Cfg::EPI_WARP0) {
    // ===================== epilogue: TMEM -> registers -> global rows ==================
    const int q = warp - Cfg::EPI_WARP0;
    float* patch = reinterpret_cast<float*>(smem + Cfg::OFF_EPI) + q * EPI_PATCH;
    const bool use_mask = (EPI == EPI_ACTGRAD) && a.mask_in != nullptr;
    const bool use_aux = (EPI != EPI_ACT) && a.aux != nullptr && !use_mask;
    const int my_row = q * 32 + lane;
    float* aux_row = sAux + my_row * AUX_PITCH;
    auto tile_at = [&](int t) { return t < a.num_tiles ? __ldg(a.tiles + t) : make_int4(0, 0, 0, 0); };
    auto count_at = [&](int t) {
      int c = 1;
      if (EPI == EPI_MSE && t < a.num_tiles) {
        const int4 ti = __ldg(a.tiles + t);
        if (my_row < ti.y) c = __ldg(a.vcount + ti.x + my_row);
      }
      return c;
    };
    // second operand (saved activations / target): each warp prefetches its own 32 rows of
    // the next tile with cp.async into rows padded to 272 B (see pipe_tc.cu)
    auto issue_aux = [&](const int4& ti) {
      const float* src = a.aux + (static_cast<size_t>(ti.x) + q * 32) * N;
      float* dst = sAux + q * 32 * AUX_PITCH;
      const int rows = min(32, ti.y - q * 32);
#pragma unroll
      for (int it = 0; it < 32 * (N / 4) / 32; ++it) {
        const int idx = it * 32 + lane;
        const int r = idx / (N / 4), c = idx - r * (N / 4);
        if (r < rows)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                           smem_u32(dst + r * AUX_PITCH + c * 4)),
                       "l"(src + r * N + c * 4)
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (use_aux && static_cast<int>(blockIdx.x) < a.num_tiles) issue_aux(tile_at(blockIdx.x));
    int count_next = count_at(blockIdx.x);
    int j = 0;
    float lsum = 0.f;
    for (int t = blockIdx.x; t < a.num_tiles; t += step, ++j) {
      const int b = j & 1;
      const int4 ti = __ldg(a.tiles + t);
      const int count = count_next;
      count_next = count_at(t + step);
      float* out_tile = a.out + static_cast<size_t>(ti.x) * N;
      const bool row_valid = my_row < ti.y;
      uint32_t min_w[N / 32] = {};
      uint32_t* mout = nullptr;
      if (row_valid) {
        const size_t grow = static_cast<size_t>(ti.x) + my_row;
        if (use_mask) {
#pragma unroll
          for (int w = 0; w < N / 32; ++w) min_w[w] = __ldg(a.mask_in + grow * (N / 32) + w);
        }
        if (EPI == EPI_ACT && a.mask_out != nullptr) mout = a.mask_out + grow * (N / 32);
      }
      mbar_wait_g(&o_full[b], (j >> 1) & 1, 5);
      tc_fence_after();
      if (use_aux) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
      }
      const uint32_t tacc = tmem + Cfg::T_O + b * N;
      const float scale =
          (EPI == EPI_MSE && row_valid) ? 1.f / static_cast<float>(N * count) : 0.f;
      const int act = use_aux || use_mask || EPI == EPI_ACT ? a.act : ATHENA_ACT_NONE;
      switch (act) {
        case ATHENA_ACT_RELU:
          lsum += epilogue_tile<ATHENA_ACT_RELU, EPI, N, false>(
              tacc, q, lane, ti.y, out_tile, aux_row, scale, patch, &o_empty[b], false, mout,
              use_mask, min_w);
          break;
        case ATHENA_ACT_LEAKY_RELU:
          lsum += epilogue_tile<ATHENA_ACT_LEAKY_RELU, EPI, N, false>(
              tacc, q, lane, ti.y, out_tile, aux_row, scale, patch, &o_empty[b], false, mout,
              use_mask, min_w);
          break;
        case ATHENA_ACT_SIGMOID:
          lsum += epilogue_tile<ATHENA_ACT_SIGMOID, EPI, N, false>(
              tacc, q, lane, ti.y, out_tile, aux_row, scale, patch, &o_empty[b], false, mout,
              use_mask, min_w);
          break;
        case ATHENA_ACT_TANH:
          lsum += epilogue_tile<ATHENA_ACT_TANH, EPI, N, false>(
              tacc, q, lane, ti.y, out_tile, aux_row, scale, patch, &o_empty[b], false, mout,
              use_mask, min_w);
          break;
        default:
          lsum += epilogue_tile<ATHENA_ACT_NONE, EPI, N, false>(
              tacc, q, lane, ti.y, out_tile, aux_row, scale, patch, &o_empty[b], false, mout,
              use_mask, min_w);
          break;
      }
      if (use_aux && t + step < a.num_tiles) {
        __syncwarp();
        issue_aux(tile_at(t + step));
      }
    }
    if (EPI == EPI_MSE) loss_red[my_row] = lsum;
  } else if (warp >= Cfg::FIX_WARP0) {
    // ===================== fix warps: P * deg_v^-1/2 -> hi / lo back into TMEM, P -> global ====
    const int q = warp - Cfg::FIX_WARP0;
    const int my_row = q * 32 + lane;
    float* patch = reinterpret_cast<float*>(smem + Cfg::OFF_FIX) + q * EPI_PATCH;
    auto rs_at = [&](int t) {
      float w = 1.f;
      if (has_coef && t < a.num_tiles) {
        const int4 ti = __ldg(a.tiles + t);
        if (my_row < ti.y) w = __ldg(a.rs + ti.x + my_row);
      }
      return w;
    };
    float w_next = rs_at(blockIdx.x);
    int j = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += step, ++j) {
      const int b = j & 1;
      const int4 ti = __ldg(a.tiles + t);
      const float wv = w_next;
      w_next = rs_at(t + step);
      mbar_wait_g(&p_full[b], (j >> 1) & 1, 6);
      tc_fence_after();
      const uint32_t tp = tmem + Cfg::T_P + b * 128 + (static_cast<uint32_t>(q * 32) << 16);
      const bool store_p = EPI != EPI_ACTGRAD && a.P != nullptr;
      float* p_tile = a.P + static_cast<size_t>(ti.x) * F;
      float* srow = patch + lane * EPI_PITCH;
#pragma unroll
      for (int h = 0; h < F / 32; ++h) {
        float v[32];
        tmem_ld32(tp + h * 32, v);
        uint32_t hi[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] *= wv;
          hi[i] = __float_as_uint(v[i]) & 0xffffe000u;
        }
        tmem_st32(tp + h * 32, hi);
#pragma unroll
        for (int i = 0; i < 32; ++i) hi[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));
        tmem_st32(tp + 64 + h * 32, hi);
        if (store_p) {
          // coalesced store of the propagated tile (the operand of dW = P^T gY)
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(srow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int idx = it * 32 + lane;
            const int r = idx >> 3, c = idx & 7;
            const int trow = q * 32 + r;
            if (trow < ti.y)
              *reinterpret_cast<float4*>(p_tile + static_cast<size_t>(trow) * F + h * 32 + c * 4) =
                  *reinterpret_cast<const float4*>(patch + r * EPI_PITCH + c * 4);
          }
          __syncwarp();
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_fixed[b]);
    }
  } else if (warp >= Cfg::BUILD_WARP0) {
    // ===================== build warps: adjacency bits -> TMEM A operand ================
    const int q = warp - Cfg::BUILD_WARP0;
    const int my_row = q * 32 + lane;
    auto bits_at = [&](int t) {
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      if (t < a.num_tiles) {
        const int4 ti = __ldg(a.tiles + t);
        if (my_row < ti.y) w = __ldg(a.abits + ti.x + my_row);
      }
      return w;
    };
    uint4 b_next = bits_at(blockIdx.x);
    int j = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += step, ++j) {
      const uint4 bits = b_next;
      b_next = bits_at(t + step);
      mbar_wait_g(g_done, (j & 1) ^ 1u, 7);  // G(j-1) has read the previous adjacency
      tc_fence_after();
      const uint32_t ta = tmem + Cfg::T_ADJ + (static_cast<uint32_t>(q * 32) << 16);
      const uint32_t words[4] = {bits.x, bits.y, bits.z, bits.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = ((words[c] >> k) & 1u) ? 0x3f800000u : 0u;
        tmem_st32(ta + c * 32, v);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(adj_ready);
    }
  } else {
    // ===================== split warps: ring -> scaled hi / lo B operand ================
    constexpr int LOADS = TILE_ROWS * (F / 4) / 128;  // 16-byte chunks per thread
    int j = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += step, ++j) {
      const int s = j % Cfg::NS;
      const uint32_t ph = (j / Cfg::NS) & 1;
      const int4 ti = __ldg(a.tiles + t);
      const int r0 = ti.x, nrows = ti.y;
      const uint8_t* st = ring + s * Cfg::STAGE_BYTES;
      const float* rss = reinterpret_cast<const float*>(st + Cfg::X_BYTES) + (r0 - (r0 & ~3));
      mbar_wait_g(&full[s], ph, 8);
      float4 x[LOADS];
#pragma unroll
      for (int i = 0; i < LOADS; ++i) {
        const int idx = tid + 128 * i;
        const int row = idx >> 4;
        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < nrows) {
          x[i] = *reinterpret_cast<const float4*>(st + idx * 16);
          if (has_coef) {
            const float w = rss[row];
            x[i].x *= w;
            x[i].y *= w;
            x[i].z *= w;
            x[i].w *= w;
          }
        }
      }
      mbar_arrive(&empty[s]);                // the stage lives in registers now
      mbar_wait_g(g_done, (j & 1) ^ 1u, 9);  // G(j-1) has read the operand buffers
#pragma unroll
      for (int i = 0; i < LOADS; ++i) {
        const int idx = tid + 128 * i;
        const int row = idx >> 4, ch = idx & 15;
        float4 hi, lo;
        split_tf32(x[i], hi, lo);
        const uint32_t off = (ch >> 3) * Cfg::OP_BLK + sw128b32_off(row, ch & 7);
        *reinterpret_cast<float4*>(sBhi + off) = hi;
        *reinterpret_cast<float4*>(sBlo + off) = lo;
      }
      fence_async_smem();
      mbar_arrive(ops_ready);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (EPI == EPI_MSE) {
    if (tid == 0) {
      float tot = 0.f;
      for (int i = 0; i < 128; ++i) tot += loss_red[i];
      a.loss_part[blockIdx.x] = tot;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::MMA_WARP) tmem_dealloc<512>(tmem);
}

template <bool TRANSB, int EPI>
int launch_tcg_t(const GatherArgs& a) {
  using Cfg = TcgCfg<EPI>;
  static bool attr = false;
  if (!attr) {
    ATH_CUDA(cudaFuncSetAttribute(k_pipe_tcg<TRANSB, EPI>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  const int grid = std::min(a.num_tiles, ctx().sm_count);
  k_pipe_tcg<TRANSB, EPI><<<grid, Cfg::THREADS, Cfg::SMEM, ctx().stream>>>(a);
  ATH_LAUNCHED_T(EPI == EPI_ACT ? "pipe_gather_fwd"
                 : EPI == EPI_MSE ? "pipe_gather_fwd_mse" : "pipe_gather_bwd");
  return ATHENA_OK;
}

}  // namespace

// The tensor-core gather needs a 0/1 adjacency: batches with a repeated (row, column) pair
// keep the list kernels.  The flag is produced by the batch build; it is read back once.
bool pipe_tcg_supported(Batch* b, int F, int N) {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("ATHENA_CUDA_DISABLE_TCG");
    off = (e && atoi(e) != 0) ? 1 : 0;
  }
  if (off || b->num_tiles == 0 || F != 64 || N != 64 || b->abits == nullptr) return false;
  if (b->multi_edges < 0) {
    int32_t st[4] = {0, 0, 0, 0};
    if (cudaMemcpyAsync(st, b->status.p, sizeof(st), cudaMemcpyDeviceToHost, ctx().stream) !=
            cudaSuccess ||
        cudaStreamSynchronize(ctx().stream) != cudaSuccess)
      return false;
    b->multi_edges = st[2] != 0 ? 1 : 0;
  }
  return b->multi_edges == 0;
}

int launch_pipe_tcg(const GatherArgs& a, bool transb, int epi) {
  if (epi == EPI_ACT) return launch_tcg_t<false, EPI_ACT>(a);
  if (epi == EPI_MSE) return launch_tcg_t<false, EPI_MSE>(a);
  ATH_REQUIRE(transb && epi == EPI_ACTGRAD, ATHENA_ERR_ARG, "pipe_tcg: unsupported variant");
  return launch_tcg_t<true, EPI_ACTGRAD>(a);
}

}  // namespace athena
