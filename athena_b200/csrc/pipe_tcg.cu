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
// Roles (one persistent CTA per SM, 24 warps):
//   producer (1 lane)   TMA bulk copies: feature rows + deg^-1/2 into a shared-memory ring
//   split warps (8)     ring -> registers (scaled by deg_u^-1/2), hi/lo split, MN-major
//                       swizzled B operand [Xhi | Xlo] (the layout k_pipe_tn uses)
//   build warps (4)     adjacency bits -> 1.0f / 0.0f -> tcgen05.st into TMEM (A operand)
//   MMA warp G          G(j):  [P(hi part) | P(lo part)] = ADJ . [Xhi | Xlo]   16 x (128 x 128 x 8)
//   fix warps (4)       tcgen05.ld P, add the halves, * deg_v^-1/2, hi -> P, lo -> Plo
//                       (tcgen05.st, in place), and the coalesced global store of P (saved
//                       for dW) through a swizzled patch
//   MMA warp T          T(j):  O = Plo . Whi + P . Wlo + P . Whi                24 x (128 x 64 x 8)
//                       G and T are issued by different warps (issue cost, not the tensor
//                       pipe, limits one issuer); both stay converged and predicate the
//                       instruction on one elected lane
//   epilogue warps (4)  the epilogue of pipe_tc.cu (activation / act' / fused MSE)
//   2nd producer (1 lane, fwd + MSE)  the target tile as two TMA TENSOR copies (128B swizzle)
// TMEM (512 columns): ADJ 0..127 | P0,Plo0 128..255 | P1,Plo1 256..383 | O0 384..447 | O1 448..511
//
// Summation order: the tensor core adds the row's terms in column order with fp32
// (truncating) accumulation, not in the reference's entry order; with <= 128 terms the
// difference is ~1e-7 relative, inside the 1e-5 parity tolerance.
// Batches in which some (row, column) pair repeats (multi-edges) keep the list kernels.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "athena_internal.h"
#include "pipe_common.cuh"
#include "tc_common.cuh"

namespace athena {

using namespace tc;
using namespace pipe;

namespace {

// mbarrier wait.  Development builds (-DATHENA_TCG_WATCHDOG) trap with the barrier's tag
// instead of hanging the GPU when a barrier never completes; production builds wait without a
// deadline (a time-sliced or profiled context may legitimately stall for seconds).
__device__ __forceinline__ void mbar_wait_g(uint64_t* bar, uint32_t parity, int tag) {
#ifdef ATHENA_TCG_WATCHDOG
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
#else
  (void)tag;
  mbar_wait(bar, parity);
#endif
}

// Consecutive kernels of a step walk the tiles in opposite directions (GatherArgs::reverse):
// what the previous kernel wrote LAST (still in the 126 MB L2, much of it not even written
// back yet) is what this kernel reads FIRST.  `t` stays the logical index everywhere.
#define TCG_TILE(t) (a.reverse ? a.num_tiles - 1 - (t) : (t))

constexpr int TCG_TRACE_TILES = 16;
#define TCG_TRACE(role, slot)                                                            \
  do {                                                                                   \
    if (a.trace != nullptr && blockIdx.x == a.dbg && j >= 0 && j < TCG_TRACE_TILES)      \
      a.trace[((role) * TCG_TRACE_TILES + j) * 8 + (slot)] = clock64();                  \
  } while (0)

template <int EPI, int FW>
struct TcgCfg {
  // square steps of width 64 or 32 (F = N = FW); the 32-wide variant supports the forward,
  // fwd + MSE with a TMA-fed target and the reverse step with sign bits / without activation
  static_assert(FW == 64 || FW == 32, "feature width");
  static constexpr int F = FW, N = FW;
  // Ring stages.  The split warps pull a stage into registers as soon as it lands, so the
  // next copy is issued almost immediately; deeper rings measured the same or slower (fwd
  // 60.1 / 61.6 / 62.0 us for 1 / 2 / 3 stages), so the shared memory is left to the L1.
  static constexpr int NS = 1;
  static constexpr int SPLIT_WARP0 = 0;                          // warps 0..7
  static constexpr int SPLIT_THREADS = 256;
  static constexpr int BUILD_WARP0 = 8;                          // warps 8..11  (warp % 4 = TMEM lane quarter)
  // fix and epilogue: one warp per TMEM lane quarter.  (Tried: two of each per quarter, one
  // per 32-column half.  No gain -- 31 warps leave 64 registers per thread and the roles
  // wait for the L1 / shared-memory data path anyway.)
  static constexpr int FIX_WARP0 = 12;                           // warps 12..15
  static constexpr int EPI_WARP0 = 16;                           // warps 16..19
  static constexpr int PRODUCER_WARP = 20;
  static constexpr int MMA_WARP = PRODUCER_WARP + 1;             // issues G (and owns the TMEM allocation)
  static constexpr int MMA_T_WARP = PRODUCER_WARP + 2;           // issues T
  static constexpr int AUX_WARP = PRODUCER_WARP + 3;             // second producer (fwd + MSE)
  static constexpr int THREADS = (PRODUCER_WARP + 4) * 32;
  static constexpr int STAGE_W_THREADS = 8 * 32;                 // fix + epilogue warps
  static constexpr int X_BYTES = TILE_ROWS * F * 4;              // 32 KB of raw feature rows
  static constexpr int RS_BYTES = (TILE_ROWS + 8) * 4;
  static constexpr int AUX_TILE = TILE_ROWS * AUX_PITCH * 4;
  static constexpr int STAGE_BYTES = (X_BYTES + RS_BYTES + 127) / 128 * 128;
  static constexpr int OP_BLK = TILE_ROWS * 128;                 // [128 vertices x 32 features]
  static constexpr int OP_BYTES = (F / 32) * OP_BLK;             // hi (or lo) feature operand
  static constexpr int W_BLK = 2 * N * 128;                      // [hi(W') ; lo(W')] x 32 k
  static constexpr int OFF_BHI = 0;
  static constexpr int OFF_BLO = OP_BYTES;
  static constexpr int OFF_W = 2 * OP_BYTES;
  static constexpr int OFF_RING = OFF_W + (F / 32) * W_BLK;
  // fwd + MSE: the target tile is fetched by the producer with TMA tensor copies (128-byte
  // swizzle, two [128 x 32] boxes per tile, double-buffered) instead of cp.async from the
  // epilogue warps: no LSU traffic for the copy and a whole tile period of prefetch distance.
  // The cp.async path (padded tile in the same region) remains for unaligned targets.
  static constexpr bool AUX_TMA = EPI == EPI_MSE;
  static constexpr int AUX_TMA_TILE = TILE_ROWS * N * 4;
  static constexpr int OFF_BAR = OFF_RING + NS * STAGE_BYTES;    // barriers (256 B) + loss_red (512 B)
  static constexpr int OFF_AUX = (OFF_BAR + 256 + 512 + 1023) / 1024 * 1024;
  static constexpr int AUX_BYTES = AUX_TMA ? 2 * AUX_TMA_TILE : EPI != EPI_ACT ? AUX_TILE : 0;
  // Output staging for the TMA tensor stores: every fix / epilogue warp owns the 32 rows of its
  // TMEM lane quarter as two [32 rows x 32 floats] boxes in the 128-byte-swizzled layout
  // (16-byte chunk c of row r at chunk c ^ (r % 8): row-per-thread STS.128 is conflict-free),
  // which one lane hands to cp.async.bulk.tensor -- no read-back through the LSU, no STG.
  // fwd + MSE writes the gradient IN PLACE over the target tile it has just read.
  static constexpr int WARP_STAGE = (N / 32) * 32 * 128;         // 8 KB at N = 64
  static constexpr int OFF_EPI = (OFF_AUX + AUX_BYTES + 1023) / 1024 * 1024;
  static constexpr int EPI_STAGE_BYTES = EPI != EPI_MSE ? 4 * WARP_STAGE : 0;
  static constexpr int OFF_FIX = OFF_EPI + EPI_STAGE_BYTES;
  static constexpr int FIX_BYTES = EPI != EPI_ACTGRAD ? 4 * WARP_STAGE : 0;  // P is stored forward only
  static constexpr int SMEM = 1024 + OFF_FIX + FIX_BYTES;
  static_assert(SMEM <= 232448, "shared memory budget");
  static constexpr uint32_t T_ADJ = 0, T_P = 128, T_O = 384;     // TMEM columns
};

// byte offset of 16-byte chunk c (0..15) of row r (0..31) inside a warp's staging area: two
// [32 rows x 128 B] boxes (columns 0..31 | 32..63) in the TMA 128-byte swizzle
__device__ __forceinline__ uint32_t stage_off(int r, int c) {
  return static_cast<uint32_t>((c >> 3) * 4096 + r * 128 + (((c & 7) ^ (r & 7)) << 4));
}

// rows [0, rows) of a warp's staging area -> global rows (the fallback of the tensor store for
// a warp whose 32 rows are not all inside the tile, or without a tensor map): 16 lanes per row,
// conflict-free LDS.128, full-line STG.128.  box_pitch: bytes between the two column boxes.
template <int N>
__device__ __forceinline__ void stage_copy_out(const uint8_t* stage, uint32_t box_pitch, int rows,
                                               float* __restrict__ grow0, int lane) {
  constexpr int NC = N / 4;  // 16-byte chunks per row
#pragma unroll
  for (int it = 0; it < NC; ++it) {
    const int idx = it * 32 + lane;
    const int r = idx / NC, c = idx % NC;
    if (r < rows)
      *reinterpret_cast<float4*>(grow0 + r * N + c * 4) = *reinterpret_cast<const float4*>(
          stage + (c >> 3) * box_pitch + r * 128 + (((c & 7) ^ (r & 7)) << 4));
  }
}

// Epilogue of one tile for the warp that owns TMEM lanes 32q .. 32q+31 (thread = row), as
// pipe::epilogue_tile, but the results leave through a swizzled staging area and TMA tensor
// stores (see TcgCfg::WARP_STAGE).
//   EPI_ACT      out = act(v)                              -> this warp's staging area
//   EPI_ACTGRAD  out = v .* act'(aux | sign bits)          -> this warp's staging area
//   EPI_MSE      out = act'(p) .* (p - target) * scale     -> IN PLACE over the target values
//                (AUX_SWZ: the TMA-loaded tile, `dst` = its base; else this thread's padded row)
// Returns this row's loss share (EPI_MSE).  `dst_rows`: base the thread's chunks are written
// relative to (staging area of the warp, or the 128-row target tile with row_in_dst = 32 q + lane).
template <int ACT, int EPI, bool AUX_SWZ, int N>
__device__ __forceinline__ float tcg_epilogue(uint32_t tacc, int q, int lane, bool dry,
                                              uint8_t* dst, uint32_t box_pitch, int row_in_dst,
                                              const float* aux_row, float row_scale,
                                              uint64_t* acc_empty, uint32_t* mask_out_row,
                                              bool use_mask, const uint32_t (&mask_in)[N / 32]) {
  float lsum = 0.f;
  const bool padded_inplace = EPI == EPI_MSE && !AUX_SWZ;
  auto chunk_ptr = [&](int col) -> float4* {
    if (padded_inplace) return reinterpret_cast<float4*>(const_cast<float*>(aux_row) + col);
    const int c = col >> 2;
    return reinterpret_cast<float4*>(dst + (c >> 3) * box_pitch + row_in_dst * 128 +
                                     (((c & 7) ^ (row_in_dst & 7)) << 4));
  };
#pragma unroll
  for (int half = 0; half < N / 32; ++half) {
    uint32_t mbits = 0;
#pragma unroll
    for (int cg = 0; cg < 2; ++cg) {
      float vh[16];
      const int col0 = half * 32 + cg * 16;
      tmem_ld16(tacc + (static_cast<uint32_t>(q * 32) << 16) + col0, vh);
      if (half == N / 32 - 1 && cg == 1) {
        tc_fence_before();
        mbar_arrive(acc_empty);  // the accumulator buffer may be overwritten now
      }
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float o[4] = {vh[i], vh[i + 1], vh[i + 2], vh[i + 3]};
        float4* ptr = chunk_ptr(col0 + i);
        if (EPI == EPI_ACT) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (ACT == ATHENA_ACT_RELU || ACT == ATHENA_ACT_LEAKY_RELU)
              mbits |= (o[k] > 0.f ? 1u : 0u) << (cg * 16 + i + k);
            o[k] = act_fwd<ACT>(o[k]);
          }
        } else if (EPI == EPI_ACTGRAD && use_mask) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool pos = (mask_in[half] >> (cg * 16 + i + k)) & 1u;
            o[k] = pos ? o[k] : (ACT == ATHENA_ACT_LEAKY_RELU ? o[k] * 0.01f : 0.f);
          }
        } else if (EPI == EPI_MSE) {
          const float4 t4 = *ptr;  // the target sits where the gradient goes
          const float t[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float pk = act_fwd<ACT>(o[k]);
            const float d = pk - t[k];
            if (row_scale != 0.f) lsum += d * d * row_scale;
            o[k] = act_bwd<ACT>(pk, d * row_scale);
          }
        } else if (ACT != ATHENA_ACT_NONE) {
          const float4 h4 = *reinterpret_cast<const float4*>(aux_row + col0 + i);
          const float h[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = act_bwd<ACT>(h[k], o[k]);
        }
        if (!dry) *ptr = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    if (EPI == EPI_ACT && mask_out_row != nullptr &&
        (ACT == ATHENA_ACT_RELU || ACT == ATHENA_ACT_LEAKY_RELU))
      mask_out_row[half] = mbits;
  }
  return lsum;
}

// Every role runs its loop body once "dry" (iteration -1: an empty tile, no barrier traffic,
// scratch TMEM buffers) before the first real tile.  A CTA only owns ~14 tiles, and the
// first pass of each role through its code misses the instruction cache (measured: 4-9
// thousand clocks per role, serialised along the tile-0 dependency chain = a third of the
// kernel).  The dry pass takes those misses in all roles AT ONCE while the first TMA copy
// is in flight; it must execute the same instructions, hence one loop, not a copy.
template <bool TRANSB, int EPI, int FW>
__global__ void __launch_bounds__(TcgCfg<EPI, FW>::THREADS, 1)
k_pipe_tcg(GatherArgs a, const __grid_constant__ CUtensorMap aux_map,
           const __grid_constant__ CUtensorMap out_map, const __grid_constant__ CUtensorMap p_map) {
  using Cfg = TcgCfg<EPI, FW>;
  constexpr int F = Cfg::F, N = Cfg::N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sBhi = smem + Cfg::OFF_BHI;
  uint8_t* sBlo = smem + Cfg::OFF_BLO;
  uint8_t* sW = smem + Cfg::OFF_W;
  uint8_t* ring = smem + Cfg::OFF_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;                  // [NS <= 3]  TMA landed
  uint64_t* empty = bars + 3;             // [NS]  split warps hold the stage in registers
  uint64_t* ops_ready = bars + 6;         // feature operand (hi/lo) of tile j staged
  uint64_t* adj_ready = bars + 7;         // adjacency of tile j in TMEM
  uint64_t* g_done = bars + 8;            // G(j) finished: operand buffers + ADJ reusable
  uint64_t* p_full = bars + 9;            // [2] P accumulator of tile j complete
  uint64_t* p_fixed = bars + 11;          // [2] scaled hi/lo P back in TMEM
  uint64_t* o_full = bars + 13;           // [2]
  uint64_t* o_empty = bars + 15;          // [2]
  uint64_t* warm = bars + 17;             // W staged + dry passes of the fix / epilogue warps done
  uint64_t* dummy = bars + 18;            // sink for the arrivals of the dry epilogue pass
  uint64_t* t_done = bars + 19;           // [2] T(j) has read P[j & 1]
  uint64_t* aux_full = bars + 21;         // [2] target tile landed (TMA)
  uint64_t* aux_empty = bars + 23;        // [2] epilogue warps are done with it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);
  float* loss_red = reinterpret_cast<float*>(smem + Cfg::OFF_BAR + 256);  // [256], EPI_MSE
  float* sAux = reinterpret_cast<float*>(smem + Cfg::OFF_AUX);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int step = gridDim.x;
  if (a.trace != nullptr && tid == 0) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.trace[6 * TCG_TRACE_TILES * 8 + 2 * blockIdx.x] = gt;
    if (blockIdx.x == a.dbg) {
      a.trace[(5 * TCG_TRACE_TILES + 14) * 8 + 0] = clock64();
      a.trace[(5 * TCG_TRACE_TILES + 14) * 8 + 1] = gt;
    }
  }

  if (warp == Cfg::MMA_WARP) tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    for (int s = 0; s < Cfg::NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], Cfg::SPLIT_THREADS);
    }
    mbar_init(ops_ready, Cfg::SPLIT_THREADS);
    mbar_init(adj_ready, 128);
    mbar_init(g_done, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&p_full[b], 1);
      mbar_init(&p_fixed[b], 128);
      mbar_init(&o_full[b], 1);
      mbar_init(&o_empty[b], 128);
      mbar_init(&t_done[b], 1);
      mbar_init(&aux_full[b], 1);
      mbar_init(&aux_empty[b], 128);
    }
    mbar_init(warm, Cfg::STAGE_W_THREADS);
    mbar_init(dummy, 1u << 19);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // programmatic dependent launch: everything above overlapped the tail of the previous
  // kernel; from here on its results (features, gradients, updated weights) are read
  pdl_wait();
  // only now may the NEXT kernel start: everything before this kernel has completed, so a
  // dependent that skips ahead never overlaps a grid older than this one
  pdl_launch_dependents();
  const bool has_coef = a.rs != nullptr;
  constexpr int ns = Cfg::NS;
  int my_tiles = 0;
  for (int t = blockIdx.x; t < a.num_tiles; t += step) ++my_tiles;

  // stacked weight operand [hi(W') ; lo(W')], W' = op(W) as [N][F] K-major (as k_pipe_gather);
  // staged by the fix + epilogue warps, which have nothing to do until the first G finishes
  if (warp >= Cfg::FIX_WARP0 && warp < Cfg::PRODUCER_WARP) {
    for (int item = tid - Cfg::FIX_WARP0 * 32; item < N * (F / 4); item += Cfg::STAGE_W_THREADS) {
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
  }

  const bool aux_tma = Cfg::AUX_TMA && a.aux_tma != 0 && a.aux != nullptr;
  if (warp == Cfg::PRODUCER_WARP) {
    // ===================== producer: TMA bulk copies into the ring =====================
    if (lane == 0) {
      const uint64_t pol_first = l2_policy_evict_first(), pol_last = l2_policy_evict_last();
      int j = 0;
      for (int t = blockIdx.x; t < a.num_tiles; t += step, ++j) {
        const int s = j % ns;
        const uint32_t ph = (j / ns) & 1;
        TCG_TRACE(0, 0);
        mbar_wait_g(&empty[s], ph ^ 1u, 0);
        TCG_TRACE(0, 1);
        const int4 ti = __ldg(a.tiles + TCG_TILE(t));
        const int r0 = ti.x, nrows = ti.y;
        const int ra = r0 & ~3, rcnt = (r0 + nrows - ra + 3) & ~3;
        uint8_t* st = ring + s * Cfg::STAGE_BYTES;
        const uint32_t xb = nrows * F * 4, rb = rcnt * 4;
        mbar_arrive_expect_tx(&full[s], xb + (has_coef ? rb : 0u));
        if (a.hint_x == 1) bulk_g2s_hint(st, a.X + static_cast<size_t>(r0) * F, xb, &full[s], pol_first);
        else if (a.hint_x == 2) bulk_g2s_hint(st, a.X + static_cast<size_t>(r0) * F, xb, &full[s], pol_last);
        else bulk_g2s(st, a.X + static_cast<size_t>(r0) * F, xb, &full[s]);
        if (has_coef) bulk_g2s(st + Cfg::X_BYTES, a.rs + ra, rb, &full[s]);
      }
    }
  } else if (warp == Cfg::AUX_WARP) {
    // ===================== second producer: the epilogue's operand tile (target) =======
    // two [128 rows x 32 floats] TMA tensor copies per tile, at most two tiles ahead
    if (lane == 0 && aux_tma) {
      uint8_t* sAuxT = smem + Cfg::OFF_AUX;
      int j = 0;
      for (int t = blockIdx.x; t < a.num_tiles; t += step, ++j) {
        const int b = j & 1;
        mbar_wait_g(&aux_empty[b], ((j >> 1) & 1) ^ 1u, 13);
        const int4 ti = __ldg(a.tiles + TCG_TILE(t));
        mbar_arrive_expect_tx(&aux_full[b], Cfg::AUX_TMA_TILE);
        if (a.hint_aux == 1) {
          const uint64_t pol = l2_policy_evict_first();
#pragma unroll
          for (int h = 0; h < N / 32; ++h)
            tma_load_2d_hint(sAuxT + b * Cfg::AUX_TMA_TILE + h * TILE_ROWS * 128, &aux_map, 32 * h,
                             ti.x, &aux_full[b], pol);
        } else {
#pragma unroll
          for (int h = 0; h < N / 32; ++h)
            tma_load_2d(sAuxT + b * Cfg::AUX_TMA_TILE + h * TILE_ROWS * 128, &aux_map, 32 * h,
                        ti.x, &aux_full[b]);
        }
      }
    }
  } else if (warp == Cfg::MMA_WARP) {
    // ===================== MMA issuer, G: [P(hi) | P(lo)] = ADJ . [Xhi | Xlo] ==========
    // Every tcgen05.mma costs tens of clocks of issue latency in the issuing thread whatever
    // its N.  Hence (a) the hi and lo halves of the feature operand are stacked along N (one
    // N = 128 instruction per k-step), (b) G and T are issued by two different warps; the
    // order between the two streams is carried by mbarriers (p_fixed, t_done), and (c) the
    // warp stays converged with the instruction predicated on one elected lane.
    {
      const uint32_t leader = elect_one();
      const uint32_t bHi = smem_u32(sBhi);
      constexpr uint32_t IDESC_G = make_idesc(128, 2 * F, false, true);  // B MN-major
      for (int j = -1; j < my_tiles; ++j) {
        const bool dry = j < 0;
        if (!dry) {
          if (lane == 0) TCG_TRACE(1, 0);
          mbar_wait_g(ops_ready, j & 1, 1);
          mbar_wait_g(adj_ready, j & 1, 2);
          if (j >= 2) mbar_wait_g(&t_done[j & 1], ((j - 2) >> 1) & 1, 11);  // T(j-2) has read P[j & 1]
          if (lane == 0) TCG_TRACE(1, 1);
        }
        tc_fence_after();
        const uint32_t tp = tmem + Cfg::T_P + (dry ? 128u : (j & 1) * 128u);
#pragma unroll
        for (int ks = 0; ks < TILE_ROWS / 8; ++ks)
          umma_tf32_ts_w(leader, tp, tmem + Cfg::T_ADJ + ks * 8,
                         make_desc_mn32(bHi + ks * 1024, Cfg::OP_BLK, 512), IDESC_G,
                         ks ? 1u : 0u);
        if (!dry) {
          umma_commit_w(leader, g_done);
          umma_commit_w(leader, &p_full[j & 1]);
          if (lane == 0) TCG_TRACE(1, 2);
        } else {
          mbar_wait_g(warm, 0, 10);  // the dry pass of the fix warps (it writes P[1]) is over
        }
        __syncwarp();
      }
    }
  } else if (warp == Cfg::MMA_T_WARP) {
    // ===================== MMA issuer, T: O = Plo . Whi + P . Wlo + P . Whi ==============
    // (small terms first: the tensor core adds into its fp32 accumulator with truncation)
    {
      const uint32_t leader = elect_one();
      const uint32_t wAddr = smem_u32(sW);
      constexpr uint32_t IDESC_T = make_idesc(128, N, false, false);  // B = W', K-major
      for (int j = -1; j < my_tiles; ++j) {
        const bool dry = j < 0;
        const int b = dry ? 1 : (j & 1);
        if (!dry) {
          if (lane == 0) TCG_TRACE(1, 3);
          mbar_wait_g(&p_fixed[b], (j >> 1) & 1, 3);
          mbar_wait_g(&o_empty[b], ((j >> 1) & 1) ^ 1u, 4);
          if (lane == 0) TCG_TRACE(1, 4);
        }
        tc_fence_after();
        const uint32_t tp = tmem + Cfg::T_P + b * 128, to = tmem + Cfg::T_O + b * N;
#pragma unroll
        for (int k8 = 0; k8 < F / 8; ++k8) {
          const uint32_t wk = wAddr + (k8 >> 2) * Cfg::W_BLK + (k8 & 3) * 32;
          umma_tf32_ts_w(leader, to, tp + F + k8 * 8, make_desc(wk, 16, 1024), IDESC_T,
                         k8 ? 1u : 0u);
          umma_tf32_ts_w(leader, to, tp + k8 * 8, make_desc(wk + N * 128, 16, 1024), IDESC_T, 1u);
        }
#pragma unroll
        for (int k8 = 0; k8 < F / 8; ++k8) {
          const uint32_t wk = wAddr + (k8 >> 2) * Cfg::W_BLK + (k8 & 3) * 32;
          umma_tf32_ts_w(leader, to, tp + k8 * 8, make_desc(wk, 16, 1024), IDESC_T, 1u);
        }
        if (!dry) {
          umma_commit_w(leader, &o_full[b]);
          umma_commit_w(leader, &t_done[b]);
          if (lane == 0) TCG_TRACE(1, 5);
        } else {
          mbar_wait_g(warm, 0, 12);  // the weights are staged
        }
        __syncwarp();
      }
    }
  } else if (warp >= Cfg::EPI_WARP0) {
    // ===================== epilogue: TMEM -> registers -> global rows ==================
    const int q = warp - Cfg::EPI_WARP0;
    uint8_t* stage = smem + Cfg::OFF_EPI + q * Cfg::WARP_STAGE;  // EPI != EPI_MSE
    const bool use_mask = (EPI == EPI_ACTGRAD) && a.mask_in != nullptr;
    const bool use_aux = (EPI != EPI_ACT) && a.aux != nullptr && !use_mask;
    const bool out_tma = (a.store_tma & 1) != 0;
    const int my_row = q * 32 + lane;
    auto tile_at = [&](int t) { return t < a.num_tiles ? __ldg(a.tiles + TCG_TILE(t)) : make_int4(0, 0, 0, 0); };
    auto count_at = [&](int t) {
      int c = 1;
      if (EPI == EPI_MSE && t < a.num_tiles) {
        const int4 ti = __ldg(a.tiles + TCG_TILE(t));
        if (my_row < ti.y) c = __ldg(a.vcount + ti.x + my_row);
      }
      return c;
    };
    // Second operand without TMA (saved activations of the backward pass; unaligned targets):
    // each warp prefetches its own 32 rows of the next tile with cp.async into rows padded to
    // 272 B (conflict-free row-per-thread reads) when it is done with the current one.
    // (Two such tiles, copy for tile j + 2 issued when tile j is done, measured slower.)
    auto issue_aux = [&](int t) {
      const int4 ti = tile_at(t);
      const float* src = a.aux + (static_cast<size_t>(ti.x) + q * 32) * N;
      float* dst = sAux + q * 32 * AUX_PITCH;
      const int rows = min(32, ti.y - q * 32);
#pragma unroll
      for (int it = 0; it < N / 4; ++it) {
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
    int count_next = count_at(blockIdx.x);
    float lsum = 0.f;
    for (int j = -1; j < my_tiles; ++j) {
      const bool dry = j < 0;
      const int t = blockIdx.x + j * step;
      const int b = dry ? 1 : (j & 1);
      const int4 ti = dry ? make_int4(0, 0, 0, 0) : __ldg(a.tiles + TCG_TILE(t));
      const int count = count_next;
      if (!dry) count_next = count_at(t + step);
      const bool row_valid = my_row < ti.y;
      const int rows_valid = min(32, ti.y - q * 32);  // rows of this warp inside the tile (<= 0: none)
      uint32_t min_w[N / 32];
#pragma unroll
      for (int h = 0; h < N / 32; ++h) min_w[h] = 0u;
      uint32_t* mout = nullptr;
      if (row_valid) {
        const size_t grow = static_cast<size_t>(ti.x) + my_row;
        if (use_mask) {
#pragma unroll
          for (int h = 0; h < N / 32; ++h) min_w[h] = __ldg(a.mask_in + grow * (N / 32) + h);
        }
        if (EPI == EPI_ACT && a.mask_out != nullptr) mout = a.mask_out + grow * (N / 32);
      }
      // the staging area is free again once the previous tile's tensor stores have read it
      if (EPI != EPI_MSE) {
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
      }
      if (!dry) {
        if (q == 0 && lane == 0) TCG_TRACE(5, 0);
        mbar_wait_g(&o_full[b], (j >> 1) & 1, 5);
        if (q == 0 && lane == 0) TCG_TRACE(5, 1);
      }
      tc_fence_after();
      if (aux_tma) {
        if (!dry) mbar_wait_g(&aux_full[j & 1], (j >> 1) & 1, 14);
      } else if (use_aux && !dry) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
      }
      uint8_t* aux_tile = smem + Cfg::OFF_AUX + (j & 1) * Cfg::AUX_TMA_TILE;  // aux_tma only
      const float* aux_row = sAux + my_row * AUX_PITCH;                       // padded tile
      const uint32_t tacc = tmem + Cfg::T_O + b * N;
      uint64_t* release = dry ? dummy : &o_empty[b];
      const float scale =
          (EPI == EPI_MSE && row_valid) ? 1.f / static_cast<float>(N * count) : 0.f;
      const int act = use_aux || use_mask || EPI == EPI_ACT ? a.act : ATHENA_ACT_NONE;
      const bool inplace_swz = EPI == EPI_MSE && aux_tma;
      uint8_t* dst = inplace_swz ? aux_tile : stage;
      const uint32_t box_pitch = inplace_swz ? TILE_ROWS * 128 : 4096;
      const int row_in_dst = inplace_swz ? my_row : lane;
#define TCG_EPI(ACT)                                                                            \
  lsum += inplace_swz ? tcg_epilogue<ACT, EPI, true, N>(tacc, q, lane, dry, dst, box_pitch,      \
                                                        row_in_dst, aux_row, scale, release,     \
                                                        mout, use_mask, min_w)                   \
                      : tcg_epilogue<ACT, EPI, false, N>(tacc, q, lane, dry, dst, box_pitch,     \
                                                         row_in_dst, aux_row, scale, release,    \
                                                         mout, use_mask, min_w)
      switch (act) {
        case ATHENA_ACT_RELU: TCG_EPI(ATHENA_ACT_RELU); break;
        case ATHENA_ACT_LEAKY_RELU: TCG_EPI(ATHENA_ACT_LEAKY_RELU); break;
        case ATHENA_ACT_SIGMOID: TCG_EPI(ATHENA_ACT_SIGMOID); break;
        case ATHENA_ACT_TANH: TCG_EPI(ATHENA_ACT_TANH); break;
        default: TCG_EPI(ATHENA_ACT_NONE); break;
      }
#undef TCG_EPI
      // hand the rows to the TMA (a full warp of rows) or copy them out (ragged tile end)
      fence_async_smem();
      __syncwarp();
      float* grow0 = a.out + (static_cast<size_t>(ti.x) + q * 32) * N;
      if (EPI == EPI_MSE && !aux_tma) {
        // padded target tile, gradient written in place: rows of 272 B
        if (!dry) {
#pragma unroll
          for (int it = 0; it < N / 4; ++it) {
            const int idx = it * 32 + lane;
            const int r = idx / (N / 4), c = idx % (N / 4);
            if (r < rows_valid)
              *reinterpret_cast<float4*>(grow0 + r * N + c * 4) =
                  *reinterpret_cast<const float4*>(sAux + (q * 32 + r) * AUX_PITCH + c * 4);
          }
          __syncwarp();
        }
      } else {
        const uint8_t* src = inplace_swz ? aux_tile + q * 4096 : stage;
        if (out_tma && rows_valid == 32) {
          if (lane == 0) {
            if (a.hint_out == 2) {
              const uint64_t pol = l2_policy_evict_last();
#pragma unroll
              for (int h = 0; h < N / 32; ++h)
                tma_store_2d_hint(&out_map, 32 * h, ti.x + q * 32, src + h * box_pitch, pol);
            } else {
#pragma unroll
              for (int h = 0; h < N / 32; ++h)
                tma_store_2d(&out_map, 32 * h, ti.x + q * 32, src + h * box_pitch);
            }
            bulk_commit();
            if (inplace_swz) bulk_wait_read0();  // the target buffer is handed back below
          }
          __syncwarp();
        } else if (rows_valid > 0) {
          stage_copy_out<N>(src, box_pitch, rows_valid, grow0, lane);
          __syncwarp();
        }
      }
      if (dry) {
        lsum = 0.f;  // whatever the scratch accumulator held
        mbar_arrive(warm);
      } else {
        if (aux_tma) mbar_arrive(&aux_empty[j & 1]);
        if (q == 0 && lane == 0) TCG_TRACE(5, 2);
      }
      if (use_aux && !aux_tma) {
        __syncwarp();
        issue_aux(t + step);  // the first tile right after the dry pass
      }
    }
    if (lane == 0) bulk_wait_read0();  // shared memory stays valid until the last store has read it
    if (EPI == EPI_MSE) loss_red[my_row] = lsum;
  } else if (warp >= Cfg::FIX_WARP0) {
    // ===================== fix warps: P * deg_v^-1/2 -> hi / lo back into TMEM, P -> global ====
    const int q = warp - Cfg::FIX_WARP0;
    const int my_row = q * 32 + lane;
    uint8_t* stage = smem + Cfg::OFF_FIX + q * Cfg::WARP_STAGE;  // P staging (forward only)
    const bool p_tma = (a.store_tma & 2) != 0;
    auto rs_at = [&](int t) {
      float w = 1.f;
      if (has_coef && t < a.num_tiles) {
        const int4 ti = __ldg(a.tiles + TCG_TILE(t));
        if (my_row < ti.y) w = __ldg(a.rs + ti.x + my_row);
      }
      return w;
    };
    float w_next = rs_at(blockIdx.x);
    for (int j = -1; j < my_tiles; ++j) {
      const bool dry = j < 0;
      const int t = blockIdx.x + j * step;
      const int b = dry ? 1 : (j & 1);
      const int4 ti = dry ? make_int4(0, 0, 0, 0) : __ldg(a.tiles + TCG_TILE(t));
      const float wv = w_next;
      if (!dry) {
        w_next = rs_at(t + step);
        if (q == 0 && lane == 0) TCG_TRACE(4, 0);
        mbar_wait_g(&p_full[b], (j >> 1) & 1, 6);
        if (q == 0 && lane == 0) TCG_TRACE(4, 1);
      }
      tc_fence_after();
      const uint32_t tp = tmem + Cfg::T_P + b * 128 + (static_cast<uint32_t>(q * 32) << 16);
      const bool store_p = EPI != EPI_ACTGRAD && a.P != nullptr;
      if (store_p) {  // the previous tile's tensor stores have read the staging area
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
      }
#pragma unroll 1
      for (int g = 0; g < F / 16; ++g) {
        // P = (A . Xhi + A . Xlo) * deg_v^-1/2 ; hi -> columns 0..63, lo -> columns 64..127
        float v[16], vl[16];
        tmem_ld16_nowait(tp + g * 16, v);
        tmem_ld16_nowait(tp + F + g * 16, vl);
        tmem_ld_wait();
        if (!dry && g == 0 && q == 0 && lane == 0) TCG_TRACE(4, 3);
        uint32_t hi[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          v[i] = (v[i] + vl[i]) * wv;
          hi[i] = __float_as_uint(v[i]) & 0xffffe000u;
        }
        tmem_st16(tp + g * 16, hi);
#pragma unroll
        for (int i = 0; i < 16; ++i) hi[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));
        tmem_st16(tp + F + g * 16, hi);
        if (!dry && g == 0 && q == 0 && lane == 0) TCG_TRACE(4, 4);
        if (store_p && !dry) {
          // the propagated tile (the operand of dW = P^T gY) goes to the swizzled staging area;
          // one TMA tensor store per column box follows below
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<float4*>(stage + stage_off(lane, g * 4 + k)) =
                make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
        if (!dry && g == 0 && q == 0 && lane == 0) TCG_TRACE(4, 5);
      }
      tmem_st_wait();
      tc_fence_before();
      if (dry) {
        mbar_arrive(warm);
      } else {
        mbar_arrive(&p_fixed[b]);
        if (q == 0 && lane == 0) TCG_TRACE(4, 2);
      }
      if (store_p && !dry) {
        const int rows_valid = min(32, ti.y - q * 32);
        fence_async_smem();
        __syncwarp();
        if (p_tma && rows_valid == 32) {
          if (lane == 0) {
            if (a.hint_p == 1) {
              const uint64_t pol = l2_policy_evict_first();
#pragma unroll
              for (int h = 0; h < F / 32; ++h)
                tma_store_2d_hint(&p_map, 32 * h, ti.x + q * 32, stage + h * 4096, pol);
            } else {
#pragma unroll
              for (int h = 0; h < F / 32; ++h)
                tma_store_2d(&p_map, 32 * h, ti.x + q * 32, stage + h * 4096);
            }
            bulk_commit();
          }
        } else if (rows_valid > 0) {
          stage_copy_out<F>(stage, 4096, rows_valid,
                            a.P + (static_cast<size_t>(ti.x) + q * 32) * F, lane);
        }
        __syncwarp();
      }
    }
    if (lane == 0) bulk_wait_read0();
  } else if (warp >= Cfg::BUILD_WARP0) {
    // ===================== build warps: adjacency bits -> TMEM A operand ================
    const int q = warp - Cfg::BUILD_WARP0;
    const int my_row = q * 32 + lane;
    auto bits_at = [&](int t) {
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      if (t < a.num_tiles) {
        const int4 ti = __ldg(a.tiles + TCG_TILE(t));
        if (my_row < ti.y) w = __ldg(a.abits + ti.x + my_row);
      }
      return w;
    };
    uint4 b_next = bits_at(blockIdx.x);
    for (int j = -1; j < my_tiles; ++j) {
      const bool dry = j < 0;  // the dry pass expands tile 0 as well (and is overwritten by it)
      const int t = blockIdx.x + j * step;
      const uint4 bits = b_next;
      if (!dry) {
        b_next = bits_at(t + step);
        if (q == 0 && lane == 0) TCG_TRACE(3, 0);
        mbar_wait_g(g_done, (j & 1) ^ 1u, 7);  // G(j-1) has read the previous adjacency
        if (q == 0 && lane == 0) TCG_TRACE(3, 1);
      }
      tc_fence_after();
      const uint32_t ta = tmem + Cfg::T_ADJ + (static_cast<uint32_t>(q * 32) << 16);
      uint4 rot = bits;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const uint32_t w = rot.x;
        rot = make_uint4(rot.y, rot.z, rot.w, 0u);
        uint32_t v[32];
        if (__any_sync(0xffffffffu, w != 0u)) {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = ((w >> k) & 1u) ? 0x3f800000u : 0u;
        } else {
          // block diagonal: no row of this warp reaches these 32 vertices
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = 0u;
        }
        tmem_st32(ta + c * 32, v);
      }
      tmem_st_wait();
      tc_fence_before();
      if (!dry) {
        mbar_arrive(adj_ready);
        if (q == 0 && lane == 0) TCG_TRACE(3, 2);
      }
    }
  } else {
    // ===================== split warps: ring -> scaled hi / lo B operand ================
    constexpr int LOADS = TILE_ROWS * (F / 4) / Cfg::SPLIT_THREADS;  // 16-byte chunks per thread
    float nf = 0.f;  // x * 0 summed over everything this thread reads: NaN iff a value is not finite
    for (int j = -1; j < my_tiles; ++j) {
      const bool dry = j < 0;
      const int t = blockIdx.x + j * step;
      const int jr = dry ? 0 : j;
      const int s = jr % ns;
      const uint32_t ph = (jr / ns) & 1;
      const int4 ti = dry ? make_int4(0, 0, 0, 0) : __ldg(a.tiles + TCG_TILE(t));
      const int r0 = ti.x, nrows = ti.y;
      const uint8_t* st = ring + s * Cfg::STAGE_BYTES;
      const float* rss = reinterpret_cast<const float*>(st + Cfg::X_BYTES) + (r0 - (r0 & ~3));
      if (!dry) {
        if (tid == 0) TCG_TRACE(2, 0);
        mbar_wait_g(&full[s], ph, 8);
        if (tid == 0) TCG_TRACE(2, 1);
      }
      float4 x[LOADS];
#pragma unroll
      for (int i = 0; i < LOADS; ++i) {
        const int idx = tid + Cfg::SPLIT_THREADS * i;
        const int row = idx / (F / 4);
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
      if (!dry) {
        mbar_arrive(&empty[s]);                // the stage lives in registers now
        if (tid == 0) TCG_TRACE(2, 2);
        mbar_wait_g(g_done, (j & 1) ^ 1u, 9);  // G(j-1) has read the operand buffers
        if (tid == 0) TCG_TRACE(2, 3);
      }
#pragma unroll
      for (int i = 0; i < LOADS; ++i) {
        const int idx = tid + Cfg::SPLIT_THREADS * i;
        const int row = idx / (F / 4), ch = idx % (F / 4);
        float4 hi, lo;
        split_tf32(x[i], hi, lo);
        if (EPI == EPI_ACT)
          nf = fmaf(x[i].x, 0.f, fmaf(x[i].y, 0.f, fmaf(x[i].z, 0.f, fmaf(x[i].w, 0.f, nf))));
        const uint32_t off = (ch >> 3) * Cfg::OP_BLK + sw128b32_off(row, ch & 7);
        *reinterpret_cast<float4*>(sBhi + off) = hi;
        *reinterpret_cast<float4*>(sBlo + off) = lo;
      }
      fence_async_smem();
      if (!dry) {
        mbar_arrive(ops_ready);
        if (tid == 0) TCG_TRACE(2, 4);
      }
    }
    if (EPI == EPI_ACT && a.nonfinite != nullptr && nf != nf) atomicOr(a.nonfinite, 1);
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
  if (a.trace != nullptr && tid == 0) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    a.trace[6 * TCG_TRACE_TILES * 8 + 2 * blockIdx.x + 1] = gt;
    if (blockIdx.x == a.dbg) {
      a.trace[(5 * TCG_TRACE_TILES + 15) * 8 + 0] = clock64();
      a.trace[(5 * TCG_TRACE_TILES + 15) * 8 + 1] = gt;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::MMA_WARP) tmem_dealloc<512>(tmem);
}

// 2-D tensor map of a dense [rows][64] fp32 array, box = [box_rows][32 floats], 128-byte
// swizzle (box_rows = 128: the target tile of fwd + MSE; 32: one warp's rows for the stores).
// The driver entry point is resolved at run time (no link-time libcuda dependency).
static bool make_row_tile_map(const float* base, long long rows, CUtensorMap* out,
                              int box_rows = TILE_ROWS, int width = 64) {
  using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                          const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                          CUtensorMapFloatOOBfill);
  static Fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<Fn>(p);
  }
  if (!fn || rows <= 0 || (reinterpret_cast<uintptr_t>(base) & 15u) != 0) return false;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstr[1] = {width * sizeof(float)};
  const cuuint32_t box[2] = {32, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool TRANSB, int EPI, int FW>
int launch_tcg_t(const GatherArgs& a) {
  using Cfg = TcgCfg<EPI, FW>;
  static bool attr = false;
  if (!attr) {
    ATH_CUDA(cudaFuncSetAttribute(k_pipe_tcg<TRANSB, EPI, FW>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  const int grid = std::min(a.num_tiles, ctx().sm_count);
  GatherArgs b = a;
  b.trace = nullptr;
  // L2 policies of the bulk copies: what this kernel reads for the last time must not evict
  // what the next kernel reads first.  ATHENA_DEBUG_L2_HINTS = bit mask (A/B runs): 1 input
  // tile evict_first (forward), 2 target evict_first, 4 P store evict_first (only the dW
  // product at the end of the reverse sweep reads it), 8 output store evict_last (the next
  // kernel's input), 16 dW product: operands evict_first.
  const int hints = l2_hint_mask();
  // 32 / 64: gradient written by fwd + MSE / by the reverse step evict_last; 128: the reverse
  // step's input gradient (read again by the dW product) evict_last
  b.hint_x = EPI == EPI_ACTGRAD ? ((hints & 128) ? 2 : 0) : ((hints & 1) ? 1 : 0);
  b.hint_aux = (hints & 2) && EPI == EPI_MSE ? 1 : 0;
  b.hint_p = (hints & 4) ? 1 : 0;
  b.hint_out = EPI == EPI_ACT ? ((hints & 8) ? 2 : 0)
               : EPI == EPI_MSE ? ((hints & 32) ? 2 : 0) : ((hints & 64) ? 2 : 0);
  b.reverse = ctx().tile_reverse ? 1 : 0;
  static int no_reverse = -1;
  if (no_reverse < 0) {
    const char* e = getenv("ATHENA_DEBUG_NO_REVERSE");  // A/B switch: always first to last
    no_reverse = (e && atoi(e) != 0) ? 1 : 0;
  }
  if (!no_reverse) ctx().tile_reverse = !ctx().tile_reverse;
  alignas(64) CUtensorMap aux_map;
  memset(&aux_map, 0, sizeof(aux_map));
  static int no_tma = -1;
  if (no_tma < 0) {
    const char* e = getenv("ATHENA_DEBUG_NO_AUX_TMA");  // A/B switch: cp.async operand prefetch
    no_tma = (e && atoi(e) != 0) ? 1 : 0;
  }
  b.aux_tma = (Cfg::AUX_TMA && !no_tma && a.aux != nullptr &&
               make_row_tile_map(a.aux, a.num_rows, &aux_map, TILE_ROWS, FW))
                  ? 1
                  : 0;
  // TMA tensor stores of the output (and of the propagated tile): [32 rows x 32 floats] boxes
  alignas(64) CUtensorMap out_map, p_map;
  memset(&out_map, 0, sizeof(out_map));
  memset(&p_map, 0, sizeof(p_map));
  static int no_store_tma = -1;
  if (no_store_tma < 0) {
    const char* e = getenv("ATHENA_DEBUG_NO_STORE_TMA");  // A/B switch: LDS / STG copy-out
    no_store_tma = (e && atoi(e) != 0) ? 1 : 0;
  }
  b.store_tma = 0;
  if (!no_store_tma) {
    if (a.out != nullptr && make_row_tile_map(a.out, a.num_rows, &out_map, 32, FW))
      b.store_tma |= 1;
    if (EPI != EPI_ACTGRAD && a.P != nullptr && make_row_tile_map(a.P, a.num_rows, &p_map, 32, FW))
      b.store_tma |= 2;
  }
  static int trace_left = -1;
  static long long* trace_buf = nullptr;
  if (trace_left < 0) {
    const char* e = getenv("ATHENA_DEBUG_TRACE");
    trace_left = e ? atoi(e) : 0;
  }
  const size_t trace_n = (size_t)6 * TCG_TRACE_TILES * 8 + 2 * 256;
  if (trace_left > 0) {
    if (!trace_buf) cudaMalloc(&trace_buf, trace_n * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, trace_n * sizeof(long long), ctx().stream);
    b.trace = trace_buf;
    const char* e = getenv("ATHENA_DEBUG_TRACE_BLOCK");
    b.dbg = e ? atoi(e) : 0;
  }
  ATH_CUDA(launch_pdl(k_pipe_tcg<TRANSB, EPI, FW>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM,
                      ctx().stream, b, aux_map, out_map, p_map));
  if (trace_left > 0) {
    --trace_left;
    std::vector<long long> h(trace_n);
    cudaStreamSynchronize(ctx().stream);
    cudaMemcpy(h.data(), trace_buf, trace_n * sizeof(long long), cudaMemcpyDeviceToHost);
    const long long t0 = h[(2 * TCG_TRACE_TILES + 0) * 8 + 0];
    static const char* names[6] = {"producer", "mma", "split", "build", "fix", "epilogue"};
    fprintf(stderr, "TRACE k_pipe_tcg EPI=%d\n", EPI);
    {
      const long long* se = h.data() + 6 * TCG_TRACE_TILES * 8;
      long long s0 = se[0];
      for (int c = 0; c < grid; ++c) s0 = std::min(s0, se[2 * c]);
      fprintf(stderr, "TRACE cta start/end (ns after first start), tiles:");
      for (int c = 0; c < grid; ++c)
        fprintf(stderr, " %d:%lld/%lld", c, se[2 * c] - s0, se[2 * c + 1] - s0);
      fprintf(stderr, "\n");
    }
    for (int role = 0; role < 6; ++role)
      for (int j = 0; j < 16; ++j) {
        fprintf(stderr, "TRACE %-8s tile %2d:", names[role], j);
        for (int k = 0; k < 6; ++k) {
          const long long v = h[(role * TCG_TRACE_TILES + j) * 8 + k];
          fprintf(stderr, " %7lld", v ? (role == 5 && j >= 14 && k == 1 ? v % 100000000ll : v - t0) : -1);
        }
        fprintf(stderr, "\n");
      }
  }
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
  if (off || b->num_tiles == 0 || F != N || (F != 64 && F != 32) || b->abits == nullptr)
    return false;
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

int launch_pipe_tcg(const GatherArgs& a, bool transb, int epi, int width) {
  ATH_REQUIRE(width == 64 || width == 32, ATHENA_ERR_ARG, "pipe_tcg: unsupported width %d", width);
  if (width == 32) {
    // no padded (non-TMA) operand tile at this width: the target must be TMA-addressable, the
    // reverse step must take act' from sign bits (or have none)
    ATH_REQUIRE(epi != EPI_ACTGRAD || a.aux == nullptr || a.mask_in != nullptr, ATHENA_ERR_ARG,
                "pipe_tcg: width 32 needs sign bits for the activation derivative");
    ATH_REQUIRE(epi != EPI_MSE || (reinterpret_cast<uintptr_t>(a.aux) & 15u) == 0, ATHENA_ERR_ARG,
                "pipe_tcg: width 32 needs a 16-byte aligned target");
    if (epi == EPI_ACT) return launch_tcg_t<false, EPI_ACT, 32>(a);
    if (epi == EPI_MSE) return launch_tcg_t<false, EPI_MSE, 32>(a);
    ATH_REQUIRE(transb && epi == EPI_ACTGRAD, ATHENA_ERR_ARG, "pipe_tcg: unsupported variant");
    return launch_tcg_t<true, EPI_ACTGRAD, 32>(a);
  }
  if (epi == EPI_ACT) return launch_tcg_t<false, EPI_ACT, 64>(a);
  if (epi == EPI_MSE) return launch_tcg_t<false, EPI_MSE, 64>(a);
  ATH_REQUIRE(transb && epi == EPI_ACTGRAD, ATHENA_ERR_ARG, "pipe_tcg: unsupported variant");
  return launch_tcg_t<true, EPI_ACTGRAD, 64>(a);
}

}  // namespace athena
