// Fused, warp-specialised tcgen05 kernels for mini-batches of small graphs
// (every graph fits one 128-vertex tile; Batch::tiles).
//
//   k_pipe_gather<F,N,TRANSB,EPI>
//       out[v,:] = epi( ( sum_{w in row v} c_w * X[col[w],:] ) . op(W) )
//     forward  (CSR, c = (deg_v deg_u)^-1/2, W_t,   epi = activation):
//         kipf_propagate + matmul + activation%apply in ONE pass
//         (athena_diffstruc_extd_sub_kipf.f90:29-46, athena_kipf_msgpass_layer.f90:943-952);
//         also stores the propagated tile P for the backward pass.
//     forward + MSE (last layer of a training step): the epilogue additionally reads the
//         target tile, writes d loss / d pre-activation instead of the output and reduces
//         the loss per CTA (mse_loss_type%compute, athena_loss.f90:393-430).
//     backward (CSC, c = 1,                  W_t^T, epi = .* act'(H_{t-1})):
//         gY_{t-1} = ( sum_{v->u} gY_t[v] ) W_t^T .* act'(H_{t-1}); the un-normalised
//         scatter of get_partial_kipf_propagate_left_val (:101-109) commutes with the
//         linear map, so dP is never materialised.
//   k_pipe_tn<N>
//       dW[64 x N] += P^T . gY over all vertices (matmul partial w.r.t. W_t)
//
// Roles of k_pipe_gather (one CTA per SM, persistent over tiles):
//   producer warp   1 lane issues TMA bulk copies (cp.async.bulk) of the tile's feature
//                   rows, row pointers, tile-local neighbour bytes and deg^-1/2 into a
//                   shared-memory ring; completion is counted on "full" mbarriers
//   gather warps    512 threads, FOUR lanes per row, 16 features per lane: walk the row's
//                   entries in ascending order (the reference's summation order); four
//                   neighbour bytes are one LDS.32, every entry is four LDS.128 whose
//                   64-byte halves alternate between the two rows of a quarter-warp, so every
//                   shared-memory access is conflict-free; split the result hi/lo and
//                   store it as the swizzled K-major A operand; store P
//   MMA warp        1 lane issues tcgen05.mma.kind::tf32 into a double-buffered TMEM
//                   accumulator and commits to mbarriers
//   epilogue warps  128 threads: tcgen05.ld (row per thread), hi+lo, activation / act' /
//                   MSE against the thread's row of a padded operand tile that the warp
//                   prefetched with cp.async, transposition through a padded patch,
//                   coalesced row stores
// so the HBM stream, the shared-memory gather, the tensor pipe and the stores of
// consecutive tiles overlap.  No float atomics; fixed tile -> CTA mapping.
//
// Debug aids (never set in production): ATHENA_DEBUG_PIPE = bitmask that disables the P
// store (1), the output store (2) or the gather loop (4) for elimination experiments;
// ATHENA_DEBUG_TRACE = n prints a clock64 timeline of the four roles of CTA 0 for the
// first n launches.  DESIGN.md section 3.1 quotes what they showed.
#include <algorithm>

#include "athena_internal.h"
#include "pipe_common.cuh"
#include "tc_common.cuh"

namespace athena {

using namespace tc;
using namespace pipe;

namespace {


// 16-byte shared-memory load that is not issued (and yields 0) when `p` is false
__device__ __forceinline__ float4 lds128_pred(uint32_t addr, bool p) {
  float4 v;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "mov.f32 %0, 0f00000000;\n\t"
      "mov.f32 %1, 0f00000000;\n\t"
      "mov.f32 %2, 0f00000000;\n\t"
      "mov.f32 %3, 0f00000000;\n\t"
      "@q ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n\t"
      "}"
      : "=&f"(v.x), "=&f"(v.y), "=&f"(v.z), "=&f"(v.w)
      : "r"(addr), "r"(static_cast<uint32_t>(p))
      : "memory");
  return v;
}


constexpr int TRACE_TILES = 32;
#define TRACE(role, slot)                                                                \
  do {                                                                                   \
    if (a.trace != nullptr && blockIdx.x == 0 && j < TRACE_TILES)                        \
      a.trace[((role) * TRACE_TILES + j) * 8 + (slot)] = clock64();                      \
  } while (0)

template <int F, int N, bool HAS_AUX>
struct GatherCfg {
  // ring stages: the kernels without a second epilogue operand spend the shared memory of
  // the operand tile on a third stage (two tiles of prefetch instead of one)
  static constexpr int NS = HAS_AUX ? 2 : 3;
  static constexpr int GATHER_THREADS = 512;                     // warps 0..15
  static constexpr int EPI_WARP0 = 16;                           // warps 16..19 (warp % 4 = lane quarter)
  static constexpr int PRODUCER_WARP = 20;
  static constexpr int MMA_WARP = 21;
  static constexpr int THREADS = 22 * 32;
  static constexpr int LPR = 4;                                  // lanes per row
  static constexpr int CPL = F / (4 * LPR);                      // 16-byte chunks per lane
  // one ring stage: feature rows | tile-local neighbour bytes | row pointers | deg^-1/2
  static constexpr int X_BYTES = TILE_ROWS * F * 4;
  static constexpr int C8_BYTES = TILE_ENTRIES + 32;
  static constexpr int RP_ELEMS = TILE_ROWS + 8;
  static constexpr int RP_BYTES = RP_ELEMS * 4;
  static constexpr int OFF_C8 = X_BYTES;
  static constexpr int OFF_RP = OFF_C8 + C8_BYTES;
  static constexpr int OFF_RS = OFF_RP + RP_BYTES;
  static constexpr int STAGE_RAW = OFF_RS + RP_BYTES;
  static constexpr int STAGE_BYTES = (STAGE_RAW + 127) / 128 * 128;
  static constexpr int KB = F / 32;
  static constexpr int A_BYTES = KB * 16384;
  static constexpr int B_BLK = 2 * N * 128;
  static constexpr int B_BYTES = KB * B_BLK;
  static constexpr int OFF_OPS = 0;                              // A hi | A lo (1024-aligned)
  static constexpr int OFF_B = 2 * A_BYTES;
  static constexpr int OFF_RING = OFF_B + B_BYTES;
  static constexpr int OFF_AUX = OFF_RING + NS * STAGE_BYTES;    // [128][N + 4] epilogue operand tile
  static constexpr int AUX_BYTES = HAS_AUX ? TILE_ROWS * (N + 4) * 4 : 0;
  static constexpr int OFF_BAR = OFF_AUX + AUX_BYTES;
  static constexpr int OFF_EPI = OFF_BAR + 256 + 512;            // 4 epilogue transposition patches
  static constexpr int SMEM = 1024 + OFF_EPI + 4 * (32 * 36) * 4;
  static_assert(SMEM <= 232448, "shared memory budget");
  static_assert(C8_BYTES % 16 == 0 && OFF_RP % 16 == 0 && OFF_RS % 16 == 0, "TMA alignment");
  static constexpr int ACC_COLS = 2 * N;                         // hi|lo stacked along N
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 64 ? 64 : 2 * ACC_COLS <= 128 ? 128
                                   : 2 * ACC_COLS <= 256 ? 256 : 512;
  static_assert(GATHER_THREADS == TILE_ROWS * LPR && CPL == 4 && F == 64, "row mapping");
  static_assert(TILE_ROWS <= 256, "tile-local neighbour index must fit one byte");
};


template <int F, int N, bool TRANSB, int EPI>
__global__ void __launch_bounds__(GatherCfg<F, N, EPI != EPI_ACT>::THREADS, 1)
k_pipe_gather(GatherArgs a) {
  using Cfg = GatherCfg<F, N, EPI != EPI_ACT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sAhi = smem + Cfg::OFF_OPS;
  uint8_t* sAlo = sAhi + Cfg::A_BYTES;
  uint8_t* sB = smem + Cfg::OFF_B;
  uint8_t* ring = smem + Cfg::OFF_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;                 // [NS]
  uint64_t* empty = bars + Cfg::NS;      // [NS]
  uint64_t* ops_ready = bars + 2 * Cfg::NS;
  uint64_t* ops_free = ops_ready + 1;
  uint64_t* acc_full = ops_free + 1;     // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* loss_red = reinterpret_cast<float*>(smem + Cfg::OFF_BAR + 256);  // [128], EPI_MSE
  float* sAux = reinterpret_cast<float*>(smem + Cfg::OFF_AUX);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == Cfg::MMA_WARP) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  if (tid == 0) {
    for (int s = 0; s < Cfg::NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], Cfg::GATHER_THREADS);
    }
    mbar_init(ops_ready, Cfg::GATHER_THREADS);
    mbar_init(ops_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
    }
    mbar_fence_init();
  }
  // stacked weight operand [hi(W') ; lo(W')] with W' = op(W) as [N][F] K-major.  One item =
  // one 16-byte chunk (4 consecutive k) of one row n; the lanes of a quarter-warp write
  // the eight chunks of one swizzled 128-byte line (conflict-free).
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
    uint8_t* blk = sB + (kc >> 3) * Cfg::B_BLK;
    *reinterpret_cast<float4*>(blk + sw128_off(n, kc & 7)) = hi;
    *reinterpret_cast<float4*>(blk + sw128_off(N + n, kc & 7)) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == Cfg::PRODUCER_WARP) {
    // ===================== producer: TMA bulk copies into the ring =====================
    if (lane == 0) {
      int j = 0;
      for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++j) {
        const int s = j % Cfg::NS;
        const uint32_t ph = (j / Cfg::NS) & 1;
        TRACE(0, 0);
        mbar_wait(&empty[s], ph ^ 1u);
        TRACE(0, 1);
        const int4 ti = __ldg(a.tiles + t);
        const int r0 = ti.x, nrows = ti.y, e0 = ti.z, nent = ti.w;
        const int ea = e0 & ~15, ebytes = (e0 + nent - ea + 15) & ~15;
        const int ra = r0 & ~3, rcnt = (r0 + nrows + 1 - ra + 3) & ~3;
        uint8_t* st = ring + s * Cfg::STAGE_BYTES;
        const uint32_t xb = nrows * F * 4, eb = ebytes, rb = rcnt * 4;
        mbar_arrive_expect_tx(&full[s], xb + eb + (a.rs ? 2 * rb : rb));
        bulk_g2s(st, a.X + static_cast<size_t>(r0) * F, xb, &full[s]);
        if (eb) bulk_g2s(st + Cfg::OFF_C8, a.col8 + ea, eb, &full[s]);
        bulk_g2s(st + Cfg::OFF_RP, a.row_ptr + ra, rb, &full[s]);
        if (a.rs) bulk_g2s(st + Cfg::OFF_RS, a.rs + ra, rb, &full[s]);
      }
    }
  } else if (warp == Cfg::MMA_WARP) {
    // ===================== MMA issuer ==================================================
    if (lane == 0) {
      const uint32_t aHi = smem_u32(sAhi), aLo = smem_u32(sAlo), bAddr = smem_u32(sB);
      constexpr uint32_t IDESC = make_idesc(128, 2 * N, false, false);
      int j = 0;
      for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++j) {
        const int b = j & 1;
        TRACE(1, 0);
        mbar_wait(ops_ready, j & 1);
        TRACE(1, 1);
        mbar_wait(&acc_empty[b], ((j >> 1) & 1) ^ 1u);
        TRACE(1, 2);
        tc_fence_after();
        const uint32_t tacc = tmem + b * Cfg::ACC_COLS;
#pragma unroll
        for (int kb = 0; kb < Cfg::KB; ++kb) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t db = make_desc(bAddr + kb * Cfg::B_BLK + kk * 32, 16, 1024);
            const uint64_t dh = make_desc(aHi + kb * 16384 + kk * 32, 16, 1024);
            const uint64_t dl = make_desc(aLo + kb * 16384 + kk * 32, 16, 1024);
            umma_tf32(tacc, dh, db, IDESC, (kb | kk) ? 1u : 0u);
            umma_tf32(tacc, dl, db, IDESC, 1u);
          }
        }
        umma_commit(ops_free);      // operand tiles may be overwritten
        umma_commit(&acc_full[b]);  // accumulator ready for the epilogue
        TRACE(1, 3);
      }
    }
  } else if (warp >= Cfg::EPI_WARP0) {
    // ===================== epilogue: TMEM -> registers -> global rows ==================
    const int q = warp - Cfg::EPI_WARP0;  // == warp % 4: TMEM lane quarter
    float* patch = reinterpret_cast<float*>(smem + Cfg::OFF_EPI) + q * EPI_PATCH;
    const bool use_mask = (EPI == EPI_ACTGRAD) && a.mask_in != nullptr;
    const bool use_aux = (EPI != EPI_ACT) && a.aux != nullptr && !use_mask;
    const int my_row = q * 32 + lane;
    float* aux_row = sAux + my_row * AUX_PITCH;
    const int step = gridDim.x;
    auto tile_at = [&](int t) { return t < a.num_tiles ? __ldg(a.tiles + t) : make_int4(0, 0, 0, 0); };
    // MSE cell size (vertices of the graph) of this thread's row, fetched one tile ahead
    auto count_at = [&](int t) {
      int c = 1;
      if (EPI == EPI_MSE && t < a.num_tiles) {
        const int4 ti = __ldg(a.tiles + t);
        if (my_row < ti.y) c = __ldg(a.vcount + ti.x + my_row);
      }
      return c;
    };
    // Second operand of the epilogue (saved activations / target): every epilogue WARP
    // prefetches its own 32 rows of the next tile with 16-byte cp.async copies (coalesced
    // global reads) into rows padded to 272 B, so that the row-per-thread reads are
    // conflict-free and no warp waits for another one before re-filling its rows.
    // (Measured alternatives: one 256-byte TMA bulk copy per padded row costs ~100+ cycles
    // of TMA issue each -- 128 per tile is more than a tile period; a single prefetch warp
    // issuing all 2048 cp.async of a tile is slower than the tile as well.)
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
      // rows past the end of a partial tile carry scale 0 and never touch the loss sum
      const bool row_valid = my_row < ti.y;
      // sign-bit words of this row: loaded before the waits so that their latency hides
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
      if (q == 0 && lane == 0) TRACE(2, 0);
      mbar_wait(&acc_full[b], (j >> 1) & 1);
      tc_fence_after();
      if (q == 0 && lane == 0) TRACE(2, 1);
      if (use_aux) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
      }
      if (q == 0 && lane == 0) TRACE(2, 2);
      const uint32_t tacc = tmem + b * Cfg::ACC_COLS;
      const float scale =
          (EPI == EPI_MSE && row_valid) ? 1.f / static_cast<float>(N * count) : 0.f;
      const int act = use_aux || use_mask || EPI == EPI_ACT ? a.act : ATHENA_ACT_NONE;
      switch (act) {
        case ATHENA_ACT_RELU:
          lsum += epilogue_tile<ATHENA_ACT_RELU, EPI, N>(tacc, q, lane, ti.y, out_tile, aux_row,
                                                         scale, patch, &acc_empty[b], (a.dbg & 2) != 0, mout, use_mask, min_w);
          break;
        case ATHENA_ACT_LEAKY_RELU:
          lsum += epilogue_tile<ATHENA_ACT_LEAKY_RELU, EPI, N>(tacc, q, lane, ti.y, out_tile,
                                                               aux_row, scale, patch,
                                                               &acc_empty[b], (a.dbg & 2) != 0, mout, use_mask, min_w);
          break;
        case ATHENA_ACT_SIGMOID:
          lsum += epilogue_tile<ATHENA_ACT_SIGMOID, EPI, N>(tacc, q, lane, ti.y, out_tile, aux_row,
                                                            scale, patch, &acc_empty[b], (a.dbg & 2) != 0, mout, use_mask, min_w);
          break;
        case ATHENA_ACT_TANH:
          lsum += epilogue_tile<ATHENA_ACT_TANH, EPI, N>(tacc, q, lane, ti.y, out_tile, aux_row,
                                                         scale, patch, &acc_empty[b], (a.dbg & 2) != 0, mout, use_mask, min_w);
          break;
        default:
          lsum += epilogue_tile<ATHENA_ACT_NONE, EPI, N>(tacc, q, lane, ti.y, out_tile, aux_row,
                                                         scale, patch, &acc_empty[b], (a.dbg & 2) != 0, mout, use_mask, min_w);
          break;
      }
      if (q == 0 && lane == 0) TRACE(2, 3);
      if (use_aux && t + step < a.num_tiles) {  // this warp is done with its operand rows
        __syncwarp();
        issue_aux(tile_at(t + step));
      }
    }
    if (EPI == EPI_MSE) loss_red[my_row] = lsum;
  } else {
    // ===================== gather warps ================================================
    // quad (4 lanes) per row; the two quads of a quarter-warp own rows r and r + 8 (equal
    // r % 8 => complementary halves of the swizzled operand line too)
    const int quad = lane >> 2, ql = lane & 3, par = quad & 1;
    const int row = (warp >> 1) * 16 + (warp & 1) * 4 + (quad >> 1) + 8 * par;
    // float offsets of this lane's four 16-byte chunks inside a feature row:
    //   step s reads chunk 4 * (s ^ par) + ql
    const int offa = par * 16 + ql * 4;        // steps 0 (and 2: + 32)
    const int offb = (par ^ 1) * 16 + ql * 4;  // steps 1 (and 3: + 32)
    int j = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++j) {
      const int s = j % Cfg::NS;
      const uint32_t ph = (j / Cfg::NS) & 1;
      const int4 ti = __ldg(a.tiles + t);
      const int r0 = ti.x, nrows = ti.y, e0 = ti.z;
      const int ea = e0 & ~15, ra = r0 & ~3;
      const uint8_t* st = ring + s * Cfg::STAGE_BYTES;
      const uint8_t* cols8 = st + Cfg::OFF_C8;
      const int32_t* rps = reinterpret_cast<const int32_t*>(st + Cfg::OFF_RP) + (r0 - ra);
      const float* rss = reinterpret_cast<const float*>(st + Cfg::OFF_RS) + (r0 - ra);
      const bool has_coef = a.rs != nullptr;
      if (tid == 0) TRACE(3, 0);
      mbar_wait(&full[s], ph);
      if (tid == 0) TRACE(3, 1);
      float4 acc[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < nrows) {
        const int beg = rps[row] - ea, end = rps[row + 1] - ea;
        const uint32_t xs = smem_u32(st);
        int e4 = beg & ~3;
        uint32_t c4 = 0;
        if (e4 < end) c4 = *reinterpret_cast<const uint32_t*>(cols8 + e4);
        // Branch-free walk over aligned groups of four entries (four neighbour bytes = one
        // LDS.32).  Entries outside [beg, end) issue no feature load (predicated off, value
        // 0) and carry coefficient 0, so every accumulator sees exactly the row's entries
        // in ascending order.  The loads of entry k + 1 are issued before the FMAs of
        // entry k.  Coefficient: deg_u^-1/2 per entry, deg_v^-1/2 once per row
        // (athena_diffstruc_extd_sub_kipf.f90:39-44 up to the rounding of the product).
        while (e4 < end && !(a.dbg & 4)) {
          const int en = e4 + 4;
          uint32_t cn = c4;
          if (en < end) cn = *reinterpret_cast<const uint32_t*>(cols8 + en);  // prefetch
          float4 x[2][4];
          float w[2];
          auto issue = [&](int k) {
            const int e = e4 + k;
            const bool v = (e >= beg) && (e < end);
            const uint32_t c = (c4 >> (8 * k)) & 0xffu;
            const uint32_t xr = xs + c * (F * 4);
            x[k & 1][0] = lds128_pred(xr + offa * 4, v);
            x[k & 1][1] = lds128_pred(xr + offb * 4, v);
            x[k & 1][2] = lds128_pred(xr + offa * 4 + 128, v);
            x[k & 1][3] = lds128_pred(xr + offb * 4 + 128, v);
            const float wu = has_coef ? rss[c & 127u] : 1.f;
            w[k & 1] = v ? wu : 0.f;
          };
          auto consume = [&](int k) {
            const float wk = w[k & 1];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 xv = x[k & 1][c];
              acc[c].x = fmaf(wk, xv.x, acc[c].x);
              acc[c].y = fmaf(wk, xv.y, acc[c].y);
              acc[c].z = fmaf(wk, xv.z, acc[c].z);
              acc[c].w = fmaf(wk, xv.w, acc[c].w);
            }
          };
          issue(0);
          issue(1);
          consume(0);
          issue(2);
          consume(1);
          issue(3);
          consume(2);
          consume(3);
          c4 = cn;
          e4 = en;
        }
        if (has_coef && end > beg) {  // an empty row sums to 0 (its deg^-1/2 is Infinity)
          const float wv = rss[row];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            acc[c].x *= wv;
            acc[c].y *= wv;
            acc[c].z *= wv;
            acc[c].w *= wv;
          }
        }
      }
      if (tid == 0) TRACE(3, 2);
      mbar_arrive(&empty[s]);  // raw stage consumed
      // chunk index of step k: 4 * (k ^ par) + ql
      const int ch0 = 4 * par + ql, ch1 = 4 * (par ^ 1) + ql;
      if (a.P != nullptr && row < nrows && !(a.dbg & 1)) {
        float* prow = a.P + (static_cast<size_t>(r0) + row) * F;
        *reinterpret_cast<float4*>(prow + ch0 * 4) = acc[0];
        *reinterpret_cast<float4*>(prow + ch1 * 4) = acc[1];
        *reinterpret_cast<float4*>(prow + ch0 * 4 + 32) = acc[2];
        *reinterpret_cast<float4*>(prow + ch1 * 4 + 32) = acc[3];
      }
      if (tid == 0) TRACE(3, 3);
      mbar_wait(ops_free, (j & 1) ^ 1u);  // MMAs of the previous tile have read the operands
      if (tid == 0) TRACE(3, 4);
      {
        const uint32_t o0 = sw128_off(row, ch0), o1 = sw128_off(row, ch1);
        float4 hi, lo;
        split_tf32_safe(acc[0], hi, lo);
        *reinterpret_cast<float4*>(sAhi + o0) = hi;
        *reinterpret_cast<float4*>(sAlo + o0) = lo;
        split_tf32_safe(acc[1], hi, lo);
        *reinterpret_cast<float4*>(sAhi + o1) = hi;
        *reinterpret_cast<float4*>(sAlo + o1) = lo;
        split_tf32_safe(acc[2], hi, lo);
        *reinterpret_cast<float4*>(sAhi + 16384 + o0) = hi;
        *reinterpret_cast<float4*>(sAlo + 16384 + o0) = lo;
        split_tf32_safe(acc[3], hi, lo);
        *reinterpret_cast<float4*>(sAhi + 16384 + o1) = hi;
        *reinterpret_cast<float4*>(sAlo + 16384 + o1) = lo;
      }
      fence_async_smem();
      mbar_arrive(ops_ready);
      if (tid == 0) TRACE(3, 5);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (EPI == EPI_MSE) {
    // fixed-order reduction of the 128 per-thread partial sums of this CTA
    if (tid == 0) {
      float tot = 0.f;
      for (int i = 0; i < 128; ++i) tot += loss_red[i];
      a.loss_part[blockIdx.x] = tot;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::MMA_WARP) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------
// dW = P^T . G, pipelined: bulk-copied raw row tiles -> split/swizzle -> MMA
// ---------------------------------------------------------------------------
// Up to TN_MAX_JOBS products dW = P^T . G of one reverse sweep (one per layer-step) run in ONE
// launch, back to back on the same pipeline: set-up, instruction-cache warm-up and tail are
// paid once.
constexpr int TN_MAX_JOBS = 4;
struct TnJobs {
  int n;
  const float* P[TN_MAX_JOBS];
  const float* G[TN_MAX_JOBS];
  float* part[TN_MAX_JOBS];   // [gridDim.x][K * N] per job
  long long M[TN_MAX_JOBS];
  int reverse;                // walk the stages of every job from the last to the first
  int evict_first;            // both operands are read for the last time: L2 evict_first
};

template <int KW, int N>
struct Tn2Cfg {
  static constexpr int K = KW;                               // 64, or 32: hi(P)^T in TMEM lanes 0..31,
                                                             // lo(P)^T in lanes 32..63, lanes 64..127 unused
  static constexpr int HI_WARPS = K / 32;                    // lane quarters of one half of the A operand
  static constexpr int RS = 64;                              // rows per stage
  static constexpr int NS = 3 * 128 / (KW + N);              // 96 KB of raw rows in flight per SM
  static constexpr int G_THREADS = 128;                      // warps 0..3: gY -> hi/lo B operand
  static constexpr int A_WARP0 = 4;                          // warps 4..7: P^T -> TMEM A operand
  static constexpr int PRODUCER_WARP = 8;
  static constexpr int MMA_WARP = 9;
  static constexpr int ACC_WARP0 = 10;                       // warps 10..13: warp % 4 = TMEM lane quarter
  static constexpr int THREADS = 14 * 32;
  // The tensor core adds into its fp32 accumulator with truncation, so a chain of m
  // accumulating MMAs carries a bias of ~m * 2^-24.  The accumulator is therefore drained
  // into registers (IEEE round-to-nearest adds) every FLUSH stages: two TMEM accumulators
  // alternate, the MMAs of group g+1 overlap the drain of group g.
  static constexpr int FLUSH = 2;
  static constexpr int P_RAW = RS * K * 4;                   // 16 KB (K = 64)
  static constexpr int G_RAW = RS * N * 4;
  static constexpr int STAGE_BYTES = P_RAW + G_RAW;
  static constexpr int BLK = RS * 128;                       // [64 rows x 32 feats]
  static constexpr int B_HALF = (N / 32) * BLK;              // hi (or lo) of gY
  static constexpr int OPS_BYTES = 2 * B_HALF;               // one B operand buffer (two exist)
  static constexpr int OFF_RING = 2 * OPS_BYTES;
  static constexpr int OFF_BAR = OFF_RING + NS * STAGE_BYTES;
  // hand-over of the lo(P)^T . G half to the warps that own the hi(P)^T . G half (one
  // partial per CTA instead of two): [K features][N + 4] floats, padded against bank conflicts
  static constexpr int COMB_PITCH = N + 4;
  static constexpr int OFF_COMB = OFF_BAR + 256;
  static constexpr int SMEM = 1024 + OFF_COMB + K * COMB_PITCH * 4;
  // TMEM: two A operand buffers [hi(P)^T ; lo(P)^T] x 64 vertices, two accumulators of 2N columns
  static constexpr uint32_t T_A = 0, T_ACC = 2 * RS;
  static constexpr int TMEM_COLS = (2 * RS + 4 * N <= 256) ? 256 : 512;
};

// All threads of the CTA meet here at the end of every product, from whatever role branch
// they are in: no role starts product q + 1 before the accumulate warps have written the
// partials of product q.  (Letting the roles run ahead across the boundary -- which the
// mbarrier protocol should allow -- gave run-to-run differences in the SECOND product of a
// launch in about half of the runs; tools/check_determinism.py.  The pipeline drain costs
// ~1 us per boundary against ~8 us for a second launch.)
template <int THREADS>
__device__ __forceinline__ void job_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}

template <int KW, int N>
__global__ void __launch_bounds__(Tn2Cfg<KW, N>::THREADS, 1)
k_pipe_tn(TnJobs jobs) {
  using Cfg = Tn2Cfg<KW, N>;
  constexpr int K = Cfg::K;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem + Cfg::OFF_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::NS;
  uint64_t* ops_ready = bars + 2 * Cfg::NS;  // [2] B operand (gY hi/lo) staged in shared memory
  uint64_t* a_ready = ops_ready + 2;         // [2] A operand (P^T hi/lo) stored in TMEM
  uint64_t* ops_free = a_ready + 2;          // [2] MMAs of the stage have read both operands
  uint64_t* acc_full = ops_free + 2;         // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == Cfg::MMA_WARP) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  if (tid == 0) {
    for (int s = 0; s < Cfg::NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], Cfg::G_THREADS + 128);
    }
    for (int ob = 0; ob < 2; ++ob) {
      mbar_init(&ops_ready[ob], Cfg::G_THREADS);
      mbar_init(&a_ready[ob], 128);
      mbar_init(&ops_free[ob], 1);
      mbar_init(&acc_full[ob], 1);
      mbar_init(&acc_empty[ob], 128);
    }
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();  // the set-up above overlapped the previous kernel's tail
  // only now may the NEXT kernel start: everything before this kernel has completed, so a
  // dependent that skips ahead never overlaps a grid older than this one
  pdl_launch_dependents();
  // stages of job q owned by this CTA; every role walks the jobs in order with running
  // counters: j over stages (ring / operand-buffer phases), gcount over accumulator groups
  auto tiles_of = [&](int q) { return (jobs.M[q] + Cfg::RS - 1) / Cfg::RS; };
  auto mine_of = [&](int q) {
    const long long nt = tiles_of(q);
    return nt > blockIdx.x ? static_cast<int>((nt - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  };

  if (warp == Cfg::PRODUCER_WARP) {
    int j = 0;
    for (int q = 0; q < jobs.n; ++q) {
      const float* P = jobs.P[q];
      const float* G = jobs.G[q];
      const long long M = jobs.M[q], ntiles = tiles_of(q);
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
        if (lane != 0) continue;
        const int s = j % Cfg::NS;
        const uint32_t ph = (j / Cfg::NS) & 1;
        mbar_wait(&empty[s], ph ^ 1u);
        const long long r0 = (jobs.reverse ? ntiles - 1 - t : t) * Cfg::RS;
        const int rows = static_cast<int>(min(static_cast<long long>(Cfg::RS), M - r0));
        uint8_t* st = ring + s * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], rows * (K + N) * 4);
        if (jobs.evict_first) {
          const uint64_t pol = l2_policy_evict_first();
          bulk_g2s_hint(st, P + r0 * K, rows * K * 4, &full[s], pol);
          bulk_g2s_hint(st + Cfg::P_RAW, G + r0 * N, rows * N * 4, &full[s], pol);
        } else {
          bulk_g2s(st, P + r0 * K, rows * K * 4, &full[s]);
          bulk_g2s(st + Cfg::P_RAW, G + r0 * N, rows * N * 4, &full[s]);
        }
      }
      __syncwarp();
      job_sync<Cfg::THREADS>();
    }
  } else if (warp == Cfg::MMA_WARP) {
    // One instruction per k-step (8 vertices): A = [hi(P)^T ; lo(P)^T] (M = 128) read from
    // TENSOR MEMORY, B = [hi(gY) | lo(gY)] (N = 2N) from shared memory -> all four partial
    // products at once.  The warp stays converged and the instruction is predicated on one
    // elected lane (see tc_common.cuh).
    const uint32_t leader = elect_one();
    constexpr uint32_t IDESC = make_idesc(128, 2 * N, false, true);
    int j = 0, gcount = 0;
    for (int q = 0; q < jobs.n; ++q) {
      const int my_tiles = mine_of(q);
      for (int jl = 0; jl < my_tiles; ++jl, ++j) {
        const int ob = j & 1;
        const int gl = jl / Cfg::FLUSH, first = jl - gl * Cfg::FLUSH;
        const int g = gcount + gl, ab = g & 1;
        const uint32_t b_hi = smem_u32(smem + ob * Cfg::OPS_BYTES);
        const uint32_t ta = tmem + Cfg::T_A + ob * Cfg::RS;
        const uint32_t tacc = tmem + Cfg::T_ACC + ab * 2 * N;
        mbar_wait(&ops_ready[ob], (j >> 1) & 1);
        mbar_wait(&a_ready[ob], (j >> 1) & 1);
        if (first == 0) mbar_wait(&acc_empty[ab], ((g >> 1) & 1) ^ 1u);  // drained by the acc warps
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < Cfg::RS / 8; ++ks)
          umma_tf32_ts_w(leader, tacc, ta + ks * 8,
                         make_desc_mn32(b_hi + ks * 1024, Cfg::BLK, 512), IDESC,
                         (first | ks) ? 1u : 0u);
        umma_commit_w(leader, &ops_free[ob]);
        if (first == Cfg::FLUSH - 1 || jl == my_tiles - 1) umma_commit_w(leader, &acc_full[ab]);
        __syncwarp();
      }
      gcount += (my_tiles + Cfg::FLUSH - 1) / Cfg::FLUSH;
      job_sync<Cfg::THREADS>();
    }
  } else if (warp >= Cfg::ACC_WARP0) {
    // accumulate warps: drain the TMEM accumulator of every finished group into registers.
    // TMEM lanes 0..63 hold hi(P)^T.G, lanes 64..127 lo(P)^T.G; the halves are added at the
    // end of the product and leave as ONE [K x N] partial per CTA; the fold over the CTAs
    // happens in the reduce / finalize kernel, in partial-index order.
    const int lq = warp & 3;
    int gcount = 0;
    for (int q = 0; q < jobs.n; ++q) {
      float acc[N];
#pragma unroll
      for (int i = 0; i < N; ++i) acc[i] = 0.f;
      const int ngroups = (mine_of(q) + Cfg::FLUSH - 1) / Cfg::FLUSH;
      for (int gl = 0; gl < ngroups; ++gl) {
        const int g = gcount + gl, ab = g & 1;
        mbar_wait(&acc_full[ab], (g >> 1) & 1);
        tc_fence_after();
        // columns n and N + n of a lane are the  . hi(G)  and  . lo(G)  products
#pragma unroll
        for (int cg = 0; cg < N / 16; ++cg) {
          float v[16], w[16];
          const uint32_t ta = tmem + Cfg::T_ACC + ab * 2 * N +
                              (static_cast<uint32_t>(lq * 32) << 16) + cg * 16;
          tmem_ld16_nowait(ta, v);
          tmem_ld16_nowait(ta + N, w);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[cg * 16 + i] += v[i] + w[i];
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[ab]);
      }
      gcount += ngroups;
      // the two halves (TMEM lanes 0..63: hi(P)^T . G, lanes 64..127: lo(P)^T . G) are added
      // here, in a fixed order, so that the CTA writes ONE partial
      // (K = 32: quarters 2 and 3 hold nothing and only keep the barriers' head counts)
      const int feat = (lq % Cfg::HI_WARPS) * 32 + lane;
      float* comb = reinterpret_cast<float*>(smem + Cfg::OFF_COMB) + feat * Cfg::COMB_PITCH;
      if (lq >= Cfg::HI_WARPS && lq < 2 * Cfg::HI_WARPS) {
#pragma unroll
        for (int i = 0; i < N; i += 4)
          *reinterpret_cast<float4*>(comb + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");  // the four accumulate warps
      if (lq < Cfg::HI_WARPS) {
        float* dst = jobs.part[q] + static_cast<size_t>(blockIdx.x) * K * N + feat * N;
#pragma unroll
        for (int i = 0; i < N; i += 4) {
          const float4 lo = *reinterpret_cast<const float4*>(comb + i);
          *reinterpret_cast<float4*>(dst + i) =
              make_float4(acc[i] + lo.x, acc[i + 1] + lo.y, acc[i + 2] + lo.z, acc[i + 3] + lo.w);
        }
      }
      job_sync<Cfg::THREADS>();
    }
  } else if (warp >= Cfg::A_WARP0) {
    // A-operand warps: thread = TMEM lane = one feature of hi(P)^T (lanes 0..63) or lo(P)^T
    // (lanes 64..127).  It reads its feature of the stage's 64 vertices with LDS.32 (a warp
    // reads 32 consecutive floats of one row: conflict-free) -- the transposition is free --
    // and stores them as 64 TMEM columns: the A operand never touches shared memory again
    // (no operand stores, no operand reads by the MMA: -25 % shared-memory traffic per stage).
    const int lq = warp & 3;
    const int f = (lq % Cfg::HI_WARPS) * 32 + lane;
    const bool lo_part = lq >= Cfg::HI_WARPS;
    const bool a_live = lq < 2 * Cfg::HI_WARPS;  // K = 32: TMEM lanes 64..127 stay unwritten
    int j = 0;
    for (int q = 0; q < jobs.n; ++q) {
      const long long M = jobs.M[q], ntiles = tiles_of(q);
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
        const int s = j % Cfg::NS;
        const uint32_t ph = (j / Cfg::NS) & 1;
        const long long r0 = (jobs.reverse ? ntiles - 1 - t : t) * Cfg::RS;
        const int rows = static_cast<int>(min(static_cast<long long>(Cfg::RS), M - r0));
        const float* raw = reinterpret_cast<const float*>(ring + s * Cfg::STAGE_BYTES);
        mbar_wait(&full[s], ph);
        float x[Cfg::RS];
#pragma unroll
        for (int r = 0; r < Cfg::RS; ++r) x[r] = (a_live && r < rows) ? raw[r * K + f] : 0.f;
        mbar_arrive(&empty[s]);
        const int ob = j & 1;
        mbar_wait(&ops_free[ob], ((j >> 1) & 1) ^ 1u);  // MMAs of stage j-2 have read this buffer
        tc_fence_after();
        const uint32_t ta =
            tmem + Cfg::T_A + ob * Cfg::RS + (static_cast<uint32_t>(lq * 32) << 16);
        if (a_live) {
#pragma unroll
          for (int h = 0; h < Cfg::RS / 32; ++h) {
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float xx = x[h * 32 + i];
              const uint32_t hi = __float_as_uint(xx) & 0xffffe000u;
              v[i] = lo_part ? __float_as_uint(xx - __uint_as_float(hi)) : hi;
            }
            tmem_st32(ta + h * 32, v);
          }
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(&a_ready[ob]);
      }
      job_sync<Cfg::THREADS>();
    }
  } else {
    // B-operand warps: raw row-major gY -> hi/lo, MN-major 32B-base swizzle
    // (two operand buffers: the stores of stage j+1 overlap the MMAs of stage j)
    constexpr int CH2 = N / 4;
    constexpr int LOADS = Cfg::RS * CH2 / Cfg::G_THREADS;
    static_assert(Cfg::RS * CH2 % Cfg::G_THREADS == 0, "loader mapping");
    int j = 0;
    for (int q = 0; q < jobs.n; ++q) {
      const long long M = jobs.M[q], ntiles = tiles_of(q);
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
        const int s = j % Cfg::NS;
        const uint32_t ph = (j / Cfg::NS) & 1;
        const long long r0 = (jobs.reverse ? ntiles - 1 - t : t) * Cfg::RS;
        const int rows = static_cast<int>(min(static_cast<long long>(Cfg::RS), M - r0));
        const uint8_t* st = ring + s * Cfg::STAGE_BYTES + Cfg::P_RAW;
        mbar_wait(&full[s], ph);
        float4 x[LOADS];
#pragma unroll
        for (int i = 0; i < LOADS; ++i) {
          const int idx = tid + Cfg::G_THREADS * i;
          const int row = idx / CH2;
          x[i] = row < rows ? *reinterpret_cast<const float4*>(st + idx * 16)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_arrive(&empty[s]);
        const int ob = j & 1;
        uint8_t* sBhi = smem + ob * Cfg::OPS_BYTES;
        uint8_t* sBlo = sBhi + Cfg::B_HALF;
        mbar_wait(&ops_free[ob], ((j >> 1) & 1) ^ 1u);  // MMAs of stage j-2 have read this buffer
#pragma unroll
        for (int i = 0; i < LOADS; ++i) {
          const int idx = tid + Cfg::G_THREADS * i;
          const int row = idx / CH2, ch = idx - row * CH2;
          float4 hi, lo;
          split_tf32(x[i], hi, lo);
          const uint32_t off = (ch >> 3) * Cfg::BLK + sw128b32_off(row, ch & 7);
          *reinterpret_cast<float4*>(sBhi + off) = hi;
          *reinterpret_cast<float4*>(sBlo + off) = lo;
        }
        fence_async_smem();
        mbar_arrive(&ops_ready[ob]);
      }
      job_sync<Cfg::THREADS>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::MMA_WARP) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

__global__ void k_pipe_tn_reduce(const float* __restrict__ part, int nparts, int KN,
                                 float* __restrict__ dW) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= KN) return;
  float s = 0.f;
  for (int c = 0; c < nparts; ++c) s += part[static_cast<size_t>(c) * KN + e];
  dW[e] += s;
}

bool pipe_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* s = getenv("ATHENA_CUDA_DISABLE_PIPE");
    const char* t = getenv("ATHENA_CUDA_DISABLE_TC");
    on = ((s && atoi(s) != 0) || (t && atoi(t) != 0)) ? 0 : 1;
  }
  return on == 1;
}

template <int F, int N, bool TRANSB, int EPI>
int launch_gather_t(const GatherArgs& a) {
  using Cfg = GatherCfg<F, N, EPI != EPI_ACT>;
  static bool attr = false;
  if (!attr) {
    ATH_CUDA(cudaFuncSetAttribute(k_pipe_gather<F, N, TRANSB, EPI>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  const int grid = std::min(a.num_tiles, ctx().sm_count);
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("ATHENA_DEBUG_PIPE");
    dbg = e ? atoi(e) : 0;
  }
  GatherArgs b = a;
  b.dbg = dbg;
  b.trace = nullptr;
  static int trace_left = -1;
  static long long* trace_buf = nullptr;
  if (trace_left < 0) {
    const char* e = getenv("ATHENA_DEBUG_TRACE");
    trace_left = e ? atoi(e) : 0;
  }
  const size_t trace_n = (size_t)4 * TRACE_TILES * 8;
  if (trace_left > 0) {
    if (!trace_buf) cudaMalloc(&trace_buf, trace_n * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, trace_n * sizeof(long long), ctx().stream);
    b.trace = trace_buf;
  }
  k_pipe_gather<F, N, TRANSB, EPI><<<grid, Cfg::THREADS, Cfg::SMEM, ctx().stream>>>(b);
  if (trace_left > 0) {
    --trace_left;
    std::vector<long long> h(trace_n);
    cudaStreamSynchronize(ctx().stream);
    cudaMemcpy(h.data(), trace_buf, trace_n * sizeof(long long), cudaMemcpyDeviceToHost);
    long long t0 = h[(3 * TRACE_TILES + 0) * 8 + 0];
    fprintf(stderr, "TRACE kernel EPI=%d TRANSB=%d\n", EPI, (int)TRANSB);
    for (int role = 0; role < 4; ++role)
      for (int j = 0; j < 16; ++j) {
        fprintf(stderr, "TRACE role %d tile %2d:", role, j);
        for (int k = 0; k < 6; ++k) {
          long long v = h[(role * TRACE_TILES + j) * 8 + k];
          fprintf(stderr, " %7lld", v ? v - t0 : -1);
        }
        fprintf(stderr, "\n");
      }
  }
  ATH_LAUNCHED_T(EPI == EPI_ACT ? "pipe_gather_fwd"
                 : EPI == EPI_MSE ? "pipe_gather_fwd_mse" : "pipe_gather_bwd");
  return ATHENA_OK;
}

}  // namespace

bool pipe_gather_supported(const Batch* b, int F, int N) {
  if (!pipe_enabled() || b->force_list || b->num_tiles == 0 || F != N) return false;
  if (F == 64) return true;  // tensor-core gather, or the list gather for multi-edge batches
  // width 32: the tensor-core gather only, and only for batches that fill the machine: below
  // two tiles per SM the FP32 tile kernels (tile_fma.cu) are as fast (latency-bound) and an
  // order of magnitude closer to the exact result (1e-7 instead of the 1e-6 of truncating
  // tensor-core accumulation), which ill-conditioned optimiser steps (classical-L2 Adam)
  // amplify
  return F == 32 && b->num_tiles >= 2 * ctx().sm_count &&
         pipe_tcg_supported(const_cast<Batch*>(b), F, N);
}

bool pipe_tn_supported(int K, int N) {
  return pipe_enabled() && (K == 64 || K == 32) && (N == 64 || N == 32);
}

// forward: out = act( (A_hat X) W ),  P = A_hat X           (W row-major [F][N])
int launch_pipe_gather_fwd(const Batch* b, const float* X, const float* W, float* P, float* out,
                           int F, int N, int act, uint32_t* mask_out) {
  GatherArgs a{};
  a.tiles = b->tiles.as<int4>();
  a.num_tiles = b->num_tiles;
  a.row_ptr = b->row_ptr;
  a.col8 = b->col8;
  a.rs = b->rsdeg;
  a.X = X;
  a.W = W;
  a.P = P;
  a.out = out;
  a.aux = nullptr;
  a.act = act;
  a.mask_out = mask_out;
  ATH_REQUIRE(F == N && (F == 64 || F == 32), ATHENA_ERR_ARG, "pipe_gather_fwd: unsupported shape");
  if (pipe_tcg_supported(const_cast<Batch*>(b), F, N)) {
    a.abits = b->abits;
    a.num_rows = b->V;
    a.nonfinite = b->status.as<int32_t>() + 3;
    const_cast<Batch*>(b)->tcg_forwards += 1;
    return launch_pipe_tcg(a, false, EPI_ACT, F);
  }
  ATH_REQUIRE(F == 64, ATHENA_ERR_ARG, "pipe_gather_fwd: width %d needs the tensor-core gather", F);
  return launch_gather_t<64, 64, false, EPI_ACT>(a);
}

// last layer of a training step: forward fused with the graph-output MSE.
//   grad[v,:]  = d loss / d pre-activation = act'(p) .* (p - target) / (N nv_s),  p = act(P W)
//   loss_part[c] = sum over the rows of CTA c of (p - target)^2 / (N nv_s)   (c < *num_parts)
int launch_pipe_gather_fwd_mse(const Batch* b, const float* X, const float* W, float* P,
                               const float* target, float* grad, int F, int N, int act,
                               float* loss_part, int* num_parts) {
  GatherArgs a{};
  a.tiles = b->tiles.as<int4>();
  a.num_tiles = b->num_tiles;
  a.row_ptr = b->row_ptr;
  a.col8 = b->col8;
  a.rs = b->rsdeg;
  a.X = X;
  a.W = W;
  a.P = P;
  a.out = grad;
  a.aux = target;
  a.act = act;
  a.vcount = b->vcount;
  a.loss_part = loss_part;
  ATH_REQUIRE(F == N && (F == 64 || F == 32), ATHENA_ERR_ARG,
              "pipe_gather_fwd_mse: unsupported shape");
  *num_parts = std::min(b->num_tiles, ctx().sm_count);
  if (pipe_tcg_supported(const_cast<Batch*>(b), F, N)) {
    a.abits = b->abits;
    a.num_rows = b->V;
    return launch_pipe_tcg(a, false, EPI_MSE, F);
  }
  ATH_REQUIRE(F == 64, ATHENA_ERR_ARG, "pipe_gather_fwd_mse: width %d needs the tensor-core gather", F);
  return launch_gather_t<64, 64, false, EPI_MSE>(a);
}

// backward: out = ( (A^T-gather of G) W^T ) .* act'(Hin)      (W row-major [N][F]: W_t as stored)
int launch_pipe_gather_bwd(const Batch* b, const float* G, const float* W, const float* Hin,
                           float* out, int F, int N, int act, const uint32_t* mask_in) {
  GatherArgs a{};
  a.tiles = b->tiles.as<int4>();
  a.num_tiles = b->num_tiles;
  a.row_ptr = b->csc_ptr;
  a.col8 = b->csc8;
  a.rs = nullptr;
  a.X = G;
  a.W = W;
  a.P = nullptr;
  a.out = out;
  a.aux = (act != ATHENA_ACT_NONE && act != ATHENA_ACT_LINEAR) ? Hin : nullptr;
  a.mask_in = (act == ATHENA_ACT_RELU || act == ATHENA_ACT_LEAKY_RELU) ? mask_in : nullptr;
  a.act = act;
  ATH_REQUIRE(F == N && (F == 64 || F == 32), ATHENA_ERR_ARG, "pipe_gather_bwd: unsupported shape");
  if (pipe_tcg_supported(const_cast<Batch*>(b), F, N)) {
    a.abits = b->atbits;
    a.num_rows = b->V;
    return launch_pipe_tcg(a, true, EPI_ACTGRAD, F);
  }
  ATH_REQUIRE(F == 64, ATHENA_ERR_ARG, "pipe_gather_bwd: width %d needs the tensor-core gather", F);
  return launch_gather_t<64, 64, true, EPI_ACTGRAD>(a);
}

template <int K, int N>
static int launch_pipe_tn_t(const TnPending* jobs, int njobs, DeferList* defer) {
  using Cfg = Tn2Cfg<K, N>;
  static bool attr = false;
  if (!attr) {
    ATH_CUDA(cudaFuncSetAttribute(k_pipe_tn<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM));
    attr = true;
  }
  int64_t max_tiles = 0;
  for (int q = 0; q < njobs; ++q) max_tiles = std::max(max_tiles, cdiv(jobs[q].M, Cfg::RS));
  const int grid = (int)std::min<int64_t>(max_tiles, (int64_t)ctx().sm_count);
  TnJobs tj{};
  tj.n = njobs;
  // the product queued LAST reads the gradient the reverse sweep has just written: it runs
  // first, and the stages are walked against the direction of the kernel that wrote them
  tj.reverse = ctx().tile_reverse ? 1 : 0;
  tj.evict_first = (l2_hint_mask() & 16) ? 1 : 0;
  for (int q = 0; q < njobs; ++q) {
    const int src = njobs - 1 - q;
    ATH_TRY(jobs[src].scratch->reserve(sizeof(float) * (size_t)grid * Cfg::K * N));
    tj.P[q] = jobs[src].P;
    tj.G[q] = jobs[src].G;
    tj.part[q] = jobs[src].scratch->template as<float>();
    tj.M[q] = jobs[src].M;
  }
  ATH_CUDA(launch_pdl(k_pipe_tn<K, N>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM, ctx().stream,
                      tj));
  ATH_LAUNCHED_T("pipe_tn");
  for (int q = 0; q < njobs; ++q) {
    float* part = jobs[q].scratch->template as<float>();
    if (defer) {
      defer->jobs.push_back(DeferJob{part, grid, Cfg::K * N, jobs[q].dW});
      continue;
    }
    k_pipe_tn_reduce<<<(unsigned)cdiv((int64_t)Cfg::K * N, 128), 128, 0, ctx().stream>>>(
        part, grid, Cfg::K * N, jobs[q].dW);
    ATH_LAUNCHED_T("pipe_tn_reduce");
  }
  return ATHENA_OK;
}

static int launch_pipe_tn_shape(int K, int N, const TnPending* jobs, int njobs, DeferList* defer) {
  if (K == 64) {
    return N == 64 ? launch_pipe_tn_t<64, 64>(jobs, njobs, defer)
                   : launch_pipe_tn_t<64, 32>(jobs, njobs, defer);
  }
  return N == 64 ? launch_pipe_tn_t<32, 64>(jobs, njobs, defer)
                 : launch_pipe_tn_t<32, 32>(jobs, njobs, defer);
}

// dW[K x N] += P^T . G     (P [M][K], G [M][N], both dense row-major; K, N in {32, 64})
// With a DeferList the product is only queued: launch_pipe_tn_pending runs every queued product
// of the reverse sweep in one launch per shape (before launch_finalize folds the partials).
// queue: P and G stay untouched until the end of the sweep, so the product itself may wait;
// otherwise only the fold of its partials is deferred.
int launch_pipe_tn(const float* P, const float* G, float* dW, int64_t M, int K, int N,
                   DevBuf& scratch, DeferList* defer, bool queue) {
  if (M == 0) return ATHENA_OK;
  ATH_REQUIRE((K == 64 || K == 32) && (N == 64 || N == 32), ATHENA_ERR_ARG,
              "pipe_tn: unsupported shape %d x %d", K, N);
  const TnPending job{P, G, dW, M, K, N, &scratch};
  static int nobatch = -1;
  if (nobatch < 0) {
    const char* e = getenv("ATHENA_DEBUG_TN_NOBATCH");  // A/B switch: one launch per product
    nobatch = e ? atoi(e) : 0;  // 1: launch at once; 2: queue, but one launch per product
  }
  if (defer && queue && nobatch != 1) {
    defer->tn.push_back(job);
    return ATHENA_OK;
  }
  return launch_pipe_tn_shape(K, N, &job, 1, defer);
}

int launch_pipe_tn_pending(DeferList* defer) {
  for (int shape = 0; shape < 4; ++shape) {
    const int K = (shape & 2) ? 32 : 64, width = (shape & 1) ? 32 : 64;
    std::vector<TnPending> batch;
    for (const TnPending& j : defer->tn)
      if (j.K == K && j.N == width) batch.push_back(j);
    const char* e = getenv("ATHENA_DEBUG_TN_NOBATCH");
    const size_t per_launch = (e && atoi(e) == 2) ? 1 : TN_MAX_JOBS;
    for (size_t i = 0; i < batch.size(); i += per_launch) {
      const int n = (int)std::min<size_t>(per_launch, batch.size() - i);
      ATH_TRY(launch_pipe_tn_shape(K, width, batch.data() + i, n, defer));
    }
  }
  defer->tn.clear();
  return ATHENA_OK;
}

}  // namespace athena
