// Fused, warp-specialised tcgen05 kernels for mini-batches of small graphs
// (every graph fits one 128-vertex tile; Batch::tiles).
//
//   k_pipe_gather<F,N,TRANSB,EPI>
//       out[v,:] = epi( ( sum_{w in row v} c_w * X[col[w],:] ) . op(W) )
//     forward  (CSR, c = (deg_v deg_u)^-1/2, W_t,   epi = activation):
//         kipf_propagate + matmul + activation%apply in ONE pass
//         (athena_diffstruc_extd_sub_kipf.f90:29-46, athena_kipf_msgpass_layer.f90:943-952);
//         also stores the propagated tile P for the backward pass.
//     backward (CSC, c = 1,                  W_t^T, epi = .* act'(H_{t-1})):
//         gY_{t-1} = ( sum_{v->u} gY_t[v] ) W_t^T .* act'(H_{t-1}); the un-normalised
//         scatter of get_partial_kipf_propagate_left_val (:101-109) commutes with the
//         linear map, so dP is never materialised.
//   k_pipe_tn<N>
//       dW[64 x N] += P^T . gY over all vertices (matmul partial w.r.t. W_t)
//
// Roles (one CTA per SM, persistent over tiles):
//   producer warp   1 lane issues TMA bulk copies (cp.async.bulk) of the tile's feature
//                   rows, CSR row pointers, column indices and coefficients into a
//                   shared-memory ring; completion is counted on "full" mbarriers
//   gather warps    256 threads: walk each row's entries in ascending order (the
//                   reference's summation order), gathering neighbour rows from the
//                   staged tile; split the result hi/lo and store it as the swizzled
//                   K-major A operand; store P
//   MMA warp        1 lane issues tcgen05.mma.kind::tf32 into a double-buffered TMEM
//                   accumulator and commits to mbarriers
//   epilogue warps  128 threads: tcgen05.ld, hi+lo, activation (or act'), row stores
// so the HBM stream, the shared-memory gather, the tensor pipe and the stores of
// consecutive tiles overlap.  No float atomics; fixed tile -> CTA mapping.
#include <algorithm>

#include "athena_internal.h"
#include "tc_common.cuh"

namespace athena {

using namespace tc;

namespace {

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

struct GatherArgs {
  const int4* tiles;
  int num_tiles;
  const int32_t* row_ptr;
  const int32_t* col;
  const float* coef;  // nullptr -> unit coefficients
  const float* X;     // [V][F]
  const float* W;
  float* P;           // optional [V][F]
  float* out;         // [V][N]
  const float* Hin;   // EPI_ACTGRAD: [V][N]
  int act;
};

template <int F, int N>
struct GatherCfg {
  static constexpr int NS = 2;                                   // ring stages
  static constexpr int GATHER_THREADS = 512;                     // warps 0..15
  static constexpr int EPI_WARP0 = 16;                           // warps 16..19 (warp % 4 = lane quarter)
  static constexpr int PRODUCER_WARP = 20;
  static constexpr int MMA_WARP = 21;
  static constexpr int THREADS = 22 * 32;
  static constexpr int LPR = F / 4;                              // lanes per row
  static constexpr int GROUPS = GATHER_THREADS / LPR;
  static constexpr int RPG = TILE_ROWS / GROUPS;                 // rows per group
  static constexpr int X_BYTES = TILE_ROWS * F * 4;
  static constexpr int IDX_ELEMS = TILE_ENTRIES + 8;
  static constexpr int IDX_BYTES = IDX_ELEMS * 4;
  static constexpr int RP_ELEMS = TILE_ROWS + 8;
  static constexpr int RP_BYTES = RP_ELEMS * 4;
  static constexpr int STAGE_RAW = X_BYTES + 2 * IDX_BYTES + RP_BYTES;
  static constexpr int STAGE_BYTES = (STAGE_RAW + 127) / 128 * 128;
  static constexpr int KB = F / 32;
  static constexpr int A_BYTES = KB * 16384;
  static constexpr int B_BLK = 2 * N * 128;
  static constexpr int B_BYTES = KB * B_BLK;
  static constexpr int OFF_OPS = 0;                              // A hi | A lo (1024-aligned)
  static constexpr int OFF_B = 2 * A_BYTES;
  static constexpr int OFF_RING = OFF_B + B_BYTES;
  static constexpr int OFF_BAR = OFF_RING + NS * STAGE_BYTES;
  static constexpr int SMEM = 1024 + OFF_BAR + 256;
  static constexpr int ACC_COLS = 2 * N;                         // hi|lo stacked along N
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 64 ? 64 : 2 * ACC_COLS <= 128 ? 128
                                   : 2 * ACC_COLS <= 256 ? 256 : 512;
  static_assert(TILE_ROWS % GROUPS == 0, "row mapping");
};

template <int ACT, int EPI, int N>
__device__ __forceinline__ void epilogue_rows(uint32_t tacc, int q, int lane, bool valid,
                                              float* __restrict__ orow,
                                              const float* __restrict__ hrow, uint64_t* acc_empty) {
#pragma unroll
  for (int cg = 0; cg < N / 16; ++cg) {
    float vh[16], vl[16];
    const uint32_t taddr = tacc + (static_cast<uint32_t>(q * 32) << 16) + cg * 16;
    tmem_ld16(taddr, vh);
    tmem_ld16(taddr + N, vl);
    if (cg == N / 16 - 1) {
      tc_fence_before();
      mbar_arrive(acc_empty);  // the accumulator buffer may be overwritten now
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 o;
        if (EPI == EPI_ACT) {
          o = make_float4(act_fwd<ACT>(vh[i] + vl[i]), act_fwd<ACT>(vh[i + 1] + vl[i + 1]),
                          act_fwd<ACT>(vh[i + 2] + vl[i + 2]), act_fwd<ACT>(vh[i + 3] + vl[i + 3]));
        } else if (ACT == ATHENA_ACT_NONE) {
          o = make_float4(vh[i] + vl[i], vh[i + 1] + vl[i + 1], vh[i + 2] + vl[i + 2],
                          vh[i + 3] + vl[i + 3]);
        } else {
          const float4 h = __ldg(reinterpret_cast<const float4*>(hrow + cg * 16 + i));
          o = make_float4(act_bwd<ACT>(h.x, vh[i] + vl[i]), act_bwd<ACT>(h.y, vh[i + 1] + vl[i + 1]),
                          act_bwd<ACT>(h.z, vh[i + 2] + vl[i + 2]),
                          act_bwd<ACT>(h.w, vh[i + 3] + vl[i + 3]));
        }
        *reinterpret_cast<float4*>(orow + cg * 16 + i) = o;
      }
    }
  }
}

template <int F, int N, bool TRANSB, int EPI>
__global__ void __launch_bounds__(GatherCfg<F, N>::THREADS, 1) k_pipe_gather(GatherArgs a) {
  using Cfg = GatherCfg<F, N>;
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
  // stacked weight operand [hi(W') ; lo(W')] with W' = op(W) as [N][F] K-major
  for (int idx = tid; idx < F * N; idx += Cfg::THREADS) {
    int k, n;
    if (!TRANSB) { k = idx / N; n = idx - k * N; } else { n = idx / F; k = idx - n * F; }
    const float w = __ldg(a.W + idx);
    const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    const float lo = w - hi;
    uint8_t* blk = sB + (k >> 5) * Cfg::B_BLK;
    const int c = k & 31;
    *reinterpret_cast<float*>(blk + sw128_off(n, c >> 2) + (c & 3) * 4) = hi;
    *reinterpret_cast<float*>(blk + sw128_off(N + n, c >> 2) + (c & 3) * 4) = lo;
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
        mbar_wait(&empty[s], ph ^ 1u);
        const int4 ti = __ldg(a.tiles + t);
        const int r0 = ti.x, nrows = ti.y, e0 = ti.z, nent = ti.w;
        const int ea = e0 & ~3, ecnt = (e0 + nent - ea + 3) & ~3;
        const int ra = r0 & ~3, rcnt = (r0 + nrows + 1 - ra + 3) & ~3;
        uint8_t* st = ring + s * Cfg::STAGE_BYTES;
        const uint32_t xb = nrows * F * 4, eb = ecnt * 4, rb = rcnt * 4;
        mbar_arrive_expect_tx(&full[s], xb + rb + (a.coef ? 2 * eb : eb));
        bulk_g2s(st, a.X + static_cast<size_t>(r0) * F, xb, &full[s]);
        if (eb) {
          bulk_g2s(st + Cfg::X_BYTES, a.col + ea, eb, &full[s]);
          if (a.coef) bulk_g2s(st + Cfg::X_BYTES + Cfg::IDX_BYTES, a.coef + ea, eb, &full[s]);
        }
        bulk_g2s(st + Cfg::X_BYTES + 2 * Cfg::IDX_BYTES, a.row_ptr + ra, rb, &full[s]);
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
        mbar_wait(ops_ready, j & 1);
        mbar_wait(&acc_empty[b], ((j >> 1) & 1) ^ 1u);
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
      }
    }
  } else if (warp >= Cfg::EPI_WARP0) {
    // ===================== epilogue: TMEM -> registers -> global rows ==================
    const int q = warp - Cfg::EPI_WARP0;  // == warp % 4: TMEM lane quarter
    int j = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++j) {
      const int b = j & 1;
      const int4 ti = __ldg(a.tiles + t);
      const int row = q * 32 + lane;
      const bool valid = row < ti.y;
      const size_t grow = static_cast<size_t>(ti.x) + row;
      float* orow = a.out + grow * N;
      const float* hrow = (EPI == EPI_ACTGRAD) ? a.Hin + grow * N : nullptr;
      mbar_wait(&acc_full[b], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem + b * Cfg::ACC_COLS;
      switch (a.act) {
        case ATHENA_ACT_RELU:
          epilogue_rows<ATHENA_ACT_RELU, EPI, N>(tacc, q, lane, valid, orow, hrow, &acc_empty[b]);
          break;
        case ATHENA_ACT_LEAKY_RELU:
          epilogue_rows<ATHENA_ACT_LEAKY_RELU, EPI, N>(tacc, q, lane, valid, orow, hrow,
                                                       &acc_empty[b]);
          break;
        case ATHENA_ACT_SIGMOID:
          epilogue_rows<ATHENA_ACT_SIGMOID, EPI, N>(tacc, q, lane, valid, orow, hrow,
                                                    &acc_empty[b]);
          break;
        case ATHENA_ACT_TANH:
          epilogue_rows<ATHENA_ACT_TANH, EPI, N>(tacc, q, lane, valid, orow, hrow, &acc_empty[b]);
          break;
        default:
          epilogue_rows<ATHENA_ACT_NONE, EPI, N>(tacc, q, lane, valid, orow, hrow, &acc_empty[b]);
          break;
      }
    }
  } else {
    // ===================== gather warps ================================================
    constexpr int LPR = Cfg::LPR;
    const int g = tid / LPR, l = tid % LPR;
    int j = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++j) {
      const int s = j % Cfg::NS;
      const uint32_t ph = (j / Cfg::NS) & 1;
      const int4 ti = __ldg(a.tiles + t);
      const int r0 = ti.x, nrows = ti.y, e0 = ti.z;
      const int ea = e0 & ~3, ra = r0 & ~3;
      const uint8_t* st = ring + s * Cfg::STAGE_BYTES;
      const float* Xs = reinterpret_cast<const float*>(st);
      const int32_t* cols = reinterpret_cast<const int32_t*>(st + Cfg::X_BYTES);
      const float* coefs = reinterpret_cast<const float*>(st + Cfg::X_BYTES + Cfg::IDX_BYTES);
      const int32_t* rps =
          reinterpret_cast<const int32_t*>(st + Cfg::X_BYTES + 2 * Cfg::IDX_BYTES) + (r0 - ra);
      const bool has_coef = a.coef != nullptr;
      mbar_wait(&full[s], ph);
      float4 acc[Cfg::RPG];
#pragma unroll
      for (int k = 0; k < Cfg::RPG; ++k) {
        const int row = g + Cfg::GROUPS * k;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < nrows) {
          const int beg = rps[row] - ea, end = rps[row + 1] - ea;
#pragma unroll 4
          for (int e = beg; e < end; ++e) {
            const int c = cols[e] - r0;
            const float w = has_coef ? coefs[e] : 1.f;
            const float4 x = *reinterpret_cast<const float4*>(Xs + c * F + l * 4);
            r.x = fmaf(w, x.x, r.x);
            r.y = fmaf(w, x.y, r.y);
            r.z = fmaf(w, x.z, r.z);
            r.w = fmaf(w, x.w, r.w);
          }
        }
        acc[k] = r;
      }
      mbar_arrive(&empty[s]);  // raw stage consumed
      if (a.P != nullptr) {
#pragma unroll
        for (int k = 0; k < Cfg::RPG; ++k) {
          const int row = g + Cfg::GROUPS * k;
          if (row < nrows)
            *reinterpret_cast<float4*>(a.P + (static_cast<size_t>(r0) + row) * F + l * 4) = acc[k];
        }
      }
      mbar_wait(ops_free, (j & 1) ^ 1u);  // MMAs of the previous tile have read the operands
#pragma unroll
      for (int k = 0; k < Cfg::RPG; ++k) {
        const int row = g + Cfg::GROUPS * k;
        float4 hi, lo;
        split_tf32(acc[k], hi, lo);
        const uint32_t off = (l >> 3) * 16384 + sw128_off(row, l & 7);
        *reinterpret_cast<float4*>(sAhi + off) = hi;
        *reinterpret_cast<float4*>(sAlo + off) = lo;
      }
      fence_async_smem();
      mbar_arrive(ops_ready);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::MMA_WARP) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------
// dW = P^T . G, pipelined: bulk-copied raw row tiles -> split/swizzle -> MMA
// ---------------------------------------------------------------------------
template <int N>
struct Tn2Cfg {
  static constexpr int K = 64;
  static constexpr int RS = 64;                              // rows per stage
  static constexpr int NS = 3;
  static constexpr int XFORM_THREADS = 256;
  static constexpr int PRODUCER_WARP = 8;
  static constexpr int MMA_WARP = 9;
  static constexpr int THREADS = 10 * 32;
  static constexpr int P_RAW = RS * K * 4;                   // 16 KB
  static constexpr int G_RAW = RS * N * 4;
  static constexpr int STAGE_BYTES = P_RAW + G_RAW;
  static constexpr int BLK = RS * 128;                       // [64 rows x 32 feats]
  static constexpr int A1_BYTES = 2 * (K / 32) * BLK;
  static constexpr int A2_HALF = (N / 32) * BLK;
  static constexpr int OPS_BYTES = A1_BYTES + 2 * A2_HALF;   // one operand buffer (two exist)
  static constexpr int OFF_RING = 2 * OPS_BYTES;
  static constexpr int OFF_BAR = OFF_RING + NS * STAGE_BYTES;
  static constexpr int SMEM = 1024 + OFF_BAR + 256;
  static constexpr int TMEM_COLS = (N <= 32 ? 32 : 64);
};

template <int N>
__global__ void __launch_bounds__(Tn2Cfg<N>::THREADS, 1)
k_pipe_tn(const float* __restrict__ P, const float* __restrict__ G, float* __restrict__ part,
          long long M) {
  using Cfg = Tn2Cfg<N>;
  constexpr int K = Cfg::K;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem + Cfg::OFF_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::NS;
  uint64_t* ops_ready = bars + 2 * Cfg::NS;  // [2]
  uint64_t* ops_free = ops_ready + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ops_free + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == Cfg::MMA_WARP) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  if (tid == 0) {
    for (int s = 0; s < Cfg::NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], Cfg::XFORM_THREADS);
    }
    for (int ob = 0; ob < 2; ++ob) {
      mbar_init(&ops_ready[ob], Cfg::XFORM_THREADS);
      mbar_init(&ops_free[ob], 1);
    }
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long ntiles = (M + Cfg::RS - 1) / Cfg::RS;
  int my_tiles = 0;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) ++my_tiles;

  if (warp == Cfg::PRODUCER_WARP) {
    if (lane == 0) {
      int j = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
        const int s = j % Cfg::NS;
        const uint32_t ph = (j / Cfg::NS) & 1;
        mbar_wait(&empty[s], ph ^ 1u);
        const long long r0 = t * Cfg::RS;
        const int rows = static_cast<int>(min(static_cast<long long>(Cfg::RS), M - r0));
        uint8_t* st = ring + s * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], rows * (K + N) * 4);
        bulk_g2s(st, P + r0 * K, rows * K * 4, &full[s]);
        bulk_g2s(st + Cfg::P_RAW, G + r0 * N, rows * N * 4, &full[s]);
      }
    }
  } else if (warp == Cfg::MMA_WARP) {
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc(128, N, true, true);
      for (int j = 0; j < my_tiles; ++j) {
        const int ob = j & 1;
        const uint32_t a1 = smem_u32(smem + ob * Cfg::OPS_BYTES);
        const uint32_t b_hi = a1 + Cfg::A1_BYTES, b_lo = b_hi + Cfg::A2_HALF;
        mbar_wait(&ops_ready[ob], (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < Cfg::RS / 8; ++ks) {
          const uint64_t da = make_desc_mn32(a1 + ks * 1024, Cfg::BLK, 512);
          const uint64_t dbh = make_desc_mn32(b_hi + ks * 1024, Cfg::BLK, 512);
          const uint64_t dbl = make_desc_mn32(b_lo + ks * 1024, Cfg::BLK, 512);
          umma_tf32(tmem, da, dbh, IDESC, (j | ks) ? 1u : 0u);
          umma_tf32(tmem, da, dbl, IDESC, 1u);
        }
        umma_commit(&ops_free[ob]);
      }
    }
  } else {
    // transform warps: raw row-major tiles -> hi/lo, MN-major 32B-base swizzle
    // (two operand buffers: the stores of tile j+1 overlap the MMAs of tile j)
    constexpr int CH1 = K / 4, CH2 = N / 4, CHT = CH1 + CH2;
    constexpr int LOADS = Cfg::RS * CHT / Cfg::XFORM_THREADS;
    static_assert(Cfg::RS * CHT % Cfg::XFORM_THREADS == 0, "loader mapping");
    int j = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
      const int s = j % Cfg::NS;
      const uint32_t ph = (j / Cfg::NS) & 1;
      const long long r0 = t * Cfg::RS;
      const int rows = static_cast<int>(min(static_cast<long long>(Cfg::RS), M - r0));
      const uint8_t* st = ring + s * Cfg::STAGE_BYTES;
      mbar_wait(&full[s], ph);
      float4 x[LOADS];
#pragma unroll
      for (int i = 0; i < LOADS; ++i) {
        const int idx = tid + Cfg::XFORM_THREADS * i;
        const int row = idx / CHT, ch = idx - row * CHT;
        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows) {
          x[i] = ch < CH1
                     ? *reinterpret_cast<const float4*>(st + (row * CH1 + ch) * 16)
                     : *reinterpret_cast<const float4*>(st + Cfg::P_RAW + (row * CH2 + ch - CH1) * 16);
        }
      }
      mbar_arrive(&empty[s]);
      const int ob = j & 1;
      uint8_t* sA1 = smem + ob * Cfg::OPS_BYTES;
      uint8_t* sA2hi = sA1 + Cfg::A1_BYTES;
      uint8_t* sA2lo = sA2hi + Cfg::A2_HALF;
      mbar_wait(&ops_free[ob], ((j >> 1) & 1) ^ 1u);  // MMAs of tile j-2 have read this buffer
#pragma unroll
      for (int i = 0; i < LOADS; ++i) {
        const int idx = tid + Cfg::XFORM_THREADS * i;
        const int row = idx / CHT, ch = idx - row * CHT;
        float4 hi, lo;
        split_tf32(x[i], hi, lo);
        if (ch < CH1) {
          const uint32_t off = (ch >> 3) * Cfg::BLK + sw128b32_off(row, ch & 7);
          *reinterpret_cast<float4*>(sA1 + off) = hi;
          *reinterpret_cast<float4*>(sA1 + (K / 32) * Cfg::BLK + off) = lo;
        } else {
          const int c2 = ch - CH1;
          const uint32_t off = (c2 >> 3) * Cfg::BLK + sw128b32_off(row, c2 & 7);
          *reinterpret_cast<float4*>(sA2hi + off) = hi;
          *reinterpret_cast<float4*>(sA2lo + off) = lo;
        }
      }
      fence_async_smem();
      mbar_arrive(&ops_ready[ob]);
    }
    // the last commit covers every MMA (in-order completion)
    if (my_tiles > 0) mbar_wait(&ops_free[(my_tiles - 1) & 1], ((my_tiles - 1) >> 1) & 1);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // final epilogue: lanes 0..63 hold hi(P)^T.G, lanes 64..127 lo(P)^T.G
  float* sOut = reinterpret_cast<float*>(smem);
  if (warp < 4 && my_tiles > 0) {
    const int q = warp;
    float v[32];
    for (int pass = 0; pass < 2; ++pass) {
      if ((q >> 1) == pass) {
        const int feat = (q & 1) * 32 + lane;
        for (int cg = 0; cg < N / 32; ++cg) {
          tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + cg * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float* dst = sOut + feat * N + cg * 32 + i;
            *dst = pass == 0 ? v[i] : *dst + v[i];
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  float* dst = part + static_cast<size_t>(blockIdx.x) * K * N;
  for (int e = tid; e < K * N; e += Cfg::THREADS) dst[e] = my_tiles > 0 ? sOut[e] : 0.f;
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
  using Cfg = GatherCfg<F, N>;
  static bool attr = false;
  if (!attr) {
    ATH_CUDA(cudaFuncSetAttribute(k_pipe_gather<F, N, TRANSB, EPI>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  const int grid = std::min(a.num_tiles, ctx().sm_count);
  k_pipe_gather<F, N, TRANSB, EPI><<<grid, Cfg::THREADS, Cfg::SMEM, ctx().stream>>>(a);
  ATH_LAUNCHED_T(EPI == EPI_ACT ? "pipe_gather_fwd" : "pipe_gather_bwd");
  return ATHENA_OK;
}

}  // namespace

bool pipe_gather_supported(const Batch* b, int F, int N) {
  return pipe_enabled() && b->num_tiles > 0 && F == 64 && N == 64;
}

bool pipe_tn_supported(int K, int N) { return pipe_enabled() && K == 64 && (N == 64 || N == 32); }

// forward: out = act( (A_hat X) W ),  P = A_hat X           (W row-major [F][N])
int launch_pipe_gather_fwd(const Batch* b, const float* X, const float* W, float* P, float* out,
                           int F, int N, int act) {
  GatherArgs a{};
  a.tiles = b->tiles.as<int4>();
  a.num_tiles = b->num_tiles;
  a.row_ptr = b->row_ptr;
  a.col = b->col;
  a.coef = b->coef;
  a.X = X;
  a.W = W;
  a.P = P;
  a.out = out;
  a.Hin = nullptr;
  a.act = act;
  ATH_REQUIRE(F == 64 && N == 64, ATHENA_ERR_ARG, "pipe_gather_fwd: unsupported shape");
  return launch_gather_t<64, 64, false, EPI_ACT>(a);
}

// backward: out = ( (A^T-gather of G) W^T ) .* act'(Hin)      (W row-major [N][F]: W_t as stored)
int launch_pipe_gather_bwd(const Batch* b, const float* G, const float* W, const float* Hin,
                           float* out, int F, int N, int act) {
  GatherArgs a{};
  a.tiles = b->tiles.as<int4>();
  a.num_tiles = b->num_tiles;
  a.row_ptr = b->csc_ptr;
  a.col = b->csc_src;
  a.coef = nullptr;
  a.X = G;
  a.W = W;
  a.P = nullptr;
  a.out = out;
  a.Hin = Hin;
  a.act = act;
  ATH_REQUIRE(F == 64 && N == 64, ATHENA_ERR_ARG, "pipe_gather_bwd: unsupported shape");
  return launch_gather_t<64, 64, true, EPI_ACTGRAD>(a);
}

template <int N>
static int launch_pipe_tn_t(const float* P, const float* G, float* dW, int64_t M, DevBuf& scratch) {
  using Cfg = Tn2Cfg<N>;
  static bool attr = false;
  if (!attr) {
    ATH_CUDA(cudaFuncSetAttribute(k_pipe_tn<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM));
    attr = true;
  }
  const int64_t ntiles = cdiv(M, Cfg::RS);
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctx().sm_count);
  ATH_TRY(scratch.reserve(sizeof(float) * (size_t)grid * Cfg::K * N));
  k_pipe_tn<N><<<grid, Cfg::THREADS, Cfg::SMEM, ctx().stream>>>(P, G, scratch.as<float>(), M);
  ATH_LAUNCHED_T("pipe_tn");
  k_pipe_tn_reduce<<<(unsigned)cdiv((int64_t)Cfg::K * N, 128), 128, 0, ctx().stream>>>(
      scratch.as<float>(), grid, Cfg::K * N, dW);
  ATH_LAUNCHED_T("pipe_tn_reduce");
  return ATHENA_OK;
}

// dW[64 x N] += P^T . G     (P [M][64], G [M][N], both dense row-major)
int launch_pipe_tn(const float* P, const float* G, float* dW, int64_t M, int N, DevBuf& scratch) {
  if (M == 0) return ATHENA_OK;
  if (N == 64) return launch_pipe_tn_t<64>(P, G, dW, M, scratch);
  if (N == 32) return launch_pipe_tn_t<32>(P, G, dW, M, scratch);
  ATH_REQUIRE(false, ATHENA_ERR_ARG, "pipe_tn: unsupported N=%d", N);
}

}  // namespace athena
