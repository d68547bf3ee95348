// tcgen05 (5th-gen tensor core) kernels for the dense part of the Kipf layers
// when the feature width makes it a real contraction (32/64 here):
//
//   k_tc_rows<K,N,TRANSB>   C[m, :] = act( A'[m, :] . W )      m over all vertices
//        forward   Y  = P . W_t            (matmul, athena_kipf_msgpass_layer.f90:951)
//        backward  dP = gY . W_t^T         (matmul partial)
//   k_tc_tn<K,N>            dW += A1^T . A2                      reduction over all vertices
//        backward  dW_t = P^T . gY         (matmul partial, summed over vertices and samples)
//
// with A' = A or A .* act'(H) (the activation derivative is fused into the
// operand load).  Operands are staged by the CUDA cores into 128B-swizzled
// shared-memory tiles (tc_common.cuh), split hi/lo for fp32-level accuracy,
// multiplied by tcgen05.mma.kind::tf32 into a TMEM accumulator, and read back
// with tcgen05.ld for the epilogue.  One elected thread issues the MMAs;
// completion is tracked with tcgen05.commit -> mbarrier.
//
// The path is HBM-bound (2 x 4VF bytes per launch against ~2VF^2 flops), so
// the design goal is to keep the memory pipes busy: k_tc_rows runs two CTAs
// per SM, k_tc_tn double-buffers its operand tiles inside one CTA so that the
// loads of tile i+1 overlap the MMAs of tile i.
#include <algorithm>

#include "athena_internal.h"
#include "tc_common.cuh"

namespace athena {

using namespace tc;

__device__ __forceinline__ float tc_act_apply(int act, float x) {
  switch (act) {
    case ATHENA_ACT_RELU: return fmaxf(x, 0.f);
    case ATHENA_ACT_LEAKY_RELU: return fmaxf(x * 0.01f, x);
    case ATHENA_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case ATHENA_ACT_TANH: return tanhf(x);
    default: return x;
  }
}
__device__ __forceinline__ float tc_act_grad(int act, float y, float g) {
  switch (act) {
    case ATHENA_ACT_RELU: return y > 0.f ? g : 0.f;
    case ATHENA_ACT_LEAKY_RELU: return y > 0.f ? g : g * 0.01f;
    case ATHENA_ACT_SIGMOID: return g * (y * (1.f - y));
    case ATHENA_ACT_TANH: return g * (1.f - y * y);
    default: return g;
  }
}
__device__ __forceinline__ float4 tc_act_grad4(int act, const float4& y, const float4& g) {
  return make_float4(tc_act_grad(act, y.x, g.x), tc_act_grad(act, y.y, g.y),
                     tc_act_grad(act, y.z, g.z), tc_act_grad(act, y.w, g.w));
}

// ---------------------------------------------------------------------------
// row-tile transform
// ---------------------------------------------------------------------------
template <int K, int N>
struct RowsCfg {
  static constexpr int THREADS = 256;
  static constexpr int KB = K / 32;                 // 32-column blocks along K
  static constexpr int A_BYTES = KB * 16384;        // [128 x K] fp32, one of hi / lo
  static constexpr int B_BLK = 2 * N * 128;         // [2N x 32] block of the stacked hi|lo weight
  static constexpr int B_BYTES = KB * B_BLK;
  static constexpr int STAGE_LD = N + 4;            // padded row of the epilogue staging tile
  static constexpr int STAGE_BYTES = 128 * STAGE_LD * 4;
  static constexpr int R0_BYTES = (2 * A_BYTES > STAGE_BYTES ? 2 * A_BYTES : STAGE_BYTES);
  static constexpr int R0_PAD = (R0_BYTES + 1023) / 1024 * 1024;
  static constexpr int SMEM = 1024 /*align*/ + R0_PAD + B_BYTES + 64;
  static constexpr int TMEM_COLS = (2 * N <= 32 ? 32 : 2 * N <= 64 ? 64 : 2 * N <= 128 ? 128 : 256);
};

template <int K, int N, bool TRANSB>
__global__ void __launch_bounds__(256, 2)
k_tc_rows(const float* __restrict__ A, int lda, const float* __restrict__ Hact, int act_in,
          const float* __restrict__ W, float* __restrict__ C, int ldc, long long M, int act_out) {
  using Cfg = RowsCfg<K, N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sAhi = smem;
  uint8_t* sAlo = smem + Cfg::A_BYTES;
  float* sStage = reinterpret_cast<float*>(smem);
  uint8_t* sB = smem + Cfg::R0_PAD;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + Cfg::B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  if (tid == 32) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  // stacked weight operand: rows [0,N) = hi(W^T), rows [N,2N) = lo(W^T); K-major
  for (int idx = tid; idx < K * N; idx += Cfg::THREADS) {
    int k, n;
    if (!TRANSB) { k = idx / N; n = idx - k * N; } else { n = idx / K; k = idx - n * K; }
    float w = __ldg(W + idx);
    float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    float lo = w - hi;
    uint8_t* blk = sB + (k >> 5) * Cfg::B_BLK;
    int c = k & 31;
    *reinterpret_cast<float*>(blk + sw128_off(n, c >> 2) + (c & 3) * 4) = hi;
    *reinterpret_cast<float*>(blk + sw128_off(N + n, c >> 2) + (c & 3) * 4) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t aHi = smem_u32(sAhi), aLo = smem_u32(sAlo), bAddr = smem_u32(sB);
  constexpr uint32_t IDESC = make_idesc(128, 2 * N, false, false);
  constexpr int CH = K / 4;                          // 16-byte chunks per row
  constexpr int LOADS = 128 * CH / Cfg::THREADS;
  uint32_t phase = 0;
  const long long ntiles = (M + 127) / 128;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long r0 = tile * 128;
    // ---- stage the A tile (hi / lo), optional fused activation derivative ----
    float4 x[LOADS];
#pragma unroll
    for (int j = 0; j < LOADS; ++j) {
      int idx = tid + Cfg::THREADS * j;
      int row = idx / CH, ch = idx - row * CH;
      long long grow = r0 + row;
      x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grow < M) {
        x[j] = __ldg(reinterpret_cast<const float4*>(A + grow * lda) + ch);
        if (Hact != nullptr) {
          float4 h = __ldg(reinterpret_cast<const float4*>(Hact + grow * lda) + ch);
          x[j] = tc_act_grad4(act_in, h, x[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < LOADS; ++j) {
      int idx = tid + Cfg::THREADS * j;
      int row = idx / CH, ch = idx - row * CH;
      float4 hi, lo;
      split_tf32_safe(x[j], hi, lo);
      uint32_t off = (ch >> 3) * 16384 + sw128_off(row, ch & 7);
      *reinterpret_cast<float4*>(sAhi + off) = hi;
      *reinterpret_cast<float4*>(sAlo + off) = lo;
    }
    fence_async_smem();
    __syncthreads();
    // ---- one thread issues the MMAs: D[128 x 2N] = (A_hi + A_lo) . [W_hi | W_lo] ----
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < Cfg::KB; ++kb) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint64_t db = make_desc(bAddr + kb * Cfg::B_BLK + kk * 32, 16, 1024);
          uint64_t dh = make_desc(aHi + kb * 16384 + kk * 32, 16, 1024);
          uint64_t dl = make_desc(aLo + kb * 16384 + kk * 32, 16, 1024);
          umma_tf32(tmem, dh, db, IDESC, (kb | kk) ? 1u : 0u);
          umma_tf32(tmem, dl, db, IDESC, 1u);
        }
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
    // ---- epilogue: TMEM -> registers -> (hi + lo, activation) -> staging -> global ----
    {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      for (int cg = warp >> 2; cg < N / 32; cg += Cfg::THREADS / 128) {
        float vh[32], vl[32];
        uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16) + cg * 32;
        tmem_ld32(taddr, vh);
        tmem_ld32(taddr + N, vl);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 o = make_float4(tc_act_apply(act_out, vh[i] + vl[i]),
                                 tc_act_apply(act_out, vh[i + 1] + vl[i + 1]),
                                 tc_act_apply(act_out, vh[i + 2] + vl[i + 2]),
                                 tc_act_apply(act_out, vh[i + 3] + vl[i + 3]));
          *reinterpret_cast<float4*>(sStage + row * Cfg::STAGE_LD + cg * 32 + i) = o;
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    constexpr int OCH = N / 4;
#pragma unroll
    for (int j = 0; j < 128 * OCH / Cfg::THREADS; ++j) {
      int idx = tid + Cfg::THREADS * j;
      int row = idx / OCH, ch = idx - row * OCH;
      long long grow = r0 + row;
      if (grow < M)
        *(reinterpret_cast<float4*>(C + grow * ldc) + ch) =
            *reinterpret_cast<const float4*>(sStage + row * Cfg::STAGE_LD + ch * 4);
    }
    __syncthreads();  // staging aliases the A tiles of the next iteration
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------
// dW = A1^T . A2 over all rows (K = 64 features of A1 stacked hi|lo -> M' = 128; K = 32:
// hi | lo | 64 zero rows, the same M' = 128 instruction shape)
// ---------------------------------------------------------------------------
template <int KW, int N>
struct TnCfg {
  static constexpr int K = KW;
  static constexpr int THREADS = 512;
  static constexpr int RS = 64;                      // rows per stage
  static constexpr int BLK = RS * 128;               // [64 rows x 32 feats] = 8192 B
  static constexpr int A1_BYTES = 4 * BLK;                   // hi blocks, lo blocks (, zero blocks)
  static constexpr int A2_HALF = (N / 32) * BLK;             // hi (or lo) of A2
  static constexpr int STAGE = A1_BYTES + 2 * A2_HALF;
  static constexpr int SMEM = 1024 + 2 * STAGE + 64;
  static constexpr int TMEM_COLS = (N <= 32 ? 32 : 64);
  // the loop is a load -> split -> MMA chain with one 64-row stage in flight per CTA: where two
  // CTAs fit one SM (N = 32) the second one doubles the bytes in flight
  static constexpr int CTAS_PER_SM = (2 * (SMEM + 1024) <= 227 * 1024) ? 2 : 1;
};

template <int KW, int N>
__global__ void __launch_bounds__(512, TnCfg<KW, N>::CTAS_PER_SM)
k_tc_tn(const float* __restrict__ A1, int lda1, const float* __restrict__ A2, int lda2,
        const float* __restrict__ Hact, int act_in, float* __restrict__ part, long long M) {
  using Cfg = TnCfg<KW, N>;
  constexpr int K = Cfg::K;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * Cfg::STAGE);  // [2] stage consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  if (tid == 32) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();  // TMEM allocation and barrier set-up overlapped the previous kernel's tail
  pdl_launch_dependents();
  constexpr uint32_t IDESC = make_idesc(128, N, true, true);
  constexpr int CH1 = K / 4, CH2 = N / 4, CHT = CH1 + CH2;
  constexpr int LOADS = Cfg::RS * CHT / Cfg::THREADS;
  static_assert(Cfg::RS * CHT % Cfg::THREADS == 0, "loader mapping");
  const long long ntiles = (M + Cfg::RS - 1) / Cfg::RS;
  if (K < 64) {
    // the unused half of the M' = 128 operand: zero rows, written once
    for (int st = 0; st < 2; ++st)
      for (int e = tid; e < (4 - 2 * (K / 32)) * Cfg::BLK / 16; e += Cfg::THREADS)
        reinterpret_cast<float4*>(smem + st * Cfg::STAGE + 2 * (K / 32) * Cfg::BLK)[e] =
            make_float4(0.f, 0.f, 0.f, 0.f);
  }
  uint32_t phase[2] = {0u, 0u};
  int it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    uint8_t* sA1 = smem + s * Cfg::STAGE;
    uint8_t* sA2hi = sA1 + Cfg::A1_BYTES;
    uint8_t* sA2lo = sA2hi + Cfg::A2_HALF;
    const long long r0 = tile * Cfg::RS;
    float4 x[LOADS];
#pragma unroll
    for (int j = 0; j < LOADS; ++j) {
      int idx = tid + Cfg::THREADS * j;
      int row = idx / CHT, ch = idx - row * CHT;
      long long grow = r0 + row;
      x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grow < M) {
        if (ch < CH1) {
          x[j] = __ldg(reinterpret_cast<const float4*>(A1 + grow * lda1) + ch);
        } else {
          x[j] = __ldg(reinterpret_cast<const float4*>(A2 + grow * lda2) + (ch - CH1));
          if (Hact != nullptr) {
            float4 h = __ldg(reinterpret_cast<const float4*>(Hact + grow * lda2) + (ch - CH1));
            x[j] = tc_act_grad4(act_in, h, x[j]);
          }
        }
      }
    }
    // the MMAs that read this stage two iterations ago must have finished
    if (it >= 2) {
      mbar_wait(&bars[s], phase[s]);
      phase[s] ^= 1u;
    }
#pragma unroll
    for (int j = 0; j < LOADS; ++j) {
      int idx = tid + Cfg::THREADS * j;
      int row = idx / CHT, ch = idx - row * CHT;
      float4 hi, lo;
      split_tf32(x[j], hi, lo);
      if (ch < CH1) {
        uint32_t off = (ch >> 3) * Cfg::BLK + sw128b32_off(row, ch & 7);
        *reinterpret_cast<float4*>(sA1 + off) = hi;
        *reinterpret_cast<float4*>(sA1 + (K / 32) * Cfg::BLK + off) = lo;
      } else {
        int c2 = ch - CH1;
        uint32_t off = (c2 >> 3) * Cfg::BLK + sw128b32_off(row, c2 & 7);
        *reinterpret_cast<float4*>(sA2hi + off) = hi;
        *reinterpret_cast<float4*>(sA2lo + off) = lo;
      }
    }
    fence_async_smem();
    __syncthreads();
    if (warp == 0) {
      // converged-warp issue (the instruction is predicated on one elected lane): half the
      // issue slots of a single-thread branch, see tc_common.cuh
      const uint32_t leader = elect_one();
      tc_fence_after();
      const uint32_t a1 = smem_u32(sA1), b_hi = smem_u32(sA2hi), b_lo = smem_u32(sA2lo);
#pragma unroll
      for (int ks = 0; ks < Cfg::RS / 8; ++ks) {
        uint64_t da = make_desc_mn32(a1 + ks * 1024, Cfg::BLK, 512);
        uint64_t dbh = make_desc_mn32(b_hi + ks * 1024, Cfg::BLK, 512);
        uint64_t dbl = make_desc_mn32(b_lo + ks * 1024, Cfg::BLK, 512);
        umma_tf32_w(leader, tmem, da, dbh, IDESC, (it | ks) ? 1u : 0u);
        umma_tf32_w(leader, tmem, da, dbl, IDESC, 1u);
      }
      umma_commit_w(leader, &bars[s]);
      __syncwarp();
    }
  }
  // drain: commits complete in issue order, so the last one covers everything
  if (it >= 1) {
    const int s = (it - 1) & 1;
    mbar_wait(&bars[s], phase[s]);
  }
  tc_fence_after();
  float* sOut = reinterpret_cast<float*>(smem);  // [64][N] (stage buffers are free now)
  constexpr int NQ = 2 * (K / 32);  // warps that hold results: hi lanes [0, K), lo lanes [K, 2K)
  if (it >= 1 && warp < NQ) {
    const int q = warp;
    float v[32];
    for (int pass = 0; pass < 2; ++pass) {
      if (q / (K / 32) == pass) {
        const int feat = (q % (K / 32)) * 32 + lane;
        for (int cg = 0; cg < N / 32; ++cg) {
          tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + cg * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float* dst = sOut + feat * N + cg * 32 + i;
            *dst = pass == 0 ? v[i] : *dst + v[i];
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NQ * 32) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  float* dst = part + static_cast<size_t>(blockIdx.x) * K * N;
  for (int e = tid; e < K * N; e += Cfg::THREADS) dst[e] = (it >= 1) ? sOut[e] : 0.f;
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

__global__ void k_tc_tn_reduce(const float* __restrict__ part, int nparts, int KN,
                               float* __restrict__ dW) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= KN) return;
  float s = 0.f;
  for (int c = 0; c < nparts; ++c) s += part[static_cast<size_t>(c) * KN + e];
  dW[e] += s;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static bool tc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* s = getenv("ATHENA_CUDA_DISABLE_TC");
    on = (s && atoi(s) != 0) ? 0 : 1;
  }
  return on == 1;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool tc_rows_supported(int K, int N, int lda, int ldc, const void* A, const void* C) {
  return tc_enabled() && (K == 32 || K == 64) && (N == 32 || N == 64) && lda == K && ldc == N &&
         aligned16(A) && aligned16(C);
}

bool tc_tn_supported(int K, int N, int lda1, int lda2, const void* A1, const void* A2) {
  return tc_enabled() && (K == 64 || K == 32) && (N == 32 || N == 64) && lda1 == K && lda2 == N &&
         aligned16(A1) && aligned16(A2);
}

template <int K, int N, bool TRANSB>
static int launch_rows_t(const float* A, int lda, const float* Hact, int act_in, const float* W,
                         float* C, int ldc, int64_t M, int act_out) {
  using Cfg = RowsCfg<K, N>;
  static bool attr = false;
  if (!attr) {
    ATH_CUDA(cudaFuncSetAttribute(k_tc_rows<K, N, TRANSB>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr = true;
  }
  int64_t ntiles = cdiv(M, 128);
  int grid = (int)std::min<int64_t>(ntiles, 2 * (int64_t)ctx().sm_count);
  k_tc_rows<K, N, TRANSB><<<grid, Cfg::THREADS, Cfg::SMEM, ctx().stream>>>(
      A, lda, Hact, act_in, W, C, ldc, M, act_out);
  ATH_LAUNCHED_T(TRANSB ? "tc_rows_nt" : "tc_rows_nn");
  return ATHENA_OK;
}

// C = act_out( (A .* act_in'(Hact)) . op(W) ); transb: W is [N][K] row-major, else [K][N]
int launch_tc_rows(bool transb, const float* A, int lda, const float* Hact, int act_in,
                   const float* W, float* C, int ldc, int64_t M, int N, int K, int act_out) {
  if (M == 0) return ATHENA_OK;
#define ATH_TC_ROWS(KK, NN)                                                                   \
  if (K == KK && N == NN)                                                                     \
    return transb ? launch_rows_t<KK, NN, true>(A, lda, Hact, act_in, W, C, ldc, M, act_out)  \
                  : launch_rows_t<KK, NN, false>(A, lda, Hact, act_in, W, C, ldc, M, act_out)
  ATH_TC_ROWS(64, 64);
  ATH_TC_ROWS(64, 32);
  ATH_TC_ROWS(32, 64);
  ATH_TC_ROWS(32, 32);
#undef ATH_TC_ROWS
  ATH_REQUIRE(false, ATHENA_ERR_ARG, "tc_rows: unsupported shape K=%d N=%d", K, N);
}

template <int K, int N>
static int launch_tn_t(const float* A1, int lda1, const float* A2, int lda2, const float* Hact,
                       int act_in, float* dW, int64_t M, DevBuf& scratch, DeferList* defer) {
  using Cfg = TnCfg<K, N>;
  static bool attr = false;
  if (!attr) {
    ATH_CUDA(cudaFuncSetAttribute(k_tc_tn<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM));
    attr = true;
  }
  int64_t ntiles = cdiv(M, Cfg::RS);
  int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctx().sm_count * Cfg::CTAS_PER_SM);
  ATH_TRY(scratch.reserve(sizeof(float) * (size_t)grid * Cfg::K * N));
  ATH_CUDA(launch_pdl(k_tc_tn<K, N>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM, ctx().stream, A1,
                      lda1, A2, lda2, Hact, act_in, scratch.as<float>(), (long long)M));
  ATH_LAUNCHED_T("tc_tn");
  if (defer) {  // the fold of the per-CTA partials rides on the finalize launch
    defer->jobs.push_back(DeferJob{scratch.as<float>(), grid, Cfg::K * N, dW});
    return ATHENA_OK;
  }
  k_tc_tn_reduce<<<(unsigned)cdiv((int64_t)Cfg::K * N, 128), 128, 0, ctx().stream>>>(
      scratch.as<float>(), grid, Cfg::K * N, dW);
  ATH_LAUNCHED_T("tc_tn_reduce");
  return ATHENA_OK;
}

// dW[K x N] += A1^T . (A2 .* act_in'(Hact))
int launch_tc_tn(const float* A1, int lda1, const float* A2, int lda2, const float* Hact,
                 int act_in, float* dW, int64_t M, int N, int K, DevBuf& scratch,
                 DeferList* defer) {
  if (M == 0) return ATHENA_OK;
  ATH_REQUIRE(K == 64 || K == 32, ATHENA_ERR_ARG, "tc_tn: K must be 32 or 64");
#define ATH_TN(KK, NN) \
  return launch_tn_t<KK, NN>(A1, lda1, A2, lda2, Hact, act_in, dW, M, scratch, defer)
  if (K == 64 && N == 64) ATH_TN(64, 64);
  if (K == 64 && N == 32) ATH_TN(64, 32);
  if (K == 32 && N == 64) ATH_TN(32, 64);
  if (K == 32 && N == 32) ATH_TN(32, 32);
#undef ATH_TN
  ATH_REQUIRE(false, ATHENA_ERR_ARG, "tc_tn: unsupported N=%d", N);
}

}  // namespace athena
