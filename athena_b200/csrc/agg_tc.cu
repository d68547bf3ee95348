// Fused Kipf layer-step for LARGE graphs (the batch is not tileable: a graph has more than
// 128 vertices, so neighbours live anywhere in HBM / L2), feature width 128:
//
//     out[v,:] = act( ( sum_{w in row v} c_w * X[col[w],:] ) . W )          P optional
//
// i.e. kipf_propagate + matmul + activation%apply in one pass
// (athena_diffstruc_extd_sub_kipf.f90:29-46, athena_kipf_msgpass_layer.f90:943-952), the
// "large-graph SpMM + tcgen05 transform" configuration of BASELINE.json (cfg3).
//
// One persistent CTA per SM, warp-specialised:
//   16 gather warps   warp-per-row SpMM over the CSR: a warp owns 8 rows of the 128-row
//                     tile; per entry every lane loads 16 bytes of the neighbour row (one
//                     coalesced 512-byte request), column indices / coefficients are read
//                     32 entries at a time and broadcast with shuffles; entries are added
//                     in ascending order (the reference's order), four loads in flight per
//                     lane.  The finished rows are split into TF32 hi/lo and written as the
//                     128B-swizzled K-major A operand, one 64-feature K-half at a time (the
//                     full K = 128 operand pair plus the weights would not fit in shared
//                     memory: W hi|lo alone is 128 KB).
//   MMA warp          tcgen05.mma.kind::tf32, M = 128, N = 256 (= [hi(W); lo(W)]), two K-halves
//                     accumulate into one TMEM accumulator.
//   4 epilogue warps  tcgen05.ld, hi + lo, activation, transposition through a padded patch,
//                     coalesced row stores (four full 128-byte lines per instruction).
// The gather dominates (about 1 MB of neighbour rows per tile), so the transform, the
// epilogue and the stores of tile j hide entirely under the gathers of tile j+1; the
// aggregate P never travels to HBM unless the backward pass needs it.
#include <algorithm>

#include "athena_internal.h"
#include "tc_common.cuh"

namespace athena {

using namespace tc;

namespace {

__device__ __forceinline__ float agg_act(int act, float x) {
  switch (act) {
    case ATHENA_ACT_RELU: return fmaxf(x, 0.f);
    case ATHENA_ACT_LEAKY_RELU: return fmaxf(x * 0.01f, x);
    case ATHENA_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case ATHENA_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

// 16-byte read-only global load that is not issued (and yields 0) when `p` is false
__device__ __forceinline__ float4 ldg128_pred(const float* ptr, bool p) {
  float4 v;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "mov.f32 %0, 0f00000000;\n\t"
      "mov.f32 %1, 0f00000000;\n\t"
      "mov.f32 %2, 0f00000000;\n\t"
      "mov.f32 %3, 0f00000000;\n\t"
      "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t"
      "}"
      : "=&f"(v.x), "=&f"(v.y), "=&f"(v.z), "=&f"(v.w)
      : "l"(ptr), "r"(static_cast<uint32_t>(p)));
  return v;
}

// TMEM as a parking lot for finished rows: thread i of warp w writes / reads 4 consecutive
// 32-bit columns of TMEM lane 32 * (w % 4) + i (layout is irrelevant: the same thread reads
// back what it wrote)
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float4& v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr),
               "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
               "r"(__float_as_uint(v.w))
               : "memory");
}
__device__ __forceinline__ float4 tmem_ld4(uint32_t taddr) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(taddr)
               : "memory");
  return make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2),
                     __uint_as_float(r3));
}

struct AggArgs {
  const int32_t* row_ptr;
  const int32_t* col;
  const float* coef;  // nullptr -> unit coefficients
  const float* X;     // [V][128]
  const float* W;     // row-major [128][128]
  float* P;           // optional [V][128]
  float* out;         // [V][128]
  long long V;
  int act;
};

struct AggCfg {
  static constexpr int F = 128, N = 128;
  static constexpr int GATHER_WARPS = 16, GATHER_THREADS = GATHER_WARPS * 32;
  static constexpr int RPW = TILE_ROWS / GATHER_WARPS;          // rows per warp and tile
  static constexpr int EPI_WARP0 = 16;                          // warps 16..19; warp 16 also issues the MMAs
  static constexpr int MMA_WARP = 16;
  static constexpr int THREADS = 20 * 32;                       // 20 warps -> 96 registers per thread
  static constexpr int UNROLL = 12;                             // neighbour rows in flight per lane
  static constexpr int A_HALF = 2 * 16384;                      // [128 x 64] fp32, hi or lo
  static constexpr int OFF_A = 0;                               // hi | lo of the current K-half
  static constexpr int B_BLK = 2 * N * 128;                     // [256 x 32] block
  static constexpr int B_HALF = 2 * B_BLK;                      // one K-half of [hi(W); lo(W)]
  static constexpr int OFF_B = 2 * A_HALF;
  static constexpr int OFF_BAR = OFF_B + 2 * B_HALF;
  static constexpr int OFF_EPI = OFF_BAR + 256;
  static constexpr int EPI_PITCH = 36;
  static constexpr int EPI_PATCH = 32 * EPI_PITCH;
  static constexpr int SMEM = 1024 + OFF_EPI + 4 * EPI_PATCH * 4;
  static constexpr int TMEM_COLS = 512;   // 256 accumulator columns + 128 parking columns
  static_assert(SMEM <= 232448, "shared memory budget");
};

__global__ void __launch_bounds__(AggCfg::THREADS, 1) k_agg_tc128(AggArgs a) {
  using Cfg = AggCfg;
  constexpr int F = Cfg::F, N = Cfg::N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sAhi = smem + Cfg::OFF_A;
  uint8_t* sAlo = sAhi + Cfg::A_HALF;
  uint8_t* sB = smem + Cfg::OFF_B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* ops_ready = bars;
  uint64_t* ops_free = bars + 1;
  uint64_t* acc_full = bars + 2;
  uint64_t* acc_empty = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (a.V + TILE_ROWS - 1) / TILE_ROWS;

  if (warp == Cfg::MMA_WARP) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  if (tid == 0) {
    mbar_init(ops_ready, Cfg::GATHER_THREADS);
    mbar_init(ops_free, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);
    mbar_fence_init();
  }
  // weight operand [hi(W^T); lo(W^T)] as [2N][K] K-major, split in two K-halves of two
  // 32-column blocks; one item = one 16-byte chunk (4 consecutive k) of one row n
  for (int item = tid; item < N * (F / 4); item += Cfg::THREADS) {
    const int n = item / (F / 4), kc = item - n * (F / 4);
    const float* src = a.W + (kc * 4) * N + n;
    const float4 w = make_float4(__ldg(src), __ldg(src + N), __ldg(src + 2 * N), __ldg(src + 3 * N));
    float4 hi, lo;
    split_tf32(w, hi, lo);
    uint8_t* blk = sB + (kc >> 4) * Cfg::B_HALF + ((kc >> 3) & 1) * Cfg::B_BLK;
    *reinterpret_cast<float4*>(blk + sw128_off(n, kc & 7)) = hi;
    *reinterpret_cast<float4*>(blk + sw128_off(N + n, kc & 7)) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= Cfg::EPI_WARP0) {
    const int q = warp - Cfg::EPI_WARP0;
    float* patch = reinterpret_cast<float*>(smem + Cfg::OFF_EPI) + q * Cfg::EPI_PATCH;
    float* srow = patch + lane * Cfg::EPI_PITCH;
    int j = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
      const long long r0 = t * TILE_ROWS;
      const int nrows = static_cast<int>(min(static_cast<long long>(TILE_ROWS), a.V - r0));
      float* out_tile = a.out + r0 * N;
      if (warp == Cfg::MMA_WARP && lane == 0) {
        // the transform of this tile: issued by one thread as soon as the gather warps hand
        // over each K-half (the epilogue below needs its result anyway)
        const uint32_t aHi = smem_u32(sAhi), aLo = smem_u32(sAlo), bAddr = smem_u32(sB);
        constexpr uint32_t IDESC = make_idesc(128, 2 * N, false, false);
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          mbar_wait(ops_ready, (2 * j + kh) & 1);
          if (kh == 0) mbar_wait(acc_empty, (j & 1) ^ 1u);  // all four warps drained tile j-1
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t db =
                  make_desc(bAddr + kh * Cfg::B_HALF + kb * Cfg::B_BLK + kk * 32, 16, 1024);
              const uint64_t dh = make_desc(aHi + kb * 16384 + kk * 32, 16, 1024);
              const uint64_t dl = make_desc(aLo + kb * 16384 + kk * 32, 16, 1024);
              umma_tf32(tmem, dh, db, IDESC, (kh | kb | kk) ? 1u : 0u);
              umma_tf32(tmem, dl, db, IDESC, 1u);
            }
          }
          umma_commit(ops_free);               // the A half may be overwritten
          if (kh == 1) umma_commit(acc_full);  // accumulator ready for the epilogue
        }
      }
      __syncwarp();
      mbar_wait(acc_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int blk = 0; blk < N / 32; ++blk) {
#pragma unroll
        for (int cg = 0; cg < 2; ++cg) {
          float vh[16], vl[16];
          const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16) + blk * 32 + cg * 16;
          tmem_ld16(taddr, vh);
          tmem_ld16(taddr + N, vl);
          if (blk == N / 32 - 1 && cg == 1) {
            tc_fence_before();
            mbar_arrive(acc_empty);
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(srow + cg * 16 + i) =
                make_float4(agg_act(a.act, vh[i] + vl[i]), agg_act(a.act, vh[i + 1] + vl[i + 1]),
                            agg_act(a.act, vh[i + 2] + vl[i + 2]),
                            agg_act(a.act, vh[i + 3] + vl[i + 3]));
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int idx = it * 32 + lane;
          const int r = idx >> 3, c = idx & 7;
          const int trow = q * 32 + r;
          if (trow < nrows)
            *reinterpret_cast<float4*>(out_tile + static_cast<size_t>(trow) * N + blk * 32 + c * 4) =
                *reinterpret_cast<const float4*>(patch + r * Cfg::EPI_PITCH + c * 4);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== gather warps: warp per row, 8 rows per tile =================
    const bool has_coef = a.coef != nullptr;
    int j = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
      const long long r0 = t * TILE_ROWS;
      // finished rows are parked in spare TMEM columns so that the registers stay free for
      // loads in flight (12 x 16 B per lane = 96 KB of neighbour rows per SM)
      const uint32_t park =
          tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 256 + (warp >> 2) * (Cfg::RPW * 4);
#pragma unroll 1
      for (int k = 0; k < Cfg::RPW; ++k) {
        const long long row = r0 + warp * Cfg::RPW + k;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < a.V) {
          const int beg = __ldg(a.row_ptr + row), end = __ldg(a.row_ptr + row + 1);
          for (int w0 = beg; w0 < end; w0 += 32) {
            const int myw = w0 + lane;
            int mycol = 0;
            float myc = 0.f;
            if (myw < end) {
              mycol = __ldg(a.col + myw);
              myc = has_coef ? __ldg(a.coef + myw) : 1.f;
            }
            const int cnt = min(32, end - w0);
#pragma unroll 1
            for (int j0 = 0; j0 < cnt; j0 += Cfg::UNROLL) {
              float4 x[Cfg::UNROLL];
              float c[Cfg::UNROLL];
#pragma unroll
              for (int u = 0; u < Cfg::UNROLL; ++u) {  // past the row end: no load, coefficient 0
                const int src = min(j0 + u, 31);
                const int cu = __shfl_sync(0xffffffffu, mycol, src);
                const float cw = __shfl_sync(0xffffffffu, myc, src);
                const bool v = j0 + u < cnt;
                c[u] = v ? cw : 0.f;
                x[u] = ldg128_pred(a.X + static_cast<size_t>(cu) * F + lane * 4, v);
              }
#pragma unroll
              for (int u = 0; u < Cfg::UNROLL; ++u) {
                r.x = fmaf(c[u], x[u].x, r.x);
                r.y = fmaf(c[u], x[u].y, r.y);
                r.z = fmaf(c[u], x[u].z, r.z);
                r.w = fmaf(c[u], x[u].w, r.w);
              }
            }
          }
          if (a.P != nullptr) *(reinterpret_cast<float4*>(a.P + static_cast<size_t>(row) * F) + lane) = r;
        }
        tmem_st4(park + k * 4, r);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      float4 acc[Cfg::RPW];
#pragma unroll
      for (int k = 0; k < Cfg::RPW; ++k) acc[k] = tmem_ld4(park + k * 4);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // hand the tile to the tensor core, one K-half at a time (lanes 0-15 hold features
      // 0..63, lanes 16-31 features 64..127)
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        mbar_wait(ops_free, ((2 * j + kh) & 1) ^ 1u);  // MMAs that read the previous half are done
        if ((lane >> 4) == kh) {
          const int ch = lane & 15;
#pragma unroll
          for (int k = 0; k < Cfg::RPW; ++k) {
            const int trow = warp * Cfg::RPW + k;
            float4 hi, lo;
            split_tf32_safe(acc[k], hi, lo);
            const uint32_t off = (ch >> 3) * 16384 + sw128_off(trow, ch & 7);
            *reinterpret_cast<float4*>(sAhi + off) = hi;
            *reinterpret_cast<float4*>(sAlo + off) = lo;
          }
        }
        fence_async_smem();
        mbar_arrive(ops_ready);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::MMA_WARP) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

bool agg_tc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* s = getenv("ATHENA_CUDA_DISABLE_AGGTC");
    const char* t = getenv("ATHENA_CUDA_DISABLE_TC");
    on = ((s && atoi(s) != 0) || (t && atoi(t) != 0)) ? 0 : 1;
  }
  return on == 1;
}

}  // namespace

bool agg_tc_supported(int F, int N, const void* X, const void* out) {
  return agg_tc_enabled() && F == 128 && N == 128 &&
         (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
}

// out = act( (A_hat X) W ), P = A_hat X (optional)      W row-major [128][128]
int launch_agg_tc_fwd(const Batch* b, const float* X, const float* W, float* P, float* out,
                      int act) {
  if (b->V == 0) return ATHENA_OK;
  AggArgs a{};
  a.row_ptr = b->row_ptr;
  a.col = b->col;
  a.coef = b->coef;
  a.X = X;
  a.W = W;
  a.P = P;
  a.out = out;
  a.V = b->V;
  a.act = act;
  static bool attr = false;
  if (!attr) {
    ATH_CUDA(cudaFuncSetAttribute(k_agg_tc128, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  AggCfg::SMEM));
    attr = true;
  }
  const int grid = (int)std::min<int64_t>(cdiv(b->V, TILE_ROWS), (int64_t)ctx().sm_count);
  k_agg_tc128<<<grid, AggCfg::THREADS, AggCfg::SMEM, ctx().stream>>>(a);
  ATH_LAUNCHED_T("agg_tc_fwd");
  return ATHENA_OK;
}

}  // namespace athena
