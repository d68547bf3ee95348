// Fused tile kernels on the FP32 pipe for mini-batches of small graphs whose feature width
// is below the tensor-core threshold (north_star: "the feature transform runs on tcgen05 only
// when the feature width makes it a real dense contraction (>= 64); otherwise it is fused into
// the aggregation epilogue"), and for the Duvenaud layer, whose per-vertex weight matrix
// W_{d(v)} makes the update a grouped product over degree buckets.
//
// A tile is a run of whole graphs with <= 128 vertices (Batch::tiles), so every neighbour of a
// tile row lies inside the tile: the tile's features are staged ONCE in shared memory and the
// whole layer -- all time steps of update_message_duvenaud and update_readout_duvenaud, or one
// Kipf step -- runs out of shared memory:
//
//   k_duv_fwd   athena_duvenaud_msgpass_layer.f90:755-859  (duvenaud_propagate
//               _sub_duvenaud.f90:34-42, duvenaud_update :204-211, activation, readout matmul,
//               softmax athena_diffstruc_extd_sub.f90:309-313, sum over vertices :848-852)
//               + optionally the [num_outputs, batch] MSE cell (athena_loss.f90:414)
//   k_duv_bwd   the reverse sweep through all of it (:284-368, :115-142, softmax :355-379),
//               recomputing A_t and S_t from the saved z_t instead of reading stored copies
//   k_kipf_fwd  kipf_propagate + matmul + activation   (_sub_kipf.f90:29-46,
//               athena_kipf_msgpass_layer.f90:943-952), any widths up to 128
//   k_kipf_bwd  act', dW = gY^T P, dP = gY W^T, un-normalised CSC scatter (_sub_kipf.f90:101-109)
//
// Arithmetic: dense products run on FFMA2 (fma.rn.f32x2: two fp32 FMAs per instruction, the
// only way to the full FP32 rate of sm_100), one vertex row per lane with the row's weight
// block read as broadcast LDS.128; per-bucket weight blocks are offset by an odd number of
// 16-byte chunks so that lanes of different buckets hit different banks.  Every sum runs in
// the reference's order (ascending CSR entry / ascending k); weight gradients are
// accumulated per CTA in a private partial vector (no atomics) that k_finalize folds in a
// fixed order.
#include <algorithm>
#include <cstdlib>

#include "athena_internal.h"

namespace athena {

namespace {

constexpr int TF_THREADS = 512;
constexpr int TF_WARPS = TF_THREADS / 32;
constexpr int TF_LPR = TF_THREADS / TILE_ROWS;  // lanes per tile row in the row-wise passes
constexpr int TF_MAX_T = 16;
constexpr int TF_NB = 8;                        // outputs per thread in the products
constexpr int TF_IDX = TILE_ENTRIES + 16;       // tile-local neighbour bytes

// row pitch (floats) of a shared-memory tile: a multiple of 4 whose quarter is odd, so that
// row-per-lane LDS.128 / STS.128 of a quarter warp cover all 32 banks
__host__ __device__ __forceinline__ int tf_pitch(int n) {
  int p = (n + 3) & ~3;
  if (p == 0) p = 4;
  if (((p >> 2) & 1) == 0) p += 4;
  return p;
}
__host__ __device__ __forceinline__ int tf_up(int n, int m) { return (n + m - 1) / m * m; }

__device__ __forceinline__ float tf_act(int act, float x) {
  switch (act) {
    case ATHENA_ACT_RELU: return fmaxf(x, 0.f);
    case ATHENA_ACT_LEAKY_RELU: return fmaxf(x * 0.01f, x);
    case ATHENA_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case ATHENA_ACT_TANH: return tanhf(x);
    default: return x;
  }
}
// derivative expressed on the saved output y
__device__ __forceinline__ float tf_act_grad(int act, float y, float g) {
  switch (act) {
    case ATHENA_ACT_RELU: return y > 0.f ? g : 0.f;
    case ATHENA_ACT_LEAKY_RELU: return y > 0.f ? g : g * 0.01f;
    case ATHENA_ACT_SIGMOID: return g * (y * (1.f - y));
    case ATHENA_ACT_TANH: return g * (1.f - y * y);
    default: return g;
  }
}

// d += a * b on both halves: one FFMA2 (the scalar is broadcast by the instruction)
__device__ __forceinline__ void fma2(float2& d, float a, float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  float2 a2 = make_float2(a, a);
  const unsigned long long aa = *reinterpret_cast<unsigned long long*>(&a2);
  const unsigned long long bb = *reinterpret_cast<unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}

// ---- tile movers ---------------------------------------------------------------------

// global [rows][F] -> shared [rows][pitch]; columns F .. 4*ceil(F/4) are zeroed
__device__ __forceinline__ void tf_load_rows(float* dst, int pitch, const float* __restrict__ src,
                                             int rows, int F) {
  const int F4 = (F + 3) >> 2;
  if ((F & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    for (int i = threadIdx.x; i < rows * F4; i += TF_THREADS) {
      const int v = i / F4, c = i - v * F4;
      *reinterpret_cast<float4*>(dst + v * pitch + 4 * c) =
          __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(v) * F) + c);
    }
  } else {
    const int Fp = F4 * 4;
    for (int i = threadIdx.x; i < rows * Fp; i += TF_THREADS) {
      const int v = i / Fp, f = i - v * Fp;
      dst[v * pitch + f] = f < F ? __ldg(src + static_cast<size_t>(v) * F + f) : 0.f;
    }
  }
}

// shared [rows][pitch] -> global [rows][F]
__device__ __forceinline__ void tf_store_rows(float* __restrict__ dst, const float* src, int pitch,
                                              int rows, int F) {
  const int F4 = (F + 3) >> 2;
  if ((F & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    for (int i = threadIdx.x; i < rows * F4; i += TF_THREADS) {
      const int v = i / F4, c = i - v * F4;
      *(reinterpret_cast<float4*>(dst + static_cast<size_t>(v) * F) + c) =
          *reinterpret_cast<const float4*>(src + v * pitch + 4 * c);
    }
  } else {
    for (int i = threadIdx.x; i < rows * F; i += TF_THREADS) {
      const int v = i / F, f = i - v * F;
      dst[static_cast<size_t>(v) * F + f] = src[v * pitch + f];
    }
  }
}

// tile-local CSR (or CSC) structure: ptr_s[0..rows] relative to the tile's first entry and
// the neighbour bytes
__device__ __forceinline__ void tf_load_struct(int* ptr_s, uint8_t* idx_s,
                                               const int32_t* __restrict__ ptr_g,
                                               const uint8_t* __restrict__ idx_g, int r0, int rows,
                                               int e0, int ents) {
  for (int i = threadIdx.x; i <= rows; i += TF_THREADS) ptr_s[i] = __ldg(ptr_g + r0 + i) - e0;
  if ((e0 & 3) == 0) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(idx_g + e0);
    uint32_t* dst = reinterpret_cast<uint32_t*>(idx_s);
    for (int i = threadIdx.x; i < (ents + 3) >> 2; i += TF_THREADS) dst[i] = __ldg(src + i);
  } else {
    for (int i = threadIdx.x; i < ents; i += TF_THREADS) idx_s[i] = __ldg(idx_g + e0 + i);
  }
}

// dst[v][4c .. 4c+3] = ( sum_{e in row v, ascending} c_e * src[idx[e]][4c .. 4c+3] ) [/ (bkt[v]+1)]
// Lanes walk the chunks of a row first: the lanes of a quarter warp read one contiguous
// piece of a source row (conflict-free); index reads are broadcasts.
template <bool COEF>
__device__ __forceinline__ void tf_gather(float* dst, int dpitch, const float* src, int spitch,
                                          int F4, const int* ptr_s, const uint8_t* idx_s,
                                          const float* coef_s, int rows, const uint8_t* bkt_s) {
  for (int i = threadIdx.x; i < rows * F4; i += TF_THREADS) {
    const int v = i / F4, c = i - v * F4;
    const int e1 = ptr_s[v + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = ptr_s[v]; e < e1; ++e) {
      const float4 x = *reinterpret_cast<const float4*>(src + idx_s[e] * spitch + 4 * c);
      if (COEF) {
        const float cw = coef_s[e];
        acc.x = fmaf(cw, x.x, acc.x);
        acc.y = fmaf(cw, x.y, acc.y);
        acc.z = fmaf(cw, x.z, acc.z);
        acc.w = fmaf(cw, x.w, acc.w);
      } else {
        acc.x += x.x;
        acc.y += x.y;
        acc.z += x.z;
        acc.w += x.w;
      }
    }
    if (bkt_s != nullptr) {
      const float d = static_cast<float>(bkt_s[v] + 1);  // A(:,v) / real(d), _sub_duvenaud.f90:208
      acc.x = acc.x / d;
      acc.y = acc.y / d;
      acc.z = acc.z / d;
      acc.w = acc.w / d;
    }
    *reinterpret_cast<float4*>(dst + v * dpitch + 4 * c) = acc;
  }
}

// C[v][n] = epi(v, n, sum_{k ascending} A[v][k] * W[grp[v]][k][n])      v < rows, n < N
// A: shared [rows][pa]; W: shared, row pitch pw (>= N rounded up to TF_NB, pad columns zero),
// group stride gs; C: shared [rows][pc] (columns N .. 4*ceil(N/4) are zeroed).
// One lane = one vertex row and TF_NB outputs; the row's weights arrive as broadcast LDS.128.
template <class Epi>
__device__ __forceinline__ void tf_gemm(float* C, int pc, const float* A, int pa, int K,
                                        const float* W, int pw, int gs, const uint8_t* grp_s,
                                        int rows, int N, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvg = (rows + 31) >> 5, nnb = (N + TF_NB - 1) / TF_NB;
  const int N4 = ((N + 3) >> 2) << 2;
  const int K4 = K >> 2;
  for (int it = warp; it < nvg * nnb; it += TF_WARPS) {
    const int vg = it % nvg, nb = it / nvg;
    const int v = vg * 32 + lane;
    const bool live = v < rows;
    const int vv = live ? v : rows - 1;
    const float* a = A + vv * pa;
    const float* w = W + (grp_s != nullptr ? grp_s[vv] * gs : 0) + nb * TF_NB;
    float2 acc[TF_NB / 2];
#pragma unroll
    for (int j = 0; j < TF_NB / 2; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int k4 = 0; k4 < K4; ++k4) {
      const float4 a4 = *reinterpret_cast<const float4*>(a + 4 * k4);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4* wr = reinterpret_cast<const float4*>(w + (4 * k4 + i) * pw);
#pragma unroll
        for (int j = 0; j < TF_NB / 4; ++j) {
          const float4 w4 = wr[j];
          fma2(acc[2 * j], av[i], make_float2(w4.x, w4.y));
          fma2(acc[2 * j + 1], av[i], make_float2(w4.z, w4.w));
        }
      }
    }
    for (int k = K4 * 4; k < K; ++k) {
      const float ak = a[k];
      const float4* wr = reinterpret_cast<const float4*>(w + k * pw);
#pragma unroll
      for (int j = 0; j < TF_NB / 4; ++j) {
        const float4 w4 = wr[j];
        fma2(acc[2 * j], ak, make_float2(w4.x, w4.y));
        fma2(acc[2 * j + 1], ak, make_float2(w4.z, w4.w));
      }
    }
    if (live) {
      float* c = C + v * pc;
#pragma unroll
      for (int j = 0; j < TF_NB; ++j) {
        const int n = nb * TF_NB + j;
        const float s = (j & 1) ? acc[j >> 1].y : acc[j >> 1].x;
        if (n < N)
          c[n] = epi(v, n, s);
        else if (n < N4)
          c[n] = 0.f;
      }
    }
  }
}

// part[(g*K + k)*N + n] += sum_{p in segment g, ascending} A[list[p]][k] * G[list[p]][n]
// (the layout of a column-major [N, K] parameter block per group: n + N*k + N*K*g).
// One warp owns four k of one group: A arrives as a broadcast LDS.128, G as one conflict-free
// LDS.32 per lane; the CTA-private partial is read before the loop and written after it.
// list_s == nullptr: identity (one segment with all rows).  NH: N <= 32 * NH.
template <int NH>
__device__ __forceinline__ void tf_outer(float* __restrict__ part, const float* A, int pa, int K,
                                         const float* G, int pg, int N, const uint8_t* list_s,
                                         const int* seg_s, int ngroups, int rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K4 = (K + 3) >> 2;
  for (int it = warp; it < ngroups * K4; it += TF_WARPS) {
    const int g = it / K4, kb = it - g * K4;
    const int p0 = seg_s != nullptr ? seg_s[g] : 0;
    const int p1 = seg_s != nullptr ? seg_s[g + 1] : rows;
    if (p0 >= p1) continue;
    float old[4][NH], acc[4][NH];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const int k = 4 * kb + i, n = lane + 32 * h;
        acc[i][h] = 0.f;
        old[i][h] = (k < K && n < N) ? part[(static_cast<size_t>(g) * K + k) * N + n] : 0.f;
      }
#pragma unroll 4
    for (int p = p0; p < p1; ++p) {
      const int v = list_s != nullptr ? list_s[p] : p;
      const float4 a4 = *reinterpret_cast<const float4*>(A + v * pa + 4 * kb);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const int n = lane + 32 * h;
        const float gv = n < N ? G[v * pg + n] : 0.f;
        acc[0][h] = fmaf(a4.x, gv, acc[0][h]);
        acc[1][h] = fmaf(a4.y, gv, acc[1][h]);
        acc[2][h] = fmaf(a4.z, gv, acc[2][h]);
        acc[3][h] = fmaf(a4.w, gv, acc[3][h]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const int k = 4 * kb + i, n = lane + 32 * h;
        if (k < K && n < N)
          part[(static_cast<size_t>(g) * K + k) * N + n] = old[i][h] + acc[i][h];
      }
  }
}

// per-row softmax, in place (athena_diffstruc_extd_sub.f90:309-313: max-subtracted)
__device__ __forceinline__ void tf_softmax_rows(float* Y, int py, int rows, int N) {
  const int v = threadIdx.x / TF_LPR, h = threadIdx.x % TF_LPR;
  const bool live = v < rows;
  float* y = Y + (live ? v : 0) * py;
  float mx = -INFINITY;
  if (live)
    for (int n = h; n < N; n += TF_LPR) mx = fmaxf(mx, y[n]);
#pragma unroll
  for (int o = 1; o < TF_LPR; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  if (live)
    for (int n = h; n < N; n += TF_LPR) {
      const float e = expf(y[n] - mx);
      y[n] = e;
      s += e;
    }
#pragma unroll
  for (int o = 1; o < TF_LPR; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (live)
    for (int n = h; n < N; n += TF_LPR) y[n] = y[n] / s;
}

// in place: Y[v][:] (holding the activation output S) <- d loss / d pre-activation for the
// upstream gradient row g(v); softmax: S*g - S*sum(S*g) (athena_diffstruc_extd_sub.f90:369-373)
template <class GRow>
__device__ __forceinline__ void tf_act_bwd_rows(int act, float* Y, int py, int rows, int N,
                                                GRow grow) {
  const int v = threadIdx.x / TF_LPR, h = threadIdx.x % TF_LPR;
  const bool live = v < rows;
  float* y = Y + (live ? v : 0) * py;
  const float* g = grow(live ? v : 0);
  if (act == ATHENA_ACT_SOFTMAX) {
    float dot = 0.f;
    if (live)
      for (int n = h; n < N; n += TF_LPR) dot += y[n] * g[n];
#pragma unroll
    for (int o = 1; o < TF_LPR; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (live)
      for (int n = h; n < N; n += TF_LPR) {
        const float s = y[n];
        y[n] = s * g[n] - s * dot;
      }
  } else if (live) {
    for (int n = h; n < N; n += TF_LPR) y[n] = tf_act_grad(act, y[n], g[n]);
  }
}

__device__ __forceinline__ float tf_block_sum(float v, float* red /* [TF_WARPS] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < TF_WARPS; ++w) r += red[w];
  return r;  // valid in thread 0
}

// ======================================================================================
// Duvenaud layer
// ======================================================================================

struct DuvLayout {  // offsets in floats from the (16-byte aligned) start of dynamic smem
  int w[TF_MAX_T], r[TF_MAX_T], rt[TF_MAX_T];
  int pw[TF_MAX_T], gs[TF_MAX_T], pr[TF_MAX_T], prt[TF_MAX_T];
  int pF, pK, pO, pG, pE;
  int buf[5], ae;
  int ints;    // int32: ptr_s[132] cptr_s[132] seg_s[260] vg_s[128] red[16]
  int bytes;   // uint8: idx_s[TF_IDX] cidx_s[TF_IDX] bkt_s[128] list_s[128]
  int total_bytes;
};

struct DuvArgs {
  const int4* tiles;
  int num_tiles, num_graphs;
  const int32_t* row_ptr;
  const uint8_t* col8;
  const int32_t* eid;
  const int32_t* csc_ptr;
  const uint8_t* csc8;
  const int32_t* vgraph;
  const int32_t* voff;
  const float* X;       // [V][F_0]
  const float* E;       // [E][F_e]
  const float* params;  // W_1..W_T, R_1..R_T (flat, the reference's packing)
  float* Z[TF_MAX_T];   // z_t, [V][F_t]
  int woff[TF_MAX_T], roff[TF_MAX_T];
  int nvf[TF_MAX_T + 1];
  int T, nef, D, min_deg, max_deg, no, act, ract;
  // forward
  float* out;            // [B][no]
  const float* target;   // fused MSE (nullptr: none)
  float* mse_grad;       // [B][no]
  float mse_denom;       // no * global batch
  float* loss_part;      // [grid]
  // backward
  const float* gout;     // [B][no]
  float* gin;            // [V][F_0] or nullptr
  float* part;           // [grid][np] CTA-private partial parameter gradients
  int np;
  DuvLayout lay;
};

struct Layer_dims_t {
  int T, nef, D, no;
  const int* nvf;
};

// forward weights: W_t as [d][k][pw] (n fastest, pad zero), R_t as [f][pr]
// backward: W_t transposed [d][o][pw'] (k fastest, k < F_{t-1} only), R_t and its transpose
static void duv_layout(const Layer_dims_t& L, bool backward, DuvLayout* o) {
  int off = 0;
  int Fmax = 0, Kmax = 0;
  for (int t = 0; t <= L.T; ++t) Fmax = std::max(Fmax, L.nvf[t]);
  for (int t = 1; t <= L.T; ++t) {
    const int Fi = L.nvf[t - 1], Fo = L.nvf[t], K = Fi + L.nef;
    Kmax = std::max(Kmax, K);
    const int i = t - 1;
    if (!backward) {
      o->pw[i] = tf_up(Fo, TF_NB);
      int gs = tf_up(K, 4) * o->pw[i];
      if (((gs >> 2) & 1) == 0) gs += 4;
      o->gs[i] = gs;
    } else {
      // transposed: rows = o (Fo of them), columns = k; only k < Fi is ever needed (dE is
      // not propagated: edge features are inputs)
      o->pw[i] = tf_up(Fi, TF_NB);
      int gs = tf_up(Fo, 4) * o->pw[i];
      if (((gs >> 2) & 1) == 0) gs += 4;
      o->gs[i] = gs;
    }
    o->w[i] = off;
    off += o->gs[i] * L.D;
    o->pr[i] = tf_up(L.no, TF_NB);
    o->r[i] = off;
    off += tf_up(Fo, 4) * o->pr[i];
    o->rt[i] = 0;
    o->prt[i] = tf_up(Fo, TF_NB);
    if (backward) {
      o->rt[i] = off;
      off += tf_up(L.no, 4) * o->prt[i];
    }
  }
  o->pF = tf_pitch(Fmax);
  o->pK = tf_pitch(Kmax);
  o->pO = tf_pitch(L.no);
  o->pG = tf_pitch(std::max(L.no, Fmax));
  o->pE = tf_pitch(L.nef);
  off = tf_up(off, 4);
  if (!backward) {
    // buf0: step input, buf1: A, buf2: z_t, buf3: readout
    const int sz[4] = {TILE_ROWS * o->pF, TILE_ROWS * o->pK, TILE_ROWS * o->pF, TILE_ROWS * o->pO};
    for (int k = 0; k < 4; ++k) {
      o->buf[k] = off;
      off += sz[k];
    }
    o->buf[4] = 0;
  } else {
    // buf0..2: rotating z_t / z_{t-1} / carry, buf3: A, buf4: readout + dA
    const int sz[5] = {TILE_ROWS * o->pF, TILE_ROWS * o->pF, TILE_ROWS * o->pF, TILE_ROWS * o->pK,
                       TILE_ROWS * o->pG};
    for (int k = 0; k < 5; ++k) {
      o->buf[k] = off;
      off += sz[k];
    }
  }
  o->ae = off;
  off += TILE_ROWS * o->pE;
  o->ints = off;
  off += 132 + 132 + 260 + 128 + 16;
  off = tf_up(off, 4);
  o->bytes = off;
  o->total_bytes = off * 4 + 2 * TF_IDX + 128 + 128 + 16;
}

// per-tile structure shared by both Duvenaud kernels
struct TileView {
  int r0, rows, e0, ents;
  int g_begin, g_end;  // graphs whose vertices lie in this tile (empty graphs included)
};

__device__ __forceinline__ TileView tf_tile(const int4* tiles, int j, int num_tiles,
                                            const int32_t* vgraph, int num_graphs) {
  const int4 ti = __ldg(tiles + j);
  TileView tv;
  tv.r0 = ti.x;
  tv.rows = ti.y;
  tv.e0 = ti.z;
  tv.ents = ti.w;
  tv.g_begin = j == 0 ? 0 : __ldg(vgraph + ti.x);
  tv.g_end = j + 1 == num_tiles ? num_graphs : __ldg(vgraph + __ldg(tiles + j + 1).x);
  return tv;
}

// Ae[v][:] = sum_w E(:, ja(2,w))  (time-step invariant part of duvenaud_propagate)
__device__ __forceinline__ void tf_edge_sum(float* ae, int pe, const float* __restrict__ E, int Fe,
                                            const int32_t* __restrict__ eid, const int* ptr_s,
                                            int e0, int rows) {
  for (int i = threadIdx.x; i < rows * Fe; i += TF_THREADS) {
    const int v = i / Fe, f = i - v * Fe;
    float s = 0.f;
    for (int e = ptr_s[v]; e < ptr_s[v + 1]; ++e) {
      const int id = __ldg(eid + e0 + e);
      if (id >= 0) s += __ldg(E + static_cast<size_t>(id) * Fe + f);
    }
    ae[v * pe + f] = s;
  }
}

// A[v][Fi .. Fi+Fe) = Ae[v][:] / d ; A[v][K .. 4*ceil(K/4)) = 0
__device__ __forceinline__ void tf_append_edges(float* A, int pa, int Fi, int Fe, const float* ae,
                                                int pe, const uint8_t* bkt_s, int rows) {
  const int K = Fi + Fe, Kp = ((K + 3) >> 2) << 2;
  const int w = Kp - Fi;
  if (w == 0) return;
  for (int i = threadIdx.x; i < rows * w; i += TF_THREADS) {
    const int v = i / w, j = i - v * w;
    A[v * pa + Fi + j] = j < Fe ? ae[v * pe + j] / static_cast<float>(bkt_s[v] + 1) : 0.f;
  }
}

__device__ __forceinline__ void duv_stage_weights(float* sm, const DuvArgs& a, bool backward) {
  const int D = a.D;
  for (int t = 1; t <= a.T; ++t) {
    const int i = t - 1;
    const int Fi = a.nvf[t - 1], Fo = a.nvf[t], K = Fi + a.nef;
    const float* Wg = a.params + a.woff[i];
    const float* Rg = a.params + a.roff[i];
    float* ws = sm + a.lay.w[i];
    const int pw = a.lay.pw[i], gs = a.lay.gs[i];
    if (!backward) {
      for (int idx = threadIdx.x; idx < D * gs; idx += TF_THREADS) {
        const int d = idx / gs, rem = idx - d * gs;
        const int k = rem / pw, n = rem - k * pw;
        ws[idx] = (k < K && n < Fo) ? __ldg(Wg + (static_cast<size_t>(d) * K + k) * Fo + n) : 0.f;
      }
    } else {
      for (int idx = threadIdx.x; idx < D * gs; idx += TF_THREADS) {
        const int d = idx / gs, rem = idx - d * gs;
        const int o = rem / pw, k = rem - o * pw;
        ws[idx] = (o < Fo && k < Fi) ? __ldg(Wg + (static_cast<size_t>(d) * K + k) * Fo + o) : 0.f;
      }
    }
    float* rs = sm + a.lay.r[i];
    const int pr = a.lay.pr[i];
    for (int idx = threadIdx.x; idx < tf_up(Fo, 4) * pr; idx += TF_THREADS) {
      const int f = idx / pr, n = idx - f * pr;
      rs[idx] = (f < Fo && n < a.no) ? __ldg(Rg + static_cast<size_t>(f) * a.no + n) : 0.f;
    }
    if (backward) {
      float* rts = sm + a.lay.rt[i];
      const int prt = a.lay.prt[i];
      for (int idx = threadIdx.x; idx < tf_up(a.no, 4) * prt; idx += TF_THREADS) {
        const int n = idx / prt, f = idx - n * prt;
        rts[idx] = (n < a.no && f < Fo) ? __ldg(Rg + static_cast<size_t>(f) * a.no + n) : 0.f;
      }
    }
  }
}

__global__ void __launch_bounds__(TF_THREADS, 1) k_duv_fwd(const DuvArgs a) {
  extern __shared__ float4 tf_smem4[];
  float* sm = reinterpret_cast<float*>(tf_smem4);
  const DuvLayout& L = a.lay;
  int* ptr_s = reinterpret_cast<int*>(sm + L.ints);
  float* red = reinterpret_cast<float*>(ptr_s + 132 + 132 + 260 + 128);
  uint8_t* idx_s = reinterpret_cast<uint8_t*>(sm + L.bytes);
  uint8_t* bkt_s = idx_s + 2 * TF_IDX;
  float* ae = sm + L.ae;
  duv_stage_weights(sm, a, false);
  float lsum = 0.f;
  for (int j = blockIdx.x; j < a.num_tiles; j += gridDim.x) {
    const TileView tv = tf_tile(a.tiles, j, a.num_tiles, a.vgraph, a.num_graphs);
    __syncthreads();  // the previous tile is done with every buffer (and the weights are staged)
    tf_load_struct(ptr_s, idx_s, a.row_ptr, a.col8, tv.r0, tv.rows, tv.e0, tv.ents);
    for (int v = threadIdx.x; v < tv.rows; v += TF_THREADS) {
      const int deg = __ldg(a.row_ptr + tv.r0 + v + 1) - __ldg(a.row_ptr + tv.r0 + v);
      bkt_s[v] = static_cast<uint8_t>(max(a.min_deg, min(deg, a.max_deg)) - a.min_deg);
    }
    float* xin = sm + L.buf[0];
    float* zout = sm + L.buf[2];
    tf_load_rows(xin, L.pF, a.X + static_cast<size_t>(tv.r0) * a.nvf[0], tv.rows, a.nvf[0]);
    __syncthreads();
    if (a.nef > 0) tf_edge_sum(ae, L.pE, a.E, a.nef, a.eid, ptr_s, tv.e0, tv.rows);
    for (int t = 1; t <= a.T; ++t) {
      const int i = t - 1;
      const int Fi = a.nvf[t - 1], Fo = a.nvf[t], K = Fi + a.nef;
      float* A = sm + L.buf[1];
      float* Y = sm + L.buf[3];
      // A = [ sum_w in(:,ja(1,w)) ; sum_w E(:,ja(2,w)) ] / d          (propagate; the division
      // belongs to duvenaud_update)
      tf_gather<false>(A, L.pK, xin, L.pF, (Fi + 3) >> 2, ptr_s, idx_s, nullptr, tv.rows, bkt_s);
      __syncthreads();  // Ae complete (first step); gather done with the pad columns
      tf_append_edges(A, L.pK, Fi, a.nef, ae, L.pE, bkt_s, tv.rows);
      __syncthreads();
      // z = act( W_d(v) . A(:,v) )
      const int act = a.act;
      tf_gemm(zout, L.pF, A, L.pK, K, sm + L.w[i], L.pw[i], L.gs[i], bkt_s, tv.rows, Fo,
              [act](int, int, float s) { return tf_act(act, s); });
      __syncthreads();
      tf_store_rows(a.Z[i] + static_cast<size_t>(tv.r0) * Fo, zout, L.pF, tv.rows, Fo);
      // readout: S = ract( R_t . z )
      const int ract = a.ract == ATHENA_ACT_SOFTMAX ? ATHENA_ACT_NONE : a.ract;
      tf_gemm(Y, L.pO, zout, L.pF, Fo, sm + L.r[i], L.pr[i], 0, nullptr, tv.rows, a.no,
              [ract](int, int, float s) { return tf_act(ract, s); });
      __syncthreads();
      if (a.ract == ATHENA_ACT_SOFTMAX) {
        tf_softmax_rows(Y, L.pO, tv.rows, a.no);
        __syncthreads();
      }
      // out(:,s) (+)= sum_v S(:,v), vertices ascending (sum(ptr2, dim=2), :848-852)
      const int ng = tv.g_end - tv.g_begin;
      for (int idx = threadIdx.x; idx < ng * a.no; idx += TF_THREADS) {
        const int gl = idx / a.no, o = idx - gl * a.no;
        const int g = tv.g_begin + gl;
        const int v0 = __ldg(a.voff + g) - tv.r0, v1 = __ldg(a.voff + g + 1) - tv.r0;
        float s = 0.f;
        for (int v = v0; v < v1; ++v) s += Y[v * L.pO + o];
        float* dst = a.out + static_cast<size_t>(g) * a.no + o;
        const float tot = t == 1 ? s : *dst + s;
        *dst = tot;
        if (t == a.T && a.target != nullptr) {
          // one [num_outputs, batch] MSE cell: mean over num_outputs * global batch, / 2
          const float d = tot - __ldg(a.target + static_cast<size_t>(g) * a.no + o);
          a.mse_grad[static_cast<size_t>(g) * a.no + o] = d / a.mse_denom;
          lsum += d * d / a.mse_denom;
        }
      }
      float* tmp = xin;
      xin = zout;
      zout = tmp;
      // (the next gather reads `xin`, which only the product above wrote: ordered by the
      //  barrier after the readout product)
    }
  }
  if (a.loss_part != nullptr) {
    __syncthreads();
    const float tot = tf_block_sum(lsum, red);
    if (threadIdx.x == 0) a.loss_part[blockIdx.x] = tot;
  }
}

__global__ void __launch_bounds__(TF_THREADS, 1) k_duv_bwd(const DuvArgs a) {
  extern __shared__ float4 tf_smem4[];
  float* sm = reinterpret_cast<float*>(tf_smem4);
  const DuvLayout& L = a.lay;
  int* ptr_s = reinterpret_cast<int*>(sm + L.ints);
  int* cptr_s = ptr_s + 132;
  int* seg_s = cptr_s + 132;
  int* vg_s = seg_s + 260;
  uint8_t* idx_s = reinterpret_cast<uint8_t*>(sm + L.bytes);
  uint8_t* cidx_s = idx_s + TF_IDX;
  uint8_t* bkt_s = cidx_s + TF_IDX;
  uint8_t* list_s = bkt_s + 128;
  float* ae = sm + L.ae;
  float* part = a.part + static_cast<size_t>(blockIdx.x) * a.np;
  for (int i = threadIdx.x; i < a.np; i += TF_THREADS) part[i] = 0.f;
  duv_stage_weights(sm, a, true);
  const int T = a.T;
  for (int j = blockIdx.x; j < a.num_tiles; j += gridDim.x) {
    const TileView tv = tf_tile(a.tiles, j, a.num_tiles, a.vgraph, a.num_graphs);
    __syncthreads();
    tf_load_struct(ptr_s, idx_s, a.row_ptr, a.col8, tv.r0, tv.rows, tv.e0, tv.ents);
    tf_load_struct(cptr_s, cidx_s, a.csc_ptr, a.csc8, tv.r0, tv.rows, tv.e0, tv.ents);
    for (int v = threadIdx.x; v < tv.rows; v += TF_THREADS) {
      const int deg = __ldg(a.row_ptr + tv.r0 + v + 1) - __ldg(a.row_ptr + tv.r0 + v);
      bkt_s[v] = static_cast<uint8_t>(max(a.min_deg, min(deg, a.max_deg)) - a.min_deg);
      vg_s[v] = __ldg(a.vgraph + tv.r0 + v);
    }
    float* zt = sm + L.buf[0];
    float* zp = sm + L.buf[1];
    float* carry = sm + L.buf[2];
    float* A = sm + L.buf[3];
    float* G = sm + L.buf[4];
    tf_load_rows(zt, L.pF, a.Z[T - 1] + static_cast<size_t>(tv.r0) * a.nvf[T], tv.rows, a.nvf[T]);
    {
      const float* prev = T >= 2 ? a.Z[T - 2] : a.X;
      tf_load_rows(zp, L.pF, prev + static_cast<size_t>(tv.r0) * a.nvf[T - 1], tv.rows, a.nvf[T - 1]);
    }
    __syncthreads();
    // vertices of the tile grouped by degree bucket (ascending vertex inside a bucket)
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      int base = 0;
      for (int d = 0; d < a.D; ++d) {
        if (lane == 0) seg_s[d] = base;
        for (int c = 0; c < TILE_ROWS / 32; ++c) {
          const int v = c * 32 + lane;
          const bool m = v < tv.rows && bkt_s[v] == d;
          const unsigned bal = __ballot_sync(0xffffffffu, m);
          if (m) list_s[base + __popc(bal & ((1u << lane) - 1u))] = static_cast<uint8_t>(v);
          base += __popc(bal);
        }
      }
      if (lane == 0) seg_s[a.D] = base;
    }
    if (a.nef > 0) tf_edge_sum(ae, L.pE, a.E, a.nef, a.eid, ptr_s, tv.e0, tv.rows);
    __syncthreads();
    for (int t = T; t >= 1; --t) {
      const int i = t - 1;
      const int Fi = a.nvf[t - 1], Fo = a.nvf[t], K = Fi + a.nef, no = a.no;
      // 1. S = ract(R_t z_t), then dY in place (upstream row = gout of the vertex's graph)
      {
        const int ract = a.ract == ATHENA_ACT_SOFTMAX ? ATHENA_ACT_NONE : a.ract;
        tf_gemm(G, L.pG, zt, L.pF, Fo, sm + L.r[i], L.pr[i], 0, nullptr, tv.rows, no,
                [ract](int, int, float s) { return tf_act(ract, s); });
      }
      // 4 (independent of 1-3). A = [gather(z_{t-1}) ; Ae] / d, recomputed
      tf_gather<false>(A, L.pK, zp, L.pF, (Fi + 3) >> 2, ptr_s, idx_s, nullptr, tv.rows, bkt_s);
      __syncthreads();
      tf_append_edges(A, L.pK, Fi, a.nef, ae, L.pE, bkt_s, tv.rows);
      if (a.ract == ATHENA_ACT_SOFTMAX) {
        tf_softmax_rows(G, L.pG, tv.rows, no);
        __syncthreads();
      }
      {
        const float* gout = a.gout;
        const int* vgs = vg_s;
        tf_act_bwd_rows(a.ract, G, L.pG, tv.rows, no,
                        [gout, vgs, no](int v) { return gout + static_cast<size_t>(vgs[v]) * no; });
      }
      __syncthreads();
      // 2. dR_t(o,f) += sum_v dY(o,v) z_t(f,v)
      if (no <= 32)
        tf_outer<1>(part + a.roff[i], zt, L.pF, Fo, G, L.pG, no, nullptr, nullptr, 1, tv.rows);
      else
        tf_outer<4>(part + a.roff[i], zt, L.pF, Fo, G, L.pG, no, nullptr, nullptr, 1, tv.rows);
      // 3. gz = ( R_t^T dY + carry ) .* act'(z_t), in place in `carry`
      {
        const int act = a.act;
        const bool has_carry = t < T;
        const float* ztc = zt;
        const float* cc = carry;
        const int pF = L.pF;
        tf_gemm(carry, L.pF, G, L.pG, no, sm + L.rt[i], L.prt[i], 0, nullptr, tv.rows, Fo,
                [act, has_carry, ztc, cc, pF](int v, int n, float s) {
                  const float dz = has_carry ? s + cc[v * pF + n] : s;
                  return tf_act_grad(act, ztc[v * pF + n], dz);
                });
      }
      __syncthreads();
      // 5. dW_{t,d}(o,k) += sum_{v in bucket d} gz(o,v) A(k,v)      (A already divided by d)
      if (Fo <= 32)
        tf_outer<1>(part + a.woff[i], A, L.pK, K, carry, L.pF, Fo, list_s, seg_s, a.D, tv.rows);
      else
        tf_outer<4>(part + a.woff[i], A, L.pK, K, carry, L.pF, Fo, list_s, seg_s, a.D, tv.rows);
      const bool need_dx = t > 1 || a.gin != nullptr;
      if (need_dx) {
        // 6. dA(k,v) = ( W_d^T gz(:,v) )(k) / d for k < Fi, into G (dY is dead)
        const uint8_t* bk = bkt_s;
        tf_gemm(G, L.pG, carry, L.pF, Fo, sm + L.w[i], L.pw[i], L.gs[i], bkt_s, tv.rows, Fi,
                [bk](int v, int, float s) { return s / static_cast<float>(bk[v] + 1); });
      }
      __syncthreads();
      if (need_dx) {
        // 7. d in(:,u) = sum over the CSC column of u of dA(1:Fi, v): into the z_t buffer
        tf_gather<false>(zt, L.pF, G, L.pG, (Fi + 3) >> 2, cptr_s, cidx_s, nullptr, tv.rows,
                         nullptr);
        __syncthreads();
        if (t == 1)
          tf_store_rows(a.gin + static_cast<size_t>(tv.r0) * Fi, zt, L.pF, tv.rows, Fi);
      }
      if (t > 1) {
        // rotate: z_{t-1} becomes the current step, the new carry sits in the old z_t buffer,
        // the old carry buffer receives z_{t-2}
        float* old_carry = carry;
        carry = zt;
        zt = zp;
        zp = old_carry;
        const float* prev = t >= 3 ? a.Z[t - 3] : a.X;
        tf_load_rows(zp, L.pF, prev + static_cast<size_t>(tv.r0) * a.nvf[t - 2], tv.rows,
                     a.nvf[t - 2]);
        __syncthreads();
      }
    }
  }
}

// ======================================================================================
// Kipf step
// ======================================================================================

struct KipfArgs {
  const int4* tiles;
  int num_tiles;
  const int32_t* ptr;    // row_ptr (forward) / csc_ptr (backward)
  const uint8_t* idx8;   // col8 / csc8
  const float* coef;     // [Z] forward coefficients (nullptr: 1)
  const float* X;        // forward: [V][Fi] input; backward: [V][Fo] upstream gradient
  const float* W;        // [Fo, Fi] column-major
  float* P;              // forward: saved aggregate (nullable); backward: the saved aggregate
  float* out;            // forward: [V][Fo]; backward: [V][Fi] input gradient (nullable)
  const float* H;        // backward: saved output of this step (act'), nullptr: gradient is
                         // already w.r.t. the pre-activation
  float* part;           // backward: [grid][Fi*Fo]
  int Fi, Fo, act;
  int pI, pO, pw;        // pitches: input-width tiles, output-width tiles, weight rows
  int offW, offA, offB, offC, offInts, offBytes;
};

__global__ void __launch_bounds__(TF_THREADS, 1) k_kipf_fwd(const KipfArgs a) {
  extern __shared__ float4 tf_smem4[];
  float* sm = reinterpret_cast<float*>(tf_smem4);
  float* Ws = sm + a.offW;
  float* xin = sm + a.offA;   // [128][pI]
  float* Pb = sm + a.offB;    // [128][pI]
  float* Hb = sm + a.offC;    // [128][pO]
  int* ptr_s = reinterpret_cast<int*>(sm + a.offInts);
  float* coef_s = reinterpret_cast<float*>(ptr_s + 132);
  uint8_t* idx_s = reinterpret_cast<uint8_t*>(sm + a.offBytes);
  const int Fi = a.Fi, Fo = a.Fo;
  // W_t [Fo, Fi] column-major = [k = i][n = o] rows
  for (int idx = threadIdx.x; idx < tf_up(Fi, 4) * a.pw; idx += TF_THREADS) {
    const int k = idx / a.pw, n = idx - k * a.pw;
    Ws[idx] = (k < Fi && n < Fo) ? __ldg(a.W + static_cast<size_t>(k) * Fo + n) : 0.f;
  }
  for (int j = blockIdx.x; j < a.num_tiles; j += gridDim.x) {
    const int4 ti = __ldg(a.tiles + j);
    const int r0 = ti.x, rows = ti.y, e0 = ti.z, ents = ti.w;
    __syncthreads();
    tf_load_struct(ptr_s, idx_s, a.ptr, a.idx8, r0, rows, e0, ents);
    if (a.coef != nullptr)
      for (int e = threadIdx.x; e < ents; e += TF_THREADS) coef_s[e] = __ldg(a.coef + e0 + e);
    tf_load_rows(xin, a.pI, a.X + static_cast<size_t>(r0) * Fi, rows, Fi);
    __syncthreads();
    // P(:,v) = sum_w (deg_v deg_u)^-1/2 X(:,u)
    if (a.coef != nullptr)
      tf_gather<true>(Pb, a.pI, xin, a.pI, (Fi + 3) >> 2, ptr_s, idx_s, coef_s, rows, nullptr);
    else
      tf_gather<false>(Pb, a.pI, xin, a.pI, (Fi + 3) >> 2, ptr_s, idx_s, nullptr, rows, nullptr);
    __syncthreads();
    if (a.P != nullptr) tf_store_rows(a.P + static_cast<size_t>(r0) * Fi, Pb, a.pI, rows, Fi);
    const int act = a.act == ATHENA_ACT_SOFTMAX ? ATHENA_ACT_NONE : a.act;
    tf_gemm(Hb, a.pO, Pb, a.pI, Fi, Ws, a.pw, 0, nullptr, rows, Fo,
            [act](int, int, float s) { return tf_act(act, s); });
    __syncthreads();
    if (a.act == ATHENA_ACT_SOFTMAX) {
      tf_softmax_rows(Hb, a.pO, rows, Fo);
      __syncthreads();
    }
    tf_store_rows(a.out + static_cast<size_t>(r0) * Fo, Hb, a.pO, rows, Fo);
  }
}

__global__ void __launch_bounds__(TF_THREADS, 1) k_kipf_bwd(const KipfArgs a) {
  extern __shared__ float4 tf_smem4[];
  float* sm = reinterpret_cast<float*>(tf_smem4);
  float* WTs = sm + a.offW;   // [o][pw], k fastest
  float* Gb = sm + a.offA;    // [128][pO]  gradient -> gY
  float* Pb = sm + a.offB;    // [128][pI]  saved aggregate, later the gathered result
  float* Db = sm + a.offC;    // [128][max(pI, pO)]  H tile, later dP
  int* ptr_s = reinterpret_cast<int*>(sm + a.offInts);
  uint8_t* idx_s = reinterpret_cast<uint8_t*>(sm + a.offBytes);
  const int Fi = a.Fi, Fo = a.Fo;
  const int pD = a.pI > a.pO ? a.pI : a.pO;
  float* part = a.part + static_cast<size_t>(blockIdx.x) * Fi * Fo;
  for (int i = threadIdx.x; i < Fi * Fo; i += TF_THREADS) part[i] = 0.f;
  for (int idx = threadIdx.x; idx < tf_up(Fo, 4) * a.pw; idx += TF_THREADS) {
    const int o = idx / a.pw, k = idx - o * a.pw;
    WTs[idx] = (o < Fo && k < Fi) ? __ldg(a.W + static_cast<size_t>(k) * Fo + o) : 0.f;
  }
  const bool need_act = a.H != nullptr;
  for (int j = blockIdx.x; j < a.num_tiles; j += gridDim.x) {
    const int4 ti = __ldg(a.tiles + j);
    const int r0 = ti.x, rows = ti.y, e0 = ti.z, ents = ti.w;
    __syncthreads();
    if (a.out != nullptr) tf_load_struct(ptr_s, idx_s, a.ptr, a.idx8, r0, rows, e0, ents);
    tf_load_rows(Gb, a.pO, a.X + static_cast<size_t>(r0) * Fo, rows, Fo);
    tf_load_rows(Pb, a.pI, a.P + static_cast<size_t>(r0) * Fi, rows, Fi);
    if (need_act) tf_load_rows(Db, pD, a.H + static_cast<size_t>(r0) * Fo, rows, Fo);
    __syncthreads();
    if (need_act) {
      // gY = gH .* act'(H)   (softmax: per-vertex Jacobian)
      const float* gb = Gb;
      const int pO = a.pO;
      // the result must land in Gb: compute into Db (which holds H), then swap roles
      tf_act_bwd_rows(a.act, Db, pD, rows, Fo, [gb, pO](int v) { return gb + v * pO; });
      __syncthreads();
    }
    const float* gy = need_act ? Db : Gb;
    const int pgy = need_act ? pD : a.pO;
    // dW_t(o,i) += sum_v gY(o,v) P(i,v)
    if (Fo <= 32)
      tf_outer<1>(part, Pb, a.pI, Fi, gy, pgy, Fo, nullptr, nullptr, 1, rows);
    else
      tf_outer<4>(part, Pb, a.pI, Fi, gy, pgy, Fo, nullptr, nullptr, 1, rows);
    if (a.out != nullptr) {
      // dP = W_t^T gY into the other gradient-width buffer, then the un-normalised scatter
      // dX(:,u) += dP(:,v) as a gather over the CSC (entries ascending)
      float* dP = need_act ? Gb : Db;
      const int pdp = need_act ? a.pO : pD;
      // dP is [rows][Fi]; Gb has pitch pO: large enough only if Fi <= Fo-padded; use a pitch
      // that fits both (the launcher sizes both buffers with max(pI, pO))
      __syncthreads();  // tf_outer reads Pb / gy; the product below only writes dP
      tf_gemm(dP, pdp, gy, pgy, Fo, WTs, a.pw, 0, nullptr, rows, Fi,
              [](int, int, float s) { return s; });
      __syncthreads();
      tf_gather<false>(Pb, a.pI, dP, pdp, (Fi + 3) >> 2, ptr_s, idx_s, nullptr, rows, nullptr);
      __syncthreads();
      tf_store_rows(a.out + static_cast<size_t>(r0) * Fi, Pb, a.pI, rows, Fi);
    }
  }
}

static bool tile_fma_enabled() {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("ATHENA_CUDA_DISABLE_TILE_FMA");
    off = (e && atoi(e) != 0) ? 1 : 0;
  }
  return off == 0;
}

template <class K>
static int tf_set_smem(K kernel, int bytes) {
  ATH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return ATHENA_OK;
}

struct KipfLayout {
  int pI, pO, pw, offW, offA, offB, offC, offInts, offBytes, total_bytes;
};
static KipfLayout kipf_layout(int Fi, int Fo, bool backward) {
  KipfLayout k;
  k.pI = tf_pitch(Fi);
  k.pO = tf_pitch(Fo);
  int off = 0;
  k.offW = 0;
  if (!backward) {
    k.pw = tf_up(Fo, TF_NB);
    off += tf_up(Fi, 4) * k.pw;
  } else {
    k.pw = tf_up(Fi, TF_NB);
    off += tf_up(Fo, 4) * k.pw;
  }
  off = tf_up(off, 4);
  const int pM = std::max(k.pI, k.pO);
  if (!backward) {
    k.offA = off; off += TILE_ROWS * k.pI;
    k.offB = off; off += TILE_ROWS * k.pI;
    k.offC = off; off += TILE_ROWS * k.pO;
  } else {
    // both gradient-width buffers are sized for max(pI, pO) (dP may land in either)
    k.pO = pM;
    k.offA = off; off += TILE_ROWS * pM;
    k.offB = off; off += TILE_ROWS * k.pI;
    k.offC = off; off += TILE_ROWS * pM;
  }
  k.offInts = off;
  off += 132 + (backward ? 0 : TILE_ENTRIES + 16);
  off = tf_up(off, 4);
  k.offBytes = off;
  k.total_bytes = off * 4 + TF_IDX + 16;
  return k;
}

}  // namespace

// ---- host side ---------------------------------------------------------------------

bool tile_kipf_supported(const Batch* b, int Fi, int Fo) {
  if (!tile_fma_enabled() || b->num_tiles == 0 || b->col8 == nullptr) return false;
  if (Fi < 1 || Fo < 1 || Fi > 128 || Fo > 128) return false;
  const size_t lim = ctx().max_smem_optin;
  return (size_t)kipf_layout(Fi, Fo, false).total_bytes <= lim &&
         (size_t)kipf_layout(Fi, Fo, true).total_bytes <= lim;
}

int launch_tile_kipf_fwd(const Batch* b, const float* X, const float* W, float* P, float* out,
                         int Fi, int Fo, int act) {
  const KipfLayout k = kipf_layout(Fi, Fo, false);
  static int smem_set = 0;
  if (k.total_bytes > smem_set) {
    ATH_TRY(tf_set_smem(k_kipf_fwd, k.total_bytes));
    smem_set = k.total_bytes;
  }
  KipfArgs a{};
  a.tiles = b->tiles.as<int4>();
  a.num_tiles = b->num_tiles;
  a.ptr = b->row_ptr;
  a.idx8 = b->col8;
  a.coef = b->coef;
  a.X = X;
  a.W = W;
  a.P = P;
  a.out = out;
  a.Fi = Fi;
  a.Fo = Fo;
  a.act = act;
  a.pI = k.pI; a.pO = k.pO; a.pw = k.pw;
  a.offW = k.offW; a.offA = k.offA; a.offB = k.offB; a.offC = k.offC;
  a.offInts = k.offInts; a.offBytes = k.offBytes;
  const int grid = std::min(b->num_tiles, ctx().sm_count);
  k_kipf_fwd<<<grid, TF_THREADS, k.total_bytes, ctx().stream>>>(a);
  ATH_LAUNCHED_T("tile_kipf_fwd");
  return ATHENA_OK;
}

// G: gradient w.r.t. the step output (H != nullptr: act'(H) is applied here) or already
// w.r.t. the pre-activation (H == nullptr).  part: [*nparts][Fi*Fo] CTA partials of dW.
int launch_tile_kipf_bwd(const Batch* b, const float* G, const float* H, const float* P,
                         const float* W, float* gin, int Fi, int Fo, int act, DevBuf& part,
                         int* nparts) {
  const KipfLayout k = kipf_layout(Fi, Fo, true);
  static int smem_set = 0;
  if (k.total_bytes > smem_set) {
    ATH_TRY(tf_set_smem(k_kipf_bwd, k.total_bytes));
    smem_set = k.total_bytes;
  }
  const int grid = std::min(b->num_tiles, ctx().sm_count);
  ATH_TRY(part.reserve(sizeof(float) * (size_t)grid * Fi * Fo));
  KipfArgs a{};
  a.tiles = b->tiles.as<int4>();
  a.num_tiles = b->num_tiles;
  a.ptr = b->csc_ptr;
  a.idx8 = b->csc8;
  a.coef = nullptr;
  a.X = G;
  a.W = W;
  a.P = const_cast<float*>(P);
  a.out = gin;
  a.H = H;
  a.part = part.as<float>();
  a.Fi = Fi;
  a.Fo = Fo;
  a.act = act;
  a.pI = k.pI; a.pO = k.pO; a.pw = k.pw;
  a.offW = k.offW; a.offA = k.offA; a.offB = k.offB; a.offC = k.offC;
  a.offInts = k.offInts; a.offBytes = k.offBytes;
  k_kipf_bwd<<<grid, TF_THREADS, k.total_bytes, ctx().stream>>>(a);
  ATH_LAUNCHED_T("tile_kipf_bwd");
  *nparts = grid;
  return ATHENA_OK;
}

static bool duv_dims_ok(int T, const int* nvf, int nef, int D, int no) {
  if (T < 1 || T > TF_MAX_T || D < 1 || D > 255 || no < 1 || no > 128 || nef < 0 || nef > 64)
    return false;
  for (int t = 0; t <= T; ++t)
    if (nvf[t] < 1 || nvf[t] > 128) return false;
  return true;
}

bool tile_duv_supported(const Batch* b, int T, const int* nvf, int nef, int D, int no) {
  if (!tile_fma_enabled() || b->num_tiles == 0 || b->col8 == nullptr) return false;
  if (!duv_dims_ok(T, nvf, nef, D, no)) return false;
  Layer_dims_t L{T, nef, D, no, nvf};
  DuvLayout f, r;
  duv_layout(L, false, &f);
  duv_layout(L, true, &r);
  const size_t lim = ctx().max_smem_optin;
  return (size_t)f.total_bytes <= lim && (size_t)r.total_bytes <= lim;
}

static void duv_fill(DuvArgs* a, const Batch* b, const TileDuvDesc& d) {
  a->tiles = b->tiles.as<int4>();
  a->num_tiles = b->num_tiles;
  a->num_graphs = b->B;
  a->row_ptr = b->row_ptr;
  a->col8 = b->col8;
  a->eid = b->eid;
  a->csc_ptr = b->csc_ptr;
  a->csc8 = b->csc8;
  a->vgraph = b->vgraph;
  a->voff = b->voff;
  a->X = d.X;
  a->E = d.E;
  a->params = d.params;
  a->T = d.T;
  a->nef = d.nef;
  a->D = d.max_deg - d.min_deg + 1;
  a->min_deg = d.min_deg;
  a->max_deg = d.max_deg;
  a->no = d.no;
  a->act = d.act;
  a->ract = d.ract;
  for (int t = 0; t <= d.T; ++t) a->nvf[t] = d.nvf[t];
  for (int t = 0; t < d.T; ++t) {
    a->Z[t] = d.Z[t];
    a->woff[t] = (int)d.poff[t];
    a->roff[t] = (int)d.poff[d.T + t];
  }
}

int launch_tile_duv_fwd(const Batch* b, const TileDuvDesc& d, float* out, const float* target,
                        float* mse_grad, float mse_denom, float* loss_part, int* num_parts) {
  DuvArgs a{};
  duv_fill(&a, b, d);
  Layer_dims_t L{d.T, d.nef, a.D, d.no, d.nvf};
  duv_layout(L, false, &a.lay);
  static int smem_set = 0;
  if (a.lay.total_bytes > smem_set) {
    ATH_TRY(tf_set_smem(k_duv_fwd, a.lay.total_bytes));
    smem_set = a.lay.total_bytes;
  }
  a.out = out;
  a.target = target;
  a.mse_grad = mse_grad;
  a.mse_denom = mse_denom;
  a.loss_part = target != nullptr ? loss_part : nullptr;
  const int grid = std::min(b->num_tiles, ctx().sm_count);
  k_duv_fwd<<<grid, TF_THREADS, a.lay.total_bytes, ctx().stream>>>(a);
  ATH_LAUNCHED_T(target != nullptr ? "tile_duv_fwd_mse" : "tile_duv_fwd");
  if (num_parts) *num_parts = grid;
  return ATHENA_OK;
}

int launch_tile_duv_bwd(const Batch* b, const TileDuvDesc& d, const float* gout, float* gin,
                        int64_t num_params, DevBuf& part, int* nparts) {
  DuvArgs a{};
  duv_fill(&a, b, d);
  Layer_dims_t L{d.T, d.nef, a.D, d.no, d.nvf};
  duv_layout(L, true, &a.lay);
  static int smem_set = 0;
  if (a.lay.total_bytes > smem_set) {
    ATH_TRY(tf_set_smem(k_duv_bwd, a.lay.total_bytes));
    smem_set = a.lay.total_bytes;
  }
  const int grid = std::min(b->num_tiles, ctx().sm_count);
  ATH_TRY(part.reserve(sizeof(float) * (size_t)grid * (size_t)num_params));
  a.gout = gout;
  a.gin = gin;
  a.part = part.as<float>();
  a.np = (int)num_params;
  k_duv_bwd<<<grid, TF_THREADS, a.lay.total_bytes, ctx().stream>>>(a);
  ATH_LAUNCHED_T("tile_duv_bwd");
  *nparts = grid;
  return ATHENA_OK;
}

}  // namespace athena
