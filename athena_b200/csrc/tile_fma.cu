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
//               recomputing A_t from the saved z_t and reading the readouts S_t back (the
//               kernels are bound by shared-memory traffic, not by HBM: DESIGN.md 3.2); the
//               activation derivative of the layer that produced the input rides on the
//               input-gradient store
//   k_kipf_fwd  kipf_propagate + matmul + activation   (_sub_kipf.f90:29-46,
//               athena_kipf_msgpass_layer.f90:943-952), any widths up to 128
//   k_kipf_bwd  act', dW = gY^T P, dP = gY W^T, un-normalised CSC scatter (_sub_kipf.f90:101-109)
//
// Execution shape: one persistent CTA per SM holding ONE copy of the layer's weights in shared
// memory and up to two independent groups of 512 threads, each walking its own tiles with
// its own three rotating [128 x pitch] buffers and its own named barrier -- two tiles in
// flight per SM hide each other's barrier and load latencies.
//
// Arithmetic: dense products run on FFMA2 (fma.rn.f32x2: two fp32 FMAs per instruction, the
// only way to the full FP32 rate of sm_100), one vertex row per lane with the row's weight
// block read as broadcast LDS.128; per-bucket weight blocks are offset by an odd number of
// 16-byte chunks so that lanes of different buckets hit different banks.  The same [k][n]
// weight block serves Y = A.W (n-blocked, scalar x pair FFMA2) and dA = G.W^T (row-blocked,
// pair x pair FFMA2 over the contraction index).  The rows of a tile are grouped by degree
// bucket (tf_bucket_list) wherever a per-bucket weight block is read, so that the lanes of a
// warp read ONE block and every weight load is a broadcast.  Sums over CSR entries run in the
// reference's order; weight gradients are accumulated per group in a private partial
// vector (no atomics) that k_finalize folds in a fixed order.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "athena_internal.h"

namespace athena {

namespace {

#ifdef TF_TRACE  // debugging build: clock64 stamps of block 0 / thread 0 at the phase boundaries
#define TFT()                                                              \
  do {                                                                     \
    if (threadIdx.x == 0 && blockIdx.x == 0 && tft_n < 96) tft_m[tft_n++] = clock64(); \
  } while (0)
#define TFT_DECL() long long tft_m[96]; int tft_n = 0; TFT()
#define TFT_DUMP(name)                                                     \
  do {                                                                     \
    if (threadIdx.x == 0 && blockIdx.x == 0) {                             \
      printf("%s:", name);                                                 \
      for (int q = 1; q < tft_n; ++q) printf(" %lld", tft_m[q] - tft_m[q - 1]); \
      printf(" | total %lld\n", tft_m[tft_n - 1] - tft_m[0]);              \
    }                                                                      \
  } while (0)
#else
#define TFT() do { } while (0)
#define TFT_DECL() do { } while (0)
#define TFT_DUMP(name) do { } while (0)
#endif

constexpr int TF_GROUP = 512;                   // threads working on one tile
constexpr int TF_GWARPS = TF_GROUP / 32;
constexpr int TF_MAXG = 2;                      // groups per CTA
constexpr int TF_LPR = TF_GROUP / TILE_ROWS;    // lanes per tile row in the row-wise passes
constexpr int TF_MAX_T = 16;
constexpr int TF_NB = 8;                        // outputs per thread in the products
constexpr int TF_IDX = TILE_ENTRIES + 16;       // tile-local neighbour bytes
constexpr int TF_OUTS = 2048;                   // per-group [graphs x outputs] readout sums
constexpr int TF_INTS = 132 + 132 + 260 + 128 + 16 + 128;  // ptr, cptr, seg, vg, red, rd

// row pitch (floats) of a shared-memory tile: a multiple of 4 whose quarter is odd, so that
// row-per-lane LDS.128 / STS.128 of a quarter warp cover all 32 banks
__host__ __device__ __forceinline__ int tf_pitch(int n) {
  int p = (n + 3) & ~3;
  if (p == 0) p = 4;
  if (((p >> 2) & 1) == 0) p += 4;
  return p;
}
__host__ __device__ __forceinline__ int tf_up(int n, int m) { return (n + m - 1) / m * m; }

// barrier of one 512-thread group (ids 1 and 2; id 0 is __syncthreads)
__device__ __forceinline__ void tf_sync(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(TF_GROUP) : "memory");
}

// exp through ex2.approx (max relative error 2^-22): two instructions instead of ~10
__device__ __forceinline__ float tf_exp(float x) {
  float y;  // (.ftz: without the denormal rescue of __expf -- 2 instructions instead of 6)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float tf_rcp(float x) { return __fdividef(1.f, x); }

struct ActNone {
  __device__ __forceinline__ float operator()(float x) const { return x; }
};
struct ActRelu {
  __device__ __forceinline__ float operator()(float x) const { return fmaxf(x, 0.f); }
};
struct ActLeaky {
  __device__ __forceinline__ float operator()(float x) const { return fmaxf(x * 0.01f, x); }
};
struct ActSigmoid {
  __device__ __forceinline__ float operator()(float x) const { return tf_rcp(1.f + tf_exp(-x)); }
};
struct ActTanh {
  __device__ __forceinline__ float operator()(float x) const { return tanhf(x); }
};

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

// d += a * b on both halves: one FFMA2
__device__ __forceinline__ void fma2s(float2& d, float a, float2 b) {  // scalar x pair
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  float2 a2 = make_float2(a, a);
  const unsigned long long aa = *reinterpret_cast<unsigned long long*>(&a2);
  const unsigned long long bb = *reinterpret_cast<unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void fma2p(float2& d, float2 a, float2 b) {  // pair x pair
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}

// floor(n / d) for 0 <= n < 4096 and 1 <= d <= 64 as one multiply and one shift: the quotients
// the tile walks need (thread -> row / chunk, item -> row / column) have run-time divisors, and
// the ~20-instruction integer division sequences were 8 % of the forward kernel's instructions.
// (n * ceil(2^20 / d)) >> 20 is exact in that range (checked exhaustively: tests/test_abi_cpu.py).
__constant__ uint32_t tf_inv20[65] = {
    0,      1048576, 524288, 349526, 262144, 209716, 174763, 149797, 131072, 116509, 104858,
    95326,  87382,   80660,  74899,  69906,  65536,  61681,  58255,  55189,  52429,  49933,
    47663,  45591,   43691,  41944,  40330,  38837,  37450,  36158,  34953,  33826,  32768,
    31776,  30841,   29960,  29128,  28340,  27595,  26887,  26215,  25576,  24967,  24386,
    23832,  23302,   22796,  22311,  21846,  21400,  20972,  20561,  20165,  19785,  19419,
    19066,  18725,   18397,  18079,  17773,  17477,  17190,  16913,  16645,  16384};
__device__ __forceinline__ int tf_div(int n, int d) {
  return (d <= 64 && n < 4096) ? static_cast<int>((static_cast<uint32_t>(n) * tf_inv20[d]) >> 20)
                               : n / d;
}

// (row, chunk) walk of a [rows][F4 chunks] tile by the 512 threads of a group without a
// division per element: when F4 divides the group size a thread keeps its chunk
struct RowWalk {
  int v, c, dv, dc;
  bool fast;
  __device__ __forceinline__ RowWalk(int tid, int F4) {
    v = tf_div(tid, F4);
    c = tid - v * F4;
    dv = tf_div(TF_GROUP, F4);
    dc = TF_GROUP - dv * F4;
    fast = dc == 0;
  }
  __device__ __forceinline__ void next(int F4) {
    v += dv;
    if (!fast) {
      c += dc;
      if (c >= F4) {
        c -= F4;
        v += 1;
      }
    }
  }
};

// ---- tile movers ---------------------------------------------------------------------

// pulls [p, p + bytes) into L2 ahead of the phase that loads it (the backward's tile loads sit
// between two group barriers: their latency is exposed)
__device__ __forceinline__ void tf_prefetch_l2(int tid, const void* p, int bytes) {
  const char* c = static_cast<const char*>(p);
  for (int o = tid * 128; o < bytes; o += TF_GROUP * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c + o));
  if (tid == 0 && bytes > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + bytes - 1));
}

// global [rows][F] -> shared [rows][pitch]; columns F .. 4*ceil(F/4) are zeroed
__device__ __forceinline__ void tf_load_rows(int tid, float* dst, int pitch,
                                             const float* __restrict__ src, int rows, int F) {
  const int F4 = (F + 3) >> 2;
  if ((F & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // the tile is contiguous in global memory: chunk i of the tile is chunk (i % F4) of row i / F4
    // (two chunks requested before the first is stored)
    RowWalk w(tid, F4);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int i = tid; w.v < rows; i += 2 * TF_GROUP) {
      float* d0 = dst + w.v * pitch + 4 * w.c;
      w.next(F4);
      const bool two = w.v < rows;
      float* d1 = dst + w.v * pitch + 4 * w.c;
      const float4 x0 = __ldg(s4 + i);
      const float4 x1 = two ? __ldg(s4 + i + TF_GROUP) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(d0) = x0;
      if (two) {
        *reinterpret_cast<float4*>(d1) = x1;
        w.next(F4);
      }
    }
  } else {
    const int Fp = F4 * 4;
    for (int i = tid; i < rows * Fp; i += TF_GROUP) {
      const int v = tf_div(i, Fp), f = i - v * Fp;
      dst[v * pitch + f] = f < F ? __ldg(src + static_cast<size_t>(v) * F + f) : 0.f;
    }
  }
}

// shared [rows][pitch] -> global [rows][F]
__device__ __forceinline__ void tf_store_rows(int tid, float* __restrict__ dst, const float* src,
                                              int pitch, int rows, int F) {
  const int F4 = (F + 3) >> 2;
  if ((F & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    RowWalk w(tid, F4);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = tid; w.v < rows; i += TF_GROUP, w.next(F4))
      d4[i] = *reinterpret_cast<const float4*>(src + w.v * pitch + 4 * w.c);
  } else {
    for (int i = tid; i < rows * F; i += TF_GROUP) {
      const int v = i / F, f = i - v * F;
      dst[static_cast<size_t>(v) * F + f] = src[v * pitch + f];
    }
  }
}

// shared [rows][pitch] -> global [rows][F], times act'(H) of the layer that produced this layer's
// input (H = its saved output): the hand-over to a Kipf layer below needs no launch of its own
__device__ __forceinline__ void tf_store_rows_actgrad(int tid, float* __restrict__ dst,
                                                      const float* src, int pitch, int rows, int F,
                                                      const float* __restrict__ H, int act) {
  const int F4 = (F + 3) >> 2;
  if ((F & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(H)) & 15) == 0) {
    RowWalk w(tid, F4);
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4* h4 = reinterpret_cast<const float4*>(H);
    for (int i = tid; w.v < rows; i += TF_GROUP, w.next(F4)) {
      const float4 g = *reinterpret_cast<const float4*>(src + w.v * pitch + 4 * w.c);
      const float4 h = __ldg(h4 + i);
      d4[i] = make_float4(tf_act_grad(act, h.x, g.x), tf_act_grad(act, h.y, g.y),
                          tf_act_grad(act, h.z, g.z), tf_act_grad(act, h.w, g.w));
    }
  } else {
    for (int i = tid; i < rows * F; i += TF_GROUP) {
      const int v = i / F, f = i - v * F;
      const size_t o = static_cast<size_t>(v) * F + f;
      dst[o] = tf_act_grad(act, __ldg(H + o), src[v * pitch + f]);
    }
  }
}

// tile-local CSR (or CSC) structure: ptr_s[0..rows] relative to the tile's first entry and
// the neighbour bytes
__device__ __forceinline__ void tf_load_struct(int tid, int* ptr_s, uint8_t* idx_s,
                                               const int32_t* __restrict__ ptr_g,
                                               const uint8_t* __restrict__ idx_g, int r0, int rows,
                                               int e0, int ents) {
  for (int i = tid; i <= rows; i += TF_GROUP) ptr_s[i] = __ldg(ptr_g + r0 + i) - e0;
  if ((e0 & 3) == 0) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(idx_g + e0);
    uint32_t* dst = reinterpret_cast<uint32_t*>(idx_s);
    for (int i = tid; i < (ents + 3) >> 2; i += TF_GROUP) dst[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < ents; i += TF_GROUP) idx_s[i] = __ldg(idx_g + e0 + i);
  }
}

// The same in two halves, so that the loads of several structures (and of the tile's rows) are
// in flight together: tf_struct_fetch requests, tf_struct_put stores and returns the address of
// the tile's first neighbour byte.  The bytes are copied as aligned words from the word that
// holds entry e0, i.e. idx_raw keeps e0 % 4 leading bytes of the previous tile.
struct StructRegs {
  int ptr;
  uint32_t w[2];  // TILE_ENTRIES + 6 bytes <= 2 * TF_GROUP words
};
static_assert((TILE_ENTRIES + 3 + 3) / 4 <= 2 * TF_GROUP, "StructRegs: two words per thread");
static_assert(TILE_ROWS < TF_GROUP, "StructRegs: one row pointer per thread");
__device__ __forceinline__ StructRegs tf_struct_fetch(int tid, const int32_t* __restrict__ ptr_g,
                                                      const uint8_t* __restrict__ idx_g, int r0,
                                                      int rows, int e0, int ents) {
  StructRegs r;
  r.ptr = tid <= rows ? __ldg(ptr_g + r0 + tid) - e0 : 0;
  const int sh = e0 & 3, words = (ents + sh + 3) >> 2;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(idx_g + (e0 - sh));
#pragma unroll
  for (int q = 0; q < 2; ++q)
    r.w[q] = tid + q * TF_GROUP < words ? __ldg(src + tid + q * TF_GROUP) : 0u;
  return r;
}
__device__ __forceinline__ uint8_t* tf_struct_put(int tid, const StructRegs& r, int* ptr_s,
                                                  uint8_t* idx_raw, int rows, int e0, int ents) {
  if (tid <= rows) ptr_s[tid] = r.ptr;
  const int sh = e0 & 3, words = (ents + sh + 3) >> 2;
  uint32_t* dst = reinterpret_cast<uint32_t*>(idx_raw);
#pragma unroll
  for (int q = 0; q < 2; ++q)
    if (tid + q * TF_GROUP < words) dst[tid + q * TF_GROUP] = r.w[q];
  return idx_raw + sh;
}

// dst[v][4c .. 4c+3] = ( sum_{e in row v, ascending} c_e * src[idx[e]][4c .. 4c+3] ) [/ (bkt[v]+1)]
// Lanes walk the chunks of a row first: the lanes of a quarter warp read one contiguous
// piece of a source row (conflict-free); index reads are broadcasts.
template <bool COEF>
__device__ __forceinline__ void tf_gather(int tid, float* dst, int dpitch, const float* src,
                                          int spitch, int F4, const int* ptr_s,
                                          const uint8_t* idx_s, const float* coef_s, int rows,
                                          const float* rd_s, const uint8_t* inv_s = nullptr) {
  RowWalk w(tid, F4);
  for (; w.v < rows; w.next(F4)) {
    const int v = w.v, c = w.c;
    const int e1 = ptr_s[v + 1];
    const float* s = src + 4 * c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int e = ptr_s[v]; e < e1; ++e) {
      const float4 x = *reinterpret_cast<const float4*>(s + idx_s[e] * spitch);
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
    if (rd_s != nullptr) {
      // A(:,v) / real(d), _sub_duvenaud.f90:208 (as a product with the rounded reciprocal, which
      // the tile prologue leaves in rd_s: one division per vertex and tile)
      const float rd = rd_s[v];
      acc.x *= rd;
      acc.y *= rd;
      acc.z *= rd;
      acc.w *= rd;
    }
    // inv_s: row v of the result is stored at position inv_s[v] (rows grouped by degree bucket)
    *reinterpret_cast<float4*>(dst + (inv_s != nullptr ? inv_s[v] : v) * dpitch + 4 * c) = acc;
  }
}

// C[v][n] = epi(v, n, sum_{k ascending} A[v][k] * W[grp[v]][k][n])      v < rows, n < N
// A: shared [rows][pa]; W: shared [k][pw] blocks (pw >= N rounded up to TF_NB, pad columns
// zero), group stride gs; C: shared [rows][pc] (columns N .. 4*ceil(N/4) are zeroed).
// One lane = one vertex row and TF_NB outputs; the row's weights arrive as broadcast LDS.128.
template <class Epi>
__device__ __forceinline__ void tf_gemm(int tid, float* C, int pc, const float* A, int pa, int K,
                                        const float* W, int pw, int gs, const uint8_t* grp_s,
                                        int rows, int N, Epi epi,
                                        const uint8_t* list_s = nullptr) {
  // list_s: A holds the rows grouped by weight block (row p of A is vertex list_s[p]): the lanes
  // of a warp then read ONE block -- every weight load is a broadcast, one shared-memory
  // wavefront instead of one per block present in the warp
  const int warp = tid >> 5, lane = tid & 31;
  const int nvg = (rows + 31) >> 5, nnb = (N + TF_NB - 1) / TF_NB;
  const int N4 = ((N + 3) >> 2) << 2;
  const int K4 = K >> 2;
  for (int it = warp; it < nvg * nnb; it += TF_GWARPS) {
    const int nb = tf_div(it, nvg), vg = it - nb * nvg;
    const int pos = vg * 32 + lane;
    const bool live = pos < rows;
    const int vv = live ? pos : rows - 1;
    const int v = list_s != nullptr ? list_s[vv] : vv;  // the row of C, the vertex of grp_s
    const float4* a = reinterpret_cast<const float4*>(A + vv * pa);
    const float* w = W + (grp_s != nullptr ? grp_s[v] * gs : 0) + nb * TF_NB;
    float2 acc[TF_NB / 2];
#pragma unroll
    for (int j = 0; j < TF_NB / 2; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int k4 = 0; k4 < K4; ++k4) {
      const float4 a4 = a[k4];
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4* wr = reinterpret_cast<const float4*>(w);
        w += pw;
#pragma unroll
        for (int j = 0; j < TF_NB / 4; ++j) {
          const float4 w4 = wr[j];
          fma2s(acc[2 * j], av[i], make_float2(w4.x, w4.y));
          fma2s(acc[2 * j + 1], av[i], make_float2(w4.z, w4.w));
        }
      }
    }
    for (int k = K4 * 4; k < K; ++k) {
      const float ak = A[vv * pa + k];
      const float4* wr = reinterpret_cast<const float4*>(w);
      w += pw;
#pragma unroll
      for (int j = 0; j < TF_NB / 4; ++j) {
        const float4 w4 = wr[j];
        fma2s(acc[2 * j], ak, make_float2(w4.x, w4.y));
        fma2s(acc[2 * j + 1], ak, make_float2(w4.z, w4.w));
      }
    }
    if (live) {
      float* c = C + v * pc + nb * TF_NB;
      const int n0 = nb * TF_NB;
      if (n0 + TF_NB <= N) {
#pragma unroll
        for (int j = 0; j < TF_NB / 4; ++j)
          *reinterpret_cast<float4*>(c + 4 * j) =
              make_float4(epi(v, n0 + 4 * j, acc[2 * j].x), epi(v, n0 + 4 * j + 1, acc[2 * j].y),
                          epi(v, n0 + 4 * j + 2, acc[2 * j + 1].x),
                          epi(v, n0 + 4 * j + 3, acc[2 * j + 1].y));
      } else {
#pragma unroll
        for (int j = 0; j < TF_NB; ++j) {
          const int n = n0 + j;
          const float s = (j & 1) ? acc[j >> 1].y : acc[j >> 1].x;
          if (n < N)
            c[j] = epi(v, n, s);
          else if (n < N4)
            c[j] = 0.f;
        }
      }
    }
  }
}

// C[v][n] = epi(v, n, sum_k A[v][k] * W[grp[v]][n][k])      v < rows, n < N   (W used transposed)
// Same weight blocks as tf_gemm ([row][pw], here the OUTPUT index selects the row and the
// contraction runs along it): pair x pair FFMA2 over k, the two halves added at the end.
// The block must have at least ceil(N / TF_NB) * TF_NB rows.
template <class Epi>
__device__ __forceinline__ void tf_gemm_nt(int tid, float* C, int pc, const float* A, int pa,
                                           int K, const float* W, int pw, int gs,
                                           const uint8_t* grp_s, int rows, int N, Epi epi,
                                           const uint8_t* list_s = nullptr) {
  // list_s: lane p works on vertex list_s[p] (vertices grouped by weight block): the lanes of a
  // warp read ONE block, every weight load is a broadcast (A and C stay in vertex order)
  const int warp = tid >> 5, lane = tid & 31;
  const int nvg = (rows + 31) >> 5, nnb = (N + TF_NB - 1) / TF_NB;
  const int N4 = ((N + 3) >> 2) << 2;
  const int K4 = K >> 2;
  for (int it = warp; it < nvg * nnb; it += TF_GWARPS) {
    const int nb = tf_div(it, nvg), vg = it - nb * nvg;
    const int pos = vg * 32 + lane;
    const bool live = pos < rows;
    const int vv = list_s != nullptr ? list_s[live ? pos : rows - 1] : (live ? pos : rows - 1);
    const int v = vv;
    const float4* a = reinterpret_cast<const float4*>(A + vv * pa);
    const float* w = W + (grp_s != nullptr ? grp_s[vv] * gs : 0) + nb * TF_NB * pw;
    float2 acc[TF_NB];
#pragma unroll
    for (int j = 0; j < TF_NB; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int k4 = 0; k4 < K4; ++k4) {
      const float4 a4 = a[k4];
      const float2 alo = make_float2(a4.x, a4.y), ahi = make_float2(a4.z, a4.w);
#pragma unroll
      for (int j = 0; j < TF_NB; ++j) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + j * pw + 4 * k4);
        fma2p(acc[j], alo, make_float2(w4.x, w4.y));
        fma2p(acc[j], ahi, make_float2(w4.z, w4.w));
      }
    }
    for (int k = K4 * 4; k < K; ++k) {
      const float ak = A[vv * pa + k];
#pragma unroll
      for (int j = 0; j < TF_NB; ++j) acc[j].x = fmaf(ak, w[j * pw + k], acc[j].x);
    }
    if (live) {
      float* c = C + v * pc + nb * TF_NB;
      const int n0 = nb * TF_NB;
      if (n0 + TF_NB <= N) {
#pragma unroll
        for (int j = 0; j < TF_NB / 4; ++j)
          *reinterpret_cast<float4*>(c + 4 * j) = make_float4(
              epi(v, n0 + 4 * j, acc[4 * j].x + acc[4 * j].y),
              epi(v, n0 + 4 * j + 1, acc[4 * j + 1].x + acc[4 * j + 1].y),
              epi(v, n0 + 4 * j + 2, acc[4 * j + 2].x + acc[4 * j + 2].y),
              epi(v, n0 + 4 * j + 3, acc[4 * j + 3].x + acc[4 * j + 3].y));
      } else {
#pragma unroll
        for (int j = 0; j < TF_NB; ++j) {
          const int n = n0 + j;
          if (n < N)
            c[j] = epi(v, n, acc[j].x + acc[j].y);
          else if (n < N4)
            c[j] = 0.f;
        }
      }
    }
  }
}

// Ungrouped products with TWO vertex rows per lane (v and v + 32 of a 64-row group) and four
// outputs: every weight LDS.128 feeds four FFMA2 instead of two.
//   C[v][n] = epi(v, n, sum_k A[v][k] * W[k][n])
template <class Epi>
__device__ __forceinline__ void tf_gemm2(int tid, float* C, int pc, const float* A, int pa, int K,
                                         const float* W, int pw, int rows, int N, Epi epi) {
  const int warp = tid >> 5, lane = tid & 31;
  const int nrg = (rows + 63) >> 6, nnb = (N + 3) >> 2;
  const int K4 = K >> 2;
  for (int it = warp; it < nrg * nnb; it += TF_GWARPS) {
    const int nb = tf_div(it, nrg), rg = it - nb * nrg;
    const int v0 = rg * 64 + lane, v1 = v0 + 32;
    const bool live0 = v0 < rows, live1 = v1 < rows;
    const float* A0 = A + (live0 ? v0 : rows - 1) * pa;
    const float* A1 = A + (live1 ? v1 : rows - 1) * pa;
    const float* w = W + nb * 4;
    float2 acc[2][2];
    acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int k4 = 0; k4 < K4; ++k4) {
      const float4 x0 = *reinterpret_cast<const float4*>(A0 + 4 * k4);
      const float4 x1 = *reinterpret_cast<const float4*>(A1 + 4 * k4);
      const float a0[4] = {x0.x, x0.y, x0.z, x0.w};
      const float a1[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w4 = *reinterpret_cast<const float4*>(w);
        w += pw;
        const float2 wlo = make_float2(w4.x, w4.y), whi = make_float2(w4.z, w4.w);
        fma2s(acc[0][0], a0[i], wlo);
        fma2s(acc[0][1], a0[i], whi);
        fma2s(acc[1][0], a1[i], wlo);
        fma2s(acc[1][1], a1[i], whi);
      }
    }
    for (int k = K4 * 4; k < K; ++k) {
      const float4 w4 = *reinterpret_cast<const float4*>(w);
      w += pw;
      const float2 wlo = make_float2(w4.x, w4.y), whi = make_float2(w4.z, w4.w);
      fma2s(acc[0][0], A0[k], wlo);
      fma2s(acc[0][1], A0[k], whi);
      fma2s(acc[1][0], A1[k], wlo);
      fma2s(acc[1][1], A1[k], whi);
    }
    const int n0 = nb * 4;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int v = rr ? v1 : v0;
      if (rr ? live1 : live0) {
        const float r4[4] = {acc[rr][0].x, acc[rr][0].y, acc[rr][1].x, acc[rr][1].y};
        float o4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o4[j] = n0 + j < N ? epi(v, n0 + j, r4[j]) : 0.f;
        *reinterpret_cast<float4*>(C + v * pc + n0) = make_float4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
}

//   C[v][n] = epi(v, n, sum_k A[v][k] * W[n][k])      (weights used transposed)
template <class Epi>
__device__ __forceinline__ void tf_gemm_nt2(int tid, float* C, int pc, const float* A, int pa,
                                            int K, const float* W, int pw, int rows, int N,
                                            Epi epi, int busy_warps = 0) {
  // busy_warps: the first warps of the group are still inside an outer product that nothing
  // separates from this call: the others take the work
  const int warp = tid >> 5, lane = tid & 31;
  const int nrg = (rows + 63) >> 6, nnb = (N + 3) >> 2;
  const int K4 = K >> 2;
  if (busy_warps > TF_GWARPS / 2) busy_warps = 0;
  for (int it = warp >= busy_warps ? warp - busy_warps : nrg * nnb; it < nrg * nnb;
       it += TF_GWARPS - busy_warps) {
    const int nb = tf_div(it, nrg), rg = it - nb * nrg;
    const int v0 = rg * 64 + lane, v1 = v0 + 32;
    const bool live0 = v0 < rows, live1 = v1 < rows;
    const float* A0 = A + (live0 ? v0 : rows - 1) * pa;
    const float* A1 = A + (live1 ? v1 : rows - 1) * pa;
    const float* w = W + nb * 4 * pw;
    float2 acc[2][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int k4 = 0; k4 < K4; ++k4) {
      const float4 x0 = *reinterpret_cast<const float4*>(A0 + 4 * k4);
      const float4 x1 = *reinterpret_cast<const float4*>(A1 + 4 * k4);
      const float2 x0lo = make_float2(x0.x, x0.y), x0hi = make_float2(x0.z, x0.w);
      const float2 x1lo = make_float2(x1.x, x1.y), x1hi = make_float2(x1.z, x1.w);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + j * pw + 4 * k4);
        const float2 wlo = make_float2(w4.x, w4.y), whi = make_float2(w4.z, w4.w);
        fma2p(acc[0][j], x0lo, wlo);
        fma2p(acc[0][j], x0hi, whi);
        fma2p(acc[1][j], x1lo, wlo);
        fma2p(acc[1][j], x1hi, whi);
      }
    }
    for (int k = K4 * 4; k < K; ++k) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float wk = w[j * pw + k];
        acc[0][j].x = fmaf(A0[k], wk, acc[0][j].x);
        acc[1][j].x = fmaf(A1[k], wk, acc[1][j].x);
      }
    }
    const int n0 = nb * 4;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int v = rr ? v1 : v0;
      if (rr ? live1 : live0) {
        float o4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          o4[j] = n0 + j < N ? epi(v, n0 + j, acc[rr][j].x + acc[rr][j].y) : 0.f;
        *reinterpret_cast<float4*>(C + v * pc + n0) = make_float4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
}

// part[(g*K + k)*N + n] += sum_{p in segment g, ascending} A[list[p]][k] * G[list[p]][n]
// (the layout of a column-major [N, K] parameter block per group: n + N*k + N*K*g).
// One warp owns 8 k x 32 n of one group: lane = (k pair, n quad), i.e. a 2 x 4 register tile
// fed by one LDS.64 of A (4 distinct addresses per warp) and one LDS.128 of G (one contiguous
// 128-byte row piece per warp) per vertex -- 4 FFMA2 per 2 loads.  The group-private partial
// is read and written after the loop.  list_s == nullptr: identity (one segment
// with all rows).  G must be zero in columns N .. 4*ceil(N/4).
// (Measured and dropped: cutting the rows of an ungrouped product into slices for otherwise
// idle warps -- the extra barrier and the combine through shared memory cost more.)
__device__ __forceinline__ void tf_outer(int tid, float* __restrict__ part, const float* A, int pa,
                                         int K, const float* G, int pg, int N,
                                         const uint8_t* list_s, const int* seg_s, int ngroups,
                                         int rows) {
  const int warp = tid >> 5, lane = tid & 31;
  const int kp = lane >> 3, nq = lane & 7;
  const int K8 = (K + 7) >> 3;
  const int NP = (N + 31) >> 5;  // passes of 32 columns
  const int N4 = ((N + 3) >> 2) << 2;
  for (int it = warp; it < ngroups * K8 * NP; it += TF_GWARPS) {
    const int np = it % NP, r = it / NP;
    const int g = r / K8, kb = r - g * K8;
    const int p0 = seg_s != nullptr ? seg_s[g] : 0;
    const int p1 = seg_s != nullptr ? seg_s[g + 1] : rows;
    if (p0 >= p1) continue;
    const int k0 = 8 * kb + 2 * kp, n0 = 32 * np + 4 * nq;
    const bool active = k0 < K && n0 < N4;
    float2 acc[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
    const float* Ak = A + (active ? k0 : 0);
    const float* Gn = G + (active ? n0 : 0);
    // four vertices per round, all eight loads requested before the first FFMA2 (left to the
    // compiler the loop ran one vertex at a time: load, load, wait, 4 x FFMA2)
    int p = p0;
    for (; p + 4 <= p1; p += 4) {
      float2 a2[4];
      float4 g4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int v = list_s != nullptr ? list_s[p + q] : p + q;
        a2[q] = *reinterpret_cast<const float2*>(Ak + v * pa);
        g4[q] = *reinterpret_cast<const float4*>(Gn + v * pg);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 glo = make_float2(g4[q].x, g4[q].y), ghi = make_float2(g4[q].z, g4[q].w);
        fma2s(acc[0][0], a2[q].x, glo);
        fma2s(acc[0][1], a2[q].x, ghi);
        fma2s(acc[1][0], a2[q].y, glo);
        fma2s(acc[1][1], a2[q].y, ghi);
      }
    }
    for (; p < p1; ++p) {
      const int v = list_s != nullptr ? list_s[p] : p;
      const float2 a2 = *reinterpret_cast<const float2*>(Ak + v * pa);
      const float4 g4 = *reinterpret_cast<const float4*>(Gn + v * pg);
      const float2 glo = make_float2(g4.x, g4.y), ghi = make_float2(g4.z, g4.w);
      fma2s(acc[0][0], a2.x, glo);
      fma2s(acc[0][1], a2.x, ghi);
      fma2s(acc[1][0], a2.y, glo);
      fma2s(acc[1][1], a2.y, ghi);
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float* row = part + (static_cast<size_t>(g) * K + k0 + i) * N + n0;
        const float r4[4] = {acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y};
        if (k0 + i < K && n0 + 4 <= N && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
          float4 o = *reinterpret_cast<const float4*>(row);
          o.x += r4[0];
          o.y += r4[1];
          o.z += r4[2];
          o.w += r4[3];
          *reinterpret_cast<float4*>(row) = o;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (k0 + i < K && n0 + j < N) row[j] += r4[j];
        }
      }
    }
  }
}

// per-row softmax, in place (athena_diffstruc_extd_sub.f90:309-313: max-subtracted).  Four
// lanes per row; rows of up to 32 columns are held in registers between the passes.
__device__ __forceinline__ void tf_softmax_rows(int tid, float* Y, int py, int rows, int N) {
  const int v = tid / TF_LPR, h = tid % TF_LPR;
  const bool live = v < rows;
  float* y = Y + (live ? v : 0) * py;
  if (N <= 8 * TF_LPR) {
    float x[8];
    float mx = -INFINITY;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int n = h + TF_LPR * q;
      x[q] = (live && n < N) ? y[n] : -INFINITY;
      mx = fmaxf(mx, x[q]);
    }
#pragma unroll
    for (int o = 1; o < TF_LPR; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      x[q] = (live && h + TF_LPR * q < N) ? tf_exp(x[q] - mx) : 0.f;
      s += x[q];
    }
#pragma unroll
    for (int o = 1; o < TF_LPR; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float rs = 1.f / s;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int n = h + TF_LPR * q;
      if (live && n < N) y[n] = x[q] * rs;
    }
    return;
  }
  float mx = -INFINITY;
  if (live)
    for (int n = h; n < N; n += TF_LPR) mx = fmaxf(mx, y[n]);
#pragma unroll
  for (int o = 1; o < TF_LPR; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  if (live)
    for (int n = h; n < N; n += TF_LPR) {
      const float e = tf_exp(y[n] - mx);
      y[n] = e;
      s += e;
    }
#pragma unroll
  for (int o = 1; o < TF_LPR; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float rs = 1.f / s;
  if (live)
    for (int n = h; n < N; n += TF_LPR) y[n] = y[n] * rs;
}

// in place: Y[v][:] (holding the activation output S) <- d loss / d pre-activation for the
// upstream gradient row g(v); softmax: S*g - S*sum(S*g) (athena_diffstruc_extd_sub.f90:369-373)
template <class GRow>
__device__ __forceinline__ void tf_act_bwd_rows(int tid, int act, float* Y, int py, int rows,
                                                int N, GRow grow) {
  const int v = tid / TF_LPR, h = tid % TF_LPR;
  const bool live = v < rows;
  float* y = Y + (live ? v : 0) * py;
  const float* g = grow(live ? v : 0);
  if (act == ATHENA_ACT_SOFTMAX && N <= 8 * TF_LPR) {
    // rows of up to 32 columns stay in registers between the two passes (S and the upstream
    // gradient row are read once)
    float sv[8], gv[8];
    float dot = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int n = h + TF_LPR * q;
      const bool in = live && n < N;
      sv[q] = in ? y[n] : 0.f;
      gv[q] = in ? g[n] : 0.f;  // (generic: the Kipf reverse kernel keeps this row in shared memory)
      dot += sv[q] * gv[q];
    }
#pragma unroll
    for (int o = 1; o < TF_LPR; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int n = h + TF_LPR * q;
      if (live && n < N) y[n] = sv[q] * gv[q] - sv[q] * dot;
    }
  } else if (act == ATHENA_ACT_SOFTMAX) {
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

// ======================================================================================
// Duvenaud layer
// ======================================================================================

struct DuvLayout {  // offsets in floats from the (16-byte aligned) start of dynamic smem
  int w[TF_MAX_T], r[TF_MAX_T];              // weight blocks (shared by the groups)
  int pw[TF_MAX_T], gs[TF_MAX_T], pr[TF_MAX_T];
  int P, pE;                                 // pitch of the tile buffers / of the edge sums
  int group0, group_stride;                  // start of group 0's region, floats per group
  int buf[3], ae, outs, ints, bytes;         // offsets inside a group's region
  int total_bytes[TF_MAXG + 1];              // dynamic smem for 1 / 2 groups
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
  float* Ae;            // [V][F_e] per-vertex edge-feature sums: written by the forward
                        // (nullable), read by the backward
  const float* params;  // W_1..W_T, R_1..R_T (flat, the reference's packing)
  float* Z[TF_MAX_T];   // z_t, [V][F_t]
  float* S[TF_MAX_T];   // S_t = ract(R_t z_t), [V][no], or nullptr (the backward recomputes it)
  int woff[TF_MAX_T], roff[TF_MAX_T];
  int nvf[TF_MAX_T + 1];
  int T, nef, D, min_deg, max_deg, no, act, ract;
  // forward
  float* out;            // [B][no]
  const float* target;   // fused MSE (nullptr: none)
  float* mse_grad;       // [B][no]
  float mse_denom;       // no * global batch
  float* loss_part;      // [grid * groups]
  // backward
  const float* gout;     // [B][no]
  float* gin;            // [V][F_0] or nullptr
  int fold_act;          // gin leaves multiplied by fold_act'(X) (X = the producing layer's output)
  float* part;           // [grid * groups][np] group-private partial parameter gradients
  int np;
  DuvLayout lay;
};

struct LayerDims {
  int T, nef, D, no;
  const int* nvf;
};

// W_t as [d][k][pw] (n fastest, pad zero, rows padded to a multiple of TF_NB), R_t as [f][pr]
static void duv_layout(const LayerDims& L, DuvLayout* o) {
  int off = 0;
  int Fmax = 0, Kmax = 0;
  for (int t = 0; t <= L.T; ++t) Fmax = std::max(Fmax, L.nvf[t]);
  for (int t = 1; t <= L.T; ++t) {
    const int Fi = L.nvf[t - 1], Fo = L.nvf[t], K = Fi + L.nef;
    Kmax = std::max(Kmax, K);
    const int i = t - 1;
    o->pw[i] = tf_up(Fo, TF_NB);
    int gs = tf_up(K, TF_NB) * o->pw[i];
    if (((gs >> 2) & 1) == 0) gs += 4;
    o->gs[i] = gs;
    o->w[i] = off;
    off += gs * L.D;
    o->pr[i] = tf_up(L.no, TF_NB);
    o->r[i] = off;
    off += tf_up(Fo, TF_NB) * o->pr[i];
  }
  o->P = tf_pitch(std::max(std::max(Fmax, Kmax), L.no));
  o->pE = tf_pitch(L.nef);
  o->group0 = tf_up(off, 4);
  int g = 0;
  for (int k = 0; k < 3; ++k) {
    o->buf[k] = g;
    g += TILE_ROWS * o->P;
  }
  o->ae = g;
  g += TILE_ROWS * o->pE;
  o->outs = g;
  g += TF_OUTS;
  o->ints = g;
  g += TF_INTS;
  g = tf_up(g, 4);
  o->bytes = g;
  g += (2 * TF_IDX + 128 + 128 + 16) / 4;
  o->group_stride = tf_up(g, 4);
  for (int n = 1; n <= TF_MAXG; ++n) o->total_bytes[n] = (o->group0 + n * o->group_stride) * 4;
  o->total_bytes[0] = 0;
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

// Ae[v][:] = sum_w E(:, ja(2,w))  (time-step invariant part of duvenaud_propagate).  The edge
// ids and rows of eight entries are requested before the first is added (two dependent global
// latencies per row of up to eight entries; the additions keep the entry order).
__device__ __forceinline__ void tf_edge_sum(int tid, float* ae, int pe,
                                            const float* __restrict__ E, int Fe,
                                            const int32_t* __restrict__ eid, const int* ptr_s,
                                            int e0, int rows) {
  for (int i = tid; i < rows * Fe; i += TF_GROUP) {
    const int v = tf_div(i, Fe), f = i - v * Fe;
    const int b = ptr_s[v], e1 = ptr_s[v + 1];
    float s = 0.f;
    for (int e = b; e < e1; e += 8) {
      int id[8];
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) id[q] = e + q < e1 ? __ldg(eid + e0 + e + q) : -1;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        x[q] = id[q] >= 0 ? __ldg(E + static_cast<size_t>(id[q]) * Fe + f) : 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (id[q] >= 0) s += x[q];
    }
    ae[v * pe + f] = s;
  }
}

// A[v][Fi .. Fi+Fe) = Ae[v][:] / d ; A[v][K .. 4*ceil(K/4)) = 0
__device__ __forceinline__ void tf_append_edges(int tid, float* A, int pa, int Fi, int Fe,
                                                const float* ae, int pe, const float* rd_s,
                                                int rows, const uint8_t* inv_s = nullptr) {
  const int K = Fi + Fe, Kp = ((K + 3) >> 2) << 2;
  const int w = Kp - Fi;
  if (w == 0) return;
  for (int i = tid; i < rows * w; i += TF_GROUP) {
    const int v = tf_div(i, w), j = i - v * w;
    A[(inv_s != nullptr ? inv_s[v] : v) * pa + Fi + j] =
        j < Fe ? ae[v * pe + j] * rd_s[v] : 0.f;
  }
}

// one warp: the vertices of the tile grouped by degree bucket, ascending inside a bucket:
// list_s[p] = vertex at position p, inv_s[v] = position of vertex v (optional), seg_s[d] = first
// position of bucket d (seg_s[D] = rows)
__device__ __forceinline__ void tf_bucket_list(int lane, const uint8_t* bkt_s, int rows, int D,
                                               uint8_t* list_s, uint8_t* inv_s, int* seg_s) {
  int base = 0;
  for (int d = 0; d < D; ++d) {
    if (lane == 0) seg_s[d] = base;
    for (int c = 0; c < TILE_ROWS / 32; ++c) {
      const int v = c * 32 + lane;
      const bool m = v < rows && bkt_s[v] == d;
      const unsigned bal = __ballot_sync(0xffffffffu, m);
      if (m) {
        const int pos = base + __popc(bal & ((1u << lane) - 1u));
        list_s[pos] = static_cast<uint8_t>(v);
        if (inv_s != nullptr) inv_s[v] = static_cast<uint8_t>(pos);
      }
      base += __popc(bal);
    }
  }
  if (lane == 0) seg_s[D] = base;
}

// one asynchronous 4-byte global -> shared copy: the staging loops below issue all of them
// without waiting for one another (as plain loads each loop exposed its own memory latency:
// 3 T loops in a row, a fifth of the forward kernel on a one-tile batch)
__device__ __forceinline__ void tf_cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(dst))),
               "l"(src)
               : "memory");
}
__device__ __forceinline__ void tf_cp_async_wait() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// (ends with tf_cp_async_wait(); the caller's __syncthreads() publishes the weights)
__device__ __forceinline__ void duv_stage_weights(float* sm, const DuvArgs& a) {
  const int D = a.D, nt = blockDim.x;
  for (int t = 1; t <= a.T; ++t) {
    const int i = t - 1;
    const int Fi = a.nvf[t - 1], Fo = a.nvf[t], K = Fi + a.nef;
    const float* Wg = a.params + a.woff[i];
    const float* Rg = a.params + a.roff[i];
    float* ws = sm + a.lay.w[i];
    const int pw = a.lay.pw[i], gs = a.lay.gs[i];
    const int rows_w = tf_up(K, TF_NB);  // rows per bucket block (pad rows are zero)
    if (nt % pw == 0) {
      // a thread keeps its column: one division per row instead of two per element
      const int row0 = tf_div(threadIdx.x, pw), n = threadIdx.x - row0 * pw;
      for (int row = row0; row < D * rows_w; row += nt / pw) {
        const int d = tf_div(row, rows_w), k = row - d * rows_w;
        float* dst = ws + d * gs + k * pw + n;
        if (k < K && n < Fo)
          tf_cp_async4(dst, Wg + (static_cast<size_t>(d) * K + k) * Fo + n);
        else
          *dst = 0.f;
      }
    } else {
      for (int idx = threadIdx.x; idx < D * rows_w * pw; idx += nt) {
        const int row = tf_div(idx, pw), n = idx - row * pw;
        const int d = tf_div(row, rows_w), k = row - d * rows_w;
        float* dst = ws + d * gs + k * pw + n;
        if (k < K && n < Fo)
          tf_cp_async4(dst, Wg + (static_cast<size_t>(d) * K + k) * Fo + n);
        else
          *dst = 0.f;
      }
    }
    // the bank-shift pad at the end of every bucket block
    for (int idx = threadIdx.x; idx < D * (gs - rows_w * pw); idx += nt) {
      const int padw = gs - rows_w * pw;
      const int d = idx / padw, q = idx - d * padw;
      ws[d * gs + rows_w * pw + q] = 0.f;
    }
    float* rs = sm + a.lay.r[i];
    const int pr = a.lay.pr[i];
    for (int idx = threadIdx.x; idx < tf_up(Fo, TF_NB) * pr; idx += nt) {
      const int f = tf_div(idx, pr), n = idx - f * pr;
      if (f < Fo && n < a.no)
        tf_cp_async4(rs + idx, Rg + static_cast<size_t>(f) * a.no + n);
      else
        rs[idx] = 0.f;
    }
  }
  tf_cp_async_wait();
}

// z = act(A . W_d) with the activation hoisted out of the epilogue's inner loop
__device__ __forceinline__ void duv_update(int tid, int act, float* Z, const float* A, int P, int K,
                                           const float* W, int pw, int gs, const uint8_t* bkt_s,
                                           int rows, int Fo, const uint8_t* list_s = nullptr) {
#define TF_UPD(ACT)                                                                          \
  do {                                                                                       \
    if (bkt_s != nullptr)                                                                    \
      tf_gemm(tid, Z, P, A, P, K, W, pw, gs, bkt_s, rows, Fo,                                \
              [](int, int, float s) { return ACT{}(s); }, list_s);                           \
    else                                                                                     \
      tf_gemm2(tid, Z, P, A, P, K, W, pw, rows, Fo, [](int, int, float s) { return ACT{}(s); }); \
  } while (0)
  switch (act) {
    case ATHENA_ACT_RELU: TF_UPD(ActRelu); break;
    case ATHENA_ACT_LEAKY_RELU: TF_UPD(ActLeaky); break;
    case ATHENA_ACT_SIGMOID: TF_UPD(ActSigmoid); break;
    case ATHENA_ACT_TANH: TF_UPD(ActTanh); break;
    default: TF_UPD(ActNone); break;  // none / linear; softmax is applied row-wise afterwards
  }
#undef TF_UPD
}

__global__ void __launch_bounds__(TF_GROUP * TF_MAXG, 1) k_duv_fwd(const DuvArgs a) {
  TFT_DECL();
  pdl_wait();  // launched programmatically dependent: nothing of the previous kernel is read
  pdl_launch_dependents();  // (or overwritten) before it has completed
  TFT();
  extern __shared__ float4 tf_smem4[];
  float* sm = reinterpret_cast<float*>(tf_smem4);
  const DuvLayout& L = a.lay;
  const int grp = threadIdx.x / TF_GROUP, tid = threadIdx.x % TF_GROUP;
  const int ngrp = blockDim.x / TF_GROUP;
  float* gsm = sm + L.group0 + grp * L.group_stride;
  int* ptr_s = reinterpret_cast<int*>(gsm + L.ints);
  float* red = reinterpret_cast<float*>(ptr_s + 132 + 132 + 260 + 128);
  uint8_t* idx_raw = reinterpret_cast<uint8_t*>(gsm + L.bytes);
  uint8_t* inv_s = idx_raw + TF_IDX;  // (the CSC bytes of the backward's layout)
  uint8_t* bkt_s = idx_raw + 2 * TF_IDX;
  uint8_t* list_s = bkt_s + 128;
  int* seg_s = ptr_s + 132 + 132;
  float* rd_s = reinterpret_cast<float*>(ptr_s + 132 + 132 + 260 + 128 + 16);  // 1 / d per vertex
  float* ae = gsm + L.ae;
  float* outs = gsm + L.outs;
  const int P = L.P;
  duv_stage_weights(sm, a);
  __syncthreads();
  TFT();
  float lsum = 0.f;
  // the descriptor of the next tile (a chain of dependent loads) is requested a whole tile ahead
  TileView tvn{};
  if (blockIdx.x * ngrp + grp < a.num_tiles)
    tvn = tf_tile(a.tiles, blockIdx.x * ngrp + grp, a.num_tiles, a.vgraph, a.num_graphs);
  for (int j = blockIdx.x * ngrp + grp; j < a.num_tiles; j += gridDim.x * ngrp) {
    const TileView tv = tvn;
    if (j + gridDim.x * ngrp < a.num_tiles)
      tvn = tf_tile(a.tiles, j + gridDim.x * ngrp, a.num_tiles, a.vgraph, a.num_graphs);
    tf_sync(grp);  // the previous tile is done with every buffer
    // every global load of the tile is requested before the first one is stored
    const StructRegs sr = tf_struct_fetch(tid, a.row_ptr, a.col8, tv.r0, tv.rows, tv.e0, tv.ents);
    int deg = 0;
    if (tid < tv.rows) deg = __ldg(a.row_ptr + tv.r0 + tid + 1) - __ldg(a.row_ptr + tv.r0 + tid);
    float* xin = gsm + L.buf[0];
    float* AY = gsm + L.buf[1];   // A, then the readout
    float* zout = gsm + L.buf[2];
    tf_load_rows(tid, xin, P, a.X + static_cast<size_t>(tv.r0) * a.nvf[0], tv.rows, a.nvf[0]);
    const uint8_t* idx_s = tf_struct_put(tid, sr, ptr_s, idx_raw, tv.rows, tv.e0, tv.ents);
    if (tid < tv.rows) {
      const int bk = max(a.min_deg, min(deg, a.max_deg)) - a.min_deg;
      bkt_s[tid] = static_cast<uint8_t>(bk);
      rd_s[tid] = 1.f / static_cast<float>(bk + 1);
    }
    tf_sync(grp);
    TFT();
    // A is built with its rows grouped by degree bucket (the barriers below order this before
    // the first use)
    if (tid < 32) tf_bucket_list(tid, bkt_s, tv.rows, a.D, list_s, inv_s, seg_s);
    if (a.nef > 0) {
      tf_edge_sum(tid, ae, L.pE, a.E, a.nef, a.eid, ptr_s, tv.e0, tv.rows);
      tf_sync(grp);
      if (a.Ae != nullptr)
        for (int i = tid; i < tv.rows * a.nef; i += TF_GROUP) {
          const int v = tf_div(i, a.nef), f = i - v * a.nef;
          a.Ae[static_cast<size_t>(tv.r0) * a.nef + i] = ae[v * L.pE + f];
        }
    } else {
      tf_sync(grp);  // the bucket list is complete
    }
    TFT();
    const int ng = tv.g_end - tv.g_begin;
    const bool outs_local = ng * a.no <= TF_OUTS;
    for (int t = 1; t <= a.T; ++t) {
      const int i = t - 1;
      const int Fi = a.nvf[t - 1], Fo = a.nvf[t], K = Fi + a.nef;
      // A = [ sum_w in(:,ja(1,w)) ; sum_w E(:,ja(2,w)) ] / d   (propagate; the division belongs
      // to duvenaud_update)
      tf_gather<false>(tid, AY, P, xin, P, (Fi + 3) >> 2, ptr_s, idx_s, nullptr, tv.rows, rd_s,
                       inv_s);
      if (Fi & 3) tf_sync(grp);  // else the two passes write disjoint 16-byte chunks
      tf_append_edges(tid, AY, P, Fi, a.nef, ae, L.pE, rd_s, tv.rows, inv_s);
      tf_sync(grp);
      TFT();
      // z = act( W_d(v) . A(:,v) ), rows of z back in vertex order
      duv_update(tid, a.act, zout, AY, P, K, sm + L.w[i], L.pw[i], L.gs[i], bkt_s, tv.rows, Fo,
                 list_s);
      tf_sync(grp);
      TFT();
      tf_store_rows(tid, a.Z[i] + static_cast<size_t>(tv.r0) * Fo, zout, P, tv.rows, Fo);
      // readout: S = ract( R_t . z )
      duv_update(tid, a.ract, AY, zout, P, Fo, sm + L.r[i], L.pr[i], 0, nullptr, tv.rows, a.no);
      tf_sync(grp);
      TFT();
      if (a.ract == ATHENA_ACT_SOFTMAX) {
        tf_softmax_rows(tid, AY, P, tv.rows, a.no);
        tf_sync(grp);
      }
      TFT();
      // S_t for the reverse sweep (the kernels are bound by shared-memory traffic, not by HBM:
      // 128 bytes per vertex and step here save the backward a product and a softmax)
      if (a.S[i] != nullptr)
        tf_store_rows(tid, a.S[i] + static_cast<size_t>(tv.r0) * a.no, AY, P, tv.rows, a.no);
      // out(:,s) (+)= sum_v S(:,v) (sum(ptr2, dim=2), :848-852).  Four lanes per (graph, output)
      // pair add every fourth vertex and are combined in a fixed order: a quarter of the
      // dependent-add chain of one lane per pair
      const int npairs = ng * a.no;
      for (int base = 0; base < npairs; base += TF_GROUP / 4) {
        const int idx = base + (tid >> 2), q = tid & 3;
        const bool live = idx < npairs;
        int g = 0, o = 0;
        float s = 0.f;
        if (live) {
          const int gl = tf_div(idx, a.no);
          o = idx - gl * a.no;
          g = tv.g_begin + gl;
          const int v0 = __ldg(a.voff + g) - tv.r0, v1 = __ldg(a.voff + g + 1) - tv.r0;
          for (int v = v0 + q; v < v1; v += 4) s += AY[v * P + o];
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (!live || q != 0) continue;
        float* dst = a.out + static_cast<size_t>(g) * a.no + o;
        float tot;
        if (outs_local) {
          tot = t == 1 ? s : outs[idx] + s;
          outs[idx] = tot;
          if (t == a.T) *dst = tot;
        } else {
          tot = t == 1 ? s : *dst + s;
          *dst = tot;
        }
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
      tf_sync(grp);  // the readout buffer becomes A again
      TFT();
    }
  }
  TFT_DUMP("duv_fwd [wait, weights, tile-load, edge-sum; per t: gather+append, update, readout, softmax, segsum]");
  if (a.loss_part != nullptr) {
    // per-group loss partial, fixed combine order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    tf_sync(grp);
    if ((tid & 31) == 0) red[tid >> 5] = lsum;
    tf_sync(grp);
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < TF_GWARPS; ++w) tot += red[w];
      a.loss_part[blockIdx.x * ngrp + grp] = tot;
    }
  }
}

__global__ void __launch_bounds__(TF_GROUP * TF_MAXG, 1) k_duv_bwd(const DuvArgs a) {
  TFT_DECL();
  pdl_wait();
  pdl_launch_dependents();
  TFT();
  extern __shared__ float4 tf_smem4[];
  float* sm = reinterpret_cast<float*>(tf_smem4);
  const DuvLayout& L = a.lay;
  const int grp = threadIdx.x / TF_GROUP, tid = threadIdx.x % TF_GROUP;
  const int ngrp = blockDim.x / TF_GROUP;
  float* gsm = sm + L.group0 + grp * L.group_stride;
  int* ptr_s = reinterpret_cast<int*>(gsm + L.ints);
  int* cptr_s = ptr_s + 132;
  int* seg_s = cptr_s + 132;
  float* rd_s = reinterpret_cast<float*>(ptr_s + 132 + 132 + 260 + 128 + 16);  // 1 / d per vertex
  int* vg_s = seg_s + 260;
  uint8_t* idx_raw = reinterpret_cast<uint8_t*>(gsm + L.bytes);
  uint8_t* cidx_raw = idx_raw + TF_IDX;
  uint8_t* bkt_s = cidx_raw + TF_IDX;
  uint8_t* list_s = bkt_s + 128;
  float* ae = gsm + L.ae;
  const int P = L.P;
  float* part = a.part + static_cast<size_t>(blockIdx.x * ngrp + grp) * a.np;
  for (int i = tid; i < a.np; i += TF_GROUP) part[i] = 0.f;
  duv_stage_weights(sm, a);
  __syncthreads();
  TFT();
  const int T = a.T;
  int4 tin = make_int4(0, 0, 0, 0);  // (first row, rows, first entry, entries) of the next tile
  if (blockIdx.x * ngrp + grp < a.num_tiles) tin = __ldg(a.tiles + blockIdx.x * ngrp + grp);
  for (int j = blockIdx.x * ngrp + grp; j < a.num_tiles; j += gridDim.x * ngrp) {
    TileView tv;
    tv.r0 = tin.x;
    tv.rows = tin.y;
    tv.e0 = tin.z;
    tv.ents = tin.w;
    if (j + gridDim.x * ngrp < a.num_tiles) tin = __ldg(a.tiles + j + gridDim.x * ngrp);
    tf_sync(grp);
    // every global load of the tile is requested before the first one is stored
    const StructRegs sr = tf_struct_fetch(tid, a.row_ptr, a.col8, tv.r0, tv.rows, tv.e0, tv.ents);
    const StructRegs sc = tf_struct_fetch(tid, a.csc_ptr, a.csc8, tv.r0, tv.rows, tv.e0, tv.ents);
    int deg = 0, vgv = 0;
    if (tid < tv.rows) {
      deg = __ldg(a.row_ptr + tv.r0 + tid + 1) - __ldg(a.row_ptr + tv.r0 + tid);
      vgv = __ldg(a.vgraph + tv.r0 + tid);
    }
    // edge-feature sums saved by the forward
    for (int i = tid; i < tv.rows * a.nef; i += TF_GROUP) {
      const int v = tf_div(i, a.nef), f = i - v * a.nef;
      ae[v * L.pE + f] = __ldg(a.Ae + static_cast<size_t>(tv.r0) * a.nef + i);
    }
    float* X = gsm + L.buf[0];  // z_t, then z_{t-1}
    float* Y = gsm + L.buf[1];  // readout / dY, then A, then dA
    float* Z = gsm + L.buf[2];  // carry -> gz -> the new carry
    tf_load_rows(tid, X, P, a.Z[T - 1] + static_cast<size_t>(tv.r0) * a.nvf[T], tv.rows, a.nvf[T]);
    if (a.S[T - 1] != nullptr)
      tf_load_rows(tid, Y, P, a.S[T - 1] + static_cast<size_t>(tv.r0) * a.no, tv.rows, a.no);
    const uint8_t* idx_s = tf_struct_put(tid, sr, ptr_s, idx_raw, tv.rows, tv.e0, tv.ents);
    const uint8_t* cidx_s = tf_struct_put(tid, sc, cptr_s, cidx_raw, tv.rows, tv.e0, tv.ents);
    if (tid < tv.rows) {
      const int bk = max(a.min_deg, min(deg, a.max_deg)) - a.min_deg;
      bkt_s[tid] = static_cast<uint8_t>(bk);
      rd_s[tid] = 1.f / static_cast<float>(bk + 1);
      vg_s[tid] = vgv;
    }
    tf_sync(grp);
    // vertices of the tile grouped by degree bucket (ascending vertex inside a bucket)
    if (tid < 32) tf_bucket_list(tid, bkt_s, tv.rows, a.D, list_s, nullptr, seg_s);
    TFT();
    {
      // the next tile of this group: its z_T rows on their way to L2 while this tile computes
      const int jn = j + gridDim.x * ngrp;
      if (jn < a.num_tiles && tid < 64)
        tf_prefetch_l2(tid * (TF_GROUP / 64), a.Z[T - 1] + static_cast<size_t>(tin.x) * a.nvf[T],
                       tin.y * a.nvf[T] * 4);
    }
    for (int t = T; t >= 1; --t) {
      const int i = t - 1;
      const int Fi = a.nvf[t - 1], Fo = a.nvf[t], K = Fi + a.nef, no = a.no;
      tf_prefetch_l2(tid, (t >= 2 ? a.Z[t - 2] : a.X) + static_cast<size_t>(tv.r0) * Fi,
                     tv.rows * Fi * 4);  // step 4 loads it
      if (t >= 2 && a.S[t - 2] != nullptr)  // step 1 of the next iteration loads it
        tf_prefetch_l2(tid, a.S[t - 2] + static_cast<size_t>(tv.r0) * no, tv.rows * no * 4);
      // 1. S = ract(R_t z_t) into Y (read back if the forward saved it), then dY in place
      //    (upstream row = gout of the vertex's graph)
      if (a.S[i] != nullptr) {
        if (t < T) {
          tf_load_rows(tid, Y, P, a.S[i] + static_cast<size_t>(tv.r0) * no, tv.rows, no);
          tf_sync(grp);
        }
      } else {
        duv_update(tid, a.ract, Y, X, P, Fo, sm + L.r[i], L.pr[i], 0, nullptr, tv.rows, no);
        tf_sync(grp);
        if (a.ract == ATHENA_ACT_SOFTMAX) {
          tf_softmax_rows(tid, Y, P, tv.rows, no);
          tf_sync(grp);
        }
      }
      {
        const float* gout = a.gout;
        const int* vgs = vg_s;
        tf_act_bwd_rows(tid, a.ract, Y, P, tv.rows, no,
                        [gout, vgs, no](int v) { return gout + static_cast<size_t>(vgs[v]) * no; });
      }
      tf_sync(grp);
      TFT();
      // 2. dR_t(o,f) += sum_v dY(o,v) z_t(f,v)
      tf_outer(tid, part + a.roff[i], X, P, Fo, Y, P, no, nullptr, nullptr, 1, tv.rows);
      // 3. gz = ( R_t^T dY + carry ) .* act'(z_t), in place in Z
      {
        const int act = a.act;
        const bool has_carry = t < T;
        const float* zt = X;
        const float* cc = Z;
        tf_gemm_nt2(tid, Z, P, Y, P, no, sm + L.r[i], L.pr[i], tv.rows, Fo,
                    [act, has_carry, zt, cc, P](int v, int n, float s) {
                      const float dz = has_carry ? s + cc[v * P + n] : s;
                      return tf_act_grad(act, zt[v * P + n], dz);
                    },
                    ((Fo + 7) >> 3) * ((no + 31) >> 5) /* warps inside step 2 */);
      }
      tf_sync(grp);
      TFT();
      // 4. z_{t-1} replaces z_t; A = [gather(z_{t-1}) ; Ae] / d is recomputed into Y (dY is dead)
      {
        const float* prev = t >= 2 ? a.Z[t - 2] : a.X;
        tf_load_rows(tid, X, P, prev + static_cast<size_t>(tv.r0) * Fi, tv.rows, Fi);
      }
      tf_sync(grp);
      TFT();
      tf_gather<false>(tid, Y, P, X, P, (Fi + 3) >> 2, ptr_s, idx_s, nullptr, tv.rows, rd_s);
      if (Fi & 3) tf_sync(grp);  // else the two passes write disjoint 16-byte chunks
      tf_append_edges(tid, Y, P, Fi, a.nef, ae, L.pE, rd_s, tv.rows);
      tf_sync(grp);
      TFT();
      // 5. dW_{t,d}(o,k) += sum_{v in bucket d} gz(o,v) A(k,v)      (A already divided by d)
      tf_outer(tid, part + a.woff[i], Y, P, K, Z, P, Fo, list_s, seg_s, a.D, tv.rows);
      const bool need_dx = t > 1 || a.gin != nullptr;
      if (need_dx) {
        // 6. dA(k,v) = ( W_d^T gz(:,v) )(k) / d for k < Fi, over A in Y once step 5 has read it:
        //    X keeps z_{t-1}, which is the next iteration's z_t (no second load of the tile)
        tf_sync(grp);
        TFT();
        const float* rd = rd_s;
        tf_gemm_nt(tid, Y, P, Z, P, Fo, sm + L.w[i], L.pw[i], L.gs[i], bkt_s, tv.rows, Fi,
                   [rd](int v, int, float s) { return s * rd[v]; },
                   list_s);
        tf_sync(grp);
        TFT();
        // 7. d in(:,u) = sum over the CSC column of u of dA(1:Fi, v): the new carry, over gz in Z
        tf_gather<false>(tid, Z, P, Y, P, (Fi + 3) >> 2, cptr_s, cidx_s, nullptr, tv.rows,
                         nullptr);
        tf_sync(grp);
        TFT();
        if (t == 1) {
          float* gdst = a.gin + static_cast<size_t>(tv.r0) * Fi;
          if (a.fold_act == ATHENA_ACT_NONE)
            tf_store_rows(tid, gdst, Z, P, tv.rows, Fi);
          else
            tf_store_rows_actgrad(tid, gdst, Z, P, tv.rows, Fi,
                                  a.X + static_cast<size_t>(tv.r0) * Fi, a.fold_act);
        }
      }
    }
  }
  TFT_DUMP("duv_bwd [wait, zero+weights, tile-load+buckets; per t: readout+softmax+dY, dR+gz, load z, gather+append, dW, dA, scatter]");
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
  float* part;           // backward: [grid * groups][Fi*Fo]
  int Fi, Fo, act;
  int P_, pw;            // pitch of the tile buffers, of the weight rows
  int group0, group_stride, buf[3], ints, bytes;
};

__global__ void __launch_bounds__(TF_GROUP * TF_MAXG, 1) k_kipf_fwd(const KipfArgs a) {
  extern __shared__ float4 tf_smem4[];
  float* sm = reinterpret_cast<float*>(tf_smem4);
  const int grp = threadIdx.x / TF_GROUP, tid = threadIdx.x % TF_GROUP;
  const int ngrp = blockDim.x / TF_GROUP;
  float* Ws = sm;
  float* gsm = sm + a.group0 + grp * a.group_stride;
  float* XH = gsm + a.buf[0];   // input tile, then the step output
  float* Pb = gsm + a.buf[1];
  int* ptr_s = reinterpret_cast<int*>(gsm + a.ints);
  float* coef_s = reinterpret_cast<float*>(ptr_s + 132);
  uint8_t* idx_s = reinterpret_cast<uint8_t*>(gsm + a.bytes);
  const int Fi = a.Fi, Fo = a.Fo, P = a.P_;
  // W_t [Fo, Fi] column-major = [k = i][n = o] rows
  for (int idx = threadIdx.x; idx < tf_up(Fi, TF_NB) * a.pw; idx += blockDim.x) {
    const int k = idx / a.pw, n = idx - k * a.pw;
    Ws[idx] = (k < Fi && n < Fo) ? __ldg(a.W + static_cast<size_t>(k) * Fo + n) : 0.f;
  }
  __syncthreads();
  for (int j = blockIdx.x * ngrp + grp; j < a.num_tiles; j += gridDim.x * ngrp) {
    const int4 ti = __ldg(a.tiles + j);
    const int r0 = ti.x, rows = ti.y, e0 = ti.z, ents = ti.w;
    tf_sync(grp);
    tf_load_struct(tid, ptr_s, idx_s, a.ptr, a.idx8, r0, rows, e0, ents);
    if (a.coef != nullptr)
      for (int e = tid; e < ents; e += TF_GROUP) coef_s[e] = __ldg(a.coef + e0 + e);
    tf_load_rows(tid, XH, P, a.X + static_cast<size_t>(r0) * Fi, rows, Fi);
    tf_sync(grp);
    // P(:,v) = sum_w (deg_v deg_u)^-1/2 X(:,u)
    if (a.coef != nullptr)
      tf_gather<true>(tid, Pb, P, XH, P, (Fi + 3) >> 2, ptr_s, idx_s, coef_s, rows, nullptr);
    else
      tf_gather<false>(tid, Pb, P, XH, P, (Fi + 3) >> 2, ptr_s, idx_s, nullptr, rows, nullptr);
    tf_sync(grp);
    if (a.P != nullptr) tf_store_rows(tid, a.P + static_cast<size_t>(r0) * Fi, Pb, P, rows, Fi);
    duv_update(tid, a.act, XH, Pb, P, Fi, Ws, a.pw, 0, nullptr, rows, Fo);
    tf_sync(grp);
    if (a.act == ATHENA_ACT_SOFTMAX) {
      tf_softmax_rows(tid, XH, P, rows, Fo);
      tf_sync(grp);
    }
    tf_store_rows(tid, a.out + static_cast<size_t>(r0) * Fo, XH, P, rows, Fo);
  }
}

__global__ void __launch_bounds__(TF_GROUP * TF_MAXG, 1) k_kipf_bwd(const KipfArgs a) {
  extern __shared__ float4 tf_smem4[];
  float* sm = reinterpret_cast<float*>(tf_smem4);
  const int grp = threadIdx.x / TF_GROUP, tid = threadIdx.x % TF_GROUP;
  const int ngrp = blockDim.x / TF_GROUP;
  float* Ws = sm;             // [i][pw] (o fastest): the forward layout, used transposed
  float* gsm = sm + a.group0 + grp * a.group_stride;
  float* Gb = gsm + a.buf[0];  // upstream gradient, later dP
  float* Pb = gsm + a.buf[1];  // saved aggregate, later the gathered input gradient
  float* Hb = gsm + a.buf[2];  // saved output -> gY
  int* ptr_s = reinterpret_cast<int*>(gsm + a.ints);
  uint8_t* idx_s = reinterpret_cast<uint8_t*>(gsm + a.bytes);
  const int Fi = a.Fi, Fo = a.Fo, P = a.P_;
  float* part = a.part + static_cast<size_t>(blockIdx.x * ngrp + grp) * Fi * Fo;
  for (int i = tid; i < Fi * Fo; i += TF_GROUP) part[i] = 0.f;
  for (int idx = threadIdx.x; idx < tf_up(Fi, TF_NB) * a.pw; idx += blockDim.x) {
    const int k = idx / a.pw, n = idx - k * a.pw;
    Ws[idx] = (k < Fi && n < Fo) ? __ldg(a.W + static_cast<size_t>(k) * Fo + n) : 0.f;
  }
  __syncthreads();
  const bool need_act = a.H != nullptr;
  for (int j = blockIdx.x * ngrp + grp; j < a.num_tiles; j += gridDim.x * ngrp) {
    const int4 ti = __ldg(a.tiles + j);
    const int r0 = ti.x, rows = ti.y, e0 = ti.z, ents = ti.w;
    tf_sync(grp);
    if (a.out != nullptr) tf_load_struct(tid, ptr_s, idx_s, a.ptr, a.idx8, r0, rows, e0, ents);
    tf_load_rows(tid, Gb, P, a.X + static_cast<size_t>(r0) * Fo, rows, Fo);
    tf_load_rows(tid, Pb, P, a.P + static_cast<size_t>(r0) * Fi, rows, Fi);
    if (need_act) tf_load_rows(tid, Hb, P, a.H + static_cast<size_t>(r0) * Fo, rows, Fo);
    tf_sync(grp);
    if (need_act) {
      // gY = gH .* act'(H) in place in Hb  (softmax: per-vertex Jacobian)
      const float* gb = Gb;
      tf_act_bwd_rows(tid, a.act, Hb, P, rows, Fo, [gb, P](int v) { return gb + v * P; });
      tf_sync(grp);
    }
    const float* gy = need_act ? Hb : Gb;
    // dW_t(o,i) += sum_v gY(o,v) P(i,v)
    tf_outer(tid, part, Pb, P, Fi, gy, P, Fo, nullptr, nullptr, 1, rows);
    if (a.out != nullptr) {
      // dP = W_t^T gY, then the un-normalised scatter dX(:,u) += dP(:,v) as a gather over the
      // CSC (entries ascending)
      float* dP = need_act ? Gb : Hb;
      tf_gemm_nt2(tid, dP, P, gy, P, Fo, Ws, a.pw, rows, Fi, [](int, int, float s) { return s; });
      tf_sync(grp);  // also orders the outer product's reads of Pb before the gather's writes
      tf_gather<false>(tid, Pb, P, dP, P, (Fi + 3) >> 2, ptr_s, idx_s, nullptr, rows, nullptr);
      tf_sync(grp);
      tf_store_rows(tid, a.out + static_cast<size_t>(r0) * Fi, Pb, P, rows, Fi);
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
  int P, pw, group0, group_stride, buf[3], ints, bytes, total_bytes[TF_MAXG + 1];
};
static KipfLayout kipf_layout(int Fi, int Fo, bool backward) {
  KipfLayout k;
  k.P = tf_pitch(std::max(Fi, Fo));
  k.pw = tf_up(Fo, TF_NB);
  k.group0 = tf_up(tf_up(Fi, TF_NB) * k.pw, 4);
  int g = 0;
  const int nbuf = backward ? 3 : 2;
  for (int i = 0; i < 3; ++i) {
    k.buf[i] = g;
    if (i < nbuf) g += TILE_ROWS * k.P;
  }
  k.ints = g;
  g += 132 + (backward ? 0 : TILE_ENTRIES + 16);
  g = tf_up(g, 4);
  k.bytes = g;
  g += (TF_IDX + 16) / 4;
  k.group_stride = tf_up(g, 4);
  k.total_bytes[0] = 0;
  for (int n = 1; n <= TF_MAXG; ++n) k.total_bytes[n] = (k.group0 + n * k.group_stride) * 4;
  return k;
}

// groups per CTA: two when both fit and there is work for them
static int tf_groups(const int* total_bytes, int num_tiles) {
  const size_t lim = ctx().max_smem_optin;
  if ((size_t)total_bytes[1] > lim) return 0;
  static int one = -1;
  if (one < 0) {
    const char* e = getenv("ATHENA_DEBUG_TILE_ONE_GROUP");  // experiments: one group per CTA
    one = (e && atoi(e) != 0) ? 1 : 0;
  }
  if (one) return 1;
  if (num_tiles >= 2 && (size_t)total_bytes[2] <= lim) return 2;
  return 1;
}

static void kipf_fill(KipfArgs* a, const KipfLayout& k) {
  a->P_ = k.P;
  a->pw = k.pw;
  a->group0 = k.group0;
  a->group_stride = k.group_stride;
  for (int i = 0; i < 3; ++i) a->buf[i] = k.buf[i];
  a->ints = k.ints;
  a->bytes = k.bytes;
}

}  // namespace

// ---- host side ---------------------------------------------------------------------

bool tile_kipf_supported(const Batch* b, int Fi, int Fo) {
  if (!tile_fma_enabled() || b->force_list || b->num_tiles == 0 || b->col8 == nullptr) return false;
  if (Fi < 1 || Fo < 1 || Fi > 128 || Fo > 128) return false;
  return tf_groups(kipf_layout(Fi, Fo, false).total_bytes, 1) > 0 &&
         tf_groups(kipf_layout(Fi, Fo, true).total_bytes, 1) > 0;
}

int launch_tile_kipf_fwd(const Batch* b, const float* X, const float* W, float* P, float* out,
                         int Fi, int Fo, int act) {
  const KipfLayout k = kipf_layout(Fi, Fo, false);
  const int ng = tf_groups(k.total_bytes, b->num_tiles);
  ATH_REQUIRE(ng > 0, ATHENA_ERR_ARG, "tile_kipf_fwd: unsupported shape %d -> %d", Fi, Fo);
  static int smem_set = 0;
  if (k.total_bytes[ng] > smem_set) {
    ATH_TRY(tf_set_smem(k_kipf_fwd, k.total_bytes[ng]));
    smem_set = k.total_bytes[ng];
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
  kipf_fill(&a, k);
  const int grid = std::min((int)cdiv(b->num_tiles, ng), ctx().sm_count);
  k_kipf_fwd<<<grid, TF_GROUP * ng, k.total_bytes[ng], ctx().stream>>>(a);
  ATH_LAUNCHED_T("tile_kipf_fwd");
  return ATHENA_OK;
}

// G: gradient w.r.t. the step output (H != nullptr: act'(H) is applied here) or already
// w.r.t. the pre-activation (H == nullptr).  part: [*nparts][Fi*Fo] group partials of dW.
int launch_tile_kipf_bwd(const Batch* b, const float* G, const float* H, const float* P,
                         const float* W, float* gin, int Fi, int Fo, int act, DevBuf& part,
                         int* nparts) {
  const KipfLayout k = kipf_layout(Fi, Fo, true);
  const int ng = tf_groups(k.total_bytes, b->num_tiles);
  ATH_REQUIRE(ng > 0, ATHENA_ERR_ARG, "tile_kipf_bwd: unsupported shape %d -> %d", Fi, Fo);
  static int smem_set = 0;
  if (k.total_bytes[ng] > smem_set) {
    ATH_TRY(tf_set_smem(k_kipf_bwd, k.total_bytes[ng]));
    smem_set = k.total_bytes[ng];
  }
  const int grid = std::min((int)cdiv(b->num_tiles, ng), ctx().sm_count);
  ATH_TRY(part.reserve(sizeof(float) * (size_t)grid * ng * Fi * Fo));
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
  kipf_fill(&a, k);
  k_kipf_bwd<<<grid, TF_GROUP * ng, k.total_bytes[ng], ctx().stream>>>(a);
  ATH_LAUNCHED_T("tile_kipf_bwd");
  *nparts = grid * ng;
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
  LayerDims L{T, nef, D, no, nvf};
  DuvLayout lay;
  duv_layout(L, &lay);
  return tf_groups(lay.total_bytes, 1) > 0;
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
  a->Ae = d.Ae;
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
    a->S[t] = d.S != nullptr ? d.S[t] : nullptr;
    a->woff[t] = (int)d.poff[t];
    a->roff[t] = (int)d.poff[d.T + t];
  }
  LayerDims L{d.T, d.nef, a->D, d.no, d.nvf};
  duv_layout(L, &a->lay);
}

int launch_tile_duv_fwd(const Batch* b, const TileDuvDesc& d, float* out, const float* target,
                        float* mse_grad, float mse_denom, float* loss_part, int* num_parts) {
  DuvArgs a{};
  duv_fill(&a, b, d);
  const int ng = tf_groups(a.lay.total_bytes, b->num_tiles);
  ATH_REQUIRE(ng > 0, ATHENA_ERR_ARG, "tile_duv_fwd: layer does not fit shared memory");
  static int smem_set = 0;
  if (a.lay.total_bytes[ng] > smem_set) {
    ATH_TRY(tf_set_smem(k_duv_fwd, a.lay.total_bytes[ng]));
    smem_set = a.lay.total_bytes[ng];
  }
  a.out = out;
  a.target = target;
  a.mse_grad = mse_grad;
  a.mse_denom = mse_denom;
  a.loss_part = target != nullptr ? loss_part : nullptr;
  const int grid = std::min((int)cdiv(b->num_tiles, ng), ctx().sm_count);
  ATH_CUDA(launch_pdl(k_duv_fwd, dim3(grid), dim3(TF_GROUP * ng), a.lay.total_bytes[ng],
                      ctx().stream, a));
  ATH_LAUNCHED_T(target != nullptr ? "tile_duv_fwd_mse" : "tile_duv_fwd");
  if (num_parts) *num_parts = grid * ng;
  return ATHENA_OK;
}

int launch_tile_duv_bwd(const Batch* b, const TileDuvDesc& d, const float* gout, float* gin,
                        int64_t num_params, DevBuf& part, int* nparts) {
  DuvArgs a{};
  duv_fill(&a, b, d);
  const int ng = tf_groups(a.lay.total_bytes, b->num_tiles);
  ATH_REQUIRE(ng > 0, ATHENA_ERR_ARG, "tile_duv_bwd: layer does not fit shared memory");
  ATH_REQUIRE(d.nef == 0 || d.Ae != nullptr, ATHENA_ERR_STATE,
              "tile_duv_bwd: the forward did not save the edge-feature sums");
  static int smem_set = 0;
  if (a.lay.total_bytes[ng] > smem_set) {
    ATH_TRY(tf_set_smem(k_duv_bwd, a.lay.total_bytes[ng]));
    smem_set = a.lay.total_bytes[ng];
  }
  const int grid = std::min((int)cdiv(b->num_tiles, ng), ctx().sm_count);
  ATH_TRY(part.reserve(sizeof(float) * (size_t)grid * ng * (size_t)num_params));
  a.gout = gout;
  a.gin = gin;
  a.fold_act = gin != nullptr ? d.fold_act : ATHENA_ACT_NONE;
  a.part = part.as<float>();
  a.np = (int)num_params;
  ATH_CUDA(launch_pdl(k_duv_bwd, dim3(grid), dim3(TF_GROUP * ng), a.lay.total_bytes[ng],
                      ctx().stream, a));
  ATH_LAUNCHED_T("tile_duv_bwd");
  *nparts = grid * ng;
  return ATHENA_OK;
}

}  // namespace athena
