// General (any shape) FP32 dense kernels of the message-passing layers: the
// feature transform Y = P.W (diffstruc matmul, call site
// athena_kipf_msgpass_layer.f90:951), the degree-bucketed Duvenaud update
// (athena_diffstruc_extd_sub_duvenaud.f90:204-211) as a grouped GEMM over the
// bucket permutation, and their backward products
//   dP = gY.W^T      (matmul partial; duvenaud :284-324)
//   dW = P^T.gY      (matmul partial; duvenaud :326-368), reduced over all
//                    vertices of the batch in two deterministic passes.
// These are the shape-generic FFMA kernels; the tcgen05 path for wide
// features lives in gemm_tc.cu.
#include <algorithm>

#include "athena_internal.h"

namespace athena {

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float act_apply(int act, float x) {
  switch (act) {
    case ATHENA_ACT_RELU: return fmaxf(x, 0.f);
    case ATHENA_ACT_LEAKY_RELU: return fmaxf(x * 0.01f, x);
    case ATHENA_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case ATHENA_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

// derivative expressed on the saved OUTPUT y
__device__ __forceinline__ float act_grad(int act, float y, float g) {
  switch (act) {
    case ATHENA_ACT_RELU: return y > 0.f ? g : 0.f;
    case ATHENA_ACT_LEAKY_RELU: return y > 0.f ? g : g * 0.01f;
    case ATHENA_ACT_SIGMOID: return g * (y * (1.f - y));
    case ATHENA_ACT_TANH: return g * (1.f - y * y);
    default: return g;
  }
}

struct TileInfo {
  int group, r0, rcnt;
};

// Map a linear tile index to (group, first permuted row, row count) such that
// tiles never straddle a group.  Returns false for surplus blocks.
__device__ __forceinline__ bool locate_tile(int tile, int tile_rows, const int32_t* __restrict__ ptr,
                                            int D, long long M, TileInfo* ti) {
  if (ptr == nullptr) {
    long long r0 = (long long)tile * tile_rows;
    if (r0 >= M) return false;
    ti->group = 0;
    ti->r0 = (int)r0;
    ti->rcnt = (int)min((long long)tile_rows, M - r0);
    return true;
  }
  for (int d = 0; d < D; ++d) {
    int b = __ldg(ptr + d), e = __ldg(ptr + d + 1);
    int nt = (e - b + tile_rows - 1) / tile_rows;
    if (tile < nt) {
      ti->group = d;
      ti->r0 = b + tile * tile_rows;
      ti->rcnt = min(tile_rows, e - ti->r0);
      return true;
    }
    tile -= nt;
  }
  return false;
}

// C[row, n0:n0+BN] = epilogue( sum_k A'[row, k] * Bt[k, n] )
//   TRANSB = false: Bt[k][n] = W[k*N + n]          (NN)
//   TRANSB = true : Bt[k][n] = W[n*K + k]          (NT)
//   scale: NN divides A rows by (group+1) on load; NT divides the result.
template <bool TRANSB>
__global__ void __launch_bounds__(256)
k_gemm_rows(const float* __restrict__ A, int lda, const float* __restrict__ W,
            float* __restrict__ C, int ldc, long long M, int N, int K, int act,
            const int32_t* __restrict__ perm, const int32_t* __restrict__ ptr, int D,
            long long wstride, int scale_by_group) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int rows[BM];
  TileInfo ti;
  if (!locate_tile(blockIdx.x, BM, ptr, D, M, &ti)) return;
  const int n0 = blockIdx.y * BN;
  const int t = threadIdx.x;
  if (t < BM) rows[t] = t < ti.rcnt ? (perm ? __ldg(perm + ti.r0 + t) : ti.r0 + t) : -1;
  const float* Wg = W + (size_t)ti.group * wstride;
  const float gdiv = (float)(ti.group + 1);
  __syncthreads();
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int a_row = t >> 2, a_k = (t & 3) * 4;
  const int grow = rows[a_row];
  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile: 64 rows x 16 k
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int k = k0 + a_k + i;
      float v = 0.f;
      if (grow >= 0 && k < K) {
        v = __ldg(A + (size_t)grow * lda + k);
        if (!TRANSB && scale_by_group) v = v / gdiv;
      }
      As[a_k + i][a_row] = v;
    }
    // B tile: 16 k x 64 n
    if (!TRANSB) {
      int kk = t >> 4, nq = (t & 15) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = k0 + kk, n = n0 + nq + i;
        Bs[kk][nq + i] = (k < K && n < N) ? __ldg(Wg + (size_t)k * N + n) : 0.f;
      }
    } else {
      int nn = t >> 2, kq = (t & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = k0 + kq + i, n = n0 + nn;
        Bs[kq + i][nn] = (k < K && n < N) ? __ldg(Wg + (size_t)n * K + k) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = rows[ty * 4 + i];
    if (r < 0) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) {
        float v = acc[i][j];
        if (TRANSB && scale_by_group) v = v / gdiv;
        C[(size_t)r * ldc + n] = act_apply(act, v);
      }
    }
  }
}

static int gemm_rows(bool transb, const float* A, int lda, const float* W, float* C, int ldc,
                     int64_t M, int N, int K, int act, const GroupDesc& gd) {
  if (M == 0) return ATHENA_OK;
  ATH_REQUIRE(N >= 1 && K >= 1, ATHENA_ERR_ARG, "gemm: bad shape N=%d K=%d", N, K);
  int64_t tiles = cdiv(M, BM) + (gd.ptr ? gd.D : 0);
  dim3 grid((unsigned)tiles, (unsigned)cdiv(N, BN));
  cudaStream_t st = ctx().stream;
  if (transb)
    k_gemm_rows<true><<<grid, 256, 0, st>>>(A, lda, W, C, ldc, M, N, K, act, gd.perm, gd.ptr,
                                            gd.D, gd.wstride, gd.scale_by_group);
  else
    k_gemm_rows<false><<<grid, 256, 0, st>>>(A, lda, W, C, ldc, M, N, K, act, gd.perm, gd.ptr,
                                             gd.D, gd.wstride, gd.scale_by_group);
  ATH_LAUNCHED_T(transb ? "gemm_nt" : "gemm_nn");
  return ATHENA_OK;
}

int launch_gemm_nn(const float* A, int lda, const float* W, float* C, int ldc, int64_t M, int N,
                   int K, int act, const GroupDesc& gd) {
  int epi = (act == ATHENA_ACT_SOFTMAX) ? ATHENA_ACT_NONE : act;
  ATH_TRY(gemm_rows(false, A, lda, W, C, ldc, M, N, K, epi, gd));
  if (act == ATHENA_ACT_SOFTMAX) {
    ATH_REQUIRE(ldc == N, ATHENA_ERR_ARG, "gemm_nn: softmax epilogue needs a dense C");
    ATH_TRY(launch_softmax_rows(C, M, N));
  }
  return ATHENA_OK;
}

int launch_gemm_nt(const float* A, int lda, const float* W, float* C, int ldc, int64_t M, int N,
                   int K, const GroupDesc& gd) {
  return gemm_rows(true, A, lda, W, C, ldc, M, N, K, ATHENA_ACT_NONE, gd);
}

// ---- dW = A^T . G, split over rows -------------------------------------------
constexpr int TN_CHUNK = 512;  // rows per partial
constexpr int TN_RK = 16;      // rows staged per step

// part[chunk][k][n] = sum_{rows of chunk} (A[row,k]/s) * G[row,n]  for the
// (k-tile, n-tile) = (blockIdx.y, blockIdx.z) of this block.
__global__ void __launch_bounds__(256)
k_gemm_tn_partial(const float* __restrict__ A, int lda, const float* __restrict__ G, int ldg,
                  float* __restrict__ part, long long M, int N, int K,
                  const int32_t* __restrict__ perm, const int32_t* __restrict__ ptr, int D,
                  int scale_by_group) {
  __shared__ float As[TN_RK][BM + 4];  // [row][k]
  __shared__ float Gs[TN_RK][BN + 4];  // [row][n]
  TileInfo ti;
  if (!locate_tile(blockIdx.x, TN_CHUNK, ptr, D, M, &ti)) return;
  const int k0 = blockIdx.y * BM, n0 = blockIdx.z * BN;
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const float gdiv = (float)(ti.group + 1);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = t >> 4;          // 0..15 : staged row
  const int lc = (t & 15) * 4;    // 0..60 : first of 4 columns
  for (int rb = 0; rb < ti.rcnt; rb += TN_RK) {
    int r = rb + lr;
    int grow = -1;
    if (r < ti.rcnt) grow = perm ? __ldg(perm + ti.r0 + r) : ti.r0 + r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int k = k0 + lc + i, n = n0 + lc + i;
      float a = 0.f, g = 0.f;
      if (grow >= 0) {
        if (k < K) {
          a = __ldg(A + (size_t)grow * lda + k);
          if (scale_by_group) a = a / gdiv;
        }
        if (n < N) g = __ldg(G + (size_t)grow * ldg + n);
      }
      As[lr][lc + i] = a;
      Gs[lr][lc + i] = g;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < TN_RK; ++rr) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[rr][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Gs[rr][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dst = part + (size_t)blockIdx.x * K * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int k = k0 + ty * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) dst[(size_t)k * N + n] = acc[i][j];
    }
  }
}

// dW[g][e] += sum over the chunks of group g of part[chunk][e].  32 elements x 8 lanes per
// block: lane l adds chunks c0+l, c0+l+8, ... into four running sums (fixed pattern, four
// loads in flight), the eight lane totals are then added in lane order -- deterministic.
constexpr int TNR_ELEMS = 32, TNR_LANES = 8;
__global__ void __launch_bounds__(TNR_ELEMS * TNR_LANES)
k_gemm_tn_reduce(const float* __restrict__ part, float* __restrict__ dW, long long M, int KN,
                 const int32_t* __restrict__ ptr, int D, long long wstride) {
  __shared__ float red[TNR_LANES][TNR_ELEMS];
  const int g = blockIdx.y;
  const int el = threadIdx.x % TNR_ELEMS, l = threadIdx.x / TNR_ELEMS;
  const int e = blockIdx.x * TNR_ELEMS + el;
  int c0 = 0, c1 = 0;
  if (ptr == nullptr) {
    c1 = (int)((M + TN_CHUNK - 1) / TN_CHUNK);
  } else {
    for (int d = 0; d <= g; ++d) {
      int nt = (__ldg(ptr + d + 1) - __ldg(ptr + d) + TN_CHUNK - 1) / TN_CHUNK;
      c0 = c1;
      c1 += nt;
    }
  }
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (e < KN) {
    const float* src = part + e;
    int c = c0 + l;
    for (; c + 3 * TNR_LANES < c1; c += 4 * TNR_LANES) {
      const float v0 = src[(size_t)c * KN], v1 = src[(size_t)(c + TNR_LANES) * KN];
      const float v2 = src[(size_t)(c + 2 * TNR_LANES) * KN];
      const float v3 = src[(size_t)(c + 3 * TNR_LANES) * KN];
      acc[0] += v0;
      acc[1] += v1;
      acc[2] += v2;
      acc[3] += v3;
    }
    for (int k = 0; c < c1; c += TNR_LANES, ++k) acc[k] += src[(size_t)c * KN];
  }
  red[l][el] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __syncthreads();
  if (l == 0 && e < KN && c1 > c0) {
    float s = red[0][el];
#pragma unroll
    for (int k = 1; k < TNR_LANES; ++k) s += red[k][el];
    dW[(size_t)g * wstride + e] += s;
  }
}

int launch_gemm_tn(const float* A, int lda, const float* G, int ldg, float* dW, int64_t M, int N,
                   int K, const GroupDesc& gd, DevBuf& scratch) {
  if (M == 0) return ATHENA_OK;
  int D = gd.ptr ? gd.D : 1;
  int64_t chunks = cdiv(M, TN_CHUNK) + (gd.ptr ? gd.D : 0);
  ATH_TRY(scratch.reserve(sizeof(float) * (size_t)chunks * K * N));
  cudaStream_t st = ctx().stream;
  dim3 grid((unsigned)chunks, (unsigned)cdiv(K, BM), (unsigned)cdiv(N, BN));
  k_gemm_tn_partial<<<grid, 256, 0, st>>>(A, lda, G, ldg, scratch.as<float>(), M, N, K, gd.perm,
                                          gd.ptr, D, gd.scale_by_group);
  ATH_LAUNCHED_T("gemm_tn_partial");
  dim3 rgrid((unsigned)cdiv((int64_t)K * N, TNR_ELEMS), (unsigned)D);
  k_gemm_tn_reduce<<<rgrid, TNR_ELEMS * TNR_LANES, 0, st>>>(scratch.as<float>(), dW, M, K * N, gd.ptr, D,
                                          gd.wstride);
  ATH_LAUNCHED_T("gemm_tn_reduce");
  return ATHENA_OK;
}

// ---- elementwise / row-wise ----------------------------------------------------

// out may alias G (in-place use): no __restrict__ on those two
__global__ void k_act_bwd(int act, const float* __restrict__ Y, const float* G, float* out,
                          long long n) {
  pdl_wait();  // launched programmatically dependent
  pdl_launch_dependents();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = act_grad(act, Y[i], G[i]);
}

// one warp per row: softmax backward  dx = y*g - y*sum(y*g)
// (athena_diffstruc_extd_sub.f90:369-373); gidx selects the upstream row
// (nullptr = same row), which lets the readout broadcast gout[vgraph[v]].
__global__ void k_softmax_bwd_rows(const float* __restrict__ Y, const float* G,
                                   const int32_t* __restrict__ gidx, float* out, long long M,
                                   int N) {
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* y = Y + (size_t)row * N;
  const float* g = G + (size_t)(gidx ? gidx[row] : row) * N;
  float s = 0.f;
  for (int i = lane; i < N; i += 32) s += y[i] * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  for (int i = lane; i < N; i += 32) out[(size_t)row * N + i] = y[i] * g[i] - y[i] * s;
}

__global__ void k_act_bwd_bcast(int act, const float* __restrict__ Y, const float* __restrict__ G,
                                const int32_t* __restrict__ gidx, float* __restrict__ out,
                                long long M, int N) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  long long row = i / N;
  int c = (int)(i - row * N);
  out[i] = act_grad(act, Y[i], G[(size_t)gidx[row] * N + c]);
}

__global__ void k_bias_act(float* __restrict__ Y, const float* __restrict__ bias, long long n,
                           int N, int act) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = Y[i];
  if (bias) v += __ldg(bias + (int)(i % N));
  Y[i] = act_apply(act, v);
}

__global__ void k_colsum_add(const float* __restrict__ G, long long M, int N,
                             float* __restrict__ dst) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (long long m = 0; m < M; ++m) s += G[m * N + n];
  dst[n] += s;
}

int launch_bias_act(float* Y, const float* bias, int64_t M, int N, int act) {
  if (M == 0) return ATHENA_OK;
  const int epi = act == ATHENA_ACT_SOFTMAX ? ATHENA_ACT_NONE : act;
  if (bias != nullptr || (epi != ATHENA_ACT_NONE && epi != ATHENA_ACT_LINEAR)) {
    k_bias_act<<<(unsigned)cdiv(M * N, 256), 256, 0, ctx().stream>>>(Y, bias, M * N, N, epi);
    ATH_LAUNCHED_T("bias_act");
  }
  if (act == ATHENA_ACT_SOFTMAX) ATH_TRY(launch_softmax_rows(Y, M, N));
  return ATHENA_OK;
}

int launch_colsum_add(const float* G, int64_t M, int N, float* dst) {
  if (M == 0) return ATHENA_OK;
  k_colsum_add<<<(unsigned)cdiv(N, 128), 128, 0, ctx().stream>>>(G, M, N, dst);
  ATH_LAUNCHED_T("colsum_add");
  return ATHENA_OK;
}

int launch_act_bwd(int act, const float* Y, const float* G, float* out, int64_t M, int N) {
  if (M == 0) return ATHENA_OK;
  cudaStream_t st = ctx().stream;
  if (act == ATHENA_ACT_SOFTMAX) {
    k_softmax_bwd_rows<<<(unsigned)cdiv(M * 32, 256), 256, 0, st>>>(Y, G, nullptr, out, M, N);
  } else {
    int64_t n = M * N;
    int blocks = (int)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx().sm_count * 16);
    ATH_CUDA(launch_pdl(k_act_bwd, dim3(blocks), dim3(256), 0, st, act, Y, G, out, (long long)n));
  }
  ATH_LAUNCHED_T("act_bwd");
  return ATHENA_OK;
}

int launch_readout_bwd(int act, const float* S, const float* gout, const int32_t* vgraph,
                       float* dY, int64_t V, int N) {
  if (V == 0) return ATHENA_OK;
  cudaStream_t st = ctx().stream;
  if (act == ATHENA_ACT_SOFTMAX)
    k_softmax_bwd_rows<<<(unsigned)cdiv(V * 32, 256), 256, 0, st>>>(S, gout, vgraph, dY, V, N);
  else
    k_act_bwd_bcast<<<(unsigned)cdiv(V * N, 256), 256, 0, st>>>(act, S, gout, vgraph, dY, V, N);
  ATH_LAUNCHED_T("readout_bwd");
  return ATHENA_OK;
}

// one warp per row, in place: y = exp(x - max) / sum   (athena_diffstruc_extd_sub.f90:309-313)
__global__ void k_softmax_rows(float* __restrict__ Y, long long M, int N) {
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= M) return;
  float* y = Y + (size_t)row * N;
  float mx = -INFINITY;
  for (int i = lane; i < N; i += 32) mx = fmaxf(mx, y[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int i = lane; i < N; i += 32) {
    float e = expf(y[i] - mx);
    y[i] = e;
    s += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  for (int i = lane; i < N; i += 32) y[i] = y[i] / s;
}

int launch_softmax_rows(float* Y, int64_t M, int N) {
  if (M == 0) return ATHENA_OK;
  k_softmax_rows<<<(unsigned)cdiv(M * 32, 256), 256, 0, ctx().stream>>>(Y, M, N);
  ATH_LAUNCHED_T("softmax_rows");
  return ATHENA_OK;
}

// thread per (graph, column): sequential over the graph's vertices, ascending
// (sum(ptr2, dim=2), athena_duvenaud_msgpass_layer.f90:848-852)
__global__ void k_segment_sum(const float* __restrict__ Y, int N, const int32_t* __restrict__ voff,
                              int B, float* __restrict__ out, int accumulate) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  int s = (int)(i / N), c = (int)(i - (long long)s * N);
  int v0 = voff[s], v1 = voff[s + 1];
  float acc = 0.f;
  for (int v = v0; v < v1; ++v) acc += Y[(size_t)v * N + c];
  out[i] = accumulate ? out[i] + acc : acc;
}

int launch_segment_sum(const float* Y, int N, const int32_t* voff, int32_t B, float* out,
                       int accumulate) {
  k_segment_sum<<<(unsigned)cdiv((int64_t)B * N, 128), 128, 0, ctx().stream>>>(Y, N, voff, B, out,
                                                                              accumulate);
  ATH_LAUNCHED_T("segment_sum");
  return ATHENA_OK;
}

// swish_array / get_partial_swish_val with beta = 1 (athena_diffstruc_extd_sub.f90:434, 480-484)
__global__ void k_swish_fwd(const float* __restrict__ X, float* __restrict__ H, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float x = X[i];
    H[i] = x * (1.f / (1.f + expf(-x)));
  }
}
__global__ void k_swish_bwd(const float* __restrict__ X, const float* G, float* out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float x = X[i], e = expf(x), d = e + 1.f;
    out[i] = G[i] * e * (x + e + 1.f) / (d * d);
  }
}
int launch_swish_fwd(const float* X, float* H, int64_t n) {
  if (n == 0) return ATHENA_OK;
  int blocks = (int)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx().sm_count * 16);
  k_swish_fwd<<<blocks, 256, 0, ctx().stream>>>(X, H, n);
  ATH_LAUNCHED_T("swish_fwd");
  return ATHENA_OK;
}
int launch_swish_bwd(const float* X, const float* G, float* out, int64_t n) {
  if (n == 0) return ATHENA_OK;
  int blocks = (int)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx().sm_count * 16);
  k_swish_bwd<<<blocks, 256, 0, ctx().stream>>>(X, G, out, n);
  ATH_LAUNCHED_T("swish_bwd");
  return ATHENA_OK;
}

__global__ void k_copy_cols(float* __restrict__ dst, int ldd, int doff, const float* __restrict__ src,
                            int lds, int soff, int w, long long M, int accumulate) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < M * w; i += stride) {
    const long long m = i / w;
    const int c = (int)(i - m * w);
    const float v = src[m * lds + soff + c];
    float* d = dst + m * ldd + doff + c;
    *d = accumulate ? *d + v : v;
  }
}
int launch_copy_cols(float* dst, int ldd, int doff, const float* src, int lds, int soff, int w,
                     int64_t M, int accumulate) {
  if (M == 0 || w == 0) return ATHENA_OK;
  int blocks = (int)std::min<int64_t>(cdiv(M * w, 256), (int64_t)ctx().sm_count * 16);
  k_copy_cols<<<blocks, 256, 0, ctx().stream>>>(dst, ldd, doff, src, lds, soff, w, M, accumulate);
  ATH_LAUNCHED_T("copy_cols");
  return ATHENA_OK;
}

__global__ void k_add_inplace(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] += src[i];
}

int launch_add_inplace(float* dst, const float* src, int64_t n) {
  if (n == 0) return ATHENA_OK;
  int blocks = (int)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx().sm_count * 16);
  k_add_inplace<<<blocks, 256, 0, ctx().stream>>>(dst, src, n);
  ATH_LAUNCHED_T("add_inplace");
  return ATHENA_OK;
}

}  // namespace athena
