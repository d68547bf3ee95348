// Neighbour aggregation: row-gather SpMM over the batch CSR (forward) or its
// CSC transpose (backward).  Deterministic: every output row is owned by one
// group of G lanes that walks the row's entries in ascending order, exactly
// the summation order of the reference loops
//   kipf_propagate              athena_diffstruc_extd_sub_kipf.f90:29-46
//   get_partial_kipf_..._val    athena_diffstruc_extd_sub_kipf.f90:101-109  (via CSC)
//   duvenaud_propagate          athena_diffstruc_extd_sub_duvenaud.f90:34-42
//   get_partial_duvenaud_..     athena_diffstruc_extd_sub_duvenaud.f90:136-141 (via CSC)
// No float atomics anywhere.
//
// Layout: features of one vertex are contiguous (Fortran val(F,V)), so a group
// of G lanes reads one neighbour row with G coalesced 16-byte (VEC=4) loads;
// the group's index/coefficient loads are one coalesced load per G entries,
// broadcast with sub-warp shuffles.
#include "athena_internal.h"

namespace athena {

template <int VEC>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
  __device__ __forceinline__ void load_rw(const float* p) { v[0] = *p; }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void load_rw(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <int VEC, int G, int MAXC, bool HAS_COEF>
__global__ void __launch_bounds__(256)
k_aggregate(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
            const float* __restrict__ coef, const float* __restrict__ X, int ldx, int F,
            float* __restrict__ out, int ldo, long long V, int accumulate,
            const float* __restrict__ tail, int tail_n) {
  const int lane = threadIdx.x & 31;
  const int lg = lane & (G - 1);
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (row >= V) return;  // the whole group leaves together
  const int beg = __ldg(row_ptr + row), end = __ldg(row_ptr + row + 1);
  const int nch = F / VEC;
  float* orow = out + (size_t)row * ldo;
  for (int cbase = 0; cbase < nch; cbase += G * MAXC) {
    float acc[MAXC][VEC];
#pragma unroll
    for (int k = 0; k < MAXC; ++k)
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[k][i] = 0.f;
    for (int w0 = beg; w0 < end; w0 += G) {
      const int myw = w0 + lg;
      int mycol = -1;
      float myc = 0.f;
      if (myw < end) {
        mycol = __ldg(col + myw);
        if (HAS_COEF) myc = __ldg(coef + myw);
      }
      const int cnt = min(G, end - w0);
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const int u = __shfl_sync(gmask, mycol, j, G);
        const float c = HAS_COEF ? __shfl_sync(gmask, myc, j, G) : 1.f;
        if (u < 0) continue;  // "no edge feature" marker
        const float* xr = X + (size_t)u * ldx;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
          const int ch = cbase + lg + k * G;
          if (ch < nch) {
            Vec<VEC> x;
            x.load(xr + ch * VEC);
#pragma unroll
            for (int i = 0; i < VEC; ++i)
              acc[k][i] = HAS_COEF ? fmaf(c, x.v[i], acc[k][i]) : acc[k][i] + x.v[i];
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      const int ch = cbase + lg + k * G;
      if (ch < nch) {
        Vec<VEC> r;
        if (accumulate) {
          r.load_rw(orow + ch * VEC);
#pragma unroll
          for (int i = 0; i < VEC; ++i) r.v[i] += acc[k][i];
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i) r.v[i] = acc[k][i];
        }
        r.store(orow + ch * VEC);
      }
    }
  }
  if (tail != nullptr)
    for (int i = lg; i < tail_n; i += G) orow[F + i] = __ldg(tail + (size_t)row * tail_n + i);
}

template <int VEC, int G, int MAXC>
static int launch_agg_t(const int32_t* row_ptr, const int32_t* col, const float* coef,
                        const float* X, int ldx, int F, float* out, int ldo, int64_t V,
                        int accumulate, const float* tail, int tail_n) {
  const int threads = 256;
  const int64_t blocks = cdiv(V * G, threads);
  cudaStream_t st = ctx().stream;
  if (coef)
    k_aggregate<VEC, G, MAXC, true><<<(unsigned)blocks, threads, 0, st>>>(
        row_ptr, col, coef, X, ldx, F, out, ldo, V, accumulate, tail, tail_n);
  else
    k_aggregate<VEC, G, MAXC, false><<<(unsigned)blocks, threads, 0, st>>>(
        row_ptr, col, coef, X, ldx, F, out, ldo, V, accumulate, tail, tail_n);
  static const std::string tag = std::string("aggregate_v") + std::to_string(VEC) + "_g" +
                                 std::to_string(G) + "_c" + std::to_string(MAXC);
  static const std::string tag_coef = tag + "_coef";
  ATH_LAUNCHED_T(coef ? tag_coef.c_str() : tag.c_str());
  return ATHENA_OK;
}

template <int VEC>
static int launch_agg_v(const int32_t* row_ptr, const int32_t* col, const float* coef,
                        const float* X, int ldx, int F, float* out, int ldo, int64_t V,
                        int accumulate, const float* tail, int tail_n) {
  const int nch = F / VEC;
#define ATH_AGG(G, MAXC)                                                                     \
  return launch_agg_t<VEC, G, MAXC>(row_ptr, col, coef, X, ldx, F, out, ldo, V, accumulate, \
                                    tail, tail_n)
  if (nch <= 4) ATH_AGG(4, 1);
  if (nch <= 8) ATH_AGG(8, 1);
  if (nch <= 16) ATH_AGG(16, 1);
  if (nch <= 32) ATH_AGG(32, 1);
  ATH_AGG(32, 4);
#undef ATH_AGG
}

// tail: optional [V, tail_n] block copied to out[:, F:F+tail_n] (the
// time-step-invariant edge-feature aggregate of the Duvenaud layer).
int launch_aggregate(const int32_t* row_ptr, const int32_t* col, const float* coef,
                          const float* X, int ldx, int F, float* out, int ldo, int64_t V,
                          int accumulate, const float* tail, int tail_n) {
  if (V == 0) return ATHENA_OK;
  ATH_REQUIRE(F >= 1 && ldx >= F && ldo >= F + tail_n, ATHENA_ERR_ARG,
              "aggregate: bad shape F=%d ldx=%d ldo=%d tail=%d", F, ldx, ldo, tail_n);
  const bool vec4 = (F % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (vec4)
    return launch_agg_v<4>(row_ptr, col, coef, X, ldx, F, out, ldo, V, accumulate, tail, tail_n);
  return launch_agg_v<1>(row_ptr, col, coef, X, ldx, F, out, ldo, V, accumulate, tail, tail_n);
}

}  // namespace athena
