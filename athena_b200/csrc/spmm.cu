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
#include <climits>

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
            const float* __restrict__ tail, int tail_n, int skip_len) {
  const int lane = threadIdx.x & 31;
  const int lg = lane & (G - 1);
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (row >= V) return;  // the whole group leaves together
  const int beg = __ldg(row_ptr + row), end = __ldg(row_ptr + row + 1);
  if (end - beg > skip_len) return;  // long row: k_aggregate_long owns it
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
      if (MAXC == 1 && G >= 4) {
        // four neighbour rows in flight per lane group: the gathers of a batch are issued
        // before the first of them is consumed (a row of a few hundred entries is otherwise a
        // chain of dependent L2 / HBM round trips); the sums still run in entry order
        const int ch = cbase + lg;
        for (int j = 0; j < cnt; j += 4) {
          int u[4];
          float c[4];
          Vec<VEC> x[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            u[jj] = __shfl_sync(gmask, mycol, (j + jj) & (G - 1), G);
            c[jj] = HAS_COEF ? __shfl_sync(gmask, myc, (j + jj) & (G - 1), G) : 1.f;
            if (j + jj >= cnt) u[jj] = -1;
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) x[jj].v[i] = 0.f;
            if (u[jj] >= 0 && ch < nch) x[jj].load(X + (size_t)u[jj] * ldx + ch * VEC);
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            if (u[jj] < 0) continue;  // "no edge feature" marker / past the row's end
#pragma unroll
            for (int i = 0; i < VEC; ++i)
              acc[0][i] = HAS_COEF ? fmaf(c[jj], x[jj].v[i], acc[0][i]) : acc[0][i] + x[jj].v[i];
          }
        }
        continue;
      }
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

// One CTA per long row (more than LONG_ROW entries; power-law hubs): the row's entries are
// cut into NG contiguous segments, one per lane group, each summed in ascending order; the
// NG partial rows are then added in segment order by group 0.  Fixed shape, no atomics:
// deterministic (the association differs from the reference's single sequential sum by
// fp32 rounding only).  VEC = 4, F / 4 <= 32 chunks.
template <int G, bool HAS_COEF>
__global__ void __launch_bounds__(256)
k_aggregate_long(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                 const float* __restrict__ coef, const float* __restrict__ X, int ldx, int F,
                 float* __restrict__ out, int ldo, int accumulate,
                 const float* __restrict__ tail, int tail_n,
                 const int32_t* __restrict__ long_list, const int32_t* __restrict__ long_count) {
  constexpr int NG = 256 / G;
  __shared__ float4 part[NG][G];
  const int g = threadIdx.x / G, lg = threadIdx.x % G;
  const int nch = F / 4;
  const int count = *long_count;
  for (int q = blockIdx.x; q < count; q += gridDim.x) {
    const int row = long_list[q];
    const int beg = __ldg(row_ptr + row), end = __ldg(row_ptr + row + 1);
    const int seg = (end - beg + NG - 1) / NG;
    const int s0 = min(end, beg + g * seg), s1 = min(end, s0 + seg);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lg < nch) {
      int w = s0;
      for (; w + 4 <= s1; w += 4) {
        int u[4];
        float c[4];
        float4 x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          u[k] = __ldg(col + w + k);
          c[k] = HAS_COEF ? __ldg(coef + w + k) : 1.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          x[k] = u[k] >= 0 ? __ldg(reinterpret_cast<const float4*>(X + (size_t)u[k] * ldx) + lg)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (u[k] < 0) continue;
          acc.x = HAS_COEF ? fmaf(c[k], x[k].x, acc.x) : acc.x + x[k].x;
          acc.y = HAS_COEF ? fmaf(c[k], x[k].y, acc.y) : acc.y + x[k].y;
          acc.z = HAS_COEF ? fmaf(c[k], x[k].z, acc.z) : acc.z + x[k].z;
          acc.w = HAS_COEF ? fmaf(c[k], x[k].w, acc.w) : acc.w + x[k].w;
        }
      }
      for (; w < s1; ++w) {
        const int u = __ldg(col + w);
        if (u < 0) continue;
        const float c = HAS_COEF ? __ldg(coef + w) : 1.f;
        const float4 x = __ldg(reinterpret_cast<const float4*>(X + (size_t)u * ldx) + lg);
        acc.x = HAS_COEF ? fmaf(c, x.x, acc.x) : acc.x + x.x;
        acc.y = HAS_COEF ? fmaf(c, x.y, acc.y) : acc.y + x.y;
        acc.z = HAS_COEF ? fmaf(c, x.z, acc.z) : acc.z + x.z;
        acc.w = HAS_COEF ? fmaf(c, x.w, acc.w) : acc.w + x.w;
      }
    }
    part[g][lg] = acc;
    __syncthreads();
    if (g == 0) {
      float* orow = out + (size_t)row * ldo;
      if (lg < nch) {
        float4 r = part[0][lg];
        for (int k = 1; k < NG; ++k) {
          const float4 p = part[k][lg];
          r.x += p.x; r.y += p.y; r.z += p.z; r.w += p.w;
        }
        float4* dst = reinterpret_cast<float4*>(orow) + lg;
        if (accumulate) {
          const float4 o = *dst;
          r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
        }
        *dst = r;
      }
      if (tail != nullptr)
        for (int i = lg; i < tail_n; i += G) orow[F + i] = __ldg(tail + (size_t)row * tail_n + i);
    }
    __syncthreads();
  }
}

template <int VEC, int G, int MAXC>
static int launch_agg_t(const int32_t* row_ptr, const int32_t* col, const float* coef,
                        const float* X, int ldx, int F, float* out, int ldo, int64_t V,
                        int accumulate, const float* tail, int tail_n, const int32_t* long_list,
                        const int32_t* long_count) {
  const int threads = 256;
  const int64_t blocks = cdiv(V * G, threads);
  cudaStream_t st = ctx().stream;
  const bool split = VEC == 4 && MAXC == 1 && long_list != nullptr && long_count != nullptr;
  const int skip_len = split ? LONG_ROW : INT_MAX;
  if (coef)
    k_aggregate<VEC, G, MAXC, true><<<(unsigned)blocks, threads, 0, st>>>(
        row_ptr, col, coef, X, ldx, F, out, ldo, V, accumulate, tail, tail_n, skip_len);
  else
    k_aggregate<VEC, G, MAXC, false><<<(unsigned)blocks, threads, 0, st>>>(
        row_ptr, col, coef, X, ldx, F, out, ldo, V, accumulate, tail, tail_n, skip_len);
  if (split) {
    if constexpr (VEC == 4 && MAXC == 1) {
      const int grid = 2 * ctx().sm_count;
      if (coef)
        k_aggregate_long<G, true><<<grid, 256, 0, st>>>(row_ptr, col, coef, X, ldx, F, out, ldo,
                                                        accumulate, tail, tail_n, long_list,
                                                        long_count);
      else
        k_aggregate_long<G, false><<<grid, 256, 0, st>>>(row_ptr, col, coef, X, ldx, F, out, ldo,
                                                         accumulate, tail, tail_n, long_list,
                                                         long_count);
      ctx().launches.fetch_add(1, std::memory_order_relaxed);
      if (ctx().profiling) prof_mark("aggregate_long");
    }
  }
  static const std::string tag = std::string("aggregate_v") + std::to_string(VEC) + "_g" +
                                 std::to_string(G) + "_c" + std::to_string(MAXC);
  static const std::string tag_coef = tag + "_coef";
  ATH_LAUNCHED_T(coef ? tag_coef.c_str() : tag.c_str());
  return ATHENA_OK;
}

template <int VEC>
static int launch_agg_v(const int32_t* row_ptr, const int32_t* col, const float* coef,
                        const float* X, int ldx, int F, float* out, int ldo, int64_t V,
                        int accumulate, const float* tail, int tail_n, const int32_t* long_list,
                        const int32_t* long_count) {
  const int nch = F / VEC;
#define ATH_AGG(G, MAXC)                                                                     \
  return launch_agg_t<VEC, G, MAXC>(row_ptr, col, coef, X, ldx, F, out, ldo, V, accumulate, \
                                    tail, tail_n, long_list, long_count)
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
                     int accumulate, const float* tail, int tail_n, const int32_t* long_list,
                     const int32_t* long_count) {
  if (V == 0) return ATHENA_OK;
  ATH_REQUIRE(F >= 1 && ldx >= F && ldo >= F + tail_n, ATHENA_ERR_ARG,
              "aggregate: bad shape F=%d ldx=%d ldo=%d tail=%d", F, ldx, ldo, tail_n);
  const bool vec4 = (F % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (vec4)
    return launch_agg_v<4>(row_ptr, col, coef, X, ldx, F, out, ldo, V, accumulate, tail, tail_n,
                           long_list, long_count);
  return launch_agg_v<1>(row_ptr, col, coef, X, ldx, F, out, ldo, V, accumulate, tail, tail_n,
                         nullptr, nullptr);
}

}  // namespace athena
