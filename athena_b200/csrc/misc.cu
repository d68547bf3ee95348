// MSE loss (+ its gradient seed) and the clip + optimiser step.
//   compute_mse            athena_loss.f90:393-430   (mean over all elements / 2,
//                          summed over cells; pinned by test/test_loss.f90:59-67,135-143)
//   clip_type%apply        athena_clipper.f90:165-207
//   minimise_sgd           athena_optimiser.f90:634-673
//   minimise_adam          athena_optimiser.f90:1027-1091
//   network%update         athena_network_sub.f90:2816-2929 (clip -> minimise -> zero grads)
// All reductions are two-pass with a fixed grid, so results are run-to-run
// identical (no float atomics).
#include <algorithm>
#include <cmath>

#include "athena_internal.h"

namespace athena {

constexpr int RED_THREADS = 256;

// Sum over a block of THREADS threads (the launch must use exactly that many); fixed
// combine pattern, result valid in warp 0.
template <int THREADS = RED_THREADS>
__device__ __forceinline__ float block_sum(float v) {
  static_assert(THREADS % 32 == 0 && THREADS <= 1024, "block_sum: whole warps, one block");
  __shared__ float ws[THREADS / 32];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) ws[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < THREADS / 32 ? ws[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

__device__ __forceinline__ float mse_act_grad(int act, float y, float g) {
  switch (act) {
    case ATHENA_ACT_RELU: return y > 0.f ? g : 0.f;
    case ATHENA_ACT_LEAKY_RELU: return y > 0.f ? g : g * 0.01f;
    case ATHENA_ACT_SIGMOID: return g * (y * (1.f - y));
    case ATHENA_ACT_TANH: return g * (1.f - y * y);
    default: return g;
  }
}

// graph-output MSE: cell s = graph s, N_cell = F * nv[s].  VEC elements per thread step
// (VEC = 4 needs F % 4 == 0 so that a float4 never straddles two vertices).
// act != NONE folds act'(pred) of the producing layer into the gradient seed.
template <int VEC>
__global__ void __launch_bounds__(RED_THREADS)
k_mse_graph(const float* __restrict__ pred, const float* __restrict__ target,
            const int32_t* __restrict__ vgraph, const int32_t* __restrict__ nv, int F, long long n,
            int act, float* __restrict__ grad, float* __restrict__ partial) {
  float local = 0.f;
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  const long long stride = (long long)gridDim.x * blockDim.x * VEC;
  for (; i < n; i += stride) {
    const long long v = i / F;
    const float denom = (float)(F * __ldg(nv + __ldg(vgraph + v)));
    float p[VEC], t[VEC], g[VEC];
    if (VEC == 4) {
      const float4 p4 = *reinterpret_cast<const float4*>(pred + i);
      const float4 t4 = *reinterpret_cast<const float4*>(target + i);
      p[0] = p4.x; p[1] = p4.y; p[2] = p4.z; p[3] = p4.w;
      t[0] = t4.x; t[1] = t4.y; t[2] = t4.z; t[3] = t4.w;
    } else {
      p[0] = pred[i];
      t[0] = target[i];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float d = p[k] - t[k];
      g[k] = mse_act_grad(act, p[k], d / denom);
      local += d * d / denom;
    }
    if (VEC == 4)
      *reinterpret_cast<float4*>(grad + i) = make_float4(g[0], g[1], g[2], g[3]);
    else
      grad[i] = g[0];
  }
  float s = block_sum(local);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(RED_THREADS)
k_mse_array(const float* __restrict__ pred, const float* __restrict__ target, long long n,
            float denom, float* __restrict__ grad, float* __restrict__ partial) {
  float local = 0.f;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float d = pred[i] - target[i];
    grad[i] = d / denom;
    local += d * d / denom;
  }
  float s = block_sum(local);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// loss_acc[0] += 0.5 * sum(partial)   (single block, fixed order)
__global__ void __launch_bounds__(RED_THREADS)
k_loss_finish(const float* __restrict__ partial, int nb, float* __restrict__ loss_acc) {
  float local = 0.f;
  for (int i = threadIdx.x; i < nb; i += RED_THREADS) local += partial[i];
  float s = block_sum(local);
  if (threadIdx.x == 0) loss_acc[0] += 0.5f * s;
}

static int red_blocks(int64_t n) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, RED_THREADS * 4), 1024));
}

int launch_mse_graph(const float* pred, const float* target, const int32_t* vgraph,
                     const int32_t* nv, int F, int64_t V, int act, float* grad, float* loss_acc,
                     DevBuf& scratch) {
  int64_t n = V * F;
  if (n == 0) return ATHENA_OK;
  ATH_TRY(scratch.reserve(sizeof(float) * 1024));
  cudaStream_t st = ctx().stream;
  const bool vec4 = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(pred) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(target) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(grad) & 15) == 0);
  int nb = red_blocks(vec4 ? n / 4 : n);
  if (vec4)
    k_mse_graph<4><<<nb, RED_THREADS, 0, st>>>(pred, target, vgraph, nv, F, n, act, grad,
                                               scratch.as<float>());
  else
    k_mse_graph<1><<<nb, RED_THREADS, 0, st>>>(pred, target, vgraph, nv, F, n, act, grad,
                                               scratch.as<float>());
  ATH_LAUNCHED_T("mse_graph");
  k_loss_finish<<<1, RED_THREADS, 0, st>>>(scratch.as<float>(), nb, loss_acc);
  ATH_LAUNCHED_T("loss_finish");
  return ATHENA_OK;
}

int launch_loss_finish(const float* partial, int nb, float* loss_acc) {
  k_loss_finish<<<1, RED_THREADS, 0, ctx().stream>>>(partial, nb, loss_acc);
  ATH_LAUNCHED_T("loss_finish");
  return ATHENA_OK;
}

int launch_mse_array(const float* pred, const float* target, int64_t n, float denom, float* grad,
                     float* loss_acc, DevBuf& scratch) {
  if (n == 0) return ATHENA_OK;
  int nb = red_blocks(n);
  ATH_TRY(scratch.reserve(sizeof(float) * 1024));
  cudaStream_t st = ctx().stream;
  k_mse_array<<<nb, RED_THREADS, 0, st>>>(pred, target, n, denom, grad, scratch.as<float>());
  ATH_LAUNCHED_T("mse_array");
  k_loss_finish<<<1, RED_THREADS, 0, st>>>(scratch.as<float>(), nb, loss_acc);
  ATH_LAUNCHED_T("loss_finish");
  return ATHENA_OK;
}

// ---- clip + step ---------------------------------------------------------------

// elementwise clamp (written back) + partial sums of squares
__global__ void __launch_bounds__(RED_THREADS)
k_clip_sumsq(float* __restrict__ g, long long n, int clamp, float cmin, float cmax,
             float* __restrict__ partial) {
  float local = 0.f;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float x = g[i];
    if (clamp) {
      x = fmaxf(cmin, fminf(cmax, x));
      g[i] = x;
    }
    local += x * x;
  }
  float s = block_sum(local);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

struct StepArgs {
  int kind;
  float lr, beta1, beta2, eps, momentum;
  int nesterov;
  int norm_on;
  float clip_norm;
  int clamp;        // clip_type min / max (athena_clipper.f90:165-207): elementwise, before the norm
  float cmin, cmax;
  float bc1, bc2;
  int reg;          // ATHENA_REG_*
  float l1, l2;
  int l2_decoupled;
};

// one parameter of minimise_sgd / _adam / _rmsprop / _adagrad
// (athena_optimiser.f90:649-672, 1043-1088, 795-803, 919-924)
__device__ __forceinline__ void step_one(float* __restrict__ p, float* __restrict__ s1,
                                         float* __restrict__ s2, long long i, float gr,
                                         const StepArgs& a) {
  // clip_type%apply, min / max part: elementwise (idempotent: callers that clipped already, or
  // scaled a clipped value by a norm factor <= 1, are unaffected)
  if (a.clamp) gr = fmaxf(a.cmin, fminf(a.cmax, gr));
  if (a.reg != ATHENA_REG_NONE) {
    // regulariser%regularise inside minimise_* (athena_regulariser.f90:99, 117, 135-136),
    // in the reference's order of operations
    const float pv = p[i], sgn = copysignf(1.f, pv);
    if (a.reg == ATHENA_REG_L1)
      gr = gr + a.lr * a.l1 * sgn;
    else if (a.reg == ATHENA_REG_L2)
      gr = gr + a.lr * 2.f * a.l2 * pv;
    else
      gr = gr + a.lr * (a.l1 * sgn + 2.f * a.l2 * pv);
  }
  if (a.kind == ATHENA_OPT_SGD) {
    gr = -a.lr * gr;
    if (a.momentum > 1e-8f) {
      float vel = a.momentum * s1[i] + gr;
      s1[i] = vel;
      p[i] = a.nesterov ? p[i] + a.momentum * vel + gr : p[i] + vel;
    } else {
      s1[i] = gr;
      p[i] = p[i] + gr;
    }
  } else if (a.kind == ATHENA_OPT_ADAM) {
    float m = a.beta1 * s1[i] + (1.f - a.beta1) * gr;
    float v = a.beta2 * s2[i] + (1.f - a.beta2) * gr * gr;
    s1[i] = m;
    s2[i] = v;
    float mh = m / a.bc1, vh = v / a.bc2;
    if (a.reg == ATHENA_REG_L2) {
      // athena_optimiser.f90:1064-1078: AdamW (decoupled decay) or L2 inside the quotient
      float pv = p[i];
      if (a.l2_decoupled) {
        pv = pv - a.lr * a.l2 * pv;
        p[i] = pv - a.lr * (mh / (sqrtf(vh) + a.eps));
      } else {
        p[i] = pv - a.lr * ((mh + a.l2 * pv) / (sqrtf(vh) + a.eps));
      }
    } else {
      p[i] = p[i] - a.lr * (mh / (sqrtf(vh) + a.eps));
    }
  } else if (a.kind == ATHENA_OPT_RMSPROP) {
    // minimise_rmsprop (athena_optimiser.f90:795-803): the moving average lives in s1
    const float avg = a.beta1 * s1[i] + (1.f - a.beta1) * (gr * gr);
    s1[i] = avg;
    p[i] = p[i] - a.lr * gr / sqrtf(avg + a.eps);
  } else {
    // minimise_adagrad (athena_optimiser.f90:919-924): the sum of squares lives in s1
    const float ss = s1[i] + gr * gr;
    s1[i] = ss;
    p[i] = p[i] - a.lr * gr / sqrtf(ss + a.eps);
  }
}

__global__ void __launch_bounds__(RED_THREADS)
k_step(float* __restrict__ p, float* __restrict__ g, float* __restrict__ s1,
       float* __restrict__ s2, long long n, const float* __restrict__ partial, int nb,
       StepArgs a) {
  __shared__ float scale_sh;
  if (a.norm_on) {
    // every block re-derives the same total in the same order
    float local = 0.f;
    for (int i = threadIdx.x; i < nb; i += RED_THREADS) local += partial[i];
    float tot = block_sum(local);
    if (threadIdx.x == 0) scale_sh = fminf(1.f, a.clip_norm / sqrtf(tot));
  } else if (threadIdx.x == 0) {
    scale_sh = 1.f;
  }
  __syncthreads();
  const float scale = scale_sh;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float gr = g[i];
    if (scale < 1.f) gr = gr * scale;
    step_one(p, s1, s2, i, gr, a);
    g[i] = 0.f;  // reset_gradients, athena_network_sub.f90:2927
  }
}

// ---- end-of-backward finalisation -----------------------------------------------
// One launch that (1) folds the per-CTA partial weight gradients of the fused dW kernels
// into the flat gradient vector, (2) folds the per-CTA loss sums into the loss slot and,
// when no gradient exchange or clipping stands between backward and update (single rank,
// clip off), (3) applies the optimiser step and zeroes the gradients.  Fixed summation
// order everywhere: partials are split over FIN_GROUPS lanes in an interleaved, fixed
// pattern and the group totals are added in group order.
constexpr int FIN_ELEMS = 32, FIN_GROUPS = 32, FIN_THREADS = FIN_ELEMS * FIN_GROUPS;
constexpr int FIN_MAX_JOBS = 8;
constexpr int64_t FIN_TAIL_MAX = 16384;  // parameters one block steps through at the end
struct FinJob {
  const float* part;  // [nparts][count]
  int nparts, count;
  long long dst;      // offset of the block in the flat gradient vector
};
struct FinArgs {
  FinJob job[FIN_MAX_JOBS];
  int njobs;
  const float* loss_part;
  int loss_nparts;
  float* loss_acc;
  int do_step;
  float* xout;  // optional copy of the finished gradients + loss slot (peer-memory exchange)
  P2PSignal sig;  // sig.world > 0: the last block to finish tells every peer "slot complete"
  // exchange fused into this launch (the whole grid is co-resident): after the signal every
  // block waits for the peers' flags, adds the staged vectors of all ranks in rank order and
  // (xstep) applies the optimiser step -- what k_p2p_sum_step does in a launch of its own
  // norm clipping without a pass of its own: the gradients are clamped here and every block
  // leaves the sum of squares of its slice for k_step (sumsq_part[gridDim.x])
  float* sumsq_part;
  int clamp;
  float cmin, cmax;
  // small parameter vectors with norm clipping: the LAST block to finish adds up the partial
  // sums of squares (in block order) and applies the whole clipped step -- no second launch
  int tail_step;
  unsigned int* tail_counter;
  int xfused, xstep;
  const float* xin[P2P_MAX_WORLD];
  long long timeout_ns;
  uint32_t* err;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// relaxed system-scope store: after ONE __threadfence_system() a run of these publishes a flag
// to every peer (fence + relaxed store = release), instead of paying a system-scope release per
// peer (measured: the exchange cost grew by ~1.8 us per peer)
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// (polling with relaxed loads and one system-scope fence behind the loop was measured 2 x slower:
// 2 048 waiting threads each pay the fence)
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(FIN_THREADS)
k_finalize(float* __restrict__ p, float* __restrict__ g, float* __restrict__ s1,
           float* __restrict__ s2, long long n, FinArgs f, StepArgs a) {
  __shared__ float red[FIN_GROUPS][FIN_ELEMS];
  pdl_wait();
  // (no early release of the next kernel: measured, it gains nothing -- 0.2041 vs 0.2031 ms per
  // cfg2 step -- and this kernel writes the parameters)
  const int e = threadIdx.x % FIN_ELEMS, grp = threadIdx.x / FIN_ELEMS;
  const long long i = (long long)blockIdx.x * FIN_ELEMS + e;
  float sum = 0.f;
  if (i < n) {
    for (int jb = 0; jb < f.njobs; ++jb) {
      const FinJob& J = f.job[jb];
      const long long off = i - J.dst;
      if (off >= 0 && off < J.count) {
        // four independent running sums (fixed pattern) keep four loads in flight
        const float* src = J.part + off;
        const size_t st = (size_t)J.count;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int c = grp;
        for (; c + 3 * FIN_GROUPS < J.nparts; c += 4 * FIN_GROUPS) {
          const float v0 = __ldg(src + c * st), v1 = __ldg(src + (c + FIN_GROUPS) * st);
          const float v2 = __ldg(src + (c + 2 * FIN_GROUPS) * st);
          const float v3 = __ldg(src + (c + 3 * FIN_GROUPS) * st);
          acc[0] += v0;
          acc[1] += v1;
          acc[2] += v2;
          acc[3] += v3;
        }
        for (int k = 0; c < J.nparts; c += FIN_GROUPS, ++k) acc[k] += __ldg(src + c * st);
        sum += (acc[0] + acc[1]) + (acc[2] + acc[3]);
      }
    }
  }
  red[grp][e] = sum;
  __syncthreads();
  float sq = 0.f;
  if (grp == 0 && i < n) {
    float gr = g[i];
#pragma unroll
    for (int k = 0; k < FIN_GROUPS; ++k) gr += red[k][e];
    if (f.do_step) {
      step_one(p, s1, s2, i, gr, a);
      g[i] = 0.f;
    } else {
      if (f.sumsq_part != nullptr && f.clamp) gr = fmaxf(f.cmin, fminf(f.cmax, gr));
      g[i] = gr;
      if (f.xout) f.xout[i] = gr;
    }
    sq = gr * gr;
  }
  if (f.sumsq_part != nullptr && grp == 0) {
    // grp == 0 is exactly warp 0 (FIN_ELEMS = 32): fixed-order butterfly over the slice (lanes
    // past the end of the vector contribute zeros)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (e == 0) f.sumsq_part[blockIdx.x] = sq;
  }
  if (f.tail_step) {
    __shared__ int is_last;
    __syncthreads();  // this block's gradients and its partial are written
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int prev = atomicAdd(f.tail_counter, 1u);
      is_last = prev + 1 == gridDim.x;
      if (is_last) *f.tail_counter = 0;
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      float local = 0.f;
      for (int k = threadIdx.x; k < (int)gridDim.x; k += FIN_THREADS) local += __ldcg(f.sumsq_part + k);
      const float tot = block_sum<FIN_THREADS>(local);
      __shared__ float scale_sh;
      if (threadIdx.x == 0) scale_sh = fminf(1.f, a.clip_norm / sqrtf(tot));
      __syncthreads();
      const float scale = scale_sh;
      for (long long k = threadIdx.x; k < n; k += FIN_THREADS) {
        float gr = __ldcg(g + k);
        if (scale < 1.f) gr = gr * scale;
        step_one(p, s1, s2, k, gr, a);
        g[k] = 0.f;  // reset_gradients, athena_network_sub.f90:2927
      }
    }
  }
  if (blockIdx.x == 0 && (f.loss_part != nullptr || f.xout != nullptr)) {
    __syncthreads();
    float local = 0.f;
    if (f.loss_part != nullptr)
      for (int k = threadIdx.x; k < f.loss_nparts; k += FIN_THREADS) local += f.loss_part[k];
    float s = block_sum<FIN_THREADS>(local);
    if (threadIdx.x == 0) {
      const float loss = f.loss_acc[0] + 0.5f * s;
      f.loss_acc[0] = loss;
      if (f.xout) f.xout[n] = loss;
    }
  }
  if (f.sig.world > 0) {
    // The staged vector is complete when every block has written its slice: the last block
    // to arrive publishes it to all peers with system-scope release stores over NVLink, so the
    // exchange kernel only waits (its progress never depends on one of its own blocks).
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int prev = atomicAdd(f.sig.done, 1u);
      if (prev + 1 == gridDim.x) {
        *f.sig.done = 0;
        __threadfence_system();
        for (int r = 0; r < f.sig.world; ++r)
          st_relaxed_sys(f.sig.flags[r] + f.sig.slot * P2P_MAX_WORLD + f.sig.rank, f.sig.epoch);
      }
    }
  }
  if (f.xfused) {
    if (threadIdx.x < f.sig.world) {
      const uint32_t* mine =
          f.sig.flags[f.sig.rank] + f.sig.slot * P2P_MAX_WORLD + threadIdx.x;
      long long t0 = 0;
      while (ld_acquire_sys(mine) != f.sig.epoch) {
        __nanosleep(64);
        long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        if (now - t0 > f.timeout_ns || *reinterpret_cast<volatile uint32_t*>(f.err) != 0u) {
          atomicExch(f.err, 1u);
          break;
        }
      }
    }
    __syncthreads();
    if (*reinterpret_cast<volatile uint32_t*>(f.err) != 0u) return;
    const bool loss_slot = blockIdx.x == 0 && threadIdx.x == FIN_ELEMS;  // one spare thread
    if ((grp == 0 && i < n) || loss_slot) {
      const long long k = loss_slot ? n : i;
      float v[P2P_MAX_WORLD];
#pragma unroll
      for (int r = 0; r < P2P_MAX_WORLD; ++r)
        v[r] = r < f.sig.world ? ld_relaxed_sys(f.xin[r] + k) : 0.f;
      float tot = 0.f;
#pragma unroll
      for (int r = 0; r < P2P_MAX_WORLD; ++r)
        if (r < f.sig.world) tot += v[r];
      if (loss_slot || !f.xstep) {
        g[k] = tot;
      } else {
        step_one(p, s1, s2, k, tot, a);
        g[k] = 0.f;
      }
    }
  }
}

// ---- peer-memory gradient exchange fused with the step -----------------------------
// Every rank staged its local gradient vector (+ loss slot) in its own exchange buffer
// (k_finalize).  This kernel (1) tells every peer "my slot is complete" with a system-scope
// flag store over NVLink, (2) waits until all peers have said the same, (3) reads the
// world_size staged vectors straight from peer memory, adds them in rank order -- every
// rank forms bitwise the same sum -- and (4) applies the optimiser step (or leaves the sums
// in the gradient buffer).  The all-reduce never exists as a separate pass: the few tens of
// kB cross NVLink as plain loads issued by the threads that consume them.
struct P2PArgs {
  const float* x[P2P_MAX_WORLD];
  uint32_t* flags[P2P_MAX_WORLD];
  int world, rank;
  uint32_t epoch;
  int slot;
  int do_step;
  int signal;            // 1: this kernel also publishes the local slot (no k_finalize before it)
  long long timeout_ns;  // bounded wait: a missing peer sets the sticky error word instead of
  uint32_t* err;         // hanging the GPU; the host then reports ATHENA_ERR_COMM
};


__global__ void __launch_bounds__(RED_THREADS)
k_p2p_sum_step(float* __restrict__ p, float* __restrict__ g, float* __restrict__ s1,
               float* __restrict__ s2, long long n, P2PArgs x, StepArgs a) {
  if (threadIdx.x < x.world) {
    const int r = threadIdx.x;
    if (x.signal && blockIdx.x == 0) {
      __threadfence_system();  // the staged vector (written by the previous kernel) first
      st_relaxed_sys(x.flags[r] + x.slot * P2P_MAX_WORLD + x.rank, x.epoch);
    }
    const uint32_t* mine = x.flags[x.rank] + x.slot * P2P_MAX_WORLD + r;
    long long t0 = 0;
    while (ld_acquire_sys(mine) != x.epoch) {
      __nanosleep(64);
      long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > x.timeout_ns || *reinterpret_cast<volatile uint32_t*>(x.err) != 0u) {
        atomicExch(x.err, 1u);  // rank r never arrived: give up, the step is not applied
        break;
      }
    }
  }
  __syncthreads();
  if (*reinterpret_cast<volatile uint32_t*>(x.err) != 0u) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  // issue all peer loads first (independent NVLink round trips overlap), then add in rank order
  float v[P2P_MAX_WORLD];
#pragma unroll
  for (int r = 0; r < P2P_MAX_WORLD; ++r) v[r] = r < x.world ? ld_relaxed_sys(x.x[r] + i) : 0.f;
  float sum = 0.f;
#pragma unroll
  for (int r = 0; r < P2P_MAX_WORLD; ++r)
    if (r < x.world) sum += v[r];
  if (i == n || !x.do_step) {
    g[i] = sum;  // index n: the global batch loss
  } else {
    step_one(p, s1, s2, i, sum, a);
    g[i] = 0.f;
  }
}

static float powi(float b, int64_t e) {  // real ** integer
  float r = 1.f, x = b;
  while (e > 0) {
    if (e & 1) r *= x;
    x *= x;
    e >>= 1;
  }
  return r;
}

static int step_prepare(int64_t n, OptimState& st, StepArgs* a) {
  cudaStream_t s = ctx().stream;
  if (!st.s1.p) {
    ATH_TRY(st.s1.reserve(sizeof(float) * (size_t)n));
    ATH_TRY(st.s2.reserve(sizeof(float) * (size_t)n));
    ATH_CUDA(cudaMemsetAsync(st.s1.p, 0, sizeof(float) * (size_t)n, s));
    ATH_CUDA(cudaMemsetAsync(st.s2.p, 0, sizeof(float) * (size_t)n, s));
  }
  ATH_TRY(st.scratch.reserve(sizeof(float) * 1024));
  // incremented BEFORE the step (athena_network_sub.f90:2834-2841) -- unless the host keeps
  // the counter (lr_decay%iterate_per_epoch advances it once per epoch, and Adam's bias
  // correction reads the same counter)
  if (!st.iter_external) st.iter += 1;
  const athena_optimiser_desc& d = st.d;
  a->kind = d.kind;
  a->lr = st.lr;
  a->beta1 = d.beta1;
  a->beta2 = d.beta2;
  a->eps = d.epsilon;
  a->momentum = d.momentum;
  a->nesterov = d.nesterov;
  a->norm_on = d.clip_norm_on;
  a->clip_norm = d.clip_norm;
  a->clamp = d.clip_min_max;
  a->cmin = d.clip_min;
  a->cmax = d.clip_max;
  a->reg = d.regulariser;
  a->l1 = d.l1;
  a->l2 = d.l2;
  a->l2_decoupled = d.l2_decoupled;
  a->bc1 = 1.f - powi(d.beta1, st.iter);
  a->bc2 = 1.f - powi(d.beta2, st.iter);
  return ATHENA_OK;
}

int launch_update(float* params, float* grads, int64_t n, OptimState& st) {
  if (n == 0) return ATHENA_OK;
  cudaStream_t s = ctx().stream;
  StepArgs a;
  ATH_TRY(step_prepare(n, st, &a));
  int nb = red_blocks(n);
  const athena_optimiser_desc& d = st.d;
  int npart = nb;
  if (st.presum_nb > 0) {
    npart = st.presum_nb;  // clamped and summed by launch_finalize
    st.presum_nb = 0;
  } else if (d.clip_norm_on) {
    // (min / max alone is applied inside the step)
    k_clip_sumsq<<<nb, RED_THREADS, 0, s>>>(grads, n, d.clip_min_max, d.clip_min, d.clip_max,
                                            st.scratch.as<float>());
    ATH_LAUNCHED_T("clip_sumsq");
  }
  k_step<<<nb, RED_THREADS, 0, s>>>(params, grads, st.s1.as<float>(), st.s2.as<float>(), n,
                                    st.scratch.as<float>(), npart, a);
  ATH_LAUNCHED_T("optimiser_step");
  return ATHENA_OK;
}

bool finalize_can_step(const OptimState& st) {
  // min / max clipping is elementwise and rides along (step_one); norm clipping needs the
  // global sum of squares first
  return !st.d.clip_norm_on && comm_world_size() == 1;
}

// Folds the deferred partial reductions (and the loss partials) into the flat gradient
// buffer; with `st` != nullptr also performs the optimiser step (caller checked
// finalize_can_step).  More than FIN_MAX_JOBS reductions are flushed in several launches,
// the step riding on the last one.
static long long p2p_timeout_ns() {
  static long long t = -1;
  if (t < 0) {
    const char* e = getenv("ATHENA_CUDA_P2P_TIMEOUT_MS");
    t = (e && atoll(e) > 0 ? atoll(e) : 20000ll) * 1000000ll;
  }
  return t;
}

void p2p_next_signal(P2PSignal* sig) {
  P2PState& P = p2p();
  *sig = P2PSignal{};
  sig->world = P.world;
  sig->rank = P.rank;
  sig->epoch = P.epoch + 1;
  sig->slot = (int)((P.epoch + 1) & 1u);
  sig->done = reinterpret_cast<unsigned int*>(P.flags[P.rank] + P2P_DONE);
  for (int r = 0; r < P.world; ++r) sig->flags[r] = P.flags[r];
}

int p2p_check() {
  P2PState& P = p2p();
  if (!P.ready) return ATHENA_OK;
  uint32_t err = 0;
  ATH_CUDA(cudaMemcpyAsync(&err, P.flags[P.rank] + P2P_ERR, sizeof(err), cudaMemcpyDeviceToHost,
                           ctx().stream));
  ATH_CUDA(cudaStreamSynchronize(ctx().stream));
  if (err != 0) P.failed = true;
  ATH_REQUIRE(!P.failed, ATHENA_ERR_COMM,
              "p2p exchange: timed out waiting for a peer's gradients (a rank is missing or "
              "out of step); parameters were not updated");
  return ATHENA_OK;
}

int launch_p2p_sum_step(float* params, float* grads, int64_t n, OptimState* st, bool signalled) {
  P2PState& P = p2p();
  ATH_REQUIRE(P.ready && (size_t)(n + 1) <= P.cap, ATHENA_ERR_STATE,
              "p2p exchange: not initialised or gradient vector too long");
  StepArgs a{};
  if (st) ATH_TRY(step_prepare(n, *st, &a));
  ATH_REQUIRE(!P.failed, ATHENA_ERR_COMM,
              "p2p exchange: an earlier exchange timed out waiting for a peer");
  P2PArgs x{};
  P.epoch += 1;
  x.world = P.world;
  x.rank = P.rank;
  x.epoch = P.epoch;
  x.slot = (int)(P.epoch & 1u);
  x.do_step = st ? 1 : 0;
  x.signal = signalled ? 0 : 1;
  x.timeout_ns = p2p_timeout_ns();
  x.err = P.flags[P.rank] + P2P_ERR;
  for (int r = 0; r < P.world; ++r) {
    x.x[r] = P.xbuf[r] + (size_t)x.slot * P.cap;
    x.flags[r] = P.flags[r];
  }
  k_p2p_sum_step<<<(unsigned)cdiv(n + 1, RED_THREADS), RED_THREADS, 0, ctx().stream>>>(
      params, grads, st ? st->s1.as<float>() : nullptr, st ? st->s2.as<float>() : nullptr, n, x,
      a);
  ATH_LAUNCHED_T(st ? "p2p_sum_step" : "p2p_sum");
  return ATHENA_OK;
}

// The exchange can ride on the finalize launch when all its blocks are resident at once (every
// block waits for the peers after the LAST local block has published the slot).
bool finalize_can_exchange(int64_t n) {
  static int cap = -1;
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("ATHENA_CUDA_NO_FUSED_EXCHANGE");
    off = (e && atoi(e) != 0) ? 1 : 0;
  }
  if (cap < 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_finalize, FIN_THREADS, 0) !=
        cudaSuccess)
      per_sm = 0;
    cap = per_sm * ctx().sm_count;
  }
  return !off && cdiv(n, FIN_ELEMS) <= cap;
}

int launch_finalize(const DeferList& dl, const float* loss_part, int loss_nparts, float* loss_acc,
                    float* params, float* grads, int64_t n, OptimState* st, float* xout,
                    const P2PSignal* sig, int exchange, OptimState* presum, bool* presum_stepped) {
  if (presum_stepped) *presum_stepped = false;
  if (n == 0) return ATHENA_OK;
  // norm clipping: this launch clamps and leaves the partial sums of squares for the step
  const int fin_grid = (int)cdiv(n, FIN_ELEMS);
  if (presum != nullptr && (st != nullptr || xout != nullptr || !presum->d.clip_norm_on ||
                            fin_grid > 2048 || dl.jobs.size() > (size_t)FIN_MAX_JOBS))
    presum = nullptr;
  if (presum) ATH_TRY(presum->scratch.reserve(sizeof(float) * (size_t)std::max(fin_grid, 1024)));
  cudaStream_t s = ctx().stream;
  StepArgs a{};
  if (st) ATH_TRY(step_prepare(n, *st, &a));
  // a short parameter vector: the clipped step rides on this launch (its last block does it)
  const bool tail = presum != nullptr && presum_stepped != nullptr && n <= FIN_TAIL_MAX;
  if (tail) {
    if (!presum->tail.p) {
      ATH_TRY(presum->tail.reserve(sizeof(unsigned int)));
      ATH_CUDA(cudaMemsetAsync(presum->tail.p, 0, sizeof(unsigned int), s));
    }
    ATH_TRY(step_prepare(n, *presum, &a));
    ATH_TRY(presum->scratch.reserve(sizeof(float) * (size_t)std::max(fin_grid, 1024)));
  }
  P2PState& PS = p2p();
  if (exchange) {
    ATH_REQUIRE(sig && xout && PS.ready && dl.jobs.size() <= (size_t)FIN_MAX_JOBS, ATHENA_ERR_STATE,
                "finalize: fused exchange without a prepared slot");
    ATH_REQUIRE(!PS.failed, ATHENA_ERR_COMM,
                "p2p exchange: an earlier exchange timed out waiting for a peer");
  }
  OptimState* sst = st ? st : (tail ? presum : nullptr);  // whose moments the launch updates
  size_t done = 0;
  do {
    FinArgs f{};
    f.njobs = (int)std::min<size_t>(FIN_MAX_JOBS, dl.jobs.size() - done);
    for (int k = 0; k < f.njobs; ++k) {
      const DeferJob& j = dl.jobs[done + k];
      f.job[k].part = j.part;
      f.job[k].nparts = j.nparts;
      f.job[k].count = j.count;
      f.job[k].dst = j.dst - grads;
    }
    done += f.njobs;
    const bool last = done == dl.jobs.size();
    f.loss_part = last ? loss_part : nullptr;
    f.loss_nparts = loss_nparts;
    f.loss_acc = loss_acc;
    f.do_step = (last && st && !exchange) ? 1 : 0;
    if (last && presum) {
      f.sumsq_part = presum->scratch.as<float>();
      f.clamp = presum->d.clip_min_max;
      f.cmin = presum->d.clip_min;
      f.cmax = presum->d.clip_max;
      presum->presum_nb = tail ? 0 : fin_grid;
      if (tail) {
        f.tail_step = 1;
        f.tail_counter = presum->tail.as<unsigned int>();
        *presum_stepped = true;
      }
    }
    f.xout = last ? xout : nullptr;
    if (last && xout && sig) f.sig = *sig;
    if (last && exchange) {
      f.xfused = 1;
      f.xstep = st ? 1 : 0;
      for (int r = 0; r < PS.world; ++r) f.xin[r] = PS.xbuf[r] + (size_t)sig->slot * PS.cap;
      f.timeout_ns = p2p_timeout_ns();
      f.err = PS.flags[PS.rank] + P2P_ERR;
      PS.epoch += 1;  // this launch IS the exchange
    }
    ATH_CUDA(launch_pdl(k_finalize, dim3((unsigned)cdiv(n, FIN_ELEMS)), dim3(FIN_THREADS), 0, s,
                        params, grads, sst ? sst->s1.as<float>() : nullptr,
                        sst ? sst->s2.as<float>() : nullptr, (long long)n, f, a));
    ATH_LAUNCHED_T(f.xfused ? (f.xstep ? "finalize_exchange_step" : "finalize_exchange")
                            : (f.do_step || f.tail_step) ? "finalize_step" : "finalize");
  } while (done < dl.jobs.size());
  return ATHENA_OK;
}

}  // namespace athena
