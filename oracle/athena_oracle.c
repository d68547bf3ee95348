/*
 * athena_oracle.c -- CPU restatement of athena's message-passing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (athena_b200/,
 * libathena_cuda) may include, link or call this file.  It is used by tests/,
 * by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs, as the checker and as the reported CPU baseline.
 *
 * Parity status: the reference (nedtaylor/athena v2.1.1, Fortran) cannot be
 * compiled in this image (no Fortran compiler, un-vendored deps diffstruc /
 * graphstruc / coreutils).  This restatement follows the in-tree Fortran
 * sources line by line (citations on every function, paths relative to
 * /root/reference) and is pinned against
 *   - the reference's own known-answer tests (identity-graph Kipf test, MSE
 *     values, clipper values, shared-operand gradient convention), and
 *   - golden vectors generated from the reference's own PyTorch restatement of
 *     the Duvenaud network (example/msgpass_chemical/pytorch_network.py),
 *     tests/golden/make_golden.py.
 * Arithmetic that lives in the out-of-tree diffstruc dependency (matmul
 * forward/backward, activation derivatives, sum(new_dim_index)) is restated
 * from its published semantics and the call sites; see DESIGN.md "parity
 * unpinned" list.
 *
 * Memory conventions (identical to the Fortran side so buffers can be shared):
 *   val(F,V) column-major  ==  C x[v*F + f]
 *   adj_ia(V+1)  1-based row pointers
 *   adj_ja(2,Z)  interleaved {neighbour, edge-id}, both 1-based;
 *                edge-id <= 0 means "no edge feature" (self-loop marker seen in
 *                test/test_diffstruc_extd_kipf.f90:30-31)
 *
 * Build: see oracle/Makefile.  -DORACLE_F64 gives a float64 shadow used to
 * separate summation-order noise from real bugs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORACLE_F64
typedef double real;
#define R_EXP exp
#define R_SQRT sqrt
#define R_COPYSIGN copysign
#define R_POW pow
#define R_TANH tanh
#else
typedef float real;
#define R_EXP expf
#define R_SQRT sqrtf
#define R_COPYSIGN copysignf
#define R_POW powf
#define R_TANH tanhf
#endif

#define API __attribute__((visibility("default")))

/* activation ids -- shared with include/athena_cuda.h */
enum {
  ACT_NONE = 0,
  ACT_LINEAR = 1,
  ACT_RELU = 2,
  ACT_LEAKY_RELU = 3,
  ACT_SIGMOID = 4,
  ACT_TANH = 5,
  ACT_SOFTMAX = 6,
  ACT_SWISH = 7
};

API int oracle_real_bytes(void) { return (int)sizeof(real); }

/* ------------------------------------------------------------------------- */
/* Kipf propagate                                                            */
/* ------------------------------------------------------------------------- */

/* kipf_propagate forward.
 * src/athena/athena_diffstruc_extd_sub_kipf.f90:29-46
 *   coeff = ( deg(v) * deg(ja(1,w)) ) ** (-0.5)   -- integer product first
 *   c(:,v) += coeff * x(:, ja(1,w))               -- ascending w            */
API void oracle_kipf_propagate(int F, int V, const real *x, const int *ia,
                               const int *ja, real *c) {
  for (int v = 0; v < V; ++v) {
    real *cv = c + (size_t)v * F;
    for (int f = 0; f < F; ++f) cv[f] = 0;
    for (int w = ia[v] - 1; w < ia[v + 1] - 1; ++w) {
      int u = ja[2 * w] - 1;
      int prod = (ia[v + 1] - ia[v]) * (ia[u + 1] - ia[u]);
      real coeff = R_POW((real)prod, (real)-0.5);
      const real *xu = x + (size_t)u * F;
      for (int f = 0; f < F; ++f) cv[f] = cv[f] + coeff * xu[f];
    }
  }
}

/* get_partial_kipf_propagate_left_val (the live backward; NO coefficient).
 * src/athena/athena_diffstruc_extd_sub_kipf.f90:85-111                      */
API void oracle_kipf_propagate_bwd(int F, int V, const real *g, const int *ia,
                                   const int *ja, real *dx) {
  memset(dx, 0, sizeof(real) * (size_t)F * V);
  for (int v = 0; v < V; ++v) {
    const real *gv = g + (size_t)v * F;
    for (int w = ia[v] - 1; w < ia[v + 1] - 1; ++w) {
      real *du = dx + (size_t)(ja[2 * w] - 1) * F;
      for (int f = 0; f < F; ++f) du[f] = du[f] + gv[f];
    }
  }
}

/* ------------------------------------------------------------------------- */
/* Duvenaud propagate / update                                               */
/* ------------------------------------------------------------------------- */

/* duvenaud_propagate forward: A(:,v) = sum_w [ x(:,ja(1,w)) ; e(:,ja(2,w)) ].
 * src/athena/athena_diffstruc_extd_sub_duvenaud.f90:34-42
 * Edge id <= 0 is out of bounds in the reference (UB); the ABI defines it as a
 * zero edge-feature contribution.                      */
API void oracle_duvenaud_propagate(int F, int Fe, int V, const real *x,
                                   const real *e, const int *ia, const int *ja,
                                   real *c) {
  int K = F + Fe;
  for (int v = 0; v < V; ++v) {
    real *cv = c + (size_t)v * K;
    for (int k = 0; k < K; ++k) cv[k] = 0;
    for (int w = ia[v] - 1; w < ia[v + 1] - 1; ++w) {
      const real *xu = x + (size_t)(ja[2 * w] - 1) * F;
      for (int f = 0; f < F; ++f) cv[f] = cv[f] + xu[f];
      int eid = ja[2 * w + 1];
      if (eid >= 1) {
        const real *ee = e + (size_t)(eid - 1) * Fe;
        for (int f = 0; f < Fe; ++f) cv[F + f] = cv[F + f] + ee[f];
      }
    }
  }
}

/* get_partial_duvenaud_propagate_left_val: dX(:,ja(1,w)) += g(1:F, v).
 * src/athena/athena_diffstruc_extd_sub_duvenaud.f90:115-142                 */
API void oracle_duvenaud_propagate_bwd_left(int F, int Fe, int V, const real *g,
                                            const int *ia, const int *ja,
                                            real *dx) {
  int K = F + Fe;
  memset(dx, 0, sizeof(real) * (size_t)F * V);
  for (int v = 0; v < V; ++v) {
    const real *gv = g + (size_t)v * K;
    for (int w = ia[v] - 1; w < ia[v + 1] - 1; ++w) {
      real *du = dx + (size_t)(ja[2 * w] - 1) * F;
      for (int f = 0; f < F; ++f) du[f] = du[f] + gv[f];
    }
  }
}

/* get_partial_duvenaud_propagate_right_val: dE(:,ja(2,w)) += g(F+1:, v).
 * src/athena/athena_diffstruc_extd_sub_duvenaud.f90:144-171                 */
API void oracle_duvenaud_propagate_bwd_right(int F, int Fe, int V, int E,
                                             const real *g, const int *ia,
                                             const int *ja, real *de) {
  int K = F + Fe;
  memset(de, 0, sizeof(real) * (size_t)Fe * E);
  for (int v = 0; v < V; ++v) {
    const real *gv = g + (size_t)v * K;
    for (int w = ia[v] - 1; w < ia[v + 1] - 1; ++w) {
      int eid = ja[2 * w + 1];
      if (eid < 1) continue;
      real *d = de + (size_t)(eid - 1) * Fe;
      for (int f = 0; f < Fe; ++f) d[f] = d[f] + gv[F + f];
    }
  }
}

static inline int bucket_of(const int *ia, int v, int min_deg, int max_deg) {
  int deg = ia[v + 1] - ia[v];
  int c = deg < max_deg ? deg : max_deg;
  c = c > min_deg ? c : min_deg;
  return c - min_deg + 1; /* 1..D, athena_diffstruc_extd_sub_duvenaud.f90:206-207 */
}

/* duvenaud_update forward: Z(:,v) = W_d . ( A(:,v) / real(d) ).
 * src/athena/athena_diffstruc_extd_sub_duvenaud.f90:204-211
 * W_d is [Fo, K] column-major at flat offset (d-1)*Fo*K.                    */
API void oracle_duvenaud_update(int K, int Fo, int V, const real *a,
                                const real *w, const int *ia, int min_deg,
                                int max_deg, real *c) {
  int interval = Fo * K;
  real *tmp = (real *)malloc(sizeof(real) * (size_t)K);
  for (int v = 0; v < V; ++v) {
    int d = bucket_of(ia, v, min_deg, max_deg);
    const real *wd = w + (size_t)interval * (d - 1);
    const real *av = a + (size_t)v * K;
    for (int k = 0; k < K; ++k) tmp[k] = av[k] / (real)d;
    real *cv = c + (size_t)v * Fo;
    for (int o = 0; o < Fo; ++o) cv[o] = 0;
    for (int k = 0; k < K; ++k)
      for (int o = 0; o < Fo; ++o) cv[o] = cv[o] + wd[o + (size_t)Fo * k] * tmp[k];
  }
  free(tmp);
}

/* get_partial_duvenaud_update_val: dA(:,v) = matmul(g(:,v), W_d) / real(d).
 * src/athena/athena_diffstruc_extd_sub_duvenaud.f90:284-324                 */
API void oracle_duvenaud_update_bwd_val(int K, int Fo, int V, const real *g,
                                        const real *w, const int *ia,
                                        int min_deg, int max_deg, real *da) {
  int interval = Fo * K;
  for (int v = 0; v < V; ++v) {
    int d = bucket_of(ia, v, min_deg, max_deg);
    const real *wd = w + (size_t)interval * (d - 1);
    const real *gv = g + (size_t)v * Fo;
    real *dv = da + (size_t)v * K;
    for (int k = 0; k < K; ++k) {
      real s = 0;
      for (int o = 0; o < Fo; ++o) s = s + gv[o] * wd[o + (size_t)Fo * k];
      dv[k] = s / (real)d;
    }
  }
}

/* get_partial_duvenaud_update_weight_val:
 *   dW_d(i,j) += g(i,v) * A(j,v) / real(d)       (ascending v)
 * src/athena/athena_diffstruc_extd_sub_duvenaud.f90:326-368
 * dw is ACCUMULATED INTO (caller zeroes), mirroring the `+` accumulation of
 * params(t)%grad over samples.                                              */
API void oracle_duvenaud_update_bwd_weight(int K, int Fo, int V, const real *g,
                                           const real *a, const int *ia,
                                           int min_deg, int max_deg, real *dw) {
  int interval = Fo * K;
  for (int v = 0; v < V; ++v) {
    int d = bucket_of(ia, v, min_deg, max_deg);
    real *wd = dw + (size_t)interval * (d - 1);
    const real *gv = g + (size_t)v * Fo;
    const real *av = a + (size_t)v * K;
    for (int j = 0; j < K; ++j)
      for (int i = 0; i < Fo; ++i)
        wd[i + (size_t)Fo * j] = wd[i + (size_t)Fo * j] + gv[i] * av[j] / (real)d;
  }
}

/* ------------------------------------------------------------------------- */
/* dense pieces that live in diffstruc (out of tree): restated from call     */
/* sites.  parity unpinned -- standard maths, sequential k order.            */
/* ------------------------------------------------------------------------- */

/* Y = W . P ; W [M,K] column-major, P [K,N], Y [M,N].
 * call sites: athena_kipf_msgpass_layer.f90:951,
 *             athena_duvenaud_msgpass_layer.f90:842                         */
API void oracle_matmul(int M, int K, int N, const real *w, const real *p,
                       real *y) {
  for (int n = 0; n < N; ++n) {
    real *yn = y + (size_t)n * M;
    const real *pn = p + (size_t)n * K;
    for (int m = 0; m < M; ++m) yn[m] = 0;
    for (int k = 0; k < K; ++k)
      for (int m = 0; m < M; ++m) yn[m] = yn[m] + w[m + (size_t)M * k] * pn[k];
  }
}

/* dP = W^T . gY */
API void oracle_matmul_bwd_right(int M, int K, int N, const real *w,
                                 const real *gy, real *dp) {
  for (int n = 0; n < N; ++n) {
    const real *gn = gy + (size_t)n * M;
    real *dn = dp + (size_t)n * K;
    for (int k = 0; k < K; ++k) {
      real s = 0;
      for (int m = 0; m < M; ++m) s = s + w[m + (size_t)M * k] * gn[m];
      dn[k] = s;
    }
  }
}

/* dW += gY . P^T  (summed over columns n ascending; shared-operand gradients
 * are summed over positions -- convention pinned by
 * test/test_diffstruc_extd.f90:41-45).                                      */
API void oracle_matmul_bwd_left(int M, int K, int N, const real *gy,
                                const real *p, real *dw) {
  for (int n = 0; n < N; ++n) {
    const real *gn = gy + (size_t)n * M;
    const real *pn = p + (size_t)n * K;
    for (int k = 0; k < K; ++k)
      for (int m = 0; m < M; ++m)
        dw[m + (size_t)M * k] = dw[m + (size_t)M * k] + gn[m] * pn[k];
  }
}

/* softmax over each column (dim=2 in athena's naming: per vertex over
 * features).  src/athena/athena_diffstruc_extd_sub.f90:309-313              */
API void oracle_softmax_cols(int F, int V, const real *x, real *y) {
  for (int v = 0; v < V; ++v) {
    const real *xv = x + (size_t)v * F;
    real *yv = y + (size_t)v * F;
    real mx = xv[0];
    for (int f = 1; f < F; ++f) mx = xv[f] > mx ? xv[f] : mx;
    real s = 0;
    for (int f = 0; f < F; ++f) {
      yv[f] = R_EXP(xv[f] - mx);
      s = s + yv[f];
    }
    for (int f = 0; f < F; ++f) yv[f] = yv[f] / s;
  }
}

/* get_partial_softmax_val, dim=2 branch -> per-column:
 *   out = y*g ; out(:,s) -= y(:,s) * sum(out(:,s))
 * src/athena/athena_diffstruc_extd_sub.f90:355-379                          */
API void oracle_softmax_cols_bwd(int F, int V, const real *y, const real *g,
                                 real *dx) {
  for (int v = 0; v < V; ++v) {
    const real *yv = y + (size_t)v * F;
    const real *gv = g + (size_t)v * F;
    real *dv = dx + (size_t)v * F;
    real s = 0;
    for (int f = 0; f < F; ++f) {
      dv[f] = yv[f] * gv[f];
      s = s + dv[f];
    }
    for (int f = 0; f < F; ++f) dv[f] = dv[f] - yv[f] * s;
  }
}

/* activation apply.  src/athena/athena_activation_{none,linear,relu,
 * leaky_relu,sigmoid,tanh,softmax}.f90 (apply functions; scale = 1,
 * threshold = 0, alpha = 0.01 defaults: athena_activation_leaky_relu.f90:83-85) */
API void oracle_activation(int kind, int F, int V, const real *x, real *y) {
  size_t n = (size_t)F * V;
  switch (kind) {
  case ACT_NONE:
  case ACT_LINEAR:
    for (size_t i = 0; i < n; ++i) y[i] = x[i];
    break;
  case ACT_RELU:
    for (size_t i = 0; i < n; ++i) y[i] = x[i] > 0 ? x[i] : 0;
    break;
  case ACT_LEAKY_RELU:
    for (size_t i = 0; i < n; ++i) {
      real a = x[i] * (real)0.01;
      y[i] = a > x[i] ? a : x[i];
    }
    break;
  case ACT_SIGMOID:
    for (size_t i = 0; i < n; ++i) y[i] = (real)1 / ((real)1 + R_EXP(-x[i]));
    break;
  case ACT_TANH:
    for (size_t i = 0; i < n; ++i) y[i] = R_TANH(x[i]);
    break;
  case ACT_SOFTMAX:
    oracle_softmax_cols(F, V, x, y);
    break;
  case ACT_SWISH:
    /* swish_array, beta = 1 (athena_activation_swish.f90:31):
     * output = input * (1 / (1 + exp(-beta * input))), athena_diffstruc_extd_sub.f90:434 */
    for (size_t i = 0; i < n; ++i)
      y[i] = x[i] * ((real)1 / ((real)1 + R_EXP(-(real)1 * x[i])));
    break;
  }
}

/* activation backward expressed on the OUTPUT y (what the device path saves):
 *   relu: g*[y>0]; leaky: g*(y>0 ? 1 : alpha); sigmoid: g*y*(1-y);
 *   tanh: g*(1-y^2); softmax: per-column Jacobian.  Derivative forms live in
 *   diffstruc (out of tree) -- standard maths assumed. */
API void oracle_activation_bwd(int kind, int F, int V, const real *y,
                               const real *g, real *dx) {
  size_t n = (size_t)F * V;
  switch (kind) {
  case ACT_NONE:
  case ACT_LINEAR:
    for (size_t i = 0; i < n; ++i) dx[i] = g[i];
    break;
  case ACT_RELU:
    for (size_t i = 0; i < n; ++i) dx[i] = y[i] > 0 ? g[i] : 0;
    break;
  case ACT_LEAKY_RELU:
    for (size_t i = 0; i < n; ++i) dx[i] = y[i] > 0 ? g[i] : g[i] * (real)0.01;
    break;
  case ACT_SIGMOID:
    for (size_t i = 0; i < n; ++i) dx[i] = g[i] * (y[i] * ((real)1 - y[i]));
    break;
  case ACT_TANH:
    for (size_t i = 0; i < n; ++i) dx[i] = g[i] * ((real)1 - y[i] * y[i]);
    break;
  case ACT_SOFTMAX:
    oracle_softmax_cols_bwd(F, V, y, g, dx);
    break;
  }
}

/* The same with the PRE-activation x at hand: swish is the one activation whose derivative
 * is written on the input, get_partial_swish_val (athena_diffstruc_extd_sub.f90:472-486):
 *   exp_term = exp(beta x) ; out = g * exp_term * (beta x + exp_term + 1) / (exp_term + 1)^2  */
API void oracle_activation_bwd_x(int kind, int F, int V, const real *x, const real *y,
                                 const real *g, real *dx) {
  if (kind != ACT_SWISH) {
    oracle_activation_bwd(kind, F, V, y, g, dx);
    return;
  }
  size_t n = (size_t)F * V;
  for (size_t i = 0; i < n; ++i) {
    real e = R_EXP((real)1 * x[i]);
    dx[i] = g[i] * e * ((real)1 * x[i] + e + (real)1) / R_POW(e + (real)1, (real)2);
  }
}

/* ------------------------------------------------------------------------- */
/* loss / clip / optimisers                                                  */
/* ------------------------------------------------------------------------- */

/* one cell of compute_mse: mean( (p-e)^2 ) / 2 over all n elements.
 * src/athena/athena_loss.f90:414 ; mean-over-all pinned by test/test_loss.f90:59-67 */
API real oracle_mse_cell(size_t n, const real *p, const real *e) {
  real s = 0;
  for (size_t i = 0; i < n; ++i) {
    real d = p[i] - e[i];
    s = s + d * d;
  }
  return s / (real)n / (real)2;
}

/* gradient of one MSE cell wrt p with upstream seed 1: (p-e)/n.             */
API void oracle_mse_cell_bwd(size_t n, const real *p, const real *e, real *g,
                             real denom) {
  for (size_t i = 0; i < n; ++i) g[i] = (p[i] - e[i]) / denom;
}

/* clip_type%apply without bias.  src/athena/athena_clipper.f90:190-203
 * flags: bit0 = l_min_max, bit1 = l_norm.                                   */
API void oracle_clip(int n, real *g, int flags, real cmin, real cmax,
                     real cnorm) {
  if (flags & 1)
    for (int i = 0; i < n; ++i) {
      real t = g[i] < cmax ? g[i] : cmax;
      g[i] = t > cmin ? t : cmin;
    }
  if (flags & 2) {
    real s = 0;
    for (int i = 0; i < n; ++i) s = s + g[i] * g[i];
    /* bias_ = [0] -> sum(bias_)**2 = 0 */
    real scale = cnorm / R_SQRT(s + (real)0);
    if (scale > (real)1) scale = 1;
    if (scale < (real)1)
      for (int i = 0; i < n; ++i) g[i] = g[i] * scale;
  }
}

/* clip with an explicit bias vector (test/test_clipper.f90:95-99).          */
API void oracle_clip_bias(int n, real *g, int nb, real *b, int flags, real cmin,
                          real cmax, real cnorm) {
  if (flags & 1) {
    for (int i = 0; i < n; ++i) {
      real t = g[i] < cmax ? g[i] : cmax;
      g[i] = t > cmin ? t : cmin;
    }
    for (int i = 0; i < nb; ++i) {
      real t = b[i] < cmax ? b[i] : cmax;
      b[i] = t > cmin ? t : cmin;
    }
  }
  if (flags & 2) {
    real s = 0, sb = 0;
    for (int i = 0; i < n; ++i) s = s + g[i] * g[i];
    for (int i = 0; i < nb; ++i) sb = sb + b[i];
    real scale = cnorm / R_SQRT(s + sb * sb);
    if (scale > (real)1) scale = 1;
    if (scale < (real)1) {
      for (int i = 0; i < n; ++i) g[i] = g[i] * scale;
      for (int i = 0; i < nb; ++i) b[i] = b[i] * scale;
    }
  }
}

/* minimise_sgd.  src/athena/athena_optimiser.f90:649-672                    */
API void oracle_sgd(int n, real *p, real *g, real *vel, real lr, real momentum,
                    int nesterov) {
  for (int i = 0; i < n; ++i) g[i] = -lr * g[i];
  if (momentum > (real)1e-8) {
    for (int i = 0; i < n; ++i) vel[i] = momentum * vel[i] + g[i];
    if (nesterov)
      for (int i = 0; i < n; ++i) p[i] = p[i] + momentum * vel[i] + g[i];
    else
      for (int i = 0; i < n; ++i) p[i] = p[i] + vel[i];
  } else {
    for (int i = 0; i < n; ++i) {
      vel[i] = g[i];
      p[i] = p[i] + vel[i];
    }
  }
}

static real powi(real b, int e) { /* real ** integer, as gfortran expands it */
  real r = 1;
  real x = b;
  while (e > 0) {
    if (e & 1) r = r * x;
    x = x * x;
    e >>= 1;
  }
  return r;
}

/* minimise_adam (no regulariser).  src/athena/athena_optimiser.f90:1043-1088
 * iter is the already-incremented counter (athena_network_sub.f90:2834-2841) */
API void oracle_adam(int n, real *p, const real *g, real *m, real *v, real lr,
                     real beta1, real beta2, real eps, int iter) {
  real bc1 = (real)1 - powi(beta1, iter);
  real bc2 = (real)1 - powi(beta2, iter);
  for (int i = 0; i < n; ++i) {
    m[i] = beta1 * m[i] + ((real)1 - beta1) * g[i];
    v[i] = beta2 * v[i] + ((real)1 - beta2) * g[i] * g[i];
  }
  for (int i = 0; i < n; ++i) {
    real mh = m[i] / bc1;
    real vh = v[i] / bc2;
    p[i] = p[i] - lr * (mh / (R_SQRT(vh) + eps));
  }
}

/* ------------------------------------------------------------------------- */
/* integer structures: host CSR batch -> device-layout CSR / CSC / buckets.  */
/* This is the bit-exact contract for libathena_cuda's batch build.          */
/* Follows the per-sample copy in athena_msgpass_layer_sub.f90:144-174 and   */
/* the bucket rule athena_diffstruc_extd_sub_duvenaud.f90:206-207.           */
/* ------------------------------------------------------------------------- */

/* Returns 0, or -(s+1) if graph s has a neighbour index outside 1..nv[s]
 * (the check in athena_duvenaud_msgpass_layer.f90:632-639).
 * Outputs (0-based, global over the block-diagonal batch):
 *   row_ptr[V+1], col[Z], eid[Z] (-1 = none), deg[V], vgraph[V],
 *   csc_ptr[V+1], csc_src[Z], csc_ent[Z]  (stable: ascending entry index)   */
API int oracle_batch_build(int B, const int *nv, const int *ne, const int *ia_cat,
                           const int *ja_cat, int *row_ptr, int *col, int *eid,
                           int *deg, int *vgraph, int *csc_ptr, int *csc_src,
                           int *csc_ent) {
  int voff = 0, zoff = 0, eoff = 0;
  const int *ia = ia_cat;
  const int *ja = ja_cat;
  for (int s = 0; s < B; ++s) {
    int nz = ia[nv[s]] - 1;
    for (int i = 0; i < nv[s]; ++i) {
      row_ptr[voff + i] = zoff + ia[i] - 1;
      vgraph[voff + i] = s;
    }
    for (int w = 0; w < nz; ++w) {
      int nb = ja[2 * w];
      if (nb < 1 || nb > nv[s]) return -(s + 1);
      col[zoff + w] = voff + nb - 1;
      int e = ja[2 * w + 1];
      eid[zoff + w] = (e >= 1 && e <= ne[s]) ? eoff + e - 1 : -1;
    }
    ia += nv[s] + 1;
    ja += 2 * (size_t)nz;
    voff += nv[s];
    zoff += nz;
    eoff += ne[s];
  }
  int V = voff, Z = zoff;
  row_ptr[V] = Z;
  for (int v = 0; v < V; ++v) deg[v] = row_ptr[v + 1] - row_ptr[v];
  /* transpose, stable in (row, entry) order */
  for (int v = 0; v <= V; ++v) csc_ptr[v] = 0;
  for (int w = 0; w < Z; ++w) csc_ptr[col[w] + 1]++;
  for (int v = 0; v < V; ++v) csc_ptr[v + 1] += csc_ptr[v];
  int *cur = (int *)malloc(sizeof(int) * (size_t)(V > 0 ? V : 1));
  for (int v = 0; v < V; ++v) cur[v] = csc_ptr[v];
  for (int v = 0; v < V; ++v)
    for (int w = row_ptr[v]; w < row_ptr[v + 1]; ++w) {
      int p = cur[col[w]]++;
      csc_src[p] = v;
      csc_ent[p] = w;
    }
  free(cur);
  return 0;
}

/* degree buckets: bkt[v] = clamp(deg,min,max)-min (0-based), stable counting
 * sort -> perm[V] (vertices of bucket b are perm[bkt_ptr[b]..bkt_ptr[b+1]) in
 * ascending vertex order).                                                  */
API void oracle_bucketize(int V, const int *deg, int min_deg, int max_deg,
                          int *bkt, int *perm, int *bkt_ptr) {
  int D = max_deg - min_deg + 1;
  for (int b = 0; b <= D; ++b) bkt_ptr[b] = 0;
  for (int v = 0; v < V; ++v) {
    int c = deg[v] < max_deg ? deg[v] : max_deg;
    c = c > min_deg ? c : min_deg;
    bkt[v] = c - min_deg;
    bkt_ptr[bkt[v] + 1]++;
  }
  for (int b = 0; b < D; ++b) bkt_ptr[b + 1] += bkt_ptr[b];
  int *cur = (int *)malloc(sizeof(int) * (size_t)D);
  for (int b = 0; b < D; ++b) cur[b] = bkt_ptr[b];
  for (int v = 0; v < V; ++v) perm[cur[bkt[v]]++] = v;
  free(cur);
}

/* ------------------------------------------------------------------------- */
/* composite: a stack of msgpass layers trained with MSE, per-sample loops   */
/* exactly as the reference runs them (serial over graphs).                  */
/* ------------------------------------------------------------------------- */

typedef struct {
  int kind;      /* 0 = kipf, 1 = duvenaud, 2 = full (dense head, T = 1) */
  int T;         /* num_time_steps */
  int nvf[17];   /* num_vertex_features(0:T), T <= 16; full: {num_inputs, num_outputs} */
  int nef;       /* num_edge_features(0) (duvenaud) */
  int min_deg, max_deg, n_out;
  int act, ract; /* message activation, readout activation */
  int use_bias;  /* full_layer_type%use_bias */
  /* network%add(layer, input_list, operator='concatenate') (athena_network_sub.f90:764-830,
   * example/msgpass_euler/src/main.f90:192-255): n_in = 0: the previous layer (or the network
   * input for the first); else the vertex features of the listed sources are concatenated
   * along the feature axis in list order (concat_layers(..., dim = 1), athena_concat_layer.f90:
   * 448): src = -1 the network input, src = k >= 0 layer k of the stack.  Kipf layers only. */
  int n_in;
  int in[4];
} oracle_layer_t;

/* width of a graph-level ([n, batch]) layer output */
static int graph_level_width(const oracle_layer_t *L) {
  return L->kind == 1 ? L->n_out : L->nvf[1];
}

API int oracle_layer_num_params(const oracle_layer_t *L) {
  int n = 0;
  if (L->kind == 2) {
    /* athena_full_layer.f90:122-138, 371-396: W [num_outputs, num_inputs] then bias */
    return L->nvf[1] * L->nvf[0] + (L->use_bias ? L->nvf[1] : 0);
  }
  if (L->kind == 0) {
    /* athena_kipf_msgpass_layer.f90:347-351 : W_t [F_t, F_{t-1}] */
    for (int t = 1; t <= L->T; ++t) n += L->nvf[t] * L->nvf[t - 1];
  } else {
    /* athena_duvenaud_msgpass_layer.f90:547-557 */
    int D = L->max_deg - L->min_deg + 1;
    for (int t = 1; t <= L->T; ++t) n += L->nvf[t] * (L->nvf[t - 1] + L->nef) * D;
    for (int t = 1; t <= L->T; ++t) n += L->n_out * L->nvf[t];
  }
  return n;
}

typedef struct {
  real **P; /* per step: propagated / aggregated input  */
  real **H; /* per step: activated output               */
  real **Y; /* per step: pre-activation (swish differentiates on it) */
} saved_t;

static void saved_free(saved_t *s, int T) {
  for (int t = 0; t < T; ++t) {
    free(s->P[t]);
    free(s->H[t]);
    if (s->Y) free(s->Y[t]);
  }
  free(s->P);
  free(s->H);
  free(s->Y);
}

/* update_message_kipf for ONE sample.  athena_kipf_msgpass_layer.f90:940-957 */
static const real *kipf_forward_sample(const oracle_layer_t *L, const real *params,
                                       int V, const int *ia, const int *ja,
                                       const real *x, saved_t *sv) {
  sv->P = (real **)calloc(L->T, sizeof(real *));
  sv->H = (real **)calloc(L->T, sizeof(real *));
  sv->Y = (real **)calloc(L->T, sizeof(real *));
  const real *in = x;
  const real *w = params;
  for (int t = 1; t <= L->T; ++t) {
    int Fi = L->nvf[t - 1], Fo = L->nvf[t];
    real *P = (real *)malloc(sizeof(real) * (size_t)Fi * V + 8);
    real *Y = (real *)malloc(sizeof(real) * (size_t)Fo * V + 8);
    real *H = (real *)malloc(sizeof(real) * (size_t)Fo * V + 8);
    oracle_kipf_propagate(Fi, V, in, ia, ja, P);
    oracle_matmul(Fo, Fi, V, w, P, Y);
    oracle_activation(L->act, Fo, V, Y, H);
    sv->Y[t - 1] = Y;
    sv->P[t - 1] = P;
    sv->H[t - 1] = H;
    in = H;
    w += (size_t)Fo * Fi;
  }
  return in;
}

/* reverse sweep through one sample of a Kipf layer.  g_out is consumed.
 * dparams accumulated (+=).  dx (may be NULL) receives the input gradient.  */
static void kipf_backward_sample(const oracle_layer_t *L, const real *params,
                                 int V, const int *ia, const int *ja,
                                 const saved_t *sv, const real *g_out,
                                 real *dparams, real *dx) {
  size_t woff = 0;
  for (int t = 1; t <= L->T; ++t) woff += (size_t)L->nvf[t] * L->nvf[t - 1];
  real *g = (real *)malloc(sizeof(real) * (size_t)L->nvf[L->T] * V + 8);
  memcpy(g, g_out, sizeof(real) * (size_t)L->nvf[L->T] * V);
  for (int t = L->T; t >= 1; --t) {
    int Fi = L->nvf[t - 1], Fo = L->nvf[t];
    woff -= (size_t)Fo * Fi;
    real *gy = (real *)malloc(sizeof(real) * (size_t)Fo * V + 8);
    oracle_activation_bwd_x(L->act, Fo, V, sv->Y[t - 1], sv->H[t - 1], g, gy);
    oracle_matmul_bwd_left(Fo, Fi, V, gy, sv->P[t - 1], dparams + woff);
    free(g);
    g = NULL;
    if (t > 1 || dx) {
      real *dp = (real *)malloc(sizeof(real) * (size_t)Fi * V + 8);
      oracle_matmul_bwd_right(Fo, Fi, V, params + woff, gy, dp);
      real *gin = (t > 1) ? (real *)malloc(sizeof(real) * (size_t)Fi * V + 8) : dx;
      oracle_kipf_propagate_bwd(Fi, V, dp, ia, ja, gin);
      free(dp);
      if (t > 1) g = gin;
    }
    free(gy);
  }
  free(g);
}

/* full_layer_type%forward for ONE sample (column) of the [num_inputs, batch] input:
 * y = act( matmul(W, x) + b ).  athena_full_layer.f90:839-874.  Saves x and y.   */
static const real *full_forward_sample(const oracle_layer_t *L, const real *params,
                                       const real *x, saved_t *sv) {
  int Ni = L->nvf[0], No = L->nvf[1];
  sv->P = (real **)calloc(1, sizeof(real *));
  sv->H = (real **)calloc(1, sizeof(real *));
  sv->Y = (real **)calloc(1, sizeof(real *));
  real *xs = (real *)malloc(sizeof(real) * (size_t)Ni + 8);
  real *z = (real *)malloc(sizeof(real) * (size_t)No + 8);
  real *y = (real *)malloc(sizeof(real) * (size_t)No + 8);
  memcpy(xs, x, sizeof(real) * (size_t)Ni);
  oracle_matmul(No, Ni, 1, params, x, z);
  if (L->use_bias) {
    const real *b = params + (size_t)No * Ni;
    for (int o = 0; o < No; ++o) z[o] = z[o] + b[o];
  }
  oracle_activation(L->act, No, 1, z, y);
  sv->Y[0] = z;
  sv->P[0] = xs;
  sv->H[0] = y;
  return y;
}

/* reverse sweep of one sample through a full layer: dW += gz x^T, db += gz,
 * dx = W^T gz with gz = act'(y) . g  (diffstruc matmul / add partials; shared operands
 * summed over samples, test/test_diffstruc_extd.f90:41-45).                          */
static void full_backward_sample(const oracle_layer_t *L, const real *params,
                                 const saved_t *sv, const real *g_out, real *dparams,
                                 real *dx) {
  int Ni = L->nvf[0], No = L->nvf[1];
  real *gz = (real *)malloc(sizeof(real) * (size_t)No + 8);
  oracle_activation_bwd_x(L->act, No, 1, sv->Y[0], sv->H[0], g_out, gz);
  oracle_matmul_bwd_left(No, Ni, 1, gz, sv->P[0], dparams);
  if (L->use_bias) {
    real *db = dparams + (size_t)No * Ni;
    for (int o = 0; o < No; ++o) db[o] = db[o] + gz[o];
  }
  if (dx) oracle_matmul_bwd_right(No, Ni, 1, params, gz, dx);
  free(gz);
}

/* update_message_duvenaud for ONE sample (athena_duvenaud_msgpass_layer.f90:
 * 792-815) followed by this sample's share of update_readout_duvenaud
 * (:838-855): out(:,s) = sum_t sum_v ract( R_t . z_t )(:,v).                */
static void duvenaud_forward_sample(const oracle_layer_t *L, const real *params,
                                    int V, const int *ia, const int *ja,
                                    const real *x, const real *e, saved_t *sv,
                                    real *out /* [n_out] */) {
  int D = L->max_deg - L->min_deg + 1;
  sv->P = (real **)calloc(L->T, sizeof(real *));
  sv->H = (real **)calloc(L->T, sizeof(real *));
  sv->Y = (real **)calloc(L->T, sizeof(real *));
  const real *in = x;
  const real *w = params;
  for (int t = 1; t <= L->T; ++t) {
    int Fi = L->nvf[t - 1], Fo = L->nvf[t], K = Fi + L->nef;
    real *A = (real *)malloc(sizeof(real) * (size_t)K * V + 8);
    real *Zp = (real *)malloc(sizeof(real) * (size_t)Fo * V + 8);
    real *Z = (real *)malloc(sizeof(real) * (size_t)Fo * V + 8);
    oracle_duvenaud_propagate(Fi, L->nef, V, in, e, ia, ja, A);
    oracle_duvenaud_update(K, Fo, V, A, w, ia, L->min_deg, L->max_deg, Zp);
    oracle_activation(L->act, Fo, V, Zp, Z);
    sv->Y[t - 1] = Zp;
    sv->P[t - 1] = A;
    sv->H[t - 1] = Z;
    in = Z;
    w += (size_t)Fo * K * D;
  }
  for (int o = 0; o < L->n_out; ++o) out[o] = 0;
  for (int t = 1; t <= L->T; ++t) {
    int Fo = L->nvf[t];
    real *Y = (real *)malloc(sizeof(real) * (size_t)L->n_out * V + 8);
    real *S = (real *)malloc(sizeof(real) * (size_t)L->n_out * V + 8);
    oracle_matmul(L->n_out, Fo, V, w, sv->H[t - 1], Y);
    oracle_activation(L->ract, L->n_out, V, Y, S);
    /* sum(ptr2, dim=2): over vertices, ascending */
    for (int o = 0; o < L->n_out; ++o) {
      real s = 0;
      for (int v = 0; v < V; ++v) s = s + S[(size_t)v * L->n_out + o];
      out[o] = out[o] + s;
    }
    free(Y);
    free(S);
    w += (size_t)L->n_out * Fo;
  }
}

static void duvenaud_backward_sample(const oracle_layer_t *L, const real *params,
                                     int V, const int *ia, const int *ja,
                                     const saved_t *sv, const real *g_out /*[n_out]*/,
                                     real *dparams, real *dx) {
  int D = L->max_deg - L->min_deg + 1;
  int T = L->T;
  size_t *woff = (size_t *)malloc(sizeof(size_t) * (size_t)(2 * T));
  size_t off = 0;
  for (int t = 1; t <= T; ++t) {
    woff[t - 1] = off;
    off += (size_t)L->nvf[t] * (L->nvf[t - 1] + L->nef) * D;
  }
  for (int t = 1; t <= T; ++t) {
    woff[T + t - 1] = off;
    off += (size_t)L->n_out * L->nvf[t];
  }
  /* gz[t] = gradient arriving at z_t: readout share first */
  real **gz = (real **)calloc(T, sizeof(real *));
  for (int t = 1; t <= T; ++t) {
    int Fo = L->nvf[t], no = L->n_out;
    const real *R = params + woff[T + t - 1];
    real *Y = (real *)malloc(sizeof(real) * (size_t)no * V + 8);
    real *S = (real *)malloc(sizeof(real) * (size_t)no * V + 8);
    real *G = (real *)malloc(sizeof(real) * (size_t)no * V + 8);
    real *dY = (real *)malloc(sizeof(real) * (size_t)no * V + 8);
    oracle_matmul(no, Fo, V, R, sv->H[t - 1], Y);
    oracle_activation(L->ract, no, V, Y, S);
    for (int v = 0; v < V; ++v)
      for (int o = 0; o < no; ++o) G[(size_t)v * no + o] = g_out[o];
    oracle_activation_bwd(L->ract, no, V, S, G, dY);
    oracle_matmul_bwd_left(no, Fo, V, dY, sv->H[t - 1], dparams + woff[T + t - 1]);
    gz[t - 1] = (real *)malloc(sizeof(real) * (size_t)Fo * V + 8);
    oracle_matmul_bwd_right(no, Fo, V, R, dY, gz[t - 1]);
    free(Y);
    free(S);
    free(G);
    free(dY);
  }
  for (int t = T; t >= 1; --t) {
    int Fi = L->nvf[t - 1], Fo = L->nvf[t], K = Fi + L->nef;
    real *gzp = (real *)malloc(sizeof(real) * (size_t)Fo * V + 8);
    oracle_activation_bwd_x(L->act, Fo, V, sv->Y[t - 1], sv->H[t - 1], gz[t - 1], gzp);
    oracle_duvenaud_update_bwd_weight(K, Fo, V, gzp, sv->P[t - 1], ia, L->min_deg,
                                      L->max_deg, dparams + woff[t - 1]);
    if (t > 1 || dx) {
      real *dA = (real *)malloc(sizeof(real) * (size_t)K * V + 8);
      oracle_duvenaud_update_bwd_val(K, Fo, V, gzp, params + woff[t - 1], ia,
                                     L->min_deg, L->max_deg, dA);
      real *gin = (real *)malloc(sizeof(real) * (size_t)Fi * V + 8);
      oracle_duvenaud_propagate_bwd_left(Fi, L->nef, V, dA, ia, ja, gin);
      if (t > 1) {
        for (size_t i = 0; i < (size_t)Fi * V; ++i) gz[t - 2][i] = gz[t - 2][i] + gin[i];
      } else {
        memcpy(dx, gin, sizeof(real) * (size_t)Fi * V);
      }
      free(gin);
      free(dA);
    }
    free(gzp);
  }
  for (int t = 0; t < T; ++t) free(gz[t]);
  free(gz);
  free(woff);
}

/*
 * Forward (+ optional loss/backward) of a stack of msgpass layers over a
 * batch, serial over samples like network%forward / loss%grad_reverse
 * (athena_network_sub.f90:2639-2768, 3637-3645).
 *
 * Stack rule: each layer's vertex input is the previous layer's node-level
 * output; a Duvenaud layer (always last) reads the ORIGINAL edge features.
 *
 * Loss (only if target != NULL):
 *   last layer Kipf     -> graph output: sum_s mean_{F,V_s}((p-e)^2)/2
 *                          (athena_loss.f90:414-427), target [V_tot][F_T]
 *   last layer Duvenaud -> one [n_out, B] cell: mean over n_out*global_B /2,
 *                          target [B][n_out]
 * global_B lets a data-parallel shard use the full-batch normalisation
 * (SURVEY 8e); pass B for single-process use.
 *
 * out      : Kipf-last: [V_tot][F_T]; Duvenaud-last: [B][n_out]
 * dparams  : flat, zeroed here, layer order x params order
 *            (athena_base_layer_sub.f90:545-571)
 * returns the loss (0 when target == NULL).
 */
API real oracle_stack_fwd_bwd(int n_layers, const oracle_layer_t *layers,
                              const real *params, int B, const int *nv,
                              const int *ne, const int *ia_cat, const int *ja_cat,
                              const real *x_cat, const real *e_cat,
                              const real *target, int global_B, real *out,
                              real *dparams) {
  const oracle_layer_t *last = &layers[n_layers - 1];
  size_t np_total = 0;
  size_t *poff = (size_t *)malloc(sizeof(size_t) * (size_t)n_layers);
  for (int l = 0; l < n_layers; ++l) {
    poff[l] = np_total;
    np_total += (size_t)oracle_layer_num_params(&layers[l]);
  }
  if (dparams) memset(dparams, 0, sizeof(real) * np_total);
  int F0 = layers[0].nvf[0];
  int Fe = 0;
  for (int l = 0; l < n_layers; ++l)
    if (layers[l].kind == 1) Fe = layers[l].nef;
  const int n_last = last->kind == 0 ? 0 : graph_level_width(last);
  real loss = 0;
  const int *ia = ia_cat;
  const int *ja = ja_cat;
  size_t voff = 0, eoff = 0;
  saved_t *sv = (saved_t *)calloc(n_layers, sizeof(saved_t));
  /* output width / rows of every layer (node-level layers: rows = V of the sample) */
  const real **outp = (const real **)calloc(n_layers, sizeof(real *));
  real **catb = (real **)calloc(n_layers, sizeof(real *));
  real **gacc = (real **)calloc(n_layers, sizeof(real *));
  for (int s = 0; s < B; ++s) {
    int V = nv[s];
    int nz = ia[V] - 1;
    const real *xs = x_cat + voff * F0;
    const real *in = xs; /* most recent node-level output */
    const real *es = e_cat ? e_cat + eoff * Fe : NULL;
    real *dout = NULL;      /* graph-level output of this sample (final layer) */
    real *duv_tmp = NULL;   /* Duvenaud output when dense layers follow it */
    const real *vec = NULL; /* current graph-level vector */
    for (int l = 0; l < n_layers; ++l) {
      const oracle_layer_t *L = &layers[l];
      catb[l] = NULL;
      if (L->kind == 0) {
        const real *lin = in;
        if (L->n_in > 0) {
          /* concatenate the sources along the feature axis, in list order */
          int Fin = L->nvf[0];
          real *cat = (real *)malloc(sizeof(real) * (size_t)Fin * V + 8);
          int off = 0;
          for (int j = 0; j < L->n_in; ++j) {
            int src = L->in[j];
            int w = src < 0 ? F0 : layers[src].nvf[layers[src].T];
            const real *sp = src < 0 ? xs : outp[src];
            for (int v = 0; v < V; ++v)
              memcpy(cat + (size_t)v * Fin + off, sp + (size_t)v * w, sizeof(real) * (size_t)w);
            off += w;
          }
          catb[l] = cat;
          lin = cat;
        }
        in = kipf_forward_sample(L, params + poff[l], V, ia, ja, lin, &sv[l]);
        outp[l] = in;
      } else if (L->kind == 1) {
        real *dst = out + (size_t)s * L->n_out;
        if (l != n_layers - 1) {
          duv_tmp = (real *)malloc(sizeof(real) * (size_t)L->n_out + 8);
          dst = duv_tmp;
        }
        duvenaud_forward_sample(L, params + poff[l], V, ia, ja, in, es, &sv[l], dst);
        vec = dst;
        outp[l] = dst;
      } else {
        vec = full_forward_sample(L, params + poff[l], vec, &sv[l]);
        outp[l] = vec;
      }
    }
    if (last->kind != 0) {
      dout = out + (size_t)s * n_last;
      if (last->kind == 2) memcpy(dout, vec, sizeof(real) * (size_t)n_last);
    }
    int FT = last->nvf[last->T];
    size_t out_off = 0;
    if (last->kind == 0) {
      /* voff counts vertices; output feature width is FT */
      out_off = voff * (size_t)FT;
      memcpy(out + out_off, in, sizeof(real) * (size_t)FT * V);
    }
    if (target) {
      real *g = NULL;
      if (last->kind == 0) {
        size_t n = (size_t)FT * V;
        loss = loss + oracle_mse_cell(n, out + out_off, target + out_off);
        g = (real *)malloc(sizeof(real) * n + 8);
        oracle_mse_cell_bwd(n, out + out_off, target + out_off, g, (real)n);
      } else {
        size_t n = (size_t)n_last;
        const real *ts = target + (size_t)s * n;
        real denom = (real)((size_t)n_last * (size_t)global_B);
        real sq = 0;
        for (size_t i = 0; i < n; ++i) {
          real d = dout[i] - ts[i];
          sq = sq + d * d;
        }
        loss = loss + sq / denom / (real)2;
        g = (real *)malloc(sizeof(real) * n + 8);
        oracle_mse_cell_bwd(n, dout, ts, g, denom);
      }
      if (dparams) {
        /* reverse sweep: the gradient of every layer output is the sum of what its consumers
         * send back (one consumer in a plain stack) */
        for (int l = 0; l < n_layers; ++l) gacc[l] = NULL;
        gacc[n_layers - 1] = g;
        g = NULL;
        for (int l = n_layers - 1; l >= 0; --l) {
          const oracle_layer_t *L = &layers[l];
          real *gl = gacc[l];
          if (!gl) continue; /* nobody consumed this layer's output */
          int nsrc = (L->kind == 0 && L->n_in > 0) ? L->n_in : 1;
          int srcs[4], need_dx = 0;
          for (int j = 0; j < nsrc; ++j) {
            srcs[j] = (L->kind == 0 && L->n_in > 0) ? L->in[j] : l - 1;
            if (srcs[j] >= 0) need_dx = 1;
          }
          int Fi = L->nvf[0];
          size_t dx_n = L->kind == 2 ? (size_t)Fi : (size_t)Fi * V;
          real *dx = need_dx ? (real *)malloc(sizeof(real) * dx_n + 8) : NULL;
          if (L->kind == 0)
            kipf_backward_sample(L, params + poff[l], V, ia, ja, &sv[l], gl, dparams + poff[l], dx);
          else if (L->kind == 1)
            duvenaud_backward_sample(L, params + poff[l], V, ia, ja, &sv[l], gl,
                                     dparams + poff[l], dx);
          else
            full_backward_sample(L, params + poff[l], &sv[l], gl, dparams + poff[l], dx);
          if (dx) {
            int off = 0;
            for (int j = 0; j < nsrc; ++j) {
              int src = srcs[j];
              int w = (L->kind == 0 && L->n_in > 0)
                          ? (src < 0 ? F0 : layers[src].nvf[layers[src].T])
                          : Fi;
              if (src >= 0) {
                size_t rows = L->kind == 2 ? 1 : (size_t)V;
                if (!gacc[src]) gacc[src] = (real *)calloc(rows * (size_t)w + 2, sizeof(real));
                for (size_t v = 0; v < rows; ++v)
                  for (int f = 0; f < w; ++f)
                    gacc[src][v * w + f] = gacc[src][v * w + f] + dx[v * (size_t)Fi + off + f];
              }
              off += w;
            }
            free(dx);
          }
          free(gl);
          gacc[l] = NULL;
        }
      }
      free(g);
    }
    for (int l = 0; l < n_layers; ++l) {
      saved_free(&sv[l], layers[l].T);
      free(catb[l]);
      catb[l] = NULL;
    }
    free(duv_tmp);
    ia += V + 1;
    ja += 2 * (size_t)nz;
    voff += (size_t)V;
    eoff += (size_t)ne[s];
  }
  free(outp);
  free(catb);
  free(gacc);
  free(sv);
  free(poff);
  return loss;
}

/* Layer-level backward with an explicit upstream gradient (no loss): used to
 * check athena_cuda_layer_backward.  g_out: Kipf [V_tot][F_T], Duvenaud
 * [B][n_out].  dx_cat may be NULL.                                          */
API void oracle_layer_fwd_bwd(const oracle_layer_t *L, const real *params, int B,
                              const int *nv, const int *ne, const int *ia_cat,
                              const int *ja_cat, const real *x_cat,
                              const real *e_cat, const real *g_out, real *out,
                              real *dparams, real *dx_cat) {
  size_t np = (size_t)oracle_layer_num_params(L);
  if (dparams) memset(dparams, 0, sizeof(real) * np);
  const int *ia = ia_cat;
  const int *ja = ja_cat;
  size_t voff = 0, eoff = 0;
  int F0 = L->nvf[0], FT = L->nvf[L->T];
  for (int s = 0; s < B; ++s) {
    int V = nv[s];
    int nz = ia[V] - 1;
    saved_t sv;
    const real *xs = x_cat + voff * F0;
    if (L->kind == 2) {
      /* dense head on its own: x_cat is the graph-level [B][num_inputs] array */
      const real *o = full_forward_sample(L, params, x_cat + (size_t)s * F0, &sv);
      if (out) memcpy(out + (size_t)s * FT, o, sizeof(real) * (size_t)FT);
      if (g_out)
        full_backward_sample(L, params, &sv, g_out + (size_t)s * FT, dparams,
                             dx_cat ? dx_cat + (size_t)s * F0 : NULL);
    } else if (L->kind == 0) {
      const real *o = kipf_forward_sample(L, params, V, ia, ja, xs, &sv);
      if (out) memcpy(out + voff * FT, o, sizeof(real) * (size_t)FT * V);
      if (g_out)
        kipf_backward_sample(L, params, V, ia, ja, &sv, g_out + voff * FT, dparams,
                             dx_cat ? dx_cat + voff * F0 : NULL);
    } else {
      real *tmp = (real *)malloc(sizeof(real) * (size_t)L->n_out + 8);
      duvenaud_forward_sample(L, params, V, ia, ja, xs,
                              e_cat ? e_cat + eoff * L->nef : NULL, &sv, tmp);
      if (out) memcpy(out + (size_t)s * L->n_out, tmp, sizeof(real) * (size_t)L->n_out);
      free(tmp);
      if (g_out)
        duvenaud_backward_sample(L, params, V, ia, ja, &sv,
                                 g_out + (size_t)s * L->n_out, dparams,
                                 dx_cat ? dx_cat + voff * F0 : NULL);
    }
    saved_free(&sv, L->T);
    ia += V + 1;
    ja += 2 * (size_t)nz;
    voff += (size_t)V;
    eoff += (size_t)ne[s];
  }
}

typedef struct {
  int kind; /* 0 = sgd, 1 = adam, 2 = rmsprop (beta in beta1), 3 = adagrad */
  real lr, beta1, beta2, eps, momentum;
  int nesterov;
  int clip_flags;
  real clip_min, clip_max, clip_norm;
  int reg; /* 0 none, 1 l1, 2 l2, 3 l1l2 (athena_regulariser.f90) */
  real l1, l2;
  int l2_decoupled;
} oracle_optim_t;

/* regulariser%regularise(param, gradient, learning_rate), called at the top of every
 * minimise_* (athena_regulariser.f90:99, 117, 135-136).                              */
API void oracle_regularise(int n, const real *p, real *g, int reg, real lr, real l1, real l2) {
  for (int i = 0; i < n; ++i) {
    real sgn = R_COPYSIGN((real)1, p[i]);
    if (reg == 1)
      g[i] = g[i] + lr * l1 * sgn;
    else if (reg == 2)
      g[i] = g[i] + lr * (real)2 * l2 * p[i];
    else if (reg == 3)
      g[i] = g[i] + lr * (l1 * sgn + (real)2 * l2 * p[i]);
  }
}

/* minimise_adam with an l2 regulariser: AdamW (decoupled) or L2 inside the quotient.
 * src/athena/athena_optimiser.f90:1049-1078                                          */
API void oracle_adam_l2(int n, real *p, const real *g, real *m, real *v, real lr, real b1,
                        real b2, real eps, int iter, real l2, int decoupled) {
  real bc1 = (real)1 - powi(b1, iter), bc2 = (real)1 - powi(b2, iter);
  for (int i = 0; i < n; ++i) {
    m[i] = b1 * m[i] + ((real)1 - b1) * g[i];
    v[i] = b2 * v[i] + ((real)1 - b2) * g[i] * g[i];
    real mh = m[i] / bc1, vh = v[i] / bc2;
    if (decoupled) {
      p[i] = p[i] - lr * l2 * p[i];
      p[i] = p[i] - lr * (mh / (R_SQRT(vh) + eps));
    } else {
      p[i] = p[i] - lr * ((mh + l2 * p[i]) / (R_SQRT(vh) + eps));
    }
  }
}

/* minimise_rmsprop: moving_avg = beta*moving_avg + (1-beta)*g^2 ;
 * param -= lr * g / sqrt(moving_avg + eps).  src/athena/athena_optimiser.f90:771-803  */
API void oracle_rmsprop(int n, real *p, const real *g, real *avg, real lr, real beta,
                        real eps) {
  for (int i = 0; i < n; ++i) {
    avg[i] = beta * avg[i] + ((real)1 - beta) * (g[i] * g[i]);
    p[i] = p[i] - lr * g[i] / R_SQRT(avg[i] + eps);
  }
}

/* minimise_adagrad: sum_squares += g^2 ; param -= lr * g / sqrt(sum_squares + eps).
 * src/athena/athena_optimiser.f90:898-925                                              */
API void oracle_adagrad(int n, real *p, const real *g, real *ss, real lr, real eps) {
  for (int i = 0; i < n; ++i) {
    ss[i] = ss[i] + g[i] * g[i];
    p[i] = p[i] - lr * g[i] / R_SQRT(ss[i] + eps);
  }
}

/* network%update: iter already incremented by the caller; grads are summed
 * over samples (single column) so no column mean applies.
 * src/athena/athena_network_sub.f90:2816-2929                               */
API void oracle_update(int n, real *params, real *grads, const oracle_optim_t *o,
                       real *state1, real *state2, int iter) {
  oracle_clip(n, grads, o->clip_flags, o->clip_min, o->clip_max, o->clip_norm);
  if (o->reg) oracle_regularise(n, params, grads, o->reg, o->lr, o->l1, o->l2);
  if (o->kind == 1 && o->reg == 2)
    oracle_adam_l2(n, params, grads, state1, state2, o->lr, o->beta1, o->beta2, o->eps, iter,
                   o->l2, o->l2_decoupled);
  else if (o->kind == 0)
    oracle_sgd(n, params, grads, state1, o->lr, o->momentum, o->nesterov);
  else if (o->kind == 1)
    oracle_adam(n, params, grads, state1, state2, o->lr, o->beta1, o->beta2, o->eps, iter);
  else if (o->kind == 2)
    oracle_rmsprop(n, params, grads, state1, o->lr, o->beta1, o->eps);
  else
    oracle_adagrad(n, params, grads, state1, o->lr, o->eps);
}

/* one full train step of the batch loop (athena_network_sub.f90:3611-3670):
 * forward, loss, grad_reverse, update.  Returns the batch loss.             */
API real oracle_train_step(int n_layers, const oracle_layer_t *layers, real *params,
                           int B, const int *nv, const int *ne, const int *ia_cat,
                           const int *ja_cat, const real *x_cat, const real *e_cat,
                           const real *target, real *out, real *grads,
                           const oracle_optim_t *o, real *state1, real *state2,
                           int iter) {
  size_t np = 0;
  for (int l = 0; l < n_layers; ++l) np += (size_t)oracle_layer_num_params(&layers[l]);
  real loss = oracle_stack_fwd_bwd(n_layers, layers, params, B, nv, ne, ia_cat,
                                   ja_cat, x_cat, e_cat, target, B, out, grads);
  oracle_update((int)np, params, grads, o, state1, state2, iter);
  return loss;
}
