/*
 * fake_athena.c -- a C stand-in for the Fortran host side of athena, calling libathena_cuda
 * through the C ABI in exactly the order the reference's network%train does (SURVEY.md
 * sections 3.1 - 3.5; athena_network_sub.f90:3611-3670):
 *
 *     get_sample / set_graph          -> athena_cuda_batch_create            (:3628, :2729)
 *     forward (per layer)             -> athena_cuda_layer_forward           (:3637, :2752)
 *     loss_eval (host, MSE)           -> per-sample cells, athena_loss.f90:393-430   (:3643)
 *     loss%grad_reverse               -> one get_partial callback per sample =
 *                                        athena_cuda_layer_backward_stage    (:3645)
 *     update                          -> get_gradients, clip + minimise, set_params  (:3667)
 * and, for a network that lives on the device as a whole,
 *     one iteration of the batch loop -> athena_cuda_network_train_step.
 *
 * TEST INFRASTRUCTURE: the results are checked against the CPU oracle (oracle/athena_oracle.c,
 * linked as liboracle_f32.so).  Built with gcc by the CPU test suite (tests/test_c_driver.py),
 * executed by the GPU suite.  Exit codes: 0 parity ok, 1 mismatch, 2 ABI error, 3 no device
 * (there is no CPU fallback).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "athena_cuda.h"

/* ---- the oracle's C interface (oracle/athena_oracle.c, float build) ---- */
typedef struct {
  int kind, T, nvf[17], nef, min_deg, max_deg, n_out, act, ract, use_bias, n_in, in[4];
} oracle_layer_t;
typedef struct {
  int kind;
  float lr, beta1, beta2, eps, momentum;
  int nesterov, clip_flags;
  float clip_min, clip_max, clip_norm;
  int reg;
  float l1, l2;
  int l2_decoupled;
} oracle_optim_t;
int oracle_layer_num_params(const oracle_layer_t* L);
void oracle_layer_fwd_bwd(const oracle_layer_t* L, const float* params, int B, const int* nv,
                          const int* ne, const int* ia, const int* ja, const float* x,
                          const float* e, const float* g_out, float* out, float* dparams,
                          float* dx);
float oracle_mse_cell(size_t n, const float* p, const float* e);
void oracle_mse_cell_bwd(size_t n, const float* p, const float* e, float* g, float denom);
void oracle_update(int n, float* params, float* grads, const oracle_optim_t* o, float* s1,
                   float* s2, int iter);
float oracle_train_step(int n_layers, const oracle_layer_t* layers, float* params, int B,
                        const int* nv, const int* ne, const int* ia, const int* ja,
                        const float* x, const float* e, const float* target, float* out,
                        float* grads, const oracle_optim_t* o, float* s1, float* s2, int iter);

#define CHECK(call)                                                                   \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != 0) {                                                                   \
      fprintf(stderr, "fake_athena: %s -> %d: %s\n", #call, rc_, athena_cuda_last_error()); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

static unsigned long long g_seed = 88172645463325252ull;
static float frand(void) { /* xorshift, uniform in (-0.5, 0.5) */
  g_seed ^= g_seed << 13;
  g_seed ^= g_seed >> 7;
  g_seed ^= g_seed << 17;
  return (float)((g_seed >> 11) * (1.0 / 9007199254740992.0)) - 0.5f;
}

/* a mini-batch of graph_type samples in the packed layout of athena_cuda_batch_create */
typedef struct {
  int B, V, Z, E;
  int *nv, *ne, *nz, *ia, *ja;
  float *x, *e;
} batch_t;

/* graph s: a ring of nv vertices (+ one chord), every undirected edge k listed in both endpoint
 * rows with edge id k, and a self loop with edge id 0 at the end of every row (add_self_loops) */
static batch_t make_batch(int B, const int* nv, int F, int Fe) {
  batch_t b;
  memset(&b, 0, sizeof(b));
  b.B = B;
  b.nv = malloc(sizeof(int) * B);
  b.ne = malloc(sizeof(int) * B);
  b.nz = malloc(sizeof(int) * B);
  for (int s = 0; s < B; ++s) {
    int n = nv[s], ne = n + (n >= 4 ? 1 : 0);
    if (n < 3) ne = n - 1; /* a path */
    b.nv[s] = n;
    b.ne[s] = ne;
    b.nz[s] = 2 * ne + n;
    b.V += n;
    b.E += ne;
    b.Z += b.nz[s];
  }
  b.ia = malloc(sizeof(int) * (b.V + B));
  b.ja = malloc(sizeof(int) * 2 * b.Z);
  int* iap = b.ia;
  int* jap = b.ja;
  for (int s = 0; s < B; ++s) {
    int n = b.nv[s], ne = b.ne[s];
    int(*edges)[2] = malloc(sizeof(int[2]) * (ne > 0 ? ne : 1));
    int k = 0;
    if (n >= 3)
      for (int i = 0; i < n; ++i, ++k) edges[k][0] = i, edges[k][1] = (i + 1) % n;
    else
      for (int i = 0; i + 1 < n; ++i, ++k) edges[k][0] = i, edges[k][1] = i + 1;
    if (n >= 4) edges[k][0] = 0, edges[k][1] = n / 2, ++k;
    int w = 0;
    for (int v = 0; v < n; ++v) {
      iap[v] = w + 1;
      for (int q = 0; q < ne; ++q) {
        if (edges[q][0] == v) jap[2 * w] = edges[q][1] + 1, jap[2 * w + 1] = q + 1, ++w;
        else if (edges[q][1] == v) jap[2 * w] = edges[q][0] + 1, jap[2 * w + 1] = q + 1, ++w;
      }
      jap[2 * w] = v + 1, jap[2 * w + 1] = 0, ++w; /* self loop, no edge feature */
    }
    iap[n] = w + 1;
    iap += n + 1;
    jap += 2 * w;
    free(edges);
  }
  b.x = malloc(sizeof(float) * b.V * F);
  for (int i = 0; i < b.V * F; ++i) b.x[i] = frand() * 2.f;
  b.e = NULL;
  if (Fe > 0) {
    b.e = malloc(sizeof(float) * b.E * Fe);
    for (int i = 0; i < b.E * Fe; ++i) b.e[i] = frand() + 0.6f;
  }
  return b;
}

static double rel_err(const float* a, const float* ref, size_t n) {
  double d = 0, m = 1e-30;
  for (size_t i = 0; i < n; ++i) {
    double x = fabs((double)a[i] - ref[i]);
    if (x > d) d = x;
    if (fabs(ref[i]) > m) m = fabs(ref[i]);
  }
  return d / m;
}

static int g_fail = 0;
static void expect(const char* what, double err, double tol) {
  printf("  %-46s rel err %.3e (tol %.0e) %s\n", what, err, tol, err <= tol ? "ok" : "MISMATCH");
  if (!(err <= tol)) g_fail = 1;
}

/* ---- scenario A: one device layer inside a host-driven train step --------------------- */
static void layer_scenario(const oracle_layer_t* spec, const batch_t* b, athena_handle_t layer,
                           const char* name) {
  printf("%s\n", name);
  const int kipf = spec->kind == 0;
  const int W = kipf ? spec->nvf[spec->T] : spec->n_out;
  const long out_n = (long)(kipf ? b->V : b->B) * W;
  const int np = oracle_layer_num_params(spec);
  int64_t np_dev = 0;
  CHECK(athena_cuda_layer_num_params(layer, &np_dev));
  if (np_dev != np) { fprintf(stderr, "num_params %lld vs %d\n", (long long)np_dev, np); exit(1); }
  float* params = malloc(sizeof(float) * np);
  for (int i = 0; i < np; ++i) params[i] = frand() * 0.8f;
  CHECK(athena_cuda_layer_set_params(layer, params, np));
  float* target = malloc(sizeof(float) * out_n);
  for (long i = 0; i < out_n; ++i) target[i] = frand();

  /* set_graph: once per forward, validated like athena_duvenaud_msgpass_layer.f90:632-639 */
  athena_handle_t batch = 0;
  CHECK(athena_cuda_batch_create(&batch, b->B, b->nv, b->ne, b->nz, b->ia, b->ja, ATHENA_MEM_HOST, 1));
  /* forward */
  float* out = malloc(sizeof(float) * out_n);
  CHECK(athena_cuda_layer_forward(layer, batch, b->x, b->e, out, ATHENA_MEM_HOST));
  /* loss_eval on the host + the upstream gradient each sample's node receives */
  float* g = malloc(sizeof(float) * out_n);
  float loss = 0.f;
  if (kipf) { /* one MSE cell per sample, mean over F_T * nv_s (athena_loss.f90:416-427) */
    long off = 0;
    for (int s = 0; s < b->B; ++s) {
      size_t n = (size_t)b->nv[s] * W;
      loss += oracle_mse_cell(n, out + off, target + off);
      oracle_mse_cell_bwd(n, out + off, target + off, g + off, (float)n);
      off += n;
    }
  } else { /* one [num_outputs, batch] cell (athena_loss.f90:414) */
    loss = oracle_mse_cell((size_t)out_n, out, target);
    oracle_mse_cell_bwd((size_t)out_n, out, target, g, (float)out_n);
  }
  /* grad_reverse: one callback per sample, in reverse sample order as a DAG walk may do */
  CHECK(athena_cuda_layer_zero_gradients(layer));
  long off_end = out_n;
  for (int s = b->B - 1; s >= 0; --s) {
    long n = kipf ? (long)b->nv[s] * W : W;
    off_end -= n;
    CHECK(athena_cuda_layer_backward_stage(layer, batch, s, g + off_end, n));
  }
  float* grads = malloc(sizeof(float) * np);
  CHECK(athena_cuda_layer_get_gradients(layer, grads, np));

  /* the oracle, fed the same upstream gradient */
  float* out_ref = malloc(sizeof(float) * out_n);
  float* dp_ref = malloc(sizeof(float) * np);
  oracle_layer_fwd_bwd(spec, params, b->B, b->nv, b->ne, b->ia, b->ja, b->x, b->e, g, out_ref,
                       dp_ref, NULL);
  expect("forward output", rel_err(out, out_ref, out_n), 1e-5);
  expect("parameter gradients (staged per sample)", rel_err(grads, dp_ref, np), 1e-5);

  /* network%update on the host (flat vectors, athena_network_sub.f90:2847-2927), new
   * parameters pushed back with set_params; the next forward must see them */
  oracle_optim_t opt;
  memset(&opt, 0, sizeof(opt));
  opt.kind = 0;
  opt.lr = 0.05f;
  float* s1 = calloc(np, sizeof(float));
  float* s2 = calloc(np, sizeof(float));
  float* p_ref = malloc(sizeof(float) * np);
  memcpy(p_ref, params, sizeof(float) * np);
  oracle_update(np, p_ref, dp_ref, &opt, s1, s2, 1);
  float* p_dev = malloc(sizeof(float) * np);
  memcpy(p_dev, params, sizeof(float) * np);
  memset(s1, 0, sizeof(float) * np);
  oracle_update(np, p_dev, grads, &opt, s1, s2, 1); /* host optimiser on the DEVICE gradients */
  CHECK(athena_cuda_layer_set_params(layer, p_dev, np));
  CHECK(athena_cuda_layer_zero_gradients(layer));
  CHECK(athena_cuda_batch_destroy(batch));
  CHECK(athena_cuda_batch_create(&batch, b->B, b->nv, b->ne, b->nz, b->ia, b->ja, ATHENA_MEM_HOST, 1));
  CHECK(athena_cuda_layer_forward(layer, batch, b->x, b->e, out, ATHENA_MEM_HOST));
  oracle_layer_fwd_bwd(spec, p_ref, b->B, b->nv, b->ne, b->ia, b->ja, b->x, b->e, NULL, out_ref,
                       NULL, NULL);
  expect("forward after the update", rel_err(out, out_ref, out_n), 1e-5);
  CHECK(athena_cuda_batch_destroy(batch));
  (void)loss;
  free(params); free(target); free(out); free(g); free(grads); free(out_ref); free(dp_ref);
  free(s1); free(s2); free(p_ref); free(p_dev);
}

/* ---- scenario B: the whole network on the device, the batch loop of network%train ------ */
static void network_scenario(int n_layers, const oracle_layer_t* specs, const athena_handle_t* layers,
                             const batch_t* b, const oracle_optim_t* oo,
                             const athena_optimiser_desc* od, const char* name) {
  printf("%s\n", name);
  const oracle_layer_t* last = &specs[n_layers - 1];
  const int kipf = last->kind == 0;
  const int W = kipf ? last->nvf[last->T] : last->n_out;
  const long out_n = (long)(kipf ? b->V : b->B) * W;
  athena_handle_t net = 0;
  CHECK(athena_cuda_network_create(&net));
  int np = 0;
  for (int l = 0; l < n_layers; ++l) {
    CHECK(athena_cuda_network_add(net, layers[l]));
    np += oracle_layer_num_params(&specs[l]);
  }
  CHECK(athena_cuda_network_compile(net, od));
  float* params = malloc(sizeof(float) * np);
  for (int i = 0; i < np; ++i) params[i] = frand() * 0.6f;
  CHECK(athena_cuda_network_set_params(net, params, np));
  float* target = malloc(sizeof(float) * out_n);
  for (long i = 0; i < out_n; ++i) target[i] = frand();
  float* s1 = calloc(np, sizeof(float));
  float* s2 = calloc(np, sizeof(float));
  float* grads = malloc(sizeof(float) * np);
  float* out_ref = malloc(sizeof(float) * out_n);
  double worst_loss = 0;
  for (int it = 1; it <= 3; ++it) {
    athena_handle_t batch = 0; /* get_sample + set_graph every iteration, as the reference does */
    CHECK(athena_cuda_batch_create(&batch, b->B, b->nv, b->ne, b->nz, b->ia, b->ja, ATHENA_MEM_HOST, 0));
    float loss = 0.f;
    CHECK(athena_cuda_network_train_step(net, batch, b->x, b->e, target, ATHENA_MEM_HOST, b->B, &loss));
    CHECK(athena_cuda_batch_destroy(batch));
    float loss_ref = oracle_train_step(n_layers, specs, params, b->B, b->nv, b->ne, b->ia, b->ja,
                                       b->x, b->e, target, out_ref, grads, oo, s1, s2, it);
    double e = fabs((double)loss - loss_ref) / fmax(fabs((double)loss_ref), 1e-30);
    if (e > worst_loss) worst_loss = e;
  }
  float* p_dev = malloc(sizeof(float) * np);
  CHECK(athena_cuda_network_get_params(net, p_dev, np));
  expect("batch loss over 3 iterations", worst_loss, 1e-5);
  expect("parameters after 3 iterations", rel_err(p_dev, params, np), 1e-4);
  CHECK(athena_cuda_network_destroy(net));
  free(params); free(target); free(s1); free(s2); free(grads); free(out_ref); free(p_dev);
}

/* ---- scenario C: example/msgpass_euler's shape of program: the reader hands over EDGE LISTS,
 * generate_adjacency + add_self_loops run on the device, the layers are added with
 * input_list = [0, -1] / operator "concatenate", the batch loop runs on the device ---------- */
static void skip_scenario(const batch_t* b, const char* name) {
  printf("%s\n", name);
  /* the edge lists make_batch built its CSR from: edge k of a graph is listed in both rows with
   * id k, so row v's entries (nb > v, id k) give back (v, nb) */
  int* il = malloc(sizeof(int) * 2 * (b->E > 0 ? b->E : 1));
  {
    const int* ia = b->ia;
    const int* ja = b->ja;
    int e0 = 0;
    for (int s = 0; s < b->B; ++s) {
      for (int v = 0; v < b->nv[s]; ++v)
        for (int w = ia[v] - 1; w < ia[v + 1] - 1; ++w) {
          const int nb = ja[2 * w], k = ja[2 * w + 1];
          if (k > 0 && nb > v + 1) il[2 * (e0 + k - 1)] = v + 1, il[2 * (e0 + k - 1) + 1] = nb;
        }
      ja += 2 * b->nz[s];
      ia += b->nv[s] + 1;
      e0 += b->ne[s];
    }
  }
  /* kipf [4 -> 6, softmax] ; kipf [4 + 6 -> 5, softmax] <- [input | previous] ;
   * kipf [4 + 5 -> 3, swish] <- [input | previous] */
  oracle_layer_t sp[3];
  memset(sp, 0, sizeof(sp));
  const int widths[3][2] = {{4, 6}, {10, 5}, {9, 3}};
  const int acts[3] = {ATHENA_ACT_SOFTMAX, ATHENA_ACT_SOFTMAX, ATHENA_ACT_SWISH};
  athena_handle_t net = 0, L[3] = {0, 0, 0};
  CHECK(athena_cuda_network_create(&net));
  int np = 0;
  for (int l = 0; l < 3; ++l) {
    sp[l].kind = 0; sp[l].T = 1; sp[l].nvf[0] = widths[l][0]; sp[l].nvf[1] = widths[l][1];
    sp[l].act = acts[l];
    if (l > 0) { sp[l].n_in = 2; sp[l].in[0] = -1; sp[l].in[1] = l - 1; }
    const int32_t nvf[2] = {widths[l][0], widths[l][1]};
    CHECK(athena_cuda_kipf_layer_create(&L[l], 1, nvf, acts[l]));
    const int32_t input_list[2] = {0, -1}; /* as example/msgpass_euler/src/main.f90:202 writes it */
    if (l == 0) CHECK(athena_cuda_network_add(net, L[l]));
    else CHECK(athena_cuda_network_add_inputs(net, L[l], 2, input_list, ATHENA_MERGE_CONCATENATE));
    np += oracle_layer_num_params(&sp[l]);
  }
  oracle_optim_t oo;
  memset(&oo, 0, sizeof(oo));
  oo.kind = 1; oo.lr = 0.02f; oo.beta1 = 0.9f; oo.beta2 = 0.999f; oo.eps = 1e-8f;
  oo.clip_flags = 1; oo.clip_min = -1.f; oo.clip_max = 1.f;
  athena_optimiser_desc od;
  memset(&od, 0, sizeof(od));
  od.kind = ATHENA_OPT_ADAM; od.learning_rate = 0.02f; od.beta1 = 0.9f; od.beta2 = 0.999f;
  od.epsilon = 1e-8f; od.clip_min_max = 1; od.clip_min = -1.f; od.clip_max = 1.f; od.l2_decoupled = 1;
  CHECK(athena_cuda_network_compile(net, &od));
  float* params = malloc(sizeof(float) * np);
  for (int i = 0; i < np; ++i) params[i] = frand() * 0.8f;
  CHECK(athena_cuda_network_set_params(net, params, np));
  const long out_n = (long)b->V * 3;
  float* target = malloc(sizeof(float) * out_n);
  for (long i = 0; i < out_n; ++i) target[i] = frand();
  float* s1 = calloc(np, sizeof(float));
  float* s2 = calloc(np, sizeof(float));
  float* grads = malloc(sizeof(float) * np);
  float* out_ref = malloc(sizeof(float) * out_n);
  double worst_loss = 0;
  for (int it = 1; it <= 3; ++it) {
    athena_handle_t batch = 0;
    CHECK(athena_cuda_batch_create_from_edges(&batch, b->B, b->nv, b->ne, il, NULL, 1,
                                              ATHENA_MEM_HOST, 1));
    float loss = 0.f;
    CHECK(athena_cuda_network_train_step(net, batch, b->x, NULL, target, ATHENA_MEM_HOST, b->B, &loss));
    CHECK(athena_cuda_batch_destroy(batch));
    float loss_ref = oracle_train_step(3, sp, params, b->B, b->nv, b->ne, b->ia, b->ja, b->x, NULL,
                                       target, out_ref, grads, &oo, s1, s2, it);
    double e = fabs((double)loss - loss_ref) / fmax(fabs((double)loss_ref), 1e-30);
    if (e > worst_loss) worst_loss = e;
  }
  float* p_dev = malloc(sizeof(float) * np);
  CHECK(athena_cuda_network_get_params(net, p_dev, np));
  expect("batch loss over 3 iterations", worst_loss, 1e-5);
  expect("parameters after 3 iterations", rel_err(p_dev, params, np), 1e-4);
  CHECK(athena_cuda_network_destroy(net));
  free(il); free(params); free(target); free(s1); free(s2); free(grads); free(out_ref); free(p_dev);
}

int main(void) {
  if (athena_cuda_init(-1) != 0) {
    fprintf(stderr, "fake_athena: %s\n", athena_cuda_last_error());
    return 3;
  }
  const int nv[6] = {5, 9, 3, 12, 2, 7};

  { /* A1: kipf_msgpass_layer_type([5, 6, 4], num_time_steps = 2, relu) */
    batch_t b = make_batch(6, nv, 5, 0);
    oracle_layer_t s;
    memset(&s, 0, sizeof(s));
    s.kind = 0; s.T = 2; s.nvf[0] = 5; s.nvf[1] = 6; s.nvf[2] = 4; s.act = ATHENA_ACT_RELU;
    const int32_t nvf[3] = {5, 6, 4};
    athena_handle_t L = 0;
    CHECK(athena_cuda_kipf_layer_create(&L, 2, nvf, ATHENA_ACT_RELU));
    layer_scenario(&s, &b, L, "A1  Kipf layer in a host-driven step (forward / loss / grad_reverse / update)");
    CHECK(athena_cuda_layer_destroy(L));
  }
  { /* A2: duvenaud_msgpass_layer_type([6], [2], T = 3, max degree 4, 5 outputs) */
    batch_t b = make_batch(6, nv, 6, 2);
    oracle_layer_t s;
    memset(&s, 0, sizeof(s));
    s.kind = 1; s.T = 3; s.nef = 2; s.min_deg = 1; s.max_deg = 4; s.n_out = 5;
    s.act = ATHENA_ACT_SIGMOID; s.ract = ATHENA_ACT_SOFTMAX;
    for (int t = 0; t <= 3; ++t) s.nvf[t] = 6;
    const int32_t nvf[4] = {6, 6, 6, 6};
    athena_handle_t L = 0;
    CHECK(athena_cuda_duvenaud_layer_create(&L, 3, nvf, 2, 1, 4, 5, ATHENA_ACT_SIGMOID, ATHENA_ACT_SOFTMAX));
    layer_scenario(&s, &b, L, "A2  Duvenaud layer in a host-driven step");
    CHECK(athena_cuda_layer_destroy(L));
  }
  { /* B: Kipf -> Duvenaud on the device as one network, Adam + clip_norm */
    batch_t b = make_batch(6, nv, 8, 2);
    oracle_layer_t s[2];
    memset(s, 0, sizeof(s));
    s[0].kind = 0; s[0].T = 1; s[0].nvf[0] = 8; s[0].nvf[1] = 8; s[0].act = ATHENA_ACT_TANH;
    s[1].kind = 1; s[1].T = 2; s[1].nef = 2; s[1].min_deg = 1; s[1].max_deg = 3; s[1].n_out = 4;
    s[1].act = ATHENA_ACT_SIGMOID; s[1].ract = ATHENA_ACT_SOFTMAX;
    for (int t = 0; t <= 2; ++t) s[1].nvf[t] = 8;
    const int32_t k_nvf[2] = {8, 8}, d_nvf[3] = {8, 8, 8};
    athena_handle_t L[2] = {0, 0};
    CHECK(athena_cuda_kipf_layer_create(&L[0], 1, k_nvf, ATHENA_ACT_TANH));
    CHECK(athena_cuda_duvenaud_layer_create(&L[1], 2, d_nvf, 2, 1, 3, 4, ATHENA_ACT_SIGMOID, ATHENA_ACT_SOFTMAX));
    oracle_optim_t oo;
    memset(&oo, 0, sizeof(oo));
    oo.kind = 1; oo.lr = 0.01f; oo.beta1 = 0.9f; oo.beta2 = 0.999f; oo.eps = 1e-8f;
    oo.clip_flags = 2; oo.clip_norm = 0.5f;
    athena_optimiser_desc od;
    memset(&od, 0, sizeof(od));
    od.kind = ATHENA_OPT_ADAM; od.learning_rate = 0.01f; od.beta1 = 0.9f; od.beta2 = 0.999f;
    od.epsilon = 1e-8f; od.clip_norm_on = 1; od.clip_norm = 0.5f; od.l2_decoupled = 1;
    network_scenario(2, s, L, &b, &oo, &od, "B   Kipf -> Duvenaud network, batch loop of network%train on the device");
  }
  { /* C: skip-connected Kipf network from edge lists (example/msgpass_euler's call sequence) */
    batch_t b = make_batch(6, nv, 4, 0);
    skip_scenario(&b, "C   edge lists -> device CSR, network%add(input_list = [0, -1], 'concatenate'), swish head");
  }
  CHECK(athena_cuda_shutdown());
  printf(g_fail ? "fake_athena: MISMATCH\n" : "fake_athena: parity ok\n");
  return g_fail;
}
