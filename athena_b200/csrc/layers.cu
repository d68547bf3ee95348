// Host orchestration of the message-passing layers and the train-step
// skeleton around them.  Mirrors (and replaces the bodies of)
//   update_message_kipf        athena_kipf_msgpass_layer.f90:915-959
//   update_message_duvenaud    athena_duvenaud_msgpass_layer.f90:755-817
//   update_readout_duvenaud    athena_duvenaud_msgpass_layer.f90:822-859
//   the reverse sweep loss%grad_reverse performs through them
//   network%forward/train/update athena_network_sub.f90:2639-2929,3611-3670
// The whole mini-batch is processed as ONE block-diagonal graph, so the
// reference's serial `do s = 1, batch` loops become single kernel launches.
#include <algorithm>

#include "athena_internal.h"

namespace athena {

struct Layer : Object {
  Layer() : Object(Kind::Layer) {}
  int kind = 0;  // 0 kipf, 1 duvenaud, 2 full (dense head: nvf = {num_inputs, num_outputs})
  int use_bias = 1;
  int T = 0;
  std::vector<int> nvf;  // (0:T)
  int nef = 0, min_deg = 1, max_deg = 1, n_out = 0, act = 0, ract = 0;
  int64_t num_params = 0;
  std::vector<int64_t> poff;  // flat offset of params(i)
  DevBuf own_params, own_grads;
  float* params = nullptr;
  float* grads = nullptr;
  bool adopted = false;  // parameters live in a network's flat buffer
  bool inference = false;  // current forward is network%predict: nothing is saved for backward
  // saved by forward for the reverse sweep
  std::vector<std::unique_ptr<DevBuf>> P, H, S, GZ, TN;  // TN: per-step dW partials (fused path)
  std::vector<std::unique_ptr<DevBuf>> MK;               // sign bits of step t's pre-activation
  std::vector<std::unique_ptr<DevBuf>> Y;                // pre-activation of step t (swish only)
  std::vector<char> mask_valid;                          // MK[t] written by the last forward
  DevBuf Ae, out_buf, g0, g1, g2, tn_scratch, stage_x, stage_e, stage_g, stage_gin;
  DevBuf tile_part;          // CTA partials of the fused Duvenaud reverse sweep (tile_fma.cu)
  bool tile_S = false;       // the fused forward saved the readouts S_t
  bool tile_fwd = false;     // the last forward ran on the fused tile kernel (saved: z_t only)
  const float* fwd_e = nullptr;  // device edge features of the last forward
  std::vector<char> staged;  // per-sample marks of athena_cuda_layer_backward_stage
  int64_t num_staged = 0;
  Batch* fwd_batch = nullptr;
  int64_t fwd_V = -1;
  const float* fwd_x = nullptr;  // device input of the last forward (valid until the next one)

  int ldA(int t) const { return (int)round_up(nvf[t - 1] + nef, 4); }
  int out_width() const { return kind == 1 ? n_out : nvf[T]; }
  int64_t out_rows(const Batch* b) const { return kind == 0 ? b->V : b->B; }
  int64_t in_rows(const Batch* b) const { return kind == 2 ? b->B : b->V; }
};

static void layer_layout(Layer* L) {
  L->poff.clear();
  int64_t off = 0;
  if (L->kind == 2) {
    // W [num_outputs, num_inputs] column-major, then the bias (athena_full_layer.f90:371-396)
    L->poff.push_back(off);
    off += (int64_t)L->nvf[1] * L->nvf[0];
    L->poff.push_back(off);
    if (L->use_bias) off += L->nvf[1];
  } else if (L->kind == 0) {
    for (int t = 1; t <= L->T; ++t) {
      L->poff.push_back(off);
      off += (int64_t)L->nvf[t] * L->nvf[t - 1];
    }
  } else {
    int D = L->max_deg - L->min_deg + 1;
    for (int t = 1; t <= L->T; ++t) {
      L->poff.push_back(off);
      off += (int64_t)L->nvf[t] * (L->nvf[t - 1] + L->nef) * D;
    }
    for (int t = 1; t <= L->T; ++t) {
      L->poff.push_back(off);
      off += (int64_t)L->n_out * L->nvf[t];
    }
  }
  L->num_params = off;
  L->P.clear();
  L->H.clear();
  L->S.clear();
  L->GZ.clear();
  L->TN.clear();
  L->MK.clear();
  L->Y.clear();
  L->mask_valid.assign(L->T, 0);
  for (int t = 0; t < L->T; ++t) {
    L->Y.emplace_back(new DevBuf);
    L->TN.emplace_back(new DevBuf);
    L->MK.emplace_back(new DevBuf);
    L->P.emplace_back(new DevBuf);
    L->H.emplace_back(new DevBuf);
    L->S.emplace_back(new DevBuf);
    L->GZ.emplace_back(new DevBuf);
  }
}

static int layer_alloc_params(Layer* L) {
  size_t bytes = sizeof(float) * (size_t)std::max<int64_t>(L->num_params, 1);
  ATH_TRY(L->own_params.reserve(bytes));
  ATH_TRY(L->own_grads.reserve(bytes));
  ATH_CUDA(cudaMemsetAsync(L->own_params.p, 0, bytes, ctx().stream));
  ATH_CUDA(cudaMemsetAsync(L->own_grads.p, 0, bytes, ctx().stream));
  L->params = L->own_params.as<float>();
  L->grads = L->own_grads.as<float>();
  return ATHENA_OK;
}

// ---- forward ---------------------------------------------------------------------

// Network-level fusion of the last Kipf step with the graph-output MSE (train step only):
// when the fused kernel applies, the layer output is never materialised; `grad` receives
// d loss / d pre-activation and `loss_part[0..*num_parts)` the per-CTA loss sums.
struct FwdOpts {
  cudaEvent_t target_ready = nullptr;  // the target's H2D copy (copy stream), if any
  const float* mse_target = nullptr;
  float* mse_grad = nullptr;
  float* loss_part = nullptr;
  int num_parts = 0;
  bool fused = false;
  float mse_denom = 0.f;  // Duvenaud-last: num_outputs * global batch (one [no, batch] cell)
  bool in_network = false;  // the network walk owns Context::tile_reverse
};

static int kipf_forward(Layer* L, Batch* b, const float* x, const float** out,
                        FwdOpts* fo = nullptr) {
  const int64_t V = b->V;
  const float* in = x;
  // swish differentiates on the pre-activation (get_partial_swish_val,
  // athena_diffstruc_extd_sub.f90:472-486): the step kernels run without activation into Y_t,
  // which is kept, and H_t = swish(Y_t) is one elementwise pass
  const bool swish = L->act == ATHENA_ACT_SWISH;
  const int kact = swish ? ATHENA_ACT_NONE : L->act;
  for (int t = 1; t <= L->T; ++t) {
    const int Fi = L->nvf[t - 1], Fo = L->nvf[t];
    DevBuf& P = *L->P[t - 1];
    DevBuf& H = *L->H[t - 1];
    ATH_TRY(P.reserve(sizeof(float) * (size_t)std::max<int64_t>(V * Fi, 1)));
    ATH_TRY(H.reserve(sizeof(float) * (size_t)std::max<int64_t>(V * Fo, 1)));
    float* Hk = H.as<float>();  // where the step kernel writes
    if (swish) {
      ATH_TRY(L->Y[t - 1]->reserve(sizeof(float) * (size_t)std::max<int64_t>(V * Fo, 1)));
      Hk = L->Y[t - 1]->as<float>();
    }
    auto step_done = [&]() -> int {
      if (swish) ATH_TRY(launch_swish_fwd(Hk, H.as<float>(), V * Fo));
      in = H.as<float>();
      return ATHENA_OK;
    };
    L->mask_valid[t - 1] = 0;
    if (fo && fo->mse_target && t == L->T && L->act != ATHENA_ACT_SOFTMAX && !swish &&
        pipe_gather_supported(b, Fi, Fo)) {
      ATH_TRY(main_wait(fo->target_ready));
      fo->target_ready = nullptr;
      ATH_TRY(launch_pipe_gather_fwd_mse(b, in, L->params + L->poff[t - 1], P.as<float>(),
                                         fo->mse_target, fo->mse_grad, Fi, Fo, L->act,
                                         fo->loss_part, &fo->num_parts));
      fo->fused = true;
      in = nullptr;  // H_T is not materialised
      continue;
    }
    if (L->act != ATHENA_ACT_SOFTMAX && pipe_gather_supported(b, Fi, Fo)) {
      // propagate + transform + activation in one fused tcgen05 kernel; for relu-type
      // activations it also records the sign bits the reverse sweep needs
      uint32_t* mk = nullptr;
      if (!L->inference && (L->act == ATHENA_ACT_RELU || L->act == ATHENA_ACT_LEAKY_RELU)) {
        ATH_TRY(L->MK[t - 1]->reserve(sizeof(uint32_t) * (size_t)std::max<int64_t>(V * (Fo / 32), 1)));
        mk = L->MK[t - 1]->as<uint32_t>();
        L->mask_valid[t - 1] = 1;
      }
      ATH_TRY(launch_pipe_gather_fwd(b, in, L->params + L->poff[t - 1], P.as<float>(), Hk, Fi, Fo,
                                     kact, mk));
      ATH_TRY(step_done());
      continue;
    }
    if (tile_kipf_supported(b, Fi, Fo)) {
      // small graphs, narrow features: propagate + transform + activation in one FP32 tile kernel
      ATH_TRY(launch_tile_kipf_fwd(b, in, L->params + L->poff[t - 1],
                                   L->inference ? nullptr : P.as<float>(), Hk, Fi, Fo, kact));
      ATH_TRY(step_done());
      continue;
    }
    if (L->act != ATHENA_ACT_SOFTMAX && agg_tc_supported(Fi, Fo, in, Hk)) {
      // large graph, wide features: SpMM + tcgen05 transform in one pass; the aggregate
      // is only written when a reverse sweep may follow
      ATH_TRY(launch_agg_tc_fwd(b, in, L->params + L->poff[t - 1],
                                L->inference ? nullptr : P.as<float>(), Hk, kact));
      ATH_TRY(step_done());
      continue;
    }
    // P = D^-1/2 A D^-1/2 . in      (kipf_propagate)
    ATH_TRY(launch_aggregate(b->row_ptr, b->col, b->coef, in, Fi, Fi, P.as<float>(), Fi, V, 0,
                             nullptr, 0, b->long_rows, b->long_counts));
    // H = act( P . W_t )            (matmul + activation%apply)
    if (L->act != ATHENA_ACT_SOFTMAX && tc_rows_supported(Fi, Fo, Fi, Fo, P.as<float>(), Hk)) {
      ATH_TRY(launch_tc_rows(false, P.as<float>(), Fi, nullptr, 0, L->params + L->poff[t - 1], Hk,
                             Fo, V, Fo, Fi, kact));
    } else {
      ATH_TRY(launch_gemm_nn(P.as<float>(), Fi, L->params + L->poff[t - 1], Hk, Fo, V, Fo, Fi, kact,
                             GroupDesc{}));
    }
    ATH_TRY(step_done());
  }
  *out = in;
  return ATHENA_OK;
}

static void duv_desc(Layer* L, const float* x, const float* e, float* const* Z,
                     TileDuvDesc* d) {
  d->T = L->T;
  d->nef = L->nef;
  d->min_deg = L->min_deg;
  d->max_deg = L->max_deg;
  d->no = L->n_out;
  d->act = L->act;
  d->ract = L->ract;
  d->nvf = L->nvf.data();
  d->poff = L->poff.data();
  d->params = L->params;
  d->X = x;
  d->E = e;
  d->Ae = L->nef > 0 ? L->Ae.as<float>() : nullptr;
  d->Z = Z;
}

static int duvenaud_forward(Layer* L, Batch* b, const float* x, const float* e,
                            const float** out, FwdOpts* fo = nullptr) {
  const int64_t V = b->V;
  L->tile_fwd = false;
  L->fwd_e = e;
  const bool swish = L->act == ATHENA_ACT_SWISH;
  if (!swish &&
      tile_duv_supported(b, L->T, L->nvf.data(), L->nef, L->max_deg - L->min_deg + 1, L->n_out)) {
    // every time step and the readout of the whole layer in ONE launch (tile_fma.cu); z_t, the
    // readouts S_t and the per-vertex edge-feature sums are kept for the reverse sweep
    ATH_REQUIRE(L->nef == 0 || e != nullptr, ATHENA_ERR_ARG,
                "duvenaud forward: edge_features is null");
    float* Z[16];
    float* S[16];
    for (int t = 1; t <= L->T; ++t) {
      DevBuf& Zt = *L->H[t - 1];
      ATH_TRY(Zt.reserve(sizeof(float) * (size_t)std::max<int64_t>(V * L->nvf[t], 1)));
      Z[t - 1] = Zt.as<float>();
      // the readouts S_t are kept for the reverse sweep (not in inference mode)
      S[t - 1] = nullptr;
      if (!L->inference) {
        DevBuf& St = *L->S[t - 1];
        ATH_TRY(St.reserve(sizeof(float) * (size_t)std::max<int64_t>(V * L->n_out, 1)));
        S[t - 1] = St.as<float>();
      }
    }
    L->tile_S = !L->inference;
    ATH_TRY(L->out_buf.reserve(sizeof(float) * (size_t)b->B * L->n_out));
    if (L->nef > 0)
      ATH_TRY(L->Ae.reserve(sizeof(float) * (size_t)std::max<int64_t>(V * L->nef, 1)));
    TileDuvDesc d;
    duv_desc(L, x, e, Z, &d);
    if (L->tile_S) d.S = S;
    const bool mse = fo != nullptr && fo->mse_target != nullptr;
    if (mse) {
      ATH_TRY(main_wait(fo->target_ready));
      fo->target_ready = nullptr;
    }
    ATH_TRY(launch_tile_duv_fwd(b, d, L->out_buf.as<float>(), mse ? fo->mse_target : nullptr,
                                mse ? fo->mse_grad : nullptr, mse ? fo->mse_denom : 1.f,
                                mse ? fo->loss_part : nullptr, mse ? &fo->num_parts : nullptr));
    if (mse) fo->fused = true;
    L->tile_fwd = true;
    *out = L->out_buf.as<float>();
    return ATHENA_OK;
  }
  BucketSet* bs = nullptr;
  ATH_TRY(batch_bucketize(b, L->min_deg, L->max_deg, &bs));
  const int D = bs->D;
  const float* tail = nullptr;
  if (L->nef > 0) {
    ATH_REQUIRE(e != nullptr, ATHENA_ERR_ARG, "duvenaud forward: edge_features is null");
    ATH_TRY(L->Ae.reserve(sizeof(float) * (size_t)std::max<int64_t>(V * L->nef, 1)));
    // time-step invariant: sum_w E(:, ja(2,w))
    ATH_TRY(launch_aggregate(b->row_ptr, b->eid, nullptr, e, L->nef, L->nef, L->Ae.as<float>(),
                             L->nef, V, 0, nullptr, 0));
    tail = L->Ae.as<float>();
  }
  const float* in = x;
  for (int t = 1; t <= L->T; ++t) {
    const int Fi = L->nvf[t - 1], Fo = L->nvf[t], K = Fi + L->nef, ld = L->ldA(t);
    DevBuf& A = *L->P[t - 1];
    DevBuf& Zt = *L->H[t - 1];
    ATH_TRY(A.reserve(sizeof(float) * (size_t)std::max<int64_t>(V * ld, 1)));
    ATH_TRY(Zt.reserve(sizeof(float) * (size_t)std::max<int64_t>(V * Fo, 1)));
    // A = [ sum_w in(:,ja(1,w)) ; sum_w E(:,ja(2,w)) ]      (duvenaud_propagate)
    ATH_TRY(launch_aggregate(b->row_ptr, b->col, nullptr, in, Fi, Fi, A.as<float>(), ld, V, 0, tail,
                             L->nef, b->long_rows, b->long_counts));
    // z = act( W_d(v) . A(:,v)/d(v) )                       (duvenaud_update + activation)
    GroupDesc gd;
    gd.perm = bs->perm.as<int32_t>();
    gd.ptr = bs->bkt_ptr.as<int32_t>();
    gd.D = D;
    gd.wstride = (int64_t)Fo * K;
    gd.scale_by_group = 1;
    if (swish) {
      DevBuf& Yt = *L->Y[t - 1];
      ATH_TRY(Yt.reserve(sizeof(float) * (size_t)std::max<int64_t>(V * Fo, 1)));
      ATH_TRY(launch_gemm_nn(A.as<float>(), ld, L->params + L->poff[t - 1], Yt.as<float>(), Fo, V,
                             Fo, K, ATHENA_ACT_NONE, gd));
      ATH_TRY(launch_swish_fwd(Yt.as<float>(), Zt.as<float>(), V * Fo));
    } else {
      ATH_TRY(launch_gemm_nn(A.as<float>(), ld, L->params + L->poff[t - 1], Zt.as<float>(), Fo, V,
                             Fo, K, L->act, gd));
    }
    in = Zt.as<float>();
  }
  // readout: out(:,s) = sum_t sum_v ract( R_t . z_t )(:,v)
  const int no = L->n_out;
  ATH_TRY(L->out_buf.reserve(sizeof(float) * (size_t)b->B * no));
  for (int t = 1; t <= L->T; ++t) {
    const int Fo = L->nvf[t];
    DevBuf& St = *L->S[t - 1];
    ATH_TRY(St.reserve(sizeof(float) * (size_t)std::max<int64_t>(V * no, 1)));
    ATH_TRY(launch_gemm_nn(L->H[t - 1]->as<float>(), Fo, L->params + L->poff[L->T + t - 1],
                           St.as<float>(), no, V, no, Fo, L->ract, GroupDesc{}));
    ATH_TRY(launch_segment_sum(St.as<float>(), no, b->voff, b->B, L->out_buf.as<float>(), t > 1));
  }
  *out = L->out_buf.as<float>();
  return ATHENA_OK;
}

// full_layer_type%forward (athena_full_layer.f90:839-874) on the [batch][num_inputs] output
// of a graph-level layer: y = act( x . W + b ).  The dense head of example/msgpass_chemical
// (main.f90:139-157); tiny, latency-bound work on the generic kernels.
static int full_forward(Layer* L, Batch* b, const float* x, const float** out) {
  const int Ni = L->nvf[0], No = L->nvf[1];
  const int64_t B = b->B;
  DevBuf& H = *L->H[0];
  ATH_TRY(H.reserve(sizeof(float) * (size_t)std::max<int64_t>(B * No, 1)));
  ATH_TRY(launch_gemm_nn(x, Ni, L->params + L->poff[0], H.as<float>(), No, B, No, Ni,
                         ATHENA_ACT_NONE, GroupDesc{}));
  if (L->act == ATHENA_ACT_SWISH) {
    DevBuf& Y = *L->Y[0];
    ATH_TRY(Y.reserve(sizeof(float) * (size_t)std::max<int64_t>(B * No, 1)));
    ATH_TRY(launch_bias_act(H.as<float>(), L->use_bias ? L->params + L->poff[1] : nullptr, B, No,
                            ATHENA_ACT_NONE));
    ATH_CUDA(cudaMemcpyAsync(Y.p, H.p, sizeof(float) * (size_t)(B * No), cudaMemcpyDeviceToDevice,
                             ctx().stream));
    ATH_TRY(launch_swish_fwd(Y.as<float>(), H.as<float>(), B * No));
  } else {
    ATH_TRY(launch_bias_act(H.as<float>(), L->use_bias ? L->params + L->poff[1] : nullptr, B, No,
                            L->act));
  }
  *out = H.as<float>();
  return ATHENA_OK;
}

static int full_backward(Layer* L, Batch* b, const float* gout, float* gin) {
  const int Ni = L->nvf[0], No = L->nvf[1];
  const int64_t B = b->B;
  ATH_TRY(L->g0.reserve(sizeof(float) * (size_t)std::max<int64_t>(B * No, 1)));
  const float* gz = gout;
  if (L->act == ATHENA_ACT_SWISH) {
    ATH_TRY(launch_swish_bwd(L->Y[0]->as<float>(), gout, L->g0.as<float>(), B * No));
    gz = L->g0.as<float>();
  } else if (L->act != ATHENA_ACT_NONE && L->act != ATHENA_ACT_LINEAR) {
    ATH_TRY(launch_act_bwd(L->act, L->H[0]->as<float>(), gout, L->g0.as<float>(), B, No));
    gz = L->g0.as<float>();
  }
  // dW(o,i) += sum_s gz(o,s) x(i,s) ; db(o) += sum_s gz(o,s) ; dx = W^T gz
  ATH_TRY(launch_gemm_tn(L->fwd_x, Ni, gz, No, L->grads + L->poff[0], B, No, Ni, GroupDesc{},
                         L->tn_scratch));
  if (L->use_bias) ATH_TRY(launch_colsum_add(gz, B, No, L->grads + L->poff[1]));
  if (gin)
    ATH_TRY(launch_gemm_nt(gz, No, L->params + L->poff[0], gin, Ni, B, Ni, No, GroupDesc{}));
  return ATHENA_OK;
}

int layer_forward_dev(Layer* L, Batch* b, const float* x, const float* e, const float** out,
                      FwdOpts* fo = nullptr) {
  if (fo == nullptr || !fo->in_network) ctx().tile_reverse = false;  // per-call determinism
  ATH_REQUIRE(x != nullptr || b->V == 0, ATHENA_ERR_ARG, "forward: vertex_features is null");
  L->fwd_batch = b;
  L->fwd_V = b->V;
  L->fwd_x = x;
  if (L->kind == 2) return full_forward(L, b, x, out);
  if (L->kind == 0) return kipf_forward(L, b, x, out, fo);
  return duvenaud_forward(L, b, x, e, out, fo);
}

// ---- backward --------------------------------------------------------------------

// Layer-boundary fusion of the reverse sweep:
//   gout_is_preact  gout already carries act'(H_T) of THIS layer (folded into the loss
//                   gradient or into the next layer's epilogue)
//   fold_act        activation of the layer that produced this layer's input; when the
//                   fused kernel is used, grad_input is returned already multiplied by
//                   fold_act'(input) and *folded is set
struct BwdOpts {
  bool gout_is_preact = false;
  int fold_act = ATHENA_ACT_NONE;
  const uint32_t* fold_mask = nullptr;  // sign bits of the producing layer's output, if recorded
  bool* folded = nullptr;
  DeferList* defer = nullptr;  // queue the dW partial folds for launch_finalize
};

static int kipf_backward(Layer* L, Batch* b, const float* gout, float* gin, const BwdOpts& opt) {
  const int64_t V = b->V;
  int Fmax = 0;
  for (int t = 0; t <= L->T; ++t) Fmax = std::max(Fmax, L->nvf[t]);
  size_t bytes = sizeof(float) * (size_t)std::max<int64_t>(V * Fmax, 1);
  ATH_TRY(L->g0.reserve(bytes));
  ATH_TRY(L->g1.reserve(bytes));
  ATH_TRY(L->g2.reserve(bytes));
  const bool swish = L->act == ATHENA_ACT_SWISH;
  const bool nonlinear = L->act != ATHENA_ACT_NONE && L->act != ATHENA_ACT_LINEAR && !swish;
  const float* g = gout;                // gradient w.r.t. the step output H_t ...
  bool preact = opt.gout_is_preact;     // ... or already w.r.t. the pre-activation (gY_t)
  if (opt.folded) *opt.folded = false;
  for (int t = L->T; t >= 1; --t) {
    const int Fi = L->nvf[t - 1], Fo = L->nvf[t];
    if (swish && !preact) {
      // gY_t = gH_t * swish'(Y_t) on the saved pre-activation; the fused kernels below then
      // see a layer without activation
      float* dst = (g == L->g0.as<float>()) ? L->g2.as<float>() : L->g0.as<float>();
      ATH_TRY(launch_swish_bwd(L->Y[t - 1]->as<float>(), g, dst, V * Fo));
      g = dst;
    }
    if (swish) preact = false;  // nonlinear == false: g is used as gY_t as it stands
    const float* Pt = L->P[t - 1]->as<float>();
    const float* Ht = L->H[t - 1]->as<float>();
    const float* Wt = L->params + L->poff[t - 1];
    float* dWt = L->grads + L->poff[t - 1];
    const bool need_dp = (t > 1 || gin);
    const bool need_act = nonlinear && !preact;
    // the fused tcgen05 reverse step at width 32 takes act' from sign bits only
    const bool pipe_bwd_ok =
        pipe_gather_supported(b, Fo, Fi) &&
        (Fi == 64 || t == 1 || !nonlinear || L->mask_valid[t - 2] != 0);
    if (tile_kipf_supported(b, Fi, Fo) && !pipe_bwd_ok) {
      // act', dW partials, dP = gY W^T and the un-normalised CSC scatter in one FP32 tile kernel
      float* dst = gin;
      if (t > 1) dst = (g == L->g2.as<float>()) ? L->g0.as<float>() : L->g2.as<float>();
      int nparts = 0;
      ATH_TRY(launch_tile_kipf_bwd(b, g, need_act ? Ht : nullptr, Pt, Wt, need_dp ? dst : nullptr,
                                   Fi, Fo, L->act, *L->TN[t - 1], &nparts));
      DeferJob job{L->TN[t - 1]->as<float>(), nparts, Fi * Fo, dWt};
      if (opt.defer) {
        opt.defer->jobs.push_back(job);
      } else {
        DeferList dl;
        dl.jobs.push_back(job);
        ATH_TRY(launch_finalize(dl, nullptr, 0, nullptr, nullptr, dWt, (int64_t)Fi * Fo, nullptr));
      }
      g = dst;
      preact = false;
      continue;
    }
    // fused path: one tcgen05 kernel gathers gY_t over the CSC, multiplies by W_t^T and applies
    // act'(H_{t-1}); the un-normalised scatter commutes with the linear map
    const bool fused_dp = need_dp && L->act != ATHENA_ACT_SOFTMAX && pipe_bwd_ok;
    const bool fused_tn = pipe_tn_supported(Fi, Fo);
    const bool tn_tc = !fused_tn && tc_tn_supported(Fi, Fo, Fi, Fo, Pt, g);
    const bool nt_tc = need_dp && !fused_dp && tc_rows_supported(Fo, Fi, Fo, Fi, g, L->g1.as<float>());
    // the unfused tensor-core kernels can apply act'(H) while loading gH
    const bool fuse_act = need_act && L->act != ATHENA_ACT_SOFTMAX && !fused_dp && !fused_tn &&
                          tn_tc && (!need_dp || nt_tc);
    const float* Hact = fuse_act ? Ht : nullptr;
    const float* gy = g;
    if (need_act && !fuse_act) {
      float* dst = (g == L->g0.as<float>()) ? L->g2.as<float>() : L->g0.as<float>();
      ATH_TRY(launch_act_bwd(L->act, Ht, g, dst, V, Fo));
      gy = dst;
    }
    // dW_t(o,i) += sum_v gY(o,v) P(i,v)
    if (fused_tn) {
      // a one-step layer never reuses its gradient buffers inside the sweep: the product can
      // join the batched launch at the end of the sweep
      ATH_TRY(launch_pipe_tn(Pt, gy, dWt, V, Fi, Fo, *L->TN[t - 1], opt.defer, L->T == 1));
    } else if (tn_tc) {
      // per-step partial buffer: with a DeferList the fold waits for the finalize launch
      ATH_TRY(launch_tc_tn(Pt, Fi, gy, Fo, Hact, L->act, dWt, V, Fo, Fi, *L->TN[t - 1], opt.defer));
    } else {
      ATH_TRY(launch_gemm_tn(Pt, Fi, gy, Fo, dWt, V, Fo, Fi, GroupDesc{}, L->tn_scratch));
    }
    if (!need_dp) continue;
    float* dst = gin;
    if (t > 1) dst = (gy == L->g2.as<float>()) ? L->g0.as<float>() : L->g2.as<float>();
    if (fused_dp) {
      const float* Hin = nullptr;
      const uint32_t* mk = nullptr;
      int act_e = ATHENA_ACT_NONE;
      if (t > 1 && nonlinear) {
        Hin = L->H[t - 2]->as<float>();
        if (L->mask_valid[t - 2]) mk = L->MK[t - 2]->as<uint32_t>();
        act_e = L->act;
      } else if (t == 1 && opt.fold_act != ATHENA_ACT_NONE && opt.fold_act != ATHENA_ACT_LINEAR &&
                 opt.fold_act != ATHENA_ACT_SOFTMAX && L->fwd_x != nullptr &&
                 (Fi == 64 || opt.fold_mask != nullptr)) {
        Hin = L->fwd_x;  // the producing layer's output
        mk = opt.fold_mask;
        act_e = opt.fold_act;
        if (opt.folded) *opt.folded = true;
      }
      ATH_TRY(launch_pipe_gather_bwd(b, gy, Wt, Hin, dst, Fo, Fi, act_e, mk));
      preact = t > 1 && !swish;  // dst already is gY_{t-1}
    } else {
      // dP = W_t^T gY, then dH(:,u) += dP(:,v) for every CSR entry (v,u): CSC gather, NO coefficient
      if (nt_tc) {
        ATH_TRY(launch_tc_rows(true, gy, Fo, Hact, L->act, Wt, L->g1.as<float>(), Fi, V, Fi, Fo,
                               ATHENA_ACT_NONE));
      } else {
        ATH_TRY(launch_gemm_nt(gy, Fo, Wt, L->g1.as<float>(), Fi, V, Fi, Fo, GroupDesc{}));
      }
      ATH_TRY(launch_aggregate(b->csc_ptr, b->csc_src, nullptr, L->g1.as<float>(), Fi, Fi, dst, Fi,
                               V, 0, nullptr, 0, b->long_cols, b->long_counts ? b->long_counts + 1 : nullptr));
      preact = false;
    }
    g = dst;
  }
  return ATHENA_OK;
}

static int duvenaud_backward(Layer* L, Batch* b, const float* gout, float* gin,
                             DeferList* defer = nullptr, int fold_act = ATHENA_ACT_NONE,
                             bool* folded = nullptr) {
  const int64_t V = b->V;
  if (L->tile_fwd) {
    // the whole reverse sweep of the layer in ONE launch; A_t and S_t are recomputed from z_t
    float* Z[16];
    float* S[16];
    for (int t = 1; t <= L->T; ++t) {
      Z[t - 1] = L->H[t - 1]->as<float>();
      S[t - 1] = L->tile_S ? L->S[t - 1]->as<float>() : nullptr;
    }
    TileDuvDesc d;
    duv_desc(L, L->fwd_x, L->fwd_e, Z, &d);
    if (L->tile_S) d.S = S;
    if (gin != nullptr && folded != nullptr && fold_act != ATHENA_ACT_NONE &&
        fold_act != ATHENA_ACT_LINEAR && fold_act != ATHENA_ACT_SOFTMAX &&
        fold_act != ATHENA_ACT_SWISH) {
      // the input gradient leaves as the gradient w.r.t. the producing layer's pre-activation
      // (its output is this layer's input X): no activation-derivative launch in between
      d.fold_act = fold_act;
      *folded = true;
    }
    int nparts = 0;
    ATH_TRY(launch_tile_duv_bwd(b, d, gout, gin, L->num_params, L->tile_part, &nparts));
    DeferJob job{L->tile_part.as<float>(), nparts, (int)L->num_params, L->grads};
    if (defer) {
      defer->jobs.push_back(job);
    } else {
      DeferList dl;
      dl.jobs.push_back(job);
      ATH_TRY(launch_finalize(dl, nullptr, 0, nullptr, nullptr, L->grads, L->num_params, nullptr));
    }
    return ATHENA_OK;
  }
  BucketSet* bs = nullptr;
  ATH_TRY(batch_bucketize(b, L->min_deg, L->max_deg, &bs));
  const int no = L->n_out, T = L->T;
  int Wmax = no;
  for (int t = 1; t <= T; ++t) Wmax = std::max(Wmax, std::max(L->nvf[t], L->ldA(t)));
  size_t bytes = sizeof(float) * (size_t)std::max<int64_t>(V * Wmax, 1);
  ATH_TRY(L->g0.reserve(bytes));
  ATH_TRY(L->g1.reserve(bytes));
  // readout share of d z_t
  for (int t = 1; t <= T; ++t) {
    const int Fo = L->nvf[t];
    const float* R = L->params + L->poff[T + t - 1];
    DevBuf& GZ = *L->GZ[t - 1];
    ATH_TRY(GZ.reserve(sizeof(float) * (size_t)std::max<int64_t>(V * Fo, 1)));
    ATH_TRY(launch_readout_bwd(L->ract, L->S[t - 1]->as<float>(), gout, b->vgraph,
                               L->g0.as<float>(), V, no));
    ATH_TRY(launch_gemm_tn(L->H[t - 1]->as<float>(), Fo, L->g0.as<float>(), no,
                           L->grads + L->poff[T + t - 1], V, no, Fo, GroupDesc{}, L->tn_scratch));
    ATH_TRY(launch_gemm_nt(L->g0.as<float>(), no, R, GZ.as<float>(), Fo, V, Fo, no, GroupDesc{}));
  }
  for (int t = T; t >= 1; --t) {
    const int Fi = L->nvf[t - 1], Fo = L->nvf[t], K = Fi + L->nef, ld = L->ldA(t);
    const float* gz = L->GZ[t - 1]->as<float>();
    const float* gzp = gz;
    if (L->act == ATHENA_ACT_SWISH) {
      ATH_TRY(launch_swish_bwd(L->Y[t - 1]->as<float>(), gz, L->g0.as<float>(), V * Fo));
      gzp = L->g0.as<float>();
    } else if (L->act != ATHENA_ACT_NONE && L->act != ATHENA_ACT_LINEAR) {
      ATH_TRY(launch_act_bwd(L->act, L->H[t - 1]->as<float>(), gz, L->g0.as<float>(), V, Fo));
      gzp = L->g0.as<float>();
    }
    GroupDesc gd;
    gd.perm = bs->perm.as<int32_t>();
    gd.ptr = bs->bkt_ptr.as<int32_t>();
    gd.D = bs->D;
    gd.wstride = (int64_t)Fo * K;
    gd.scale_by_group = 1;
    // dW_d(i,j) += gZ(i,v) A(j,v)/d
    ATH_TRY(launch_gemm_tn(L->P[t - 1]->as<float>(), ld, gzp, Fo, L->grads + L->poff[t - 1], V, Fo,
                           K, gd, L->tn_scratch));
    if (t > 1 || gin) {
      // dA(:,v) = (W_d^T gZ(:,v)) / d
      ATH_TRY(launch_gemm_nt(gzp, Fo, L->params + L->poff[t - 1], L->g1.as<float>(), ld, V, K, Fo,
                             gd));
      // d in(:,u) += dA(1:F,v) over the CSC
      if (t > 1)
        ATH_TRY(launch_aggregate(b->csc_ptr, b->csc_src, nullptr, L->g1.as<float>(), ld, Fi,
                                 L->GZ[t - 2]->as<float>(), Fi, V, 1, nullptr, 0, b->long_cols,
                                 b->long_counts ? b->long_counts + 1 : nullptr));
      else
        ATH_TRY(launch_aggregate(b->csc_ptr, b->csc_src, nullptr, L->g1.as<float>(), ld, Fi, gin,
                                 Fi, V, 0, nullptr, 0, b->long_cols,
                                 b->long_counts ? b->long_counts + 1 : nullptr));
    }
  }
  return ATHENA_OK;
}

int layer_backward_dev(Layer* L, Batch* b, const float* gout, float* gin,
                       const BwdOpts& opt = BwdOpts{}) {
  ATH_REQUIRE(L->fwd_batch == b && L->fwd_V == b->V, ATHENA_ERR_STATE,
              "backward: no forward pass on this batch");
  ATH_REQUIRE(gout != nullptr, ATHENA_ERR_ARG, "backward: grad_output is null");
  if (opt.folded) *opt.folded = false;
  if (L->kind == 2) return full_backward(L, b, gout, gin);
  if (L->kind == 0) return kipf_backward(L, b, gout, gin, opt);
  ATH_REQUIRE(!opt.gout_is_preact, ATHENA_ERR_STATE, "duvenaud backward: unexpected pre-activation gradient");
  return duvenaud_backward(L, b, gout, gin, opt.defer, opt.fold_act, opt.folded);
}

// ---- network ---------------------------------------------------------------------

struct Network : Object {
  Network() : Object(Kind::Network) {}
  std::vector<athena_handle_t> handles;
  std::vector<Layer*> layers;
  bool compiled = false;
  int64_t n = 0;
  DevBuf flat_params, flat_grads;  // grads has n + 1 (+pad) floats: [n] = batch loss
  OptimState opt;
  DevBuf stage_x, stage_e, stage_t, gbuf, loss_scratch;
  std::vector<std::unique_ptr<DevBuf>> gin;  // input gradient of layer l (l >= 1)
  // network%add(layer, input_list, operator = 'concatenate') (athena_network_sub.f90:764-830):
  // sources of layer l, in list order: -1 the network input, k >= 0 layer k; empty = the plain
  // chain (layer l - 1, or the network input for l = 0)
  std::vector<std::vector<int>> inputs;
  std::vector<int> consumers;                 // how many layers read layer l's output
  std::vector<std::unique_ptr<DevBuf>> cat;   // concatenated input of layer l
  std::vector<std::unique_ptr<DevBuf>> gsum;  // summed output gradient of layer l (skip links)
  bool has_skips = false;
  float* pinned_loss = nullptr;
  int edge_width() const {  // edge features are consumed by the Duvenaud layer (if any)
    for (const Layer* L : layers)
      if (L->kind == 1) return L->nef;
    return 0;
  }
  ~Network() {
    if (pinned_loss) cudaFreeHost(pinned_loss);
  }
};

static int stage_in(DevBuf& buf, const float* src, int64_t count, int mem, const float** out) {
  if (src == nullptr || mem == ATHENA_MEM_DEVICE) {
    *out = src;
    return ATHENA_OK;
  }
  ATH_TRY(buf.reserve(sizeof(float) * (size_t)std::max<int64_t>(count, 1)));
  if (count > 0)
    ATH_CUDA(cudaMemcpyAsync(buf.p, src, sizeof(float) * (size_t)count, cudaMemcpyHostToDevice,
                             ctx().stream));
  *out = buf.as<float>();
  return ATHENA_OK;
}

// host -> device staging on the copy stream (no-op for device / null inputs)
static int stage_in_side(DevBuf& buf, const float* src, int64_t count, int mem,
                         const float** out) {
  if (src == nullptr || mem == ATHENA_MEM_DEVICE) {
    *out = src;
    return ATHENA_OK;
  }
  ATH_TRY(buf.reserve(sizeof(float) * (size_t)std::max<int64_t>(count, 1)));
  if (count > 0) ATH_TRY(side_copy(buf.p, src, sizeof(float) * (size_t)count));
  *out = buf.as<float>();
  return ATHENA_OK;
}

static int src_width(const Network* N, int src) {
  const Layer* S = src < 0 ? N->layers.front() : N->layers[src];
  return src < 0 ? S->nvf[0] : S->nvf[S->T];
}

static int net_forward_dev(Network* N, Batch* b, const float* x, const float* e,
                           const float** out, FwdOpts* fo = nullptr) {
  const float* in = x;
  std::vector<const float*> outs(N->layers.size(), nullptr);
  ctx().tile_reverse = false;  // the walk alternates from here: results reproducible per call
  FwdOpts inner;
  inner.in_network = true;
  if (fo) fo->in_network = true;
  for (size_t l = 0; l < N->layers.size(); ++l) {
    Layer* L = N->layers[l];
    if (N->has_skips && !N->inputs[l].empty()) {
      // concat_layer_type%combine (athena_concat_layer.f90:413-456): the sources' vertex
      // features side by side, in list order
      const int Fin = L->nvf[0];
      DevBuf& cat = *N->cat[l];
      ATH_TRY(cat.reserve(sizeof(float) * (size_t)std::max<int64_t>(b->V * Fin, 1)));
      int off = 0;
      for (int src : N->inputs[l]) {
        const int w = src_width(N, src);
        ATH_TRY(launch_copy_cols(cat.as<float>(), Fin, off, src < 0 ? x : outs[src], w, 0, w, b->V, 0));
        off += w;
      }
      in = cat.as<float>();
    }
    const float* o = nullptr;
    ATH_TRY(layer_forward_dev(L, b, in, e, &o, (l + 1 == N->layers.size() && fo) ? fo : &inner));
    outs[l] = o;
    in = o;
  }
  *out = in;
  return ATHENA_OK;
}

// step_too: also perform network%update when nothing (gradient exchange, clipping) has to
// happen between the reverse sweep and the step; *stepped reports whether it did.
static int net_loss_grads(Network* N, Batch* b, const float* x, const float* e, const float* tgt,
                          int mem, int global_batch, float* loss, bool step_too = false,
                          bool* stepped = nullptr) {
  if (stepped) *stepped = false;
  ATH_REQUIRE(N->compiled, ATHENA_ERR_STATE, "network is not compiled");
  ATH_REQUIRE(tgt != nullptr, ATHENA_ERR_ARG, "train: target is null");
  Layer* first = N->layers.front();
  Layer* last = N->layers.back();
  const float *dx = nullptr, *de = nullptr, *dt = nullptr, *out = nullptr;
  // inputs on the copy stream: the features first (the first layer waits for them), then the
  // target, which only the loss needs -- the layers before it run under its copy
  cudaEvent_t ev_in = nullptr, ev_tgt = nullptr;
  const int64_t out_n = last->out_rows(b) * last->out_width();
  if (mem == ATHENA_MEM_HOST) {
    ATH_TRY(side_begin());
    ATH_TRY(stage_in_side(N->stage_x, x, b->V * first->nvf[0], mem, &dx));
    ATH_TRY(stage_in_side(N->stage_e, e, b->E * N->edge_width(), mem, &de));
    ATH_TRY(side_fence(&ev_in));
    ATH_TRY(stage_in_side(N->stage_t, tgt, out_n, mem, &dt));
    ATH_TRY(side_fence(&ev_tgt));
    ATH_TRY(main_wait(ev_in));
  } else {
    dx = x;
    de = e;
    dt = tgt;
  }
  float* gflat = N->flat_grads.as<float>();
  cudaStream_t st = ctx().stream;
  ATH_CUDA(cudaMemsetAsync(gflat + N->n, 0, sizeof(float), st));
  ATH_TRY(N->gbuf.reserve(sizeof(float) * (size_t)std::max<int64_t>(out_n, 1)));
  ATH_TRY(N->loss_scratch.reserve(sizeof(float) * 1024));
  FwdOpts fo;
  fo.target_ready = ev_tgt;
  if ((last->kind == 0 || last->kind == 1) && b->V > 0) {
    fo.mse_target = dt;
    fo.mse_grad = N->gbuf.as<float>();
    fo.loss_part = N->loss_scratch.as<float>();
    const int gb = global_batch > 0 ? global_batch : b->B;
    fo.mse_denom = (float)((int64_t)last->out_width() * gb);
  }
  ATH_TRY(net_forward_dev(N, b, dx, de, &out, &fo));
  ATH_TRY(main_wait(fo.target_ready));  // unfused loss: wait here
  // the activation derivative of the last Kipf layer is folded into the loss gradient
  auto foldable = [](const Layer* L) {
    return L->kind == 0 && L->act != ATHENA_ACT_NONE && L->act != ATHENA_ACT_LINEAR &&
           L->act != ATHENA_ACT_SOFTMAX && L->act != ATHENA_ACT_SWISH;
  };
  bool g_preact = false;
  DeferList defer;
  if (fo.fused) {
    // the last layer's kernel already produced the loss gradient (Kipf: w.r.t. the
    // pre-activation; Duvenaud: w.r.t. the [num_outputs, batch] output) and the per-CTA loss
    // sums, which launch_finalize folds into the loss slot
    g_preact = last->kind == 0;
  } else if (last->kind == 0) {
    g_preact = foldable(last);
    ATH_TRY(launch_mse_graph(out, dt, b->vgraph, b->nv, last->nvf[last->T], b->V,
                             g_preact ? last->act : ATHENA_ACT_NONE, N->gbuf.as<float>(),
                             gflat + N->n, N->loss_scratch));
  } else {
    int gb = global_batch > 0 ? global_batch : b->B;
    ATH_TRY(launch_mse_array(out, dt, out_n, (float)((int64_t)last->out_width() * gb),
                             N->gbuf.as<float>(), gflat + N->n, N->loss_scratch));
  }
  const float* g = N->gbuf.as<float>();
  const int nl = (int)N->layers.size();
  // with skip links the gradient of a layer output is the sum of what its consumers send back,
  // added in the order of the reverse sweep (last consumer first)
  std::vector<char> gset(N->has_skips ? nl : 0, 0);
  for (int l = nl - 1; l >= 0; --l) {
    Layer* L = N->layers[l];
    const bool concat = N->has_skips && !N->inputs[l].empty();
    // the plain hand-over (this layer's input gradient IS the previous layer's output
    // gradient, activation derivatives folded across the boundary) needs a sole consumer
    const bool chain = !concat && (!N->has_skips || l == 0 || N->consumers[l - 1] == 1);
    if (N->has_skips && l < nl - 1) {
      if (!gset[l]) continue;  // nobody consumed this layer's output
      if (N->consumers[l] != 1 || !N->inputs[l + 1].empty()) g = N->gsum[l]->as<float>();
    }
    bool needs_gin = l > 0;
    if (concat) {
      needs_gin = false;
      for (int src : N->inputs[l]) needs_gin = needs_gin || src >= 0;
    }
    float* gi = nullptr;
    if (needs_gin) {
      DevBuf& buf = *N->gin[l];
      ATH_TRY(buf.reserve(sizeof(float) * (size_t)std::max<int64_t>(L->in_rows(b) * L->nvf[0], 1)));
      gi = buf.as<float>();
    }
    BwdOpts opt;
    bool folded = false;
    opt.gout_is_preact = g_preact;
    opt.fold_act = (chain && l > 0 && foldable(N->layers[l - 1])) ? N->layers[l - 1]->act
                                                                   : ATHENA_ACT_NONE;
    if (chain && l > 0) {
      const Layer* prev = N->layers[l - 1];
      if (prev->kind == 0 && prev->mask_valid[prev->T - 1])
        opt.fold_mask = prev->MK[prev->T - 1]->as<uint32_t>();
    }
    opt.folded = &folded;
    opt.defer = &defer;
    ATH_TRY(layer_backward_dev(L, b, g, gi, opt));
    g_preact = folded;
    if (chain) {
      if (N->has_skips && l > 0) gset[l - 1] = 1;
      g = gi;
      continue;
    }
    // split the input gradient back over the sources (concat_layer: the reverse of combine)
    const int Fin = L->nvf[0];
    int off = 0;
    std::vector<int> plain{l - 1};
    for (int src : concat ? N->inputs[l] : plain) {
      const int w = src_width(N, src);
      if (src >= 0) {
        DevBuf& acc = *N->gsum[src];
        ATH_TRY(acc.reserve(sizeof(float) * (size_t)std::max<int64_t>(b->V * w, 1)));
        ATH_TRY(launch_copy_cols(acc.as<float>(), w, 0, gi, Fin, off, w, b->V, gset[src] ? 1 : 0));
        gset[src] = 1;
      }
      off += w;
    }
    g = nullptr;
  }
  ATH_TRY(launch_pipe_tn_pending(&defer));  // all queued dW products in one launch
  // gradient exchange over peer memory when it is set up (comm_p2p_*), else NCCL
  P2PState& P = p2p();
  const bool use_p2p = P.ready && P.world > 1 && (size_t)(N->n + 1) <= P.cap;
  float* xout = use_p2p ? P.xbuf[P.rank] + (size_t)((P.epoch + 1) & 1u) * P.cap : nullptr;
  P2PSignal sig;
  if (use_p2p) p2p_next_signal(&sig);
  const bool p2p_step = step_too && !N->opt.d.clip_norm_on;  // min / max rides along (step_one)
  // one launch folds the partials, publishes the slot, waits for the peers, sums and steps
  const bool fuse_exchange = use_p2p && defer.jobs.size() <= 8 && finalize_can_exchange(N->n);
  if (!defer.jobs.empty() || fo.fused || use_p2p) {
    const bool fuse_step = step_too && finalize_can_step(N->opt);
    OptimState* fst = fuse_exchange ? (p2p_step ? &N->opt : nullptr) : (fuse_step ? &N->opt : nullptr);
    // norm clipping on one GPU: the finalize launch clamps and pre-sums for the step that follows
    OptimState* presum = (step_too && fst == nullptr && !use_p2p && comm_world_size() == 1 &&
                          N->opt.d.clip_norm_on) ? &N->opt : nullptr;
    bool tail_stepped = false;  // short parameter vectors: the clipped step rode along
    ATH_TRY(launch_finalize(defer, fo.fused ? fo.loss_part : nullptr, fo.num_parts, gflat + N->n,
                            N->flat_params.as<float>(), gflat, N->n, fst, xout,
                            use_p2p ? &sig : nullptr, fuse_exchange ? 1 : 0, presum,
                            stepped ? &tail_stepped : nullptr));
    if (stepped) *stepped = (fuse_exchange ? p2p_step : fuse_step) || tail_stepped;
  }
  if (use_p2p && !fuse_exchange) {
    // signal (done by the finalize launch) + wait + sum over NVLink + (when no clipping
    // intervenes) the step, in one kernel
    ATH_TRY(launch_p2p_sum_step(N->flat_params.as<float>(), gflat, N->n,
                                p2p_step ? &N->opt : nullptr, /*signalled=*/true));
    if (stepped) *stepped = p2p_step;
  } else if (!use_p2p) {
    ATH_TRY(comm_allreduce_sum(gflat, N->n + 1));
  }
  ATH_TRY(record_mark());
  if (loss) {
    ATH_CUDA(cudaMemcpyAsync(N->pinned_loss, gflat + N->n, sizeof(float), cudaMemcpyDeviceToHost,
                             st));
    ATH_CUDA(cudaStreamSynchronize(st));
    *loss = *N->pinned_loss;
    if (use_p2p) ATH_TRY(p2p_check());
  }
  return ATHENA_OK;
}

}  // namespace athena

using namespace athena;

// Non-finite features on the tensor-core gather: see Batch::force_list.  begin() arms the
// detector, retry() says whether the forward has to be repeated with the list gather.
static int nonfinite_begin(Batch* b, int64_t* mark) {
  b->force_list = false;
  *mark = b->tcg_forwards;
  if (b->abits != nullptr)
    ATH_CUDA(cudaMemsetAsync(b->status.as<int32_t>() + 3, 0, sizeof(int32_t), ctx().stream));
  return ATHENA_OK;
}
static int nonfinite_retry(Batch* b, int64_t mark, bool* again) {
  *again = false;
  if (b->tcg_forwards == mark) return ATHENA_OK;  // no tensor-core forward ran
  int32_t flag = 0;
  ATH_CUDA(cudaMemcpyAsync(&flag, b->status.as<int32_t>() + 3, sizeof(int32_t),
                           cudaMemcpyDeviceToHost, ctx().stream));
  ATH_CUDA(cudaStreamSynchronize(ctx().stream));
  if (flag != 0) {
    b->force_list = true;
    *again = true;
  }
  return ATHENA_OK;
}

// ---- layer ABI ---------------------------------------------------------------------

static int check_act(int a) { return a >= ATHENA_ACT_NONE && a <= ATHENA_ACT_SWISH; }

ATHENA_API int athena_cuda_kipf_layer_create(athena_handle_t* layer, int32_t num_time_steps,
                                             const int32_t* num_vertex_features,
                                             int32_t activation) {
  ATH_TRY(ensure_init());
  ATH_REQUIRE(layer && num_vertex_features, ATHENA_ERR_ARG, "kipf_layer_create: null argument");
  // "Number of time steps must be at least 1": athena_kipf_msgpass_layer.f90:271-274
  ATH_REQUIRE(num_time_steps >= 1, ATHENA_ERR_ARG,
              "kipf_layer_create: num_time_steps must be >= 1 (got %d)", num_time_steps);
  ATH_REQUIRE(check_act(activation), ATHENA_ERR_ARG, "kipf_layer_create: unknown activation %d",
              activation);
  std::unique_ptr<Layer> L(new Layer);
  L->kind = 0;
  L->T = num_time_steps;
  L->act = activation;
  for (int t = 0; t <= num_time_steps; ++t) {
    ATH_REQUIRE(num_vertex_features[t] >= 1, ATHENA_ERR_ARG,
                "kipf_layer_create: num_vertex_features(%d) = %d", t, num_vertex_features[t]);
    L->nvf.push_back(num_vertex_features[t]);
  }
  layer_layout(L.get());
  ATH_TRY(layer_alloc_params(L.get()));
  *layer = register_object(L.release());
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_duvenaud_layer_create(athena_handle_t* layer, int32_t num_time_steps,
                                                 const int32_t* num_vertex_features,
                                                 int32_t num_edge_features,
                                                 int32_t min_vertex_degree,
                                                 int32_t max_vertex_degree, int32_t num_outputs,
                                                 int32_t message_activation,
                                                 int32_t readout_activation) {
  ATH_TRY(ensure_init());
  ATH_REQUIRE(layer && num_vertex_features, ATHENA_ERR_ARG,
              "duvenaud_layer_create: null argument");
  ATH_REQUIRE(num_time_steps >= 1, ATHENA_ERR_ARG,
              "duvenaud_layer_create: num_time_steps must be >= 1 (got %d)", num_time_steps);
  ATH_REQUIRE(num_edge_features >= 0 && num_outputs >= 1, ATHENA_ERR_ARG,
              "duvenaud_layer_create: bad num_edge_features/num_outputs");
  ATH_REQUIRE(min_vertex_degree >= 1 && max_vertex_degree >= min_vertex_degree, ATHENA_ERR_ARG,
              "duvenaud_layer_create: need 1 <= min_vertex_degree <= max_vertex_degree");
  ATH_REQUIRE(check_act(message_activation) && check_act(readout_activation), ATHENA_ERR_ARG,
              "duvenaud_layer_create: unknown activation");
  ATH_REQUIRE(readout_activation != ATHENA_ACT_SWISH, ATHENA_ERR_ARG,
              "duvenaud_layer_create: swish is supported as message activation only");
  std::unique_ptr<Layer> L(new Layer);
  L->kind = 1;
  L->T = num_time_steps;
  L->nef = num_edge_features;
  L->min_deg = min_vertex_degree;
  L->max_deg = max_vertex_degree;
  L->n_out = num_outputs;
  L->act = message_activation;
  L->ract = readout_activation;
  for (int t = 0; t <= num_time_steps; ++t) {
    ATH_REQUIRE(num_vertex_features[t] >= 1, ATHENA_ERR_ARG,
                "duvenaud_layer_create: num_vertex_features(%d) = %d", t, num_vertex_features[t]);
    L->nvf.push_back(num_vertex_features[t]);
  }
  layer_layout(L.get());
  ATH_TRY(layer_alloc_params(L.get()));
  *layer = register_object(L.release());
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_full_layer_create(athena_handle_t* layer, int32_t num_inputs,
                                             int32_t num_outputs, int32_t activation,
                                             int32_t use_bias) {
  ATH_TRY(ensure_init());
  ATH_REQUIRE(layer, ATHENA_ERR_ARG, "full_layer_create: null argument");
  ATH_REQUIRE(num_inputs >= 1 && num_outputs >= 1, ATHENA_ERR_ARG,
              "full_layer_create: num_inputs = %d, num_outputs = %d", num_inputs, num_outputs);
  ATH_REQUIRE(check_act(activation), ATHENA_ERR_ARG, "full_layer_create: unknown activation %d",
              activation);
  std::unique_ptr<Layer> L(new Layer);
  L->kind = 2;
  L->T = 1;
  L->act = activation;
  L->use_bias = use_bias ? 1 : 0;
  L->nvf = {num_inputs, num_outputs};
  layer_layout(L.get());
  ATH_TRY(layer_alloc_params(L.get()));
  *layer = register_object(L.release());
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_layer_destroy(athena_handle_t layer) {
  Layer* L = static_cast<Layer*>(lookup_object(layer, Kind::Layer));
  if (!L) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(!L->adopted, ATHENA_ERR_STATE,
              "layer_destroy: layer belongs to a network; destroy the network instead");
  return destroy_object(layer, Kind::Layer);
}

ATHENA_API int athena_cuda_layer_num_params(athena_handle_t layer, int64_t* n) {
  Layer* L = static_cast<Layer*>(lookup_object(layer, Kind::Layer));
  if (!L) return ATHENA_ERR_HANDLE;
  if (n) *n = L->num_params;
  return ATHENA_OK;
}

static int layer_copy(athena_handle_t layer, float* host_out, const float* host_in, int64_t n,
                      bool grads) {
  Layer* L = static_cast<Layer*>(lookup_object(layer, Kind::Layer));
  if (!L) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(n == L->num_params, ATHENA_ERR_ARG, "layer has %lld parameters, caller passed %lld",
              (long long)L->num_params, (long long)n);
  float* dev = grads ? L->grads : L->params;
  cudaStream_t st = ctx().stream;
  if (host_in) {
    ATH_CUDA(cudaMemcpyAsync(dev, host_in, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st));
    ATH_CUDA(cudaStreamSynchronize(st));
  } else {
    ATH_CUDA(cudaMemcpyAsync(host_out, dev, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, st));
    ATH_CUDA(cudaStreamSynchronize(st));
  }
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_layer_set_params(athena_handle_t layer, const float* host, int64_t n) {
  ATH_REQUIRE(host, ATHENA_ERR_ARG, "set_params: null");
  return layer_copy(layer, nullptr, host, n, false);
}
ATHENA_API int athena_cuda_layer_get_params(athena_handle_t layer, float* host, int64_t n) {
  ATH_REQUIRE(host, ATHENA_ERR_ARG, "get_params: null");
  return layer_copy(layer, host, nullptr, n, false);
}
ATHENA_API int athena_cuda_layer_set_gradients(athena_handle_t layer, const float* host,
                                               int64_t n) {
  ATH_REQUIRE(host, ATHENA_ERR_ARG, "set_gradients: null");
  return layer_copy(layer, nullptr, host, n, true);
}
ATHENA_API int athena_cuda_layer_get_gradients(athena_handle_t layer, float* host, int64_t n) {
  ATH_REQUIRE(host, ATHENA_ERR_ARG, "get_gradients: null");
  return layer_copy(layer, host, nullptr, n, true);
}
ATHENA_API int athena_cuda_layer_zero_gradients(athena_handle_t layer) {
  Layer* L = static_cast<Layer*>(lookup_object(layer, Kind::Layer));
  if (!L) return ATHENA_ERR_HANDLE;
  ATH_CUDA(cudaMemsetAsync(L->grads, 0, sizeof(float) * (size_t)L->num_params, ctx().stream));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_layer_forward(athena_handle_t layer, athena_handle_t batch,
                                         const float* vertex_features, const float* edge_features,
                                         float* output, int32_t mem) {
  Layer* L = static_cast<Layer*>(lookup_object(layer, Kind::Layer));
  Batch* b = static_cast<Batch*>(lookup_object(batch, Kind::Batch));
  if (!L || !b) return ATHENA_ERR_HANDLE;
  const float *dx = nullptr, *de = nullptr, *out = nullptr;
  ATH_TRY(stage_in(L->stage_x, vertex_features, L->in_rows(b) * L->nvf[0], mem, &dx));
  ATH_TRY(stage_in(L->stage_e, edge_features, b->E * L->nef, mem, &de));
  int64_t nf_mark = 0;
  bool again = false;
  ATH_TRY(nonfinite_begin(b, &nf_mark));
  ATH_TRY(layer_forward_dev(L, b, dx, de, &out));
  ATH_TRY(nonfinite_retry(b, nf_mark, &again));
  if (again) {
    int rc = layer_forward_dev(L, b, dx, de, &out);
    b->force_list = false;
    ATH_TRY(rc);
  }
  ATH_TRY(record_mark());
  if (output) {
    size_t bytes = sizeof(float) * (size_t)(L->out_rows(b) * L->out_width());
    cudaStream_t st = ctx().stream;
    if (mem == ATHENA_MEM_HOST) {
      ATH_CUDA(cudaMemcpyAsync(output, out, bytes, cudaMemcpyDeviceToHost, st));
      ATH_CUDA(cudaStreamSynchronize(st));
    } else {
      ATH_CUDA(cudaMemcpyAsync(output, out, bytes, cudaMemcpyDeviceToDevice, st));
    }
  }
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_layer_backward(athena_handle_t layer, athena_handle_t batch,
                                          const float* grad_output, float* grad_input,
                                          int32_t mem) {
  Layer* L = static_cast<Layer*>(lookup_object(layer, Kind::Layer));
  Batch* b = static_cast<Batch*>(lookup_object(batch, Kind::Batch));
  if (!L || !b) return ATHENA_ERR_HANDLE;
  const float* dg;
  ATH_TRY(stage_in(L->stage_g, grad_output, L->out_rows(b) * L->out_width(), mem, &dg));
  float* dgi = grad_input;
  int64_t gin_n = L->in_rows(b) * L->nvf[0];
  if (grad_input && mem == ATHENA_MEM_HOST) {
    ATH_TRY(L->stage_gin.reserve(sizeof(float) * (size_t)std::max<int64_t>(gin_n, 1)));
    dgi = L->stage_gin.as<float>();
  }
  ATH_TRY(layer_backward_dev(L, b, dg, dgi));
  ATH_TRY(record_mark());
  if (grad_input && mem == ATHENA_MEM_HOST) {
    cudaStream_t st = ctx().stream;
    ATH_CUDA(cudaMemcpyAsync(grad_input, dgi, sizeof(float) * (size_t)gin_n,
                             cudaMemcpyDeviceToHost, st));
    ATH_CUDA(cudaStreamSynchronize(st));
  }
  return ATHENA_OK;
}

// per-sample staging of the upstream gradient (the autodiff seam)
static int stage_run(Layer* L, Batch* b) {
  ATH_TRY(layer_backward_dev(L, b, L->stage_g.as<float>(), nullptr));
  L->staged.assign(L->staged.size(), 0);
  L->num_staged = 0;
  return record_mark();
}

ATHENA_API int athena_cuda_layer_backward_stage(athena_handle_t layer, athena_handle_t batch,
                                                int32_t sample, const float* grad_output,
                                                int64_t count) {
  Layer* L = static_cast<Layer*>(lookup_object(layer, Kind::Layer));
  Batch* b = static_cast<Batch*>(lookup_object(batch, Kind::Batch));
  if (!L || !b) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(L->fwd_batch == b && L->fwd_V == b->V, ATHENA_ERR_STATE,
              "backward_stage: no forward pass on this batch");
  ATH_REQUIRE(sample >= 0 && sample < b->B && grad_output, ATHENA_ERR_ARG,
              "backward_stage: bad sample %d (batch of %d) or null gradient", sample, b->B);
  const int W = L->out_width();
  const int64_t row0 = L->kind == 0 ? b->h_voff[sample] : sample;
  const int64_t rows = L->kind == 0 ? b->h_voff[sample + 1] - b->h_voff[sample] : 1;
  ATH_REQUIRE(count == rows * W, ATHENA_ERR_ARG,
              "backward_stage: sample %d has %lld gradient values, caller passed %lld", sample,
              (long long)(rows * W), (long long)count);
  const int64_t total = L->out_rows(b) * W;
  cudaStream_t st = ctx().stream;
  if ((int64_t)L->staged.size() != b->B || L->num_staged == 0) {
    ATH_TRY(L->stage_g.reserve(sizeof(float) * (size_t)std::max<int64_t>(total, 1)));
    ATH_CUDA(cudaMemsetAsync(L->stage_g.p, 0, sizeof(float) * (size_t)std::max<int64_t>(total, 1), st));
    L->staged.assign((size_t)b->B, 0);
    L->num_staged = 0;
  }
  ATH_REQUIRE(!L->staged[sample], ATHENA_ERR_STATE,
              "backward_stage: sample %d was already staged for this sweep", sample);
  if (count > 0)
    ATH_CUDA(cudaMemcpyAsync(L->stage_g.as<float>() + row0 * W, grad_output,
                             sizeof(float) * (size_t)count, cudaMemcpyHostToDevice, st));
  L->staged[sample] = 1;
  L->num_staged += 1;
  if (L->num_staged == b->B) return stage_run(L, b);
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_layer_backward_flush(athena_handle_t layer, athena_handle_t batch) {
  Layer* L = static_cast<Layer*>(lookup_object(layer, Kind::Layer));
  Batch* b = static_cast<Batch*>(lookup_object(batch, Kind::Batch));
  if (!L || !b) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(L->fwd_batch == b && L->fwd_V == b->V, ATHENA_ERR_STATE,
              "backward_flush: no forward pass on this batch");
  if (L->num_staged == 0) return ATHENA_OK;
  return stage_run(L, b);
}

// ---- network ABI -------------------------------------------------------------------

ATHENA_API int athena_cuda_network_create(athena_handle_t* net) {
  ATH_TRY(ensure_init());
  ATH_REQUIRE(net, ATHENA_ERR_ARG, "network_create: null");
  *net = register_object(new Network);
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_network_destroy(athena_handle_t net) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  if (!N) return ATHENA_ERR_HANDLE;
  std::vector<athena_handle_t> hs = N->handles;
  ATH_TRY(destroy_object(net, Kind::Network));
  for (athena_handle_t h : hs) destroy_object(h, Kind::Layer);
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_network_add(athena_handle_t net, athena_handle_t layer) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  Layer* L = static_cast<Layer*>(lookup_object(layer, Kind::Layer));
  if (!N || !L) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(!N->compiled, ATHENA_ERR_STATE, "network_add: network already compiled");
  ATH_REQUIRE(!L->adopted, ATHENA_ERR_STATE, "network_add: layer already belongs to a network");
  if (!N->layers.empty()) {
    Layer* prev = N->layers.back();
    if (L->kind == 2) {
      // dense head: consumes the [num_outputs, batch] output of a graph-level layer
      ATH_REQUIRE(prev->kind != 0, ATHENA_ERR_ARG,
                  "network_add: a full layer must follow a Duvenaud or another full layer");
      ATH_REQUIRE(prev->out_width() == L->nvf[0], ATHENA_ERR_ARG,
                  "network_add: full layer expects %d inputs, previous layer emits %d", L->nvf[0],
                  prev->out_width());
    } else {
      ATH_REQUIRE(prev->kind == 0, ATHENA_ERR_ARG,
                  "network_add: only full layers may follow a graph-level (Duvenaud) output");
      ATH_REQUIRE(prev->nvf[prev->T] == L->nvf[0], ATHENA_ERR_ARG,
                  "network_add: layer expects %d vertex features, previous layer emits %d",
                  L->nvf[0], prev->nvf[prev->T]);
    }
  } else {
    ATH_REQUIRE(L->kind != 2, ATHENA_ERR_ARG,
                "network_add: the first layer must be a message-passing layer");
  }
  L->adopted = true;
  N->layers.push_back(L);
  N->handles.push_back(layer);
  N->inputs.resize(N->layers.size());
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_network_add_inputs(athena_handle_t net, athena_handle_t layer,
                                              int32_t num_inputs, const int32_t* input_list,
                                              int32_t merge_operator) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  Layer* L = static_cast<Layer*>(lookup_object(layer, Kind::Layer));
  if (!N || !L) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(!N->compiled, ATHENA_ERR_STATE, "network_add: network already compiled");
  ATH_REQUIRE(!L->adopted, ATHENA_ERR_STATE, "network_add: layer already belongs to a network");
  ATH_REQUIRE(num_inputs >= 1 && num_inputs <= 8 && input_list, ATHENA_ERR_ARG,
              "network_add: input_list must hold 1..8 entries");
  // operator 1 = concatenate; 2 (add) is accepted by the reference's add() but has no
  // message-passing user; anything else is "invalid operator" (athena_network_sub.f90:820-823)
  ATH_REQUIRE(merge_operator == ATHENA_MERGE_CONCATENATE, ATHENA_ERR_ARG,
              "network_add: invalid operator %d (only concatenate is supported)", merge_operator);
  ATH_REQUIRE(L->kind == 0, ATHENA_ERR_ARG,
              "network_add: only Kipf layers take an input_list (vertex-level concatenation)");
  const int nl = (int)N->layers.size();  // layers added before this one
  ATH_REQUIRE(nl >= 1 || (num_inputs == 1 && input_list[0] == 0), ATHENA_ERR_ARG,
              "network_add: the first layer can only read the network input");
  std::vector<int> srcs;
  int width = 0;
  for (int i = 0; i < num_inputs; ++i) {
    const int id = input_list[i];
    // ids as network%add resolves them (athena_network_sub.f90:832-853): 0 the input layer,
    // k > 0 the k-th added layer, k < 0 counted back from this layer (-1 = the previous one)
    ATH_REQUIRE(id > -(nl + 1) && id <= nl, ATHENA_ERR_ARG,
                "network_add: input vertex index %d out of range (%d:%d)", id, -nl, nl);
    const int src = id == 0 ? -1 : (id < 0 ? nl + id : id - 1);
    if (src >= 0) {
      ATH_REQUIRE(N->layers[src]->kind == 0, ATHENA_ERR_ARG,
                  "network_add: source layer %d is not a Kipf layer", src + 1);
      width += N->layers[src]->nvf[N->layers[src]->T];
    } else {
      ATH_REQUIRE(nl == 0 || N->layers.front()->kind == 0, ATHENA_ERR_ARG,
                  "network_add: the network input is not a vertex-feature input");
      width += nl == 0 ? L->nvf[0] : N->layers.front()->nvf[0];
    }
    srcs.push_back(src);
  }
  ATH_REQUIRE(width == L->nvf[0], ATHENA_ERR_ARG,
              "network_add: layer expects %d vertex features, the listed sources supply %d",
              L->nvf[0], width);
  L->adopted = true;
  N->layers.push_back(L);
  N->handles.push_back(layer);
  N->inputs.resize(N->layers.size());
  if (!(nl == 0) && !(srcs.size() == 1 && srcs[0] == nl - 1)) {
    N->inputs.back() = srcs;
    N->has_skips = true;
  }
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_network_compile(athena_handle_t net,
                                           const athena_optimiser_desc* optimiser) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  if (!N) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(optimiser, ATHENA_ERR_ARG, "network_compile: null optimiser");
  ATH_REQUIRE(!N->layers.empty(), ATHENA_ERR_STATE, "network_compile: no layers");
  ATH_REQUIRE(!N->compiled, ATHENA_ERR_STATE, "network_compile: already compiled");
  ATH_REQUIRE(optimiser->kind >= ATHENA_OPT_SGD && optimiser->kind <= ATHENA_OPT_ADAGRAD,
              ATHENA_ERR_ARG, "network_compile: unknown optimiser kind %d", optimiser->kind);
  ATH_REQUIRE(optimiser->regulariser >= ATHENA_REG_NONE && optimiser->regulariser <= ATHENA_REG_L1L2,
              ATHENA_ERR_ARG, "network_compile: unknown regulariser %d", optimiser->regulariser);
  int64_t n = 0;
  for (Layer* L : N->layers) n += L->num_params;
  N->n = n;
  size_t bytes = sizeof(float) * (size_t)(n + 4);
  ATH_TRY(N->flat_params.reserve(bytes));
  ATH_TRY(N->flat_grads.reserve(bytes));
  cudaStream_t st = ctx().stream;
  ATH_CUDA(cudaMemsetAsync(N->flat_grads.p, 0, bytes, st));
  // re-home every layer's parameters into the flat buffer, layer order x params order
  // (athena_base_layer_sub.f90:545-571, athena_network_sub.f90:2847-2903)
  int64_t off = 0;
  for (Layer* L : N->layers) {
    ATH_CUDA(cudaMemcpyAsync(N->flat_params.as<float>() + off, L->params,
                             sizeof(float) * (size_t)L->num_params, cudaMemcpyDeviceToDevice, st));
    L->params = N->flat_params.as<float>() + off;
    L->grads = N->flat_grads.as<float>() + off;
    off += L->num_params;
  }
  ATH_CUDA(cudaStreamSynchronize(st));
  for (Layer* L : N->layers) {
    L->own_params.release();
    L->own_grads.release();
  }
  N->gin.clear();
  N->cat.clear();
  N->gsum.clear();
  N->inputs.resize(N->layers.size());
  N->consumers.assign(N->layers.size(), 0);
  for (size_t l = 0; l < N->layers.size(); ++l) {
    N->gin.emplace_back(new DevBuf);
    N->cat.emplace_back(new DevBuf);
    N->gsum.emplace_back(new DevBuf);
    if (N->inputs[l].empty()) {
      if (l > 0) N->consumers[l - 1] += 1;
    } else {
      for (int src : N->inputs[l])
        if (src >= 0) N->consumers[src] += 1;
    }
  }
  N->opt.d = *optimiser;
  N->opt.lr = optimiser->learning_rate;
  N->opt.iter = 0;
  ATH_CUDA(cudaHostAlloc((void**)&N->pinned_loss, sizeof(float) * 4, cudaHostAllocDefault));
  N->compiled = true;
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_network_num_params(athena_handle_t net, int64_t* n) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  if (!N) return ATHENA_ERR_HANDLE;
  int64_t s = 0;
  for (Layer* L : N->layers) s += L->num_params;
  if (n) *n = s;
  return ATHENA_OK;
}

static int net_copy(athena_handle_t net, float* host_out, const float* host_in, int64_t n,
                    bool grads) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  if (!N) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(N->compiled, ATHENA_ERR_STATE, "network is not compiled");
  ATH_REQUIRE(n == N->n, ATHENA_ERR_ARG, "network has %lld parameters, caller passed %lld",
              (long long)N->n, (long long)n);
  float* dev = grads ? N->flat_grads.as<float>() : N->flat_params.as<float>();
  cudaStream_t st = ctx().stream;
  if (host_in)
    ATH_CUDA(cudaMemcpyAsync(dev, host_in, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st));
  else
    ATH_CUDA(cudaMemcpyAsync(host_out, dev, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, st));
  ATH_CUDA(cudaStreamSynchronize(st));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_network_set_params(athena_handle_t net, const float* host, int64_t n) {
  ATH_REQUIRE(host, ATHENA_ERR_ARG, "set_params: null");
  return net_copy(net, nullptr, host, n, false);
}
ATHENA_API int athena_cuda_network_get_params(athena_handle_t net, float* host, int64_t n) {
  ATH_REQUIRE(host, ATHENA_ERR_ARG, "get_params: null");
  return net_copy(net, host, nullptr, n, false);
}
ATHENA_API int athena_cuda_network_get_gradients(athena_handle_t net, float* host, int64_t n) {
  ATH_REQUIRE(host, ATHENA_ERR_ARG, "get_gradients: null");
  return net_copy(net, host, nullptr, n, true);
}

ATHENA_API int athena_cuda_network_set_learning_rate(athena_handle_t net, float lr) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  if (!N) return ATHENA_ERR_HANDLE;
  N->opt.lr = lr;
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_network_set_iteration(athena_handle_t net, int64_t iteration) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  if (!N) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(iteration >= 1, ATHENA_ERR_ARG, "set_iteration: iteration = %lld (must be >= 1)",
              (long long)iteration);
  N->opt.iter = iteration;
  N->opt.iter_external = true;
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_network_forward(athena_handle_t net, athena_handle_t batch,
                                           const float* vertex_features,
                                           const float* edge_features, float* output,
                                           int32_t mem) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  Batch* b = static_cast<Batch*>(lookup_object(batch, Kind::Batch));
  if (!N || !b) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(!N->layers.empty(), ATHENA_ERR_STATE, "network_forward: no layers");
  Layer* first = N->layers.front();
  Layer* last = N->layers.back();
  const float *dx = nullptr, *de = nullptr, *out = nullptr;
  ATH_TRY(stage_in(N->stage_x, vertex_features, b->V * first->nvf[0], mem, &dx));
  ATH_TRY(stage_in(N->stage_e, edge_features, b->E * N->edge_width(), mem, &de));
  // network%predict runs in inference mode (athena_network_sub.f90:4226-4303): layers keep
  // nothing for a reverse sweep
  for (Layer* Lr : N->layers) Lr->inference = true;
  int64_t nf_mark = 0;
  bool again = false;
  int frc = nonfinite_begin(b, &nf_mark);
  if (frc == ATHENA_OK) frc = net_forward_dev(N, b, dx, de, &out);
  if (frc == ATHENA_OK) frc = nonfinite_retry(b, nf_mark, &again);
  if (frc == ATHENA_OK && again) {
    frc = net_forward_dev(N, b, dx, de, &out);
    b->force_list = false;
  }
  for (Layer* Lr : N->layers) {
    Lr->inference = false;
    Lr->fwd_batch = nullptr;  // a backward call must be preceded by a training forward
  }
  ATH_TRY(frc);
  ATH_TRY(record_mark());
  if (output) {
    size_t bytes = sizeof(float) * (size_t)(last->out_rows(b) * last->out_width());
    cudaStream_t st = ctx().stream;
    if (mem == ATHENA_MEM_HOST) {
      ATH_CUDA(cudaMemcpyAsync(output, out, bytes, cudaMemcpyDeviceToHost, st));
      ATH_CUDA(cudaStreamSynchronize(st));
    } else {
      ATH_CUDA(cudaMemcpyAsync(output, out, bytes, cudaMemcpyDeviceToDevice, st));
    }
  }
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_network_loss_and_gradients(athena_handle_t net, athena_handle_t batch,
                                                      const float* vertex_features,
                                                      const float* edge_features,
                                                      const float* target, int32_t mem,
                                                      int32_t global_batch, float* loss) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  Batch* b = static_cast<Batch*>(lookup_object(batch, Kind::Batch));
  if (!N || !b) return ATHENA_ERR_HANDLE;
  return net_loss_grads(N, b, vertex_features, edge_features, target, mem, global_batch, loss);
}

ATHENA_API int athena_cuda_network_update(athena_handle_t net) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  if (!N) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(N->compiled, ATHENA_ERR_STATE, "network is not compiled");
  ATH_TRY(launch_update(N->flat_params.as<float>(), N->flat_grads.as<float>(), N->n, N->opt));
  return record_mark();
}

ATHENA_API int athena_cuda_network_train_step(athena_handle_t net, athena_handle_t batch,
                                              const float* vertex_features,
                                              const float* edge_features, const float* target,
                                              int32_t mem, int32_t global_batch, float* loss) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  Batch* b = static_cast<Batch*>(lookup_object(batch, Kind::Batch));
  if (!N || !b) return ATHENA_ERR_HANDLE;
  // the loss read-back (if requested) is deferred until the step is queued,
  // so the host does not stall between backward and update
  bool stepped = false;
  ATH_TRY(net_loss_grads(N, b, vertex_features, edge_features, target, mem, global_batch,
                         nullptr, true, &stepped));
  if (!stepped)
    ATH_TRY(launch_update(N->flat_params.as<float>(), N->flat_grads.as<float>(), N->n, N->opt));
  ATH_TRY(record_mark());
  if (loss) return athena_cuda_network_last_loss(net, loss);
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_network_last_loss(athena_handle_t net, float* loss) {
  Network* N = static_cast<Network*>(lookup_object(net, Kind::Network));
  if (!N) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(N->compiled && loss, ATHENA_ERR_STATE, "last_loss: network not compiled / null");
  cudaStream_t st = ctx().stream;
  ATH_CUDA(cudaMemcpyAsync(N->pinned_loss, N->flat_grads.as<float>() + N->n, sizeof(float),
                           cudaMemcpyDeviceToHost, st));
  ATH_CUDA(cudaStreamSynchronize(st));
  *loss = *N->pinned_loss;
  ATH_TRY(p2p_check());  // a gradient exchange that timed out surfaces here
  return ATHENA_OK;
}
