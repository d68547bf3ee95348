// Internal declarations shared by the translation units of libathena_cuda.
// Nothing here is part of the ABI (see include/athena_cuda.h).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "athena_cuda.h"

#define ATHENA_API extern "C" __attribute__((visibility("default")))

namespace athena {

void set_error(const char* fmt, ...);

#define ATH_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t e_ = (expr);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      ::athena::set_error("%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                          __LINE__);                                                \
      return ATHENA_ERR_CUDA;                                                       \
    }                                                                               \
  } while (0)

#define ATH_REQUIRE(cond, code, ...)     \
  do {                                   \
    if (!(cond)) {                       \
      ::athena::set_error(__VA_ARGS__);  \
      return (code);                     \
    }                                    \
  } while (0)

#define ATH_TRY(expr)          \
  do {                         \
    int rc_ = (expr);          \
    if (rc_ != 0) return rc_;  \
  } while (0)

struct Context {
  bool ready = false;
  int device = 0;
  int sm_count = 0;
  size_t total_mem = 0;
  size_t max_smem_optin = 0;
  cudaStream_t stream = nullptr;
  std::atomic<int64_t> launches{0};
  cudaEvent_t ev_start[16] = {};
  cudaEvent_t ev_stop[16] = {};
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  bool profiling = false;
  // Host->device input copies travel on their own stream so that they overlap the kernels of
  // the same call (batch build under the feature copy, first layer under the target copy).
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t mark = nullptr;       // main-stream position after the last compute call
  bool mark_valid = false;
  cudaEvent_t ev_pool[32] = {};
  int ev_next = 0;
  // direction in which the next fused tile kernel walks its tiles; flips with every launch so
  // that a kernel starts with what its predecessor wrote last (L2-resident), and is reset at
  // the start of every forward / train call (results stay bitwise reproducible per call)
  bool tile_reverse = false;
};
Context& ctx();
int ensure_init();
void pool_trim();  // return every cached device block to the driver

// side (copy) stream helpers, context.cu
int side_begin();                                         // copy stream waits for the last mark
int side_copy(void* dst, const void* src, size_t bytes);  // async H2D on the copy stream
int side_fence(cudaEvent_t* ev);                          // event after the copies queued so far
int main_wait(cudaEvent_t ev);                            // main stream waits for `ev` (nullptr: no-op)
int record_mark();                                        // remember the main-stream position

void prof_mark(const char* tag);  // no-op unless profiling is on

// Count + check a kernel launch.  Use right after <<<>>>.
#define ATH_LAUNCHED() ATH_LAUNCHED_T(__func__)
#define ATH_LAUNCHED_T(tag)                                                    \
  do {                                                                         \
    ::athena::ctx().launches.fetch_add(1, std::memory_order_relaxed);          \
    if (::athena::ctx().profiling) ::athena::prof_mark(tag);                   \
    cudaError_t e_ = cudaPeekAtLastError();                                    \
    if (e_ != cudaSuccess) {                                                   \
      ::athena::set_error("kernel launch: %s (%s:%d)", cudaGetErrorString(e_), \
                          __FILE__, __LINE__);                                 \
      return ATHENA_ERR_CUDA;                                                  \
    }                                                                          \
  } while (0)

// Programmatic dependent launch: the kernel may start (block scheduling, TMEM allocation,
// barrier set-up, instruction-cache warm-up) while the previous kernel of the stream drains;
// it calls pdl_wait() before it touches anything the previous kernel wrote.  Kernels launched
// this way MUST call pdl_wait().  ATHENA_CUDA_DISABLE_PDL=1 launches them plainly.
bool pdl_enabled();
constexpr int L2_HINTS_DEFAULT = 23;  // measured best: read-once loads and P stores evict_first, nothing evict_last
int l2_hint_mask();  // ATHENA_DEBUG_L2_HINTS (see pipe_tcg.cu)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

// Grow-only device buffer: layers keep their activations in these so that a
// training loop performs no cudaMalloc after the first (largest) batch.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

enum class Kind : int { Batch = 1, Layer = 2, Network = 3 };

struct Object {
  Kind kind;
  explicit Object(Kind k) : kind(k) {}
  virtual ~Object() {}
};

athena_handle_t register_object(Object* obj);  // takes ownership
Object* lookup_object(athena_handle_t h, Kind kind);
int destroy_object(athena_handle_t h, Kind kind);

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

// ---------------------------------------------------------------------------
// graph batch (batch.cu)
// ---------------------------------------------------------------------------
constexpr int TILE_ROWS = 128;      // vertices per fused-kernel tile (= UMMA M)
constexpr int TILE_ENTRIES = 3072;
constexpr int LONG_ROW = 512;       // CSR rows / CSC columns longer than this are split over a CTA  // CSR entries per tile that fit the shared-memory ring

struct BucketSet {
  int min_deg = 0, max_deg = 0, D = 0;
  DevBuf bkt;      // [V]   0-based bucket id
  DevBuf perm;     // [V]   stable bucket permutation
  DevBuf bkt_ptr;  // [D+1]
  DevBuf scratch;  // block histograms
};

struct Batch : Object {
  Batch() : Object(Kind::Batch) {}
  int32_t B = 0;
  int64_t V = 0, Z = 0, E = 0;
  std::vector<int32_t> h_voff;  // host copy of voff[B+1] (per-sample staging offsets)
  DevBuf meta;  // int32: nv[B], ne[B], voff[B+1], zoff[B+1], eoff[B+1]
  const int32_t* nv = nullptr;
  const int32_t* ne = nullptr;
  const int32_t* voff = nullptr;
  const int32_t* zoff = nullptr;
  const int32_t* eoff = nullptr;
  DevBuf raw;  // staging of adj_ia / adj_ja when they arrive from the host
  DevBuf ints; // row_ptr, col, eid, deg, vgraph, csc_ptr, csc_src, csc_ent, cursor
  int32_t* row_ptr = nullptr;
  int32_t* col = nullptr;
  int32_t* eid = nullptr;
  int32_t* deg = nullptr;
  int32_t* vgraph = nullptr;
  int32_t* csc_ptr = nullptr;
  int32_t* csc_src = nullptr;
  int32_t* csc_ent = nullptr;
  DevBuf coef_buf;
  float* coef = nullptr;
  DevBuf scratch;
  DevBuf status;  // int32[4]: [0] first bad graph + 1, [1] long-column count, [2] multi-edge /
                  // degree-0 flags, [3] non-finite feature seen by a tensor-core forward
  // Graph-aligned row tiles for the fused gather kernels (pipe_tc.cu): a tile is
  // a run of whole graphs with <= TILE_ROWS vertices and <= TILE_ENTRIES CSR
  // entries, so every neighbour of a tile row lies inside the tile (the batch is
  // block diagonal) and the tile's features can be staged once in shared memory.
  // Built on the host from num_vertices / num_entries (no device sync); valid for
  // the CSR and the CSC alike.  num_tiles == 0: batch not tileable (a graph too big).
  DevBuf tiles;   // int4 {first row, rows, first entry, entries}
  int32_t num_tiles = 0;
  // per-tile compact operands of the fused kernels (built only when num_tiles > 0):
  // neighbour index RELATIVE to the tile's first row (< TILE_ROWS, one byte per entry)
  // for the CSR and for the CSC, and deg(v)^-1/2 per vertex
  DevBuf tile_ops;
  uint8_t* col8 = nullptr;   // [Z] CSR neighbour - tile first row
  uint8_t* csc8 = nullptr;   // [Z] CSC source    - tile first row
  // adjacency rows as 128-bit masks over the tile's vertices (CSR / CSC direction) for the
  // tensor-core gather; multi_edges: -1 unknown (status not read yet), 0 the bit masks are a
  // faithful adjacency, != 0 not (bit 0: some pair repeats -- bits cannot carry a multiplicity;
  // bit 1: some vertex has degree 0, i.e. deg^-1/2 = Infinity, and 0 * Infinity = NaN): the list
  // kernels are used
  uint4* abits = nullptr;
  uint4* atbits = nullptr;
  int multi_edges = -1;
  // The dense adjacency product also spreads a NaN / Inf FEATURE over its whole 128-row tile
  // (0 * Inf), where the reference confines it to the vertex's neighbours.  The forward
  // kernel reports such values in status[3]; the forward entry points then repeat the pass
  // on the generic FP32 kernels (force_list: no fused tile kernel is chosen), which gather
  // entry by entry and multiply in fp32 like the reference.  tcg_forwards counts tensor-core
  // forward launches.
  bool force_list = false;
  int64_t tcg_forwards = 0;
  float* rsdeg = nullptr;    // [V]
  int32_t* vcount = nullptr; // [V] number of vertices of the vertex's graph (MSE cell size)
  // rows (CSR) / columns (CSC) with more than LONG_ROW entries: aggregated by a whole CTA
  // (row split with a fixed-order combine) instead of one lane group
  DevBuf long_buf;           // int32: [0] #long rows, [1] #long columns, then the two lists
  const int32_t* long_rows = nullptr;
  const int32_t* long_cols = nullptr;
  const int32_t* long_counts = nullptr;
  std::vector<std::unique_ptr<BucketSet>> buckets;
  BucketSet* find_buckets(int min_deg, int max_deg) const;
};
int batch_bucketize(Batch* b, int min_deg, int max_deg, BucketSet** out);

// ---------------------------------------------------------------------------
// kernels (spmm.cu / dense.cu / misc.cu): host launchers, all on ctx().stream
// ---------------------------------------------------------------------------
struct DeferList;
struct P2PSignal;

// out[v, 0:F] (+)= sum_{w in row v} c_w * X[col[w], 0:F]   (entries with col < 0 skipped)
// coef == nullptr -> c_w = 1.  accumulate != 0 -> adds to the existing out.
// tail != nullptr  -> out[v, F:F+tail_n] = tail[v, 0:tail_n]  (plain copy).
// long_list / long_count (device; may be null): rows with more than LONG_ROW entries, which
// are then skipped by the row-per-lane-group kernel and split over one CTA each.
int launch_aggregate(const int32_t* row_ptr, const int32_t* col, const float* coef,
                     const float* X, int ldx, int F, float* out, int ldo, int64_t V,
                     int accumulate, const float* tail, int tail_n,
                     const int32_t* long_list = nullptr, const int32_t* long_count = nullptr);

struct GroupDesc {
  // rows are visited through perm (nullptr = identity) in tiles that never
  // straddle a group boundary; group g owns rows [ptr[g], ptr[g+1]) of the
  // permuted order and weight matrix Wbase + g * wstride.
  const int32_t* perm = nullptr;
  const int32_t* ptr = nullptr;  // device [D+1]; nullptr = single group [0, M)
  int D = 1;
  int64_t wstride = 0;
  int scale_by_group = 0;  // divide the A rows (NN/TN) or the result (NT) by (g+1)
};

// C[m, 0:N] = act( (A[m, 0:K] / s_m) . W_g[K, N] )          (row-major W)
int launch_gemm_nn(const float* A, int lda, const float* W, float* C, int ldc, int64_t M,
                   int N, int K, int act, const GroupDesc& gd);
// C[m, 0:N] = ( A[m, 0:K] . W_g[N, K]^T ) / s_m              (row-major W, transposed use)
int launch_gemm_nt(const float* A, int lda, const float* W, float* C, int ldc, int64_t M,
                   int N, int K, const GroupDesc& gd);
// dW_g[k, n] += sum_{m in group g} (A[m, k] / s_m) * G[m, n]  (deterministic two-pass)
int launch_gemm_tn(const float* A, int lda, const float* G, int ldg, float* dW, int64_t M,
                   int N, int K, const GroupDesc& gd, DevBuf& scratch);

// tcgen05 path (gemm_tc.cu): dense row-major operands, feature widths 32/64.
// Hact != nullptr fuses the activation derivative into the operand load:
// the operand becomes A .* act_in'(Hact).
bool tc_rows_supported(int K, int N, int lda, int ldc, const void* A, const void* C);
bool tc_tn_supported(int K, int N, int lda1, int lda2, const void* A1, const void* A2);
int launch_tc_rows(bool transb, const float* A, int lda, const float* Hact, int act_in,
                   const float* W, float* C, int ldc, int64_t M, int N, int K, int act_out);
// defer != nullptr: the fold of the per-CTA partials is queued for launch_finalize (the
// scratch buffer must then stay untouched until the end of the reverse sweep)
int launch_tc_tn(const float* A1, int lda1, const float* A2, int lda2, const float* Hact,
                 int act_in, float* dW, int64_t M, int N, int K, DevBuf& scratch,
                 DeferList* defer = nullptr);

// fused warp-specialised tcgen05 kernels for batches of small graphs (pipe_tc.cu)
bool pipe_gather_supported(const Batch* b, int F, int N);
bool pipe_tn_supported(int K, int N);
// mask_out / mask_in: sign bits of the pre-activation, [V][N/32] words (relu, leaky_relu):
// written by the forward, read by the reverse sweep instead of the saved activations
int launch_pipe_gather_fwd(const Batch* b, const float* X, const float* W, float* P, float* out,
                           int F, int N, int act, uint32_t* mask_out = nullptr);
int launch_pipe_gather_fwd_mse(const Batch* b, const float* X, const float* W, float* P,
                               const float* target, float* grad, int F, int N, int act,
                               float* loss_part, int* num_parts);
int launch_pipe_gather_bwd(const Batch* b, const float* G, const float* W, const float* Hin,
                           float* out, int F, int N, int act, const uint32_t* mask_in = nullptr);
// defer != nullptr: the fold of the per-CTA partials into dW is queued instead of launched
int launch_pipe_tn(const float* P, const float* G, float* dW, int64_t M, int K, int N,
                   DevBuf& scratch, DeferList* defer = nullptr, bool queue = false);

// fused FP32 (FFMA2) tile kernels for batches of small graphs (tile_fma.cu): one Kipf step of
// any width up to 128, and the whole Duvenaud layer (all time steps + readout) per launch
bool tile_kipf_supported(const Batch* b, int Fi, int Fo);
int launch_tile_kipf_fwd(const Batch* b, const float* X, const float* W, float* P, float* out,
                         int Fi, int Fo, int act);
// G: gradient w.r.t. the step output (H != nullptr: act'(H) is applied) or w.r.t. the
// pre-activation (H == nullptr); part receives [*nparts][Fi*Fo] CTA partials of dW
int launch_tile_kipf_bwd(const Batch* b, const float* G, const float* H, const float* P,
                         const float* W, float* gin, int Fi, int Fo, int act, DevBuf& part,
                         int* nparts);
struct TileDuvDesc {
  int T, nef, min_deg, max_deg, no, act, ract;
  const int* nvf;         // (0:T)
  const int64_t* poff;    // offsets of W_1..W_T, R_1..R_T inside the layer's parameter block
  const float* params;
  const float* X;         // [V][F_0]
  const float* E;         // [E][F_e]
  float* Ae;              // [V][F_e] per-vertex edge-feature sums (forward writes, backward reads)
  float* const* Z;        // z_t, t = 1..T (written by the forward, read by the backward)
  float* const* S = nullptr;  // optional readouts S_t = ract(R_t z_t), [V][no]: the forward saves
                              // them, the backward reads them back instead of recomputing them
  int fold_act = 0;       // backward: the input gradient leaves multiplied by fold_act'(X)
};
bool tile_duv_supported(const Batch* b, int T, const int* nvf, int nef, int D, int no);
// target != nullptr: also the [num_outputs, batch] MSE cell -- mse_grad = (out - target) /
// mse_denom and loss_part[0..*num_parts) = per-CTA sums of (out - target)^2 / mse_denom
int launch_tile_duv_fwd(const Batch* b, const TileDuvDesc& d, float* out, const float* target,
                        float* mse_grad, float mse_denom, float* loss_part, int* num_parts);
// part receives [*nparts][num_params] CTA partials of the layer's parameter gradients
int launch_tile_duv_bwd(const Batch* b, const TileDuvDesc& d, const float* gout, float* gin,
                        int64_t num_params, DevBuf& part, int* nparts);

// fused SpMM + tcgen05 transform for large graphs, feature width 128 (agg_tc.cu)
bool agg_tc_supported(int F, int N, const void* X, const void* out);
int launch_agg_tc_fwd(const Batch* b, const float* X, const float* W, float* P, float* out,
                      int act);

// elementwise / row-wise helpers
int launch_act_bwd(int act, const float* Y, const float* G, float* out, int64_t M, int N);
int launch_softmax_rows(float* Y, int64_t M, int N);
// out[s, 0:N] (+)= sum_{v in graph s} Y[v, 0:N]
int launch_segment_sum(const float* Y, int N, const int32_t* voff, int32_t B, float* out,
                       int accumulate);
// dY[v, :] = act_bwd(S[v, :], gout[vgraph[v], :])
int launch_readout_bwd(int act, const float* S, const float* gout, const int32_t* vgraph,
                       float* dY, int64_t V, int N);
int launch_add_inplace(float* dst, const float* src, int64_t n);
// swish (beta = 1): H = X / (1 + exp(-X)); backward on the saved pre-activation X
int launch_swish_fwd(const float* X, float* H, int64_t n);
int launch_swish_bwd(const float* X, const float* G, float* out, int64_t n);
// dst[m, doff : doff + w] (=|+=) src[m, soff : soff + w]   (concatenation and its reverse)
int launch_copy_cols(float* dst, int ldd, int doff, const float* src, int lds, int soff, int w,
                     int64_t M, int accumulate);
// Y[m, n] = act( Y[m, n] + bias[n] )   (bias may be null; act may be softmax over n)
int launch_bias_act(float* Y, const float* bias, int64_t M, int N, int act);
// dst[n] += sum_m G[m, n]   (ascending m, one thread per column: deterministic)
int launch_colsum_add(const float* G, int64_t M, int N, float* dst);

// loss (misc.cu)
// graph output: L = sum_s mean_{F,V_s}((p-e)^2)/2 ; g = (p-e)/(F*V_s).  loss_acc[0] += L.
// act != NONE: grad is additionally multiplied by act'(pred) (layer-boundary fusion)
int launch_mse_graph(const float* pred, const float* target, const int32_t* vgraph,
                     const int32_t* nv, int F, int64_t V, int act, float* grad, float* loss_acc,
                     DevBuf& scratch);
// loss_acc[0] += 0.5 * sum(partial[0..nb))   (single block, fixed order)
int launch_loss_finish(const float* partial, int nb, float* loss_acc);
// array output [B, N]: L = sum((p-e)^2)/(2*N*global_B) ; g = (p-e)/(N*global_B)
int launch_mse_array(const float* pred, const float* target, int64_t n, float denom,
                     float* grad, float* loss_acc, DevBuf& scratch);

struct OptimState {
  athena_optimiser_desc d{};
  float lr = 0.f;
  int64_t iter = 0;
  bool iter_external = false;  // the host owns the iteration counter (set_iteration)
  DevBuf s1, s2;   // velocity | m, v
  DevBuf scratch;  // norm partials
  // > 0: launch_finalize already clamped the gradients and left this many partial sums of
  // squares in `scratch`; launch_update then goes straight to the step
  int presum_nb = 0;
  DevBuf tail;     // block counter of the finalize launch that also steps (short vectors)
};
// clip + optimiser step + zero the gradients (athena_network_sub.f90:2904-2927)
int launch_update(float* params, float* grads, int64_t n, OptimState& st);

// Partial weight-gradient reductions whose final fold into the flat gradient vector is
// deferred to ONE launch at the end of the reverse sweep (launch_finalize).
struct DeferJob {
  const float* part;  // [nparts][count] per-CTA partial sums
  int nparts, count;
  float* dst;         // += into this block of the flat gradient vector
};
// dW = P^T . G products (k_pipe_tn) queued during the reverse sweep and run in one launch
struct TnPending {
  const float* P;
  const float* G;
  float* dW;
  int64_t M;
  int K, N;
  DevBuf* scratch;  // per-CTA partials of this product
};
struct DeferList {
  std::vector<DeferJob> jobs;
  std::vector<TnPending> tn;
};
int launch_pipe_tn_pending(DeferList* defer);
bool finalize_can_step(const OptimState& st);
// xout != nullptr: the finished local gradients (and the loss slot, index n) are also
// written to xout[0..n] -- the staging buffer of the peer-memory exchange.
int launch_finalize(const DeferList& dl, const float* loss_part, int loss_nparts, float* loss_acc,
                    float* params, float* grads, int64_t n, OptimState* st,
                    float* xout = nullptr, const P2PSignal* sig = nullptr, int exchange = 0,
                    OptimState* presum = nullptr, bool* presum_stepped = nullptr);
// exchange != 0: the peer-memory sum (+ the step when st != nullptr) rides on the same launch
bool finalize_can_exchange(int64_t n);

// comm (comm.cu)
int comm_allreduce_sum(float* buf, int64_t n);  // no-op when no communicator
int comm_world_size();

// Peer-memory gradient exchange (comm.cu): every rank exposes a two-slot staging buffer and a
// flag array through CUDA IPC; k_p2p_sum_step (misc.cu) signals, waits and sums over NVLink.
constexpr int P2P_MAX_WORLD = 8;
struct P2PState {
  bool ready = false;
  int world = 1, rank = 0;
  size_t cap = 0;                      // floats per slot
  uint32_t epoch = 0;                  // exchanges done so far
  float* xbuf[P2P_MAX_WORLD] = {};     // [2][cap] staging buffer of every rank (own = local)
  uint32_t* flags[P2P_MAX_WORLD] = {}; // [2][P2P_MAX_WORLD] arrival flags of every rank, then
                                       // (own array only) [P2P_DONE] block counter of the
                                       // signalling kernel, [P2P_ERR] sticky time-out flag
  bool failed = false;                 // a wait timed out: every later exchange is ATHENA_ERR_COMM
};
constexpr int P2P_DONE = 2 * P2P_MAX_WORLD, P2P_ERR = P2P_DONE + 1, P2P_FLAG_WORDS = P2P_DONE + 8;
// what the kernel that finishes the local gradients needs to tell the peers "slot complete"
struct P2PSignal {
  uint32_t* flags[P2P_MAX_WORLD];
  unsigned int* done;  // blocks of the signalling kernel that have finished (last one signals)
  int world = 0, rank = 0, slot = 0;
  uint32_t epoch = 0;
};
// fills `sig` for the NEXT exchange (the one launch_p2p_sum_step will wait for)
void p2p_next_signal(P2PSignal* sig);
// reads the sticky time-out flag (synchronises the stream); ATHENA_ERR_COMM if set
int p2p_check();
P2PState& p2p();
// sums the staged gradients of all ranks in rank order (bitwise identical everywhere) into
// `grads` [0, n] -- or, with st != nullptr, applies the optimiser step and zeroes them
// signalled: the kernel that staged the gradients (launch_finalize with a P2PSignal) already
// published the slot; otherwise this kernel does it itself
int launch_p2p_sum_step(float* params, float* grads, int64_t n, OptimState* st,
                        bool signalled = false);

}  // namespace athena
