/*
 * athena_cuda.h -- C ABI of libathena_cuda, the B200 (sm_100a) implementation
 * of athena's graph message-passing hot path.
 *
 * The reference (nedtaylor/athena v2.1.1, 100 % Fortran) has no FFI/plugin
 * seam for this path; the seam is cut at the Fortran procedures listed below
 * (SURVEY.md section 8b).  Every entry point names the reference procedure it
 * replaces (file:line relative to the athena source tree).  The Fortran side
 * binds these with ISO_C_BINDING -- see fortran/athena_cuda_bindings.f90 and
 * INTEGRATION.md.
 *
 * Conventions
 *   - plain C: pointers, sizes, int32/int64/float.  No C++ or torch types.
 *   - every function returns 0 on success, <0 on error; the message is
 *     available from athena_cuda_last_error() (the shim maps it to
 *     coreutils' stop_program(msg)).
 *   - handles are opaque int64 values (storable in a Fortran integer(c_int64_t)).
 *   - memory order is the reference's: a Fortran val(F,V) column-major array is
 *     passed as-is (== C row-major [V][F]); adj_ja(2,Z) is passed as-is
 *     (interleaved {neighbour, edge id} pairs); indices are 1-based int32.
 *   - the host keeps ownership of every host pointer; the library keeps no
 *     reference to it.  Copies from PINNED host buffers are asynchronous: such a
 *     buffer must stay unchanged until the next call that returns host data (a
 *     loss, an output, *_get_*) or athena_cuda_synchronize(); pageable buffers
 *     are consumed before the call returns.  Device memory is owned by the library.
 *   - there is NO CPU fallback: a call that needs the GPU fails with
 *     ATHENA_ERR_CUDA if no device is usable.
 *   - one compute stream per process (the library's own) plus a copy stream for the
 *     host->device input copies; calls are asynchronous with respect to the host
 *     unless they return host data.
 */
#ifndef ATHENA_CUDA_H
#define ATHENA_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t athena_handle_t;

/* error codes */
#define ATHENA_OK 0
#define ATHENA_ERR_CUDA (-1)      /* CUDA runtime / driver failure, or no device */
#define ATHENA_ERR_ARG (-2)       /* invalid argument / shape mismatch */
#define ATHENA_ERR_HANDLE (-3)    /* unknown or stale handle */
#define ATHENA_ERR_GRAPH (-4)     /* adjacency index outside 1..num_vertices */
#define ATHENA_ERR_STATE (-5)     /* call order violated (e.g. backward before forward) */
#define ATHENA_ERR_COMM (-6)      /* NCCL failure / NCCL not loadable */

/* activation ids: athena_activation_{none,linear,relu,leaky_relu,sigmoid,tanh,softmax,swish}.f90
 * (scale = 1, threshold = 0, leaky alpha = 0.01: athena_activation_leaky_relu.f90:83-85;
 *  softmax is per vertex over features, dim=2: athena_activation_softmax.f90:183-203) */
#define ATHENA_ACT_NONE 0
#define ATHENA_ACT_LINEAR 1
#define ATHENA_ACT_RELU 2
#define ATHENA_ACT_LEAKY_RELU 3
#define ATHENA_ACT_SIGMOID 4
#define ATHENA_ACT_TANH 5
#define ATHENA_ACT_SOFTMAX 6
/* swish, beta = 1 (athena_activation_swish.f90:29-34; x / (1 + exp(-x)),
 * athena_diffstruc_extd_sub.f90:424-486); message / dense activations only */
#define ATHENA_ACT_SWISH 7

/* optimiser kinds: athena_optimiser.f90:634-673 (sgd), :1027-1091 (adam) */
#define ATHENA_OPT_SGD 0
#define ATHENA_OPT_ADAM 1
#define ATHENA_OPT_RMSPROP 2 /* athena_optimiser.f90:771-803; beta in `beta1`, epsilon */
#define ATHENA_OPT_ADAGRAD 3 /* athena_optimiser.f90:898-925; epsilon */

#define ATHENA_REG_NONE 0
#define ATHENA_REG_L1 1
#define ATHENA_REG_L2 2
#define ATHENA_REG_L1L2 3

/* memory space of data pointers handed to *_forward/_backward/_train_step */
#define ATHENA_MEM_HOST 0
#define ATHENA_MEM_DEVICE 1

/* ------------------------------------------------------------------------ */
/* context                                                                   */
/* ------------------------------------------------------------------------ */

/* Select the device and create the library stream.  Idempotent per process.
 * device < 0: use env ATHENA_CUDA_DEVICE, else LOCAL_RANK, else 0. */
int athena_cuda_init(int32_t device);
int athena_cuda_shutdown(void);
/* Thread-local message of the last failing call (never NULL). */
const char* athena_cuda_last_error(void);
int athena_cuda_version(int32_t* major, int32_t* minor);
int athena_cuda_device_info(int32_t* device, int32_t* sm_count, int64_t* total_mem_bytes);
/* Block the host until everything queued on the library stream has finished. */
int athena_cuda_synchronize(void);

/* Device/pinned memory helpers (for callers that keep data resident, e.g.
 * the benchmark's device-resident leg and the tests). */
int athena_cuda_malloc(void** dptr, size_t bytes);
int athena_cuda_free(void* dptr);
int athena_cuda_host_alloc(void** hptr, size_t bytes); /* pinned */
int athena_cuda_host_free(void* hptr);
int athena_cuda_memcpy_h2d(void* dst, const void* src, size_t bytes); /* async on the library stream */
int athena_cuda_memcpy_d2h(void* dst, const void* src, size_t bytes); /* synchronous */
int athena_cuda_memset(void* dst, int value, size_t bytes);

/* CUDA-event timers on the library stream (bench.py; torch.cuda.Event only
 * sees torch's stream). slot in [0,16). */
int athena_cuda_timer_start(int32_t slot);
int athena_cuda_timer_stop(int32_t slot, float* elapsed_ms); /* synchronises on the stop event */
/* Number of kernels this library has launched since init (monotonic). */
int athena_cuda_launch_count(int64_t* n);
/* Write 256 MB of device scratch to evict the 126 MB L2 between timed iterations. */
int athena_cuda_flush_l2(void);

/* Per-kernel device timing for bench.py's roofline line: between begin and
 * end every kernel launch of this library is followed by a CUDA event on the
 * library stream; a kernel's duration is the gap to the previous event
 * (kernels on one stream run back to back).  Results are grouped by kernel tag. */
int athena_cuda_profile_begin(void);
int athena_cuda_profile_end(int32_t* num_tags); /* synchronises */
int athena_cuda_profile_get(int32_t index, char* name, int32_t name_capacity, int64_t* launches,
                            float* total_ms);

/* ------------------------------------------------------------------------ */
/* graph batch  (replaces msgpass_layer_type%set_graph,                      */
/*   athena_msgpass_layer_sub.f90:144-174; Duvenaud override                 */
/*   athena_duvenaud_msgpass_layer.f90:604-641; and the per-sample deep      */
/*   copies of get_sample, athena_network_sub.f90:2145-2164)                 */
/* ------------------------------------------------------------------------ */

/*
 * Build the device representation of a mini-batch of graph_type samples ONCE;
 * every layer and both passes share it:
 *   block-diagonal 0-based CSR (row_ptr, col, eid), per-entry symmetric
 *   normalisation coefficient, degree vector, vertex->graph map, the
 *   transposed CSC (stable in source order) for the backward pass, and --
 *   lazily per (min_degree, max_degree) -- the degree-bucket ids and the
 *   stable bucket permutation that groups vertices by Duvenaud weight matrix.
 *
 *   num_graphs      B
 *   num_vertices    [B]  graph(s)%num_vertices
 *   num_edges       [B]  graph(s)%num_edges  (columns of edge_features)
 *   num_entries     [B]  size(graph(s)%adj_ja, 2)  (CSR entries, == adj_ia(nv+1)-1)
 *   adj_ia          concatenation of graph(s)%adj_ia, each num_vertices(s)+1, 1-based
 *   adj_ja          concatenation of graph(s)%adj_ja(2,:) as stored (interleaved), 1-based;
 *                   adj_ja(2,w) <= 0 marks "no edge feature" (contributes zero)
 *   mem             ATHENA_MEM_HOST or ATHENA_MEM_DEVICE for adj_ia/adj_ja
 *                   (num_vertices/num_edges/num_entries are always host arrays)
 *   validate        != 0: wait for the build and return ATHENA_ERR_GRAPH if any
 *                   neighbour index is outside 1..num_vertices(s) (the reference
 *                   stops with "graph adjacency matrix has indices greater than
 *                   the number of vertices", athena_duvenaud_msgpass_layer.f90:632-639).
 *                   == 0: the check result is deferred to athena_cuda_batch_status().
 */
int athena_cuda_batch_create(athena_handle_t* batch, int32_t num_graphs,
                             const int32_t* num_vertices, const int32_t* num_edges,
                             const int32_t* num_entries, const int32_t* adj_ia,
                             const int32_t* adj_ja, int32_t mem, int32_t validate);

/*
 * The same batch built from the EDGE LISTS, i.e. with graph%generate_adjacency(index_list) and
 * (optionally) graph%add_self_loops() done on the device -- the two graphstruc calls every
 * reader of the reference makes before set_graph (example/example_library/src/
 * mod_read_chemical_graphs.f90:275-276, example/msgpass_euler/src/mod_read_euler.f90:48,
 * example/msgpass_chemical/src/main.f90:108, test/test_kipf_msgpass_layer.f90:93).
 *   index_list      concatenation of the graphs' index_list(2, num_edges(s)): 1-based vertex
 *                   pairs (i, j) of the undirected edges, edge k of a graph = its k-th pair
 * Each edge k = (i, j) is listed in row i and, if i /= j, in row j, with edge id k; a row is in
 * ascending edge id.  add_self_loops != 0 appends (v, edge id 0) to every row that lists no v.
 * (graphstruc v0.2.1 is not vendored with the reference, fpm.toml:21, and no athena test pins
 * its neighbour order: this is the order of the host restatement athena_b200/graph.py, against
 * which the device build is bit-exact.)  An index outside 1..num_vertices(s) is
 * ATHENA_ERR_GRAPH as above.  The per-graph entry counts (size(adj_ja, 2)) are data dependent,
 * so this call synchronises once (num_graphs integers come back) -- unless the caller passes
 * them in num_entries_hint ([B], may be NULL; e.g. remembered from an earlier epoch over the
 * same graphs): the build is then fully asynchronous and a wrong hint is ATHENA_ERR_GRAPH.
 */
int athena_cuda_batch_create_from_edges(athena_handle_t* batch, int32_t num_graphs,
                                        const int32_t* num_vertices, const int32_t* num_edges,
                                        const int32_t* index_list,
                                        const int32_t* num_entries_hint, int32_t add_self_loops,
                                        int32_t mem, int32_t validate);

/*
 * The same batch from the graph inputs of athena's ONNX message-passing export
 * (emit_msgpass_graph_inputs, athena_onnx_msgpass_utils.f90:53-92; built from the CSR by
 * example/msgpass_chemical/validate_onnx.py:38-57):
 *   edge_index      per graph an int64 [3, num_entries(s)] block, row-major, 0-based:
 *                   row 0 = neighbour (source), row 1 = edge-feature index (adj_ja(2,w) - 1;
 *                   negative: none), row 2 = the vertex whose row holds the entry (target);
 *                   entries in CSR order (grouped by target, ascending); blocks concatenated
 *   degree          int64 [sum num_vertices]: CSR row lengths
 * ATHENA_ERR_GRAPH if the degrees of a graph do not add up to its num_entries, an entry does
 * not lie in its target's row, or an index is out of range.
 */
int athena_cuda_batch_create_from_edge_index(athena_handle_t* batch, int32_t num_graphs,
                                             const int32_t* num_vertices,
                                             const int32_t* num_edges,
                                             const int32_t* num_entries,
                                             const int64_t* edge_index, const int64_t* degree,
                                             int32_t mem, int32_t validate);
int athena_cuda_batch_destroy(athena_handle_t batch);
int athena_cuda_batch_status(athena_handle_t batch); /* synchronises; 0 or ATHENA_ERR_GRAPH */
int athena_cuda_batch_info(athena_handle_t batch, int32_t* num_graphs, int64_t* num_vertices,
                           int64_t* num_entries, int64_t* num_edges);
/* Build (or fetch) the degree-bucket structures for (min_degree, max_degree):
 * d(v) = max(min, min(deg(v), max)) - min + 1, athena_diffstruc_extd_sub_duvenaud.f90:206-207 */
int athena_cuda_batch_bucketize(athena_handle_t batch, int32_t min_degree, int32_t max_degree);

/* Copy one of the integer structures back to the host (bit-exact parity tests). */
#define ATHENA_BATCH_ROW_PTR 0   /* [V+1] */
#define ATHENA_BATCH_COL 1       /* [Z]   */
#define ATHENA_BATCH_EID 2       /* [Z]   -1 = none */
#define ATHENA_BATCH_DEG 3       /* [V]   */
#define ATHENA_BATCH_VGRAPH 4    /* [V]   */
#define ATHENA_BATCH_CSC_PTR 5   /* [V+1] */
#define ATHENA_BATCH_CSC_SRC 6   /* [Z]   */
#define ATHENA_BATCH_CSC_ENT 7   /* [Z]   */
#define ATHENA_BATCH_BUCKET 8    /* [V]   0-based bucket id (needs bucketize) */
#define ATHENA_BATCH_PERM 9      /* [V]   */
#define ATHENA_BATCH_BUCKET_PTR 10 /* [D+1] */
#define ATHENA_BATCH_COEF 11     /* [Z] float bits: (deg_v*deg_u)^-1/2 */
int athena_cuda_batch_export(athena_handle_t batch, int32_t what, void* host_out, int64_t count);

/* ------------------------------------------------------------------------ */
/* layers                                                                    */
/* ------------------------------------------------------------------------ */

/* kipf_msgpass_layer_type(num_vertex_features, num_time_steps, activation)
 * athena_kipf_msgpass_layer.f90:80-97,143-212; parameters W_t [F_t, F_{t-1}]
 * column-major, t = 1..T (:347-362).
 *   num_vertex_features: [T+1] = num_vertex_features(0:T) */
int athena_cuda_kipf_layer_create(athena_handle_t* layer, int32_t num_time_steps,
                                  const int32_t* num_vertex_features, int32_t activation);

/* duvenaud_msgpass_layer_type(num_vertex_features, num_edge_features,
 *   num_time_steps, max_vertex_degree, num_outputs, min_vertex_degree,
 *   message_activation, readout_activation)
 * athena_duvenaud_msgpass_layer.f90:88-120,256-365; parameters
 *   params(t)   = W_t [F_t, F_{t-1}+F_e, D] column-major, D = max-min+1  (:550-557)
 *   params(T+t) = R_t [num_outputs, F_t]                                 (:558-560)
 * flat order W_1..W_T, R_1..R_T (athena_base_layer_sub.f90:545-571). */
int athena_cuda_duvenaud_layer_create(athena_handle_t* layer, int32_t num_time_steps,
                                      const int32_t* num_vertex_features,
                                      int32_t num_edge_features, int32_t min_vertex_degree,
                                      int32_t max_vertex_degree, int32_t num_outputs,
                                      int32_t message_activation, int32_t readout_activation);
/* full_layer_type(num_inputs, num_outputs, use_bias, activation) -- the dense head that
 * follows the Duvenaud readout in example/msgpass_chemical (main.f90:139-157).
 * athena_full_layer.f90:147-160 (constructor), :839-874 (forward: act(matmul(W, x) + b));
 * parameters W [num_outputs, num_inputs] column-major, then the bias [num_outputs]
 * (:371-396); num_params = (num_inputs + 1) * num_outputs (:122-138) with a bias.
 * In a network it consumes the [num_outputs, batch] output of a Duvenaud (or another
 * full) layer; forward/backward through athena_cuda_layer_forward/_backward take
 * vertex_features = [B][num_inputs] and return [B][num_outputs]. */
int athena_cuda_full_layer_create(athena_handle_t* layer, int32_t num_inputs,
                                  int32_t num_outputs, int32_t activation, int32_t use_bias);
int athena_cuda_layer_destroy(athena_handle_t layer);
int athena_cuda_layer_num_params(athena_handle_t layer, int64_t* n);

/* learnable_layer_type get/set_params, get/set_gradients: flat real32 vectors
 * in the reference's packing order (athena_base_layer_sub.f90:545-691). */
int athena_cuda_layer_set_params(athena_handle_t layer, const float* host, int64_t n);
int athena_cuda_layer_get_params(athena_handle_t layer, float* host, int64_t n);
int athena_cuda_layer_set_gradients(athena_handle_t layer, const float* host, int64_t n);
int athena_cuda_layer_get_gradients(athena_handle_t layer, float* host, int64_t n);
int athena_cuda_layer_zero_gradients(athena_handle_t layer);

/*
 * layer%forward(input) = update_message + update_readout
 * (athena_msgpass_layer_sub.f90:184-198; Kipf athena_kipf_msgpass_layer.f90:915-959;
 *  Duvenaud athena_duvenaud_msgpass_layer.f90:755-859).
 *   vertex_features  input(1,s)%val concatenated over s: [V_tot][F_0]
 *   edge_features    input(2,s)%val concatenated:        [E_tot][F_e]  (Duvenaud; NULL for Kipf)
 *   output           Kipf: output(1,s)%val concatenated  [V_tot][F_T]
 *                    Duvenaud: output(1,1)%val           [B][num_outputs]
 *                    may be NULL (result stays on the device for backward)
 *   mem              memory space of the three data pointers
 * Activations needed by the backward pass are kept on the device.
  * Non-finite inputs: a NaN / Inf feature value reaches exactly the rows that list its vertex,
 * as in the reference's entry-by-entry gather (athena_diffstruc_extd_sub_kipf.f90:36-45).  The
 * tensor-core gather of small-graph batches would spread it over the vertex's 128-row tile, so
 * the forward kernels flag such values and layer_forward / network_forward repeat the pass
 * with the list gather (one extra synchronisation per call on tileable batches).  A TRAINING
 * step is not repeated: its loss sums over all samples and its weight gradients over all
 * vertices, so loss and parameters are NaN in the reference and here alike.
 */
int athena_cuda_layer_forward(athena_handle_t layer, athena_handle_t batch,
                              const float* vertex_features, const float* edge_features,
                              float* output, int32_t mem);

/*
 * Reverse sweep of the layer (what loss%grad_reverse does through
 * get_partial_kipf_propagate_left_val athena_diffstruc_extd_sub_kipf.f90:85-111,
 * get_partial_duvenaud_* athena_diffstruc_extd_sub_duvenaud.f90:115-171,284-368,
 * get_partial_softmax_val athena_diffstruc_extd_sub.f90:355-379 and diffstruc's
 * matmul/activation partials).  Must follow a forward on the same batch.
 *   grad_output   same shape as `output` of forward
 *   grad_input    [V_tot][F_0] or NULL (the input layer never requires a
 *                 gradient, athena_input_layer.f90:541)
 * Parameter gradients are ACCUMULATED into the layer's gradient buffer (as the
 * reference accumulates into params(t)%grad); zero them with
 * athena_cuda_layer_zero_gradients or a network update.
 */
int athena_cuda_layer_backward(athena_handle_t layer, athena_handle_t batch,
                               const float* grad_output, float* grad_input, int32_t mem);

/*
 * The same reverse sweep, fed one SAMPLE at a time: the shape in which diffstruc's
 * loss%grad_reverse reaches a layer whose output(1,s) are separate autodiff nodes (one
 * get_partial_*_val callback per sample, athena_diffstruc_extd_sub_kipf.f90:85-95).
 *   sample        0-based index into the batch
 *   grad_output   that sample's upstream gradient: Kipf [nv_s][F_T]; Duvenaud / full
 *                 [num_outputs]; host memory
 *   count         number of floats in grad_output (checked)
 * The gradient is parked in the layer's staging buffer; when the last sample of the batch
 * has been staged the reverse sweep of the whole batch runs (parameter gradients
 * accumulate on the device, no input gradient is produced).  _flush runs it with zero
 * gradients for the samples that were not staged (a loss that ignores some samples).
 */
int athena_cuda_layer_backward_stage(athena_handle_t layer, athena_handle_t batch, int32_t sample,
                                     const float* grad_output, int64_t count);
int athena_cuda_layer_backward_flush(athena_handle_t layer, athena_handle_t batch);

/* ------------------------------------------------------------------------ */
/* network: the train-step skeleton around the layers                        */
/* (network_type%add/compile/forward/train/update/predict,                   */
/*  athena_network.f90:142-204, athena_network_sub.f90:2639-2929,3611-3670)  */
/* ------------------------------------------------------------------------ */

int athena_cuda_network_create(athena_handle_t* net);
int athena_cuda_network_destroy(athena_handle_t net);
/* network%add(layer): layers run in order; each layer's vertex input is the
 * previous layer's node-level output.  A Duvenaud layer reads the ORIGINAL edge
 * features (SURVEY finding 7) and emits a graph-level [num_outputs, batch] array, after
 * which only full layers may follow.  The network takes ownership of the layer's
 * parameters (re-homed into one flat buffer, layer order x params order). */
int athena_cuda_network_add(athena_handle_t net, athena_handle_t layer);
/* network%add(layer, input_list, operator = 'concatenate') (athena_network_sub.f90:764-830; the
 * skip-connected Kipf stack of example/msgpass_euler/src/main.f90:192-255): the layer's vertex
 * input is the concatenation, along the feature axis and in list order, of the vertex outputs
 * of the listed sources (concat_layer_type%combine, athena_concat_layer.f90:413-456).
 *   input_list[i] = 0    the network's input features (the input layer)
 *                 = k>0  the k-th layer added to this network (1-based)
 *                 = k<0  counted back from the layer being added: -1 = the layer added last
 *                        (vertex_index = num_vertices + input_list(i), :848-849)
 * ids outside (-n, n], n = layers added so far, are ATHENA_ERR_ARG ("input vertex index out of
 * range", :835-847), as is any operator but concatenate ("invalid operator", :820-823; the
 * reference also accepts 'add', which no message-passing network uses).
 * Sources and the layer itself must be Kipf layers; the sum of the source widths must equal
 * the layer's num_vertex_features(0).  In the reverse sweep the input gradient is split back
 * over the sources; a layer read by several consumers receives the sum of their shares, added
 * in the order of the reverse sweep (the network input receives none). */
#define ATHENA_MERGE_CONCATENATE 1
int athena_cuda_network_add_inputs(athena_handle_t net, athena_handle_t layer, int32_t num_inputs,
                                   const int32_t* input_list, int32_t merge_operator);

typedef struct athena_optimiser_desc {
  int32_t kind;          /* ATHENA_OPT_* */
  float learning_rate;
  float beta1, beta2, epsilon; /* adam; rmsprop: beta1 = beta; adagrad: epsilon */
  float momentum;        /* sgd */
  int32_t nesterov;      /* sgd */
  int32_t clip_min_max;  /* clip_type%l_min_max, athena_clipper.f90:190-193 */
  float clip_min, clip_max;
  int32_t clip_norm_on;  /* clip_type%l_norm, :196-203 */
  float clip_norm;
  /* regulariser%regularise(param, gradient, learning_rate), applied inside minimise_* before
   * the step (athena_regulariser.f90:85-137): gradient += lr * (l1 sign(1,p) + 2 l2 p).
   * With an l2 regulariser minimise_adam additionally decays the parameter, decoupled
   * (AdamW) or inside the quotient (athena_optimiser.f90:1064-1084). */
  int32_t regulariser;   /* ATHENA_REG_* */
  float l1, l2;
  int32_t l2_decoupled;  /* l2_regulariser_type%decoupled (default .true.) */
} athena_optimiser_desc;

/* network%compile(optimiser, loss_method="mse").  Loss: athena_loss.f90:393-430. */
int athena_cuda_network_compile(athena_handle_t net, const athena_optimiser_desc* optimiser);
int athena_cuda_network_num_params(athena_handle_t net, int64_t* n);
int athena_cuda_network_set_params(athena_handle_t net, const float* host, int64_t n);
int athena_cuda_network_get_params(athena_handle_t net, float* host, int64_t n);
int athena_cuda_network_get_gradients(athena_handle_t net, float* host, int64_t n);
/* learning rate for the next update (host-side lr_decay%get_lr result,
 * athena_lr_decay.f90:200-216, stays a Fortran scalar). */
int athena_cuda_network_set_learning_rate(athena_handle_t net, float lr);
/* The optimiser's iteration counter as the host keeps it (this%optimiser%iter): by default the
 * library increments its own counter before every step (athena_network_sub.f90:2834-2841); a
 * network whose lr_decay iterates per epoch (step_lr_decay_type, athena_lr_decay.f90:169)
 * advances it once per epoch only (:2834-2838) -- and Adam's bias corrections read the same
 * counter (athena_optimiser.f90:1058-1059).  After this call the library uses the value given
 * (>= 1) for the next steps and no longer increments it: call it before every step. */
int athena_cuda_network_set_iteration(athena_handle_t net, int64_t iteration);

/* network%forward / predict (athena_network_sub.f90:2639-2768, 4226-4303). */
int athena_cuda_network_forward(athena_handle_t net, athena_handle_t batch,
                                const float* vertex_features, const float* edge_features,
                                float* output, int32_t mem);

/*
 * One iteration of the batch loop of network%train
 * (athena_network_sub.f90:3611-3670): forward, MSE loss, reverse sweep,
 * [gradient all-reduce over the communicator], clip + optimiser step
 * (network%update, :2816-2929; iter is incremented before the step),
 * zero gradients.
 *   target        Kipf-last network: graph target, [V_tot][F_T]
 *                 Duvenaud-last network: [B][num_outputs]
 *   global_batch  number of graphs in the whole (all ranks) mini-batch; used
 *                 for the mean of the [num_outputs, batch] MSE cell.  <= 0: B.
 *   loss          host pointer, may be NULL.  Receives the GLOBAL batch loss
 *                 (forces a synchronisation when non-NULL).
 */
int athena_cuda_network_train_step(athena_handle_t net, athena_handle_t batch,
                                   const float* vertex_features, const float* edge_features,
                                   const float* target, int32_t mem, int32_t global_batch,
                                   float* loss);
/* Same, without the optimiser step: leaves (all-reduced) gradients in place. */
int athena_cuda_network_loss_and_gradients(athena_handle_t net, athena_handle_t batch,
                                           const float* vertex_features,
                                           const float* edge_features, const float* target,
                                           int32_t mem, int32_t global_batch, float* loss);
/* network%update() on the current gradients. */
int athena_cuda_network_update(athena_handle_t net);
/* Loss of the most recent step (synchronises). */
int athena_cuda_network_last_loss(athena_handle_t net, float* loss);

/* ------------------------------------------------------------------------ */
/* data parallelism: one process per GPU, graphs sharded across ranks,       */
/* gradient all-reduce with NCCL over NVLink (semantic template:             */
/* network_type%reduce, athena_network_sub.f90:36-62)                        */
/* ------------------------------------------------------------------------ */

#define ATHENA_COMM_ID_BYTES 128
/* Rank 0 creates the id and ships it to the other ranks by any host channel
 * (torch.distributed store, MPI_Bcast, a file). */
int athena_cuda_comm_unique_id(char id[ATHENA_COMM_ID_BYTES]);
int athena_cuda_comm_init(int32_t world_size, int32_t rank, const char id[ATHENA_COMM_ID_BYTES]);
int athena_cuda_comm_destroy(void);
int athena_cuda_comm_info(int32_t* world_size, int32_t* rank);

/*
 * Optional peer-memory gradient exchange (one box, NVLink / NVSwitch): instead of a separate
 * NCCL all-reduce, every rank exposes a staging buffer through CUDA IPC and ONE kernel on each
 * rank signals, waits, reads the peers' staged gradients over NVLink, adds them in rank
 * order and applies the optimiser step.  Every rank calls _export, the 128-byte handles are
 * all-gathered by any host channel, every rank calls _import with the world_size handles in
 * rank order.  Falls back to NCCL when not set up (or for gradient vectors > 1 Mi floats).
 */
#define ATHENA_P2P_HANDLE_BYTES 128
int athena_cuda_comm_p2p_export(char handle[ATHENA_P2P_HANDLE_BYTES]);
int athena_cuda_comm_p2p_import(int32_t world_size, int32_t rank, const char* handles);

/* Host-side helper: contiguous partition of B graphs over world_size ranks,
 * balanced by CSR entries.  first_graph has world_size+1 entries. Pure host code. */
int athena_cuda_shard_graphs(int32_t num_graphs, const int64_t* entries_per_graph,
                             int32_t world_size, int32_t* first_graph);

#ifdef __cplusplus
}
#endif
#endif /* ATHENA_CUDA_H */
