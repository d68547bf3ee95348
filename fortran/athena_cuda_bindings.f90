!> ISO_C_BINDING interfaces to libathena_cuda (include/athena_cuda.h).
!>
!> This is the reference-side binding: athena's Fortran API (layer constructors,
!> set_graph, forward, network%train/update/predict) stays byte-for-byte the same
!> and the bodies named in INTEGRATION.md marshal to these entry points.
!>
!> Memory conventions match the C header one to one:
!>   - a Fortran real(real32) val(F,V) array is passed as-is (C row-major [V][F]);
!>   - adj_ja(2,Z) is passed as-is (interleaved {neighbour, edge id}), 1-based;
!>   - handles are integer(c_int64_t); every function returns integer(c_int),
!>     0 = ok, <0 = error with the text available from athena_cuda_last_error().
!>
!> NOTE: no Fortran compiler exists in the image this repository is developed in
!> (SURVEY.md section 0.2), so this file is kept mechanical: one interface block
!> per C prototype, argument order and kinds copied from the header.  The same
!> ABI is exercised from Python/ctypes by tests/ (every exported symbol is
!> checked against the header by tests/test_abi_cpu.py).
module athena__cuda_bindings
  use, intrinsic :: iso_c_binding
  implicit none
  private

  integer(c_int), parameter, public :: ATHENA_OK = 0
  integer(c_int), parameter, public :: ATHENA_ERR_CUDA = -1
  integer(c_int), parameter, public :: ATHENA_ERR_ARG = -2
  integer(c_int), parameter, public :: ATHENA_ERR_HANDLE = -3
  integer(c_int), parameter, public :: ATHENA_ERR_GRAPH = -4
  integer(c_int), parameter, public :: ATHENA_ERR_STATE = -5
  integer(c_int), parameter, public :: ATHENA_ERR_COMM = -6

  integer(c_int32_t), parameter, public :: ATHENA_ACT_NONE = 0
  integer(c_int32_t), parameter, public :: ATHENA_ACT_LINEAR = 1
  integer(c_int32_t), parameter, public :: ATHENA_ACT_RELU = 2
  integer(c_int32_t), parameter, public :: ATHENA_ACT_LEAKY_RELU = 3
  integer(c_int32_t), parameter, public :: ATHENA_ACT_SIGMOID = 4
  integer(c_int32_t), parameter, public :: ATHENA_ACT_TANH = 5
  integer(c_int32_t), parameter, public :: ATHENA_ACT_SOFTMAX = 6, ATHENA_ACT_SWISH = 7

  integer(c_int32_t), parameter, public :: ATHENA_OPT_SGD = 0
  integer(c_int32_t), parameter, public :: ATHENA_OPT_ADAM = 1
  integer(c_int32_t), parameter, public :: ATHENA_OPT_RMSPROP = 2
  integer(c_int32_t), parameter, public :: ATHENA_OPT_ADAGRAD = 3
  integer(c_int32_t), parameter, public :: ATHENA_REG_NONE = 0, ATHENA_REG_L1 = 1
  integer(c_int32_t), parameter, public :: ATHENA_REG_L2 = 2, ATHENA_REG_L1L2 = 3
  integer(c_int32_t), parameter, public :: ATHENA_MEM_HOST = 0
  integer(c_int32_t), parameter, public :: ATHENA_MEM_DEVICE = 1
  integer, parameter, public :: ATHENA_COMM_ID_BYTES = 128

  !> struct athena_optimiser_desc (include/athena_cuda.h)
  type, bind(C), public :: athena_optimiser_desc
     integer(c_int32_t) :: kind
     real(c_float) :: learning_rate
     real(c_float) :: beta1, beta2, epsilon
     real(c_float) :: momentum
     integer(c_int32_t) :: nesterov
     integer(c_int32_t) :: clip_min_max
     real(c_float) :: clip_min, clip_max
     integer(c_int32_t) :: clip_norm_on
     real(c_float) :: clip_norm
     integer(c_int32_t) :: regulariser
     real(c_float) :: l1, l2
     integer(c_int32_t) :: l2_decoupled
  end type athena_optimiser_desc

  public :: athena_cuda_init, athena_cuda_shutdown, athena_cuda_last_error
  public :: athena_cuda_synchronize, athena_cuda_device_info
  public :: athena_cuda_batch_create, athena_cuda_batch_destroy, athena_cuda_batch_status
  public :: athena_cuda_batch_bucketize
  public :: athena_cuda_kipf_layer_create, athena_cuda_duvenaud_layer_create
  public :: athena_cuda_full_layer_create
  public :: athena_cuda_layer_destroy, athena_cuda_layer_num_params
  public :: athena_cuda_layer_set_params, athena_cuda_layer_get_params
  public :: athena_cuda_layer_set_gradients, athena_cuda_layer_get_gradients
  public :: athena_cuda_layer_zero_gradients
  public :: athena_cuda_layer_forward, athena_cuda_layer_backward
  public :: athena_cuda_network_create, athena_cuda_network_destroy, athena_cuda_network_add
  public :: athena_cuda_network_add_inputs, athena_cuda_network_set_iteration
  public :: athena_cuda_batch_create_from_edges, athena_cuda_batch_create_from_edge_index
  public :: athena_cuda_network_compile, athena_cuda_network_num_params
  public :: athena_cuda_network_set_params, athena_cuda_network_get_params
  public :: athena_cuda_network_get_gradients, athena_cuda_network_set_learning_rate
  public :: athena_cuda_network_forward, athena_cuda_network_train_step
  public :: athena_cuda_network_loss_and_gradients, athena_cuda_network_update
  public :: athena_cuda_network_last_loss
  public :: athena_cuda_comm_unique_id, athena_cuda_comm_init, athena_cuda_comm_destroy
  public :: athena_cuda_comm_info, athena_cuda_shard_graphs
  public :: athena_cuda_comm_p2p_export, athena_cuda_comm_p2p_import
  public :: athena_cuda_layer_backward_stage, athena_cuda_layer_backward_flush
  public :: athena_cuda_layer_backward_stage_pure
  public :: athena_cuda_check

  interface
     ! ---- context -----------------------------------------------------------
     function athena_cuda_init(device) bind(C, name="athena_cuda_init") result(rc)
       import :: c_int, c_int32_t
       integer(c_int32_t), value :: device
       integer(c_int) :: rc
     end function
     function athena_cuda_shutdown() bind(C, name="athena_cuda_shutdown") result(rc)
       import :: c_int
       integer(c_int) :: rc
     end function
     function athena_cuda_last_error_c() bind(C, name="athena_cuda_last_error") result(msg)
       import :: c_ptr
       type(c_ptr) :: msg
     end function
     function athena_cuda_synchronize() bind(C, name="athena_cuda_synchronize") result(rc)
       import :: c_int
       integer(c_int) :: rc
     end function
     function athena_cuda_device_info(device, sm_count, total_mem_bytes) &
          bind(C, name="athena_cuda_device_info") result(rc)
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int32_t), intent(out) :: device, sm_count
       integer(c_int64_t), intent(out) :: total_mem_bytes
       integer(c_int) :: rc
     end function

     ! ---- graph batch: replaces msgpass_layer_type%set_graph
     !      (athena_msgpass_layer_sub.f90:144-174) ---------------------------------
     function athena_cuda_batch_create(batch, num_graphs, num_vertices, num_edges, &
          num_entries, adj_ia, adj_ja, mem, validate) &
          bind(C, name="athena_cuda_batch_create") result(rc)
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int64_t), intent(out) :: batch
       integer(c_int32_t), value :: num_graphs
       integer(c_int32_t), intent(in) :: num_vertices(*), num_edges(*), num_entries(*)
       integer(c_int32_t), intent(in) :: adj_ia(*)   ! concatenated adj_ia of every sample
       integer(c_int32_t), intent(in) :: adj_ja(2,*) ! concatenated adj_ja(2,:) of every sample
       integer(c_int32_t), value :: mem, validate
       integer(c_int) :: rc
     end function
     !! graph%generate_adjacency(index_list) (+ graph%add_self_loops()) on the device
     function athena_cuda_batch_create_from_edges(batch, num_graphs, num_vertices, num_edges, &
          index_list, num_entries_hint, add_self_loops, mem, validate) &
          bind(C, name="athena_cuda_batch_create_from_edges") result(rc)
       import :: c_int, c_int32_t, c_int64_t, c_ptr
       integer(c_int64_t), intent(out) :: batch
       integer(c_int32_t), value :: num_graphs, add_self_loops, mem, validate
       integer(c_int32_t), intent(in) :: num_vertices(*), num_edges(*), index_list(*)
       type(c_ptr), value :: num_entries_hint   !! c_null_ptr, or c_loc of the known entry counts
       integer(c_int) :: rc
     end function
     !! ONNX graph inputs: edge_index [3, ncsr] + degree (athena_onnx_msgpass_utils.f90:53-92)
     function athena_cuda_batch_create_from_edge_index(batch, num_graphs, num_vertices, &
          num_edges, num_entries, edge_index, degree, mem, validate) &
          bind(C, name="athena_cuda_batch_create_from_edge_index") result(rc)
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int64_t), intent(out) :: batch
       integer(c_int32_t), value :: num_graphs, mem, validate
       integer(c_int32_t), intent(in) :: num_vertices(*), num_edges(*), num_entries(*)
       integer(c_int64_t), intent(in) :: edge_index(*), degree(*)
       integer(c_int) :: rc
     end function
     function athena_cuda_batch_destroy(batch) bind(C, name="athena_cuda_batch_destroy") result(rc)
       import :: c_int, c_int64_t
       integer(c_int64_t), value :: batch
       integer(c_int) :: rc
     end function
     function athena_cuda_batch_status(batch) bind(C, name="athena_cuda_batch_status") result(rc)
       import :: c_int, c_int64_t
       integer(c_int64_t), value :: batch
       integer(c_int) :: rc
     end function
     function athena_cuda_batch_bucketize(batch, min_degree, max_degree) &
          bind(C, name="athena_cuda_batch_bucketize") result(rc)
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int64_t), value :: batch
       integer(c_int32_t), value :: min_degree, max_degree
       integer(c_int) :: rc
     end function

     ! ---- layers -------------------------------------------------------------
     function athena_cuda_kipf_layer_create(layer, num_time_steps, num_vertex_features, &
          activation) bind(C, name="athena_cuda_kipf_layer_create") result(rc)
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int64_t), intent(out) :: layer
       integer(c_int32_t), value :: num_time_steps
       integer(c_int32_t), intent(in) :: num_vertex_features(*)  ! (0:T)
       integer(c_int32_t), value :: activation
       integer(c_int) :: rc
     end function
     function athena_cuda_duvenaud_layer_create(layer, num_time_steps, num_vertex_features, &
          num_edge_features, min_vertex_degree, max_vertex_degree, num_outputs, &
          message_activation, readout_activation) &
          bind(C, name="athena_cuda_duvenaud_layer_create") result(rc)
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int64_t), intent(out) :: layer
       integer(c_int32_t), value :: num_time_steps
       integer(c_int32_t), intent(in) :: num_vertex_features(*)  ! (0:T)
       integer(c_int32_t), value :: num_edge_features, min_vertex_degree, max_vertex_degree
       integer(c_int32_t), value :: num_outputs, message_activation, readout_activation
       integer(c_int) :: rc
     end function
     function athena_cuda_full_layer_create(layer, num_inputs, num_outputs, activation, &
          use_bias) bind(C, name="athena_cuda_full_layer_create") result(rc)
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int64_t), intent(out) :: layer
       integer(c_int32_t), value :: num_inputs, num_outputs, activation, use_bias
       integer(c_int) :: rc
     end function
     function athena_cuda_layer_destroy(layer) bind(C, name="athena_cuda_layer_destroy") result(rc)
       import :: c_int, c_int64_t
       integer(c_int64_t), value :: layer
       integer(c_int) :: rc
     end function
     function athena_cuda_layer_num_params(layer, n) &
          bind(C, name="athena_cuda_layer_num_params") result(rc)
       import :: c_int, c_int64_t
       integer(c_int64_t), value :: layer
       integer(c_int64_t), intent(out) :: n
       integer(c_int) :: rc
     end function
     function athena_cuda_layer_set_params(layer, host, n) &
          bind(C, name="athena_cuda_layer_set_params") result(rc)
       import :: c_int, c_int64_t, c_float
       integer(c_int64_t), value :: layer, n
       real(c_float), intent(in) :: host(*)
       integer(c_int) :: rc
     end function
     function athena_cuda_layer_get_params(layer, host, n) &
          bind(C, name="athena_cuda_layer_get_params") result(rc)
       import :: c_int, c_int64_t, c_float
       integer(c_int64_t), value :: layer, n
       real(c_float), intent(out) :: host(*)
       integer(c_int) :: rc
     end function
     function athena_cuda_layer_set_gradients(layer, host, n) &
          bind(C, name="athena_cuda_layer_set_gradients") result(rc)
       import :: c_int, c_int64_t, c_float
       integer(c_int64_t), value :: layer, n
       real(c_float), intent(in) :: host(*)
       integer(c_int) :: rc
     end function
     function athena_cuda_layer_get_gradients(layer, host, n) &
          bind(C, name="athena_cuda_layer_get_gradients") result(rc)
       import :: c_int, c_int64_t, c_float
       integer(c_int64_t), value :: layer, n
       real(c_float), intent(out) :: host(*)
       integer(c_int) :: rc
     end function
     function athena_cuda_layer_zero_gradients(layer) &
          bind(C, name="athena_cuda_layer_zero_gradients") result(rc)
       import :: c_int, c_int64_t
       integer(c_int64_t), value :: layer
       integer(c_int) :: rc
     end function
     !> layer%forward(input): vertex_features = input(1,:) concatenated, edge_features =
     !> input(2,:) concatenated (c_null_ptr for Kipf).  Pointers are type(c_ptr) so that
     !> either host arrays (c_loc) or device addresses can be passed, selected by `mem`.
     function athena_cuda_layer_forward(layer, batch, vertex_features, edge_features, output, &
          mem) bind(C, name="athena_cuda_layer_forward") result(rc)
       import :: c_int, c_int32_t, c_int64_t, c_ptr
       integer(c_int64_t), value :: layer, batch
       type(c_ptr), value :: vertex_features, edge_features, output
       integer(c_int32_t), value :: mem
       integer(c_int) :: rc
     end function
     function athena_cuda_layer_backward(layer, batch, grad_output, grad_input, mem) &
          bind(C, name="athena_cuda_layer_backward") result(rc)
       import :: c_int, c_int32_t, c_int64_t, c_ptr
       integer(c_int64_t), value :: layer, batch
       type(c_ptr), value :: grad_output, grad_input
       integer(c_int32_t), value :: mem
       integer(c_int) :: rc
     end function

     ! per-sample staging of the upstream gradient (the autodiff seam)
     function athena_cuda_layer_backward_stage(layer, batch, sample, grad_output, count) &
          bind(C, name="athena_cuda_layer_backward_stage") result(rc)
       import :: c_int, c_int32_t, c_int64_t, c_float
       integer(c_int64_t), value :: layer, batch
       integer(c_int32_t), value :: sample
       real(c_float), intent(in) :: grad_output(*)
       integer(c_int64_t), value :: count
       integer(c_int) :: rc
     end function
     !> The same C symbol under a PURE interface: diffstruc declares its get_partial_*_val
     !> callbacks pure (athena_diffstruc_extd_sub_kipf.f90:85), and a pure procedure may only
     !> reference pure procedures.  The C function has no Fortran-visible side effect: it
     !> copies `grad_output` to the device and touches device state only.
     pure function athena_cuda_layer_backward_stage_pure(layer, batch, sample, grad_output, &
          count) bind(C, name="athena_cuda_layer_backward_stage") result(rc)
       import :: c_int, c_int32_t, c_int64_t, c_float
       integer(c_int64_t), value :: layer, batch
       integer(c_int32_t), value :: sample
       real(c_float), intent(in) :: grad_output(*)
       integer(c_int64_t), value :: count
       integer(c_int) :: rc
     end function
     function athena_cuda_layer_backward_flush(layer, batch) &
          bind(C, name="athena_cuda_layer_backward_flush") result(rc)
       import :: c_int, c_int64_t
       integer(c_int64_t), value :: layer, batch
       integer(c_int) :: rc
     end function

     ! ---- network --------------------------------------------------------------
     function athena_cuda_network_create(net) bind(C, name="athena_cuda_network_create") result(rc)
       import :: c_int, c_int64_t
       integer(c_int64_t), intent(out) :: net
       integer(c_int) :: rc
     end function
     function athena_cuda_network_destroy(net) &
          bind(C, name="athena_cuda_network_destroy") result(rc)
       import :: c_int, c_int64_t
       integer(c_int64_t), value :: net
       integer(c_int) :: rc
     end function
     function athena_cuda_network_add(net, layer) bind(C, name="athena_cuda_network_add") result(rc)
       import :: c_int, c_int64_t
       integer(c_int64_t), value :: net, layer
       integer(c_int) :: rc
     end function
     !! network%add(layer, input_list, operator) (athena_network_sub.f90:764-830): the ids are
     !! passed as the caller wrote them (0 = input layer, k > 0, k < 0); operator 1 = concatenate
     function athena_cuda_network_add_inputs(net, layer, num_inputs, input_list, merge_operator) &
          bind(C, name="athena_cuda_network_add_inputs") result(rc)
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int64_t), value :: net, layer
       integer(c_int32_t), value :: num_inputs, merge_operator
       integer(c_int32_t), intent(in) :: input_list(*)
       integer(c_int) :: rc
     end function
     function athena_cuda_network_compile(net, optimiser) &
          bind(C, name="athena_cuda_network_compile") result(rc)
       import :: c_int, c_int64_t, athena_optimiser_desc
       integer(c_int64_t), value :: net
       type(athena_optimiser_desc), intent(in) :: optimiser
       integer(c_int) :: rc
     end function
     function athena_cuda_network_num_params(net, n) &
          bind(C, name="athena_cuda_network_num_params") result(rc)
       import :: c_int, c_int64_t
       integer(c_int64_t), value :: net
       integer(c_int64_t), intent(out) :: n
       integer(c_int) :: rc
     end function
     function athena_cuda_network_set_params(net, host, n) &
          bind(C, name="athena_cuda_network_set_params") result(rc)
       import :: c_int, c_int64_t, c_float
       integer(c_int64_t), value :: net, n
       real(c_float), intent(in) :: host(*)
       integer(c_int) :: rc
     end function
     function athena_cuda_network_get_params(net, host, n) &
          bind(C, name="athena_cuda_network_get_params") result(rc)
       import :: c_int, c_int64_t, c_float
       integer(c_int64_t), value :: net, n
       real(c_float), intent(out) :: host(*)
       integer(c_int) :: rc
     end function
     function athena_cuda_network_get_gradients(net, host, n) &
          bind(C, name="athena_cuda_network_get_gradients") result(rc)
       import :: c_int, c_int64_t, c_float
       integer(c_int64_t), value :: net, n
       real(c_float), intent(out) :: host(*)
       integer(c_int) :: rc
     end function
     function athena_cuda_network_set_learning_rate(net, lr) &
          bind(C, name="athena_cuda_network_set_learning_rate") result(rc)
       import :: c_int, c_int64_t, c_float
       integer(c_int64_t), value :: net
       real(c_float), value :: lr
       integer(c_int) :: rc
     end function
     function athena_cuda_network_set_iteration(net, iteration) &
          bind(C, name="athena_cuda_network_set_iteration") result(rc)
       import :: c_int, c_int64_t
       integer(c_int64_t), value :: net, iteration
       integer(c_int) :: rc
     end function
     function athena_cuda_network_forward(net, batch, vertex_features, edge_features, output, &
          mem) bind(C, name="athena_cuda_network_forward") result(rc)
       import :: c_int, c_int32_t, c_int64_t, c_ptr
       integer(c_int64_t), value :: net, batch
       type(c_ptr), value :: vertex_features, edge_features, output
       integer(c_int32_t), value :: mem
       integer(c_int) :: rc
     end function
     function athena_cuda_network_train_step(net, batch, vertex_features, edge_features, &
          target, mem, global_batch, loss) &
          bind(C, name="athena_cuda_network_train_step") result(rc)
       import :: c_int, c_int32_t, c_int64_t, c_ptr
       integer(c_int64_t), value :: net, batch
       type(c_ptr), value :: vertex_features, edge_features, target
       integer(c_int32_t), value :: mem, global_batch
       type(c_ptr), value :: loss   ! c_loc(real(c_float)) or c_null_ptr
       integer(c_int) :: rc
     end function
     function athena_cuda_network_loss_and_gradients(net, batch, vertex_features, &
          edge_features, target, mem, global_batch, loss) &
          bind(C, name="athena_cuda_network_loss_and_gradients") result(rc)
       import :: c_int, c_int32_t, c_int64_t, c_ptr
       integer(c_int64_t), value :: net, batch
       type(c_ptr), value :: vertex_features, edge_features, target
       integer(c_int32_t), value :: mem, global_batch
       type(c_ptr), value :: loss
       integer(c_int) :: rc
     end function
     function athena_cuda_network_update(net) bind(C, name="athena_cuda_network_update") result(rc)
       import :: c_int, c_int64_t
       integer(c_int64_t), value :: net
       integer(c_int) :: rc
     end function
     function athena_cuda_network_last_loss(net, loss) &
          bind(C, name="athena_cuda_network_last_loss") result(rc)
       import :: c_int, c_int64_t, c_float
       integer(c_int64_t), value :: net
       real(c_float), intent(out) :: loss
       integer(c_int) :: rc
     end function

     ! ---- data parallelism ---------------------------------------------------------
     function athena_cuda_comm_unique_id(id) bind(C, name="athena_cuda_comm_unique_id") result(rc)
       import :: c_int, c_char
       character(kind=c_char), intent(out) :: id(128)
       integer(c_int) :: rc
     end function
     function athena_cuda_comm_init(world_size, rank, id) &
          bind(C, name="athena_cuda_comm_init") result(rc)
       import :: c_int, c_int32_t, c_char
       integer(c_int32_t), value :: world_size, rank
       character(kind=c_char), intent(in) :: id(128)
       integer(c_int) :: rc
     end function
     function athena_cuda_comm_destroy() bind(C, name="athena_cuda_comm_destroy") result(rc)
       import :: c_int
       integer(c_int) :: rc
     end function
     function athena_cuda_comm_info(world_size, rank) &
          bind(C, name="athena_cuda_comm_info") result(rc)
       import :: c_int, c_int32_t
       integer(c_int32_t), intent(out) :: world_size, rank
       integer(c_int) :: rc
     end function
     function athena_cuda_comm_p2p_export(handle) &
          bind(C, name="athena_cuda_comm_p2p_export") result(rc)
       import :: c_int, c_char
       character(kind=c_char), intent(out) :: handle(128)
       integer(c_int) :: rc
     end function
     function athena_cuda_comm_p2p_import(world_size, rank, handles) &
          bind(C, name="athena_cuda_comm_p2p_import") result(rc)
       import :: c_int, c_int32_t, c_char
       integer(c_int32_t), value :: world_size, rank
       character(kind=c_char), intent(in) :: handles(*)   ! world_size x 128 bytes, rank order
       integer(c_int) :: rc
     end function
     function athena_cuda_shard_graphs(num_graphs, entries_per_graph, world_size, first_graph) &
          bind(C, name="athena_cuda_shard_graphs") result(rc)
       import :: c_int, c_int32_t, c_int64_t
       integer(c_int32_t), value :: num_graphs, world_size
       integer(c_int64_t), intent(in) :: entries_per_graph(*)
       integer(c_int32_t), intent(out) :: first_graph(*)   ! world_size + 1
       integer(c_int) :: rc
     end function
  end interface

contains

  !> Thread-local message of the last failing call as a Fortran string.
  function athena_cuda_last_error() result(msg)
    character(len=:), allocatable :: msg
    type(c_ptr) :: p
    character(kind=c_char), pointer :: s(:)
    integer :: n
    p = athena_cuda_last_error_c()
    call c_f_pointer(p, s, [1024])
    n = 0
    do while (n < 1024)
       if (s(n + 1) == c_null_char) exit
       n = n + 1
    end do
    allocate(character(len=n) :: msg)
    msg = transfer(s(1:n), msg)
  end function athena_cuda_last_error

  !> Map a non-zero status to coreutils' stop_program(msg), the reference's
  !> error convention (e.g. athena_kipf_msgpass_layer.f90:271-274).
  subroutine athena_cuda_check(rc)
    use coreutils, only: stop_program
    integer(c_int), intent(in) :: rc
    if (rc /= ATHENA_OK) call stop_program("libathena_cuda: " // athena_cuda_last_error())
  end subroutine athena_cuda_check

end module athena__cuda_bindings
