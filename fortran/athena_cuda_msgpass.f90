!> Glue between athena's message-passing layer types and libathena_cuda.
!>
!> These are the bodies a maintainer drops into the reference (INTEGRATION.md
!> lists the exact replacement points).  The public Fortran API does not change:
!>   set_graph_msgpass        athena_msgpass_layer_sub.f90:144-174
!>   update_message_kipf      athena_kipf_msgpass_layer.f90:915-959
!>   update_message_duvenaud  athena_duvenaud_msgpass_layer.f90:755-817
!>   update_readout_duvenaud  athena_duvenaud_msgpass_layer.f90:822-859
!>   network%compile          optimiser object -> athena_optimiser_desc (cuda_optimiser_desc)
!>   network%update           athena_network_sub.f90:2816-2929
!>   train batch loop         athena_network_sub.f90:3611-3670
!>
!> Written against athena v2.1.1 / graphstruc v0.2.1 / diffstruc v1.2.0.  Not
!> compiled in the development image (no Fortran compiler there); the same ABI calls in the
!> same order are compiled and run from C by tests/c_driver/fake_athena.c (per-sample
!> gradient staging included) and from Python by tests/.
module athena__cuda_msgpass
  use, intrinsic :: iso_c_binding
  use coreutils, only: real32
  use graphstruc, only: graph_type
  use diffstruc, only: array_type
  use athena__cuda_bindings
  implicit none
  private

  public :: cuda_graph_batch_type
  public :: cuda_msgpass_forward, cuda_msgpass_backward
  public :: cuda_register_output_nodes, get_partial_cuda_layer_val
  public :: cuda_network_update, cuda_train_batch
  public :: cuda_network_add, cuda_batch_from_edges
  public :: cuda_optimiser_desc

  !> Device twin of graph(:) -- built once per mini-batch and shared by every
  !> message-passing layer of the network (replaces the per-layer deep copies
  !> of set_graph_msgpass).
  type :: cuda_graph_batch_type
     integer(c_int64_t) :: handle = 0_c_int64_t
     integer :: num_graphs = 0
     integer :: num_vertices = 0   !! sum over samples
     integer :: num_edges = 0
   contains
     procedure :: create => batch_create
     procedure :: destroy => batch_destroy
  end type cuda_graph_batch_type

contains

  !> Pack graph(:)%adj_ia / adj_ja (both stay 1-based) and build the device CSR/CSC.
  subroutine batch_create(this, graph)
    class(cuda_graph_batch_type), intent(inout) :: this
    type(graph_type), dimension(:), intent(in) :: graph
    integer(c_int32_t), allocatable :: nv(:), ne(:), nz(:), ia(:), ja(:,:)
    integer :: s, ia_pos, ja_pos, b

    call this%destroy()
    b = size(graph)
    allocate(nv(b), ne(b), nz(b))
    do s = 1, b
       nv(s) = graph(s)%num_vertices
       ne(s) = graph(s)%num_edges
       nz(s) = size(graph(s)%adj_ja, 2)
    end do
    allocate(ia(sum(nv) + b), ja(2, sum(nz)))
    ia_pos = 0
    ja_pos = 0
    do s = 1, b
       ia(ia_pos + 1 : ia_pos + nv(s) + 1) = graph(s)%adj_ia(1 : nv(s) + 1)
       ja(:, ja_pos + 1 : ja_pos + nz(s)) = graph(s)%adj_ja(:, 1 : nz(s))
       ia_pos = ia_pos + nv(s) + 1
       ja_pos = ja_pos + nz(s)
    end do
    ! validate = 1: "graph adjacency matrix has indices greater than the number of
    ! vertices" (athena_duvenaud_msgpass_layer.f90:632-639) surfaces here
    call athena_cuda_check(athena_cuda_batch_create(this%handle, int(b, c_int32_t), nv, ne, nz, &
         ia, ja, ATHENA_MEM_HOST, 1_c_int32_t))
    this%num_graphs = b
    this%num_vertices = sum(nv)
    this%num_edges = sum(ne)
  end subroutine batch_create

  subroutine batch_destroy(this)
    class(cuda_graph_batch_type), intent(inout) :: this
    if (this%handle /= 0_c_int64_t) then
       call athena_cuda_check(athena_cuda_batch_destroy(this%handle))
       this%handle = 0_c_int64_t
    end if
  end subroutine batch_destroy

  !> graph%generate_adjacency(index_list) [+ graph%add_self_loops()] for a whole mini-batch on
  !> the device: what the readers do per sample before set_graph
  !> (example_library/src/mod_read_chemical_graphs.f90:275-276, example/msgpass_euler/src/
  !> mod_read_euler.f90:48, example/msgpass_chemical/src/main.f90:108).  index_list is the
  !> concatenation of the samples' index_list(2, num_edges(s)).
  subroutine cuda_batch_from_edges(this, num_vertices, num_edges, index_list, self_loops)
    class(cuda_graph_batch_type), intent(inout) :: this
    integer(c_int32_t), intent(in) :: num_vertices(:), num_edges(:)
    integer(c_int32_t), intent(in) :: index_list(:,:)
    logical, intent(in) :: self_loops
    call this%destroy()
    call athena_cuda_check(athena_cuda_batch_create_from_edges(this%handle, &
         int(size(num_vertices), c_int32_t), num_vertices, num_edges, index_list, c_null_ptr, &
         merge(1_c_int32_t, 0_c_int32_t, self_loops), ATHENA_MEM_HOST, 1_c_int32_t))
    this%num_graphs = size(num_vertices)
    this%num_vertices = sum(num_vertices)
    this%num_edges = sum(num_edges)
  end subroutine cuda_batch_from_edges

  !> network%add(layer, input_list, operator) (athena_network_sub.f90:764-830) for a device
  !> layer: the ids go through as the caller wrote them (0 = input layer, k > 0 the k-th added
  !> layer, k < 0 counted back from this one); "||" / "concat" / "concatenate" / "append" = 1
  !> as in :808-811, anything the device path does not implement stops like the reference's
  !> "invalid operator".
  subroutine cuda_network_add(net_handle, layer_handle, input_list, operator)
    integer(c_int64_t), intent(in) :: net_handle, layer_handle
    integer, dimension(:), optional, intent(in) :: input_list
    character(*), optional, intent(in) :: operator
    integer(c_int32_t) :: op
    if (.not. present(input_list)) then
       call athena_cuda_check(athena_cuda_network_add(net_handle, layer_handle))
       return
    end if
    op = 1_c_int32_t
    if (present(operator)) then
       select case (trim(operator))
       case ("||", "concat", "concatenate", "append")
          op = 1_c_int32_t
       case ("+", "add")
          op = 2_c_int32_t
       case default
          op = 0_c_int32_t
       end select
    end if
    call athena_cuda_check(athena_cuda_network_add_inputs(net_handle, layer_handle, &
         int(size(input_list), c_int32_t), int(input_list, c_int32_t), op))
  end subroutine cuda_network_add

  !> Body of update_message_* + update_readout_*: input(1,s)%val(F,V_s) and
  !> input(2,s)%val(Fe,E_s) are concatenated over s (already contiguous per
  !> sample in Fortran's column-major order) and sent through one call; the
  !> result is split back into this%output(1,s)%val (Kipf) or stored as
  !> output(1,1)%val(num_outputs, batch) (Duvenaud).
  subroutine cuda_msgpass_forward(layer_handle, batch, input, num_out_features, output, &
       graph_level)
    integer(c_int64_t), intent(in) :: layer_handle
    type(cuda_graph_batch_type), intent(in) :: batch
    class(array_type), dimension(:,:), intent(in) :: input
    integer, intent(in) :: num_out_features
    real(real32), allocatable, target, intent(out) :: output(:,:)
    logical, intent(in) :: graph_level
    real(real32), allocatable, target :: x(:,:), e(:,:)
    type(c_ptr) :: e_ptr
    integer :: s, fv, fe, v0, e0, nvs, nes

    fv = size(input(1,1)%val, 1)
    allocate(x(fv, batch%num_vertices))
    v0 = 0
    do s = 1, size(input, 2)
       nvs = size(input(1,s)%val, 2)
       x(:, v0 + 1 : v0 + nvs) = input(1,s)%val
       v0 = v0 + nvs
    end do
    e_ptr = c_null_ptr
    if (size(input, 1) >= 2 .and. graph_level) then
       fe = size(input(2,1)%val, 1)
       allocate(e(fe, batch%num_edges))
       e0 = 0
       do s = 1, size(input, 2)
          nes = size(input(2,s)%val, 2)
          e(:, e0 + 1 : e0 + nes) = input(2,s)%val
          e0 = e0 + nes
       end do
       e_ptr = c_loc(e)
    end if
    if (graph_level) then
       allocate(output(num_out_features, batch%num_graphs))
    else
       allocate(output(num_out_features, batch%num_vertices))
    end if
    call athena_cuda_check(athena_cuda_layer_forward(layer_handle, batch%handle, c_loc(x), &
         e_ptr, c_loc(output), ATHENA_MEM_HOST))
  end subroutine cuda_msgpass_forward

  !> Reverse sweep: called from the get_partial_left_val callback of the single
  !> autodiff node that stands for the whole layer (INTEGRATION.md section 3), with
  !> the upstream gradient of the layer output.  Parameter gradients accumulate on
  !> the device; grad_input is only requested when the producer requires a gradient
  !> (the input layer never does, athena_input_layer.f90:541).
  subroutine cuda_msgpass_backward(layer_handle, batch, upstream_grad, grad_input)
    integer(c_int64_t), intent(in) :: layer_handle
    type(cuda_graph_batch_type), intent(in) :: batch
    real(real32), dimension(:,:), intent(in), target :: upstream_grad
    real(real32), dimension(:,:), intent(out), target, optional :: grad_input
    type(c_ptr) :: gi
    gi = c_null_ptr
    if (present(grad_input)) gi = c_loc(grad_input)
    call athena_cuda_check(athena_cuda_layer_backward(layer_handle, batch%handle, &
         c_loc(upstream_grad), gi, ATHENA_MEM_HOST))
  end subroutine cuda_msgpass_backward

  !> The autodiff seam.  After the device forward every output(1,s) (Kipf) or output(1,1)
  !> (Duvenaud, one [num_outputs, batch] array) is turned into a leaf-like node of the
  !> expression graph whose operand is the layer's first parameter array -- the operand
  !> that requires a gradient, so loss%grad_reverse (athena_network_sub.f90:3645) visits the
  !> node and calls its get_partial_left_val with the upstream gradient of THAT sample.
  !> The two handles and the sample index travel in the node's integer `indices` array
  !> (the field kipf_propagate uses for adj_ia, athena_diffstruc_extd_sub_kipf.f90:48).
  !>   indices(1:2) = layer handle, indices(3:4) = batch handle (int64 as two int32),
  !>   indices(5)   = 0-based sample index, or -1 for a graph-level [num_outputs, batch] array
  subroutine cuda_register_output_nodes(output, params1, layer_handle, batch, graph_level)
    class(array_type), dimension(:,:), intent(inout), target :: output
    type(array_type), intent(in), target :: params1
    integer(c_int64_t), intent(in) :: layer_handle
    type(cuda_graph_batch_type), intent(in) :: batch
    logical, intent(in) :: graph_level
    integer :: s
    integer(c_int32_t) :: h(2)

    do s = 1, size(output, 2)
       if (allocated(output(1,s)%indices)) deallocate(output(1,s)%indices)
       allocate(output(1,s)%indices(5))
       h = transfer(layer_handle, h)
       output(1,s)%indices(1:2) = h
       h = transfer(batch%handle, h)
       output(1,s)%indices(3:4) = h
       output(1,s)%indices(5) = merge(-1, s - 1, graph_level)
       output(1,s)%get_partial_left_val => get_partial_cuda_layer_val
       output(1,s)%left_operand => params1
       output(1,s)%owns_left_operand = .false.
       output(1,s)%requires_grad = .true.
       output(1,s)%is_forward = params1%is_forward
       output(1,s)%operation = 'cuda_msgpass'
       output(1,s)%is_temporary = .false.
    end do
  end subroutine cuda_register_output_nodes

  !> get_partial_left_val of those nodes (same signature as
  !> get_partial_kipf_propagate_left_val, athena_diffstruc_extd_sub_kipf.f90:85-95).
  !> The upstream gradient of the sample is staged on the device; when the last sample of
  !> the batch arrives the library runs the reverse sweep of the whole batch and
  !> accumulates the parameter gradients there.  `output` is the gradient diffstruc adds
  !> to params(1)%grad on the host: zero -- the device holds the gradients, and the
  !> replacement body of network%update (cuda_network_update) steps from them.
  !> Errors cannot stop a pure procedure; a failed call leaves the error text for the next
  !> checked call (athena_cuda_network_update reports ATHENA_ERR_STATE).
  pure subroutine get_partial_cuda_layer_val(this, upstream_grad, output)
    class(array_type), intent(in) :: this
    real(real32), dimension(:,:), intent(in) :: upstream_grad
    real(real32), dimension(:,:), intent(out) :: output
    integer(c_int64_t) :: layer_handle, batch_handle
    integer(c_int32_t) :: sample
    integer(c_int) :: rc
    integer :: s

    layer_handle = transfer(this%indices(1:2), layer_handle)
    batch_handle = transfer(this%indices(3:4), batch_handle)
    output = 0._real32
    if (this%indices(5) >= 0) then
       sample = int(this%indices(5), c_int32_t)
       rc = athena_cuda_layer_backward_stage_pure(layer_handle, batch_handle, sample, &
            upstream_grad, int(size(upstream_grad), c_int64_t))
    else
       ! graph-level output [num_outputs, batch]: column s is sample s
       do s = 1, size(upstream_grad, 2)
          rc = athena_cuda_layer_backward_stage_pure(layer_handle, batch_handle, &
               int(s - 1, c_int32_t), upstream_grad(:, s), &
               int(size(upstream_grad, 1), c_int64_t))
       end do
    end if
  end subroutine get_partial_cuda_layer_val

  !> Body of network%update (athena_network_sub.f90:2816-2929) for a network whose learnable
  !> layers live on the device: the iteration counter, the learning-rate schedule and the
  !> clip / minimise / zero-gradients sequence keep their order; the flat parameter and
  !> gradient vectors never visit the host.
  subroutine cuda_network_update(net_handle, learning_rate, iteration)
    integer(c_int64_t), intent(in) :: net_handle
    real(real32), intent(in) :: learning_rate   !! lr_decay%get_lr(lr0, iter), host side
    integer, optional, intent(in) :: iteration  !! this%optimiser%iter when lr_decay iterates
                                                !! per epoch (athena_network_sub.f90:2834-2838)
    if (present(iteration)) call athena_cuda_check( &
         athena_cuda_network_set_iteration(net_handle, int(iteration, c_int64_t)))
    call athena_cuda_check(athena_cuda_network_set_learning_rate(net_handle, learning_rate))
    call athena_cuda_check(athena_cuda_network_update(net_handle))
  end subroutine cuda_network_update

  !> One iteration of the batch loop of network%train (athena_network_sub.f90:3611-3670) for a
  !> pure message-passing network: get_sample + set_graph (batch%create), forward, loss_eval
  !> (MSE), grad_reverse, update -- one library call; returns batch_loss (:3650).
  !>   graph    this%input_graph(start_index:end_index) of the batch
  !>   target   expected output of the batch: Kipf-last [F_T, sum of vertices]; Duvenaud-last
  !>            [num_outputs, batch]
  function cuda_train_batch(net_handle, graph, target, global_batch, learning_rate) &
       result(batch_loss)
    integer(c_int64_t), intent(in) :: net_handle
    type(graph_type), dimension(:), intent(in) :: graph
    real(real32), dimension(:,:), intent(in), target :: target
    integer, intent(in) :: global_batch
    real(real32), intent(in) :: learning_rate
    real(real32) :: batch_loss
    type(cuda_graph_batch_type) :: batch
    real(real32), allocatable, target :: x(:,:), e(:,:)
    real(c_float), target :: loss
    type(c_ptr) :: e_ptr
    integer :: s, v0, e0, fv, fe

    call batch%create(graph)
    fv = graph(1)%num_vertex_features
    fe = graph(1)%num_edge_features
    allocate(x(fv, batch%num_vertices))
    v0 = 0
    do s = 1, size(graph)
       x(:, v0 + 1 : v0 + graph(s)%num_vertices) = graph(s)%vertex_features
       v0 = v0 + graph(s)%num_vertices
    end do
    e_ptr = c_null_ptr
    if (fe > 0) then
       allocate(e(fe, batch%num_edges))
       e0 = 0
       do s = 1, size(graph)
          e(:, e0 + 1 : e0 + graph(s)%num_edges) = graph(s)%edge_features
          e0 = e0 + graph(s)%num_edges
       end do
       e_ptr = c_loc(e)
    end if
    call athena_cuda_check(athena_cuda_network_set_learning_rate(net_handle, learning_rate))
    call athena_cuda_check(athena_cuda_network_train_step(net_handle, batch%handle, c_loc(x), &
         e_ptr, c_loc(target), ATHENA_MEM_HOST, int(global_batch, c_int32_t), c_loc(loss)))
    batch_loss = loss
    call batch%destroy()
  end function cuda_train_batch

  !> network%compile: describe athena's optimiser object to athena_cuda_network_compile.
  !> The decayed learning rate is passed per update with
  !> athena_cuda_network_set_learning_rate (lr_decay%get_lr stays on the host).
  !> Optimisers / regularisers outside the device path stop the program, as an unknown
  !> option does in the reference.
  function cuda_optimiser_desc(optimiser) result(d)
    use athena__optimiser, only: base_optimiser_type, sgd_optimiser_type, &
         adam_optimiser_type, rmsprop_optimiser_type, adagrad_optimiser_type
    use athena__regulariser, only: l1_regulariser_type, l2_regulariser_type, &
         l1l2_regulariser_type
    use coreutils, only: stop_program
    class(base_optimiser_type), intent(in) :: optimiser
    type(athena_optimiser_desc) :: d

    d%kind = ATHENA_OPT_SGD
    d%learning_rate = optimiser%learning_rate
    d%beta1 = 0.9_c_float;  d%beta2 = 0.999_c_float;  d%epsilon = 1.e-8_c_float
    d%momentum = 0._c_float;  d%nesterov = 0
    select type(optimiser)
    type is (sgd_optimiser_type)          ! athena_optimiser.f90:119-132
       d%momentum = optimiser%momentum
       d%nesterov = merge(1, 0, optimiser%nesterov)
    type is (adam_optimiser_type)         ! :236-250
       d%kind = ATHENA_OPT_ADAM
       d%beta1 = optimiser%beta1;  d%beta2 = optimiser%beta2;  d%epsilon = optimiser%epsilon
    type is (rmsprop_optimiser_type)      ! :160-173
       d%kind = ATHENA_OPT_RMSPROP
       d%beta1 = optimiser%beta;  d%epsilon = optimiser%epsilon
    type is (adagrad_optimiser_type)      ! :199-210
       d%kind = ATHENA_OPT_ADAGRAD
       d%epsilon = optimiser%epsilon
    class default
       call stop_program("optimiser type has no device implementation")
    end select
    ! clip_type (athena_clipper.f90:20-30)
    d%clip_min_max = merge(1, 0, optimiser%clip_dict%l_min_max)
    d%clip_min = optimiser%clip_dict%min;  d%clip_max = optimiser%clip_dict%max
    d%clip_norm_on = merge(1, 0, optimiser%clip_dict%l_norm)
    d%clip_norm = optimiser%clip_dict%norm
    ! regulariser (athena_regulariser.f90:40-76)
    d%regulariser = ATHENA_REG_NONE;  d%l1 = 0._c_float;  d%l2 = 0._c_float;  d%l2_decoupled = 1
    if (optimiser%regularisation .and. allocated(optimiser%regulariser)) then
       select type(r => optimiser%regulariser)
       type is (l1_regulariser_type)
          d%regulariser = ATHENA_REG_L1;  d%l1 = r%l1
       type is (l2_regulariser_type)
          d%regulariser = ATHENA_REG_L2;  d%l2 = r%l2
          d%l2_decoupled = merge(1, 0, r%decoupled)
       type is (l1l2_regulariser_type)
          d%regulariser = ATHENA_REG_L1L2;  d%l1 = r%l1;  d%l2 = r%l2
       class default
          call stop_program("regulariser type has no device implementation")
       end select
    end if
  end function cuda_optimiser_desc

end module athena__cuda_msgpass
