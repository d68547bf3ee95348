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
!> compiled in the development image (no Fortran compiler there); the C ABI it
!> calls is what tests/ exercise through ctypes in the same call order.
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
