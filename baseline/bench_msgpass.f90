!> Driver for baseline/run_reference.sh: the bench.py workload (BASELINE.json configs[1]) through
!> the reference's own public API -- network_type%add / compile / train on graph_type inputs.
!>   Kipf GCN 2-layer (64 -> 64 relu, 64 -> 64), GRAPHS graphs x 64 vertices, every vertex 12
!>   distinct neighbours (circulant offsets) + a self loop, 64 features, MSE, SGD lr 0.01,
!>   one mini-batch of all graphs per epoch (batch_size = GRAPHS), EPOCHS timed epochs.
!> Prints {"impl": "reference-fortran", "value": edges/s, ...}: an edge is one CSR entry.
!> API usage follows test/test_msgpass_network.f90:40-88,249-276 (graph set-up, add, compile,
!> train) and example/msgpass_chemical/src/main.f90:106-109 (add_self_loops).
program bench_msgpass
  use coreutils, only: real32
  use graphstruc, only: graph_type
  use athena
  implicit none
  integer, parameter :: nv = 64, half_degree = 6, nf = 64
  integer :: num_graphs, num_epochs, s, v, k, e, nargs
  integer(8) :: t0, t1, rate, entries
  character(32) :: arg
  type(network_type) :: network
  type(graph_type), allocatable :: graphs(:,:), targets(:,:)
  integer, allocatable :: index_list(:,:)
  real(real32) :: seconds

  num_graphs = 4096
  num_epochs = 3
  nargs = command_argument_count()
  if (nargs >= 1) then
     call get_command_argument(1, arg); read(arg, *) num_graphs
  end if
  if (nargs >= 2) then
     call get_command_argument(2, arg); read(arg, *) num_epochs
  end if

  allocate(graphs(1, num_graphs), targets(1, num_graphs))
  allocate(index_list(2, nv * half_degree))
  entries = 0
  do s = 1, num_graphs
     call graphs(1,s)%set_num_vertices(nv, nf)
     call graphs(1,s)%set_num_edges(nv * half_degree)
     graphs(1,s)%is_sparse = .true.
     call random_number(graphs(1,s)%vertex_features)
     graphs(1,s)%vertex_features = 2._real32 * graphs(1,s)%vertex_features - 1._real32
     e = 0
     do v = 1, nv
        do k = 1, half_degree          ! circulant: v -- v + k (mod nv)
           e = e + 1
           index_list(:, e) = [v, mod(v - 1 + k, nv) + 1]
        end do
     end do
     call graphs(1,s)%generate_adjacency(index_list)
     call graphs(1,s)%add_self_loops()
     entries = entries + size(graphs(1,s)%adj_ja, 2)
     ! graph-level target: same structure, random vertex features of the output width
     targets(1,s) = graphs(1,s)
     call random_number(targets(1,s)%vertex_features)
  end do

  call network%add(kipf_msgpass_layer_type( &
       num_vertex_features = [nf, nf], num_time_steps = 1, activation = 'relu'))
  call network%add(kipf_msgpass_layer_type( &
       num_vertex_features = [nf, nf], num_time_steps = 1, activation = 'none'))
  call network%compile( &
       optimiser = sgd_optimiser_type(learning_rate = 0.01), &
       loss_method = 'mse', accuracy_method = 'mse', metrics = ['loss'], &
       batch_size = num_graphs, verbose = 0)

  ! one untimed epoch (allocation of the autodiff buffers), then the timed ones
  call network%train(graphs, targets, num_epochs = 1, shuffle_batches = .false.)
  call system_clock(t0, rate)
  call network%train(graphs, targets, num_epochs = num_epochs, shuffle_batches = .false.)
  call system_clock(t1)
  seconds = real(t1 - t0, real32) / real(rate, real32)
  write(*,'(A,ES14.6,A,I0,A,I0,A,I0,A,ES12.4,A)') &
       '{"impl": "reference-fortran", "metric": "msgpass train edges/sec (fwd+bwd)", "value": ', &
       real(entries, real32) * real(num_epochs, real32) / seconds, &
       ', "unit": "edges/s", "cores": 1, "graphs": ', num_graphs, ', "entries": ', entries, &
       ', "epochs": ', num_epochs, ', "seconds": ', seconds, '}'
end program bench_msgpass
