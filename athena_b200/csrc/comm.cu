// Data-parallel plumbing: one process per GPU, graphs sharded across ranks,
// one in-stream NCCL all-reduce over the flat gradient vector (+ the loss
// slot).  The reference has no collective at all; its semantic template is
// network_type%reduce (athena_network_sub.f90:36-62), which sums the
// gradients of two replicas.
//
// NCCL is resolved with dlopen at comm_init time so that libathena_cuda has
// no link-time dependency on it (single-GPU users never load it, and inside
// a PyTorch process the already-loaded libnccl.so.2 is reused).
#include <dlfcn.h>

#include "athena_internal.h"

namespace athena {

// Minimal mirror of the NCCL C API (stable since NCCL 2.0).
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7 };
enum { ncclSum = 0 };

struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
};
static Nccl g_nccl;

static int nccl_load() {
  if (g_nccl.lib) return ATHENA_OK;
  const char* names[] = {getenv("ATHENA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    if (!n) continue;
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  ATH_REQUIRE(lib, ATHENA_ERR_COMM, "cannot dlopen libnccl.so.2 (%s); set ATHENA_NCCL_LIB",
              dlerror());
#define ATH_SYM(field, name)                                                   \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name));   \
  ATH_REQUIRE(g_nccl.field, ATHENA_ERR_COMM, "libnccl: missing symbol %s", name)
  ATH_SYM(GetUniqueId, "ncclGetUniqueId");
  ATH_SYM(CommInitRank, "ncclCommInitRank");
  ATH_SYM(AllReduce, "ncclAllReduce");
  ATH_SYM(CommDestroy, "ncclCommDestroy");
  ATH_SYM(GetErrorString, "ncclGetErrorString");
#undef ATH_SYM
  g_nccl.lib = lib;
  return ATHENA_OK;
}

#define ATH_NCCL(expr)                                                                   \
  do {                                                                                   \
    int r_ = (expr);                                                                     \
    if (r_ != ncclSuccess) {                                                             \
      set_error("%s: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
      return ATHENA_ERR_COMM;                                                            \
    }                                                                                    \
  } while (0)

int comm_world_size() { return (g_nccl.comm || p2p().ready) ? g_nccl.world : 1; }

P2PState& p2p() {
  static P2PState s;
  return s;
}
static void* g_p2p_local_x = nullptr;
static void* g_p2p_local_f = nullptr;
constexpr size_t P2P_CAP_FLOATS = size_t(1) << 20;  // 4 MB per slot: any msgpass network here

int comm_allreduce_sum(float* buf, int64_t n) {
  if (!g_nccl.comm || g_nccl.world == 1 || n == 0) return ATHENA_OK;
  ATH_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, g_nccl.comm, ctx().stream));
  ctx().launches.fetch_add(1, std::memory_order_relaxed);
  return ATHENA_OK;
}

}  // namespace athena

using namespace athena;

ATHENA_API int athena_cuda_comm_unique_id(char id[ATHENA_COMM_ID_BYTES]) {
  ATH_REQUIRE(id, ATHENA_ERR_ARG, "comm_unique_id: null");
  ATH_TRY(ensure_init());
  ATH_TRY(nccl_load());
  ncclUniqueId u;
  ATH_NCCL(g_nccl.GetUniqueId(&u));
  static_assert(sizeof(u) == ATHENA_COMM_ID_BYTES, "id size");
  memcpy(id, &u, sizeof(u));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_comm_init(int32_t world_size, int32_t rank,
                                     const char id[ATHENA_COMM_ID_BYTES]) {
  ATH_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, ATHENA_ERR_ARG,
              "comm_init: bad world_size/rank %d/%d", world_size, rank);
  ATH_TRY(ensure_init());
  ATH_REQUIRE(!g_nccl.comm, ATHENA_ERR_STATE, "comm_init: communicator already exists");
  g_nccl.world = world_size;
  g_nccl.rank = rank;
  if (world_size == 1) return ATHENA_OK;
  ATH_REQUIRE(id, ATHENA_ERR_ARG, "comm_init: null id");
  ATH_TRY(nccl_load());
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ATH_NCCL(g_nccl.CommInitRank(&g_nccl.comm, world_size, u, rank));
  return ATHENA_OK;
}

// ---- peer-memory exchange: CUDA IPC set-up ------------------------------------------
// handle = { cudaIpcMemHandle_t of the staging buffer, cudaIpcMemHandle_t of the flags }
ATHENA_API int athena_cuda_comm_p2p_export(char handle[ATHENA_P2P_HANDLE_BYTES]) {
  ATH_REQUIRE(handle, ATHENA_ERR_ARG, "comm_p2p_export: null");
  ATH_TRY(ensure_init());
  static_assert(2 * sizeof(cudaIpcMemHandle_t) == ATHENA_P2P_HANDLE_BYTES, "handle size");
  if (!g_p2p_local_x) {
    // plain cudaMalloc (not the pool): IPC handles cover whole allocations
    ATH_CUDA(cudaMalloc(&g_p2p_local_x, 2 * P2P_CAP_FLOATS * sizeof(float)));
    ATH_CUDA(cudaMalloc(&g_p2p_local_f, P2P_FLAG_WORDS * sizeof(uint32_t)));
    ATH_CUDA(cudaMemset(g_p2p_local_x, 0, 2 * P2P_CAP_FLOATS * sizeof(float)));
    ATH_CUDA(cudaMemset(g_p2p_local_f, 0, P2P_FLAG_WORDS * sizeof(uint32_t)));
  }
  cudaIpcMemHandle_t hx, hf;
  ATH_CUDA(cudaIpcGetMemHandle(&hx, g_p2p_local_x));
  ATH_CUDA(cudaIpcGetMemHandle(&hf, g_p2p_local_f));
  memcpy(handle, &hx, sizeof(hx));
  memcpy(handle + sizeof(hx), &hf, sizeof(hf));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_comm_p2p_import(int32_t world_size, int32_t rank, const char* handles) {
  ATH_REQUIRE(handles && world_size >= 2 && world_size <= P2P_MAX_WORLD && rank >= 0 &&
                  rank < world_size,
              ATHENA_ERR_ARG, "comm_p2p_import: bad argument (world %d, rank %d, max world %d)",
              world_size, rank, P2P_MAX_WORLD);
  ATH_REQUIRE(g_p2p_local_x, ATHENA_ERR_STATE, "comm_p2p_import: call comm_p2p_export first");
  P2PState& P = p2p();
  ATH_REQUIRE(!P.ready, ATHENA_ERR_STATE, "comm_p2p_import: already initialised");
  for (int r = 0; r < world_size; ++r) {
    if (r == rank) {
      P.xbuf[r] = static_cast<float*>(g_p2p_local_x);
      P.flags[r] = static_cast<uint32_t*>(g_p2p_local_f);
      continue;
    }
    cudaIpcMemHandle_t hx, hf;
    memcpy(&hx, handles + (size_t)r * ATHENA_P2P_HANDLE_BYTES, sizeof(hx));
    memcpy(&hf, handles + (size_t)r * ATHENA_P2P_HANDLE_BYTES + sizeof(hx), sizeof(hf));
    void *px = nullptr, *pf = nullptr;
    ATH_CUDA(cudaIpcOpenMemHandle(&px, hx, cudaIpcMemLazyEnablePeerAccess));
    ATH_CUDA(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
    P.xbuf[r] = static_cast<float*>(px);
    P.flags[r] = static_cast<uint32_t*>(pf);
  }
  P.world = world_size;
  P.rank = rank;
  P.cap = P2P_CAP_FLOATS;
  P.epoch = 0;
  g_nccl.world = world_size;
  g_nccl.rank = rank;
  P.ready = true;
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_comm_destroy(void) {
  P2PState& P = p2p();
  if (P.ready) {
    if (ctx().ready) cudaStreamSynchronize(ctx().stream);
    for (int r = 0; r < P.world; ++r) {
      if (r == P.rank) continue;
      cudaIpcCloseMemHandle(P.xbuf[r]);
      cudaIpcCloseMemHandle(P.flags[r]);
    }
    P = P2PState{};
  }
  if (g_nccl.comm) {
    if (ctx().ready) cudaStreamSynchronize(ctx().stream);
    g_nccl.CommDestroy(g_nccl.comm);
    g_nccl.comm = nullptr;
  }
  g_nccl.world = 1;
  g_nccl.rank = 0;
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_comm_info(int32_t* world_size, int32_t* rank) {
  if (world_size) *world_size = g_nccl.world;
  if (rank) *rank = g_nccl.rank;
  return ATHENA_OK;
}

// Contiguous partition of B graphs over `world_size` ranks, balanced by CSR
// entries (the per-graph cost of aggregation): boundary r is placed at the
// first graph whose running entry count reaches r/world of the total.
// Deterministic, pure host code; every rank computes the same answer.
ATHENA_API int athena_cuda_shard_graphs(int32_t num_graphs, const int64_t* entries_per_graph,
                                        int32_t world_size, int32_t* first_graph) {
  ATH_REQUIRE(num_graphs >= 0 && world_size >= 1 && first_graph &&
                  (entries_per_graph || num_graphs == 0),
              ATHENA_ERR_ARG, "shard_graphs: bad argument");
  int64_t total = 0;
  for (int32_t s = 0; s < num_graphs; ++s) {
    ATH_REQUIRE(entries_per_graph[s] >= 0, ATHENA_ERR_ARG, "shard_graphs: negative weight");
    total += entries_per_graph[s] + 1;  // +1 keeps empty graphs spread out
  }
  first_graph[0] = 0;
  int64_t run = 0;
  int32_t s = 0;
  for (int32_t r = 1; r < world_size; ++r) {
    // smallest s with run >= total * r / world
    while (s < num_graphs && run * world_size < total * r) {
      run += entries_per_graph[s] + 1;
      ++s;
    }
    // When there are at least as many graphs as ranks every rank gets one: boundary r is
    // kept inside [r, num_graphs - (world - r)].  (A skewed batch -- one graph holding most
    // of the entries -- would otherwise leave ranks without work, and athena_cuda_batch_create
    // rejects an empty batch while the peers wait in the gradient exchange.)
    if (num_graphs >= world_size) {
      const int32_t lo = r, hi = num_graphs - (world_size - r);
      while (s < lo) {
        run += entries_per_graph[s] + 1;
        ++s;
      }
      while (s > hi) {
        --s;
        run -= entries_per_graph[s] + 1;
      }
    }
    first_graph[r] = s;
  }
  first_graph[world_size] = num_graphs;
  return ATHENA_OK;
}
