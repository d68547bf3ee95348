// Process-wide context: device selection, the library stream, error strings,
// the handle registry, memory helpers and event timers.
#include <cstdlib>
#include <map>

#include "athena_internal.h"

namespace athena {

static thread_local char g_err[1024] = "no error";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int l2_hint_mask() {
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("ATHENA_DEBUG_L2_HINTS");
    m = e ? atoi(e) : L2_HINTS_DEFAULT;
  }
  return m;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("ATHENA_CUDA_DISABLE_PDL");
    on = (e && atoi(e) != 0) ? 0 : 1;
  }
  return on == 1;
}

Context& ctx() {
  static Context c;
  return c;
}

static std::mutex g_mu;
static std::unordered_map<athena_handle_t, std::unique_ptr<Object>> g_objects;
static athena_handle_t g_next = 0x1000;

athena_handle_t register_object(Object* obj) {
  std::lock_guard<std::mutex> lk(g_mu);
  athena_handle_t h = g_next++;
  g_objects[h] = std::unique_ptr<Object>(obj);
  return h;
}

Object* lookup_object(athena_handle_t h, Kind kind) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_objects.find(h);
  if (it == g_objects.end() || it->second->kind != kind) {
    set_error("invalid handle %lld (expected kind %d)", (long long)h, (int)kind);
    return nullptr;
  }
  return it->second.get();
}

int destroy_object(athena_handle_t h, Kind kind) {
  std::unique_ptr<Object> victim;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_objects.find(h);
    if (it == g_objects.end() || it->second->kind != kind) {
      set_error("invalid handle %lld (expected kind %d)", (long long)h, (int)kind);
      return ATHENA_ERR_HANDLE;
    }
    victim = std::move(it->second);
    g_objects.erase(it);
  }
  // the object's device buffers return to the stream-ordered pool (no synchronisation);
  // the copy stream must not touch them before everything queued so far has run
  if (ctx().ready) record_mark();
  victim.reset();
  return ATHENA_OK;
}

// ---- caching device allocator -----------------------------------------------
// Every DevBuf draws from a process-wide free list instead of cudaMalloc /
// cudaFree: a training loop that builds and destroys one graph batch per step
// (network%train, athena_network_sub.f90:3611-3670) would otherwise pay a
// device-wide synchronisation plus an unmap per buffer per step.  All work of
// the library is ordered on one stream, so a block may be handed to its next
// owner as soon as the previous owner lets go of it: kernels of the new owner
// are queued behind the kernels that still read the old contents.
namespace {
struct DevPool {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;
  size_t cached = 0;
};
DevPool& pool() {
  static DevPool* p = new DevPool;  // leaked on purpose: outlives every static DevBuf owner
  return *p;
}
size_t pool_round(size_t bytes) {
  if (bytes < (size_t(1) << 20)) return (bytes + 511) & ~size_t(511);
  return (bytes + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);  // 2 MB pages
}
}  // namespace

void pool_trim() {
  DevPool& P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  if (P.free_blocks.empty()) return;
  if (ctx().ready) {
    // both streams: a host->device copy queued by a call that failed before its consumer was
    // launched may still be writing into a cached block
    cudaStreamSynchronize(ctx().copy_stream);
    cudaStreamSynchronize(ctx().stream);
  }
  for (auto& kv : P.free_blocks) cudaFree(kv.second);
  P.free_blocks.clear();
  P.cached = 0;
}

static int pool_alloc(size_t bytes, void** out, size_t* cap) {
  const size_t want = pool_round(bytes ? bytes : 1);
  DevPool& P = pool();
  {
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.free_blocks.lower_bound(want);
    // accept a cached block that wastes at most 25 % (+1 MB)
    if (it != P.free_blocks.end() && it->first <= want + want / 4 + (size_t(1) << 20)) {
      *out = it->second;
      *cap = it->first;
      P.cached -= it->first;
      P.free_blocks.erase(it);
      return ATHENA_OK;
    }
  }
  cudaError_t e = cudaMalloc(out, want);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    pool_trim();  // give the cached blocks back and retry once
    e = cudaMalloc(out, want);
  }
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes): %s", want, cudaGetErrorString(e));
    return ATHENA_ERR_CUDA;
  }
  *cap = want;
  return ATHENA_OK;
}

static void pool_free(void* p, size_t cap) {
  DevPool& P = pool();
  std::lock_guard<std::mutex> lk(P.mu);
  P.free_blocks.emplace(cap, p);
  P.cached += cap;
}

int DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return ATHENA_OK;
  ATH_TRY(ensure_init());
  if (p) {
    pool_free(p, cap);
    p = nullptr;
    cap = 0;
  }
  return pool_alloc(bytes, &p, &cap);
}

void DevBuf::release() {
  if (p) {
    pool_free(p, cap);
    p = nullptr;
    cap = 0;
  }
}

// ---- copy stream ---------------------------------------------------------------
// Hazards and how they are closed:
//  * a destination may still be read by kernels queued by EARLIER calls (staging buffers are
//    reused, pool blocks are recycled): the copy stream first waits for `mark`, the main-stream
//    position recorded at the end of every compute call and at every batch destruction;
//  * consumers on the main stream wait for the event recorded behind the copies;
//  * athena_cuda_synchronize drains both streams.
int side_begin() {
  Context& c = ctx();
  if (c.mark_valid) ATH_CUDA(cudaStreamWaitEvent(c.copy_stream, c.mark, 0));
  return ATHENA_OK;
}
int side_copy(void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return ATHENA_OK;
  ATH_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx().copy_stream));
  return ATHENA_OK;
}
int side_fence(cudaEvent_t* ev) {
  Context& c = ctx();
  cudaEvent_t e = c.ev_pool[c.ev_next];
  c.ev_next = (c.ev_next + 1) % 32;
  ATH_CUDA(cudaEventRecord(e, c.copy_stream));
  *ev = e;
  return ATHENA_OK;
}
int main_wait(cudaEvent_t ev) {
  if (ev) ATH_CUDA(cudaStreamWaitEvent(ctx().stream, ev, 0));
  return ATHENA_OK;
}
int record_mark() {
  Context& c = ctx();
  ATH_CUDA(cudaEventRecord(c.mark, c.stream));
  c.mark_valid = true;
  return ATHENA_OK;
}

// ---- per-kernel profiling ---------------------------------------------------
struct ProfRec {
  std::string tag;
  cudaEvent_t ev;
};
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t g_prof_begin = nullptr;
struct ProfAgg {
  std::string tag;
  int64_t launches;
  float ms;
};
static std::vector<ProfAgg> g_prof_result;

void prof_mark(const char* tag) {
  cudaEvent_t ev;
  if (!g_prof_pool.empty()) {
    ev = g_prof_pool.back();
    g_prof_pool.pop_back();
  } else if (cudaEventCreate(&ev) != cudaSuccess) {
    return;
  }
  cudaEventRecord(ev, ctx().stream);
  g_prof.push_back({tag, ev});
}

int ensure_init() {
  if (ctx().ready) return ATHENA_OK;
  return athena_cuda_init(-1);
}

}  // namespace athena

using namespace athena;

ATHENA_API int athena_cuda_init(int32_t device) {
  Context& c = ctx();
  if (c.ready) {
    if (device >= 0 && device != c.device) {
      set_error("athena_cuda_init: already initialised on device %d", c.device);
      return ATHENA_ERR_STATE;
    }
    return ATHENA_OK;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("athena_cuda_init: no usable CUDA device (%s); libathena_cuda has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return ATHENA_ERR_CUDA;
  }
  if (device < 0) {
    const char* s = getenv("ATHENA_CUDA_DEVICE");
    if (!s) s = getenv("LOCAL_RANK");
    device = s ? atoi(s) : 0;
  }
  ATH_REQUIRE(device < count, ATHENA_ERR_ARG, "athena_cuda_init: device %d of %d", device, count);
  ATH_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ATH_CUDA(cudaGetDeviceProperties(&prop, device));
  ATH_REQUIRE(prop.major == 10, ATHENA_ERR_CUDA,
              "athena_cuda_init: device %d is sm_%d%d; this library is built for sm_100a only",
              device, prop.major, prop.minor);
  c.device = device;
  c.sm_count = prop.multiProcessorCount;
  c.total_mem = prop.totalGlobalMem;
  c.max_smem_optin = prop.sharedMemPerBlockOptin;
  ATH_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  ATH_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
  ATH_CUDA(cudaEventCreateWithFlags(&c.mark, cudaEventDisableTiming));
  for (int i = 0; i < 32; ++i)
    ATH_CUDA(cudaEventCreateWithFlags(&c.ev_pool[i], cudaEventDisableTiming));
  c.mark_valid = false;
  for (int i = 0; i < 16; ++i) {
    ATH_CUDA(cudaEventCreate(&c.ev_start[i]));
    ATH_CUDA(cudaEventCreate(&c.ev_stop[i]));
  }
  c.ready = true;
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_shutdown(void) {
  Context& c = ctx();
  if (!c.ready) return ATHENA_OK;
  cudaStreamSynchronize(c.copy_stream);
  cudaStreamSynchronize(c.stream);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_objects.clear();
  }
  pool_trim();
  athena_cuda_comm_destroy();
  if (c.flush_buf) cudaFree(c.flush_buf);
  c.flush_buf = nullptr;
  for (int i = 0; i < 16; ++i) {
    cudaEventDestroy(c.ev_start[i]);
    cudaEventDestroy(c.ev_stop[i]);
  }
  cudaStreamSynchronize(c.copy_stream);
  cudaStreamDestroy(c.copy_stream);
  c.copy_stream = nullptr;
  cudaEventDestroy(c.mark);
  for (int i = 0; i < 32; ++i) cudaEventDestroy(c.ev_pool[i]);
  cudaStreamDestroy(c.stream);
  c.stream = nullptr;
  c.ready = false;
  return ATHENA_OK;
}

ATHENA_API const char* athena_cuda_last_error(void) { return g_err; }

ATHENA_API int athena_cuda_version(int32_t* major, int32_t* minor) {
  if (major) *major = 0;
  if (minor) *minor = 1;
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_device_info(int32_t* device, int32_t* sm_count,
                                       int64_t* total_mem_bytes) {
  ATH_TRY(ensure_init());
  if (device) *device = ctx().device;
  if (sm_count) *sm_count = ctx().sm_count;
  if (total_mem_bytes) *total_mem_bytes = (int64_t)ctx().total_mem;
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_synchronize(void) {
  ATH_TRY(ensure_init());
  ATH_CUDA(cudaStreamSynchronize(ctx().copy_stream));
  ATH_CUDA(cudaStreamSynchronize(ctx().stream));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_malloc(void** dptr, size_t bytes) {
  ATH_TRY(ensure_init());
  ATH_REQUIRE(dptr, ATHENA_ERR_ARG, "athena_cuda_malloc: null out pointer");
  ATH_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_free(void* dptr) {
  ATH_TRY(ensure_init());
  ATH_CUDA(cudaStreamSynchronize(ctx().stream));
  ATH_CUDA(cudaFree(dptr));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_host_alloc(void** hptr, size_t bytes) {
  ATH_TRY(ensure_init());
  ATH_REQUIRE(hptr, ATHENA_ERR_ARG, "athena_cuda_host_alloc: null out pointer");
  ATH_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_host_free(void* hptr) {
  ATH_TRY(ensure_init());
  ATH_CUDA(cudaFreeHost(hptr));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_memcpy_h2d(void* dst, const void* src, size_t bytes) {
  ATH_TRY(ensure_init());
  ATH_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx().stream));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_memcpy_d2h(void* dst, const void* src, size_t bytes) {
  ATH_TRY(ensure_init());
  ATH_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx().stream));
  ATH_CUDA(cudaStreamSynchronize(ctx().stream));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_memset(void* dst, int value, size_t bytes) {
  ATH_TRY(ensure_init());
  ATH_CUDA(cudaMemsetAsync(dst, value, bytes, ctx().stream));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_timer_start(int32_t slot) {
  ATH_TRY(ensure_init());
  ATH_REQUIRE(slot >= 0 && slot < 16, ATHENA_ERR_ARG, "timer slot %d out of range", slot);
  ATH_CUDA(cudaEventRecord(ctx().ev_start[slot], ctx().stream));
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_timer_stop(int32_t slot, float* elapsed_ms) {
  ATH_TRY(ensure_init());
  ATH_REQUIRE(slot >= 0 && slot < 16, ATHENA_ERR_ARG, "timer slot %d out of range", slot);
  ATH_CUDA(cudaEventRecord(ctx().ev_stop[slot], ctx().stream));
  ATH_CUDA(cudaEventSynchronize(ctx().ev_stop[slot]));
  float ms = 0.f;
  ATH_CUDA(cudaEventElapsedTime(&ms, ctx().ev_start[slot], ctx().ev_stop[slot]));
  if (elapsed_ms) *elapsed_ms = ms;
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_launch_count(int64_t* n) {
  if (n) *n = ctx().launches.load();
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_profile_begin(void) {
  ATH_TRY(ensure_init());
  for (auto& r : g_prof) g_prof_pool.push_back(r.ev);
  g_prof.clear();
  g_prof_result.clear();
  if (!g_prof_begin) ATH_CUDA(cudaEventCreate(&g_prof_begin));
  ATH_CUDA(cudaEventRecord(g_prof_begin, ctx().stream));
  ctx().profiling = true;
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_profile_end(int32_t* num_tags) {
  ATH_TRY(ensure_init());
  ctx().profiling = false;
  ATH_CUDA(cudaStreamSynchronize(ctx().stream));
  g_prof_result.clear();
  cudaEvent_t prev = g_prof_begin;
  for (auto& r : g_prof) {
    float ms = 0.f;
    ATH_CUDA(cudaEventElapsedTime(&ms, prev, r.ev));
    prev = r.ev;
    bool found = false;
    for (auto& a : g_prof_result)
      if (a.tag == r.tag) {
        a.launches += 1;
        a.ms += ms;
        found = true;
        break;
      }
    if (!found) g_prof_result.push_back({r.tag, 1, ms});
  }
  if (num_tags) *num_tags = (int32_t)g_prof_result.size();
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_profile_get(int32_t index, char* name, int32_t name_capacity,
                                       int64_t* launches, float* total_ms) {
  ATH_REQUIRE(index >= 0 && index < (int32_t)g_prof_result.size(), ATHENA_ERR_ARG,
              "profile_get: index %d out of range", index);
  const ProfAgg& a = g_prof_result[index];
  if (name && name_capacity > 0) {
    strncpy(name, a.tag.c_str(), (size_t)name_capacity - 1);
    name[name_capacity - 1] = 0;
  }
  if (launches) *launches = a.launches;
  if (total_ms) *total_ms = a.ms;
  return ATHENA_OK;
}

__global__ void k_flush(float4* buf, size_t n4, float v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) buf[i] = make_float4(v, v, v, v);
}

ATHENA_API int athena_cuda_flush_l2(void) {
  ATH_TRY(ensure_init());
  Context& c = ctx();
  if (!c.flush_buf) {
    c.flush_bytes = size_t(256) << 20;  // 2x the 126 MB L2
    ATH_CUDA(cudaMalloc(&c.flush_buf, c.flush_bytes));
  }
  k_flush<<<c.sm_count * 8, 256, 0, c.stream>>>((float4*)c.flush_buf, c.flush_bytes / 16, 0.f);
  ATH_LAUNCHED();
  return ATHENA_OK;
}
