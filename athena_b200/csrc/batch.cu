// K1: graph_type batch -> device CSR / CSC / degrees / normalisation
// coefficients / degree buckets.  Replaces the per-sample host deep copies of
// msgpass_layer_type%set_graph (athena_msgpass_layer_sub.f90:144-174) with one
// device build per mini-batch that all layers and both passes share.
//
// Integer results are bit-exact against oracle_batch_build / oracle_bucketize
// (oracle/athena_oracle.c): counting uses integer atomics (order-independent
// totals); the CSC fill claims slots in arbitrary order and every column is
// then sorted by CSR entry index, which yields the unique stable order.
#include <algorithm>
#include <climits>

#include "athena_internal.h"

namespace athena {

// largest s in [0, n) with off[s] <= x   (off is non-decreasing, off[0] == 0)
__device__ __forceinline__ int find_segment(const int32_t* __restrict__ off, int n, int x) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (__ldg(off + mid) <= x) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void k_convert_rows(int B, int V, int Z, const int32_t* __restrict__ nz,
                               const int32_t* __restrict__ voff,
                               const int32_t* __restrict__ zoff,
                               const int32_t* __restrict__ ia_cat, int32_t* __restrict__ row_ptr,
                               int32_t* __restrict__ deg, int32_t* __restrict__ vgraph,
                               int32_t* __restrict__ status) {
  int gv = blockIdx.x * blockDim.x + threadIdx.x;
  if (gv == 0) row_ptr[V] = Z;
  if (gv >= V) return;
  int s = find_segment(voff, B, gv);
  int i = gv - voff[s];
  const int32_t* ia = ia_cat + voff[s] + s;  // each graph contributes nv+1 row pointers
  int a0 = ia[i], a1 = ia[i + 1];
  int nvs = voff[s + 1] - voff[s];
  bool bad = (a1 < a0) || (a0 < 1) || (a1 - 1 > nz[s]) || (i == 0 && a0 != 1) ||
             (i == nvs - 1 && a1 - 1 != nz[s]);
  if (bad) {
    atomicMin(status, s);
    a0 = 1;
    a1 = 1;
  }
  row_ptr[gv] = zoff[s] + a0 - 1;
  deg[gv] = a1 - a0;
  vgraph[gv] = s;
}

__global__ void k_convert_entries(int B, int Z, const int32_t* __restrict__ nv,
                                  const int32_t* __restrict__ ne,
                                  const int32_t* __restrict__ voff,
                                  const int32_t* __restrict__ zoff,
                                  const int32_t* __restrict__ eoff,
                                  const int2* __restrict__ ja_cat, int32_t* __restrict__ col,
                                  int32_t* __restrict__ eid, int32_t* __restrict__ csc_cnt,
                                  int32_t* __restrict__ status) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= Z) return;
  int s = find_segment(zoff, B, w);
  int2 p = ja_cat[w];
  int nb = p.x;
  if (nb < 1 || nb > nv[s]) {
    atomicMin(status, s);
    nb = 1;  // keep later kernels in bounds; the batch is flagged invalid
  }
  int c = voff[s] + nb - 1;
  col[w] = c;
  eid[w] = (p.y >= 1 && p.y <= ne[s]) ? eoff[s] + p.y - 1 : -1;
  atomicAdd(csc_cnt + c, 1);
}

// ---- exclusive scan of int32 (three small kernels) -------------------------
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int x, int* total) {
  __shared__ int warp_sums[32];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int nw = blockDim.x >> 5;
    int ws = lane < nw ? warp_sums[lane] : 0;
    int wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += y;
    }
    warp_sums[lane] = wi - ws;  // exclusive warp offsets
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  int res = incl - x + warp_sums[warp];
  __syncthreads();
  return res;
}

// in may alias out (no __restrict__ on purpose)
__global__ void k_scan_tiles(const int32_t* in, int32_t* out, int n,
                             int32_t* __restrict__ tile_sums) {
  __shared__ int total;
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  int ex = block_exclusive_scan(s, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the tile sums in place; writes the grand total
__global__ void k_scan_sums(int32_t* __restrict__ tile_sums, int ntiles,
                            int32_t* __restrict__ grand_total) {
  __shared__ int total;
  int carry = 0;
  for (int base = 0; base < ntiles; base += blockDim.x) {
    int i = base + threadIdx.x;
    int x = i < ntiles ? tile_sums[i] : 0;
    int ex = block_exclusive_scan(x, &total);
    if (i < ntiles) tile_sums[i] = ex + carry;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void k_scan_add(int32_t* __restrict__ out, int n,
                           const int32_t* __restrict__ tile_sums) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += tile_sums[i / SCAN_TILE];
}

// out[0..n) = exclusive scan of in[0..n);  out[n] = total.  in may alias out.
static int exclusive_scan(const int32_t* in, int32_t* out, int n, int32_t* tile_sums) {
  cudaStream_t st = ctx().stream;
  int ntiles = (int)cdiv(n, SCAN_TILE);
  if (ntiles == 0) {
    ATH_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t), st));
    return ATHENA_OK;
  }
  k_scan_tiles<<<ntiles, SCAN_THREADS, 0, st>>>(in, out, n, tile_sums);
  ATH_LAUNCHED();
  k_scan_sums<<<1, SCAN_THREADS, 0, st>>>(tile_sums, ntiles, out + n);
  ATH_LAUNCHED();
  k_scan_add<<<(int)cdiv(n, 256), 256, 0, st>>>(out, n, tile_sums);
  ATH_LAUNCHED();
  return ATHENA_OK;
}

// ---- coefficients + CSC fill ------------------------------------------------
// 8 lanes per CSR row.  coef follows athena_diffstruc_extd_sub_kipf.f90:39-42:
// integer degree product, converted to real32, raised to -1/2.
__global__ void k_coef_and_csc_fill(int V, const int32_t* __restrict__ row_ptr,
                                    const int32_t* __restrict__ col,
                                    const int32_t* __restrict__ deg, float* __restrict__ coef,
                                    int32_t* __restrict__ cursor, int32_t* __restrict__ csc_ent,
                                    int32_t* __restrict__ csc_src) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int lane = (int)(t & 7);
  if ((t >> 3) >= V) return;
  int v = (int)(t >> 3);
  int beg = row_ptr[v], end = row_ptr[v + 1];
  int dv = deg[v];
  for (int w = beg + lane; w < end; w += 8) {
    int u = col[w];
    int prod = dv * __ldg(deg + u);
    coef[w] = 1.0f / sqrtf((float)prod);
    int pos = atomicAdd(cursor + u, 1);
    csc_ent[pos] = w;
    csc_src[pos] = v;
  }
}

// warp per CSC column: rank sort by entry index for columns of <= 32 entries;
// longer columns are queued for k_csc_sort_long.
__global__ void k_csc_sort_short(int V, const int32_t* __restrict__ csc_ptr,
                                 int32_t* __restrict__ csc_ent, int32_t* __restrict__ csc_src,
                                 int32_t* __restrict__ long_list, int32_t* __restrict__ long_count) {
  long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (gw >= V) return;
  int warp = (int)gw;
  int beg = csc_ptr[warp];
  int len = csc_ptr[warp + 1] - beg;
  if (len <= 1) return;
  if (len > 32) {
    if (lane == 0) long_list[atomicAdd(long_count, 1)] = warp;
    return;
  }
  int key = lane < len ? csc_ent[beg + lane] : INT_MAX;
  int val = lane < len ? csc_src[beg + lane] : 0;
  int rank = 0;
  for (int j = 0; j < len; ++j) rank += (__shfl_sync(0xffffffffu, key, j) < key) ? 1 : 0;
  __syncwarp();
  if (lane < len) {
    csc_ent[beg + rank] = key;
    csc_src[beg + rank] = val;
  }
}

// Ascending-only bitonic network (mirror step + half cleaners) so that virtual
// +inf padding above n never has to move.
template <class Swap>
__device__ __forceinline__ void bitonic_network(int n, Swap swap_if) {
  int P = 1;
  while (P < n) P <<= 1;
  for (int k = 2; k <= P; k <<= 1) {
    int half = k >> 1;
    for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
      int grp = t / half, off = t - grp * half;
      int i = grp * k + off, l = grp * k + k - 1 - off;
      if (l < n) swap_if(i, l);
    }
    __syncthreads();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        int i = 2 * j * (t / j) + (t % j), l = i + j;
        if (l < n) swap_if(i, l);
      }
      __syncthreads();
    }
  }
}

constexpr int LONG_SMEM_KEYS = 16384;  // 128 KB of 64-bit keys

__global__ void __launch_bounds__(1024)
k_csc_sort_long(const int32_t* __restrict__ csc_ptr, int32_t* __restrict__ csc_ent,
                int32_t* __restrict__ csc_src, const int32_t* __restrict__ long_list,
                const int32_t* __restrict__ long_count) {
  extern __shared__ unsigned long long keys[];
  int count = *long_count;
  for (int q = blockIdx.x; q < count; q += gridDim.x) {
    int u = long_list[q];
    int beg = csc_ptr[u];
    int n = csc_ptr[u + 1] - beg;
    int32_t* ent = csc_ent + beg;
    int32_t* src = csc_src + beg;
    if (n <= LONG_SMEM_KEYS) {
      for (int i = threadIdx.x; i < n; i += blockDim.x)
        keys[i] = ((unsigned long long)(unsigned)ent[i] << 32) | (unsigned)src[i];
      __syncthreads();
      bitonic_network(
          n, [&](int i, int l) {
            unsigned long long a = keys[i], b = keys[l];
            if (b < a) {
              keys[i] = b;
              keys[l] = a;
            }
          });
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        ent[i] = (int)(keys[i] >> 32);
        src[i] = (int)(keys[i] & 0xffffffffu);
      }
      __syncthreads();
    } else {
      __syncthreads();
      bitonic_network(
          n, [&](int i, int l) {
            int a = ent[i], b = ent[l];
            if (b < a) {
              ent[i] = b;
              ent[l] = a;
              int sa = src[i];
              src[i] = src[l];
              src[l] = sa;
            }
          });
    }
  }
}

// ---- compact per-tile operands of the fused kernels ----------------------------
// one CTA per tile: tile-local neighbour indices (CSR and CSC entry ranges of a tile of
// whole graphs coincide), and deg^-1/2 of the tile's rows
__global__ void k_tile_operands(const int4* __restrict__ tiles, const int32_t* __restrict__ col,
                                const int32_t* __restrict__ csc_src,
                                const int32_t* __restrict__ deg, uint8_t* __restrict__ col8,
                                uint8_t* __restrict__ csc8, float* __restrict__ rsdeg,
                                const int32_t* __restrict__ vgraph,
                                const int32_t* __restrict__ nv, int32_t* __restrict__ vcount,
                                const int32_t* __restrict__ row_ptr,
                                const int32_t* __restrict__ csc_ptr, uint4* __restrict__ abits,
                                uint4* __restrict__ atbits, int32_t* __restrict__ multi) {
  const int4 ti = tiles[blockIdx.x];
  // adjacency rows of the tile as 128-bit masks (bit c = an entry to tile-local vertex c),
  // CSR and CSC direction: the operand of the tensor-core gather (pipe_tcg.cu).  A repeated
  // (row, column) pair cannot be a bit; *multi reports it and the batch keeps the list path.
  for (int i = threadIdx.x; i < 2 * ti.y; i += blockDim.x) {
    const int r = i >> 1, dir = i & 1;
    const int32_t* ptr = dir ? csc_ptr : row_ptr;
    const int32_t* idx = dir ? csc_src : col;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    bool dup = false;
    for (int e = ptr[ti.x + r]; e < ptr[ti.x + r + 1]; ++e) {
      const int c = idx[e] - ti.x;
      const uint32_t bit = 1u << (c & 31);
      dup |= (w[(c >> 5) & 3] & bit) != 0u;
      w[(c >> 5) & 3] |= bit;
    }
    (dir ? atbits : abits)[ti.x + r] = make_uint4(w[0], w[1], w[2], w[3]);
    if (dup) atomicOr(multi, 1);
    // a vertex without CSR entries has deg^-1/2 = Infinity (as in the reference, where it only
    // matters to the rows that list the vertex): in the dense adjacency product 0 * Infinity =
    // NaN would reach every row of the tile -- such batches keep the list kernels
    if (dir == 1 && deg[ti.x + r] == 0) atomicOr(multi, 2);
  }
  for (int e = threadIdx.x; e < ti.w; e += blockDim.x) {
    col8[ti.z + e] = static_cast<uint8_t>(col[ti.z + e] - ti.x);
    csc8[ti.z + e] = static_cast<uint8_t>(csc_src[ti.z + e] - ti.x);
  }
  for (int r = threadIdx.x; r < ti.y; r += blockDim.x) {
    rsdeg[ti.x + r] = 1.0f / sqrtf(static_cast<float>(deg[ti.x + r]));
    vcount[ti.x + r] = nv[vgraph[ti.x + r]];
  }
}

// rows of a CSR/CSC pointer array with more than `limit` entries (order irrelevant)
__global__ void k_find_long(int V, const int32_t* __restrict__ ptr, int limit,
                            int32_t* __restrict__ list, int32_t* __restrict__ count) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < V && ptr[v + 1] - ptr[v] > limit) list[atomicAdd(count, 1)] = v;
}

// ---- degree buckets -----------------------------------------------------------
constexpr int BKT_THREADS = 1024;
constexpr int BKT_MAX_D = 256;

__device__ __forceinline__ int bucket_of(int deg, int min_deg, int max_deg) {
  return max(min_deg, min(deg, max_deg)) - min_deg;  // 0-based
}

// pass 0: bucket ids + per-block histograms.  pass 1: stable scatter.
template <int PASS>
__global__ void __launch_bounds__(BKT_THREADS)
k_bucket_pass(int V, int D, int min_deg, int max_deg, const int32_t* __restrict__ deg,
              int32_t* __restrict__ bkt, int32_t* __restrict__ block_hist,
              int32_t* __restrict__ perm) {
  extern __shared__ int warp_cnt[];  // [32][D]
  int v = blockIdx.x * BKT_THREADS + threadIdx.x;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * D; i += BKT_THREADS) warp_cnt[i] = 0;
  __syncthreads();
  bool live = v < V;
  int b = live ? bucket_of(deg[v], min_deg, max_deg) : -1;
  unsigned peers = __match_any_sync(0xffffffffu, b);
  int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
  if (live && rank_in_warp == 0) warp_cnt[warp * D + b] = __popc(peers);
  __syncthreads();
  if (PASS == 0) {
    if (live) bkt[v] = b;
    for (int d = threadIdx.x; d < D; d += BKT_THREADS) {
      int s = 0;
      for (int w = 0; w < 32; ++w) s += warp_cnt[w * D + d];
      block_hist[(size_t)blockIdx.x * D + d] = s;
    }
  } else if (live) {
    int off = block_hist[(size_t)blockIdx.x * D + b];  // global base of (block, bucket)
    for (int w = 0; w < warp; ++w) off += warp_cnt[w * D + b];
    perm[off + rank_in_warp] = v;
  }
}

// single block: block_hist[blk][d] -> global base offsets, bkt_ptr[D+1]
__global__ void k_bucket_scan(int nblocks, int D, int32_t* __restrict__ block_hist,
                              int32_t* __restrict__ bkt_ptr) {
  __shared__ int totals[BKT_MAX_D + 1];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    int run = 0;
    for (int b = 0; b < nblocks; ++b) {
      int c = block_hist[(size_t)b * D + d];
      block_hist[(size_t)b * D + d] = run;
      run += c;
    }
    totals[d] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int d = 0; d < D; ++d) {
      int c = totals[d];
      totals[d] = run;
      bkt_ptr[d] = run;
      run += c;
    }
    bkt_ptr[D] = run;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    int base = totals[d];
    for (int b = 0; b < nblocks; ++b) block_hist[(size_t)b * D + d] += base;
  }
}

BucketSet* Batch::find_buckets(int min_deg, int max_deg) const {
  for (auto& bs : buckets)
    if (bs->min_deg == min_deg && bs->max_deg == max_deg) return bs.get();
  return nullptr;
}

int batch_bucketize(Batch* b, int min_deg, int max_deg, BucketSet** out) {
  ATH_REQUIRE(max_deg >= min_deg && min_deg >= 1, ATHENA_ERR_ARG,
              "bucketize: need 1 <= min_vertex_degree <= max_vertex_degree (got %d, %d)", min_deg,
              max_deg);
  int D = max_deg - min_deg + 1;
  ATH_REQUIRE(D <= BKT_MAX_D, ATHENA_ERR_ARG, "bucketize: %d degree buckets > supported %d", D,
              BKT_MAX_D);
  if (BucketSet* hit = b->find_buckets(min_deg, max_deg)) {
    if (out) *out = hit;
    return ATHENA_OK;
  }
  std::unique_ptr<BucketSet> bs(new BucketSet);
  bs->min_deg = min_deg;
  bs->max_deg = max_deg;
  bs->D = D;
  int V = (int)b->V;
  int nblocks = (int)cdiv(V, BKT_THREADS);
  ATH_TRY(bs->bkt.reserve(sizeof(int32_t) * (size_t)(V + 1)));
  ATH_TRY(bs->perm.reserve(sizeof(int32_t) * (size_t)(V + 1)));
  ATH_TRY(bs->bkt_ptr.reserve(sizeof(int32_t) * (size_t)(D + 1)));
  ATH_TRY(bs->scratch.reserve(sizeof(int32_t) * (size_t)(nblocks + 1) * D));
  cudaStream_t st = ctx().stream;
  size_t smem = sizeof(int) * 32 * (size_t)D;
  if (V > 0) {
    k_bucket_pass<0><<<nblocks, BKT_THREADS, smem, st>>>(V, D, min_deg, max_deg, b->deg,
                                                         bs->bkt.as<int32_t>(),
                                                         bs->scratch.as<int32_t>(), nullptr);
    ATH_LAUNCHED();
  }
  k_bucket_scan<<<1, 256, 0, st>>>(nblocks, D, bs->scratch.as<int32_t>(),
                                   bs->bkt_ptr.as<int32_t>());
  ATH_LAUNCHED();
  if (V > 0) {
    k_bucket_pass<1><<<nblocks, BKT_THREADS, smem, st>>>(V, D, min_deg, max_deg, b->deg,
                                                         bs->bkt.as<int32_t>(),
                                                         bs->scratch.as<int32_t>(),
                                                         bs->perm.as<int32_t>());
    ATH_LAUNCHED();
  }
  if (out) *out = bs.get();
  b->buckets.push_back(std::move(bs));
  return ATHENA_OK;
}

}  // namespace athena

using namespace athena;

ATHENA_API int athena_cuda_batch_create(athena_handle_t* batch, int32_t num_graphs,
                                        const int32_t* num_vertices, const int32_t* num_edges,
                                        const int32_t* num_entries, const int32_t* adj_ia,
                                        const int32_t* adj_ja, int32_t mem, int32_t validate) {
  ATH_TRY(ensure_init());
  ATH_REQUIRE(batch && num_vertices && num_edges && num_entries && adj_ia, ATHENA_ERR_ARG,
              "batch_create: null argument");
  ATH_REQUIRE(num_graphs >= 1, ATHENA_ERR_ARG, "batch_create: num_graphs = %d", num_graphs);
  ATH_REQUIRE(mem == ATHENA_MEM_HOST || mem == ATHENA_MEM_DEVICE, ATHENA_ERR_ARG,
              "batch_create: bad mem %d", mem);
  const int B = num_graphs;
  std::vector<int32_t> meta((size_t)5 * B + 3);
  int32_t* h_nv = meta.data();
  int32_t* h_ne = h_nv + B;
  int32_t* h_nz = h_ne + B;
  int32_t* h_voff = h_nz + B;
  int32_t* h_zoff = h_voff + B + 1;
  // eoff appended below (needs its own B+1)
  std::vector<int32_t> eoffv((size_t)B + 1);
  int64_t V = 0, Z = 0, E = 0;
  for (int s = 0; s < B; ++s) {
    ATH_REQUIRE(num_vertices[s] >= 0 && num_edges[s] >= 0 && num_entries[s] >= 0, ATHENA_ERR_ARG,
                "batch_create: negative size in graph %d", s);
    h_nv[s] = num_vertices[s];
    h_ne[s] = num_edges[s];
    h_nz[s] = num_entries[s];
    h_voff[s] = (int32_t)V;
    h_zoff[s] = (int32_t)Z;
    eoffv[s] = (int32_t)E;
    V += num_vertices[s];
    Z += num_entries[s];
    E += num_edges[s];
  }
  ATH_REQUIRE(V < INT32_MAX - 2 && Z < INT32_MAX - 2 && E < INT32_MAX - 2, ATHENA_ERR_ARG,
              "batch_create: batch exceeds int32 indexing (V=%lld Z=%lld E=%lld)", (long long)V,
              (long long)Z, (long long)E);
  ATH_REQUIRE(Z == 0 || adj_ja, ATHENA_ERR_ARG, "batch_create: null adj_ja");
  h_voff[B] = (int32_t)V;
  h_zoff[B] = (int32_t)Z;
  eoffv[B] = (int32_t)E;

  std::unique_ptr<Batch> b(new Batch);
  b->h_voff.assign(h_voff, h_voff + B + 1);
  b->B = B;
  b->V = V;
  b->Z = Z;
  b->E = E;
  cudaStream_t st = ctx().stream;

  // graph-aligned tiles (greedy packing of whole graphs), see Batch::tiles
  {
    std::vector<int32_t> tiles;
    tiles.reserve((size_t)(V / TILE_ROWS + B / 8 + 4) * 4);
    bool ok = true;
    int32_t r0 = 0, e0 = 0, rows = 0, ents = 0;
    for (int s = 0; s < B && ok; ++s) {
      const int32_t nvs = num_vertices[s], nzs = num_entries[s];
      if (nvs > TILE_ROWS || nzs > TILE_ENTRIES) ok = false;
      if (rows + nvs > TILE_ROWS || ents + nzs > TILE_ENTRIES) {
        if (rows > 0) tiles.insert(tiles.end(), {r0, rows, e0, ents});
        r0 += rows;
        e0 += ents;
        rows = 0;
        ents = 0;
      }
      rows += nvs;
      ents += nzs;
    }
    if (ok && rows > 0) tiles.insert(tiles.end(), {r0, rows, e0, ents});
    if (ok && !tiles.empty()) {
      b->num_tiles = (int32_t)(tiles.size() / 4);
      ATH_TRY(b->tiles.reserve(sizeof(int32_t) * tiles.size()));
      ATH_CUDA(cudaMemcpyAsync(b->tiles.p, tiles.data(), sizeof(int32_t) * tiles.size(),
                               cudaMemcpyHostToDevice, st));
    }
  }

  // meta on the device: nv | ne | nz | voff | zoff | eoff
  size_t meta_ints = (size_t)3 * B + 3 * ((size_t)B + 1);
  ATH_TRY(b->meta.reserve(sizeof(int32_t) * meta_ints));
  int32_t* d_meta = b->meta.as<int32_t>();
  ATH_CUDA(cudaMemcpyAsync(d_meta, meta.data(), sizeof(int32_t) * ((size_t)5 * B + 2),
                           cudaMemcpyHostToDevice, st));
  ATH_CUDA(cudaMemcpyAsync(d_meta + 5 * (size_t)B + 2, eoffv.data(),
                           sizeof(int32_t) * ((size_t)B + 1), cudaMemcpyHostToDevice, st));
  b->nv = d_meta;
  b->ne = d_meta + B;
  const int32_t* d_nz = d_meta + 2 * (size_t)B;
  b->voff = d_meta + 3 * (size_t)B;
  b->zoff = b->voff + B + 1;
  b->eoff = b->zoff + B + 1;

  // raw adjacency on the device
  const int32_t* d_ia = adj_ia;
  const int32_t* d_ja = adj_ja;
  if (mem == ATHENA_MEM_HOST) {
    size_t ia_ints = (size_t)V + B, ja_ints = 2 * (size_t)Z;
    size_t ia_pad = (size_t)round_up((int64_t)ia_ints, 4);
    ATH_TRY(b->raw.reserve(sizeof(int32_t) * (ia_pad + ja_ints + 4)));
    int32_t* r = b->raw.as<int32_t>();
    // the adjacency (the bulk of the bytes) travels on the copy stream; the build kernels
    // below wait for it, whatever the caller queues next on the copy stream does not
    cudaEvent_t ev_adj = nullptr;
    ATH_TRY(side_begin());
    ATH_TRY(side_copy(r, adj_ia, sizeof(int32_t) * ia_ints));
    if (Z > 0) ATH_TRY(side_copy(r + ia_pad, adj_ja, sizeof(int32_t) * ja_ints));
    ATH_TRY(side_fence(&ev_adj));
    ATH_TRY(main_wait(ev_adj));
    d_ia = r;
    d_ja = r + ia_pad;
  } else {
    ATH_REQUIRE((reinterpret_cast<uintptr_t>(adj_ja) & 7) == 0, ATHENA_ERR_ARG,
                "batch_create: device adj_ja must be 8-byte aligned");
  }

  // integer structures, one allocation
  size_t Vp = (size_t)round_up(V + 1, 4), Zp = (size_t)round_up(Z + 1, 4);
  size_t total = 5 * Vp /* row_ptr deg vgraph csc_ptr cursor */ + 4 * Zp /* col eid src ent */ +
                 Vp /* long list */;
  ATH_TRY(b->ints.reserve(sizeof(int32_t) * total));
  int32_t* p = b->ints.as<int32_t>();
  b->row_ptr = p; p += Vp;
  b->deg = p; p += Vp;
  b->vgraph = p; p += Vp;
  b->csc_ptr = p; p += Vp;
  int32_t* cursor = p; p += Vp;
  int32_t* long_list = p; p += Vp;
  b->col = p; p += Zp;
  b->eid = p; p += Zp;
  b->csc_src = p; p += Zp;
  b->csc_ent = p; p += Zp;
  ATH_TRY(b->coef_buf.reserve(sizeof(float) * Zp));
  b->coef = b->coef_buf.as<float>();
  int ntiles = (int)cdiv(V + 1, SCAN_TILE) + 1;
  ATH_TRY(b->scratch.reserve(sizeof(int32_t) * (size_t)ntiles));
  ATH_TRY(b->status.reserve(sizeof(int32_t) * 4));
  int32_t* status = b->status.as<int32_t>();
  const int32_t status_init[4] = {INT_MAX, 0, 0, 0};
  ATH_CUDA(cudaMemcpyAsync(status, status_init, sizeof(status_init), cudaMemcpyHostToDevice, st));

  ATH_CUDA(cudaMemsetAsync(b->csc_ptr, 0, sizeof(int32_t) * Vp, st));
  k_convert_rows<<<(int)cdiv(V + 1, 256), 256, 0, st>>>(B, (int)V, (int)Z, d_nz, b->voff, b->zoff,
                                                       d_ia, b->row_ptr, b->deg, b->vgraph,
                                                       status);
  ATH_LAUNCHED();
  if (Z > 0) {
    k_convert_entries<<<(int)cdiv(Z, 256), 256, 0, st>>>(
        B, (int)Z, b->nv, b->ne, b->voff, b->zoff, b->eoff, reinterpret_cast<const int2*>(d_ja),
        b->col, b->eid, b->csc_ptr, status);
    ATH_LAUNCHED();
  }
  // csc_ptr currently holds per-column counts in [0, V)
  ATH_TRY(exclusive_scan(b->csc_ptr, b->csc_ptr, (int)V, b->scratch.as<int32_t>()));
  if (V > 0 && Z > 0) {
    ATH_CUDA(cudaMemcpyAsync(cursor, b->csc_ptr, sizeof(int32_t) * (size_t)V,
                             cudaMemcpyDeviceToDevice, st));
    k_coef_and_csc_fill<<<(int)cdiv(V * 8, 256), 256, 0, st>>>((int)V, b->row_ptr, b->col, b->deg,
                                                              b->coef, cursor, b->csc_ent,
                                                              b->csc_src);
    ATH_LAUNCHED();
    k_csc_sort_short<<<(int)cdiv(V * 32, 256), 256, 0, st>>>((int)V, b->csc_ptr, b->csc_ent,
                                                            b->csc_src, long_list, status + 1);
    ATH_LAUNCHED();
    static bool attr_set = false;
    size_t smem = sizeof(unsigned long long) * LONG_SMEM_KEYS;
    if (!attr_set) {
      ATH_CUDA(cudaFuncSetAttribute(k_csc_sort_long, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      attr_set = true;
    }
    k_csc_sort_long<<<ctx().sm_count, 1024, smem, st>>>(b->csc_ptr, b->csc_ent, b->csc_src,
                                                        long_list, status + 1);
    ATH_LAUNCHED();
  }
  // a row / column can only be long if its graph has that many entries (host-known bound)
  int32_t max_graph_entries = 0;
  for (int s = 0; s < B; ++s) max_graph_entries = std::max(max_graph_entries, num_entries[s]);
  if (V > 0 && Z > 0 && max_graph_entries > LONG_ROW) {
    const size_t cap = (size_t)(Z / LONG_ROW + 4);
    ATH_TRY(b->long_buf.reserve(sizeof(int32_t) * (4 + 2 * cap)));
    int32_t* lb = b->long_buf.as<int32_t>();
    ATH_CUDA(cudaMemsetAsync(lb, 0, sizeof(int32_t) * 4, st));
    b->long_counts = lb;
    b->long_rows = lb + 4;
    b->long_cols = lb + 4 + cap;
    k_find_long<<<(int)cdiv(V, 256), 256, 0, st>>>((int)V, b->row_ptr, LONG_ROW, lb + 4, lb);
    ATH_LAUNCHED();
    k_find_long<<<(int)cdiv(V, 256), 256, 0, st>>>((int)V, b->csc_ptr, LONG_ROW, lb + 4 + cap,
                                                  lb + 1);
    ATH_LAUNCHED();
  }
  if (b->num_tiles > 0) {
    const size_t z16 = (size_t)round_up(Z + 32, 16);
    const size_t v4 = (size_t)round_up(V + 8, 4);
    ATH_TRY(b->tile_ops.reserve(2 * z16 + 2 * sizeof(float) * v4 + 2 * sizeof(uint4) * v4));
    b->col8 = b->tile_ops.as<uint8_t>();
    b->csc8 = b->col8 + z16;
    b->rsdeg = reinterpret_cast<float*>(b->csc8 + z16);
    b->vcount = reinterpret_cast<int32_t*>(b->rsdeg + v4);
    b->abits = reinterpret_cast<uint4*>(b->vcount + v4);
    b->atbits = b->abits + v4;
    k_tile_operands<<<b->num_tiles, 256, 0, st>>>(b->tiles.as<int4>(), b->col, b->csc_src, b->deg,
                                                 b->col8, b->csc8, b->rsdeg, b->vgraph, b->nv,
                                                 b->vcount, b->row_ptr, b->csc_ptr, b->abits,
                                                 b->atbits, status + 2);
    ATH_LAUNCHED();
  }
  Batch* raw = b.release();
  athena_handle_t h = register_object(raw);
  *batch = h;
  if (validate) {
    int rc = athena_cuda_batch_status(h);
    if (rc != ATHENA_OK) {
      destroy_object(h, Kind::Batch);
      *batch = 0;
      return rc;
    }
  }
  return ATHENA_OK;
}

// ---- graphstruc-side construction on the device -------------------------------------------
// graph%generate_adjacency(index_list) (+ graph%add_self_loops()), the calls every reader of the
// reference makes before set_graph (example_library/src/mod_read_chemical_graphs.f90:275-276,
// example/msgpass_euler/src/mod_read_euler.f90:48, example/msgpass_chemical/src/main.f90:108).
// graphstruc v0.2.1 is an un-vendored dependency (fpm.toml:21) and no athena test pins the
// order of its neighbour lists; the order built here is the one of the host restatement
// (athena_b200/graph.py): undirected edge k = (i, j) is listed in row i and -- if i /= j -- in
// row j with edge id k, rows in ascending edge id; add_self_loops appends (v, 0) to every row
// that lists no v.  Bit-exact against that restatement (tests/test_gpu_parity.py).
namespace athena {

constexpr int LOOP_KEY = INT_MAX - 1;  // sorts behind every edge id

__global__ void k_edges_count(int B, int E, const int32_t* __restrict__ voff,
                              const int32_t* __restrict__ eoff, const int2* __restrict__ il,
                              int32_t* __restrict__ cnt, int32_t* __restrict__ hasself,
                              int32_t* __restrict__ status) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  int s = find_segment(eoff, B, k);
  int2 e = il[k];
  int nvs = voff[s + 1] - voff[s];
  if (e.x < 1 || e.x > nvs || e.y < 1 || e.y > nvs) {
    atomicMin(status, s);
    return;
  }
  int gi = voff[s] + e.x - 1, gj = voff[s] + e.y - 1;
  atomicAdd(cnt + gi, 1);
  if (gi != gj) atomicAdd(cnt + gj, 1);
  else hasself[gi] = 1;
}

__global__ void k_edges_loops(int V, const int32_t* __restrict__ hasself, int32_t* __restrict__ cnt) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < V && !hasself[v]) cnt[v] += 1;
}

__global__ void k_edges_fill(int B, int E, int V, int loops, const int32_t* __restrict__ voff,
                             const int32_t* __restrict__ eoff, const int2* __restrict__ il,
                             const int32_t* __restrict__ hasself, int32_t* __restrict__ cursor,
                             int32_t* __restrict__ key, int32_t* __restrict__ nb) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < E) {
    int s = find_segment(eoff, B, k);
    int2 e = il[k];
    int nvs = voff[s + 1] - voff[s];
    if (!(e.x < 1 || e.x > nvs || e.y < 1 || e.y > nvs)) {
      int gi = voff[s] + e.x - 1, gj = voff[s] + e.y - 1;
      int kl = k - eoff[s] + 1;  // edge id inside its graph, 1-based
      int pos = atomicAdd(cursor + gi, 1);
      key[pos] = kl;
      nb[pos] = gj;
      if (gi != gj) {
        pos = atomicAdd(cursor + gj, 1);
        key[pos] = kl;
        nb[pos] = gi;
      }
    }
  }
  if (loops && k < V && !hasself[k]) {
    int pos = atomicAdd(cursor + k, 1);
    key[pos] = LOOP_KEY;
    nb[pos] = k;
  }
}

// adj_ia / adj_ja of every graph in the layout athena_cuda_batch_create takes; 8 lanes per row
__global__ void k_edges_emit(int B, int V, const int32_t* __restrict__ voff,
                             const int32_t* __restrict__ off, const int32_t* __restrict__ key,
                             const int32_t* __restrict__ nb, int32_t* __restrict__ ia_cat,
                             int2* __restrict__ ja_cat, int32_t* __restrict__ nz) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int lane = (int)(t & 7);
  if ((t >> 3) >= V) return;
  int gv = (int)(t >> 3);
  int s = find_segment(voff, B, gv);
  int v0 = voff[s], z0 = off[v0];
  int beg = off[gv], end = off[gv + 1];
  if (lane == 0) {
    ia_cat[gv + s] = beg - z0 + 1;
    if (gv + 1 == voff[s + 1]) {
      ia_cat[gv + s + 1] = end - z0 + 1;
      nz[s] = end - z0;
    }
  }
  for (int w = beg + lane; w < end; w += 8) {
    int kk = key[w];
    ja_cat[w] = make_int2(nb[w] - v0 + 1, kk == LOOP_KEY ? 0 : kk);
  }
}

__global__ void k_merge_status(int32_t* __restrict__ dst, const int32_t* __restrict__ src) {
  atomicMin(dst, *src);
}

// ONNX graph inputs of a message-passing layer (athena_onnx_msgpass_utils.f90:53-92,
// example/msgpass_chemical/validate_onnx.py:38-57): edge_index [3, ncsr] int64, 0-based --
// row 0 the neighbour (source), row 1 the edge-feature index (-1: none), row 2 the vertex whose
// row the entry belongs to (target), entries in CSR order -- and degree [num_nodes] int64.
__global__ void k_onnx_rows(int B, int V, const int32_t* __restrict__ voff,
                            const int32_t* __restrict__ zoff, const long long* __restrict__ degree,
                            int32_t* __restrict__ cnt) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  int s = find_segment(voff, B, v);
  long long d = degree[v];
  long long cap = (long long)zoff[s + 1] - zoff[s];
  cnt[v] = (int)(d < 0 ? 0 : (d > cap ? cap + 1 : d));
}

__global__ void k_onnx_emit(int B, int V, int Z, const int32_t* __restrict__ voff,
                            const int32_t* __restrict__ zoff, const int32_t* __restrict__ off,
                            const long long* __restrict__ ei, int32_t* __restrict__ ia_cat,
                            int2* __restrict__ ja_cat, int32_t* __restrict__ status) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < V) {
    int s = find_segment(voff, B, t);
    // the degrees of graph s must add up to exactly its ncsr entries
    if (off[voff[s]] != zoff[s] || off[voff[s + 1]] != zoff[s + 1]) atomicMin(status, s);
    ia_cat[t + s] = off[t] - zoff[s] + 1;
    if (t + 1 == voff[s + 1]) ia_cat[t + s + 1] = off[t + 1] - zoff[s] + 1;
  }
  if (t < Z) {
    int s = find_segment(zoff, B, t);
    int n = zoff[s + 1] - zoff[s], wl = t - zoff[s];
    int nvs = voff[s + 1] - voff[s];
    const long long* blk = ei + 3ll * zoff[s];
    long long src = blk[wl], ef = blk[n + wl], tgt = blk[2ll * n + wl];
    bool ok = tgt >= 0 && tgt < nvs;
    if (ok) {
      int gv = voff[s] + (int)tgt;
      ok = off[gv] <= t && t < off[gv + 1];  // the entry lies inside its target's row
    }
    if (!ok) atomicMin(status, s);
    int nbv = (src >= 0 && src < nvs) ? (int)src + 1 : 0;  // 0: flagged by the CSR build
    int eidv = (ef >= 0 && ef < INT_MAX - 1) ? (int)ef + 1 : 0;
    ja_cat[t] = make_int2(nbv, eidv);
  }
}

// hands the scratch that holds adj_ia / adj_ja to the batch (the build is asynchronous)
static void adopt(DevBuf& dst, DevBuf& src) {
  dst.release();
  dst.p = src.p;
  dst.cap = src.cap;
  src.p = nullptr;
  src.cap = 0;
}

}  // namespace athena

static int graph_offsets(int B, const int32_t* num_vertices, const int32_t* counts,
                         std::vector<int32_t>& voff, std::vector<int32_t>& coff, const char* who) {
  int64_t V = 0, C = 0;
  voff.resize((size_t)B + 1);
  coff.resize((size_t)B + 1);
  for (int s = 0; s < B; ++s) {
    ATH_REQUIRE(num_vertices[s] >= 0 && counts[s] >= 0, ATHENA_ERR_ARG,
                "%s: negative size in graph %d", who, s);
    voff[s] = (int32_t)V;
    coff[s] = (int32_t)C;
    V += num_vertices[s];
    C += counts[s];
    ATH_REQUIRE(V < INT32_MAX / 2 && C < INT32_MAX / 4, ATHENA_ERR_ARG,
                "%s: batch exceeds int32 indexing", who);
  }
  voff[B] = (int32_t)V;
  coff[B] = (int32_t)C;
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_batch_create_from_edges(athena_handle_t* batch, int32_t num_graphs,
                                                   const int32_t* num_vertices,
                                                   const int32_t* num_edges,
                                                   const int32_t* index_list,
                                                   const int32_t* num_entries_hint,
                                                   int32_t add_self_loops, int32_t mem,
                                                   int32_t validate) {
  ATH_TRY(ensure_init());
  ATH_REQUIRE(batch && num_vertices && num_edges, ATHENA_ERR_ARG,
              "batch_create_from_edges: null argument");
  ATH_REQUIRE(num_graphs >= 1, ATHENA_ERR_ARG, "batch_create_from_edges: num_graphs = %d",
              num_graphs);
  ATH_REQUIRE(mem == ATHENA_MEM_HOST || mem == ATHENA_MEM_DEVICE, ATHENA_ERR_ARG,
              "batch_create_from_edges: bad mem %d", mem);
  const int B = num_graphs;
  std::vector<int32_t> voff, eoff;
  ATH_TRY(graph_offsets(B, num_vertices, num_edges, voff, eoff, "batch_create_from_edges"));
  const int64_t V = voff[B], E = eoff[B];
  ATH_REQUIRE(E == 0 || index_list, ATHENA_ERR_ARG, "batch_create_from_edges: null index_list");
  const int64_t Zmax = 2 * E + (add_self_loops ? V : 0);
  cudaStream_t st = ctx().stream;
  // scratch: voff | eoff | nz | status | cnt(V+1) | hasself | cursor | key | nb | ia | (pad) ja
  const size_t Vp = (size_t)round_up(V + 2, 4), Zp = (size_t)round_up(Zmax + 2, 4);
  const size_t Bp = (size_t)round_up(B + 2, 4), Ep = (size_t)round_up(2 * E + 2, 4);
  const size_t ia_ints = (size_t)round_up(V + B + 2, 4);
  const int ntiles = (int)cdiv(V + 1, SCAN_TILE) + 1;
  size_t total = 3 * Bp + 4 + 3 * Vp + 2 * Zp + ia_ints + 2 * Zp + (size_t)round_up(ntiles, 4) +
                 (mem == ATHENA_MEM_HOST ? Ep : 0);
  DevBuf work;
  ATH_TRY(work.reserve(sizeof(int32_t) * total));
  int32_t* p = work.as<int32_t>();
  int32_t* d_voff = p; p += Bp;
  int32_t* d_eoff = p; p += Bp;
  int32_t* d_nz = p; p += Bp;
  int32_t* status = p; p += 4;
  int32_t* cnt = p; p += Vp;
  int32_t* hasself = p; p += Vp;
  int32_t* cursor = p; p += Vp;
  int32_t* key = p; p += Zp;
  int32_t* nb = p; p += Zp;
  int32_t* d_ia = p; p += ia_ints;
  int32_t* d_ja = p; p += 2 * Zp;
  int32_t* tile_sums = p; p += round_up(ntiles, 4);
  const int32_t* d_il = index_list;
  ATH_CUDA(cudaMemcpyAsync(d_voff, voff.data(), sizeof(int32_t) * (B + 1), cudaMemcpyHostToDevice, st));
  ATH_CUDA(cudaMemcpyAsync(d_eoff, eoff.data(), sizeof(int32_t) * (B + 1), cudaMemcpyHostToDevice, st));
  if (mem == ATHENA_MEM_HOST && E > 0) {
    // on the copy stream, like the adjacency of athena_cuda_batch_create: whatever the caller
    // uploads next (the features of the training step) queues right behind it instead of
    // waiting for the build kernels
    cudaEvent_t ev_il = nullptr;
    ATH_TRY(side_begin());
    ATH_TRY(side_copy(p, index_list, sizeof(int32_t) * 2 * (size_t)E));
    ATH_TRY(side_fence(&ev_il));
    ATH_TRY(main_wait(ev_il));
    d_il = p;
  } else if (E > 0) {
    ATH_REQUIRE((reinterpret_cast<uintptr_t>(index_list) & 7) == 0, ATHENA_ERR_ARG,
                "batch_create_from_edges: device index_list must be 8-byte aligned");
  }
  const int32_t status_init[4] = {INT_MAX, 0, 0, 0};
  ATH_CUDA(cudaMemcpyAsync(status, status_init, sizeof(status_init), cudaMemcpyHostToDevice, st));
  ATH_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * 2 * Vp, st));  // cnt and hasself
  ATH_CUDA(cudaMemsetAsync(d_nz, 0, sizeof(int32_t) * Bp, st));
  const int2* il2 = reinterpret_cast<const int2*>(d_il);
  if (E > 0) {
    k_edges_count<<<(int)cdiv(E, 256), 256, 0, st>>>(B, (int)E, d_voff, d_eoff, il2, cnt, hasself,
                                                    status);
    ATH_LAUNCHED();
  }
  if (add_self_loops && V > 0) {
    k_edges_loops<<<(int)cdiv(V, 256), 256, 0, st>>>((int)V, hasself, cnt);
    ATH_LAUNCHED();
  }
  ATH_TRY(exclusive_scan(cnt, cnt, (int)V, tile_sums));  // cnt -> row offsets, cnt[V] = Z
  if (V > 0) {
    ATH_CUDA(cudaMemcpyAsync(cursor, cnt, sizeof(int32_t) * (size_t)V, cudaMemcpyDeviceToDevice, st));
    const int64_t n = std::max<int64_t>(E, add_self_loops ? V : 0);
    if (n > 0) {
      k_edges_fill<<<(int)cdiv(n, 256), 256, 0, st>>>(B, (int)E, (int)V, add_self_loops ? 1 : 0,
                                                     d_voff, d_eoff, il2, hasself, cursor, key, nb);
      ATH_LAUNCHED();
    }
    // rows in ascending edge id (the fill claimed slots in arbitrary order); `cursor` is
    // free again and holds the list of rows longer than a warp
    k_csc_sort_short<<<(int)cdiv(V * 32, 256), 256, 0, st>>>((int)V, cnt, key, nb, cursor, status + 1);
    ATH_LAUNCHED();
    size_t smem = sizeof(unsigned long long) * LONG_SMEM_KEYS;
    ATH_CUDA(cudaFuncSetAttribute(k_csc_sort_long, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    k_csc_sort_long<<<ctx().sm_count, 1024, smem, st>>>(cnt, key, nb, cursor, status + 1);
    ATH_LAUNCHED();
    k_edges_emit<<<(int)cdiv(V * 8, 256), 256, 0, st>>>(B, (int)V, d_voff, cnt, key, nb, d_ia,
                                                       reinterpret_cast<int2*>(d_ja), d_nz);
    ATH_LAUNCHED();
  }
  // The entry counts per graph are data (self edges, missing loops): B integers come back
  // (through a pinned buffer: a pageable destination would stage the copy) -- unless the caller
  // already knows them (num_entries_hint, e.g. from an earlier epoch over the same graphs): the
  // build then stays asynchronous, and a wrong hint is caught by the row-pointer check of the
  // CSR build (ATHENA_ERR_GRAPH through validate / athena_cuda_batch_status).
  std::vector<int32_t> h_nz;
  if (num_entries_hint != nullptr) {
    h_nz.assign(num_entries_hint, num_entries_hint + B);
  } else {
    static int32_t* pinned_nz = nullptr;
    static size_t pinned_cap = 0;
    if ((size_t)B > pinned_cap) {
      if (pinned_nz) cudaFreeHost(pinned_nz);
      pinned_cap = (size_t)round_up(B, 1024);
      ATH_CUDA(cudaHostAlloc((void**)&pinned_nz, sizeof(int32_t) * pinned_cap, cudaHostAllocDefault));
    }
    ATH_CUDA(cudaMemcpyAsync(pinned_nz, d_nz, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
    ATH_CUDA(cudaStreamSynchronize(st));
    h_nz.assign(pinned_nz, pinned_nz + B);
  }
  athena_handle_t h = 0;
  ATH_TRY(athena_cuda_batch_create(&h, B, num_vertices, num_edges, h_nz.data(), d_ia, d_ja,
                                   ATHENA_MEM_DEVICE, 0));
  Batch* b = static_cast<Batch*>(lookup_object(h, Kind::Batch));
  k_merge_status<<<1, 1, 0, st>>>(b->status.as<int32_t>(), status);
  ATH_LAUNCHED();
  adopt(b->raw, work);
  *batch = h;
  if (validate) {
    int rc = athena_cuda_batch_status(h);
    if (rc != ATHENA_OK) {
      destroy_object(h, Kind::Batch);
      *batch = 0;
      return rc;
    }
  }
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_batch_create_from_edge_index(athena_handle_t* batch, int32_t num_graphs,
                                                        const int32_t* num_vertices,
                                                        const int32_t* num_edges,
                                                        const int32_t* num_entries,
                                                        const int64_t* edge_index,
                                                        const int64_t* degree, int32_t mem,
                                                        int32_t validate) {
  ATH_TRY(ensure_init());
  ATH_REQUIRE(batch && num_vertices && num_edges && num_entries, ATHENA_ERR_ARG,
              "batch_create_from_edge_index: null argument");
  ATH_REQUIRE(num_graphs >= 1, ATHENA_ERR_ARG, "batch_create_from_edge_index: num_graphs = %d",
              num_graphs);
  ATH_REQUIRE(mem == ATHENA_MEM_HOST || mem == ATHENA_MEM_DEVICE, ATHENA_ERR_ARG,
              "batch_create_from_edge_index: bad mem %d", mem);
  const int B = num_graphs;
  std::vector<int32_t> voff, zoff;
  ATH_TRY(graph_offsets(B, num_vertices, num_entries, voff, zoff, "batch_create_from_edge_index"));
  const int64_t V = voff[B], Z = zoff[B];
  ATH_REQUIRE((V == 0 || degree) && (Z == 0 || edge_index), ATHENA_ERR_ARG,
              "batch_create_from_edge_index: null edge_index / degree");
  cudaStream_t st = ctx().stream;
  const size_t Vp = (size_t)round_up(V + 2, 4), Zp = (size_t)round_up(Z + 2, 4);
  const size_t Bp = (size_t)round_up(B + 2, 4);
  const size_t ia_ints = (size_t)round_up(V + B + 2, 4);
  const int ntiles = (int)cdiv(V + 1, SCAN_TILE) + 1;
  size_t total = 2 * Bp + 4 + Vp + ia_ints + 2 * Zp + (size_t)round_up(ntiles, 4) +
                 (mem == ATHENA_MEM_HOST ? 2 * (3 * Zp + Vp) : 0);
  DevBuf work;
  ATH_TRY(work.reserve(sizeof(int32_t) * total));
  int32_t* p = work.as<int32_t>();
  int32_t* d_voff = p; p += Bp;
  int32_t* d_zoff = p; p += Bp;
  int32_t* status = p; p += 4;
  int32_t* cnt = p; p += Vp;
  int32_t* d_ia = p; p += ia_ints;
  int32_t* d_ja = p; p += 2 * Zp;
  int32_t* tile_sums = p; p += round_up(ntiles, 4);
  const long long* d_ei = reinterpret_cast<const long long*>(edge_index);
  const long long* d_deg = reinterpret_cast<const long long*>(degree);
  if (mem == ATHENA_MEM_HOST) {
    long long* q = reinterpret_cast<long long*>(p);
    if (Z > 0)
      ATH_CUDA(cudaMemcpyAsync(q, edge_index, sizeof(int64_t) * 3 * (size_t)Z, cudaMemcpyHostToDevice, st));
    d_ei = q;
    q += 3 * Zp;
    if (V > 0)
      ATH_CUDA(cudaMemcpyAsync(q, degree, sizeof(int64_t) * (size_t)V, cudaMemcpyHostToDevice, st));
    d_deg = q;
  }
  ATH_CUDA(cudaMemcpyAsync(d_voff, voff.data(), sizeof(int32_t) * (B + 1), cudaMemcpyHostToDevice, st));
  ATH_CUDA(cudaMemcpyAsync(d_zoff, zoff.data(), sizeof(int32_t) * (B + 1), cudaMemcpyHostToDevice, st));
  const int32_t status_init[4] = {INT_MAX, 0, 0, 0};
  ATH_CUDA(cudaMemcpyAsync(status, status_init, sizeof(status_init), cudaMemcpyHostToDevice, st));
  if (V > 0) {
    k_onnx_rows<<<(int)cdiv(V, 256), 256, 0, st>>>(B, (int)V, d_voff, d_zoff, d_deg, cnt);
    ATH_LAUNCHED();
  }
  ATH_TRY(exclusive_scan(cnt, cnt, (int)V, tile_sums));
  const int64_t n = std::max(V, Z);
  if (n > 0) {
    k_onnx_emit<<<(int)cdiv(n, 256), 256, 0, st>>>(B, (int)V, (int)Z, d_voff, d_zoff, cnt, d_ei,
                                                  d_ia, reinterpret_cast<int2*>(d_ja), status);
    ATH_LAUNCHED();
  }
  athena_handle_t h = 0;
  ATH_TRY(athena_cuda_batch_create(&h, B, num_vertices, num_edges, num_entries, d_ia, d_ja,
                                   ATHENA_MEM_DEVICE, 0));
  Batch* b = static_cast<Batch*>(lookup_object(h, Kind::Batch));
  k_merge_status<<<1, 1, 0, st>>>(b->status.as<int32_t>(), status);
  ATH_LAUNCHED();
  adopt(b->raw, work);
  *batch = h;
  if (validate) {
    int rc = athena_cuda_batch_status(h);
    if (rc != ATHENA_OK) {
      destroy_object(h, Kind::Batch);
      *batch = 0;
      return rc;
    }
  }
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_batch_destroy(athena_handle_t batch) {
  return destroy_object(batch, Kind::Batch);
}

ATHENA_API int athena_cuda_batch_status(athena_handle_t batch) {
  Batch* b = static_cast<Batch*>(lookup_object(batch, Kind::Batch));
  if (!b) return ATHENA_ERR_HANDLE;
  int32_t st[4];
  ATH_CUDA(cudaMemcpyAsync(st, b->status.p, sizeof(st), cudaMemcpyDeviceToHost, ctx().stream));
  ATH_CUDA(cudaStreamSynchronize(ctx().stream));
  b->multi_edges = st[2] != 0 ? 1 : 0;
  ATH_REQUIRE(st[0] == INT_MAX, ATHENA_ERR_GRAPH,
              "graph adjacency matrix has indices greater than the number of vertices "
              "(or inconsistent adj_ia) in sample %d",
              st[0] + 1);
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_batch_info(athena_handle_t batch, int32_t* num_graphs,
                                      int64_t* num_vertices, int64_t* num_entries,
                                      int64_t* num_edges) {
  Batch* b = static_cast<Batch*>(lookup_object(batch, Kind::Batch));
  if (!b) return ATHENA_ERR_HANDLE;
  if (num_graphs) *num_graphs = b->B;
  if (num_vertices) *num_vertices = b->V;
  if (num_entries) *num_entries = b->Z;
  if (num_edges) *num_edges = b->E;
  return ATHENA_OK;
}

ATHENA_API int athena_cuda_batch_bucketize(athena_handle_t batch, int32_t min_degree,
                                           int32_t max_degree) {
  Batch* b = static_cast<Batch*>(lookup_object(batch, Kind::Batch));
  if (!b) return ATHENA_ERR_HANDLE;
  return batch_bucketize(b, min_degree, max_degree, nullptr);
}

ATHENA_API int athena_cuda_batch_export(athena_handle_t batch, int32_t what, void* host_out,
                                        int64_t count) {
  Batch* b = static_cast<Batch*>(lookup_object(batch, Kind::Batch));
  if (!b) return ATHENA_ERR_HANDLE;
  ATH_REQUIRE(host_out, ATHENA_ERR_ARG, "batch_export: null output");
  const void* src = nullptr;
  int64_t n = 0;
  BucketSet* bs = b->buckets.empty() ? nullptr : b->buckets.back().get();
  switch (what) {
    case ATHENA_BATCH_ROW_PTR: src = b->row_ptr; n = b->V + 1; break;
    case ATHENA_BATCH_COL: src = b->col; n = b->Z; break;
    case ATHENA_BATCH_EID: src = b->eid; n = b->Z; break;
    case ATHENA_BATCH_DEG: src = b->deg; n = b->V; break;
    case ATHENA_BATCH_VGRAPH: src = b->vgraph; n = b->V; break;
    case ATHENA_BATCH_CSC_PTR: src = b->csc_ptr; n = b->V + 1; break;
    case ATHENA_BATCH_CSC_SRC: src = b->csc_src; n = b->Z; break;
    case ATHENA_BATCH_CSC_ENT: src = b->csc_ent; n = b->Z; break;
    case ATHENA_BATCH_COEF: src = b->coef; n = b->Z; break;
    case ATHENA_BATCH_BUCKET:
    case ATHENA_BATCH_PERM:
    case ATHENA_BATCH_BUCKET_PTR:
      ATH_REQUIRE(bs, ATHENA_ERR_STATE, "batch_export: call athena_cuda_batch_bucketize first");
      if (what == ATHENA_BATCH_BUCKET) { src = bs->bkt.p; n = b->V; }
      else if (what == ATHENA_BATCH_PERM) { src = bs->perm.p; n = b->V; }
      else { src = bs->bkt_ptr.p; n = bs->D + 1; }
      break;
    default:
      ATH_REQUIRE(false, ATHENA_ERR_ARG, "batch_export: unknown selector %d", what);
  }
  ATH_REQUIRE(count == n, ATHENA_ERR_ARG, "batch_export: selector %d has %lld elements, got %lld",
              what, (long long)n, (long long)count);
  if (n > 0)
    ATH_CUDA(cudaMemcpyAsync(host_out, src, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost,
                             ctx().stream));
  ATH_CUDA(cudaStreamSynchronize(ctx().stream));
  return ATHENA_OK;
}
