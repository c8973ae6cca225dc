// Device-side h-hop local-subgraph extraction (the producer of the hot path's input):
// replaces Subgraphs.generate_subgraph / generate_subgraph_link_pred
// (reference G-Meta/subgraph_data_processing.py:295-346) for a whole meta-batch of requests, and
// writes the result directly in the packed-set layout of include/gmeta_b200.h (CSR over the
// concatenated subgraphs with batch offsets applied, parent ids, centre rows), so the meta-batch
// never leaves HBM.
//
//   request r = (centre a, optional second centre b, node range [lo, hi) of its graph)
//   1. closure: <= h hops over IN-edges from a (link prediction: 2 hops from a, 1 hop from b --
//      the reference's inner comprehension at :332 re-reads G.in_edges(j)), a and b included;
//   2. if |closure| > sample_nodes: a uniform sample of sample_nodes nodes without replacement
//      (the sample_nodes smallest values of a counter-based hash of (seed, r, node)), then the
//      centre(s) re-added (:312-314, :337-339);
//   3. nodes sorted by parent id (np.unique order), node-induced subgraph with every parent edge
//      whose endpoints are both selected (multiplicity kept), local ids = rank in the sorted list.
//
// One CTA per request.  Kernel 1 (select) leaves the sorted node list, the row degrees and the
// counts in per-request slabs; a one-block scan turns the counts into packed offsets; kernel 2
// (build) writes indptr / indices / parent ids / centre rows at those offsets.  All integer work:
// results are bit-exact w.r.t. the host extractor whenever the closure fits sample_nodes.
#include "common.cuh"

namespace gmeta {
namespace {

constexpr int KH_THREADS = 256;
constexpr int KH_MAX_NODES = 2048;      // sample_nodes + 2 must fit (shared-memory node list / bitonic sort)
constexpr int KH_BINS = 2048;

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t sample_key(uint64_t seed, int req, int node) {
  return mix32((uint32_t)node * 0x9E3779B9U ^ mix32((uint32_t)seed ^ (uint32_t)req * 0x85EBCA6BU) ^ (uint32_t)(seed >> 32));
}

struct KhopParams {
  const int32_t* indptr;      // parent CSR by destination over the concatenated graphs
  const int32_t* indices;
  const int32_t* req_a;       // [R] first centre (global id)
  const int32_t* req_b;       // [R] second centre or NULL
  const int32_t* req_lo;      // [R] first node of the request's graph
  const int32_t* req_hi;      // [R] one past its last node
  int n_req;
  int hops_a, hops_b;
  int sample_nodes;
  uint64_t seed;
  int slab;                   // ints per request in nodes_slab / deg_slab (>= sample_nodes + 2)
  int32_t* nodes_slab;        // [R][slab] sorted selected nodes (global ids)
  int32_t* deg_slab;          // [R][slab] induced in-degree of each selected node
  int32_t* n_nodes;           // [R]
  int32_t* n_edges;           // [R]
  int32_t* closure_size;      // [R] size of the closure before sampling (diagnostics / tests)
  uint32_t* scratch;          // per CTA: bitmap words + closure list
  long long scratch_stride;   // uint32 per CTA
  int max_graph_nodes;
};

// membership / rank of `u` in the sorted shared-memory list nodes[0..n)
__device__ __forceinline__ int rank_of(const int* nodes, int n, int u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (nodes[mid] < u) lo = mid + 1; else hi = mid;
  }
  return (lo < n && nodes[lo] == u) ? lo : -1;
}

__device__ void bitonic_sort(int* a, int n_pow2) {     // ascending, whole CTA
  for (int k = 2; k <= n_pow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const int x = a[i], y = a[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
}

__global__ void __launch_bounds__(KH_THREADS) khop_select_kernel(const KhopParams p) {
  __shared__ int s_nodes[KH_MAX_NODES];
  __shared__ int s_hist[KH_BINS];
  __shared__ int s_cnt, s_n, s_ties[64], s_nties, s_bin, s_below;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = KH_THREADS / 32;
  uint32_t* bitmap = p.scratch + (size_t)blockIdx.x * p.scratch_stride;
  int* list = reinterpret_cast<int*>(bitmap + (p.max_graph_nodes + 31) / 32);
  for (int r = blockIdx.x; r < p.n_req; r += gridDim.x) {
    const int lo = p.req_lo[r], hi = p.req_hi[r], a = p.req_a[r], b = p.req_b ? p.req_b[r] : -1;
    const int words = (hi - lo + 31) / 32;
    for (int w = threadIdx.x; w < words; w += blockDim.x) bitmap[w] = 0u;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    // ---- closure: BFS over in-edges from a, frontier by frontier; then b and (hops_b == 1) its in-neighbours ----
    // in-neighbours of the `n_front` nodes front(f) that are not in the set yet are appended to the list
    auto expand = [&](int n_front, auto front) {
      for (int f = warp; f < n_front; f += nwarps) {
        const int v = front(f);
        const int e0 = p.indptr[v], e1 = p.indptr[v + 1];
        for (int base = e0; base < e1; base += 32) {
          const int e = base + lane;
          bool fresh = false;
          int u = 0;
          if (e < e1) {
            u = p.indices[e];
            const uint32_t bit = 1u << ((u - lo) & 31);
            const uint32_t old = atomicOr(&bitmap[(u - lo) >> 5], bit);
            fresh = !(old & bit);
          }
          const unsigned m = __ballot_sync(0xffffffffu, fresh);
          int base_out = 0;
          if (lane == 0 && m) base_out = atomicAdd(&s_cnt, __popc(m));
          base_out = __shfl_sync(0xffffffffu, base_out, 0);
          if (fresh) list[base_out + __popc(m & ((1u << lane) - 1u))] = u;
        }
      }
      __syncthreads();
    };
    if (threadIdx.x == 0) {
      bitmap[(a - lo) >> 5] |= 1u << ((a - lo) & 31);
      list[0] = a;
      s_cnt = 1;
    }
    __syncthreads();
    {
      int beg = 0, end = 1;
      for (int hop = 0; hop < p.hops_a && beg < end; ++hop) {
        expand(end - beg, [&](int f) { return list[beg + f]; });
        beg = end;
        end = s_cnt;
      }
    }
    if (b >= 0) {
      if (threadIdx.x == 0) {
        const uint32_t bit = 1u << ((b - lo) & 31);
        if (!(bitmap[(b - lo) >> 5] & bit)) { bitmap[(b - lo) >> 5] |= bit; list[s_cnt++] = b; }
      }
      __syncthreads();
      if (p.hops_b >= 1) expand(1, [&](int) { return b; });
    }
    const int count = s_cnt;
    if (threadIdx.x == 0) p.closure_size[r] = count;
    // ---- selection ----
    int n_sel;
    if (count <= p.sample_nodes) {
      for (int i = threadIdx.x; i < count; i += blockDim.x) s_nodes[i] = list[i];
      n_sel = count;
      __syncthreads();
    } else {
      // radix select of the sample_nodes-th smallest key: 11 + 11 + 10 bits
      uint32_t prefix = 0, prefix_mask = 0;
      int need = p.sample_nodes;                 // keys still to take among those matching the prefix
      for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
        const int bits = pass == 2 ? 10 : 11;
        for (int i = threadIdx.x; i < KH_BINS; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < count; i += blockDim.x) {
          const uint32_t k = sample_key(p.seed, r, list[i]);
          if ((k & prefix_mask) == prefix) atomicAdd(&s_hist[(k >> shift) & ((1u << bits) - 1u)], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          int cum = 0, bin = 0;
          for (; bin < (1 << bits); ++bin) {
            if (cum + s_hist[bin] >= need) break;
            cum += s_hist[bin];
          }
          s_bin = bin; s_below = cum;
        }
        __syncthreads();
        prefix |= (uint32_t)s_bin << shift;
        prefix_mask |= ((1u << bits) - 1u) << shift;
        need -= s_below;
        __syncthreads();
      }
      // keys < prefix are all taken; `need` of the keys == prefix (ties), lowest node ids first
      if (threadIdx.x == 0) { s_n = 0; s_nties = 0; }
      __syncthreads();
      for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const int u = list[i];
        const uint32_t k = sample_key(p.seed, r, u);
        if (k < prefix) s_nodes[atomicAdd(&s_n, 1)] = u;
        else if (k == prefix) { const int t = atomicAdd(&s_nties, 1); if (t < 64) s_ties[t] = u; }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        const int nt = s_nties < 64 ? s_nties : 64;
        for (int i = 1; i < nt; ++i) {           // insertion sort of the (almost always single) tie
          const int x = s_ties[i];
          int j = i - 1;
          for (; j >= 0 && s_ties[j] > x; --j) s_ties[j + 1] = s_ties[j];
          s_ties[j + 1] = x;
        }
        for (int i = 0; i < need && i < nt; ++i) s_nodes[s_n++] = s_ties[i];
        // centre(s) re-added (np.unique(np.append(...)) -- duplicates removed by the sort + unique below)
        s_nodes[s_n++] = a;
        if (b >= 0) s_nodes[s_n++] = b;
      }
      __syncthreads();
      n_sel = s_n;
    }
    // ---- sort ascending + unique ----
    int n_pow2 = 1;
    while (n_pow2 < n_sel) n_pow2 <<= 1;
    for (int i = n_sel + threadIdx.x; i < n_pow2; i += blockDim.x) s_nodes[i] = 0x7fffffff;
    __syncthreads();
    bitonic_sort(s_nodes, n_pow2);
    if (count > p.sample_nodes) {                // at most two duplicates (the re-added centres)
      if (threadIdx.x == 0) {
        int w = 0;
        for (int i = 0; i < n_sel; ++i)
          if (i == 0 || s_nodes[i] != s_nodes[i - 1]) s_nodes[w++] = s_nodes[i];
        s_n = w;
      }
      __syncthreads();
      n_sel = s_n;
    }
    // ---- induced in-degrees ----
    int32_t* nodes_out = p.nodes_slab + (size_t)r * p.slab;
    int32_t* deg_out = p.deg_slab + (size_t)r * p.slab;
    int my_edges = 0;
    for (int k = warp; k < n_sel; k += nwarps) {
      const int v = s_nodes[k];
      const int e0 = p.indptr[v], e1 = p.indptr[v + 1];
      int c = 0;
      for (int e = e0 + lane; e < e1; e += 32) c += rank_of(s_nodes, n_sel, p.indices[e]) >= 0;
#pragma unroll
      for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0) { nodes_out[k] = v; deg_out[k] = c; my_edges += c; }
    }
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    if (lane == 0 && my_edges) atomicAdd(&s_cnt, my_edges);
    __syncthreads();
    if (threadIdx.x == 0) { p.n_nodes[r] = n_sel; p.n_edges[r] = s_cnt; }
    __syncthreads();
  }
}

// exclusive scans of the per-request counts -> packed offsets (one block)
__global__ void __launch_bounds__(1024) khop_scan_kernel(const int32_t* __restrict__ n_nodes,
                                                         const int32_t* __restrict__ n_edges, int n_req,
                                                         int32_t* __restrict__ node_ptr, int32_t* __restrict__ edge_ptr) {
  __shared__ int sn[1024], se[1024];
  __shared__ int carry_n, carry_e;
  if (threadIdx.x == 0) { carry_n = 0; carry_e = 0; }
  __syncthreads();
  for (int base = 0; base < n_req; base += 1024) {
    const int i = base + threadIdx.x;
    const int vn = i < n_req ? n_nodes[i] : 0, ve = i < n_req ? n_edges[i] : 0;
    sn[threadIdx.x] = vn; se[threadIdx.x] = ve;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      const int an = threadIdx.x >= d ? sn[threadIdx.x - d] : 0, ae = threadIdx.x >= d ? se[threadIdx.x - d] : 0;
      __syncthreads();
      sn[threadIdx.x] += an; se[threadIdx.x] += ae;
      __syncthreads();
    }
    if (i < n_req) { node_ptr[i] = carry_n + sn[threadIdx.x] - vn; edge_ptr[i] = carry_e + se[threadIdx.x] - ve; }
    __syncthreads();
    if (threadIdx.x == 1023) { carry_n += sn[1023]; carry_e += se[1023]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { node_ptr[n_req] = carry_n; edge_ptr[n_req] = carry_e; }
}

struct BuildParams {
  const int32_t* indptr;
  const int32_t* indices;
  const int32_t* req_a;
  const int32_t* req_b;
  const int32_t* req_lo;
  int n_req;
  int slab;
  const int32_t* nodes_slab;
  const int32_t* deg_slab;
  const int32_t* n_nodes;
  const int32_t* node_ptr;
  const int32_t* edge_ptr;
  int32_t* out_indptr;     // [N_total + 1] packed CSR by destination
  int32_t* out_indices;    // [E_total] packed row ids
  int32_t* out_parent;     // [N_total] node id inside its graph
  int32_t* out_global;     // [N_total] global id (= feature-table row), or NULL
  int32_t* out_centre;     // [R * (1 or 2)] packed row of the centre node(s)
};

__global__ void __launch_bounds__(KH_THREADS) khop_build_kernel(const BuildParams p) {
  __shared__ int s_nodes[KH_MAX_NODES];
  __shared__ int s_off[KH_MAX_NODES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = KH_THREADS / 32;
  for (int r = blockIdx.x; r < p.n_req; r += gridDim.x) {
    const int n = p.n_nodes[r], nb = p.node_ptr[r], eb = p.edge_ptr[r], lo = p.req_lo[r];
    const int32_t* nodes = p.nodes_slab + (size_t)r * p.slab;
    const int32_t* deg = p.deg_slab + (size_t)r * p.slab;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_nodes[i] = nodes[i];
    __syncthreads();
    if (warp == 0) {                    // exclusive scan of the row degrees (n <= 2048)
      int carry = 0;
      for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const int v = i < n ? deg[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += y;
        }
        if (i < n) s_off[i] = carry + x - v;
        carry += __shfl_sync(0xffffffffu, x, 31);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      p.out_indptr[nb + i] = eb + s_off[i];
      p.out_parent[nb + i] = s_nodes[i] - lo;
      if (p.out_global) p.out_global[nb + i] = s_nodes[i];
    }
    if (r == p.n_req - 1 && threadIdx.x == 0) p.out_indptr[nb + n] = p.edge_ptr[p.n_req];
    if (threadIdx.x == 0) {
      if (p.req_b) {
        p.out_centre[2 * r] = nb + rank_of(s_nodes, n, p.req_a[r]);
        p.out_centre[2 * r + 1] = nb + rank_of(s_nodes, n, p.req_b[r]);
      } else {
        p.out_centre[r] = nb + rank_of(s_nodes, n, p.req_a[r]);
      }
    }
    for (int k = warp; k < n; k += nwarps) {      // one warp per row, parent CSR order kept
      const int v = s_nodes[k];
      const int e0 = p.indptr[v], e1 = p.indptr[v + 1];
      int w = eb + s_off[k];
      for (int base = e0; base < e1; base += 32) {
        const int e = base + lane;
        int loc = -1;
        if (e < e1) loc = rank_of(s_nodes, n, p.indices[e]);
        const unsigned m = __ballot_sync(0xffffffffu, loc >= 0);
        if (loc >= 0) p.out_indices[w + __popc(m & ((1u << lane) - 1u))] = nb + loc;
        w += __popc(m);
      }
    }
    __syncthreads();
  }
}

// one CTA per request, eight resident per SM: the kernels are chains of dependent index loads and barriers, so
// they are paid for in latency and want every warp slot of the SM (16 KB of shared memory and 32 registers each)
inline int khop_grid(int n_req) { return n_req < 8 * kNumSMs ? n_req : 8 * kNumSMs; }
inline int64_t khop_scratch_stride(int max_graph_nodes) {
  return ((int64_t)(max_graph_nodes + 31) / 32 + max_graph_nodes + 64 + 63) / 64 * 64;   // uint32 per CTA
}

}  // namespace
}  // namespace gmeta

using namespace gmeta;

extern "C" int64_t gmeta_khop_workspace_bytes(int32_t n_req, int32_t sample_nodes, int32_t max_graph_nodes) {
  if (n_req < 0 || sample_nodes <= 0 || max_graph_nodes <= 0) return -1;
  const int64_t slab = (sample_nodes + 2 + 3) / 4 * 4;
  return 2 * (int64_t)n_req * slab * 4 + 3 * ((int64_t)n_req * 4 + 256) +
         (int64_t)khop_grid(n_req > 0 ? n_req : 1) * khop_scratch_stride(max_graph_nodes) * 4 + 1024;
}

namespace {
struct KhopWs {
  int32_t *nodes_slab, *deg_slab, *n_nodes, *n_edges, *closure;
  uint32_t* scratch;
  int slab;
};
KhopWs khop_carve(void* ws, int n_req, int sample_nodes, int max_graph_nodes) {
  KhopWs k;
  k.slab = (sample_nodes + 2 + 3) / 4 * 4;
  char* p = reinterpret_cast<char*>(ws);
  auto take = [&](int64_t bytes) { char* q = p; p += (bytes + 255) / 256 * 256; return q; };
  k.nodes_slab = reinterpret_cast<int32_t*>(take((int64_t)n_req * k.slab * 4));
  k.deg_slab = reinterpret_cast<int32_t*>(take((int64_t)n_req * k.slab * 4));
  k.n_nodes = reinterpret_cast<int32_t*>(take((int64_t)n_req * 4));
  k.n_edges = reinterpret_cast<int32_t*>(take((int64_t)n_req * 4));
  k.closure = reinterpret_cast<int32_t*>(take((int64_t)n_req * 4));
  k.scratch = reinterpret_cast<uint32_t*>(take(0));
  (void)max_graph_nodes;
  return k;
}
}  // namespace

extern "C" int gmeta_khop_select(const int32_t* indptr, const int32_t* indices, const int32_t* req_a,
                                 const int32_t* req_b, const int32_t* req_lo, const int32_t* req_hi, int32_t n_req,
                                 int32_t hops_a, int32_t hops_b, int32_t sample_nodes, int32_t max_graph_nodes,
                                 uint64_t seed, int32_t* node_ptr, int32_t* edge_ptr, int32_t* closure_size,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
  if (!indptr || !indices || !req_a || !req_lo || !req_hi || !node_ptr || !edge_ptr || !workspace) return GMETA_ERR_BAD_ARG;
  if (n_req < 0 || hops_a < 0 || hops_a > GMETA_MAX_LAYERS || hops_b < 0 || sample_nodes <= 0) return GMETA_ERR_BAD_ARG;
  if (hops_b > 1) return GMETA_ERR_UNSUPPORTED;   // the reference only ever takes ONE hop from the second endpoint (:332)
  if (sample_nodes + 2 > KH_MAX_NODES) return GMETA_ERR_UNSUPPORTED;
  if (workspace_bytes < gmeta_khop_workspace_bytes(n_req, sample_nodes, max_graph_nodes)) return GMETA_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return GMETA_ERR_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_req == 0) {
    return cudaMemsetAsync(node_ptr, 0, 4, s) == cudaSuccess && cudaMemsetAsync(edge_ptr, 0, 4, s) == cudaSuccess
               ? GMETA_OK : GMETA_ERR_LAUNCH;
  }
  KhopWs k = khop_carve(workspace, n_req, sample_nodes, max_graph_nodes);
  KhopParams p;
  p.indptr = indptr; p.indices = indices; p.req_a = req_a; p.req_b = req_b; p.req_lo = req_lo; p.req_hi = req_hi;
  p.n_req = n_req; p.hops_a = hops_a; p.hops_b = hops_b; p.sample_nodes = sample_nodes; p.seed = seed;
  p.slab = k.slab; p.nodes_slab = k.nodes_slab; p.deg_slab = k.deg_slab; p.n_nodes = k.n_nodes; p.n_edges = k.n_edges;
  p.closure_size = closure_size ? closure_size : k.closure;
  p.scratch = k.scratch; p.scratch_stride = khop_scratch_stride(max_graph_nodes); p.max_graph_nodes = max_graph_nodes;
  khop_select_kernel<<<khop_grid(n_req), KH_THREADS, 0, s>>>(p);
  int rc = check_launch();
  if (rc != GMETA_OK) return rc;
  khop_scan_kernel<<<1, 1024, 0, s>>>(k.n_nodes, k.n_edges, n_req, node_ptr, edge_ptr);
  return check_launch();
}

extern "C" int gmeta_khop_build(const int32_t* indptr, const int32_t* indices, const int32_t* req_a,
                                const int32_t* req_b, const int32_t* req_lo, int32_t n_req, int32_t sample_nodes,
                                int32_t max_graph_nodes, const int32_t* node_ptr, const int32_t* edge_ptr,
                                int32_t* out_indptr, int32_t* out_indices, int32_t* out_parent, int32_t* out_global,
                                int32_t* out_centre, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!indptr || !indices || !req_a || !req_lo || !node_ptr || !edge_ptr || !out_indptr || !out_indices ||
      !out_parent || !out_centre || !workspace)
    return GMETA_ERR_BAD_ARG;
  if (n_req < 0 || sample_nodes <= 0) return GMETA_ERR_BAD_ARG;
  if (workspace_bytes < gmeta_khop_workspace_bytes(n_req, sample_nodes, max_graph_nodes)) return GMETA_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_req == 0) return cudaMemsetAsync(out_indptr, 0, 4, s) == cudaSuccess ? GMETA_OK : GMETA_ERR_LAUNCH;
  KhopWs k = khop_carve(workspace, n_req, sample_nodes, max_graph_nodes);
  BuildParams p;
  p.indptr = indptr; p.indices = indices; p.req_a = req_a; p.req_b = req_b; p.req_lo = req_lo; p.n_req = n_req;
  p.slab = k.slab; p.nodes_slab = k.nodes_slab; p.deg_slab = k.deg_slab; p.n_nodes = k.n_nodes;
  p.node_ptr = node_ptr; p.edge_ptr = edge_ptr;
  p.out_indptr = out_indptr; p.out_indices = out_indices; p.out_parent = out_parent; p.out_global = out_global;
  p.out_centre = out_centre;
  khop_build_kernel<<<khop_grid(n_req), KH_THREADS, 0, s>>>(p);
  return check_launch();
}
