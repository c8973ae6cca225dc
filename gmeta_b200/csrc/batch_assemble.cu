// Packed-set assembly on the device: everything gmeta_packed_set_t needs on top of what the h-hop extractor
// (khop.cu) emits -- the CSR by source, the task row pointers and row-tile tables, the active-row lists of the
// pruned backward with their task pointers / tile tables, and the centre positions -- as a handful of CUDA passes
// over the packed int32 arrays, without a host round trip.  Replaces the structure work of dgl.batch
// (reference G-Meta/subgraph_data_processing.py:399-406) and of packing.pack_meta_batch on the host path; the
// results are bit-identical to the host packer's (integer work, tests/test_gpu_device_batch.py).
//
// Determinism: atomics are used only on counters whose final value does not depend on the order (degree
// histogram) or whose order is erased afterwards (the fill cursor: every by-source list is sorted ascending, the
// rule of packed.csr_transpose: destinations ascending inside a row).
#include "common.cuh"

namespace gmeta {
namespace {

constexpr int TILE = GMETA_TILE_ROWS;
constexpr int SCAN_ITEMS = 4;                    // elements per thread of a scan block
constexpr int SCAN_BLOCK = 1024 * SCAN_ITEMS;

__device__ __forceinline__ int warp_incl_scan_i(int x, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  return x;
}
// exclusive scan of one value per thread over a 1024-thread CTA; returns the CTA total through `total`
__device__ __forceinline__ int block_excl_scan_1024(int x, int* sh /* [33] */, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int incl = warp_incl_scan_i(x, lane);
  if (lane == 31) sh[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = sh[lane];
    const int wi = warp_incl_scan_i(w, lane);
    sh[lane] = wi - w;
    if (lane == 31) sh[32] = wi;
  }
  __syncthreads();
  const int r = sh[warp] + incl - x;
  total = sh[32];
  __syncthreads();
  return r;
}

// ---- device-wide exclusive scan of int32 (or of a 0/1 byte-flag array): block sums, their scan, apply ----
template <typename T>
__global__ void __launch_bounds__(1024) scan_sums_kernel(const T* __restrict__ a, int n, int* __restrict__ block_sum) {
  __shared__ int sh[33];
  const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) s += base + j < n ? (int)a[base + j] : 0;
  int total;
  block_excl_scan_1024(s, sh, total);
  if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(int* __restrict__ block_sum, int n_blocks) {
  __shared__ int sh[33];
  int carry = 0;
  for (int b0 = 0; b0 < n_blocks; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const int x = i < n_blocks ? block_sum[i] : 0;
    int total;
    const int ex = block_excl_scan_1024(x, sh, total);
    if (i < n_blocks) block_sum[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) block_sum[n_blocks] = carry;
}
// out[i] = sum a[0..i), i <= n (out has n + 1 entries)
template <typename T>
__global__ void __launch_bounds__(1024) scan_apply_kernel(const T* __restrict__ a, int n, const int* __restrict__ block_sum,
                                                          int32_t* __restrict__ out) {
  __shared__ int sh[33];
  const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    v[j] = base + j < n ? (int)a[base + j] : 0;
    s += v[j];
  }
  int total;
  int run = block_sum[blockIdx.x] + block_excl_scan_1024(s, sh, total);
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    if (base + j < n) out[base + j] = run;
    run += v[j];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 1023) out[n] = block_sum[gridDim.x];
}
template <typename T>
int excl_scan(const T* a, int n, int32_t* out, int* block_sum, cudaStream_t s) {
  const int nb = n > 0 ? ceil_div(n, SCAN_BLOCK) : 1;
  scan_sums_kernel<T><<<nb, 1024, 0, s>>>(a, n, block_sum);
  int rc = check_launch();
  if (rc != GMETA_OK) return rc;
  scan_block_sums_kernel<<<1, 1024, 0, s>>>(block_sum, nb);
  if ((rc = check_launch()) != GMETA_OK) return rc;
  scan_apply_kernel<T><<<nb, 1024, 0, s>>>(a, n, block_sum, out);
  return check_launch();
}

// ---- task row pointers and row tiles (learner.tile_table: tiles of <= 128 rows that never straddle a task) ----
// ptr_src: [n_tasks + 1] row pointers; when `via` is given the pointer of task t is ptr_src[via[t]] (first row of
// the task's first subgraph) and it is also written to ptr_out.  counts[0] = number of tiles, counts[1] = largest
// task (rows), when counts is given.  One CTA; one warp per task writes the task's tiles.
__global__ void __launch_bounds__(1024) tile_table_kernel(const int32_t* __restrict__ ptr_src, const int32_t* __restrict__ via,
                                                          int n_tasks, int32_t* __restrict__ ptr_out,
                                                          int32_t* __restrict__ tile_row0, int32_t* __restrict__ tile_nrows,
                                                          int32_t* __restrict__ tile_task, int32_t* __restrict__ n_tiles_out,
                                                          int32_t* __restrict__ max_rows_out) {
  __shared__ int sh[33];
  __shared__ int s_first[1024];
  __shared__ int s_max;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  auto rowptr = [&](int t) { return via ? ptr_src[via[t]] : ptr_src[t]; };
  if (ptr_out)
    for (int t = threadIdx.x; t <= n_tasks; t += 1024) ptr_out[t] = rowptr(t);
  int carry = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = 0; t0 < n_tasks; t0 += 1024) {
    const int t = t0 + threadIdx.x;
    int rows = 0;
    if (t < n_tasks) rows = rowptr(t + 1) - rowptr(t);
    const int nt = (rows + TILE - 1) / TILE;
    if (rows > 0) atomicMax(&s_max, rows);
    int total;
    const int ex = block_excl_scan_1024(nt, sh, total);
    s_first[threadIdx.x] = carry + ex;
    __syncthreads();
    for (int j = warp; j < 1024 && t0 + j < n_tasks; j += 32) {
      const int tt = t0 + j;
      const int r0 = rowptr(tt), r1 = rowptr(tt + 1), first = s_first[j];
      for (int k = lane; k * TILE < r1 - r0; k += 32) {
        tile_row0[first + k] = r0 + k * TILE;
        tile_nrows[first + k] = min(TILE, r1 - r0 - k * TILE);
        tile_task[first + k] = tt;
      }
    }
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (n_tiles_out) *n_tiles_out = carry;
    if (max_rows_out) *max_rows_out = s_max;
  }
}

// ---- CSR by source ----
__global__ void out_degree_kernel(const int32_t* __restrict__ indices, int n_edges, int32_t* __restrict__ cnt) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += gridDim.x * blockDim.x) atomicAdd(cnt + indices[e], 1);
}
// one warp per destination row: its in-edges are dropped into their sources' lists at an atomic cursor
__global__ void transpose_fill_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int n_rows,
                                      int32_t* __restrict__ cursor, int32_t* __restrict__ tmp) {
  const int lane = threadIdx.x & 31;
  for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < n_rows; v += (gridDim.x * blockDim.x) >> 5) {
    const int beg = indptr[v], end = indptr[v + 1];
    for (int e = beg + lane; e < end; e += 32) tmp[atomicAdd(cursor + indices[e], 1)] = v;
  }
}
// By-source lists sorted ascending from tmp into t_indices -- erases the order the atomic cursor produced.
// Pass 1, one warp per source row: lists of <= 32 entries are ranked by counting (element i goes to the number of
// smaller elements, ties by position); longer ones are queued.  Pass 2, one CTA per queued row: bitonic sort in
// shared memory (<= SORT_SMEM entries; a packed subgraph has at most 2048 nodes), rank by counting beyond that.
constexpr int SORT_SMEM = 4096;
__global__ void sort_short_lists_kernel(const int32_t* __restrict__ t_indptr, int n_rows, const int32_t* __restrict__ tmp,
                                        int32_t* __restrict__ t_indices, int32_t* __restrict__ long_rows,
                                        int32_t* __restrict__ n_long) {
  const int lane = threadIdx.x & 31;
  for (int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n_rows; u += (gridDim.x * blockDim.x) >> 5) {
    const int beg = t_indptr[u], n = t_indptr[u + 1] - beg;
    if (n <= 0) continue;
    if (n > 32) {
      if (lane == 0) long_rows[atomicAdd(n_long, 1)] = u;      // the order of the queue affects no value
      continue;
    }
    const int x = lane < n ? tmp[beg + lane] : 0x7fffffff;
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      const int yk = __shfl_sync(0xffffffffu, x, k);
      rank += (yk < x) || (yk == x && k < lane);
    }
    if (lane < n) t_indices[beg + rank] = x;
  }
}
__global__ void __launch_bounds__(256) sort_long_lists_kernel(const int32_t* __restrict__ t_indptr,
                                                              const int32_t* __restrict__ long_rows,
                                                              const int32_t* __restrict__ n_long,
                                                              const int32_t* __restrict__ tmp,
                                                              int32_t* __restrict__ t_indices) {
  __shared__ int a[SORT_SMEM];
  const int total = *n_long;
  for (int q = blockIdx.x; q < total; q += gridDim.x) {
    const int u = long_rows[q];
    const int beg = t_indptr[u], n = t_indptr[u + 1] - beg;
    if (n <= SORT_SMEM) {
      int np2 = 64;
      while (np2 < n) np2 <<= 1;
      for (int i = threadIdx.x; i < np2; i += blockDim.x) a[i] = i < n ? tmp[beg + i] : 0x7fffffff;
      __syncthreads();
      for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = threadIdx.x; i < np2; i += blockDim.x) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const int x = a[i], y = a[ixj];
              if ((x > y) == ((i & k) == 0)) { a[i] = y; a[ixj] = x; }
            }
          }
          __syncthreads();
        }
      for (int i = threadIdx.x; i < n; i += blockDim.x) t_indices[beg + i] = a[i];
      __syncthreads();
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {      // any length: quadratic, every thread ranks its elements
        const int x = tmp[beg + i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
          const int y = tmp[beg + j];
          rank += (y < x) || (y == x && j < i);
        }
        t_indices[beg + rank] = x;
      }
    }
  }
}

// ---- active rows (rows whose gradient is not structurally zero, packing.active_rows) ----
__global__ void flag_centres_kernel(const int32_t* __restrict__ centre_row, int n, uint8_t* __restrict__ flag) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) flag[centre_row[i]] = 1;
}
// in-neighbours of the flagged rows: a warp looks at 32 rows at a time and walks the lists of the flagged ones
__global__ void flag_in_neighbours_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int n_rows,
                                          const uint8_t* __restrict__ flag_hi, uint8_t* __restrict__ flag_lo) {
  const int lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int v0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; v0 < n_rows; v0 += nw * 32) {
    const int v = v0 + lane;
    const bool on = v < n_rows && flag_hi[v];
    int beg = 0, end = 0;
    if (on) { beg = indptr[v]; end = indptr[v + 1]; }
    unsigned todo = __ballot_sync(0xffffffffu, on);
    while (todo) {
      const int l = __ffs(todo) - 1;
      todo &= todo - 1;
      const int b = __shfl_sync(0xffffffffu, beg, l), e = __shfl_sync(0xffffffffu, end, l);
      for (int k = b + lane; k < e; k += 32) flag_lo[indices[k]] = 1;
    }
  }
}
// flagged rows in ascending order (pos = exclusive scan of the flags), the number of active rows in front of every
// task, and -- for the last layer -- the position of every centre among the active rows
__global__ void compact_rows_kernel(const uint8_t* __restrict__ flag, const int32_t* __restrict__ pos, int n_rows,
                                    const int32_t* __restrict__ task_row_ptr, int n_tasks, const int32_t* __restrict__ centre_row,
                                    int n_centres, int32_t* __restrict__ act_rows, int32_t* __restrict__ act_task_ptr,
                                    int32_t* __restrict__ centre_pos, int32_t* __restrict__ n_act_out) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int v = tid; v < n_rows; v += nt)
    if (flag[v]) act_rows[pos[v]] = v;
  for (int t = tid; t <= n_tasks; t += nt) act_task_ptr[t] = pos[task_row_ptr[t]];
  if (centre_pos)
    for (int i = tid; i < n_centres; i += nt) centre_pos[i] = pos[centre_row[i]];
  if (tid == 0) *n_act_out = pos[n_rows];
}

inline int64_t al256(int64_t x) { return (x + 255) / 256 * 256; }
struct Ws {
  int32_t* cnt;        // [N + 1] out-degrees, then the fill cursor
  int32_t* tmp;        // [E] by-source lists in cursor order
  int32_t* pos;        // [N + 1] scan of a flag array
  int* block_sum;      // [ceil(N / SCAN_BLOCK) + 2]
  uint8_t* flag;       // [n_layers][N]
  int32_t* long_rows;  // [E / 33 + 1] source rows with more than 32 out-edges
  int32_t* n_long;
  int64_t total;
};
Ws carve(void* base, int n_nodes, int n_edges, int n_layers) {
  Ws w;
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) { char* q = p ? p + off : nullptr; off += al256(bytes); return q; };
  w.cnt = reinterpret_cast<int32_t*>(take(((int64_t)n_nodes + 1) * 4));
  w.tmp = reinterpret_cast<int32_t*>(take((int64_t)(n_edges > 0 ? n_edges : 1) * 4));
  w.pos = reinterpret_cast<int32_t*>(take(((int64_t)n_nodes + 1) * 4));
  w.block_sum = reinterpret_cast<int*>(take(((int64_t)n_nodes / SCAN_BLOCK + 3) * 4));
  w.flag = reinterpret_cast<uint8_t*>(take((int64_t)(n_layers > 0 ? n_layers : 1) * (n_nodes > 0 ? n_nodes : 1)));
  w.long_rows = reinterpret_cast<int32_t*>(take(((int64_t)n_edges / 33 + 1) * 4));
  w.n_long = reinterpret_cast<int32_t*>(take(4));
  w.total = off;
  return w;
}
inline int grid_for(int64_t items, int per_block) {
  const int64_t g = (items + per_block - 1) / per_block;
  return (int)(g < 1 ? 1 : (g < 16 * kNumSMs ? g : 16 * kNumSMs));
}

}  // namespace
}  // namespace gmeta

using namespace gmeta;

extern "C" int64_t gmeta_packed_set_finish_workspace_bytes(int32_t n_nodes, int32_t n_edges, int32_t n_layers) {
  if (n_nodes < 0 || n_edges < 0 || n_layers < 0) return -1;
  return carve(nullptr, n_nodes, n_edges, n_layers).total;
}

extern "C" int gmeta_packed_set_finish(const int32_t* indptr, const int32_t* indices, int32_t n_nodes, int32_t n_edges,
                                       const int32_t* sub_node_ptr, const int32_t* task_sub_ptr, int32_t n_tasks,
                                       const int32_t* centre_row, int32_t n_centres, int32_t n_layers,
                                       int32_t* t_indptr, int32_t* t_indices, int32_t* task_row_ptr, int32_t* tile_row0,
                                       int32_t* tile_nrows, int32_t* tile_task, int32_t* const* act_rows,
                                       int32_t* const* act_task_ptr, int32_t* const* act_tile_row0,
                                       int32_t* const* act_tile_nrows, int32_t* const* act_tile_task,
                                       int32_t* centre_pos, int32_t* counts, void* workspace, int64_t workspace_bytes,
                                       void* stream) {
  if (n_nodes <= 0 || n_edges < 0 || n_tasks <= 0 || n_layers < 0 || n_layers > GMETA_MAX_LAYERS) return GMETA_ERR_BAD_ARG;
  if (!indptr || !sub_node_ptr || !centre_row || !t_indptr || !task_row_ptr || !tile_row0 ||
      !tile_nrows || !tile_task || !counts || (n_edges > 0 && (!indices || !t_indices)))
    return GMETA_ERR_BAD_ARG;
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return GMETA_ERR_WORKSPACE;
  const Ws w = carve(workspace, n_nodes, n_edges, n_layers);
  if (workspace_bytes < w.total) return GMETA_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  // task row pointers + row tiles: counts[0] = n_tiles, counts[1] = rows of the largest task
  tile_table_kernel<<<1, 1024, 0, s>>>(sub_node_ptr, task_sub_ptr, n_tasks, task_row_ptr, tile_row0, tile_nrows, tile_task,
                                       counts + 0, counts + 1);
  if ((rc = check_launch()) != GMETA_OK) return rc;
  // CSR by source
  if (cudaMemsetAsync(w.cnt, 0, ((size_t)n_nodes + 1) * 4, s) != cudaSuccess) return GMETA_ERR_LAUNCH;
  if (n_edges > 0) {
    out_degree_kernel<<<grid_for(n_edges, 256), 256, 0, s>>>(indices, n_edges, w.cnt);
    if ((rc = check_launch()) != GMETA_OK) return rc;
  }
  if ((rc = excl_scan<int32_t>(w.cnt, n_nodes, t_indptr, w.block_sum, s)) != GMETA_OK) return rc;
  if (n_edges > 0) {
    if (cudaMemcpyAsync(w.cnt, t_indptr, (size_t)n_nodes * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return GMETA_ERR_LAUNCH;
    transpose_fill_kernel<<<grid_for(n_nodes, 8), 256, 0, s>>>(indptr, indices, n_nodes, w.cnt, w.tmp);
    if ((rc = check_launch()) != GMETA_OK) return rc;
    if (cudaMemsetAsync(w.n_long, 0, 4, s) != cudaSuccess) return GMETA_ERR_LAUNCH;
    sort_short_lists_kernel<<<grid_for(n_nodes, 8), 256, 0, s>>>(t_indptr, n_nodes, w.tmp, t_indices, w.long_rows, w.n_long);
    if ((rc = check_launch()) != GMETA_OK) return rc;
    sort_long_lists_kernel<<<8 * kNumSMs, 256, 0, s>>>(t_indptr, w.long_rows, w.n_long, w.tmp, t_indices);
    if ((rc = check_launch()) != GMETA_OK) return rc;
  }
  // active rows per layer: centres at the last layer, the in-neighbours of the layer above below it
  if (n_layers > 0) {
    if (!act_rows || !act_task_ptr || !act_tile_row0 || !act_tile_nrows || !act_tile_task || !centre_pos) return GMETA_ERR_BAD_ARG;
    if (cudaMemsetAsync(w.flag, 0, (size_t)n_layers * n_nodes, s) != cudaSuccess) return GMETA_ERR_LAUNCH;
    uint8_t* top = w.flag + (size_t)(n_layers - 1) * n_nodes;
    flag_centres_kernel<<<grid_for(n_centres, 256), 256, 0, s>>>(centre_row, n_centres, top);
    if ((rc = check_launch()) != GMETA_OK) return rc;
    for (int l = n_layers - 1; l >= 0; --l) {
      uint8_t* fl = w.flag + (size_t)l * n_nodes;
      if (l < n_layers - 1) {
        flag_in_neighbours_kernel<<<grid_for(n_nodes, 256), 256, 0, s>>>(indptr, indices, n_nodes, fl + n_nodes, fl);
        if ((rc = check_launch()) != GMETA_OK) return rc;
      }
      if ((rc = excl_scan<uint8_t>(fl, n_nodes, w.pos, w.block_sum, s)) != GMETA_OK) return rc;
      compact_rows_kernel<<<grid_for(n_nodes, 256), 256, 0, s>>>(fl, w.pos, n_nodes, task_row_ptr, n_tasks, centre_row, n_centres,
                                                                 act_rows[l], act_task_ptr[l],
                                                                 l == n_layers - 1 ? centre_pos : nullptr, counts + 2 + l);
      if ((rc = check_launch()) != GMETA_OK) return rc;
      tile_table_kernel<<<1, 1024, 0, s>>>(act_task_ptr[l], nullptr, n_tasks, nullptr, act_tile_row0[l], act_tile_nrows[l],
                                           act_tile_task[l], counts + 2 + n_layers + l, nullptr);
      if ((rc = check_launch()) != GMETA_OK) return rc;
    }
  }
  return GMETA_OK;
}
