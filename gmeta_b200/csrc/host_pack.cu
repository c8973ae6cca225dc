// Host-side helper of the packing step (no device code): concatenates the per-task CSR arrays of a
// meta-batch into the packed-set layout (include/gmeta_b200.h) with node / edge offsets applied, on a
// few host threads.  Replaces the per-task numpy loops of gmeta_b200/packing.py:fill_set, which were the
// largest part of the end-to-end step once the device work had shrunk below them.  Pure integer work:
// the reference does the equivalent inside dgl.batch (subgraph_data_processing.py:399-406).
#include <algorithm>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {
// dst[i] = src[i] + add  (int32).  The destination (tens of MB of pinned staging memory, written once and then
// read by the DMA engine only) is written with non-temporal stores where SSE2 is available: no read-for-ownership
// of the destination lines and no cache pollution.
inline void add_copy(int32_t* __restrict__ dst, const int32_t* __restrict__ src, int64_t n, int32_t add) {
  int64_t i = 0;
#if defined(__SSE2__)
  for (; i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15u); ++i) dst[i] = src[i] + add;
  const __m128i va = _mm_set1_epi32(add);
  for (; i + 16 <= n; i += 16) {
    const __m128i a0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i a1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 4));
    const __m128i a2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 8));
    const __m128i a3 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 12));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_add_epi32(a0, va));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 4), _mm_add_epi32(a1, va));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 8), _mm_add_epi32(a2, va));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 12), _mm_add_epi32(a3, va));
  }
#endif
  for (; i < n; ++i) dst[i] = src[i] + add;
}
}  // namespace

extern "C" int gmeta_host_pack_csr(int32_t n_tasks, const int32_t* const* indptr, const int32_t* const* indices,
                                   const int32_t* const* t_indptr, const int32_t* const* t_indices,
                                   const int64_t* node_off, const int64_t* edge_off, int32_t* out_indptr,
                                   int32_t* out_indices, int32_t* out_t_indptr, int32_t* out_t_indices,
                                   int32_t n_threads) {
  // the by-source arrays are optional (all four of t_indptr / t_indices / out_t_indptr / out_t_indices NULL): the
  // caller derives them on the device instead (gmeta_packed_set_finish)
  const bool with_t = t_indptr || t_indices || out_t_indptr || out_t_indices;
  if (n_tasks < 0 || !node_off || !edge_off || !out_indptr || (with_t && !out_t_indptr)) return GMETA_ERR_BAD_ARG;
  if (n_tasks > 0 && (!indptr || !indices || (with_t && (!t_indptr || !t_indices)))) return GMETA_ERR_BAD_ARG;
  if (node_off[n_tasks] > 0x7fffffffLL || edge_off[n_tasks] > 0x7fffffffLL) return GMETA_ERR_UNSUPPORTED;
  out_indptr[0] = 0;
  if (with_t) out_t_indptr[0] = 0;
  std::atomic<int> next(0);
  const int n_units = 4 * n_tasks;      // (task, array) units, handed out dynamically
  auto work = [&]() {
    for (int u = next.fetch_add(1); u < n_units; u = next.fetch_add(1)) {
      const int t = u >> 2;
      const int64_t a = node_off[t], n = node_off[t + 1] - a, ea = edge_off[t], e = edge_off[t + 1] - ea;
      switch (u & 3) {
        case 0: add_copy(out_indptr + a + 1, indptr[t] + 1, n, (int32_t)ea); break;
        case 1: if (with_t) add_copy(out_t_indptr + a + 1, t_indptr[t] + 1, n, (int32_t)ea); break;
        case 2: if (e) add_copy(out_indices + ea, indices[t], e, (int32_t)a); break;
        default: if (with_t && e) add_copy(out_t_indices + ea, t_indices[t], e, (int32_t)a); break;
      }
    }
#if defined(__SSE2__)
    _mm_sfence();     // non-temporal stores are globally visible before the caller starts the H2D copy
#endif
  };
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt > 8) nt = 8;            // memory-bound: a few threads saturate the host memory system
  if (nt > n_units) nt = n_units;
  if (nt <= 1) {
    work();
    return GMETA_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(nt - 1);
  for (int i = 0; i < nt - 1; ++i) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return GMETA_OK;
}

// feat_row[node_off[t] + i] = ids[t][i] + goff, goff = the feature-table row of node 0 of the graph the node's
// subgraph came from: sub_goff[t][k] for the nodes [sub_ptr[t][k], sub_ptr[t][k+1]) (NULL sub_goff: one graph,
// offset 0).  What meta.py:119-120 does with numpy fancy indexing + vstack per task.
extern "C" int gmeta_host_pack_feat_rows(int32_t n_tasks, const int64_t* const* ids, const int64_t* const* sub_ptr,
                                         const int64_t* const* sub_goff, const int32_t* n_sub, const int64_t* node_off,
                                         int32_t* out_feat_row, int32_t n_threads) {
  if (n_tasks < 0 || !node_off || !out_feat_row || (n_tasks > 0 && !ids)) return GMETA_ERR_BAD_ARG;
  if (sub_goff && (!sub_ptr || !n_sub)) return GMETA_ERR_BAD_ARG;
  std::atomic<int> next(0);
  auto work = [&]() {
    for (int t = next.fetch_add(1); t < n_tasks; t = next.fetch_add(1)) {
      int32_t* dst = out_feat_row + node_off[t];
      const int64_t n = node_off[t + 1] - node_off[t];
      const int64_t* src = ids[t];
      if (!sub_goff) {
        for (int64_t i = 0; i < n; ++i) dst[i] = (int32_t)src[i];
      } else {
        for (int k = 0; k < n_sub[t]; ++k) {
          const int64_t g = sub_goff[t][k];
          for (int64_t i = sub_ptr[t][k]; i < sub_ptr[t][k + 1]; ++i) dst[i] = (int32_t)(src[i] + g);
        }
      }
    }
  };
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt > 16) nt = 16;
  if (nt > n_tasks) nt = n_tasks;
  if (nt <= 1) {
    work();
    return GMETA_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(nt - 1);
  for (int i = 0; i < nt - 1; ++i) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  return GMETA_OK;
}

// Active rows of one layer from those of the layer above: the sorted distinct in-neighbours of `rows`
// (packed CSR by destination), through a caller-owned zeroed byte map of n_nodes entries (left zeroed again).
// Returns the count (<= capacity of out_rows = n_nodes), or a negative error.
extern "C" int64_t gmeta_host_active_in_neighbours(const int32_t* indptr, const int32_t* indices, const int64_t* rows,
                                                   int64_t n_rows, int64_t n_nodes, uint8_t* flags, int64_t* out_rows) {
  if (n_rows < 0 || n_nodes < 0 || (n_rows > 0 && (!indptr || !indices || !rows)) || !flags || !out_rows)
    return GMETA_ERR_BAD_ARG;
  // collect the distinct neighbours in visiting order (the byte map de-duplicates); the lists of ascending
  // centre rows in disjoint subgraph ranges usually come out sorted already, otherwise sort
  int64_t n = 0, prev = -1;
  bool sorted = true;
  for (int64_t i = 0; i < n_rows; ++i) {
    const int64_t v = rows[i];
    for (int32_t e = indptr[v]; e < indptr[v + 1]; ++e) {
      const int32_t u = indices[e];
      if (flags[u]) continue;
      flags[u] = 1;
      out_rows[n++] = u;
      if (u < prev) sorted = false;
      prev = u;
    }
  }
  if (!sorted) std::sort(out_rows, out_rows + n);
  for (int64_t i = 0; i < n; ++i) flags[out_rows[i]] = 0;
  return n;
}
