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

// ------------------------------------------------------------------------------------------------------------------
// The remaining per-batch host work of the packer as single calls (no Python between the steps, and -- being ctypes
// calls -- none of it holds the interpreter lock: with three packer threads beside the thread that launches the steps,
// the ~3 ms of numpy bookkeeping per batch were what the end-to-end step waited for).
// ------------------------------------------------------------------------------------------------------------------
namespace {
inline int64_t al4(int64_t n) { return (n + 3) / 4 * 4; }

// tiles of <= GMETA_TILE_ROWS rows that never straddle a task (learner.tile_table); returns the number of tiles
int64_t write_tile_table(const int64_t* ptr, int32_t n_tasks, int32_t* row0, int32_t* nrows, int32_t* task) {
  int64_t k = 0;
  for (int32_t t = 0; t < n_tasks; ++t)
    for (int64_t r = ptr[t]; r < ptr[t + 1]; r += GMETA_TILE_ROWS, ++k) {
      if (row0) {
        row0[k] = (int32_t)r;
        nrows[k] = (int32_t)std::min<int64_t>(GMETA_TILE_ROWS, ptr[t + 1] - r);
        task[k] = t;
      }
    }
  return k;
}
}  // namespace

// Centre rows (learner.py:161-170: centre index inside its subgraph + first packed row of the subgraph), labels, task
// pointers and (unless NULL) the row-tile table of one set.  bnn[t]: nodes per subgraph of task t (n_sub[t] entries),
// centres[t]: n_sub[t] * cps local centre indices, labels[t]: n_sub[t] labels.
extern "C" int gmeta_host_pack_small(int32_t n_tasks, int32_t cps, const int64_t* const* bnn, const int32_t* n_sub,
                                     const int64_t* const* centres, const int64_t* const* labels,
                                     const int64_t* node_off, const int64_t* sub_off, int32_t* out_centre_row,
                                     int32_t* out_labels, int32_t* out_task_row_ptr, int32_t* out_task_sub_ptr,
                                     int32_t* out_tile_row0, int32_t* out_tile_nrows, int32_t* out_tile_task) {
  if (n_tasks < 0 || cps < 1 || cps > 2 || !node_off || !sub_off || !out_centre_row || !out_labels || !out_task_row_ptr ||
      !out_task_sub_ptr || (n_tasks > 0 && (!bnn || !n_sub || !centres || !labels)))
    return GMETA_ERR_BAD_ARG;
  for (int32_t t = 0; t < n_tasks; ++t) {
    int64_t first = node_off[t];
    const int64_t s0 = sub_off[t];
    if (sub_off[t + 1] - s0 != n_sub[t]) return GMETA_ERR_BAD_ARG;
    for (int32_t k = 0; k < n_sub[t]; ++k) {
      for (int32_t c = 0; c < cps; ++c) out_centre_row[(s0 + k) * cps + c] = (int32_t)(centres[t][(int64_t)k * cps + c] + first);
      out_labels[s0 + k] = (int32_t)labels[t][k];
      first += bnn[t][k];
    }
    if (first != node_off[t + 1]) return GMETA_ERR_BAD_ARG;
  }
  for (int32_t t = 0; t <= n_tasks; ++t) {
    out_task_row_ptr[t] = (int32_t)node_off[t];
    out_task_sub_ptr[t] = (int32_t)sub_off[t];
  }
  if (out_tile_row0) write_tile_table(node_off, n_tasks, out_tile_row0, out_tile_nrows, out_tile_task);
  return GMETA_OK;
}

// Active rows of every layer of one set (centres at the last layer, the in-neighbours of the layer above below it,
// packing.active_rows), their task pointers and tile tables, and the position of every centre among the active rows of
// the last layer -- written behind each other at buf[off ...] in the order of packing._plan_fill_act: per layer
// act_rows | act_task_ptr | act_tile_row0 | act_tile_nrows | act_tile_task, every segment starting at a multiple of 4
// ints.  seg_off / seg_n [n_layers * 5] receive the offsets and lengths.  flags: >= n_nodes zeroed bytes (left zeroed),
// scratch: >= 2 * (n_nodes + 1) int64.  Returns the first free offset, or a negative error.
extern "C" int64_t gmeta_host_active_rows(const int32_t* indptr, const int32_t* indices, int64_t n_nodes,
                                          const int32_t* centre_row, int64_t n_centres, const int64_t* node_off,
                                          int32_t n_tasks, int32_t n_layers, uint8_t* flags, int64_t* scratch, int32_t* buf,
                                          int64_t off, int64_t* seg_off, int64_t* seg_n, int32_t* out_centre_pos) {
  if (n_layers < 1 || n_layers > GMETA_MAX_LAYERS || n_nodes < 0 || n_centres < 0 || n_tasks < 0 || !indptr || !centre_row ||
      !node_off || !flags || !scratch || !buf || !seg_off || !seg_n || !out_centre_pos)
    return GMETA_ERR_BAD_ARG;
  std::vector<std::vector<int64_t>> rows((size_t)n_layers);
  // last layer: the distinct centre rows, ascending
  {
    std::vector<int64_t>& r = rows[(size_t)n_layers - 1];
    r.reserve((size_t)n_centres);
    for (int64_t i = 0; i < n_centres; ++i) {
      const int32_t c = centre_row[i];
      if (c < 0 || c >= n_nodes) return GMETA_ERR_BAD_ARG;
      if (!flags[c]) { flags[c] = 1; r.push_back(c); }
    }
    std::sort(r.begin(), r.end());
    for (int64_t v : r) flags[v] = 0;
  }
  for (int32_t l = n_layers - 1; l > 0; --l) {
    const std::vector<int64_t>& up = rows[(size_t)l];
    const int64_t n = gmeta_host_active_in_neighbours(indptr, indices, up.data(), (int64_t)up.size(), n_nodes, flags, scratch);
    if (n < 0) return n;
    rows[(size_t)l - 1].assign(scratch, scratch + n);
  }
  int64_t* tptr = scratch;                       // [n_tasks + 1]
  for (int32_t l = 0; l < n_layers; ++l) {
    const std::vector<int64_t>& r = rows[(size_t)l];
    const int64_t n = (int64_t)r.size();
    for (int32_t t = 0; t <= n_tasks; ++t) tptr[t] = std::lower_bound(r.begin(), r.end(), node_off[t]) - r.begin();
    const int64_t nt = write_tile_table(tptr, n_tasks, nullptr, nullptr, nullptr);
    const int64_t lens[5] = {n, (int64_t)n_tasks + 1, nt, nt, nt};
    int64_t o[5];
    for (int k = 0; k < 5; ++k) {
      o[k] = off;
      seg_off[l * 5 + k] = off;
      seg_n[l * 5 + k] = lens[k];
      off += al4(lens[k]);
    }
    for (int64_t i = 0; i < n; ++i) buf[o[0] + i] = (int32_t)r[(size_t)i];
    for (int32_t t = 0; t <= n_tasks; ++t) buf[o[1] + t] = (int32_t)tptr[t];
    write_tile_table(tptr, n_tasks, buf + o[2], buf + o[3], buf + o[4]);
  }
  const std::vector<int64_t>& last = rows[(size_t)n_layers - 1];
  for (int64_t i = 0; i < n_centres; ++i)
    out_centre_pos[i] = (int32_t)(std::lower_bound(last.begin(), last.end(), (int64_t)centre_row[i]) - last.begin());
  return off;
}

// The equal-count requirements the reference enforces implicitly through torch.stack (meta.py:42,65-66) for every task
// of a meta-batch: every support class has >= k_spt members, the query classes are balanced and equal the support
// classes.  Returns the largest number of classes of a task (>= 1), or -1 / -2 / -3 for the three violations.
extern "C" int gmeta_host_validate_labels(int32_t n_tasks, const int64_t* const* y_spt, const int32_t* n_spt,
                                          const int64_t* const* y_qry, const int32_t* n_qry, int32_t k_spt) {
  int max_classes = 1;
  std::vector<std::pair<int64_t, int>> cs, cq;
  auto count = [](const int64_t* y, int32_t n, std::vector<std::pair<int64_t, int>>& out) {
    std::vector<int64_t> v(y, y + n);
    std::sort(v.begin(), v.end());
    out.clear();
    for (int32_t i = 0; i < n;) {
      int32_t j = i;
      while (j < n && v[(size_t)j] == v[(size_t)i]) ++j;
      out.emplace_back(v[(size_t)i], j - i);
      i = j;
    }
  };
  for (int32_t t = 0; t < n_tasks; ++t) {
    count(y_spt[t], n_spt[t], cs);
    count(y_qry[t], n_qry[t], cq);
    if (cs.empty() || cq.empty()) return -1;
    for (const auto& c : cs)
      if (c.second < k_spt) return -1;
    for (const auto& c : cq)
      if (c.second != cq[0].second) return -2;
    if (cs.size() != cq.size()) return -3;
    for (size_t i = 0; i < cs.size(); ++i)
      if (cs[i].first != cq[i].first) return -3;
    if ((int)cs.size() > max_classes) max_classes = (int)cs.size();
  }
  return max_classes;
}
