// C-ABI entry points that dispatch between implementations, plus version / error strings.
#include "common.cuh"
#include "internal.cuh"

#include <cstdlib>

using namespace gmeta;

namespace gmeta {
thread_local const float* g_norm_dst = nullptr;
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("GMETA_B200_PDL");
    return e && e[0] && e[0] != '0';
  }();
  return on;
}
}  // namespace gmeta

extern "C" int gmeta_version(void) { return 100; }

extern "C" const char* gmeta_error_string(int code) {
  switch (code) {
    case GMETA_OK: return "ok";
    case GMETA_ERR_BAD_ARG: return "bad argument (null pointer, negative size or inconsistent dims)";
    case GMETA_ERR_ALIGN: return "pointer or leading dimension not aligned as required";
    case GMETA_ERR_UNSUPPORTED: return "shape or mode not supported by this implementation";
    case GMETA_ERR_LAUNCH: return "CUDA launch failed";
    case GMETA_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown error";
  }
}

extern "C" int gmeta_gcn_layer_fwd_ex(const float* in, int32_t ld_in, const int32_t* in_row_map,
                                      const int32_t* dst_rows, const int32_t* indptr, const int32_t* indices,
                                      const float* norm, const int32_t* tile_row0, const int32_t* tile_nrows,
                                      const int32_t* tile_task, int32_t n_tiles, int32_t n_tasks, const float* W,
                                      int64_t w_task_stride, int32_t ldw, int32_t trans_w, const float* bias,
                                      int64_t b_task_stride, int32_t f_in, int32_t f_out, int32_t relu,
                                      const float* relu_mask, float* out, int32_t ld_out, int32_t impl,
                                      void* workspace, int64_t workspace_bytes, int32_t n_rows, int32_t n_edges,
                                      const float* in_rowmax, float* out_rowmax, const void* plan, void* stream) {
  if (!in || !indptr || !norm || !tile_row0 || !tile_nrows || !tile_task || !W || !out) return GMETA_ERR_BAD_ARG;
  if (n_tiles < 0 || n_tasks <= 0 || f_in <= 0 || f_out <= 0 || ld_in < f_in || ld_out < f_out) return GMETA_ERR_BAD_ARG;
  if (ldw < (trans_w ? f_in : f_out)) return GMETA_ERR_BAD_ARG;
  if (out_rowmax && n_rows <= 0) return GMETA_ERR_BAD_ARG;
  if (n_tiles == 0) return GMETA_OK;
  GatherSrc g;
  g.in = in; g.in_row_map = in_row_map; g.dst_rows = dst_rows; g.indptr = indptr; g.indices = indices; g.norm = norm;
  g.ld_in = ld_in; g.f_in = f_in; g.norm_dst = g_norm_dst;
  cudaStream_t s = (cudaStream_t)stream;
  const int n_copies = w_task_stride == 0 ? 1 : n_tasks;
  const bool ws_aligned = workspace && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0;
  const bool pair_ok = in_rowmax && n_rows > 0 && n_edges >= 0 &&
                       gcn_layer_fwd_pair_supported(g, f_out, bias, b_task_stride, relu_mask, out, ld_out, n_tasks);
  if (impl == GMETA_IMPL_TCPAIR && !pair_ok) return GMETA_ERR_UNSUPPORTED;
  if (impl == GMETA_IMPL_TCPAIR ||
      (impl == GMETA_IMPL_AUTO && pair_ok && ws_aligned &&
       workspace_bytes >= gcn_layer_fwd_pair_workspace_bytes(n_copies, n_tiles, n_tasks, n_rows, n_edges, f_in, f_out)))
    return gcn_layer_fwd_pair(g, tile_row0, tile_nrows, tile_task, n_tiles, n_tasks, n_copies, n_rows, n_edges,
                              in_rowmax, plan, W, w_task_stride, ldw, trans_w, bias, b_task_stride, f_out, relu,
                              relu_mask, out, ld_out, out_rowmax, workspace, workspace_bytes, s);
  const bool tc_ok = gcn_layer_fwd_tc_supported(g, ldw, trans_w, f_out, out, ld_out);
  if (impl == GMETA_IMPL_TCGEN05 && !tc_ok) return GMETA_ERR_UNSUPPORTED;
  // AUTO falls back to the FFMA kernel when no workspace for the weight image was provided
  const bool ws_ok = workspace && workspace_bytes >= gcn_layer_fwd_tc_workspace_bytes(n_copies, f_in, f_out);
  int rc;
  if (impl == GMETA_IMPL_TCGEN05 || (impl == GMETA_IMPL_AUTO && tc_ok && ws_ok))
    rc = gcn_layer_fwd_tc(g, tile_row0, tile_nrows, tile_task, n_tiles, n_copies, W, w_task_stride, ldw,
                          trans_w, bias, b_task_stride, f_out, relu, relu_mask, out, ld_out, workspace,
                          workspace_bytes, nullptr, 0, s);
  else if (impl != GMETA_IMPL_AUTO && impl != GMETA_IMPL_SIMT)
    return GMETA_ERR_BAD_ARG;
  else
    rc = gcn_layer_fwd_simt(g, tile_row0, tile_nrows, tile_task, n_tiles, W, w_task_stride, ldw, trans_w, bias,
                            b_task_stride, f_out, relu, relu_mask, out, ld_out, s);
  if (rc == GMETA_OK && out_rowmax) rc = row_absmax(out, ld_out, n_rows, f_out, out_rowmax, s);
  return rc;
}

extern "C" int gmeta_gcn_layer_fwd(const float* in, int32_t ld_in, const int32_t* in_row_map,
                                   const int32_t* dst_rows, const int32_t* indptr, const int32_t* indices, const float* norm,
                                   const int32_t* tile_row0, const int32_t* tile_nrows,
                                   const int32_t* tile_task, int32_t n_tiles, int32_t n_tasks, const float* W,
                                   int64_t w_task_stride, int32_t ldw, int32_t trans_w, const float* bias,
                                   int64_t b_task_stride, int32_t f_in, int32_t f_out, int32_t relu,
                                   const float* relu_mask, float* out, int32_t ld_out, int32_t impl,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
  if (impl == GMETA_IMPL_TCPAIR) return GMETA_ERR_UNSUPPORTED;   // needs the _ex arguments
  return gmeta_gcn_layer_fwd_ex(in, ld_in, in_row_map, dst_rows, indptr, indices, norm, tile_row0, tile_nrows,
                                tile_task, n_tiles, n_tasks, W, w_task_stride, ldw, trans_w, bias, b_task_stride, f_in,
                                f_out, relu, relu_mask, out, ld_out, impl, workspace, workspace_bytes, 0, 0, nullptr,
                                nullptr, nullptr, stream);
}

// Same layer with separate scales for a row as a source (`norm`) and as a destination (`norm_dst`):
//   out[i,:] = act( norm_dst[v] * (sum_u norm[u] * in[map(u),:]) . B + bias )
// norm_dst == NULL is gmeta_gcn_layer_fwd_ex.  Mean aggregation: norm = 1, norm_dst = 1 / max(in_deg, 1); plain sum: both 1
// (gmeta_aggregation_norms fills them).  A plan built with gmeta_layer_plan_build holds the SOURCE scales.
extern "C" int gmeta_gcn_layer_fwd_nd(const float* in, int32_t ld_in, const int32_t* in_row_map,
                                      const int32_t* dst_rows, const int32_t* indptr, const int32_t* indices,
                                      const float* norm, const float* norm_dst, const int32_t* tile_row0,
                                      const int32_t* tile_nrows, const int32_t* tile_task, int32_t n_tiles,
                                      int32_t n_tasks, const float* W, int64_t w_task_stride, int32_t ldw,
                                      int32_t trans_w, const float* bias, int64_t b_task_stride, int32_t f_in,
                                      int32_t f_out, int32_t relu, const float* relu_mask, float* out, int32_t ld_out,
                                      int32_t impl, void* workspace, int64_t workspace_bytes, int32_t n_rows,
                                      int32_t n_edges, const float* in_rowmax, float* out_rowmax, const void* plan,
                                      void* stream) {
  NormDstScope scope(norm_dst);
  return gmeta_gcn_layer_fwd_ex(in, ld_in, in_row_map, dst_rows, indptr, indices, norm, tile_row0, tile_nrows, tile_task,
                                n_tiles, n_tasks, W, w_task_stride, ldw, trans_w, bias, b_task_stride, f_in, f_out, relu,
                                relu_mask, out, ld_out, impl, workspace, workspace_bytes, n_rows, n_edges, in_rowmax,
                                out_rowmax, plan, stream);
}

extern "C" int64_t gmeta_gcn_layer_fwd_workspace_bytes(int32_t n_tasks, int64_t w_task_stride, int32_t f_in,
                                                       int32_t f_out, int32_t impl) {
  if (n_tasks <= 0 || f_in <= 0 || f_out <= 0) return 0;
  if (impl == GMETA_IMPL_SIMT) return 0;
  return gcn_layer_fwd_tc_workspace_bytes(w_task_stride == 0 ? 1 : n_tasks, f_in, f_out);
}

extern "C" int64_t gmeta_gcn_layer_fwd_ex_workspace_bytes(int32_t n_tasks, int64_t w_task_stride, int32_t n_tiles,
                                                          int32_t n_rows, int32_t n_edges, int32_t f_in,
                                                          int32_t f_out, int32_t impl) {
  if (n_tasks <= 0 || f_in <= 0 || f_out <= 0) return 0;
  if (impl == GMETA_IMPL_SIMT) return 0;
  const int n_copies = w_task_stride == 0 ? 1 : n_tasks;
  int64_t b = impl == GMETA_IMPL_TCPAIR ? 0 : gcn_layer_fwd_tc_workspace_bytes(n_copies, f_in, f_out);
  if (impl != GMETA_IMPL_TCGEN05 && f_in % 64 == 0 && f_out % 16 == 0 && n_rows > 0) {
    const int64_t p = gcn_layer_fwd_pair_workspace_bytes(n_copies, n_tiles, n_tasks, n_rows, n_edges, f_in, f_out);
    if (p > b) b = p;
  }
  return b;
}

extern "C" int gmeta_row_absmax(const float* x, int32_t ld, int32_t n_rows, int32_t f, float* out, void* stream) {
  if (!x || !out || n_rows < 0 || f <= 0 || ld < f) return GMETA_ERR_BAD_ARG;
  return row_absmax(x, ld, n_rows, f, out, (cudaStream_t)stream);
}

extern "C" int64_t gmeta_layer_plan_bytes(int32_t n_tiles, int32_t n_tasks, int32_t n_rows, int32_t n_edges) {
  if (n_tiles < 0 || n_tasks <= 0 || n_rows < 0 || n_edges < 0) return -1;
  return layer_plan_bytes(n_tiles, n_tasks, n_rows, n_edges);
}

extern "C" int gmeta_layer_plan_build(const int32_t* indptr, const int32_t* indices, const float* norm,
                                      const int32_t* in_row_map, const int32_t* dst_rows, const int32_t* tile_row0,
                                      const int32_t* tile_nrows, const int32_t* tile_task, int32_t n_tiles,
                                      int32_t n_tasks, int32_t n_rows, int32_t n_edges, void* plan, void* stream) {
  if (!indptr || !norm || !tile_row0 || !tile_nrows || !tile_task || !plan) return GMETA_ERR_BAD_ARG;
  if (n_tiles < 0 || n_tasks <= 0 || n_rows < 0 || n_edges < 0) return GMETA_ERR_BAD_ARG;
  return layer_plan_build(indptr, indices, norm, in_row_map, dst_rows, tile_row0, tile_nrows, tile_task, n_tiles,
                          n_tasks, n_rows, n_edges, plan, (cudaStream_t)stream);
}
