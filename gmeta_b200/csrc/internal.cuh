// Declarations shared between the translation units of libgmeta_b200.so (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace gmeta {

int gcn_layer_fwd_simt(const GatherSrc& g, const int32_t* tile_row0, const int32_t* tile_nrows,
                       const int32_t* tile_task, int n_tiles, const float* W, int64_t w_task_stride,
                       int ldw, int trans_w, const float* bias, int64_t b_task_stride, int f_out,
                       int relu, const float* relu_mask, float* out, int ld_out, cudaStream_t stream);

int aggregate_rows_impl(const float* in, int32_t ld_in, const int32_t* in_row_map, const int32_t* dst_rows,
                        const int32_t* indptr, const int32_t* indices, const float* norm, int32_t n_rows,
                        int32_t f_in, int32_t scale_dst, float* out, int32_t ld_out, const int32_t* pos_indptr,
                        cudaStream_t stream);
int active_out_lists_build(const int32_t* rows, int n_rows, const int32_t* t_indptr, const int32_t* t_indices,
                           const int32_t* keep, int32_t* count, int32_t* ptr, int32_t* out_idx, cudaStream_t stream);
int gcn_layer_wgrad_impl(const float* in, int32_t ld_in, const int32_t* in_row_map, const int32_t* dst_rows,
                         const int32_t* indptr, const int32_t* indices, const float* norm, const int32_t* task_row_ptr,
                         int32_t n_tasks, const float* dZ, int32_t ld_dz, int32_t f_in, int32_t f_out, float* dW,
                         int64_t dw_task_stride, float* db, int64_t db_task_stride, void* workspace,
                         int64_t workspace_bytes, int64_t rows_hint, cudaStream_t s, int identity_graph = 0);

// ---- streamed-weight tensor-core layer kernel (gcn_layer_tc.cu) ----
bool gcn_layer_fwd_tc_supported(const GatherSrc& g, int ldw, int trans_w, int f_out, const float* out,
                                int ld_out);
int64_t gcn_layer_fwd_tc_workspace_bytes(int n_copies, int f_in, int f_out);
// `prepacked`: the weights' hi/lo TF32 operand image made by gcn_tc_sgd_pack for exactly these (W, orientation), copies
// `prepacked_stride` floats apart (0 = one shared copy); NULL = split W into the workspace on every call.
int gcn_layer_fwd_tc(const GatherSrc& g, const int32_t* tile_row0, const int32_t* tile_nrows,
                     const int32_t* tile_task, int n_tiles, int n_copies, const float* W, int64_t w_task_stride,
                     int ldw, int trans_w, const float* bias, int64_t b_task_stride, int f_out, int relu,
                     const float* relu_mask, float* out, int ld_out, void* workspace, int64_t workspace_bytes,
                     const float* prepacked, int64_t prepacked_stride, cudaStream_t stream);
// Per-CTA scratch the kernel needs besides the weight image (long rows), for callers that bring their own image.
int64_t gcn_layer_fwd_tc_scratch_bytes(int f_in);

// One weight matrix inside the flat parameter buffer and where its operand image goes.
struct TcPackSeg {
  int w_off;            // offset of the matrix in a parameter copy (floats)
  int K, N;             // contraction / output width of the layer launch that will use the image
  int ldw, trans;       // B[k][n] = W[k*ldw + n] (trans == 0) or W[n*ldw + k]
  long long img_off;    // offset of the image inside one copy's image block (floats); 2*K*N floats long
};
struct TcPackPlan {
  int n_seg;
  TcPackSeg seg[2 * GMETA_MAX_LAYERS];
  long long img_copy_stride;   // floats between the image blocks of consecutive copies
};
// w_out[c][p] = w_in[c * w_in_stride + p] - lr * grad[c][p]  (meta.py:126,151; grad == NULL: no update, w_out unused)
// AND the operand images of the UPDATED weights for every segment of `plan`, in ONE launch.
int gcn_tc_sgd_pack(const float* w_in, int64_t w_in_stride, const float* grad, float lr, int n_copies, int n_params,
                    float* w_out, const TcPackPlan& plan, float* image, cudaStream_t stream);

// ---- CTA-pair tensor-core layer kernel (gcn_layer_pair.cu) ----
bool gcn_layer_fwd_pair_supported(const GatherSrc& g, int f_out, const float* bias, int64_t b_task_stride,
                                  const float* relu_mask, const float* out, int ld_out, int n_tasks);
int64_t gcn_layer_fwd_pair_workspace_bytes(int n_copies, int n_tiles, int n_tasks, int n_rows, int n_edges, int f_in,
                                           int f_out);
int gcn_layer_fwd_pair(const GatherSrc& g, const int32_t* tile_row0, const int32_t* tile_nrows,
                       const int32_t* tile_task, int n_tiles, int n_tasks, int n_copies, int n_rows, int n_edges,
                       const float* in_rowmax, const void* plan, const float* W, int64_t w_task_stride, int ldw,
                       int trans_w, const float* bias, int64_t b_task_stride, int f_out, int relu,
                       const float* relu_mask, float* out, int ld_out, float* out_rowmax, void* workspace,
                       int64_t workspace_bytes, cudaStream_t stream);
int64_t layer_plan_bytes(int n_tiles, int n_tasks, int n_rows, int n_edges);
int layer_plan_build(const int32_t* indptr, const int32_t* indices, const float* norm, const int32_t* in_row_map,
                     const int32_t* dst_rows, const int32_t* tile_row0, const int32_t* tile_nrows,
                     const int32_t* tile_task, int n_tiles, int n_tasks, int n_rows, int n_edges, void* plan,
                     cudaStream_t stream);
int row_absmax(const float* x, int ld, int n_rows, int f, float* out, cudaStream_t stream);

}  // namespace gmeta
