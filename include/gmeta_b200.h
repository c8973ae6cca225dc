/*
 * gmeta_b200 -- C ABI of the B200-native G-Meta inner-loop hot path.
 *
 * The reference (mims-harvard/G-Meta) has no FFI layer: its device arithmetic is
 * reached through the Python surface of G-Meta/learner.py and G-Meta/meta.py.  Each
 * entry point below replaces one of those call sites (cited per function) and is what
 * a maintainer would bind from that file (ctypes stubs in INTEGRATION.md).
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless named host_*;
 *   - fp32 row-major data with an explicit leading dimension (`ld*`, in floats);
 *     int32 indices; rows of activations/features are 16-byte aligned when ld % 4 == 0
 *     (the vectorised paths are selected at run time from ld and pointer alignment);
 *   - returns GMETA_OK or a negative GMETA_ERR_* code; never throws, never allocates or
 *     frees device memory (workspaces are passed in, sized by the *_workspace_bytes twin),
 *     never synchronises the stream; launch errors are picked up with cudaGetLastError();
 *   - `stream` is a cudaStream_t passed as void*;
 *   - stateless, hence thread-safe per stream.
 *
 * Packed meta-batch layout ("packed set"): the support (or query) subgraphs of ALL tasks
 * of a meta-batch are concatenated into one node range [0, N).  Rows of one task are
 * contiguous (task_row_ptr), a row tile (<= GMETA_TILE_ROWS rows) never straddles two
 * tasks, and every kernel picks the task's weight copy by tile_task[tile].
 */
#ifndef GMETA_B200_H_
#define GMETA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMETA_OK 0
#define GMETA_ERR_BAD_ARG (-1)      /* null pointer / negative size / inconsistent dims */
#define GMETA_ERR_ALIGN (-2)        /* pointer or leading dimension not aligned as required */
#define GMETA_ERR_UNSUPPORTED (-3)  /* shape outside what the kernels implement */
#define GMETA_ERR_LAUNCH (-4)       /* cudaGetLastError() != cudaSuccess after a launch */
#define GMETA_ERR_WORKSPACE (-5)    /* workspace too small */

#define GMETA_TILE_ROWS 128
#define GMETA_MAX_LAYERS 3          /* h in {1,2,3}: subgraph_data_processing.py:300-311 */

/* neighbourhood aggregation of a GCN layer: out_v = r_v * sum_u s_u x_u.  GCN: s = r = clamp(in_deg,1)^-1/2 -- the
 * reference's GraphConv (learner.py:29-49) and the only mode it has; MEAN: s = 1, r = 1/clamp(in_deg,1) (GraphSAGE-style
 * mean, named by north_star, no reference counterpart); SUM: s = r = 1. */
#define GMETA_AGG_GCN 0
#define GMETA_AGG_MEAN 1
#define GMETA_AGG_SUM 2

#define GMETA_IMPL_AUTO 0
#define GMETA_IMPL_SIMT 1           /* fp32 FFMA, any shape */
#define GMETA_IMPL_TCGEN05 2        /* tcgen05.mma kind::tf32, 3xTF32 error-compensated, weights streamed per tile */
#define GMETA_IMPL_TCPAIR 3         /* tcgen05.mma.cta_group::2 kind::f16, scaled FP16 hi/lo split, weights resident per CTA pair */

int gmeta_version(void);
const char* gmeta_error_string(int code);

/* ---------------------------------------------------------------------------------------
 * Packed set: adjacency + tiling of all tasks' subgraphs (device pointers).
 * ------------------------------------------------------------------------------------- */
typedef struct gmeta_packed_set {
  int32_t n_nodes;               /* N */
  int32_t n_edges;               /* E (directed, multi-edges counted) */
  int32_t n_tiles;
  int32_t n_tasks;               /* T */
  int32_t n_subgraphs;           /* S (all tasks) */
  int32_t centres_per_subgraph;  /* 1, or 2 in link-prediction mode (learner.py:165-168) */
  const int32_t* indptr;         /* [N+1] CSR by destination: in-neighbours of v */
  const int32_t* indices;        /* [E]   packed row ids */
  const int32_t* t_indptr;       /* [N+1] CSR by source: out-neighbours of u (backward) */
  const int32_t* t_indices;      /* [E] */
  const int32_t* tile_row0;      /* [n_tiles] first row of the tile */
  const int32_t* tile_nrows;     /* [n_tiles] 1..GMETA_TILE_ROWS */
  const int32_t* tile_task;      /* [n_tiles] */
  const int32_t* task_row_ptr;   /* [T+1] */
  const int32_t* task_sub_ptr;   /* [T+1] */
  const int32_t* centre_row;     /* [S * centres_per_subgraph] packed row of the centre node(s) */
  const int32_t* feat_row;       /* [N] row of the device feature table feeding layer 1, or NULL */
  const int32_t* labels;         /* [S] raw class labels */
  float* norm;                   /* [N] scale of a row as a source; GCN: clamp(in_deg,1)^-1/2 (filled by the step) */
  float* norm_dst;               /* [N] scale of a row as a destination, or NULL = norm (required when model.aggregation != GCN) */
  int32_t* class_pos;            /* [S] rank of the label among the task's sorted unique labels */
  int32_t* class_occ;            /* [S] how many earlier subgraphs of the task share the label */
  int32_t* n_classes;            /* [T] */
  /* Active rows of the backward pass, per GCN layer l (0-based): the rows where dL/dZ_l can be
   * non-zero.  Only the centre rows are read out (learner.py:166-170), so act[L-1] = the centre
   * rows and act[l-1] = the in-neighbours of act[l]; every other row of dZ_l is structurally 0
   * and the backward kernels skip it (exact).  Sorted by row, tasks contiguous. */
  int32_t n_act[GMETA_MAX_LAYERS];
  int32_t n_act_tiles[GMETA_MAX_LAYERS];
  const int32_t* act_rows[GMETA_MAX_LAYERS];       /* [n_act]  packed row ids */
  const int32_t* act_task_ptr[GMETA_MAX_LAYERS];   /* [T+1]    into act_rows */
  const int32_t* act_tile_row0[GMETA_MAX_LAYERS];  /* tile table over positions of act_rows */
  const int32_t* act_tile_nrows[GMETA_MAX_LAYERS];
  const int32_t* act_tile_task[GMETA_MAX_LAYERS];
  int32_t* row_pos[GMETA_MAX_LAYERS];              /* [N] scratch: row -> position in act_rows or -1 */
  const int32_t* centre_pos;     /* [S * centres_per_subgraph] position of each centre row in act_rows[L-1] */
} gmeta_packed_set_t;

/* Model topology = the reference's `config` list (train.py:67-75, learner.py:81-97) flattened.
 * Parameters live in ONE flat fp32 buffer in the reference's creation order
 * [W1 (in,out) | b1 | W2 | b2 | ... | Wlin (n_out, hid*(2 if link_pred)) | blin], each tensor
 * starting at a multiple of 4 floats (offsets below); per-task copies are strided by
 * n_params_padded. */
typedef struct gmeta_model {
  int32_t n_layers;
  int32_t f_in[GMETA_MAX_LAYERS];
  int32_t f_out[GMETA_MAX_LAYERS];
  int32_t n_out;                 /* logits width (labels_num, train.py:58-61) */
  int32_t link_pred;
  int32_t n_params_padded;       /* P */
  int32_t w_off[GMETA_MAX_LAYERS];
  int32_t b_off[GMETA_MAX_LAYERS];
  int32_t wlin_off;
  int32_t blin_off;
  int32_t aggregation;           /* GMETA_AGG_* (0 = the reference's symmetric normalisation) */
} gmeta_model_t;

/* norm[v] = 1/sqrt(max(indptr[v+1]-indptr[v], 1)), IEEE-rounded.
 * Replaces learner.py:29 `torch.pow(graph.in_degrees().float().clamp(min=1), -0.5)`. */
int gmeta_degree_norm(const int32_t* indptr, int32_t n_nodes, float* norm, void* stream);
/* Source / destination scales of an aggregation mode (GMETA_AGG_*) from the in-degrees; mode GCN writes the
 * gmeta_degree_norm values into both arrays. */
int gmeta_aggregation_norms(const int32_t* indptr, int32_t n_nodes, int32_t mode, float* norm_src, float* norm_dst,
                            void* stream);

/* One fused GCN layer over a packed set (replaces GraphConv.forward, learner.py:25-56, and the
 * features gather of meta.py:119-120 when in_row_map != NULL):
 *   M[v,:]   = sum_{u in indices[indptr[v]:indptr[v+1]]} norm[u] * in[map(u), :f_in]
 *   out[i,j] = act( norm[v] * sum_k M[v,k] * B[k,j] + bias[j] ),   j < f_out
 * where i runs over the tile rows and v = dst_rows ? dst_rows[i] : i (dst_rows selects a subset
 * of rows and makes `out` compact); map(u) = in_row_map ? in_row_map[u] : u, and a negative
 * in_row_map entry drops that neighbour;
 * with B[k,j] = W[k*ldw + j] (trans_w == 0) or W[j*ldw + k] (trans_w != 0), W/bias taken from
 * task t = tile_task[tile] at W + t*w_task_stride / bias + t*b_task_stride (stride 0 = shared),
 * act = ReLU iff (relu & 1); if relu_mask != NULL the result is zeroed where relu_mask[m,j] <= 0
 * (ld_out layout), m = the real row v, or the output row i when (relu & 2) (a compact mask stored in
 * the order of dst_rows).  Columns [f_out, round_up(f_out,4)) of out are written as 0.
 * The same entry point run on (t_indptr, t_indices) with trans_w=1, bias=NULL, relu=0 and
 * relu_mask = the lower layer's activations is the layer's data-gradient.
 * impl: GMETA_IMPL_SIMT (fp32 FFMA, any shape), GMETA_IMPL_TCGEN05 (tcgen05.mma 3xTF32 with TMEM
 * accumulators; needs f_in % 32 == 0, f_out % 16 == 0, f_out <= 256, ld % 4 == 0 and the
 * workspace below) or GMETA_IMPL_AUTO. */
int gmeta_gcn_layer_fwd(const float* in, int32_t ld_in, const int32_t* in_row_map,
                        const int32_t* dst_rows, const int32_t* indptr, const int32_t* indices, const float* norm,
                        const int32_t* tile_row0, const int32_t* tile_nrows,
                        const int32_t* tile_task, int32_t n_tiles, int32_t n_tasks,
                        const float* W, int64_t w_task_stride, int32_t ldw, int32_t trans_w,
                        const float* bias, int64_t b_task_stride,
                        int32_t f_in, int32_t f_out, int32_t relu, const float* relu_mask,
                        float* out, int32_t ld_out, int32_t impl,
                        void* workspace, int64_t workspace_bytes, void* stream);
/* Scratch for the tensor-core path's pre-split (hi/lo TF32), pre-swizzled weight image:
 * 2 * f_in * f_out floats per weight copy (n_tasks copies, or 1 when w_task_stride == 0).
 * 0 for GMETA_IMPL_SIMT.  With GMETA_IMPL_AUTO and a NULL/too small workspace the FFMA kernel runs. */
int64_t gmeta_gcn_layer_fwd_workspace_bytes(int32_t n_tasks, int64_t w_task_stride, int32_t f_in,
                                            int32_t f_out, int32_t impl);

/* Extended form of gmeta_gcn_layer_fwd.  Additional arguments:
 *   n_rows     rows covered by the tile table (N, or the length of dst_rows);
 *   n_edges    >= number of CSR entries of those rows (E of the packed set is always enough);
 *   in_rowmax  [rows of `in`] max_k |in[r,k]| of every input row, or NULL;
 *   out_rowmax [n_rows] receives max_j |out[i,j]| (the in_rowmax of the next layer), or NULL.
 * With in_rowmax given, GMETA_IMPL_AUTO selects GMETA_IMPL_TCPAIR when the shape allows
 * (f_in % 64 == 0, f_out % 16 == 0, f_out <= 256, 2*f_in*f_out + 64 KB <= 227 KB of shared memory,
 * n_tasks <= 2048): the row abs-max vector gives the rigorous per-row bound from which the
 * FP16 operand scaling is derived (csrc/gcn_layer_pair.cu).  The tile table must list the tiles
 * of a task contiguously.  Workspace: gmeta_gcn_layer_fwd_ex_workspace_bytes, 256-byte aligned.
 *   plan       a layer plan built by gmeta_layer_plan_build for exactly these (indptr, indices,
 *              norm, in_row_map, dst_rows, tile table) arguments, or NULL (then it is rebuilt
 *              inside the workspace on every call).  Only GMETA_IMPL_TCPAIR uses it. */
int gmeta_gcn_layer_fwd_ex(const float* in, int32_t ld_in, const int32_t* in_row_map,
                           const int32_t* dst_rows, const int32_t* indptr, const int32_t* indices, const float* norm,
                           const int32_t* tile_row0, const int32_t* tile_nrows,
                           const int32_t* tile_task, int32_t n_tiles, int32_t n_tasks,
                           const float* W, int64_t w_task_stride, int32_t ldw, int32_t trans_w,
                           const float* bias, int64_t b_task_stride,
                           int32_t f_in, int32_t f_out, int32_t relu, const float* relu_mask,
                           float* out, int32_t ld_out, int32_t impl,
                           void* workspace, int64_t workspace_bytes,
                           int32_t n_rows, int32_t n_edges, const float* in_rowmax, float* out_rowmax,
                           const void* plan, void* stream);
/* gmeta_gcn_layer_fwd_ex with separate scales for a row as a source (`norm`) and as a destination (`norm_dst`):
 *   out[i,:] = act( norm_dst[v] * (sum_u norm[u] * in[map(u),:]) . B + bias );  norm_dst == NULL: norm (the _ex call).
 * Mean / sum aggregation (GMETA_AGG_*; arrays from gmeta_aggregation_norms).  The data gradient of such a layer is the
 * same call on the transposed graph with the two arrays SWAPPED; a plan holds the source scales. */
int gmeta_gcn_layer_fwd_nd(const float* in, int32_t ld_in, const int32_t* in_row_map,
                           const int32_t* dst_rows, const int32_t* indptr, const int32_t* indices, const float* norm,
                           const float* norm_dst, const int32_t* tile_row0, const int32_t* tile_nrows,
                           const int32_t* tile_task, int32_t n_tiles, int32_t n_tasks,
                           const float* W, int64_t w_task_stride, int32_t ldw, int32_t trans_w,
                           const float* bias, int64_t b_task_stride,
                           int32_t f_in, int32_t f_out, int32_t relu, const float* relu_mask,
                           float* out, int32_t ld_out, int32_t impl,
                           void* workspace, int64_t workspace_bytes,
                           int32_t n_rows, int32_t n_edges, const float* in_rowmax, float* out_rowmax,
                           const void* plan, void* stream);
int64_t gmeta_gcn_layer_fwd_ex_workspace_bytes(int32_t n_tasks, int64_t w_task_stride, int32_t n_tiles,
                                               int32_t n_rows, int32_t n_edges, int32_t f_in,
                                               int32_t f_out, int32_t impl);

/* Layer plan: everything the CTA-pair layer kernel derives from the graph STRUCTURE and the
 * tiling alone (per-row source rows / norms of the first two in-neighbours, hub rows and their
 * flattened edge lists, pairs of same-task tiles).  Build once per packed set and operand mapping,
 * reuse for every layer call of the meta-step.  `plan` is caller-owned, 256-byte aligned. */
int64_t gmeta_layer_plan_bytes(int32_t n_tiles, int32_t n_tasks, int32_t n_rows, int32_t n_edges);
int gmeta_layer_plan_build(const int32_t* indptr, const int32_t* indices, const float* norm,
                           const int32_t* in_row_map, const int32_t* dst_rows,
                           const int32_t* tile_row0, const int32_t* tile_nrows, const int32_t* tile_task,
                           int32_t n_tiles, int32_t n_tasks, int32_t n_rows, int32_t n_edges,
                           void* plan, void* stream);

/* out[r] = max_k |x[r*ld + k]|, k < f  (prepares in_rowmax for gmeta_gcn_layer_fwd_ex). */
int gmeta_row_absmax(const float* x, int32_t ld, int32_t n_rows, int32_t f, float* out, void* stream);

/* Debug hook: device buffer [148][16] of int64 cycle counters that subsequent tensor-core layer
 * launches fill per role (producer prologue/wait/body, MMA waits, epilogue); NULL = off. */
void gmeta_debug_set_tc_profile(long long* device_buffer);
/* Debug ablation flags for performance triage (outputs are wrong while non-zero): 1 skip output
 * stores, 2 skip gather loads, 4 issue 1/4 of the MMAs, 8 skip the weight-chunk copies. */
void gmeta_debug_set_tc_flags(int flags);
/* Same for the CTA-pair kernel: 1 skip output stores, 2 skip gather loads, 4 issue 1/4 of the MMAs. */
void gmeta_debug_set_pair_flags(int flags);
void gmeta_debug_set_pair_profile(long long* device_buffer);

/* Weight/bias gradient of one GCN layer (the autograd.grad of meta.py:125,149 for that layer):
 *   dW[t][k,j] = sum_{v in task t} norm[v] * M[v,k] * dZ[v,j],   db[t][j] = sum_v dZ[v,j]
 * M as in gmeta_gcn_layer_fwd (re-gathered, not stored).  dZ is the gradient w.r.t. the
 * pre-activation (already ReLU-masked).  dW / db are written (not accumulated) at
 * dW + t*dw_task_stride (row-major [f_in, f_out], ld = f_out) and db + t*db_task_stride.
 * Deterministic: row-range partials in `workspace` are summed in a fixed order.
 * With dst_rows != NULL the sums run over the listed rows only (task_row_ptr then indexes the
 * list and dZ is compact: row i of dZ belongs to row dst_rows[i]). */
int64_t gmeta_gcn_layer_wgrad_workspace_bytes(int32_t n_tasks, int32_t f_in, int32_t f_out);
int gmeta_gcn_layer_wgrad(const float* in, int32_t ld_in, const int32_t* in_row_map,
                          const int32_t* dst_rows, const int32_t* indptr, const int32_t* indices, const float* norm,
                          const int32_t* task_row_ptr, int32_t n_tasks,
                          const float* dZ, int32_t ld_dz, int32_t f_in, int32_t f_out,
                          float* dW, int64_t dw_task_stride, float* db, int64_t db_task_stride,
                          void* workspace, int64_t workspace_bytes, void* stream);
/* same with a separate destination scale: dW[t] = sum_v (norm_dst[v] M_v)^T dZ_v  (norm_dst == NULL: norm) */
int gmeta_gcn_layer_wgrad_nd(const float* in, int32_t ld_in, const int32_t* in_row_map,
                             const int32_t* dst_rows, const int32_t* indptr, const int32_t* indices, const float* norm,
                             const float* norm_dst, const int32_t* task_row_ptr, int32_t n_tasks,
                             const float* dZ, int32_t ld_dz, int32_t f_in, int32_t f_out,
                             float* dW, int64_t dw_task_stride, float* db, int64_t db_task_stride,
                             void* workspace, int64_t workspace_bytes, void* stream);

/* HOST helper (no device work, no stream): packs the per-task CSR arrays of one set of a meta-batch into the
 * packed-set layout -- out_indptr[node_off[t] + 1 + i] = indptr[t][1 + i] + edge_off[t],
 * out_indices[edge_off[t] + e] = indices[t][e] + node_off[t], same for the transposed arrays -- on up to
 * n_threads host threads (0 = hardware concurrency, capped at 8).  What dgl.batch does for the reference
 * (subgraph_data_processing.py:399-406).  Output pointers are host memory (the pinned staging buffer).  The four
 * by-source arguments (t_indptr, t_indices, out_t_indptr, out_t_indices) may all be NULL: the caller then derives
 * the CSR by source on the device (gmeta_packed_set_finish). */
int gmeta_host_pack_csr(int32_t n_tasks, const int32_t* const* indptr, const int32_t* const* indices,
                        const int32_t* const* t_indptr, const int32_t* const* t_indices,
                        const int64_t* node_off, const int64_t* edge_off, int32_t* out_indptr,
                        int32_t* out_indices, int32_t* out_t_indptr, int32_t* out_t_indices, int32_t n_threads);

/* HOST helpers of the same packing step.  gmeta_host_pack_feat_rows: feat_row[node_off[t] + i] = ids[t][i] + the
 * feature-table offset of the node's graph (sub_goff[t][k] for the nodes [sub_ptr[t][k], sub_ptr[t][k+1]) of task t;
 * NULL sub_goff = one graph) -- the per-task feature gather of meta.py:119-120 as an index.
 * gmeta_host_active_in_neighbours: sorted distinct in-neighbours of `rows` in the packed CSR (the rows of layer l-1
 * the rows of layer l depend on); `flags` is a zeroed n_nodes-byte scratch map, returned zeroed; returns the count. */
int gmeta_host_pack_feat_rows(int32_t n_tasks, const int64_t* const* ids, const int64_t* const* sub_ptr,
                              const int64_t* const* sub_goff, const int32_t* n_sub, const int64_t* node_off,
                              int32_t* out_feat_row, int32_t n_threads);
int64_t gmeta_host_active_in_neighbours(const int32_t* indptr, const int32_t* indices, const int64_t* rows,
                                        int64_t n_rows, int64_t n_nodes, uint8_t* flags, int64_t* out_rows);

/* Normalised neighbourhood sum alone -- the aggregation half of GraphConv.forward (learner.py:29-32,41-45):
 *   out[i, :f_in] = (scale_dst ? norm[v] : 1) * sum_{(u -> v)} norm[u] * in[map(u), :],  v = dst_rows ? dst_rows[i] : i
 * (columns f_in..ld_out-1 are zeroed).  For the FIRST layer this depends on the graph and the features only,
 * not on the weights, so the ProtoMAML driver computes it once per meta-step and runs every first-layer
 * forward / weight gradient of the inner loop on the cached rows ("cached A.X"). */
int gmeta_aggregate_rows(const float* in, int32_t ld_in, const int32_t* in_row_map, const int32_t* dst_rows,
                         const int32_t* indptr, const int32_t* indices, const float* norm, int32_t n_rows,
                         int32_t f_in, int32_t scale_dst, float* out, int32_t ld_out, void* stream);
/* same with a separate destination scale (used when scale_dst != 0; norm_dst == NULL: norm) */
int gmeta_aggregate_rows_nd(const float* in, int32_t ld_in, const int32_t* in_row_map, const int32_t* dst_rows,
                            const int32_t* indptr, const int32_t* indices, const float* norm, const float* norm_dst,
                            int32_t n_rows, int32_t f_in, int32_t scale_dst, float* out, int32_t ld_out, void* stream);

/* HOST helpers of the packer (no device work): the small segments of one set -- centre rows (learner.py:161-170), labels,
 * task pointers, row tiles (out_tile_* may be NULL) -- from the per-task arrays; the active rows of every layer with their
 * task pointers / tile tables / centre positions written behind each other at buf[off...] (returns the next free
 * offset; seg_off / seg_n [n_layers * 5]: act_rows, act_task_ptr, act_tile_row0, act_tile_nrows, act_tile_task per
 * layer); and the label requirements of meta.py:42,65-66 for all tasks (returns the largest class count, or -1: a
 * support class with fewer than k_spt members, -2: unbalanced query classes, -3: support / query classes differ). */
int gmeta_host_pack_small(int32_t n_tasks, int32_t cps, const int64_t* const* bnn, const int32_t* n_sub,
                          const int64_t* const* centres, const int64_t* const* labels, const int64_t* node_off,
                          const int64_t* sub_off, int32_t* out_centre_row, int32_t* out_labels,
                          int32_t* out_task_row_ptr, int32_t* out_task_sub_ptr, int32_t* out_tile_row0,
                          int32_t* out_tile_nrows, int32_t* out_tile_task);
int64_t gmeta_host_active_rows(const int32_t* indptr, const int32_t* indices, int64_t n_nodes, const int32_t* centre_row,
                               int64_t n_centres, const int64_t* node_off, int32_t n_tasks, int32_t n_layers,
                               uint8_t* flags, int64_t* scratch, int32_t* buf, int64_t off, int64_t* seg_off,
                               int64_t* seg_n, int32_t* out_centre_pos);
int gmeta_host_validate_labels(int32_t n_tasks, const int64_t* const* y_spt, const int32_t* n_spt,
                               const int64_t* const* y_qry, const int32_t* n_qry, int32_t k_spt);

/* Centre-row readout + linear head (learner.py:159-175):
 *   r_s = H[centre_row[s]]   (link_pred: H[centre_row[2s]] || H[centre_row[2s+1]])
 *   logits[s,c] = sum_k r_s[k] * Wlin[t][c,k] + blin[t][c],  t = task of subgraph s. */
int gmeta_readout_linear_fwd(const float* H, int32_t ld_h, int32_t hid,
                             const int32_t* centre_row, int32_t centres_per_subgraph,
                             const int32_t* task_sub_ptr, int32_t n_tasks, int32_t n_subgraphs,
                             const float* Wlin, int64_t w_task_stride,
                             const float* blin, int64_t b_task_stride, int32_t n_out,
                             float* logits, void* stream);

/* Backward of the above plus the ReLU mask of the last GCN layer:
 *   dWlin[t][c,k] = sum_s dlogits[s,c] r_s[k];  dblin[t][c] = sum_s dlogits[s,c];
 *   dZ[n_dz_rows, ld_h] = 0 everywhere except dZ[pos(centre rows)] += (H > 0) ? dlogits[s,:] . Wlin[t][:,k] : 0
 * with pos(v) = row_pos ? row_pos[v] : v (row_pos: compact dZ over the active rows). */
int gmeta_readout_linear_bwd(const float* H, int32_t ld_h, int32_t hid, int32_t n_dz_rows,
                             const int32_t* row_pos, const int32_t* centre_row, int32_t centres_per_subgraph,
                             const int32_t* task_sub_ptr, int32_t n_tasks, int32_t n_subgraphs,
                             const float* Wlin, int64_t w_task_stride, int32_t n_out,
                             const float* dlogits,
                             float* dWlin, int64_t dw_task_stride,
                             float* dblin, int64_t db_task_stride,
                             float* dZ, void* stream);

/* row_pos[0..n_nodes) = -1, then row_pos[rows[i]] = i  (row -> position in an active-row list). */
int gmeta_build_row_pos(const int32_t* rows, int32_t n_rows, int32_t n_nodes, int32_t* row_pos,
                        void* stream);

/* class_pos / class_occ / n_classes from raw labels, per task (the `torch.unique` +
 * `eq(c).nonzero()` bookkeeping of meta.py:32-42,60-66), on device. */
int gmeta_proto_label_prep(const int32_t* labels, const int32_t* task_sub_ptr, int32_t n_tasks,
                           int32_t* class_pos, int32_t* class_occ, int32_t* n_classes,
                           void* stream);

/* proto_loss_spt (meta.py:28-54) for every task at once, forward + backward:
 *   protos[t][c,:] = mean of the first n_support logits of class c; loss/acc over those rows;
 *   dlogits = grad_scale * dloss/dlogits (including the path through the prototypes).
 * protos is [T, max_classes, n_out]; loss, acc are [T] written at stride out_stride;
 * max_rows_per_task = the largest number of subgraphs any one task has (sizes shared memory). */
int gmeta_proto_loss_spt(const float* logits, int32_t n_out, const int32_t* task_sub_ptr,
                         int32_t n_tasks, const int32_t* class_pos, const int32_t* class_occ,
                         const int32_t* n_classes, int32_t n_support, int32_t max_classes,
                         int32_t max_rows_per_task, float grad_scale, float* protos, float* loss,
                         float* acc, int32_t out_stride, float* dlogits, void* stream);

/* proto_loss_qry (meta.py:56-79): loss/acc of query logits against given prototypes
 * (n_classes[t] = number of prototypes of task t, i.e. the SUPPORT set's class count);
 * dlogits and dprotos (both optional, may be NULL) are grad_scale * dloss/d(.). */
int gmeta_proto_loss_qry(const float* logits, int32_t n_out, const int32_t* task_sub_ptr,
                         int32_t n_tasks, const int32_t* class_pos, const int32_t* n_classes,
                         const float* protos, int32_t max_classes, int32_t max_rows_per_task,
                         float grad_scale, float* loss, float* acc, int32_t out_stride,
                         float* dlogits, float* dprotos, void* stream);

/* Gradient reaching the SUPPORT logits through the prototypes used by the final query loss
 * (meta.py:54 returns them un-detached):
 *   dlogits_spt[s,:] = class_occ[s] < n_support ? dprotos[t][class_pos[s],:] / n_support : 0. */
int gmeta_proto_grad_to_support(const float* dprotos, int32_t n_out, int32_t max_classes,
                                const int32_t* task_sub_ptr, int32_t n_tasks,
                                const int32_t* class_pos, const int32_t* class_occ,
                                int32_t n_support, int32_t n_subgraphs, float* dlogits_spt,
                                void* stream);

/* fast[t][p] = w_in[t*w_in_task_stride + p] - lr * grad[t][p]   (meta.py:126,151). */
int gmeta_sgd_update(const float* w_in, int64_t w_in_task_stride, const float* grad,
                     float lr, int32_t n_tasks, int32_t n_params, float* w_out, void* stream);

/* out[p] = sum_t a[t][p] (+ sum_t b[t][p] if b != NULL), tasks summed in index order. */
int gmeta_sum_over_tasks(const float* a, const float* b, int32_t n_tasks, int32_t n_params,
                         float* out, void* stream);

/* torch.optim.Adam step (meta.py:97,169; betas/eps/lr given, no weight decay, no amsgrad) on
 * the flat parameter buffer.  `step` is the 1-based step count AFTER this update.  If
 * loss_gate != NULL and *loss_gate is NaN the update is skipped entirely (meta.py:163-164);
 * *skipped (optional) is set to 1/0 accordingly.  grad is multiplied by grad_scale first. */
int gmeta_adam_update(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                      int32_t n_params, double lr, double beta1, double beta2, double eps,
                      int32_t step, float grad_scale, const float* loss_gate, int32_t* skipped,
                      void* stream);

/* The same Adam step with every changing scalar on the DEVICE, so that the launches are identical from step to step
 * and can be replayed from a CUDA graph: `state` is a caller-owned, zero-initialised int32[8] -- [0] the step count
 * (incremented only when the update is applied, like torch.optim.Adam under the reference's NaN skip,
 * meta.py:163-169), [1] the skipped flag of the last call, [2..3] bias-correction scalars.  The gate is
 * *loss_sum * loss_scale (meta.py:161: sum of the tasks' last query losses / task_num); loss_sum == NULL never skips.
 * If step_out != NULL: step_out[k] = acc_sums[k] * loss_scale for k < n_acc (meta.py:171), step_out[n_acc] = the
 * gate value, step_out[n_acc + 1] = skipped (0 / 1).  Two launches. */
int gmeta_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int32_t n_params,
                    double lr, double beta1, double beta2, double eps, int32_t* state, float grad_scale,
                    const float* loss_sum, float loss_scale, const float* acc_sums, int32_t n_acc,
                    float* step_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Whole first-order ProtoMAML inner loop for every task of a packed meta-batch, enqueued
 * from C++ in one call (replaces the body of Meta.forward_ProtoMAML, meta.py:118-161, and of
 * finetunning_ProtoMAML, meta.py:193-230, minus the optimizer step).
 * ------------------------------------------------------------------------------------- */
typedef struct gmeta_step_args {
  gmeta_model_t model;
  gmeta_packed_set_t spt;
  gmeta_packed_set_t qry;
  const float* feat_table;       /* [rows, ld_feat] device-resident features of all graphs */
  int32_t ld_feat;
  const float* theta;            /* [P] meta-parameters */
  int32_t update_step;           /* K >= 1 (training needs K >= 2, as in the reference) */
  int32_t n_support;             /* k_spt */
  int32_t max_classes;
  int32_t spt_max_rows_per_task; /* most subgraphs any one task has in spt / qry (0 = use S) */
  int32_t qry_max_rows_per_task;
  float update_lr;
  float grad_scale;              /* 1 / global task_num (meta.py:161) */
  int32_t compute_meta_grad;     /* 1: training step; 0: finetunning (no backward of the query loss) */
  int32_t dense_backward;        /* 1: back-propagate over every row like the reference's autograd;
                                    0: skip the structurally-zero rows (packed set's act_* lists) */
  int32_t impl;                  /* GMETA_IMPL_* for the GCN layers */
  int32_t pruned_forward;        /* 1: every forward computes only the rows the read-out depends on (the same
                                    active-row lists: layer l over act_rows[l], compact activations) -- exact,
                                    because only centre rows are read out (learner.py:166-170); needs
                                    dense_backward == 0.  0: every row of every layer, like the reference. */
  /* outputs */
  float* meta_grad;              /* [P] sum over this call's tasks of d(grad_scale*loss_q^K)/d(theta) */
  float* loss_q;                 /* [T, K+1] query loss per task per step */
  float* acc_q;                  /* [T, K+1] query accuracy per task per step */
  float* loss_s;                 /* [T, K]   support loss per task per step */
  float* logits_spt0;            /* optional [S_spt, n_out]: support logits of step 0 (debug/parity) */
  void* workspace;
  int64_t workspace_bytes;
  const float* feat_rowmax;      /* optional [rows of feat_table]: max |feat_table[r, :]| (gmeta_row_absmax, once per
                                    table).  With it the full-formulation forwards (pruned_forward == 0) take the
                                    CTA-pair tensor-core layer path where the shape allows. */
  void* aux_stream;              /* optional second cudaStream_t (not the one passed to gmeta_maml_step): the query
                                    forwards that need no gradient are enqueued on it, beside the support chain of
                                    the following steps; forked from / joined back into the main stream with events,
                                    so the call still behaves like work on the main stream only (and can be captured
                                    into a CUDA graph).  NULL: everything on the main stream. */
  float* step_stats;             /* optional [K+2]: sum_t loss_q[t][K], then sum_t acc_q[t][0..K] (the scalars of the
                                    meta-step's all-reduce, meta.py:161,171), written by one extra launch */
} gmeta_step_args_t;

int64_t gmeta_maml_step_workspace_bytes(const gmeta_step_args_t* args);
int gmeta_maml_step(const gmeta_step_args_t* args, void* stream);
/* number of kernel launches the last gmeta_maml_step call on this thread enqueued */
int gmeta_last_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * Device-side h-hop local-subgraph extraction for a whole meta-batch of requests (replaces
 * Subgraphs.generate_subgraph / generate_subgraph_link_pred, subgraph_data_processing.py:295-346),
 * emitting the packed-set CSR directly.  The parent graphs are ONE int32 CSR by destination over
 * their concatenated node ranges; request r = (req_a[r], optional req_b[r], node range
 * [req_lo[r], req_hi[r]) of its graph), all global ids.
 *   closure  = <= hops_a in-hops from a, plus b and (hops_b == 1) b's in-neighbours -- the reference
 *              takes 2 hops from the first endpoint and, through :332, ONE from the second;
 *   sampling = if |closure| > sample_nodes: the sample_nodes nodes with the smallest values of a
 *              counter-based hash of (seed, r, node) -- a uniform sample without replacement --
 *              then the centre(s) re-added (:312-314, :337-339);
 *   output   = nodes in ascending id order (np.unique), node-induced edges in parent-CSR order with
 *              multiplicity, local ids = rank: identical to the host extractor whenever the closure
 *              fits (integer work, bit-exact).
 * Two calls: gmeta_khop_select writes node_ptr / edge_ptr [n_req + 1] (exclusive sums; the last
 * entries are the packed totals the caller reads to size the outputs), gmeta_khop_build fills
 *   out_indptr [N+1], out_indices [E] (packed row ids), out_parent [N] (id inside the graph),
 *   out_global [N] or NULL (global id = feature-table row), out_centre [n_req * (req_b ? 2 : 1)].
 * Same 256-byte aligned workspace for both (gmeta_khop_workspace_bytes); sample_nodes <= 2046. */
int64_t gmeta_khop_workspace_bytes(int32_t n_req, int32_t sample_nodes, int32_t max_graph_nodes);
int gmeta_khop_select(const int32_t* indptr, const int32_t* indices, const int32_t* req_a,
                      const int32_t* req_b, const int32_t* req_lo, const int32_t* req_hi, int32_t n_req,
                      int32_t hops_a, int32_t hops_b, int32_t sample_nodes, int32_t max_graph_nodes,
                      uint64_t seed, int32_t* node_ptr, int32_t* edge_ptr, int32_t* closure_size,
                      void* workspace, int64_t workspace_bytes, void* stream);
int gmeta_khop_build(const int32_t* indptr, const int32_t* indices, const int32_t* req_a,
                     const int32_t* req_b, const int32_t* req_lo, int32_t n_req, int32_t sample_nodes,
                     int32_t max_graph_nodes, const int32_t* node_ptr, const int32_t* edge_ptr,
                     int32_t* out_indptr, int32_t* out_indices, int32_t* out_parent, int32_t* out_global,
                     int32_t* out_centre, void* workspace, int64_t workspace_bytes, void* stream);

/* The meta-step as an updatable CUDA graph, for batches whose shapes change from step to step.  `prepare` captures
 * what gmeta_maml_step(args) would enqueue (no device work is done; capture is thread-local, `stream` must not be the
 * legacy default stream) and updates the handle's executable graph in place -- re-instantiating it only when the
 * launch topology changed; `launch` runs it on any stream.  Prepare batch i+1 while step i runs and the ~200 launches
 * of a step leave the critical path.  The argument buffers must stay valid until the launch has completed. */
int gmeta_step_graph_create(void** handle);
void gmeta_step_graph_destroy(void* handle);
int gmeta_step_graph_prepare(void* handle, const gmeta_step_args_t* args, void* stream);
int gmeta_step_graph_launch(void* handle, void* stream);
int gmeta_step_graph_stats(void* handle, int32_t* updates, int32_t* instantiations);

/* Packed-set assembly on the device (csrc/batch_assemble.cu): what gmeta_packed_set_t needs on top of the
 * extractor's output, derived in HBM without a host round trip -- the structure work of dgl.batch
 * (subgraph_data_processing.py:399-406) and of the host packer.  Inputs: the packed CSR by destination
 * (indptr [N+1], indices [E]), sub_node_ptr [S+1] (first packed row of every subgraph: node_ptr of
 * gmeta_khop_select), task_sub_ptr [T+1] (device; NULL: sub_node_ptr holds the T+1 task row pointers themselves),
 * centre_row [n_centres].  Outputs (device, caller-sized):
 *   t_indptr [N+1], t_indices [E]         CSR by source, destinations ascending inside a row
 *   task_row_ptr [T+1]; tile_row0/nrows/task [<= ceil(N/128) + T]   tiles of <= 128 rows inside one task
 *   per GCN layer l < n_layers: act_rows[l] [<= N] (ascending), act_task_ptr[l] [T+1],
 *     act_tile_row0/nrows/task[l] [<= ceil(N/128) + T]: the rows whose gradient is not structurally zero
 *     (centres at the last layer, the in-neighbours of the layer above below it, learner.py:166-170)
 *   centre_pos [n_centres]                position of every centre among act_rows[n_layers - 1]
 *   counts [2 + 2 * n_layers]             n_tiles, rows of the largest task, n_act[l]..., n_act_tiles[l]...
 * The pointer arrays are HOST arrays of device pointers.  Integer work: identical to the host packer. */
int64_t gmeta_packed_set_finish_workspace_bytes(int32_t n_nodes, int32_t n_edges, int32_t n_layers);
int gmeta_packed_set_finish(const int32_t* indptr, const int32_t* indices, int32_t n_nodes, int32_t n_edges,
                            const int32_t* sub_node_ptr, const int32_t* task_sub_ptr, int32_t n_tasks,
                            const int32_t* centre_row, int32_t n_centres, int32_t n_layers,
                            int32_t* t_indptr, int32_t* t_indices, int32_t* task_row_ptr, int32_t* tile_row0,
                            int32_t* tile_nrows, int32_t* tile_task, int32_t* const* act_rows,
                            int32_t* const* act_task_ptr, int32_t* const* act_tile_row0,
                            int32_t* const* act_tile_nrows, int32_t* const* act_tile_task,
                            int32_t* centre_pos, int32_t* counts, void* workspace, int64_t workspace_bytes,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GMETA_B200_H_ */
