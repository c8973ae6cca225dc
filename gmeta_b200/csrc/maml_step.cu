// C++ driver that enqueues the whole first-order ProtoMAML inner loop for every task of a
// packed meta-batch (reference G-Meta/meta.py:118-161 forward_ProtoMAML and :193-230
// finetunning_ProtoMAML, minus the optimizer step): K support forward/backward + SGD steps,
// K+1 query forwards with their prototype losses, and -- for training -- the backward of the
// last query loss through the query forward (weights fast_K) and, via the prototypes, through
// the last support forward (weights fast_{K-1}); SURVEY 3.2.  No host synchronisation, no
// allocation: every buffer is carved out of the caller's workspace.
#include "common.cuh"

namespace gmeta {

int fill_identity_graph(int32_t* iota, float* ones, int n, cudaStream_t stream);

int proto_loss_launch(bool spt, const float* logits, int n_out, const int32_t* task_sub_ptr, int n_tasks,
                      const int32_t* class_pos, const int32_t* class_occ, const int32_t* n_classes,
                      int n_support, int max_classes, int max_rows, float grad_scale, float* protos,
                      float* loss, float* acc, int out_stride, float* dlogits, float* dprotos,
                      cudaStream_t s);

namespace {

inline int64_t align_up(int64_t x) { return (x + 255) / 256 * 256; }

struct Carver {
  char* base;
  int64_t off;
  template <typename T>
  T* take(int64_t count) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += align_up(count * (int64_t)sizeof(T));
    return p;
  }
};

struct StepBuffers {
  float* fast[2];
  float* g_spt;
  float* g_qry;
  float* act_spt[GMETA_MAX_LAYERS];
  float* act_qry[GMETA_MAX_LAYERS];
  float* dz_spt[2];          // dense: [N, ld]; sparse backward: [max n_act, ld]
  float* dz_qry[2];
  float* logits_s;
  float* logits_q;
  float* dlogits_s;
  float* dlogits_q;
  float* protos;
  float* dprotos;
  float* acc_s;
  void* wgrad_ws;
  int64_t wgrad_ws_bytes;
  // pruned forward: the first layer's normalised neighbourhood sums over its active rows (weights do not
  // enter them: computed once per meta-step) and the identity graph the layer kernels then run on
  float* agg_spt[GMETA_MAX_LAYERS];   // [n_act[l], ld_in(l)]: aggregated input of layer l (l = 0: of the features, once per step)
  float* agg_qry[GMETA_MAX_LAYERS];
  float* dagg;                        // aggregated dZ of the data-gradient step (one at a time)
  int n_ident;                        // rows of the identity graph
  int32_t* iota;
  float* ones;
  // full-formulation forwards through the CTA-pair path: structure plans (layer 0 maps rows through feat_row,
  // the upper layers do not) and the row abs-max chain
  bool use_ex;
  void* plan_spt[2];
  void* plan_qry[2];
  float* rmax_spt[2];
  float* rmax_qry[2];
  void* layer_ws;            // weight image of the tensor-core layer kernel
  int64_t layer_ws_bytes;
  int ld[GMETA_MAX_LAYERS];
  int64_t total;
};

int validate(const gmeta_step_args_t* a) {
  if (!a) return GMETA_ERR_BAD_ARG;
  const gmeta_model_t& m = a->model;
  if (m.n_layers < 1 || m.n_layers > GMETA_MAX_LAYERS || m.n_out <= 0 || m.n_params_padded <= 0)
    return GMETA_ERR_BAD_ARG;
  for (int l = 0; l < m.n_layers; ++l) {
    if (m.f_in[l] <= 0 || m.f_out[l] <= 0) return GMETA_ERR_BAD_ARG;
    if (l > 0 && m.f_in[l] != m.f_out[l - 1]) return GMETA_ERR_BAD_ARG;
  }
  if (a->spt.n_tasks <= 0 || a->spt.n_tasks != a->qry.n_tasks) return GMETA_ERR_BAD_ARG;
  if (a->update_step < 1 || a->n_support < 1 || a->max_classes < 1) return GMETA_ERR_BAD_ARG;
  if (a->compute_meta_grad && a->update_step < 2) return GMETA_ERR_UNSUPPORTED;  // meta.py:161 has no grad path at K=1
  if (a->pruned_forward && a->dense_backward) return GMETA_ERR_UNSUPPORTED;
  if (a->pruned_forward && (!a->spt.centre_pos || !a->qry.centre_pos)) return GMETA_ERR_BAD_ARG;
  return GMETA_OK;
}

void carve(const gmeta_step_args_t* a, void* ws, StepBuffers& b) {
  const gmeta_model_t& m = a->model;
  const int64_t T = a->spt.n_tasks, P = m.n_params_padded, C = m.n_out, MC = a->max_classes;
  const int64_t Ns = a->spt.n_nodes, Nq = a->qry.n_nodes, Ss = a->spt.n_subgraphs, Sq = a->qry.n_subgraphs;
  Carver c{reinterpret_cast<char*>(ws), 0};
  int ld_max = 4;
  for (int l = 0; l < m.n_layers; ++l) {
    b.ld[l] = round_up(m.f_out[l], 4);
    if (b.ld[l] > ld_max) ld_max = b.ld[l];
  }
  b.fast[0] = c.take<float>(T * P);
  b.fast[1] = c.take<float>(T * P);
  b.g_spt = c.take<float>(T * P);
  b.g_qry = c.take<float>(T * P);
  // pruned forward: activations are compact over the active rows of each layer
  for (int l = 0; l < m.n_layers; ++l)
    b.act_spt[l] = c.take<float>((a->pruned_forward ? (int64_t)a->spt.n_act[l] : Ns) * b.ld[l]);
  for (int l = 0; l < m.n_layers; ++l)
    b.act_qry[l] = c.take<float>((a->pruned_forward ? (int64_t)a->qry.n_act[l] : Nq) * b.ld[l]);
  int64_t rows_s = Ns, rows_q = Nq;   // rows of the dZ buffers
  if (!a->dense_backward) {
    rows_s = rows_q = 1;
    for (int l = 0; l < m.n_layers; ++l) {
      if (a->spt.n_act[l] > rows_s) rows_s = a->spt.n_act[l];
      if (a->qry.n_act[l] > rows_q) rows_q = a->qry.n_act[l];
    }
  }
  const int64_t n0s = a->pruned_forward ? a->spt.n_act[0] : 0, n0q = a->pruned_forward ? a->qry.n_act[0] : 0;
  b.agg_spt[0] = c.take<float>(n0s * a->ld_feat);
  b.agg_qry[0] = c.take<float>(n0q * a->ld_feat);
  int64_t n_max = n0s > n0q ? n0s : n0q;
  for (int l = 1; l < m.n_layers; ++l) {
    const int64_t ns = a->pruned_forward ? a->spt.n_act[l] : 0, nq = a->pruned_forward ? a->qry.n_act[l] : 0;
    b.agg_spt[l] = c.take<float>(ns * b.ld[l - 1]);
    b.agg_qry[l] = c.take<float>(nq * b.ld[l - 1]);
    if (ns > n_max) n_max = ns;
    if (nq > n_max) n_max = nq;
  }
  b.dagg = c.take<float>(a->pruned_forward && m.n_layers > 1 ? n_max * ld_max : 0);
  b.n_ident = a->pruned_forward ? (int)n_max : 0;
  b.iota = c.take<int32_t>(a->pruned_forward ? n_max + 1 : 0);
  b.ones = c.take<float>(a->pruned_forward ? n_max : 0);
  b.dz_spt[0] = c.take<float>(rows_s * ld_max);
  b.dz_spt[1] = c.take<float>(m.n_layers > 1 ? rows_s * ld_max : 0);
  b.dz_qry[0] = c.take<float>(a->compute_meta_grad ? rows_q * ld_max : 0);
  b.dz_qry[1] = c.take<float>(a->compute_meta_grad && m.n_layers > 1 ? rows_q * ld_max : 0);
  b.logits_s = c.take<float>(Ss * C);
  b.logits_q = c.take<float>(Sq * C);
  b.dlogits_s = c.take<float>(Ss * C);
  b.dlogits_q = c.take<float>(Sq * C);
  b.protos = c.take<float>(T * MC * C);
  b.dprotos = c.take<float>(T * MC * C);
  b.acc_s = c.take<float>(T);
  b.wgrad_ws_bytes = 0;
  for (int l = 0; l < m.n_layers; ++l) {
    const int64_t w = gmeta_gcn_layer_wgrad_workspace_bytes((int)T, m.f_in[l], m.f_out[l]);
    if (w > b.wgrad_ws_bytes) b.wgrad_ws_bytes = w;
  }
  b.wgrad_ws = c.take<char>(b.wgrad_ws_bytes);
  b.use_ex = !a->pruned_forward && a->feat_rowmax && (a->impl == GMETA_IMPL_AUTO || a->impl == GMETA_IMPL_TCPAIR);
  for (int i = 0; i < 2; ++i) {
    const bool need = b.use_ex && (i == 0 || m.n_layers > 1);
    b.plan_spt[i] = c.take<char>(need ? gmeta_layer_plan_bytes(a->spt.n_tiles, (int)T, (int)Ns, a->spt.n_edges) : 0);
    b.plan_qry[i] = c.take<char>(need ? gmeta_layer_plan_bytes(a->qry.n_tiles, (int)T, (int)Nq, a->qry.n_edges) : 0);
    b.rmax_spt[i] = c.take<float>(need ? Ns : 0);
    b.rmax_qry[i] = c.take<float>(need ? Nq : 0);
  }
  b.layer_ws_bytes = 0;
  for (int l = 0; l < m.n_layers; ++l) {
    int64_t w = gmeta_gcn_layer_fwd_workspace_bytes((int)T, P, m.f_in[l], m.f_out[l], a->impl);
    if (b.use_ex) {
      const int64_t ws = gmeta_gcn_layer_fwd_ex_workspace_bytes((int)T, P, a->spt.n_tiles, (int)Ns, a->spt.n_edges,
                                                                 m.f_in[l], m.f_out[l], a->impl);
      const int64_t wq = gmeta_gcn_layer_fwd_ex_workspace_bytes((int)T, P, a->qry.n_tiles, (int)Nq, a->qry.n_edges,
                                                                 m.f_in[l], m.f_out[l], a->impl);
      if (ws > w) w = ws;
      if (wq > w) w = wq;
    }
    if (w > b.layer_ws_bytes) b.layer_ws_bytes = w;
  }
  b.layer_ws = c.take<char>(b.layer_ws_bytes);
  b.total = c.off;
}

struct Runner {
  const gmeta_step_args_t* a;
  StepBuffers b;
  cudaStream_t s;
  int rc = GMETA_OK;

  bool ok() const { return rc == GMETA_OK; }
  void run(int code) { if (rc == GMETA_OK) rc = code; }

  void forward(const gmeta_packed_set_t& set, float* const* act, const float* W, int64_t stride, float* logits) {
    const gmeta_model_t& m = a->model;
    const bool pruned = a->pruned_forward != 0;
    for (int l = 0; l < m.n_layers && ok(); ++l) {
      if (pruned) {
        // The neighbourhood sums of the layer's active rows come from a chip-wide gather (one warp per row;
        // layer 0: cached for the whole meta-step, the weights do not enter it), the contraction then runs
        // dense on the pre-summed rows: identity graph, unit norms.  Every in-neighbour of an active row of
        // layer l is an active row of layer l-1 by construction, so row_pos[l-1] maps it to its compact row.
        float* agg = (&set == &a->spt ? b.agg_spt : b.agg_qry)[l];
        const int ld_agg = l == 0 ? a->ld_feat : b.ld[l - 1];
        if (l > 0)
          run(gmeta_aggregate_rows(act[l - 1], b.ld[l - 1], set.row_pos[l - 1], set.act_rows[l], set.indptr, set.indices,
                                   set.norm, set.n_act[l], m.f_in[l], 1, agg, ld_agg, s));
        run(gmeta_gcn_layer_fwd(agg, ld_agg, nullptr, nullptr, b.iota, b.iota, b.ones, set.act_tile_row0[l],
                                set.act_tile_nrows[l], set.act_tile_task[l], set.n_act_tiles[l], set.n_tasks,
                                W + m.w_off[l], stride, m.f_out[l], 0, W + m.b_off[l], stride, m.f_in[l], m.f_out[l], 1,
                                nullptr, act[l], b.ld[l], a->impl, b.layer_ws, b.layer_ws_bytes, s));
        continue;
      }
      const float* in = l == 0 ? a->feat_table : act[l - 1];
      const int ld_in = l == 0 ? a->ld_feat : b.ld[l - 1];
      // pruned: layer l over its active rows only (compact output); its inputs are the compact
      // activations of layer l-1, addressed through row_pos[l-1] (every in-neighbour of an active
      // row of layer l is an active row of layer l-1 by construction)
      const int32_t* map = l == 0 ? set.feat_row : (pruned ? set.row_pos[l - 1] : nullptr);
      if (b.use_ex) {
        // every row of every layer (the reference's formulation): CTA-pair tensor-core path where the shape allows,
        // with the structure plan built once per step and the row abs-max handed from layer to layer
        const bool is_spt = &set == &a->spt;
        float* const* rmax = is_spt ? b.rmax_spt : b.rmax_qry;
        void* const* plan = is_spt ? b.plan_spt : b.plan_qry;
        run(gmeta_gcn_layer_fwd_ex(in, ld_in, map, nullptr, set.indptr, set.indices, set.norm, set.tile_row0,
                                   set.tile_nrows, set.tile_task, set.n_tiles, set.n_tasks, W + m.w_off[l], stride,
                                   m.f_out[l], 0, W + m.b_off[l], stride, m.f_in[l], m.f_out[l], 1, nullptr, act[l],
                                   b.ld[l], a->impl, b.layer_ws, b.layer_ws_bytes, set.n_nodes, set.n_edges,
                                   l == 0 ? a->feat_rowmax : rmax[(l - 1) & 1], l + 1 < m.n_layers ? rmax[l & 1] : nullptr,
                                   plan[l == 0 ? 0 : 1], s));
        continue;
      }
      run(gmeta_gcn_layer_fwd(in, ld_in, map, pruned ? set.act_rows[l] : nullptr, set.indptr, set.indices, set.norm,
                              pruned ? set.act_tile_row0[l] : set.tile_row0,
                              pruned ? set.act_tile_nrows[l] : set.tile_nrows,
                              pruned ? set.act_tile_task[l] : set.tile_task,
                              pruned ? set.n_act_tiles[l] : set.n_tiles, set.n_tasks,
                              W + m.w_off[l], stride, m.f_out[l], 0, W + m.b_off[l], stride, m.f_in[l],
                              m.f_out[l], 1, nullptr, act[l], b.ld[l], a->impl, b.layer_ws, b.layer_ws_bytes, s));
    }
    const int L = m.n_layers;
    run(gmeta_readout_linear_fwd(act[L - 1], b.ld[L - 1], m.f_out[L - 1], pruned ? set.centre_pos : set.centre_row,
                                 set.centres_per_subgraph, set.task_sub_ptr, set.n_tasks, set.n_subgraphs,
                                 W + m.wlin_off, stride, W + m.blin_off, stride, m.n_out, logits, s));
  }

  void backward(const gmeta_packed_set_t& set, float* const* act, float* const* dz, const float* W,
                int64_t stride, const float* dlogits, float* gout) {
    const gmeta_model_t& m = a->model;
    const int L = m.n_layers;
    const int64_t P = m.n_params_padded;
    const bool sparse = !a->dense_backward;
    const bool pruned = a->pruned_forward != 0;
    int cur = 0;
    // dZ of the last GCN layer: zero except at the centre rows (dense [N, ld] like the reference's
    // autograd, or compact over the active rows)
    // pruned: H is compact in the order of act_rows[L-1], and so is dZ -> centre positions index both
    run(gmeta_readout_linear_bwd(act[L - 1], b.ld[L - 1], m.f_out[L - 1], sparse ? set.n_act[L - 1] : set.n_nodes,
                                 pruned ? nullptr : (sparse ? set.row_pos[L - 1] : nullptr),
                                 pruned ? set.centre_pos : set.centre_row, set.centres_per_subgraph,
                                 set.task_sub_ptr, set.n_tasks, set.n_subgraphs, W + m.wlin_off, stride, m.n_out,
                                 dlogits, gout + m.wlin_off, P, gout + m.blin_off, P, dz[cur], s));
    for (int l = L - 1; l >= 0 && ok(); --l) {
      const float* in = l == 0 ? a->feat_table : act[l - 1];
      const int ld_in = l == 0 ? a->ld_feat : b.ld[l - 1];
      if (pruned) {
        // dW_l = (n_v M_v)^T dZ_l with the aggregated rows the forward kept; the data gradient is the same two
        // steps as a forward on the transposed graph: gather n_v dZ_v over the out-neighbours that are active at
        // layer l (the others are dropped through row_pos[l]), then the dense contraction with W^T, masked by
        // the ReLU of the layer below.
        const float* agg = (&set == &a->spt ? b.agg_spt : b.agg_qry)[l];
        run(gmeta_gcn_layer_wgrad(agg, ld_in, nullptr, nullptr, b.iota, b.iota, b.ones, set.act_task_ptr[l], set.n_tasks,
                                  dz[cur], b.ld[l], m.f_in[l], m.f_out[l], gout + m.w_off[l], P, gout + m.b_off[l], P,
                                  b.wgrad_ws, b.wgrad_ws_bytes, s));
        if (l > 0) {
          run(gmeta_aggregate_rows(dz[cur], b.ld[l], set.row_pos[l], set.act_rows[l - 1], set.t_indptr, set.t_indices,
                                   set.norm, set.n_act[l - 1], m.f_out[l], 1, b.dagg, b.ld[l], s));
          run(gmeta_gcn_layer_fwd(b.dagg, b.ld[l], nullptr, nullptr, b.iota, b.iota, b.ones, set.act_tile_row0[l - 1],
                                  set.act_tile_nrows[l - 1], set.act_tile_task[l - 1], set.n_act_tiles[l - 1], set.n_tasks,
                                  W + m.w_off[l], stride, m.f_out[l], 1, nullptr, 0, m.f_out[l], m.f_in[l], 2, act[l - 1],
                                  dz[cur ^ 1], b.ld[l - 1], a->impl, b.layer_ws, b.layer_ws_bytes, s));
          cur ^= 1;
        }
        continue;
      }
      run(gmeta_gcn_layer_wgrad(in, ld_in, l == 0 ? set.feat_row : (pruned ? set.row_pos[l - 1] : nullptr),
                                sparse ? set.act_rows[l] : nullptr,
                                set.indptr, set.indices, set.norm, sparse ? set.act_task_ptr[l] : set.task_row_ptr,
                                set.n_tasks, dz[cur], b.ld[l], m.f_in[l], m.f_out[l], gout + m.w_off[l], P,
                                gout + m.b_off[l], P, b.wgrad_ws, b.wgrad_ws_bytes, s));
      if (l > 0) {
        // data gradient = the forward kernel on the transposed graph with W^T, masked by the
        // ReLU of the layer below (features carry no gradient, so layer 0 stops here).  Sparse:
        // only rows with an out-edge into an active row of layer l are computed, reading the
        // compact dZ_l through row_pos[l] (inactive out-neighbours are dropped).
        run(gmeta_gcn_layer_fwd(dz[cur], b.ld[l], sparse ? set.row_pos[l] : nullptr,
                                sparse ? set.act_rows[l - 1] : nullptr, set.t_indptr, set.t_indices, set.norm,
                                sparse ? set.act_tile_row0[l - 1] : set.tile_row0,
                                sparse ? set.act_tile_nrows[l - 1] : set.tile_nrows,
                                sparse ? set.act_tile_task[l - 1] : set.tile_task,
                                sparse ? set.n_act_tiles[l - 1] : set.n_tiles, set.n_tasks, W + m.w_off[l], stride,
                                m.f_out[l], 1, nullptr, 0, m.f_out[l], m.f_in[l], pruned ? 2 : 0, act[l - 1], dz[cur ^ 1],
                                b.ld[l - 1], a->impl, b.layer_ws, b.layer_ws_bytes, s));
        cur ^= 1;
      }
    }
  }

  void qry_loss(int k, bool want_grad) {
    const gmeta_packed_set_t& q = a->qry;
    // prototypes (and their count) come from the support set of the same task (meta.py:132,154)
    run(proto_loss_launch(false, b.logits_q, a->model.n_out, q.task_sub_ptr, q.n_tasks, q.class_pos, nullptr,
                          a->spt.n_classes, 0, a->max_classes, max_rows_q, a->grad_scale, b.protos,
                          a->loss_q + k, a->acc_q + k, a->update_step + 1, want_grad ? b.dlogits_q : nullptr,
                          want_grad ? b.dprotos : nullptr, s));
  }
  int max_rows_s = 1, max_rows_q = 1;
};

}  // namespace
}  // namespace gmeta

using namespace gmeta;

// At the step level GMETA_IMPL_TCPAIR means "the CTA-pair path wherever a launch has the structure plan and the
// row abs-max it needs" -- which is what AUTO selects; the launches without them (pruned forwards on the identity
// graph, data gradients) cannot take it and fall back inside the library like AUTO does.
static gmeta_step_args_t step_level_args(const gmeta_step_args_t* a) {
  gmeta_step_args_t c = *a;
  if (c.impl == GMETA_IMPL_TCPAIR) c.impl = GMETA_IMPL_AUTO;
  return c;
}

extern "C" int64_t gmeta_maml_step_workspace_bytes(const gmeta_step_args_t* args) {
  if (validate(args) != GMETA_OK) return -1;
  const gmeta_step_args_t c = step_level_args(args);
  StepBuffers b;
  carve(&c, nullptr, b);
  return b.total;
}

extern "C" int gmeta_last_launch_count(void) { return g_launch_count; }

extern "C" int gmeta_maml_step(const gmeta_step_args_t* a_in, void* stream) {
  int rc = validate(a_in);
  if (rc != GMETA_OK) return rc;
  const gmeta_step_args_t a_copy = step_level_args(a_in);
  const gmeta_step_args_t* a = &a_copy;
  if (!a->workspace || !a->theta || !a->feat_table || !a->loss_q || !a->acc_q || !a->loss_s)
    return GMETA_ERR_BAD_ARG;
  if (a->compute_meta_grad && !a->meta_grad) return GMETA_ERR_BAD_ARG;
  if (!aligned16(a->workspace)) return GMETA_ERR_ALIGN;
  Runner r;
  r.a = a;
  r.s = (cudaStream_t)stream;
  carve(a, a->workspace, r.b);
  if (r.b.total > a->workspace_bytes) return GMETA_ERR_WORKSPACE;
  g_launch_count = 0;

  const gmeta_model_t& m = a->model;
  const gmeta_packed_set_t& sp = a->spt;
  const gmeta_packed_set_t& qr = a->qry;
  const int T = sp.n_tasks, K = a->update_step;
  const int64_t P = m.n_params_padded;
  StepBuffers& b = r.b;
  cudaStream_t s = r.s;
  // shared memory of the loss kernels is sized for the task with the most subgraphs
  r.max_rows_s = a->spt_max_rows_per_task > 0 ? a->spt_max_rows_per_task : sp.n_subgraphs;
  r.max_rows_q = a->qry_max_rows_per_task > 0 ? a->qry_max_rows_per_task : qr.n_subgraphs;

  if (cudaMemsetAsync(b.g_spt, 0, (size_t)T * P * sizeof(float), s) != cudaSuccess) return GMETA_ERR_LAUNCH;
  if (cudaMemsetAsync(b.g_qry, 0, (size_t)T * P * sizeof(float), s) != cudaSuccess) return GMETA_ERR_LAUNCH;
  r.run(gmeta_degree_norm(sp.indptr, sp.n_nodes, sp.norm, s));
  r.run(gmeta_degree_norm(qr.indptr, qr.n_nodes, qr.norm, s));
  r.run(gmeta_proto_label_prep(sp.labels, sp.task_sub_ptr, T, sp.class_pos, sp.class_occ, sp.n_classes, s));
  r.run(gmeta_proto_label_prep(qr.labels, qr.task_sub_ptr, T, qr.class_pos, qr.class_occ, qr.n_classes, s));
  if (!a->dense_backward) {   // row -> position maps of the active-row lists (structure only: once per step)
    for (int l = 0; l < m.n_layers; ++l) {
      r.run(gmeta_build_row_pos(sp.act_rows[l], sp.n_act[l], sp.n_nodes, sp.row_pos[l], s));
      if (a->compute_meta_grad || a->pruned_forward)
        r.run(gmeta_build_row_pos(qr.act_rows[l], qr.n_act[l], qr.n_nodes, qr.row_pos[l], s));
    }
  }

  if (b.use_ex) {
    for (int i = 0; i < (m.n_layers > 1 ? 2 : 1); ++i) {
      r.run(gmeta_layer_plan_build(sp.indptr, sp.indices, sp.norm, i == 0 ? sp.feat_row : nullptr, nullptr, sp.tile_row0,
                                   sp.tile_nrows, sp.tile_task, sp.n_tiles, T, sp.n_nodes, sp.n_edges, b.plan_spt[i], s));
      r.run(gmeta_layer_plan_build(qr.indptr, qr.indices, qr.norm, i == 0 ? qr.feat_row : nullptr, nullptr, qr.tile_row0,
                                   qr.tile_nrows, qr.tile_task, qr.n_tiles, T, qr.n_nodes, qr.n_edges, b.plan_qry[i], s));
    }
  }
  if (a->pruned_forward) {
    const int n0s = sp.n_act[0], n0q = qr.n_act[0];
    r.run(fill_identity_graph(b.iota, b.ones, b.n_ident, s));
    r.run(gmeta_aggregate_rows(a->feat_table, a->ld_feat, sp.feat_row, sp.act_rows[0], sp.indptr, sp.indices, sp.norm,
                               n0s, m.f_in[0], 1, b.agg_spt[0], a->ld_feat, s));
    r.run(gmeta_aggregate_rows(a->feat_table, a->ld_feat, qr.feat_row, qr.act_rows[0], qr.indptr, qr.indices, qr.norm,
                               n0q, m.f_in[0], 1, b.agg_qry[0], a->ld_feat, s));
  }

  for (int k = 0; k < K && r.ok(); ++k) {
    const float* Wcur = k == 0 ? a->theta : b.fast[(k - 1) & 1];
    const int64_t stride = k == 0 ? 0 : P;
    r.forward(sp, b.act_spt, Wcur, stride, b.logits_s);
    if (k == 0 && a->logits_spt0 && r.ok())
      if (cudaMemcpyAsync(a->logits_spt0, b.logits_s, (size_t)sp.n_subgraphs * m.n_out * sizeof(float),
                          cudaMemcpyDeviceToDevice, s) != cudaSuccess) r.rc = GMETA_ERR_LAUNCH;
    r.run(proto_loss_launch(true, b.logits_s, m.n_out, sp.task_sub_ptr, T, sp.class_pos, sp.class_occ,
                            sp.n_classes, a->n_support, a->max_classes, r.max_rows_s, 1.0f, b.protos,
                            a->loss_s + k, b.acc_s, K, b.dlogits_s, nullptr, s));
    r.backward(sp, b.act_spt, b.dz_spt, Wcur, stride, b.dlogits_s, b.g_spt);
    r.run(gmeta_sgd_update(Wcur, stride, b.g_spt, a->update_lr, T, (int)P, b.fast[k & 1], s));
    if (k == 0) {  // query loss / accuracy before the first update (meta.py:129-134)
      r.forward(qr, b.act_qry, a->theta, 0, b.logits_q);
      r.qry_loss(0, false);
    }
    r.forward(qr, b.act_qry, b.fast[k & 1], P, b.logits_q);
    r.qry_loss(k + 1, a->compute_meta_grad && k == K - 1);
  }
  if (a->compute_meta_grad && r.ok()) {
    r.backward(qr, b.act_qry, b.dz_qry, b.fast[(K - 1) & 1], P, b.dlogits_q, b.g_qry);
    r.run(gmeta_proto_grad_to_support(b.dprotos, m.n_out, a->max_classes, sp.task_sub_ptr, T, sp.class_pos,
                                      sp.class_occ, a->n_support, sp.n_subgraphs, b.dlogits_s, s));
    r.backward(sp, b.act_spt, b.dz_spt, b.fast[(K - 2) & 1], P, b.dlogits_s, b.g_spt);
    r.run(gmeta_sum_over_tasks(b.g_qry, b.g_spt, T, (int)P, a->meta_grad, s));
  }
  return r.rc;
}
