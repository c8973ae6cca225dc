// C++ driver that enqueues the whole first-order ProtoMAML inner loop for every task of a
// packed meta-batch (reference G-Meta/meta.py:118-161 forward_ProtoMAML and :193-230
// finetunning_ProtoMAML, minus the optimizer step): K support forward/backward + SGD steps,
// K+1 query forwards with their prototype losses, and -- for training -- the backward of the
// last query loss through the query forward (weights fast_K) and, via the prototypes, through
// the last support forward (weights fast_{K-1}); SURVEY 3.2.  No host synchronisation, no
// allocation: every buffer is carved out of the caller's workspace.
//
// Schedule.  The query forward of step k only produces an accuracy / loss (meta.py:131-141,152-157); nothing on the
// support chain depends on it.  With an auxiliary stream (gmeta_step_args_t::aux_stream) the K query forwards that
// need no gradient run on it, beside the support forward/backward/SGD chain of the following steps (event fork /
// join, prototypes and fast weights double buffered) -- under stream capture this becomes a two-branch CUDA graph.
// The fast weights' tensor-core operand images (both orientations of every layer that takes the tensor-core
// kernel) are written by the same launch as the SGD update, once per step, instead of once per layer call.
#include <vector>

#include "common.cuh"
#include "internal.cuh"

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
  float* protos[2];          // prototypes of the support forward of step k live in protos[k & 1]
  float* dprotos;
  float* acc_s;
  void* wgrad_ws;
  int64_t wgrad_ws_bytes;
  // pruned forward: the first layer's normalised neighbourhood sums over its active rows (weights do not
  // enter them: computed once per meta-step) and the identity graph the layer kernels then run on
  float* agg_spt[GMETA_MAX_LAYERS];   // [n_act[l], ld_in(l)]: aggregated input of layer l (l = 0: of the features, once per step)
  float* agg_qry[GMETA_MAX_LAYERS];
  float* dagg;                        // aggregated dZ of the data-gradient step (one at a time)
  // active out-neighbour lists of the pruned backward (layer l >= 1: rows of act[l-1] -> their out-neighbours in act[l])
  int32_t* aout_ptr_spt[GMETA_MAX_LAYERS];
  int32_t* aout_idx_spt[GMETA_MAX_LAYERS];
  int32_t* aout_ptr_qry[GMETA_MAX_LAYERS];
  int32_t* aout_idx_qry[GMETA_MAX_LAYERS];
  int32_t* aout_count;
  int n_ident;                        // rows of the identity graph
  float* agg_bwd;                     // full-formulation forwards + sparse backward: aggregated input of the layer whose
                                      // weight gradient is being taken, [max n_act[l], max ld_in]
  int32_t* iota;
  float* ones;
  // full-formulation forwards through the CTA-pair path: structure plans (layer 0 maps rows through feat_row,
  // the upper layers do not) and the row abs-max chain
  bool use_ex;
  void* plan_spt[2];
  void* plan_qry[2];
  float* rmax_spt[2];
  float* rmax_qry[2];
  void* layer_ws;            // weight image / scratch of the tensor-core layer kernels (main stream)
  void* layer_ws_aux;        // same for launches on the auxiliary stream
  int64_t layer_ws_bytes;
  // prepacked tensor-core operand images: [fwd orientation of layer l | transposed orientation of layer l >= 1]
  bool tc_img[GMETA_MAX_LAYERS][2];      // [l][orientation]: that launch takes the streamed-weight tensor-core kernel
  TcPackPlan pack;
  float* img_theta;          // one copy (theta)
  float* img_fast[2];        // T copies each (fast[k & 1])
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
  for (int l = 0; l < GMETA_MAX_LAYERS; ++l) {
    const bool need = a->pruned_forward && l >= 1 && l < m.n_layers;
    b.aout_ptr_spt[l] = c.take<int32_t>(need ? (int64_t)a->spt.n_act[l - 1] + 1 : 0);
    b.aout_idx_spt[l] = c.take<int32_t>(need ? (int64_t)a->spt.n_edges : 0);
    b.aout_ptr_qry[l] = c.take<int32_t>(need && a->compute_meta_grad ? (int64_t)a->qry.n_act[l - 1] + 1 : 0);
    b.aout_idx_qry[l] = c.take<int32_t>(need && a->compute_meta_grad ? (int64_t)a->qry.n_edges : 0);
  }
  b.aout_count = c.take<int32_t>(a->pruned_forward ? 2 * n_max : 0);
  // full-formulation forwards with the structurally-sparse backward: the weight gradient takes the same two steps as in
  // pruned mode (chip-wide neighbourhood sums of the active rows, then the dense contraction), into this scratch
  const bool sparse_full = !a->pruned_forward && !a->dense_backward;
  if (sparse_full) {
    for (int l = 0; l < m.n_layers; ++l) {
      if (a->spt.n_act[l] > n_max) n_max = a->spt.n_act[l];
      if (a->qry.n_act[l] > n_max) n_max = a->qry.n_act[l];
    }
  }
  int64_t ld_in_max = a->ld_feat;
  for (int l = 1; l < m.n_layers; ++l)
    if (b.ld[l - 1] > ld_in_max) ld_in_max = b.ld[l - 1];
  b.agg_bwd = c.take<float>(sparse_full ? n_max * ld_in_max : 0);
  const bool ident = a->pruned_forward || sparse_full;
  b.n_ident = ident ? (int)n_max : 0;
  b.iota = c.take<int32_t>(ident ? n_max + 1 : 0);
  b.ones = c.take<float>(ident ? n_max : 0);
  b.dz_spt[0] = c.take<float>(rows_s * ld_max);
  b.dz_spt[1] = c.take<float>(m.n_layers > 1 ? rows_s * ld_max : 0);
  b.dz_qry[0] = c.take<float>(a->compute_meta_grad ? rows_q * ld_max : 0);
  b.dz_qry[1] = c.take<float>(a->compute_meta_grad && m.n_layers > 1 ? rows_q * ld_max : 0);
  b.logits_s = c.take<float>(Ss * C);
  b.logits_q = c.take<float>(Sq * C);
  b.dlogits_s = c.take<float>(Ss * C);
  b.dlogits_q = c.take<float>(Sq * C);
  b.protos[0] = c.take<float>(T * MC * C);
  b.protos[1] = c.take<float>(T * MC * C);
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
  b.layer_ws_aux = c.take<char>(a->aux_stream ? b.layer_ws_bytes : 0);
  // Operand images of the weights for the launches that take the streamed-weight tensor-core kernel: the pruned
  // mode's dense contractions (identity graph) and every data gradient.  Activations and dZ buffers are carved
  // 256-byte aligned with ld % 4 == 0, so the shape alone decides.
  b.pack.n_seg = 0;
  b.pack.img_copy_stride = 0;
  const bool tc_wanted = a->impl == GMETA_IMPL_AUTO || a->impl == GMETA_IMPL_TCGEN05;
  for (int l = 0; l < m.n_layers; ++l) {
    for (int o = 0; o < 2; ++o) {
      const int K = o ? m.f_out[l] : m.f_in[l], N = o ? m.f_in[l] : m.f_out[l];
      const int ld_in = o ? b.ld[l] : (l == 0 ? a->ld_feat : b.ld[l - 1]);
      bool use = tc_wanted && K % 32 == 0 && K >= 32 && K <= 2048 && N % 16 == 0 && N >= 16 && N <= 256 && ld_in % 4 == 0;
      if (o == 0 && !a->pruned_forward) use = false;               // full formulation: forwards go through gmeta_gcn_layer_fwd[_ex]
      if (o == 1 && (l == 0 || !a->pruned_forward)) use = false;   // features carry no gradient; unpruned: generic path
      b.tc_img[l][o] = use;
      if (use) {
        TcPackSeg& sg = b.pack.seg[b.pack.n_seg++];
        sg.w_off = m.w_off[l]; sg.K = K; sg.N = N; sg.ldw = m.f_out[l]; sg.trans = o;
        sg.img_off = b.pack.img_copy_stride;
        b.pack.img_copy_stride += 2LL * K * N;
      }
    }
  }
  b.img_theta = c.take<float>(b.pack.img_copy_stride);
  b.img_fast[0] = c.take<float>(T * b.pack.img_copy_stride);
  b.img_fast[1] = c.take<float>(T * b.pack.img_copy_stride);
  b.total = c.off;
}

// Events for the fork / join between the main and the auxiliary stream: created once per host thread, reused by
// every step (timing disabled; recording an event does not synchronise anything).
struct EventPool {
  std::vector<cudaEvent_t> ev;
  cudaEvent_t get(size_t i) {
    while (ev.size() <= i) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      ev.push_back(e);
    }
    return ev[i];
  }
};
thread_local EventPool g_events;

__global__ void step_stats_kernel(const float* __restrict__ loss_q, const float* __restrict__ acc_q, int n_tasks,
                                  int n_cols, float* __restrict__ out) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  // out[0] = sum_t loss_q[t][K], out[1 + k] = sum_t acc_q[t][k]; tasks added in index order (deterministic)
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_cols) return;
  float s = 0.f;
  if (k == 0) {
    for (int t = 0; t < n_tasks; ++t) s += loss_q[(size_t)t * n_cols + n_cols - 1];
  } else {
    for (int t = 0; t < n_tasks; ++t) s += acc_q[(size_t)t * n_cols + k - 1];
  }
  out[k] = s;
}

struct Runner {
  const gmeta_step_args_t* a;
  StepBuffers b;
  cudaStream_t s;            // stream of the launches being enqueued (main or auxiliary)
  void* ws;                  // tensor-core scratch of that stream
  int rc = GMETA_OK;

  bool ok() const { return rc == GMETA_OK; }
  // scales of a transposed pass (data gradient): the forward's destination scale acts on the sources and vice versa
  static const float* t_src(const gmeta_packed_set_t& set) { return set.norm_dst ? set.norm_dst : set.norm; }
  static const float* t_dst(const gmeta_packed_set_t& set) { return set.norm_dst ? set.norm : nullptr; }
  void run(int code) { if (rc == GMETA_OK) rc = code; }
  void on(cudaStream_t stream, void* scratch) { s = stream; ws = scratch; }

  // Weight version: -1 = theta (one shared copy), 0 / 1 = fast[v] (one copy per task)
  struct Weights {
    const float* W;
    int64_t stride;
    const float* img;        // operand images of this version, or NULL
  };
  Weights theta() const { return Weights{a->theta, 0, b.img_theta}; }
  Weights fast(int v) const { return Weights{b.fast[v], (int64_t)a->model.n_params_padded, b.img_fast[v]}; }

  // dense contraction of pre-summed rows (identity graph): out = act(in . B + bias), B = W_l or W_l^T
  void dense(const float* in, int ld_in, const int32_t* t_row0, const int32_t* t_nrows, const int32_t* t_task,
             int n_tiles, int n_tasks, const Weights& w, int l, int orient, int relu, const float* relu_mask, float* out,
             int ld_out) {
    const gmeta_model_t& m = a->model;
    const int K = orient ? m.f_out[l] : m.f_in[l], N = orient ? m.f_in[l] : m.f_out[l];
    const float* bias = orient ? nullptr : w.W + m.b_off[l];
    if (b.tc_img[l][orient] && n_tiles > 0) {
      int seg = 0;       // images are laid out in (layer, orientation) order
      for (int i = 0; i < 2 * l + orient; ++i)
        if (b.tc_img[i >> 1][i & 1]) ++seg;
      GatherSrc g;
      g.in = in; g.in_row_map = nullptr; g.dst_rows = nullptr; g.indptr = b.iota; g.indices = b.iota; g.norm = b.ones;
      g.ld_in = ld_in; g.f_in = K; g.identity = 1;
      run(gcn_layer_fwd_tc(g, t_row0, t_nrows, t_task, n_tiles, w.stride == 0 ? 1 : n_tasks, w.W + m.w_off[l], w.stride,
                           m.f_out[l], orient, bias, w.stride, N, relu, relu_mask, out, ld_out, ws, b.layer_ws_bytes,
                           w.img + b.pack.seg[seg].img_off, w.stride == 0 ? 0 : b.pack.img_copy_stride, s));
      return;
    }
    run(gmeta_gcn_layer_fwd(in, ld_in, nullptr, nullptr, b.iota, b.iota, b.ones, t_row0, t_nrows, t_task, n_tiles, n_tasks,
                            w.W + m.w_off[l], w.stride, m.f_out[l], orient, bias, w.stride, K, N, relu, relu_mask, out,
                            ld_out, a->impl, ws, b.layer_ws_bytes, s));
  }

  void forward(const gmeta_packed_set_t& set, float* const* act, const Weights& w, float* logits) {
    const gmeta_model_t& m = a->model;
    const float* W = w.W;
    const int64_t stride = w.stride;
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
          {
            NormDstScope nd(set.norm_dst);
            run(gmeta_aggregate_rows(act[l - 1], b.ld[l - 1], set.row_pos[l - 1], set.act_rows[l], set.indptr, set.indices,
                                     set.norm, set.n_act[l], m.f_in[l], 1, agg, ld_agg, s));
          }
        dense(agg, ld_agg, set.act_tile_row0[l], set.act_tile_nrows[l], set.act_tile_task[l], set.n_act_tiles[l],
              set.n_tasks, w, l, 0, 1, nullptr, act[l], b.ld[l]);
        continue;
      }
      const float* in = l == 0 ? a->feat_table : act[l - 1];
      const int ld_in = l == 0 ? a->ld_feat : b.ld[l - 1];
      const int32_t* map = l == 0 ? set.feat_row : nullptr;
      if (b.use_ex) {
        // every row of every layer (the reference's formulation): CTA-pair tensor-core path where the shape allows,
        // with the structure plan built once per step and the row abs-max handed from layer to layer
        const bool is_spt = &set == &a->spt;
        float* const* rmax = is_spt ? b.rmax_spt : b.rmax_qry;
        void* const* plan = is_spt ? b.plan_spt : b.plan_qry;
        {
          NormDstScope nd(set.norm_dst);
          run(gmeta_gcn_layer_fwd_ex(in, ld_in, map, nullptr, set.indptr, set.indices, set.norm, set.tile_row0,
                                     set.tile_nrows, set.tile_task, set.n_tiles, set.n_tasks, W + m.w_off[l], stride,
                                     m.f_out[l], 0, W + m.b_off[l], stride, m.f_in[l], m.f_out[l], 1, nullptr, act[l],
                                     b.ld[l], a->impl, ws, b.layer_ws_bytes, set.n_nodes, set.n_edges,
                                     l == 0 ? a->feat_rowmax : rmax[(l - 1) & 1], l + 1 < m.n_layers ? rmax[l & 1] : nullptr,
                                     plan[l == 0 ? 0 : 1], s));
        }
        continue;
      }
      {
        NormDstScope nd(set.norm_dst);
        run(gmeta_gcn_layer_fwd(in, ld_in, map, nullptr, set.indptr, set.indices, set.norm, set.tile_row0, set.tile_nrows,
                                set.tile_task, set.n_tiles, set.n_tasks, W + m.w_off[l], stride, m.f_out[l], 0,
                                W + m.b_off[l], stride, m.f_in[l], m.f_out[l], 1, nullptr, act[l], b.ld[l], a->impl, ws,
                                b.layer_ws_bytes, s));
      }
    }
    const int L = m.n_layers;
    run(gmeta_readout_linear_fwd(act[L - 1], b.ld[L - 1], m.f_out[L - 1], pruned ? set.centre_pos : set.centre_row,
                                 set.centres_per_subgraph, set.task_sub_ptr, set.n_tasks, set.n_subgraphs,
                                 W + m.wlin_off, stride, W + m.blin_off, stride, m.n_out, logits, s));
  }

  void backward(const gmeta_packed_set_t& set, float* const* act, float* const* dz, const Weights& w,
                const float* dlogits, float* gout) {
    const gmeta_model_t& m = a->model;
    const float* W = w.W;
    const int64_t stride = w.stride;
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
        run(gcn_layer_wgrad_impl(agg, ld_in, nullptr, nullptr, b.iota, b.iota, b.ones, set.act_task_ptr[l], set.n_tasks,
                                 dz[cur], b.ld[l], m.f_in[l], m.f_out[l], gout + m.w_off[l], P, gout + m.b_off[l], P,
                                 b.wgrad_ws, b.wgrad_ws_bytes, set.n_act[l], s, /*identity_graph=*/1));
        if (l > 0) {
          const bool is_spt = &set == &a->spt;
          {
            NormDstScope nd(t_dst(set));      // transposed pass: the two scale arrays swap roles
            run(aggregate_rows_impl(dz[cur], b.ld[l], set.row_pos[l], set.act_rows[l - 1], nullptr,
                                    (is_spt ? b.aout_idx_spt : b.aout_idx_qry)[l], t_src(set), set.n_act[l - 1], m.f_out[l], 1,
                                    b.dagg, b.ld[l], (is_spt ? b.aout_ptr_spt : b.aout_ptr_qry)[l], s));
          }
          dense(b.dagg, b.ld[l], set.act_tile_row0[l - 1], set.act_tile_nrows[l - 1], set.act_tile_task[l - 1],
                set.n_act_tiles[l - 1], set.n_tasks, w, l, 1, 2, act[l - 1], dz[cur ^ 1], b.ld[l - 1]);
          cur ^= 1;
        }
        continue;
      }
      if (sparse) {
        // dW_l = (n_v M_v)^T dZ_l over the active rows: their neighbourhood sums by the chip-wide gather (one warp per
        // row), then the dense contraction -- instead of re-gathering M inside every (row chunk, column block) item
        {
          NormDstScope nd(set.norm_dst);
          run(gmeta_aggregate_rows(in, ld_in, l == 0 ? set.feat_row : nullptr, set.act_rows[l], set.indptr, set.indices,
                                   set.norm, set.n_act[l], m.f_in[l], 1, b.agg_bwd, ld_in, s));
        }
        run(gcn_layer_wgrad_impl(b.agg_bwd, ld_in, nullptr, nullptr, b.iota, b.iota, b.ones, set.act_task_ptr[l],
                                 set.n_tasks, dz[cur], b.ld[l], m.f_in[l], m.f_out[l], gout + m.w_off[l], P,
                                 gout + m.b_off[l], P, b.wgrad_ws, b.wgrad_ws_bytes, set.n_act[l], s,
                                 /*identity_graph=*/1));
      } else {
        {
          NormDstScope nd(set.norm_dst);
          run(gcn_layer_wgrad_impl(in, ld_in, l == 0 ? set.feat_row : nullptr, nullptr, set.indptr, set.indices, set.norm,
                                   set.task_row_ptr, set.n_tasks, dz[cur], b.ld[l], m.f_in[l], m.f_out[l],
                                   gout + m.w_off[l], P, gout + m.b_off[l], P, b.wgrad_ws, b.wgrad_ws_bytes, set.n_nodes,
                                   s, 0));
        }
      }
      if (l > 0) {
        // data gradient = the forward kernel on the transposed graph with W^T, masked by the
        // ReLU of the layer below (features carry no gradient, so layer 0 stops here).  Sparse:
        // only rows with an out-edge into an active row of layer l are computed, reading the
        // compact dZ_l through row_pos[l] (inactive out-neighbours are dropped).
        {
          NormDstScope nd(t_dst(set));      // transposed pass: the two scale arrays swap roles
          run(gmeta_gcn_layer_fwd(dz[cur], b.ld[l], sparse ? set.row_pos[l] : nullptr,
                                  sparse ? set.act_rows[l - 1] : nullptr, set.t_indptr, set.t_indices, t_src(set),
                                  sparse ? set.act_tile_row0[l - 1] : set.tile_row0,
                                  sparse ? set.act_tile_nrows[l - 1] : set.tile_nrows,
                                  sparse ? set.act_tile_task[l - 1] : set.tile_task,
                                  sparse ? set.n_act_tiles[l - 1] : set.n_tiles, set.n_tasks, W + m.w_off[l], stride,
                                  m.f_out[l], 1, nullptr, 0, m.f_out[l], m.f_in[l], 0, act[l - 1], dz[cur ^ 1],
                                  b.ld[l - 1], a->impl, ws, b.layer_ws_bytes, s));
        }
        cur ^= 1;
      }
    }
  }

  // query loss / accuracy of step `k` against the prototypes of the support forward that used weight version pv
  void qry_loss(int k, int pv, bool want_grad) {
    const gmeta_packed_set_t& q = a->qry;
    // prototypes (and their count) come from the support set of the same task (meta.py:132,154)
    run(proto_loss_launch(false, b.logits_q, a->model.n_out, q.task_sub_ptr, q.n_tasks, q.class_pos, nullptr,
                          a->spt.n_classes, 0, a->max_classes, max_rows_q, a->grad_scale, b.protos[pv],
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
  carve(a, a->workspace, r.b);
  if (r.b.total > a->workspace_bytes) return GMETA_ERR_WORKSPACE;
  g_launch_count = 0;

  const gmeta_model_t& m = a->model;
  const gmeta_packed_set_t& sp = a->spt;
  const gmeta_packed_set_t& qr = a->qry;
  const int T = sp.n_tasks, K = a->update_step;
  const int64_t P = m.n_params_padded;
  StepBuffers& b = r.b;
  cudaStream_t s = (cudaStream_t)stream;
  cudaStream_t sq = a->aux_stream ? (cudaStream_t)a->aux_stream : s;
  const bool two = sq != s;
  r.on(s, b.layer_ws);
  // fork / join helpers (no-ops on one stream).  Events: 0 prologue, 1 + k support step k done, 1 + K + k query k done
  auto signal = [&](int ev, cudaStream_t from) {
    if (!two || !r.ok()) return;
    cudaEvent_t e = g_events.get((size_t)ev);
    if (!e || cudaEventRecord(e, from) != cudaSuccess) r.rc = GMETA_ERR_LAUNCH;
  };
  auto wait = [&](int ev, cudaStream_t on) {
    if (!two || !r.ok()) return;
    cudaEvent_t e = g_events.get((size_t)ev);
    if (!e || cudaStreamWaitEvent(on, e, 0) != cudaSuccess) r.rc = GMETA_ERR_LAUNCH;
  };
  // shared memory of the loss kernels is sized for the task with the most subgraphs
  r.max_rows_s = a->spt_max_rows_per_task > 0 ? a->spt_max_rows_per_task : sp.n_subgraphs;
  r.max_rows_q = a->qry_max_rows_per_task > 0 ? a->qry_max_rows_per_task : qr.n_subgraphs;

  // ---- prologue: structure-only bookkeeping and the layer-0 neighbourhood sums -- the support set's and the operand
  // images of theta on the main stream, the query set's on the auxiliary stream (forked here; its first forward
  // waits for the first support step anyway) ----
  if (cudaMemsetAsync(b.g_spt, 0, (size_t)T * P * sizeof(float), s) != cudaSuccess) return GMETA_ERR_LAUNCH;
  if (cudaMemsetAsync(b.g_qry, 0, (size_t)T * P * sizeof(float), s) != cudaSuccess) return GMETA_ERR_LAUNCH;
  signal(0, s);
  wait(0, sq);
  for (int side = 0; side < 2 && r.ok(); ++side) {
    const gmeta_packed_set_t& set = side ? qr : sp;
    cudaStream_t st = side ? sq : s;
    void* const* plan = side ? b.plan_qry : b.plan_spt;
    if (m.aggregation == GMETA_AGG_GCN && !set.norm_dst) r.run(gmeta_degree_norm(set.indptr, set.n_nodes, set.norm, st));
    else r.run(gmeta_aggregation_norms(set.indptr, set.n_nodes, m.aggregation, set.norm, set.norm_dst, st));
    r.run(gmeta_proto_label_prep(set.labels, set.task_sub_ptr, T, set.class_pos, set.class_occ, set.n_classes, st));
    if (!a->dense_backward && (side == 0 || a->compute_meta_grad || a->pruned_forward))
      for (int l = 0; l < m.n_layers; ++l)   // row -> position maps of the active-row lists (structure only)
        r.run(gmeta_build_row_pos(set.act_rows[l], set.n_act[l], set.n_nodes, set.row_pos[l], st));
    if (b.use_ex)
      for (int i = 0; i < (m.n_layers > 1 ? 2 : 1); ++i)
        r.run(gmeta_layer_plan_build(set.indptr, set.indices, set.norm, i == 0 ? set.feat_row : nullptr, nullptr,
                                     set.tile_row0, set.tile_nrows, set.tile_task, set.n_tiles, T, set.n_nodes,
                                     set.n_edges, plan[i], st));
    if (a->pruned_forward && (side == 0 || a->compute_meta_grad))
      for (int l = 1; l < m.n_layers; ++l)   // out-neighbour lists of the data gradients (structure only)
        r.run(active_out_lists_build(set.act_rows[l - 1], set.n_act[l - 1], set.t_indptr, set.t_indices, set.row_pos[l],
                                     b.aout_count + (side ? b.n_ident : 0), (side ? b.aout_ptr_qry : b.aout_ptr_spt)[l],
                                     (side ? b.aout_idx_qry : b.aout_idx_spt)[l], st));
    if (a->pruned_forward)   // layer-0 sums: features and structure only, once per meta-step
      {
        NormDstScope nd(set.norm_dst);
        r.run(gmeta_aggregate_rows(a->feat_table, a->ld_feat, set.feat_row, set.act_rows[0], set.indptr, set.indices,
                                   set.norm, set.n_act[0], m.f_in[0], 1, (side ? b.agg_qry : b.agg_spt)[0], a->ld_feat, st));
      }
  }
  if (b.n_ident > 0) r.run(fill_identity_graph(b.iota, b.ones, b.n_ident, s));
  if (b.pack.n_seg > 0)
    r.run(gcn_tc_sgd_pack(a->theta, 0, nullptr, 0.f, 1, (int)P, nullptr, b.pack, b.img_theta, s));

  // ---- K inner steps.  Main stream: support forward -> loss -> backward -> SGD (+ operand images).  The query
  // forward of step k (weights fast_k, prototypes of support forward k) follows on the auxiliary stream; the one
  // whose loss is back-propagated (k = K-1 when training) stays on the main stream. ----
  for (int k = 0; k < K && r.ok(); ++k) {
    const Runner::Weights wcur = k == 0 ? r.theta() : r.fast((k - 1) & 1);
    if (k >= 2) wait(1 + K + (k - 2), s);       // query k-2 has read fast[k & 1] and protos[k & 1]
    r.on(s, b.layer_ws);
    r.forward(sp, b.act_spt, wcur, b.logits_s);
    if (k == 0 && a->logits_spt0 && r.ok())
      if (cudaMemcpyAsync(a->logits_spt0, b.logits_s, (size_t)sp.n_subgraphs * m.n_out * sizeof(float),
                          cudaMemcpyDeviceToDevice, s) != cudaSuccess) r.rc = GMETA_ERR_LAUNCH;
    r.run(proto_loss_launch(true, b.logits_s, m.n_out, sp.task_sub_ptr, T, sp.class_pos, sp.class_occ,
                            sp.n_classes, a->n_support, a->max_classes, r.max_rows_s, 1.0f, b.protos[k & 1],
                            a->loss_s + k, b.acc_s, K, b.dlogits_s, nullptr, s));
    r.backward(sp, b.act_spt, b.dz_spt, wcur, b.dlogits_s, b.g_spt);
    // fast_k = w - lr * g (meta.py:126,151) and its operand images, one launch
    r.run(gcn_tc_sgd_pack(wcur.W, wcur.stride, b.g_spt, a->update_lr, T, (int)P, b.fast[k & 1], b.pack, b.img_fast[k & 1], s));
    signal(1 + k, s);
    const bool last_on_main = a->compute_meta_grad && k == K - 1;
    if (last_on_main) {
      wait(1 + K + (k - 1), s);                 // the auxiliary stream is done with the query buffers
      r.on(s, b.layer_ws);
    } else {
      wait(1 + k, sq);
      r.on(sq, two ? b.layer_ws_aux : b.layer_ws);
    }
    if (k == 0) {  // query loss / accuracy before the first update (meta.py:129-134)
      r.forward(qr, b.act_qry, r.theta(), b.logits_q);
      r.qry_loss(0, 0, false);
    }
    r.forward(qr, b.act_qry, r.fast(k & 1), b.logits_q);
    r.qry_loss(k + 1, k & 1, last_on_main);
    if (!last_on_main) signal(1 + K + k, sq);
  }
  r.on(s, b.layer_ws);
  if (a->compute_meta_grad && r.ok()) {
    r.backward(qr, b.act_qry, b.dz_qry, r.fast((K - 1) & 1), b.dlogits_q, b.g_qry);
    r.run(gmeta_proto_grad_to_support(b.dprotos, m.n_out, a->max_classes, sp.task_sub_ptr, T, sp.class_pos,
                                      sp.class_occ, a->n_support, sp.n_subgraphs, b.dlogits_s, s));
    r.backward(sp, b.act_spt, b.dz_spt, r.fast((K - 2) & 1), b.dlogits_s, b.g_spt);
    r.run(gmeta_sum_over_tasks(b.g_qry, b.g_spt, T, (int)P, a->meta_grad, s));
  } else {
    wait(1 + K + (K - 1), s);                   // join: every query forward has finished
  }
  if (a->step_stats && r.ok()) {
    launch_pdl(step_stats_kernel, dim3(1), dim3(64), 0, s, a->loss_q, a->acc_q, T, K + 1, a->step_stats);
    r.run(check_launch());
  }
  return r.rc;
}

// ------------------------------------------------------------------------------------------
// The meta-step as an UPDATABLE CUDA graph, for batches whose shapes change from step to step (host batches): the
// ~200 launches of a step are captured (no device work), the executable graph of the previous batch is updated in
// place with the new launch parameters (same topology: cudaGraphExecUpdate, ~0.2 ms on the host) and launched as one
// unit.  The caller prepares batch i+1 while step i runs on the device, so the per-launch cost of an eager step
// (~4 us of gap behind each of the 200 small kernels) leaves the critical path.  Capture is thread-local: packer
// threads keep issuing copies and allocations meanwhile.  `stream` of prepare must not be the legacy default stream.
// ------------------------------------------------------------------------------------------
namespace {
struct StepGraph {
  cudaGraphExec_t exec = nullptr;
  cudaEvent_t uploaded = nullptr;     // the executable graph's (updated) launch data is resident on the device
  int launches = 0;
  int updates = 0, instantiations = 0;
};
}  // namespace

extern "C" int gmeta_step_graph_create(void** handle) {
  if (!handle) return GMETA_ERR_BAD_ARG;
  *handle = new StepGraph();
  return GMETA_OK;
}

extern "C" void gmeta_step_graph_destroy(void* handle) {
  StepGraph* g = reinterpret_cast<StepGraph*>(handle);
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->uploaded) cudaEventDestroy(g->uploaded);
  delete g;
}

extern "C" int gmeta_step_graph_prepare(void* handle, const gmeta_step_args_t* args, void* stream) {
  StepGraph* g = reinterpret_cast<StepGraph*>(handle);
  if (!g || !args || !stream) return GMETA_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return GMETA_ERR_LAUNCH;
  }
  const int rc = gmeta_maml_step(args, stream);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(s, &graph);
  if (rc != GMETA_OK || ce != cudaSuccess || !graph) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    return rc != GMETA_OK ? rc : GMETA_ERR_LAUNCH;
  }
  g->launches = g_launch_count;
  if (g->exec) {
    cudaGraphExecUpdateResultInfo info;
    if (cudaGraphExecUpdate(g->exec, graph, &info) == cudaSuccess) {
      ++g->updates;
    } else {                              // topology changed (e.g. a reduction launch appeared): rebuild
      cudaGetLastError();
      cudaGraphExecDestroy(g->exec);
      g->exec = nullptr;
    }
  }
  if (!g->exec) {
    if (cudaGraphInstantiate(&g->exec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      g->exec = nullptr;
      cudaGraphDestroy(graph);
      return GMETA_ERR_LAUNCH;
    }
    ++g->instantiations;
  }
  cudaGraphDestroy(graph);
  // push the (patched) launch data to the device now, on the preparing stream, instead of in front of the launch
  if (!g->uploaded && cudaEventCreateWithFlags(&g->uploaded, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    g->uploaded = nullptr;
  }
  if (g->uploaded) {
    if (cudaGraphUpload(g->exec, s) != cudaSuccess || cudaEventRecord(g->uploaded, s) != cudaSuccess) cudaGetLastError();
  }
  return GMETA_OK;
}

extern "C" int gmeta_step_graph_launch(void* handle, void* stream) {
  StepGraph* g = reinterpret_cast<StepGraph*>(handle);
  if (!g || !g->exec) return GMETA_ERR_BAD_ARG;
  if (g->uploaded && cudaStreamWaitEvent((cudaStream_t)stream, g->uploaded, 0) != cudaSuccess) cudaGetLastError();
  if (cudaGraphLaunch(g->exec, (cudaStream_t)stream) != cudaSuccess) {
    cudaGetLastError();
    return GMETA_ERR_LAUNCH;
  }
  g_launch_count = g->launches;
  return GMETA_OK;
}

extern "C" int gmeta_step_graph_stats(void* handle, int32_t* updates, int32_t* instantiations) {
  StepGraph* g = reinterpret_cast<StepGraph*>(handle);
  if (!g) return GMETA_ERR_BAD_ARG;
  if (updates) *updates = g->updates;
  if (instantiations) *instantiations = g->instantiations;
  return GMETA_OK;
}
