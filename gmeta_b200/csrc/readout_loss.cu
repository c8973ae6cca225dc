// Centre-row readout + linear head (reference G-Meta/learner.py:159-175) and the prototype /
// centroid loss (G-Meta/meta.py:14-79), forward and backward, for every task of a packed
// meta-batch at once.  The reference moves logits to the CPU for every loss evaluation
// (meta.py:29-30,57-58); here the loss, its accuracy and its gradient never leave HBM.
#include <math.h>

#include "common.cuh"

namespace gmeta {
namespace {

__device__ __forceinline__ int find_task(const int32_t* ptr, int n_tasks, int s) {
  int lo = 0, hi = n_tasks;  // largest t with ptr[t] <= s
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (ptr[mid] <= s) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one CTA (128 threads) per subgraph: r_s staged in shared memory, one warp per output class
__global__ void readout_linear_fwd_kernel(const float* __restrict__ H, int ld_h, int hid,
                                          const int32_t* __restrict__ centre_row, int cps,
                                          const int32_t* __restrict__ task_sub_ptr, int n_tasks,
                                          const float* __restrict__ Wlin, long long w_stride,
                                          const float* __restrict__ blin, long long b_stride,
                                          int n_out, float* __restrict__ logits) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  extern __shared__ float r[];  // [cps * hid]
  const int s = blockIdx.x;
  const int t = find_task(task_sub_ptr, n_tasks, s);
  const int width = cps * hid;
  for (int k = threadIdx.x; k < width; k += blockDim.x) {
    const int half = k / hid;
    r[k] = H[(size_t)centre_row[s * cps + half] * ld_h + (k - half * hid)];
  }
  __syncthreads();
  const float* W = Wlin + t * w_stride;
  const float* b = blin + t * b_stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = warp; c < n_out; c += nw) {
    float acc = 0.f;
    for (int k = lane; k < width; k += 32) acc = fmaf(r[k], W[(size_t)c * width + k], acc);
    acc = warp_sum(acc);
    if (lane == 0) logits[(size_t)s * n_out + c] = acc + b[c];
  }
}

// grid (T, 2): blockIdx.y == 0 -> dWlin / dblin of the task (fixed summation order over its
// subgraphs); blockIdx.y == 1 -> scatter of the masked data-gradient into dZ at the centre rows.
__global__ void readout_linear_bwd_kernel(const float* __restrict__ H, int ld_h, int hid,
                                          const int32_t* __restrict__ centre_row, int cps,
                                          const int32_t* __restrict__ task_sub_ptr,
                                          const float* __restrict__ Wlin, long long w_stride,
                                          int n_out, const float* __restrict__ dlogits,
                                          float* __restrict__ dWlin, long long dw_stride,
                                          float* __restrict__ dblin, long long db_stride,
                                          const int32_t* __restrict__ row_pos, float* __restrict__ dZ) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  const int t = blockIdx.x;
  const int s0 = task_sub_ptr[t], s1 = task_sub_ptr[t + 1];
  const int width = cps * hid;
  if (blockIdx.y == 0) {
    float* dW = dWlin + t * dw_stride;
    for (int idx = threadIdx.x; idx < n_out * width; idx += blockDim.x) {
      const int c = idx / width, k = idx - c * width;
      const int half = k / hid, kk = k - half * hid;
      float acc = 0.f;
      for (int s = s0; s < s1; ++s)
        acc = fmaf(dlogits[(size_t)s * n_out + c], H[(size_t)centre_row[s * cps + half] * ld_h + kk], acc);
      dW[idx] = acc;
    }
    float* db = dblin + t * db_stride;
    for (int c = threadIdx.x; c < n_out; c += blockDim.x) {
      float acc = 0.f;
      for (int s = s0; s < s1; ++s) acc += dlogits[(size_t)s * n_out + c];
      db[c] = acc;
    }
  } else {
    const float* W = Wlin + t * w_stride;
    const int total = (s1 - s0) * width;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int s = s0 + idx / width, k = idx % width;
      const int half = k / hid, kk = k - half * hid;
      const size_t row = (size_t)centre_row[s * cps + half];
      if (H[row * ld_h + kk] > 0.f) {  // ReLU of the last GCN layer (learner.py:53-54)
        float g = 0.f;
        for (int c = 0; c < n_out; ++c) g = fmaf(dlogits[(size_t)s * n_out + c], W[(size_t)c * width + k], g);
        const size_t orow = row_pos ? (size_t)row_pos[row] : row;   // compact row of the active-row list
        atomicAdd(dZ + orow * ld_h + kk, g);  // both endpoints of a pair may name the same row
      }
    }
  }
}

// class_pos / class_occ / n_classes: one CTA per task, O(S_t^2) integer compares in shared memory
__global__ void proto_label_prep_kernel(const int32_t* __restrict__ labels,
                                        const int32_t* __restrict__ task_sub_ptr,
                                        int32_t* __restrict__ class_pos,
                                        int32_t* __restrict__ class_occ,
                                        int32_t* __restrict__ n_classes) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  const int t = blockIdx.x;
  const int s0 = task_sub_ptr[t], n = task_sub_ptr[t + 1] - s0;
  int my_first_count = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int y = labels[s0 + i];
    int occ = 0, pos = 0;
    for (int j = 0; j < n; ++j) {
      const int yj = labels[s0 + j];
      if (j < i && yj == y) ++occ;
      if (yj < y) {  // count distinct smaller labels: only the first occurrence of each
        bool first = true;
        for (int q = 0; q < j; ++q)
          if (labels[s0 + q] == yj) { first = false; break; }
        if (first) ++pos;
      }
    }
    class_pos[s0 + i] = pos;
    class_occ[s0 + i] = occ;
    if (occ == 0) ++my_first_count;
  }
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  if (my_first_count) atomicAdd(&total, my_first_count);
  __syncthreads();
  if (threadIdx.x == 0) n_classes[t] = total;
}

// Shared device routine: per query row q (thread-per-row), distances to the prototypes,
// log-softmax(-d), NLL term, argmax hit, and a[q][c] = dL/dd_qc.
__device__ __forceinline__ void proto_row(const float* z, const float* mu, int D, int ncls, int target,
                                          float inv_q, bool squeeze_quirk, float* a_row, float& nll,
                                          float& hit) {
  float dmin = INFINITY;
  for (int c = 0; c < ncls; ++c) {
    float d = 0.f;
    for (int k = 0; k < D; ++k) {
      const float diff = z[k] - mu[c * D + k];
      d = fmaf(diff, diff, d);
    }
    a_row[c] = d;
    dmin = fminf(dmin, d);
  }
  float se = 0.f;
  for (int c = 0; c < ncls; ++c) se += expf(dmin - a_row[c]);  // sum exp(-d + max(-d))
  const float lse = logf(se) - dmin;                            // logsumexp(-d)
  int best = 0;
  float best_lp = -INFINITY;
  float lp_t = 0.f;
  for (int c = 0; c < ncls; ++c) {
    const float lp = -a_row[c] - lse;
    if (lp > best_lp) { best_lp = lp; best = c; }  // first maximum, like torch.max
    if (c == target) lp_t = lp;
    a_row[c] = ((c == target ? 1.f : 0.f) - expf(lp)) * inv_q;
  }
  nll = -lp_t;
  // Accuracy as the reference computes it (meta.py:52-53,77-78): y_hat.eq(target_inds.squeeze()).
  // With ONE row per class (1-shot support / 1 query per class) squeeze() drops that axis and the
  // comparison broadcasts [n_cls,1] against [n_cls]: every row then scores 1/n_cls whatever it
  // predicts.  Reproduced so the accuracy vector stays identical to the reference's.
  if (squeeze_quirk) hit = 1.f / (float)ncls;
  else hit = (best == target) ? 1.f : 0.f;
}

// one CTA per task.  smem: z[n*D] | mu[C*D] | a[n*C] | red[n*2]
template <bool SPT>
__global__ void proto_loss_kernel(const float* __restrict__ logits, int D,
                                  const int32_t* __restrict__ task_sub_ptr,
                                  const int32_t* __restrict__ class_pos,
                                  const int32_t* __restrict__ class_occ,
                                  const int32_t* __restrict__ n_classes_spt_or_qry,
                                  int n_support, int max_classes, float grad_scale,
                                  float* __restrict__ protos_io, float* __restrict__ loss,
                                  float* __restrict__ acc, int out_stride,
                                  float* __restrict__ dlogits, float* __restrict__ dprotos) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  extern __shared__ float sm[];
  const int t = blockIdx.x;
  const int s0 = task_sub_ptr[t], n = task_sub_ptr[t + 1] - s0;
  const int ncls = SPT ? n_classes_spt_or_qry[t] : min(n_classes_spt_or_qry[t], max_classes);
  float* z = sm;
  float* mu = z + n * D;
  float* a = mu + max_classes * D;
  float* red = a + n * max_classes;
  float* P = protos_io + (size_t)t * max_classes * D;

  for (int i = threadIdx.x; i < n * D; i += blockDim.x) z[i] = logits[(size_t)s0 * D + i];
  __syncthreads();
  if (SPT) {
    // prototypes: mean of the first n_support rows of each class, rows visited in index order
    for (int i = threadIdx.x; i < ncls * D; i += blockDim.x) {
      const int c = i / D, k = i - c * D;
      float sum = 0.f;
      int cnt = 0;
      for (int s = 0; s < n; ++s)
        if (class_pos[s0 + s] == c && class_occ[s0 + s] < n_support) { sum += z[s * D + k]; ++cnt; }
      const float m = sum / (float)max(cnt, 1);
      mu[i] = m;
      P[i] = m;
    }
  } else {
    for (int i = threadIdx.x; i < ncls * D; i += blockDim.x) mu[i] = P[i];
  }
  __syncthreads();

  // rows that enter the loss: support -> the first n_support of each class; query -> all
  int n_q = 0;
  for (int s = 0; s < n; ++s) n_q += (!SPT || class_occ[s0 + s] < n_support) ? 1 : 0;
  const float inv_q = 1.f / (float)max(n_q, 1);
  const bool squeeze_quirk = ncls > 1 && (SPT ? n_support == 1 : n_q == ncls);
  for (int s = threadIdx.x; s < n; s += blockDim.x) {
    float nll = 0.f, hit = 0.f;
    const bool used = !SPT || class_occ[s0 + s] < n_support;
    if (used) {
      proto_row(z + s * D, mu, D, ncls, class_pos[s0 + s], inv_q, squeeze_quirk, a + s * max_classes, nll, hit);
    } else {
      for (int c = 0; c < ncls; ++c) a[s * max_classes + c] = 0.f;
    }
    red[2 * s] = nll;
    red[2 * s + 1] = hit;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, h = 0.f;
    for (int s = 0; s < n; ++s) { l += red[2 * s]; h += red[2 * s + 1]; }
    loss[(size_t)t * out_stride] = l / (float)max(n_q, 1);
    acc[(size_t)t * out_stride] = h / (float)max(n_q, 1);
  }
  if (dlogits == nullptr && dprotos == nullptr) return;

  // d loss / d mu[c][k] = -2 sum_s a[s][c] (z[s][k] - mu[c][k]); stored in red2 = reuse of `red`?
  // (red is too small) -> recomputed per consumer below; ncls*D and n*D are tiny.
  if (dprotos) {
    float* dP = dprotos + (size_t)t * max_classes * D;
    for (int i = threadIdx.x; i < ncls * D; i += blockDim.x) {
      const int c = i / D, k = i - c * D;
      float g = 0.f;
      for (int s = 0; s < n; ++s) g = fmaf(a[s * max_classes + c], -2.f * (z[s * D + k] - mu[i]), g);
      dP[i] = g * grad_scale;
    }
  }
  if (dlogits) {
    for (int i = threadIdx.x; i < n * D; i += blockDim.x) {
      const int s = i / D, k = i - s * D;
      float g = 0.f;
      for (int c = 0; c < ncls; ++c) g = fmaf(a[s * max_classes + c], 2.f * (z[i] - mu[c * D + k]), g);
      if (SPT && class_occ[s0 + s] < n_support) {
        // this row is also a member of its class prototype: + (1/n_support) dL/dmu[class]
        const int c = class_pos[s0 + s];
        float gm = 0.f;
        for (int q = 0; q < n; ++q) gm = fmaf(a[q * max_classes + c], -2.f * (z[q * D + k] - mu[c * D + k]), gm);
        g += gm / (float)n_support;
      }
      dlogits[(size_t)s0 * D + i] = g * grad_scale;
    }
  }
}

__global__ void proto_grad_to_support_kernel(const float* __restrict__ dprotos, int D, int max_classes,
                                             const int32_t* __restrict__ task_sub_ptr, int n_tasks,
                                             const int32_t* __restrict__ class_pos,
                                             const int32_t* __restrict__ class_occ, int n_support,
                                             int n_sub, float* __restrict__ dlogits) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_sub * D; i += gridDim.x * blockDim.x) {
    const int s = i / D, k = i - s * D;
    const int t = find_task(task_sub_ptr, n_tasks, s);
    float g = 0.f;
    if (class_occ[s] < n_support && class_pos[s] < max_classes)
      g = dprotos[((size_t)t * max_classes + class_pos[s]) * D + k] / (float)n_support;
    dlogits[i] = g;
  }
}

size_t proto_smem_bytes(int max_rows, int D, int max_classes) {
  return (size_t)(max_rows * D + max_classes * D + max_rows * max_classes + 2 * max_rows) * sizeof(float);
}

}  // namespace
}  // namespace gmeta

using namespace gmeta;

extern "C" int gmeta_readout_linear_fwd(const float* H, int32_t ld_h, int32_t hid,
                                        const int32_t* centre_row, int32_t cps,
                                        const int32_t* task_sub_ptr, int32_t n_tasks,
                                        int32_t n_subgraphs, const float* Wlin, int64_t w_task_stride,
                                        const float* blin, int64_t b_task_stride, int32_t n_out,
                                        float* logits, void* stream) {
  if (!H || !centre_row || !task_sub_ptr || !Wlin || !blin || !logits) return GMETA_ERR_BAD_ARG;
  if (hid <= 0 || ld_h < hid || (cps != 1 && cps != 2) || n_tasks <= 0 || n_out <= 0 || n_subgraphs < 0)
    return GMETA_ERR_BAD_ARG;
  if (n_subgraphs == 0) return GMETA_OK;
  const size_t smem = (size_t)cps * hid * sizeof(float);
  if (smem > 48 * 1024) return GMETA_ERR_UNSUPPORTED;
  launch_pdl(readout_linear_fwd_kernel, dim3(n_subgraphs), dim3(128), smem, (cudaStream_t)stream, 
      H, ld_h, hid, centre_row, cps, task_sub_ptr, n_tasks, Wlin, w_task_stride, blin, b_task_stride,
      n_out, logits);
  return check_launch();
}

extern "C" int gmeta_readout_linear_bwd(const float* H, int32_t ld_h, int32_t hid, int32_t n_dz_rows,
                                        const int32_t* row_pos, const int32_t* centre_row, int32_t cps,
                                        const int32_t* task_sub_ptr, int32_t n_tasks,
                                        int32_t n_subgraphs, const float* Wlin, int64_t w_task_stride,
                                        int32_t n_out, const float* dlogits, float* dWlin,
                                        int64_t dw_task_stride, float* dblin, int64_t db_task_stride,
                                        float* dZ, void* stream) {
  if (!H || !centre_row || !task_sub_ptr || !Wlin || !dlogits || !dWlin || !dblin || !dZ)
    return GMETA_ERR_BAD_ARG;
  if (hid <= 0 || ld_h < hid || (cps != 1 && cps != 2) || n_tasks <= 0 || n_out <= 0 || n_dz_rows < 0)
    return GMETA_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(dZ, 0, (size_t)n_dz_rows * ld_h * sizeof(float), s) != cudaSuccess) return GMETA_ERR_LAUNCH;
  launch_pdl(readout_linear_bwd_kernel, dim3(dim3(n_tasks, 2)), dim3(256), 0, s, H, ld_h, hid, centre_row, cps, task_sub_ptr,
                                                            Wlin, w_task_stride, n_out, dlogits, dWlin,
                                                            dw_task_stride, dblin, db_task_stride, row_pos, dZ);
  return check_launch();
}

namespace gmeta {
namespace {
__global__ void scatter_row_pos_kernel(const int32_t* __restrict__ rows, int n_rows, int32_t* __restrict__ row_pos) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += gridDim.x * blockDim.x) row_pos[rows[i]] = i;
}
}  // namespace
}  // namespace gmeta

extern "C" int gmeta_build_row_pos(const int32_t* rows, int32_t n_rows, int32_t n_nodes, int32_t* row_pos,
                                   void* stream) {
  if (n_rows < 0 || n_nodes < 0 || (n_nodes > 0 && !row_pos) || (n_rows > 0 && !rows)) return GMETA_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_nodes == 0) return GMETA_OK;
  if (cudaMemsetAsync(row_pos, 0xFF, (size_t)n_nodes * sizeof(int32_t), s) != cudaSuccess) return GMETA_ERR_LAUNCH;
  if (n_rows == 0) return GMETA_OK;
  launch_pdl(scatter_row_pos_kernel, dim3(ceil_div(n_rows, 256)), dim3(256), 0, s, rows, n_rows, row_pos);
  return check_launch();
}

extern "C" int gmeta_proto_label_prep(const int32_t* labels, const int32_t* task_sub_ptr, int32_t n_tasks,
                                      int32_t* class_pos, int32_t* class_occ, int32_t* n_classes,
                                      void* stream) {
  if (!labels || !task_sub_ptr || !class_pos || !class_occ || !n_classes || n_tasks <= 0)
    return GMETA_ERR_BAD_ARG;
  launch_pdl(proto_label_prep_kernel, dim3(n_tasks), dim3(128), 0, (cudaStream_t)stream, labels, task_sub_ptr, class_pos,
                                                                   class_occ, n_classes);
  return check_launch();
}

// The per-task row count is only known on the device; shared memory is sized for the
// largest task the caller can have: max_rows_per_task is passed through n_support's sibling
// below (host knows S and T; it passes the max over tasks).
static int launch_proto(bool spt, const float* logits, int n_out, const int32_t* task_sub_ptr, int n_tasks,
                        const int32_t* class_pos, const int32_t* class_occ, const int32_t* n_classes,
                        int n_support, int max_classes, int max_rows, float grad_scale, float* protos,
                        float* loss, float* acc, int out_stride, float* dlogits, float* dprotos,
                        cudaStream_t s) {
  const size_t smem = proto_smem_bytes(max_rows, n_out, max_classes);
  if (smem > 200 * 1024) return GMETA_ERR_UNSUPPORTED;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(proto_loss_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(proto_loss_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  if (spt)
    launch_pdl(proto_loss_kernel<true>, dim3(n_tasks), dim3(128), smem, s, logits, n_out, task_sub_ptr, class_pos, class_occ,
                                                      n_classes, n_support, max_classes, grad_scale,
                                                      protos, loss, acc, out_stride, dlogits, dprotos);
  else
    launch_pdl(proto_loss_kernel<false>, dim3(n_tasks), dim3(128), smem, s, logits, n_out, task_sub_ptr, class_pos, class_occ,
                                                       n_classes, n_support, max_classes, grad_scale,
                                                       protos, loss, acc, out_stride, dlogits, dprotos);
  return check_launch();
}

namespace gmeta {
int proto_loss_launch(bool spt, const float* logits, int n_out, const int32_t* task_sub_ptr, int n_tasks,
                      const int32_t* class_pos, const int32_t* class_occ, const int32_t* n_classes,
                      int n_support, int max_classes, int max_rows, float grad_scale, float* protos,
                      float* loss, float* acc, int out_stride, float* dlogits, float* dprotos,
                      cudaStream_t s) {
  return launch_proto(spt, logits, n_out, task_sub_ptr, n_tasks, class_pos, class_occ, n_classes,
                      n_support, max_classes, max_rows, grad_scale, protos, loss, acc, out_stride,
                      dlogits, dprotos, s);
}
}  // namespace gmeta

extern "C" int gmeta_proto_loss_spt(const float* logits, int32_t n_out, const int32_t* task_sub_ptr,
                                    int32_t n_tasks, const int32_t* class_pos, const int32_t* class_occ,
                                    const int32_t* n_classes, int32_t n_support, int32_t max_classes,
                                    int32_t max_rows_per_task, float grad_scale, float* protos,
                                    float* loss, float* acc, int32_t out_stride, float* dlogits,
                                    void* stream) {
  if (!logits || !task_sub_ptr || !class_pos || !class_occ || !n_classes || !protos || !loss || !acc)
    return GMETA_ERR_BAD_ARG;
  if (n_out <= 0 || n_tasks <= 0 || n_support <= 0 || max_classes <= 0 || max_rows_per_task <= 0)
    return GMETA_ERR_BAD_ARG;
  return launch_proto(true, logits, n_out, task_sub_ptr, n_tasks, class_pos, class_occ, n_classes,
                      n_support, max_classes, max_rows_per_task, grad_scale, protos, loss, acc,
                      out_stride, dlogits, nullptr, (cudaStream_t)stream);
}

extern "C" int gmeta_proto_loss_qry(const float* logits, int32_t n_out, const int32_t* task_sub_ptr,
                                    int32_t n_tasks, const int32_t* class_pos, const int32_t* n_classes,
                                    const float* protos, int32_t max_classes, int32_t max_rows_per_task,
                                    float grad_scale, float* loss, float* acc, int32_t out_stride,
                                    float* dlogits, float* dprotos, void* stream) {
  if (!logits || !task_sub_ptr || !class_pos || !n_classes || !protos || !loss || !acc)
    return GMETA_ERR_BAD_ARG;
  if (n_out <= 0 || n_tasks <= 0 || max_classes <= 0 || max_rows_per_task <= 0) return GMETA_ERR_BAD_ARG;
  return launch_proto(false, logits, n_out, task_sub_ptr, n_tasks, class_pos, nullptr, n_classes, 0,
                      max_classes, max_rows_per_task, grad_scale, const_cast<float*>(protos), loss, acc,
                      out_stride, dlogits, dprotos, (cudaStream_t)stream);
}

extern "C" int gmeta_proto_grad_to_support(const float* dprotos, int32_t n_out, int32_t max_classes,
                                           const int32_t* task_sub_ptr, int32_t n_tasks,
                                           const int32_t* class_pos, const int32_t* class_occ,
                                           int32_t n_support, int32_t n_subgraphs, float* dlogits_spt,
                                           void* stream) {
  if (!dprotos || !task_sub_ptr || !class_pos || !class_occ || !dlogits_spt) return GMETA_ERR_BAD_ARG;
  if (n_out <= 0 || max_classes <= 0 || n_tasks <= 0 || n_support <= 0 || n_subgraphs < 0)
    return GMETA_ERR_BAD_ARG;
  if (n_subgraphs == 0) return GMETA_OK;
  const int total = n_subgraphs * n_out;
  launch_pdl(proto_grad_to_support_kernel, dim3(ceil_div(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
      dprotos, n_out, max_classes, task_sub_ptr, n_tasks, class_pos, class_occ, n_support, n_subgraphs,
      dlogits_spt);
  return check_launch();
}
