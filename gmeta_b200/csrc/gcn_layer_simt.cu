// Fused GCN layer over a packed meta-batch of local subgraphs -- fp32 FFMA implementation
// (any shape; the tcgen05 3xTF32 implementation in gcn_layer_tc.cu covers the wide shapes).
//
// Replaces GraphConv.forward (reference G-Meta/learner.py:25-56): in-degree normalisation,
// CSR neighbour gather + sum aggregation, feature x weight contraction, bias, ReLU -- one
// kernel, per-task weights selected per row tile.  Also the layer's weight gradient
// (autograd.grad at meta.py:125,149 restricted to this layer), which re-gathers the
// aggregated rows instead of storing them.
#include "common.cuh"

namespace gmeta {

thread_local int g_launch_count = 0;

namespace {

constexpr int TM = GMETA_TILE_ROWS;  // rows per tile
constexpr int BN = 128;              // output columns per work item
constexpr int KP = 128;              // K panel staged in shared memory
constexpr int KC = 16;               // K chunk of the weight staging
constexpr int LDA = KP + 4;
constexpr int LDB = BN + 4;
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
constexpr size_t kFwdSmem = (size_t)(TM * LDA + 2 * KC * LDB) * sizeof(float);

struct FwdParams {
  GatherSrc g;
  const int32_t* tile_row0;
  const int32_t* tile_nrows;
  const int32_t* tile_task;
  int n_tiles;
  const float* W;
  long long w_task_stride;
  int ldw;
  int trans_w;
  const float* bias;
  long long b_task_stride;
  int f_out;
  int relu;
  const float* relu_mask;
  float* out;
  int ld_out;
  int vec_out;  // 16-byte stores allowed
};

// B[k][j] of the contraction, bounds-masked.
__device__ __forceinline__ float load_w(const float* W, int ldw, int trans, int k, int j, int f_in,
                                        int f_out) {
  if (k >= f_in || j >= f_out) return 0.f;
  return trans ? W[(size_t)j * ldw + k] : W[(size_t)k * ldw + j];
}

template <bool VEC>
__global__ void __launch_bounds__(NTHREADS, 2) gcn_layer_fwd_simt_kernel(const FwdParams p) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                // [TM][LDA]   aggregated rows, one K panel
  float* Bs = smem + TM * LDA;     // [2][KC][LDB] weight chunk, double buffered
  const int tid = threadIdx.x;
  const int tx = tid & 15;         // 8 output columns: 4*tx.. and 64+4*tx..
  const int ty = tid >> 4;         // 8 output rows:    ty + 16*i
  const int f_in = p.g.f_in, f_out = p.f_out;
  const int n_cb = (f_out + BN - 1) / BN;
  const int n_kp = (f_in + KP - 1) / KP;
  const int n_work = p.n_tiles * n_cb;

  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
    const int tile = work / n_cb, cb = work - tile * n_cb;
    const int row0 = p.tile_row0[tile], nrows = p.tile_nrows[tile], task = p.tile_task[tile];
    const float* W = p.W + (long long)task * p.w_task_stride;
    const int col0 = cb * BN;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int kp = 0; kp < n_kp; ++kp) {
      const int k0 = kp * KP;
      const int klen = min(KP, f_in - k0);
      const int n_kc = (klen + KC - 1) / KC;
      __syncthreads();  // previous panel / work item finished reading As and Bs
      gather_rows<VEC, TM, NWARPS, false>(p.g, row0, nrows, k0, As, LDA);

      // weight chunk: 16 x 128 values, 8 per thread, staged through registers
      float wreg[8];
      auto fetch_w = [&](int kc) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int e = tid + NTHREADS * i;
          int kk, jj;
          if (p.trans_w) { jj = e >> 4; kk = e & 15; } else { kk = e >> 7; jj = e & 127; }
          wreg[i] = load_w(W, p.ldw, p.trans_w, k0 + kc * KC + kk, col0 + jj, f_in, f_out);
        }
      };
      auto stash_w = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int e = tid + NTHREADS * i;
          int kk, jj;
          if (p.trans_w) { jj = e >> 4; kk = e & 15; } else { kk = e >> 7; jj = e & 127; }
          Bs[(buf * KC + kk) * LDB + jj] = wreg[i];
        }
      };
      fetch_w(0);
      stash_w(0);
      __syncthreads();  // As panel and Bs[0] visible
      for (int kc = 0; kc < n_kc; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < n_kc) fetch_w(kc + 1);
        const float* Ab = As + kc * KC;
        const float* Bb = Bs + buf * KC * LDB;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
          float4 a[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = ld_f4(Ab + (ty + 16 * i) * LDA + kk);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b0 = ld_f4(Bb + (kk + q) * LDB + 4 * tx);
            const float4 b1 = ld_f4(Bb + (kk + q) * LDB + 64 + 4 * tx);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float av = q == 0 ? a[i].x : q == 1 ? a[i].y : q == 2 ? a[i].z : a[i].w;
              acc[i][0] = fmaf(av, b0.x, acc[i][0]);
              acc[i][1] = fmaf(av, b0.y, acc[i][1]);
              acc[i][2] = fmaf(av, b0.z, acc[i][2]);
              acc[i][3] = fmaf(av, b0.w, acc[i][3]);
              acc[i][4] = fmaf(av, b1.x, acc[i][4]);
              acc[i][5] = fmaf(av, b1.y, acc[i][5]);
              acc[i][6] = fmaf(av, b1.z, acc[i][6]);
              acc[i][7] = fmaf(av, b1.w, acc[i][7]);
            }
          }
        }
        if (kc + 1 < n_kc) stash_w(buf ^ 1);
        __syncthreads();
      }
    }

    // epilogue: norm[v] * acc + bias, ReLU / mask, store
    const float* bias = p.bias ? p.bias + (long long)task * p.b_task_stride : nullptr;
    const int f_out4 = min((f_out + 3) & ~3, p.ld_out);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = ty + 16 * i;
      if (r >= nrows) continue;
      const int orow = row0 + r;                                     // output row (compact or dense)
      const int v = p.g.dst_rows ? p.g.dst_rows[orow] : orow;        // real row: norm and mask
      const float nv = dst_norm(p.g, v);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = col0 + 64 * h + 4 * tx;
        if (c >= f_out4) continue;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float val = 0.f;
          if (c + j < f_out) {
            val = nv * acc[i][4 * h + j];
            if (bias) val += bias[c + j];
            if (p.relu & 1) val = relu_keep_nan(val);
            if (p.relu_mask && !(p.relu_mask[(size_t)((p.relu & 2) ? orow : v) * p.ld_out + c + j] > 0.f)) val = 0.f;
          }
          o[j] = val;
        }
        float* dst = p.out + (size_t)orow * p.ld_out + c;
        if (p.vec_out && c + 4 <= f_out4) {
          st_f4(dst, make_float4(o[0], o[1], o[2], o[3]));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < f_out4) dst[j] = o[j];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------
constexpr int WG_ROWS = 32;  // rows per staged chunk
constexpr int WG_MAX_SPLIT = 16;
constexpr size_t kWgSmem = (size_t)(2 * WG_ROWS * LDA) * sizeof(float);

struct WgradParams {
  GatherSrc g;
  const int32_t* task_row_ptr;
  int n_tasks;
  const float* dZ;
  int ld_dz;
  int f_out;
  int n_split;
  int vec_dz;
  float* part_w;  // [T][n_split][f_in][f_out], or dW itself when n_split == 1
  float* part_b;  // [T][n_split][f_out], or db itself
  long long pw_task_stride, pw_split_stride, pb_task_stride, pb_split_stride;
};

// DENSE (p.g.identity, 16-byte aligned rows, f_in % 4 == 0): the rows of `in` ARE the aggregated rows -- a chunk is
// loaded with plain coalesced 16-byte loads, all of them in flight at once, and the next chunk's loads are issued
// before the current chunk is multiplied (registers as the second buffer).  The gather path walks indptr -> indices ->
// norm -> row per row, four rows in sequence per warp: ~12 us of dependent latency per 32-row chunk against ~2 us of
// arithmetic, which is what the short support-set launches of the pruned meta-step consisted of.
template <bool VEC, bool DENSE>
__global__ void __launch_bounds__(NTHREADS, 2) gcn_layer_wgrad_simt_kernel(const WgradParams p) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                  // [32][LDA]  norm[v] * M[v, kblock]
  float* Bs = smem + WG_ROWS * LDA;  // [32][LDA]  dZ[v, jblock]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int f_in = p.g.f_in, f_out = p.f_out;
  const int n_kb = (f_in + KP - 1) / KP, n_jb = (f_out + BN - 1) / BN;
  const int n_work = p.n_tasks * n_kb * n_jb * p.n_split;

  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
    int w = work;
    const int jb = w % n_jb; w /= n_jb;
    const int kb = w % n_kb; w /= n_kb;
    const int sp = w % p.n_split;
    const int task = w / p.n_split;
    const int rs = p.task_row_ptr[task], re = p.task_row_ptr[task + 1];
    const int n_chunks = (re - rs + WG_ROWS - 1) / WG_ROWS;
    const int c_beg = (int)((long long)n_chunks * sp / p.n_split);
    const int c_end = (int)((long long)n_chunks * (sp + 1) / p.n_split);
    const int k0 = kb * KP, j0 = jb * BN;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float bsum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bsum[j] = 0.f;

    auto multiply = [&]() {
#pragma unroll 4
      for (int r = 0; r < WG_ROWS; ++r) {
        const float4 a0 = ld_f4(As + r * LDA + 4 * ty);
        const float4 a1 = ld_f4(As + r * LDA + 64 + 4 * ty);
        const float4 b0 = ld_f4(Bs + r * LDA + 4 * tx);
        const float4 b1 = ld_f4(Bs + r * LDA + 64 + 4 * tx);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        if (ty == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bsum[j] += b[j];
        }
      }
    };
    // dZ chunk: 32 rows x 128 columns; thread -> (row = tid/32 + 8*i, 4 columns at 4*(tid%32))
    auto load_dz = [&](int row0, int nrows, int i) {
      const int r = (tid >> 5) + NWARPS * i;
      const int cc = j0 + 4 * (tid & 31);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows) {
        const float* src = p.dZ + (size_t)(row0 + r) * p.ld_dz + cc;
        if (p.vec_dz && cc + 4 <= f_out) {
          v = ld_f4(src);
        } else {
          if (cc + 0 < f_out) v.x = src[0];
          if (cc + 1 < f_out) v.y = src[1];
          if (cc + 2 < f_out) v.z = src[2];
          if (cc + 3 < f_out) v.w = src[3];
        }
      }
      return v;
    };
    if constexpr (DENSE) {
      float4 ra[WG_ROWS / NWARPS], rb[WG_ROWS / NWARPS];
      auto load_chunk = [&](int c) {
        const int row0 = rs + c * WG_ROWS;
        const int nrows = min(WG_ROWS, re - row0);
        const int kc = k0 + 4 * (tid & 31);
#pragma unroll
        for (int i = 0; i < WG_ROWS / NWARPS; ++i) {
          const int r = (tid >> 5) + NWARPS * i;
          ra[i] = (r < nrows && kc < f_in) ? ld_f4(p.g.in + (size_t)(row0 + r) * p.g.ld_in + kc) : make_float4(0.f, 0.f, 0.f, 0.f);
          rb[i] = load_dz(row0, nrows, i);
        }
      };
      if (c_beg < c_end) load_chunk(c_beg);
      for (int c = c_beg; c < c_end; ++c) {
        __syncthreads();                       // the previous chunk has been multiplied
#pragma unroll
        for (int i = 0; i < WG_ROWS / NWARPS; ++i) {
          const int r = (tid >> 5) + NWARPS * i;
          st_f4(As + r * LDA + 4 * (tid & 31), ra[i]);
          st_f4(Bs + r * LDA + 4 * (tid & 31), rb[i]);
        }
        __syncthreads();
        if (c + 1 < c_end) load_chunk(c + 1);  // in flight while this chunk is multiplied
        multiply();
      }
    } else {
      for (int c = c_beg; c < c_end; ++c) {
        const int row0 = rs + c * WG_ROWS;
        const int nrows = min(WG_ROWS, re - row0);
        __syncthreads();
        gather_rows<VEC, WG_ROWS, NWARPS, true>(p.g, row0, nrows, k0, As, LDA);
#pragma unroll
        for (int i = 0; i < WG_ROWS / NWARPS; ++i)
          st_f4(Bs + ((tid >> 5) + NWARPS * i) * LDA + 4 * (tid & 31), load_dz(row0, nrows, i));
        __syncthreads();
        multiply();
      }
    }

    float* pw = p.part_w + task * p.pw_task_stride + sp * p.pw_split_stride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + (i - 4));
      if (k >= f_in) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int jj = j0 + (j < 4 ? 4 * tx + j : 64 + 4 * tx + (j - 4));
        if (jj < f_out) pw[(size_t)k * f_out + jj] = acc[i][j];
      }
    }
    if (kb == 0 && ty == 0) {
      float* pb = p.part_b + task * p.pb_task_stride + sp * p.pb_split_stride;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int jj = j0 + (j < 4 ? 4 * tx + j : 64 + 4 * tx + (j - 4));
        if (jj < f_out) pb[jj] = bsum[j];
      }
    }
  }
}

// dW[t] = sum_sp part_w[t][sp], db[t] = sum_sp part_b[t][sp]  (fixed order)
__global__ void wgrad_reduce_kernel(const float* part_w, const float* part_b, int n_tasks,
                                    int n_split, int n_w, int f_out, float* dW,
                                    long long dw_stride, float* db, long long db_stride) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  const long long total = (long long)n_tasks * (n_w + f_out);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / (n_w + f_out));
    const int e = (int)(i - (long long)t * (n_w + f_out));
    float s = 0.f;
    if (e < n_w) {
      const float* src = part_w + (size_t)t * n_split * n_w + e;
      for (int sp = 0; sp < n_split; ++sp) s += src[(size_t)sp * n_w];
      dW[t * dw_stride + e] = s;
    } else {
      const int j = e - n_w;
      const float* src = part_b + (size_t)t * n_split * f_out + j;
      for (int sp = 0; sp < n_split; ++sp) s += src[(size_t)sp * f_out];
      db[t * db_stride + j] = s;
    }
  }
}

__global__ void degree_norm_kernel(const int32_t* indptr, int n, float* norm) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    const int d = max(indptr[v + 1] - indptr[v], 1);
    norm[v] = __fdiv_rn(1.0f, __fsqrt_rn((float)d));  // == torch CPU pow(x, -0.5) -> rsqrt path
  }
}

int pick_wgrad_split(int n_tasks, int f_in, int f_out) {
  const int base = n_tasks * ceil_div(f_in, KP) * ceil_div(f_out, BN);
  int s = ceil_div(4 * kNumSMs, base > 0 ? base : 1);
  if (s < 1) s = 1;
  if (s > WG_MAX_SPLIT) s = WG_MAX_SPLIT;
  return s;
}

}  // namespace

int gcn_layer_fwd_simt(const GatherSrc& g, const int32_t* tile_row0, const int32_t* tile_nrows,
                       const int32_t* tile_task, int n_tiles, const float* W,
                       int64_t w_task_stride, int ldw, int trans_w, const float* bias,
                       int64_t b_task_stride, int f_out, int relu, const float* relu_mask,
                       float* out, int ld_out, cudaStream_t stream) {
  FwdParams p;
  p.g = g;
  p.tile_row0 = tile_row0; p.tile_nrows = tile_nrows; p.tile_task = tile_task; p.n_tiles = n_tiles;
  p.W = W; p.w_task_stride = w_task_stride; p.ldw = ldw; p.trans_w = trans_w;
  p.bias = bias; p.b_task_stride = b_task_stride; p.f_out = f_out; p.relu = relu;
  p.relu_mask = relu_mask; p.out = out; p.ld_out = ld_out;
  p.vec_out = (ld_out % 4 == 0) && aligned16(out);
  const bool vec_in = (g.ld_in % 4 == 0) && aligned16(g.in) && g.ld_in >= round_up(g.f_in, 4);
  const int n_work = n_tiles * ceil_div(f_out, BN);
  const int grid = n_work < 16 * kNumSMs ? n_work : 16 * kNumSMs;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gcn_layer_fwd_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
    cudaFuncSetAttribute(gcn_layer_fwd_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
    attr_done = true;
  }
  if (vec_in)
    launch_pdl(gcn_layer_fwd_simt_kernel<true>, dim3(grid), dim3(NTHREADS), kFwdSmem, stream, p);
  else
    launch_pdl(gcn_layer_fwd_simt_kernel<false>, dim3(grid), dim3(NTHREADS), kFwdSmem, stream, p);
  return check_launch();
}

}  // namespace gmeta

using namespace gmeta;

namespace gmeta {
namespace {
// out[i, c] = (scale_dst ? norm[v] : 1) * sum_e norm[u_e] * in[map(u_e), c], v = dst_rows ? dst_rows[i] : i.
// A CTA (8 warps) takes 8 consecutive rows.  Rows with at most AGG_LONG in-edges: one warp per row, edge records
// loaded coalesced (one per lane) and broadcast, eight row segments in flight per lane.  Longer rows (the hubs of
// a 2-hop subgraph: up to ~1000 in-edges) would serialise one warp for tens of microseconds, so they are
// summed afterwards by the whole CTA -- warp w takes the 32-record blocks w, w+8, ... and the eight partials are
// added in warp order.  Summation order is fixed either way (deterministic).  Columns [f_in, ld_out) are zeroed.
constexpr int AGG_LONG = 64;

// acc += sum over the 32-record blocks b0, b0 + stride, ... < end of norm[u] * in[map(u), kcol .. kcol+3]
template <bool VEC>
__device__ __forceinline__ void agg_accumulate(const GatherSrc& g, int b0, int end, int stride, int kcol, int lane,
                                               float4& acc) {
  for (int base = b0; base < end; base += stride) {
    int u_src = 0;
    float u_norm = 0.f;
    if (base + lane < end) {
      const int u = g.indices[base + lane];
      u_norm = g.norm[u];
      u_src = g.in_row_map ? g.in_row_map[u] : u;
      if (u_src < 0) { u_src = 0; u_norm = 0.f; }
    }
    const int cnt = min(32, end - base);
    for (int j0 = 0; j0 < cnt; j0 += 8) {
      float4 x[8];
      float w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int src = __shfl_sync(0xffffffffu, u_src, (j0 + j) & 31);
        w[j] = __shfl_sync(0xffffffffu, u_norm, (j0 + j) & 31);
        if (j0 + j >= cnt) w[j] = 0.f;
        x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w[j] != 0.f && kcol < g.f_in) {
          const float* p = g.in + (size_t)src * g.ld_in + kcol;
          if (VEC) {
            x[j] = ld_f4(p);
          } else {
            x[j].x = p[0];
            if (kcol + 1 < g.f_in) x[j].y = p[1];
            if (kcol + 2 < g.f_in) x[j].z = p[2];
            if (kcol + 3 < g.f_in) x[j].w = p[3];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc.x = fmaf(w[j], x[j].x, acc.x);
        acc.y = fmaf(w[j], x[j].y, acc.y);
        acc.z = fmaf(w[j], x[j].z, acc.z);
        acc.w = fmaf(w[j], x[j].w, acc.w);
      }
    }
  }
}

__device__ __forceinline__ void agg_store(const GatherSrc& g, float* __restrict__ out, int ld_out, int i, int kcol,
                                          float4 acc, float nv) {
  if (kcol >= ld_out) return;
  const float r[4] = {acc.x * nv, acc.y * nv, acc.z * nv, acc.w * nv};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (kcol + k < ld_out) out[(size_t)i * ld_out + kcol + k] = kcol + k < g.f_in ? r[k] : 0.f;
}

template <bool VEC>
__global__ void __launch_bounds__(256) aggregate_rows_kernel(GatherSrc g, int n_rows, int scale_dst,
                                                             float* __restrict__ out, int ld_out,
                                                             const int32_t* __restrict__ pos_ptr) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  __shared__ int long_rows[8];
  __shared__ int n_long;
  __shared__ float4 part[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int row0 = blockIdx.x * 8; row0 < n_rows; row0 += gridDim.x * 8) {
    if (threadIdx.x == 0) n_long = 0;
    __syncthreads();
    const int i = row0 + warp;
    if (i < n_rows) {
      const int v = g.dst_rows ? g.dst_rows[i] : i;
      const int beg = pos_ptr ? pos_ptr[i] : g.indptr[v], end = pos_ptr ? pos_ptr[i + 1] : g.indptr[v + 1];
      if (end - beg > AGG_LONG) {
        if (lane == 0) long_rows[atomicAdd(&n_long, 1)] = i;     // the order of this list does not affect any value
      } else {
        const float nv = scale_dst ? dst_norm(g, v) : 1.f;
        for (int c0 = 0; c0 < ld_out; c0 += 128) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          agg_accumulate<VEC>(g, beg, end, 32, c0 + 4 * lane, lane, acc);
          agg_store(g, out, ld_out, i, c0 + 4 * lane, acc, nv);
        }
      }
    }
    __syncthreads();
    const int nl = n_long;
    for (int q = 0; q < nl; ++q) {
      const int il = long_rows[q];
      const int v = g.dst_rows ? g.dst_rows[il] : il;
      const int beg = pos_ptr ? pos_ptr[il] : g.indptr[v], end = pos_ptr ? pos_ptr[il + 1] : g.indptr[v + 1];
      const float nv = scale_dst ? dst_norm(g, v) : 1.f;
      for (int c0 = 0; c0 < ld_out; c0 += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        agg_accumulate<VEC>(g, beg + 32 * warp, end, 256, c0 + 4 * lane, lane, acc);
        part[warp][lane] = acc;
        __syncthreads();
        if (warp == 0) {
          float4 t = part[0][lane];
#pragma unroll
          for (int w = 1; w < 8; ++w) {
            const float4 p = part[w][lane];
            t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
          }
          agg_store(g, out, ld_out, il, c0 + 4 * lane, t, nv);
        }
        __syncthreads();
      }
    }
  }
}

__global__ void identity_graph_kernel(int32_t* __restrict__ iota, float* __restrict__ ones, int n) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
    iota[i] = i;
    if (i < n) ones[i] = 1.f;
  }
}
}  // namespace

// iota[0..n] = 0..n and ones[0..n) = 1: the CSR (indptr == indices == iota) and degree norm of a graph
// whose every row has exactly itself as in-neighbour with weight 1
int fill_identity_graph(int32_t* iota, float* ones, int n, cudaStream_t stream) {
  if (n < 0) return GMETA_ERR_BAD_ARG;
  const int grid = ceil_div(n + 1, 256) < 4 * kNumSMs ? ceil_div(n + 1, 256) : 4 * kNumSMs;
  launch_pdl(identity_graph_kernel, dim3(grid), dim3(256), 0, stream, iota, ones, n);
  return check_launch();
}
}  // namespace gmeta

namespace gmeta {
// pos_indptr != NULL: the in-neighbour list of output row i is indices[pos_indptr[i] .. pos_indptr[i+1]) (a CSR indexed
// by output POSITION, e.g. the active out-neighbour lists of active_out_lists_build) instead of the graph row's.
int aggregate_rows_impl(const float* in, int32_t ld_in, const int32_t* in_row_map, const int32_t* dst_rows,
                        const int32_t* indptr, const int32_t* indices, const float* norm, int32_t n_rows,
                        int32_t f_in, int32_t scale_dst, float* out, int32_t ld_out, const int32_t* pos_indptr,
                        cudaStream_t stream) {
  if (n_rows < 0 || f_in <= 0 || ld_in < f_in || ld_out < f_in) return GMETA_ERR_BAD_ARG;
  if (n_rows == 0) return GMETA_OK;
  if (!in || !(indptr || pos_indptr) || !indices || !norm || !out) return GMETA_ERR_BAD_ARG;
  GatherSrc g;
  g.in = in; g.in_row_map = in_row_map; g.dst_rows = dst_rows; g.indptr = indptr; g.indices = indices; g.norm = norm;
  g.ld_in = ld_in; g.f_in = f_in; g.norm_dst = g_norm_dst;
  const int grid = ceil_div(n_rows, 8) < 16 * kNumSMs ? ceil_div(n_rows, 8) : 16 * kNumSMs;
  if (ld_in % 4 == 0 && f_in % 4 == 0 && aligned16(in))
    launch_pdl(aggregate_rows_kernel<true>, dim3(grid), dim3(256), 0, stream, g, n_rows, scale_dst, out, ld_out, pos_indptr);
  else
    launch_pdl(aggregate_rows_kernel<false>, dim3(grid), dim3(256), 0, stream, g, n_rows, scale_dst, out, ld_out, pos_indptr);
  return check_launch();
}

namespace {
// One warp per listed row u = rows[i]: its out-neighbours v (CSR by source) with keep[v] >= 0, counted (fill == 0:
// count[i]) or written in CSR order to out_idx[ptr[i] ..) (fill != 0).
__global__ void active_out_kernel(const int32_t* __restrict__ rows, int n_rows, const int32_t* __restrict__ t_indptr,
                                  const int32_t* __restrict__ t_indices, const int32_t* __restrict__ keep, int fill,
                                  int32_t* __restrict__ count, const int32_t* __restrict__ ptr,
                                  int32_t* __restrict__ out_idx) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  const int lane = threadIdx.x & 31;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_rows; i += (gridDim.x * blockDim.x) >> 5) {
    const int u = rows[i];
    const int beg = t_indptr[u], end = t_indptr[u + 1];
    int n = 0;
    const int base = fill ? ptr[i] : 0;
    for (int e0 = beg; e0 < end; e0 += 128) {          // four independent 32-edge batches in flight
      int v[4];
      bool k[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = e0 + 32 * q + lane;
        v[q] = e < end ? t_indices[e] : -1;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) k[q] = v[q] >= 0 && keep[v[q]] >= 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned bal = __ballot_sync(0xffffffffu, k[q]);
        if (fill && k[q]) out_idx[base + n + __popc(bal & ((1u << lane) - 1u))] = v[q];
        n += __popc(bal);
      }
    }
    if (!fill && lane == 0) count[i] = n;
  }
}

// exclusive prefix sum of a[0..n) into out[0..n], out[n] = total; one CTA
__global__ void __launch_bounds__(1024) excl_scan_kernel(const int32_t* __restrict__ a, int n, int32_t* __restrict__ out) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  __shared__ int sh[1024];
  int carry = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int x = i < n ? a[i] : 0;
    sh[threadIdx.x] = x;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      const int y = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += y;
      __syncthreads();
    }
    if (i < n) out[i] = carry + sh[threadIdx.x] - x;
    carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}
}  // namespace

// Active out-neighbour lists: for every row u of `rows` (the active rows of layer l-1) the rows v it has an edge TO
// that are active at layer l (keep[v] >= 0), as a CSR indexed by the position of u in `rows`.  The data gradient of
// the pruned backward sums over exactly these; walking the full out-edge list of a hub (thousands of edges, one or
// two of them active) on every backward is what this replaces -- built once per meta-step.
// count: scratch [n_rows]; ptr: [n_rows + 1]; out_idx: capacity = number of edges of the set.
int active_out_lists_build(const int32_t* rows, int n_rows, const int32_t* t_indptr, const int32_t* t_indices,
                           const int32_t* keep, int32_t* count, int32_t* ptr, int32_t* out_idx, cudaStream_t stream) {
  if (n_rows == 0) return GMETA_OK;
  const int grid = ceil_div(n_rows, 8) < 8 * kNumSMs ? ceil_div(n_rows, 8) : 8 * kNumSMs;
  launch_pdl(active_out_kernel, dim3(grid), dim3(256), 0, stream, rows, n_rows, t_indptr, t_indices, keep, 0, count, nullptr, nullptr);
  int rc = check_launch();
  if (rc != GMETA_OK) return rc;
  launch_pdl(excl_scan_kernel, dim3(1), dim3(1024), 0, stream, count, n_rows, ptr);
  if ((rc = check_launch()) != GMETA_OK) return rc;
  launch_pdl(active_out_kernel, dim3(grid), dim3(256), 0, stream, rows, n_rows, t_indptr, t_indices, keep, 1, nullptr, ptr, out_idx);
  return check_launch();
}
}  // namespace gmeta

extern "C" int gmeta_aggregate_rows(const float* in, int32_t ld_in, const int32_t* in_row_map, const int32_t* dst_rows,
                                    const int32_t* indptr, const int32_t* indices, const float* norm, int32_t n_rows,
                                    int32_t f_in, int32_t scale_dst, float* out, int32_t ld_out, void* stream) {
  if (n_rows > 0 && !indptr) return GMETA_ERR_BAD_ARG;
  return aggregate_rows_impl(in, ld_in, in_row_map, dst_rows, indptr, indices, norm, n_rows, f_in, scale_dst, out, ld_out,
                             nullptr, (cudaStream_t)stream);
}

extern "C" int gmeta_aggregate_rows_nd(const float* in, int32_t ld_in, const int32_t* in_row_map, const int32_t* dst_rows,
                                       const int32_t* indptr, const int32_t* indices, const float* norm,
                                       const float* norm_dst, int32_t n_rows, int32_t f_in, int32_t scale_dst, float* out,
                                       int32_t ld_out, void* stream) {
  NormDstScope scope(norm_dst);
  return gmeta_aggregate_rows(in, ld_in, in_row_map, dst_rows, indptr, indices, norm, n_rows, f_in, scale_dst, out, ld_out,
                              stream);
}

namespace gmeta {
namespace {
__global__ void aggregation_norms_kernel(const int32_t* indptr, int n, int mode, float* norm_src, float* norm_dst) {
  pdl_prologue();
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    const int d = max(indptr[v + 1] - indptr[v], 1);
    float s = 1.f, r = 1.f;
    if (mode == GMETA_AGG_GCN) s = r = __fdiv_rn(1.0f, __fsqrt_rn((float)d));
    else if (mode == GMETA_AGG_MEAN) r = __fdiv_rn(1.0f, (float)d);
    norm_src[v] = s;
    norm_dst[v] = r;
  }
}
}  // namespace
}  // namespace gmeta

extern "C" int gmeta_aggregation_norms(const int32_t* indptr, int32_t n_nodes, int32_t mode, float* norm_src,
                                       float* norm_dst, void* stream) {
  if (n_nodes < 0 || (n_nodes > 0 && (!indptr || !norm_src || !norm_dst))) return GMETA_ERR_BAD_ARG;
  if (mode != GMETA_AGG_GCN && mode != GMETA_AGG_MEAN && mode != GMETA_AGG_SUM) return GMETA_ERR_BAD_ARG;
  if (n_nodes == 0) return GMETA_OK;
  const int grid = ceil_div(n_nodes, 256) < 8 * kNumSMs ? ceil_div(n_nodes, 256) : 8 * kNumSMs;
  launch_pdl(aggregation_norms_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, indptr, n_nodes, mode, norm_src, norm_dst);
  return check_launch();
}

extern "C" int gmeta_degree_norm(const int32_t* indptr, int32_t n_nodes, float* norm, void* stream) {
  if (n_nodes < 0 || (n_nodes > 0 && (!indptr || !norm))) return GMETA_ERR_BAD_ARG;
  if (n_nodes == 0) return GMETA_OK;
  const int grid = ceil_div(n_nodes, 256) < 8 * kNumSMs ? ceil_div(n_nodes, 256) : 8 * kNumSMs;
  launch_pdl(degree_norm_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, indptr, n_nodes, norm);
  return check_launch();
}

extern "C" int64_t gmeta_gcn_layer_wgrad_workspace_bytes(int32_t n_tasks, int32_t f_in, int32_t f_out) {
  if (n_tasks <= 0 || f_in <= 0 || f_out <= 0) return 0;
  const int s = pick_wgrad_split(n_tasks, f_in, f_out);
  return (int64_t)n_tasks * s * ((int64_t)f_in * f_out + f_out) * (int64_t)sizeof(float);
}

namespace gmeta {
// rows_hint: total number of rows the task_row_ptr covers (-1 = unknown).  With few rows per task the row-range split
// would only produce empty partials: the split count is capped by the average number of 32-row chunks per task,
// and with a single range per task the kernel writes dW / db directly (no partials, no reduction launch).
int gcn_layer_wgrad_impl(const float* in, int32_t ld_in, const int32_t* in_row_map, const int32_t* dst_rows,
                         const int32_t* indptr, const int32_t* indices, const float* norm, const int32_t* task_row_ptr,
                         int32_t n_tasks, const float* dZ, int32_t ld_dz, int32_t f_in, int32_t f_out, float* dW,
                         int64_t dw_task_stride, float* db, int64_t db_task_stride, void* workspace,
                         int64_t workspace_bytes, int64_t rows_hint, cudaStream_t s, int identity_graph) {
  if (!in || !indptr || !norm || !task_row_ptr || !dZ || !dW || !db || !workspace)
    return GMETA_ERR_BAD_ARG;
  if (n_tasks <= 0 || f_in <= 0 || f_out <= 0 || ld_in < f_in || ld_dz < f_out) return GMETA_ERR_BAD_ARG;
  if (workspace_bytes < gmeta_gcn_layer_wgrad_workspace_bytes(n_tasks, f_in, f_out)) return GMETA_ERR_WORKSPACE;
  if (!aligned16(workspace)) return GMETA_ERR_ALIGN;
  WgradParams p;
  p.g.in = in; p.g.in_row_map = in_row_map; p.g.dst_rows = dst_rows; p.g.indptr = indptr; p.g.indices = indices;
  p.g.norm = norm; p.g.ld_in = ld_in; p.g.f_in = f_in; p.g.norm_dst = g_norm_dst;
  p.task_row_ptr = task_row_ptr; p.n_tasks = n_tasks; p.dZ = dZ; p.ld_dz = ld_dz; p.f_out = f_out;
  p.n_split = pick_wgrad_split(n_tasks, f_in, f_out);
  if (rows_hint >= 0) {
    // a 32-row chunk costs a CTA ~2 us, a reduction launch ~13 us: up to 8 chunks per task run unsplit
    const int64_t avg_chunks = (rows_hint / n_tasks + WG_ROWS - 1) / WG_ROWS;
    if (avg_chunks <= 8) p.n_split = 1;
    else if (avg_chunks / 4 < p.n_split) p.n_split = (int)(avg_chunks / 4);
  }
  p.vec_dz = (ld_dz % 4 == 0) && aligned16(dZ);
  const int n_w = f_in * f_out;
  if (p.n_split == 1) {
    p.part_w = dW; p.pw_task_stride = dw_task_stride; p.pw_split_stride = 0;
    p.part_b = db; p.pb_task_stride = db_task_stride; p.pb_split_stride = 0;
  } else {
    p.part_w = (float*)workspace;
    p.part_b = p.part_w + (size_t)n_tasks * p.n_split * n_w;
    p.pw_task_stride = (long long)p.n_split * n_w; p.pw_split_stride = n_w;
    p.pb_task_stride = (long long)p.n_split * f_out; p.pb_split_stride = f_out;
  }
  const bool vec_in = (ld_in % 4 == 0) && aligned16(in) && ld_in >= round_up(f_in, 4);
  const int n_work = n_tasks * ceil_div(f_in, KP) * ceil_div(f_out, BN) * p.n_split;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gcn_layer_wgrad_simt_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmem);
    cudaFuncSetAttribute(gcn_layer_wgrad_simt_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmem);
    cudaFuncSetAttribute(gcn_layer_wgrad_simt_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmem);
    attr_done = true;
  }
  p.g.identity = (identity_graph && vec_in && f_in % 4 == 0 && !in_row_map && !dst_rows) ? 1 : 0;
  if (p.g.identity)
    launch_pdl(gcn_layer_wgrad_simt_kernel<true, true>, dim3(n_work), dim3(NTHREADS), kWgSmem, s, p);
  else if (vec_in)
    launch_pdl(gcn_layer_wgrad_simt_kernel<true, false>, dim3(n_work), dim3(NTHREADS), kWgSmem, s, p);
  else
    launch_pdl(gcn_layer_wgrad_simt_kernel<false, false>, dim3(n_work), dim3(NTHREADS), kWgSmem, s, p);
  int rc = check_launch();
  if (rc != GMETA_OK || p.n_split == 1) return rc;
  const long long total = (long long)n_tasks * (n_w + f_out);
  const int grid = (int)((total + 255) / 256 < 8 * kNumSMs ? (total + 255) / 256 : 8 * kNumSMs);
  launch_pdl(wgrad_reduce_kernel, dim3(grid), dim3(256), 0, s, p.part_w, p.part_b, n_tasks, p.n_split, n_w, f_out, dW,
                                          dw_task_stride, db, db_task_stride);
  return check_launch();
}
}  // namespace gmeta

extern "C" int gmeta_gcn_layer_wgrad(const float* in, int32_t ld_in, const int32_t* in_row_map,
                                     const int32_t* dst_rows, const int32_t* indptr, const int32_t* indices, const float* norm,
                                     const int32_t* task_row_ptr, int32_t n_tasks, const float* dZ,
                                     int32_t ld_dz, int32_t f_in, int32_t f_out, float* dW,
                                     int64_t dw_task_stride, float* db, int64_t db_task_stride,
                                     void* workspace, int64_t workspace_bytes, void* stream) {
  return gcn_layer_wgrad_impl(in, ld_in, in_row_map, dst_rows, indptr, indices, norm, task_row_ptr, n_tasks, dZ, ld_dz,
                              f_in, f_out, dW, dw_task_stride, db, db_task_stride, workspace, workspace_bytes, -1,
                              (cudaStream_t)stream, 0);
}

extern "C" int gmeta_gcn_layer_wgrad_nd(const float* in, int32_t ld_in, const int32_t* in_row_map, const int32_t* dst_rows,
                                        const int32_t* indptr, const int32_t* indices, const float* norm,
                                        const float* norm_dst, const int32_t* task_row_ptr, int32_t n_tasks,
                                        const float* dZ, int32_t ld_dz, int32_t f_in, int32_t f_out, float* dW,
                                        int64_t dw_task_stride, float* db, int64_t db_task_stride, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
  NormDstScope scope(norm_dst);
  return gmeta_gcn_layer_wgrad(in, ld_in, in_row_map, dst_rows, indptr, indices, norm, task_row_ptr, n_tasks, dZ, ld_dz,
                               f_in, f_out, dW, dw_task_stride, db, db_task_stride, workspace, workspace_bytes, stream);
}
