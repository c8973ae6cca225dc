// Shared helpers for the gmeta_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gmeta_b200.h"

namespace gmeta {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// per-thread launch counter: gmeta_last_launch_count() reports what one driver call enqueued
extern thread_local int g_launch_count;

inline int check_launch() {
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? GMETA_OK : GMETA_ERR_LAUNCH;
}

// Destination-side scales of the layer call being issued on this host thread (NULL = symmetric, the reference's GraphConv):
// set by the *_nd entry points and by the step driver around the calls they forward to, read where a GatherSrc is filled.
extern thread_local const float* g_norm_dst;
struct NormDstScope {
  const float* saved;
  explicit NormDstScope(const float* p) : saved(g_norm_dst) { g_norm_dst = p; }
  ~NormDstScope() { g_norm_dst = saved; }
};

// ---- programmatic dependent launch (opt-in: GMETA_B200_PDL=1) ----
// Every kernel of the library can be launched with the programmatic-serialization attribute and begins with
// pdl_prologue(): "my dependents may be scheduled" (they park at their own wait) and "wait until the kernel before
// me has completed and its writes are visible".  Stream order is unchanged -- a kernel touches memory only after
// its wait, and it completes only after that, so completion stays transitive along the chain -- but the launch
// latency and block scheduling of kernel i+1 overlap kernel i.  Measured on the C2 meta-step (~200 dependent
// launches on two streams, round 2): 2.99 ms per step without, 3.21 ms with it -- the parked blocks of one stream's
// next kernel take SM slots from the other stream's kernels -- and no measurable change of the full-layer launch
// or of the eager end-to-end step, so it is OFF unless GMETA_B200_PDL=1.  Both instructions are no-ops in a
// launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait();
}
bool pdl_enabled();     // api.cu
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors surface in check_launch()
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int ceil_div(int x, int m) { return (x + m - 1) / m; }

// ReLU as torch computes it (learner.py:49 F.relu): a NaN pre-activation stays NaN, so that a NaN anywhere reaches the
// mean query loss and the reference's NaN skip (meta.py:163-164) fires.  fmaxf(NaN, 0) would return 0.  One FMNMX.NAN.
__device__ __forceinline__ float relu_keep_nan(float v) {
  float r;
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Adjacency + row source of a gather: M[v,:] = sum_{e in [indptr[v],indptr[v+1])} norm[u_e] * in[map(u_e),:]
struct GatherSrc {
  const float* in;
  const int32_t* in_row_map;  // nullable; a negative entry means "this neighbour contributes nothing"
  const int32_t* dst_rows;    // nullable; real row of compact row i (adjacency, norm, mask use the real row)
  const int32_t* indptr;
  const int32_t* indices;
  const float* norm;          // scale of a row as a SOURCE (norm[u] above)
  int ld_in;
  int f_in;
  // 1: the graph is the identity (row v's only in-neighbour is v, indptr = indices = 0..N, norm = 1, no row map /
  // row list): the pre-summed rows of the pruned meta-step.  Kernels that honour it skip the index loads.
  int identity = 0;
  // scale of a row as a DESTINATION (the factor in front of the sum); NULL = `norm` (the reference's symmetric
  // GraphConv normalisation, learner.py:29-49).  Mean aggregation: norm = 1, norm_dst = 1 / in-degree.
  const float* norm_dst = nullptr;
};
__device__ __forceinline__ float dst_norm(const GatherSrc& g, int v) { return (g.norm_dst ? g.norm_dst : g.norm)[v]; }

// One warp aggregates, for each of its rows r = warp, warp+NW, ... < R, the 4 columns
// [kcol, kcol+4) (kcol = k0 + 4*lane) of M[row0+r] into As[r*lda + 4*lane .. +3]; rows >= nrows
// (tile tail) are written as zero.  Neighbour ids, their norms and source rows are loaded
// coalesced (one edge per lane) and broadcast with shuffles; the next row's edge batch is
// prefetched while the current row's feature segments are in flight.  Summation follows edge
// order, so results are run-to-run deterministic.  VEC: 16-byte loads (ld_in % 4 == 0 and
// 16-byte aligned base); otherwise scalar loads with exact column bounds.
template <bool VEC, int R, int NW, bool SCALE_DST>
__device__ __forceinline__ void gather_rows(const GatherSrc& g, int row0, int nrows, int k0,
                                            float* As, int lda) {
  constexpr int RPW = R / NW;  // rows per warp
  static_assert(RPW <= 32, "one lane per row for the indptr preload");
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int kcol = k0 + 4 * lane;

  // lane i holds [beg,end) of this warp's i-th row
  int my_beg = 0, my_end = 0;
  {
    const int r = warp + NW * lane;
    if (lane < RPW && r < nrows) {
      const int v = g.dst_rows ? g.dst_rows[row0 + r] : row0 + r;
      my_beg = g.indptr[v];
      my_end = g.indptr[v + 1];
    }
  }
  auto load_batch = [&](int base, int end, int& u_src, float& u_norm) {
    const int e = base + lane;
    u_src = 0;
    u_norm = 0.f;
    if (e < end) {
      const int u = g.indices[e];
      u_norm = g.norm[u];
      u_src = g.in_row_map ? g.in_row_map[u] : u;
      if (u_src < 0) { u_src = 0; u_norm = 0.f; }   // skipped neighbour: weight 0 on a valid row
    }
  };
  int nb = __shfl_sync(0xffffffffu, my_beg, 0), ne = __shfl_sync(0xffffffffu, my_end, 0);
  int p_src;
  float p_norm;
  load_batch(nb, ne, p_src, p_norm);

#pragma unroll 1
  for (int i = 0; i < RPW; ++i) {
    const int r = warp + NW * i;
    const int beg = nb, end = ne;
    int c_src = p_src;
    float c_norm = p_norm;
    if (i + 1 < RPW) {  // prefetch the next row's first edge batch
      nb = __shfl_sync(0xffffffffu, my_beg, i + 1);
      ne = __shfl_sync(0xffffffffu, my_end, i + 1);
      load_batch(nb, ne, p_src, p_norm);
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = beg; base < end; base += 32) {
      if (base != beg) load_batch(base, end, c_src, c_norm);
      const int cnt = min(32, end - base);
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const int src = __shfl_sync(0xffffffffu, c_src, j);
        const float nu = __shfl_sync(0xffffffffu, c_norm, j);
        const float* p = g.in + (size_t)src * g.ld_in + kcol;
        float4 x;
        if (VEC) {
          x = (kcol < g.f_in) ? ld_f4(p) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (kcol + 4 > g.f_in) {  // ragged tail: columns >= f_in may hold anything
            if (kcol + 1 >= g.f_in) x.y = 0.f;
            if (kcol + 2 >= g.f_in) x.z = 0.f;
            if (kcol + 3 >= g.f_in) x.w = 0.f;
          }
        } else {
          x.x = (kcol + 0 < g.f_in) ? p[0] : 0.f;
          x.y = (kcol + 1 < g.f_in) ? p[1] : 0.f;
          x.z = (kcol + 2 < g.f_in) ? p[2] : 0.f;
          x.w = (kcol + 3 < g.f_in) ? p[3] : 0.f;
        }
        acc.x = fmaf(nu, x.x, acc.x);
        acc.y = fmaf(nu, x.y, acc.y);
        acc.z = fmaf(nu, x.z, acc.z);
        acc.w = fmaf(nu, x.w, acc.w);
      }
    }
    if (SCALE_DST && r < nrows) {
      const float nv = dst_norm(g, g.dst_rows ? g.dst_rows[row0 + r] : row0 + r);
      acc.x *= nv; acc.y *= nv; acc.z *= nv; acc.w *= nv;
    }
    st_f4(As + r * lda + 4 * lane, acc);
  }
}

}  // namespace gmeta
