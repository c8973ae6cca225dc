// Fused GCN layer over a packed meta-batch -- CTA-pair tensor-core implementation (sm_100a).
//
//   out[i,:] = act( norm[v] * (sum_{u in N_in(v)} norm[u] * in[map(u),:]) . B_task + bias_task )
//
// replaces GraphConv.forward (reference G-Meta/learner.py:25-56) for f_in % 64 == 0,
// f_out % 16 == 0, f_out <= 256 and 4*f_in*f_out/2 bytes of weights fitting next to two operand
// stages in shared memory; other shapes take gcn_layer_tc.cu (3xTF32, streamed weights) or the
// FFMA kernel.
//
// Design (why it looks like this):
//   * The contraction must match a true-fp32 GEMM to ~1e-6, so both operands are split into an
//     error-compensated pair of 11-bit-significand halves x = hi + lo and hi*hi + lo*hi + hi*lo
//     is accumulated in fp32 in tensor memory.  The halves are FP16 (kind::f16 runs at twice the
//     kind::tf32 rate and needs half the shared memory); FP16's narrow exponent range is handled
//     by exact power-of-two scaling: every aggregated row is scaled by 2^e(row) chosen from a
//     rigorous bound (sum_u norm[u] * max|in[u,:]| from a per-row abs-max vector of the input),
//     every task's weight matrix by 2^e(task) from its abs-max; the epilogue multiplies the
//     inverse back (exact), so no value can overflow and the split keeps ~2^-22 relative
//     accuracy w.r.t. the row magnitude.
//   * With per-task fast weights a 256x256 weight matrix would have to be re-streamed from L2 for
//     every 128-row tile (4x the HBM traffic of the layer).  Instead two CTAs (one cluster =
//     one SM pair) issue tcgen05.mma.cta_group::2: each CTA keeps HALF of the output columns of
//     the task's weights resident in shared memory (K x N/2, hi+lo = 128 KB at 256x256) and
//     gathers its own 128 rows; the weights are re-loaded only when the cluster moves on to the
//     next task (pairs of tiles are assigned to clusters in contiguous runs).
//   * Everything that depends only on the graph STRUCTURE is hoisted into a "plan": per output
//     row the mapped source rows and norms of its first two in-neighbours (96% of the rows of a
//     2-hop subgraph batch have <= 2), and for longer rows (hubs) a slot in a flattened edge list
//     that a small edge-parallel kernel aggregates first; the fused kernel then reads a hub like
//     a single neighbour with weight 1.
//
// Warp roles per CTA (704 threads): warps 0..15 gather producers (quarter-warp per row, 64-float
// K chunks, fp32 sum -> scaled FP16 hi/lo -> 128B-swizzled K-major operand stage), warp 16 weight
// loader (cp.async.bulk of the pre-split, pre-swizzled image), warp 17 MMA issuer (leader CTA
// only; one thread), warps 18..21 epilogue (tcgen05.ld, * norm * 2^-e + bias, ReLU / mask, row
// abs-max for the next layer, 16-byte stores).  Hand-offs are mbarriers; the peer CTA signals
// the leader's barriers through the cluster address space, the MMA thread releases operand
// stages / accumulators in both CTAs with multicast commits.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace gmeta {
namespace {

using namespace ptx;

constexpr int TM = GMETA_TILE_ROWS;            // 128 rows per CTA tile; the pair's MMA has M = 256
constexpr int KCH = 64;                        // fp16 elements per K chunk = one 128-byte swizzle row
constexpr int A_HALF_BYTES = TM * 128;         // 16 KB: hi (or lo) operand tile of one chunk
constexpr int STAGE_BYTES = 2 * A_HALF_BYTES;  // 32 KB
constexpr int N_PROD_WARPS = 16;
constexpr int WARP_LOAD = 16;
constexpr int WARP_MMA = 17;
constexpr int WARP_EPI0 = 18;
constexpr int NTHREADS = 22 * 32;
constexpr int MAX_STAGES = 4;
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = 256;
constexpr int PRE = 2;                         // in-neighbours per row the fused kernel gathers itself
constexpr int SMEM_FIXED = 256 /*barriers*/ + 4 * TM /*row scale exponents*/;
constexpr int SMEM_MAX = 227 * 1024;
constexpr int PT_MAXT = 2048;                  // tasks the pair-table kernel handles
constexpr int SCALE_TARGET = 13;               // scaled bound in [2^13, 2^14): 4x below the FP16 maximum
constexpr int SCALE_CLAMP = 100;

// ---- plan (structure only) ----
struct PlanRec {      // 16 bytes per output row
  int r0, r1;         // mapped source rows of in-neighbours 0/1; hub row: r0 = slot, r1 = -1
  float n0, n1;       // their norms (0 = absent or dropped)
};
struct Plan {
  int* hdr;           // [0] n_long  [1] n_long_edges  [2] n_pairs
  PlanRec* rec;       // [n_rows]
  int2* pair_tiles;   // [cap_pairs] the two tiles of a pair (same task); .y = -1 when the task has an odd tile count
  int* pair_task;     // [cap_pairs]
  int* long_row;      // [cap_long] real row of each hub slot
  int* long_beg;      // [cap_long] first record of the slot in long_src / long_nrm
  int* long_deg;      // [cap_long]
  int* long_src;      // [n_edges] mapped source row of each hub edge
  float* long_nrm;    // [n_edges]
};
struct Workspace {
  Plan plan;
  unsigned* w_absmax;   // [n_copies] bit pattern of max|W_c|
  float* w_inv_scale;   // [n_copies] 2^-e(c)
  __half* w_image;      // [n_copies][rank 2][K/64][hi|lo][N/2 rows][64 halves, 128B swizzle]
  float* mlong;         // [cap_long][f_in] aggregated hub rows
  float* mlong_rowmax;  // [cap_long]
  int64_t total;
};

inline int64_t al(int64_t x) { return (x + 255) / 256 * 256; }
inline int cap_pairs_for(int n_tiles, int n_tasks) { return (n_tiles + n_tasks) / 2 + 1; }
inline int cap_long_for(int n_rows, int n_edges) {
  const int64_t by_edges = (int64_t)n_edges / (PRE + 1) + 1;
  return (int)(by_edges < n_rows ? by_edges : n_rows) + 1;
}

Workspace carve(void* base, int n_copies, int n_tiles, int n_tasks, int n_rows, int n_edges, int K, int N) {
  Workspace w;
  char* p = reinterpret_cast<char*>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) { char* q = p ? p + off : nullptr; off += al(bytes); return q; };
  const int cp = cap_pairs_for(n_tiles, n_tasks), cl = cap_long_for(n_rows, n_edges);
  w.plan.hdr = reinterpret_cast<int*>(take(256));
  w.w_absmax = reinterpret_cast<unsigned*>(take((int64_t)n_copies * 4));
  w.w_inv_scale = reinterpret_cast<float*>(take((int64_t)n_copies * 4));
  w.w_image = reinterpret_cast<__half*>(take((int64_t)n_copies * 2 * K * N * 2));
  w.plan.rec = reinterpret_cast<PlanRec*>(take((int64_t)n_rows * 16));
  w.plan.pair_tiles = reinterpret_cast<int2*>(take((int64_t)cp * 8));
  w.plan.pair_task = reinterpret_cast<int*>(take((int64_t)cp * 4));
  w.plan.long_row = reinterpret_cast<int*>(take((int64_t)cl * 4));
  w.plan.long_beg = reinterpret_cast<int*>(take((int64_t)cl * 4));
  w.plan.long_deg = reinterpret_cast<int*>(take((int64_t)cl * 4));
  w.plan.long_src = reinterpret_cast<int*>(take((int64_t)n_edges * 4 + 4));
  w.plan.long_nrm = reinterpret_cast<float*>(take((int64_t)n_edges * 4 + 4));
  w.mlong = reinterpret_cast<float*>(take((int64_t)cl * K * 4));
  w.mlong_rowmax = reinterpret_cast<float*>(take((int64_t)cl * 4));
  w.total = off;
  return w;
}

// exponent e such that bound * 2^e lies in [2^SCALE_TARGET, 2^(SCALE_TARGET+1)); 0 for bound == 0
__device__ __forceinline__ int scale_exponent(float bound) {
  if (!(bound > 0.f)) return 0;
  int e = SCALE_TARGET - (int)((__float_as_uint(bound) >> 23) & 0xFFu) + 127;
  e = e > SCALE_CLAMP ? SCALE_CLAMP : e;
  return e < -SCALE_CLAMP ? -SCALE_CLAMP : e;
}

struct PairParams {
  const float* in;
  int ld_in;
  int f_in;
  const float* in_rowmax;
  const float* mlong;
  const float* mlong_rowmax;
  const PlanRec* rec;
  const int* hdr;
  const int2* pair_tiles;
  const int* pair_task;
  const int32_t* tile_row0;
  const int32_t* tile_nrows;
  const int32_t* dst_rows;
  const float* norm;
  const __half* w_image;
  long long image_task_stride;   // halves between task copies (0 = shared weights)
  const float* w_inv_scale;
  const float* bias;
  long long b_task_stride;
  int f_out;
  int relu;
  const float* relu_mask;
  float* out;
  int ld_out;
  float* out_rowmax;
  int n_stages;
  int w_bytes;                   // this CTA's resident weight image: 2 * f_in * f_out bytes
  int dbg;                       // debug ablation flags: 1 skip output stores, 2 skip gather loads, 4 issue 1/4 of the MMAs
};

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// v[0..7] (already scaled) -> 8 FP16 hi and 8 FP16 lo = v - hi
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 hh = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    const float2 back = __half22float2(hh);
    h[j] = *reinterpret_cast<const uint32_t*>(&hh);
    l[j] = pack_half2(v[2 * j] - back.x, v[2 * j + 1] - back.y);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
gcn_layer_fwd_pair_kernel(const PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int N = p.f_out, K = p.f_in;
  const int nkc = K / KCH;
  const int NS = p.n_stages;
  const int half_n_bytes = (N / 2) * 128;                 // one chunk of W hi (or lo) in this CTA
  uint8_t* w_s = smem;
  uint8_t* a_s = smem + p.w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_s + (size_t)NS * STAGE_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };                    // leader's: 32 producer warps of the pair
  auto empty = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };      // per CTA: multicast commit
  auto acc_full = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + b); };       // per CTA: multicast commit
  auto acc_empty = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 2 + b); };  // leader's: 8 epilogue warps of the pair
  const uint32_t w_local = bar0 + 8u * (2 * MAX_STAGES + 4);             // per CTA: bulk copy landed
  const uint32_t w_ready = bar0 + 8u * (2 * MAX_STAGES + 5);             // leader's: both CTAs hold the task's weights
  const uint32_t w_free = bar0 + 8u * (2 * MAX_STAGES + 6);              // per CTA: MMAs of the previous task are done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 8);
  int8_t* scale_e = reinterpret_cast<int8_t*>(bars) + 256;               // [4][TM]

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { printf("gmeta pair kernel: shared memory base not 1024-byte aligned\n"); __trap(); }
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(a_full(s), 2 * N_PROD_WARPS);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 8);
    }
    mbar_init(w_local, 1);
    mbar_init(w_ready, 2);
    mbar_init(w_free, 1);
    fence_mbar_init_cluster();
  }
  if (warp == WARP_LOAD) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();        // peers' barriers are initialised and TMEM is allocated in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // contiguous run of tile pairs for this cluster
  const int n_pairs = p.hdr[2];
  const int n_cl = gridDim.x >> 1, cid = blockIdx.x >> 1;
  const int ppc = (n_pairs + n_cl - 1) / n_cl;
  const int p_beg = cid * ppc < n_pairs ? cid * ppc : n_pairs;
  const int p_end = p_beg + ppc < n_pairs ? p_beg + ppc : n_pairs;

  if (warp < N_PROD_WARPS) {
    // ===================== gather producers =====================
    const int q = warp * 4 + (lane >> 3);   // a quarter-warp owns tile rows q and q + 64
    const int sub = lane & 7;               // and, within a chunk, floats [8*sub, 8*sub + 8)
    int it = 0, ti = 0;
    for (int pr = p_beg; pr < p_end; ++pr, ++ti) {
      const int2 pt = p.pair_tiles[pr];
      const int tile = rank ? pt.y : pt.x;
      int nrows = 0, row0 = 0;
      if (tile >= 0) { nrows = p.tile_nrows[tile]; row0 = p.tile_row0[tile]; }
      const float* src0[2];
      const float* src1[2];
      float n0[2], n1[2];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = q + 64 * rr;
        src0[rr] = src1[rr] = p.in;
        n0[rr] = n1[rr] = 0.f;
        if (r < nrows) {
          const int4 rc = __ldg(reinterpret_cast<const int4*>(p.rec + row0 + r));
          const bool hub = rc.y < 0;
          float a0 = __int_as_float(rc.z), a1 = __int_as_float(rc.w);
          float rm0 = 0.f, rm1 = 0.f;
          if (a0 != 0.f) {
            src0[rr] = hub ? p.mlong + (size_t)rc.x * K : p.in + (size_t)rc.x * p.ld_in;
            rm0 = hub ? p.mlong_rowmax[rc.x] : p.in_rowmax[rc.x];
          }
          if (a1 != 0.f) {
            src1[rr] = p.in + (size_t)rc.y * p.ld_in;
            rm1 = p.in_rowmax[rc.y];
          }
          const int e = scale_exponent(a0 * rm0 + a1 * rm1);
          const float sc = exp2i(e);
          n0[rr] = a0 * sc;
          n1[rr] = a1 * sc;
          if (sub == 0) scale_e[(ti & 3) * TM + r] = (int8_t)e;
        }
      }
      for (int kc = 0; kc < nkc; ++kc, ++it) {
        const int s = it % NS;
        const uint32_t ph = (uint32_t)((it / NS) & 1);
        // issue this chunk's loads before waiting for the stage: their latency overlaps the wait
        float4 x0[2][2], x1[2][2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int off = kc * KCH + sub * 8 + 4 * h;
            x0[rr][h] = (n0[rr] != 0.f && !(p.dbg & 2)) ? ld_f4(src0[rr] + off) : make_float4(0.f, 0.f, 0.f, 0.f);
            x1[rr][h] = (n1[rr] != 0.f && !(p.dbg & 2)) ? ld_f4(src1[rr] + off) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        mbar_wait(empty(s), ph ^ 1u, 1);
        uint8_t* stage = a_s + (size_t)s * STAGE_BYTES;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = q + 64 * rr;
          if (r < nrows) {
            float v[8];
            v[0] = fmaf(n1[rr], x1[rr][0].x, n0[rr] * x0[rr][0].x);
            v[1] = fmaf(n1[rr], x1[rr][0].y, n0[rr] * x0[rr][0].y);
            v[2] = fmaf(n1[rr], x1[rr][0].z, n0[rr] * x0[rr][0].z);
            v[3] = fmaf(n1[rr], x1[rr][0].w, n0[rr] * x0[rr][0].w);
            v[4] = fmaf(n1[rr], x1[rr][1].x, n0[rr] * x0[rr][1].x);
            v[5] = fmaf(n1[rr], x1[rr][1].y, n0[rr] * x0[rr][1].y);
            v[6] = fmaf(n1[rr], x1[rr][1].z, n0[rr] * x0[rr][1].z);
            v[7] = fmaf(n1[rr], x1[rr][1].w, n0[rr] * x0[rr][1].w);
            uint4 hi, lo;
            split8(v, hi, lo);
            const int off = r * 128 + ((sub ^ (r & 7)) << 4);   // 128B swizzle: 16-byte unit ^ (row % 8)
            *reinterpret_cast<uint4*>(stage + off) = hi;
            *reinterpret_cast<uint4*>(stage + A_HALF_BYTES + off) = lo;
          }
        }
        fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(a_full(s), 0);
      }
    }
  } else if (warp == WARP_LOAD) {
    // ===================== weight loader: one bulk copy per task change =====================
    if (lane == 0) {
      int cur = -1, n = 0;
      for (int pr = p_beg; pr < p_end; ++pr) {
        const int task = p.pair_task[pr];
        if (task == cur) continue;
        if (n > 0) mbar_wait(w_free, (uint32_t)((n - 1) & 1), 2);   // MMAs reading the old image are complete
        const uint8_t* img = reinterpret_cast<const uint8_t*>(p.w_image + (long long)task * p.image_task_stride) +
                             (size_t)rank * p.w_bytes;
        mbar_arrive_expect_tx(w_local, (uint32_t)p.w_bytes);
        for (int o = 0; o < p.w_bytes; o += 16384) {
          const int nb = p.w_bytes - o < 16384 ? p.w_bytes - o : 16384;
          bulk_copy_g2s(smem_u32(w_s + o), img + o, (uint32_t)nb, w_local);
        }
        mbar_wait(w_local, (uint32_t)(n & 1), 3);
        mbar_arrive_cluster(w_ready, 0);
        cur = task;
        ++n;
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc_f16(2 * TM, N);
      int it = 0, ti = 0, cur = -1, nw = 0;
      for (int pr = p_beg; pr < p_end; ++pr, ++ti) {
        const int task = p.pair_task[pr];
        if (task != cur) {
          mbar_wait_cluster(w_ready, (uint32_t)(nw & 1), 4);
          cur = task;
          ++nw;
        }
        const int buf = ti & 1;
        mbar_wait_cluster(acc_empty(buf), (uint32_t)(((ti >> 1) & 1) ^ 1), 5);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ACC_COLS);
        for (int kc = 0; kc < nkc; ++kc, ++it) {
          const int s = it % NS;
          const uint32_t ph = (uint32_t)((it / NS) & 1);
          mbar_wait_cluster(a_full(s), ph, 6);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_s + (size_t)s * STAGE_BYTES);
          const uint32_t b_addr = smem_u32(w_s + (size_t)kc * 2 * half_n_bytes);
          const uint64_t da_hi = umma_desc_k_sw128(a_addr);
          const uint64_t da_lo = umma_desc_k_sw128(a_addr + A_HALF_BYTES);
          const uint64_t db_hi = umma_desc_k_sw128(b_addr);
          const uint64_t db_lo = umma_desc_k_sw128(b_addr + half_n_bytes);
#pragma unroll
          for (int k = 0; k < KCH / 16; ++k) {        // UMMA_K = 16 halves = 32 bytes = 2 descriptor units
            if ((p.dbg & 4) && k) break;
            const uint64_t adv = (uint64_t)(2 * k);
            tc_mma_f16_pair(d_tmem, da_hi + adv, db_hi + adv, idesc, (kc | k) != 0 ? 1u : 0u);
            tc_mma_f16_pair(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
            tc_mma_f16_pair(d_tmem, da_hi + adv, db_lo + adv, idesc, 1u);
          }
          tc_commit_pair(empty(s));          // frees the stage in both CTAs once these MMAs have read it
        }
        tc_commit_pair(acc_full(buf));       // accumulators complete -> both epilogues
        if (pr + 1 < p_end && p.pair_task[pr + 1] != task) tc_commit_pair(w_free);
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;            // TMEM lanes 32*quarter .. +31 are the ones this warp may read
    const int r = quarter * 32 + lane;
    int ti = 0;
    for (int pr = p_beg; pr < p_end; ++pr, ++ti) {
      const int buf = ti & 1;
      const int2 pt = p.pair_tiles[pr];
      const int tile = rank ? pt.y : pt.x;
      const int task = p.pair_task[pr];
      int nrows = 0, row0 = 0;
      if (tile >= 0) { nrows = p.tile_nrows[tile]; row0 = p.tile_row0[tile]; }
      const bool live = r < nrows;
      const int oi = row0 + (live ? r : 0);                           // output row (compact or dense)
      const int v = p.dst_rows ? p.dst_rows[oi] : oi;                 // real row: norm and mask
      const float nv = p.norm[v];
      const float wis = p.w_inv_scale[p.image_task_stride ? task : 0];
      const float* bias = p.bias ? p.bias + (long long)task * p.b_task_stride : nullptr;
      const float* mrow = p.relu_mask ? p.relu_mask + (size_t)v * p.ld_out : nullptr;
      float* orow = p.out + (size_t)oi * p.ld_out;
      mbar_wait(acc_full(buf), (uint32_t)((ti >> 1) & 1), 7);
      tc_fence_after();
      const float f = live ? nv * exp2i(-(int)scale_e[(ti & 3) * TM + r]) * wis : 0.f;
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * ACC_COLS);
      float rmax = 0.f;
      uint32_t acc_n[16];
      tmem_ld16(t_addr, acc_n);
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t acc[16];
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = acc_n[j];
        if (c0 + 16 < N) tmem_ld16(t_addr + (uint32_t)(c0 + 16), acc_n);   // next chunk in flight
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          o[j] = fmaf(f, __uint_as_float(acc[j]), b4.x);
          o[j + 1] = fmaf(f, __uint_as_float(acc[j + 1]), b4.y);
          o[j + 2] = fmaf(f, __uint_as_float(acc[j + 2]), b4.z);
          o[j + 3] = fmaf(f, __uint_as_float(acc[j + 3]), b4.w);
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
        }
        if (mrow && live) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 m4 = ld_f4(mrow + c0 + j);
            if (!(m4.x > 0.f)) o[j] = 0.f;
            if (!(m4.y > 0.f)) o[j + 1] = 0.f;
            if (!(m4.z > 0.f)) o[j + 2] = 0.f;
            if (!(m4.w > 0.f)) o[j + 3] = 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) rmax = fmaxf(rmax, fabsf(o[j]));
        if (live && !(p.dbg & 1)) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) st_f4(orow + c0 + j, make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]));
        }
      }
      if (p.out_rowmax && live) p.out_rowmax[oi] = rmax;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty(buf), 0);
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();       // nobody leaves while the peer may still read this CTA's shared memory
  if (warp == WARP_LOAD) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// weights: abs-max per copy, then the scaled FP16 hi/lo image in the operand layout
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) w_absmax_kernel(const float* __restrict__ W, long long w_stride, int ldw,
                                                       int trans, int K, int N, unsigned* __restrict__ absmax) {
  __shared__ float red[8];
  const float* w = W + (long long)blockIdx.y * w_stride;
  const int inner = trans ? K : N, total = K * N;
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(w[(size_t)(i / inner) * ldw + (i % inner)]));
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    atomicMax(absmax + blockIdx.y, __float_as_uint(m));   // non-negative floats order like their bit patterns
  }
}

__global__ void pack_w_pair_kernel(const float* __restrict__ W, long long w_stride, int ldw, int trans, int K, int N,
                                   int n_copies, const unsigned* __restrict__ absmax, __half* __restrict__ image,
                                   long long image_stride, float* __restrict__ w_inv_scale) {
  const int ku = K / 8, nkc = K / KCH, hn = N / 2;
  const long long total = (long long)n_copies * ku * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const int u = (int)((i / N) % ku);
    const int c = (int)(i / ((long long)N * ku));
    const int e = scale_exponent(__uint_as_float(absmax[c]));
    const float sc = exp2i(e);
    if (n == 0 && u == 0) w_inv_scale[c] = exp2i(-e);
    const float* w = W + c * w_stride;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = u * 8 + j;
      v[j] = sc * (trans ? w[(size_t)n * ldw + k] : w[(size_t)k * ldw + n]);
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    const int rank = n / hn, nn = n - rank * hn, kc = u >> 3, unit = u & 7;
    __half* dst = image + c * image_stride + ((((size_t)rank * nkc + kc) * 2) * hn + nn) * 64 + ((unit ^ (nn & 7)) << 3);
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + (size_t)hn * 64) = lo;
  }
}

// per-row abs-max of a row-major matrix (one warp per row)
__global__ void row_absmax_kernel(const float* __restrict__ x, int ld, int n_rows, int f, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += (gridDim.x * blockDim.x) >> 5) {
    const float* row = x + (size_t)r * ld;
    float m = 0.f;
    for (int k = lane; k < f; k += 32) m = fmaxf(m, fabsf(row[k]));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) out[r] = m;
  }
}

// ------------------------------------------------------------------------------------------
// plan construction (structure only) and the hub-row pre-aggregation
// ------------------------------------------------------------------------------------------
__global__ void plan_classify_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                     const float* __restrict__ norm, const int32_t* __restrict__ in_row_map,
                                     const int32_t* __restrict__ dst_rows, int n_rows, Plan pl) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += gridDim.x * blockDim.x) {
    const int v = dst_rows ? dst_rows[i] : i;
    const int beg = indptr[v], deg = indptr[v + 1] - beg;
    PlanRec r = {0, 0, 0.f, 0.f};
    if (deg > PRE) {
      const int slot = atomicAdd(pl.hdr + 0, 1);          // slot order does not affect any value
      pl.long_row[slot] = v;
      pl.long_beg[slot] = atomicAdd(pl.hdr + 1, deg);
      pl.long_deg[slot] = deg;
      r.r0 = slot; r.r1 = -1; r.n0 = 1.f;
    } else {
      if (deg > 0) {
        const int u = indices[beg];
        const int s = in_row_map ? in_row_map[u] : u;
        if (s >= 0) { r.r0 = s; r.n0 = norm[u]; }          // negative map entry: neighbour dropped
      }
      if (deg > 1) {
        const int u = indices[beg + 1];
        const int s = in_row_map ? in_row_map[u] : u;
        if (s >= 0) { r.r1 = s; r.n1 = norm[u]; }
      }
    }
    pl.rec[i] = r;
  }
}

// one warp per hub slot: its edges, coalesced, in CSR order
__global__ void plan_long_edges_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                       const float* __restrict__ norm, const int32_t* __restrict__ in_row_map,
                                       Plan pl) {
  const int n_long = pl.hdr[0];
  const int lane = threadIdx.x & 31;
  for (int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slot < n_long; slot += (gridDim.x * blockDim.x) >> 5) {
    const int beg = indptr[pl.long_row[slot]], deg = pl.long_deg[slot], dst = pl.long_beg[slot];
    for (int e = lane; e < deg; e += 32) {
      const int u = indices[beg + e];
      const int s = in_row_map ? in_row_map[u] : u;
      pl.long_src[dst + e] = s < 0 ? 0 : s;
      pl.long_nrm[dst + e] = s < 0 ? 0.f : norm[u];
    }
  }
}

// tiles -> pairs of tiles of the same task (tiles of a task are contiguous in the tile table)
__global__ void __launch_bounds__(1024) pair_table_kernel(const int32_t* __restrict__ tile_task, int n_tiles,
                                                          int n_tasks, Plan pl) {
  __shared__ int first[PT_MAXT], cnt[PT_MAXT], base[PT_MAXT], tmp[PT_MAXT];
  for (int t = threadIdx.x; t < n_tasks; t += blockDim.x) { first[t] = 0x7fffffff; cnt[t] = 0; }
  __syncthreads();
  for (int i = threadIdx.x; i < n_tiles; i += blockDim.x) {
    const int t = tile_task[i];
    atomicMin(&first[t], i);
    atomicAdd(&cnt[t], 1);
  }
  __syncthreads();
  // exclusive scan of ceil(cnt/2) over tasks (Hillis-Steele, double buffered)
  for (int t = threadIdx.x; t < n_tasks; t += blockDim.x) base[t] = (cnt[t] + 1) >> 1;
  __syncthreads();
  int* a = base;
  int* b = tmp;
  for (int d = 1; d < n_tasks; d <<= 1) {
    for (int t = threadIdx.x; t < n_tasks; t += blockDim.x) b[t] = a[t] + (t >= d ? a[t - d] : 0);
    __syncthreads();
    int* c = a; a = b; b = c;
  }
  const int total = n_tasks > 0 ? a[n_tasks - 1] : 0;     // inclusive sums in a[]
  for (int i = threadIdx.x; i < total; i += blockDim.x) pl.pair_tiles[i] = make_int2(-1, -1);
  __syncthreads();
  for (int i = threadIdx.x; i < n_tiles; i += blockDim.x) {
    const int t = tile_task[i];
    const int j = i - first[t];
    const int pr = a[t] - ((cnt[t] + 1) >> 1) + (j >> 1);
    if (j & 1) pl.pair_tiles[pr].y = i;
    else { pl.pair_tiles[pr].x = i; pl.pair_task[pr] = t; }
  }
  if (threadIdx.x == 0) pl.hdr[2] = total;
}

// mlong[slot][:] = sum_e long_nrm[e] * in[long_src[e]][:] and its abs-max, for every hub slot.  One CTA
// per slot at a time; thread groups of f_in/4 threads take edges round-robin (8 in flight each), partial
// sums are added in group order -> deterministic.
constexpr int LR_THREADS = 256;
__global__ void __launch_bounds__(LR_THREADS) long_rows_kernel(const float* __restrict__ in, int ld_in, int f_in,
                                                               Plan pl, float* __restrict__ mlong,
                                                               float* __restrict__ mlong_rowmax) {
  __shared__ __align__(16) float part[LR_THREADS * 4];
  __shared__ float wmax[LR_THREADS / 32];
  const int n_long = pl.hdr[0];
  const int tpr = f_in >> 2, G = LR_THREADS / tpr;
  const int g = threadIdx.x / tpr, cu = threadIdx.x - g * tpr;
  for (int slot = blockIdx.x; slot < n_long; slot += gridDim.x) {
    const int beg = pl.long_beg[slot], deg = pl.long_deg[slot];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g < G) {
      for (int e = g; e < deg; e += 8 * G) {
        float4 xv[8];
        float nn[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int ee = e + j * G;
          const bool ok = ee < deg;
          nn[j] = ok ? pl.long_nrm[beg + ee] : 0.f;
          xv[j] = ok ? ld_f4(in + (size_t)pl.long_src[beg + ee] * ld_in + 4 * cu) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc.x = fmaf(nn[j], xv[j].x, acc.x);
          acc.y = fmaf(nn[j], xv[j].y, acc.y);
          acc.z = fmaf(nn[j], xv[j].z, acc.z);
          acc.w = fmaf(nn[j], xv[j].w, acc.w);
        }
      }
      st_f4(part + threadIdx.x * 4, acc);
    }
    __syncthreads();
    float m = 0.f;
    if (g == 0) {
      for (int g2 = 1; g2 < G; ++g2) {
        const float4 v = ld_f4(part + (g2 * tpr + cu) * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      st_f4(mlong + (size_t)slot * f_in + 4 * cu, acc);
      m = fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w)));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < LR_THREADS / 32; ++i) m = fmaxf(m, wmax[i]);
      mlong_rowmax[slot] = m;
    }
  }
}

int g_pair_dbg = 0;

int stages_for(int K, int N) {
  const int w_bytes = 2 * K * N;
  int s = (SMEM_MAX - SMEM_FIXED - w_bytes) / STAGE_BYTES;
  return s > MAX_STAGES ? MAX_STAGES : s;
}

}  // namespace

bool gcn_layer_fwd_pair_supported(const GatherSrc& g, int f_out, const float* bias, int64_t b_task_stride,
                                  const float* relu_mask, const float* out, int ld_out, int n_tasks) {
  if (bias && (!aligned16(bias) || b_task_stride % 4 != 0)) return false;
  if (relu_mask && !aligned16(relu_mask)) return false;
  if (g.f_in % KCH != 0 || g.f_in < KCH || g.f_in > 1024) return false;
  if (f_out % 16 != 0 || f_out < 16 || f_out > 256) return false;
  if (g.ld_in % 4 != 0 || !aligned16(g.in) || g.ld_in < g.f_in) return false;
  if (ld_out % 4 != 0 || !aligned16(out)) return false;
  if (n_tasks > PT_MAXT) return false;
  return stages_for(g.f_in, f_out) >= 2;
}

int64_t gcn_layer_fwd_pair_workspace_bytes(int n_copies, int n_tiles, int n_tasks, int n_rows, int n_edges, int f_in,
                                           int f_out) {
  return carve(nullptr, n_copies, n_tiles, n_tasks, n_rows, n_edges, f_in, f_out).total;
}

int row_absmax(const float* x, int ld, int n_rows, int f, float* out, cudaStream_t stream) {
  if (n_rows == 0) return GMETA_OK;
  const int grid = ceil_div(n_rows, 8) < 16 * kNumSMs ? ceil_div(n_rows, 8) : 16 * kNumSMs;
  row_absmax_kernel<<<grid, 256, 0, stream>>>(x, ld, n_rows, f, out);
  return check_launch();
}

int gcn_layer_fwd_pair(const GatherSrc& g, const int32_t* tile_row0, const int32_t* tile_nrows,
                       const int32_t* tile_task, int n_tiles, int n_tasks, int n_copies, int n_rows, int n_edges,
                       const float* in_rowmax, const float* W, int64_t w_task_stride, int ldw, int trans_w,
                       const float* bias, int64_t b_task_stride, int f_out, int relu, const float* relu_mask,
                       float* out, int ld_out, float* out_rowmax, void* workspace, int64_t workspace_bytes,
                       cudaStream_t stream) {
  const int K = g.f_in, N = f_out;
  if (!in_rowmax || n_rows <= 0 || n_edges < 0) return GMETA_ERR_BAD_ARG;
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return GMETA_ERR_WORKSPACE;
  Workspace ws = carve(workspace, n_copies, n_tiles, n_tasks, n_rows, n_edges, K, N);
  if (workspace_bytes < ws.total) return GMETA_ERR_WORKSPACE;
  int rc;
  // header + weight abs-max are adjacent: one memset
  if (cudaMemsetAsync(ws.plan.hdr, 0, (size_t)(reinterpret_cast<char*>(ws.w_inv_scale) - reinterpret_cast<char*>(ws.plan.hdr)),
                      stream) != cudaSuccess)
    return GMETA_ERR_LAUNCH;
  {
    w_absmax_kernel<<<dim3(8, n_copies), 256, 0, stream>>>(W, w_task_stride, ldw, trans_w, K, N, ws.w_absmax);
    if ((rc = check_launch()) != GMETA_OK) return rc;
    const long long total = (long long)n_copies * (K / 8) * N;
    const int grid = (int)((total + 255) / 256 < 8 * kNumSMs ? (total + 255) / 256 : 8 * kNumSMs);
    pack_w_pair_kernel<<<grid, 256, 0, stream>>>(W, w_task_stride, ldw, trans_w, K, N, n_copies, ws.w_absmax,
                                                ws.w_image, 2LL * K * N, ws.w_inv_scale);
    if ((rc = check_launch()) != GMETA_OK) return rc;
  }
  {
    const int grid = ceil_div(n_rows, 256) < 8 * kNumSMs ? ceil_div(n_rows, 256) : 8 * kNumSMs;
    plan_classify_kernel<<<grid, 256, 0, stream>>>(g.indptr, g.indices, g.norm, g.in_row_map, g.dst_rows, n_rows,
                                                  ws.plan);
    if ((rc = check_launch()) != GMETA_OK) return rc;
    plan_long_edges_kernel<<<4 * kNumSMs, 256, 0, stream>>>(g.indptr, g.indices, g.norm, g.in_row_map, ws.plan);
    if ((rc = check_launch()) != GMETA_OK) return rc;
    pair_table_kernel<<<1, 1024, 0, stream>>>(tile_task, n_tiles, n_tasks, ws.plan);
    if ((rc = check_launch()) != GMETA_OK) return rc;
    long_rows_kernel<<<8 * kNumSMs, LR_THREADS, 0, stream>>>(g.in, g.ld_in, K, ws.plan, ws.mlong, ws.mlong_rowmax);
    if ((rc = check_launch()) != GMETA_OK) return rc;
  }
  PairParams p;
  p.in = g.in; p.ld_in = g.ld_in; p.f_in = K; p.in_rowmax = in_rowmax;
  p.mlong = ws.mlong; p.mlong_rowmax = ws.mlong_rowmax;
  p.rec = ws.plan.rec; p.hdr = ws.plan.hdr; p.pair_tiles = ws.plan.pair_tiles; p.pair_task = ws.plan.pair_task;
  p.tile_row0 = tile_row0; p.tile_nrows = tile_nrows; p.dst_rows = g.dst_rows; p.norm = g.norm;
  p.w_image = ws.w_image; p.image_task_stride = n_copies > 1 ? 2LL * K * N : 0; p.w_inv_scale = ws.w_inv_scale;
  p.bias = bias; p.b_task_stride = b_task_stride; p.f_out = N; p.relu = relu; p.relu_mask = relu_mask;
  p.out = out; p.ld_out = ld_out; p.out_rowmax = out_rowmax;
  p.n_stages = stages_for(K, N);
  p.w_bytes = 2 * K * N;   // bytes per CTA: K x N/2 halves, hi + lo
  p.dbg = g_pair_dbg;
  const size_t smem = (size_t)p.w_bytes + (size_t)p.n_stages * STAGE_BYTES + SMEM_FIXED;
  static int n_clusters = -1;
  if (n_clusters < 0) {
    if (cudaFuncSetAttribute(gcn_layer_fwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX) != cudaSuccess)
      return GMETA_ERR_LAUNCH;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNumSMs); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = SMEM_MAX;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, gcn_layer_fwd_pair_kernel, &cfg) != cudaSuccess || nc < 1) {
      cudaGetLastError();
      return GMETA_ERR_UNSUPPORTED;
    }
    n_clusters = nc < kNumSMs / 2 ? nc : kNumSMs / 2;
  }
  gcn_layer_fwd_pair_kernel<<<2 * n_clusters, NTHREADS, smem, stream>>>(p);
  return check_launch();
}

}  // namespace gmeta

// Debug ablation switches of the CTA-pair kernel for performance triage (results are WRONG with any flag set).
extern "C" void gmeta_debug_set_pair_flags(int flags) { gmeta::g_pair_dbg = flags; }
