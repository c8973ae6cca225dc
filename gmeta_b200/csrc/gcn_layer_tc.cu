// Fused GCN layer over a packed meta-batch -- Blackwell tensor-core implementation (sm_100a).
//
// One persistent CTA per SM, warp-specialised:
//   warps 0..15  gather producers: CSR neighbour gather + norm-weighted sum of the input rows for a
//                128-row tile, one 32-float K chunk at a time, written as an error-compensated
//                TF32 pair (hi = top 19 bits, lo = exact remainder) into 128B-swizzled K-major
//                shared-memory operand tiles;
//   warp 16      bulk-async (TMA) copies of the pre-split, pre-swizzled weight chunk of the
//                tile's TASK (per-task fast weights) into shared memory;
//   warp 17      one elected thread issues tcgen05.mma kind::tf32 -- three MMAs per K step
//                (hi*hi + lo*hi + hi*lo, "3xTF32") accumulating fp32 in tensor memory (TMEM);
//   warps 18..21 epilogue: tcgen05.ld of the accumulator rows, * norm[v] + bias, ReLU / mask,
//                16-byte stores; TMEM is double buffered so the epilogue of tile i overlaps the
//                MMAs of tile i+1.
// Stages are handed over with mbarriers (full/empty per ring stage, full/empty per accumulator).
// Replaces GraphConv.forward (reference G-Meta/learner.py:25-56) for K % 32 == 0 and
// N % 16 == 0, N <= 256; everything else takes the FFMA kernel in gcn_layer_simt.cu.
#include <cstdio>

#include "common.cuh"
#include "internal.cuh"

namespace gmeta {
namespace {

constexpr int TM = GMETA_TILE_ROWS;     // 128 = UMMA M
constexpr int KCH = 32;                 // floats per K chunk = one 128-byte swizzle row
constexpr int A_TILE_BYTES = TM * KCH * 4;   // 16 KB (hi or lo)
constexpr int N_PROD_WARPS = 16;
constexpr int WARP_TMA = 16;
constexpr int WARP_MMA = 17;
constexpr int WARP_EPI0 = 18;
constexpr int NTHREADS_TC = 22 * 32;
constexpr int MAX_STAGES = 4;
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = 256;           // columns per accumulator buffer
constexpr int PRE = 2;                  // neighbours per row whose (row, norm) stay in registers
constexpr int N_PROD_THREADS = N_PROD_WARPS * 32;
constexpr int LCAP = 512;               // long-edge records staged per window
constexpr int EPI_LD = 20;              // floats per row of the epilogue transpose tile (16 + pad, conflict-free)
constexpr int ST_TILES = 4;             // row tiles per super-tile (one prologue per super-tile)
constexpr int ST_ROWS = ST_TILES * TM;  // 512 = one row per producer thread
constexpr long long WAIT_TIMEOUT_CYCLES = 4000000000LL;  // ~2 s: trap instead of hanging the GPU

struct TcParams {
  GatherSrc g;
  const int32_t* tile_row0;
  const int32_t* tile_nrows;
  const int32_t* tile_task;
  int n_tiles;
  const float* w_image;      // [copies][K/32][hi|lo][N rows][32 floats, 128B swizzle]
  long long image_task_stride;  // floats between task copies (0 = shared weights)
  const float* bias;
  long long b_task_stride;
  int f_out;
  int relu;
  const float* relu_mask;
  float* out;
  int ld_out;
  int n_stages;
  int dbg;                   // debug ablation flags (0 in production): 1 skip output stores, 2 skip gather loads, 4 skip MMAs, 8 skip weight copies
  long long* prof;           // optional [gridDim.x][16] cycle counters per role (debug), or NULL
  int st_tiles;              // row tiles per super-tile (<= ST_TILES); 1 for small launches so that they spread over the SMs
  float* long_scratch;       // per CTA: [ST_ROWS][f_in] aggregated long (hub) rows + 4096 floats of slice partials, L2 resident
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) {
      printf("gmeta tc kernel: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, M=128
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory operand descriptor (cute::UMMA::SmemDescriptor layout):
// start address >> 4 [0,14), LBO (ignored for swizzled K-major, set 1) [16,30), SBO = 1024 B
// between 8-row groups [32,46), version 1 [46,48), layout SWIZZLE_128B = 2 at [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 at [4,6), a/b format
// TF32 = 2 at [7,10)/[10,13), a/b K-major (0) at 15/16, N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tf32_hi(float a) { return __uint_as_float(__float_as_uint(a) & 0xFFFFE000u); }

// producer-side bookkeeping in shared memory (after the pipeline stages and barriers), for one
// SUPER-TILE = ST_TILES consecutive row tiles whose structure-dependent prologue is done at once
struct alignas(16) ProdSmem {
  int src[LCAP];                        // source row offset (elements) of each staged long edge
  float nrm[LCAP];                      // its norm
  int beg[ST_ROWS];                     // first edge of each row
  int deg[ST_ROWS];
  int lpos[ST_ROWS + 1];                // start of each row in the flattened long-edge list
  int s_off[ST_ROWS][PRE];              // short rows: element offset of neighbour i's row
  float s_nrm[ST_ROWS][PRE];            //             and its norm (0 = absent / dropped)
  uint16_t row[LCAP];                   // super-tile row of each staged long edge
  uint8_t is_long[ST_ROWS];
  int wsum[N_PROD_WARPS];
  float bias_s[2][ACC_COLS];            // bias of the tile's task, double buffered by tile parity
  alignas(16) float epi[4][32 * EPI_LD];  // per epilogue warp: 32 rows x 16 columns, transposed on the way out
};

__device__ __forceinline__ void producer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N_PROD_THREADS) : "memory"); }

// every role walks the tiles in the same order: super-tiles round-robin over CTAs, tiles inside
#define TILE_LOOP_BEGIN                                                                        \
  for (int st_ = blockIdx.x; st_ < (p.n_tiles + p.st_tiles - 1) / p.st_tiles; st_ += gridDim.x)    \
    for (int tile = st_ * p.st_tiles; tile < min((st_ + 1) * p.st_tiles, p.n_tiles); ++tile) {
#define TILE_LOOP_END }

struct Smem {
  // dynamic shared memory, 1024-byte aligned: [stage][A_hi | A_lo | B_hi | B_lo], then barriers
  uint8_t* base;
  int stage_bytes;
  int b_bytes;  // one of B_hi / B_lo
  __device__ uint8_t* a_hi(int s) const { return base + (size_t)s * stage_bytes; }
  __device__ uint8_t* a_lo(int s) const { return a_hi(s) + A_TILE_BYTES; }
  __device__ uint8_t* b_hi(int s) const { return a_hi(s) + 2 * A_TILE_BYTES; }
};

__global__ void __launch_bounds__(NTHREADS_TC, 1) gcn_layer_fwd_tc_kernel(const TcParams p) {
  pdl_launch_dependents();     // programmatic dependent launch (common.cuh); the wait follows the set-up below
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.f_out;
  const int nkc = p.g.f_in / KCH;
  const int NS = p.n_stages;
  Smem sm;
  sm.base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B needs 1024-byte alignment
  sm.b_bytes = N * KCH * 4;
  sm.stage_bytes = 2 * A_TILE_BYTES + 2 * sm.b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm.base + (size_t)NS * sm.stage_bytes);
  // barrier slots: a_full[4] b_full[4] empty[4] acc_full[2] acc_empty[2]
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto b_full = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  auto empty = [&](int s) { return bar0 + 8u * (2 * MAX_STAGES + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (3 * MAX_STAGES + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (3 * MAX_STAGES + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);
  ProdSmem* ps = reinterpret_cast<ProdSmem*>(reinterpret_cast<uint8_t*>(bars) + 256);

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(a_full(s), N_PROD_WARPS);
      mbar_init(b_full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_TMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();       // barriers and tensor memory were set up under the tail of the kernel before this one

  if (warp < N_PROD_WARPS) {
    // ===================== gather producers =====================
    // Rows with at most PRE in-neighbours (97% of the rows of a 2-hop subgraph batch) are gathered
    // chunk by chunk by a quarter-warp each.  The few LONG rows (hubs: ~3% of rows, ~half of the
    // edges) would serialise a quarter-warp for hundreds of dependent loads, so they are
    // aggregated ONCE per tile over the full feature width by all 512 producer threads,
    // edge-parallel and balanced (contiguous slices of the flattened long-edge list per thread
    // group, partial sums of rows that straddle slices combined in slice order -> deterministic),
    // into an L2-resident per-CTA scratch row that the chunk loop then reads like a single
    // neighbour with weight 1.
    const int tid = threadIdx.x;
    const int q = warp * 4 + (lane >> 3);
    const int sub = lane & 7;
    const int f_in = p.g.f_in;
    const int tpr = f_in >> 2;                      // threads covering one full row (16 B each)
    const int G = N_PROD_THREADS / tpr;             // thread groups working on different edges
    const int g = tid / tpr, cu = tid - g * tpr;    // my group / my 16-byte column unit
    float* scratch = p.long_scratch + (size_t)blockIdx.x * ((size_t)ST_ROWS * f_in + 4 * N_PROD_THREADS * 2);
    float* part = scratch + (size_t)ST_ROWS * f_in; // [group][from-left | to-right][f_in] slice partials
    const int n_super = (p.n_tiles + p.st_tiles - 1) / p.st_tiles;
    int it = 0;
    long long t_pro = 0, t_wait = 0, t_body = 0, t_mark = clock64();
    auto lap = [&](long long& acc) { const long long now = clock64(); acc += now - t_mark; t_mark = now; };
    for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
      const int tile0 = st * p.st_tiles;
      const int nt = min(p.st_tiles, p.n_tiles - tile0);
      // ---- super-tile setup, one row per thread: extents, long flag, short-row neighbour records,
      //      positions in the flattened long-edge list ----
      int w = 0;
      {
        const int j = tid >> 7, r = tid & (TM - 1);
        int rb = 0, d = 0;
        if (j < nt && r < p.tile_nrows[tile0 + j]) {
          const int oi = p.tile_row0[tile0 + j] + r;
          const int v = p.g.dst_rows ? p.g.dst_rows[oi] : oi;
          if (p.g.identity) {              // pre-summed rows: row v's only "neighbour" is v itself, weight 1
            rb = v;
            d = 1;
          } else {
            rb = p.g.indptr[v];
            d = p.g.indptr[v + 1] - rb;
          }
        }
        const bool lg = d > PRE;
        ps->beg[tid] = rb;
        ps->deg[tid] = d;
        ps->is_long[tid] = lg ? 1 : 0;
        w = lg ? d : 0;
#pragma unroll
        for (int i = 0; i < PRE; ++i) {
          int off = 0;
          float nn = 0.f;
          if (!lg && i < d) {
            if (p.g.identity) {
              nn = 1.f;
              off = rb * p.g.ld_in;
            } else {
              const int u = p.g.indices[rb + i];
              const int srow = p.g.in_row_map ? p.g.in_row_map[u] : u;
              nn = srow < 0 ? 0.f : p.g.norm[u];        // negative map entry: neighbour dropped
              off = (srow < 0 ? 0 : srow) * p.g.ld_in;
            }
          }
          ps->s_off[tid][i] = off;
          ps->s_nrm[tid][i] = nn;
        }
        if (lg) {                                     // long row: one "neighbour" = its scratch row
          ps->s_off[tid][0] = tid * f_in;
          ps->s_nrm[tid][0] = 1.f;
        }
      }
      int incl = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v2 = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v2;
      }
      if (lane == 31) ps->wsum[warp] = incl;
      producer_sync();
      {
        int off = 0;
        for (int w2 = 0; w2 < warp; ++w2) off += ps->wsum[w2];
        ps->lpos[tid] = off + incl - w;
        if (tid == ST_ROWS - 1) ps->lpos[ST_ROWS] = off + incl;
      }
      producer_sync();
      const int nLE = ps->lpos[ST_ROWS];
      for (int wbeg = 0; wbeg < nLE; wbeg += LCAP) {
        const int wlen = min(LCAP, nLE - wbeg);
        // ---- edge records of this window, one per thread, coalesced ----
        for (int i = tid; i < wlen; i += N_PROD_THREADS) {
          const int pos = wbeg + i;
          int lo = 0, hi = ST_ROWS;                 // largest r with lpos[r] <= pos (a long row)
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (ps->lpos[mid] <= pos) lo = mid; else hi = mid;
          }
          const int u = p.g.indices[ps->beg[lo] + (pos - ps->lpos[lo])];
          const int srow = p.g.in_row_map ? p.g.in_row_map[u] : u;
          ps->nrm[i] = srow < 0 ? 0.f : p.g.norm[u];
          ps->src[i] = (srow < 0 ? 0 : srow) * p.g.ld_in;
          ps->row[i] = (uint16_t)lo;
        }
        producer_sync();
        // ---- balanced accumulation: group g owns list slice [sb, se) ----
        int own_r = -1;
        if (g < G) {
          const int per = (wlen + G - 1) / G;
          const int sb = min(g * per, wlen), se = min(sb + per, wlen);
          int pos = sb;
          while (pos < se) {
            const int r = ps->row[pos];
            const int rbeg = ps->lpos[r] - wbeg, rend = rbeg + ps->deg[r];
            const int wrb = max(rbeg, 0), wre = min(rend, wlen);   // the row's span inside this window
            const int run_end = min(se, wre);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int e = pos; e < run_end; e += 8) {
              float4 xv[8];
              float nn[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const bool ok = e + j < run_end;
                nn[j] = ok ? ps->nrm[e + j] : 0.f;
                xv[j] = ok ? ld_f4(p.g.in + ps->src[e + j] + 4 * cu) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                acc.x = fmaf(nn[j], xv[j].x, acc.x);
                acc.y = fmaf(nn[j], xv[j].y, acc.y);
                acc.z = fmaf(nn[j], xv[j].z, acc.z);
                acc.w = fmaf(nn[j], xv[j].w, acc.w);
              }
            }
            const bool from_left = pos > wrb, to_right = run_end < wre;
            if (!from_left && !to_right) {            // whole row (within this window) in my slice
              float* dst = scratch + r * f_in + 4 * cu;
              if (rbeg < 0) {                         // row began in an earlier window: accumulate
                const float4 old = __ldcg(reinterpret_cast<const float4*>(dst));
                acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
              }
              __stcg(reinterpret_cast<float4*>(dst), acc);
            } else {
              __stcg(reinterpret_cast<float4*>(part + ((g * 2 + (from_left ? 0 : 1)) * f_in + 4 * cu)), acc);
              if (!from_left) own_r = r;              // I own the row: it starts in my slice
            }
            pos = run_end;
          }
        }
        producer_sync();
        // ---- rows straddling slices: the owner adds the partials in slice order ----
        if (own_r >= 0) {
          const int per = (wlen + G - 1) / G;
          const int rbeg = ps->lpos[own_r] - wbeg;
          const int wre = min(rbeg + ps->deg[own_r], wlen);
          const int g_last = (wre - 1) / per;
          float4 acc = __ldcg(reinterpret_cast<const float4*>(part + ((g * 2 + 1) * f_in + 4 * cu)));
          for (int g2 = g + 1; g2 <= g_last; ++g2) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(part + ((g2 * 2 + 0) * f_in + 4 * cu)));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
          float* dst = scratch + own_r * f_in + 4 * cu;
          if (rbeg < 0) {
            const float4 old = __ldcg(reinterpret_cast<const float4*>(dst));
            acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
          }
          __stcg(reinterpret_cast<float4*>(dst), acc);
        }
        producer_sync();
      }
      lap(t_pro);
      // ---- the row tiles of the super-tile, chunk by chunk ----
      for (int j = 0; j < nt; ++j) {
        int src_off[2][PRE];       // element offset of the neighbour's row (+ this lane's 16-byte unit)
        float src_norm[2][PRE];
        const float* base0[2];     // base of neighbour 0: the input, or the scratch row of a long row
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int sr = j * TM + q + 64 * rr;         // row index inside the super-tile
          base0[rr] = ps->is_long[sr] ? scratch : p.g.in;
#pragma unroll
          for (int i = 0; i < PRE; ++i) {
            src_off[rr][i] = ps->s_off[sr][i] + 4 * sub;
            src_norm[rr][i] = ps->s_nrm[sr][i];
          }
        }
        // the feature segments of chunk kc+1 are requested before chunk kc is reduced and stored,
        // so their latency overlaps the barrier wait, the FMAs and the stores
        float4 xn[2][PRE];
        auto request = [&](int kc) {
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            xn[rr][0] = (src_norm[rr][0] != 0.f && !(p.dbg & 2))
                            ? __ldcg(reinterpret_cast<const float4*>(base0[rr] + src_off[rr][0] + kc * KCH))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 1; i < PRE; ++i)
              xn[rr][i] = (src_norm[rr][i] != 0.f && !(p.dbg & 2)) ? ld_f4(p.g.in + src_off[rr][i] + kc * KCH)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        request(0);
        for (int kc = 0; kc < nkc; ++kc, ++it) {
          const int s = it % NS;
          const uint32_t ph = (uint32_t)((it / NS) & 1);
          float4 x[2][PRE];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int i = 0; i < PRE; ++i) x[rr][i] = xn[rr][i];
          if (kc + 1 < nkc) request(kc + 1);
          lap(t_body);
          mbar_wait(empty(s), ph ^ 1u);
          lap(t_wait);
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const int r = q + 64 * rr;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < PRE; ++i) {
              acc.x = fmaf(src_norm[rr][i], x[rr][i].x, acc.x);
              acc.y = fmaf(src_norm[rr][i], x[rr][i].y, acc.y);
              acc.z = fmaf(src_norm[rr][i], x[rr][i].z, acc.z);
              acc.w = fmaf(src_norm[rr][i], x[rr][i].w, acc.w);
            }
            const float4 hi = make_float4(tf32_hi(acc.x), tf32_hi(acc.y), tf32_hi(acc.z), tf32_hi(acc.w));
            const float4 lo = make_float4(acc.x - hi.x, acc.y - hi.y, acc.z - hi.z, acc.w - hi.w);
            const int off = r * 128 + ((sub ^ (r & 7)) << 4);   // 128B swizzle: 16-byte unit ^ (row % 8)
            *reinterpret_cast<float4*>(sm.a_hi(s) + off) = hi;
            *reinterpret_cast<float4*>(sm.a_lo(s) + off) = lo;
          }
          fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(a_full(s));
        }
      }
      lap(t_body);
      producer_sync();   // ps / scratch of this super-tile are dead: the next prologue may overwrite them
      lap(t_wait);
    }
    if (p.prof && (tid == 0 || tid == 255)) {
      long long* o = p.prof + blockIdx.x * 16 + (tid == 0 ? 0 : 3);
      o[0] = t_pro; o[1] = t_wait; o[2] = t_body;
    }
  } else if (warp == WARP_TMA) {
    // ===================== weight chunk loader (bulk async copy) =====================
    if (lane == 0) {
      int it = 0;
      const uint32_t bytes = 2u * (uint32_t)sm.b_bytes;
      TILE_LOOP_BEGIN
        const float* img = p.w_image + (long long)p.tile_task[tile] * p.image_task_stride;
        for (int kc = 0; kc < nkc; ++kc, ++it) {
          const int s = it % NS;
          const uint32_t ph = (uint32_t)((it / NS) & 1);
          mbar_wait(empty(s), ph ^ 1u);
          if (p.dbg & 8) { mbar_arrive(b_full(s)); continue; }
          mbar_arrive_expect_tx(b_full(s), bytes);
          bulk_copy_g2s(smem_u32(sm.b_hi(s)), img + (size_t)kc * 2 * N * KCH, bytes, b_full(s));
        }
      TILE_LOOP_END
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(TM, N);
      int it = 0;
      int ti = 0;
      long long t_acc = 0, t_a = 0, t_b = 0, t_issue = 0, t_mark = clock64();
      auto lap = [&](long long& acc) { const long long now = clock64(); acc += now - t_mark; t_mark = now; };
      TILE_LOOP_BEGIN
        const int buf = ti & 1;
        lap(t_issue);
        mbar_wait(acc_empty(buf), (uint32_t)(((ti >> 1) & 1) ^ 1));
        lap(t_acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ACC_COLS);
        for (int kc = 0; kc < nkc; ++kc, ++it) {
          const int s = it % NS;
          const uint32_t ph = (uint32_t)((it / NS) & 1);
          lap(t_issue);
          mbar_wait(b_full(s), ph);
          lap(t_b);
          mbar_wait(a_full(s), ph);
          lap(t_a);
          tc_fence_after();
          const uint64_t da_hi = umma_desc_k_sw128(smem_u32(sm.a_hi(s)));
          const uint64_t da_lo = umma_desc_k_sw128(smem_u32(sm.a_lo(s)));
          const uint64_t db_hi = umma_desc_k_sw128(smem_u32(sm.b_hi(s)));
          const uint64_t db_lo = umma_desc_k_sw128(smem_u32(sm.b_hi(s) + sm.b_bytes));
#pragma unroll
          for (int k = 0; k < ((p.dbg & 4) ? 1 : KCH / 8); ++k) {     // UMMA_K = 8 tf32 = 32 bytes = 2 descriptor units
            const uint64_t adv = (uint64_t)(2 * k);
            tc_mma_tf32(d_tmem, da_hi + adv, db_hi + adv, idesc, (kc | k) != 0 ? 1u : 0u);
            tc_mma_tf32(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
            tc_mma_tf32(d_tmem, da_hi + adv, db_lo + adv, idesc, 1u);
          }
          tc_commit(empty(s));          // frees the stage once these MMAs have read it
        }
        tc_commit(acc_full(buf));       // accumulator complete -> epilogue
        ++ti;
      TILE_LOOP_END
      lap(t_issue);
      if (p.prof) {
        long long* o = p.prof + blockIdx.x * 16 + 6;
        o[0] = t_acc; o[1] = t_a; o[2] = t_b; o[3] = t_issue;
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;       // TMEM lanes 32*quarter .. +31 are the ones this warp may read
    const int r = quarter * 32 + lane;
    int ti = 0;
    long long t_wacc = 0, t_epi = 0, t_ldtm = 0, t_store = 0, t_mark = clock64();
    auto lap = [&](long long& acc) { const long long now = clock64(); acc += now - t_mark; t_mark = now; };
    TILE_LOOP_BEGIN
      const int buf = ti & 1;
      const int row0 = p.tile_row0[tile], nrows = p.tile_nrows[tile], task = p.tile_task[tile];
      const float* bias = p.bias ? p.bias + (long long)task * p.b_task_stride : nullptr;
      const bool live = r < nrows;
      const int oi = row0 + (live ? r : 0);                          // output row (compact or dense)
      const int v = p.g.dst_rows ? p.g.dst_rows[oi] : oi;            // real row: norm and mask
      const float nv = dst_norm(p.g, v);
      const float* mrow = p.relu_mask ? p.relu_mask + (size_t)((p.relu & 2) ? oi : v) * p.ld_out : nullptr;
      lap(t_epi);
      mbar_wait(acc_full(buf), (uint32_t)((ti >> 1) & 1));
      lap(t_wacc);
      tc_fence_after();
      float* stg = ps->epi[quarter];
      // bias of this tile's task -> shared memory once per tile (one coalesced load per lane), so the
      // chunk loop below has no dependent global load; double buffered + one 128-thread barrier
      float* bias_s = ps->bias_s[ti & 1];
      for (int c = (warp - WARP_EPI0) * 32 + lane; c < N; c += 128) bias_s[c] = bias ? bias[c] : 0.f;
      asm volatile("bar.sync 2, 128;" ::: "memory");
      // rows this lane stores after the transpose: 8*j + lane/4, 16-byte unit lane%4 of the chunk
      const int orow_base = p.tile_row0[tile] + quarter * 32 + (lane >> 2);
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * ACC_COLS);
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
        for (int j = 0; j < 16; ++j) {
          float val = fmaf(nv, __uint_as_float(acc[j]), bias_s[c0 + j]);
          if (p.relu & 1) val = relu_keep_nan(val);
          o[j] = val;
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
        // transpose through shared memory so that a store instruction writes 8 rows x 64 contiguous
        // bytes (whole 32-byte sectors) instead of 32 rows x 16 bytes
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; j += 4) st_f4(stg + lane * EPI_LD + j, make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]));
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rr = 8 * j + (lane >> 2);
          if (quarter * 32 + rr < nrows && !(p.dbg & 1)) {
            const float4 w4 = ld_f4(stg + rr * EPI_LD + 4 * (lane & 3));
            st_f4(p.out + (size_t)(orow_base + 8 * j) * p.ld_out + c0 + 4 * (lane & 3), w4);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      ++ti;
    TILE_LOOP_END
    lap(t_epi);
    if (p.prof && warp == WARP_EPI0 && lane == 0) {
      long long* o = p.prof + blockIdx.x * 16 + 10;
      o[0] = t_wacc; o[1] = t_epi; o[2] = t_ldtm; o[3] = t_store;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_TMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// raw fp32 weights (either orientation) -> per-copy UMMA image [K/32][hi|lo][N][32, swizzled]
__global__ void pack_w_umma_kernel(const float* __restrict__ W, long long w_stride, int ldw, int trans, int K,
                                   int N, int n_copies, float* __restrict__ image, long long image_stride) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  const long long total = (long long)n_copies * (K / KCH) * N * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int unit = (int)(i & 7);
    long long rest = i >> 3;
    const int n = (int)(rest % N); rest /= N;
    const int kc = (int)(rest % (K / KCH));
    const int c = (int)(rest / (K / KCH));
    const float* w = W + c * w_stride;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kc * KCH + unit * 4 + j;
      v[j] = trans ? w[(size_t)n * ldw + k] : w[(size_t)k * ldw + n];
    }
    const float4 hi = make_float4(tf32_hi(v[0]), tf32_hi(v[1]), tf32_hi(v[2]), tf32_hi(v[3]));
    const float4 lo = make_float4(v[0] - hi.x, v[1] - hi.y, v[2] - hi.z, v[3] - hi.w);
    float* chunk = image + c * image_stride + (size_t)kc * 2 * N * KCH;
    const int off = n * KCH + ((unit ^ (n & 7)) << 2);
    st_f4(chunk + off, hi);
    st_f4(chunk + (size_t)N * KCH + off, lo);
  }
}

// Inner SGD step on the per-task fast weights (meta.py:126,151) fused with the operand images of the updated weights
// for every tensor-core layer launch that will use them (forward orientation of each layer, transposed orientation
// for the data gradients): the images are recomputed from (w_in, grad), so the two halves of the kernel are
// independent and one launch replaces the update plus one split per layer call.
__global__ void sgd_pack_kernel(const float* __restrict__ w_in, long long w_in_stride, const float* __restrict__ grad,
                                float lr, int n_copies, int n_params, float* __restrict__ w_out, const TcPackPlan plan,
                                float* __restrict__ image, long long n_update, long long n_units) {
  pdl_prologue();     // programmatic dependent launch: see common.cuh
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_update + n_units; i += stride) {
    if (i < n_update) {
      const int t = (int)(i / n_params);
      const int p = (int)(i - (long long)t * n_params);
      // p - lr * g, rounded as torch does it (mul, then sub): meta.py:126
      w_out[i] = w_in[t * w_in_stride + p] - __fmul_rn(lr, grad[i]);
      continue;
    }
    long long r = i - n_update;
    int sidx = 0;
    long long per_copy = 0;
#pragma unroll
    for (int k = 0; k < 2 * GMETA_MAX_LAYERS; ++k)
      if (k < plan.n_seg) per_copy += (long long)(plan.seg[k].K / KCH) * plan.seg[k].N * 8;
    const int c = (int)(r / per_copy);
    r -= (long long)c * per_copy;
    while (true) {
      const long long n = (long long)(plan.seg[sidx].K / KCH) * plan.seg[sidx].N * 8;
      if (r < n) break;
      r -= n;
      ++sidx;
    }
    const TcPackSeg sg = plan.seg[sidx];
    const int unit = (int)(r & 7);
    long long rest = r >> 3;
    const int n = (int)(rest % sg.N);
    const int kc = (int)(rest / sg.N);
    const float* w = w_in + c * w_in_stride + sg.w_off;
    const float* gr = grad ? grad + (long long)c * n_params + sg.w_off : nullptr;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kc * KCH + unit * 4 + j;
      const long long e = sg.trans ? (long long)n * sg.ldw + k : (long long)k * sg.ldw + n;
      v[j] = gr ? w[e] - __fmul_rn(lr, gr[e]) : w[e];
    }
    const float4 hi = make_float4(tf32_hi(v[0]), tf32_hi(v[1]), tf32_hi(v[2]), tf32_hi(v[3]));
    const float4 lo = make_float4(v[0] - hi.x, v[1] - hi.y, v[2] - hi.z, v[3] - hi.w);
    float* chunk = image + c * plan.img_copy_stride + sg.img_off + (size_t)kc * 2 * sg.N * KCH;
    const int off = n * KCH + ((unit ^ (n & 7)) << 2);
    st_f4(chunk + off, hi);
    st_f4(chunk + (size_t)sg.N * KCH + off, lo);
  }
}

constexpr int kSmemFixed = 1024 /*alignment slack*/ + 256 /*barriers*/ + (int)sizeof(ProdSmem);

long long* g_tc_prof = nullptr;   // set through gmeta_debug_set_tc_profile
int g_tc_dbg = 0;                 // set through gmeta_debug_set_tc_flags

int stages_for(int N) {
  const int stage = 2 * A_TILE_BYTES + 2 * N * KCH * 4;
  int s = (227 * 1024 - kSmemFixed) / stage;
  return s > MAX_STAGES ? MAX_STAGES : s;
}

}  // namespace

bool gcn_layer_fwd_tc_supported(const GatherSrc& g, int ldw, int trans_w, int f_out, const float* out,
                                int ld_out) {
  (void)ldw; (void)trans_w;
  if (g.f_in % KCH != 0 || g.f_in < KCH || g.f_in > 2048) return false;
  if (f_out % 16 != 0 || f_out < 16 || f_out > 256) return false;
  if (g.ld_in % 4 != 0 || !aligned16(g.in) || g.ld_in < g.f_in) return false;
  if (ld_out % 4 != 0 || !aligned16(out)) return false;
  return true;
}

static int64_t image_bytes(int n_copies, int f_in, int f_out) {
  return ((int64_t)n_copies * 2 * f_in * f_out * (int64_t)sizeof(float) + 255) / 256 * 256;
}

int64_t gcn_layer_fwd_tc_scratch_bytes(int f_in) {
  return (int64_t)kNumSMs * ((int64_t)ST_ROWS * f_in + 4 * N_PROD_THREADS * 2) * (int64_t)sizeof(float);
}

int64_t gcn_layer_fwd_tc_workspace_bytes(int n_copies, int f_in, int f_out) {
  // weight image + per-CTA scratch rows for long (hub) rows
  return image_bytes(n_copies, f_in, f_out) + gcn_layer_fwd_tc_scratch_bytes(f_in);
}

int gcn_layer_fwd_tc(const GatherSrc& g, const int32_t* tile_row0, const int32_t* tile_nrows,
                     const int32_t* tile_task, int n_tiles, int n_copies, const float* W, int64_t w_task_stride,
                     int ldw, int trans_w, const float* bias, int64_t b_task_stride, int f_out, int relu,
                     const float* relu_mask, float* out, int ld_out, void* workspace, int64_t workspace_bytes,
                     const float* prepacked, int64_t prepacked_stride, cudaStream_t stream) {
  const int K = g.f_in, N = f_out;
  if (!workspace || !aligned16(workspace)) return GMETA_ERR_WORKSPACE;
  if (workspace_bytes < (prepacked ? gcn_layer_fwd_tc_scratch_bytes(K) : gcn_layer_fwd_tc_workspace_bytes(n_copies, K, N)))
    return GMETA_ERR_WORKSPACE;
  const float* image = prepacked ? prepacked : reinterpret_cast<float*>(workspace);
  const long long image_stride = prepacked ? prepacked_stride : 2LL * K * N;
  if (!prepacked) {
    const long long total = (long long)n_copies * (K / KCH) * N * 8;
    const int grid = (int)((total + 255) / 256 < 8 * kNumSMs ? (total + 255) / 256 : 8 * kNumSMs);
    launch_pdl(pack_w_umma_kernel, dim3(grid), dim3(256), 0, stream, W, w_task_stride, ldw, trans_w, K, N, n_copies,
                                                 reinterpret_cast<float*>(workspace), image_stride);
    int rc = check_launch();
    if (rc != GMETA_OK) return rc;
  }
  TcParams p;
  p.g = g;
  p.tile_row0 = tile_row0; p.tile_nrows = tile_nrows; p.tile_task = tile_task; p.n_tiles = n_tiles;
  p.w_image = image; p.image_task_stride = n_copies > 1 ? image_stride : 0;
  p.bias = bias; p.b_task_stride = b_task_stride; p.f_out = N; p.relu = relu; p.relu_mask = relu_mask;
  p.out = out; p.ld_out = ld_out;
  p.n_stages = stages_for(N);
  p.prof = g_tc_prof;
  p.dbg = g_tc_dbg;
  p.long_scratch = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + (prepacked ? 0 : image_bytes(n_copies, K, N)));
  const size_t smem = (size_t)p.n_stages * (2 * A_TILE_BYTES + 2 * N * KCH * 4) + kSmemFixed;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gcn_layer_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_done = true;
  }
  // small launches (the pruned meta-step's layers over a few thousand active rows): one tile per CTA, so that
  // they use as many SMs as they have tiles; large ones amortise the structure prologue over 4 tiles
  p.st_tiles = n_tiles >= 2 * ST_TILES * kNumSMs ? ST_TILES : (n_tiles >= 2 * kNumSMs ? 2 : 1);
  const int n_super = (n_tiles + p.st_tiles - 1) / p.st_tiles;
  const int grid = n_super < kNumSMs ? n_super : kNumSMs;
  launch_pdl(gcn_layer_fwd_tc_kernel, dim3(grid), dim3(NTHREADS_TC), smem, stream, p);
  return check_launch();
}

int gcn_tc_sgd_pack(const float* w_in, int64_t w_in_stride, const float* grad, float lr, int n_copies, int n_params,
                    float* w_out, const TcPackPlan& plan, float* image, cudaStream_t stream) {
  if (!w_in || n_copies <= 0 || n_params <= 0 || (grad && !w_out) || (plan.n_seg > 0 && !image)) return GMETA_ERR_BAD_ARG;
  long long per_copy = 0;
  for (int k = 0; k < plan.n_seg; ++k) per_copy += (long long)(plan.seg[k].K / KCH) * plan.seg[k].N * 8;
  const long long n_update = grad ? (long long)n_copies * n_params : 0;
  const long long n_units = per_copy * n_copies;
  if (n_update + n_units == 0) return GMETA_OK;
  const long long blocks = (n_update + n_units + 255) / 256;
  const int grid = (int)(blocks < 16 * kNumSMs ? blocks : 16 * kNumSMs);
  launch_pdl(sgd_pack_kernel, dim3(grid), dim3(256), 0, stream, w_in, w_in_stride, grad, lr, n_copies, n_params, w_out, plan, image,
                                            n_update, n_units);
  return check_launch();
}

}  // namespace gmeta

// Debug hook (not part of the reference-facing surface): device buffer of 148*16 int64 cycle
// counters filled by subsequent tensor-core layer launches; NULL switches it off.
extern "C" void gmeta_debug_set_tc_profile(long long* device_buffer) { gmeta::g_tc_prof = device_buffer; }
// Debug ablation switches for performance triage (results are WRONG with any flag set).
extern "C" void gmeta_debug_set_tc_flags(int flags) { gmeta::g_tc_dbg = flags; }
