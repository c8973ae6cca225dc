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
//   * Local subgraphs are star-like: most rows are leaves whose only in-neighbour is a hub they share
//     with hundreds of other leaves, so their aggregated rows M[v,:] = n_u * in[u,:] are IDENTICAL
//     (a C2 query set: 1.28 M rows, 0.17 M distinct (sources, norms) records per task).  The plan
//     therefore cuts the rows of a task into "compute tiles": a run of up to 1024 consecutive
//     output rows holding at most 128 distinct records ("slots").  Only the slots are gathered,
//     split and contracted on the tensor cores (one 128-row MMA tile per compute tile, 6.5x fewer
//     than rows / 128); the epilogue stages the accumulator block in shared memory and EXPANDS it:
//     out[v,:] = act(norm[v] * 2^-e * acc[slot(v),:] + bias), every output row still written once,
//     every input row still read once -- the same arithmetic per row as without the sharing, so
//     results are bit-identical to the undeduplicated kernel.
//
// Warp roles per CTA (896 threads = 28 warps; register budgets re-balanced per warpgroup with setmaxnreg: 112 / 24 / 64):
// warps 0..7 gather producers (a lane quad owns two slots, 32-float K chunks, the next chunk's segments requested before
// the current one is reduced: fp32 sum -> scaled FP16 hi/lo -> 64B-swizzled K-major operand stage), warp 12
// weight loader (cp.async.bulk of the pre-split, pre-swizzled image), warp 13 MMA issuer (leader CTA only; one
// thread), warps 8..11 and 16..27 epilogue (sixteen warps): per 64-column block the accumulators of the 128 slots go
// tcgen05.ld -> swizzled shared-memory block, then all sixteen warps expand it to the tile's output rows -- a lane
// octet per row, one row per iteration, 128 contiguous bytes per streaming store: * norm * 2^-e + bias, ReLU / mask,
// row abs-max for the next layer.  The launch before this kernel (hub pre-pass) and this one are chained with
// programmatic dependent launch: everything up to the first global read runs under the pre-pass tail.  Hand-offs are mbarriers waited on with the hardware
// suspend hint; the peer CTA signals the leader's barriers through the cluster address space, the MMA thread
// releases operand stages / accumulators in both CTAs with multicast commits.  Hub rows are summed beforehand
// by hub_prepass_kernel (whole chip, one warp per small hub, one CTA per big hub).
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace gmeta {
namespace {

using namespace ptx;

constexpr int TM = GMETA_TILE_ROWS;            // 128 rows per CTA tile; the pair's MMA has M = 256
constexpr int KCH = 32;                        // K elements per operand stage (64-byte swizzle rows of FP16)
constexpr int WCH = 64;                        // K elements per resident weight chunk (128-byte swizzle rows)
constexpr int A_HALF_BYTES = TM * 64;          // 8 KB: hi (or lo) operand tile of one chunk
constexpr int STAGE_BYTES = 2 * A_HALF_BYTES;  // 16 KB
// warp roles (register budgets are re-balanced per warpgroup with setmaxnreg)
constexpr int N_PROD_WARPS = 8;                // warps 0..7   gather producers
// warps 8..11: epilogue (first four of sixteen)
constexpr int WARP_LOAD = 12;                  // warp 12      weight loader (+ TMEM allocation)
constexpr int WARP_MMA = 13;                   // warp 13      MMA issuer (14, 15 idle)
constexpr int WARP_EPI0 = 16;                  // warps 16..27 epilogue as well (the expansion is latency-bound: 16 warps)
constexpr int N_EPI_WARPS = 16;
constexpr int N_EPI_THREADS = N_EPI_WARPS * 32;
constexpr int NTHREADS = 28 * 32;
constexpr int REGS_ENTRY = 72;                 // registers per thread at launch: 64K / 896 threads, rounded down to 8
constexpr int REGS_PROD = 112, REGS_CTRL = 24, REGS_EPI = 64;
// setmaxnreg re-distributes the registers the CTA got at launch (896 threads x 72 = 64512)
static_assert(8 * 32 * REGS_PROD + 4 * 32 * REGS_CTRL + N_EPI_WARPS * 32 * REGS_EPI <= NTHREADS * REGS_ENTRY,
              "register budget");
#ifndef GMETA_PAIR_PROF
#define GMETA_PAIR_PROF 0      // 1: per-role cycle counters (costs registers; debug builds only)
#endif
constexpr int MAX_STAGES = 6;
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = 256;
constexpr int PRE = 2;                         // in-neighbours per row the fused kernel gathers itself
#ifndef GMETA_RMAX
#define GMETA_RMAX 1024
#endif
constexpr int RMAX = GMETA_RMAX;                     // output rows per compute tile (<= 128 distinct slots among them)
constexpr int EBLK = 64;                       // accumulator columns per expansion block
constexpr int EPI_STAGE_BYTES = TM * EBLK * 4; // one [128 slots][64 columns] fp32 block (16-byte units XOR-swizzled by slot)
constexpr int BAR_BYTES = 256;
constexpr int SMEM_FIXED = BAR_BYTES /*barriers*/ + 4 * TM /*slot scale exponents*/ + ACC_COLS * 4 /*bias*/ +
                           RMAX * 8 /*row -> (slot, factor)*/ + RMAX * 4 /*row abs-max*/ + EPI_STAGE_BYTES;
#ifndef GMETA_GROUP_TILES
#define GMETA_GROUP_TILES 32
#endif
constexpr int GROUP_TILES = GMETA_GROUP_TILES;  // caller tiles (<= 4096 rows) one warp of the dedupe pass walks through
constexpr int GROUP_CAP = 80;                  // compute-tile entries reserved per group (<= 4096/97 + 32 forced closes)
#ifndef GMETA_PAIR_FIXED_COST
#define GMETA_PAIR_FIXED_COST 320
#endif
constexpr int PAIR_FIXED_COST = GMETA_PAIR_FIXED_COST;           // per-pair overhead of the cluster schedule, in output rows (measured: ~17k of ~61k cycles per tile)
constexpr int SMEM_MAX = 227 * 1024;
#ifndef GMETA_HUB_BIG
#define GMETA_HUB_BIG 64        // swept on the C2 query set (with GMETA_EXPAND_UNR 1): 32: 0.635 ms, 64: 0.605, 96: 0.612, 128: 0.620
#endif
constexpr int HUB_BIG = GMETA_HUB_BIG;                    // hubs with more (padded) edge records are aggregated by a whole CTA
constexpr int PT_MAXT = 2048;                  // tasks the pair-table kernel handles
constexpr int SCALE_TARGET = 13;               // scaled bound in [2^13, 2^14): 4x below the FP16 maximum
constexpr int SCALE_CLAMP = 100;

// ---- plan (graph structure + tiling only; built once per packed set and operand mapping) ----
struct PlanRec {      // 16 bytes
  int r0, r1;         // mapped source rows of in-neighbours 0/1; hub row: r0 = slot, r1 = -1
  float n0, n1;       // their norms (0 = absent or dropped)
};
struct CTile { int row0, nrows, nslots, task; };   // compute tile: output rows [row0, row0 + nrows), nslots <= 128 distinct records
struct PairEnt {      // 64 bytes; nrows[1] = nslots[1] = 0 when the task has an odd tile count
  int row0[2], nslots[2];     // first 16 bytes: what the gather producers need (slot s of a tile: srec[row0 + s])
  int nrows[2], task, cost0;  // cost0: schedule cost of all pairs before this one
  int pad[8];
};
struct Plan {
  int* hdr;           // [0] n_hubs  [1] n_hub_edges  [2] n_pairs  [3] n_big_hubs  [4] n_compute_tiles
  PlanRec* rec;       // [n_rows] record of every output row (input of the dedupe pass)
  PlanRec* srec;      // [n_rows] records of the slots of a compute tile at srec[tile.row0 + slot]
  uint8_t* row_slot;  // [n_rows] slot of every output row inside its compute tile
  CTile* ctiles;      // [n_groups * GROUP_CAP] compute tiles, per dedupe group
  int* group_nt;      // [n_groups] compute tiles of the group; after pair_table_kernel: exclusive prefix
  PairEnt* pairs;     // [cap_pairs] two compute tiles of the same task each
  int* cl_beg;        // [kNumSMs / 2 + 1] first pair of every cluster: contiguous runs of equal schedule cost
  int* hub_row;       // [cap_hub] real row of each hub slot
  int* hub_beg;       // [cap_hub] first record of the slot
  int* hub_deg;       // [cap_hub]
  int* big_list;      // [cap_big] slots of the hubs with more than HUB_BIG records (aggregated by a whole CTA)
  // hub edge list, grouped by tile then hub, every hub padded to a multiple of 4 records (pads: norm 0)
  int* hub_src;       // [cap_edges] mapped source row
  float* hub_nrm;     // [cap_edges]
  int64_t total;
};
struct Workspace {
  unsigned* w_absmax;   // [n_copies] bit pattern of max|W_c|
  float* w_inv_scale;   // [n_copies] 2^-e(c)
  __half* w_image;      // [n_copies][rank 2][K/64][hi|lo][N/2 rows][64 halves, 128B swizzle]
  float* mlong;         // [cap_hub][f_in] aggregated hub rows
  float* mlong_bound;   // [cap_hub] sum_e norm_e * max|in[src_e,:]| >= max|mlong[slot,:]|
  void* plan;           // plan built per call when the caller passes none
  int64_t total;
};

inline int64_t al(int64_t x) { return (x + 255) / 256 * 256; }
inline int n_groups_for(int n_tiles) { return (n_tiles + GROUP_TILES - 1) / GROUP_TILES; }
inline int cap_pairs_for(int n_tiles, int n_tasks) { return (n_groups_for(n_tiles) * GROUP_CAP + n_tasks) / 2 + 1; }
inline int cap_hub_for(int n_rows, int n_edges) {
  const int64_t by_edges = (int64_t)n_edges / (PRE + 1) + 1;
  return (int)(by_edges < n_rows ? by_edges : n_rows) + 1;
}

struct Carver {
  char* p;
  int64_t off;
  char* take(int64_t bytes) { char* q = p ? p + off : nullptr; off += al(bytes); return q; }
};

Plan carve_plan(void* base, int n_tiles, int n_tasks, int n_rows, int n_edges) {
  Plan pl;
  Carver c{reinterpret_cast<char*>(base), 0};
  const int cp = cap_pairs_for(n_tiles, n_tasks), ch = cap_hub_for(n_rows, n_edges), ng = n_groups_for(n_tiles);
  pl.hdr = reinterpret_cast<int*>(c.take(256));
  pl.rec = reinterpret_cast<PlanRec*>(c.take((int64_t)n_rows * 16));
  pl.srec = reinterpret_cast<PlanRec*>(c.take((int64_t)n_rows * 16));
  pl.row_slot = reinterpret_cast<uint8_t*>(c.take((int64_t)n_rows));
  pl.ctiles = reinterpret_cast<CTile*>(c.take((int64_t)ng * GROUP_CAP * 16));
  pl.group_nt = reinterpret_cast<int*>(c.take((int64_t)ng * 4));
  pl.pairs = reinterpret_cast<PairEnt*>(c.take((int64_t)cp * 64));
  pl.cl_beg = reinterpret_cast<int*>(c.take((kNumSMs / 2 + 1) * 4));
  pl.hub_row = reinterpret_cast<int*>(c.take((int64_t)ch * 4));
  pl.hub_beg = reinterpret_cast<int*>(c.take((int64_t)ch * 4));
  pl.hub_deg = reinterpret_cast<int*>(c.take((int64_t)ch * 4));
  pl.big_list = reinterpret_cast<int*>(c.take(((int64_t)n_edges / HUB_BIG + 2) * 4));
  const int64_t ce = (int64_t)n_edges + 3LL * ch + 64;
  pl.hub_src = reinterpret_cast<int*>(c.take(ce * 4));
  pl.hub_nrm = reinterpret_cast<float*>(c.take(ce * 4));
  pl.total = c.off;
  return pl;
}

Workspace carve_ws(void* base, int n_copies, int n_tiles, int n_tasks, int n_rows, int n_edges, int K, int N) {
  Workspace w;
  Carver c{reinterpret_cast<char*>(base), 0};
  const int ch = cap_hub_for(n_rows, n_edges);
  w.w_absmax = reinterpret_cast<unsigned*>(c.take((int64_t)n_copies * 4));
  w.w_inv_scale = reinterpret_cast<float*>(c.take((int64_t)n_copies * 4));
  w.w_image = reinterpret_cast<__half*>(c.take((int64_t)n_copies * 2 * K * N * 2));
  w.mlong = reinterpret_cast<float*>(c.take((int64_t)ch * K * 4));
  w.mlong_bound = reinterpret_cast<float*>(c.take((int64_t)ch * 4));
  w.plan = c.take(carve_plan(nullptr, n_tiles, n_tasks, n_rows, n_edges).total);
  w.total = c.off;
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
  const float* mlong_bound;
  const PlanRec* srec;           // slot records of the compute tiles
  const uint8_t* row_slot;       // output row -> slot of its compute tile
  const int* hdr;
  const PairEnt* pairs;
  const int* cl_beg;             // first pair of every cluster (kNumSMs / 2 clusters), or NULL = equal pair counts
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
  int dbg;                       // debug ablation flags: 1 skip output stores, 2 skip gather loads, 4 issue 1/2 of the MMAs
  long long* prof;               // optional [gridDim.x][16] cycle counters per role (debug), or NULL
};

__device__ __forceinline__ void st_f8(float* p, const float (&v)[8]) {   // 256-bit global store: one full sector per lane
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
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

__device__ __forceinline__ void ld8(float* d, const float* p) {      // 256-bit global load (32-byte aligned)
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]), "=f"(d[4]), "=f"(d[5]), "=f"(d[6]), "=f"(d[7])
               : "l"(p));
}
// One output row's gather state for one tile: where its (<= 2) sources live, their norms and abs-max.
struct RowCtx {
  const float* s0;
  const float* s1;
  float a0, a1, rm0, rm1;
  bool live;
  __device__ __forceinline__ void clear(const float* base) {
    s0 = s1 = base;
    a0 = a1 = rm0 = rm1 = 0.f;
    live = false;
  }
  __device__ __forceinline__ void decode(const int4 rc, bool lv, const PairParams& p, int sub) {
    clear(p.in);
    live = lv;
    if (!lv) return;
    const bool hub = rc.y < 0;
    a0 = __int_as_float(rc.z);
    a1 = __int_as_float(rc.w);
    if (a0 != 0.f) {
      s0 = (hub ? p.mlong + (size_t)rc.x * p.f_in : p.in + (size_t)rc.x * p.ld_in) + 8 * sub;
      rm0 = hub ? p.mlong_bound[rc.x] : p.in_rowmax[rc.x];
    }
    if (a1 != 0.f) {
      s1 = p.in + (size_t)rc.y * p.ld_in + 8 * sub;
      rm1 = p.in_rowmax[rc.y];
    }
  }
};
// buf[0..7] = 8 floats of source 0 at column offset `off`, buf[8..15] = of source 1 (zeros when absent)
__device__ __forceinline__ void gather_request(float* buf, const RowCtx& c, int off, int dbg) {
  if (c.a0 != 0.f && !(dbg & 2)) ld8(buf, c.s0 + off);
  else {
#pragma unroll
    for (int j = 0; j < 8; ++j) buf[j] = 0.f;
  }
  if (c.a1 != 0.f && !(dbg & 2)) ld8(buf + 8, c.s1 + off);
  else {
#pragma unroll
    for (int j = 8; j < 16; ++j) buf[j] = 0.f;
  }
}
// v = n0 * src0 + n1 * src1 over 8 columns -> scaled FP16 hi/lo
__device__ __forceinline__ void combine(const float* m, float n0, float n1, uint4& hi, uint4& lo) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = fmaf(n1, m[8 + j], n0 * m[j]);
  split8(v, hi, lo);
}

template <int R>
__device__ __forceinline__ void reg_set() {      // setmaxnreg for the executing warpgroup
  if constexpr (R > REGS_ENTRY) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R));
  else if constexpr (R < REGS_ENTRY) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R));
}

template <int VEC> struct VecLd;
template <> struct VecLd<4> {
  static __device__ __forceinline__ void ld(float (&d)[4], const float* p) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&d)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
  }
};
template <> struct VecLd<8> {
  static __device__ __forceinline__ void ld(float (&d)[8], const float* p) { ld8(d, p); }
  static __device__ __forceinline__ void st(float* p, const float (&d)[8]) { st_f8(p, d); }
};
template <> struct VecLd<2> {
  static __device__ __forceinline__ void ld(float (&d)[2], const float* p) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    d[0] = v.x; d[1] = v.y;
  }
  static __device__ __forceinline__ void st(float* p, const float (&d)[2]) {
    *reinterpret_cast<float2*>(p) = make_float2(d[0], d[1]);
  }
};

// output rows are written once and not read again by this launch: streaming (evict-first) stores
__device__ __forceinline__ void st_f4_stream(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Expansion of one staged accumulator block to the output rows of a compute tile (epilogue warps).  `stg` is the
// [TM slots][EBLK] block (16-byte units XOR-swizzled by slot), rinfo[r] = (slot, factor) of tile row r.  Octet `oct`
// of epilogue warp `ew` owns rows ew*4 + oct + 64*i; its lane u holds columns [c0, c0+4) and [c0+32, c0+36).
struct ExpandArgs {
  const float* stg;
  const float2* rinfo;
  float* rmax_s;
  const float* bias_s;
  float* out;            // + row0 * ld_out already applied
  float* out_rowmax;     // + row0 already applied (RMX)
  const float* mask;     // MASK: relu_mask base
  const int32_t* dst_rows;
  int ld_out, nrows, row0, n_cols, blk, nblk, relu, store;
};
template <bool MASK, bool RMX>
__device__ __forceinline__ void expand_block(const ExpandArgs& a, int ew, int lane) {
#ifndef GMETA_EXPAND_UNR
#define GMETA_EXPAND_UNR 1      // rows per octet and iteration (swept on the C2 query set: 1: 0.606 ms, 2: 0.626, 4: 0.669 per launch)
#endif
  constexpr int UNR = GMETA_EXPAND_UNR;
  constexpr int RSTEP = N_EPI_WARPS * 4;       // rows the epilogue warps cover per pass
  const int oct = lane >> 3, u = lane & 7;
  const int c0 = a.blk * EBLK + 4 * u;
  const bool ok0 = c0 < a.n_cols, ok1 = c0 + 32 < a.n_cols;
  const float4 b0 = ok0 ? ld_f4(a.bias_s + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b1 = ok1 ? ld_f4(a.bias_s + c0 + 32) : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool relu = (a.relu & 1) != 0, first = a.blk == 0, last = a.blk + 1 == a.nblk;
  const int nrows = a.nrows, ldo = a.ld_out;
  const bool st0 = ok0 && a.store, st1 = ok1 && a.store;
  const float* const stg_u = a.stg;
  float* const outp = a.out + c0;
  for (int rb = ew * 4 + oct; rb < nrows + oct; rb += RSTEP * UNR) {   // rb - oct < nrows: uniform trip count per warp
    float2 ri[UNR];
    float4 x0[UNR], x1[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      const int r = rb + RSTEP * j;
      ri[j] = a.rinfo[r < nrows ? r : 0];
    }
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      const int slot = __float_as_int(ri[j].x);
      const float* srow = stg_u + slot * EBLK + ((u ^ (slot & 7)) << 2);
      x0[j] = ld_f4(srow);
      x1[j] = ld_f4(srow + 32);
    }
    float mx[UNR];
#pragma unroll
    for (int j = 0; j < UNR; ++j) {
      const int r = rb + RSTEP * j;
      const bool live = r < nrows;
      const float f = ri[j].y;
      float4 w0 = make_float4(fmaf(f, x0[j].x, b0.x), fmaf(f, x0[j].y, b0.y), fmaf(f, x0[j].z, b0.z), fmaf(f, x0[j].w, b0.w));
      float4 w1 = make_float4(fmaf(f, x1[j].x, b1.x), fmaf(f, x1[j].y, b1.y), fmaf(f, x1[j].z, b1.z), fmaf(f, x1[j].w, b1.w));
      if (relu) {
        w0.x = relu_keep_nan(w0.x); w0.y = relu_keep_nan(w0.y); w0.z = relu_keep_nan(w0.z); w0.w = relu_keep_nan(w0.w);
        w1.x = relu_keep_nan(w1.x); w1.y = relu_keep_nan(w1.y); w1.z = relu_keep_nan(w1.z); w1.w = relu_keep_nan(w1.w);
      }
      if (MASK) {
        if (live) {
          const int oi = a.row0 + r;
          const int mr = (a.relu & 2) ? oi : (a.dst_rows ? a.dst_rows[oi] : oi);    // compact mask / mask by real row
          const float* mp = a.mask + (size_t)mr * ldo + c0;
          if (ok0) {
            const float4 m4 = ld_f4(mp);
            if (!(m4.x > 0.f)) w0.x = 0.f;
            if (!(m4.y > 0.f)) w0.y = 0.f;
            if (!(m4.z > 0.f)) w0.z = 0.f;
            if (!(m4.w > 0.f)) w0.w = 0.f;
          }
          if (ok1) {
            const float4 m4 = ld_f4(mp + 32);
            if (!(m4.x > 0.f)) w1.x = 0.f;
            if (!(m4.y > 0.f)) w1.y = 0.f;
            if (!(m4.z > 0.f)) w1.z = 0.f;
            if (!(m4.w > 0.f)) w1.w = 0.f;
          }
        }
      }
      float* o = outp + (size_t)r * ldo;
      if (live && st0) st_f4_stream(o, w0);
      if (live && st1) st_f4_stream(o + 32, w1);
      if (RMX) {
        float m = 0.f;
        if (ok0) m = fmaxf(fmaxf(fabsf(w0.x), fabsf(w0.y)), fmaxf(fabsf(w0.z), fabsf(w0.w)));
        if (ok1) m = fmaxf(m, fmaxf(fmaxf(fabsf(w1.x), fabsf(w1.y)), fmaxf(fabsf(w1.z), fabsf(w1.w))));
        mx[j] = m;
      }
    }
    if (RMX) {
      // row abs-max over the octet's lanes: the butterflies of the UNR rows are independent and overlap
#pragma unroll
      for (int o = 1; o < 8; o <<= 1)
#pragma unroll
        for (int j = 0; j < UNR; ++j) mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], o));
      if (u == 0) {          // the same lane owns row r in every block: no synchronisation needed
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
          const int r = rb + RSTEP * j;
          if (r < nrows) {
            float m = mx[j];
            if (!first) m = fmaxf(m, a.rmax_s[r]);
            if (!last) a.rmax_s[r] = m;
            else a.out_rowmax[r] = m;
          }
        }
      }
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
gcn_layer_fwd_pair_kernel(const PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int N = p.f_out, K = p.f_in;
  const int nkc = K / KCH;
  const int NS = p.n_stages;
  const int half_n_bytes = (N / 2) * 128;                 // one 64-wide chunk of W hi (or lo) in this CTA
  uint8_t* w_s = smem;
  uint8_t* a_s = smem + p.w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_s + (size_t)NS * STAGE_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };                    // leader's: producer warps of the pair
  auto empty = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };      // per CTA: multicast commit
  auto acc_full = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + b); };       // per CTA: multicast commit
  auto acc_empty = [&](int b) { return bar0 + 8u * (2 * MAX_STAGES + 2 + b); };  // leader's: 16 epilogue warps of the pair
  const uint32_t w_local = bar0 + 8u * (2 * MAX_STAGES + 4);             // per CTA: bulk copy landed
  const uint32_t w_ready = bar0 + 8u * (2 * MAX_STAGES + 5);             // leader's: both CTAs hold the task's weights
  const uint32_t w_free = bar0 + 8u * (2 * MAX_STAGES + 6);              // per CTA: MMAs of the previous task are done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES - 16);
  int8_t* scale_e = reinterpret_cast<int8_t*>(bars) + BAR_BYTES;               // [4][TM] per slot
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES + 4 * TM);   // [ACC_COLS]
  float2* rinfo_s = reinterpret_cast<float2*>(bias_s + ACC_COLS);        // [RMAX] (slot as int bits, output factor) of a tile's rows
  float* rmax_s = reinterpret_cast<float*>(rinfo_s + RMAX);              // [RMAX] running row abs-max over the column blocks
  float* epi_s = rmax_s + RMAX;                                          // [TM slots][EBLK floats]

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();   // SWIZZLE_128B operand tiles need a 1024-byte aligned base
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(a_full(s), 2 * N_PROD_WARPS);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 2 * N_EPI_WARPS);
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
  // everything above overlapped the tail of the launch before this one (the hub pre-pass, launched -- like this
  // kernel -- with programmatic serialization); its results and the weight images are read from here on
  pdl_wait();

  // contiguous run of tile pairs for this cluster: equal schedule cost (output rows + a per-pair constant) from the
  // plan when the grid is the full chip, equal pair counts otherwise
  const int n_pairs = p.hdr[2];
  const int n_cl = gridDim.x >> 1, cid = blockIdx.x >> 1;
  int p_beg, p_end;
  if (p.cl_beg && n_cl == kNumSMs / 2) {
    p_beg = p.cl_beg[cid];
    p_end = p.cl_beg[cid + 1];
  } else {
    const int ppc = (n_pairs + n_cl - 1) / n_cl;
    p_beg = cid * ppc < n_pairs ? cid * ppc : n_pairs;
    p_end = p_beg + ppc < n_pairs ? p_beg + ppc : n_pairs;
  }
  // the gather side of this CTA's compute tile in a pair entry: its slots live at srec[row0 + s], s < nslots
  // (the producers' "rows" below are slots: the distinct (sources, norms) records of the tile's output rows)
  auto ent_tile = [&](int pr, int& row0, int& nrows) {
    const int4 e = __ldg(reinterpret_cast<const int4*>(p.pairs + pr));
    row0 = rank ? e.y : e.x;
    nrows = rank ? e.w : e.z;
  };

  if (warp < N_PROD_WARPS) {
    // ===================== gather producers =====================
    reg_set<REGS_PROD>();
    // a lane quad owns tile rows rA = 8*warp + lane/4 and rA + 64; within a 32-float chunk lane `sub` holds floats
    // [8*sub, 8*sub + 8): the quad's loads cover one whole 128-byte line per source row
    const int rA = warp * 8 + (lane >> 2);
    const int sub = lane & 3;
    const int soffA = rA * 64 + ((sub ^ ((rA >> 1) & 3)) << 4);          // 64B swizzle: 16-byte unit ^ ((row / 2) % 4)
    const int soffB = soffA + 64 * 64;                                    // row + 64: same swizzle phase
#if GMETA_PAIR_PROF
    long long t_setup = 0, t_wait = 0, t_body = 0, t_mark = clock64();
#define PLAP(acc) { const long long now_ = clock64(); acc += now_ - t_mark; t_mark = now_; }
#else
#define PLAP(acc)
#endif
    RowCtx curA, curB, nxtA, nxtB;
    float buf0[32], buf1[32];               // two chunks: [rowA src0 | rowA src1 | rowB src0 | rowB src1] x 8 floats
    int ti = 0;
    int row0_n = 0, nrows_n = 0, row0_nn = 0, nrows_nn = 0;
    curA.clear(p.in);
    curB.clear(p.in);
    if (p_beg < p_end) {
      int row0, nrows;
      ent_tile(p_beg, row0, nrows);
      if (p_beg + 1 < p_end) ent_tile(p_beg + 1, row0_n, nrows_n);
      int4 rcA = make_int4(0, 0, 0, 0), rcB = make_int4(0, 0, 0, 0);
      if (rA < nrows) rcA = __ldg(reinterpret_cast<const int4*>(p.srec + row0 + rA));
      if (rA + 64 < nrows) rcB = __ldg(reinterpret_cast<const int4*>(p.srec + row0 + rA + 64));
      curA.decode(rcA, rA < nrows, p, sub);
      curB.decode(rcB, rA + 64 < nrows, p, sub);
      gather_request(buf0, curA, 0, p.dbg);
      gather_request(buf0 + 16, curB, 0, p.dbg);
    }
    float nA0 = 0.f, nA1 = 0.f, nB0 = 0.f, nB1 = 0.f;
    int st_i = 0;
    uint32_t st_ph = 0u;
    // one chunk: `mine` holds this chunk's segments, `other` receives the next chunk's (requested first,
    // so their latency overlaps the stage wait, the conversion and the stores)
    auto step = [&](float (&mine)[32], float (&other)[32], int kc) {
      const int s = st_i;
      const uint32_t ph = st_ph;
      if (++st_i == NS) { st_i = 0; st_ph ^= 1u; }
      if (kc + 1 < nkc) {
        gather_request(other, curA, (kc + 1) * KCH, p.dbg);
        gather_request(other + 16, curB, (kc + 1) * KCH, p.dbg);
      } else {                                                  // first chunk of the next tile
        gather_request(other, nxtA, 0, p.dbg);
        gather_request(other + 16, nxtB, 0, p.dbg);
      }
      PLAP(t_body);
      mbar_wait(empty(s), ph ^ 1u, 1);
      PLAP(t_wait);
      uint8_t* stage = a_s + (size_t)s * STAGE_BYTES;
      uint4 hi, lo;
      if (curA.live) {
        combine(mine, nA0, nA1, hi, lo);
        *reinterpret_cast<uint4*>(stage + soffA) = hi;
        *reinterpret_cast<uint4*>(stage + A_HALF_BYTES + soffA) = lo;
      }
      if (curB.live) {
        combine(mine + 16, nB0, nB1, hi, lo);
        *reinterpret_cast<uint4*>(stage + soffB) = hi;
        *reinterpret_cast<uint4*>(stage + A_HALF_BYTES + soffB) = lo;
      }
      // generic-proxy stores -> visible to the tensor core (async proxy).  The arrive is RELAXED: a release
      // would also wait for this thread's outstanding global loads (the next chunk is always in flight).
      if (!(p.dbg & 32)) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(a_full(s), 0);
    };
    for (int pr = p_beg; pr < p_end; ++pr, ++ti) {
      // scale of this tile's rows from the rigorous bound (see header comment)
      {
        const int eA = scale_exponent(curA.a0 * curA.rm0 + curA.a1 * curA.rm1);
        const int eB = scale_exponent(curB.a0 * curB.rm0 + curB.a1 * curB.rm1);
        nA0 = curA.a0 * exp2i(eA); nA1 = curA.a1 * exp2i(eA);
        nB0 = curB.a0 * exp2i(eB); nB1 = curB.a1 * exp2i(eB);
        if (sub == 0 && curA.live) scale_e[(ti & 3) * TM + rA] = (int8_t)eA;
        if (sub == 0 && curB.live) scale_e[(ti & 3) * TM + rA + 64] = (int8_t)eB;
      }
      // context of the next tile, fetched while this one streams: pair entry two tiles ahead, row records
      // at chunk 0, source abs-max at the middle chunk, first feature segments at the last chunk
      nrows_nn = 0;
      if (pr + 2 < p_end) ent_tile(pr + 2, row0_nn, nrows_nn);
      const bool liveA_n = rA < nrows_n, liveB_n = rA + 64 < nrows_n;
      int4 rcA_n = make_int4(0, 0, 0, 0), rcB_n = make_int4(0, 0, 0, 0);
      nxtA.clear(p.in);
      nxtB.clear(p.in);
      PLAP(t_setup);
      for (int kc = 0; kc < nkc; kc += 2) {
        if (kc == 0) {
          if (liveA_n) rcA_n = __ldg(reinterpret_cast<const int4*>(p.srec + row0_n + rA));
          if (liveB_n) rcB_n = __ldg(reinterpret_cast<const int4*>(p.srec + row0_n + rA + 64));
        }
        if (kc == ((nkc >> 2) << 1)) {
          nxtA.decode(rcA_n, liveA_n, p, sub);
          nxtB.decode(rcB_n, liveB_n, p, sub);
        }
        step(buf0, buf1, kc);
        step(buf1, buf0, kc + 1);
      }
      curA = nxtA;
      curB = nxtB;
      row0_n = row0_nn; nrows_n = nrows_nn;
      PLAP(t_body);
    }
#if GMETA_PAIR_PROF
    if (p.prof && (threadIdx.x == 0 || threadIdx.x == 255)) {
      long long* o = p.prof + blockIdx.x * 16 + (threadIdx.x == 0 ? 0 : 3);
      o[0] = t_setup; o[1] = t_wait; o[2] = t_body;
    }
#endif
  } else if (warp >= WARP_LOAD && warp < WARP_EPI0) {
    reg_set<REGS_CTRL>();
    if (warp == WARP_LOAD) {
      // ===================== weight loader: one bulk copy per task change =====================
      if (lane == 0) {
        int cur = -1, n = 0;
        for (int pr = p_beg; pr < p_end; ++pr) {
          const int task = p.pairs[pr].task;
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
        int ti = 0, cur = -1, nw = 0;
        int st_i = 0;
        uint32_t st_ph = 0u;
#if GMETA_PAIR_PROF
        long long t_w = 0, t_acc = 0, t_a = 0, t_issue = 0, t_mark = clock64();
#define MLAP(acc) { const long long now_ = clock64(); acc += now_ - t_mark; t_mark = now_; }
#else
#define MLAP(acc)
#endif
        int task_n = p_beg < p_end ? p.pairs[p_beg].task : 0;
        for (int pr = p_beg; pr < p_end; ++pr, ++ti) {
          const int task = task_n;
          if (pr + 1 < p_end) task_n = p.pairs[pr + 1].task;
          MLAP(t_issue);
          if (task != cur) {
            mbar_wait_cluster(w_ready, (uint32_t)(nw & 1), 4);
            cur = task;
            ++nw;
          }
          MLAP(t_w);
          const int buf = ti & 1;
          mbar_wait_cluster(acc_empty(buf), (uint32_t)(((ti >> 1) & 1) ^ 1), 5);
          MLAP(t_acc);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ACC_COLS);
          for (int kc = 0; kc < nkc; ++kc) {
            const int s = st_i;
            const uint32_t ph = st_ph;
            if (++st_i == NS) { st_i = 0; st_ph ^= 1u; }
            MLAP(t_issue);
            mbar_wait_cluster(a_full(s), ph, 6);
            MLAP(t_a);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(a_s + (size_t)s * STAGE_BYTES);
            const uint32_t b_addr = smem_u32(w_s + (size_t)(kc >> 1) * 2 * half_n_bytes);
            const uint64_t da_hi = umma_desc_k_sw64(a_addr);
            const uint64_t da_lo = umma_desc_k_sw64(a_addr + A_HALF_BYTES);
            const uint64_t db_hi = umma_desc_k_sw128(b_addr);
            const uint64_t db_lo = umma_desc_k_sw128(b_addr + half_n_bytes);
#pragma unroll
            for (int k = 0; k < KCH / 16; ++k) {        // UMMA_K = 16 halves = 32 bytes = 2 descriptor units
              if ((p.dbg & 4) && k) break;
              const uint64_t adv_a = (uint64_t)(2 * k);
              const uint64_t adv_b = (uint64_t)(2 * ((kc & 1) * 2 + k));
              tc_mma_f16_pair(d_tmem, da_hi + adv_a, db_hi + adv_b, idesc, (kc | k) != 0 ? 1u : 0u);
              tc_mma_f16_pair(d_tmem, da_lo + adv_a, db_hi + adv_b, idesc, 1u);
              tc_mma_f16_pair(d_tmem, da_hi + adv_a, db_lo + adv_b, idesc, 1u);
            }
            if (!(p.dbg & 64)) tc_commit_pair(empty(s));          // frees the stage in both CTAs once these MMAs have read it
          }
          tc_commit_pair(acc_full(buf));       // accumulators complete -> both epilogues
          if (pr + 1 < p_end && task_n != task) tc_commit_pair(w_free);
        }
        MLAP(t_issue);
#if GMETA_PAIR_PROF
        if (p.prof) {
          long long* o = p.prof + blockIdx.x * 16 + 6;
          o[0] = t_w; o[1] = t_acc; o[2] = t_a; o[3] = t_issue;
        }
#endif
      }
    }
  } else {
    // ===================== epilogue (16 warps): accumulator slots -> output rows =====================
    // Warp w may read TMEM lanes 32*(w%4)..+31 (= slots): the four warps of a lane quarter take 16 of a block's 64
    // columns each.  Per 64-column block: tcgen05.ld -> raw accumulators into the [128 slots][64 columns]
    // shared-memory block (16-byte units XOR-swizzled by slot) -> named barrier -> all sixteen warps expand the
    // block to the tile's output rows: a lane octet per row reads its slot's 256 bytes, * norm[v] * 2^-e(slot) *
    // 2^-e(W) + bias, ReLU / mask, running row abs-max, two 128-byte row segments per octet.  The expansion is
    // bound by shared-memory and store latency, not by issue slots or bytes: hence 16 warps with 64 registers.
    reg_set<REGS_EPI>();
    const int quarter = warp & 3;
    const int cpart = warp < WARP_EPI0 ? 0 : 1 + ((warp - WARP_EPI0) >> 2);   // which 16 columns of a block: 0..3
    const int ew = cpart * 4 + quarter;      // epilogue warp 0..15
    const int et = ew * 32 + lane;           // index among the 512 epilogue threads
    const int slot_l = quarter * 32 + lane;  // staging: the accumulator row (slot) this lane drains
    const int nblk = (N + EBLK - 1) / EBLK;
    int ti = 0, bias_task = -1;
#if GMETA_PAIR_PROF
    long long t_wacc = 0, t_epi = 0, t_stage = 0, t_expand = 0, t_mark = clock64();
#define ELAP(acc) { const long long now_ = clock64(); acc += now_ - t_mark; t_mark = now_; }
#else
#define ELAP(acc)
#endif
    for (int pr = p_beg; pr < p_end; ++pr, ++ti) {
      const int buf = ti & 1;
      const int4 e0 = __ldg(reinterpret_cast<const int4*>(p.pairs + pr));
      const int4 e1 = __ldg(reinterpret_cast<const int4*>(p.pairs + pr) + 1);
      const int row0 = rank ? e0.y : e0.x, nrows = rank ? e1.y : e1.x, task = e1.z;
      // row -> (slot, norm) of this tile's output rows, requested before the accumulators are waited for
      int rs[RMAX / N_EPI_THREADS];
      float rn[RMAX / N_EPI_THREADS];
#pragma unroll
      for (int j = 0; j < RMAX / N_EPI_THREADS; ++j) {
        const int r = et + N_EPI_THREADS * j;
        rs[j] = 0;
        rn[j] = 0.f;
        if (r < nrows) {
          rs[j] = (int)__ldg(p.row_slot + row0 + r);
          rn[j] = p.norm[p.dst_rows ? p.dst_rows[row0 + r] : row0 + r];
        }
      }
      const float wis = p.w_inv_scale[p.image_task_stride ? task : 0];
      // every epilogue warp is done with the previous tile's row table, bias and staging blocks
      asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_THREADS) : "memory");
      if (task != bias_task) {     // the task's bias -> shared memory, once per task
        bias_task = task;
        const float* bias = p.bias ? p.bias + (long long)task * p.b_task_stride : nullptr;
        if (et < ACC_COLS) bias_s[et] = (bias && et < N) ? bias[et] : 0.f;
      }
      ELAP(t_epi);
      mbar_wait(acc_full(buf), (uint32_t)((ti >> 1) & 1), 7);
      ELAP(t_wacc);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < RMAX / N_EPI_THREADS; ++j) {
        const int r = et + N_EPI_THREADS * j;
        if (r < nrows)
          rinfo_s[r] = make_float2(__int_as_float(rs[j]), rn[j] * exp2i(-(int)scale_e[(ti & 3) * TM + rs[j]]) * wis);
      }
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * ACC_COLS + 16 * cpart);
      ExpandArgs xa;
      xa.stg = epi_s; xa.rinfo = rinfo_s; xa.rmax_s = rmax_s; xa.bias_s = bias_s;
      xa.out = p.out + (size_t)row0 * p.ld_out;
      xa.out_rowmax = p.out_rowmax ? p.out_rowmax + row0 : nullptr;
      xa.mask = p.relu_mask; xa.dst_rows = p.dst_rows;
      xa.ld_out = p.ld_out; xa.nrows = nrows; xa.row0 = row0; xa.n_cols = N; xa.nblk = nblk; xa.relu = p.relu;
      xa.store = !(p.dbg & 1);
      ELAP(t_epi);
      for (int blk = 0; blk < nblk; ++blk) {
        const bool mine = blk * EBLK + 16 * cpart < N;    // this warp's 16 columns of the block exist
        if (blk) asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_THREADS) : "memory");   // the previous block has been expanded by everybody
        if (mine) {
          uint32_t acc[16];
          tmem_ld16(t_addr + (uint32_t)(blk * EBLK), acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int x = 4 * cpart + j;      // logical 16-byte unit of the 64-column row
            st_f4(epi_s + slot_l * EBLK + (((x & 8) | ((x & 7) ^ (slot_l & 7))) << 2),
                  make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                              __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3])));
          }
        }
        if (blk + 1 == nblk) tc_fence_before();           // the tile's accumulators have been read: hand the buffer back
        asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_THREADS) : "memory");
        if (blk + 1 == nblk && lane == 0) mbar_arrive_cluster(acc_empty(buf), 0);
        ELAP(t_stage);
        xa.blk = blk;
        if (p.relu_mask) {
          if (p.out_rowmax) expand_block<true, true>(xa, ew, lane);
          else expand_block<true, false>(xa, ew, lane);
        } else {
          if (p.out_rowmax) expand_block<false, true>(xa, ew, lane);
          else expand_block<false, false>(xa, ew, lane);
        }
        ELAP(t_expand);
      }
    }
    ELAP(t_epi);
#if GMETA_PAIR_PROF
    if (p.prof && warp == WARP_EPI0 && lane == 0) {
      long long* o = p.prof + blockIdx.x * 16 + 10;
      o[0] = t_wacc; o[1] = t_epi; o[2] = t_stage; o[3] = t_expand;
    }
#endif
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
// One CTA per weight copy: abs-max of the copy (block reduction), then its scaled FP16 hi/lo image.
__global__ void __launch_bounds__(1024) pack_w_pair_kernel(const float* __restrict__ W, long long w_stride, int ldw,
                                                           int trans, int K, int N, __half* __restrict__ image,
                                                           long long image_stride, float* __restrict__ w_inv_scale) {
  __shared__ float red[32];
  __shared__ float s_max;
  // wait BEFORE releasing the dependents: the hub pre-pass behind this kernel runs beside it (it does not read the
  // weight images) and waits only at its end, so it must not start before the producer of `in` has completed
  pdl_wait();
  pdl_launch_dependents();
  const int c = blockIdx.x;
  const float* w = W + c * w_stride;
  const int inner = trans ? K : N, total = K * N;
  float m = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) m = fmaxf(m, fabsf(w[(size_t)(i / inner) * ldw + (i % inner)]));
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) s_max = m;
  }
  __syncthreads();
  const int e = scale_exponent(s_max);
  const float sc = exp2i(e);
  if (threadIdx.x == 0) w_inv_scale[c] = exp2i(-e);
  const int ku = K / 8, nwc = K / WCH, hn = N / 2;
  for (int i = threadIdx.x; i < ku * N; i += blockDim.x) {
    const int n = i % N, u = i / N;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = u * 8 + j;
      v[j] = sc * (trans ? w[(size_t)n * ldw + k] : w[(size_t)k * ldw + n]);
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    const int rank = n / hn, nn = n - rank * hn, kc = u >> 3, unit = u & 7;
    __half* dst = image + c * image_stride + ((((size_t)rank * nwc + kc) * 2) * hn + nn) * 64 + ((unit ^ (nn & 7)) << 3);
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + (size_t)hn * 64) = lo;
  }
}

// per-row abs-max of a row-major matrix (one warp per row)
__global__ void row_absmax_kernel(const float* __restrict__ x, int ld, int n_rows, int f, float* __restrict__ out) {
  pdl_prologue();
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
// hub rows: mlong[slot][:] = sum_e nrm[e] * in[src[e]][:] and the bound sum_e nrm[e] * max|in[src[e],:]|
// Runs before the fused kernel with the whole chip's memory-level parallelism (a hub row has up to
// ~1000 in-neighbours: half of all edges of a 2-hop subgraph batch sit in ~3% of its rows); the fused
// kernel then reads a hub like a single neighbour with weight 1.  Summation order is fixed.
// ------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void hub_accumulate(const float* __restrict__ in, int ld_in, const float* __restrict__ in_rowmax,
                                               const int* __restrict__ src, const float* __restrict__ nrm, int n,
                                               int lane, float (&acc)[VEC], float& bacc, bool pre = false, int s_pre = 0,
                                               float n_pre = 0.f) {
  // records [0, n) of one hub (n a multiple of 4), 32 at a time; lane covers columns [lane*VEC, lane*VEC + VEC).
  // `pre`: the first block's record of this lane (source, norm) was fetched ahead by the caller.
  for (int b = 0; b < n; b += 32) {
    const int idx = b + lane;
    int s_l = 0;
    float n_l = 0.f;
    if (b == 0 && pre) { s_l = s_pre; n_l = n_pre; }
    else if (idx < n) { s_l = src[idx]; n_l = nrm[idx]; }
    if (n_l != 0.f) bacc += n_l * in_rowmax[s_l];
    const int cnt = n - b < 32 ? n - b : 32;
    for (int j0 = 0; j0 < cnt; j0 += 8) {
      float x[8][VEC], w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        w[j] = __shfl_sync(0xffffffffu, n_l, (j0 + j) & 31);
        const int sj = __shfl_sync(0xffffffffu, s_l, (j0 + j) & 31);
        if (j0 + j < cnt && w[j] != 0.f) VecLd<VEC>::ld(x[j], in + (size_t)sj * ld_in + lane * VEC);
        else {
          w[j] = 0.f;
#pragma unroll
          for (int k = 0; k < VEC; ++k) x[j][k] = 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = fmaf(w[j], x[j][k], acc[k]);
    }
  }
}

// One launch for all hub rows.  CTAs [0, HUB_BIG_CTAS): one CTA (8 warps) per hub with more than HUB_BIG
// records -- warp w sums the 32-record blocks w, w+8, ..., the partials are added in warp order; the other
// CTAs: one warp per smaller hub.  The big hubs are scheduled first (launch order) and the small ones fill in
// behind them, so the launch has one tail instead of two.  (Interleaving the two kinds was measured slower.)
#ifndef GMETA_HUB_BIG_CTAS
#define GMETA_HUB_BIG_CTAS 3
#endif
constexpr int HUB_BIG_CTAS = GMETA_HUB_BIG_CTAS * kNumSMs;
constexpr int HUB_SMALL_CTAS = 8 * kNumSMs;
template <int VEC>
__global__ void __launch_bounds__(256) hub_prepass_kernel(const float* __restrict__ in, int ld_in,
                                                          const float* __restrict__ in_rowmax, const int* __restrict__ hdr,
                                                          const int* __restrict__ big_list, const int* __restrict__ hub_beg,
                                                          const int* __restrict__ hub_deg, const int* __restrict__ hub_src,
                                                          const float* __restrict__ hub_nrm, float* __restrict__ mlong,
                                                          float* __restrict__ mlong_bound) {
  constexpr int K = 32 * VEC;
  __shared__ float part[8][K];
  __shared__ float bpart[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Launched with programmatic serialization behind the weight split, which it does not depend on: the two run side
  // by side; the layer kernel behind it sets up (barriers, tensor memory) as soon as all CTAs of this grid are
  // running.  The wait at the end makes "this grid complete" imply "weight split complete" for the layer kernel's wait.
  pdl_launch_dependents();
  if (blockIdx.x >= HUB_BIG_CTAS) {
    const int n_hub = hdr[0];
    const int nw = (gridDim.x - HUB_BIG_CTAS) * 8;
    // A small hub is a chain of three dependent loads (degree / first record -> its records -> the rows) for 3..128
    // records of work, so the chain is started ahead: (degree, first record) two hubs ahead, the first 32 records one
    // hub ahead -- only the row loads of the current hub are waited for.
    const int slot0 = (blockIdx.x - HUB_BIG_CTAS) * 8 + warp;
    auto meta = [&](int slot, int& n, int& beg) {
      n = 0; beg = 0;
      if (slot < n_hub) { n = (hub_deg[slot] + 3) & ~3; beg = hub_beg[slot]; }
    };
    auto first_records = [&](int n, int beg, int& s_l, float& n_l) {
      s_l = 0; n_l = 0.f;
      if (n <= HUB_BIG && lane < n) { s_l = hub_src[beg + lane]; n_l = hub_nrm[beg + lane]; }
    };
    int n1, beg1, n2, beg2, s1;
    float w1;
    meta(slot0, n1, beg1);
    meta(slot0 + nw, n2, beg2);
    first_records(n1, beg1, s1, w1);
    for (int slot = slot0; slot < n_hub; slot += nw) {
      const int n = n1, beg = beg1, s_cur = s1;
      const float w_cur = w1;
      n1 = n2; beg1 = beg2;
      first_records(n1, beg1, s1, w1);
      meta(slot + 2 * nw, n2, beg2);
      if (n > HUB_BIG) continue;
      float acc[VEC], bacc = 0.f;
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
      hub_accumulate<VEC>(in, ld_in, in_rowmax, hub_src + beg, hub_nrm + beg, n, lane, acc, bacc, true, s_cur, w_cur);
      VecLd<VEC>::st(mlong + (size_t)slot * K + lane * VEC, acc);
#pragma unroll
      for (int o = 16; o; o >>= 1) bacc += __shfl_xor_sync(0xffffffffu, bacc, o);
      if (lane == 0) mlong_bound[slot] = bacc;
    }
    pdl_wait();
    return;
  }
  const int n_big = hdr[3];
  for (int i = blockIdx.x; i < n_big; i += HUB_BIG_CTAS) {
    const int slot = big_list[i];
    const int n = (hub_deg[slot] + 3) & ~3, beg = hub_beg[slot];
    float acc[VEC], bacc = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    for (int b = warp * 32; b < n; b += 256) {
      const int m = n - b < 32 ? n - b : 32;
      hub_accumulate<VEC>(in, ld_in, in_rowmax, hub_src + beg + b, hub_nrm + beg + b, m, lane, acc, bacc);
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) part[warp][lane * VEC + k] = acc[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) bacc += __shfl_xor_sync(0xffffffffu, bacc, o);
    if (lane == 0) bpart[warp] = bacc;
    __syncthreads();
    if (threadIdx.x < K) {
      float v = part[0][threadIdx.x];
#pragma unroll
      for (int w = 1; w < 8; ++w) v += part[w][threadIdx.x];
      mlong[(size_t)slot * K + threadIdx.x] = v;
    }
    if (threadIdx.x == 0) {
      float v = bpart[0];
      for (int w = 1; w < 8; ++w) v += bpart[w];
      mlong_bound[slot] = v;
    }
    __syncthreads();
  }
  pdl_wait();
}

template <int VEC>
int hub_prepass_launch(const float* in, int ld_in, const float* in_rowmax, const Plan& pl, float* mlong,
                       float* mlong_bound, cudaStream_t stream) {
  // programmatic serialization (always, unlike the library-wide switch of common.cuh: measured 0.71 -> 0.665 ms per
  // C2 layer launch): runs beside the weight split in front of it, the layer kernel's set-up runs under its tail
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(HUB_BIG_CTAS + HUB_SMALL_CTAS); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, hub_prepass_kernel<VEC>, in, ld_in, in_rowmax, (const int*)pl.hdr, (const int*)pl.big_list,
                     (const int*)pl.hub_beg, (const int*)pl.hub_deg, (const int*)pl.hub_src, (const float*)pl.hub_nrm,
                     mlong, mlong_bound);
  return check_launch();
}

// ------------------------------------------------------------------------------------------
// plan construction (structure only) and the hub-row pre-aggregation
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int x, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  return x;
}

// one warp per tile: row records, and the tile's hub rows as one contiguous run of slots
__global__ void plan_tiles_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                  const float* __restrict__ norm, const int32_t* __restrict__ in_row_map,
                                  const int32_t* __restrict__ dst_rows, const int32_t* __restrict__ tile_row0,
                                  const int32_t* __restrict__ tile_nrows, int n_tiles, Plan pl) {
  const int lane = threadIdx.x & 31;
  for (int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < n_tiles; tile += (gridDim.x * blockDim.x) >> 5) {
    const int row0 = tile_row0[tile], nrows = tile_nrows[tile];
    int v[4], beg[4], deg[4], hrank[4], eoff[4];
    bool hub[4];
    int n_h = 0, n_e = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 32 * j + lane;
      v[j] = 0; beg[j] = 0; deg[j] = 0;
      if (r < nrows) {
        v[j] = dst_rows ? dst_rows[row0 + r] : row0 + r;
        beg[j] = indptr[v[j]];
        deg[j] = indptr[v[j] + 1] - beg[j];
      }
      hub[j] = deg[j] > PRE;
      const unsigned bal = __ballot_sync(0xffffffffu, hub[j]);
      hrank[j] = n_h + __popc(bal & ((1u << lane) - 1u));
      n_h += __popc(bal);
      const int pdeg = hub[j] ? (deg[j] + 3) & ~3 : 0;     // records of this hub, padded to whole rounds of 4
      const int incl = warp_incl_scan(pdeg, lane);
      eoff[j] = n_e + incl - pdeg;
      n_e += __shfl_sync(0xffffffffu, incl, 31);
    }
    int hbase = 0, ebase = 0;
    if (lane == 0 && n_h > 0) {
      hbase = atomicAdd(pl.hdr + 0, n_h);        // the order of tiles among the slots does not affect any value
      ebase = atomicAdd(pl.hdr + 1, n_e);
    }
    hbase = __shfl_sync(0xffffffffu, hbase, 0);
    ebase = __shfl_sync(0xffffffffu, ebase, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 32 * j + lane;
      if (r >= nrows) continue;
      PlanRec rc = {0, 0, 0.f, 0.f};
      if (hub[j]) {
        const int slot = hbase + hrank[j];
        pl.hub_row[slot] = v[j];
        pl.hub_beg[slot] = ebase + eoff[j];
        pl.hub_deg[slot] = deg[j];
        if (((deg[j] + 3) & ~3) > HUB_BIG) pl.big_list[atomicAdd(pl.hdr + 3, 1)] = slot;   // order does not affect any value
        rc.r0 = slot; rc.r1 = -1; rc.n0 = 1.f;
      } else {
        if (deg[j] > 0) {
          const int u = indices[beg[j]];
          const int s = in_row_map ? in_row_map[u] : u;
          if (s >= 0) { rc.r0 = s; rc.n0 = norm[u]; }          // negative map entry: neighbour dropped
        }
        if (deg[j] > 1) {
          const int u = indices[beg[j] + 1];
          const int s = in_row_map ? in_row_map[u] : u;
          if (s >= 0) { rc.r1 = s; rc.n1 = norm[u]; }
        }
      }
      pl.rec[row0 + r] = rc;
    }
  }
}

// one warp per hub slot: its edge records, coalesced, in CSR order, padded with zero-weight records
__global__ void plan_hub_edges_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                      const float* __restrict__ norm, const int32_t* __restrict__ in_row_map,
                                      Plan pl) {
  const int n_hub = pl.hdr[0];
  const int lane = threadIdx.x & 31;
  for (int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slot < n_hub; slot += (gridDim.x * blockDim.x) >> 5) {
    const int beg = indptr[pl.hub_row[slot]], deg = pl.hub_deg[slot], dst = pl.hub_beg[slot];
    const int pdeg = (deg + 3) & ~3;
    for (int e = lane; e < pdeg; e += 32) {
      int s = -1;
      float nr = 0.f;
      if (e < deg) {
        const int u = indices[beg + e];
        s = in_row_map ? in_row_map[u] : u;
        nr = norm[u];
      }
      pl.hub_src[dst + e] = s < 0 ? 0 : s;
      pl.hub_nrm[dst + e] = s < 0 ? 0.f : nr;
    }
  }
}

// ------------------------------------------------------------------------------------------
// compute tiles: runs of consecutive output rows with at most TM distinct records
// ------------------------------------------------------------------------------------------
// One warp walks GROUP_TILES consecutive caller tiles (<= 4096 rows) in batches of 32 rows and assigns every row
// the slot of its record (r0, r1, n0, n1) inside the current compute tile: the distinct records of a batch are
// taken in lane order, each looked up among the tile's slots by the whole warp (<= 4 comparisons per lane) and
// appended when new -- deterministic, no atomics.  A compute tile is closed before a batch that could overflow
// TM slots or RMAX rows, at a task change, and at the end of the group.  Rows sharing a record produce the same
// aggregated row, so only the slots are gathered and contracted; the epilogue expands them (see the header).
__global__ void __launch_bounds__(128) plan_dedupe_kernel(const int32_t* __restrict__ tile_row0,
                                                          const int32_t* __restrict__ tile_nrows,
                                                          const int32_t* __restrict__ tile_task, int n_tiles, Plan pl) {
  __shared__ int4 skeys[4][TM];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = blockIdx.x * 4 + w;
  if (g * GROUP_TILES >= n_tiles) return;
  int4* keys = skeys[w];
  CTile* out = pl.ctiles + (size_t)g * GROUP_CAP;
  int nt = 0;                 // compute tiles emitted by this group
  bool open = false;
  int c_row0 = 0, c_rows = 0, c_cnt = 0, c_task = 0;
  auto close = [&]() {
    if (open && c_rows > 0) {
      if (lane == 0) out[nt] = CTile{c_row0, c_rows, c_cnt, c_task};
      ++nt;
    }
    open = false;
  };
  const int t_end = min(n_tiles, (g + 1) * GROUP_TILES);
  for (int tile = g * GROUP_TILES; tile < t_end; ++tile) {
    const int row0 = tile_row0[tile], nrows = tile_nrows[tile], task = tile_task[tile];
    if (open && (task != c_task || row0 != c_row0 + c_rows)) close();
    for (int b = 0; b < nrows; b += 32) {
      if (open && (c_cnt + 32 > TM || c_rows + 32 > RMAX)) close();
      if (!open) { open = true; c_row0 = row0 + b; c_rows = 0; c_cnt = 0; c_task = task; }
      const bool valid = b + lane < nrows;
      int4 key = make_int4(0, 0, 0, 0);
      if (valid) key = *reinterpret_cast<const int4*>(pl.rec + row0 + b + lane);
      int my_slot = 0;
      unsigned todo = __ballot_sync(0xffffffffu, valid);
      while (todo) {
        const int l = __ffs(todo) - 1;
        int4 kl;
        kl.x = __shfl_sync(0xffffffffu, key.x, l);
        kl.y = __shfl_sync(0xffffffffu, key.y, l);
        kl.z = __shfl_sync(0xffffffffu, key.z, l);
        kl.w = __shfl_sync(0xffffffffu, key.w, l);
        const bool same = valid && key.x == kl.x && key.y == kl.y && key.z == kl.z && key.w == kl.w;
        const unsigned grp = __ballot_sync(0xffffffffu, same);
        int found = -1;
        for (int s = lane; s < c_cnt; s += 32) {
          const int4 k = keys[s];
          if (k.x == kl.x && k.y == kl.y && k.z == kl.z && k.w == kl.w) found = s;
        }
        const unsigned fb = __ballot_sync(0xffffffffu, found >= 0);
        int slot;
        if (fb) {
          slot = __shfl_sync(0xffffffffu, found, __ffs(fb) - 1);
        } else {
          slot = c_cnt;
          if (lane == 0) {
            keys[c_cnt] = kl;
            *reinterpret_cast<int4*>(pl.srec + c_row0 + c_cnt) = kl;
          }
          ++c_cnt;
          __syncwarp();
        }
        if (same) my_slot = slot;
        todo &= ~grp;
      }
      if (valid) pl.row_slot[row0 + b + lane] = (uint8_t)my_slot;
      c_rows += min(32, nrows - b);
    }
  }
  close();
  if (lane == 0) pl.group_nt[g] = nt;
}

// in-place exclusive prefix sum of a[0..n) by one CTA; returns the total (in every thread)
__device__ int block_excl_scan(int* __restrict__ a, int n, int* sh /* [blockDim.x + 1] */) {
  int carry = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int x = i < n ? a[i] : 0;
    sh[threadIdx.x] = x;
    __syncthreads();
    for (int d = 1; d < blockDim.x; d <<= 1) {
      const int y = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += y;
      __syncthreads();
    }
    if (i < n) a[i] = carry + sh[threadIdx.x] - x;
    carry += sh[blockDim.x - 1];
    __syncthreads();
  }
  return carry;
}

// compute tiles -> pairs of tiles of the same task (the tiles of a task are contiguous: groups follow the row
// order), the schedule cost in front of every pair and the first pair of every cluster
__global__ void __launch_bounds__(1024) pair_table_kernel(int n_groups, int n_tasks, Plan pl) {
  __shared__ int first[PT_MAXT], cnt[PT_MAXT], base[PT_MAXT], tmp[PT_MAXT];
  __shared__ int sh[1025];
  for (int t = threadIdx.x; t < n_tasks; t += blockDim.x) { first[t] = 0x7fffffff; cnt[t] = 0; }
  __syncthreads();
  // dense index of every compute tile = (tiles of the groups before it) + position inside its group
  const int n_ct = block_excl_scan(pl.group_nt, n_groups, sh);
  for (int g = threadIdx.x; g < n_groups; g += blockDim.x) {
    const int d0 = pl.group_nt[g], d1 = g + 1 < n_groups ? pl.group_nt[g + 1] : n_ct;
    for (int j = 0; j < d1 - d0; ++j) {
      const int t = pl.ctiles[(size_t)g * GROUP_CAP + j].task;
      atomicMin(&first[t], d0 + j);
      atomicAdd(&cnt[t], 1);
    }
  }
  __syncthreads();
  // inclusive scan of ceil(cnt/2) over tasks (Hillis-Steele, double buffered)
  for (int t = threadIdx.x; t < n_tasks; t += blockDim.x) base[t] = (cnt[t] + 1) >> 1;
  __syncthreads();
  int* a = base;
  int* b = tmp;
  for (int d = 1; d < n_tasks; d <<= 1) {
    for (int t = threadIdx.x; t < n_tasks; t += blockDim.x) b[t] = a[t] + (t >= d ? a[t - d] : 0);
    __syncthreads();
    int* c = a; a = b; b = c;
  }
  const int total = n_tasks > 0 ? a[n_tasks - 1] : 0;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    PairEnt e;
    e.row0[0] = e.row0[1] = e.nslots[0] = e.nslots[1] = e.nrows[0] = e.nrows[1] = e.task = e.cost0 = 0;
    for (int k = 0; k < 8; ++k) e.pad[k] = 0;
    pl.pairs[i] = e;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < n_groups; g += blockDim.x) {
    const int d0 = pl.group_nt[g], d1 = g + 1 < n_groups ? pl.group_nt[g + 1] : n_ct;
    for (int j = 0; j < d1 - d0; ++j) {
      const CTile ct = pl.ctiles[(size_t)g * GROUP_CAP + j];
      const int t = ct.task;
      const int jj = d0 + j - first[t];
      PairEnt* e = pl.pairs + a[t] - ((cnt[t] + 1) >> 1) + (jj >> 1);
      e->row0[jj & 1] = ct.row0;
      e->nslots[jj & 1] = ct.nslots;
      e->nrows[jj & 1] = ct.nrows;
      if (!(jj & 1)) e->task = t;
    }
  }
  __syncthreads();
  // schedule cost in front of every pair (output rows of its larger tile + a per-pair constant), then the cluster boundaries:
  // pair i goes to cluster floor(cost0[i] * n_cl / total_cost) -- contiguous runs, tasks stay together
  int carry = 0;
  for (int b0 = 0; b0 < total; b0 += blockDim.x) {
    const int i = b0 + threadIdx.x;
    // the two CTAs of a cluster expand their tiles side by side: a pair costs its larger tile
    const int x = i < total ? max(pl.pairs[i].nrows[0], pl.pairs[i].nrows[1]) + PAIR_FIXED_COST : 0;
    sh[threadIdx.x] = x;
    __syncthreads();
    for (int d = 1; d < blockDim.x; d <<= 1) {
      const int y = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += y;
      __syncthreads();
    }
    if (i < total) pl.pairs[i].cost0 = carry + sh[threadIdx.x] - x;
    carry += sh[blockDim.x - 1];
    __syncthreads();
  }
  const long long total_cost = carry;
  constexpr int n_cl = kNumSMs / 2;
  for (int c = threadIdx.x; c <= n_cl; c += blockDim.x) {
    // first pair whose cost0 >= c * total_cost / n_cl
    const long long want = (total_cost * c + n_cl - 1) / n_cl;
    int lo = 0, hi = total;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (pl.pairs[mid].cost0 < want) lo = mid + 1; else hi = mid;
    }
    pl.cl_beg[c] = c == n_cl ? total : lo;
  }
  if (threadIdx.x == 0) { pl.hdr[2] = total; pl.hdr[4] = n_ct; }
}

int g_pair_dbg = 0;
long long* g_pair_prof = nullptr;

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
  if (g.f_in != 64 && g.f_in != 128 && g.f_in != 256) return false;   // hub pre-pass: K/32 columns per lane
  if (f_out % 16 != 0 || f_out < 16 || f_out > 256) return false;
  if (g.ld_in % 8 != 0 || (reinterpret_cast<uintptr_t>(g.in) & 31u) || g.ld_in < g.f_in) return false;   // 256-bit loads
  if (ld_out % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 31u)) return false;
  if (relu_mask && (reinterpret_cast<uintptr_t>(relu_mask) & 31u)) return false;
  if (n_tasks > PT_MAXT) return false;
  return stages_for(g.f_in, f_out) >= 3;
}

int64_t layer_plan_bytes(int n_tiles, int n_tasks, int n_rows, int n_edges) {
  return carve_plan(nullptr, n_tiles, n_tasks, n_rows, n_edges).total;
}

int layer_plan_build(const int32_t* indptr, const int32_t* indices, const float* norm, const int32_t* in_row_map,
                     const int32_t* dst_rows, const int32_t* tile_row0, const int32_t* tile_nrows,
                     const int32_t* tile_task, int n_tiles, int n_tasks, int n_rows, int n_edges, void* plan,
                     cudaStream_t stream) {
  if (n_tasks > PT_MAXT) return GMETA_ERR_UNSUPPORTED;
  if (!plan || (reinterpret_cast<uintptr_t>(plan) & 255u)) return GMETA_ERR_ALIGN;
  Plan pl = carve_plan(plan, n_tiles, n_tasks, n_rows, n_edges);
  if (cudaMemsetAsync(pl.hdr, 0, 256, stream) != cudaSuccess) return GMETA_ERR_LAUNCH;
  if (n_tiles == 0) return GMETA_OK;
  int rc;
  const int grid = ceil_div(n_tiles, 8) < 8 * kNumSMs ? ceil_div(n_tiles, 8) : 8 * kNumSMs;
  plan_tiles_kernel<<<grid, 256, 0, stream>>>(indptr, indices, norm, in_row_map, dst_rows, tile_row0, tile_nrows,
                                             n_tiles, pl);
  if ((rc = check_launch()) != GMETA_OK) return rc;
  plan_hub_edges_kernel<<<4 * kNumSMs, 256, 0, stream>>>(indptr, indices, norm, in_row_map, pl);
  if ((rc = check_launch()) != GMETA_OK) return rc;
  const int n_groups = n_groups_for(n_tiles);
  plan_dedupe_kernel<<<ceil_div(n_groups, 4), 128, 0, stream>>>(tile_row0, tile_nrows, tile_task, n_tiles, pl);
  if ((rc = check_launch()) != GMETA_OK) return rc;
  pair_table_kernel<<<1, 1024, 0, stream>>>(n_groups, n_tasks, pl);
  return check_launch();
}

int64_t gcn_layer_fwd_pair_workspace_bytes(int n_copies, int n_tiles, int n_tasks, int n_rows, int n_edges, int f_in,
                                           int f_out) {
  return carve_ws(nullptr, n_copies, n_tiles, n_tasks, n_rows, n_edges, f_in, f_out).total;
}

int row_absmax(const float* x, int ld, int n_rows, int f, float* out, cudaStream_t stream) {
  if (n_rows == 0) return GMETA_OK;
  const int grid = ceil_div(n_rows, 8) < 16 * kNumSMs ? ceil_div(n_rows, 8) : 16 * kNumSMs;
  launch_pdl(row_absmax_kernel, dim3(grid), dim3(256), 0, stream, x, ld, n_rows, f, out);
  return check_launch();
}

int gcn_layer_fwd_pair(const GatherSrc& g, const int32_t* tile_row0, const int32_t* tile_nrows,
                       const int32_t* tile_task, int n_tiles, int n_tasks, int n_copies, int n_rows, int n_edges,
                       const float* in_rowmax, const void* plan, const float* W, int64_t w_task_stride, int ldw,
                       int trans_w, const float* bias, int64_t b_task_stride, int f_out, int relu,
                       const float* relu_mask, float* out, int ld_out, float* out_rowmax, void* workspace,
                       int64_t workspace_bytes, cudaStream_t stream) {
  const int K = g.f_in, N = f_out;
  if (!in_rowmax || n_rows <= 0 || n_edges < 0) return GMETA_ERR_BAD_ARG;
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return GMETA_ERR_WORKSPACE;
  if (plan && (reinterpret_cast<uintptr_t>(plan) & 255u)) return GMETA_ERR_ALIGN;
  Workspace ws = carve_ws(workspace, n_copies, n_tiles, n_tasks, n_rows, n_edges, K, N);
  if (workspace_bytes < ws.total) return GMETA_ERR_WORKSPACE;
  int rc;
  if (!plan) {
    rc = layer_plan_build(g.indptr, g.indices, g.norm, g.in_row_map, g.dst_rows, tile_row0, tile_nrows, tile_task,
                          n_tiles, n_tasks, n_rows, n_edges, ws.plan, stream);
    if (rc != GMETA_OK) return rc;
    plan = ws.plan;
  }
  const Plan pl = carve_plan(const_cast<void*>(plan), n_tiles, n_tasks, n_rows, n_edges);
  launch_pdl(pack_w_pair_kernel, dim3(n_copies), dim3(1024), 0, stream, W, w_task_stride, ldw, trans_w, K, N, ws.w_image,
             2LL * K * N, ws.w_inv_scale);
  if ((rc = check_launch()) != GMETA_OK) return rc;
  if (!(g_pair_dbg & 16)) {     // hub rows first (debug flag 16: skip, results are wrong)
    if (K == 256) rc = hub_prepass_launch<8>(g.in, g.ld_in, in_rowmax, pl, ws.mlong, ws.mlong_bound, stream);
    else if (K == 128) rc = hub_prepass_launch<4>(g.in, g.ld_in, in_rowmax, pl, ws.mlong, ws.mlong_bound, stream);
    else rc = hub_prepass_launch<2>(g.in, g.ld_in, in_rowmax, pl, ws.mlong, ws.mlong_bound, stream);
    if (rc != GMETA_OK) return rc;
  }
  PairParams p;
  p.in = g.in; p.ld_in = g.ld_in; p.f_in = K; p.in_rowmax = in_rowmax;
  p.mlong = ws.mlong; p.mlong_bound = ws.mlong_bound;
  p.srec = pl.srec; p.row_slot = pl.row_slot; p.hdr = pl.hdr; p.pairs = pl.pairs; p.cl_beg = pl.cl_beg;
  p.dst_rows = g.dst_rows; p.norm = g.norm_dst ? g.norm_dst : g.norm;     // the epilogue scales destinations
  p.w_image = ws.w_image; p.image_task_stride = n_copies > 1 ? 2LL * K * N : 0; p.w_inv_scale = ws.w_inv_scale;
  p.bias = bias; p.b_task_stride = b_task_stride; p.f_out = N; p.relu = relu; p.relu_mask = relu_mask;
  p.out = out; p.ld_out = ld_out; p.out_rowmax = out_rowmax;
  p.n_stages = stages_for(K, N);
  p.w_bytes = 2 * K * N;   // bytes per CTA: K x N/2 halves, hi + lo
  p.dbg = g_pair_dbg;
  p.prof = g_pair_prof;
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
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * n_clusters); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;     // see pdl_wait in the kernel
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, gcn_layer_fwd_pair_kernel, p);
  return check_launch();
}

}  // namespace gmeta

// Debug ablation switches of the CTA-pair kernel for performance triage (results are WRONG with any flag set).
extern "C" void gmeta_debug_set_pair_flags(int flags) { gmeta::g_pair_dbg = flags; }
extern "C" void gmeta_debug_set_pair_profile(long long* device_buffer) { gmeta::g_pair_prof = device_buffer; }
