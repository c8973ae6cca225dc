// Inline-PTX wrappers (sm_100a) shared by the CTA-pair tensor-core layer kernel: mbarriers with
// cluster-scope variants, bulk async copies, tcgen05 (cta_group::2) MMA / commit / TMEM access.
#pragma once
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gmeta {
namespace ptx {

constexpr long long kWaitTimeoutCycles = 4000000000LL;  // ~2 s: trap instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init_cluster() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive (release at cluster scope) on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta) : "memory");
}
// same without release semantics: the caller has already ordered its shared-memory writes with
// fence.proxy.async, and a release here would also wait for the thread's outstanding global loads
// (the gather producers always have the next chunk's loads in flight)
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes or the
// hint expires, instead of burning issue slots and shared-memory bandwidth in a software spin loop
constexpr uint32_t kSuspendHintNs = 20000u;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(kSuspendHintNs) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(kSuspendHintNs) : "memory");
  return ok != 0;
}
// A printf here would be an ABI call, which makes ptxas cap EVERY warp role at the launch-bound
// register count (setmaxnreg.inc regions included); debug builds only.
#ifndef GMETA_TC_DEBUG_PRINT
#define GMETA_TC_DEBUG_PRINT 0
#endif
__device__ __forceinline__ void wait_timed_out(uint32_t bar, uint32_t parity, int what) {
#if GMETA_TC_DEBUG_PRINT
  printf("gmeta pair kernel: mbarrier wait timed out (block %d thread %d bar %u parity %u site %d)\n",
         (int)blockIdx.x, (int)threadIdx.x, bar, parity, what);
#else
  (void)bar; (void)parity; (void)what;
#endif
  __trap();
}
// `what` tags the call site in the time-out message
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int what = 0) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  unsigned n = 0;
  while (!mbar_try_wait(bar, parity))
    if ((++n & 255u) == 0 && clock64() - t0 > kWaitTimeoutCycles) wait_timed_out(bar, parity, what);
}
// whole-warp wait: one lane polls (with back-off), the warp re-converges on __syncwarp.  Hundreds of
// threads spinning on try_wait would otherwise flood the shared-memory pipe of the SM.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int what = 0) {
  if ((threadIdx.x & 31) == 0 && !mbar_try_wait(bar, parity)) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
      __nanosleep(32);
      if (clock64() - t0 > kWaitTimeoutCycles) wait_timed_out(bar, parity, what);
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int what = 0) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  unsigned n = 0;
  while (!mbar_try_wait_cluster(bar, parity))
    if ((++n & 255u) == 0 && clock64() - t0 > kWaitTimeoutCycles) wait_timed_out(bar, parity, what);
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

// arrive on the barrier at this offset in BOTH CTAs of the pair once all MMAs issued so far are done
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem desc, own rows of each CTA] * B[smem desc, N/2 rows from each CTA]; kind::f16
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, SWIZZLE_128B shared-memory operand descriptor (cute::UMMA::SmemDescriptor): start
// address >> 4 [0,14), LBO (ignored for swizzled K-major, 1) [16,30), SBO = 1024 B between 8-row
// groups [32,46), version 1 [46,48), layout SWIZZLE_128B = 2 at [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// same for SWIZZLE_64B operand tiles (64-byte rows, 512 B between 8-row groups, layout type 4)
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 at [4,6), a/b format
// F16 = 0 at [7,10)/[10,13), a/b K-major (0) at 15/16, N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 2^e as a float, e in [-126, 127]
__device__ __forceinline__ float exp2i(int e) { return __int_as_float((e + 127) << 23); }

}  // namespace ptx
}  // namespace gmeta
