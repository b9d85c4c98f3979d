// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// descriptor builders and small numeric utilities.  Everything here is inline PTX written for
// B200 (compute_100a); nothing is portable to older parts and nothing falls back.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/autognothi_b200.h"

namespace agb {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// status codes returned by every C-ABI entry point (include/autognothi_b200.h mirrors these)
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define AGB_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      agb::set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return AGB_ERR_CUDA;                                                           \
    }                                                                                     \
  } while (0)

#define AGB_REQUIRE(cond, msg)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      agb::set_last_error("%s:%d requirement failed: %s (%s)", __FILE__, __LINE__, #cond, msg); \
      return AGB_ERR_INVALID;                                                        \
    }                                                                                     \
  } while (0)

// ---------------------------------------------------------------------------------------------
// generic
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (an error the host sees), never as
// a hung GPU.  ~4e9 SM cycles is > 2 s at any clock the part runs at.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ffu) == 0 && (clock64() - t0) > 4000000000LL) {
      printf("agb: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// Busy-polling variant (mbarrier.test_wait never suspends the thread): for the few latency-critical single hand-offs of a
// pipeline (a tcgen05 issuer waiting for its operand), where the wake-up latency of try_wait's suspension is on the
// critical path.  Same bounded wait.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++spins & 0xfffu) == 0 && (clock64() - t0) > 4000000000LL) {
      printf("agb: mbarrier spin-wait timed out (block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x, (int)threadIdx.x,
             bar, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — 2D / 3D tiled loads completing on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap,
                                            uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tmap,
                                            uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_result_addr),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued MMAs of this thread arrive (once) on `bar` when they retire.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 32 lanes x 32 bit, 16 consecutive columns -> 16 registers per thread (thread = TMEM lane).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA stores (smem -> global, bulk async-group completion) and the CTA-pair (cta_group::2) forms
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's most recent bulk groups may still be READING their smem source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// full completion (global writes performed)
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// address of the same smem offset in CTA `rank` of this cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float ld_shared_cluster_f32(uint32_t cluster_addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// The pair's TMA loads land in the issuing CTA's smem but complete on the LEADER's mbarrier: clearing
// the peer bit of a shared::cta address names the same offset in the even CTA of the pair.
constexpr uint32_t PAIR_LEADER_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// ---------------------------------------------------------------------------------------------
// "Elected" forms for the single-thread roles.  The whole warp executes these convergently with
// warp-uniform operands and only the instruction itself is predicated on the elected lane (e != 0 in
// exactly one lane, from elect_one()).  Keeping the address / descriptor arithmetic outside a
// divergent `if (lane == 0)` lets ptxas hold it in uniform registers instead of emitting an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop around every TMA and MMA instruction (round-1 finding:
// the producer and MMA-issue loops cost ~750 cycles per 512-cycle k-block).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_idx_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ void mbar_arrive_expect_tx_e(uint32_t e, uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
               "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes), "r"(e)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_e(uint32_t e, uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar,
                                              int c0, int c1) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
               "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%3, %4}], [%2];\n\t}\n"
               :
               : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(e)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_e(uint32_t e, uint32_t smem_dst, const CUtensorMap* tmap,
                                                   uint32_t bar, int c0, int c1) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
               "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%3, %4}], [%2];\n\t}\n"
               :
               : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & PAIR_LEADER_MASK), "r"(c0),
                 "r"(c1), "r"(e)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d_e(uint32_t e, uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar,
                                              int c0, int c1, int c2) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %6, 0;\n\t"
               "@q cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%3, %4, %5}], [%2];\n\t}\n"
               :
               : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(e)
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem], elected form
__device__ __forceinline__ void umma_ts_e(uint32_t e, uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               :
               : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(e)
               : "memory");
}
// 32 lanes x 32 bit x 32 columns register -> TMEM store (thread = TMEM lane)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// wait for the outstanding TMEM loads AND tie the destination registers of an asynchronous tmem_ld32 to the wait: between
// the load and this call the compiler must neither read them (the values are not there yet) nor reuse them
__device__ __forceinline__ void tmem_wait_ld_dep32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_ss_e(uint32_t e, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  if (CG == 2) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :
                 : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(e)
                 : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :
                 : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(e)
                 : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void umma_commit_e(uint32_t e, uint32_t bar) {
  if (CG == 2) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n"
                 :
                 : "r"(bar), "r"(e)
                 : "memory");
  } else {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
                 :
                 : "r"(bar), "r"(e)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// descriptors
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B canonical layouts (Blackwell "version 1" format):
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1
//   bits [61,64) layout type (2 = SWIZZLE_128B)
// K-major tile  (rows x 64 bf16, 128 B per row, 8-row groups of 1024 B): LBO = 1 (ignored),
//   SBO = 1024 B.  MN-major tile (64 mn-elements contiguous per 128 B row, one row per k):
//   LBO = byte stride between 64-element mn atoms, SBO = 1024 B (8 k-rows).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16 with bf16 inputs, fp32 accumulation.
//   [4,6) c_format=1(F32)  [7,10) a_format=1(BF16)  [10,13) b_format=1(BF16)
//   [15] a_major (0=K,1=MN)  [16] b_major  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int a_mn_major,
                                                       int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// numerics
// ---------------------------------------------------------------------------------------------
// erf-exact GELU (nn.GELU() default, reference models/vanilla_vit.py:488): fp32 path uses erff.
__device__ __forceinline__ float gelu_erf_exact(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// Cheap erf for the bf16 epilogue (Abramowitz-Stegun 7.1.26, |err| <= 1.5e-7 — far below bf16
// resolution): ~14 issue slots instead of erff's branchy ~30.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-z * z * 1.4426950408889634f);
  const float erf_abs = fmaf(-p, e, 1.0f);
  const float erf_v = copysignf(erf_abs, x);
  const float hx = 0.5f * x;
  return fmaf(hx, erf_v, hx);
}

// erf-GELU with ONE MUFU and 5 FMA-pipe ops:  gelu(x) = 0.5 x (1 + tanh(x (a + b x^2))), the tanh-form REFITTED to the
// erf definition (minimax over [-10, 10]; not the classic 0.044715 "tanh GELU", whose error is 4.7e-4): max |error|
// 2.7e-4, max relative error 5.8e-4 for x > 0 — a quarter of a bf16 ulp, the precision of the tensor the epilogue
// writes.  Both coefficients are positive, so the argument is monotone and needs no clamp.  (A 3-coefficient fit
// reaches 2.5e-5 but costs a clamp + one more FMA per element in an epilogue that is issue-bound; the fp32-exact
// mode uses erff.)
// ---- dropout (training mode of nn.Dropout; reference models/vanilla_vit.py:253,457,501-503, vanilla_bert.py:325,530,559,603)
// keep / drop decided by a counter hash, so the adjoint regenerates the mask instead of storing it.  One 32-bit hash
// serves two neighbouring elements (16 bits each): element kept iff its 16 bits >= thr16 = round(p * 65536).
__host__ __device__ __forceinline__ uint32_t agb_hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t agb_drop_key(unsigned long long seed, uint32_t a, uint32_t b) {
  return agb_hash32(agb_hash32((uint32_t)seed ^ (a * 0x85EBCA6Bu)) ^ (uint32_t)(seed >> 32) ^ (b * 0xC2B2AE35u));
}
__host__ __device__ __forceinline__ uint32_t agb_drop_bits(uint32_t key, uint32_t pair) {
  return agb_hash32(pair * 0x9E3779B1u + key);
}
// attention-probability dropout: stream of (row * heads + head, query index); pair = key index / 2
__host__ __device__ __forceinline__ bool agb_attn_keep(unsigned long long seed, uint32_t unit, uint32_t query, uint32_t key_idx,
                                                       uint32_t thr16) {
  const uint32_t x = agb_drop_bits(agb_drop_key(seed, unit, query), key_idx >> 1);
  return ((key_idx & 1u) ? (x >> 16) : (x & 0xFFFFu)) >= thr16;
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf_tanhform(float x) {
  const float q = fmaf(0.03470090309328562f, x * x, 0.8001570568972525f);
  const float th = tanh_approx(x * q);
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}

// Derivative of the SAME approximant (bf16 training adjoints: gelu_bwd on bf16 data, the dgrad GEMM's fused epilogue), so
// forward and adjoint stay consistent:  g'(x) = 0.5 (1 + t) + 0.5 x (1 - t^2) (c0 + 3 c1 x^2);  |g' - exact| <= 9e-4 (bf16
// resolves 4e-3 near 1).  The erf / exp forms made the elementwise kernels compute-bound (85 / 105 us per 12608 x 3072 launch
// against 26 / 39 us of HBM time).
__device__ __forceinline__ float gelu_grad_tanhform(float x) {
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(0.03470090309328562f, x2, 0.8001570568972525f));
  const float hx = 0.5f * x;
  const float du = fmaf(3.0f * 0.03470090309328562f, x2, 0.8001570568972525f);
  return fmaf(hx * fmaf(-t, t, 1.0f), du, fmaf(0.5f, t, 0.5f));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// host side: cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency)
int encode_tmap_2d_bf16(CUtensorMap* out, const void* gptr, uint64_t inner, uint64_t outer,
                        uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer);
int encode_tmap_2d(CUtensorMap* out, const void* gptr, int elem_bytes, uint64_t inner, uint64_t outer,
                   uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer);
int encode_tmap_3d_bf16(CUtensorMap* out, const void* gptr, uint64_t d0, uint64_t d1, uint64_t d2,
                        uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0,
                        uint32_t box1, uint32_t box2);
int sm_count();

// ---------------------------------------------------------------------------------------------
// One request to the CTA-pair tcgen05 GEMM (agb_gemm_pair.cu):  C = epilogue(alpha * A * B^T + bias).
// Fields left at their defaults are off.  gemm_bf16_pair_call returns AGB_ERR_UNSUPPORTED for combinations / shapes the
// kernel does not cover (narrow N, bf16 residual, ...): the generic entry point then runs the first-generation kernel.
// ---------------------------------------------------------------------------------------------
struct PairGemmCall {
  const bf16* A = nullptr;  int lda = 0, a_mn = 0;       // operand majorness: 0 = K-major [rows, K], 1 = MN-major [K, rows]
  const bf16* B = nullptr;  int ldb = 0, b_mn = 0;
  int M = 0, N = 0, K = 0;
  float alpha = 1.0f;
  const float* bias = nullptr;
  int act = 0;                                           // 0 none, 1 GELU, 2 GELU + pre-activation to z_out (training)
  const bf16* res_bf16 = nullptr;                        // (not covered by this kernel)
  const float* res_f32 = nullptr;  int ldr = 0;          // fp32 residual (fp32 output)
  void* out = nullptr;  int ldo = 0, out_f32 = 0;
  // LayerNorm of the A rows folded into the epilogue (bf16 output, no residual): row statistics + column sums of B
  const float* ln_stats = nullptr;  int ln_parts = 0;  const float* ln_colsum = nullptr;  float ln_eps = 0.f;
  // fp32 residual epilogue: also emit a bf16 copy of the output and per-row partial (sum, sum of squares) per 128-column slab
  bf16* out16 = nullptr;  int ldo16 = 0;
  float* stats_out = nullptr;                            // [M][2 * ceil(N / 256)][2]; mandatory with hl_hi
  // residual stream as two bf16 planes updated in place (x = hi + lo); `out`, `res_f32`, `out16` unused
  bf16* hl_hi = nullptr;  bf16* hl_lo = nullptr;  int ld_hl = 0;
  // nn.Dropout on the GEMM output ahead of the fp32 residual add (threshold 0 = off; key of agb_dropout's stream)
  unsigned drop_thr = 0, drop_key = 0;
  bf16* z_out = nullptr;  int ldz_out = 0;               // act == 2: pre-activation
  const bf16* z_in = nullptr;  int ldz_in = 0;           // out = acc * GELU'(z_in)  (bf16 output, no bias)
};
int gemm_bf16_pair_call(const PairGemmCall& call, cudaStream_t stream);

}  // namespace agb
