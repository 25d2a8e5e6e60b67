// ptx_async.cuh — thin inline-PTX wrappers for the sm_100a asynchronous machinery the kernels use:
// transaction barriers (mbarrier), 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP), asynchronous gathers
// (cp.async, SASS LDGSTS) reported to an mbarrier, L2 eviction policies, and asynchronous stores into a peer
// CTA's shared memory that complete a barrier there (st.async over distributed shared memory).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rdisgpu {

// ---- PTX wrappers (mbarrier + TMA bulk copy) ----------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_addr_u32(bar);
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
// producer-side wait: it is one thread with nothing else to do, so it sleeps between probes instead
// of competing with the consumer warps for issue slots
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_addr_u32(bar);
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(100);
  }
}
// L2 policy of the streamed slices: evict-first, so that 268 MB of single-use stream does not push the
// 17 MB of gathered variable values (re-used ~8 times each, loaded evict-last below) out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar,
                                             uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_addr_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr_u32(bar)), "l"(policy)
               : "memory");
}
// shared -> global bulk store (bulk async-group completion).  Source, destination and size are multiples of 16 bytes.
__device__ __forceinline__ void tma_bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_addr_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// every committed bulk store has finished READING its shared-memory source (the buffer may be overwritten)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... has been performed entirely
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory made visible to the asynchronous proxy (before a barrier that precedes a bulk store)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// the variable gather: 16 B {value, direction slot}, kept in L2 with evict-last priority
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// asynchronous 8-byte gather global -> shared (LDGSTS), L2 evict-last; completion is reported to an mbarrier
__device__ __forceinline__ void cp_async_gather8(double* dst_smem, const double* src_gmem, uint64_t policy) {
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(smem_addr_u32(dst_smem)), "l"(src_gmem),
               "l"(policy)
               : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_addr_u32(bar)) : "memory");
}

// ---- distributed shared memory: peer addresses, asynchronous stores, cluster-scope barrier wait ----
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta_rank) {  // shared::cta -> shared::cluster of a peer
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_async_v2(uint32_t remote_addr, double a, double b, uint32_t remote_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(remote_addr),
               "d"(a), "d"(b), "r"(remote_mbar)
               : "memory");
}
// wait that also acquires at cluster scope: the data was written by peers' st.async
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_addr_u32(bar);
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}

}  // namespace rdisgpu
