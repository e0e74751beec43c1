// cluster_halo.cuh -- point-to-point halo exchange between neighbouring CTAs of a thread-block
// cluster, used by the forward sweep kernels when the rows of a source are split over several SMs.
//
// A cluster barrier per level (barrier.cluster arrive.release / wait.acquire) is a full memory fence:
// it drains the CTA's outstanding global stores and invalidates L1 -- measured 1.8x slower sweeps.
// Here the boundary row travels with st.async (DSMEM store that completes a transaction count on an
// mbarrier in the DESTINATION CTA), so data + "ready" signal are one operation and no fence is needed;
// a second pair of mbarriers carries the back-pressure ("halo buffer free") from consumer to producer.
// Per level the CTA still runs its own __syncthreads only.
//
// Protocol for one sweep (step s writes sheet buffer p = s & 1; `last` = number of levels - 1):
//   producer, step s <= last-1 : [s >= 2: wait empty[p]]; thread 0: remote arrive.expect_tx(consumer full[p], bytes of
//                                my edge row at this level, possibly 0); edge nodes: st.async -> consumer sheet[p] halo
//   consumer, step s >= 1      : wait full[1-p]  (signal + halo of step s-1)
//   consumer, after step s in [1, last-2] : remote arrive on producer's empty[1-p]
// The producer signals EVERY step, also when its edge row is not in the level: this keeps the two CTAs within
// one (consumer) / two (producer) steps of each other, so no mbarrier ever sees two arrivals in one phase.
#pragma once
#include <cstdint>

namespace adtomo {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned cluster_map(unsigned local_addr, int cta_rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arm_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_LOOP;\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_remote_arrive(unsigned remote_bar) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
// remote: announce `bytes` of st.async traffic for the current phase and arrive (count 1)
__device__ __forceinline__ void mbar_remote_arrive_tx(unsigned remote_bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(remote_bar), "r"(bytes)
                 : "memory");
}
// store one double into another CTA's shared memory and count 8 bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_f64(unsigned remote_addr, double v, unsigned remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(remote_addr),
                 "d"(v), "r"(remote_bar)
                 : "memory");
}

// State of the exchange for one sweep, identical in all threads of the CTA.
struct HaloLink {
    uint64_t *bars;          // shared: full[0], full[1], empty[0], empty[1]
    bool hasUp, hasDown;     // a neighbour pushes to me / I push to a neighbour
    unsigned rmFull[2];      // downstream neighbour's full[] (I complete transactions on them)
    unsigned rmEmpty[2];     // upstream neighbour's empty[] (I arrive on them)
    unsigned rmSheets;       // downstream neighbour's sheet base (shared::cluster address)
    unsigned phFull[2], phEmpty[2];
};

}  // namespace adtomo
