// benchmarks/microbench_level.cu -- what a level of the team forward kernel is made of (sm_100a, B200).
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb benchmarks/microbench_level.cu && /tmp/mb
// All numbers are clock64() cycles of one warp (SM clock), medians over repetitions.
//  (1) dependent-load latency: L1 hit, L2 hit (ld.cg), DRAM (cold line, > L2 footprint)
//  (2) prefetch.global.L1 followed (after a delay) by a plain load of the same line: L1 hit or not?
//  (3) __syncthreads of a 512-thread CTA: bare, and with one 8-byte global store per thread issued just before
//  (4) packet hop: CTA A on one SM writes a tagged 16-byte packet (st.relaxed.gpu.v2.u64), CTA B on another SM
//      spins on it (ld.relaxed.gpu.v2.u64) and answers; half the ping-pong round trip
//  (5) STS -> __syncthreads -> LDS of another warp's value (the sheet hand-over)
//  (6) the same ping-pong as (4) inside a thread-block cluster of 2: the producer writes the 8-byte token straight
//      into the consumer's shared memory (st.shared::cluster via mapa), the consumer spins on its OWN shared memory
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
typedef unsigned long long u64;

__global__ void k_chase(const u64 *p, int iters, int mode, u64 *out) {   // mode 0: ld.ca, 1: ld.cg
    u64 idx = 0;
    // warm one pass for the L1 / L2 cases is done by the host choosing the footprint
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (mode == 0) idx = p[idx];
        else idx = __ldcg(p + idx);
    }
    long long t1 = clock64();
    out[0] = (u64)(t1 - t0);
    out[1] = idx;
}

__global__ void k_prefetch(const double *p, long long stride, int n, int delay, u64 *out, int do_prefetch) {
    // for n distinct cold lines: prefetch.L1, spin `delay` cycles, then a plain (L1-cached) load.  The address of
    // iteration i+1 depends on the value loaded in iteration i (all zeros), so an iteration cannot start before
    // its predecessor's load has returned: time per iteration = delay + load latency + a few cycles of loop overhead.
    long long extra = 0;
    long long spun = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        const double *q = p + (long long)i * stride + extra;
        if (do_prefetch) asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
        long long s = clock64();
        while (clock64() - s < delay) {}
        spun += clock64() - s;
        double v;
        asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(q) : "memory");
        extra = (long long)v;          // 0, but only known once the load is back
    }
    long long t1 = clock64();
    out[0] = (u64)((t1 - t0 - spun) / n);
    out[1] = (u64)extra;
}

__global__ void k_barrier(double *g, int iters, int with_store, u64 *out) {
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (with_store) g[(size_t)i * blockDim.x + threadIdx.x] = (double)i;
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (u64)(t1 - t0) / iters;
}

__global__ void k_sheet(int iters, u64 *out) {
    __shared__ double sh[2][512];
    double v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        sh[i & 1][threadIdx.x] = v;
        __syncthreads();
        v = sh[i & 1][(threadIdx.x + 32) & 511] + 1.0;      // another warp's value
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (u64)(t1 - t0) / iters; out[1] = (u64)v; }
}

__device__ __forceinline__ void st_pk(u64 *s, u64 a, u64 b) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(s), "l"(a), "l"(b) : "memory"); }
__device__ __forceinline__ void ld_pk(const u64 *s, u64 &a, u64 &b) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(s) : "memory"); }

// 2 CTAs (cooperative launch so that both are resident), one warp each does the ping-pong, lane i uses packet i
__global__ void k_hop(u64 *box, int iters, u64 *out, unsigned *smid) {
    const int me = blockIdx.x, lane = threadIdx.x;
    u64 *mine = box + (size_t)me * 64 + 2 * lane, *other = box + (size_t)(1 - me) * 64 + 2 * lane;
    if (lane == 0) { unsigned s; asm("mov.u32 %0, %%smid;" : "=r"(s)); smid[me] = s; }
    long long t0 = clock64();
    for (int i = 1; i <= iters; i++) {
        u64 a, b;
        if (me == 0) {
            st_pk(other, (u64)i, (u64)i);
            do { ld_pk(mine, a, b); } while (a != (u64)i || b != (u64)i);
        } else {
            do { ld_pk(mine, a, b); } while (a != (u64)i || b != (u64)i);
            st_pk(other, (u64)i, (u64)i);
        }
    }
    long long t1 = clock64();
    if (me == 0 && lane == 0) out[0] = (u64)(t1 - t0) / iters / 2;
}

// cluster of 2 CTAs, one warp each; lane 0 does the ping-pong with an 8-byte token in distributed shared memory
__global__ void __cluster_dims__(2, 1, 1) k_hop_dsmem(int iters, u64 *out, unsigned *smid) {
    __shared__ u64 slot;
    unsigned rank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        slot = 0;
        unsigned s; asm("mov.u32 %0, %%smid;" : "=r"(s)); smid[rank] = s;
    }
    asm volatile("barrier.cluster.arrive.release.aligned; barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) {
        const unsigned local = (unsigned)__cvta_generic_to_shared(&slot);
        unsigned remote;
        asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(1u - rank));
        long long t0 = clock64();
        for (int i = 1; i <= iters; i++) {
            u64 v;
            if (rank == 0) {
                asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(remote), "l"((u64)i) : "memory");
                do { asm volatile("ld.relaxed.cluster.shared::cta.u64 %0, [%1];" : "=l"(v) : "r"(local) : "memory"); } while (v != (u64)i);
            } else {
                do { asm volatile("ld.relaxed.cluster.shared::cta.u64 %0, [%1];" : "=l"(v) : "r"(local) : "memory"); } while (v != (u64)i);
                asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(remote), "l"((u64)i) : "memory");
            }
        }
        long long t1 = clock64();
        if (rank == 0) out[0] = (u64)(t1 - t0) / iters / 2;
    }
    asm volatile("barrier.cluster.arrive.release.aligned; barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

static u64 run_chase(size_t bytes, size_t stride, int mode, int iters, bool warm) {
    size_t n = bytes / 8, step = stride / 8;
    std::vector<u64> h(n, 0);
    // cyclic chain with a fixed stride (wraps), one element per line
    for (size_t i = 0; i < n; i += step) h[i] = (i + step) % n - ((i + step) % n) % step;
    u64 *d, *o;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&o, 16));
    CK(cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice));
    if (warm) { k_chase<<<1, 1>>>(d, (int)(n / step), mode, o); CK(cudaDeviceSynchronize()); }
    k_chase<<<1, 1>>>(d, iters, mode, o);
    CK(cudaDeviceSynchronize());
    u64 r[2];
    CK(cudaMemcpy(r, o, 16, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(o);
    return r[0] / iters;
}

static double *g_big = nullptr;
static void flush_l2() {          // write a buffer larger than L2 (126 MB)
    if (!g_big) CK(cudaMalloc(&g_big, (size_t)512 << 20));
    CK(cudaMemset(g_big, 1, (size_t)512 << 20));
    CK(cudaDeviceSynchronize());
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    printf("{\"device\": \"%s\", \"sm_clock_khz\": %d,\n", pr.name, khz);
    // (1) dependent-load latency
    const u64 l1 = run_chase(16 << 10, 128, 0, 4096, true);
    const u64 l2 = run_chase(8 << 20, 128, 1, 20000, true);
    flush_l2();
    const u64 dr = run_chase((size_t)2 << 30, 4096 + 128, 1, 20000, false);
    printf(" \"load_latency_cycles\": {\"l1_hit_ld_ca_16KB\": %llu, \"l2_hit_ld_cg_8MB\": %llu, \"dram_ld_cg_2GB_cold\": %llu},\n", l1, l2, dr);
    u64 *o;
    CK(cudaMalloc(&o, 64));
    u64 r[2];
    {   // (2) prefetch.global.L1 then a plain load of the same (cold) line
        const long long stride = 4096 + 128;       // bytes
        const int n = 2000;
        double *d;
        CK(cudaMalloc(&d, (size_t)n * stride + 4096));
        CK(cudaMemset(d, 0, (size_t)n * stride + 4096));
        printf(" \"plain_load_after_prefetch_L1_cycles\": {");
        const int delays[4] = {0, 500, 2000, 6000};
        for (int k = 0; k < 4; k++) {
            flush_l2();
            k_prefetch<<<1, 1>>>(d, stride / 8, n, delays[k], o, 1);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(r, o, 16, cudaMemcpyDeviceToHost));
            printf("\"prefetch_then_wait_%d\": %llu, ", delays[k], r[0]);
        }
        flush_l2();
        k_prefetch<<<1, 1>>>(d, stride / 8, n, 0, o, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(r, o, 16, cudaMemcpyDeviceToHost));
        printf("\"no_prefetch_cold\": %llu, ", r[0]);
        // warm in L2 (not flushed), no prefetch: the L2-hit cost of the same code
        k_prefetch<<<1, 1>>>(d, stride / 8, n, 0, o, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(r, o, 16, cudaMemcpyDeviceToHost));
        printf("\"no_prefetch_l2_warm\": %llu},\n", r[0]);
        cudaFree(d);
    }
    {   // (3) __syncthreads, 512 threads
        double *g;
        const int iters = 2000;
        CK(cudaMalloc(&g, (size_t)iters * 512 * 8));
        k_barrier<<<1, 512>>>(g, iters, 0, o);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(r, o, 8, cudaMemcpyDeviceToHost));
        printf(" \"syncthreads_512_cycles\": {\"bare\": %llu, ", r[0]);
        k_barrier<<<1, 512>>>(g, iters, 1, o);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(r, o, 8, cudaMemcpyDeviceToHost));
        printf("\"after_one_global_store_per_thread\": %llu},\n", r[0]);
        cudaFree(g);
    }
    {   // (5) sheet hand-over
        k_sheet<<<1, 512>>>(2000, o);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(r, o, 8, cudaMemcpyDeviceToHost));
        printf(" \"sts_syncthreads_lds_512_cycles\": %llu,\n", r[0]);
    }
    {   // (4) packet hop between two SMs
        u64 *box;
        unsigned *smid;
        CK(cudaMalloc(&box, 2 * 64 * 8));
        CK(cudaMemset(box, 0, 2 * 64 * 8));
        CK(cudaMalloc(&smid, 8));
        int iters = 5000;
        void *args[] = {&box, &iters, &o, &smid};
        CK(cudaLaunchCooperativeKernel((const void *)k_hop, dim3(2), dim3(32), args, 0, 0));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(r, o, 8, cudaMemcpyDeviceToHost));
        unsigned sm[2];
        CK(cudaMemcpy(sm, smid, 8, cudaMemcpyDeviceToHost));
        printf(" \"packet_hop_cycles_one_way\": %llu, \"hop_between_sms\": [%u, %u],\n", r[0], sm[0], sm[1]);
        // (6) distributed shared memory inside a cluster of 2
        k_hop_dsmem<<<2, 32>>>(iters, o, smid);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(r, o, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(sm, smid, 8, cudaMemcpyDeviceToHost));
        printf(" \"dsmem_hop_cycles_one_way\": %llu, \"dsmem_hop_between_sms\": [%u, %u]}\n", r[0], sm[0], sm[1]);
    }
    return 0;
}
