// benchmarks/microbench_l1store.cu -- MEASUREMENT AID.  Does a global store leave the line readable from L1?
// The batch forward kernel reads three UPWIND values per node that the same CTA stored one level (one __syncthreads)
// earlier; ncu shows an L1 hit rate of 49 % for its loads.  One warp: (a) load a line (L1 now holds it), (b) store to it,
// (c) barrier, (d) timed load with a dependent use.  Reported: latency of (d) against a plain L1 hit and an L2 hit (ld.cg).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb_l1store benchmarks/microbench_l1store.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double ld_ca(const double *p) { double v; asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ double ld_cg(const double *p) { double v; asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_wb(double *p, double v) { asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

__global__ void k(double *buf, long long *out, double *sink, int mode) {
    const int lane = threadIdx.x;
    double *p = buf + lane * 16;            // each lane its own 128-byte line
    double v = ld_ca(p);                    // warm: L1 holds the line
    v += ld_ca(p + 1);
    if (mode == 1) st_wb(p, v);             // store to the word that is read below
    if (mode == 3) st_wb(p + 1, v);         // store to a neighbouring word of the same sector
    if (mode == 4) st_wb(p + 8, v);         // store to another sector of the same line
    __syncthreads();
    __shared__ volatile double sh[32];
    long long t0 = clock64();
    double w = (mode == 2) ? ld_cg(p) : ld_ca(p);
    sh[lane] = w;                           // dependent use: the clock is read after the value has arrived
    long long t1 = clock64();
    out[mode * 32 + lane] = t1 - t0;
    sink[lane] = w + v;
}

int main() {
    double *d, *s; long long *o;
    cudaMalloc(&d, 8 * 32 * 16 * 4);
    cudaMalloc(&s, 8 * 64);
    cudaMalloc(&o, 8 * 32 * 8);
    static double h[32 * 16 * 4];
    for (int i = 0; i < 32 * 16 * 4; i++) h[i] = (double)i;
    const char *names[5] = {"load after load (L1 hit)", "load after store to the same word", "ld.cg (L2 hit)",
                            "load after store to a neighbouring word of the sector", "load after store to another sector of the line"};
    for (int rep = 0; rep < 2; rep++)
        for (int mode = 0; mode < 5; mode++) {
            cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
            k<<<1, 32>>>(d, o, s, mode);
            cudaDeviceSynchronize();
            long long r[32];
            cudaMemcpy(r, o + mode * 32, sizeof r, cudaMemcpyDeviceToHost);
            if (rep == 1) printf("{\"case\": \"%s\", \"cycles\": %lld}\n", names[mode], r[0]);
        }
    return 0;
}
