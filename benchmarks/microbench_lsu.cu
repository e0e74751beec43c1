// benchmarks/microbench_lsu.cu -- MEASUREMENT AID.  Do warp shuffles share the L1 data pipe with loads?
// The batch forward kernel is limited by l1tex__data_pipe_lsu_wavefronts (47 % busy on average, bursty); four of its
// eight loads per node could be replaced by shuffles of a neighbouring lane's value.  Cycles per warp instruction and SM
// for: SHFL alone, LDS.64 alone, LDG.64 (L1 hits) alone, and SHFL interleaved 1:1 with LDS / LDG.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb_lsu benchmarks/microbench_lsu.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, const double *g, int iters) {
    __shared__ double sm[1024 * 2];
    const int tid = threadIdx.x;
    sm[tid] = tid; sm[tid + 1024] = tid * 2.0;
    __syncthreads();
    double acc0 = tid, acc1 = 1.0, acc2 = 2.0, acc3 = 3.0;
    const double *gp = g + tid;
    const volatile double *sp = sm + tid;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 3 || MODE == 4) {          // 4 shuffles of 32-bit halves = two 64-bit values
            acc0 = __shfl_up_sync(0xffffffffu, acc0, 1);
            acc1 = __shfl_up_sync(0xffffffffu, acc1, 8);
        }
        if (MODE == 1 || MODE == 3) { acc2 += sp[0]; acc3 += sp[1024]; }                    // 2 LDS.64
        if (MODE == 2 || MODE == 4) { acc2 += __ldcg(gp) ; acc3 += gp[(it & 7) * 1024 + 1024]; }   // 2 LDG.64 (one L2, one L1)
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = acc0 + acc1 + acc2 + acc3;
    if (tid == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}

template <int MODE>
static void run(const char *name, double *d, const double *g) {
    const int iters = 4096, nt = 1024;
    k<MODE><<<148, nt>>>(d, g, iters);
    k<MODE><<<148, nt>>>(d, g, iters);
    cudaDeviceSynchronize();
    double cyc; cudaMemcpy(&cyc, d + 148 * nt, 8, cudaMemcpyDeviceToHost);
    printf("{\"mode\": \"%s\", \"cycles_per_iteration_per_SM_with_32_warps\": %.2f, \"cycles_per_warp_iteration\": %.3f}\n", name, cyc / iters, cyc / iters / 32.0);
}

int main() {
    double *d, *g;
    cudaMalloc(&d, sizeof(double) * (148 * 1024 + 8));
    cudaMalloc(&g, sizeof(double) * 1024 * 16);
    cudaMemset(g, 0, sizeof(double) * 1024 * 16);
    run<0>("4 SHFL.32 (two 64-bit values)", d, g);
    run<1>("2 LDS.64", d, g);
    run<2>("2 LDG.64", d, g);
    run<3>("4 SHFL.32 + 2 LDS.64", d, g);
    run<4>("4 SHFL.32 + 2 LDG.64", d, g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
