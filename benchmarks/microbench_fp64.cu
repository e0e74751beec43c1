// benchmarks/microbench_fp64.cu -- MEASUREMENT AID.  What do the fp64 instructions of the exact Godunov update cost
// on this GPU?  (a) issue cadence and latency of DFMA / DADD / DMUL / DSETP+FSEL / MUFU.RSQ64H, (b) the solve itself
// (v2_prep + eik_solve3_sorted from registers, no memory) at 16 and 32 warps per SM: the arithmetic-only ceiling of the
// batch forward kernel in warp slots per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -o /tmp/mb_fp64 benchmarks/microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../adtomo.jl_b200/csrc/kernels_fwd_v2.cuh"
using namespace adtomo;

template <int OP, int ILP>
__global__ void k_op(double *out, int iters, double seed) {
    double x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = seed + threadIdx.x * 1e-9 + i;
    const double a = 1.0000001, b = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) x[i] = __fma_rn(x[i], a, b);
            else if (OP == 1) x[i] = __dadd_rn(x[i], b);
            else if (OP == 2) x[i] = __dmul_rn(x[i], a);
            else if (OP == 3) x[i] = (x[i] < a) ? b : x[i];                      // DSETP + 2 FSEL
            else if (OP == 4) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i])); x[i] = y; }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}

__global__ void __launch_bounds__(512, 2) k_solve(double *out, int iters, double h, double seed) {
    double a = seed + (threadIdx.x & 7) * 0.01, b = seed + 0.3 + (threadIdx.x & 3) * 0.02, c = seed + 0.5, own = 1000.0, f = 0.2;
    double acc = 0.0, err = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        V2Vals V;
        V.off = 0; V.own = own; V.fv = f; V.dA = a; V.uA = a + 0.1; V.dW = b; V.uW = b + 0.05; V.dC = c; V.uC = c + 0.01; V.ref = 0.0;
        V2Prep Q;
        v2_prep(V, Q);
        double res = Q.own;
        if (Q.a1 < Q.own) {
            const double un = eik_solve3_sorted(Q.a1, Q.a2, Q.a3, Q.fv * h, Q.fv * Q.fv * h * h);
            if (un < Q.own) res = un;
        }
        acc += res;
        a += 1e-7; b += 2e-7; c += 3e-7;      // new inputs every iteration (3 DADD of overhead)
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + err;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}

template <int OP, int ILP>
static void run_op(const char *name, int ctas_per_sm, int nt, double *d) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4096;
    k_op<OP, ILP><<<sms * ctas_per_sm, nt>>>(d, iters, 1.5);
    k_op<OP, ILP><<<sms * ctas_per_sm, nt>>>(d, iters, 1.5);
    cudaDeviceSynchronize();
    double cyc; cudaMemcpy(&cyc, d + (size_t)sms * ctas_per_sm * nt, 8, cudaMemcpyDeviceToHost);
    const double warps_per_smsp = ctas_per_sm * nt / 32.0 / 4.0;
    const double inst = (double)iters * ILP * warps_per_smsp;   // warp instructions per SM sub-partition
    printf("{\"op\": \"%s\", \"ilp\": %d, \"warps_per_smsp\": %.0f, \"cycles_per_warp_inst_per_smsp\": %.3f, \"latency_if_ilp1_1warp\": %.2f}\n", name, ILP,
           warps_per_smsp, cyc / inst, cyc / iters / ILP);
}

int main() {
    double *d; cudaMalloc(&d, sizeof(double) * (148 * 2 * 1024 + 8));
    // latency: one warp per SMSP, one chain
    run_op<0, 1>("DFMA", 1, 128, d); run_op<1, 1>("DADD", 1, 128, d); run_op<2, 1>("DMUL", 1, 128, d);
    run_op<3, 1>("DSETP+2FSEL", 1, 128, d); run_op<4, 1>("MUFU.RSQ64H", 1, 128, d);
    // throughput: 8 warps per SMSP, 4 chains each
    run_op<0, 4>("DFMA", 2, 512, d); run_op<1, 4>("DADD", 2, 512, d); run_op<2, 4>("DMUL", 2, 512, d);
    run_op<3, 4>("DSETP+2FSEL", 2, 512, d); run_op<4, 4>("MUFU.RSQ64H", 2, 512, d);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int cps = 1; cps <= 2; cps++) {
        const int iters = 2048;
        k_solve<<<sms * cps, 512>>>(d, iters, 1.0, 0.7);
        k_solve<<<sms * cps, 512>>>(d, iters, 1.0, 0.7);
        cudaDeviceSynchronize();
        double cyc; cudaMemcpy(&cyc, d + (size_t)sms * cps * 512, 8, cudaMemcpyDeviceToHost);
        const double warps_per_smsp = cps * 4.0;
        printf("{\"op\": \"prep+solve (registers only)\", \"warps_per_smsp\": %.0f, \"cycles_per_warp_slot_per_smsp\": %.1f, \"cycles_per_iteration_of_one_warp\": %.1f}\n",
               warps_per_smsp, cyc / (iters * warps_per_smsp), cyc / iters);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
