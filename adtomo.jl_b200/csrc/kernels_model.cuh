// kernels_model.cuh -- the inversion drivers' model parametrisation, chain rule and regulariser on the device
// (scripts/inversion.jl:42-43,61,107-121; inversion_joint.jl:49-51,80,140-166).  Elementwise / small-stencil passes over
// the N model cells, once per loss evaluation: they exist so that a loss/gradient evaluation moves N optimiser variables
// in and N+1 doubles out and nothing else (the slowness field and its gradient never cross PCIe).
//
//   fvar   = 2 * sigmoid(x) - 1 + vel0                       inversion.jl:42-43
//   f_p    = scale_p / fvar                                  :61 (P: 1 ./ fvar), inversion_joint.jl:80 (S: pvs ./ fvar)
//   reg    = lambda * sum |fvar - box(fvar)|                 :107-121, box = mean over a periodic sh x sh x sv window
//   d/dx   = (sum_p -grad_f_p * scale_p / fvar^2 + lambda * (s - box(s))) * 2 sigmoid (1 - sigmoid),  s = sign(fvar - box(fvar))
//   d/dscale_p = sum_q grad_f_p[q] / fvar[q]
// Sums are two-stage with a fixed block count and a fixed order: results do not depend on scheduling.
#pragma once
#include "kernels_v0.cuh"

namespace adtomo {

constexpr int MODEL_RED_BLOCKS = 512;    // partial sums per reduction
constexpr int MODEL_NT = 256;

__global__ void k_model_fwd(const double *__restrict__ x, const double *__restrict__ vel0, double *__restrict__ fvar,
                            double *__restrict__ sig, const long long N) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += (long long)gridDim.x * blockDim.x) {
        const double s = 1.0 / (1.0 + exp(-x[q]));
        sig[q] = s;
        fvar[q] = 2.0 * s - 1.0 + vel0[q];
    }
}

__global__ void k_model_slowness(const double *__restrict__ fvar, double *__restrict__ f, const double scale, const long long N) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += (long long)gridDim.x * blockDim.x)
        f[q] = scale / fvar[q];
}

// g_fvar += -gf * f / fvar  (f = scale / fvar);  part[block] = sum gf / fvar.  grid = MODEL_RED_BLOCKS x MODEL_NT
__global__ void __launch_bounds__(MODEL_NT) k_model_phase_acc(const double *__restrict__ gf, const double *__restrict__ f,
                                                              const double *__restrict__ fvar, double *__restrict__ g_fvar,
                                                              double *__restrict__ part, const long long N) {
    __shared__ double red[MODEL_NT / 32];
    double acc = 0.0;
    for (long long q = (long long)blockIdx.x * MODEL_NT + threadIdx.x; q < N; q += (long long)gridDim.x * MODEL_NT) {
        const double g = gf[q], v = fvar[q];
        g_fvar[q] += -g * f[q] / v;
        acc += g / v;
    }
    acc = block_sum<MODEL_NT>(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// out[0] (+)= sum of part[0..n) in index order.  One block of 32 threads; lane 0 adds the lanes' strided sums in order.
__global__ void k_model_sum_parts(const double *__restrict__ part, const int n, double *__restrict__ out, const int accumulate,
                                  const double factor) {
    double acc = 0.0;
    for (int q = threadIdx.x; q < n; q += 32) acc += part[q];
    __shared__ double lanes[32];
    lanes[threadIdx.x] = acc;
    __syncwarp();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < 32; q++) t += lanes[q];
        out[0] = (accumulate ? out[0] : 0.0) + factor * t;
    }
}

// periodic box mean: out = box(a), window sh x sh x sv centred on the node (inversion.jl:107-118)
__global__ void k_model_box(const double *__restrict__ a, double *__restrict__ out, const int m, const int n, const int l,
                            const int sh, const int sv) {
    const long long N = (long long)m * n * l;
    const int h2 = (sh - 1) / 2, v2 = (sv - 1) / 2;
    const double cnt = (double)sh * sh * sv;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(q % l);
        const long long t = q / l;
        const int j = (int)(t % n), i = (int)(t / n);
        double acc = 0.0;
        for (int di = -h2; di <= h2; di++) {
            const int ii = ((i + di) % m + m) % m;
            for (int dj = -h2; dj <= h2; dj++) {
                const int jj = ((j + dj) % n + n) % n;
                const double *row = a + ((long long)ii * n + jj) * l;
                for (int dk = -v2; dk <= v2; dk++) acc += row[((k + dk) % l + l) % l];
            }
        }
        out[q] = acc / cnt;
    }
}

// s = sign(fvar - nvel);  part[block] = sum |fvar - nvel|
__global__ void __launch_bounds__(MODEL_NT) k_model_reg(const double *__restrict__ fvar, const double *__restrict__ nvel,
                                                        double *__restrict__ s, double *__restrict__ part, const long long N) {
    __shared__ double red[MODEL_NT / 32];
    double acc = 0.0;
    for (long long q = (long long)blockIdx.x * MODEL_NT + threadIdx.x; q < N; q += (long long)gridDim.x * MODEL_NT) {
        const double dlt = fvar[q] - nvel[q];
        s[q] = (dlt > 0.0) ? 1.0 : ((dlt < 0.0) ? -1.0 : 0.0);
        acc += fabs(dlt);
    }
    acc = block_sum<MODEL_NT>(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// grad_x = (g_fvar + lambda * (s - box(s))) * 2 sig (1 - sig); s / bs may be NULL (no regulariser)
__global__ void k_model_finish(const double *__restrict__ g_fvar, const double *__restrict__ s, const double *__restrict__ bs,
                               const double *__restrict__ sig, const double lambda, double *__restrict__ grad_x, const long long N) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += (long long)gridDim.x * blockDim.x) {
        double g = g_fvar[q];
        if (s) g += lambda * (s[q] - bs[q]);
        const double sg = sig[q];
        grad_x[q] = g * 2.0 * sg * (1.0 - sg);
    }
}

}  // namespace adtomo
