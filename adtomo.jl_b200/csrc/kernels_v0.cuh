// kernels_v0.cuh -- shared device helpers, the row-major FALLBACK 3D forward kernel (grids whose
// sheets do not fit the level-major kernel), the fused-step helpers (sparse sources, receiver
// sampling, misfit) and the 2D kernels.
//
// k_fwd3d_v0: one CTA per source, hyperplane (level-set) ordering of every directional Gauss-Seidel
// sweep, fields in the reference's row-major layout in global memory.
//
// Why hyperplanes: in a sweep with directions (di,dj,dk) node (I,J,K) (sweep coordinates, i.e.
// reflected so that the sweep ascends) reads the NEW values of (I-1,J,K),(I,J-1,K),(I,J,K-1)
// and the OLD values of the three opposite neighbours (Eikonal3D.cpp:44-55).  All nodes with
// I+J+K = s are mutually independent and depend only on levels s-1 (new) and s+1 (old), so
// processing levels in increasing s reproduces the serial lexicographic sweep bit for bit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "eik_core.h"

namespace adtomo {

struct Dims3 {
    int m, n, l;
    long long N;
};

__constant__ int c_dirs3[8][3] = {   // Eikonal3D.cpp:59-68
    {1, 1, 1}, {-1, 1, 1}, {-1, -1, 1}, {1, -1, 1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, -1}};
__constant__ int c_dirs2[4][2] = {{1, 1}, {-1, 1}, {-1, -1}, {1, -1}};   // Eikonal.h:74-77

template <int NT>
__device__ __forceinline__ double block_max(double v, double *red) {
    for (int o = 16; o > 0; o >>= 1) {
        double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = (w > v) ? w : v;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < NT / 32; w++) r = (red[w] > r) ? red[w] : r;
    return r;
}

template <int NT>
__device__ __forceinline__ double block_sum(double v, double *red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < NT / 32; w++) r += red[w];
    return r;
}

template <int NT>
__device__ __forceinline__ int block_sum_int(int v, int *red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = 0;
    for (int w = 0; w < NT / 32; w++) r += red[w];
    return r;
}

// ------------------------------------------------------------------------------------------
// 3D forward, v0
// ------------------------------------------------------------------------------------------
template <int NT>
__device__ void sweep3d_v0(double *u, const double *__restrict__ f, const Dims3 d, const double h, const int di,
                           const int dj, const int dk) {
    const int m = d.m, n = d.n, l = d.l;
    const int nlev = m + n + l - 2;
    for (int s = 0; s < nlev; s++) {
        const int Jlo = max(0, s - (m - 1) - (l - 1));
        const int Jhi = min(n - 1, s);
        const int cnt = (Jhi - Jlo + 1) * l;
        for (int idx = threadIdx.x; idx < cnt; idx += NT) {
            const int Jr = idx / l;
            const int K = idx - Jr * l;
            const int J = Jlo + Jr;
            const int I = s - J - K;
            if (I < 0 || I >= m) continue;
            const int i = di > 0 ? I : m - 1 - I;
            const int j = dj > 0 ? J : n - 1 - J;
            const int k = dk > 0 ? K : l - 1 - K;
            const long long id = ((long long)i * n + j) * l + k;
            const long long si = (long long)n * l;
            const double ux = i == 0 ? u[id + si] : (i == m - 1 ? u[id - si] : eik_min(u[id + si], u[id - si]));
            const double uy = j == 0 ? u[id + l] : (j == n - 1 ? u[id - l] : eik_min(u[id + l], u[id - l]));
            const double uz = k == 0 ? u[id + 1] : (k == l - 1 ? u[id - 1] : eik_min(u[id + 1], u[id - 1]));
            const double uo = u[id];
            const double un = eik_solve3(ux, uy, uz, f[id], h);
            if (un < uo) u[id] = un;   // u = std::min(u_new, u)
        }
        __syncthreads();
    }
}

// One CTA per source (grid-strided over sources).  u holds u0 on entry.  scratch: gridDim.x * N.
template <int NT>
__global__ void __launch_bounds__(NT) k_fwd3d_v0(double *__restrict__ U, double *__restrict__ scratch,
                                                 const double *__restrict__ f, const Dims3 d, const double h,
                                                 const double tol, const int max_rounds, const int S,
                                                 int *__restrict__ rounds, double *__restrict__ errs) {
    __shared__ double red[NT / 32];
    double *uo = scratch + (long long)blockIdx.x * d.N;
    for (int src = blockIdx.x; src < S; src += gridDim.x) {
        double *u = U + (long long)src * d.N;
        int r = 0;
        bool conv = false;
        while (r < max_rounds) {
            for (long long q = threadIdx.x; q < d.N; q += NT) uo[q] = u[q];
            __syncthreads();
            for (int sw = 0; sw < 8; sw++) sweep3d_v0<NT>(u, f, d, h, c_dirs3[sw][0], c_dirs3[sw][1], c_dirs3[sw][2]);
            double e = 0.0;
            for (long long q = threadIdx.x; q < d.N; q += NT) {
                const double dd = fabs(u[q] - uo[q]);
                e = (e < dd) ? dd : e;
            }
            e = block_max<NT>(e, red);
            if (threadIdx.x == 0 && errs) errs[(long long)src * max_rounds + r] = e;
            r++;
            if (e < tol) { conv = true; break; }
        }
        if (threadIdx.x == 0 && rounds) rounds[src] = conv ? r : -r;   // negative: cap hit
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// adjoint parent code (shared by the 3D wavefront kernels and the 2D kernel).  code byte per node: bits [2a,2a+1] for axis a in {i,j,k}: 0 = axis inactive,
// 1 = upwind parent is the -1 neighbour, 2 = parent is the +1 neighbour; bit 6 = value final;
// bit 7 = pinned (row of Z in Eikonal3D.cpp:126-130,168-171: x = 0).
// ------------------------------------------------------------------------------------------
#define ADJ_DONE 0x40
#define ADJ_PIN 0x80

__device__ __forceinline__ unsigned adj_axis_code(const double *u, long long id, int c, int lim, long long stride,
                                                  double ui) {
    // parent selection of Eikonal3D.cpp:139-147: mirror at the edges, interior tie -> +1.
    int side;   // 1: -1 neighbour, 2: +1 neighbour
    if (c == 0) side = 2;
    else if (c == lim - 1) side = 1;
    else side = (u[id + stride] > u[id - stride]) ? 1 : 2;
    const double a = side == 1 ? u[id - stride] : u[id + stride];
    return (ui > a) ? (unsigned)side : 0u;
}

// ------------------------------------------------------------------------------------------
// sparse source initialisation, receiver sampling, misfit (scripts/inversion.jl:46-105)
// ------------------------------------------------------------------------------------------
__global__ void k_fill(double *__restrict__ p, const long long n, const double v) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
        p[q] = v;
}

// One thread per sparse entry, in list order per source (later entries win, like the Julia
// assignments of inversion.jl:52-60 -- duplicates carry identical values there).
// Device-resident source / receiver tables cannot be checked on the host: flag[0] |= 1 for a source index outside the
// grid, |= 2 for a receiver outside it (the scatter / sampling kernels would write or read out of bounds).
__global__ void k_validate_tables(const int *__restrict__ src_idx, const int nnz, const long long N,
                                  const double *__restrict__ rcv, const int E, const int m, const int n, const int l,
                                  int *__restrict__ flag) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0;
    if (q < nnz && (src_idx[q] < 0 || src_idx[q] >= N)) bad |= 1;
    if (q < E) {
        const double x = rcv[3 * q], y = rcv[3 * q + 1], z = rcv[3 * q + 2];
        if (!(x >= 0 && x <= m - 1 && y >= 0 && y <= n - 1 && z >= 0 && z <= l - 1)) bad |= 2;
    }
    if (bad) atomicOr(flag, bad);
}

__global__ void k_scatter_sources(double *__restrict__ U0, const int *__restrict__ src_ptr,
                                  const int *__restrict__ src_idx, const double *__restrict__ src_val,
                                  const long long N, const int S) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    for (int q = src_ptr[s]; q < src_ptr[s + 1]; q++) U0[(long long)s * N + src_idx[q]] = src_val[q];
}

// Trilinear rule of inversion.jl:64-95: floor/ceil corners; an integer coordinate uses that
// node alone (weight 1), otherwise weights (x2-x) and (x-x1).
__device__ __forceinline__ void axis_w(double x, int &a, int &b, double &wa, double &wb) {
    const double fl = floor(x), ce = ceil(x);
    a = (int)fl;
    b = (int)ce;
    if (a == b) { wa = 1.0; wb = 0.0; }
    else { wa = ce - x; wb = x - fl; }
}

// grid: (ceil(E/NT), S).  Accumulates misfit per source into mis[S] and scatters d(misfit)/du
// into G (S*N, zero-initialised).
template <int NT>
__global__ void __launch_bounds__(NT) k_misfit(const double *__restrict__ U, double *__restrict__ G,
                                               const double *__restrict__ rcv, const double *__restrict__ uobs,
                                               const double *__restrict__ qua, double *__restrict__ mis,
                                               const Dims3 d, const int E, const int want_grad) {
    __shared__ double red[NT / 32];
    const int s = blockIdx.y;
    const int e = blockIdx.x * NT + threadIdx.x;
    const double *u = U + (long long)s * d.N;
    double part = 0.0;
    if (e < E) {
        const double ob = uobs[(long long)s * E + e];
        if (ob != -1.0) {
            int x1, x2, y1, y2, z1, z2;
            double wx1, wx2, wy1, wy2, wz1, wz2;
            axis_w(rcv[3 * e + 0], x1, x2, wx1, wx2);
            axis_w(rcv[3 * e + 1], y1, y2, wy1, wy2);
            axis_w(rcv[3 * e + 2], z1, z2, wz1, wz2);
            const long long n = d.n, l = d.l;
#define UAT(a, b, c) u[((long long)(a) * n + (b)) * l + (c)]
            const double tx11 = (x1 == x2) ? UAT(x1, y1, z1) : wx1 * UAT(x1, y1, z1) + wx2 * UAT(x2, y1, z1);
            const double tx12 = (x1 == x2) ? UAT(x1, y1, z2) : wx1 * UAT(x1, y1, z2) + wx2 * UAT(x2, y1, z2);
            const double tx21 = (x1 == x2) ? UAT(x1, y2, z1) : wx1 * UAT(x1, y2, z1) + wx2 * UAT(x2, y2, z1);
            const double tx22 = (x1 == x2) ? UAT(x1, y2, z2) : wx1 * UAT(x1, y2, z2) + wx2 * UAT(x2, y2, z2);
#undef UAT
            const double txy1 = (y1 == y2) ? tx11 : wy1 * tx11 + wy2 * tx21;
            const double txy2 = (y1 == y2) ? tx12 : wy1 * tx12 + wy2 * tx22;
            const double t = (z1 == z2) ? txy1 : wz1 * txy1 + wz2 * txy2;
            const double w = qua[(long long)s * E + e];
            const double r = ob - t;
            part = w * (r * r);
            if (want_grad) {
                const double dt = -2.0 * w * r;   // d/dt of w*(ob-t)^2
                double *g = G + (long long)s * d.N;
                const int xs[2] = {x1, x2}, ys[2] = {y1, y2}, zs[2] = {z1, z2};
                const double wxs[2] = {wx1, wx2}, wys[2] = {wy1, wy2}, wzs[2] = {wz1, wz2};
                for (int a = 0; a < 2; a++)
                    for (int b = 0; b < 2; b++)
                        for (int c = 0; c < 2; c++) {
                            const double ww = wxs[a] * wys[b] * wzs[c];
                            if (ww != 0.0) atomicAdd(&g[((long long)xs[a] * n + ys[b]) * l + zs[c]], dt * ww);
                        }
            }
        }
    }
    part = block_sum<NT>(part, red);
    if (threadIdx.x == 0 && part != 0.0) atomicAdd(&mis[s], part);
}

// ------------------------------------------------------------------------------------------
// 2D forward / adjoint: one CTA per source, the field lives in shared memory when it fits.
// ------------------------------------------------------------------------------------------
// Eikonal.h:54-93.  work: per-CTA, 2*N2 doubles (u, u_old) in shared or global memory.
template <int NT>
__global__ void __launch_bounds__(NT) k_fwd2d(double *__restrict__ Uout, const double *__restrict__ f,
                                              const int m, const int n, const double h,
                                              const int *__restrict__ IX, const int *__restrict__ JX, const int S,
                                              int *__restrict__ rounds, double *__restrict__ gwork,
                                              const int use_smem) {
    extern __shared__ double sm2[];
    __shared__ double red[NT / 32];
    const int w = m + 1, N2 = (m + 1) * (n + 1);
    double *u = use_smem ? sm2 : gwork + (long long)blockIdx.x * 2 * N2;
    double *uo = u + N2;
    for (int src = blockIdx.x; src < S; src += gridDim.x) {
        const int ix = IX[src], jx = JX[src];
        for (int q = threadIdx.x; q < N2; q += NT) {
            const double v = (q == jx * w + ix) ? 0.0 : 100000.0;
            u[q] = v;
            uo[q] = v;
        }
        __syncthreads();
        int r = 0;
        bool conv = false;
        while (r < 100) {
            for (int sw = 0; sw < 4; sw++) {
                const int di = c_dirs2[sw][0], dj = c_dirs2[sw][1];
                // sweep coordinates (I,J); the reference loops i (x) outer, j inner: any order that
                // respects (I-1,J),(I,J-1) first is identical -> anti-diagonals I+J = s.
                for (int s = 0; s <= m + n; s++) {
                    const int Ilo = max(0, s - n), Ihi = min(m, s);
                    for (int I = Ilo + threadIdx.x; I <= Ihi; I += NT) {
                        const int J = s - I;
                        const int i = di > 0 ? I : m - I;
                        const int j = dj > 0 ? J : n - J;
                        if (i == ix && j == jx) continue;
                        const int id = j * w + i;
                        const double a = i == 0 ? u[id + 1] : (i == m ? u[id - 1] : eik_min(u[id + 1], u[id - 1]));
                        const double b = j == 0 ? u[id + w] : (j == n ? u[id - w] : eik_min(u[id - w], u[id + w]));
                        const double un = eik_solve2(a, b, f[id], h);
                        const double uc = u[id];
                        if (un < uc) u[id] = un;
                    }
                    __syncthreads();
                }
            }
            double num = 0.0, den = 0.0;
            for (int q = threadIdx.x; q < N2; q += NT) {
                const double dd = u[q] - uo[q];
                num += dd * dd;
                den += uo[q] * uo[q];
            }
            num = block_sum<NT>(num, red);
            den = block_sum<NT>(den, red);
            r++;
            const double err = sqrt(num) / sqrt(den);
            if (err < 1e-8) { conv = true; break; }
            for (int q = threadIdx.x; q < N2; q += NT) uo[q] = u[q];
            __syncthreads();
        }
        for (int q = threadIdx.x; q < N2; q += NT) Uout[(long long)src * N2 + q] = u[q];
        if (threadIdx.x == 0 && rounds) rounds[src] = conv ? r : -r;
        __syncthreads();
    }
}

// Eikonal.h:95-200 by causal sweeps.  work per CTA: N2 doubles (x) + N2 bytes (code).
template <int NT>
__global__ void __launch_bounds__(NT) k_adj2d(double *__restrict__ GF, const double *__restrict__ G,
                                              const double *__restrict__ U, const double *__restrict__ f,
                                              const int m, const int n, const double h,
                                              const int *__restrict__ IX, const int *__restrict__ JX, const int S,
                                              int *__restrict__ status, double *__restrict__ gwork,
                                              const int use_smem) {
    extern __shared__ double sm2[];
    __shared__ int redi[NT / 32];
    const int w = m + 1, N2 = (m + 1) * (n + 1);
    const long long per = (long long)N2 + (N2 + 7) / 8;   // doubles per CTA of workspace
    double *x = use_smem ? sm2 : gwork + (long long)blockIdx.x * per;
    unsigned char *code = (unsigned char *)(x + N2);
    for (int src = blockIdx.x; src < S; src += gridDim.x) {
        const int ix = IX[src], jx = JX[src];
        const double *u = U + (long long)src * N2;
        const double *g = G + (long long)src * N2;
        int rem = 0, flagged = 0;
        for (int q = threadIdx.x; q < N2; q += NT) {
            const int j = q / w, i = q - j * w;
            unsigned cd;
            if (i == ix && j == jx) cd = ADJ_PIN | ADJ_DONE;   // identity row; dFdf[src] = 0
            else {
                cd = adj_axis_code(u, q, i, m + 1, 1, u[q]) | (adj_axis_code(u, q, j, n + 1, w, u[q]) << 2);
                if (cd == 0) { cd = ADJ_PIN | ADJ_DONE; flagged++; }   // empty row: singular in the reference
                else rem++;
            }
            code[q] = (unsigned char)cd;
            x[q] = 0.0;
        }
        rem = block_sum_int<NT>(rem, redi);
        flagged = block_sum_int<NT>(flagged, redi);
        int sweeps = 0;
        bool stalled = false;
        while (rem > 0 && !stalled) {
            int round_fin = 0;
            for (int sw = 3; sw >= 0 && rem > 0; sw--) {
                const int di = -c_dirs2[sw][0], dj = -c_dirs2[sw][1];
                int fin = 0;
                for (int s = 0; s <= m + n; s++) {
                    const int Ilo = max(0, s - n), Ihi = min(m, s);
                    for (int I = Ilo + threadIdx.x; I <= Ihi; I += NT) {
                        const int J = s - I;
                        const int i = di > 0 ? I : m - I;
                        const int j = dj > 0 ? J : n - J;
                        const int id = j * w + i;
                        const unsigned cd = code[id];
                        if (cd & ADJ_DONE) continue;
                        const double ui = u[id];
                        double acc = 0.0;
                        bool ok = true;
#define ADJ_CHILD2(cond, off, shift, want)                                           \
    if (cond) {                                                                      \
        const unsigned cc = code[id + (off)];                                        \
        if (((cc >> (shift)) & 3u) == (want)) {                                      \
            if (!(cc & ADJ_DONE)) ok = false;                                        \
            else acc += 2 * (u[id + (off)] - ui) * x[id + (off)];                    \
        }                                                                            \
    }
                        ADJ_CHILD2(i > 0, -1, 0, 2u)
                        ADJ_CHILD2(i < m, 1, 0, 1u)
                        ADJ_CHILD2(j > 0, -w, 2, 2u)
                        ADJ_CHILD2(j < n, w, 2, 1u)
#undef ADJ_CHILD2
                        if (!ok) continue;
                        double D = 0.0;
                        const unsigned ci = cd & 3u, cj = (cd >> 2) & 3u;
                        if (ci) D += 2 * (ui - u[ci == 1 ? id - 1 : id + 1]);
                        if (cj) D += 2 * (ui - u[cj == 1 ? id - w : id + w]);
                        x[id] = (g[id] + acc) / D;
                        code[id] = (unsigned char)(cd | ADJ_DONE);
                        fin++;
                    }
                    __syncthreads();
                }
                fin = block_sum_int<NT>(fin, redi);
                rem -= fin;
                round_fin += fin;
                sweeps++;
            }
            if (round_fin == 0) stalled = true;
        }
        for (int q = threadIdx.x; q < N2; q += NT) {
            double dFdf = -2 * f[q] * h * h;
            if (q == jx * w + ix) dFdf = 0.0;
            GF[(long long)src * N2 + q] = -x[q] * dFdf;
        }
        if (threadIdx.x == 0 && status) status[src] = (rem > 0 || flagged) ? -sweeps - 1 : sweeps;
        __syncthreads();
    }
}

// sum over sources: out[q] = sum_s in[s*N+q]
__global__ void k_sum_sources(const double *__restrict__ in, double *__restrict__ out, const long long N, const int S) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int s = 0; s < S; s++) acc += in[(long long)s * N + q];
        out[q] = acc;
    }
}

// a += b
__global__ void k_axpy(double *__restrict__ a, const double *__restrict__ b, const long long n) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
        a[q] += b[q];
}

}  // namespace adtomo
