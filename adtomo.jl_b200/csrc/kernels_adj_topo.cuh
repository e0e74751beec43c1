// kernels_adj_topo.cuh -- the 3D adjoint as ONE causality-ordered pass (level-synchronous
// topological wavefront), replacing the reference's sparse LU (Eikonal3D.cpp:186-193).
//
// A^T x = g with the reference's assembly rules is a permuted triangular system: node p receives
// from node c only if p is c's selected upwind parent on some axis, and then u_c > u_p strictly.
// So x_p = (g_p + sum_{children c} 2(u_c-u_p) x_c) / D_p can be evaluated as soon as all children
// are final.  We count children per node, start from the nodes without children, and process
// "ready" nodes wave by wave; every node is evaluated exactly once (O(N) work, independent of the
// number of sweeps the forward solve needed), children are gathered in a fixed neighbour order,
// so the result is deterministic.  One CTA per source; the ready queue lives in global memory.
#pragma once
#include "kernels_v0.cuh"

namespace adtomo {

// pass 2 of the setup: children counts and the initial ready set.  cnt is a byte per node;
// pinned nodes get 0xFF so that decrements never make them "ready".
__global__ void k_adj3d_count(const unsigned char *__restrict__ code, unsigned char *__restrict__ cnt,
                              int *__restrict__ Q, int *__restrict__ qtail,
                              const Dims3 d, const int S) {
    const int src = blockIdx.y;
    const long long base = (long long)src * d.N;
    const unsigned char *cd_ = code + base;
    const int N = (int)d.N, n = d.n, l = d.l, nl = d.n * d.l;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < N; id += gridDim.x * blockDim.x) {
        const unsigned cd = cd_[id];
        if (cd & ADJ_PIN) { cnt[base + id] = 0xFF; continue; }
        const int i = id / nl;
        const int r = id - i * nl;
        const int j = r / l;
        const int k = r - j * l;
        int c = 0;
        if (i > 0 && ((cd_[id - nl] >> 0) & 3u) == 2u) c++;
        if (i < d.m - 1 && ((cd_[id + nl] >> 0) & 3u) == 1u) c++;
        if (j > 0 && ((cd_[id - l] >> 2) & 3u) == 2u) c++;
        if (j < n - 1 && ((cd_[id + l] >> 2) & 3u) == 1u) c++;
        if (k > 0 && ((cd_[id - 1] >> 4) & 3u) == 2u) c++;
        if (k < l - 1 && ((cd_[id + 1] >> 4) & 3u) == 1u) c++;
        cnt[base + id] = (unsigned char)c;
        if (c == 0) {
            const int pos = atomicAdd(&qtail[src], 1);
            Q[base + pos] = id;
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(NT) k_adj3d_topo(const double *__restrict__ U, const double *__restrict__ G,
                                                   double *X, const unsigned char *__restrict__ code,
                                                   unsigned int *cnt32, int *Q, const int *__restrict__ qtail,
                                                   const int *__restrict__ nfree, const Dims3 d, const int S,
                                                   int *__restrict__ status) {
    __shared__ int s_tail;
    const int m = d.m, n = d.n, l = d.l;
    const long long si = (long long)n * l;
    for (int src = blockIdx.x; src < S; src += gridDim.x) {
        const long long base = (long long)src * d.N;
        const double *u = U + base;
        const double *g = G + base;
        double *x = X + base;
        const unsigned char *cd_ = code + base;
        int *q = Q + base;
        int head = 0, tail = qtail[src], waves = 0;
        if (threadIdx.x == 0) s_tail = tail;
        __syncthreads();
        while (head < tail) {
            for (int t = head + threadIdx.x; t < tail; t += NT) {
                const int id = q[t];
                const int k = id % l;
                const int tt = id / l;
                const int j = tt % n;
                const int i = tt / n;
                const unsigned cd = cd_[id];
                const double ui = u[id];
                double acc = 0.0;
                // children, fixed order: i-1, i+1, j-1, j+1, k-1, k+1
#define TOPO_CHILD(cond, off, shift, want)                                       \
    if (cond) {                                                                  \
        const unsigned cc = cd_[id + (off)];                                     \
        if (((cc >> (shift)) & 3u) == (want)) acc += 2.0 * (u[id + (off)] - ui) * x[id + (off)]; \
    }
                TOPO_CHILD(i > 0, -si, 0, 2u)
                TOPO_CHILD(i < m - 1, si, 0, 1u)
                TOPO_CHILD(j > 0, -(long long)l, 2, 2u)
                TOPO_CHILD(j < n - 1, (long long)l, 2, 1u)
                TOPO_CHILD(k > 0, -1LL, 4, 2u)
                TOPO_CHILD(k < l - 1, 1LL, 4, 1u)
#undef TOPO_CHILD
                const unsigned ci = cd & 3u, cj = (cd >> 2) & 3u, ck = (cd >> 4) & 3u;
                const long long pi = ci == 1 ? id - si : id + si;
                const long long pj = cj == 1 ? id - l : id + l;
                const long long pk = ck == 1 ? id - 1 : id + 1;
                double D = 0.0;
                if (ci) D += 2.0 * (ui - u[pi]);
                if (cj) D += 2.0 * (ui - u[pj]);
                if (ck) D += 2.0 * (ui - u[pk]);
                x[id] = (g[id] + acc) / D;
                // release the parents
#define TOPO_RELEASE(active, p)                                                  \
    if (active) {                                                                \
        const long long gq = base + (p);                                         \
        const unsigned sh = (unsigned)(gq & 3) * 8u;                             \
        const unsigned old = atomicSub(&cnt32[gq >> 2], 1u << sh);               \
        if (((old >> sh) & 0xFFu) == 1u) {                                       \
            const int pos = atomicAdd(&s_tail, 1);                               \
            q[pos] = (int)(p);                                                   \
        }                                                                        \
    }
                TOPO_RELEASE(ci, pi)
                TOPO_RELEASE(cj, pj)
                TOPO_RELEASE(ck, pk)
#undef TOPO_RELEASE
            }
            __syncthreads();
            const int nt = s_tail;
            __syncthreads();
            head = tail;
            tail = nt;
            waves++;
        }
        if (threadIdx.x == 0 && status) status[src] = (tail == nfree[src]) ? waves + 1 : -(waves + 1);
        __syncthreads();
    }
}

}  // namespace adtomo
