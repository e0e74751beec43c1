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
//
// Data layout (the wavefront touches scattered nodes, so what counts is 32-byte sectors per node):
//   UX  double2 {u, x}   a child contributes through ONE sector; x is stored next to the node's u
//   GD  double2 {g, D}   right-hand side and diagonal, precomputed by the coalesced setup pass
//   CM  uint16  low byte: parent side per axis + pinned flag (reference rules, tie -> +1);
//               high byte: which of the 6 neighbours are children (no index arithmetic, no bounds
//               checks and no neighbour-code reads in the wavefront)
//   CNT uint8   children still pending (decremented with 32-bit atomics on the containing word)
#pragma once
#include "kernels_v0.cuh"

namespace adtomo {

// pass 1 (coalesced): parent code, diagonal, interleaved copies, grad_u0.  grid = (blocks, S)
__global__ void k_adj3d_setup2(const double *__restrict__ U, const double *__restrict__ U0,
                               const double *__restrict__ G, double2 *__restrict__ UX, double2 *__restrict__ GD,
                               double *__restrict__ GU0, unsigned char *__restrict__ code,
                               int *__restrict__ remaining, const Dims3 d, const int S) {
    const int src = blockIdx.y;
    const long long base = (long long)src * d.N;
    const double *u = U + base;
    const int N = (int)d.N, n = d.n, l = d.l, nl = d.n * d.l;
    int mycount = 0;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < N; id += gridDim.x * blockDim.x) {
        const double ui = u[id];
        const double gi = G[base + id];
        const bool same = (ui == U0[base + id]);
        if (GU0) GU0[base + id] = same ? gi : 0.0;       // Eikonal3D.cpp:106-110
        unsigned cd;
        double D = 0.0;
        if (same) cd = ADJ_PIN | ADJ_DONE;
        else {
            const int i = id / nl;
            const int r = id - i * nl;
            const int j = r / l;
            const int k = r - j * l;
            const unsigned ci = adj_axis_code(u, id, i, d.m, nl, ui), cj = adj_axis_code(u, id, j, n, l, ui),
                           ck = adj_axis_code(u, id, k, l, 1, ui);
            cd = ci | (cj << 2) | (ck << 4);
            if (cd == 0) cd = ADJ_PIN | ADJ_DONE;
            else {
                if (ci) D += 2.0 * (ui - u[ci == 1 ? id - nl : id + nl]);
                if (cj) D += 2.0 * (ui - u[cj == 1 ? id - l : id + l]);
                if (ck) D += 2.0 * (ui - u[ck == 1 ? id - 1 : id + 1]);
            }
        }
        code[base + id] = (unsigned char)cd;
        UX[base + id] = make_double2(ui, 0.0);
        GD[base + id] = make_double2(gi, D);
        if (!(cd & ADJ_DONE)) mycount++;
    }
    for (int o = 16; o > 0; o >>= 1) mycount += __shfl_xor_sync(0xffffffffu, mycount, o);
    if ((threadIdx.x & 31) == 0 && mycount) atomicAdd(&remaining[src], mycount);
}

// pass 2 (coalesced): children mask + count, initial ready set.  Pinned nodes get count 0xFF so
// that decrements never make them "ready".  grid = (blocks, S)
__global__ void k_adj3d_count2(const unsigned char *__restrict__ code, unsigned short *__restrict__ CM,
                               unsigned char *__restrict__ cnt, int *__restrict__ Q, int *__restrict__ qtail,
                               const Dims3 d, const int S) {
    const int src = blockIdx.y;
    const long long base = (long long)src * d.N;
    const unsigned char *cd_ = code + base;
    const int N = (int)d.N, n = d.n, l = d.l, nl = d.n * d.l;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < N; id += gridDim.x * blockDim.x) {
        const unsigned cd = cd_[id];
        if (cd & ADJ_PIN) { cnt[base + id] = 0xFF; CM[base + id] = (unsigned short)cd; continue; }
        const int i = id / nl;
        const int r = id - i * nl;
        const int j = r / l;
        const int k = r - j * l;
        unsigned mask = 0;
        // the -1 neighbour is my child iff its parent on this axis is its +1 side (2); the +1 neighbour iff 1
        if (i > 0 && ((cd_[id - nl] >> 0) & 3u) == 2u) mask |= 1u;
        if (i < d.m - 1 && ((cd_[id + nl] >> 0) & 3u) == 1u) mask |= 2u;
        if (j > 0 && ((cd_[id - l] >> 2) & 3u) == 2u) mask |= 4u;
        if (j < n - 1 && ((cd_[id + l] >> 2) & 3u) == 1u) mask |= 8u;
        if (k > 0 && ((cd_[id - 1] >> 4) & 3u) == 2u) mask |= 16u;
        if (k < l - 1 && ((cd_[id + 1] >> 4) & 3u) == 1u) mask |= 32u;
        const int c = __popc(mask);
        cnt[base + id] = (unsigned char)c;
        CM[base + id] = (unsigned short)(cd | (mask << 8));
        if (c == 0) {
            const int pos = atomicAdd(&qtail[src], 1);
            Q[base + pos] = id;
        }
    }
}

template <int NT, bool AGG>
__global__ void __launch_bounds__(NT) k_adj3d_topo2(double2 *UX, const double2 *__restrict__ GD,
                                                    const unsigned short *__restrict__ CM, unsigned int *cnt32,
                                                    int *Q, const int *__restrict__ qtail,
                                                    const int *__restrict__ nfree, const Dims3 d, const int S,
                                                    int *__restrict__ status) {
    __shared__ int s_tail;
    const int l = d.l;
    const int nl = d.n * d.l;
    for (int src = blockIdx.x; src < S; src += gridDim.x) {
        const long long base = (long long)src * d.N;
        double2 *ux = UX + base;
        const double2 *gd = GD + base;
        const unsigned short *cm_ = CM + base;
        int *q = Q + base;
        int head = 0, tail = qtail[src], waves = 0;
        if (threadIdx.x == 0) s_tail = tail;
        __syncthreads();
        while (head < tail) {
            // warp-uniform trip count: the lanes of a warp push their newly ready parents with ONE shared-memory
            // atomic per axis.  (No measurable effect on the 256-source bench batch, 35 ms either way: a wave moves
            // ~12 scattered 32-byte sectors per node -- ~100 GB per step -- and is bound by that traffic.)
            for (int t0 = head + (threadIdx.x & ~31); t0 < tail; t0 += NT) {
                const int t = t0 + (threadIdx.x & 31);
                int rp0 = -1, rp1 = -1, rp2 = -1;      // parents that became ready through this node
                if (t < tail) {
                    const int id = q[t];
                    const unsigned cm = cm_[id];
                    const double ui = ux[id].x;
                    const double2 g = gd[id];
                    double acc = 0.0;
                    // children in the fixed order i-1, i+1, j-1, j+1, k-1, k+1 (all final by construction)
#define TOPO_CHILD(bit, off)                                        \
    if (cm & ((bit) << 8)) {                                        \
        const double2 c = ux[id + (off)];                           \
        acc += 2.0 * (c.x - ui) * c.y;                              \
    }
                    TOPO_CHILD(1u, -nl)
                    TOPO_CHILD(2u, nl)
                    TOPO_CHILD(4u, -l)
                    TOPO_CHILD(8u, l)
                    TOPO_CHILD(16u, -1)
                    TOPO_CHILD(32u, 1)
#undef TOPO_CHILD
                    ux[id].y = (g.x + acc) / g.y;
                    // release the parents
                    const unsigned ci = cm & 3u, cj = (cm >> 2) & 3u, ck = (cm >> 4) & 3u;
#define TOPO_RELEASE(active, p, rp)                                              \
    if (active) {                                                                \
        const long long gq = base + (p);                                         \
        const unsigned sh = (unsigned)(gq & 3) * 8u;                             \
        const unsigned old = atomicSub(&cnt32[gq >> 2], 1u << sh);               \
        if (((old >> sh) & 0xFFu) == 1u) {                                       \
            if (AGG) rp = (int)(p);                                              \
            else q[atomicAdd(&s_tail, 1)] = (int)(p);                            \
        }                                                                        \
    }
                    TOPO_RELEASE(ci, ci == 1 ? id - nl : id + nl, rp0)
                    TOPO_RELEASE(cj, cj == 1 ? id - l : id + l, rp1)
                    TOPO_RELEASE(ck, ck == 1 ? id - 1 : id + 1, rp2)
#undef TOPO_RELEASE
                }
                const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
#define TOPO_PUSH(rp)                                                                      \
    {                                                                                      \
        const unsigned m = __ballot_sync(0xffffffffu, rp >= 0);                            \
        if (m) {                                                                           \
            const int leader = __ffs((int)m) - 1;                                          \
            int pos = 0;                                                                   \
            if ((int)(threadIdx.x & 31) == leader) pos = atomicAdd(&s_tail, __popc(m));    \
            pos = __shfl_sync(0xffffffffu, pos, leader);                                   \
            if (rp >= 0) q[pos + __popc(m & lt)] = rp;                                     \
        }                                                                                  \
    }
                if (AGG) {
                    TOPO_PUSH(rp0)
                    TOPO_PUSH(rp1)
                    TOPO_PUSH(rp2)
                }
#undef TOPO_PUSH
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

// grad_f[s][q] = -x * (-2 f h h)  (Eikonal3D.cpp:113-116,194-196); optional sum over sources (fixed order).
__global__ void k_adj3d_finish2(const double2 *__restrict__ UX, const double *__restrict__ f,
                                double *__restrict__ GF, double *__restrict__ GFsum, const long long N, const int S,
                                const double h) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += (long long)gridDim.x * blockDim.x) {
        const double rhs = -2 * f[q] * h * h;
        double acc = 0.0;
        for (int s = 0; s < S; s++) {
            const double v = -UX[(long long)s * N + q].y * rhs;
            if (GF) GF[(long long)s * N + q] = v;
            acc += v;
        }
        if (GFsum) GFsum[q] = acc;
    }
}

}  // namespace adtomo
