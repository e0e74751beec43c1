// kernels_adj_sparse.cuh -- the 3D adjoint restricted to the nodes that can carry a non-zero adjoint value.
//
// Reference semantics: Eikonal3D.cpp:96-198 (assemble A, solve A^T x = grad_u, grad_f = 2 f h^2 x); the wavefront
// formulation of kernels_adj_topo.cuh (x_p = (g_p + sum_children 2 (u_c - u_p) x_c) / D_p, children first).
//
// Why.  In an inversion step the right-hand side is SPARSE: d(misfit)/du lives on the 8 corners of the receiver
// cells (scripts/inversion.jl:64-105; 512 receivers -> ~3.9 k of 1 M nodes on the bench batch).  x is non-zero only
// on the ANCESTORS of those nodes (the nodes reached by following upwind-parent links, i.e. the ray tubes back to
// the source): 37 % of the grid on the bench model.  The dense wavefront (k_adj3d_topo2) evaluates every node, and it
// is bound by scattered 32-byte sectors per node (ncu: profiles/r02_ncu_summary_adj.json), so work per node is what
// counts.  Here:
//   setup  (coalesced, k_adj3d_setup3): parent code, diagonal, {u,x=0}, {g,D}; one 32-bit word W per node; the nodes
//          with g != 0 ("seeds") start the mark queue;
//   mark   (wavefront over parent links, phase A of k_adj3d_sparse): every node reached for the first time is queued;
//          the child leaves its bit in the parent's child mask and +1 in the parent's pending count -- ONE atomicAdd on
//          W[parent]; its return value tells whether the parent was reached before;
//   solve  (Kahn wavefront, phase B): exactly k_adj3d_topo2's node update on the marked nodes only; children outside
//          the marked set have x = 0 exactly, so leaving them out changes no bit of x.
// Every marked node is visited twice (mark, solve), every other node only by the coalesced setup / finish passes.
//
//   W: bits 0-5 children mask (1: i-1, 2: i+1, 4: j-1, 8: j+1, 16: k-1, 32: k+1 is a child), bit 7 seed,
//      bits 8-15 children still pending, bits 16-21 parent code (2 bits per axis: 0 none, 1: -1 side, 2: +1 side),
//      bit 31 pinned (u == u0 or no active axis: x = 0, Eikonal3D.cpp:126-130,168-171)
#pragma once
#include "kernels_adj_topo.cuh"

namespace adtomo {

#define AW_MASK 0x3Fu
#define AW_SEED 0x80u
#define AW_CNT1 0x100u
#define AW_CNT 0xFF00u
#define AW_PIN 0x80000000u
#define AW_TOUCHED 0x8000FFFFu   // pinned, seed, or reached by a child before

// grid = (blocks, S).  qtail[src]: number of seeds written to Q1 of that source.
__global__ void k_adj3d_setup3(const double *__restrict__ U, const double *__restrict__ U0,
                               const double *__restrict__ G, double2 *__restrict__ UX, double2 *__restrict__ GD,
                               double *__restrict__ GU0, unsigned *__restrict__ W, int *__restrict__ Q1,
                               int *__restrict__ qtail, const Dims3 d, const int S) {
    const int src = blockIdx.y;
    const long long base = (long long)src * d.N;
    const double *u = U + base;
    const int N = (int)d.N, n = d.n, l = d.l, nl = d.n * d.l;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < N; id += gridDim.x * blockDim.x) {
        const double ui = u[id];
        const double gi = G[base + id];
        const bool same = (ui == U0[base + id]);
        if (GU0) GU0[base + id] = same ? gi : 0.0;       // Eikonal3D.cpp:106-110
        unsigned w = AW_PIN;
        double D = 0.0;
        if (!same) {
            const int i = id / nl;
            const int r = id - i * nl;
            const int j = r / l;
            const int k = r - j * l;
            const unsigned ci = adj_axis_code(u, id, i, d.m, nl, ui), cj = adj_axis_code(u, id, j, n, l, ui),
                           ck = adj_axis_code(u, id, k, l, 1, ui);
            const unsigned cd = ci | (cj << 2) | (ck << 4);
            if (cd != 0) {
                if (ci) D += 2.0 * (ui - u[ci == 1 ? id - nl : id + nl]);
                if (cj) D += 2.0 * (ui - u[cj == 1 ? id - l : id + l]);
                if (ck) D += 2.0 * (ui - u[ck == 1 ? id - 1 : id + 1]);
                w = cd << 16;
                if (gi != 0.0) {
                    w |= AW_SEED;
                    Q1[base + atomicAdd(&qtail[src], 1)] = id;
                }
            }
        }
        W[base + id] = w;
        UX[base + id] = make_double2(ui, 0.0);
        GD[base + id] = make_double2(gi, D);
    }
}

// One CTA per source.  Q1: mark queue (seeds on entry), Q2: ready queue of the solve.
template <int NT>
__global__ void __launch_bounds__(NT) k_adj3d_sparse(double2 *UX, const double2 *__restrict__ GD, unsigned *W, int *Q1, int *Q2,
                                                     const int *__restrict__ qtail, const Dims3 d, const int S,
                                                     int *__restrict__ status) {
    __shared__ int s_tail;
    const int l = d.l;
    const int nl = d.n * d.l;
    const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
    // the lanes of a warp append their entries with ONE shared-memory atomic (warp-uniform call)
#define ASP_PUSH(queue, rp)                                                                \
    {                                                                                      \
        const unsigned m__ = __ballot_sync(0xffffffffu, (rp) >= 0);                        \
        if (m__) {                                                                         \
            const int leader = __ffs((int)m__) - 1;                                        \
            int pos = 0;                                                                   \
            if ((int)(threadIdx.x & 31) == leader) pos = atomicAdd(&s_tail, __popc(m__));  \
            pos = __shfl_sync(0xffffffffu, pos, leader);                                   \
            if ((rp) >= 0) (queue)[pos + __popc(m__ & lt)] = (rp);                         \
        }                                                                                  \
    }
    for (int src = blockIdx.x; src < S; src += gridDim.x) {
        const long long base = (long long)src * d.N;
        double2 *ux = UX + base;
        const double2 *gd = GD + base;
        unsigned *w_ = W + base;
        int *q1 = Q1 + base, *q2 = Q2 + base;
        const int nseeds = qtail[src];
        // ---- phase A: mark the ancestors of the seeds, leave child masks and pending counts behind ----
        int head = 0, tail = nseeds;
        if (threadIdx.x == 0) s_tail = tail;
        __syncthreads();
        while (head < tail) {
            for (int t0 = head + (threadIdx.x & ~31); t0 < tail; t0 += NT) {
                const int t = t0 + (threadIdx.x & 31);
                int np0 = -1, np1 = -1, np2 = -1;      // parents reached for the first time through this node
                if (t < tail) {
                    const int id = q1[t];
                    const unsigned w = __ldcg(&w_[id]);            // the parent code bits are static
                    const unsigned ci = (w >> 16) & 3u, cj = (w >> 18) & 3u, ck = (w >> 20) & 3u;
#define ASP_MARK(active, p, childbit, np)                                                  \
    if (active) {                                                                          \
        const unsigned old = atomicAdd(&w_[p], (childbit) | AW_CNT1);                      \
        if ((old & AW_TOUCHED) == 0u) np = (int)(p);                                       \
    }
                    ASP_MARK(ci, ci == 1 ? id - nl : id + nl, ci == 1 ? 2u : 1u, np0)
                    ASP_MARK(cj, cj == 1 ? id - l : id + l, cj == 1 ? 8u : 4u, np1)
                    ASP_MARK(ck, ck == 1 ? id - 1 : id + 1, ck == 1 ? 32u : 16u, np2)
#undef ASP_MARK
                }
                ASP_PUSH(q1, np0)
                ASP_PUSH(q1, np1)
                ASP_PUSH(q1, np2)
            }
            __syncthreads();
            const int nt = s_tail;
            __syncthreads();
            head = tail;
            tail = nt;
        }
        const int nmarked = tail;
        // ---- the solve starts from the marked nodes without marked children: they are all seeds ----
        if (threadIdx.x == 0) s_tail = 0;
        __syncthreads();
        for (int t0 = threadIdx.x & ~31; t0 < nseeds; t0 += NT) {
            const int t = t0 + (threadIdx.x & 31);
            int rp = -1;
            if (t < nseeds) {
                const int id = q1[t];
                if ((__ldcg(&w_[id]) & AW_CNT) == 0u) rp = id;
            }
            ASP_PUSH(q2, rp)
        }
        __syncthreads();
        head = 0;
        tail = s_tail;
        __syncthreads();
        // ---- phase B: Kahn wavefront over the marked nodes (node update of k_adj3d_topo2) ----
        int waves = 0;
        while (head < tail) {
            for (int t0 = head + (threadIdx.x & ~31); t0 < tail; t0 += NT) {
                const int t = t0 + (threadIdx.x & 31);
                int rp0 = -1, rp1 = -1, rp2 = -1;      // parents that became ready through this node
                if (t < tail) {
                    const int id = q2[t];
                    const unsigned w = __ldcg(&w_[id]);            // mask complete: phase A is over
                    const double ui = ux[id].x;
                    const double2 g = gd[id];
                    double acc = 0.0;
                    // children in the fixed order i-1, i+1, j-1, j+1, k-1, k+1 (all final by construction)
#define ASP_CHILD(bit, off)                                         \
    if (w & (bit)) {                                                \
        const double2 c = ux[id + (off)];                           \
        acc += 2.0 * (c.x - ui) * c.y;                              \
    }
                    ASP_CHILD(1u, -nl)
                    ASP_CHILD(2u, nl)
                    ASP_CHILD(4u, -l)
                    ASP_CHILD(8u, l)
                    ASP_CHILD(16u, -1)
                    ASP_CHILD(32u, 1)
#undef ASP_CHILD
                    ux[id].y = (g.x + acc) / g.y;
                    const unsigned ci = (w >> 16) & 3u, cj = (w >> 18) & 3u, ck = (w >> 20) & 3u;
#define ASP_RELEASE(active, p, rp)                                                         \
    if (active) {                                                                          \
        const unsigned old = atomicSub(&w_[p], AW_CNT1);                                   \
        if ((old & (AW_PIN | AW_CNT)) == AW_CNT1) rp = (int)(p);                           \
    }
                    ASP_RELEASE(ci, ci == 1 ? id - nl : id + nl, rp0)
                    ASP_RELEASE(cj, cj == 1 ? id - l : id + l, rp1)
                    ASP_RELEASE(ck, ck == 1 ? id - 1 : id + 1, rp2)
#undef ASP_RELEASE
                }
                ASP_PUSH(q2, rp0)
                ASP_PUSH(q2, rp1)
                ASP_PUSH(q2, rp2)
            }
            __syncthreads();
            const int nt = s_tail;
            __syncthreads();
            head = tail;
            tail = nt;
            waves++;
        }
        if (threadIdx.x == 0 && status) status[src] = (tail == nmarked) ? waves + 1 : -(waves + 1);
        __syncthreads();
    }
#undef ASP_PUSH
}

}  // namespace adtomo
