// kernels_adj_team.cuh -- the causality-ordered adjoint wavefront (kernels_adj_topo.cuh) for FEW sources:
// a TEAM of CTAs per source.
//
// Same algorithm, data layout and per-node arithmetic as k_adj3d_topo2 (reference: Eikonal3D.cpp:96-198,
// the transposed triangular system solved in decreasing-u order), so the result is the single-CTA kernel's,
// bit for bit: a node's value depends only on its children's values, gathered in a fixed order.  With one
// CTA per source a single large grid (BASELINE config C5) keeps one SM busy for ~100 ms at 256^3.  Here the
// ready queue of a wave is split over the CTAs of a team; newly ready parents are staged in shared memory
// (one shared-memory atomic per warp and axis) and appended to the queue with ONE global atomic per CTA and
// wave (one per warp and axis -- ~4000 on the same address per wave at 512^3 -- serialised in L2), and ONE
// team barrier (kernels_fwd_team.cuh: tm_barrier, whose fences also invalidate L1) separates the waves.  The counter of wave w is D[w % 3]: it
// counts the pushes of wave w relative to the wave's tail, is read by every CTA after barrier w, and is
// cleared by CTA 0 after barrier w+1 -- when every CTA has read it -- for its next use in wave w+3.
// grid = S x nC CTAs, all co-resident (cooperative launch).
#pragma once
#include "kernels_adj_topo.cuh"
#include "kernels_fwd_team.cuh"

namespace adtomo {

// tail0: queue tails after k_adj3d_count2; D: S x 4 push counters (3 used), zero on entry; bar: S barrier
// counters, zero on entry.
// CAP: staged pushes per CTA and wave; beyond that a push goes straight to the queue (CAP = 64 exists only so that the
// tests can exercise that path: ADTOMO_ADJ_CAP_SMALL=1).
template <int NT, int CAP = 6144>
__global__ void __launch_bounds__(NT) k_adj3d_topo_team(double2 *UX, const double2 *__restrict__ GD,
                                                        const unsigned short *__restrict__ CM, unsigned int *cnt32,
                                                        int *Q, const int *__restrict__ tail0, int *D,
                                                        const int *__restrict__ nfree, const Dims3 d, const int nC,
                                                        unsigned *bar, int *__restrict__ status) {
    __shared__ int s_stage[CAP];
    __shared__ int s_n, s_base;
    const int l = d.l;
    const int nl = d.n * d.l;
    const int src = blockIdx.x / nC, t = blockIdx.x - src * nC;
    const long long base = (long long)src * d.N;
    double2 *ux = UX + base;
    const double2 *gd = GD + base;
    const unsigned short *cm_ = CM + base;
    int *q = Q + base;
    int *Dsrc = D + 4 * src;
    unsigned epoch = 0;
    int head = 0, tail = tail0[src], waves = 0;
    const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
    while (head < tail) {
        int *Dw = Dsrc + waves % 3;
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        for (int t0 = head + t * NT + (threadIdx.x & ~31); t0 < tail; t0 += nC * NT) {
            const int tq = t0 + (threadIdx.x & 31);
            int rp0 = -1, rp1 = -1, rp2 = -1;      // parents that became ready through this node
            if (tq < tail) {
                const int id = __ldcg(q + tq);
                const unsigned cm = cm_[id];
                const double ui = ux[id].x;
                const double2 g = gd[id];
                double acc = 0.0;
                // children in the fixed order i-1, i+1, j-1, j+1, k-1, k+1 (all final by construction)
#define TOPO_CHILD(bit, off)                                        \
    if (cm & ((bit) << 8)) {                                        \
        const double2 c = __ldcg(ux + id + (off));                  \
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
                const unsigned ci = cm & 3u, cj = (cm >> 2) & 3u, ck = (cm >> 4) & 3u;
#define TOPO_RELEASE(active, p, rp)                                              \
    if (active) {                                                                \
        const long long gq = base + (p);                                         \
        const unsigned sh = (unsigned)(gq & 3) * 8u;                             \
        const unsigned old = atomicSub(&cnt32[gq >> 2], 1u << sh);               \
        if (((old >> sh) & 0xFFu) == 1u) rp = (int)(p);                          \
    }
                TOPO_RELEASE(ci, ci == 1 ? id - nl : id + nl, rp0)
                TOPO_RELEASE(cj, cj == 1 ? id - l : id + l, rp1)
                TOPO_RELEASE(ck, ck == 1 ? id - 1 : id + 1, rp2)
#undef TOPO_RELEASE
            }
#define TOPO_PUSH(rp)                                                                      \
    {                                                                                      \
        const unsigned m = __ballot_sync(0xffffffffu, rp >= 0);                            \
        if (m) {                                                                           \
            const int leader = __ffs((int)m) - 1;                                          \
            int pos = 0;                                                                   \
            if ((int)(threadIdx.x & 31) == leader) pos = atomicAdd(&s_n, __popc(m));       \
            pos = __shfl_sync(0xffffffffu, pos, leader);                                   \
            if (rp >= 0) {                                                                 \
                const int my = pos + __popc(m & lt);                                       \
                if (my < CAP) s_stage[my] = rp;                                            \
                else q[tail + atomicAdd(Dw, 1)] = rp;          /* staging full (rare) */   \
            }                                                                              \
        }                                                                                  \
    }
            TOPO_PUSH(rp0)
            TOPO_PUSH(rp1)
            TOPO_PUSH(rp2)
#undef TOPO_PUSH
        }
        __syncthreads();
        const int staged = s_n < CAP ? s_n : CAP;
        if (threadIdx.x == 0 && staged > 0) s_base = atomicAdd(Dw, staged);
        __syncthreads();
        for (int i = threadIdx.x; i < staged; i += NT) q[tail + s_base + i] = s_stage[i];
        tm_barrier(bar + src, epoch, nC);      // x values, counters and queue entries of this wave are visible
        const int pushed = __ldcg(Dw);
        if (t == 0 && threadIdx.x == 0) Dsrc[(waves + 2) % 3] = 0;   // wave w-1's counter: everyone read it before this barrier
        head = tail;
        tail += pushed;
        waves++;
    }
    if (t == 0 && threadIdx.x == 0 && status) status[src] = (tail == nfree[src]) ? waves + 1 : -(waves + 1);
}

}  // namespace adtomo
