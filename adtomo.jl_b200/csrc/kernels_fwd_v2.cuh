// kernels_fwd_v2.cuh -- 3D forward fast sweeping on SKEWED-PENCIL layouts.
//
// Reference semantics: Eikonal3D.cpp:35-57 (one directional Gauss-Seidel sweep), :59-68 (the 8
// sweeps of a round), :71-88 (rounds until max|u - u_old| < tol, cap 20).  A sweep is executed level
// by level (level = A'+W'+C' with every coordinate counted in the sweep's direction): the nodes of a
// level are independent, take NEW values from level-1 and OLD values from level+1, which reproduces
// the serial lexicographic sweep bit for bit.
//
// The level-major kernel (kernels_fwd_v1.cuh) spends ~185 of its ~290 instructions per 32 node
// updates on index arithmetic (packed level enumeration, boundary predicates, next-layout offsets).
// Here the three grid axes get fixed ROLES (A = row axis, W = walk axis, C = lane axis) and a field
// is stored skewed:   slot(A, mu, C),  mu = W + C  (layout P)   or   mu = W + (dC-1-C)  (layout M),
// offset ((A+1)*RS + mu+1)*PC + C.  A thread owns PENCILS (A, C) and walks them along mu, one node
// per level:
//   * all six neighbours sit at CONSTANT offsets from the node (global: A+-1 -> +-RS*PC,
//     W+-1 -> +-PC, C+-1 -> +-(PC+-1)), the offset of a
//     pencil advances by a constant per level, and the sweep direction only changes the SIGNS of
//     those constants: one code path, no per-node index decode;
//   * every slot that is not a grid node (skew padding, one pad row/column/slab around the box) holds
//     +inf forever, so the reference's mirror rule (Eikonal3D.cpp:47-52: a missing neighbour is
//     replaced by the existing one) is min(x, +inf) and needs NO boundary predicate;
//   * sweeps whose (W,C) signs agree run on layout P, the others on M; consecutive sweeps on the same
//     layout update the field IN PLACE.  The reference order needs 4 layout changes per round
//     (P P M M P M M P); each is a re-skew of every A-slab through a shared-memory plane, coalesced
//     and with full lanes on both sides.  (Measured alternative: letting the sweep before a change
//     scatter its results into the other layout -- the 4 doubles of a sector arrive 2 levels apart,
//     and with 256 sources in flight the partially written sectors overflow L2: 3.5x the DRAM bytes.)
//   * lanes of a warp cover a patch of 4 rows x 8 columns of pencils (a warp slot), so a slot is live
//     for dW + 10 levels of which dW are fully used (128 -> 93 %); the live slots of a level are dealt
//     round-robin to the warps.
// Upwind (new) values are read back from global memory: they were stored by threads of the same CTA
// one level (one __syncthreads) earlier, so they are visible, and they sit at the mirrored constant
// offsets.  The sweep therefore needs NO shared memory (the level-major kernel needs two sheets of
// (dA+2) x pitch doubles, 137 KB for 128x128x64, which pins it to one CTA per SM and to grids whose
// sheets fit): two CTAs (= two sources) share an SM and fill each other's barrier and load stalls,
// and any grid size runs on a single CTA per source.
// Buffers per source: three fields o/a (layout P) and z (layout M).  Sweep 1 of a round reads o and
// writes a (o stays as the round-start field for the L-inf stopping test, fused into sweep 8).
#pragma once
#include <cstdlib>
#include "eik_core.h"

namespace adtomo {

// Lane patch of a warp slot: LA rows x LC columns of pencils (LA * LC = 32).  Build-time knobs for A/B runs
// (-DADTOMO_V2_LA=2 -DADTOMO_V2_LC=16): wider patches halve the L1 wavefronts per load (rows of 128 instead of 64 bytes)
// and keep a slot live for LA + LC - 2 more levels than the pencils are long.
#ifndef ADTOMO_V2_LA
#define ADTOMO_V2_LA 4
#endif
#ifndef ADTOMO_V2_LC
#define ADTOMO_V2_LC 8
#endif
constexpr int V2_LA = ADTOMO_V2_LA;   // rows of the lane patch
constexpr int V2_LC = ADTOMO_V2_LC;   // columns of the lane patch
static_assert(V2_LA * V2_LC == 32, "a warp slot is one node per lane");

struct Plan2 {
    int ext[3];          // m, n, l
    int role[3];         // grid axis (0=i,1=j,2=k) playing A, W, C
    int dA, dW, dC;
    int nmu;             // dW + dC - 1 skewed rows per slab
    int RS;              // rows per slab incl. one pad row each side = nmu + 2
    int PC;              // row pitch (>= dC+1, multiple of 4)
    int nlev;            // dA + dW + dC - 2
    int G;               // column groups (of LC columns) per row
    int NT;              // threads per CTA
    int WCH, PS;         // re-skew: W rows per pass, plane pitch (even)
    long long M;         // slots per field buffer = (dA+2)*RS*PC
    long long N;
    int sg[8][3];        // (sA, sW, sC) of the reference's 8 sweeps (Eikonal3D.cpp:59-68) by role
};

#define V2_INF_BITS 0x7ff0000000000000LL

EIK_HD double v2_inf() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(V2_INF_BITS);
#else
    return INFINITY;
#endif
}

// slot of node (A, W, C) in layout sigma (+1: P, -1: M)
EIK_HD long long v2_offset(const Plan2 &P, int A, int W, int C, int sigma) {
    const int mu = W + (sigma > 0 ? C : P.dC - 1 - C);
    return ((long long)(A + 1) * P.RS + (mu + 1)) * P.PC + C;
}

// slot of grid node (i,j,k)
EIK_HD long long v2_offset_ijk(const Plan2 &P, int i, int j, int k, int sigma) {
    const int x[3] = {i, j, k};
    return v2_offset(P, x[P.role[0]], x[P.role[1]], x[P.role[2]], sigma);
}

// Live row blocks of column group g at level lam: rb in [lo, lo+n).  A row block is LA rows of
// pencils, a column group LC columns; the block is live when some pencil (A', C') in it has
// 0 <= lam - A' - C' < dW (coordinates counted in the sweep's direction).
template <int SC>
EIK_HD void v2_window(const Plan2 &P, const int g, const int lam, int &lo, int &n) {
    const int c_lo = g * V2_LC, c_hi = (g * V2_LC + V2_LC - 1 < P.dC - 1) ? g * V2_LC + V2_LC - 1 : P.dC - 1;
    const int cpmin = SC > 0 ? c_lo : P.dC - 1 - c_hi;       // C' range of the group (columns beyond dC excluded)
    const int cpmax = SC > 0 ? c_hi : P.dC - 1 - c_lo;
    const int nrb = (P.dA + V2_LA - 1) / V2_LA;
    const int t = lam - cpmin;
    int hi = t >= 0 ? t / V2_LA : -1;
    if (hi > nrb - 1) hi = nrb - 1;
    const int lo_n = lam - (P.dW - 1 + cpmax + V2_LA - 1);
    lo = lo_n > 0 ? (lo_n + V2_LA - 1) / V2_LA : 0;
    n = hi - lo + 1;
    if (n < 0 || g >= P.G) n = 0;
}

// Per-lane constants of one sweep: lane = (la, lc) inside the LA x LC patch of pencils.
struct V2Lane {
    int offc;   // slot = offc + rb*offRB + lam*offW + g*LC
    int wqc;    // W'   = wqc + lam - rb*LA - SC*g*LC
    int la, lc;
};

template <int SA, int SW, int SC>
EIK_HD V2Lane v2_lane_setup(const Plan2 &P, const int lane) {
    V2Lane L;
    L.la = lane / V2_LC;
    L.lc = lane % V2_LC;
    L.offc = ((SA > 0 ? L.la + 1 : P.dA - L.la) * P.RS + (SW > 0 ? 1 - L.la : P.nmu + L.la)) * P.PC + L.lc;
    L.wqc = -L.la - (SC > 0 ? L.lc : P.dC - 1 - L.lc);
    return L;
}

// The eight values one node update reads.  off < 0: the lane has no node in this warp slot.
struct V2Vals {
    double own, fv, dA, dW, dC, uA, uW, uC;
    double ref;   // CMP sweeps: the round-start value of the node
    int off;
};

// Loads of one lane's node of warp slot (rb, g) at level lam of a sweep with signs (SA, SW, SC) by
// role.  OOP: rd != wr (sweep 1 of a round: old values in rd, new ones in wr), else in place.
// Separate from the arithmetic so that the kernel can issue the NEXT slot's loads before it computes
// the current slot (all of them are level-1 / level+1 / own values: nothing this level writes).
template <int SA, int SW, int SC, bool OOP, bool CMP>
EIK_HD void v2_load(const Plan2 &P, const V2Lane &L, const int lam, const int rb, const int g, const double *rd,
                    const double *wr, const double *__restrict__ fl, const double *cmp, V2Vals &V) {
    const int offA = SA * P.RS * P.PC, offW = SW * P.PC, offC = SW * P.PC + SC;   // downwind (old, level+1)
    const int offRB = V2_LA * (SA * P.RS - SW) * P.PC;
    const int wq = L.wqc + lam - rb * V2_LA - SC * g * V2_LC;
    const bool act = (unsigned)wq < (unsigned)P.dW && rb * V2_LA + L.la < P.dA && g * V2_LC + L.lc < P.dC;
    // a lane without a node loads from a harmless slot (A = 0, mu = 0: all six neighbour slots exist) instead of
    // branching around the loads; it neither stores nor contributes to err
    const int off = act ? L.offc + rb * offRB + lam * offW + g * V2_LC : (P.RS + 1) * P.PC + 1;
    V.off = act ? off : -1;
    const double *p = rd + off;
    V.own = p[0];
    V.fv = fl[off];
    V.dA = p[offA];
    V.dW = p[offW];
    V.dC = p[offC];
    // upwind neighbours (new values of level-1, stored one barrier ago by this CTA); OOP: they live in wr
    const double *pu = OOP ? wr + off : p;
    V.uA = pu[-offA];
    V.uW = pu[-offW];
    V.uC = pu[-offC];
    V.ref = CMP ? cmp[off] : 0.0;
}

// The update itself (Eikonal3D.cpp:47-54), in two steps so that the kernel can issue the next slot's
// loads in between: v2_prep consumes the eight loaded values (per-axis minima, sorted), v2_solve does
// the arithmetic and the store.  CMP: fold |new - cmp| into err (sweep 8).
struct V2Prep {
    double a1, a2, a3, own, fv, ref;
    int off;
};

EIK_HD void v2_prep(const V2Vals &V, V2Prep &Q) {
    Q.off = V.off;
    Q.own = V.own;
    Q.fv = V.fv;
    Q.ref = V.ref;
    Q.a1 = eik_min(V.uA, V.dA);
    Q.a2 = eik_min(V.uW, V.dW);
    Q.a3 = eik_min(V.uC, V.dC);
    eik_sort3(Q.a1, Q.a2, Q.a3);
}

template <bool OOP, bool CMP>
EIK_HD void v2_solve(const V2Prep &Q, double *wr, const double h, double &err) {
    if (Q.off < 0) return;
    double res = Q.own;
    bool changed = false;
    if (Q.a1 < Q.own) {   // otherwise the candidate (> a1) cannot win the min: exact skip
        const double un = eik_solve3_sorted(Q.a1, Q.a2, Q.a3, Q.fv * h, Q.fv * Q.fv * h * h);
        if (un < Q.own) { res = un; changed = true; }
    }
    if (OOP || changed) wr[Q.off] = res;
    if (CMP) {
        const double dd = fabs(res - Q.ref);
        err = (err < dd) ? dd : err;
    }
}

template <bool OOP, bool CMP>
EIK_HD void v2_finish(const V2Vals &V, double *wr, const double h, double &err) {
    V2Prep Q;
    v2_prep(V, Q);
    v2_solve<OOP, CMP>(Q, wr, h, err);
}

template <int SA, int SW, int SC, bool OOP, bool CMP>
EIK_HD void v2_node(const Plan2 &P, const V2Lane &L, const int lam, const int rb, const int g, const double *rd,
                    double *wr, const double *__restrict__ fl, const double *cmp, const double h, double &err) {
    V2Vals V;
    v2_load<SA, SW, SC, OOP, CMP>(P, L, lam, rb, g, rd, wr, fl, cmp, V);
    v2_finish<OOP, CMP>(V, wr, h, err);
}

// dispatch on the signs of sweep sw (sweeps 0 and 7 are (+,+,+) and (-,-,-) under every role assignment)
#define V2_DISPATCH(P, sw, CALL)                                                           \
    do {                                                                                   \
        if ((sw) == 0) { CALL(1, 1, 1, true, false); }                                     \
        else if ((sw) == 7) { CALL(-1, -1, -1, false, true); }                             \
        else {                                                                             \
            const int code__ = ((P).sg[sw][0] > 0 ? 4 : 0) | ((P).sg[sw][1] > 0 ? 2 : 0) | ((P).sg[sw][2] > 0 ? 1 : 0); \
            switch (code__) {                                                              \
                case 1: CALL(-1, -1, 1, false, false); break;                              \
                case 2: CALL(-1, 1, -1, false, false); break;                              \
                case 3: CALL(-1, 1, 1, false, false); break;                               \
                case 4: CALL(1, -1, -1, false, false); break;                              \
                case 5: CALL(1, -1, 1, false, false); break;                               \
                default: CALL(1, 1, -1, false, false); break;                              \
            }                                                                              \
        }                                                                                  \
    } while (0)

// Re-skew of W-chunk [w0, w0+wc) of slab A between the layouts, through `plane` (wc x PS doubles).
// The chunk is enumerated by VIRTUAL rows v in [0, wc) x columns C: element (v, C) is the node
// W = w0 + ((v - cc) mod wc), cc = C (layout P) or dC-1-C (layout M), which lies in row mu = W + cc of
// that layout.  A virtual row is made of pieces of the physical rows w0+v, w0+v+wc, ...: every lane
// has a node, and each piece is a contiguous run of C, i.e. coalesced.
// phase 0: rows of layout sigmaFrom -> plane;  phase 1: plane -> rows of the other layout.
// Index map of element (v, C): plane slot and offset inside the slab (layout sigma).
EIK_HD void v2_reskew_index(const Plan2 &P, const int sigma, const int w0, const int wc, const int v, const int C,
                            int &pl, int &go) {
    const int cc = sigma > 0 ? C : P.dC - 1 - C;
    int Wl = v - cc;                       // local W in [0, wc)
    while (Wl < 0) Wl += wc;
    pl = Wl * P.PS + C;
    go = (w0 + Wl + cc + 1) * P.PC + C;
}

EIK_HD void v2_reskew_elem(const Plan2 &P, const double *src, double *dst, const int sigmaFrom, double *plane,
                           const long long slab, const int w0, const int wc, const int phase, const int v, const int C) {
    int pl, go;
    v2_reskew_index(P, phase == 0 ? sigmaFrom : -sigmaFrom, w0, wc, v, C, pl, go);
    if (phase == 0) plane[pl] = src[slab + go];
    else dst[slab + go] = plane[pl];
}

#if defined(__CUDACC__)

__device__ __forceinline__ double v2_block_max(double v, double *red) {
    for (int o = 16; o > 0; o >>= 1) {
        double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = (w > v) ? w : v;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = red[0];
    const int nw = blockDim.x >> 5;
    for (int w = 1; w < nw; w++) r = (red[w] > r) ? red[w] : r;
    return r;
}

// One sweep.  Per level the live warp slots (row block x column group) are enumerated group by group
// and dealt round-robin to the warps: lane g of every warp computes group g's window, a warp scan gives
// the slot ranges, and slot q is located with one ballot (fixed pencil ownership left the slowest warp
// with 1.5x the mean work per level).
template <int SA, int SW, int SC, bool OOP, bool CMP>
__device__ __forceinline__ void v2_sweep(const Plan2 &P, const double *rd, double *wr,
                                         const double *__restrict__ fl, const double *cmp, const double h,
                                         double &err) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const V2Lane L = v2_lane_setup<SA, SW, SC>(P, lane);
    __syncthreads();     // the previous sweep wrote the field this sweep reads
    for (int lam = 0; lam < P.nlev; lam++) {
        int lo, n;
        v2_window<SC>(P, lane, lam, lo, n);
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            if (d >= P.G) break;                   // groups live in lanes 0..G-1 (uniform exit)
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, P.G - 1);
        const int excl = incl - n;
        const int rb0 = lo - excl;                 // slot q of group g is row block rb0[g] + q
        // slot q -> (g, rb).  The loads of the warp's next slot are issued before the current slot's
        // arithmetic (software pipelining by hand: the values come from L2 / DRAM and the ~120 instructions
        // of the solve hide their latency).
#define V2_MAP(q_, g_, rb_)                                                                           \
    do {                                                                                              \
        const unsigned m__ = __ballot_sync(0xffffffffu, excl <= (q_) && n > 0);                       \
        g_ = 31 - __clz((int)m__);                                                                    \
        rb_ = __shfl_sync(0xffffffffu, rb0, g_) + (q_);                                               \
    } while (0)
#define V2_LOAD(q_, V_)                                                    \
    do {                                                                   \
        int g__, rb__;                                                     \
        V2_MAP(q_, g__, rb__);                                             \
        v2_load<SA, SW, SC, OOP, CMP>(P, L, lam, rb__, g__, rd, wr, fl, cmp, V_); \
    } while (0)
        int q = warp;
        if (q < total) {
            V2Vals V;
            V2_LOAD(q, V);
#pragma unroll 1
            for (;;) {
                V2Prep Q;
                v2_prep(V, Q);                     // consumes V: its registers take the next slot's loads
                q += nw;
                if (q < total) V2_LOAD(q, V);
                v2_solve<OOP, CMP>(Q, wr, h, err);
                if (q >= total) break;
            }
        }
#undef V2_LOAD
#undef V2_MAP
        __syncthreads();
    }
}

// Re-skew of a whole field: slab by slab, chunk by chunk; two barriers per chunk.  Lane = column C,
// a warp takes virtual rows v = warp, warp+nw, ...: plane slot and slab offset advance by constants
// (with a wrap), four elements are in flight per thread.  Same map as v2_reskew_index.
template <int PHASE>
__device__ __forceinline__ void v2_reskew_pass(const Plan2 &P, const double *s, double *d, const int sigma,
                                               double *plane, const int w0, const int wc) {
    constexpr int U = 8;                  // elements in flight per thread (the loads come from DRAM)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int dpl = nw * P.PS, dgo = nw * P.PC, wpl = wc * P.PS, wgo = wc * P.PC;
    for (int C = lane; C < P.dC; C += 32) {
        const int cc = sigma > 0 ? C : P.dC - 1 - C;
        int Wl = warp - cc;
        while (Wl < 0) Wl += wc;
        int pl = Wl * P.PS + C, go = (w0 + Wl + cc + 1) * P.PC + C;
        for (int v = warp; v < wc; v += U * nw) {
            int pls[U], gos[U];
            double x[U];
#pragma unroll
            for (int j = 0; j < U; j++) {
                pls[j] = pl; gos[j] = go;
                Wl += nw; pl += dpl; go += dgo;
                while (Wl >= wc) { Wl -= wc; pl -= wpl; go -= wgo; }
            }
#pragma unroll
            for (int j = 0; j < U; j++) {
                x[j] = 0.0;
                if (v + j * nw < wc) x[j] = PHASE == 0 ? s[gos[j]] : plane[pls[j]];
            }
#pragma unroll
            for (int j = 0; j < U; j++)
                if (v + j * nw < wc) {
                    if (PHASE == 0) plane[pls[j]] = x[j];
                    else d[gos[j]] = x[j];
                }
        }
    }
}

// slabs A in [A0, A1): the team kernel (kernels_fwd_team.cuh) gives every CTA of a team its own slabs
__device__ __forceinline__ void v2_reskew(const Plan2 &P, const double *src, double *dst, const int sigmaFrom,
                                          double *plane, const int A0, const int A1) {
    for (int A = A0; A < A1; A++) {
        const double *s = src + (long long)(A + 1) * P.RS * P.PC;
        double *d = dst + (long long)(A + 1) * P.RS * P.PC;
        for (int w0 = 0; w0 < P.dW; w0 += P.WCH) {
            const int wc = (P.dW - w0 < P.WCH) ? P.dW - w0 : P.WCH;
            v2_reskew_pass<0>(P, s, d, sigmaFrom, plane, w0, wc);
            __syncthreads();
            v2_reskew_pass<1>(P, s, d, -sigmaFrom, plane, w0, wc);
            __syncthreads();
        }
    }
}

// bufs: S x 3 x M doubles; buffer 0 of every source holds u0 in layout P on entry, every slot that
// is not a grid node holds +inf in all three buffers.  fP/fM: the slowness in layouts P and M.
// where[slot] receives the index (0..2) of the buffer holding the result (layout P).
// Buffer slot b holds source order[b] (see k2_make_order); rounds / errs are indexed by source.
// One CTA per source at a time; sources are assigned statically (block-uniform control flow).
// Dynamic shared memory: the re-skew plane, WCH x PS doubles.
template <int NTMAX, int MINB>
__global__ void __launch_bounds__(NTMAX, MINB) k_fwd3d_v2(const Plan2 P, double *bufs,
                                                          const double *__restrict__ fP, const double *__restrict__ fM,
                                                          const double h, const double tol, const int max_rounds,
                                                          const int S, int *__restrict__ rounds, double *__restrict__ errs,
                                                          int *__restrict__ where, const int *__restrict__ order,
                                                          int *__restrict__ spent) {
    extern __shared__ double plane[];
    __shared__ double red[32];
    for (int src = blockIdx.x; src < S; src += gridDim.x) {
        double *B3 = bufs + (long long)src * 3 * P.M;
        int o = 0, a = 1;          // layout P: round-start field, working field;  buffer 2: layout M
        double *Bz = B3 + 2 * P.M;
        int r = 0;
        bool conv = false;
        while (r < max_rounds) {
            double err = 0.0;
            double *Bo = B3 + o * P.M, *Ba = B3 + a * P.M;
            int state = 1;                        // layout of the working field
            double *w = Ba;
            for (int sw = 0; sw < 8; sw++) {
                const int sigma = P.sg[sw][1] * P.sg[sw][2];
                if (sw > 0 && sigma != state) {
                    double *dst = state > 0 ? Bz : Ba;
                    __syncthreads();
                    v2_reskew(P, w, dst, state, plane, 0, P.dA);
                    w = dst;
                    state = sigma;
                }
#define V2_CALL(a_, w_, c_, oop_, cmp_) v2_sweep<a_, w_, c_, oop_, cmp_>(P, oop_ ? Bo : w, w, sigma > 0 ? fP : fM, Bo, h, err)
                V2_DISPATCH(P, sw, V2_CALL);
#undef V2_CALL
            }
            const double e = v2_block_max(err, red);
            if (threadIdx.x == 0 && errs) errs[(long long)order[src] * max_rounds + r] = e;
            r++;
            const int oo = o; o = a; a = oo;      // the result (in a) becomes next round's round-start field
            if (__any_sync(0xffffffffu, e < tol)) { conv = true; break; }   // e is block-uniform; the vote makes that visible
        }
        if (threadIdx.x == 0) {
            if (rounds) rounds[order[src]] = conv ? r : -r;
            spent[order[src]] = r;      // remembered: the next call with this batch pairs long sources with short ones
            {
                unsigned sm__;
                asm("mov.u32 %0, %%smid;" : "=r"(sm__));
                spent[S + src] = (int)sm__;   // ... using the SM this CTA ran on
            }
            where[src] = o;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// layout conversion (elementwise, once per call)
// ---------------------------------------------------------------------------------------------
__global__ void k2_fill(double *__restrict__ p, const long long n, const double v) {
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < n; id += (long long)gridDim.x * blockDim.x) p[id] = v;
}

// f (row-major) -> layouts P and M
__global__ void k2_f_to_layouts(const Plan2 P, const double *__restrict__ f, double *__restrict__ fP, double *__restrict__ fM) {
    const int n = P.ext[1], l = P.ext[2];
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < P.N; id += gridDim.x * blockDim.x) {
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        const double v = f[id];
        fP[v2_offset_ijk(P, i, j, k, +1)] = v;
        fM[v2_offset_ijk(P, i, j, k, -1)] = v;
    }
}

// Which source goes into which buffer slot (= which CTA).  The number of rounds differs between sources
// (4..7 on the bench batch) and two CTAs share an SM, so an SM that happens to get two long sources finishes
// last.  The library remembers, per batch, the rounds every source needed in the previous call (spent[s];
// 0: unknown) and on which SM every CTA of that call ran (smid[b]; the placement of a launch with the same
// configuration repeats).  When both are known and all CTAs are resident at once (S <= 2 x SMs), the
// sources that needed the most rounds go to the CTAs that have an SM to themselves and the others are
// paired longest-with-shortest.  Anything else: caller order.  One CTA.
__global__ void k2_make_order(const int *__restrict__ spent, const int *__restrict__ smid, const int S,
                              const int nsm, int *__restrict__ order) {
    __shared__ int known, n1;
    if (threadIdx.x == 0) { known = 1; n1 = 0; }
    __syncthreads();
    for (int s = threadIdx.x; s < S; s += blockDim.x)
        if (spent[s] <= 0 || smid[s] < 0) known = 0;
    __syncthreads();
    const bool paired = known && S > nsm && S <= 2 * nsm;
    if (!paired) {
        for (int s = threadIdx.x; s < S; s += blockDim.x) order[s] = s;
        return;
    }
    // unit[b]: which rank CTA b takes.  Singles (in CTA order) take ranks 0..n1-1; pair number jp (pairs in the
    // order of their lower CTA index) takes ranks n1+jp (lower CTA) and S-1-jp (upper CTA).
    __shared__ int partner[1024], takes[1024];
    if (S > 1024) {      // not reached: paired implies S <= 2 x SMs
        for (int s = threadIdx.x; s < S; s += blockDim.x) order[s] = s;
        return;
    }
    for (int b = threadIdx.x; b < S; b += blockDim.x) {
        int same = 0, other = -1;
        for (int t = 0; t < S; t++)
            if (smid[t] == smid[b]) { same++; if (t != b) other = t; }
        partner[b] = same == 2 ? other : -1;
        if (same != 2) atomicAdd(&n1, 1);
    }
    __syncthreads();
    const int nsingle = n1;
    for (int b = threadIdx.x; b < S; b += blockDim.x) {
        const int lowb = (partner[b] >= 0 && partner[b] < b) ? partner[b] : b;   // the unit's lower CTA
        int js = 0, jp = 0;
        for (int t = 0; t < lowb; t++) {
            if (partner[t] < 0) js++;
            else if (t < partner[t]) jp++;
        }
        takes[b] = partner[b] < 0 ? js : (b == lowb ? nsingle + jp : S - 1 - jp);
    }
    __syncthreads();
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        int rank = 0;                                  // descending by rounds, ties by index
        const int ks = spent[s];
        for (int t = 0; t < S; t++) {
            const int kt = spent[t];
            rank += (kt > ks || (kt == ks && t < s)) ? 1 : 0;
        }
        for (int b = 0; b < S; b++)
            if (takes[b] == rank) order[b] = s;
    }
}

// u0 = `value` at every grid node of buffer 0 of each slot, written row by row of layout P (coalesced).
// grid: (blocks, S), 256 threads.  Used with k2_scatter_P when the sources are given as sparse lists.
__global__ void k2_fill_valid(const Plan2 P, double *__restrict__ bufs, const double value) {
    double *b0 = bufs + (long long)blockIdx.y * 3 * P.M;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nrows = P.dA * P.nmu;
    for (int row = blockIdx.x * nw + warp; row < nrows; row += gridDim.x * nw) {
        const int A = row / P.nmu, mu = row - A * P.nmu;
        const int c0 = mu - P.dW + 1 > 0 ? mu - P.dW + 1 : 0, c1 = mu < P.dC - 1 ? mu : P.dC - 1;
        double *r = b0 + ((long long)(A + 1) * P.RS + (mu + 1)) * P.PC;
        for (int C = c0 + lane; C <= c1; C += 32) r[C] = value;
    }
}

// sparse source values into buffer 0 of each slot (layout P); slot b holds source order[b].  One thread per
// slot, entries in list order (later entries win, like the Julia assignments of inversion.jl:52-60).
__global__ void k2_scatter_P(const Plan2 P, double *__restrict__ bufs, const int *__restrict__ order,
                             const int *__restrict__ src_ptr, const int *__restrict__ src_idx,
                             const double *__restrict__ src_val, const int S) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= S) return;
    const int s = order[slot];
    const int n = P.ext[1], l = P.ext[2];
    double *b0 = bufs + (long long)slot * 3 * P.M;
    for (int q = src_ptr[s]; q < src_ptr[s + 1]; q++) {
        const int id = src_idx[q];
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        b0[v2_offset_ijk(P, i, j, k, +1)] = src_val[q];
    }
}

// dense row-major u0 (S x N) -> buffer 0 of each source, layout P.  grid: (blocks, S)
__global__ void k2_u0_to_P(const Plan2 P, const double *__restrict__ U0, double *__restrict__ bufs,
                           const int *__restrict__ order) {
    const int n = P.ext[1], l = P.ext[2];
    const int slot = blockIdx.y;
    const double *u0 = U0 + (long long)order[slot] * P.N;
    double *b0 = bufs + (long long)slot * 3 * P.M;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < P.N; id += gridDim.x * blockDim.x) {
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        b0[v2_offset_ijk(P, i, j, k, +1)] = u0[id];
    }
}

// result (layout P, buffer where[src]) -> dense row-major U (S x N).  grid: (blocks, S)
__global__ void k2_P_to_rowmajor(const Plan2 P, const double *__restrict__ bufs, const int *__restrict__ where,
                                 double *__restrict__ U, const int *__restrict__ order) {
    const int n = P.ext[1], l = P.ext[2];
    const int slot = blockIdx.y;
    const double *b = bufs + ((long long)slot * 3 + where[slot]) * P.M;
    double *u = U + (long long)order[slot] * P.N;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < P.N; id += gridDim.x * blockDim.x) {
        const int k = id % l, t = id / l, j = t % n, i = t / n;
        u[id] = b[v2_offset_ijk(P, i, j, k, +1)];
    }
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// host: plan construction
// ---------------------------------------------------------------------------------------------
// Chooses the axis roles (A, W, C) for an m x n x l grid: lanes fill best when dC is a multiple of 8
// and dA of 4, a warp slot is live for dW + 10 levels of which dW are fully used, and role assignments
// with A = k need 6 instead of 4 layout changes per round.  nwarps: warps per CTA.
// Returns false when the grid cannot be handled (more than 32 column groups, 32-bit slots).
// LA x LC: the lane patch the plan is scored for (the team kernel uses 1 x 32 and has no limit on the groups).
inline bool v2_build_plan(Plan2 &P, int m, int n, int l, int nwarps, size_t plane_bytes, int LA = V2_LA,
                          int LC = V2_LC, int maxG = 32) {
    static const int SG[8][3] = {{1, 1, 1}, {-1, 1, 1}, {-1, -1, 1}, {1, -1, 1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, -1}};
    const int ext[3] = {m, n, l};
    double best = -1.0;
    int bestRole[3] = {0, 1, 2};
    static const int perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    static const int forced = getenv("ADTOMO_V2_PERM") ? atoi(getenv("ADTOMO_V2_PERM")) : -1;   // tuning aid: role assignment 0..5
    for (int p = 0; p < 6; p++) {
        if (forced >= 0 && p != forced) continue;
        const int dA = ext[perms[p][0]], dW = ext[perms[p][1]], dC = ext[perms[p][2]];
        const int G = (dC + LC - 1) / LC;
        if (G > maxG) continue;    // v2: lane g of a warp computes the window of column group g
        // layout changes per round (sigma = sW*sC along the reference's sweep order, cyclic)
        int changes = 0;
        for (int s = 0; s < 8; s++) {
            const int s0 = SG[s][perms[p][1]] * SG[s][perms[p][2]];
            const int s1 = SG[(s + 1) % 8][perms[p][1]] * SG[(s + 1) % 8][perms[p][2]];
            changes += (s0 != s1);
        }
        const double fill = (double)dC / (LC * G);
        const double live = (double)dW / (dW + LC + LA - 2);
        const int nrb = (dA + LA - 1) / LA;
        const double afill = (double)dA / (nrb * LA);
        double score = fill * live * afill * (1.0 - 0.02 * changes);
        if (perms[p][2] == 2) score *= 1.01;      // tie-break: lanes along the grid's fastest axis
        if (score > best) { best = score; for (int q = 0; q < 3; q++) bestRole[q] = perms[p][q]; }
    }
    if (best < 0) return false;
    for (int q = 0; q < 3; q++) { P.ext[q] = ext[q]; P.role[q] = bestRole[q]; }
    P.dA = ext[P.role[0]]; P.dW = ext[P.role[1]]; P.dC = ext[P.role[2]];
    P.nmu = P.dW + P.dC - 1;
    P.RS = P.nmu + 2;
    P.PC = ((P.dC + 1 + 3) / 4) * 4;
    P.nlev = P.dA + P.dW + P.dC - 2;
    P.G = (P.dC + LC - 1) / LC;
    P.NT = 32 * nwarps;
    P.PS = (P.dC + 1) & ~1;
    P.WCH = (int)(plane_bytes / (sizeof(double) * P.PS));
    if (P.WCH > P.dW) P.WCH = P.dW;
    if (P.WCH < 1) return false;
    P.N = (long long)m * n * l;
    P.M = (long long)(P.dA + 2) * P.RS * P.PC;
    if (P.M >= (1LL << 31) - 4 * (long long)P.RS * P.PC) return false;   // 32-bit slot arithmetic
    for (int s = 0; s < 8; s++)
        for (int q = 0; q < 3; q++) P.sg[s][q] = SG[s][P.role[q]];
    return true;
}

}  // namespace adtomo
