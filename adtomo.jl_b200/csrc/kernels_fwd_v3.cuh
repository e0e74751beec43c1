// kernels_fwd_v3.cuh -- 3D forward fast sweeping for BATCHES: the skewed-pencil layouts of
// kernels_fwd_v2.cuh with a sweep loop rebuilt around the instruction count.
//
// Reference semantics: Eikonal3D.cpp:35-57 (one directional Gauss-Seidel sweep), :59-68 (the 8 sweeps of
// a round), :71-88 (rounds until max|u - u_old| < tol, cap 20); level-by-level execution as in v2
// (bit-identical to the serial sweep).
//
// Why.  ncu on k_fwd3d_v2 (profiles/r01_ncu_summary_v2.json): the kernel is bound by instruction issue,
// not by HBM -- 95 G warp instructions for 403 M warp slots, i.e. ~206 executed SASS instructions per 32
// node updates (cuobjdump of the round-1 build), of which only ~56 are the fp64 work of the exact update.
// ~75 went into addresses (every neighbour offset is a run-time value: IADD3 + LEA + LEA.HI.X per load,
// 64-bit buffer bases re-derived per slot) and ~25 into locating the warp's next slot (ballot + find-leading-one +
// shuffle per slot on top of a window computation and a scan per level).  Skipping whole slots whose inputs did not
// change (bit-exact) was measured first and saves only 5-9 % of the slot evaluations on the
// bench model (benchmarks/skip_potential.c, profiles/r02_skip_potential.json): with tol = 1e-3 about 65 % of ALL
// node evaluations still lower the node's value by a tiny amount, so almost every 4x8x8 box is touched in
// every sweep.  The lever is instructions per slot:
//   * the row pitch PC is a COMPILE-TIME constant taken from a small menu (the plan pads the rows; pad
//     columns hold +inf like every other non-node slot), so the W and C neighbours are immediate
//     offsets of ONE address register; only the A neighbours (slab stride) need a 64-bit add each;
//   * warp slots are owned STATICALLY: the slots of a sweep are ranked by the level at which they become
//     live and dealt cyclically to the warps (rank mod #warps).  Every slot is live for the same number of
//     levels, so the live slots of a level are a contiguous range of ranks and every warp holds the same
//     number of them (+-1): the balance of v2's per-level dealing without its per-slot search.  A warp's
//     slots sit in a small shared-memory table (base offset, W' origin, first level); per level the warp
//     advances a [head, tail) window over its own list;
//   * loads of the warp's next slot are issued before the arithmetic of the current one, as in v2.
#pragma once
#include "kernels_fwd_v2.cuh"

namespace adtomo {

// Row pitches the batch kernel is compiled for (doubles; multiples of 8 so that a lane patch row is one
// 64-byte run).  dC + 1 <= PC.  0 = run-time pitch (any grid).
#if ADTOMO_V2_LC == 16
#define V3_PC_MENU(X) X(48) X(80) X(112) X(144) X(272)
#elif ADTOMO_V2_LC == 32
#define V3_PC_MENU(X) X(64) X(96) X(160) X(288)
#else
#define V3_PC_MENU(X) X(40) X(72) X(104) X(136) X(264)
#endif

inline int v3_menu_pitch(int dC) {
#define V3_PICK(pc_) if (dC + 1 <= pc_) return pc_;
    V3_PC_MENU(V3_PICK)
#undef V3_PICK
    return 0;
}

// Table entry of a warp slot: base = rb * offRB + g * LC (added to the lane/level part of the node's offset),
// meta = (sprime + V3_BIAS) | flags << 28 with sprime = rb * LA + SC * g * LC (W' origin) and flags = bit 0: last row
// block of a grid with dA % LA != 0, bit 1: last column group with dC % LC != 0 (some lanes have no pencil there).
struct V3Slot { int base, meta; };                // one int2 in shared memory
constexpr int V3_BIAS = 1 << 20;

// C' range (columns counted in the sweep's direction) of column group g
template <int SC>
EIK_HD void v3_cprange(const Plan2 &P, const int g, int &cpmin, int &cpmax) {
    const int c_lo = g * V2_LC, c_hi = (g * V2_LC + V2_LC - 1 < P.dC - 1) ? g * V2_LC + V2_LC - 1 : P.dC - 1;
    cpmin = SC > 0 ? c_lo : P.dC - 1 - c_hi;
    cpmax = SC > 0 ? c_hi : P.dC - 1 - c_lo;
}

// Rank of warp slot (rb, g) in the order (first live level, g): the number of slots that come before it.
template <int SC>
EIK_HD int v3_rank(const Plan2 &P, const int rb, const int g, int &s_out) {
    const int nrb = (P.dA + V2_LA - 1) / V2_LA;
    int cm, cx;
    v3_cprange<SC>(P, g, cm, cx);
    const int s = V2_LA * rb + cm;
    int r = 0;
    for (int g2 = 0; g2 < P.G; g2++) {
        int cm2, cx2;
        v3_cprange<SC>(P, g2, cm2, cx2);
        const int t = s - cm2;                    // row blocks rb2 of group g2 with V2_LA * rb2 < t start earlier
        int c = t > 0 ? (t + V2_LA - 1) / V2_LA : 0;
        if (c > nrb) c = nrb;
        r += c;
        if (g2 < g && t >= 0 && t % V2_LA == 0 && t / V2_LA < nrb) r++;    // same level, smaller g
    }
    s_out = s;
    return r;
}

// Table entry of slot (rb, g) for a sweep with signs (SA, SW, SC); PC: row pitch.
template <int SA, int SW, int SC>
EIK_HD V3Slot v3_make_slot(const Plan2 &P, const int PC, const int rb, const int g) {
    const int nrb = (P.dA + V2_LA - 1) / V2_LA;
    V3Slot d;
    d.base = rb * (V2_LA * (SA * P.RS - SW) * PC) + g * V2_LC;
    const int flags = ((rb == nrb - 1 && P.dA % V2_LA) ? 1 : 0) | ((g == P.G - 1 && P.dC % V2_LC) ? 2 : 0);
    d.meta = (rb * V2_LA + SC * g * V2_LC + V3_BIAS) | (flags << 28);
    return d;
}

// levels a slot stays live after its first one (the longest slot; shorter ones just find no node)
EIK_HD int v3_duration(const Plan2 &P) {
    return (P.dA < V2_LA ? P.dA - 1 : V2_LA - 1) + (P.dC < V2_LC ? P.dC - 1 : V2_LC - 1) + P.dW - 1;
}

// Lane mask applied to V3Slot::meta: keeps the flag bits of the edges where this lane has NO pencil, so that a
// flagged slot pushes the lane's W' far out of range.
inline bool v3_ragged(const Plan2 &P) { return P.dA % V2_LA != 0 || P.dC % V2_LC != 0; }
EIK_HD int v3_lane_mask(const Plan2 &P, const int la, const int lc) {
    const int nrb = (P.dA + V2_LA - 1) / V2_LA;
    const int bad = ((V2_LA * (nrb - 1) + la >= P.dA) ? 1 : 0) | ((V2_LC * (P.G - 1) + lc >= P.dC) ? 2 : 0);
    return 0x0fffffff | (bad << 28);
}

// Node of a lane in slot d at level lam.  lamOff = L.offc + lam * SW * PC, lamWb = L.wqc + lam + V3_BIAS.
// off: the slot of the lane's pencil position -- ALWAYS a loadable address: for a lane without a node it lies up to
// LA + LC rows outside the slab or up to LA - 1 slabs outside the field (the buffers are allocated with
// v3_slack() doubles on both sides); act: the position is a grid node.
// RG = false: the plan has no ragged edge (dA % LA == 0 and dC % LC == 0): no flag is ever set and the lane mask --
// one more live register in a loop that already spills at 64 -- is not needed.
template <bool RG = true>
EIK_HD void v3_node(const Plan2 &P, const V3Slot &d, const int lamOff, const int lamWb, const int lmask, int &off, bool &act) {
    const int wq = lamWb - (RG ? (d.meta & lmask) : d.meta);
    act = (unsigned)wq < (unsigned)P.dW;
    off = lamOff + d.base;
}

// doubles of slack the field / slowness buffers need before and after (idle lanes load, never store, there)
inline long long v3_slack(const Plan2 &P) { return (long long)(V2_LA + 1) * P.RS * P.PC; }

// L2 eviction policies (createpolicy; they travel in the memory descriptor of the load / store, i.e. cost uniform
// registers only).  ncu on the bench batch: the kernel reads 244 GB from DRAM per launch, 2.5x the fields it sweeps --
// with 256 sources in flight the 126 MB L2 holds neither the level fronts nor the slowness field that ALL sources read
// once per sweep.  ADTOMO_F_POLICY=1: slowness loads are evict_last (the two layouts of f, 29 MB at C3, stay resident);
// ADTOMO_RESKEW_POLICY=1: the re-skew's loads and stores and the L-inf reference loads (one touch per round) are
// evict_first.  Performance hints only.
#ifndef ADTOMO_F_POLICY
#define ADTOMO_F_POLICY 0
#endif
#ifndef ADTOMO_RESKEW_POLICY
#define ADTOMO_RESKEW_POLICY 0
#endif
struct V3Pol { unsigned long long first, last; };
#if defined(__CUDACC__)
__device__ __forceinline__ V3Pol v3_policies() {
    V3Pol q;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(q.first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(q.last));
    return q;
}
#endif
// slowness of a node (read-only for the whole launch)
EIK_HD double v3_ld_f(const double *p, const V3Pol &pol) {
#if defined(__CUDA_ARCH__) && ADTOMO_F_POLICY == 1
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol.last));
    return v;
#else
    return *p;
#endif
}
// a value that is touched once (re-skew source, L-inf reference)
EIK_HD double v3_ld_once(const double *p, const V3Pol &pol) {
#if defined(__CUDA_ARCH__) && ADTOMO_RESKEW_POLICY >= 1
    double v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol.first));
    return v;
#else
    return *p;
#endif
}
EIK_HD void v3_st_once(double *p, const double v, const V3Pol &pol) {
#if defined(__CUDA_ARCH__) && ADTOMO_RESKEW_POLICY == 1
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol.first) : "memory");
#else
    *p = v;
#endif
}

// Loads of one node (same values as v2_load).  PCT: compile-time pitch or 0.  sAb: slab stride in bytes (RS * PC * 8).
template <int SA, int SW, int SC, bool OOP, bool CMP, int PCT>
EIK_HD void v3_load_old(const Plan2 &P, const long long sAb, const int off, const bool act, const double *rd,
                        const double *__restrict__ fl, const double *cmp, V2Vals &V, const V3Pol &pol) {
    // the node's own value, the slowness and the three DOWNWIND values: nothing this sweep has written when the node's
    // level starts (they belong to this level and the next), so they may be loaded before the previous level's barrier
    const int PC = PCT ? PCT : P.PC;
    const int offW = SW * PC, offC = SW * PC + SC;
    V.off = act ? off : -1;
    const char *p = reinterpret_cast<const char *>(rd + off);
#define V3_AT(ptr_, bytes_) (*reinterpret_cast<const double *>((ptr_) + (bytes_)))
    V.own = V3_AT(p, 0);
    V.fv = v3_ld_f(fl + off, pol);
    V.dW = V3_AT(p, offW * 8);
    V.dC = V3_AT(p, offC * 8);
    V.dA = V3_AT(p, SA * sAb);
    V.ref = CMP ? v3_ld_once(cmp + off, pol) : 0.0;
}

// the three UPWIND values (results of the previous level: only after that level's barrier).  off: the slot's offset
// (v3_node), also for lanes without a node.
template <int SA, int SW, int SC, bool OOP, int PCT>
EIK_HD void v3_load_up(const Plan2 &P, const long long sAb, const int off, const double *rd, const double *wr, V2Vals &V) {
    const int PC = PCT ? PCT : P.PC;
    const int offW = SW * PC, offC = SW * PC + SC;
    const char *pu = reinterpret_cast<const char *>((OOP ? wr : rd) + off);
    V.uW = V3_AT(pu, -offW * 8);
    V.uC = V3_AT(pu, -offC * 8);
    V.uA = V3_AT(pu, -SA * sAb);
#undef V3_AT
}

template <int SA, int SW, int SC, bool OOP, bool CMP, int PCT>
EIK_HD void v3_load(const Plan2 &P, const long long sAb, const int off, const bool act, const double *rd, const double *wr,
                    const double *__restrict__ fl, const double *cmp, V2Vals &V, const V3Pol &pol) {
    v3_load_old<SA, SW, SC, OOP, CMP, PCT>(P, sAb, off, act, rd, fl, cmp, V, pol);
    v3_load_up<SA, SW, SC, OOP, PCT>(P, sAb, off, rd, wr, V);
}

// Re-skew of the W-chunk [w0, w0 + wc) of one slab through the plane (wc x PS doubles, un-skewed rows Wl).  Element
// (Wl, C) lies in row w0 + 1 + Wl + cc of layout sigma, cc = C (P) or dC-1-C (M).  Thread (warp, lane = C mod 32) moves
// the elements Wl = t0 + j * nw, j = 0, 1, ..., with t0 = (warp - cc) mod nw: at step j the lanes of a warp touch at
// most three rows of the layout, each in a contiguous run of C (coalesced), every element exactly once, and both the
// plane slot and the slab offset advance by constants -- no wrap, no per-element index arithmetic.
EIK_HD void v3_reskew_start(const Plan2 &P, const int PC, const int sigma, const int w0, const int nw, const int warp,
                            const int C, int &t0, int &pl0, int &go0) {
    const int cc = sigma > 0 ? C : P.dC - 1 - C;
    t0 = (warp - cc) % nw;
    if (t0 < 0) t0 += nw;
    pl0 = t0 * P.PS + C;
    go0 = (w0 + 1 + cc + t0) * PC + C;
}

#if defined(__CUDACC__)

// Builds the CTA's slot table for a sweep: tab[w * maxPer + j] = j-th slot of warp w (ranks w, w + nw, ...),
// tabS[same] = its first live level.
template <int SA, int SW, int SC>
__device__ __forceinline__ void v3_build_table(const Plan2 &P, const int PC, V3Slot *tab, int *tabS, const int maxPer) {
    const int nw = blockDim.x >> 5;
    const int nrb = (P.dA + V2_LA - 1) / V2_LA, nslots = nrb * P.G;
    for (int idx = threadIdx.x; idx < nslots; idx += blockDim.x) {
        const int rb = idx / P.G, g = idx - rb * P.G;
        int s;
        const int r = v3_rank<SC>(P, rb, g, s);
        const int at = (r % nw) * maxPer + r / nw;
        tab[at] = v3_make_slot<SA, SW, SC>(P, PC, rb, g);
        tabS[at] = s;
    }
}

// One phase of the re-skew of a chunk (see v3_reskew_start).  PHASE 0: rows of layout sigma -> plane; 1: plane -> rows
// of layout sigma.  NW: warps of the CTA (0 = run time), PCT: row pitch (0 = run time): with both known the eight
// elements a thread has in flight sit at immediate offsets of one address.  The round-2 profile
// (profiles/r02_ncu_summary_v3b.json) had the previous form of this pass at 22 % of the kernel's instructions
// (~50 per element and phase: the index map was rematerialised for every element under the 64-register cap).
template <int PHASE, int NW, int PCT>
__device__ __forceinline__ void v3_reskew_rows(const Plan2 &P, const double *s, double *d, const int sigma, double *plane,
                                               const int w0, const int wc, const V3Pol &pol) {
    constexpr int U = 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = NW ? NW : (int)(blockDim.x >> 5);
    const int PC = PCT ? PCT : P.PC;
    const int stepP = nw * P.PS, stepG = nw * PC;
    for (int C = lane; C < P.dC; C += 32) {
        int t, pl, go;
        v3_reskew_start(P, PC, sigma, w0, nw, warp, C, t, pl, go);
        const double *gs = s + go;
        double *gd = d + go;
        double *pp = plane + pl;
        for (; t < wc; t += U * nw, gs += U * stepG, gd += U * stepG, pp += U * stepP) {
            double x[U];
#pragma unroll
            for (int j = 0; j < U; j++) {
                x[j] = 0.0;
                if (t + j * nw < wc) x[j] = PHASE == 0 ? v3_ld_once(gs + j * stepG, pol) : pp[j * stepP];
            }
#pragma unroll
            for (int j = 0; j < U; j++)
                if (t + j * nw < wc) {
                    if (PHASE == 0) pp[j * stepP] = x[j];
                    else v3_st_once(gd + j * stepG, x[j], pol);
                }
        }
    }
}

// slabs A in [A0, A1); same contract as v2_reskew
template <int PCT>
__device__ __forceinline__ void v3_reskew(const Plan2 &P, const double *src, double *dst, const int sigmaFrom,
                                          double *plane, const int A0, const int A1, const V3Pol &pol) {
    const int PC = PCT ? PCT : P.PC;
    for (int A = A0; A < A1; A++) {
        const double *s = src + (long long)(A + 1) * P.RS * PC;
        double *d = dst + (long long)(A + 1) * P.RS * PC;
        for (int w0 = 0; w0 < P.dW; w0 += P.WCH) {
            const int wc = (P.dW - w0 < P.WCH) ? P.dW - w0 : P.WCH;
            if (blockDim.x == 512) v3_reskew_rows<0, 16, PCT>(P, s, d, sigmaFrom, plane, w0, wc, pol);
            else v3_reskew_rows<0, 0, PCT>(P, s, d, sigmaFrom, plane, w0, wc, pol);
            __syncthreads();
            if (blockDim.x == 512) v3_reskew_rows<1, 16, PCT>(P, s, d, -sigmaFrom, plane, w0, wc, pol);
            else v3_reskew_rows<1, 0, PCT>(P, s, d, -sigmaFrom, plane, w0, wc, pol);
            __syncthreads();
        }
    }
}

template <int SA, int SW, int SC, bool OOP, bool CMP, int PCT, bool RG>
__device__ __forceinline__ void v3_sweep(const Plan2 &P, V3Slot *tab, int *tabS, const int maxPer, const double *rd, double *wr,
                                         const double *__restrict__ fl, const double *cmp, const double h, double &err,
                                         const V3Pol &pol) {
    const int PC = PCT ? PCT : P.PC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const V2Lane L = v2_lane_setup<SA, SW, SC>(P, lane);      // P.PC == PC (the plan was built for this pitch)
    const int lmask = v3_lane_mask(P, L.la, L.lc);
    const int nslots = ((P.dA + V2_LA - 1) / V2_LA) * P.G;
    const int cnt = warp < nslots ? (nslots - warp + nw - 1) / nw : 0;
    const int dur = v3_duration(P);
    const long long sAb = (long long)P.RS * PC * 8;
    __syncthreads();     // the previous sweep (its field writes, its use of the table) is complete
    v3_build_table<SA, SW, SC>(P, PC, tab, tabS, maxPer);
    __syncthreads();
    const int2 *mine = reinterpret_cast<const int2 *>(tab) + warp * maxPer;
    const int *mineS = tabS + warp * maxPer;
    int head = 0, tail = 0;
    int sTail = cnt > 0 ? mineS[0] : 0x7fffffff;     // first level of the next slot to enter the window
    int sHead = sTail;                                // first level of the oldest slot in the window
    int lamOff = L.offc, lamWb = L.wqc + V3_BIAS;
    for (int lam = 0; lam < P.nlev; lam++, lamOff += SW * PC, lamWb++) {
        while (sTail <= lam) {
            tail++;
            sTail = tail < cnt ? mineS[tail] : 0x7fffffff;
        }
        while (head < tail && sHead + dur < lam) {
            head++;
            sHead = head < cnt ? mineS[head] : 0x7fffffff;
        }
        // The loads of the warp's next slot are issued before the current slot's arithmetic (software pipelining by
        // hand).  Measured and rejected (profiles/r02_v3_phase_cycles.json): a second register set (spills at 64
        // registers; CTAs of 384 / 256 threads with 80 / 114 registers lose more to the missing warps), carrying the
        // look-ahead across the level barrier, L2 prefetch of the next level's rows; round 2, second session
        // (profiles/r02_forward_experiments.json): the three COLD values of a slot (slowness, downwind W / A) asked for
        // two slots ahead and across the barrier with the five warm ones loaded at use (161.8 vs 137.5 ms: the warm
        // values are not reliable L1 hits), L2 eviction hints and a persisting window for the slowness (no change in
        // DRAM bytes), a smaller re-skew plane for more L1 (slower).
#define V3_LOAD(V_)                                                                              \
    do {                                                                                         \
        const int2 d2__ = *dp++;                                                                 \
        V3Slot d__;                                                                              \
        d__.base = d2__.x; d__.meta = d2__.y;                                                    \
        int off__; bool act__;                                                                   \
        v3_node<RG>(P, d__, lamOff, lamWb, lmask, off__, act__);                                 \
        v3_load<SA, SW, SC, OOP, CMP, PCT>(P, sAb, off__, act__, rd, wr, fl, cmp, V_, pol);      \
    } while (0)
        int left = tail - head;
        if (left > 0) {
            const int2 *dp = mine + head;
            V2Vals V;
            V3_LOAD(V);
#pragma unroll 1
            for (;;) {
                V2Prep Q;
                v2_prep(V, Q);                     // consumes V: its registers take the next slot's loads
                if (--left > 0) V3_LOAD(V);
                v2_solve<OOP, CMP>(Q, wr, h, err);
                if (left <= 0) break;
            }
        }
#undef V3_LOAD
        __syncthreads();
    }
}

// ---- staged variant: the eight values of a node travel global -> shared memory with cp.async, two warp slots ahead ----
// ncu on the register-pipelined sweep (profiles/r02_ncu_summary_v3a.json): 32 % of the resident warps wait on a load
// (long scoreboard at the first use of the next slot's values) although those loads were issued a whole solve earlier:
// with ~7 of 8 warps per scheduler needed in the arithmetic to fill the issue slots, one slot of look-ahead is not
// enough, and a second register set does not fit 64 registers.  Here the look-ahead lives in shared memory (the re-skew
// plane is idle during a sweep): 2 stages x 8 values x 32 lanes x 8 B = 4 KB per warp, lane-private, so cp.async's
// per-thread completion (wait_group) is the only synchronisation.
template <int SOFF, int GOFF>
__device__ __forceinline__ void v3_cp8(const unsigned s, const void *g) {
    asm volatile("cp.async.ca.shared.global [%0 + %2], [%1 + %3], 8;" ::"r"(s), "l"(g), "n"(SOFF), "n"(GOFF) : "memory");
}
__device__ __forceinline__ void v3_cp8r(const unsigned s, const void *g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void v3_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void v3_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int V3_STAGE_BYTES_PER_WARP = 2 * 8 * 32 * 8;

template <int SA, int SW, int SC, bool OOP, bool CMP, int PCT, bool RG>
__device__ __forceinline__ void v3_sweep_staged(const Plan2 &P, V3Slot *tab, int *tabS, const int maxPer, double *stage,
                                                const double *rd, double *wr, const double *__restrict__ fl,
                                                const double *cmp, const double h, double &err) {
    static_assert(PCT != 0, "the staged sweep needs a compile-time pitch");
    constexpr int PC = PCT;
    constexpr int offW8 = SW * PC * 8, offC8 = (SW * PC + SC) * 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const V2Lane L = v2_lane_setup<SA, SW, SC>(P, lane);
    const int lmask = v3_lane_mask(P, L.la, L.lc);
    const int nslots = ((P.dA + V2_LA - 1) / V2_LA) * P.G;
    const int cnt = warp < nslots ? (nslots - warp + nw - 1) / nw : 0;
    const int dur = v3_duration(P);
    const long long sAb = (long long)SA * P.RS * PC * 8;
    __syncthreads();     // the previous sweep / re-skew (field writes, table, plane) is complete
    v3_build_table<SA, SW, SC>(P, PC, tab, tabS, maxPer);
    __syncthreads();
    const int2 *mine = reinterpret_cast<const int2 *>(tab) + warp * maxPer;
    const int *mineS = tabS + warp * maxPer;
    double *sp = stage + warp * (V3_STAGE_BYTES_PER_WARP / 8) + lane;              // value k of stage d: sp[(d * 8 + k) * 32]
    const unsigned sb = (unsigned)__cvta_generic_to_shared(sp);
    int head = 0, tail = 0;
    int sTail = cnt > 0 ? mineS[0] : 0x7fffffff;
    int sHead = sTail;
    int lamOff = L.offc, lamWb = L.wqc + V3_BIAS;
    for (int lam = 0; lam < P.nlev; lam++, lamOff += SW * PC, lamWb++) {
        while (sTail <= lam) {
            tail++;
            sTail = tail < cnt ? mineS[tail] : 0x7fffffff;
        }
        while (head < tail && sHead + dur < lam) {
            head++;
            sHead = head < cnt ? mineS[head] : 0x7fffffff;
        }
        // asks for the eight values of the warp's next slot (stage par_: 0 or 2048 bytes); off_ receives V.off
#define V3_ISSUE(par_, off_)                                                                     \
    do {                                                                                         \
        const int2 d2__ = *dp++;                                                                 \
        V3Slot d__;                                                                              \
        d__.base = d2__.x; d__.meta = d2__.y;                                                    \
        int off__; bool act__;                                                                   \
        v3_node<RG>(P, d__, lamOff, lamWb, lmask, off__, act__);                                 \
        off_ = act__ ? off__ : -1;                                                               \
        const unsigned s__ = sb + (par_);                                                        \
        const char *p__ = reinterpret_cast<const char *>(rd + off__);                            \
        const char *pu__ = OOP ? reinterpret_cast<const char *>(wr + off__) : p__;               \
        v3_cp8<0 * 256, 0>(s__, p__);                                                            \
        v3_cp8r(s__ + 1 * 256, fl + off__);                                                      \
        v3_cp8<2 * 256, offW8>(s__, p__);                                                        \
        v3_cp8<3 * 256, offC8>(s__, p__);                                                        \
        v3_cp8r(s__ + 4 * 256, p__ + sAb);                                                       \
        v3_cp8<5 * 256, -offW8>(s__, pu__);                                                      \
        v3_cp8<6 * 256, -offC8>(s__, pu__);                                                      \
        v3_cp8r(s__ + 7 * 256, pu__ - sAb);                                                      \
        v3_cp_commit();                                                                          \
    } while (0)
        int left = tail - head;
        if (left > 0) {
            const int2 *dp = mine + head;
            int off0, off1 = -1;               // V.off of the two slots in flight (oldest first)
            V3_ISSUE(0, off0);
            if (left > 1) V3_ISSUE(2048, off1);
            int par = 0;
#pragma unroll 1
            for (;;) {
                if (left > 1) v3_cp_wait<1>(); else v3_cp_wait<0>();
                V2Vals V;
                const double *q = sp + par * 8;              // par = 0 / 32: stage 0 / 1 (8 values x 32 lanes apart)
                V.own = q[0 * 32]; V.fv = q[1 * 32]; V.dW = q[2 * 32]; V.dC = q[3 * 32];
                V.dA = q[4 * 32]; V.uW = q[5 * 32]; V.uC = q[6 * 32]; V.uA = q[7 * 32];
                V.off = off0;
                V.ref = CMP ? cmp[off0 < 0 ? 0 : off0] : 0.0;      // used last, after the solve
                V2Prep Q;
                v2_prep(V, Q);                 // the stage is free again: its values sit in registers
                off0 = off1;
                if (left > 2) V3_ISSUE(par * 64, off1);
                v2_solve<OOP, CMP>(Q, wr, h, err);
                par ^= 32;
                if (--left <= 0) break;
            }
        }
#undef V3_ISSUE
        __syncthreads();
    }
}

// One round (the 8 sweeps with their re-skews, Eikonal3D.cpp:59-68) of the source whose buffers start at B3; o / a: the
// buffers that hold the round-start field / receive the result.  Returns the round's L-inf change (block-uniform).
template <int PCT, bool STG, bool RG>
__device__ __forceinline__ double v3_round(const Plan2 &P, V3Slot *tab, int *tabS, const int maxPer, double *plane, double *red,
                                           double *B3, const int o, const int a, const double *__restrict__ fP,
                                           const double *__restrict__ fM, const double h, const V3Pol &pol) {
    double err = 0.0;
    double *Bo = B3 + o * P.M, *Ba = B3 + a * P.M, *Bz = B3 + 2 * P.M;
    int state = 1;                        // layout of the working field
    double *w = Ba;
    for (int sw = 0; sw < 8; sw++) {
        const int sigma = P.sg[sw][1] * P.sg[sw][2];
        if (sw > 0 && sigma != state) {
            double *dst = state > 0 ? Bz : Ba;
            __syncthreads();
            v3_reskew<PCT>(P, w, dst, state, plane, 0, P.dA, pol);
            w = dst;
            state = sigma;
        }
#define V3_CALL(a_, w_, c_, oop_, cmp_)                                                                                  \
    do {                                                                                                                 \
        if constexpr (STG && PCT != 0)                                                                                   \
            v3_sweep_staged<a_, w_, c_, oop_, cmp_, PCT, RG>(P, tab, tabS, maxPer, plane, oop_ ? Bo : w, w, sigma > 0 ? fP : fM, Bo, h, err); \
        else                                                                                                             \
            v3_sweep<a_, w_, c_, oop_, cmp_, PCT, RG>(P, tab, tabS, maxPer, oop_ ? Bo : w, w, sigma > 0 ? fP : fM, Bo, h, err, pol); \
    } while (0)
        V2_DISPATCH(P, sw, V3_CALL);
#undef V3_CALL
    }
    return v2_block_max(err, red);
}

// Same contract as k_fwd3d_v2 (buffers, order, rounds, errs, where, spent); the field and slowness buffers have
// v3_slack() loadable doubles on both sides.  Dynamic shared memory: the re-skew plane (WCH x PS doubles) followed by
// the slot table (nw x maxPer int2, then nw x maxPer int).
template <int NTMAX, int MINB, int PCT, bool STG, bool RG>
__global__ void __launch_bounds__(NTMAX, MINB) k_fwd3d_v3(const Plan2 P, const int tabOffset, const int maxPer, double *bufs,
                                                          const double *__restrict__ fP, const double *__restrict__ fM,
                                                          const double h, const double tol, const int max_rounds,
                                                          const int S, int *__restrict__ rounds, double *__restrict__ errs,
                                                          int *__restrict__ where, const int *__restrict__ order,
                                                          int *__restrict__ spent) {
    extern __shared__ double plane[];
    __shared__ double red[32];
    V3Slot *tab = reinterpret_cast<V3Slot *>(reinterpret_cast<char *>(plane) + tabOffset);
    int *tabS = reinterpret_cast<int *>(tab + (blockDim.x >> 5) * maxPer);
    const V3Pol pol = v3_policies();
    for (int src = blockIdx.x; src < S; src += gridDim.x) {
        double *B3 = bufs + (long long)src * 3 * P.M;
        int o = 0, a = 1;          // layout P: round-start field, working field;  buffer 2: layout M
        int r = 0;
        bool conv = false;
        while (r < max_rounds) {
            const double e = v3_round<PCT, STG, RG>(P, tab, tabS, maxPer, plane, red, B3, o, a, fP, fM, h, pol);
            if (threadIdx.x == 0 && errs) errs[(long long)order[src] * max_rounds + r] = e;
            r++;
            const int oo = o; o = a; a = oo;      // the result (in a) becomes next round's round-start field
            if (__any_sync(0xffffffffu, e < tol)) { conv = true; break; }   // e is block-uniform
        }
        if (threadIdx.x == 0) {
            if (rounds) rounds[order[src]] = conv ? r : -r;
            spent[order[src]] = r;
            {
                unsigned sm__;
                asm("mov.u32 %0, %%smid;" : "=r"(sm__));
                spent[S + src] = (int)sm__;
            }
            where[src] = o;
        }
        __syncthreads();
    }
}

#endif  // __CUDACC__

// Plan of the batch kernel: v2's plan with the row pitch taken from the menu (PCT of the kernel instantiation to
// launch is returned in *pct; 0 = no menu entry fits, run-time pitch).  The padded rows cost memory, not traffic.
inline bool v3_build_plan(Plan2 &P, int m, int n, int l, int nwarps, size_t plane_bytes, int *pct, bool use_menu = true) {
    if (!v2_build_plan(P, m, n, l, nwarps, plane_bytes)) return false;
    const int pc = use_menu ? v3_menu_pitch(P.dC) : 0;
    *pct = 0;
    if (pc) {
        const long long M = (long long)(P.dA + 2) * P.RS * pc;
        if (M < (1LL << 31) - 4 * (long long)P.RS * pc) { P.PC = pc; P.M = M; *pct = pc; }
    }
    return true;
}

// slots per warp in the shared-memory table
inline int v3_max_per_warp(const Plan2 &P) {
    const int nslots = ((P.dA + V2_LA - 1) / V2_LA) * P.G, nw = P.NT / 32;
    return (nslots + nw - 1) / nw;
}

}  // namespace adtomo
