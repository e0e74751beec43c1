// kernels_fwd_team.cuh -- 3D forward fast sweeping for FEW sources: a TEAM of CTAs per source.
//
// Reference semantics: Eikonal3D.cpp:35-57 (one directional Gauss-Seidel sweep), :59-68 (the 8 sweeps of
// a round), :71-88 (rounds until max|u - u_old| < tol).  Same level-by-level execution and the same
// skewed-pencil layouts as kernels_fwd_v2.cuh (constant-offset neighbours, +inf padding, in-place
// sweeps, re-skew between the layouts P and M), so the result is the serial sweep's, bit for bit.
//
// Why.  With one CTA (v2) or one cluster of <= 8 SMs (level-major kernel) per source, a single large
// grid (BASELINE config C5: 256^3 .. 512^3, one source) uses 1-8 of the 148 SMs and one level costs a
// full load->solve->store->barrier latency for a handful of warps.  Here the A rows of a source are
// split over up to 2 x #SM co-resident CTAs (cooperative launch).  A sweep's data dependence between
// two CTAs is one-directional: the first row of CTA t at level lam needs the last row of CTA t-1 at
// level lam-1 (new value), and CTA t may overwrite that node's level-lam value only after CTA t-1 has
// read it (in-place update).  Both are carried by a MAILBOX in the style of NCCL's LL protocol: the
// warp that updates a node of a CTA's last row also writes the new value as two 8-byte packets
// {32 value bits | 32-bit tag}, tag = (sweep serial, level); the consumer takes its upwind-A value from
// the mailbox and spins on the packet until both tags match.  A packet is written after its node was
// computed, i.e. after the producer read the old downwind value, so its arrival also licenses the
// overwrite.  No fence, no flag and no grid-wide barrier on the level path (a first version with one
// progress word per CTA spent 0.9 us per level in MEMBAR.SC.GPU and ~10 polls of LDG.STRONG+CCTL.IVALL
// per warp and level): the CTAs form a systolic pipeline skewed by one L2 round trip per CTA.  A mailbox
// has one slot per node of the boundary row (a slab of the skewed layout), so a producer can run
// arbitrarily far ahead; tags only grow, across sweeps and across launches, so slots are never reset.
// Team-wide barriers (a counter in global memory, with fences) separate the sweeps, the re-skews and
// the rounds: ~13 per round.
//
// Lanes of a warp cover 32 consecutive columns C of ONE row A (a warp slot): at a fixed level these are
// 32 consecutive doubles of one skewed row mu -- a 256-byte contiguous run for the node and for each of
// its six neighbours (512 bytes for its mailbox packets).
#pragma once
#include <cstring>
#include "kernels_fwd_v2.cuh"

namespace adtomo {

constexpr int TM_LC = 32;          // columns per warp slot
constexpr int TM_LEVEL_BITS = 12;  // packet tag = (sweep serial << 12) | (level + 1); nlev < 4095
constexpr int TM_SYNC_WORDS = 8;   // per-source sync area (unsigned words): [0] barrier counter, [2..3],[4..5] err slots

struct TeamCfg {
    int nC;              // CTAs per source
    int R;               // rows (A') per CTA
    int G32;             // column groups of 32 per row
    long long mbStride;  // packets (16 bytes each) of one CTA's inbox = RS * PC (one slab of the skewed layout)
};

#if defined(__CUDA_ARCH__)
#define TM_LDU(p) __ldcg(p)        // fields other CTAs write: bypass L1
#else
#define TM_LDU(p) (*(p))
#endif

typedef unsigned long long tm_u64;

// A value as two packets {low 32 value bits | tag << 32}, {high 32 value bits | tag << 32}.
EIK_HD void tm_pack(const double v, const unsigned tag, tm_u64 &p0, tm_u64 &p1) {
    tm_u64 b;
#if defined(__CUDA_ARCH__)
    b = (tm_u64)__double_as_longlong(v);
#else
    memcpy(&b, &v, 8);
#endif
    p0 = (b & 0xffffffffULL) | ((tm_u64)tag << 32);
    p1 = (b >> 32) | ((tm_u64)tag << 32);
}
EIK_HD bool tm_unpack(const tm_u64 p0, const tm_u64 p1, const unsigned tag, double &v) {
    const tm_u64 b = (p0 & 0xffffffffULL) | (p1 << 32);
#if defined(__CUDA_ARCH__)
    v = __longlong_as_double((long long)b);
#else
    memcpy(&v, &b, 8);
#endif
    return (unsigned)(p0 >> 32) == tag && (unsigned)(p1 >> 32) == tag;
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void tm_mb_store(tm_u64 *slot, const tm_u64 p0, const tm_u64 p1) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(p0), "l"(p1) : "memory");
}
__device__ __forceinline__ void tm_mb_load(const tm_u64 *slot, tm_u64 &p0, tm_u64 &p1) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(p0), "=l"(p1) : "l"(slot) : "memory");
}
#else
inline void tm_mb_store(tm_u64 *slot, const tm_u64 p0, const tm_u64 p1) { slot[0] = p0; slot[1] = p1; }
inline void tm_mb_load(const tm_u64 *slot, tm_u64 &p0, tm_u64 &p1) { p0 = slot[0]; p1 = slot[1]; }
#endif

// Loads of one lane's node of warp slot (row Ap, column group g) at level lam; coordinates with a
// prime are counted in the sweep's direction.  Same offsets as v2_load, computed per slot.
// inbox != nullptr: Ap is the CTA's first row and its upwind-A values come from the mailbox (packets of
// level lam-1, tag `tag_in`).  mb receives the node's mailbox slot (offset inside its slab).
// Host build: a packet that has not arrived yields NaN (the emulation's scheduler must prevent that).
template <int SA, int SW, int SC, bool OOP, bool CMP>
EIK_HD void tm_load(const Plan2 &P, const int lane, const int lam, const int Ap, const int g, const double *rd,
                    const double *wr, const double *__restrict__ fl, const double *cmp, const tm_u64 *inbox,
                    const unsigned tag_in, V2Vals &V, int &mb) {
    const int offA = SA * P.RS * P.PC, offW = SW * P.PC, offC = SW * P.PC + SC;   // downwind (old, level+1)
    const int Cp = g * TM_LC + lane;
    const int Wp = lam - Ap - Cp;
    const bool act = (unsigned)Wp < (unsigned)P.dW && Cp < P.dC;
    const int A = SA > 0 ? Ap : P.dA - 1 - Ap;
    const int C = SC > 0 ? Cp : P.dC - 1 - Cp;
    const int mu = SW > 0 ? lam - Ap : P.nmu - 1 - (lam - Ap);      // uniform over the warp: one skewed row
    mb = (mu + 1) * P.PC + C;
    const int off = act ? (A + 1) * P.RS * P.PC + mb : (P.RS + 1) * P.PC + 1;
    V.off = act ? off : -1;
    const double *p = rd + off;
    V.own = TM_LDU(p);
    V.fv = fl[off];
    V.dA = TM_LDU(p + offA);
    V.dW = TM_LDU(p + offW);
    V.dC = TM_LDU(p + offC);
    const double *pu = OOP ? wr + off : p;
    V.uW = TM_LDU(pu - offW);
    V.uC = TM_LDU(pu - offC);
    V.ref = CMP ? TM_LDU(cmp + off) : 0.0;
    if (inbox) {
        V.uA = 0.0;
        if (act) {
            // the A-neighbour has the same (W, C), i.e. the same slot inside ITS slab
            tm_u64 p0, p1;
#if defined(__CUDA_ARCH__)
            do { tm_mb_load(inbox + 2 * (long long)mb, p0, p1); } while (!tm_unpack(p0, p1, tag_in, V.uA));
#else
            tm_mb_load(inbox + 2 * (long long)mb, p0, p1);
            if (!tm_unpack(p0, p1, tag_in, V.uA)) V.uA = NAN;
#endif
        }
    } else {
        V.uA = TM_LDU(pu - offA);
    }
}

// v2_finish + the packet for the downstream CTA (outbox != nullptr: the node is in the CTA's last row).
// The packet is written for EVERY node of the row, changed or not: the consumer waits for it.
template <bool OOP, bool CMP>
EIK_HD void tm_finish(const V2Vals &V, double *wr, const double h, double &err, tm_u64 *outbox, const int mb,
                      const unsigned tag_out) {
    if (V.off < 0) return;
    V2Prep Q;
    v2_prep(V, Q);
    double res = Q.own;
    bool changed = false;
    if (Q.a1 < Q.own) {
        const double un = eik_solve3_sorted(Q.a1, Q.a2, Q.a3, Q.fv * h, Q.fv * Q.fv * h * h);
        if (un < Q.own) { res = un; changed = true; }
    }
    if (OOP || changed) wr[Q.off] = res;
    if (outbox) {
        tm_u64 p0, p1;
        tm_pack(res, tag_out, p0, p1);
        tm_mb_store(outbox + 2 * (long long)mb, p0, p1);
    }
    if (CMP) {
        const double dd = fabs(res - Q.ref);
        err = (err < dd) ? dd : err;
    }
}

// rows [a0, a1) of team member t, first and last level at which one of them has a node
EIK_HD void tm_rows(const Plan2 &P, const TeamCfg &T, const int t, int &a0, int &a1, int &lam0, int &lam1) {
    a0 = t * T.R;
    a1 = a0 + T.R < P.dA ? a0 + T.R : P.dA;
    lam0 = a0;
    lam1 = a1 - 1 + P.dW - 1 + P.dC - 1;
}

// warp-uniform: does slot (Ap, g) have a node at level lam?
EIK_HD bool tm_slot_live(const Plan2 &P, const int lam, const int Ap, const int g) {
    const int top = lam - Ap - g * TM_LC;          // W' of lane 0; lane j has W' = top - j
    return top >= 0 && top - (TM_LC - 1) < P.dW;
}

// One warp slot of team member t: inbox / outbox selection, loads, update, packet.
// base = sweep serial << TM_LEVEL_BITS; a packet of level L carries tag base | (L + 1).
template <int SA, int SW, int SC, bool OOP, bool CMP>
EIK_HD void tm_slot(const Plan2 &P, const TeamCfg &T, const int t, const int a0, const int a1, const int lane,
                    const int lam, const int Ap, const int g, const double *rd, double *wr,
                    const double *__restrict__ fl, const double *cmp, const double h, double &err, tm_u64 *mbox,
                    const unsigned base) {
    // mbox: the team's inboxes, inbox of member t at mbox + t * 2 * mbStride
    const tm_u64 *inbox = (t > 0 && Ap == a0) ? mbox + (long long)t * 2 * T.mbStride : nullptr;
    tm_u64 *outbox = (t < T.nC - 1 && Ap == a1 - 1) ? mbox + (long long)(t + 1) * 2 * T.mbStride : nullptr;
    V2Vals V;
    int mb;
    tm_load<SA, SW, SC, OOP, CMP>(P, lane, lam, Ap, g, rd, wr, fl, cmp, inbox, base | (unsigned)lam, V, mb);
    tm_finish<OOP, CMP>(V, wr, h, err, outbox, mb, base | (unsigned)(lam + 1));
}

#if defined(__CUDACC__)

__device__ __forceinline__ unsigned tm_ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Barrier over the nC CTAs of a team.  ctr only grows; epoch is the value it reaches when every CTA has
// arrived (tracked identically by every CTA).
__device__ __forceinline__ void tm_barrier(unsigned *ctr, unsigned &epoch, const int nC) {
    epoch += (unsigned)nC;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (tm_ld_acquire(ctr) < epoch) {}
        __threadfence();
    }
    __syncthreads();
}

// One sweep of team member t.
template <int SA, int SW, int SC, bool OOP, bool CMP>
__device__ __forceinline__ void tm_sweep(const Plan2 &P, const TeamCfg &T, const int t, const double *rd, double *wr,
                                         const double *__restrict__ fl, const double *cmp, const double h,
                                         double &err, tm_u64 *mbox, const unsigned base) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int a0, a1, lam0, lam1;
    tm_rows(P, T, t, a0, a1, lam0, lam1);
    const int nslot = (a1 - a0) * T.G32;
    for (int lam = lam0; lam <= lam1; lam++) {
        for (int q = warp; q < nslot; q += nw) {
            const int r = q / T.G32, g = q - r * T.G32;
            const int Ap = a0 + r;
            if (!tm_slot_live(P, lam, Ap, g)) continue;
            tm_slot<SA, SW, SC, OOP, CMP>(P, T, t, a0, a1, lane, lam, Ap, g, rd, wr, fl, cmp, h, err, mbox, base);
        }
        __syncthreads();
    }
}

// bufs: S x 3 x M doubles as in k_fwd3d_v2 (buffer 0 of every source: u0 in layout P; every slot that is
// not a grid node: +inf in all three buffers).  grid = S x nC CTAs, ALL co-resident (cooperative launch).
// sync: S x TM_SYNC_WORDS unsigned words, zero on entry.  mbox: S x nC inboxes of mbStride packets; no tag in
// it is >= (serial0 + 1) << TM_LEVEL_BITS (the host hands out serial ranges and clears the mailbox on wrap).
template <int NTMAX, int MINB>
__global__ void __launch_bounds__(NTMAX, MINB) k_fwd3d_team(const Plan2 P, const TeamCfg T, double *bufs,
                                                            const double *__restrict__ fP, const double *__restrict__ fM,
                                                            const double h, const double tol, const int max_rounds,
                                                            int *__restrict__ rounds, double *__restrict__ errs,
                                                            int *__restrict__ where, unsigned *sync, tm_u64 *mbox_all,
                                                            const unsigned serial0) {
    extern __shared__ double plane[];
    __shared__ double red[32];
    const int src = blockIdx.x / T.nC, t = blockIdx.x - src * T.nC;
    unsigned *sy = sync + (long long)src * TM_SYNC_WORDS;
    unsigned *ctr = sy;
    unsigned long long *errslot = (unsigned long long *)(sy + 2);
    tm_u64 *mbox = mbox_all + (long long)src * T.nC * 2 * T.mbStride;
    unsigned epoch = 0, serial = serial0;
    double *B3 = bufs + (long long)src * 3 * P.M;
    double *Bz = B3 + 2 * P.M;
    const int A0 = t * T.R, A1 = A0 + T.R < P.dA ? A0 + T.R : P.dA;   // slabs this CTA re-skews
    int o = 0, a = 1, r = 0;
    bool conv = false;
    while (r < max_rounds) {
        double err = 0.0;
        double *Bo = B3 + o * P.M, *Ba = B3 + a * P.M;
        int state = 1;
        double *w = Ba;
        for (int sw = 0; sw < 8; sw++) {
            const int sigma = P.sg[sw][1] * P.sg[sw][2];
            if (sw > 0 && sigma != state) {
                double *dst = state > 0 ? Bz : Ba;
                tm_barrier(ctr, epoch, T.nC);         // the sweep that wrote w is complete everywhere
                v2_reskew(P, w, dst, state, plane, A0, A1);
                w = dst;
                state = sigma;
            }
            tm_barrier(ctr, epoch, T.nC);             // previous sweep / re-skew complete everywhere
            serial++;
            const unsigned base = serial << TM_LEVEL_BITS;
#define TM_CALL(a_, w_, c_, oop_, cmp_) \
    tm_sweep<a_, w_, c_, oop_, cmp_>(P, T, t, oop_ ? Bo : w, w, sigma > 0 ? fP : fM, Bo, h, err, mbox, base)
            V2_DISPATCH(P, sw, TM_CALL);
#undef TM_CALL
        }
        const double eb = v2_block_max(err, red);
        if (threadIdx.x == 0) atomicMax(errslot + (r & 1), (unsigned long long)__double_as_longlong(eb));   // eb >= 0
        tm_barrier(ctr, epoch, T.nC);
        const double e = __longlong_as_double((long long)__ldcg(errslot + (r & 1)));
        if (t == 0 && threadIdx.x == 0) {
            errslot[(r + 1) & 1] = 0ULL;              // next round's slot: last read one round (>= 9 barriers) ago
            if (errs) errs[(long long)src * max_rounds + r] = e;
        }
        r++;
        const int oo = o; o = a; a = oo;
        if (e < tol) { conv = true; break; }          // e is team-uniform
    }
    if (t == 0 && threadIdx.x == 0) {
        if (rounds) rounds[src] = conv ? r : -r;
        where[src] = o;
    }
}

__global__ void k2_identity(int *__restrict__ p, const int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

#endif  // __CUDACC__

// Plan for the team kernel: 1 x 32 lane patch, any number of column groups.
inline bool team_build_plan(Plan2 &P, int m, int n, int l, int nwarps, size_t plane_bytes) {
    if (!v2_build_plan(P, m, n, l, nwarps, plane_bytes, 1, TM_LC, 1 << 30)) return false;
    return P.nlev < (1 << TM_LEVEL_BITS) - 1;
}

// Team shape for S sources on a device that can hold max_ctas CTAs at once.  R rows per CTA: as few as the
// CTA budget allows, but enough that a level offers every warp a slot (R x G32 >= nwarps) unless that
// would leave fewer than 8 CTAs per source.  Rforce > 0 overrides (tuning aid).
inline bool team_config(const Plan2 &P, int S, int max_ctas, int nwarps, int Rforce, TeamCfg &T) {
    const int budget = max_ctas / S;
    if (budget < 1) return false;
    T.G32 = (P.dC + TM_LC - 1) / TM_LC;
    int R = (P.dA + budget - 1) / budget;
    const int Rfill = (nwarps + T.G32 - 1) / T.G32;
    if (R < Rfill && (P.dA + Rfill - 1) / Rfill >= 8) R = Rfill;
    if (Rforce > 0 && (P.dA + Rforce - 1) / Rforce <= budget) R = Rforce;
    T.R = R;
    T.nC = (P.dA + R - 1) / R;
    T.mbStride = (long long)P.RS * P.PC;
    return true;
}

}  // namespace adtomo
