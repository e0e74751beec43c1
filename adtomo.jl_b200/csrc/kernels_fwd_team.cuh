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
// Between sweeps the pipeline is NOT drained: every CTA owns fixed physical rows (its rank in the pipeline is
// p or nC-1-p, by the sign of the sweep along A), re-skews its own slabs when the next sweep needs the other
// layout, publishes "step k done" (fence + one word) and starts sweep k+1 as soon as both physical neighbours
// have published step k: their rows (read as old downwind values) are final and re-skewed, and they no longer
// read the mailbox a neighbour is about to refill.  Neighbours are therefore at most one sweep apart, which
// is also why a re-skew may overwrite the buffer of the layout before last.  Sweeps that keep their direction
// along A (4 of 8 in the reference order, 6 of 8 when A is the grid's k axis) follow each other through the
// pipeline back to back.  ONE team-wide barrier per round (a counter in global memory, with fences) remains:
// the stopping test needs the maximum over all CTAs.
//
// Lanes of a warp cover 32 consecutive columns C of ONE row A (a warp slot): at a fixed level these are
// 32 consecutive doubles of one skewed row mu -- a 256-byte contiguous run for the node and for each of
// its neighbours (512 bytes for its mailbox packets).  A warp keeps its slots for the whole sweep, so
//   * the NEW (upwind, level lam-1) values never come from global memory: every slot writes its results
//     into a shared-memory sheet (R rows x dC columns, two sheets alternating by level parity, +inf where
//     there is no node) and reads its three upwind neighbours from the other sheet;
//   * the OLD values (own, downwind, slowness, round-start value) of level lam+1 are prefetched into L1
//     while level lam is computed, and read through L1: nothing in the sweep writes them before level
//     lam+1 (own) or lam+2 (downwind), the downstream CTA cannot overwrite its first row before it has our
//     packet, which we send after reading it, no line is read again after another CTA may have written it,
//     and every team barrier invalidates L1 (its __threadfence is MEMBAR.SC.GPU + CCTL.IVALL).  (Holding
//     them in registers instead spilled at 64 registers per thread.)
// The level's critical path is then L1 / sheet read -> solve -> store -> barrier (the first version read
// everything back from L2 / DRAM: 1.9 us per level at 256^3, 53 % of the warp time in the level barrier).
#pragma once
#include <cstring>
#include "kernels_fwd_v3.cuh"

namespace adtomo {

constexpr int TM_LC = 32;          // columns per warp slot
constexpr int TM_LEVEL_BITS = 12;  // packet tag = (sweep serial << 12) | (level + 1); nlev < 4095
constexpr int TM_SYNC_HDR = 8;     // per-source sync area (unsigned words): [0] barrier counter, [2..3],[4..5] err slots, [8..8+nC) steps done

struct TeamCfg {
    int nC;              // CTAs per source
    int R;               // rows (A') per CTA
    int G32;             // column groups of 32 per row
    long long mbStride;  // packets (16 bytes each) of one CTA's inbox = RS * PC (one slab of the skewed layout)
    int SP;              // pitch of a sheet row = 32 * G32 + 2 doubles
    int stride;          // unsigned words of one source's sync area
};

#define TM_LDU(p) (*(p))           // old values: through L1 (see above)
#if defined(__CUDA_ARCH__)
#define TM_PREFETCH(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#else
#define TM_PREFETCH(p) ((void)(p))
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

// Team member p owns the PHYSICAL rows A in [p R, min((p+1) R, dA)).  In a sweep with sign SA along A these are
// the rows A' in [a0, a1) counted in the sweep's direction; lam0 / lam1: first and last level at which one of
// them has a node.  The member's upstream neighbour (whose packets it receives) is p-1 for SA > 0, p+1 otherwise.
EIK_HD void tm_rows(const Plan2 &P, const TeamCfg &T, const int p, const int SA, int &a0, int &a1, int &lam0, int &lam1) {
    const int lo = p * T.R, hi = lo + T.R < P.dA ? lo + T.R : P.dA;
    a0 = SA > 0 ? lo : P.dA - hi;
    a1 = SA > 0 ? hi : P.dA - lo;
    lam0 = a0;
    lam1 = a1 - 1 + P.dW - 1 + P.dC - 1;
}

// Per-lane constants of one warp slot (row r of the CTA = row Ap of the sweep, column group g) for a whole
// sweep; coordinates with a prime are counted in the sweep's direction.  The lane's node at level lam is
// (Ap, W' = lam - wsum, C' = Cp); its slot in the field buffers is offc + lam * (SW * PC) -- affine in the level
// (same layout arithmetic as v2_load) -- and its slot inside its slab (= mailbox slot) is that minus slab.
struct TmSlotC {
    int wsum;    // Ap + Cp
    int offc;    // field slot at level 0
    int slab;    // (A + 1) * RS * PC
    int sidx;    // sheet index r * SP + Cp + 1
    int flags;
};
enum { TM_COL = 1, TM_FIRST = 2, TM_LAST = 4, TM_RPOS = 8 };   // column exists; uA from the mailbox; packet to send; uA from the sheet

template <int SA, int SW, int SC>
EIK_HD void tm_slot_setup(const Plan2 &P, const TeamCfg &T, const int p, const int a0, const int nrow, const int lane,
                          const int q, TmSlotC &K) {
    const bool has_up = SA > 0 ? p > 0 : p < T.nC - 1, has_down = SA > 0 ? p < T.nC - 1 : p > 0;
    const int r = q / T.G32, g = q - r * T.G32;
    const int Ap = a0 + r, Cp = g * TM_LC + lane;
    const int A = SA > 0 ? Ap : P.dA - 1 - Ap;
    const int C = SC > 0 ? Cp : P.dC - 1 - Cp;
    K.wsum = Ap + Cp;
    K.slab = (A + 1) * P.RS * P.PC;
    // mu + 1 = lam - Ap + 1 (SW > 0) or nmu - lam + Ap (SW < 0): one skewed row per level, uniform over the warp
    K.offc = K.slab + (SW > 0 ? 1 - Ap : P.nmu + Ap) * P.PC + C;
    K.sidx = r * T.SP + Cp + 1;
    K.flags = (Cp < P.dC ? TM_COL : 0) | ((has_up && r == 0) ? TM_FIRST : 0) |
              ((has_down && r == nrow - 1) ? TM_LAST : 0) | (r > 0 ? TM_RPOS : 0);
}

// does the lane have a node at level lam?
EIK_HD bool tm_act(const Plan2 &P, const TmSlotC &K, const int lam) {
    return (unsigned)(lam - K.wsum) < (unsigned)P.dW && (K.flags & TM_COL);
}
// warp-uniform: does the slot have a node at level lam?  (lane j has W' = top - j)
EIK_HD bool tm_slot_live(const Plan2 &P, const TmSlotC &K, const int lane, const int lam) {
    const int top = lam - (K.wsum - lane);
    return top >= 0 && top - (TM_LC - 1) < P.dW;
}
template <int SW>
EIK_HD int tm_off(const Plan2 &P, const TmSlotC &K, const int lam) { return K.offc + lam * (SW * P.PC); }

// OLD values of the node: nothing in the sweep writes them before level lam (own) / lam + 1 (downwind).
struct TmOld {
    double own, fv, dA, dW, dC;
    double ref;   // CMP sweeps: the round-start value of the node
};

template <int SA, int SW, int SC, bool CMP>
EIK_HD void tm_load_old(const Plan2 &P, const TmSlotC &K, const int lam, const double *rd,
                        const double *__restrict__ fl, const double *cmp, TmOld &O) {
    const int offA = SA * P.RS * P.PC, offW = SW * P.PC, offC = SW * P.PC + SC;   // downwind (old, level+1)
    // a lane without a node loads from a harmless slot (A = 0, mu = 0: all neighbour slots exist)
    const int off = tm_act(P, K, lam) ? tm_off<SW>(P, K, lam) : (P.RS + 1) * P.PC + 1;
    const double *p = rd + off;
    O.own = TM_LDU(p);
    O.fv = fl[off];
    O.dA = TM_LDU(p + offA);
    O.dW = TM_LDU(p + offW);
    O.dC = TM_LDU(p + offC);
    O.ref = CMP ? TM_LDU(cmp + off) : 0.0;
}

// L1 prefetch of what tm_load_old will read TM_PF levels ahead and no earlier level touches: the row of level
// lam + TM_PF + 1 of this slab (dW, dC), the A-neighbour's row, slowness, round-start value.
// Measured at 256^3: no prefetch 18.8 ms, distance 1 or 3 16.7 ms (one level ahead already covers the miss).
#ifndef TM_PF
#define TM_PF 1
#endif
template <int SA, int SW, int SC, bool CMP>
EIK_HD void tm_prefetch_old(const Plan2 &P, const TmSlotC &K, const int lam, const double *rd,
                            const double *__restrict__ fl, const double *cmp) {
    if (TM_PF <= 0 || !tm_act(P, K, lam + TM_PF)) return;
    const int offA = SA * P.RS * P.PC, offW = SW * P.PC;
    const int off = tm_off<SW>(P, K, lam + TM_PF);
    TM_PREFETCH(rd + off + offW);
    TM_PREFETCH(rd + off + offA);
    TM_PREFETCH(fl + off);
    if (CMP) TM_PREFETCH(cmp + off);
}

// One warp slot at level lam, for one lane, in two steps (loads + ordering, then arithmetic + stores).
// base = sweep serial << TM_LEVEL_BITS; a packet of level L carries tag base | (L + 1).
// sheets: [2][R][SP] doubles, sheet (lam & 1) receives this level, the other one holds level lam-1:
// column Cp at index Cp+1, index 0 = +inf.  The A neighbour of the CTA's first row comes from the mailbox
// (inbox = mbox + p * 2 * mbStride, packets tagged base | lam) or is +inf (the grid's first row).
// Host build: a packet that has not arrived yields NaN (the emulation's scheduler must prevent that).
struct TmPrep {
    double a1, a2, a3, own, fv, ref;
    int off;      // field slot, < 0: no node
};

template <int SA, int SW, int SC>
EIK_HD void tm_prep(const Plan2 &P, const TeamCfg &T, const TmSlotC &K, const int lam, const TmOld &O,
                    const tm_u64 *inbox, const unsigned base, const double *sheets, TmPrep &Q) {
    const bool act = tm_act(P, K, lam);
    const int off = tm_off<SW>(P, K, lam);
    Q.off = act ? off : -1;
    const double *row = sheets + ((lam & 1) ^ 1) * T.R * T.SP + K.sidx;
    const double uC = row[-1], uW = row[0];
    double uA;
    if (K.flags & TM_RPOS) {
        uA = row[-T.SP];
    } else if (K.flags & TM_FIRST) {
        uA = 0.0;
        if (act) {
            // the A-neighbour has the same (W, C), i.e. the same slot inside ITS slab
            const tm_u64 *pk = inbox + 2 * (long long)(off - K.slab);
            const unsigned tag = base | (unsigned)lam;
            tm_u64 p0, p1;
#if defined(__CUDA_ARCH__)
            do { tm_mb_load(pk, p0, p1); } while (!tm_unpack(p0, p1, tag, uA));
#else
            tm_mb_load(pk, p0, p1);
            if (!tm_unpack(p0, p1, tag, uA)) uA = NAN;
#endif
        }
    } else {
        uA = v2_inf();
    }
    Q.own = O.own;
    Q.fv = O.fv;
    Q.ref = O.ref;
    Q.a1 = eik_min(uA, O.dA);
    Q.a2 = eik_min(uW, O.dW);
    Q.a3 = eik_min(uC, O.dC);
    eik_sort3(Q.a1, Q.a2, Q.a3);
}

// The update (Eikonal3D.cpp:47-54), its store, its sheet entry (+inf when the lane has no node) and the packet
// for the downstream CTA (outbox = the inbox of member p + 1 or p - 1; written for EVERY node of the CTA's last row,
// changed or not: the consumer waits for it).
template <bool OOP, bool CMP>
EIK_HD void tm_solve(const TeamCfg &T, const TmSlotC &K, const int lam, const TmPrep &Q, double *wr, const double h,
                     double &err, tm_u64 *outbox, const unsigned base, double *sheets) {
    double res = v2_inf();
    if (Q.off >= 0) {
        res = Q.own;
        bool changed = false;
        if (Q.a1 < Q.own) {   // otherwise the candidate (> a1) cannot win the min: exact skip
            const double un = eik_solve3_sorted(Q.a1, Q.a2, Q.a3, Q.fv * h, Q.fv * Q.fv * h * h);
            if (un < Q.own) { res = un; changed = true; }
        }
        if (OOP || changed) wr[Q.off] = res;
        if (K.flags & TM_LAST) {
            tm_u64 p0, p1;
            tm_pack(res, base | (unsigned)(lam + 1), p0, p1);
            tm_mb_store(outbox + 2 * (long long)(Q.off - K.slab), p0, p1);
        }
        if (CMP) {
            const double dd = fabs(res - Q.ref);
            err = (err < dd) ? dd : err;
        }
    }
    sheets[(lam & 1) * T.R * T.SP + K.sidx] = res;
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

// "Steps 1..step of this CTA are done": its rows are final (and re-skewed if the next sweep needs it).
__device__ __forceinline__ void tm_publish(unsigned *done_p, const unsigned step) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(done_p), "r"(step) : "memory");
    }
}
// Wait until both physical neighbours have published `step`.  The acquire loads also invalidate L1
// (LDG.STRONG.GPU + CCTL.IVALL), which the L1 reads of the sweep rely on.
__device__ __forceinline__ void tm_wait_neighbours(const unsigned *done, const int p, const int nC, const unsigned step) {
    if (threadIdx.x == 0) {
        if (p > 0) while (tm_ld_acquire(done + p - 1) < step) {}
        if (p < nC - 1) while (tm_ld_acquire(done + p + 1) < step) {}
    }
    __syncthreads();
}

// One sweep of team member t.  A warp owns slots q = warp, warp + nw, ... for the whole sweep; the constants of
// its first KS slots stay in registers, further slots (grids wider than KS x 512 columns per row) recompute them.
template <int SA, int SW, int SC, bool OOP, bool CMP, int KS>
__device__ __forceinline__ void tm_sweep(const Plan2 &P, const TeamCfg &T, const int t, const double *rd, double *wr,
                                         const double *__restrict__ fl, const double *cmp, const double h,
                                         double &err, tm_u64 *mbox, const unsigned base, double *sheets,
                                         tm_u64 *volatile *s_box) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int a0, a1, lam0, lam1;
    tm_rows(P, T, t, SA, a0, a1, lam0, lam1);
    const int nrow = a1 - a0, nslot = nrow * T.G32;
    // The two mailbox base pointers are per-sweep constants, but at 64 registers the compiler re-derived them
    // (64-bit multiplies, ~30 instructions each) on every level: park them in shared memory, one LDS.64 per use.
    if (threadIdx.x == 0) {
        s_box[0] = mbox + (long long)t * 2 * T.mbStride;
        s_box[1] = mbox + (long long)(SA > 0 ? t + 1 : t - 1) * 2 * T.mbStride;
    }
    for (int i = threadIdx.x; i < 2 * T.R * T.SP; i += blockDim.x) sheets[i] = v2_inf();
    TmSlotC K[KS];
#pragma unroll
    for (int k = 0; k < KS; k++)
        if (warp + k * nw < nslot) tm_slot_setup<SA, SW, SC>(P, T, t, a0, nrow, lane, warp + k * nw, K[k]);
    __syncthreads();
#define TM_STEP(K_)                                                                                   \
    do {                                                                                              \
        if (!tm_slot_live(P, K_, lane, lam)) {      /* no node: the sheet still needs its +inf */    \
            sheets[(lam & 1) * T.R * T.SP + (K_).sidx] = v2_inf();                                    \
        } else {                                                                                      \
            TmOld O;                                                                                  \
            TmPrep Q;                                                                                 \
            tm_prefetch_old<SA, SW, SC, CMP>(P, K_, lam, rd, fl, cmp);                                \
            tm_load_old<SA, SW, SC, CMP>(P, K_, lam, rd, fl, cmp, O);                                 \
            tm_prep<SA, SW, SC>(P, T, K_, lam, O, s_box[0], base, sheets, Q);                         \
            tm_solve<OOP, CMP>(T, K_, lam, Q, wr, h, err, s_box[1], base, sheets);                    \
        }                                                                                             \
    } while (0)
    for (int lam = lam0; lam <= lam1; lam++) {
#pragma unroll
        for (int k = 0; k < KS; k++)
            if (warp + k * nw < nslot) TM_STEP(K[k]);
        for (int q = warp + KS * nw; q < nslot; q += nw) {
            TmSlotC G;
            tm_slot_setup<SA, SW, SC>(P, T, t, a0, nrow, lane, q, G);
            TM_STEP(G);
        }
        __syncthreads();
    }
#undef TM_STEP
}

// bufs: S x 3 x M doubles as in k_fwd3d_v2 (buffer 0 of every source: u0 in layout P; every slot that is
// not a grid node: +inf in all three buffers).  grid = S x nC CTAs, ALL co-resident (cooperative launch).
// sync: S x T.stride unsigned words, zero on entry.  mbox: S x nC inboxes of mbStride packets; no tag in
// it is >= (serial0 + 1) << TM_LEVEL_BITS (the host hands out serial ranges and clears the mailbox on wrap).
// Dynamic shared memory: max(re-skew plane, two sheets of R x SP doubles); they are never live together.
template <int NTMAX, int MINB, int KS>
__global__ void __launch_bounds__(NTMAX, MINB) k_fwd3d_team(const Plan2 P, const TeamCfg T, double *bufs,
                                                            const double *__restrict__ fP, const double *__restrict__ fM,
                                                            const double h, const double tol, const int max_rounds,
                                                            int *__restrict__ rounds, double *__restrict__ errs,
                                                            int *__restrict__ where, unsigned *sync, tm_u64 *mbox_all,
                                                            const unsigned serial0) {
    extern __shared__ double plane[];
    __shared__ double red[32];
    __shared__ tm_u64 *volatile s_box[2];
    const int src = blockIdx.x / T.nC, t = blockIdx.x - src * T.nC;
    unsigned *sy = sync + (long long)src * T.stride;
    unsigned *ctr = sy, *done = sy + TM_SYNC_HDR;
    unsigned long long *errslot = (unsigned long long *)(sy + 2);
    tm_u64 *mbox = mbox_all + (long long)src * T.nC * 2 * T.mbStride;
    unsigned epoch = 0, serial = serial0, step = 0;
    const V3Pol pol = v3_policies();
    double *B3 = bufs + (long long)src * 3 * P.M;
    double *Bz = B3 + 2 * P.M;
    const int A0 = t * T.R, A1 = A0 + T.R < P.dA ? A0 + T.R : P.dA;   // the slabs this CTA owns (and re-skews)
    int o = 0, a = 1, r = 0;
    bool conv = false;
    while (r < max_rounds) {
        double err = 0.0;
        double *Bo = B3 + o * P.M, *Ba = B3 + a * P.M;
        int state = 1;                                // layout of the working field: sweep 0 is always on P
        double *w = Ba;
        for (int sw = 0; sw < 8; sw++) {
            const int sigma = P.sg[sw][1] * P.sg[sw][2];
            tm_wait_neighbours(done, t, T.nC, step);  // both neighbours finished the previous sweep (and its re-skew)
            serial++;
            const unsigned base = serial << TM_LEVEL_BITS;
#define TM_CALL(a_, w_, c_, oop_, cmp_) \
    tm_sweep<a_, w_, c_, oop_, cmp_, KS>(P, T, t, oop_ ? Bo : w, w, sigma > 0 ? fP : fM, Bo, h, err, mbox, base, plane, s_box)
            V2_DISPATCH(P, sw, TM_CALL);
#undef TM_CALL
            if (sw < 7) {
                const int next = P.sg[sw + 1][1] * P.sg[sw + 1][2];
                if (next != state) {                  // the next sweep runs on the other layout: re-skew my slabs
                    double *dst = state > 0 ? Bz : Ba;
                    v3_reskew<0>(P, w, dst, state, plane, A0, A1, pol);      // run-time pitch; thread-owned columns (kernels_fwd_v3.cuh)
                    w = dst;
                    state = next;
                }
            }
            tm_publish(done + t, ++step);
        }
        const double eb = v2_block_max(err, red);
        if (threadIdx.x == 0) atomicMax(errslot + (r & 1), (unsigned long long)__double_as_longlong(eb));   // eb >= 0
        tm_barrier(ctr, epoch, T.nC);
        const double e = __longlong_as_double((long long)__ldcg(errslot + (r & 1)));
        if (t == 0 && threadIdx.x == 0) {
            errslot[(r + 1) & 1] = 0ULL;              // next round's slot: last read one round (>= 1 barrier) ago
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
// CTA budget allows -- a level is latency-bound (one ~250-instruction update per lane), so more CTAs with fewer
// slots each win at every size measured (64^3: R = 1 3.3 ms, R = 8 4.2 ms; 256^3: 19.5 / 19.9 / 35.3 ms for
// R = 1 / 2 / 4).  Rforce > 0 overrides (tuning aid).
inline bool team_config(const Plan2 &P, int S, int max_ctas, int nwarps, int Rforce, TeamCfg &T) {
    const int budget = max_ctas / S;
    if (budget < 1) return false;
    T.G32 = (P.dC + TM_LC - 1) / TM_LC;
    int R = (P.dA + budget - 1) / budget;
    (void)nwarps;
    if (Rforce > 0 && (P.dA + Rforce - 1) / Rforce <= budget) R = Rforce;
    T.R = R;
    T.nC = (P.dA + R - 1) / R;
    T.mbStride = (long long)P.RS * P.PC;
    T.SP = TM_LC * T.G32 + 2;
    T.stride = TM_SYNC_HDR + ((T.nC + 1) & ~1);
    return true;
}

}  // namespace adtomo
