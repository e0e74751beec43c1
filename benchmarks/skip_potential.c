// benchmarks/skip_potential.c -- MEASUREMENT AID (CPU, not product code).
// Question: how many warp-slot evaluations of the batch forward kernel could be skipped
// BIT-EXACTLY by tracking, per box of BA x BW x BC nodes, the serial of the last sweep in which a node of the box
// changed, and evaluating a warp slot (4 A x 8 C pencils at one level) only when a box it touches or one of their face
// neighbours changed in the previous or the current sweep?  The emulation runs the reference's sweeps level by level
// (Eikonal3D.cpp:35-88) with exactly that rule, checks the result and the round count against the unskipped solve,
// and counts evaluated vs live slots and the nodes whose value changed.  Driver: benchmarks/skip_potential.py.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static double sol(double a1, double a2, double a3, double f, double h) {
    double t;
    if (a1 > a2) { t = a1; a1 = a2; a2 = t; }
    if (a1 > a3) { t = a1; a1 = a3; a3 = t; }
    if (a2 > a3) { t = a2; a2 = a3; a3 = t; }
    double x = a1 + f * h;
    if (x <= a2) return x;
    double B = -(a1 + a2);
    double C = (a1 * a1 + a2 * a2 - f * f * h * h) / 2.0;
    x = (-B + sqrt(B * B - 4 * C)) / 2.0;
    if (x <= a3) return x;
    B = -2.0 * (a1 + a2 + a3) / 3.0;
    C = (a1 * a1 + a2 * a2 + a3 * a3 - f * f * h * h) / 3.0;
    return (-B + sqrt(B * B - 4 * C)) / 2.0;
}
static inline double dmin(double a, double b) { return b < a ? b : a; }
static const int DIRS[8][3] = {{1,1,1},{-1,1,1},{-1,-1,1},{1,-1,1},{1,-1,-1},{1,1,-1},{-1,1,-1},{-1,-1,-1}};
#define ID(i,j,k) (((size_t)(i)*n+(j))*l+(k))
static int m, n, l;
static double hh;
static inline int upd(double *u, const double *f, int i, int j, int k) {
    double ux = i == 0 ? u[ID(i+1,j,k)] : (i == m-1 ? u[ID(i-1,j,k)] : dmin(u[ID(i+1,j,k)], u[ID(i-1,j,k)]));
    double uy = j == 0 ? u[ID(i,j+1,k)] : (j == n-1 ? u[ID(i,j-1,k)] : dmin(u[ID(i,j+1,k)], u[ID(i,j-1,k)]));
    double uz = k == 0 ? u[ID(i,j,k+1)] : (k == l-1 ? u[ID(i,j,k-1)] : dmin(u[ID(i,j,k+1)], u[ID(i,j,k-1)]));
    double un = sol(ux, uy, uz, f[ID(i,j,k)], hh);
    if (un < u[ID(i,j,k)]) { u[ID(i,j,k)] = un; return 1; }
    return 0;
}
// roles: A=i, W=j, C=k.  slot (rb,g) at level lam: A' in 4rb.., C phys in 8g.., W' = lam - A' - C'
// boxes: BA x BW x BC physical
long long ev_slots, live_slots, ev_nodes_changed, copy_slots;
long long ev_by_sweep[8*32], live_by_sweep[8*32];
int skip_forward(double *u, const double *u0, const double *f, double tol, int max_rounds, int BA, int BW, int BC, int mode, int *rounds_out) {
    size_t N = (size_t)m*n*l;
    memcpy(u, u0, 8*N);
    double *uo = malloc(8*N);
    int nba = (m+BA-1)/BA, nbw = (n+BW-1)/BW, nbc = (l+BC-1)/BC;
    int *chg = malloc(sizeof(int)*nba*nbw*nbc);
    for (int q = 0; q < nba*nbw*nbc; q++) chg[q] = 0;   // serial 0: "changed at time 0" => everything active in sweep t=1 (t-1=0)
#define CH(a,w,c) chg[((a)*nbw+(w))*nbc+(c)]
    int t = 0, it = 0;
    int nrb = (m+3)/4, G = (l+7)/8, nlev = m+n+l-2;
    unsigned char *act = malloc(nrb*G);
    for (int r = 0; r < max_rounds; r++) {
        memcpy(uo, u, 8*N);
        for (int s = 0; s < 8; s++) {
            t++;
            int SA = DIRS[s][0], SW = DIRS[s][1], SC = DIRS[s][2];
            for (int lam = 0; lam < nlev; lam++) {
                // decide
                for (int rb = 0; rb < nrb; rb++) for (int g = 0; g < G; g++) {
                    // live? node ranges
                    int a0p = rb*4, a1p = rb*4+3 < m-1 ? rb*4+3 : m-1;         // A' range
                    int c0 = g*8, c1 = g*8+7 < l-1 ? g*8+7 : l-1;               // physical C
                    int cp0 = SC>0 ? c0 : l-1-c1, cp1 = SC>0 ? c1 : l-1-c0;
                    int wlo = lam - a1p - cp1, whi = lam - a0p - cp0;            // W' range
                    if (whi < 0 || wlo > n-1) { act[rb*G+g] = 0; continue; }
                    if (wlo < 0) wlo = 0; if (whi > n-1) whi = n-1;
                    act[rb*G+g] = 1;
                    live_slots++; live_by_sweep[t-1 < 256 ? t-1 : 255]++;
                    if (mode == 0) { act[rb*G+g] = 2; continue; }
                    int A0 = SA>0 ? a0p : m-1-a1p, A1 = SA>0 ? a1p : m-1-a0p;
                    int W0 = SW>0 ? wlo : n-1-whi, W1 = SW>0 ? whi : n-1-wlo;
                    int ba0 = A0/BA, ba1 = A1/BA, bw0 = W0/BW, bw1 = W1/BW, bc0 = c0/BC, bc1 = c1/BC;
                    int mx = -1;
                    if (mode == 1) {   // full 3x3x3-ish range
                        for (int a = ba0-1; a <= ba1+1; a++) for (int w = bw0-1; w <= bw1+1; w++) for (int c = bc0-1; c <= bc1+1; c++) {
                            if (a<0||a>=nba||w<0||w>=nbw||c<0||c>=nbc) continue;
                            if (CH(a,w,c) > mx) mx = CH(a,w,c);
                        }
                    } else {           // cross
                        for (int a = ba0; a <= ba1; a++) for (int w = bw0; w <= bw1; w++) for (int c = bc0; c <= bc1; c++) {
                            if (CH(a,w,c) > mx) mx = CH(a,w,c);
                        }
                        for (int w = bw0; w <= bw1; w++) for (int c = bc0; c <= bc1; c++) {
                            if (ba0>0 && CH(ba0-1,w,c) > mx) mx = CH(ba0-1,w,c);
                            if (ba1<nba-1 && CH(ba1+1,w,c) > mx) mx = CH(ba1+1,w,c);
                        }
                        for (int a = ba0; a <= ba1; a++) for (int c = bc0; c <= bc1; c++) {
                            if (bw0>0 && CH(a,bw0-1,c) > mx) mx = CH(a,bw0-1,c);
                            if (bw1<nbw-1 && CH(a,bw1+1,c) > mx) mx = CH(a,bw1+1,c);
                        }
                        for (int a = ba0; a <= ba1; a++) for (int w = bw0; w <= bw1; w++) {
                            if (bc0>0 && CH(a,w,bc0-1) > mx) mx = CH(a,w,bc0-1);
                            if (bc1<nbc-1 && CH(a,w,bc1+1) > mx) mx = CH(a,w,bc1+1);
                        }
                    }
                    if (mx >= t-1) act[rb*G+g] = 2;
                }
                // evaluate
                for (int rb = 0; rb < nrb; rb++) for (int g = 0; g < G; g++) {
                    if (act[rb*G+g] != 2) continue;
                    ev_slots++; ev_by_sweep[t-1 < 256 ? t-1 : 255]++;
                    for (int la = 0; la < 4; la++) for (int lc = 0; lc < 8; lc++) {
                        int ap = rb*4+la, c = g*8+lc;
                        if (ap >= m || c >= l) continue;
                        int cp = SC>0 ? c : l-1-c;
                        int wp = lam - ap - cp;
                        if (wp < 0 || wp >= n) continue;
                        int A = SA>0 ? ap : m-1-ap, W = SW>0 ? wp : n-1-wp;
                        if (upd(u, f, A, W, c)) { CH(A/BA, W/BW, c/BC) = t; ev_nodes_changed++; }
                    }
                }
            }
        }
        double err = 0;
        for (size_t q = 0; q < N; q++) { double d = fabs(u[q]-uo[q]); err = d > err ? d : err; }
        it = r+1;
        if (err < tol) break;
    }
    *rounds_out = it;
    free(uo); free(chg); free(act);
    return 0;
}
void set_dims(int m_, int n_, int l_, double h_) { m = m_; n = n_; l = l_; hh = h_; }
void get_counts(long long *o) { o[0] = ev_slots; o[1] = live_slots; o[2] = ev_nodes_changed; }
void get_by_sweep(long long *e, long long *lv) { memcpy(e, ev_by_sweep, sizeof(ev_by_sweep)); memcpy(lv, live_by_sweep, sizeof(live_by_sweep)); }
void reset_counts(void) { ev_slots = live_slots = ev_nodes_changed = 0; memset(ev_by_sweep,0,sizeof(ev_by_sweep)); memset(live_by_sweep,0,sizeof(live_by_sweep)); }
