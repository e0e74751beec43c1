// oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE.  C entry points around the REFERENCE's own Eikonal solvers, which are
// compiled unmodified from /root/reference/deps/CustomOps (Eikonal/Eikonal.h is header-only; Eikonal3D/Eikonal3D.cpp
// is compiled next to this file by oracle/Makefile) against oracle/eigen_stub (the third-party Eigen is absent).
// Nothing of the reference is copied: this file only includes its headers by path and forwards the calls that
// the TensorFlow shims make (Eikonal.cpp:125,217-219; EikonalThreeD.cpp:136-138,245-248).
#include "Eikonal3D/Eikonal3D.h"
#include "Eikonal/Eikonal.h"

extern "C" {

void ref_eikonal3d_forward(double *u, const double *u0, const double *f, double h, int m, int n, int l, double tol,
                           int verbose) {
    Eikonal3D::forward(u, u0, f, h, m, n, l, tol, verbose != 0);
}

void ref_eikonal3d_backward(double *grad_u0, double *grad_f, const double *grad_u, const double *u, const double *u0,
                            const double *f, double h, int m, int n, int l) {
    Eikonal3D::backward(grad_u0, grad_f, grad_u, u, u0, f, h, m, n, l);
}

void ref_eikonal2d_forward(double *u, const double *f, int m, int n, double h, int ix, int jx) {
    forward(u, f, m, n, h, ix, jx);
}

void ref_eikonal2d_backward(double *grad_f, const double *grad_u, const double *u, const double *f, int m, int n,
                            double h, int ix, int jx) {
    backward(grad_f, grad_u, u, f, m, n, h, ix, jx);
}

}
