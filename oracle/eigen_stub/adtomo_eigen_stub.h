// oracle/eigen_stub/adtomo_eigen_stub.h -- TEST INFRASTRUCTURE, ours (not Eigen, not reference code).
//
// The reference's Eikonal hot path (deps/CustomOps/Eikonal/Eikonal.h, deps/CustomOps/Eikonal3D/Eikonal3D.{h,cpp})
// includes Eigen, which is an un-vendored, unpinned third-party dependency that is absent from this image
// (SURVEY 8c).  Its forward solvers do not use Eigen at all (2D: two vector norms), its adjoints use it for
// "assemble triplets -> sparse matrix -> transpose -> SparseLU -> solve".  This header declares exactly the
// handful of Eigen names those files touch, with the documented semantics (setFromTriplets SUMS duplicate
// entries; SparseLU::solve returns the solution of A x = b), so that the reference's own sources compile
// UNMODIFIED, from where they lie under /root/reference, into oracle/_ref/ (recipe: oracle/Makefile).
// The linear solve is a dense LU with partial pivoting: exact enough (<= 1e-13 relative on these permuted
// triangular systems) but O(n^3), so the compiled reference adjoint is only usable for n <= ~4000 unknowns; the
// forward solvers are the reference's code line for line at any size.
// Differences from real Eigen that can matter: VectorXd::norm() sums sequentially (Eigen vectorises the
// reduction), so the 2D stopping test `err < 1e-8` could in principle flip on a value within ~1e-16 relative of the
// threshold; the LU's rounding differs from SparseLU's.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <utility>
#include <vector>

namespace Eigen {

template <typename Derived>
struct MapBase_ {};

class VectorXd {
public:
    VectorXd() {}
    explicit VectorXd(long n) : v_((size_t)n, 0.0) {}
    double *data() { return v_.data(); }
    const double *data() const { return v_.data(); }
    long size() const { return (long)v_.size(); }
    double &operator[](long i) { return v_[(size_t)i]; }
    const double &operator[](long i) const { return v_[(size_t)i]; }
    double &operator()(long i) { return v_[(size_t)i]; }
    const double &operator()(long i) const { return v_[(size_t)i]; }
    double norm() const {
        double s = 0.0;
        for (double x : v_) s += x * x;
        return std::sqrt(s);
    }
    VectorXd operator-(const VectorXd &o) const {
        VectorXd r((long)v_.size());
        for (size_t i = 0; i < v_.size(); i++) r.v_[i] = v_[i] - o.v_[i];
        return r;
    }
    std::vector<double> v_;
};

template <typename T>
class Map;
template <>
class Map<const VectorXd> {
public:
    Map(const double *p, long n) : p_(p), n_(n) {}
    operator VectorXd() const {
        VectorXd r(n_);
        for (long i = 0; i < n_; i++) r[i] = p_[i];
        return r;
    }
private:
    const double *p_;
    long n_;
};

template <typename Scalar>
class Triplet {
public:
    Triplet() : r_(0), c_(0), v_(0) {}
    Triplet(int r, int c, Scalar v) : r_(r), c_(c), v_(v) {}
    int row() const { return r_; }
    int col() const { return c_; }
    Scalar value() const { return v_; }
private:
    int r_, c_;
    Scalar v_;
};

template <typename Scalar>
class SparseMatrix {
public:
    SparseMatrix() : rows_(0), cols_(0) {}
    SparseMatrix(long r, long c) : rows_(r), cols_(c) {}
    template <typename It>
    void setFromTriplets(It b, It e) {
        a_.clear();
        for (It t = b; t != e; ++t) a_[std::make_pair((long)t->row(), (long)t->col())] += t->value();   // duplicates are summed
    }
    SparseMatrix transpose() const {
        SparseMatrix r(cols_, rows_);
        for (const auto &kv : a_) r.a_[std::make_pair(kv.first.second, kv.first.first)] = kv.second;
        return r;
    }
    long rows() const { return rows_; }
    long cols() const { return cols_; }
    std::map<std::pair<long, long>, Scalar> a_;
    long rows_, cols_;
};

template <typename Mat>
class SparseLU {
public:
    void analyzePattern(const Mat &) {}
    void factorize(const Mat &A) {
        n_ = A.rows();
        if (n_ != A.cols() || n_ > 6000) {
            std::fprintf(stderr, "adtomo_eigen_stub: dense LU limited to square systems with <= 6000 unknowns (got %ld x %ld)\n",
                         A.rows(), A.cols());
            std::abort();
        }
        lu_.assign((size_t)n_ * n_, 0.0);
        for (const auto &kv : A.a_) lu_[(size_t)kv.first.first * n_ + kv.first.second] = kv.second;
        piv_.resize((size_t)n_);
        for (long k = 0; k < n_; k++) {
            long p = k;
            double best = std::fabs(lu_[(size_t)k * n_ + k]);
            for (long i = k + 1; i < n_; i++) {
                const double a = std::fabs(lu_[(size_t)i * n_ + k]);
                if (a > best) { best = a; p = i; }
            }
            piv_[(size_t)k] = p;
            if (p != k)
                for (long j = 0; j < n_; j++) std::swap(lu_[(size_t)k * n_ + j], lu_[(size_t)p * n_ + j]);
            const double d = lu_[(size_t)k * n_ + k];
            if (d == 0.0) continue;               // singular column: like the reference, no guard (result will be inf/nan)
            for (long i = k + 1; i < n_; i++) {
                double &lik = lu_[(size_t)i * n_ + k];
                if (lik == 0.0) continue;
                lik /= d;
                const double m = lik;
                const double *rk = &lu_[(size_t)k * n_];
                double *ri = &lu_[(size_t)i * n_];
                for (long j = k + 1; j < n_; j++) ri[j] -= m * rk[j];
            }
        }
    }
    VectorXd solve(const VectorXd &b) const {
        VectorXd x = b;
        for (long k = 0; k < n_; k++) {
            if (piv_[(size_t)k] != k) std::swap(x[k], x[piv_[(size_t)k]]);
            for (long i = k + 1; i < n_; i++) x[i] -= lu_[(size_t)i * n_ + k] * x[k];
        }
        for (long k = n_ - 1; k >= 0; k--) {
            double s = x[k];
            for (long j = k + 1; j < n_; j++) s -= lu_[(size_t)k * n_ + j] * x[j];
            x[k] = s / lu_[(size_t)k * n_ + k];
        }
        return x;
    }
private:
    long n_ = 0;
    std::vector<double> lu_;
    std::vector<long> piv_;
};

}  // namespace Eigen
