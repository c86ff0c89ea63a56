// smallmat.hpp -- host-side small dense matrix functions for the Krylov projection step.
//
// These act on the m x m (m <= ~130) Hessenberg / tridiagonal matrix that arnoldi!/lanczos! leave
// on the host (the reference keeps Ks.H in a host Matrix and does this work with LAPACK on the
// host even for GPU vectors: src/krylov_phiv.jl:283-315).  Everything is column-major.
//
// Reference semantics followed (paths relative to the reference repository):
//   exponential!(A, ExpMethodHigham2005Base)   src/exp_baseexp.jl:112-161  (balance -> 1-norm switch
//        Pade 3/5/7/9/13 -> squaring -> unbalance; generic even/odd power loop :84-105; LU solve :44-59)
//   phiv_dense!                                src/phi.jl:84-115
//   expv! symmetric branch                     src/krylov_phiv.jl:225-229  (eigen!(SymTridiagonal))
// Balancing follows the published LAPACK xGEBAL job='B' algorithm (permute, then power-of-two
// scaling with 2-norms), which is what PureGebal.balance! restates; the LU is Gaussian elimination
// with partial pivoting (xGETRF/xGETRS semantics); the tridiagonal eigensolver is the implicit QL
// iteration (EISPACK tql2 scheme) instead of LAPACK's MRRR -- eigenpairs agree to rounding.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <limits>
#include <vector>

namespace b200k {
namespace smallmat {

using std::size_t;

struct Mat {  // owning column-major square/rectangular matrix
    int r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    double &operator()(int i, int j) { return a[(size_t)j * r + i]; }
    double operator()(int i, int j) const { return a[(size_t)j * r + i]; }
    double *col(int j) { return a.data() + (size_t)j * r; }
    const double *col(int j) const { return a.data() + (size_t)j * r; }
};

// The functions below are templates on the scalar S = double or std::complex<double> (the reference's
// exponential! covers all BlasFloat types; the complex instance serves the ComplexF64 Krylov path).
typedef std::complex<double> cplx;
inline double absval(double x) { return std::fabs(x); }
inline double absval(const cplx &x) { return std::abs(x); }
inline double abs2(double x) { return x * x; }
inline double abs2(const cplx &x) { return std::norm(x); }

// C = A * B, all n x n.  j-k-i loop order: unit stride on the inner loop (column-major).
template <typename S>
inline void matmul(int n, const S *A, const S *B, S *C) {
    std::fill(C, C + (size_t)n * n, S(0.0));
    for (int j = 0; j < n; ++j) {
        S *cj = C + (size_t)j * n;
        for (int k = 0; k < n; ++k) {
            const S bkj = B[(size_t)j * n + k];
            if (bkj == S(0.0)) continue;
            const S *ak = A + (size_t)k * n;
            for (int i = 0; i < n; ++i) cj[i] += ak[i] * bkj;
        }
    }
}

template <typename S>
inline double norm1(int n, const S *A) {
    double best = 0.0;
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += absval(A[(size_t)j * n + i]);
        if (s > best || s != s) best = s;
    }
    return best;
}

// ---- balancing (xGEBAL job 'B') ------------------------------------------------------------------
struct Balance {
    int ilo = 0, ihi = -1;       // 0-based inclusive active block
    std::vector<double> scale;   // permutation targets outside [ilo, ihi], scale factors inside
};

template <typename S>
inline void swap_rc(int n, S *A, int i, int j, int rows_hi /*swap columns over rows 0..rows_hi*/,
                    int cols_lo /*swap rows over columns cols_lo..n-1*/) {
    if (i == j) return;
    for (int r = 0; r <= rows_hi; ++r) std::swap(A[(size_t)i * n + r], A[(size_t)j * n + r]);
    for (int c = cols_lo; c < n; ++c) std::swap(A[(size_t)c * n + i], A[(size_t)c * n + j]);
}

template <typename S>
inline void balance(int n, S *A, Balance &bal) {
    bal.scale.assign(n, 1.0);
    int k = 0, l = n - 1;
    if (n == 0) { bal.ilo = 0; bal.ihi = -1; return; }
    // rows isolating an eigenvalue are pushed down
    bool noconv = true;
    while (noconv) {
        noconv = false;
        for (int i = l; i >= 0; --i) {
            bool canswap = true;
            for (int j = 0; j <= l; ++j)
                if (i != j && A[(size_t)j * n + i] != S(0.0)) { canswap = false; break; }
            if (canswap) {
                bal.scale[l] = (double)i;
                swap_rc(n, A, i, l, l, k);
                noconv = true;
                if (l == 0) { bal.ilo = 0; bal.ihi = 0; return; }
                --l;
                break;
            }
        }
    }
    // columns isolating an eigenvalue are pushed left
    noconv = true;
    while (noconv) {
        noconv = false;
        for (int j = k; j <= l; ++j) {
            bool canswap = true;
            for (int i = k; i <= l; ++i)
                if (i != j && A[(size_t)j * n + i] != S(0.0)) { canswap = false; break; }
            if (canswap) {
                bal.scale[k] = (double)j;
                swap_rc(n, A, j, k, l, k);
                noconv = true;
                ++k;
                break;
            }
        }
    }
    for (int i = k; i <= l; ++i) bal.scale[i] = 1.0;
    const double radix = 2.0, sclfac = 2.0, factor = 0.95;
    const double sfmin1 = std::numeric_limits<double>::min() / std::numeric_limits<double>::epsilon();
    const double sfmax1 = 1.0 / sfmin1;
    const double sfmin2 = sfmin1 * sclfac, sfmax2 = 1.0 / sfmin2;
    (void)radix;
    noconv = true;
    while (noconv) {
        noconv = false;
        for (int i = k; i <= l; ++i) {
            double c = 0.0, r = 0.0;
            for (int q = k; q <= l; ++q) {
                c += abs2(A[(size_t)i * n + q]);
                r += abs2(A[(size_t)q * n + i]);
            }
            c = std::sqrt(c);
            r = std::sqrt(r);
            double ca = 0.0, ra = 0.0;
            for (int q = 0; q <= l; ++q) ca = std::max(ca, absval(A[(size_t)i * n + q]));
            for (int q = k; q < n; ++q) ra = std::max(ra, absval(A[(size_t)q * n + i]));
            if (c == 0.0 || r == 0.0) continue;
            if (!(c + ca + r + ra == c + ca + r + ra)) { bal.ilo = k; bal.ihi = l; return; }  // NaN guard
            double g = r / sclfac, f = 1.0;
            const double s = c + r;
            while (c < g && std::max(f, std::max(c, ca)) < sfmax2 &&
                   std::min(r, std::min(g, ra)) > sfmin2) {
                f *= sclfac; c *= sclfac; ca *= sclfac;
                r /= sclfac; g /= sclfac; ra /= sclfac;
            }
            g = c / sclfac;
            while (g >= r && std::max(r, ra) < sfmax2 &&
                   std::min(std::min(f, c), std::min(g, ca)) > sfmin2) {
                f /= sclfac; c /= sclfac; g /= sclfac; ca /= sclfac;
                r *= sclfac; ra *= sclfac;
            }
            if (c + r >= factor * s) continue;
            if (f < 1.0 && bal.scale[i] < 1.0 && f * bal.scale[i] <= sfmin1) continue;
            if (f > 1.0 && bal.scale[i] > 1.0 && bal.scale[i] >= sfmax1 / f) continue;
            g = 1.0 / f;
            bal.scale[i] *= f;
            noconv = true;
            for (int q = k; q < n; ++q) A[(size_t)q * n + i] *= g;   // row i
            for (int q = 0; q <= l; ++q) A[(size_t)i * n + q] *= f;  // column i
        }
    }
    bal.ilo = k;
    bal.ihi = l;
}

// X <- (D P) X (D P)^{-1}: undo the scaling, then the permutations in reverse order.
template <typename S>
inline void unbalance(int n, S *X, const Balance &bal) {
    // ilo == ihi: the block is 1 x 1 and scale[ilo] holds a permutation index, not a factor (xGEBAK).
    for (int j = bal.ilo; j <= bal.ihi && bal.ilo < bal.ihi; ++j) {
        const double s = bal.scale[j];
        if (s == 1.0) continue;
        for (int q = 0; q < n; ++q) X[(size_t)q * n + j] *= s;  // row j
        for (int q = 0; q < n; ++q) X[(size_t)j * n + q] /= s;  // column j
    }
    auto rcswap = [&](int i, int j) {
        if (i == j) return;
        for (int q = 0; q < n; ++q) std::swap(X[(size_t)q * n + i], X[(size_t)q * n + j]);
        for (int q = 0; q < n; ++q) std::swap(X[(size_t)i * n + q], X[(size_t)j * n + q]);
    };
    for (int j = bal.ilo - 1; j >= 0; --j) rcswap(j, (int)bal.scale[j]);
    for (int j = bal.ihi + 1; j < n; ++j) rcswap(j, (int)bal.scale[j]);
}

// ---- LU solve: A X = B in place (B <- X).  Returns false on an exactly singular pivot. -----------
template <typename S>
inline bool lu_solve(int n, S *A, S *B, int nrhs) {
    std::vector<int> piv(n);
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = absval(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            const double v = absval(A[(size_t)k * n + i]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (best == 0.0) return false;
        if (p != k)
            for (int j = 0; j < n; ++j) std::swap(A[(size_t)j * n + k], A[(size_t)j * n + p]);
        const S inv = S(1.0) / A[(size_t)k * n + k];
        for (int i = k + 1; i < n; ++i) A[(size_t)k * n + i] *= inv;
        for (int j = k + 1; j < n; ++j) {
            const S akj = A[(size_t)j * n + k];
            if (akj == S(0.0)) continue;
            S *cj = A + (size_t)j * n;
            const S *lk = A + (size_t)k * n;
            for (int i = k + 1; i < n; ++i) cj[i] -= lk[i] * akj;
        }
    }
    for (int c = 0; c < nrhs; ++c) {
        S *b = B + (size_t)c * n;
        for (int k = 0; k < n; ++k)
            if (piv[k] != k) std::swap(b[k], b[piv[k]]);
        for (int k = 0; k < n; ++k) {  // L y = b (unit lower)
            const S bk = b[k];
            if (bk == S(0.0)) continue;
            const S *lk = A + (size_t)k * n;
            for (int i = k + 1; i < n; ++i) b[i] -= lk[i] * bk;
        }
        for (int k = n - 1; k >= 0; --k) {  // U x = y
            b[k] /= A[(size_t)k * n + k];
            const S bk = b[k];
            const S *uk = A + (size_t)k * n;
            for (int i = 0; i < k; ++i) b[i] -= uk[i] * bk;
        }
    }
    return true;
}

// ---- Higham 2005 scaling & squaring, Base-compatible variant ------------------------------------
static const double PADE_C3[] = {120.0, 60.0, 12.0, 1.0};
static const double PADE_C5[] = {30240.0, 15120.0, 3360.0, 420.0, 30.0, 1.0};
static const double PADE_C7[] = {17297280.0, 8648640.0, 1995840.0, 277200.0, 25200.0, 1512.0, 56.0, 1.0};
static const double PADE_C9[] = {17643225600.0, 8821612800.0, 2075673600.0, 302702400.0, 30270240.0,
                                 2162160.0, 110880.0, 3960.0, 90.0, 1.0};
static const double PADE_C13[] = {64764752532480000.0, 32382376266240000.0, 7771770303897600.0,
                                  1187353796428800.0, 129060195264000.0, 10559470521600.0,
                                  670442572800.0, 33522128640.0, 1323241920.0, 40840800.0, 960960.0,
                                  16380.0, 182.0, 1.0};

template <typename S>
struct ExpWorkT {  // the (A2, P, U, V, temp) tuple of alloc_mem (exp_baseexp.jl:14-40)
    int n = 0;
    std::vector<S> A2, P, U, V, T;
    void reserve(int n_) {
        if (n_ == n) return;
        n = n_;
        const size_t s = (size_t)n * n;
        A2.resize(s); P.resize(s); U.resize(s); V.resize(s); T.resize(s);
    }
};

typedef ExpWorkT<double> ExpWork;

// X (= A, in place) <- r_N(A); returns false if the denominator is singular.
template <typename S>
inline bool pade_evaluate(int n, S *A, const double *C, int N, ExpWorkT<S> &w) {
    const size_t s = (size_t)n * n;
    S *A2 = w.A2.data(), *P = w.P.data(), *U = w.U.data(), *V = w.V.data(), *T = w.T.data();
    matmul(n, A, A, A2);
    std::fill(P, P + s, S(0.0));
    for (int i = 0; i < n; ++i) P[(size_t)i * n + i] = S(1.0);
    for (size_t i = 0; i < s; ++i) { U[i] = C[1] * P[i]; V[i] = C[0] * P[i]; }
    for (int k = 1; k <= N / 2 - 1; ++k) {
        matmul(n, P, A2, T);
        std::swap(P, T);
        const double cu = C[2 * k + 1], cv = C[2 * k];
        for (size_t i = 0; i < s; ++i) { U[i] += cu * P[i]; V[i] += cv * P[i]; }
    }
    matmul(n, A, U, T);  // U = A * U
    std::swap(U, T);
    for (size_t i = 0; i < s; ++i) { A[i] = V[i] + U[i]; T[i] = V[i] - U[i]; }
    return lu_solve(n, T, A, n);
}

// exponential!(A) in place on a dense n x n column-major matrix with leading dimension n.
// Returns 0, or 3 (ESINGULAR).
template <typename S>
inline int expm_higham2005base(int n, S *A, ExpWorkT<S> &w) {
    if (n == 0) return 0;
    w.reserve(n);
    Balance bal;
    balance(n, A, bal);
    const double nA = norm1(n, A);
    bool ok;
    if (nA <= 2.1) {
        if (nA > 0.95) ok = pade_evaluate(n, A, PADE_C9, 10, w);
        else if (nA > 0.25) ok = pade_evaluate(n, A, PADE_C7, 8, w);
        else if (nA > 0.015) ok = pade_evaluate(n, A, PADE_C5, 6, w);
        else ok = pade_evaluate(n, A, PADE_C3, 4, w);
    } else {
        const double s = std::log2(nA / 5.4);
        int si = 0;
        const size_t sz = (size_t)n * n;
        if (s > 0) {
            si = (int)std::ceil(std::min(s, 1100.0));  // (Inf norm: bounded loop, the result is NaN anyway)
            const double f = std::ldexp(1.0, si);
            for (size_t i = 0; i < sz; ++i) A[i] /= f;
        }
        ok = pade_evaluate(n, A, PADE_C13, 14, w);
        if (ok && s > 0) {
            S *T = w.T.data();
            for (int q = 0; q < si; ++q) {
                matmul(n, A, A, T);
                std::copy(T, T + sz, A);
            }
        }
    }
    if (!ok) return 3;
    unbalance(n, A, bal);
    return 0;
}

// ---- symmetric tridiagonal eigen-decomposition (implicit QL) ------------------------------------
// d[0..n-1] diagonal, e[0..n-2] sub-diagonal.  On return d holds the eigenvalues and Z (n x n
// column-major) the orthonormal eigenvectors.  Returns false if an eigenvalue fails to converge.
inline bool symtridiag_eig(int n, std::vector<double> &d, std::vector<double> e_in,
                           std::vector<double> &Z) {
    Z.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) Z[(size_t)i * n + i] = 1.0;
    if (n <= 1) return true;
    std::vector<double> e(n, 0.0);
    for (int i = 0; i + 1 < n; ++i) e[i] = e_in[i];
    const double eps = std::numeric_limits<double>::epsilon();
    for (int l = 0; l < n; ++l) {
        int iter = 0;
        int m;
        do {
            for (m = l; m < n - 1; ++m) {
                const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= eps * dd) break;
            }
            if (m != l) {
                if (iter++ == 60) return false;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    e[i + 1] = (r = std::hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    double *zi = Z.data() + (size_t)i * n, *zi1 = Z.data() + (size_t)(i + 1) * n;
                    for (int k = 0; k < n; ++k) {
                        f = zi1[k];
                        zi1[k] = s * zi[k] + c * f;
                        zi[k] = c * zi[k] - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    return true;
}

// Same QL iteration, but only rows 0 and n-1 of the eigenvector matrix are accumulated (O(n) per rotation):
// enough for e_n' exp(tT) e_1 = sum_k exp(t lambda_k) Z[0,k] Z[n-1,k], the error estimate of
// krylov_phiv_error_estimate.jl:189-191.
inline bool symtridiag_eig_firstlast(int n, std::vector<double> &d, std::vector<double> e_in,
                                     std::vector<double> &zfirst, std::vector<double> &zlast) {
    zfirst.assign(n, 0.0);
    zlast.assign(n, 0.0);
    if (n == 0) return true;
    // row 0 of Z starts as e_0', row n-1 as e_{n-1}'
    zfirst[0] = 1.0;
    zlast[n - 1] = 1.0;
    if (n == 1) return true;
    std::vector<double> e(n, 0.0);
    for (int i = 0; i + 1 < n; ++i) e[i] = e_in[i];
    const double eps = std::numeric_limits<double>::epsilon();
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= eps * dd) break;
            }
            if (m != l) {
                if (iter++ == 60) return false;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    e[i + 1] = (r = std::hypot(f, g));
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    f = zfirst[i + 1];
                    zfirst[i + 1] = s * zfirst[i] + c * f;
                    zfirst[i] = c * zfirst[i] - s * f;
                    f = zlast[i + 1];
                    zlast[i + 1] = s * zlast[i] + c * f;
                    zlast[i] = c * zlast[i] - s * f;
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    return true;
}

// expHe = exp(t T) e1 for the symmetric tridiagonal T = tridiag(e, d, e)  (krylov_phiv.jl:227-229):
//   expHe = Z * (exp.(t * lambda) .* Z[1, :]).
inline bool exp_symtridiag_e1(int m, const double *H, int ldh, double t, double *out) {
    std::vector<double> d(m), e(std::max(m - 1, 0)), Z;
    for (int i = 0; i < m; ++i) d[i] = H[(size_t)i * ldh + i];
    for (int i = 0; i + 1 < m; ++i) e[i] = H[(size_t)i * ldh + i + 1];
    if (!symtridiag_eig(m, d, e, Z)) return false;
    for (int i = 0; i < m; ++i) out[i] = 0.0;
    for (int k = 0; k < m; ++k) {
        const double wk = std::exp(t * d[k]) * Z[(size_t)k * m + 0];
        const double *zk = Z.data() + (size_t)k * m;
        for (int i = 0; i < m; ++i) out[i] += zk[i] * wk;
    }
    return true;
}

// ishermitian(Hcopy) for a real m x m block: exact symmetry (krylov_phiv.jl:225).
inline bool is_exactly_symmetric(int m, const double *H, int ldh) {
    for (int j = 0; j < m; ++j)
        for (int i = j + 1; i < m; ++i)
            if (H[(size_t)j * ldh + i] != H[(size_t)i * ldh + j]) return false;
    return true;
}

// phiv_dense!(w, A, v, k): w (m x (k+1), ld ldw) = [phi_0(A) v, ..., phi_k(A) v]  (phi.jl:84-115).
template <typename S>
inline int phiv_dense(int m, const S *A, int lda, const S *v, int k, S *w, int ldw, ExpWorkT<S> &work) {
    const int N = m + k;
    std::vector<S> C((size_t)N * N, S(0.0));
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) C[(size_t)j * N + i] = A[(size_t)j * lda + i];
    for (int i = 0; i < m; ++i) C[(size_t)m * N + i] = v[i];
    for (int i = m; i < m + k - 1; ++i) C[(size_t)(i + 1) * N + i] = S(1.0);
    const int st = expm_higham2005base(N, C.data(), work);
    if (st) return st;
    for (int i = 0; i < m; ++i) {
        S s = S(0.0);
        for (int j = 0; j < m; ++j) s += C[(size_t)j * N + i] * v[j];
        w[i] = s;
    }
    for (int c = 1; c <= k; ++c)
        for (int i = 0; i < m; ++i) w[(size_t)c * ldw + i] = C[(size_t)(m + c - 1) * N + i];
    return 0;
}

}  // namespace smallmat
}  // namespace b200k
