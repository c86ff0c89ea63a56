/*
 * cpu_krylov.c -- "best-effort CPU" restatement of the reference's Krylov factorisation in C + OpenMP.
 *
 * THIS IS TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it (through oracle/cpu_fast.py).
 *
 * Same algorithm as the reference (paths relative to /root/reference, v1.35.0):
 *   firststep!      src/arnoldi.jl:230-250      beta = ||b||, v_1 = b / beta
 *   applyA!         src/arnoldi.jl:183-187      y = A x            (CSR, one OpenMP loop over rows)
 *   arnoldi_step!   src/arnoldi.jl:289-308      SEQUENTIAL MODIFIED Gram-Schmidt over the IOP window,
 *                                               h_ij = <v_i, y>; y -= h_ij v_i; h_{j+1,j} = ||y||; y /= h_{j+1,j}
 *   arnoldi!        src/arnoldi.jl:345-377      absolute happy-breakdown test beta_j < tol
 *   lanczos_step!   src/arnoldi.jl:388-403      alpha = <v_j, A v_j>; y -= alpha v_j; y -= beta_{j-1} v_{j-1}
 *   lanczos!        src/arnoldi.jl:456-490      mirror copy of the sub-diagonal
 *   expv! (gemv)    src/krylov_phiv.jl:244      w = beta * V[:, 1:m] * y
 * What "best effort" means: every pass over a length-n vector is an OpenMP
 * parallel loop on all cores, and the axpy of window column i is fused with the dot product of column i + 1
 * (one pass instead of two), which is the most a CPU can do without changing the MGS recurrence.  Julia's own
 * SparseMatrixCSC mul! is serial; this is therefore faster than what a user of the package gets.
 * Reductions use a fixed static schedule, so results are deterministic for a given thread count.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int cpuk_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void cpuk_set_threads(int nt) {
#ifdef _OPENMP
    if (nt > 0) omp_set_num_threads(nt);
#else
    (void)nt;
#endif
}

/* y = A x (CSR, 0-based) */
static void spmv(int64_t n, const int32_t *rowptr, const int32_t *colind, const double *val, const double *x,
                 double *y) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        double s = 0.0;
        for (int32_t e = rowptr[r]; e < rowptr[r + 1]; ++e) s += val[e] * x[colind[e]];
        y[r] = s;
    }
}

static double dot(int64_t n, const double *x, const double *y) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

/* y -= a x, returns <z, y_new> (z may be NULL: returns ||y_new||^2) -- the fused MGS pass */
static double axpy_dot(int64_t n, double a, const double *x, double *y, const double *z) {
    double s = 0.0;
    if (z) {
#pragma omp parallel for schedule(static) reduction(+ : s)
        for (int64_t i = 0; i < n; ++i) {
            const double v = y[i] - a * x[i];
            y[i] = v;
            s += z[i] * v;
        }
    } else {
#pragma omp parallel for schedule(static) reduction(+ : s)
        for (int64_t i = 0; i < n; ++i) {
            const double v = y[i] - a * x[i];
            y[i] = v;
            s += v * v;
        }
    }
    return s;
}

/* y = (y - a x) - c z in one pass (the two axpys of lanczos_step!, same per-element rounding order); returns ||y||^2 */
static double axpy2_nrm(int64_t n, double a, const double *x, double c, const double *z, double *y) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t i = 0; i < n; ++i) {
        const double v = (y[i] - a * x[i]) - c * z[i];
        y[i] = v;
        s += v * v;
    }
    return s;
}

static void scale_to(int64_t n, double a, const double *x, double *y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) y[i] = x[i] * a;
}

static void div_inplace(int64_t n, double beta, double *y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) y[i] /= beta;
}

/*
 * arnoldi!(Ks, A, b; tol, m, iop) / lanczos! for a CSR operator.
 *   V: n x (m+1) column-major (ld = n);  H: (m+1) x m column-major (ld = m+1), zero-filled here.
 *   Returns 0; *beta = ||b||, *m_out = Ks.m (smaller on happy breakdown), *breakdown = 0/1.
 */
int cpuk_arnoldi(int64_t n, const int32_t *rowptr, const int32_t *colind, const double *val, const double *b, int m,
                 double tol, int iop, int lanczos, double *V, double *H, double *beta, int *m_out, int *breakdown) {
    const int ldh = m + 1;
    memset(H, 0, sizeof(double) * (size_t)ldh * (size_t)m);
    *m_out = m;
    *breakdown = 0;
    const double bn = sqrt(dot(n, b, b));
    *beta = bn;
    if (bn == 0.0) return 0; /* V stays untouched (arnoldi.jl:240-244) */
    scale_to(n, 1.0 / bn, b, V);
    if (iop == 0) iop = m;
    for (int j = 1; j <= m; ++j) {
        const double *x = V + (size_t)(j - 1) * (size_t)n;
        double *y = V + (size_t)j * (size_t)n;
        spmv(n, rowptr, colind, val, x, y);
        double bj;
        if (lanczos) {
            /* alpha = <x, y>; y -= alpha x; y -= beta_{j-1} v_{j-1}; beta_j = ||y|| (arnoldi.jl:396-401) */
            const double alpha = dot(n, x, y);
            H[(size_t)(j - 1) * ldh + (j - 1)] = alpha;
            double nrm2;
            if (j > 1) {
                const double bp = H[(size_t)(j - 2) * ldh + (j - 1)];
                nrm2 = axpy2_nrm(n, alpha, x, bp, V + (size_t)(j - 2) * (size_t)n, y);
            } else {
                nrm2 = axpy_dot(n, alpha, x, y, NULL);
            }
            bj = sqrt(nrm2);
        } else {
            int i0 = j - iop + 1;
            if (i0 < 1) i0 = 1;
            /* sequential MGS; the update with column i is fused with the dot against column i + 1 */
            double h = dot(n, V + (size_t)(i0 - 1) * (size_t)n, y);
            double nrm2 = 0.0;
            for (int i = i0; i <= j; ++i) {
                H[(size_t)(j - 1) * ldh + (i - 1)] = h;
                const double *vi = V + (size_t)(i - 1) * (size_t)n;
                if (i < j) h = axpy_dot(n, h, vi, y, V + (size_t)i * (size_t)n);
                else nrm2 = axpy_dot(n, h, vi, y, NULL);
            }
            bj = sqrt(nrm2);
        }
        H[(size_t)(j - 1) * ldh + j] = bj;
        div_inplace(n, bj, y);
        if (bj < tol) {
            *m_out = j;
            *breakdown = 1;
            break;
        }
    }
    if (lanczos) /* copyto!(@diagview(H, 1), sub-diagonal) (arnoldi.jl:488) */
        for (int i = 0; i + 1 < m; ++i) H[(size_t)(i + 1) * ldh + i] = H[(size_t)i * ldh + i + 1];
    return 0;
}

/* w = beta * V[:, 0:m] * y */
void cpuk_project(int64_t n, int m, const double *V, const double *y, double beta, double *w) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = 0; k < m; ++k) s += V[(size_t)k * (size_t)n + i] * y[k];
        w[i] = beta * s;
    }
}
