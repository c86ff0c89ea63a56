// smallexp_kernel.cuh -- the small dense phase of expv! on the device: y = exp(t * H[1:m,1:m]) e1.
//
// One CTA per problem, matrices in shared memory.  Launched on the same stream right behind the Krylov
// kernel, it reads the device copy of H / beta / (m, breakdown) and writes the projection coefficients for
// project_kernel, so the fused one-shot expv (and every problem of a batch) needs no host round trip.
//
// Algorithm = exponential!(A, ExpMethodHigham2005Base()) of the reference (src/exp_baseexp.jl:112-161), as in
// smallmat.hpp: balance (power-of-two diagonal scaling; the permutation phase of xGEBAL is skipped -- it is
// the identity for an unreduced Hessenberg matrix) -> 1-norm switch Pade 3/5/7/9/13 (generic even/odd power
// loop) -> LU solve with partial pivoting -> squaring -> unbalance.  The reference takes an eigen-decomposition
// branch when H is exactly symmetric (krylov_phiv.jl:225-229); on the device the Pade path is used for both
// (agrees with the eigen branch to rounding, see tests); the host path b200k_expv_ks keeps both branches.
#pragma once
#include <cuda_runtime.h>

namespace b200k {

constexpr int SE_NT = 1024;
constexpr int SE_MAXM = 48;  // 6 m^2 doubles of shared memory (110 KB at m = 48)

struct SmallExpParams {
    const double *Hd;   // [nprob][ldh * (hcols)] device H written by the Krylov kernel
    int ldh;
    long long H_stride;
    const double *scal;  // [nprob*4]: beta
    const int *stat;     // [nprob*4]: m_out, breakdown
    const double *tvec;  // [nprob] device, or nullptr -> t
    double t;
    int m;               // requested dimension (used when beta == 0)
    int lanczos;         // H holds only diagonal + sub-diagonal: mirror it
    double *Y;           // [nprob][ldy] out
    int ldy;
    double *betavec;     // [nprob] out
    int *mvec;           // [nprob] out
    int *err;            // set to 1 if a Pade denominator is singular
};

__device__ __forceinline__ void se_matmul(int n, const double *A, const double *B, double *C) {
    // C = A * B, column-major n x n in shared memory; thread -> (i, j)
    for (int idx = threadIdx.x; idx < n * n; idx += SE_NT) {
        const int i = idx % n, j = idx / n;
        double s = 0.0;
#pragma unroll 6
        for (int k = 0; k < n; ++k) s = fma(A[k * n + i], B[j * n + k], s);
        C[idx] = s;
    }
    __syncthreads();
}

// exp of the n x n column-major matrix in sm[0 .. n*n) (shared memory, 6 n^2 doubles of workspace behind it).
// Returns a pointer (into sm) to the exponential of the BALANCED matrix: exp(A)[i, j] = sc[i] * X[i, j] / sc[j];
// nullptr if the Pade denominator is singular.  `full` = all columns are needed (otherwise only column 0).
__device__ double *se_expm_core(int n, double *sm, double *sc, double *colsum, bool full) {
    __shared__ double s_nA;
    __shared__ int s_piv, s_flag;
    const int tid = threadIdx.x;
    const int nn = n * n;
    double *A = sm, *A2 = sm + nn, *Pm = sm + 2 * nn, *U = sm + 3 * nn, *V = sm + 4 * nn, *T = sm + 5 * nn;
    if (tid < n) sc[tid] = 1.0;
    __syncthreads();

    // ---- balance (xGEBAL scaling phase).  A Krylov Hessenberg matrix built on an orthonormal basis is almost
    // always already balanced, so first ask in parallel whether ANY row would be rescaled; only then run the
    // (inherently sequential, Gauss-Seidel) sweeps on warp 0.
    int need = 0;
    if (tid < n) {
        double c = 0.0, r = 0.0;
        for (int q = 0; q < n; ++q) {
            c = fma(A[tid * n + q], A[tid * n + q], c);
            r = fma(A[q * n + tid], A[q * n + tid], r);
        }
        c = sqrt(c);
        r = sqrt(r);
        if (c != 0.0 && r != 0.0 && (c + r == c + r)) {
            double g = r / 2.0, f = 1.0;
            const double s0 = c + r;
            int guard = 0;
            while (c < g && guard++ < 1100) { f *= 2.0; c *= 2.0; r /= 2.0; g /= 2.0; }
            g = c / 2.0;
            guard = 0;
            while (g >= r && guard++ < 1100) { f /= 2.0; c /= 2.0; g /= 2.0; r *= 2.0; }
            if (c + r < 0.95 * s0 && f != 1.0) need = 1;
        }
    }
    need = __syncthreads_or(need);
    if (need && tid < 32) {
        const int lane = tid;
        for (int sweep = 0; sweep < 32; ++sweep) {
            bool noconv = false;
            for (int i = 0; i < n; ++i) {
                double c = 0.0, r = 0.0;
                for (int q = lane; q < n; q += 32) {
                    c = fma(A[i * n + q], A[i * n + q], c);
                    r = fma(A[q * n + i], A[q * n + i], r);
                }
                for (int o = 16; o > 0; o >>= 1) {
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                    r += __shfl_xor_sync(0xffffffffu, r, o);
                }
                c = sqrt(c);
                r = sqrt(r);
                if (c == 0.0 || r == 0.0 || !(c + r == c + r)) continue;
                double g = r / 2.0, f = 1.0;
                const double s = c + r;
                int guard = 0;
                while (c < g && guard++ < 1100) { f *= 2.0; c *= 2.0; r /= 2.0; g /= 2.0; }
                g = c / 2.0;
                guard = 0;
                while (g >= r && guard++ < 1100) { f /= 2.0; c /= 2.0; g /= 2.0; r *= 2.0; }
                if (c + r >= 0.95 * s || f == 1.0) continue;
                noconv = true;
                if (lane == 0) sc[i] *= f;
                const double gi = 1.0 / f;
                for (int q = lane; q < n; q += 32) A[q * n + i] *= gi;  // row i
                __syncwarp();
                for (int q = lane; q < n; q += 32) A[i * n + q] *= f;   // column i
                __syncwarp();
            }
            if (!noconv) break;
        }
    }
    __syncthreads();

    // ---- nA = opnorm(A, 1) ----
    if (tid < n) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += fabs(A[tid * n + i]);
        colsum[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        double best = 0.0;
        for (int j = 0; j < n; ++j)
            if (colsum[j] > best || colsum[j] != colsum[j]) best = colsum[j];
        s_nA = best;
    }
    __syncthreads();
    const double nA = s_nA;
    const double C3[] = {120.0, 60.0, 12.0, 1.0};
    const double C5[] = {30240.0, 15120.0, 3360.0, 420.0, 30.0, 1.0};
    const double C7[] = {17297280.0, 8648640.0, 1995840.0, 277200.0, 25200.0, 1512.0, 56.0, 1.0};
    const double C9[] = {17643225600.0, 8821612800.0, 2075673600.0, 302702400.0, 30270240.0,
                         2162160.0, 110880.0, 3960.0, 90.0, 1.0};
    const double C13[] = {64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0,
                          129060195264000.0, 10559470521600.0, 670442572800.0, 33522128640.0, 1323241920.0,
                          40840800.0, 960960.0, 16380.0, 182.0, 1.0};
    const double *C;
    int N, si = 0;
    if (nA <= 2.1) {
        if (nA > 0.95) { C = C9; N = 10; }
        else if (nA > 0.25) { C = C7; N = 8; }
        else if (nA > 0.015) { C = C5; N = 6; }
        else { C = C3; N = 4; }
    } else {
        C = C13;
        N = 14;
        const double s = log2(nA / 5.4);
        if (s > 0) {
            si = (int)ceil(s);
            const double f = ldexp(1.0, si);
            for (int idx = tid; idx < nn; idx += SE_NT) A[idx] /= f;
            __syncthreads();
        }
    }

    // ---- Pade numerator / denominator (generic even/odd power loop, exp_baseexp.jl:84-105) ----
    se_matmul(n, A, A, A2);
    for (int idx = tid; idx < nn; idx += SE_NT) {
        const double p = (idx % n == idx / n) ? 1.0 : 0.0;
        Pm[idx] = p;
        U[idx] = C[1] * p;
        V[idx] = C[0] * p;
    }
    __syncthreads();
    for (int k = 1; k <= N / 2 - 1; ++k) {
        se_matmul(n, Pm, A2, T);
        double *tmp = Pm; Pm = T; T = tmp;
        const double cu = C[2 * k + 1], cv = C[2 * k];
        for (int idx = tid; idx < nn; idx += SE_NT) {
            U[idx] = fma(cu, Pm[idx], U[idx]);
            V[idx] = fma(cv, Pm[idx], V[idx]);
        }
        __syncthreads();
    }
    se_matmul(n, A, U, T);  // U = A * U  (in T)
    // X (in A) = V + U ; D (in A2) = V - U
    for (int idx = tid; idx < nn; idx += SE_NT) {
        const double u = T[idx], v = V[idx];
        A[idx] = v + u;
        A2[idx] = v - u;
    }
    __syncthreads();

    // ---- LU with partial pivoting on D (A2), applied to the right-hand sides X (A) ----
    double *D = A2, *X = A;
    for (int k = 0; k < n; ++k) {
        if (tid < 32) {
            double best = -1.0;
            int bi = k;
            for (int i = k + tid; i < n; i += 32) {
                const double v = fabs(D[k * n + i]);
                if (v > best) { best = v; bi = i; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (tid == 0) {
                s_piv = bi;
                s_flag = (best == 0.0 || best != best) ? 1 : 0;
            }
        }
        __syncthreads();
        if (s_flag) return nullptr;  // singular Pade denominator (uniform for the CTA)
        const int pv = s_piv;
        if (pv != k) {  // swap rows k and pv of D and X
            for (int j = tid; j < 2 * n; j += SE_NT) {
                double *M = j < n ? D : X;
                const int jj = j < n ? j : j - n;
                const double a = M[jj * n + k];
                M[jj * n + k] = M[jj * n + pv];
                M[jj * n + pv] = a;
            }
            __syncthreads();
        }
        // trailing update of D (columns k+1..) and forward elimination of X (all columns); the multipliers
        // l_ik = D(i,k) / D(k,k) are applied on the fly (L itself is not needed afterwards)
        const double inv = 1.0 / D[k * n + k];
        const int rows = n - k - 1;
        for (int idx = tid; idx < rows * (rows + n); idx += SE_NT) {
            const int i = k + 1 + idx % rows;
            const int jc = idx / rows;
            const double l = D[k * n + i] * inv;
            if (jc < rows) {
                const int j = k + 1 + jc;
                D[j * n + i] = fma(-l, D[j * n + k], D[j * n + i]);
            } else {
                const int j = jc - rows;
                X[j * n + i] = fma(-l, X[j * n + k], X[j * n + i]);
            }
        }
        __syncthreads();
    }
    // back substitution U x = y: sequential in k, parallel over (row, right-hand side).  Only column 0 is
    // needed when no squaring follows.
    const int nrhs = (si > 0 || full) ? n : 1;
    for (int k = n - 1; k >= 0; --k) {
        const double dkk = D[k * n + k];
        for (int j = tid; j < nrhs; j += SE_NT) X[j * n + k] /= dkk;
        __syncthreads();
        for (int idx = tid; idx < k * nrhs; idx += SE_NT) {
            const int i = idx % k, j = idx / k;
            X[j * n + i] = fma(-D[k * n + i], X[j * n + k], X[j * n + i]);
        }
        __syncthreads();
    }
    // ---- squaring ----
    double *Xc = X, *Xn = U;
    for (int q = 0; q < si; ++q) {
        se_matmul(n, Xc, Xc, Xn);
        double *tmp = Xc; Xc = Xn; Xn = tmp;
    }
    return Xc;
}

__global__ void __launch_bounds__(SE_NT) small_exp_kernel(const SmallExpParams P) {
    extern __shared__ double sm[];
    __shared__ double sc[SE_MAXM];   // balancing scale factors
    __shared__ double colsum[SE_MAXM];
    const int prob = blockIdx.x;
    const int tid = threadIdx.x;
    const double beta = P.scal[prob * 4];
    int n = P.stat[prob * 4 + 0];
    if (beta == 0.0) n = P.m;
    if (tid == 0) {
        P.betavec[prob] = beta;
        P.mvec[prob] = n;
    }
    double *yout = P.Y + (long long)prob * P.ldy;
    if (beta == 0.0 || n < 1 || n > SE_MAXM) {
        for (int i = tid; i < P.ldy; i += SE_NT) yout[i] = 0.0;
        return;
    }
    const double t = P.tvec ? P.tvec[prob] : P.t;
    const double *H = P.Hd + (long long)prob * P.H_stride;
    const int nn = n * n;
    for (int idx = tid; idx < nn; idx += SE_NT) {
        const int i = idx % n, j = idx / n;
        double v = H[(long long)j * P.ldh + i];
        if (P.lanczos && j == i + 1) v = H[(long long)i * P.ldh + j];  // mirror the sub-diagonal
        sm[idx] = t * v;
    }
    __syncthreads();
    const double *Xc = se_expm_core(n, sm, sc, colsum, false);
    if (!Xc) {
        if (tid == 0) *P.err = 1;
        for (int i = tid; i < P.ldy; i += SE_NT) yout[i] = 0.0;
        return;
    }
    // unbalance and take the first column: exp(tH)[i, 0] = sc[i] * X[i, 0] / sc[0]
    for (int i = tid; i < P.ldy; i += SE_NT) yout[i] = i < n ? Xc[i] * sc[i] / sc[0] : 0.0;
}

// Batched exponential!(A_b, ExpMethodHigham2005Base()) for many small matrices resident on the device
// (SURVEY 8f-4): one CTA per matrix, in place.  A: [nbatch] matrices, n x n column-major with leading
// dimension lda and `stride` doubles between matrices.
__global__ void __launch_bounds__(SE_NT) small_exp_batched_kernel(int n, double *A, int lda, long long stride, int *err) {
    extern __shared__ double sm[];
    __shared__ double sc[SE_MAXM];
    __shared__ double colsum[SE_MAXM];
    const int tid = threadIdx.x;
    double *Ab = A + (long long)blockIdx.x * stride;
    const int nn = n * n;
    for (int idx = tid; idx < nn; idx += SE_NT) sm[idx] = Ab[(long long)(idx / n) * lda + idx % n];
    __syncthreads();
    const double *Xc = se_expm_core(n, sm, sc, colsum, true);
    if (!Xc) {
        if (tid == 0) atomicExch(err, 1);
        return;
    }
    for (int idx = tid; idx < nn; idx += SE_NT) {
        const int i = idx % n, j = idx / n;
        Ab[(long long)j * lda + i] = Xc[idx] * sc[i] / sc[j];
    }
}

}  // namespace b200k
