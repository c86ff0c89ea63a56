// smallexp_kernel.cuh -- the small dense phase of expv! on the device: y = exp(t * H[1:m,1:m]) e1.
//
// One CTA per problem, matrices in shared memory.  Launched on the same stream right behind the Krylov
// kernel, it reads the device copy of H / beta / (m, breakdown) and writes the projection coefficients for
// project_kernel, so the fused one-shot expv (and every problem of a batch) needs no host round trip.
//
// Algorithm = exponential!(A, ExpMethodHigham2005Base()) of the reference (src/exp_baseexp.jl:112-161), as in
// smallmat.hpp: balance (power-of-two diagonal scaling; the permutation phase of xGEBAL is skipped -- it is
// the identity for an unreduced Hessenberg matrix) -> 1-norm switch Pade 3/5/7/9/13 (generic even/odd power
// loop) -> LU solve with (implicit) partial pivoting -> squaring -> unbalance.  The reference takes an eigen-decomposition
// branch when H is exactly symmetric (krylov_phiv.jl:225-229); on the device the Pade path is used for both
// (agrees with the eigen branch to rounding, see tests); the host path b200k_expv_ks keeps both branches.
#pragma once
#include <cuda_runtime.h>

namespace b200k {

#ifdef B200K_PHASE_TIMING
__device__ long long g_se_ts[16];
#define SE_MARK(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_se_ts[k] = clock64(); } while (0)
#else
#define SE_MARK(k) do { } while (0)
#endif

__device__ __forceinline__ double se_lds(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void se_sts(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

constexpr int SE_NT = 1024;
constexpr int SE_MAXM = 48;  // 6 m^2 doubles of shared memory (110 KB at m = 48)

// Pade coefficient tuples of exp_baseexp.jl:65-77, concatenated: [C3 | C5 | C7 | C9 | C13]
__constant__ double SE_PADE[4 + 6 + 8 + 10 + 14] = {
    120.0, 60.0, 12.0, 1.0,
    30240.0, 15120.0, 3360.0, 420.0, 30.0, 1.0,
    17297280.0, 8648640.0, 1995840.0, 277200.0, 25200.0, 1512.0, 56.0, 1.0,
    17643225600.0, 8821612800.0, 2075673600.0, 302702400.0, 30270240.0, 2162160.0, 110880.0, 3960.0, 90.0, 1.0,
    64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0, 129060195264000.0,
    10559470521600.0, 670442572800.0, 33522128640.0, 1323241920.0, 40840800.0, 960960.0, 16380.0, 182.0, 1.0};

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
    int force_pade;      // tests / A-B: never take the Chebyshev path for symmetric tridiagonal H
    double *Y;           // [nprob][ldy] out
    int ldy;
    double *betavec;     // [nprob] out
    int *mvec;           // [nprob] out
    int *err;            // set to 1 if a Pade denominator is singular
};

// C = A * B, column-major n x n in shared memory, on the fp64 tensor-core path: one warp per 8 x 8 output tile,
// mma.sync.m8n8k4.f64 over k (zero-padded at the edges by predicated loads).  The scalar version (one thread per
// entry, 2 LDS per FMA) was shared-memory-bandwidth bound: 4.2k cycles per 30 x 30 product, nine of them per
// exponential; this one needs 2 LDS per 256 FMAs.  Fragment layout (PTX ISA, m8n8k4 .f64): A[g][t], B[t][g],
// C[g][2t], C[g][2t+1] with g = lane / 4, t = lane % 4.  Summation order differs from the scalar loop (rounding).
__device__ __forceinline__ void se_matmul(int n, const double *A, const double *B, double *C) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nt = (n + 7) >> 3;
    for (int tile = warp; tile < nt * nt; tile += SE_NT / 32) {
        const int ti = tile % nt, tj = tile / nt;
        const int row = ti * 8 + g;          // row of the A fragment / of both C entries
        const int colb = tj * 8 + g;         // column of the B fragment
        double c0 = 0.0, c1 = 0.0;
        const bool rok = row < n, cok = colb < n;
        for (int k0 = 0; k0 < n; k0 += 4) {
            const int k = k0 + t;
            const double a = (rok && k < n) ? A[k * n + row] : 0.0;
            const double b = (cok && k < n) ? B[colb * n + k] : 0.0;
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c0), "+d"(c1)
                         : "d"(a), "d"(b));
        }
        const int cc = tj * 8 + 2 * t;
        if (rok) {
            if (cc < n) C[cc * n + row] = c0;
            if (cc + 1 < n) C[(cc + 1) * n + row] = c1;
        }
    }
    __syncthreads();
}

// exp of the n x n column-major matrix in sm[0 .. n*n) (shared memory, 6 n^2 doubles of workspace behind it).
// Returns a pointer (into sm) to the exponential of the BALANCED matrix: exp(A)[i, j] = sc[i] * X[i, j] / sc[j];
// nullptr if the Pade denominator is singular.  `full` = all columns are needed (otherwise only column 0).
__device__ double *se_expm_core(int n, double *sm, double *sc, double *colsum, bool full) {
    __shared__ int s_cfg[3];
    const int tid = threadIdx.x;
    const int nn = n * n;
    double *A = sm, *A2 = sm + nn, *Pm = sm + 2 * nn, *U = sm + 3 * nn, *V = sm + 4 * nn, *T = sm + 5 * nn;
    SE_MARK(1);
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

    SE_MARK(2);
    // ---- nA = opnorm(A, 1) ----
    if (tid < n) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += fabs(A[tid * n + i]);
        colsum[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {  // 1-norm, Pade order and number of squarings: one thread, broadcast through shared memory
        double best = 0.0;
        for (int j = 0; j < n; ++j)
            if (colsum[j] > best || colsum[j] != colsum[j]) best = colsum[j];
        int off, N_, si_ = 0;
        if (best <= 2.1) {
            if (best > 0.95) { off = 18; N_ = 10; }
            else if (best > 0.25) { off = 10; N_ = 8; }
            else if (best > 0.015) { off = 4; N_ = 6; }
            else { off = 0; N_ = 4; }
        } else {
            off = 28;
            N_ = 14;
            const double s2 = log2(best / 5.4);
            if (s2 > 0) si_ = (int)ceil(fmin(s2, 1100.0));  // (Inf norm: bounded squaring loop, the result is NaN anyway)
        }
        s_cfg[0] = off;
        s_cfg[1] = N_;
        s_cfg[2] = si_;
    }
    __syncthreads();
    const double *C = SE_PADE + s_cfg[0];
    const int N = s_cfg[1], si = s_cfg[2];
    if (si > 0) {
        const double f = ldexp(1.0, si);
        for (int idx = tid; idx < nn; idx += SE_NT) A[idx] /= f;
        __syncthreads();
    }

    SE_MARK(3);
    // ---- Pade numerator / denominator (generic even/odd power loop, exp_baseexp.jl:84-105) ----
    se_matmul(n, A, A, A2);
    SE_MARK(4);
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
    SE_MARK(5);
    se_matmul(n, A, U, T);  // U = A * U  (in T)
    // X (in A) = V + U ; D (in A2) = V - U
    for (int idx = tid; idx < nn; idx += SE_NT) {
        const double u = T[idx], v = V[idx];
        A[idx] = v + u;
        A2[idx] = v - u;
    }
    __syncthreads();

    SE_MARK(6);
    // ---- LU with partial pivoting on D (A2), applied to the right-hand sides X (A) ----
    // Pivoting is IMPLICIT: rows are never swapped, a row that has served as pivot is skipped from then on, and the
    // pivot sequence is remembered for the back substitution.  Per step: warp 0 finds the pivot (arg-max of |.| on
    // the INTEGER image of the doubles -- monotone for non-negative values -- because fp64 compares are scarce on
    // this SM: 32 warps each doing the search redundantly cost 1400 cycles per step, measured) and its reciprocal,
    // publishes both through shared memory; then thread (row = lane, column group = warp) eliminates, loading
    // everything it needs before its first store.  The pivot row and column k are not written during step k.
    // Same arithmetic as xGETF2/xGETRS: multipliers l_ik = D(i,k) * (1 / D(p,k)), partial pivoting by max |.|.
    // (History: warp-0 search + physical swap + idx % rows: ~2500 cycles per step; this: see profiles/.)
    double *D = A2, *X = A;
    const int nrhs = (si > 0 || full) ? n : 1;
    __shared__ int s_pivrow[SE_MAXM];     // s_pivrow[k] = row used as the k-th pivot
    __shared__ int s_pivpos[SE_MAXM];     // inverse: s_pivpos[row] = k, -1 while the row is still active
    __shared__ double s_rdiag[SE_MAXM];   // 1 / U(k, k)
    __shared__ int s_pv;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = SE_NT / 32;
    if (tid < SE_MAXM) s_pivpos[tid] = -1;
    __syncthreads();
    const uint32_t sm_a = (uint32_t)__cvta_generic_to_shared(sm);
    for (int k = 0; k < n; ++k) {
        if (k == 5) SE_MARK(11);
        const uint32_t dk_a = sm_a + 8u * (uint32_t)(nn + k * n);  // column k of D
        if (warp == 0) {
            unsigned long long key = 0ull;  // bits of |d| + 1 (0 = no candidate); NaN sorts above everything
            int bi = n;
            double bval = 0.0;
            for (int r = lane; r < n; r += 32) {
                if (s_pivpos[r] < 0) {
                    const double d = D[k * n + r];
                    const unsigned long long kk = ((unsigned long long)__double_as_longlong(d) & 0x7fffffffffffffffull) + 1ull;
                    if (kk > key) { key = kk; bi = r; bval = d; }
                }
            }
            unsigned long long mx = key;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(0xffffffffu, mx, o);
                mx = ok > mx ? ok : mx;
            }
            const unsigned win = __ballot_sync(0xffffffffu, key == mx);
            const int src = __ffs(win) - 1;
            bi = __shfl_sync(0xffffffffu, bi, src);
            bval = __shfl_sync(0xffffffffu, bval, src);
            // mx == 1: every candidate is +-0 (singular); NaN pivot: singular as well (SingularException)
            const bool bad = mx <= 1ull || bval != bval || bi >= n;
            if (lane == 0) {
                s_pv = bad ? -1 : bi;
                if (!bad) {
                    s_pivrow[k] = bi;
                    s_pivpos[bi] = k;
                    s_rdiag[k] = 1.0 / bval;
                }
            }
        }
        __syncthreads();
        if (k == 5) SE_MARK(13);
        const int pv = s_pv;
        if (pv < 0) return nullptr;  // singular Pade denominator (uniform for the CTA)
        const double inv = s_rdiag[k];
        // eliminate column k from the active rows.  [D | X] is one logical matrix of n + nrhs columns: logical
        // column c lives at sm[(c < n ? nn : -nn) + c * n ...] (D = sm + nn, X = sm).  The phase is instruction-issue
        // bound (32 warps, tiny work), so warps without a column leave at once and addresses are 32-bit shared.
        const int ctot = n + nrhs;
#pragma unroll 1
        for (int c = k + 1 + warp; c < ctot; c += NWARP) {
            const uint32_t cb = sm_a + 8u * (uint32_t)(c * n + (c < n ? nn : -nn));
            const double pvv = se_lds(cb + 8u * (uint32_t)pv);
#pragma unroll 1
            for (int r = lane; r < n; r += 32) {
                if (s_pivpos[r] >= 0) continue;
                const double l = se_lds(dk_a + 8u * (uint32_t)r) * inv;
                const uint32_t a = cb + 8u * (uint32_t)r;
                se_sts(a, fma(-l, pvv, se_lds(a)));
            }
        }
        if (k == 5) SE_MARK(14);
        __syncthreads();
        if (k == 5) SE_MARK(15);
    }
    SE_MARK(7);
    // back substitution U x = y in pivot order: x_k = y(p_k) / U(p_k, k); rows that became pivots earlier
    // (position < k) subtract U(row, k) * x_k.  x_k goes to its natural place k of the output T, which also
    // undoes the row permutation.  One barrier per step; reciprocals of the diagonal come from the factorisation.
    const uint32_t t_a = (uint32_t)__cvta_generic_to_shared(T);
    for (int k = n - 1; k >= 0; --k) {
        const int pr = s_pivrow[k];
        const double rk = s_rdiag[k];
        const uint32_t dk = sm_a + 8u * (uint32_t)(nn + k * n);
#pragma unroll 1
        for (int j = warp; j < nrhs; j += NWARP) {
            const uint32_t xb = sm_a + 8u * (uint32_t)(j * n);
            const double xk = se_lds(xb + 8u * (uint32_t)pr) * rk;
#pragma unroll 1
            for (int r = lane; r < n; r += 32) {
                const int pos = s_pivpos[r];
                if (pos > k) continue;
                if (pos == k) se_sts(t_a + 8u * (uint32_t)(j * n + k), xk);
                else {
                    const uint32_t a = xb + 8u * (uint32_t)r;
                    se_sts(a, fma(-se_lds(dk + 8u * (uint32_t)r), xk, se_lds(a)));
                }
            }
        }
        __syncthreads();
    }
    X = T;
    SE_MARK(8);
    // ---- squaring ----
    double *Xc = X, *Xn = U;
    for (int q = 0; q < si; ++q) {
        se_matmul(n, Xc, Xc, Xn);
        double *tmp = Xc; Xc = Xn; Xn = tmp;
    }
    SE_MARK(9);
    return Xc;
}

// ---- symmetric tridiagonal H (Lanczos): exp(t T) e1 by ONE warp, no matrix function -------------------------------
// The reference takes eigen!(SymTridiagonal) here (krylov_phiv.jl:225-229).  A QL / QR eigensolver is a chain of ~m^2
// dependent rotations (sqrt + 2 divisions each: ~400 k cycles for m = 30 on one SM) and the Pade path costs 65 us
// (nine 30^3 products + a pivoted LU), one CTA busy while 147 SMs idle -- 10 % of a Lanczos expv at C2.  What expv!
// needs is only the VECTOR exp(t T) e1, which a Chebyshev expansion on the spectral interval [a, b] gives with
// three-term recurrences on m-vectors:
//     exp(t x) = e^{t b} * [ i_0(z) + 2 sum_k i_k(z) T_k(xi) ],   x = c + h xi,  z = t h,  i_k = e^{-z} I_k(z)   (t >= 0)
// (t < 0: T -> -T).  K = z + 12 z^(1/3) + 25 terms reach 1e-17 (Bessel coefficients by Miller's backward recurrence).
// The interval must be TIGHT at the upper end: rounding errors are eps * e^{t b}, the result is ~ e^{t lambda_max}.
// So b comes from Sturm-sequence multisection (32 shifts per round, one per lane: each round shrinks the bracket of
// lambda_max 33x) started from the Gershgorin bounds until t (b - lambda_max) <= 0.02; the lower end a only costs
// terms and stays at its Gershgorin value.  Same function of the same matrix as the reference's eigen branch,
// agreeing to ~K eps (tests compare with the host QL path); z > 150 (or NaN) falls back to the Pade path below.
// Measured: ~10 k cycles (5 us) instead of 130 k.
constexpr int LC_KMAX = 256;
constexpr double LC_ZMAX = 150.0;

__device__ bool se_lanczos_cheb(int n, const double *H, int ldh, double t, double *yout, int ldy) {
    __shared__ double s_a[SE_MAXM], s_b[SE_MAXM + 1], s_d[SE_MAXM], s_e[SE_MAXM + 1];
    __shared__ double s_u[3][SE_MAXM + 2];
    __shared__ double s_ck[LC_KMAX + 2];
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const double sg = t < 0.0 ? -1.0 : 1.0;
    const double tt = fabs(t);
    for (int i = lane; i < n; i += 32) {
        s_a[i] = sg * H[(long long)i * ldh + i];
        s_b[i] = (i < n - 1) ? sg * H[(long long)i * ldh + i + 1] : 0.0;  // couples i and i + 1
    }
    __syncwarp();
    if (n == 1) {
        if (lane == 0) yout[0] = exp(tt * s_a[0]);
        for (int i = 1 + lane; i < ldy; i += 32) yout[i] = 0.0;
        return true;
    }
    // Gershgorin bounds
    double glo = 1.0e300, ghi = -1.0e300;
    bool bad = false;
    for (int i = lane; i < n; i += 32) {
        const double r = (i > 0 ? fabs(s_b[i - 1]) : 0.0) + fabs(s_b[i]);
        glo = fmin(glo, s_a[i] - r);
        ghi = fmax(ghi, s_a[i] + r);
        bad = bad || !(s_a[i] - r == s_a[i] - r) || !(fabs(s_a[i]) + r < 1.0e300);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        glo = fmin(glo, __shfl_xor_sync(full, glo, o));
        ghi = fmax(ghi, __shfl_xor_sync(full, ghi, o));
    }
    if (__any_sync(full, bad)) return false;  // NaN / Inf in H: the Pade path propagates it
    // multisection for a tight upper bound of lambda_max: count(x) = #eigenvalues < x (Sturm sequence of T - x I)
    double lo = glo, hi = ghi;
    for (int round = 0; round < 6 && tt * (hi - lo) > 0.02; ++round) {
        const double x = lo + (hi - lo) * (double)(lane + 1) * (1.0 / 33.0);
        double d = s_a[0] - x;
        int cnt = d < 0.0 ? 1 : 0;
        for (int i = 1; i < n; ++i) {
            if (d == 0.0) d = 1.0e-300;
            const double bq = s_b[i - 1];
            d = (s_a[i] - x) - bq * bq / d;
            cnt += d < 0.0 ? 1 : 0;
        }
        const unsigned m_all = __ballot_sync(full, cnt == n);
        if (m_all == 0u) {
            lo = __shfl_sync(full, x, 31);
        } else {
            const int jf = __ffs(m_all) - 1;
            hi = __shfl_sync(full, x, jf);
            if (jf > 0) lo = __shfl_sync(full, x, jf - 1);
        }
    }
    const double a = glo, b = hi;
    const double c = 0.5 * (a + b), h = 0.5 * (b - a);
    const double z = tt * h;
    if (!(z <= LC_ZMAX)) return false;
    double *u0 = &s_u[0][1], *u1 = &s_u[1][1], *u2 = &s_u[2][1];
    for (int i = lane; i < n + 2; i += 32) s_u[0][i] = s_u[1][i] = s_u[2][i] = 0.0;
    __syncwarp();
    double y0 = 0.0, y1 = 0.0;  // this lane's entries lane and lane + 32 of the result (n <= 64)
    if (z < 0.5) {
        // Taylor series of exp(t (T - c I)) e1: ||t (T - c I)|| <= z < 1/2, 20 terms
        for (int i = lane; i < n; i += 32) {
            s_d[i] = tt * (s_a[i] - c);
            s_e[i] = tt * s_b[i];
        }
        if (lane == 0) u0[0] = 1.0;
        __syncwarp();
        y0 = lane == 0 ? 1.0 : 0.0;
        for (int k = 1; k <= 20; ++k) {
            const double rk = 1.0 / (double)k;
            for (int i = lane, q = 0; i < n; i += 32, ++q) {
                const double v = (s_d[i] * u0[i] + (i > 0 ? s_e[i - 1] : 0.0) * u0[i - 1] + s_e[i] * u0[i + 1]) * rk;
                u1[i] = v;
                if (q == 0) y0 += v; else y1 += v;
            }
            __syncwarp();
            double *tp = u0; u0 = u1; u1 = tp;
        }
        const double scale = exp(tt * c);
        y0 *= scale;
        y1 *= scale;
    } else {
        const int N = (int)(z + 12.0 * cbrt(z) + 25.0);  // <= 239
        if (lane == 0) {  // Miller: p_{k-1} = (2k / z) p_k + p_{k+1}, normalised by 1 = i_0 + 2 sum i_k
            const double tz = 2.0 / z;
            double pk1 = 0.0, pk = 1.0e-100, sum = 0.0;
            s_ck[N] = pk;
            for (int k = N; k >= 1; --k) {
                const double pm = fma(tz * (double)k, pk, pk1);
                sum += pk;
                pk1 = pk;
                pk = pm;
                s_ck[k - 1] = pm;
            }
            s_ck[N + 1] = 1.0 / (pk + 2.0 * sum);
        }
        const double rh = 1.0 / h;
        for (int i = lane; i < n; i += 32) {
            s_d[i] = (s_a[i] - c) * rh;
            s_e[i] = s_b[i] * rh;
        }
        __syncwarp();
        const double nrm = s_ck[N + 1];
        // u0 = e1, u1 = M e1
        if (lane == 0) u0[0] = 1.0;
        __syncwarp();
        for (int i = lane; i < n; i += 32)
            u1[i] = s_d[i] * u0[i] + (i > 0 ? s_e[i - 1] : 0.0) * u0[i - 1] + s_e[i] * u0[i + 1];
        __syncwarp();
        {
            const double c0 = s_ck[0] * nrm, c1 = 2.0 * s_ck[1] * nrm;
            y0 = (lane == 0 ? c0 : 0.0) + (lane < n ? c1 * u1[lane] : 0.0);
            y1 = lane + 32 < n ? c1 * u1[lane + 32] : 0.0;
        }
        for (int k = 2; k <= N; ++k) {
            const double ckk = 2.0 * s_ck[k] * nrm;
            for (int i = lane, q = 0; i < n; i += 32, ++q) {
                const double mv = s_d[i] * u1[i] + (i > 0 ? s_e[i - 1] : 0.0) * u1[i - 1] + s_e[i] * u1[i + 1];
                const double v = 2.0 * mv - u0[i];
                u2[i] = v;
                if (q == 0) y0 = fma(ckk, v, y0); else y1 = fma(ckk, v, y1);
            }
            __syncwarp();
            double *tp = u0; u0 = u1; u1 = u2; u2 = tp;
        }
        const double scale = exp(tt * b);
        y0 *= scale;
        y1 *= scale;
    }
    if (lane < n) yout[lane] = y0;
    if (lane + 32 < n) yout[lane + 32] = y1;
    for (int i = n + lane; i < ldy; i += 32) yout[i] = 0.0;
    return true;
}


__global__ void __launch_bounds__(SE_NT) small_exp_kernel(const SmallExpParams P) {
    extern __shared__ double sm[];
    __shared__ double sc[SE_MAXM];   // balancing scale factors
    __shared__ double colsum[SE_MAXM];
    const int prob = blockIdx.x;
    const int tid = threadIdx.x;
    SE_MARK(0);
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
    if (P.lanczos && !P.force_pade) {  // symmetric tridiagonal: Chebyshev on one warp (uniform per CTA)
        __shared__ int s_done;
        if (tid < 32) {
            const bool ok = se_lanczos_cheb(n, H, P.ldh, t, yout, P.ldy);
            if (tid == 0) s_done = ok ? 1 : 0;
        }
        __syncthreads();
        if (s_done) return;
    }
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
    SE_MARK(10);
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
