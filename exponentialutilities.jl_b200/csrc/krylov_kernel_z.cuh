// krylov_kernel_z.cuh -- persistent fused Arnoldi / Lanczos kernel for ComplexF64 bases.
//
// Same algorithm and team-barrier protocol as krylov_persistent_kernel (krylov_kernel.cuh), with complex
// arithmetic: conjugating inner products (dot(v_i, y), src/arnoldi.jl:302), Hermitian operators take the Lanczos
// recurrence with REAL coefficients (coeff(::Type{U<:Real}, alpha) = real(alpha), src/arnoldi.jl:412-413).
// A complex number is one 16-byte double2 {re, im}, so every basis access is a single vector load.  The mat-vec
// reads the operator with direct loads (4 lanes per row for short rows, a warp per row otherwise; dense: one row
// per thread x column groups) -- this path is about parity and API coverage first; the TMA ring of the fp64
// kernel is the obvious next step for it.  One problem per launch, no augmentation, no row sharding.
#pragma once
#include "krylov_kernel.cuh"

namespace b200k {

struct KrylovParamsZ {
    int op_kind;  // OP_CSR_WARP or OP_DENSE
    int n;
    const int *rowptr;
    const int *colind;
    const double2 *val;
    int lanes_per_row;  // CSR: 1 (rows of <= 8 entries, one row per thread), 4 or 32
    const double2 *Ad;
    long long lda;
    int team_size;
    int slice;
    const double2 *b;
    double2 *V;
    long long ldv;
    double2 *Hd;  // complex (m+1) x (m+1), column-major, ld = ldh
    int ldh;
    double *scal;
    int *stat;
    int m, j0, iop, lanczos;
    double tol;
    double2 *xbuf;   // [2][xlen]
    long long xlen;
    double2 *part;   // [2][MAXCOL][CPAD]
    double *partn;   // [4][CPAD]
    unsigned *bar;
    double2 *wglob;
    int w_in_smem;
    // krylov_tma_z_kernel (krylov_kernel_tma_z.cuh) only
    int nnz_cap;     // entries one ring slot holds (val 16 B | colind 4 B | rowptr segment)
    int ch_rows;     // rows per CSR chunk (<= 256: two consumer lanes per row)
    int nslot;       // ring depth
    int slot_bytes;  // bytes per ring slot (multiple of 128; >= one basis tile and >= one CSR chunk)
    int tile_rows;   // complex rows per basis tile (multiple of 16, <= 2048)
    int hintA_cols;  // operator chunks get L2::evict_first in steps whose window has >= this many columns
    // krylov_z_kernel launched behind the TMA instance: find the first step whose update removed most of the vector
    // from the stored H and redo the factorisation from there with the two-pass loop; exit at once if there is none
    int safe_scan;
};

struct __align__(16) SmemZ {
    int j0_found;
    int pad_[3];
    double2 hs[MAXCOL];
    double2 red[2][NW][CB];
    double redn[NW];
    double2 dscratch[NT];
};

__device__ __forceinline__ double2 zmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 zfma(double2 a, double2 b, double2 c) {  // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ double2 zcfma(double2 a, double2 b, double2 c) {  // conj(a)*b + c
    return make_double2(fma(a.x, b.x, fma(a.y, b.y, c.x)), fma(a.x, b.y, fma(-a.y, b.x, c.y)));
}
__device__ __forceinline__ double2 ldz(const double2 *p) {
    double2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ldz_cg(const double2 *p) {
    double2 r;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

__device__ __forceinline__ double2 team_sum_z(const double2 *pp, int C, int lane) {
    double2 s = make_double2(0.0, 0.0);
    for (int q = lane; q < C; q += 32) {
        const double2 v = ldz_cg(pp + q);
        s.x += v.x;
        s.y += v.y;
    }
    s.x = warp_sum(s.x);
    s.y = warp_sum(s.y);
    return s;
}

__device__ __forceinline__ void block_sum_to_z(SmemZ *S, int tid, int lane, int warp, double v, double *out) {
    v = warp_sum(v);
    if (lane == 0) S->redn[warp] = v;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += S->redn[w];
        *out = s;
    }
}

__global__ void __launch_bounds__(NT, 1) krylov_z_kernel(const __grid_constant__ KrylovParamsZ P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemZ *S = reinterpret_cast<SmemZ *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Team tm;
    tm.rank = blockIdx.x;
    tm.C = P.team_size;
    tm.bar = P.bar;
    tm.target = 0;
    const int n = P.n;
    const int r0 = min(n, tm.rank * P.slice);
    const int nrows = min(n, r0 + P.slice) - r0;
    double2 *ws = P.w_in_smem ? reinterpret_cast<double2 *>(smem_raw + sizeof(SmemZ)) : (P.wglob + r0);
    double2 *V = P.V;
    const long long ldv = P.ldv;
    double2 *xb0 = P.xbuf, *xb1 = P.xbuf + P.xlen;
    const double2 *xsrc;
    double xscale;
    int jstart;
    int m_out = P.m, breakdown = 0;

    int j0 = P.j0;
    if (P.safe_scan) {
        // H[j+1, j]^2 < eta^2 (||H[lo..j, j]||^2 + H[j+1, j]^2) (Pythagoras form of the DGKS test); columns the
        // fast instance never reached are zero and never trigger, NaN never triggers
        if (tid < 32) {
            const int iopw0 = P.iop > 0 ? P.iop : P.m;
            int first = 0x7fffffff;
            for (int jc = tid; jc < P.m; jc += 32) {
                double hsq = 0.0;
                for (int i = max(0, jc - iopw0 + 1); i <= jc; ++i) {
                    const double2 hv = P.Hd[(long long)jc * P.ldh + i];
                    hsq = fma(hv.x, hv.x, fma(hv.y, hv.y, hsq));
                }
                const double bj = P.Hd[(long long)jc * P.ldh + jc + 1].x;
                if (bj * bj < 0.0625 * (hsq + bj * bj)) first = min(first, jc + 1);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            if (tid == 0) S->j0_found = first == 0x7fffffff ? 0 : first;
        }
        __syncthreads();
        j0 = S->j0_found;
        if (j0 == 0) return;  // (every CTA takes the same decision: H is complete when this launch starts)
    }
    if (j0 == 0) {  // firststep!
        double nrm = 0.0;
        for (int i = tid; i < nrows; i += NT) {
            const double2 b1 = P.b[r0 + i];
            ws[i] = b1;
            nrm = fma(b1.x, b1.x, fma(b1.y, b1.y, nrm));
        }
        block_sum_to_z(S, tid, lane, warp, nrm, P.partn + 2 * CPAD + tm.rank);
        team_barrier(tm);
        const double beta = sqrt(team_sum(P.partn + 2 * CPAD, tm.C, lane));
        if (tm.rank == 0 && tid == 0) P.scal[0] = beta;
        if (beta == 0.0) {
            if (tm.rank == 0 && tid == 0) {
                P.stat[0] = P.m;
                P.stat[1] = 0;
            }
            return;
        }
        const double inv = 1.0 / beta;
        for (int i = tid; i < nrows; i += NT) {
            double2 b1 = ws[i];
            b1.x *= inv;
            b1.y *= inv;
            V[r0 + i] = b1;
        }
        __syncthreads();
        xsrc = P.b;
        xscale = inv;
        jstart = 1;
    } else {
        xsrc = V + (long long)(j0 - 1) * ldv;
        xscale = 1.0;
        jstart = j0;
    }

    double beta_prev = 0.0;
    const int iopw = P.iop > 0 ? P.iop : P.m;
    for (int j = jstart; j <= P.m; ++j) {
        const int jc = j - 1;
        const int par = j & 1;
        double2 *xout = par ? xb1 : xb0;
        double2 *part = P.part + (long long)par * MAXCOL * CPAD;
        double *partn = P.partn + par * CPAD;

        // ---- mat-vec: ws = xscale * (A x)[slice] ----
        if (P.op_kind == OP_CSR_WARP && P.lanes_per_row == 1) {
            // short rows (<= 8 entries): one row per thread, all gathers of the row in flight before the first FMA
            for (int rl = tid; rl < nrows; rl += NT) {
                const int e0 = P.rowptr[r0 + rl], e1 = P.rowptr[r0 + rl + 1];
                double2 av[8], xv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool ok = e0 + u < e1;
                    av[u] = make_double2(0.0, 0.0);
                    xv[u] = make_double2(0.0, 0.0);
                    if (ok) {
                        av[u] = P.val[e0 + u];
                        xv[u] = xsrc[P.colind[e0 + u]];
                    }
                }
                double2 sum = make_double2(0.0, 0.0);
#pragma unroll
                for (int u = 0; u < 8; ++u) sum = zfma(av[u], xv[u], sum);
                ws[rl] = make_double2(sum.x * xscale, sum.y * xscale);
            }
        } else if (P.op_kind == OP_CSR_WARP) {
            const int G = P.lanes_per_row;  // 4 or 32
            const int gl = lane % G, gi = lane / G, rows_per_it = NW * (32 / G);
            for (int rbase = 0; rbase < nrows; rbase += rows_per_it) {
                const int rl = rbase + warp * (32 / G) + gi;
                double2 sum = make_double2(0.0, 0.0);
                if (rl < nrows) {
                    const int e0 = P.rowptr[r0 + rl], e1 = P.rowptr[r0 + rl + 1];
                    for (int e = e0 + gl; e < e1; e += G) sum = zfma(P.val[e], xsrc[P.colind[e]], sum);
                }
                for (int o = G >> 1; o > 0; o >>= 1) {
                    sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
                    sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
                }
                if (rl < nrows && gl == 0) ws[rl] = make_double2(sum.x * xscale, sum.y * xscale);
            }
        } else {  // dense column-major
            int RL = 32;
            while (RL < nrows && RL < NT) RL <<= 1;
            const int Gc = NT / RL;
            const int ul = tid % RL, g = tid / RL;
            for (int ubase = 0; ubase < nrows; ubase += RL) {
                const int u = ubase + ul;
                double2 acc = make_double2(0.0, 0.0);
                if (u < nrows) {
                    const double2 *ap = P.Ad + r0 + u;
#pragma unroll 4
                    for (int c = g; c < n; c += Gc) acc = zfma(ldz(ap + (long long)c * P.lda), xsrc[c], acc);
                }
                S->dscratch[g * RL + ul] = acc;
                __syncthreads();
                if (g == 0 && u < nrows) {
                    double2 s = make_double2(0.0, 0.0);
                    for (int q = 0; q < Gc; ++q) {
                        s.x += S->dscratch[q * RL + ul].x;
                        s.y += S->dscratch[q * RL + ul].y;
                    }
                    ws[u] = make_double2(s.x * xscale, s.y * xscale);
                }
                __syncthreads();
            }
        }
        __syncthreads();

        // ---- inner products h_i = <v_i, w> = sum conj(v_i) w over the window, update, norm -- at most twice:
        // the second classical Gram-Schmidt pass runs when ||w_after|| < ||w_before|| / 4 (DGKS re-orthogonalisation,
        // see krylov_kernel_tma.cuh) and adds its coefficients to H ----
        const int lo = P.lanczos ? jc : max(0, jc - iopw + 1);
        const int hi = jc;
        const int nc = hi - lo + 1;
        const int ulo = (P.lanczos && jc >= 1) ? jc - 1 : lo;
        const bool dgks = !P.lanczos;
        double beta2 = 0.0, wsq_before = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            int batch = 0;
            for (int cb = lo; cb <= hi; cb += CB, ++batch) {
                const int nb = min(CB, hi - cb + 1);
                double are[CB], aim[CB];
#pragma unroll
                for (int u = 0; u < CB; ++u) are[u] = aim[u] = 0.0;
                const double2 *vb = V + (long long)cb * ldv + r0;
                for (int i = tid; i < nrows; i += NT) {
                    const double2 w1 = ws[i];
#pragma unroll
                    for (int u = 0; u < CB; ++u)
                        if (u < nb) {
                            const double2 v1 = ldz(vb + (long long)u * ldv + i);
                            are[u] = fma(v1.x, w1.x, fma(v1.y, w1.y, are[u]));
                            aim[u] = fma(v1.x, w1.y, fma(-v1.y, w1.x, aim[u]));
                        }
                }
                const double rr = warp_reduce8(are, lane);
                const double ri = warp_reduce8(aim, lane);
                const int buf = batch & 1;
                if ((lane & 3) == 0)
                    S->red[buf][warp][((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = make_double2(rr, ri);
                __syncthreads();
                if (tid < nb) {
                    double2 s = make_double2(0.0, 0.0);
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        s.x += S->red[buf][w][tid].x;
                        s.y += S->red[buf][w][tid].y;
                    }
                    part[(long long)(cb - lo + tid) * CPAD + tm.rank] = s;
                }
            }
            team_barrier(tm);

            for (int ci = warp; ci < nc; ci += NW) {
                double2 s = team_sum_z(part + (long long)ci * CPAD, tm.C, lane);
                if (P.lanczos) s.y = 0.0;  // coeff(U <: Real, alpha) = real(alpha)
                if (lane == 0) {
                    if (pass == 0) {
                        S->hs[lo + ci - ulo] = s;
                        if (tm.rank == 0) P.Hd[(long long)jc * P.ldh + lo + ci] = s;
                    } else {
                        if (tm.rank == 0) {
                            const double2 h1 = S->hs[ci];
                            P.Hd[(long long)jc * P.ldh + lo + ci] = make_double2(h1.x + s.x, h1.y + s.y);
                        }
                        S->hs[ci] = s;
                    }
                }
            }
            if (P.lanczos && jc >= 1 && tid == 0) S->hs[0] = make_double2(beta_prev, 0.0);
            __syncthreads();
            if (dgks && pass == 0)  // ||w_before||^2 = ||h||^2 + ||w_after||^2 (orthonormal window)
                for (int ci = 0; ci < nc; ++ci) wsq_before += S->hs[ci].x * S->hs[ci].x + S->hs[ci].y * S->hs[ci].y;

            // ---- update w -= sum_c h_c v_c (reverse order), partial ||w||^2, publish the unnormalised w ----
            double nrm = 0.0;
            for (int i = tid; i < nrows; i += NT) {
                double2 w1 = ws[i];
                for (int c = hi; c >= ulo; --c) {
                    const double2 hc = S->hs[c - ulo];
                    const double2 v1 = ldz(V + (long long)c * ldv + r0 + i);
                    w1 = zfma(make_double2(-hc.x, -hc.y), v1, w1);
                }
                ws[i] = w1;
                xout[r0 + i] = w1;
                nrm = fma(w1.x, w1.x, fma(w1.y, w1.y, nrm));
            }
            block_sum_to_z(S, tid, lane, warp, nrm, partn + tm.rank);
            team_barrier(tm);
            beta2 = team_sum(partn, tm.C, lane);
            if (!(dgks && pass == 0 && beta2 < 0.0625 * (wsq_before + beta2))) break;
            __syncthreads();
        }

        const double beta = sqrt(beta2);
        if (tm.rank == 0 && tid == 0) P.Hd[(long long)jc * P.ldh + jc + 1] = make_double2(beta, 0.0);
        {
            double2 *vn = V + (long long)(jc + 1) * ldv;
            for (int i = tid; i < nrows; i += NT) {
                double2 w1 = ws[i];
                w1.x /= beta;
                w1.y /= beta;
                vn[r0 + i] = w1;
            }
        }
        __syncthreads();
        xsrc = xout;
        xscale = 1.0 / beta;
        beta_prev = beta;
        if (beta < P.tol) {
            m_out = j;
            breakdown = 1;
            break;
        }
    }
    if (tm.rank == 0 && tid == 0) {
        P.stat[0] = m_out;
        P.stat[1] = breakdown;
    }
}

// w (complex n) = beta * V[:, 0:m] * y (complex m)
__global__ void project_z_kernel(const double2 *V, long long ldv, long long n, int m, double beta, const double2 *y,
                                 double2 *w) {
    __shared__ double2 ys[MAXCOL];
    for (int i = threadIdx.x; i < m; i += blockDim.x) ys[i] = y[i];
    __syncthreads();
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll 4
        for (int i = 0; i < m; ++i) acc = zfma(ldz(V + (long long)i * ldv + r), ys[i], acc);
        w[r] = beta == 0.0 ? make_double2(0.0, 0.0) : make_double2(beta * acc.x, beta * acc.y);
    }
}

// ishermitian(A) (A == A') and opnorm(A, Inf) for a complex CSR operator
__global__ void csr_analyze_z_kernel(int n, const int *rowptr, const int *colind, const double2 *val, int *max_row_nnz,
                                     int *nonherm, double *norminf) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const int e0 = rowptr[r], e1 = rowptr[r + 1];
        atomicMax(max_row_nnz, e1 - e0);
        double s = 0.0;
        for (int e = e0; e < e1; ++e) {
            const double2 v = val[e];
            s += hypot(v.x, v.y);
            const int c = colind[e];
            if (c == r) {
                if (v.y != 0.0) *nonherm = 1;
                continue;
            }
            double2 vt = make_double2(0.0, 0.0);
            for (int f = rowptr[c]; f < rowptr[c + 1]; ++f)
                if (colind[f] == r) {
                    vt.x += val[f].x;
                    vt.y += val[f].y;
                }
            if (!(v.x == vt.x && v.y == -vt.y)) *nonherm = 1;
        }
        atomicMax(reinterpret_cast<unsigned long long *>(norminf), (unsigned long long)__double_as_longlong(s));
    }
}

__global__ void dense_analyze_z_kernel(int n, const double2 *A, long long lda, int *nonherm, double *rowsum) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        double s = 0.0;
        bool bad = false;
        for (int c = 0; c < n; ++c) {
            const double2 v = A[(long long)c * lda + r];
            s += hypot(v.x, v.y);
            const double2 vt = A[(long long)r * lda + c];
            if (!(v.x == vt.x && v.y == -vt.y)) bad = true;
        }
        if (bad) *nonherm = 1;
        atomicMax(reinterpret_cast<unsigned long long *>(rowsum), (unsigned long long)__double_as_longlong(s));
    }
}

}  // namespace b200k
