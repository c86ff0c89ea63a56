// aux_kernels.cuh -- kernels around the persistent Krylov kernel:
//   * project_kernel:     w = beta * V[:, 1:m] * Y (+ last-vector correction)   krylov_phiv.jl:229,242-244,640-650
//   * operator ingestion: index rebasing, row statistics, ishermitian(A), opnorm(A, Inf)
//   * plain mat-vec (mul!) for b200k_op_apply
//   * small helpers for kiops (column flip/scale, 1-norm)
#pragma once
#include "ptx.cuh"

namespace b200k {

constexpr int PROJ_NT = 256;
constexpr int PROJ_NC = 8;     // output columns per pass
constexpr int PROJ_MAXM = 256;

struct ProjectParams {
    const double *V;
    long long ldv;
    long long V_stride;  // per problem (batched)
    long long nrows;
    const double *Y;  // device: per problem ldy x nc, column-major
    int ldy;
    long long Y_stride;
    const int *mvec;       // per-problem m (device) or nullptr -> m
    const double *betavec; // per-problem beta (device) or nullptr -> beta
    int m;
    double beta;
    int nc;
    double *W;
    long long ldw;
    long long W_stride;
    const double *corr;  // nc coefficients of v_{m+1} (already multiplied by beta*h*t), or nullptr
    int vec2;
};

// grid = (row tiles, ceil(nc / PROJ_NC), nprob)
template <int NCT>
__global__ void __launch_bounds__(PROJ_NT) project_kernel_t(const ProjectParams P) {
    __shared__ double Ys[NCT][PROJ_MAXM];
    __shared__ double cs[NCT];
    const int prob = blockIdx.z;
    const int c0 = blockIdx.y * NCT;
    const int ncl = min(NCT, P.nc - c0);
    const int m = P.mvec ? P.mvec[prob] : P.m;
    const double beta = P.betavec ? P.betavec[prob] : P.beta;
    const double *V = P.V + (long long)prob * P.V_stride;
    double *W = P.W + (long long)prob * P.W_stride;
    const double *Y = P.Y + (long long)prob * P.Y_stride;
    for (int idx = threadIdx.x; idx < NCT * m; idx += PROJ_NT) {
        const int c = idx / m, i = idx % m;
        Ys[c][i] = (c < ncl) ? Y[(long long)(c0 + c) * P.ldy + i] : 0.0;
    }
    if (threadIdx.x < NCT)
        cs[threadIdx.x] = (P.corr && threadIdx.x < ncl) ? P.corr[c0 + threadIdx.x] : 0.0;
    __syncthreads();
    const bool has_corr = P.corr != nullptr;
    const double *vlast = V + (long long)m * P.ldv;
    if (beta == 0.0) {  // expv!: beta == 0 -> w .= 0, V may be uninitialised (krylov_phiv.jl:206-213)
        for (long long r = (long long)blockIdx.x * PROJ_NT + threadIdx.x; r < P.nrows;
             r += (long long)gridDim.x * PROJ_NT)
            for (int c = 0; c < ncl; ++c) W[(long long)(c0 + c) * P.ldw + r] = 0.0;
        return;
    }
    if (P.vec2) {
        const long long units = P.nrows >> 1;
        for (long long u = (long long)blockIdx.x * PROJ_NT + threadIdx.x; u < units;
             u += (long long)gridDim.x * PROJ_NT) {
            double2 acc[NCT];
#pragma unroll
            for (int c = 0; c < NCT; ++c) acc[c] = make_double2(0.0, 0.0);
            const double *vp = V + 2 * u;
            constexpr int LB = NCT == 1 ? 10 : 4;  // loads in flight per thread
            for (int ib = 0; ib < m; ib += LB) {
                double2 v2[LB];
#pragma unroll
                for (int u = 0; u < LB; ++u)
                    if (ib + u < m) v2[u] = ld_stream2(vp + (long long)(ib + u) * P.ldv);
#pragma unroll
                for (int u = 0; u < LB; ++u)
                    if (ib + u < m) {
#pragma unroll
                        for (int c = 0; c < NCT; ++c) {
                            acc[c].x = fma(v2[u].x, Ys[c][ib + u], acc[c].x);
                            acc[c].y = fma(v2[u].y, Ys[c][ib + u], acc[c].y);
                        }
                    }
            }
            double2 vl = make_double2(0.0, 0.0);
            if (has_corr) vl = ld_stream2(vlast + 2 * u);
#pragma unroll
            for (int c = 0; c < NCT; ++c)
                if (c < ncl) {
                    double2 o;
                    o.x = beta * acc[c].x;
                    o.y = beta * acc[c].y;
                    if (has_corr) {
                        o.x = fma(cs[c], vl.x, o.x);
                        o.y = fma(cs[c], vl.y, o.y);
                    }
                    *reinterpret_cast<double2 *>(W + (long long)(c0 + c) * P.ldw + 2 * u) = o;
                }
        }
    } else {
        for (long long r = (long long)blockIdx.x * PROJ_NT + threadIdx.x; r < P.nrows;
             r += (long long)gridDim.x * PROJ_NT) {
            double acc[NCT];
#pragma unroll
            for (int c = 0; c < NCT; ++c) acc[c] = 0.0;
            const double *vp = V + r;
#pragma unroll 4
            for (int i = 0; i < m; ++i) {
                const double v1 = ld_stream1(vp + (long long)i * P.ldv);
#pragma unroll
                for (int c = 0; c < NCT; ++c) acc[c] = fma(v1, Ys[c][i], acc[c]);
            }
            const double vl = has_corr ? vlast[r] : 0.0;
#pragma unroll
            for (int c = 0; c < NCT; ++c)
                if (c < ncl) {
                    double o = beta * acc[c];
                    if (has_corr) o = fma(cs[c], vl, o);
                    W[(long long)(c0 + c) * P.ldw + r] = o;
                }
        }
    }
}

// Launch helper: one output column (expv, kiops) uses the specialised instance.
inline void launch_project_kernel(const ProjectParams &P, dim3 grid, cudaStream_t stream) {
    if (P.nc == 1) project_kernel_t<1><<<grid, PROJ_NT, 0, stream>>>(P);
    else project_kernel_t<PROJ_NC><<<grid, PROJ_NT, 0, stream>>>(P);
}

// ---- operator ingestion ---------------------------------------------------------------------------
__global__ void rebase_kernel(const int *in, int *out, long long count, int base) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = in[i] - base;
}

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
    // non-negative doubles order like their bit patterns; NaN maps above +Inf and sticks.
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// stats[0] = max row nnz (as int), flags[0] = 1 if a non-Hermitian entry was found, norm[0] = opnorm(A, Inf).
__global__ void csr_analyze_kernel(int n, const int *rowptr, const int *colind, const double *val,
                                   int *max_row_nnz, int *nonsym, double *norminf) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const int e0 = rowptr[r], e1 = rowptr[r + 1];
        atomicMax(max_row_nnz, e1 - e0);
        double s = 0.0;
        for (int e = e0; e < e1; ++e) {
            const double v = val[e];
            s += fabs(v);
            const int c = colind[e];
            if (c == r || c >= n) continue;  // c >= n: halo column of a row-sharded block
            // find the transposed entry (c, r): sum duplicates, absent means 0
            double vt = 0.0;
            const int f0 = rowptr[c], f1 = rowptr[c + 1];
            for (int f = f0; f < f1; ++f)
                if (colind[f] == r) vt += val[f];
            // duplicates of (r, c) in this row: compare the summed value
            double vs = 0.0;
            for (int f = e0; f < e1; ++f)
                if (colind[f] == c) vs += val[f];
            if (!(vs == vt)) *nonsym = 1;
        }
        atomic_max_nonneg(norminf, s);
    }
}

__global__ void dense_analyze_kernel(int n, const double *A, long long lda, int *nonsym, double *rowsum) {
    // one thread per row: row abs-sum and symmetry against the transposed entries
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        double s = 0.0;
        bool bad = false;
        for (int c = 0; c < n; ++c) {
            const double v = A[(long long)c * lda + r];
            s += fabs(v);
            if (c > r && !(v == A[(long long)r * lda + c])) bad = true;
        }
        if (bad) *nonsym = 1;
        atomic_max_nonneg(rowsum, s);
    }
}

// ---- mul!(y, A, x) ----------------------------------------------------------------------------------
__global__ void csr_apply_kernel(int n, const int *rowptr, const int *colind, const double *val, const double *x,
                                 double *y) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n; r += gridDim.x * wpb) {
        double s = 0.0;
        for (int e = rowptr[r] + lane; e < rowptr[r + 1]; e += 32) s = fma(val[e], x[colind[e]], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[r] = s;
    }
}

__global__ void dense_apply_kernel(int n, const double *A, long long lda, const double *x, double *y) {
    // one CTA per 64-row tile, 4 column groups of 64 threads
    __shared__ double red[4][64];
    const int rl = threadIdx.x & 63, g = threadIdx.x >> 6;
    for (int tile = blockIdx.x; tile * 64 < n; tile += gridDim.x) {
        const int r = tile * 64 + rl;
        double s = 0.0;
        if (r < n)
            for (int c = g; c < n; c += 4) s = fma(A[(long long)c * lda + r], x[c], s);
        red[g][rl] = s;
        __syncthreads();
        if (g == 0 && r < n) y[r] = red[0][rl] + red[1][rl] + red[2][rl] + red[3][rl];
        __syncthreads();
    }
}

// ---- kiops helpers ------------------------------------------------------------------------------------
// out[:, k] = nu * U[:, ppo-1-k], k = 0..p-1   (u_flip = nu * reverse(u[:, 2:end], dims = 2), kiops.jl:104-106)
__global__ void flip_scale_kernel(long long n, int p, const double *U, long long ldu, double nu, double *out,
                                  long long ldo) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * p;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i % n;
        const int k = (int)(i / n);
        out[(long long)k * ldo + r] = nu * U[(long long)(p - k) * ldu + r];
    }
}

// per-block partial sums of |x| over a strided n x ncol block; finished on the host (deterministic).
__global__ void abs_sum_kernel(long long n, int ncol, const double *X, long long ldx, double *partial) {
    __shared__ double red[32];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * ncol;
         i += (long long)gridDim.x * blockDim.x)
        s += fabs(X[(i / n) * ldx + (i % n)]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}

// y += a * x  (axpy!), y = a * x (lmul!(a, copyto!(y, x))), per-block max |x| (norm(x, Inf))
__global__ void axpy_kernel(long long n, double a, const double *x, double *y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        y[i] = fma(a, x[i], y[i]);
}
__global__ void scalecopy_kernel(long long n, double a, const double *x, double *y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        y[i] = a * x[i];
}
__global__ void abs_max_kernel(long long n, const double *x, double *partial) {
    __shared__ double red[32];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        s = fmax(s, fabs(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmax(t, red[w]);
        partial[blockIdx.x] = t;
    }
}

__global__ void scale_kernel(long long n, double *x, double a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        x[i] *= a;
}

}  // namespace b200k
