/*
 * b200krylov.h -- C ABI of the B200-native Krylov expmv / phiv engine.
 *
 * Drop-in boundary for the hot path of SciML/ExponentialUtilities.jl v1.35.0
 * (arnoldi!/lanczos! -> expv/phiv -> kiops).  Every entry point below names the reference
 * interface it replaces (paths relative to the reference repository root).  A Julia host binds
 * these with `ccall`; the Python host in
 * exponentialutilities.jl_b200/ binds them with ctypes.
 *
 * Conventions
 *   - Plain C: pointers, sizes, doubles.  No C++ / torch types cross this boundary.
 *   - "device" pointers are CUDA device pointers owned by the caller; "host" pointers are ordinary
 *     host memory.  Matrices are column-major with an explicit leading dimension (Julia layout).
 *   - Every function returns a b200k_status (0 = OK).  Nothing throws across the ABI.  Happy
 *     breakdown and beta == 0 are NOT errors (reference: src/arnoldi.jl:370-374,
 *     src/krylov_phiv.jl:206-213); they are reported through out-parameters.
 *   - One handle <-> one CUDA stream <-> one host thread at a time.
 *   - There is no CPU fallback: without a CUDA device b200k_create fails with B200K_ECUDA.
 */
#ifndef B200KRYLOV_H
#define B200KRYLOV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200K_VERSION 200 /* 0.2.0 */

typedef enum {
    B200K_OK = 0,
    B200K_EDIM = 1,         /* DimensionMismatch / "Dimension mismatch" asserts (arnoldi.jl:217, krylov_phiv.jl:205,625) */
    B200K_EARG = 2,         /* ArgumentError (krylov_phiv.jl:132,221,630) */
    B200K_ESINGULAR = 3,    /* SingularException(0) from the Pade solve (exp_baseexp.jl:55) */
    B200K_ECUDA = 4,        /* CUDA runtime failure, or no device */
    B200K_ECOMM = 5,        /* multi-GPU communicator failure */
    B200K_EUNSUPPORTED = 6, /* valid in the reference but not implemented here */
    B200K_ENOMEM = 7
} b200k_status;

typedef struct b200k_context *b200k_handle_t;
typedef struct b200k_operator *b200k_op_t;

/* ---- library / handle ------------------------------------------------------------------- */
int b200k_version(void);
/* sizeof() of the option structs as THIS library was compiled, so that a foreign-language binding can verify its own
 * struct definitions once at load time (a Julia `struct` that lacks a trailing field would otherwise make the library
 * read past it).  which: B200K_STRUCT_*; returns -1 for an unknown id. */
#define B200K_STRUCT_KRYLOV_OPTS 1
#define B200K_STRUCT_KIOPS_OPTS 2
#define B200K_STRUCT_TIMESTEP_OPTS 3
int b200k_sizeof(int which);
/* Static description of an error code (never NULL). */
const char *b200k_status_string(int status);
/* device: CUDA ordinal; stream: a cudaStream_t (CUstream) or NULL for the legacy default stream. */
int b200k_create(b200k_handle_t *h, int device, void *stream);
int b200k_destroy(b200k_handle_t h);
int b200k_set_stream(b200k_handle_t h, void *stream);
int b200k_synchronize(b200k_handle_t h);
/* Message of the last failure on this handle ("" if none). */
const char *b200k_last_error(b200k_handle_t h);
/* Number of SMs / CTAs the persistent kernel uses on this device, and kernels launched so far. */
int b200k_device_info(b200k_handle_t h, int *sm_count, int *max_team, int64_t *launches);

/* ---- operators: the reference's operator interface (docs/src/interfaces.md:9-36) ------------
 * NOT offered: matrix-free operators (a `mul!` callback, SURVEY.md 8b `b200k_op_callback`).  The whole design rests on
 * fusing the mat-vec into the persistent Krylov kernel (the operator is streamed by the kernel's own TMA producer);
 * a host- or device-side callback per step would break that fusion and fall back to one launch per step -- the
 * reference's generic path already does that.  Concrete CSR and dense operators, real and ComplexF64, are covered. */
/*
 * size / eltype / mul! / ishermitian / opnorm of a concrete fp64 matrix.  The arrays are copied
 * into library-owned, 0-based, padded device storage ("operator ingestion"), so the caller may
 * free its copies afterwards.  `location`: 0 = the pointers are device pointers, 1 = host.
 * `index_base`: 0 (C/SciPy) or 1 (Julia CuSparseMatrixCSR). */
int b200k_op_csr_create(b200k_handle_t h, int64_t n, int64_t nnz, const int32_t *rowptr,
                        const int32_t *colind, const double *val, int index_base, int location,
                        b200k_op_t *op);
/* Dense column-major n x n with leading dimension lda (mul!(y, A::Matrix, x), arnoldi.jl:185). */
int b200k_op_dense_create(b200k_handle_t h, int64_t n, const double *A, int64_t lda, int location,
                          b200k_op_t *op);
int b200k_op_destroy(b200k_op_t op);
/* size(A,1), nnz, kind (0 CSR, 1 dense), LinearAlgebra.ishermitian(A), opnorm(A, Inf). */
int b200k_op_info(b200k_op_t op, int64_t *n, int64_t *nnz, int *kind, int *is_hermitian,
                  double *opnorm_inf);
/* y = A x on device vectors (mul!; used by tests and the matrix-free style callers). */
int b200k_op_apply(b200k_handle_t h, b200k_op_t op, const double *x, double *y);

/* ---- Krylov factorisation: arnoldi! / lanczos! (src/arnoldi.jl:345-377, 456-490) -----------
 * Keyword arguments of the reference, same names and defaults. */
typedef struct {
    int m;         /* requested Krylov dimension (default min(maxiter, n))               */
    double tol;    /* happy-breakdown threshold, absolute test beta_j < tol (1e-7)        */
    int iop;       /* incomplete-orthogonalisation length, 0 = full Arnoldi              */
    int hermitian; /* 1: lanczos!, 0: arnoldi!, -1: LinearAlgebra.ishermitian(A)         */
    int init;      /* continue an existing factorisation from column init (0 = start)    */
    /* augmented operator (A, B) of kiops (arnoldi.jl:191-205, 257-279); p = 0: plain    */
    int p;             /* number of augmented rows                                       */
    const double *B;   /* device, n x p column-major (u_flip)                            */
    int64_t ldb;
    double t;          /* tau_now                                                        */
    double mu;
} b200k_krylov_opts;
void b200k_krylov_opts_default(b200k_krylov_opts *o);

/* arnoldi!(Ks, A, b; ...).  KrylovSubspace fields are passed piecewise (arnoldi.jl:50-61):
 *   V      device, (n + p) x (maxiter + 1) column-major, leading dimension ldv   (Ks.V)
 *   H      HOST,   (maxiter + 1) x (maxiter + [p != 0]) column-major, ld ldh     (Ks.H)
 *   beta   in (init != 0) / out                                                  (Ks.beta)
 *   m_out, breakdown                                                             (Ks.m, Ks.wasbreakdown)
 * b: device vector of length n (for the augmented form: the column w[:, l]).
 * Requires opts->m <= maxiter (the host-side KrylovSubspace does the resize!, arnoldi.jl:357).
 * Orthogonalisation is classical Gram-Schmidt fused with the mat-vec (the reference uses
 * sequential modified Gram-Schmidt, arnoldi.jl:301-304); results agree to rounding, see DESIGN.md.
 * Returns after H, beta, m_out are valid on the host. */
int b200k_arnoldi(b200k_handle_t h, b200k_op_t op, const double *b, const b200k_krylov_opts *opts,
                  double *V, int64_t ldv, int maxiter, double *H, int ldh, double *beta,
                  int *m_out, int *breakdown);

/* expv!(w, t, Ks) (src/krylov_phiv.jl:200-247): w = beta * V[:, 1:m] * exp(t H[1:m,1:m]) e1.
 * Exactly-symmetric H -> symmetric-tridiagonal eigen branch (:225-229), else Higham-2005 Pade.
 * nrows = size(V, 1) (n + p).  w: device, nrows. */
int b200k_expv_ks(b200k_handle_t h, double t, const double *V, int64_t ldv, int64_t nrows,
                  const double *H, int ldh, int m, double beta, double *w);

/* _phiv!(w, t, Ks, k, cache, correct) (src/krylov_phiv.jl:620-653).  W: device nrows x (k+1),
 * leading dimension ldw.  errest (may be NULL) receives |beta * h_{m+1,m} * t * C2[m, k+1]|.
 * H must hold rows 1..m+1 (h_{m+1,m} is read). */
int b200k_phiv_ks(b200k_handle_t h, double t, const double *V, int64_t ldv, int64_t nrows,
                  const double *H, int ldh, int m, double beta, int k, int correct, double *W,
                  int64_t ldw, double *errest);

/* expv(t, A, b; m, tol, iop, ishermitian) one-shot (src/krylov_phiv.jl:125-144): arnoldi + expv!
 * with library-owned Krylov storage.  b, w: device vectors of length n.  m_out / breakdown /
 * beta_out may be NULL. */
int b200k_expv(b200k_handle_t h, b200k_op_t op, double t, const double *b,
               const b200k_krylov_opts *opts, double *w, int *m_out, int *breakdown,
               double *beta_out);
/* expv(t, A, b; mode = :error_estimate) (src/krylov_phiv.jl:145-160, src/krylov_phiv_error_estimate.jl:149-207):
 * Lanczos with Saad's a-posteriori estimate sigma_j = beta_j * beta * |e_j' exp(t T_j) e_1| and the stop
 * sigma_j < atol + rtol * beta.  Hermitian operators only (the reference raises otherwise -> B200K_EUNSUPPORTED).
 * The m-step Lanczos factorisation runs in one launch (its coefficients do not depend on the stopping index);
 * the estimates are then evaluated for j = 1, 2, ... on the host and the subspace is truncated at the first hit,
 * so w and m_out equal the reference's. */
int b200k_expv_ee(b200k_handle_t h, b200k_op_t op, double t, const double *b, int m, double atol, double rtol,
                  double *w, int *m_out);
/* Same call with HOST vectors: copies b in and w out on the handle's stream (end-to-end path). */
int b200k_expv_host(b200k_handle_t h, b200k_op_t op, double t, const double *b_host,
                    const b200k_krylov_opts *opts, double *w_host, int *m_out, int *breakdown);
/* The same without the final synchronisation: everything (H2D of b, the three kernels, D2H of w) is queued on the
 * handle's stream and the call returns; w_host is valid after b200k_synchronize(h), which also reports a deferred
 * SingularException.  b_host / w_host should be pinned and must stay untouched until then.  A server keeps two
 * handles on two streams in flight so that the copies of one request overlap the kernels of the other (an operator
 * may be used through any handle of its device, one call at a time). */
int b200k_expv_host_async(b200k_handle_t h, b200k_op_t op, double t, const double *b_host,
                          const b200k_krylov_opts *opts, double *w_host);
/* phiv(t, A, b, k; correct, errest) one-shot (src/krylov_phiv.jl:563-570). */
int b200k_phiv(b200k_handle_t h, b200k_op_t op, double t, const double *b, int k,
               const b200k_krylov_opts *opts, int correct, double *W, int64_t ldw, double *errest,
               int *m_out, int *breakdown);

/* nb independent expv(t_i, A, b_i) on one shared operator, one launch (batched replicas of
 * krylov_phiv.jl:125-144).  t: HOST nb; Bv, W: device n x nb column-major.  m_out/breakdown: HOST
 * arrays of nb ints or NULL. */
int b200k_expv_batched(b200k_handle_t h, b200k_op_t op, int nb, const double *t, const double *Bv,
                       int64_t ldbv, const b200k_krylov_opts *opts, double *W, int64_t ldw,
                       int *m_out, int *breakdown);

/* ---- kiops (src/kiops.jl:57-326) ------------------------------------------------------------ */
typedef struct {
    int mmin;       /* 10  */
    int mmax;       /* 128 */
    int m;          /* min(mmin, mmax) */
    double tol;     /* 1e-7 */
    int iop;        /* 2 */
    int hermitian;  /* -1: ishermitian(A) */
    int task1;      /* false */
    double opnorm;  /* accepted and unused, as in the reference */
    double normU;   /* norm(u[:, 2:end], 1); NaN = computed by the library (row-sharded callers pass the global sum) */
} b200k_kiops_opts;
void b200k_kiops_opts_default(b200k_kiops_opts *o);
/* tau_out: HOST, ntau values; numSteps follows the reference (`size(tau_out, 2)`): pass
 * tau_is_row = 1 for a 1 x ntau row (numSteps = ntau), 0 for a scalar / column (numSteps = 1).
 * U: device n x ppo (u, ppo = p + 1 columns).  W: device n x numSteps.  stats: HOST int64[5] =
 * (steps, rejected, krylov_steps(always 0), exps, m). */
int b200k_kiops(b200k_handle_t h, b200k_op_t op, int ntau, const double *tau_out, int tau_is_row,
                const double *U, int64_t ldu, int ppo, const b200k_kiops_opts *opts, double *W,
                int64_t ldw, int64_t *stats);

/* ---- ComplexF64 element types and complex t (src/arnoldi.jl:412-421, src/krylov_phiv.jl:252-280) ----
 * Complex data is passed as interleaved (re, im) doubles, i.e. Julia's Vector/Matrix{ComplexF64} memory; leading
 * dimensions and lengths count COMPLEX elements.  H is always complex here (for a Hermitian operator the Lanczos
 * coefficients are real: the imaginary parts are zero and the host mirror stores the real parts, as the reference's
 * KrylovSubspace{T, U = real(T)} does).  Not available for the augmented (kiops) operator or row sharding. */
int b200k_op_csr_create_z(b200k_handle_t h, int64_t n, int64_t nnz, const int32_t *rowptr, const int32_t *colind,
                          const double *val, int index_base, int location, b200k_op_t *op);
int b200k_op_dense_create_z(b200k_handle_t h, int64_t n, const double *A, int64_t lda, int location, b200k_op_t *op);
/* arnoldi!/lanczos! on a complex basis (b: device, n complex values; V: device, complex, ldv; H: HOST, complex, ldh). */
int b200k_arnoldi_z(b200k_handle_t h, b200k_op_t op, const double *b, const b200k_krylov_opts *opts, double *V,
                    int64_t ldv, int maxiter, double *H, int ldh, double *beta, int *m_out, int *breakdown);
/* expv!(w, t, Ks) with complex w and real or complex t = t_re + i t_im (krylov_phiv.jl:200-280). */
int b200k_expv_ks_z(b200k_handle_t h, double t_re, double t_im, const double *V, int64_t ldv, int64_t nrows,
                    const double *H, int ldh, int m, double beta, double *w);
/* _phiv!(w, t, Ks, k, cache, correct) on a ComplexF64 subspace (src/krylov_phiv.jl:620-653): W device, complex,
 * nrows x (k+1), leading dimension ldw (complex elements); errest may be NULL. */
int b200k_phiv_ks_z(b200k_handle_t h, double t_re, double t_im, const double *V, int64_t ldv, int64_t nrows,
                    const double *H, int ldh, int m, double beta, int k, int correct, double *W, int64_t ldw,
                    double *errest);
/* expv(t, A, b) one-shot on a complex operator. */
int b200k_expv_z(b200k_handle_t h, b200k_op_t op, double t_re, double t_im, const double *b,
                 const b200k_krylov_opts *opts, double *w, int *m_out, int *breakdown);
/* Host small dense phase for complex H and/or complex t: y = exp(t H[1:m,1:m]) e1 (branch as b200k_expv_small). */
int b200k_expv_small_z(int m, const double *H, int ldh, double t_re, double t_im, double *y, int *branch);
/* W (device, nrows x nc) = beta * V[:, 1:m] * Y for a REAL basis V and a HOST coefficient matrix Y (m x nc, ldy):
 * the projection step on its own.  Used for a real Krylov subspace with complex t (Y = [re(y) im(y)]). */
int b200k_project(b200k_handle_t h, const double *V, int64_t ldv, int64_t nrows, int m, double beta, const double *Y,
                  int ldy, int nc, double *W, int64_t ldw);

/* ---- phiv_timestep! / expv_timestep! (src/krylov_phiv_adaptive.jl:57-114, 260-501) ------------------------
 * u(t) = phi_0(tA) b_0 + t phi_1(tA) b_1 + ... + t^p phi_p(tA) b_p at the times ts, by internal time stepping
 * with the Niesen-Wright adaptation of (tau, m) when `adaptive`.  Keyword arguments of the reference, same
 * defaults; opnorm = NaN means `nothing` (scale estimated from the Arnoldi Hessenberg, :371-385). */
typedef struct {
    double tau;     /* 0.0: estimate */
    int m;          /* min(10, n) */
    double tol;     /* 1e-7 */
    double opnorm;  /* NaN = nothing */
    int iop;        /* 0 */
    int correct;    /* false */
    int adaptive;   /* false */
    double delta;   /* 1.2 */
    int hermitian;  /* -1: ishermitian(A); only used for the flops model, as in the reference */
    double gamma;   /* 0.8 */
    int64_t NA;     /* 0: nnz(A) */
} b200k_timestep_opts;
void b200k_timestep_opts_default(b200k_timestep_opts *o);
/* ts: HOST, nts times (sorted in place, as the reference does).  B: device n x ncoef (ncoef = p + 1; ncoef = 1 is
 * expv_timestep).  U: device n x nts.  num_timesteps (may be NULL) receives the number of internal steps. */
int b200k_phiv_timestep(b200k_handle_t h, b200k_op_t op, int nts, double *ts, const double *B, int64_t ldb,
                        int ncoef, const b200k_timestep_opts *opts, double *U, int64_t ldu, int *num_timesteps);

/* ---- row sharding of ONE vector across the GPUs of a node (one process per GPU) --------------------------
 * No reference equivalent (the reference has no distributed code); this is how
 * arnoldi!/expv/phiv/kiops scale past one GPU.  Rank r owns a contiguous block of rows of every vector and of
 * the operator; H, beta and the p augmented rows are replicated (every rank computes bitwise identical values).
 * Per Krylov step the only exchanges are (i) the halo entries of the gather vector and (ii) the <= m+1 partial
 * inner products / the norm.  Both are done INSIDE the persistent kernel: every CTA stores its contributions
 * straight into the peers' buffers (peer-mapped memory over NVLink) and the team barrier is extended across
 * GPUs with system-scope atomics -- there is no NCCL call on the data path.  The host only exchanges the
 * 64-byte CUDA IPC handles once (torch.distributed / MPI / anything). */
typedef struct b200k_comm *b200k_comm_t;
#define B200K_IPC_HANDLE_BYTES 64
/* Allocate this rank's peer-visible buffer (barrier word, partial-sum inboxes, two gather buffers of xlen
 * doubles: nloc local entries | halo landing zone | augmented tail) and return its IPC handle.  xlen must be
 * the same on every rank (max over ranks of nloc + nhalo, plus 16). */
int b200k_comm_create(b200k_handle_t h, int rank, int nranks, int64_t xlen, unsigned char *handle_out,
                      b200k_comm_t *comm);
/* all_handles: nranks x 64 bytes in rank order.  Maps every peer's buffer (cudaIpcOpenMemHandle). */
int b200k_comm_connect(b200k_comm_t comm, const unsigned char *all_handles);
int b200k_comm_destroy(b200k_comm_t comm);
/* This rank's row block of a CSR operator.  colind are LOCAL gather indices: [0, nloc) = own rows,
 * [nloc, nloc + nhalo) = the rank's sorted list of remote columns.  The send list (HOST arrays, sorted by
 * send_row) names, for every own row some other rank gathers, the destination rank and the position in its
 * gather buffer.  is_hermitian: LinearAlgebra.ishermitian of the GLOBAL operator (cannot be decided locally). */
int b200k_op_csr_create_sharded(b200k_handle_t h, b200k_comm_t comm, int64_t nloc, int64_t nhalo, int64_t nnz,
                                const int32_t *rowptr, const int32_t *colind, const double *val, int index_base,
                                int location, int is_hermitian, int64_t nsend, const int32_t *send_row,
                                const int32_t *send_peer, const int32_t *send_pos, b200k_op_t *op);

/* This rank's row block of a DENSE operator (the dense analogue; BASELINE config 3 beyond one GPU, reference call site
 * src/arnoldi.jl:185).  row_starts: HOST, nranks + 1 global row offsets (rank r owns rows [row_starts[r],
 * row_starts[r+1]), all even); A_block: this rank's rows, (row_starts[rank+1] - row_starts[rank]) x n_global
 * column-major with leading dimension lda, columns in GLOBAL order (the library copies it and permutes the columns
 * into its gather order: own entries first).  The communicator's xlen must be >= n_global + 16.  Every step each rank
 * pushes its block of the new vector to every peer (the "all-gather of x") inside the persistent kernel. */
int b200k_op_dense_create_sharded(b200k_handle_t h, b200k_comm_t comm, int64_t n_global, const int64_t *row_starts,
                                  const double *A_block, int64_t lda, int location, int is_hermitian, b200k_op_t *op);

/* ---- small dense matrix functions (host, m <= ~130) ----------------------------------------- */
/* exponential!(A, ExpMethodHigham2005Base()) in place (src/exp_baseexp.jl:112-161). */
int b200k_exponential(int n, double *A, int lda);
/* Batched exponential!(A_b, ExpMethodHigham2005Base()) of nbatch n x n matrices RESIDENT ON THE DEVICE, in place, one CTA
 * per matrix (n <= 48).  A: device, matrix b at A + b*stride, column-major with leading dimension lda.
 * Synchronous with respect to the status: returns B200K_ESINGULAR if any Pade denominator was singular. */
int b200k_exponential_batched(b200k_handle_t h, int nbatch, int n, double *A, int lda, int64_t stride);
/* The small dense phase of expv! on its own (src/krylov_phiv.jl:223-244): y = exp(t*H[1:m,1:m]) e1,
 * taking the SymTridiagonal eigen branch when H[1:m,1:m] is exactly symmetric, else the Pade branch.
 * branch (may be NULL) receives 1 for the symmetric branch, 0 for Pade. */
int b200k_expv_small(int m, const double *H, int ldh, double t, double *y, int *branch);
/* phiv_dense!(w, A, v, k) (src/phi.jl:84-115): w is m x (k+1), ldw. */
int b200k_phiv_dense(int m, const double *A, int lda, const double *v, int k, double *w, int ldw);

/* ---- instrumentation ------------------------------------------------------------------------ */
/* Device time in ms of the last Krylov-factorisation kernel and of the last projection kernel
 * (CUDA events on the handle's stream; valid after a synchronising call). */
int b200k_last_timing(b200k_handle_t h, float *krylov_ms, float *project_ms);
/* Enable (1) / disable (0) the event timing above; off by default. */
int b200k_set_timing(b200k_handle_t h, int enabled);
/* Which Krylov kernel the last factorisation used: 1 = LDG kernel (krylov_persistent_kernel, any layout),
 * 2 = TMA-ring kernel (krylov_tma_kernel; needs even n / ldv and 16-byte aligned bases), 3 = complex kernel,
 * 4 = the short-window (Lanczos / IOP) instance of the TMA-ring kernel (resident basis vector), 5 = the lock-step
 * multi-vector Lanczos kernel of batched expv, 6 = the short-window instance with the one-reduction Lanczos step,
 * 7 = the complex kernel on the TMA ring (krylov_tma_z_kernel; CSR operators with short rows).  The environment
 * variable B200K_KERNEL=ldg, read at b200k_create, forces 1 (A/B measurements). */
int b200k_last_kernel(b200k_handle_t h, int *which);
/* Runtime switches (A/B measurements and tests): B200K_FLAG_FORCE_LDG = 1 uses krylov_persistent_kernel even
 * where the TMA-ring kernel applies; B200K_FLAG_HOST_SMALLEXP = 1 makes the fused one-shot / batched expv do the
 * small exponential on the host (both branches of krylov_phiv.jl:225-244) instead of small_exp_kernel. */
#define B200K_FLAG_FORCE_LDG 1
#define B200K_FLAG_HOST_SMALLEXP 2
#define B200K_FLAG_L2HINT 3 /* L2::evict_first on the CSR operator stream: -1 automatic (default: only in steps
                               whose working set exceeds the L2), 0 never, 1 always */
#define B200K_FLAG_NO_XL 4  /* 1: never use the short-window (Lanczos / IOP) instance of the TMA-ring kernel that keeps
                               the current basis vector in shared memory and reduces with packet all-reduces */
#define B200K_FLAG_NO_MV 5 /* batched Lanczos and the lock-step multi-vector kernel (several problems per team): 1 = never,
                              2 = whenever possible, 0 (default) = when the cost model says it is faster */
#define B200K_FLAG_SYM_PADE 6 /* 1: the device-side small exponential of a symmetric tridiagonal (Lanczos) H uses the
                                 Pade path instead of the one-warp Chebyshev evaluation of exp(tT) e1 (A/B, tests) */
#define B200K_FLAG_NO_LZ1 7 /* Lanczos on the short-window instance: 0 (default) = the one-reduction step (krylov_kernel_tma.cuh,
                               "One-reduction Lanczos step") for row-sharded operators, the two-reduction step on one GPU;
                               1 = never the one-reduction step, 2 = always */
int b200k_set_flag(b200k_handle_t h, int flag, int value);

#ifdef __cplusplus
}
#endif
#endif /* B200KRYLOV_H */
