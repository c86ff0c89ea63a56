// krylov_kernel.cuh -- the persistent fused Arnoldi / Lanczos / IOP kernel (sm_100a).
//
// Replaces the j-loop of arnoldi!/lanczos! (reference: src/arnoldi.jl:345-377, 456-490) including
// firststep! (:230-279), applyA! plain and augmented (:183-205), arnoldi_step! (:289-308) and
// lanczos_step! (:388-403).  One cooperative launch runs ALL Krylov steps of one problem (or of a
// batch of independent problems sharing the operator):
//
//   * The grid is split into teams of C CTAs (single problem: one team = every SM).  A team owns one
//     problem at a time; CTA `rank` of the team owns the contiguous row slice
//     [rank*slice, (rank+1)*slice) of every length-n vector for the whole factorisation.
//   * Per step j:   w = A v_j (slice kept in shared memory)            -- operator streamed from HBM:
//                                                                        CSR chunks staged by 1-D TMA
//                                                                        bulk copies (UBLKCP) + mbarrier
//                   h = V_j^T w   classical Gram-Schmidt, fused         -- V slice streamed, 16-B loads,
//                                                                        butterfly warp reduction
//                   == team barrier #1 (deterministic tree sum of the per-CTA partials) ==
//                   w -= V_j h ; partial ||w||^2 ; w -> x scratch       -- second pass over the V slice
//                   == team barrier #2 ==
//                   beta = ||w|| ; H[j+1,j] = beta ; v_{j+1} = w / beta -- written once to V
//     The next mat-vec gathers from the *unnormalised* scratch copy and folds 1/beta into the row
//     sums, so normalisation needs no third barrier.
//   * Lanczos (hermitian) and IOP-q (kiops) are the same loop with a shorter orthogonalisation window.
//   * Augmented operator [A B; 0 K] of kiops: the p extra rows live in a tiny replicated tail.
//
// Numerical note: the reference orthogonalises with sequential *modified* Gram-Schmidt; this kernel
// uses classical Gram-Schmidt so that all inner products of a step need one reduction.
#pragma once
#include "ptx.cuh"

namespace b200k {

constexpr int NT = 512;          // threads per CTA
constexpr int NW = NT / 32;      // warps per CTA
constexpr int CB = 8;            // basis columns per accumulation batch
constexpr int MAXCOL = 256;      // max orthogonalisation window (Krylov dimension m <= 255)
constexpr int STAGES = 3;        // TMA stages for the CSR stream
constexpr int CH_NNZ = 3072;     // nnz capacity of one stage
constexpr int CH_PAD = 8;
constexpr int MAXCH = 512;       // chunk-table capacity (chunks per CTA slice)
constexpr int MAXP = 16;         // max augmented rows (kiops p)
constexpr int CPAD = 160;        // padded team size for partial-sum rows (>= 148, multiple of 32)

enum OpKind { OP_CSR_STREAM = 0, OP_CSR_WARP = 1, OP_DENSE = 2 };

struct KrylovParams {
    // operator
    int op_kind;
    int n;
    const int *rowptr;
    const int *colind;
    const double *val;
    int ch_rows;  // rows per TMA chunk (CSR stream)
    const double *Ad;
    long long lda;
    int ncols;  // dense: number of columns (= n, or the GLOBAL dimension for a row block of a row-sharded dense operator)
    // augmentation (kiops)
    int p;
    const double *Bm;
    long long ldb;
    const double *btail;  // device, p values (b_aug of firststep!)
    // team geometry
    int team_size;
    int nteams;
    int nprob;
    int slice;  // rows per CTA, multiple of 16
    int vec2;   // 1: 16-byte path (n even, ldv even, 16-B aligned bases)
    // per-problem, strided
    const double *b;
    long long b_stride;
    double *V;
    long long ldv;
    long long V_stride;
    double *Hd;
    int ldh;
    long long H_stride;
    double *scal;  // [prob*4]: beta
    int *stat;     // [prob*4]: m_out, breakdown
    // algorithm
    int m;
    int j0;  // 0: firststep!, else continue with step j0 (1-based)
    int j0_from_stat;  // SAFE instance of krylov_tma_kernel: per problem, resume at stat[4 prob + 3] (0: skip the problem)
    int iop;
    int lanczos;
    double tol;
    // scratch
    double *xbuf;  // [nteams][2][xlen]
    long long xlen;
    double *part;   // [nteams][2][MAXCOL][CPAD]
    double *partn;  // [nteams][4][CPAD]
    unsigned *bar;  // [nteams]
    double *wglob;  // [nteams][n] when the w slice does not fit in shared memory
    int w_in_smem;
    // TMA-ring kernel (krylov_kernel_tma.cuh) only
    int nnz_cap;       // nnz capacity of one ring slot holding a CSR chunk (val | colind | rowptr segment)
    int nslot;         // ring depth
    int tile_rows;     // rows per basis tile (multiple of 16, <= 4096)
    int l2hint;        // bit 1: basis tiles L2::evict_last (experiment; off)
    int hintA_cols;    // operator chunks get L2::evict_first in steps whose window has >= this many columns
    int dense_cpt;     // dense operator: columns per ring slot (0: direct loads)
    int dense_box_rows;  // rows per tensor-map box (<= 256); a tile = slice/dense_box_rows boxes of dense_cpt columns
    int dscratch_off;  // byte offset of the dense mat-vec reduction scratch in dynamic shared memory
    uint4 *llpkt;      // single-GPU packet all-reduce inboxes [2 parities][LLQ][CPAD dest CTA][CPAD source rank]
    uint4 *llloc;      // row-sharded XL: this GPU's local packet inbox [2 parities][LLQ][CPAD source CTAs] (in the comm buffer)
    int xl;            // XL instance: the current basis vector's slice stays in shared memory (second w-sized buffer)
    // row sharding across GPUs (krylov_kernel_tma.cuh; nranks == 1: single GPU, peers point to local buffers).
    // Every CTA of every GPU writes its partial sums into the inbox of EVERY GPU (peer stores over NVLink) and
    // arrives on every GPU's barrier word with a system-scope atomic: the all-reduce is the team barrier.
    int nranks;
    int myrank;
    unsigned bar_base;      // local barrier arrivals accumulated by earlier launches (multi-GPU counters are never reset)
    unsigned seq_base;      // team barriers passed by earlier launches
    unsigned *comm_state;   // row-sharded: {sequence number, barrier target} carried from launch to launch on the device
    uint4 *peer_pkt[8];     // per GPU: LL inbox [2 parities][MAXCOL + 1 quantities][8 source ranks] of {lo, seq, hi, seq}
    int nhalo;              // remote x entries this GPU gathers (appended after the n local entries)
    int cpad;               // row length of the partial-sum tables: round_up(team_size * nranks, 32)
    double *peer_part[8];   // [2][MAXCOL][cpad] on each GPU
    double *peer_partn[8];  // [4][cpad]
    unsigned *peer_bar[8];
    double *peer_xbuf[8];   // [2][xlen]: gather source incl. the halo landing zone [nloc, nloc + nhalo)
    const int *send_row;    // halo push list sorted by local row: row, destination rank, position in its xbuf
    const int *send_peer;
    const int *send_pos;
    const int *send_ofs;    // [team_size + 1] range of the list owned by each CTA slice
};

struct __align__(128) SmemFixed {
    double val_s[STAGES][CH_NNZ + CH_PAD];
    int col_s[STAGES][CH_NNZ + CH_PAD];
    double hs[MAXCOL];
    double red[2][NW][CB];
    double redn[NW];
    double wtail[MAXP];
    double xtail[MAXP];
    int chunk_a0[MAXCH];
    int chunk_cnt[MAXCH];
    int stage_a0[STAGES];
    uint64_t full[STAGES];
};

// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reduce 8 per-lane accumulators over the warp with 9 shuffles (transposing butterfly).  Returns the
// full sum of accumulator index ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1).
__device__ __forceinline__ double warp_reduce8(double (&a)[CB], int lane) {
    bool hi = (lane & 16) != 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const double send = hi ? a[u] : a[u + 4];
        const double keep = hi ? a[u + 4] : a[u];
        a[u] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    hi = (lane & 8) != 0;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const double send = hi ? a[u] : a[u + 2];
        const double keep = hi ? a[u + 2] : a[u];
        a[u] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    hi = (lane & 4) != 0;
    {
        const double send = hi ? a[0] : a[1];
        const double keep = hi ? a[1] : a[0];
        a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 2);
    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
    return a[0];
}

struct Team {
    unsigned *bar;
    unsigned target;
    unsigned seq;  // number of team barriers passed (multi-GPU flags carry this)
    int C;
    int rank;
};

// All CTAs of a team arrive; returns when every one has.  Same protocol as a cooperative-groups grid
// sync (fence / arrive / spin / fence by one thread, bracketed by CTA barriers); co-residency is
// guaranteed by the cooperative launch.
__device__ __forceinline__ void team_barrier(Team &tm) {
    __syncthreads();
    if (threadIdx.x == 0) {
        tm.target += (unsigned)tm.C;
        __threadfence();
        atomicAdd(tm.bar, 1u);
        while ((int)(ld_acquire_u32(tm.bar) - tm.target) < 0) {
        }
        __threadfence();
    }
    __syncthreads();
}

// Deterministic sum of the C per-CTA partials in row `pp` (every lane returns the same value).
__device__ __forceinline__ double team_sum(const double *pp, int C, int lane) {
    double s = 0.0;
    for (int q = lane; q < C; q += 32) s += ld_cg(pp + q);
    return warp_sum(s);
}

struct Ctx {
    SmemFixed *S;
    double *ws;  // this CTA's slice of w (shared memory, or global scratch for very large n)
    int tid, lane, warp;
    int r0, nrows;  // slice [r0, r0 + nrows)
    int nch;        // CSR chunks in the slice
    unsigned gchunk;  // running TMA chunk counter (stage / parity bookkeeping)
};

// ---- mat-vec phase ---------------------------------------------------------------------------------
__device__ __forceinline__ void csr_issue_chunk(const KrylovParams &P, Ctx &cx, int c, unsigned g) {
    SmemFixed *S = cx.S;
    const int stage = (int)(g % STAGES);
    int a0, cnt;
    if (cx.nch <= MAXCH) {
        a0 = S->chunk_a0[c];
        cnt = S->chunk_cnt[c];
    } else {
        const int rs = cx.r0 + c * P.ch_rows;
        const int re = min(cx.r0 + cx.nrows, rs + P.ch_rows);
        const int e0 = P.rowptr[rs], e1 = P.rowptr[re];
        a0 = e0 & ~3;
        cnt = ((e1 + 3) & ~3) - a0;
    }
    S->stage_a0[stage] = a0;
    if (cnt > 0) {
        mbar_arrive_expect_tx(&S->full[stage], (uint32_t)cnt * 12u);
        bulk_g2s(S->val_s[stage], P.val + a0, (uint32_t)cnt * 8u, &S->full[stage]);
        bulk_g2s(S->col_s[stage], P.colind + a0, (uint32_t)cnt * 4u, &S->full[stage]);
    } else {
        mbar_arrive(&S->full[stage]);
    }
}

// w[0..nrows) = xscale * (A x)[slice] (+ xscale * B x_tail for the augmented operator); tail rows shift.
template <int VEC>
__device__ __forceinline__ void matvec_phase(const KrylovParams &P, Ctx &cx, const double *xsrc,
                                             double xscale) {
    SmemFixed *S = cx.S;
    const int tid = cx.tid, lane = cx.lane, warp = cx.warp;
    const int n = P.n, p = P.p;
    double *ws = cx.ws;
    if (p > 0) {
        if (tid < p) S->xtail[tid] = xsrc[n + tid];
        __syncthreads();
        if (tid < p) S->wtail[tid] = (tid < p - 1) ? S->xtail[tid + 1] * xscale : 0.0;
    }
    if (P.op_kind == OP_CSR_STREAM) {
        const int nch = cx.nch;
        if (tid == 0) {
            const int pre = min(STAGES - 1, nch);
            for (int c = 0; c < pre; ++c) csr_issue_chunk(P, cx, c, cx.gchunk + c);
        }
        for (int c = 0; c < nch; ++c) {
            const unsigned g = cx.gchunk + c;
            const int stage = (int)(g % STAGES);
            const unsigned parity = (g / STAGES) & 1u;
            if (tid == 0 && c + STAGES - 1 < nch) csr_issue_chunk(P, cx, c + STAGES - 1, g + STAGES - 1);
            const int rl = c * P.ch_rows + tid;  // local row
            const bool active = tid < P.ch_rows && rl < cx.nrows;
            int e0 = 0, e1 = 0;
            if (active) {
                e0 = P.rowptr[cx.r0 + rl];
                e1 = P.rowptr[cx.r0 + rl + 1];
            }
            mbar_wait(&S->full[stage], parity);
            if (active) {
                const int a0 = S->stage_a0[stage];
                const double *vs = S->val_s[stage];
                const int *cs = S->col_s[stage];
                double sum = 0.0;
                for (int e = e0 - a0; e < e1 - a0; ++e) sum = fma(vs[e], xsrc[cs[e]], sum);
                if (p > 0) {
                    const double *brow = P.Bm + (cx.r0 + rl);
                    for (int k = 0; k < p; ++k) sum = fma(brow[(long long)k * P.ldb], S->xtail[k], sum);
                }
                ws[rl] = sum * xscale;
            }
            __syncthreads();
        }
        cx.gchunk += (unsigned)nch;
    } else if (P.op_kind == OP_CSR_WARP) {
        for (int rl = warp; rl < cx.nrows; rl += NW) {
            const int row = cx.r0 + rl;
            const int e0 = P.rowptr[row], e1 = P.rowptr[row + 1];
            double sum = 0.0;
            for (int e = e0 + lane; e < e1; e += 32) sum = fma(ld_ro1(P.val + e), xsrc[P.colind[e]], sum);
            sum = warp_sum(sum);
            if (lane == 0) {
                if (p > 0)
                    for (int k = 0; k < p; ++k) sum = fma(P.Bm[row + (long long)k * P.ldb], S->xtail[k], sum);
                ws[rl] = sum * xscale;
            }
        }
        __syncthreads();
    } else {  // dense column-major: thread = (row group of VEC rows) x (column group)
        const int units = (cx.nrows + VEC - 1) / VEC;  // VEC==2 implies nrows even
        int RL = 32;
        while (RL < units && RL < NT) RL <<= 1;
        const int G = NT / RL;
        const int ul = tid % RL, g = tid / RL;
        double *scratch = &S->val_s[0][0];  // NT * VEC doubles of reduction scratch (stages unused here)
        for (int ubase = 0; ubase < units; ubase += RL) {
            const int u = ubase + ul;
            const bool valid = u < units;
            double a0 = 0.0, a1 = 0.0;
            if (valid) {
                const double *ap = P.Ad + cx.r0 + (long long)VEC * u;
#pragma unroll 8
                for (int c = g; c < n; c += G) {
                    const double xc = xsrc[c];
                    if (VEC == 2) {
                        const double2 a2 = ld_ro2(ap + (long long)c * P.lda);
                        a0 = fma(a2.x, xc, a0);
                        a1 = fma(a2.y, xc, a1);
                    } else {
                        a0 = fma(ld_ro1(ap + (long long)c * P.lda), xc, a0);
                    }
                }
            }
            scratch[(g * RL + ul) * 2 + 0] = a0;
            scratch[(g * RL + ul) * 2 + 1] = a1;
            __syncthreads();
            if (g == 0 && valid) {
                double s0 = 0.0, s1 = 0.0;
                for (int q = 0; q < G; ++q) {
                    s0 += scratch[(q * RL + ul) * 2 + 0];
                    s1 += scratch[(q * RL + ul) * 2 + 1];
                }
                const int rl = VEC * u;
                if (p > 0) {
                    for (int k = 0; k < p; ++k) {
                        s0 = fma(P.Bm[cx.r0 + rl + (long long)k * P.ldb], S->xtail[k], s0);
                        if (VEC == 2) s1 = fma(P.Bm[cx.r0 + rl + 1 + (long long)k * P.ldb], S->xtail[k], s1);
                    }
                }
                ws[rl] = s0 * xscale;
                if (VEC == 2) ws[rl + 1] = s1 * xscale;
            }
            __syncthreads();
        }
    }
}

// ---- Gram-Schmidt: inner products of w with basis columns [lo, hi] ------------------------------------
// Per-CTA partials go to part[(ci)*CPAD + rank], ci = column - lo.
template <int VEC>
__device__ __forceinline__ void dots_phase(const KrylovParams &P, Ctx &cx, const Team &tm, const double *V,
                                           int lo, int hi, double *part) {
    SmemFixed *S = cx.S;
    const int tid = cx.tid, lane = cx.lane, warp = cx.warp;
    const long long ldv = P.ldv;
    const double *ws = cx.ws;
    const int units = cx.nrows / VEC;
    int batch = 0;
    for (int cb = lo; cb <= hi; cb += CB, ++batch) {
        const int nb = min(CB, hi - cb + 1);
        double acc[CB];
#pragma unroll
        for (int u = 0; u < CB; ++u) acc[u] = 0.0;
        const double *vb = V + (long long)cb * ldv + cx.r0;
        if (nb == CB) {
            for (int i = tid; i < units; i += NT) {
                if (VEC == 2) {
                    const double2 w2 = reinterpret_cast<const double2 *>(ws)[i];
                    double2 v2[CB];
#pragma unroll
                    for (int u = 0; u < CB; ++u) v2[u] = ld_stream2(vb + (long long)u * ldv + 2 * i);
#pragma unroll
                    for (int u = 0; u < CB; ++u) acc[u] = fma(v2[u].x, w2.x, fma(v2[u].y, w2.y, acc[u]));
                } else {
                    const double w1 = ws[i];
                    double v1[CB];
#pragma unroll
                    for (int u = 0; u < CB; ++u) v1[u] = ld_stream1(vb + (long long)u * ldv + i);
#pragma unroll
                    for (int u = 0; u < CB; ++u) acc[u] = fma(v1[u], w1, acc[u]);
                }
            }
        } else {
            for (int i = tid; i < units; i += NT) {
                if (VEC == 2) {
                    const double2 w2 = reinterpret_cast<const double2 *>(ws)[i];
#pragma unroll
                    for (int u = 0; u < CB; ++u)
                        if (u < nb) {
                            const double2 v2 = ld_stream2(vb + (long long)u * ldv + 2 * i);
                            acc[u] = fma(v2.x, w2.x, fma(v2.y, w2.y, acc[u]));
                        }
                } else {
                    const double w1 = ws[i];
#pragma unroll
                    for (int u = 0; u < CB; ++u)
                        if (u < nb) acc[u] = fma(ld_stream1(vb + (long long)u * ldv + i), w1, acc[u]);
                }
            }
        }
        if (P.p > 0 && tm.rank == 0 && tid == 0) {  // augmented tail rows
#pragma unroll
            for (int u = 0; u < CB; ++u)
                if (u < nb)
                    for (int k = 0; k < P.p; ++k)
                        acc[u] = fma(V[(long long)(cb + u) * ldv + P.n + k], S->wtail[k], acc[u]);
        }
        const double r = warp_reduce8(acc, lane);
        const int buf = batch & 1;
        if ((lane & 3) == 0) S->red[buf][warp][((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = r;
        __syncthreads();
        if (tid < nb) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += S->red[buf][w][tid];
            part[(long long)(cb - lo + tid) * CPAD + tm.rank] = s;
        }
    }
}

// ---- Gram-Schmidt update: w -= sum_c hu[c-ulo] V[:,c], c = uhi..ulo; returns this thread's partial ||w||^2;
// the unnormalised w slice is also written to xout (the next mat-vec's gather source).
template <int VEC>
__device__ __forceinline__ double update_phase(const KrylovParams &P, Ctx &cx, const Team &tm, const double *V,
                                               int ulo, int uhi, double *xout) {
    SmemFixed *S = cx.S;
    const int tid = cx.tid;
    const long long ldv = P.ldv;
    double *ws = cx.ws;
    const int units = cx.nrows / VEC;
    const double *hs = S->hs;
    double nrm = 0.0;
    for (int i = tid; i < units; i += NT) {
        if (VEC == 2) {
            double2 w2 = reinterpret_cast<double2 *>(ws)[i];
            const double *vrow = V + cx.r0 + 2 * i;
            for (int ctop = uhi; ctop >= ulo; ctop -= CB) {
                const int nb = min(CB, ctop - ulo + 1);
                double2 v2[CB];
#pragma unroll
                for (int u = 0; u < CB; ++u)
                    if (u < nb) v2[u] = ld_stream2(vrow + (long long)(ctop - u) * ldv);
#pragma unroll
                for (int u = 0; u < CB; ++u)
                    if (u < nb) {
                        const double hc = hs[ctop - u - ulo];
                        w2.x = fma(-hc, v2[u].x, w2.x);
                        w2.y = fma(-hc, v2[u].y, w2.y);
                    }
            }
            reinterpret_cast<double2 *>(ws)[i] = w2;
            reinterpret_cast<double2 *>(xout + cx.r0)[i] = w2;
            nrm = fma(w2.x, w2.x, fma(w2.y, w2.y, nrm));
        } else {
            double w1 = ws[i];
            const double *vrow = V + cx.r0 + i;
            for (int ctop = uhi; ctop >= ulo; ctop -= CB) {
                const int nb = min(CB, ctop - ulo + 1);
                double v1[CB];
#pragma unroll
                for (int u = 0; u < CB; ++u)
                    if (u < nb) v1[u] = ld_stream1(vrow + (long long)(ctop - u) * ldv);
#pragma unroll
                for (int u = 0; u < CB; ++u)
                    if (u < nb) w1 = fma(-hs[ctop - u - ulo], v1[u], w1);
            }
            ws[i] = w1;
            xout[cx.r0 + i] = w1;
            nrm = fma(w1, w1, nrm);
        }
    }
    if (P.p > 0 && tid < P.p) {  // tail rows: every CTA keeps its own copy, rank 0 publishes
        double wt = S->wtail[tid];
        for (int c = uhi; c >= ulo; --c) wt = fma(-hs[c - ulo], V[(long long)c * ldv + P.n + tid], wt);
        S->wtail[tid] = wt;
        if (tm.rank == 0) {
            xout[P.n + tid] = wt;
            nrm = fma(wt, wt, nrm);
        }
    }
    return nrm;
}

// CTA-wide deterministic sum of one double per thread; the result is written by thread 0 to *out.
__device__ __forceinline__ void block_sum_to(Ctx &cx, double v, double *out) {
    v = warp_sum(v);
    if (cx.lane == 0) cx.S->redn[cx.warp] = v;
    __syncthreads();
    if (cx.tid == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += cx.S->redn[w];
        *out = s;
    }
}

// ----------------------------------------------------------------------------------------------------
template <int VEC>
__device__ void krylov_body(const KrylovParams &P, SmemFixed *S, double *ws_smem) {
    Ctx cx;
    cx.S = S;
    cx.tid = threadIdx.x;
    cx.lane = threadIdx.x & 31;
    cx.warp = threadIdx.x >> 5;
    cx.gchunk = 0;
    Team tm;
    const int team = blockIdx.x / P.team_size;
    tm.rank = blockIdx.x % P.team_size;
    tm.C = P.team_size;
    tm.bar = P.bar + team;
    tm.target = 0;
    const int n = P.n, p = P.p;
    cx.r0 = min(n, tm.rank * P.slice);
    cx.nrows = min(n, cx.r0 + P.slice) - cx.r0;
    cx.ws = P.w_in_smem ? ws_smem : (P.wglob + (long long)team * n + cx.r0);
    cx.nch = 0;
    const int tid = cx.tid;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&S->full[s], 1);
        mbar_fence_init();
    }
    if (P.op_kind == OP_CSR_STREAM) {
        cx.nch = (cx.nrows + P.ch_rows - 1) / P.ch_rows;
        if (cx.nch <= MAXCH) {
            for (int c = tid; c < cx.nch; c += NT) {
                const int rs = cx.r0 + c * P.ch_rows;
                const int re = min(cx.r0 + cx.nrows, rs + P.ch_rows);
                const int e0 = P.rowptr[rs], e1 = P.rowptr[re];
                const int a0 = e0 & ~3;
                S->chunk_a0[c] = a0;
                S->chunk_cnt[c] = ((e1 + 3) & ~3) - a0;
            }
        }
    }
    __syncthreads();

    double *xb0 = P.xbuf + (long long)team * 2 * P.xlen;
    double *xb1 = xb0 + P.xlen;
    double *part0 = P.part + (long long)team * 2 * MAXCOL * CPAD;
    double *partn0 = P.partn + (long long)team * 4 * CPAD;
    const int units = cx.nrows / VEC;

    int nlocal = -1;
    for (int prob = team; prob < P.nprob; prob += P.nteams) {
        ++nlocal;
        double *V = P.V + (long long)prob * P.V_stride;
        double *Hd = P.Hd + (long long)prob * P.H_stride;
        const double *b = P.b + (long long)prob * P.b_stride;
        const long long ldv = P.ldv;
        const int ldh = P.ldh;
        const double *xsrc;
        double xscale;
        int jstart;
        int m_out = P.m, breakdown = 0;

        if (P.j0 == 0) {
            // firststep! (arnoldi.jl:230-250 / 257-279): beta = ||[b; b_aug]||, v_1 = b / beta.
            // The b slice is parked in the w buffer so it is read from HBM once.
            double nrm = 0.0;
            double *ws = cx.ws;
            for (int i = tid; i < units; i += NT) {
                if (VEC == 2) {
                    const double2 b2 = reinterpret_cast<const double2 *>(b + cx.r0)[i];
                    reinterpret_cast<double2 *>(ws)[i] = b2;
                    if (p > 0) reinterpret_cast<double2 *>(xb0 + cx.r0)[i] = b2;
                    nrm = fma(b2.x, b2.x, fma(b2.y, b2.y, nrm));
                } else {
                    const double b1 = b[cx.r0 + i];
                    ws[i] = b1;
                    if (p > 0) xb0[cx.r0 + i] = b1;
                    nrm = fma(b1, b1, nrm);
                }
            }
            if (p > 0 && tm.rank == 0 && tid < p) {
                const double bt = P.btail[tid];
                xb0[n + tid] = bt;
                nrm = fma(bt, bt, nrm);
            }
            double *pslot = partn0 + (2 + (nlocal & 1)) * CPAD;
            block_sum_to(cx, nrm, pslot + tm.rank);
            team_barrier(tm);
            const double beta = sqrt(team_sum(pslot, tm.C, cx.lane));
            if (tm.rank == 0 && tid == 0) P.scal[prob * 4] = beta;
            if (beta == 0.0) {  // zero start vector: Ks.m stays m, V untouched (arnoldi.jl:366)
                if (tm.rank == 0 && tid == 0) {
                    P.stat[prob * 4 + 0] = P.m;
                    P.stat[prob * 4 + 1] = 0;
                }
                continue;
            }
            if (p == 0) {
                const double inv = 1.0 / beta;  // V[i,1] = b[i] * invbeta (arnoldi.jl:240)
                for (int i = tid; i < units; i += NT) {
                    if (VEC == 2) {
                        double2 b2 = reinterpret_cast<const double2 *>(ws)[i];
                        b2.x *= inv;
                        b2.y *= inv;
                        reinterpret_cast<double2 *>(V + cx.r0)[i] = b2;
                    } else {
                        V[cx.r0 + i] = ws[i] * inv;
                    }
                }
                xsrc = b;  // gather from the caller's b, scaled by 1/beta inside the mat-vec
            } else {  // V[1:n,1] = bl / beta; V[n+1:n+p,1] = b_aug / beta (arnoldi.jl:275-276)
                for (int i = tid; i < units; i += NT) {
                    if (VEC == 2) {
                        double2 b2 = reinterpret_cast<const double2 *>(ws)[i];
                        b2.x /= beta;
                        b2.y /= beta;
                        reinterpret_cast<double2 *>(V + cx.r0)[i] = b2;
                    } else {
                        V[cx.r0 + i] = ws[i] / beta;
                    }
                }
                if (tm.rank == 0 && tid < p) V[n + tid] = P.btail[tid] / beta;
                xsrc = xb0;
            }
            __syncthreads();  // the w buffer is overwritten by the first mat-vec
            xscale = 1.0 / beta;
            jstart = 1;
        } else {
            xsrc = V + (long long)(P.j0 - 1) * ldv;  // normalised basis column, tail rows included
            xscale = 1.0;
            jstart = P.j0;
        }

        double beta_prev = 0.0;  // Lanczos: beta_{j-1}
        for (int j = jstart; j <= P.m; ++j) {
            const int jc = j - 1;  // 0-based column of x; the new vector goes to column jc + 1
            const int par = j & 1;
            double *xout = par ? xb1 : xb0;
            double *part = part0 + (long long)par * MAXCOL * CPAD;
            double *partn = partn0 + par * CPAD;

            matvec_phase<VEC>(P, cx, xsrc, xscale);

            const int iopw = P.iop > 0 ? P.iop : P.m;
            const int lo = P.lanczos ? jc : max(0, jc - iopw + 1);
            const int hi = jc;
            dots_phase<VEC>(P, cx, tm, V, lo, hi, part);
            const int nc = hi - lo + 1;
            const int ulo = (P.lanczos && jc >= 1) ? jc - 1 : lo;
            // DGKS re-orthogonalisation test (see krylov_kernel_tma.cuh): ||w_before||^2 = ||h||^2 + ||w_after||^2
            const bool dgks = !P.lanczos;
            team_barrier(tm);

            for (int ci = cx.warp; ci < nc; ci += NW) {
                const double s = team_sum(part + (long long)ci * CPAD, tm.C, cx.lane);
                if (cx.lane == 0) {
                    S->hs[lo + ci - ulo] = s;
                    if (tm.rank == 0) Hd[(long long)jc * ldh + lo + ci] = s;
                }
            }
            if (P.lanczos && jc >= 1 && tid == 0) S->hs[0] = beta_prev;
            __syncthreads();

            double nrm = update_phase<VEC>(P, cx, tm, V, ulo, hi, xout);
            block_sum_to(cx, nrm, partn + tm.rank);
            team_barrier(tm);

            double beta2 = team_sum(partn, tm.C, cx.lane);
            double hsq = 0.0;
            if (dgks)
                for (int ci = 0; ci < nc; ++ci) hsq = fma(S->hs[ci], S->hs[ci], hsq);
            if (dgks && beta2 < 0.0625 * (hsq + beta2)) {  // second classical Gram-Schmidt pass (eta = 1/4)
                __syncthreads();
                dots_phase<VEC>(P, cx, tm, V, lo, hi, part);
                team_barrier(tm);
                for (int ci = cx.warp; ci < nc; ci += NW) {
                    const double s = team_sum(part + (long long)ci * CPAD, tm.C, cx.lane);
                    if (cx.lane == 0) {
                        if (tm.rank == 0) Hd[(long long)jc * ldh + lo + ci] = S->hs[ci] + s;
                        S->hs[ci] = s;
                    }
                }
                __syncthreads();
                nrm = update_phase<VEC>(P, cx, tm, V, lo, hi, xout);
                block_sum_to(cx, nrm, partn + tm.rank);
                team_barrier(tm);
                beta2 = team_sum(partn, tm.C, cx.lane);
            }
            const double beta = sqrt(beta2);
            if (tm.rank == 0 && tid == 0) Hd[(long long)jc * ldh + jc + 1] = beta;
            {  // y /= beta (arnoldi.jl:306): v_{j+1}
                double *vn = V + (long long)(jc + 1) * ldv;
                const double *ws = cx.ws;
                for (int i = tid; i < units; i += NT) {
                    if (VEC == 2) {
                        double2 w2 = reinterpret_cast<const double2 *>(ws)[i];
                        w2.x /= beta;
                        w2.y /= beta;
                        reinterpret_cast<double2 *>(vn + cx.r0)[i] = w2;
                    } else {
                        vn[cx.r0 + i] = ws[i] / beta;
                    }
                }
                if (p > 0 && tm.rank == 0 && tid < p) vn[n + tid] = S->wtail[tid] / beta;
            }
            __syncthreads();  // the next mat-vec overwrites the w slice that was just read
            xsrc = xout;
            xscale = 1.0 / beta;
            beta_prev = beta;
            if (beta < P.tol) {  // happy breakdown (arnoldi.jl:370-374): absolute test
                m_out = j;
                breakdown = 1;
                break;
            }
        }
        if (tm.rank == 0 && tid == 0) {
            P.stat[prob * 4 + 0] = m_out;
            P.stat[prob * 4 + 1] = breakdown;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NT, 1) krylov_persistent_kernel(const __grid_constant__ KrylovParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemFixed *S = reinterpret_cast<SmemFixed *>(smem_raw);
    double *ws = reinterpret_cast<double *>(smem_raw + sizeof(SmemFixed));
    if (P.vec2)
        krylov_body<2>(P, S, ws);
    else
        krylov_body<1>(P, S, ws);
}

}  // namespace b200k
