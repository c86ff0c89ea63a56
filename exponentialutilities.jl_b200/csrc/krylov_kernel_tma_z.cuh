// krylov_kernel_tma_z.cuh -- ComplexF64 Arnoldi / Lanczos / IOP on the producer / consumer ring of krylov_kernel_tma.cuh.
//
// Same algorithm, reduction order per quantity and outputs as krylov_z_kernel (krylov_kernel_z.cuh; reference
// src/arnoldi.jl:230-308, 345-377, 388-403 with conjugating inner products and REAL Lanczos coefficients,
// arnoldi.jl:412-413), but everything that comes from HBM -- the CSR operator (16-byte values) and the basis slice
// (16-byte elements) -- is streamed by one producer lane through the shared-memory ring with 1-D TMA bulk copies, and
// the 16 consumer warps compute out of shared memory.  The LDG kernel kept 16 warps x a few 16-byte loads in flight per
// SM and measured 69 % of the HBM roofline on the n = 10^6 general Arnoldi case; here the bytes in flight are set by the
// ring depth.
//
// Tile schedule of step j (identical on the producer and the consumers of a CTA):
//   [A chunks 0..nch-1]                 one slot = val (16 B / entry) | colind (4 B / entry) | rowptr segment
//   [dots  : for cb in lo..hi step 8 : for k < ntk : for u < nb : basis tile (column cb+u, rows k)]
//   [update: the same tiles in exactly the reverse order (batches, row tiles and columns descending; L2 reuse, see update_z)]
// A basis tile holds <= 2048 complex rows (32 KB); the ring slots are sized to the tile (RingZ).  Two consumer lanes
// share a CSR row (even / odd entries), so a chunk of <= 256 rows keeps all 512 consumer threads busy although a complex
// entry needs 20 bytes of slot space.
//
// Scope: one problem per launch, CSR rows short enough for >= 64 rows per chunk, w slice resident in shared memory, no
// augmentation, no row sharding; everything else stays on krylov_z_kernel.  This instance carries no DGKS code: like
// the real kernel the re-orthogonalisation test is evaluated afterwards from the stored H by
// krylov_z_kernel launched behind it in `safe_scan` mode, which exits at once in the normal case and otherwise redoes
// the factorisation from the first failing step with its two-pass loop.
#pragma once
#include "krylov_kernel_tma.cuh"
#include "krylov_kernel_z.cuh"

namespace b200k {

constexpr int TILE_ROWS_Z = SLOT_BYTES / 16;   // 2048 complex rows per basis tile
constexpr int PPTZ = TILE_ROWS_Z / NTC;        // complex rows per consumer thread per tile (4)
constexpr int CHZ_TPR = 2;                     // consumer lanes per CSR row
constexpr int CHZ_ROWS_MAX = NTC / CHZ_TPR;    // 256

struct __align__(128) SmemTmaZ {
    uint64_t full[MAXSLOT];
    uint64_t empty[MAXSLOT];
    double2 hs[MAXCOL];
    double2 red[2][NW][CB];
    double redn[NW];
    int chunk_a0[MAXCH2];
    int chunk_cnt[MAXCH2];
    int slot_a0[MAXSLOT];
    double bc[2];
    int cols_ready;  // number of complete basis columns (polled by the producer lane: flag_set / flag_get)
    int stop_seq;    // consumers are done (1)
};

// Ring with a run-time slot size: the w slice of a complex problem takes twice the shared memory of a real one, so the
// host sizes the slots to the basis tile (and the CSR chunk to the slot) instead of a fixed 32 KB -- one more slot in
// flight at n = 10^6 (4 x 27 KB instead of 3 x 32 KB with 27 KB used).
struct RingZ {
    unsigned char *base;
    int nslot;
    int slot;
    unsigned phase;
    uint32_t slot_bytes;
    __device__ __forceinline__ void advance() {
        if (++slot == nslot) {
            slot = 0;
            phase ^= 1u;
        }
    }
    __device__ __forceinline__ unsigned char *ptr() const { return base + (size_t)slot * slot_bytes; }
};

struct ConsZ {
    SmemTmaZ *S;
    double2 *ws;
    int tid, lane, warp;
    RingZ rg;
    __device__ __forceinline__ void wait_full() { mbar_wait(&S->full[rg.slot], rg.phase); }
    __device__ __forceinline__ void release() {
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->empty[rg.slot]);
        rg.advance();
    }
};

__device__ __forceinline__ bool prodz_acquire(SmemTmaZ *S, const RingZ &rg) {
    unsigned spins = 0;
    while (!mbar_try_wait(&S->empty[rg.slot], rg.phase ^ 1u)) {
        if ((++spins & 7u) == 0u && flag_get(&S->stop_seq) >= 1) return false;
    }
    return true;
}
__device__ __forceinline__ bool prodz_wait_col(SmemTmaZ *S, int col) {
    while (flag_get(&S->cols_ready) <= col) {
        if (flag_get(&S->stop_seq) >= 1) return false;
    }
    return true;
}

// ---- producer (one lane) ----------------------------------------------------------------------------------------------
__device__ void producer_z(const KrylovParamsZ &P, SmemTmaZ *S, RingZ &rg, const TmaGeom &G) {
    const long long ldv = P.ldv;
    const int jstart = P.j0 == 0 ? 1 : P.j0;
    const int iopw = P.iop > 0 ? P.iop : P.m;
    const int nnz_cap = P.nnz_cap;
    const uint64_t polA = policy_evict_first();
    unsigned issued = 0;
    bool stopped = false;
    for (int j = jstart; j <= P.m && !stopped; ++j) {
        const int jc = j - 1;
        const int lo = P.lanczos ? jc : max(0, jc - iopw + 1);
        const int hi = jc;
        const int ulo = (P.lanczos && jc >= 1) ? jc - 1 : lo;
        // wide windows: the operator cannot survive in L2 until the next step anyway (same rule as the real kernel)
        const bool hintA = (hi - lo + 1) >= P.hintA_cols;
        for (int c = 0; c < G.nch; ++c) {
            if (!prodz_acquire(S, rg)) { stopped = true; break; }
            const int rs = G.r0 + c * P.ch_rows;
            const int re = min(G.r0 + G.nrows, rs + P.ch_rows);
            int a0, cnt;
            if (G.nch <= MAXCH2) {
                a0 = S->chunk_a0[c];
                cnt = S->chunk_cnt[c];
            } else {
                const int e0 = P.rowptr[rs], e1 = P.rowptr[re];
                a0 = e0 & ~3;
                cnt = ((e1 + 3) & ~3) - a0;
            }
            const int rpc = (re - rs + 1 + 3) & ~3;
            S->slot_a0[rg.slot] = a0;
            unsigned char *dst = rg.ptr();
            mbar_arrive_expect_tx(&S->full[rg.slot], (uint32_t)cnt * 20u + (uint32_t)rpc * 4u);
            if (hintA) {
                if (cnt > 0) {
                    bulk_g2s_hint(dst, P.val + a0, (uint32_t)cnt * 16u, &S->full[rg.slot], polA);
                    bulk_g2s_hint(dst + (size_t)nnz_cap * 16, P.colind + a0, (uint32_t)cnt * 4u, &S->full[rg.slot], polA);
                }
                bulk_g2s_hint(dst + (size_t)nnz_cap * 20, P.rowptr + rs, (uint32_t)rpc * 4u, &S->full[rg.slot], polA);
            } else {
                if (cnt > 0) {
                    bulk_g2s(dst, P.val + a0, (uint32_t)cnt * 16u, &S->full[rg.slot]);
                    bulk_g2s(dst + (size_t)nnz_cap * 16, P.colind + a0, (uint32_t)cnt * 4u, &S->full[rg.slot]);
                }
                bulk_g2s(dst + (size_t)nnz_cap * 20, P.rowptr + rs, (uint32_t)rpc * 4u, &S->full[rg.slot]);
            }
            rg.advance();
            ++issued;
        }
        for (int cb = lo; cb <= hi && !stopped; cb += CB) {
            const int nb = min(CB, hi - cb + 1);
            for (int k = 0; k < G.ntk && !stopped; ++k) {
                const int rows = min(G.TR, G.nrows - k * G.TR);
                for (int u = 0; u < nb; ++u) {
                    const int col = cb + u;
                    if (!prodz_wait_col(S, col) || !prodz_acquire(S, rg)) { stopped = true; break; }
                    mbar_arrive_expect_tx(&S->full[rg.slot], (uint32_t)rows * 16u);
                    bulk_g2s(rg.ptr(), P.V + (long long)col * ldv + G.r0 + (long long)k * G.TR, (uint32_t)rows * 16u,
                             &S->full[rg.slot]);
                    rg.advance();
                    ++issued;
                }
            }
        }
        // update tiles: the exact reverse of the dots order (batches, tiles and columns descending) -- see update_z
        const int nbatch = (hi - lo) / CB + 1;
        for (int bi = nbatch - 1; bi >= 0 && !stopped; --bi) {
            const int c0 = bi == 0 ? ulo : lo + bi * CB;
            const int c1 = min(lo + bi * CB + CB - 1, hi);
            for (int k = G.ntk - 1; k >= 0 && !stopped; --k) {
                const int rows = min(G.TR, G.nrows - k * G.TR);
                for (int col = c1; col >= c0; --col) {
                    if (!prodz_acquire(S, rg)) { stopped = true; break; }
                    mbar_arrive_expect_tx(&S->full[rg.slot], (uint32_t)rows * 16u);
                    bulk_g2s(rg.ptr(), P.V + (long long)col * ldv + G.r0 + (long long)k * G.TR, (uint32_t)rows * 16u,
                             &S->full[rg.slot]);
                    rg.advance();
                    ++issued;
                }
            }
        }
    }
    while (flag_get(&S->stop_seq) < 1) __nanosleep(256);
    // every copy that was issued must have landed before the CTA exits
    const unsigned ns = (unsigned)rg.nslot;
    const unsigned first = issued > ns ? issued - ns : 0u;
    for (unsigned t = first; t < issued; ++t) mbar_wait(&S->full[t % ns], (t / ns) & 1u);
}

// ---- consumers (512 threads) --------------------------------------------------------------------------------------------
__device__ __forceinline__ void team_barrier_zc(Team &tm) {
    consumer_sync();
    if (threadIdx.x == 0) {
        tm.target += (unsigned)tm.C;
        __threadfence();
        atomicAdd(tm.bar, 1u);
        while ((int)(ld_acquire_u32(tm.bar) - tm.target) < 0) {
        }
        __threadfence();
    }
    consumer_sync();
}

__device__ __forceinline__ void block_sum_to_zc(ConsZ &cx, double v, double *out) {
    v = warp_sum(v);
    if (cx.lane == 0) cx.S->redn[cx.warp] = v;
    consumer_sync();
    if (cx.tid == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += cx.S->redn[w];
        *out = s;
    }
}

// ws = xscale * (A x)[slice]; two lanes per row (even / odd entries), gathers of a batch in flight before the first FMA
__device__ void matvec_z(const KrylovParamsZ &P, ConsZ &cx, const TmaGeom &G, const double2 *xsrc, double xscale) {
    SmemTmaZ *S = cx.S;
    const int tid = cx.tid;
    const int nnz_cap = P.nnz_cap;
    const int rc = tid >> 1, sub = tid & 1;
    for (int c = 0; c < G.nch; ++c) {
        const int rl = c * P.ch_rows + rc;
        const bool active = rc < P.ch_rows && rl < G.nrows;
        cx.wait_full();
        double2 sum = make_double2(0.0, 0.0);
        if (active) {
            const unsigned char *base = cx.rg.ptr();
            const double2 *vs = reinterpret_cast<const double2 *>(base);
            const int *cs = reinterpret_cast<const int *>(base + (size_t)nnz_cap * 16);
            const int *rp = reinterpret_cast<const int *>(base + (size_t)nnz_cap * 20);
            const int a0 = S->slot_a0[cx.rg.slot];
            const int e0 = rp[rc] - a0, e1 = rp[rc + 1] - a0;
            for (int eb = e0 + sub; eb < e1; eb += 2 * 4) {
                double2 av[4], xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool ok = eb + 2 * u < e1;
                    av[u] = make_double2(0.0, 0.0);
                    xv[u] = make_double2(0.0, 0.0);
                    if (ok) {
                        av[u] = vs[eb + 2 * u];
                        xv[u] = xsrc[cs[eb + 2 * u]];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) sum = zfma(av[u], xv[u], sum);
            }
        }
        // (pairs never straddle a warp; inactive pairs add zeros)
        sum.x += __shfl_xor_sync(0xffffffffu, sum.x, 1);
        sum.y += __shfl_xor_sync(0xffffffffu, sum.y, 1);
        if (active && sub == 0) cx.ws[rl] = make_double2(sum.x * xscale, sum.y * xscale);
        cx.release();
    }
    consumer_sync();
}

// per-CTA partials of h_c = <v_c, w> = sum conj(v_c) w for c = lo..hi -> part[(c - lo) * CPAD + rank]
__device__ void dots_z(const KrylovParamsZ &P, ConsZ &cx, const TmaGeom &G, const Team &tm, int lo, int hi, double2 *part) {
    SmemTmaZ *S = cx.S;
    const int tid = cx.tid, lane = cx.lane, warp = cx.warp;
    int batch = 0;
    for (int cb = lo; cb <= hi; cb += CB, ++batch) {
        const int nb = min(CB, hi - cb + 1);
        double are[CB], aim[CB];
#pragma unroll
        for (int u = 0; u < CB; ++u) are[u] = aim[u] = 0.0;
        for (int k = 0; k < G.ntk; ++k) {
            const int rows = min(G.TR, G.nrows - k * G.TR);
            const int rbase = k * G.TR;
            double2 wr[PPTZ];
#pragma unroll
            for (int q = 0; q < PPTZ; ++q) {
                const int idx = tid + q * NTC;
                wr[q] = idx < rows ? cx.ws[rbase + idx] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < CB; ++u) {
                if (u < nb) {
                    cx.wait_full();
                    const double2 *vt = reinterpret_cast<const double2 *>(cx.rg.ptr());
#pragma unroll
                    for (int q = 0; q < PPTZ; ++q) {
                        const int idx = tid + q * NTC;
                        if (idx < rows) {
                            const double2 v1 = vt[idx];
                            are[u] = fma(v1.x, wr[q].x, fma(v1.y, wr[q].y, are[u]));
                            aim[u] = fma(v1.x, wr[q].y, fma(-v1.y, wr[q].x, aim[u]));
                        }
                    }
                    cx.release();
                }
            }
        }
        const double rr = warp_reduce8(are, lane);
        const double ri = warp_reduce8(aim, lane);
        const int buf = batch & 1;
        if ((lane & 3) == 0)
            S->red[buf][warp][((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = make_double2(rr, ri);
        consumer_sync();
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
}

// w -= sum_c h_c v_c (c = hi..ulo), partial ||w||^2, unnormalised w to the gather buffer.
// Traversal = the exact reverse of the dots phase at tile granularity (column batches descending, row tiles descending,
// columns descending), so the update starts with what the dots phase touched last: under LRU everything among the last
// ~100 MB of the dots traffic is an L2 hit (six 16 MB columns at n = 10^6).  A row-tile-outer loop (the real kernel's
// order) lost almost all of that with four tiles per slice: L2 hit rate 8 %, DRAM traffic 18.1 GB of 19.0 GB algorithmic
// (profiles/r2_ncu_full_complex_arnoldi_summary.json).  The price is one shared-memory load + store of the w tile per
// (batch, tile) instead of per tile.
__device__ double update_z(const KrylovParamsZ &P, ConsZ &cx, const TmaGeom &G, int lo, int ulo, int hi, double2 *xout) {
    SmemTmaZ *S = cx.S;
    const int tid = cx.tid;
    double nrm = 0.0;
    const int nbatch = (hi - lo) / CB + 1;
    for (int bi = nbatch - 1; bi >= 0; --bi) {
        const int c0 = bi == 0 ? ulo : lo + bi * CB;
        const int c1 = min(lo + bi * CB + CB - 1, hi);
        for (int k = G.ntk - 1; k >= 0; --k) {
            const int rows = min(G.TR, G.nrows - k * G.TR);
            const int rbase = k * G.TR;
            double2 wr[PPTZ];
#pragma unroll
            for (int q = 0; q < PPTZ; ++q) {
                const int idx = tid + q * NTC;
                wr[q] = idx < rows ? cx.ws[rbase + idx] : make_double2(0.0, 0.0);
            }
            for (int col = c1; col >= c0; --col) {
                const double2 hc = S->hs[col - ulo];
                const double2 mh = make_double2(-hc.x, -hc.y);
                cx.wait_full();
                const double2 *vt = reinterpret_cast<const double2 *>(cx.rg.ptr());
#pragma unroll
                for (int q = 0; q < PPTZ; ++q) {
                    const int idx = tid + q * NTC;
                    if (idx < rows) wr[q] = zfma(mh, vt[idx], wr[q]);
                }
                cx.release();
            }
            if (bi == 0) {  // last batch: the tile is final
#pragma unroll
                for (int q = 0; q < PPTZ; ++q) {
                    const int idx = tid + q * NTC;
                    if (idx < rows) {
                        cx.ws[rbase + idx] = wr[q];
                        xout[G.r0 + rbase + idx] = wr[q];
                        nrm = fma(wr[q].x, wr[q].x, fma(wr[q].y, wr[q].y, nrm));
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < PPTZ; ++q) {
                    const int idx = tid + q * NTC;
                    if (idx < rows) cx.ws[rbase + idx] = wr[q];
                }
            }
        }
    }
    return nrm;
}

__device__ void consumer_z(const KrylovParamsZ &P, ConsZ &cx, const TmaGeom &G, Team &tm) {
    SmemTmaZ *S = cx.S;
    const int tid = cx.tid, lane = cx.lane, warp = cx.warp;
    double2 *V = P.V;
    const long long ldv = P.ldv;
    const int nrows = G.nrows, r0 = G.r0;
    double2 *xb0 = P.xbuf, *xb1 = P.xbuf + P.xlen;
    const double2 *xsrc;
    double xscale;
    int jstart;
    int m_out = P.m, breakdown = 0;

    if (P.j0 == 0) {  // firststep!
        double nrm = 0.0;
        for (int i = tid; i < nrows; i += NTC) {
            const double2 b1 = P.b[r0 + i];
            cx.ws[i] = b1;
            nrm = fma(b1.x, b1.x, fma(b1.y, b1.y, nrm));
        }
        block_sum_to_zc(cx, nrm, P.partn + 2 * CPAD + tm.rank);
        team_barrier_zc(tm);
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
        for (int i = tid; i < nrows; i += NTC) {
            double2 b1 = cx.ws[i];
            b1.x *= inv;
            b1.y *= inv;
            V[r0 + i] = b1;
        }
        fence_proxy_async();
        consumer_sync();
        if (tid == 0) flag_set(&S->cols_ready, 1);
        xsrc = P.b;
        xscale = inv;
        jstart = 1;
    } else {
        xsrc = V + (long long)(P.j0 - 1) * ldv;
        xscale = 1.0;
        jstart = P.j0;
    }

    double beta_prev = 0.0;
    const int iopw = P.iop > 0 ? P.iop : P.m;
    for (int j = jstart; j <= P.m; ++j) {
        const int jc = j - 1;
        const int par = j & 1;
        double2 *xout = par ? xb1 : xb0;
        double2 *part = P.part + (long long)par * MAXCOL * CPAD;
        double *partn = P.partn + par * CPAD;

        matvec_z(P, cx, G, xsrc, xscale);

        const int lo = P.lanczos ? jc : max(0, jc - iopw + 1);
        const int hi = jc;
        const int nc = hi - lo + 1;
        const int ulo = (P.lanczos && jc >= 1) ? jc - 1 : lo;
        dots_z(P, cx, G, tm, lo, hi, part);
        team_barrier_zc(tm);
        for (int ci = warp; ci < nc; ci += NW) {
            double2 s = team_sum_z(part + (long long)ci * CPAD, tm.C, lane);
            if (P.lanczos) s.y = 0.0;  // coeff(U <: Real, alpha) = real(alpha)
            if (lane == 0) {
                S->hs[lo + ci - ulo] = s;
                if (tm.rank == 0) P.Hd[(long long)jc * P.ldh + lo + ci] = s;
            }
        }
        if (P.lanczos && jc >= 1 && tid == 0) S->hs[0] = make_double2(beta_prev, 0.0);
        consumer_sync();

        const double nrm = update_z(P, cx, G, lo, ulo, hi, xout);
        block_sum_to_zc(cx, nrm, partn + tm.rank);
        team_barrier_zc(tm);
        const double beta = sqrt(team_sum(partn, tm.C, lane));
        if (tm.rank == 0 && tid == 0) P.Hd[(long long)jc * P.ldh + jc + 1] = make_double2(beta, 0.0);
        {
            double2 *vn = V + (long long)(jc + 1) * ldv;
            for (int i = tid; i < nrows; i += NTC) {
                double2 w1 = cx.ws[i];
                w1.x /= beta;
                w1.y /= beta;
                vn[r0 + i] = w1;
            }
        }
        fence_proxy_async();  // the producer's TMA reads of this column must see these generic-proxy stores
        consumer_sync();
        if (tid == 0) flag_set(&S->cols_ready, jc + 2);
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

__global__ void __launch_bounds__(NT2, 1) krylov_tma_z_kernel(const __grid_constant__ KrylovParamsZ P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemTmaZ *S = reinterpret_cast<SmemTmaZ *>(smem_raw);
    const size_t ws_bytes = ((size_t)P.slice * 16 + 127) & ~(size_t)127;
    double2 *ws_smem = reinterpret_cast<double2 *>(smem_raw + sizeof(SmemTmaZ));
    unsigned char *ring = smem_raw + sizeof(SmemTmaZ) + ws_bytes;

    const int tid = threadIdx.x;
    Team tm;
    tm.rank = blockIdx.x;
    tm.C = P.team_size;
    tm.bar = P.bar;
    tm.target = 0;
    tm.seq = 0;
    TmaGeom G;
    G.r0 = min(P.n, tm.rank * P.slice);
    G.nrows = min(P.n, G.r0 + P.slice) - G.r0;
    G.TR = P.tile_rows;
    G.ntk = (G.nrows + G.TR - 1) / G.TR;
    G.nch = (G.nrows + P.ch_rows - 1) / P.ch_rows;

    if (G.nch > 0 && G.nch <= MAXCH2) {
        for (int c = tid; c < G.nch; c += NT2) {
            const int rs = G.r0 + c * P.ch_rows;
            const int re = min(G.r0 + G.nrows, rs + P.ch_rows);
            const int e0 = P.rowptr[rs], e1 = P.rowptr[re];
            const int a0 = e0 & ~3;
            S->chunk_a0[c] = a0;
            S->chunk_cnt[c] = ((e1 + 3) & ~3) - a0;
        }
    }
    if (tid == 0) {
        for (int s = 0; s < P.nslot; ++s) {
            mbar_init(&S->full[s], 1);
            mbar_init(&S->empty[s], NW);
        }
        S->cols_ready = P.j0;  // continuation: columns 0..j0-1 already exist
        S->stop_seq = 0;
        mbar_fence_init();
    }
    __syncthreads();

    if (tid >= NTC) {
        if (tid == NTC) {
            RingZ rg{ring, P.nslot, 0, 0u, (uint32_t)P.slot_bytes};
            producer_z(P, S, rg, G);
        }
        __syncwarp();
    } else {
        ConsZ cx;
        cx.S = S;
        cx.ws = ws_smem;
        cx.tid = tid;
        cx.lane = tid & 31;
        cx.warp = tid >> 5;
        cx.rg = RingZ{ring, P.nslot, 0, 0u, (uint32_t)P.slot_bytes};
        consumer_z(P, cx, G, tm);
        consumer_sync();
        if (tid == 0) flag_set(&S->stop_seq, 1);
    }
}

}  // namespace b200k
