// krylov_kernel_tma.cuh -- v2 of the persistent fused Arnoldi / Lanczos / IOP kernel: warp-specialised,
// everything that comes from HBM (CSR chunks AND the basis slice) is streamed by one producer warp through
// a shared-memory ring of 32 KB slots with 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) and
// full/empty mbarriers; 16 consumer warps compute out of shared memory.
//
// Why (profiles/r1_v1_*): the LDG version kept only 16 warps x 8 loads in flight per SM and sat at 40 % DRAM
// utilisation, stalled on long_scoreboard and on the two team barriers per step.  With the ring, bytes in
// flight are set by the ring depth (up to 160 KB per SM) instead of by registers, and the producer runs
// ahead of the consumers across phase boundaries and team barriers (the next phase's first tiles land
// while the CTAs are still synchronising).
//
// Same algorithm, same reduction order, same team-barrier protocol and the same outputs as
// krylov_kernel.cuh (which remains the path for odd n / odd ldv / unaligned bases).  Reference semantics:
// src/arnoldi.jl:230-308, 345-377, 388-403, 456-490.
//
// Tile schedule of step j (identical on the producer and on the consumers of a CTA):
//   [A chunks 0..nch-1]                       (CSR stream only; val | colind | rowptr segment per slot)
//   [dots : for cb in lo..hi step 8 : for k in 0..ntk-1 : for u < nb : basis tile (col cb+u, rows k)]
//   [update: for k in 0..ntk-1 : for col = uhi..ulo : basis tile (col, rows k)]
// A basis tile of column c may only be fetched once the consumers have written that column
// (cols_ready > c, published after a generic->async proxy fence).
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "krylov_kernel.cuh"

namespace b200k {

constexpr int NTC = 512;                         // consumer threads (16 warps)
constexpr int NT2 = NTC + 32;                    // + 1 producer warp
constexpr int SLOT_BYTES = 32768;
constexpr int MAXSLOT = 6;
constexpr int TILE_ROWS_MAX = SLOT_BYTES / 8;    // 4096 rows per basis tile
constexpr int PPT = TILE_ROWS_MAX / 2 / NTC;     // row pairs per consumer thread per tile (4)
constexpr int MAXCH2 = 64;                       // chunk-table capacity

struct __align__(128) SmemTma {
    uint64_t full[MAXSLOT];
    uint64_t empty[MAXSLOT];
    double hs[MAXCOL];
    double red[2][NW][CB];
    double redn[NW];
    double wtail[MAXP];
    double xtail[MAXP];
    int chunk_a0[MAXCH2];
    int chunk_cnt[MAXCH2];
    int slot_a0[MAXSLOT];
    double bc[2];             // reduced scalars (squared norm) broadcast to the CTA
    volatile int cols_ready;  // number of complete basis columns of the current problem
    volatile int stop_seq;    // consumers finished local problem #stop_seq (1-based)
};

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTC) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- team barrier and team-wide reduction -----------------------------------------------------------------------
// Single GPU: the cooperative-groups grid-sync protocol on one counter, then every CTA sums the per-CTA partials in
// a fixed order.  Row-sharded over R GPUs (NVLink peer memory, no NCCL): two levels --
//   1. the CTAs of a GPU synchronise on their OWN counter (they stored their partials locally; CTAs that pushed
//      halo values into a peer's gather buffer first make them visible with a system-scope fence);
//   2. for each reduced quantity one CTA sums the GPU-local partials and stores {value, seq} packets straight into
//      every GPU's inbox: one self-validating 16-byte store {lo32, seq, hi32, seq} per peer, the NCCL "LL" idea, so
//      no fence / flag round trip sits between data and notification;
//   3. every CTA polls the R packets of each quantity and adds them in rank order (bitwise identical everywhere).
__device__ __forceinline__ void ll_push(uint4 *p, double v, unsigned seq) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)b), "r"(seq),
                 "r"((unsigned)(b >> 32)), "r"(seq)
                 : "memory");
}
__device__ __forceinline__ double ll_poll(const uint4 *p, unsigned seq) {
    unsigned lo, f1, hi, f2;
    do {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2)
                     : "l"(p)
                     : "memory");
    } while (f1 != seq || f2 != seq);
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

__device__ __forceinline__ void team_barrier_c(Team &tm, const KrylovParams &P, bool pushed_to_peers) {
    consumer_sync();
    if (threadIdx.x == 0) {
        tm.target += (unsigned)tm.C;
        if (pushed_to_peers) __threadfence_system();  // halo stores into peer memory are performed first
        else __threadfence();
        atomicAdd(tm.bar, 1u);
        while ((int)(ld_acquire_u32(tm.bar) - tm.target) < 0) {
        }
        __threadfence();
    }
    consumer_sync();
}

struct Ring {
    unsigned char *base;
    int nslot;
    int slot;
    unsigned phase;
    __device__ __forceinline__ void advance() {
        if (++slot == nslot) {
            slot = 0;
            phase ^= 1u;
        }
    }
    __device__ __forceinline__ unsigned char *ptr() const { return base + (size_t)slot * SLOT_BYTES; }
};

struct TmaGeom {
    int r0, nrows;  // this CTA's slice
    int nch;        // CSR chunks in the slice
    int ntk;        // basis tiles per column in the slice
    int TR;         // rows per basis tile
};

// ---------------------------------------------------------------------------------------------------
// producer (one lane)
// ---------------------------------------------------------------------------------------------------
// The producer is ONE lane of the producer warp (every copy, including the 2-D tensor-map boxes, is a single
// instruction, so there is nothing for the other lanes to do).
__device__ __forceinline__ bool prod_acquire(SmemTma *S, const Ring &rg, int seq, int) {
    while (!mbar_try_wait(&S->empty[rg.slot], rg.phase ^ 1u)) {
        if (S->stop_seq >= seq) return false;
    }
    return true;
}

__device__ __forceinline__ bool prod_wait_col(SmemTma *S, int col, int seq, int) {
    while (S->cols_ready <= col) {
        if (S->stop_seq >= seq) return false;
    }
    return true;
}

template <int OPK, bool AUG>
__device__ void producer_problem(const KrylovParams &P, const CUtensorMap *tmA, SmemTma *S, Ring &rg,
                                 const TmaGeom &G, const double *V, int seq, unsigned &issued, int lane) {
    const long long ldv = P.ldv;
    const int jstart = P.j0 == 0 ? 1 : P.j0;
    const int iopw = P.iop > 0 ? P.iop : P.m;
    const int nnz_cap = P.nnz_cap;
    // L2 eviction priorities (P.l2hint): the operator is streamed once per step and would otherwise push the
    // basis out of the 126 MB L2 between the two Gram-Schmidt passes.
    const uint64_t polA = policy_evict_first();
    const uint64_t polV = policy_evict_last();
    const bool hintV = (P.l2hint & 2) != 0;
    bool stopped = false;
    for (int j = jstart; j <= P.m && !stopped; ++j) {
        const int jc = j - 1;
        // operator chunks are marked evict_first in the steps whose orthogonalisation window is so wide that the
        // operator could not survive in L2 until the next step anyway (hintA_cols from the host: L2 size vs
        // operator + two basis passes); in narrow-window steps (Lanczos, IOP, early Arnoldi) it stays resident.
        const bool hintA = (jc - (P.lanczos ? jc : max(0, jc - iopw + 1)) + 1) >= P.hintA_cols;
        if (OPK == OP_CSR_STREAM) {
            for (int c = 0; c < G.nch; ++c) {
                if (!prod_acquire(S, rg, seq, lane)) { stopped = true; break; }
                if (lane == 0) {
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
                    mbar_arrive_expect_tx(&S->full[rg.slot], (uint32_t)cnt * 12u + (uint32_t)rpc * 4u);
                    if (hintA) {
                        if (cnt > 0) {
                            bulk_g2s_hint(dst, P.val + a0, (uint32_t)cnt * 8u, &S->full[rg.slot], polA);
                            bulk_g2s_hint(dst + (size_t)nnz_cap * 8, P.colind + a0, (uint32_t)cnt * 4u,
                                          &S->full[rg.slot], polA);
                        }
                        bulk_g2s_hint(dst + (size_t)nnz_cap * 12, P.rowptr + rs, (uint32_t)rpc * 4u, &S->full[rg.slot],
                                      polA);
                    } else {
                        if (cnt > 0) {
                            bulk_g2s(dst, P.val + a0, (uint32_t)cnt * 8u, &S->full[rg.slot]);
                            bulk_g2s(dst + (size_t)nnz_cap * 8, P.colind + a0, (uint32_t)cnt * 4u, &S->full[rg.slot]);
                        }
                        bulk_g2s(dst + (size_t)nnz_cap * 12, P.rowptr + rs, (uint32_t)rpc * 4u, &S->full[rg.slot]);
                    }
                }
                rg.advance();
                ++issued;
            }
            if (stopped) break;
        } else if (OPK == OP_DENSE && P.dense_cpt > 0 && G.nrows > 0) {
            // dense operator: a tile = dense_cpt columns x the CTA's row slice, fetched as slice/box_rows
            // tensor-map boxes (rows past n are zero-filled by the TMA unit and still count as bytes)
            const int nrb = P.slice / P.dense_box_rows;
            const uint32_t tbytes = (uint32_t)P.slice * (uint32_t)P.dense_cpt * 8u;
            for (int c0 = 0; c0 < P.n; c0 += P.dense_cpt) {
                if (!prod_acquire(S, rg, seq, lane)) { stopped = true; break; }
                if (lane == 0) {
                    mbar_arrive_expect_tx(&S->full[rg.slot], tbytes);
                    unsigned char *dst = rg.ptr();
                    for (int rb = 0; rb < nrb; ++rb)
                        tma_load_2d(dst + (size_t)rb * P.dense_box_rows * P.dense_cpt * 8, tmA,
                                    G.r0 + rb * P.dense_box_rows, c0, &S->full[rg.slot]);
                }
                rg.advance();
                ++issued;
            }
            if (stopped) break;
        }
        const int lo = P.lanczos ? jc : max(0, jc - iopw + 1);
        const int hi = jc;
        const int ulo = (P.lanczos && jc >= 1) ? jc - 1 : lo;
        for (int cb = lo; cb <= hi && !stopped; cb += CB) {
            const int nb = min(CB, hi - cb + 1);
            for (int k = 0; k < G.ntk && !stopped; ++k) {
                const int rows = min(G.TR, G.nrows - k * G.TR);
                for (int u = 0; u < nb; ++u) {
                    const int col = cb + u;
                    if (!prod_wait_col(S, col, seq, lane) || !prod_acquire(S, rg, seq, lane)) { stopped = true; break; }
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&S->full[rg.slot], (uint32_t)rows * 8u);
                        {
                        const double *src = V + (long long)col * ldv + G.r0 + (long long)k * G.TR;
                        if (hintV) bulk_g2s_hint(rg.ptr(), src, (uint32_t)rows * 8u, &S->full[rg.slot], polV);
                        else bulk_g2s(rg.ptr(), src, (uint32_t)rows * 8u, &S->full[rg.slot]);
                    }
                    }
                    rg.advance();
                    ++issued;
                }
            }
        }
        for (int k = 0; k < G.ntk && !stopped; ++k) {
            const int rows = min(G.TR, G.nrows - k * G.TR);
            for (int col = hi; col >= ulo; --col) {
                if (!prod_acquire(S, rg, seq, lane)) { stopped = true; break; }
                if (lane == 0) {
                    mbar_arrive_expect_tx(&S->full[rg.slot], (uint32_t)rows * 8u);
                    {
                        const double *src = V + (long long)col * ldv + G.r0 + (long long)k * G.TR;
                        if (hintV) bulk_g2s_hint(rg.ptr(), src, (uint32_t)rows * 8u, &S->full[rg.slot], polV);
                        else bulk_g2s(rg.ptr(), src, (uint32_t)rows * 8u, &S->full[rg.slot]);
                    }
                }
                rg.advance();
                ++issued;
            }
        }
    }
    // the consumers decide when the problem is over (m steps, happy breakdown, or beta == 0)
    while (S->stop_seq < seq) {
    }
    // every copy that was issued must have landed before the ring is re-initialised / the CTA exits
    const unsigned ns = (unsigned)rg.nslot;
    const unsigned first = issued > ns ? issued - ns : 0u;
    for (unsigned t = first; t < issued; ++t) mbar_wait(&S->full[t % ns], (t / ns) & 1u);
}

// ---------------------------------------------------------------------------------------------------
// consumers (512 threads)
// ---------------------------------------------------------------------------------------------------
struct Cons {
    SmemTma *S;
    double *ws;
    int tid, lane, warp;
    unsigned seq;  // team barriers passed so far (the LL packets of barrier #seq carry it)
    Ring rg;
    __device__ __forceinline__ void wait_full() { mbar_wait(&S->full[rg.slot], rg.phase); }
    __device__ __forceinline__ void release() {
        __syncwarp();
        if (lane == 0) mbar_arrive(&S->empty[rg.slot]);
        rg.advance();
    }
};

// CTA-wide deterministic sum; thread 0 stores it at offset `off` of this GPU's norm table.
__device__ __forceinline__ void block_sum_to_c(const KrylovParams &P, Cons &cx, double v, long long off) {
    v = warp_sum(v);
    if (cx.lane == 0) cx.S->redn[cx.warp] = v;
    consumer_sync();
    if (cx.tid == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += cx.S->redn[w];
        P.peer_partn[P.myrank][off] = s;
    }
}

// Team barrier + reduction of `ncols` quantities whose per-CTA partials sit in this GPU's table `ltab`
// (row ci at ltab + ci*cpad, entry = CTA rank).  out[ci] (shared memory) = sum over the whole team.
__device__ void team_reduce_c(const KrylovParams &P, Cons &cx, Team &tm, const double *ltab, int ncols, double *out,
                              bool pushed_to_peers) {
    team_barrier_c(tm, P, pushed_to_peers);
    cx.seq += 1u;
    if (P.nranks == 1) {
        for (int ci = cx.warp; ci < ncols; ci += NW) {
            const double s = team_sum(ltab + (long long)ci * P.cpad, tm.C, cx.lane);
            if (cx.lane == 0) out[ci] = s;
        }
    } else {
        const unsigned seq = cx.seq;
        const long long pbase = (long long)(seq & 1u) * (MAXCOL + 1) * 8;
        if (cx.warp == 0) {  // this GPU's sum of quantity ci goes to every GPU's inbox
            for (int ci = tm.rank; ci < ncols; ci += tm.C) {
                const double s = team_sum(ltab + (long long)ci * P.cpad, tm.C, cx.lane);
                if (cx.lane < P.nranks) ll_push(P.peer_pkt[cx.lane] + pbase + (long long)ci * 8 + P.myrank, s, seq);
            }
        }
        for (int ci = cx.warp; ci < ncols; ci += NW) {
            double v = 0.0;
            if (cx.lane < P.nranks) v = ll_poll(P.peer_pkt[P.myrank] + pbase + (long long)ci * 8 + cx.lane, seq);
            double s = 0.0;
            for (int r = 0; r < P.nranks; ++r) s += __shfl_sync(0xffffffffu, v, r);
            if (cx.lane == 0) out[ci] = s;
        }
        consumer_sync();
        if (cx.tid == 0) __threadfence_system();  // L1 is invalidated before anyone gathers pushed halo values
    }
    consumer_sync();
}

// Push this CTA's rows that other GPUs gather (halo) into their gather buffers (peer stores over NVLink).
__device__ __forceinline__ void push_halo(const KrylovParams &P, Cons &cx, const TmaGeom &G, const Team &tm,
                                          long long xoff) {
    if (P.nranks == 1) return;
    const int e1 = P.send_ofs[tm.rank + 1];
    for (int e = P.send_ofs[tm.rank] + cx.tid; e < e1; e += NTC)
        P.peer_xbuf[P.send_peer[e]][xoff + P.send_pos[e]] = cx.ws[P.send_row[e] - G.r0];
}

template <int OPK, bool AUG>
__device__ void matvec_phase_c(const KrylovParams &P, Cons &cx, const TmaGeom &G, const double *xsrc, double xscale) {
    SmemTma *S = cx.S;
    const int tid = cx.tid, lane = cx.lane, warp = cx.warp;
    const int n = P.n, p = AUG ? P.p : 0;
    double *ws = cx.ws;
    if (p > 0) {
        if (tid < p) S->xtail[tid] = xsrc[n + P.nhalo + tid];
        consumer_sync();
        if (tid < p) S->wtail[tid] = (tid < p - 1) ? S->xtail[tid + 1] * xscale : 0.0;
    }
    if (OPK == OP_CSR_STREAM) {
        const int nnz_cap = P.nnz_cap;
        for (int c = 0; c < G.nch; ++c) {
            const int rl = c * P.ch_rows + tid;
            const bool active = tid < P.ch_rows && rl < G.nrows;
            cx.wait_full();
            if (active) {
                const unsigned char *base = cx.rg.ptr();
                const double *vs = reinterpret_cast<const double *>(base);
                const int *cs = reinterpret_cast<const int *>(base + (size_t)nnz_cap * 8);
                const int *rp = reinterpret_cast<const int *>(base + (size_t)nnz_cap * 12);
                const int a0 = S->slot_a0[cx.rg.slot];
                const int e0 = rp[tid] - a0, e1 = rp[tid + 1] - a0;
                double sum = 0.0;
                // gather in batches of 8: all x loads of a batch are in flight before the first FMA needs one.
                // (Software-pipelining the gathers across chunks was measured and is slower: 1.542 vs 1.520 ms on
                // C2 Arnoldi, 0.687 vs 0.658 ms Lanczos -- the phase is bound by the operator stream, not by gather
                // latency -- and carrying both loops in one kernel cost 3-5 % everywhere through code size.)
                for (int eb = e0; eb < e1; eb += 8) {
                    double av[8], xv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const bool ok = eb + u < e1;
                        av[u] = ok ? vs[eb + u] : 0.0;
                        xv[u] = 0.0;
                        if (ok) xv[u] = xsrc[cs[eb + u]];
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) sum = fma(av[u], xv[u], sum);
                }
                if (p > 0) {
                    const double *brow = P.Bm + (G.r0 + rl);
                    for (int k = 0; k < p; ++k) sum = fma(brow[(long long)k * P.ldb], S->xtail[k], sum);
                }
                ws[rl] = sum * xscale;
            }
            cx.release();
        }
    } else if (OPK == OP_CSR_WARP) {
        for (int rl = warp; rl < G.nrows; rl += NW) {
            const int row = G.r0 + rl;
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
    } else {  // dense column-major
        const int units = G.nrows / 2;
        int RL = 32;
        while (RL < units && RL < NTC) RL <<= 1;
        const int Gc = NTC / RL;
        const int ul = tid % RL, g = tid / RL;
        // NTC*2 doubles of reduction scratch behind the ring (the ring slots may have basis tiles in flight)
        double *dscratch = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(S) + P.dscratch_off);
        if (P.dense_cpt > 0) {
            // operator tiles arrive through the TMA ring: [dense_cpt columns][nrows] per slot
            double a0 = 0.0, a1 = 0.0;
            if (G.nrows > 0) {
                for (int c0 = 0; c0 < n; c0 += P.dense_cpt) {
                    const int nc = min(P.dense_cpt, n - c0);
                    cx.wait_full();
                    if (ul < units) {
                        // smem tile: [row box][column][box rows]; this thread's row pair sits in box rb
                        const int bh = P.dense_box_rows >> 1;  // row pairs per box
                        const int rb = ul / bh;
                        const double2 *t2 = reinterpret_cast<const double2 *>(cx.rg.ptr()) +
                                            (size_t)rb * bh * P.dense_cpt + (ul - rb * bh);
#pragma unroll 4
                        for (int cc = g; cc < nc; cc += Gc) {
                            const double2 a2 = t2[(size_t)cc * bh];
                            const double xc = xsrc[c0 + cc];
                            a0 = fma(a2.x, xc, a0);
                            a1 = fma(a2.y, xc, a1);
                        }
                    }
                    cx.release();
                }
            }
            dscratch[(g * RL + ul) * 2 + 0] = a0;
            dscratch[(g * RL + ul) * 2 + 1] = a1;
            consumer_sync();
            if (g == 0 && ul < units) {
                double s0 = 0.0, s1 = 0.0;
                for (int q = 0; q < Gc; ++q) {
                    s0 += dscratch[(q * RL + ul) * 2 + 0];
                    s1 += dscratch[(q * RL + ul) * 2 + 1];
                }
                const int rl = 2 * ul;
                if (p > 0) {
                    for (int k = 0; k < p; ++k) {
                        s0 = fma(P.Bm[G.r0 + rl + (long long)k * P.ldb], S->xtail[k], s0);
                        s1 = fma(P.Bm[G.r0 + rl + 1 + (long long)k * P.ldb], S->xtail[k], s1);
                    }
                }
                ws[rl] = s0 * xscale;
                ws[rl + 1] = s1 * xscale;
            }
        } else {
            // direct 16-byte loads (slices with more than 1024 rows per CTA)
            for (int ubase = 0; ubase < units; ubase += RL) {
                const int u = ubase + ul;
                const bool valid = u < units;
                double a0 = 0.0, a1 = 0.0;
                if (valid) {
                    const double *ap = P.Ad + G.r0 + 2LL * u;
#pragma unroll 8
                    for (int c = g; c < n; c += Gc) {
                        const double xc = xsrc[c];
                        const double2 a2 = ld_ro2(ap + (long long)c * P.lda);
                        a0 = fma(a2.x, xc, a0);
                        a1 = fma(a2.y, xc, a1);
                    }
                }
                dscratch[(g * RL + ul) * 2 + 0] = a0;
                dscratch[(g * RL + ul) * 2 + 1] = a1;
                consumer_sync();
                if (g == 0 && valid) {
                    double s0 = 0.0, s1 = 0.0;
                    for (int q = 0; q < Gc; ++q) {
                        s0 += dscratch[(q * RL + ul) * 2 + 0];
                        s1 += dscratch[(q * RL + ul) * 2 + 1];
                    }
                    const int rl = 2 * u;
                    if (p > 0) {
                        for (int k = 0; k < p; ++k) {
                            s0 = fma(P.Bm[G.r0 + rl + (long long)k * P.ldb], S->xtail[k], s0);
                            s1 = fma(P.Bm[G.r0 + rl + 1 + (long long)k * P.ldb], S->xtail[k], s1);
                        }
                    }
                    ws[rl] = s0 * xscale;
                    ws[rl + 1] = s1 * xscale;
                }
                consumer_sync();
            }
        }
    }
    consumer_sync();
}

template <int OPK, bool AUG>
__device__ void dots_phase_c(const KrylovParams &P, Cons &cx, const TmaGeom &G, const Team &tm, const double *V,
                             int lo, int hi, long long part_off) {
    SmemTma *S = cx.S;
    const int tid = cx.tid, lane = cx.lane, warp = cx.warp;
    const double2 *ws2 = reinterpret_cast<const double2 *>(cx.ws);
    int batch = 0;
    for (int cb = lo; cb <= hi; cb += CB, ++batch) {
        const int nb = min(CB, hi - cb + 1);
        double acc[CB];
#pragma unroll
        for (int u = 0; u < CB; ++u) acc[u] = 0.0;
        for (int k = 0; k < G.ntk; ++k) {
            const int pairs = min(G.TR, G.nrows - k * G.TR) >> 1;
            const int pbase = (k * G.TR) >> 1;
            double2 wr[PPT];
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                const int idx = tid + q * NTC;
                wr[q] = idx < pairs ? ws2[pbase + idx] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < CB; ++u) {
                if (u < nb) {
                    cx.wait_full();
                    const double2 *vt = reinterpret_cast<const double2 *>(cx.rg.ptr());
#pragma unroll
                    for (int q = 0; q < PPT; ++q) {
                        const int idx = tid + q * NTC;
                        if (idx < pairs) {
                            const double2 v2 = vt[idx];
                            acc[u] = fma(v2.x, wr[q].x, fma(v2.y, wr[q].y, acc[u]));
                        }
                    }
                    cx.release();
                }
            }
        }
        if (AUG && P.p > 0 && tm.rank == 0 && P.myrank == 0 && tid == 0) {  // augmented tail rows (direct loads)
#pragma unroll
            for (int u = 0; u < CB; ++u)
                if (u < nb)
                    for (int kk = 0; kk < P.p; ++kk)
                        acc[u] = fma(V[(long long)(cb + u) * P.ldv + P.n + kk], S->wtail[kk], acc[u]);
        }
        const double r = warp_reduce8(acc, lane);
        const int buf = batch & 1;
        if ((lane & 3) == 0) S->red[buf][warp][((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = r;
        consumer_sync();
        if (tid < nb) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += S->red[buf][w][tid];
            P.peer_part[P.myrank][part_off + (long long)(cb - lo + tid) * P.cpad + tm.rank] = s;
        }
    }
}

template <int OPK, bool AUG>
__device__ double update_phase_c(const KrylovParams &P, Cons &cx, const TmaGeom &G, const Team &tm, const double *V,
                                 int ulo, int uhi, double *xout) {
    SmemTma *S = cx.S;
    const int tid = cx.tid;
    double2 *ws2 = reinterpret_cast<double2 *>(cx.ws);
    double2 *xo2 = reinterpret_cast<double2 *>(xout + G.r0);
    const double *hs = S->hs;
    double nrm = 0.0;
    for (int k = 0; k < G.ntk; ++k) {
        const int pairs = min(G.TR, G.nrows - k * G.TR) >> 1;
        const int pbase = (k * G.TR) >> 1;
        double2 wr[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            const int idx = tid + q * NTC;
            wr[q] = idx < pairs ? ws2[pbase + idx] : make_double2(0.0, 0.0);
        }
        for (int col = uhi; col >= ulo; --col) {
            const double hc = hs[col - ulo];
            cx.wait_full();
            const double2 *vt = reinterpret_cast<const double2 *>(cx.rg.ptr());
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                const int idx = tid + q * NTC;
                if (idx < pairs) {
                    const double2 v2 = vt[idx];
                    wr[q].x = fma(-hc, v2.x, wr[q].x);
                    wr[q].y = fma(-hc, v2.y, wr[q].y);
                }
            }
            cx.release();
        }
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            const int idx = tid + q * NTC;
            if (idx < pairs) {
                ws2[pbase + idx] = wr[q];
                xo2[pbase + idx] = wr[q];
                nrm = fma(wr[q].x, wr[q].x, fma(wr[q].y, wr[q].y, nrm));
            }
        }
    }
    if (AUG && P.p > 0 && tid < P.p) {
        double wt = S->wtail[tid];
        for (int c = uhi; c >= ulo; --c) wt = fma(-hs[c - ulo], V[(long long)c * P.ldv + P.n + tid], wt);
        S->wtail[tid] = wt;
        if (tm.rank == 0) {
            xout[P.n + P.nhalo + tid] = wt;
            if (P.myrank == 0) nrm = fma(wt, wt, nrm);
        }
    }
    return nrm;
}

// One problem on the consumer side.  Mirrors krylov_body<2> of krylov_kernel.cuh.
template <int OPK, bool AUG>
__device__ void consumer_problem(const KrylovParams &P, Cons &cx, const TmaGeom &G, Team &tm, int prob, int nlocal,
                                 double *xb0, double *xb1, long long xoff0, long long part0, long long partn0) {
    SmemTma *S = cx.S;
    const int tid = cx.tid;
    const int n = P.n, p = AUG ? P.p : 0;
    double *V = P.V + (long long)prob * P.V_stride;
    double *Hd = P.Hd + (long long)prob * P.H_stride;
    const double *b = P.b + (long long)prob * P.b_stride;
    const long long ldv = P.ldv;
    const int ldh = P.ldh;
    const int units = G.nrows >> 1;
    double2 *ws2 = reinterpret_cast<double2 *>(cx.ws);
    const double *xsrc;
    double xscale;
    int jstart;
    int m_out = P.m, breakdown = 0;
    const bool sharded = P.nranks > 1;
    const bool via_xb0 = p > 0 || sharded;            // first gather source must carry tail / halo entries
    const double *lpart = P.peer_part[P.myrank];      // this GPU's inboxes
    const double *lpartn = P.peer_partn[P.myrank];
    const int xt = n + P.nhalo;                       // offset of the augmented tail in the gather buffers

    if (P.j0 == 0) {  // firststep! (arnoldi.jl:230-250 / 257-279)
        double nrm = 0.0;
        for (int i = tid; i < units; i += NTC) {
            const double2 b2 = reinterpret_cast<const double2 *>(b + G.r0)[i];
            ws2[i] = b2;
            if (via_xb0) reinterpret_cast<double2 *>(xb0 + G.r0)[i] = b2;
            nrm = fma(b2.x, b2.x, fma(b2.y, b2.y, nrm));
        }
        if (p > 0 && tm.rank == 0 && tid < p) {
            const double bt = P.btail[tid];
            xb0[xt + tid] = bt;
            if (P.myrank == 0) nrm = fma(bt, bt, nrm);
        }
        const long long pslot = partn0 + (long long)(2 + (nlocal & 1)) * P.cpad;
        block_sum_to_c(P, cx, nrm, pslot + tm.rank);
        push_halo(P, cx, G, tm, xoff0);
        team_reduce_c(P, cx, tm, lpartn + pslot, 1, S->bc, sharded);
        const double beta = sqrt(S->bc[0]);
        if (tm.rank == 0 && tid == 0) P.scal[prob * 4] = beta;
        if (beta == 0.0) {
            if (tm.rank == 0 && tid == 0) {
                P.stat[prob * 4 + 0] = P.m;
                P.stat[prob * 4 + 1] = 0;
            }
            return;
        }
        if (p == 0) {
            const double inv = 1.0 / beta;
            for (int i = tid; i < units; i += NTC) {
                double2 b2 = ws2[i];
                b2.x *= inv;
                b2.y *= inv;
                reinterpret_cast<double2 *>(V + G.r0)[i] = b2;
            }
            xsrc = via_xb0 ? xb0 : b;
        } else {
            for (int i = tid; i < units; i += NTC) {
                double2 b2 = ws2[i];
                b2.x /= beta;
                b2.y /= beta;
                reinterpret_cast<double2 *>(V + G.r0)[i] = b2;
            }
            if (tm.rank == 0 && tid < p) V[n + tid] = P.btail[tid] / beta;
            xsrc = xb0;
        }
        fence_proxy_async();
        consumer_sync();
        if (tid == 0) S->cols_ready = 1;
        xscale = 1.0 / beta;
        jstart = 1;
    } else {
        const double *vj = V + (long long)(P.j0 - 1) * ldv;
        if (sharded) {  // the resumed column has no halo: stage it in the gather buffer and push the halo
            for (int i = tid; i < units; i += NTC) {
                const double2 v2 = reinterpret_cast<const double2 *>(vj + G.r0)[i];
                ws2[i] = v2;
                reinterpret_cast<double2 *>(xb0 + G.r0)[i] = v2;
            }
            if (p > 0 && tm.rank == 0 && tid < p) xb0[xt + tid] = vj[n + tid];
            block_sum_to_c(P, cx, 0.0, partn0 + 2LL * P.cpad + tm.rank);  // (has the CTA barrier the push needs)
            push_halo(P, cx, G, tm, xoff0);
            team_reduce_c(P, cx, tm, lpartn + partn0 + 2LL * P.cpad, 1, S->bc, true);  // every GPU's halo is in place
            xsrc = xb0;
        } else {
            xsrc = vj;
        }
        xscale = 1.0;
        jstart = P.j0;
    }

    double beta_prev = 0.0;
    const int iopw = P.iop > 0 ? P.iop : P.m;
    for (int j = jstart; j <= P.m; ++j) {
        const int jc = j - 1;
        const int par = j & 1;
        double *xout = par ? xb1 : xb0;
        const long long xoff = xoff0 + (par ? P.xlen : 0);
        const long long part = part0 + (long long)par * MAXCOL * P.cpad;
        const long long partn = partn0 + (long long)par * P.cpad;

        matvec_phase_c<OPK, AUG>(P, cx, G, xsrc, xscale);

        const int lo = P.lanczos ? jc : max(0, jc - iopw + 1);
        const int hi = jc;
        dots_phase_c<OPK, AUG>(P, cx, G, tm, V, lo, hi, part);
        const int nc = hi - lo + 1;
        const int ulo = (P.lanczos && jc >= 1) ? jc - 1 : lo;
        team_reduce_c(P, cx, tm, lpart + part, nc, S->hs + (lo - ulo), false);
        if (tm.rank == 0)
            for (int ci = tid; ci < nc; ci += NTC) Hd[(long long)jc * ldh + lo + ci] = S->hs[lo + ci - ulo];
        if (P.lanczos && jc >= 1 && tid == 0) S->hs[0] = beta_prev;
        consumer_sync();

        const double nrm = update_phase_c<OPK, AUG>(P, cx, G, tm, V, ulo, hi, xout);
        block_sum_to_c(P, cx, nrm, partn + tm.rank);
        push_halo(P, cx, G, tm, xoff);  // after block_sum's CTA barrier: the whole w slice is in place
        team_reduce_c(P, cx, tm, lpartn + partn, 1, S->bc, sharded);

        const double beta = sqrt(S->bc[0]);
        if (tm.rank == 0 && tid == 0) Hd[(long long)jc * ldh + jc + 1] = beta;
        {
            double *vn = V + (long long)(jc + 1) * ldv;
            for (int i = tid; i < units; i += NTC) {
                double2 w2 = ws2[i];
                w2.x /= beta;
                w2.y /= beta;
                reinterpret_cast<double2 *>(vn + G.r0)[i] = w2;
            }
            if (p > 0 && tm.rank == 0 && tid < p) vn[n + tid] = S->wtail[tid] / beta;
        }
        fence_proxy_async();  // the producer's TMA reads of this column must see these generic-proxy stores
        consumer_sync();
        if (tid == 0) S->cols_ready = jc + 2;
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
        P.stat[prob * 4 + 0] = m_out;
        P.stat[prob * 4 + 1] = breakdown;
    }
}

// One instance per (operator kind, augmented or not): the persistent kernel is sensitive to code size (an unused
// extra mat-vec loop cost 3-5 % everywhere), so each instance carries only the paths it can take.
template <int OPK, bool AUG>
__global__ void __launch_bounds__(NT2, 1) krylov_tma_kernel(const __grid_constant__ KrylovParams P,
                                                            const __grid_constant__ CUtensorMap tmA) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemTma *S = reinterpret_cast<SmemTma *>(smem_raw);
    const size_t ws_bytes = P.w_in_smem ? (((size_t)P.slice * 8 + 127) & ~(size_t)127) : 0;
    double *ws_smem = reinterpret_cast<double *>(smem_raw + sizeof(SmemTma));
    unsigned char *ring = smem_raw + sizeof(SmemTma) + ws_bytes;

    const int tid = threadIdx.x;
    const int team = blockIdx.x / P.team_size;
    Team tm;
    tm.rank = blockIdx.x % P.team_size;
    tm.C = P.team_size;
    tm.bar = P.peer_bar[P.myrank] + team;
    tm.target = P.bar_base;
    tm.seq = P.seq_base;
    TmaGeom G;
    G.r0 = min(P.n, tm.rank * P.slice);
    G.nrows = min(P.n, G.r0 + P.slice) - G.r0;
    G.TR = P.tile_rows;
    G.ntk = (G.nrows + G.TR - 1) / G.TR;
    G.nch = OPK == OP_CSR_STREAM ? (G.nrows + P.ch_rows - 1) / P.ch_rows : 0;

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

    const long long xoff0 = (long long)team * 2 * P.xlen;
    double *xb0 = P.peer_xbuf[P.myrank] + xoff0;
    double *xb1 = xb0 + P.xlen;
    const long long part0 = (long long)team * 2 * MAXCOL * P.cpad;
    const long long partn0 = (long long)team * 4 * P.cpad;

    const bool is_producer = tid >= NTC;
    Cons cx;
    cx.S = S;
    cx.ws = P.w_in_smem ? ws_smem : (P.wglob + (long long)team * P.n + G.r0);
    cx.tid = tid;
    cx.lane = tid & 31;
    cx.warp = tid >> 5;
    cx.seq = P.seq_base;

    int nlocal = -1;
    for (int prob = team; prob < P.nprob; prob += P.nteams) {
        ++nlocal;
        if (is_producer) {
            if (tid == NTC) {
                Ring rg{ring, P.nslot, 0, 0u};
                unsigned issued = 0;
                producer_problem<OPK, AUG>(P, &tmA, S, rg, G, P.V + (long long)prob * P.V_stride, nlocal + 1, issued, 0);
            }
            __syncwarp();
        } else {
            cx.rg = Ring{ring, P.nslot, 0, 0u};
            consumer_problem<OPK, AUG>(P, cx, G, tm, prob, nlocal, xb0, xb1, xoff0, part0, partn0);
            consumer_sync();
            if (tid == 0) S->stop_seq = nlocal + 1;
        }
        // CTA-wide resynchronisation: the ring is re-initialised between problems
        __syncthreads();
        if (prob + P.nteams < P.nprob) {
            if (tid == 0) {
                for (int s = 0; s < P.nslot; ++s) {
                    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&S->full[s])) : "memory");
                    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&S->empty[s])) : "memory");
                    mbar_init(&S->full[s], 1);
                    mbar_init(&S->empty[s], NW);
                }
                S->cols_ready = P.j0;
                mbar_fence_init();
            }
            __syncthreads();
        }
    }
}

}  // namespace b200k
