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
//   [update: the same tiles in exactly the reverse order -- batches, row tiles and columns descending (L2 reuse under
//            LRU, see update_phase_c); XL instance: for k in 0..ntk-1 : for col = uhi..ulo]
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
constexpr int LLQ = 4;                           // quantities one packet all-reduce (ll_publish / ll_collect) carries

struct __align__(128) SmemTma {
    uint64_t full[MAXSLOT];
    uint64_t empty[MAXSLOT];
    double hs[MAXCOL];
    double red[2][NW][CB];
    double redn[NW];
    double wtail[MAXP];
    double xtail[MAXP];
    double ptail[MAXP];  // one-reduction Lanczos: tail rows of the previous vector
    int chunk_a0[MAXCH2];
    int chunk_cnt[MAXCH2];
    int slot_a0[MAXSLOT];
    int chunk_local[MAXCH2];  // XL: every column of the chunk lies in this CTA's slice (learnt by the first mat-vec)
    double bc[2];             // reduced scalars (squared norm) broadcast to the CTA
    double llv[LLQ][CPAD];    // packets of a fused barrier + all-reduce (ll_collect), one value per CTA of the team
    // Flags the producer lane polls while the consumers run (single writer, single reader).  flag_set / flag_get are
    // volatile accesses in the product and shared-memory atomics in -DB200K_ATOMIC_FLAGS builds: a polled flag is a
    // data race by definition for plain loads and stores, which compute-sanitizer racecheck reports.
    int j0_found;    // SAFE instance: step at which it takes over (0: nowhere)
    int cols_ready;  // number of complete basis columns of the current problem
    int stop_seq;    // consumers finished local problem #stop_seq (1-based)
};
#ifdef B200K_ATOMIC_FLAGS  // sanitizer builds: shared-memory atomics, so that racecheck sees a race-free protocol
__device__ __forceinline__ void flag_set(int *f, int v) { atomicExch(f, v); }
__device__ __forceinline__ int flag_get(int *f) { return atomicAdd(f, 0); }
#else  // product: volatile accesses (measured: the atomics cost 1.1 % of the C2 Arnoldi kernel)
__device__ __forceinline__ void flag_set(int *f, int v) { *reinterpret_cast<volatile int *>(f) = v; }
__device__ __forceinline__ int flag_get(int *f) { return *reinterpret_cast<volatile int *>(f); }
#endif

// Per-phase timestamps (profiling builds only: -DB200K_PHASE_TIMING, scripts/phase_timing.py).
#ifdef B200K_PHASE_TIMING
constexpr int PT_STEPS = 64, PT_MARKS = 16, PT_CTAS = 160;
__device__ long long g_phase_ts[PT_CTAS * PT_STEPS * PT_MARKS];
#define PT_MARK(rank, step, k)                                                                    \
    do {                                                                                          \
        if (threadIdx.x == 0 && (rank) < PT_CTAS && (step) < PT_STEPS)                            \
            g_phase_ts[((rank) * PT_STEPS + (step)) * PT_MARKS + (k)] = clock64();                \
    } while (0)
#else
#define PT_MARK(rank, step, k) do { } while (0)
#endif

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTC) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Explicit shared-memory accesses by 32-bit shared address.  The two slice buffers of the XL instance swap roles
// every step, so the compiler cannot prove their address space and emits generic LD / ST with 64-bit address
// arithmetic (measured: the mat-vec is instruction-issue bound); these keep them LDS / STS.
__device__ __forceinline__ double lds1(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 lds2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts1(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
__device__ __forceinline__ void sts2(uint32_t a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y));
}

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
        // acquire poll: orders everything after it (the other threads through the CTA barrier below) and invalidates the
        // L1 (LDG.STRONG + CCTL.IVALL), which is what makes the L1-cached gathers of the next mat-vec safe.  No fence
        // behind it: a trailing __threadfence cost 1.5 % of the C2 Arnoldi kernel (profiles/r2_s31_ab_barrier.log;
        // release on the atomic itself -- red.release instead of fence + atomicAdd -- measured 2 % slower).
        while ((int)(ld_acquire_u32(tm.bar) - tm.target) < 0) {
        }
    }
    consumer_sync();
}

// ---- single-GPU fused barrier + all-reduce of <= LLQ scalars (short orthogonalisation windows) --------------------
// The counter barrier costs three dependent L2 round trips and two MEMBARs per reduction (partial store + fence +
// atomic, poll, fence, partial loads: ~2.4 us measured, one third of a Lanczos step).  Here every CTA sends its
// partial as a self-validating 16-byte packet {lo32, seq, hi32, seq} to the inbox of EVERY CTA of its team (C plain
// stores, no hot spot) and polls only its own inbox, adding the C packets in CTA order (bitwise identical
// everywhere): data and notification travel together, one L2 round trip.  (A single shared packet table polled by
// all 148 CTAs was measured first: 148 x 148 spinning loads on 19 cache lines made it slower than the counter.)
// Two packet sets (parity of seq) suffice: a CTA can only be one reduction ahead of the slowest CTA of its team.
// `release` / `acquire` = the reduction also orders this step's gather-buffer stores before the next mat-vec's
// gathers: MEMBAR before the packet stores, ld.acquire polls (LDG.STRONG + CCTL.IVALL, no MEMBAR) on the other side.
__device__ __forceinline__ uint4 *ll_slot(const KrylovParams &P, unsigned seq, int ci, int dest_cta, int src_rank) {
    return P.llpkt + ((((long long)(seq & 1u)) * LLQ + ci) * CPAD + dest_cta) * CPAD + src_rank;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ double ll_poll_acquire(const uint4 *p, unsigned seq) {
    unsigned lo, f1, hi, f2;
    do {
        asm volatile("ld.acquire.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2)
                     : "l"(p)
                     : "memory");
    } while (f1 != seq || f2 != seq);
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
// Called by ALL lanes of one warp with the same value v (CTA partial of quantity ci).
__device__ __forceinline__ void ll_publish_warp(const KrylovParams &P, int team, const Team &tm, unsigned seq, int ci,
                                                double v, int lane, bool release) {
    if (release) fence_acq_rel_gpu();
    const int base = team * tm.C;
    for (int d = lane; d < tm.C; d += 32) ll_push(ll_slot(P, seq, ci, base + d, tm.rank), v, seq);
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
    // (the stop flag is only looked at every 8th failed wait: it matters once per problem, the wait latency matters
    // every slot, and the flag read is an atomic)
    unsigned spins = 0;
    while (!mbar_try_wait(&S->empty[rg.slot], rg.phase ^ 1u)) {
        if ((++spins & 7u) == 0u && flag_get(&S->stop_seq) >= seq) return false;
    }
    return true;
}

__device__ __forceinline__ bool prod_wait_col(SmemTma *S, int col, int seq, int) {
    while (flag_get(&S->cols_ready) <= col) {
        if (flag_get(&S->stop_seq) >= seq) return false;
    }
    return true;
}

template <int OPK, bool AUG, bool XL>
__device__ void producer_problem(const KrylovParams &P, const CUtensorMap *tmA, SmemTma *S, Ring &rg,
                                 const TmaGeom &G, const double *V, int seq, unsigned &issued, int lane, int j0) {
    const long long ldv = P.ldv;
    const int jstart = j0 == 0 ? 1 : j0;
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
#ifdef B200K_PHASE_TIMING
            long long pt_acq = 0;
            const long long pt_begin = clock64();
#endif
            for (int c = 0; c < G.nch; ++c) {
#ifdef B200K_PHASE_TIMING
                const long long pa0 = clock64();
                const bool got = prod_acquire(S, rg, seq, lane);
                pt_acq += clock64() - pa0;
                if (!got) { stopped = true; break; }
#else
                if (!prod_acquire(S, rg, seq, lane)) { stopped = true; break; }
#endif
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
#ifdef B200K_PHASE_TIMING
            if (blockIdx.x < PT_CTAS && j < PT_STEPS) {
                long long *q = g_phase_ts + ((long long)blockIdx.x * PT_STEPS + j) * PT_MARKS + 13;
                q[0] = pt_acq;
                q[1] = clock64() - pt_begin;
            }
#endif
            if (stopped) break;
        } else if (OPK == OP_DENSE && P.dense_cpt > 0 && G.nrows > 0) {
            // dense operator: a tile = dense_cpt columns x the CTA's row slice, fetched as slice/box_rows
            // tensor-map boxes (rows past n are zero-filled by the TMA unit and still count as bytes)
            const int nrb = P.slice / P.dense_box_rows;
            const uint32_t tbytes = (uint32_t)P.slice * (uint32_t)P.dense_cpt * 8u;
            for (int c0 = 0; c0 < P.ncols; c0 += P.dense_cpt) {
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
        // XL: column jc is the shared-memory resident vector -- no tiles for it in either pass
        const int hi = XL ? jc - 1 : jc;
        // (XL Lanczos: v_{j-1} is folded into the mat-vec out of shared memory -- the ring carries the operator only)
        const int ulo = (XL && P.lanczos) ? jc : ((P.lanczos && jc >= 1) ? jc - 1 : lo);
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
        // update tiles.  General instance: the exact reverse of the dots order (column batches, row tiles and columns
        // descending; see update_phase_c).  XL: one pass over the row tiles (short windows).
        const int nbatch = XL ? 1 : (hi - lo) / CB + 1;
        for (int bi = nbatch - 1; bi >= 0 && !stopped; --bi) {
            const int c0 = (XL || bi == 0) ? ulo : lo + bi * CB;
            const int c1 = XL ? hi : min(lo + bi * CB + CB - 1, hi);
            for (int kk = 0; kk < G.ntk && !stopped; ++kk) {
                const int k = XL ? kk : G.ntk - 1 - kk;
                const int rows = min(G.TR, G.nrows - k * G.TR);
                for (int col = c1; col >= c0; --col) {
                    // (XL: column jc - 1 is written lazily during the previous step's update phase)
                    if ((XL && !prod_wait_col(S, col, seq, lane)) || !prod_acquire(S, rg, seq, lane)) { stopped = true; break; }
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&S->full[rg.slot], (uint32_t)rows * 8u);
                        const double *src = V + (long long)col * ldv + G.r0 + (long long)k * G.TR;
                        if (hintV) bulk_g2s_hint(rg.ptr(), src, (uint32_t)rows * 8u, &S->full[rg.slot], polV);
                        else bulk_g2s(rg.ptr(), src, (uint32_t)rows * 8u, &S->full[rg.slot]);
                    }
                    rg.advance();
                    ++issued;
                }
            }
        }
    }
    // the consumers decide when the problem is over (m steps, happy breakdown, or beta == 0)
    while (flag_get(&S->stop_seq) < seq) __nanosleep(256);  // (do not steal issue slots from the consumers while waiting)
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
    double *xin;  // XL: this CTA's slice of the (unnormalised) current basis vector, resident in shared memory
    uint32_t xin_a, ws_a;  // XL: shared addresses of xin / ws (lds1 / sts1 ...)
    int team;
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

// Collect side of the packet all-reduce: out[ci] (shared memory) = sum over the team of quantity ci.
__device__ void ll_collect(const KrylovParams &P, Cons &cx, const Team &tm, int ncols, double *out, bool acquire) {
    SmemTma *S = cx.S;
    cx.seq += 1u;
    const unsigned seq = cx.seq;
    const int total = ncols * tm.C;
    for (int idx = cx.tid; idx < total; idx += NTC) {
        const int ci = idx / tm.C, r = idx - ci * tm.C;
        const uint4 *slot = ll_slot(P, seq, ci, (int)blockIdx.x, r);
        S->llv[ci][r] = acquire ? ll_poll_acquire(slot, seq) : ll_poll(slot, seq);
    }
    consumer_sync();
    for (int ci = cx.warp; ci < ncols; ci += NW) {
        double s = 0.0;
        for (int q = cx.lane; q < tm.C; q += 32) s += S->llv[ci][q];
        s = warp_sum(s);
        if (cx.lane == 0) out[ci] = s;
    }
    consumer_sync();
}

// ---- row-sharded short windows: two-level packet all-reduce (no counter barrier, no trailing fence) ------------------
// Level 1 (inside the GPU): every CTA sends its partial of quantity ci to the inbox of ONE owner CTA (CTA ci); the
// owner adds the C packets in CTA order.  Level 2 (NVLink): the owner pushes the GPU's sum to every GPU's inbox, every
// CTA polls the R packets of each quantity of its own GPU's inbox and adds them in rank order.  `ordered` = the
// reduction also publishes this step's gather-buffer stores and halo pushes: a CTA fences (system scope if it pushed
// rows to a peer) before its level-1 packet, the owner acquires, fences at system scope before the level-2 pushes,
// and the final polls are ld.acquire.sys (which also invalidates L1 for the halo rows peers wrote into this GPU).
// Replaces per reduction: counter barrier (atomic + spin + 2 fences), partial-sum loads, trailing __threadfence_system.
__device__ __forceinline__ double ll_poll_acquire_sys(const uint4 *p, unsigned seq) {
    unsigned lo, f1, hi, f2;
    do {
        asm volatile("ld.acquire.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2)
                     : "l"(p)
                     : "memory");
    } while (f1 != seq || f2 != seq);
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
constexpr int LLLOCQ = MAXCOL + 1;  // quantities the per-GPU level-1 inbox holds per parity (full Arnoldi windows)
__device__ __forceinline__ uint4 *llloc_slot(const KrylovParams &P, unsigned seq, int ci, int src) {
    return P.llloc + (((long long)(seq & 1u)) * LLLOCQ + ci) * CPAD + src;
}
// Called by ALL lanes of warp 0 with the CTA partial v of quantity ci.
__device__ __forceinline__ void shard_publish_warp(const KrylovParams &P, const Team &tm, unsigned seq, int ci, double v,
                                                   int lane, bool ordered, bool pushed) {
    if (ordered) {
        if (pushed) asm volatile("fence.acq_rel.sys;" ::: "memory");
        else fence_acq_rel_gpu();
    }
    if (lane == 0) ll_push(llloc_slot(P, seq, ci, tm.rank), v, seq);
}
__device__ void shard_collect(const KrylovParams &P, Cons &cx, const Team &tm, int ncols, double *out, bool ordered) {
    SmemTma *S = cx.S;
    cx.seq += 1u;
    const unsigned seq = cx.seq;
    const long long pbase = (long long)(seq & 1u) * (MAXCOL + 1) * 8;
    for (int ci = tm.rank; ci < ncols; ci += tm.C) {  // quantities this CTA owns (uniform for the CTA)
        for (int r = cx.tid; r < tm.C; r += NTC) {
            const uint4 *slot = llloc_slot(P, seq, ci, r);
            S->llv[0][r] = ordered ? ll_poll_acquire(slot, seq) : ll_poll(slot, seq);
        }
        consumer_sync();
        if (cx.warp == 0) {
            double s = 0.0;
            for (int q = cx.lane; q < tm.C; q += 32) s += S->llv[0][q];
            s = warp_sum(s);
            if (ordered) asm volatile("fence.acq_rel.sys;" ::: "memory");
            if (cx.lane < P.nranks) ll_push(P.peer_pkt[cx.lane] + pbase + (long long)ci * 8 + P.myrank, s, seq);
        }
        if (ci + tm.C < ncols) consumer_sync();  // the staging row is reused by the next owned quantity
    }
    for (int ci = cx.warp; ci < ncols; ci += NW) {
        double v = 0.0;
        if (cx.lane < P.nranks) {
            const uint4 *slot = P.peer_pkt[P.myrank] + pbase + (long long)ci * 8 + cx.lane;
            v = ordered ? ll_poll_acquire_sys(slot, seq) : ll_poll(slot, seq);
        }
        double s = 0.0;
        for (int r = 0; r < P.nranks; ++r) s += __shfl_sync(0xffffffffu, v, r);
        if (cx.lane == 0) out[ci] = s;
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

// Row-sharded norm-type reduction of the general instance: block sum of `v`, halo push of the finished w slice, then a
// two-level packet all-reduce that also publishes the gather-buffer stores and the pushed rows (replaces counter
// barrier + __threadfence_system; the same scheme the short-window instance uses).  Result in out[0].
__device__ void shard_norm_reduce(const KrylovParams &P, Cons &cx, const TmaGeom &G, const Team &tm, double v,
                                  long long xoff, double *out) {
    v = warp_sum(v);
    if (cx.lane == 0) cx.S->redn[cx.warp] = v;
    consumer_sync();  // (also: the whole w slice of this CTA is in place)
    const bool pushed = P.send_ofs[tm.rank + 1] > P.send_ofs[tm.rank];
    push_halo(P, cx, G, tm, xoff);
    consumer_sync();
    if (cx.warp == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += cx.S->redn[w];
        shard_publish_warp(P, tm, cx.seq + 1u, 0, s, cx.lane, true, pushed);
    }
    shard_collect(P, cx, tm, 1, out, true);
}

// XL mat-vec (CSR stream): entries whose column lies in this CTA's own row slice are gathered from the shared-memory
// resident slice cx.xin instead of L2 (for stencil-like operators that is almost every entry), the result goes to the
// other buffer cx.ws, and the inner product of the new w with the current basis vector is accumulated on the fly --
// returned UNSCALED per thread: sum over this thread's rows of xin[row] * (A x)[row].
// The phase is instruction-issue bound (16 warps, one row per thread: measured 820 cycles per 512-row chunk without
// any x load), so (a) the gather batch width GW is a template parameter (5 for operators with <= 5 entries per row),
// and (b) chunks whose columns ALL lie in the slice take a loop without the local/remote select or any global
// address arithmetic; which chunks those are is learnt by the first mat-vec of a launch (`learn`).
// (Negative result, measured in one A/B call: reading a SELL-16 copy of the operator straight from L2 into registers
// with coalesced read-only loads, two rows per thread in flight and no shared-memory staging at all, took 30.6 k
// cycles per C2 mat-vec against 17.0 k for this ring version -- like the v1 LDG kernel, direct loads do not keep
// enough bytes in flight per SM.  The code was removed again.)
template <bool AUG, int GW>
__device__ double matvec_xl(const KrylovParams &P, Cons &cx, const TmaGeom &G, const double *xsrc, double xscale,
                            bool learn, bool fold, double foldc, int pt_step = 0) {
    double selfacc = 0.0;
    SmemTma *S = cx.S;
    const int tid = cx.tid;
    const int n = P.n, p = AUG ? P.p : 0;
    const uint32_t ws_a = cx.ws_a, xin_a = cx.xin_a;
    const uint32_t xl_a = xin_a - 8u * (uint32_t)G.r0;  // shared address of x(column) for own-slice columns
    if (p > 0) {
        if (tid < p) S->xtail[tid] = xsrc[n + P.nhalo + tid];
        consumer_sync();
        if (tid < p) S->wtail[tid] = (tid < p - 1) ? S->xtail[tid + 1] * xscale : 0.0;
    }
    const int nnz_cap = P.nnz_cap;
#ifdef B200K_PHASE_TIMING
    long long pt_wait = 0, pt_comp = 0;
#endif
    for (int c = 0; c < G.nch; ++c) {
        const int rl = c * P.ch_rows + tid;
        const bool active = tid < P.ch_rows && rl < G.nrows;
        const bool fast = !learn && c < MAXCH2 && S->chunk_local[c] != 0;
#ifdef B200K_PHASE_TIMING
        const long long pt0 = clock64();
#endif
        cx.wait_full();
#ifdef B200K_PHASE_TIMING
        const long long pt1 = clock64();
#endif
        bool loc = true;
#ifdef B200K_EXP_NOCOMPUTE  // profiling experiment: pure operator-stream rate of the mat-vec phase
        if (false) {
#else
        if (active) {
#endif
            const unsigned char *base = cx.rg.ptr();
            const double *vs = reinterpret_cast<const double *>(base);
            const int *cs = reinterpret_cast<const int *>(base + (size_t)nnz_cap * 8);
            const int *rp = reinterpret_cast<const int *>(base + (size_t)nnz_cap * 12);
            const int a0 = S->slot_a0[cx.rg.slot];
            const int e0 = rp[tid] - a0, e1 = rp[tid + 1] - a0;
            double sum = 0.0;
            if (fast) {
#pragma unroll 1
                for (int eb = e0; eb < e1; eb += GW) {
                    double av[GW], xv[GW];
#pragma unroll
                    for (int u = 0; u < GW; ++u) {
                        const bool ok = eb + u < e1;
                        av[u] = ok ? vs[eb + u] : 0.0;
                        xv[u] = 0.0;
                        if (ok) xv[u] = lds1(xl_a + 8u * (uint32_t)cs[eb + u]);
                    }
#pragma unroll
                    for (int u = 0; u < GW; ++u) sum = fma(av[u], xv[u], sum);
                }
            } else {
#pragma unroll 1
                for (int eb = e0; eb < e1; eb += GW) {
                    double av[GW], xv[GW];
#pragma unroll
                    for (int u = 0; u < GW; ++u) {
                        const bool ok = eb + u < e1;
                        av[u] = ok ? vs[eb + u] : 0.0;
                        xv[u] = 0.0;
#ifdef B200K_EXP_NOGATHER  // profiling experiment: no x loads at all
                        if (ok) xv[u] = 1.0 + (double)cs[eb + u];
                        else
#endif
                        if (ok) {
                            const int col = cs[eb + u];
                            const unsigned lc = (unsigned)(col - G.r0);
                            const bool here = lc < (unsigned)G.nrows;
                            loc = loc && here;
                            if (here) xv[u] = lds1(xin_a + 8u * lc);
                            else xv[u] = xsrc[col];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < GW; ++u) sum = fma(av[u], xv[u], sum);
                }
            }
            if (p > 0) {
                const double *brow = P.Bm + (G.r0 + rl);
                for (int k = 0; k < p; ++k) sum = fma(brow[(long long)k * P.ldb], S->xtail[k], sum);
            }
            // Lanczos: the output buffer still holds the unnormalised v_{j-1} (it was the resident vector of the
            // previous step) and beta_{j-1} is known, so its term of the three-term recurrence is subtracted here,
            // by the thread that overwrites the entry -- no basis tile, no second pass.  The inner product with
            // v_j is taken before that, as lanczos_step! does (arnoldi.jl:396-399).
            double wv = sum * xscale;
            selfacc = fma(lds1(xin_a + 8u * (uint32_t)rl), wv, selfacc);
            if (fold) wv = fma(-foldc, lds1(ws_a + 8u * (uint32_t)rl), wv);
            sts1(ws_a + 8u * (uint32_t)rl, wv);
        }
        if (learn && c < MAXCH2) {
            if (!__all_sync(0xffffffffu, loc) && cx.lane == 0) S->chunk_local[c] = 0;
        }
        cx.release();
#ifdef B200K_PHASE_TIMING
        pt_wait += pt1 - pt0;
        pt_comp += clock64() - pt1;
#endif
    }
#ifdef B200K_PHASE_TIMING
    if ((tid == 0 || tid == 480) && blockIdx.x < PT_CTAS && pt_step < PT_STEPS) {
        long long *q = g_phase_ts + ((long long)blockIdx.x * PT_STEPS + pt_step) * PT_MARKS + (tid == 0 ? 9 : 11);
        q[0] = pt_wait;
        q[1] = pt_comp;
    }
#endif
    return selfacc;  // (no trailing barrier: the block reduction of the fused inner product is the barrier)
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
                for (int c0 = 0; c0 < P.ncols; c0 += P.dense_cpt) {
                    const int nc = min(P.dense_cpt, P.ncols - c0);
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
                    for (int c = g; c < P.ncols; c += Gc) {
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
            if (P.nranks > 1) ll_push(llloc_slot(P, cx.seq + 1u, cb - lo + tid, tm.rank), s, cx.seq + 1u);  // level-1 packet
            else P.peer_part[P.myrank][part_off + (long long)(cb - lo + tid) * P.cpad + tm.rank] = s;
        }
    }
}

// w -= sum_c h_c v_c (c = uhi..ulo), partial ||w||^2, unnormalised w to the gather buffer.
// Traversal = the exact reverse of the dots phase at tile granularity (column batches descending, row tiles descending,
// columns descending): the update starts with what the dots phase touched last, so under LRU whatever is among the last
// ~100 MB of the dots traffic is an L2 hit.  A row-tile-outer loop evicts the later row tiles of the recent columns
// while it misses on the old columns of the first row tile.  Per element the columns are still subtracted in the order
// uhi..ulo (bitwise identical results); the price is one shared-memory load + store of the w tile per (batch, tile).
template <int OPK, bool AUG>
__device__ double update_phase_c(const KrylovParams &P, Cons &cx, const TmaGeom &G, const Team &tm, const double *V,
                                 int lo, int ulo, int uhi, double *xout) {
    SmemTma *S = cx.S;
    const int tid = cx.tid;
    double2 *ws2 = reinterpret_cast<double2 *>(cx.ws);
    double2 *xo2 = reinterpret_cast<double2 *>(xout + G.r0);
    const double *hs = S->hs;
    double nrm = 0.0;
    const int nbatch = (uhi - lo) / CB + 1;
    for (int bi = nbatch - 1; bi >= 0; --bi) {
        const int c0 = bi == 0 ? ulo : lo + bi * CB;
        const int c1 = min(lo + bi * CB + CB - 1, uhi);
        for (int k = G.ntk - 1; k >= 0; --k) {
            const int pairs = min(G.TR, G.nrows - k * G.TR) >> 1;
            const int pbase = (k * G.TR) >> 1;
            double2 wr[PPT];
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                const int idx = tid + q * NTC;
                wr[q] = idx < pairs ? ws2[pbase + idx] : make_double2(0.0, 0.0);
            }
            for (int col = c1; col >= c0; --col) {
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
            if (bi == 0) {  // last batch: the tile is final
#pragma unroll
                for (int q = 0; q < PPT; ++q) {
                    const int idx = tid + q * NTC;
                    if (idx < pairs) {
                        ws2[pbase + idx] = wr[q];
                        xo2[pbase + idx] = wr[q];
                        nrm = fma(wr[q].x, wr[q].x, fma(wr[q].y, wr[q].y, nrm));
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < PPT; ++q) {
                    const int idx = tid + q * NTC;
                    if (idx < pairs) ws2[pbase + idx] = wr[q];
                }
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


// ---- DGKS re-orthogonalisation ("twice is enough") -------------------------------------------------------------------
// Classical Gram-Schmidt loses orthogonality like eps * (||A v_j|| / ||w||)^2 where the reference's modified
// Gram-Schmidt (arnoldi.jl:301-304) loses eps * ||A v_j|| / ||w||.  Whenever the update removed most of the vector,
// ||w_after|| < REORTH_ETA * ||w_before||, the step runs a second classical pass on the updated w and adds its
// coefficients to H -- after which the basis is orthogonal to rounding (Daniel, Gragg, Kaufman, Stewart 1976; Giraud,
// Langou, Rozloznik 2005).  Well-conditioned Krylov sequences (every BASELINE config) never trigger it, so the hot
// path pays one extra block reduction per step for ||w_before||^2 and nothing else; the second pass itself is rare, so
// it reads the basis slice with direct loads instead of going through the producer's tile schedule, and lives in
// out-of-line functions to keep the hot loop's code size and register allocation unchanged.
constexpr double REORTH_ETA2 = 0.0625;  // eta = 1/4 (squared norms are compared)

// (All of these take plain values, not the Cons / TmaGeom / Team structs of the caller: an address that escapes into an
// out-of-line call would pin those structs in local memory for the whole hot loop.)
// Second-pass inner products <v_c, w> for c = lo..hi with direct loads of the CTA's basis slice.
__device__ __noinline__ void reorth_dots_c(const KrylovParams &P, SmemTma *S, const double *ws, int r0, int nrows, int rank,
                                           const double *V, int lo, int hi, long long part_off, bool aug, unsigned seq_next) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double2 *ws2 = reinterpret_cast<const double2 *>(ws);
    const int units = nrows >> 1;
    int batch = 0;
    for (int cb = lo; cb <= hi; cb += CB, ++batch) {
        const int nb = min(CB, hi - cb + 1);
        double acc[CB];
#pragma unroll
        for (int u = 0; u < CB; ++u) acc[u] = 0.0;
        const double *vb = V + (long long)cb * P.ldv + r0;
        for (int i = tid; i < units; i += NTC) {
            const double2 w2 = ws2[i];
#pragma unroll
            for (int u = 0; u < CB; ++u)
                if (u < nb) {
                    const double2 v2 = *reinterpret_cast<const double2 *>(vb + (long long)u * P.ldv + 2 * i);
                    acc[u] = fma(v2.x, w2.x, fma(v2.y, w2.y, acc[u]));
                }
        }
        if (aug && P.p > 0 && rank == 0 && P.myrank == 0 && tid == 0) {
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
            if (P.nranks > 1) ll_push(llloc_slot(P, seq_next, cb - lo + tid, rank), s, seq_next);
            else P.peer_part[P.myrank][part_off + (long long)(cb - lo + tid) * P.cpad + rank] = s;
        }
    }
}

// Second-pass update w -= sum_c h2[c - lo] v_c (direct loads); rewrites the w slice and the gather-buffer copy and
// returns this thread's partial ||w||^2.
__device__ __noinline__ double reorth_update_c(const KrylovParams &P, SmemTma *S, double *ws, int r0, int nrows, int rank,
                                               const double *V, int lo, int hi, const double *h2, double *xout, bool aug) {
    const int tid = threadIdx.x;
    double2 *ws2 = reinterpret_cast<double2 *>(ws);
    double2 *xo2 = reinterpret_cast<double2 *>(xout + r0);
    const int units = nrows >> 1;
    double nrm = 0.0;
    for (int i = tid; i < units; i += NTC) {
        double2 w2 = ws2[i];
        const double *vrow = V + r0 + 2 * i;
        for (int c = hi; c >= lo; --c) {
            const double hc = h2[c - lo];
            const double2 v2 = *reinterpret_cast<const double2 *>(vrow + (long long)c * P.ldv);
            w2.x = fma(-hc, v2.x, w2.x);
            w2.y = fma(-hc, v2.y, w2.y);
        }
        ws2[i] = w2;
        xo2[i] = w2;
        nrm = fma(w2.x, w2.x, fma(w2.y, w2.y, nrm));
    }
    if (aug && P.p > 0 && tid < P.p) {
        double wt = S->wtail[tid];
        for (int c = hi; c >= lo; --c) wt = fma(-h2[c - lo], V[(long long)c * P.ldv + P.n + tid], wt);
        S->wtail[tid] = wt;
        if (rank == 0) {
            xout[P.n + P.nhalo + tid] = wt;
            if (P.myrank == 0) nrm = fma(wt, wt, nrm);
        }
    }
    return nrm;
}

// The whole second pass: inner products, reduction, H += h2, update, norm reduction (result in S->bc[0]).  Passes two
// team reductions: the caller advances its barrier target by 2 C and its sequence number by 2 afterwards.  Few scalar
// arguments (everything else is re-derived from P and blockIdx exactly as the kernel does): they travel in registers.
__device__ __noinline__ void reorth_step_c(const KrylovParams &P, unsigned target, unsigned seq, int prob, int jc, bool aug) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemTma *S = reinterpret_cast<SmemTma *>(smem_raw);
    const int team = blockIdx.x / P.team_size;
    Cons cx;
    cx.S = S;
    cx.xin = nullptr;
    cx.ws_a = cx.xin_a = 0;
    cx.team = team;
    cx.tid = threadIdx.x;
    cx.lane = threadIdx.x & 31;
    cx.warp = threadIdx.x >> 5;
    cx.seq = seq;
    Team tm;
    tm.rank = blockIdx.x % P.team_size;
    tm.C = P.team_size;
    tm.bar = P.peer_bar[P.myrank] + team;
    tm.target = target;
    tm.seq = seq;
    TmaGeom G;
    G.r0 = min(P.n, tm.rank * P.slice);
    G.nrows = min(P.n, G.r0 + P.slice) - G.r0;
    G.nch = G.ntk = G.TR = 0;
    cx.ws = P.w_in_smem ? reinterpret_cast<double *>(smem_raw + sizeof(SmemTma)) : (P.wglob + (long long)team * P.n + G.r0);
    const int j = jc + 1, par = j & 1;
    const int iopw = P.iop > 0 ? P.iop : P.m;
    const int lo = max(0, jc - iopw + 1), hi = jc, nc = hi - lo + 1;
    const long long xoff = (long long)team * 2 * P.xlen + (par ? P.xlen : 0);
    double *xout = P.peer_xbuf[P.myrank] + xoff;
    const long long part = (long long)team * 2 * MAXCOL * P.cpad + (long long)par * MAXCOL * P.cpad;
    const long long partn = (long long)team * 4 * P.cpad + (long long)par * P.cpad;
    const double *V = P.V + (long long)prob * P.V_stride;
    double *Hcol = P.Hd + (long long)prob * P.H_stride + (long long)jc * P.ldh + lo;
    const bool sharded = P.nranks > 1;
    double *h2 = &S->llv[1][0];  // (rows 1..3 of the packet staging area: 480 doubles; row 0 is shard_collect's)
    consumer_sync();
    reorth_dots_c(P, S, cx.ws, G.r0, G.nrows, tm.rank, V, lo, hi, part, aug, cx.seq + 1u);
    if (sharded) shard_collect(P, cx, tm, nc, h2, false);
    else team_reduce_c(P, cx, tm, P.peer_part[P.myrank] + part, nc, h2, false);
    if (tm.rank == 0)
        for (int ci = cx.tid; ci < nc; ci += NTC) Hcol[ci] = S->hs[ci] + h2[ci];
    const double nrm2 = reorth_update_c(P, S, cx.ws, G.r0, G.nrows, tm.rank, V, lo, hi, h2, xout, aug);
    if (sharded) {
        shard_norm_reduce(P, cx, G, tm, nrm2, xoff, S->bc);
    } else {
        block_sum_to_c(P, cx, nrm2, partn + tm.rank);
        team_reduce_c(P, cx, tm, P.peer_partn[P.myrank] + partn, 1, S->bc, false);
    }
}

// One problem on the consumer side.  Mirrors krylov_body<2> of krylov_kernel.cuh.
// SAFE = false: the fast instance -- no re-orthogonalisation code at all (any such code inside the step loop, even a
// never-taken branch, cost 4-5 % of the C2 Arnoldi kernel through its effect on code generation,
// profiles/r2_dgks_ab.md).  The test needs nothing the fast instance does not already leave behind: ||w_after|| is
// H[j+1, j] and ||h||^2 the sum of squares of the column above it, so it is evaluated AFTERWARDS from the stored
// Hessenberg matrix by the SAFE instance (launched right behind the fast one): if no step fails the test it exits at
// once; otherwise it resumes the factorisation at the first failing step like arnoldi!(...; init = j), with the second
// Gram-Schmidt pass wherever the test fires, and overwrites what the fast instance computed from there on.
template <int OPK, bool AUG, bool SAFE>
__device__ void consumer_problem(const KrylovParams &P, Cons &cx, const TmaGeom &G, Team &tm, int prob, int nlocal,
                                 double *xb0, double *xb1, long long xoff0, long long part0, long long partn0, int j0) {
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
    int m_out = P.m, breakdown = 0, nreorth = 0;
    const bool sharded = P.nranks > 1;
    const bool via_xb0 = p > 0 || sharded;            // first gather source must carry tail / halo entries
    const double *lpart = P.peer_part[P.myrank];      // this GPU's inboxes
    const double *lpartn = P.peer_partn[P.myrank];
    const int xt = n + P.nhalo;                       // offset of the augmented tail in the gather buffers

    if (j0 == 0) {  // firststep! (arnoldi.jl:230-250 / 257-279)
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
        if (sharded) {
            shard_norm_reduce(P, cx, G, tm, nrm, xoff0, S->bc);
        } else {
            block_sum_to_c(P, cx, nrm, pslot + tm.rank);
            team_reduce_c(P, cx, tm, lpartn + pslot, 1, S->bc, false);
        }
        const double beta = sqrt(S->bc[0]);
        if (tm.rank == 0 && tid == 0) P.scal[prob * 4] = beta;
        if (beta == 0.0) {
            if (tm.rank == 0 && tid == 0) {
                P.stat[prob * 4 + 0] = P.m;
                P.stat[prob * 4 + 1] = 0;
                P.stat[prob * 4 + 2] = 0;
                P.stat[prob * 4 + 3] = 0;
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
        if (tid == 0) flag_set(&S->cols_ready, 1);
        xscale = 1.0 / beta;
        jstart = 1;
    } else {
        const double *vj = V + (long long)(j0 - 1) * ldv;
        if (sharded) {  // the resumed column has no halo: stage it in the gather buffer and push the halo
            for (int i = tid; i < units; i += NTC) {
                const double2 v2 = reinterpret_cast<const double2 *>(vj + G.r0)[i];
                ws2[i] = v2;
                reinterpret_cast<double2 *>(xb0 + G.r0)[i] = v2;
            }
            if (p > 0 && tm.rank == 0 && tid < p) xb0[xt + tid] = vj[n + tid];
            shard_norm_reduce(P, cx, G, tm, 0.0, xoff0, S->bc);  // every GPU's halo is in place afterwards
            xsrc = xb0;
        } else {
            xsrc = vj;
        }
        xscale = 1.0;
        jstart = j0;
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

        PT_MARK(blockIdx.x, j, 0);
        matvec_phase_c<OPK, AUG>(P, cx, G, xsrc, xscale);
        PT_MARK(blockIdx.x, j, 1);

        const int lo = P.lanczos ? jc : max(0, jc - iopw + 1);
        const int hi = jc;
        const bool dgks = SAFE && !P.lanczos;
        const int nc = hi - lo + 1;
        const int ulo = (P.lanczos && jc >= 1) ? jc - 1 : lo;
        dots_phase_c<OPK, AUG>(P, cx, G, tm, V, lo, hi, part);
        PT_MARK(blockIdx.x, j, 2);
        if (sharded) shard_collect(P, cx, tm, nc, S->hs + (lo - ulo), false);
        else team_reduce_c(P, cx, tm, lpart + part, nc, S->hs + (lo - ulo), false);
        PT_MARK(blockIdx.x, j, 3);
        if (tm.rank == 0)
            for (int ci = tid; ci < nc; ci += NTC) Hd[(long long)jc * ldh + lo + ci] = S->hs[lo + ci - ulo];
        if (P.lanczos && jc >= 1 && tid == 0) S->hs[0] = beta_prev;
        // Re-orthogonalisation test: ||w_before||^2 = ||h||^2 + ||w_after||^2 for an orthonormal window (Pythagoras), so
        // no extra reduction is needed.  One warp sums the coefficients now and parks the result in shared memory;
        // nothing is carried in registers through the update phase.  (SAFE instance only.)
        if (dgks && cx.warp == 1) {
            double q = 0.0;
            for (int ci = cx.lane; ci < nc; ci += 32) q = fma(S->hs[ci], S->hs[ci], q);
            q = warp_sum(q);
            if (cx.lane == 0) S->bc[1] = q;
        }
        consumer_sync();

        const double nrm = update_phase_c<OPK, AUG>(P, cx, G, tm, V, lo, ulo, hi, xout);
        PT_MARK(blockIdx.x, j, 4);
        if (sharded) {
            shard_norm_reduce(P, cx, G, tm, nrm, xoff, S->bc);
        } else {
            block_sum_to_c(P, cx, nrm, partn + tm.rank);
            team_reduce_c(P, cx, tm, lpartn + partn, 1, S->bc, false);
        }
        PT_MARK(blockIdx.x, j, 5);

        if (dgks && S->bc[0] < REORTH_ETA2 * (S->bc[1] + S->bc[0])) {
            // second classical Gram-Schmidt pass, out of line; two more team reductions (every CTA of every rank takes
            // the same decision: the reduced values are bitwise identical everywhere; NaN never triggers it)
            reorth_step_c(P, tm.target, cx.seq, prob, jc, AUG);
            tm.target += 2u * (unsigned)tm.C;
            cx.seq += 2u;
            ++nreorth;
        }
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
        if (tid == 0) flag_set(&S->cols_ready, jc + 2);
        PT_MARK(blockIdx.x, j, 6);
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
        P.stat[prob * 4 + 2] = nreorth;  // steps that took the second Gram-Schmidt pass (host: barrier accounting)
    }
}

template <int OPK, bool AUG, bool XL>
__device__ void dots_phase_xl(const KrylovParams &P, Cons &cx, const TmaGeom &G, const Team &tm, const double *V,
                             int lo, int hi, long long part_off, bool use_ll) {
    SmemTma *S = cx.S;
    const int tid = cx.tid, lane = cx.lane, warp = cx.warp;
    const uint32_t ws_a = cx.ws_a;
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
                wr[q] = idx < pairs ? lds2(ws_a + 16u * (uint32_t)(pbase + idx)) : make_double2(0.0, 0.0);
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
        if (XL && use_ll) {
            if (warp == 0) {
                double s = 0.0;
                if (lane < nb) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) s += S->red[buf][w][lane];
                }
                for (int u = 0; u < nb; ++u) {
                    const double su = __shfl_sync(0xffffffffu, s, u);
                    if (P.nranks > 1) shard_publish_warp(P, tm, cx.seq + 1u, cb - lo + u, su, lane, false, false);
                    else ll_publish_warp(P, cx.team, tm, cx.seq + 1u, cb - lo + u, su, lane, false);
                }
            }
        } else if (tid < nb) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += S->red[buf][w][tid];
            P.peer_part[P.myrank][part_off + (long long)(cb - lo + tid) * P.cpad + tm.rank] = s;
        }
    }
}

// XL: column uhi + 1 (the current basis vector) is cx.xin * xscale in shared memory (`vlazy`: where its augmented
// tail rows are stored, nullptr = already in V; the slice itself is stored by the caller, see consumer_problem).
template <int OPK, bool AUG, bool XL>
__device__ double update_phase_xl(const KrylovParams &P, Cons &cx, const TmaGeom &G, const Team &tm, const double *V,
                                 int ulo, int uhi, double *xout, double xscale, double *vlazy) {
    SmemTma *S = cx.S;
    const int tid = cx.tid;
    const uint32_t ws_a = cx.ws_a, xin_a = cx.xin_a;
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
            wr[q] = idx < pairs ? lds2(ws_a + 16u * (uint32_t)(pbase + idx)) : make_double2(0.0, 0.0);
        }
        if (XL) {
            const double hc = hs[uhi + 1 - ulo] * xscale;
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                const int idx = tid + q * NTC;
                if (idx < pairs) {
                    const double2 x2 = lds2(xin_a + 16u * (uint32_t)(pbase + idx));
                    wr[q].x = fma(-hc, x2.x, wr[q].x);
                    wr[q].y = fma(-hc, x2.y, wr[q].y);
                }
            }
        }
        for (int col = P.lanczos ? ulo - 1 : uhi; col >= ulo; --col) {  // (Lanczos: v_{j-1} was folded into the mat-vec)
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
                sts2(ws_a + 16u * (uint32_t)(pbase + idx), wr[q]);
                xo2[pbase + idx] = wr[q];
                nrm = fma(wr[q].x, wr[q].x, fma(wr[q].y, wr[q].y, nrm));
            }
        }
    }
    if (AUG && P.p > 0 && tid < P.p) {
        double wt = S->wtail[tid];
        if (XL) {
            const double vt = S->xtail[tid] * xscale;  // tail of the resident column
            wt = fma(-hs[uhi + 1 - ulo], vt, wt);
            if (vlazy && tm.rank == 0) vlazy[P.n + tid] = vt;
        }
        for (int c = uhi; c >= ulo; --c) wt = fma(-hs[c - ulo], V[(long long)c * P.ldv + P.n + tid], wt);
        S->wtail[tid] = wt;
        if (tm.rank == 0) {
            xout[P.n + P.nhalo + tid] = wt;
            if (P.myrank == 0) nrm = fma(wt, wt, nrm);
        }
    }
    return nrm;
}

// One problem on the consumer side, short-window (XL) instance.
//
// XL instance (short windows: Lanczos / IOP-q on a CSR-stream operator whose slice fits twice in shared memory):
//   * two slice buffers: xin = unnormalised current basis vector (v_j = xin * xscale), ws = the new w; swapped per step;
//   * mat-vec gathers own-slice columns from xin (shared memory) and accumulates <v_j, w> on the fly;
//   * v_j is stored to V during the update phase of step j (one step late), the last column by an epilogue;
//   * on one GPU both reductions of a step are packet all-reduces (ll_publish / ll_collect).
// Measured per Lanczos step at C2 before / after: see DESIGN.md 3.1c.
template <int OPK, bool AUG, bool XL, int GW>
__device__ void consumer_problem_xl(const KrylovParams &P, Cons &cx, const TmaGeom &G, Team &tm, int prob, int nlocal,
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
    const uint32_t stage_a = cx.xin_a;                // firststep! stages b in the x buffer
    const double *xsrc;
    double xscale;
    int jstart;
    int m_out = P.m, breakdown = 0;
    const bool sharded = P.nranks > 1;
    const bool via_xb0 = p > 0 || sharded;            // first gather source must carry tail / halo entries
    const double *lpart = P.peer_part[P.myrank];      // this GPU's inboxes
    const double *lpartn = P.peer_partn[P.myrank];
    const int xt = n + P.nhalo;                       // offset of the augmented tail in the gather buffers
    bool lazy_first = true;                           // XL: the first step's resident column still has to be stored
    if (XL && nlocal == 0)                            // chunk flags are learnt once per launch (geometry only)
        for (int c = tid; c < MAXCH2; c += NTC) S->chunk_local[c] = 1;

    if (P.j0 == 0) {  // firststep! (arnoldi.jl:230-250 / 257-279)
        double nrm = 0.0;
        for (int i = tid; i < units; i += NTC) {
            const double2 b2 = reinterpret_cast<const double2 *>(b + G.r0)[i];
            sts2(stage_a + 16u * (uint32_t)i, b2);
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
        if (XL && sharded) {  // push_halo reads the staged slice through cx.ws
            double *t = cx.ws; cx.ws = cx.xin; push_halo(P, cx, G, tm, xoff0); cx.ws = t;
        } else {
            push_halo(P, cx, G, tm, xoff0);
        }
        team_reduce_c(P, cx, tm, lpartn + pslot, 1, S->bc, sharded);
        const double beta = sqrt(S->bc[0]);
        if (tm.rank == 0 && tid == 0) P.scal[prob * 4] = beta;
        if (beta == 0.0) {
            if (tm.rank == 0 && tid == 0) {
                P.stat[prob * 4 + 0] = P.m;
                P.stat[prob * 4 + 1] = 0;
                P.stat[prob * 4 + 2] = 0;
            }
            return;
        }
        xsrc = via_xb0 ? xb0 : b;  // v_1 = b / beta is stored by the first step's update phase
        xscale = 1.0 / beta;
        jstart = 1;
    } else {
        const double *vj = V + (long long)(P.j0 - 1) * ldv;
        if (sharded) {  // the resumed column has no halo: stage it in the gather buffer and push the halo
            for (int i = tid; i < units; i += NTC) {
                const double2 v2 = reinterpret_cast<const double2 *>(vj + G.r0)[i];
                sts2(stage_a + 16u * (uint32_t)i, v2);
                reinterpret_cast<double2 *>(xb0 + G.r0)[i] = v2;
            }
            if (p > 0 && tm.rank == 0 && tid < p) xb0[xt + tid] = vj[n + tid];
            block_sum_to_c(P, cx, 0.0, partn0 + 2LL * P.cpad + tm.rank);  // (has the CTA barrier the push needs)
            if (XL) {
                double *t = cx.ws; cx.ws = cx.xin; push_halo(P, cx, G, tm, xoff0); cx.ws = t;
            } else {
                push_halo(P, cx, G, tm, xoff0);
            }
            team_reduce_c(P, cx, tm, lpartn + partn0 + 2LL * P.cpad, 1, S->bc, true);  // every GPU's halo is in place
            xsrc = xb0;
        } else {
            if (XL) {
                for (int i = tid; i < units; i += NTC)
                    sts2(stage_a + 16u * (uint32_t)i, reinterpret_cast<const double2 *>(vj + G.r0)[i]);
                consumer_sync();
            }
            xsrc = vj;
        }
        lazy_first = false;  // the resumed column is already in V
        xscale = 1.0;
        jstart = P.j0;
    }

    double beta_prev = 0.0, xscale_prev = 0.0;
    const int iopw = P.iop > 0 ? P.iop : P.m;
    for (int j = jstart; j <= P.m; ++j) {
        const int jc = j - 1;
        const int par = j & 1;
        double *xout = par ? xb1 : xb0;
        const long long xoff = xoff0 + (par ? P.xlen : 0);
        const long long part = part0 + (long long)par * MAXCOL * P.cpad;
        const int lo = P.lanczos ? jc : max(0, jc - iopw + 1);
        const int hi = jc;
        const int nc = hi - lo + 1;
        const int ulo = (P.lanczos && jc >= 1) ? jc - 1 : lo;

        PT_MARK(blockIdx.x, j, 0);
        const bool fold = P.lanczos && j > jstart;  // (the first step has no v_{j-1}: lanczos! restarts the recurrence)
        const double selfacc = matvec_xl<AUG, GW>(P, cx, G, xsrc, xscale, nlocal == 0 && j == jstart, fold,
                                                  beta_prev * xscale_prev, j);
        PT_MARK(blockIdx.x, j, 1);

        double beta;
        {
            const bool use_ll = sharded || nc <= LLQ;  // packet all-reduce (one GPU: ll_*, <= LLQ columns; row-sharded: shard_*, any)
            // <v_jc, w> was accumulated by the mat-vec (unscaled); its block reduction is also the barrier that
            // completes the w slice.  Columns lo..jc-1 come through the ring as before.
            {
                const double v = warp_sum(selfacc);
                if (cx.lane == 0) S->redn[cx.warp] = v;
                consumer_sync();
                if (cx.warp == 0) {  // (every lane computes the same sum)
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) s += S->redn[w];
                    s *= xscale;
                    if (AUG && p > 0 && tm.rank == 0 && P.myrank == 0)
                        for (int kk = 0; kk < p; ++kk) s = fma(S->xtail[kk] * xscale, S->wtail[kk], s);
                    if (use_ll && sharded) shard_publish_warp(P, tm, cx.seq + 1u, nc - 1, s, cx.lane, false, false);
                    else if (use_ll) ll_publish_warp(P, cx.team, tm, cx.seq + 1u, nc - 1, s, cx.lane, false);
                    else if (cx.lane == 0) P.peer_part[P.myrank][part + (long long)(nc - 1) * P.cpad + tm.rank] = s;
                }
            }
            if (nc > 1) dots_phase_xl<OPK, AUG, XL>(P, cx, G, tm, V, lo, hi - 1, part, use_ll);
            PT_MARK(blockIdx.x, j, 2);
            if (use_ll && sharded) shard_collect(P, cx, tm, nc, S->hs + (lo - ulo), false);
            else if (use_ll) ll_collect(P, cx, tm, nc, S->hs + (lo - ulo), false);
            else team_reduce_c(P, cx, tm, lpart + part, nc, S->hs + (lo - ulo), false);
            PT_MARK(blockIdx.x, j, 3);
            if (tm.rank == 0)
                for (int ci = tid; ci < nc; ci += NTC) Hd[(long long)jc * ldh + lo + ci] = S->hs[lo + ci - ulo];
            if (P.lanczos && jc >= 1 && tid == 0) S->hs[0] = beta_prev;
            consumer_sync();

            double *vlazy = (j > jstart || lazy_first) ? V + (long long)jc * ldv : nullptr;
            const double nrm = update_phase_xl<OPK, AUG, XL>(P, cx, G, tm, V, ulo, hi - 1, xout, xscale, vlazy);
            PT_MARK(blockIdx.x, j, 4);
            // block reduction of the squared norm; the release fence covers exactly the gather-buffer stores
            {
                const double v = warp_sum(nrm);
                if (cx.lane == 0) S->redn[cx.warp] = v;
                consumer_sync();
                bool pushed = false;
                if (sharded) {  // halo rows go to the peers' gather buffers before this CTA's packet may leave
                    pushed = P.send_ofs[tm.rank + 1] > P.send_ofs[tm.rank];
                    push_halo(P, cx, G, tm, xoff);
                    consumer_sync();
                }
                if (cx.warp == 0) {
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) s += S->redn[w];
                    if (!sharded) ll_publish_warp(P, cx.team, tm, cx.seq + 1u, 0, s, cx.lane, true);
                    else shard_publish_warp(P, tm, cx.seq + 1u, 0, s, cx.lane, true, pushed);
                }
            }
            PT_MARK(blockIdx.x, j, 7);
            // the resident column v_jc = xin * xscale goes to V now, one step late, between publishing this CTA's
            // norm partial and collecting everybody's: the stores overlap the packet flight time.  (After the
            // collect they cost 2.2k cycles of their own per step -- measured; in front of the release fence
            // they would have to drain before the packets may leave.)
            if (vlazy) {
                const uint32_t xin_a = cx.xin_a;
                double2 *vl2 = reinterpret_cast<double2 *>(vlazy + G.r0);
                for (int i = tid; i < units; i += NTC) {
                    const double2 x2 = lds2(xin_a + 16u * (uint32_t)i);
                    vl2[i] = make_double2(x2.x * xscale, x2.y * xscale);
                }
                if (!P.lanczos) fence_proxy_async();  // fetched by this CTA's TMA producer in later steps
            }
            PT_MARK(blockIdx.x, j, 8);
            if (!sharded) ll_collect(P, cx, tm, 1, S->bc, true);
            else shard_collect(P, cx, tm, 1, S->bc, true);
            PT_MARK(blockIdx.x, j, 5);
            if (tid == 0) flag_set(&S->cols_ready, jc + 1);
            beta = sqrt(S->bc[0]);
            if (tm.rank == 0 && tid == 0) Hd[(long long)jc * ldh + jc + 1] = beta;
            {  // the new w becomes the resident vector of the next step
                double *t = cx.ws;
                cx.ws = cx.xin;
                cx.xin = t;
                const uint32_t ta = cx.ws_a;
                cx.ws_a = cx.xin_a;
                cx.xin_a = ta;
            }
            PT_MARK(blockIdx.x, j, 6);
        }
        xsrc = xout;
        xscale_prev = xscale;
        xscale = 1.0 / beta;
        beta_prev = beta;
        if (beta < P.tol) {
            m_out = j;
            breakdown = 1;
        }
        if (XL && (breakdown || j == P.m)) {
            // epilogue: the last column v_{j+1} = w / beta (true division: beta may be tiny on breakdown,
            // arnoldi.jl:306 runs before the breakdown test)
            double *vn = V + (long long)(jc + 1) * ldv;
            const uint32_t xin_a = cx.xin_a;
            for (int i = tid; i < units; i += NTC) {
                double2 w2 = lds2(xin_a + 16u * (uint32_t)i);
                w2.x /= beta;
                w2.y /= beta;
                reinterpret_cast<double2 *>(vn + G.r0)[i] = w2;
            }
            if (p > 0 && tm.rank == 0 && tid < p) vn[n + tid] = S->wtail[tid] / beta;
        }
        if (breakdown) break;
    }
    if (tm.rank == 0 && tid == 0) {
        P.stat[prob * 4 + 0] = m_out;
        P.stat[prob * 4 + 1] = breakdown;
        P.stat[prob * 4 + 2] = 0;
    }
}

// =====================================================================================================================
// One-reduction Lanczos step (short-window instance, Hermitian operators; reference: lanczos_step!, arnoldi.jl:388-403)
// =====================================================================================================================
// The two-reduction step above waits for an all-reduce twice: alpha = <v_j, A v_j>, then ||w||^2 after the update --
// each a full latency chain over the team (4-6 k cycles on one GPU, 7-8 us over NVLink), and the second one is also the
// barrier that publishes the updated vector to the CTAs that gather it.  Here ONE all-reduce per step carries
//     S = <X_j, A v_j> (-> alpha),   Q = ||w'||^2,   G = <X_j, w'>,   N = ||X_j||^2      with  w' = A v_j - beta_{j-1} v_{j-1},
// and ||w' - alpha v_j||^2 = Q - 2 alpha g + alpha^2 nu gives beta_j without a second pass (g = xscale G,
// nu = xscale^2 N = ||v_j||^2, alpha = xscale S / nu).  nu is what makes this usable: v_j was normalised with the
// PREDICTED beta_{j-1}, so ||v_j|| = 1 + O(eps Q / beta^2); assuming nu = 1 feeds that error back into the next
// prediction and it grows ~4x per step (measured: garbage after 25 steps); with the measured nu every prediction is
// exact for the vector actually used and the error stays at rounding level (1e-14 after 100 steps).  A cancellation
// guard -- beta^2 < Q / 100, which is also where breakdown decisions are taken -- falls back to an explicit norm
// reduction.  Used for row-sharded operators (an all-reduce over NVLink costs 7-8 us); on one GPU the release fence
// behind 54 KB of gather-buffer stores per CTA makes the single reduction as expensive as the two it replaces.  What the second reduction used to publish is reconstructed by the
// readers instead: w' is written to a gather buffer DURING the mat-vec (published by the step's single all-reduce),
// and a CTA that needs entry c of the next vector outside its own slice forms
//     X_{j+1}[c] = w'_j[c] - (alpha_j xscale_j) X_j[c]
// from two published arrays (GW = w' and GX = X, both double buffered by step parity; X_{j+1} is stored to GX during
// the local update and becomes visible with the NEXT all-reduce, one step before anyone needs it).  For stencil-like
// operators only the few out-of-slice entries pay the second load.  The p augmented tail rows of kiops are kept
// redundantly by every CTA in shared memory (every CTA applies the same update with the same reduced scalars).
template <bool AUG, int GW>
__device__ void matvec_xl1(const KrylovParams &P, Cons &cx, const TmaGeom &G, const double *gw, const double *gx,
                           double gam, double xscale, bool learn, bool fold, double foldc, double *gwout,
                           double &sacc, double &qacc, double &gacc, double &nacc) {
    SmemTma *S = cx.S;
    const int tid = cx.tid;
    const int p = AUG ? P.p : 0;
    const uint32_t ws_a = cx.ws_a, xin_a = cx.xin_a;
    const uint32_t xl_a = xin_a - 8u * (uint32_t)G.r0;  // shared address of x(column) for own-slice columns
    const int nnz_cap = P.nnz_cap;
    double s1 = 0.0, q1 = 0.0, g1 = 0.0, n1 = 0.0;
    for (int c = 0; c < G.nch; ++c) {
        const int rl = c * P.ch_rows + tid;
        const bool active = tid < P.ch_rows && rl < G.nrows;
        const bool fast = !learn && c < MAXCH2 && S->chunk_local[c] != 0;
        cx.wait_full();
        bool loc = true;
        if (active) {
            const unsigned char *base = cx.rg.ptr();
            const double *vs = reinterpret_cast<const double *>(base);
            const int *cs = reinterpret_cast<const int *>(base + (size_t)nnz_cap * 8);
            const int *rp = reinterpret_cast<const int *>(base + (size_t)nnz_cap * 12);
            const int a0 = S->slot_a0[cx.rg.slot];
            const int e0 = rp[tid] - a0, e1 = rp[tid + 1] - a0;
            double sum = 0.0;
            if (fast) {
#pragma unroll 1
                for (int eb = e0; eb < e1; eb += GW) {
                    double av[GW], xv[GW];
#pragma unroll
                    for (int u = 0; u < GW; ++u) {
                        const bool ok = eb + u < e1;
                        av[u] = ok ? vs[eb + u] : 0.0;
                        xv[u] = 0.0;
                        if (ok) xv[u] = lds1(xl_a + 8u * (uint32_t)cs[eb + u]);
                    }
#pragma unroll
                    for (int u = 0; u < GW; ++u) sum = fma(av[u], xv[u], sum);
                }
            } else {
#pragma unroll 1
                for (int eb = e0; eb < e1; eb += GW) {
                    double av[GW], xv[GW];
#pragma unroll
                    for (int u = 0; u < GW; ++u) {
                        const bool ok = eb + u < e1;
                        av[u] = ok ? vs[eb + u] : 0.0;
                        xv[u] = 0.0;
                        if (ok) {
                            const int col = cs[eb + u];
                            const unsigned lc = (unsigned)(col - G.r0);
                            const bool here = lc < (unsigned)G.nrows;
                            loc = loc && here;
                            if (here) xv[u] = lds1(xin_a + 8u * lc);
                            else xv[u] = fma(-gam, gx[col], gw[col]);  // X_j[col] rebuilt from the two published arrays
                        }
                    }
#pragma unroll
                    for (int u = 0; u < GW; ++u) sum = fma(av[u], xv[u], sum);
                }
            }
            if (p > 0) {
                const double *brow = P.Bm + (G.r0 + rl);
                for (int k = 0; k < p; ++k) sum = fma(brow[(long long)k * P.ldb], S->xtail[k], sum);
            }
            double wv = sum * xscale;
            const double xr = lds1(xin_a + 8u * (uint32_t)rl);
            s1 = fma(xr, wv, s1);
            n1 = fma(xr, xr, n1);
            if (fold) wv = fma(-foldc, lds1(ws_a + 8u * (uint32_t)rl), wv);
            q1 = fma(wv, wv, q1);
            g1 = fma(xr, wv, g1);
            sts1(ws_a + 8u * (uint32_t)rl, wv);
            gwout[G.r0 + rl] = wv;
        }
        if (learn && c < MAXCH2) {
            if (!__all_sync(0xffffffffu, loc) && cx.lane == 0) S->chunk_local[c] = 0;
        }
        cx.release();
    }
    sacc = s1;
    qacc = q1;
    gacc = g1;
    nacc = n1;
}

template <int OPK, bool AUG, int GW>
__device__ void consumer_problem_xl1(const KrylovParams &P, Cons &cx, const TmaGeom &G, Team &tm, int prob, int nlocal,
                                     double *xbase, long long xoffbase, long long partn0) {
    SmemTma *S = cx.S;
    const int tid = cx.tid;
    const int n = P.n, p = AUG ? P.p : 0;
    double *V = P.V + (long long)prob * P.V_stride;
    double *Hd = P.Hd + (long long)prob * P.H_stride;
    const double *b = P.b + (long long)prob * P.b_stride;
    const long long ldv = P.ldv;
    const int ldh = P.ldh;
    const int units = G.nrows >> 1;
    const bool sharded = P.nranks > 1;
    const double *lpartn = P.peer_partn[P.myrank];
    const int xt = n + P.nhalo;  // offset of the augmented tail in a gather buffer (only the staging of step 1 uses it)
    // gather buffers of this team: GW[par] = w' of the step with parity par, GX[par] = X_j with j & 1 == par
    auto GWb = [&](int par) { return xbase + (long long)par * P.xlen; };
    auto GXb = [&](int par) { return xbase + (long long)(2 + par) * P.xlen; };
    auto GWo = [&](int par) { return xoffbase + (long long)par * P.xlen; };
    auto GXo = [&](int par) { return xoffbase + (long long)(2 + par) * P.xlen; };
    if (nlocal == 0)
        for (int c = tid; c < MAXCH2; c += NTC) S->chunk_local[c] = 1;

    // ---- X_1: b (firststep!, arnoldi.jl:230-250 / 257-279) or the normalised first column of a resumed subspace
    const double *src1;  // out-of-slice entries of X_1 are read from here
    double xscale;
    // X_1 is always stored to GX[1]: step 2 rebuilds out-of-slice entries of X_2 from w'_1 and X_1
    if (P.j0 == 0) {
        double nrm = 0.0;
        for (int i = tid; i < units; i += NTC) {
            const double2 b2 = reinterpret_cast<const double2 *>(b + G.r0)[i];
            sts2(cx.xin_a + 16u * (uint32_t)i, b2);
            reinterpret_cast<double2 *>(GXb(1) + G.r0)[i] = b2;
            nrm = fma(b2.x, b2.x, fma(b2.y, b2.y, nrm));
        }
        if (p > 0 && tid < p) {
            const double bt = P.btail[tid];
            S->xtail[tid] = bt;
            if (tm.rank == 0 && P.myrank == 0) nrm = fma(bt, bt, nrm);
        }
        const long long pslot = partn0 + (long long)(2 + (nlocal & 1)) * P.cpad;
        block_sum_to_c(P, cx, nrm, pslot + tm.rank);
        if (sharded) {
            double *t = cx.ws; cx.ws = cx.xin; push_halo(P, cx, G, tm, GXo(1)); cx.ws = t;
        }
        team_reduce_c(P, cx, tm, lpartn + pslot, 1, S->bc, sharded);
        const double beta = sqrt(S->bc[0]);
        if (tm.rank == 0 && tid == 0) P.scal[prob * 4] = beta;
        if (beta == 0.0) {
            if (tm.rank == 0 && tid == 0) {
                P.stat[prob * 4 + 0] = P.m;
                P.stat[prob * 4 + 1] = 0;
                P.stat[prob * 4 + 2] = 0;
                P.stat[prob * 4 + 3] = 0;
            }
            return;
        }
        src1 = GXb(1);  // (published by the reduction above)
        xscale = 1.0 / beta;
    } else {
        const double *vj = V;  // lanczos! restarts at column 1 (it ignores init, arnoldi.jl:480)
        for (int i = tid; i < units; i += NTC) {
            const double2 v2 = reinterpret_cast<const double2 *>(vj + G.r0)[i];
            sts2(cx.xin_a + 16u * (uint32_t)i, v2);
            reinterpret_cast<double2 *>(GXb(1) + G.r0)[i] = v2;  // (visible to the team from step 2 on)
        }
        if (p > 0 && tid < p) S->xtail[tid] = vj[n + tid];
        if (sharded) {
            block_sum_to_c(P, cx, 0.0, partn0 + 2LL * P.cpad + tm.rank);  // (has the CTA barrier the push needs)
            double *t = cx.ws; cx.ws = cx.xin; push_halo(P, cx, G, tm, GXo(1)); cx.ws = t;
            team_reduce_c(P, cx, tm, lpartn + partn0 + 2LL * P.cpad, 1, S->bc, true);
            src1 = GXb(1);
        } else {
            consumer_sync();
            src1 = vj;
        }
        xscale = 1.0;
    }
    (void)xt;

    double beta_prev = 0.0, xscale_prev = 0.0, gam = 0.0, beta = 0.0;
    int m_out = P.m, breakdown = 0, nfall = 0;
    const bool lazy_v1 = P.j0 == 0;  // (a resumed first column is already in V)
    for (int j = 1; j <= P.m; ++j) {
        const int jc = j - 1;
        const int par = j & 1, parp = par ^ 1;
        const bool fold = j > 1;
        PT_MARK(blockIdx.x, j, 0);
        // ---- tail rows of w' (every CTA keeps them): (K x)_k = x_{k+1}, last row 0
        double st = 0.0, qt = 0.0, gt = 0.0, nt0 = 0.0;
        if (p > 0) {
            consumer_sync();
            if (tid < p) {
                double wt = (tid < p - 1) ? S->xtail[tid + 1] * xscale : 0.0;
                st = S->xtail[tid] * wt;
                nt0 = S->xtail[tid] * S->xtail[tid];
                if (fold) wt = fma(-(beta_prev * xscale_prev), S->ptail[tid], wt);
                qt = wt * wt;
                gt = S->xtail[tid] * wt;
                S->wtail[tid] = wt;
            }
            if (!(tm.rank == 0 && P.myrank == 0)) st = qt = gt = nt0 = 0.0;  // counted once
        }
        double sacc, qacc, gacc, nacc;
        matvec_xl1<AUG, GW>(P, cx, G, j == 1 ? src1 : GWb(parp), j == 1 ? src1 : GXb(parp), j == 1 ? 0.0 : gam, xscale,
                            nlocal == 0 && j == 1, fold, beta_prev * xscale_prev, GWb(par), sacc, qacc, gacc, nacc);
        sacc += st;
        qacc += qt;
        gacc += gt;
        nacc += nt0;
        PT_MARK(blockIdx.x, j, 1);
        PT_MARK(blockIdx.x, j, 2);
        // ---- the step's single all-reduce (release / acquire: publishes the w' stores, the previous update's X stores
        // and, row-sharded, the halo rows pushed to the peers)
        {
            double v0 = warp_sum(sacc), v1 = warp_sum(qacc), v2 = warp_sum(gacc), v3 = warp_sum(nacc);
            if (cx.lane == 0) {
                S->red[0][cx.warp][0] = v0;
                S->red[0][cx.warp][1] = v1;
                S->red[0][cx.warp][2] = v2;
                S->red[0][cx.warp][3] = v3;
            }
            consumer_sync();  // (also: the whole w' slice of this CTA is in place)
            bool pushed = false;
            if (sharded) {
                pushed = P.send_ofs[tm.rank + 1] > P.send_ofs[tm.rank];
                push_halo(P, cx, G, tm, GWo(par));
                consumer_sync();
            }
            if (cx.warp == 0) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    s0 += S->red[0][w][0];
                    s1 += S->red[0][w][1];
                    s2 += S->red[0][w][2];
                    s3 += S->red[0][w][3];
                }
                if (!sharded) {
                    ll_publish_warp(P, cx.team, tm, cx.seq + 1u, 0, s0, cx.lane, true);
                    ll_publish_warp(P, cx.team, tm, cx.seq + 1u, 1, s1, cx.lane, false);
                    ll_publish_warp(P, cx.team, tm, cx.seq + 1u, 2, s2, cx.lane, false);
                    ll_publish_warp(P, cx.team, tm, cx.seq + 1u, 3, s3, cx.lane, false);
                } else {
                    shard_publish_warp(P, tm, cx.seq + 1u, 0, s0, cx.lane, true, pushed);
                    shard_publish_warp(P, tm, cx.seq + 1u, 1, s1, cx.lane, false, false);
                    shard_publish_warp(P, tm, cx.seq + 1u, 2, s2, cx.lane, false, false);
                    shard_publish_warp(P, tm, cx.seq + 1u, 3, s3, cx.lane, false, false);
                }
            }
        }
        PT_MARK(blockIdx.x, j, 7);
        // v_j = X_j * xscale goes to V while the packets fly
        if (j > 1 || lazy_v1) {
            double2 *vl2 = reinterpret_cast<double2 *>(V + (long long)jc * ldv + G.r0);
            const uint32_t xin_a = cx.xin_a;
            for (int i = tid; i < units; i += NTC) {
                const double2 x2 = lds2(xin_a + 16u * (uint32_t)i);
                vl2[i] = make_double2(x2.x * xscale, x2.y * xscale);
            }
            if (p > 0 && tm.rank == 0 && tid < p) V[(long long)jc * ldv + n + tid] = S->xtail[tid] * xscale;
        }
        PT_MARK(blockIdx.x, j, 8);
        if (!sharded) ll_collect(P, cx, tm, 4, S->hs, true);
        else shard_collect(P, cx, tm, 4, S->hs, true);
        PT_MARK(blockIdx.x, j, 3);
        const double nu = xscale * xscale * S->hs[3];  // ||v_j||^2 as actually used (1 + O(eps))
        const double alpha = S->hs[0] * xscale / nu, Q = S->hs[1], g = S->hs[2] * xscale;
        double beta2 = (Q - 2.0 * alpha * g) + alpha * alpha * nu;
        const bool fallback = !(beta2 > 0.01 * Q);  // cancellation (or NaN): take the norm explicitly
        consumer_sync();  // S->hs is rewritten by the next collect
        // ---- local update X_{j+1} = w' - alpha v_j (shared memory + GX[(j+1) & 1]); no reduction needed
        const double coef = alpha * xscale;
        double nrm = 0.0;
        {
            const uint32_t ws_a = cx.ws_a, xin_a = cx.xin_a;
            double2 *xo2 = reinterpret_cast<double2 *>(GXb(parp) + G.r0);
            for (int i = tid; i < units; i += NTC) {
                double2 w2 = lds2(ws_a + 16u * (uint32_t)i);
                const double2 x2 = lds2(xin_a + 16u * (uint32_t)i);
                w2.x = fma(-coef, x2.x, w2.x);
                w2.y = fma(-coef, x2.y, w2.y);
                sts2(ws_a + 16u * (uint32_t)i, w2);
                xo2[i] = w2;
                nrm = fma(w2.x, w2.x, fma(w2.y, w2.y, nrm));
            }
        }
        if (p > 0) {
            consumer_sync();
            if (tid < p) {
                const double nt = fma(-coef, S->xtail[tid], S->wtail[tid]);
                S->ptail[tid] = S->xtail[tid];
                S->wtail[tid] = nt;  // X_{j+1} tail (copied to xtail below)
                if (tm.rank == 0 && P.myrank == 0) nrm = fma(nt, nt, nrm);
            }
            consumer_sync();
            if (tid < p) S->xtail[tid] = S->wtail[tid];
        }
        if (sharded) {  // halo rows of X_{j+1}: published by the NEXT step's all-reduce, needed one step after that
            consumer_sync();
            push_halo(P, cx, G, tm, GXo(parp));
        }
        PT_MARK(blockIdx.x, j, 4);
        if (fallback) {  // (uniform: every CTA of every rank sees the same reduced values)
            const double v = warp_sum(nrm);
            if (cx.lane == 0) S->redn[cx.warp] = v;
            consumer_sync();
            if (cx.warp == 0) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) s += S->redn[w];
                if (!sharded) ll_publish_warp(P, cx.team, tm, cx.seq + 1u, 0, s, cx.lane, false);
                else shard_publish_warp(P, tm, cx.seq + 1u, 0, s, cx.lane, false, false);
            }
            if (!sharded) ll_collect(P, cx, tm, 1, S->bc, false);
            else shard_collect(P, cx, tm, 1, S->bc, false);
            beta2 = S->bc[0];
            ++nfall;
        }
        PT_MARK(blockIdx.x, j, 5);
        beta = sqrt(beta2);
        if (tm.rank == 0 && tid == 0) {
            Hd[(long long)jc * ldh + jc] = alpha;
            Hd[(long long)jc * ldh + jc + 1] = beta;
        }
        {  // the new vector becomes the resident one
            double *t = cx.ws; cx.ws = cx.xin; cx.xin = t;
            const uint32_t ta = cx.ws_a; cx.ws_a = cx.xin_a; cx.xin_a = ta;
        }
        PT_MARK(blockIdx.x, j, 6);
        gam = coef;
        xscale_prev = xscale;
        xscale = 1.0 / beta;
        beta_prev = beta;
        if (beta < P.tol) {
            m_out = j;
            breakdown = 1;
        }
        if (breakdown || j == P.m) {
            // epilogue: v_{j+1} = X_{j+1} / beta (true division: beta may be tiny on breakdown, arnoldi.jl:306)
            consumer_sync();
            double *vn = V + (long long)(jc + 1) * ldv;
            const uint32_t xin_a = cx.xin_a;
            for (int i = tid; i < units; i += NTC) {
                double2 w2 = lds2(xin_a + 16u * (uint32_t)i);
                w2.x /= beta;
                w2.y /= beta;
                reinterpret_cast<double2 *>(vn + G.r0)[i] = w2;
            }
            if (p > 0 && tm.rank == 0 && tid < p) vn[n + tid] = S->xtail[tid] / beta;
            break;
        }
    }
    if (tm.rank == 0 && tid == 0) {
        P.stat[prob * 4 + 0] = m_out;
        P.stat[prob * 4 + 1] = breakdown;
        P.stat[prob * 4 + 2] = nfall;
        P.stat[prob * 4 + 3] = 0;
    }
}

// One instance per (operator kind, augmented or not): the persistent kernel is sensitive to code size (an unused
// extra mat-vec loop cost 3-5 % everywhere), so each instance carries only the paths it can take.
template <int OPK, bool AUG, bool XL, int GW = 8, bool SAFE = false, bool LZ1 = false>
__global__ void __launch_bounds__(NT2, 1) krylov_tma_kernel(const __grid_constant__ KrylovParams P,
                                                            const __grid_constant__ CUtensorMap tmA) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemTma *S = reinterpret_cast<SmemTma *>(smem_raw);
    const size_t ws_bytes = P.w_in_smem ? (((size_t)P.slice * 8 + 127) & ~(size_t)127) : 0;
    double *ws_smem = reinterpret_cast<double *>(smem_raw + sizeof(SmemTma));
    unsigned char *ring = smem_raw + sizeof(SmemTma) + ws_bytes * (XL ? 2 : 1);

    const int tid = threadIdx.x;
    const int team = blockIdx.x / P.team_size;
    Team tm;
    tm.rank = blockIdx.x % P.team_size;
    tm.C = P.team_size;
    tm.bar = P.peer_bar[P.myrank] + team;
    // Row-sharded: the barrier target / packet sequence number continue where the previous launch on this communicator
    // stopped.  How far a launch gets is data dependent (breakdown, hand-over to the SAFE instance), so the state lives
    // in the communicator buffer on the device -- every GPU keeps an identical copy -- not in host bookkeeping.
    tm.target = P.comm_state ? P.comm_state[1] : P.bar_base;
    tm.seq = P.comm_state ? P.comm_state[0] : P.seq_base;
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

    const long long xoff0 = (long long)team * (LZ1 ? 4 : 2) * P.xlen;  // (LZ1: four gather buffers per team)
    double *xb0 = P.peer_xbuf[P.myrank] + xoff0;
    double *xb1 = xb0 + P.xlen;
    const long long part0 = (long long)team * 2 * MAXCOL * P.cpad;
    const long long partn0 = (long long)team * 4 * P.cpad;

    const bool is_producer = tid >= NTC;
    Cons cx;
    cx.S = S;
    cx.ws = P.w_in_smem ? ws_smem : (P.wglob + (long long)team * P.n + G.r0);
    cx.xin = XL ? ws_smem + ws_bytes / 8 : nullptr;
    cx.ws_a = smem_u32(ws_smem);
    cx.xin_a = cx.ws_a + (uint32_t)ws_bytes;
    cx.team = team;
    cx.tid = tid;
    cx.lane = tid & 31;
    cx.warp = tid >> 5;
    cx.seq = tm.seq;

    int nlocal = -1;
    for (int prob = team; prob < P.nprob; prob += P.nteams) {
        // SAFE instance launched behind the fast one: the step to resume at comes from the fast instance's status
        // word (0: this problem needed no re-orthogonalisation -- nothing to do)
        int j0 = P.j0;
        if (SAFE && P.j0_from_stat) {
            // first step j whose update removed most of the vector: H[j+1, j]^2 < eta^2 (||H[lo..j, j]||^2 + H[j+1, j]^2).
            // Columns the fast instance never reached are zero (no trigger); so is every column in the normal case.
            if (tid < 32) {
                const double *Hp = P.Hd + (long long)prob * P.H_stride;
                const int iopw = P.iop > 0 ? P.iop : P.m;
                int first = 0x7fffffff;
                for (int jc = tid; jc < P.m; jc += 32) {
                    const int lo = max(0, jc - iopw + 1);
                    double hsq = 0.0;
                    for (int i = lo; i <= jc; ++i) {
                        const double hv = Hp[(long long)jc * P.ldh + i];
                        hsq = fma(hv, hv, hsq);
                    }
                    const double bj = Hp[(long long)jc * P.ldh + jc + 1];
                    if (bj * bj < REORTH_ETA2 * (hsq + bj * bj)) first = min(first, jc + 1);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
                if (tid == 0) S->j0_found = first == 0x7fffffff ? 0 : first;
            }
            __syncthreads();
            j0 = S->j0_found;
            __syncthreads();
            if (j0 == 0) continue;  // (uniform for the CTA; the ring stays as the last processed problem left it)
            if (tid == 0) S->cols_ready = j0;
            __syncthreads();
        }
        ++nlocal;
        if (is_producer) {
            if (tid == NTC) {
                Ring rg{ring, P.nslot, 0, 0u};
                unsigned issued = 0;
                producer_problem<OPK, AUG, XL>(P, &tmA, S, rg, G, P.V + (long long)prob * P.V_stride, nlocal + 1, issued, 0, j0);
            }
            __syncwarp();
        } else {
            cx.rg = Ring{ring, P.nslot, 0, 0u};
            if constexpr (XL && LZ1) consumer_problem_xl1<OPK, AUG, GW>(P, cx, G, tm, prob, nlocal, xb0, xoff0, partn0);
            else if constexpr (XL) consumer_problem_xl<OPK, AUG, true, GW>(P, cx, G, tm, prob, nlocal, xb0, xb1, xoff0, part0, partn0);
            else consumer_problem<OPK, AUG, SAFE>(P, cx, G, tm, prob, nlocal, xb0, xb1, xoff0, part0, partn0, j0);
            consumer_sync();
            if (tid == 0) flag_set(&S->stop_seq, nlocal + 1);
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
    // (no CTA can get here before every CTA of every GPU has read the state above: all of them take part in the first
    // reduction of the launch; a launch without any reduction leaves the state as it is)
    if (P.comm_state && blockIdx.x == 0 && tid == 0) {
        P.comm_state[0] = cx.seq;
        P.comm_state[1] = tm.target;
    }
}

}  // namespace b200k
