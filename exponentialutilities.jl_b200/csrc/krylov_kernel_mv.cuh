// krylov_kernel_mv.cuh -- lock-step Lanczos for KV = 4 independent problems that share one CSR operator (batched expv,
// BASELINE config 5; src/arnoldi.jl:388-403, 456-490 per problem).
//
// Why: a batch on the short-window instance (krylov_kernel_tma.cuh, XL) gives every problem its own team, so the
// operator is streamed from L2 and staged through shared memory once per problem and step, and a step of every team
// pays its own two all-reduce latencies.  Here a team advances FOUR problems together: their vectors are interleaved
// (gather buffer: entry i of problem v at 4 i + v; shared memory: two planes of (v0, v1) / (v2, v3) pairs, 16 bytes per
// row and plane, so the 16-byte loads of consecutive rows are bank-conflict free), one pass over the operator chunk
// serves four mat-vecs (320 instead of 784 shared-memory bytes per row and four problems), and the four inner
// products / norms travel in one packet all-reduce (LLQ = 4 quantities).
//
// Same step structure as the XL instance: resident interleaved slice `xin` (unnormalised, v_j = xin * xscale_v), new w
// in `ws`, inner product fused into the mat-vec, beta_{j-1} v_{j-1} folded from the old contents of `ws`, basis
// columns stored one step late between publishing and collecting the norm, packet all-reduces.  Problems are
// independent: a problem whose start vector is zero, or that breaks down (beta_j < tol), or that pads the last group,
// is "dead": its scale becomes 0, it computes zeros from then on and nothing of it is recorded any more.
#pragma once
#include "krylov_kernel_tma.cuh"

namespace b200k {

constexpr int KV = 4;
static_assert(KV == LLQ, "one packet set carries the KV inner products of a step");

// CTA-wide deterministic sums of KV per-thread values; every lane of warp 0 returns them in s[].
__device__ __forceinline__ void mv_block_sum(Cons &cx, double (&x)[KV], double (&s)[KV], int buf) {
    SmemTma *S = cx.S;
#pragma unroll
    for (int v = 0; v < KV; ++v) x[v] = warp_sum(x[v]);
    if (cx.lane == 0) {
#pragma unroll
        for (int v = 0; v < KV; ++v) S->red[buf][cx.warp][v] = x[v];
    }
    consumer_sync();
    if (cx.warp == 0) {
#pragma unroll
        for (int v = 0; v < KV; ++v) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) t += S->red[buf][w][v];
            s[v] = t;
        }
    }
}

// producer: the ring carries operator chunks only
__device__ void mv_producer(const KrylovParams &P, SmemTma *S, Ring &rg, const TmaGeom &G, int seq, unsigned &issued) {
    const int nnz_cap = P.nnz_cap;
    bool stopped = false;
    for (int j = 1; j <= P.m && !stopped; ++j) {
        for (int c = 0; c < G.nch; ++c) {
            if (!prod_acquire(S, rg, seq, 0)) { stopped = true; break; }
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
            if (cnt > 0) {
                bulk_g2s(dst, P.val + a0, (uint32_t)cnt * 8u, &S->full[rg.slot]);
                bulk_g2s(dst + (size_t)nnz_cap * 8, P.colind + a0, (uint32_t)cnt * 4u, &S->full[rg.slot]);
            }
            bulk_g2s(dst + (size_t)nnz_cap * 12, P.rowptr + rs, (uint32_t)rpc * 4u, &S->full[rg.slot]);
            rg.advance();
            ++issued;
        }
    }
    while (flag_get(&S->stop_seq) < seq) __nanosleep(256);
    const unsigned ns = (unsigned)rg.nslot;
    const unsigned first = issued > ns ? issued - ns : 0u;
    for (unsigned t = first; t < issued; ++t) mbar_wait(&S->full[t % ns], (t / ns) & 1u);
}

template <int GW>
__device__ void mv_consumer(const KrylovParams &P, Cons &cx, const TmaGeom &G, Team &tm, int grp, int nlocal,
                            double *xb0, double *xb1) {
    SmemTma *S = cx.S;
    const int tid = cx.tid, lane = cx.lane;
    const int nrows = G.nrows, r0 = G.r0;
    const long long ldv = P.ldv;
    const int ldh = P.ldh;
    const uint32_t pb = 16u * (uint32_t)P.slice;  // bytes of one (v0, v1) / (v2, v3) plane of a slice buffer
    bool valid[KV], dead[KV], run[KV];
    const double *bp[KV];
    double *Vp[KV];
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        const int prob = grp * KV + v;
        valid[v] = prob < P.nprob;
        const int pq = valid[v] ? prob : grp * KV;  // (padding problems alias the first one for reads; nothing is written)
        bp[v] = P.b + (long long)pq * P.b_stride;
        Vp[v] = P.V + (long long)pq * P.V_stride;
    }
    if (nlocal == 0)
        for (int c = tid; c < MAXCH2; c += NTC) S->chunk_local[c] = 1;

    // ---- firststep! (arnoldi.jl:230-250) of the KV problems: interleave b into the resident slice and the gather buffer
    double acc[KV], s4[KV];
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        double a = 0.0;
        for (int i = tid; i < nrows; i += NTC) {
            const double bv = valid[v] ? bp[v][r0 + i] : 0.0;
            sts1(cx.xin_a + (uint32_t)(v >> 1) * pb + 16u * (uint32_t)i + 8u * (uint32_t)(v & 1), bv);
            xb0[(long long)(r0 + i) * KV + v] = bv;
            a = fma(bv, bv, a);
        }
        acc[v] = a;
    }
    mv_block_sum(cx, acc, s4, 0);
    if (cx.warp == 0) {
#pragma unroll
        for (int v = 0; v < KV; ++v) ll_publish_warp(P, cx.team, tm, cx.seq + 1u, v, s4[v], lane, v == 0);
    }
    ll_collect(P, cx, tm, KV, S->hs, true);
    double beta[KV], xscale[KV], vscale[KV], beta_prev[KV], xscale_prev[KV];
    int m_out[KV], brk[KV];
    bool any_alive = false;
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        beta[v] = sqrt(S->hs[v]);
        dead[v] = !valid[v] || beta[v] == 0.0;
        xscale[v] = dead[v] ? 0.0 : 1.0 / beta[v];
        vscale[v] = xscale[v];
        beta_prev[v] = 0.0;
        xscale_prev[v] = 0.0;
        m_out[v] = P.m;
        brk[v] = 0;
        any_alive = any_alive || !dead[v];
        if (tm.rank == 0 && tid == v && valid[v]) P.scal[(grp * KV + v) * 4] = beta[v];
    }
    consumer_sync();  // S->hs is rewritten by the next collect
#pragma unroll
    for (int v = 0; v < KV; ++v) run[v] = !dead[v];  // problems that have a basis at all (valid, beta_0 != 0)

    const double *xsrc = xb0;
    const int nnz_cap = P.nnz_cap;
    int jlast = 0;
    for (int j = 1; j <= P.m && any_alive; ++j) {
        jlast = j;
        const int jc = j - 1;
        double *xout = (j & 1) ? xb1 : xb0;
        const bool fold = j > 1;
        const bool learn = nlocal == 0 && j == 1;
        double foldc[KV];
#pragma unroll
        for (int v = 0; v < KV; ++v) {
            foldc[v] = beta_prev[v] * xscale_prev[v];
            acc[v] = 0.0;
        }
        const uint32_t xin_a = cx.xin_a, ws_a = cx.ws_a;
        PT_MARK(blockIdx.x, j, 0);

        // ---- mat-vec of the KV problems + <v_j, A v_j> + fold of beta_{j-1} v_{j-1}
        for (int c = 0; c < G.nch; ++c) {
            const int rl = c * P.ch_rows + tid;
            const bool active = tid < P.ch_rows && rl < nrows;
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
                double sum[KV];
#pragma unroll
                for (int v = 0; v < KV; ++v) sum[v] = 0.0;
#pragma unroll 1
                for (int eb = e0; eb < e1; eb += GW) {
                    double av[GW];
                    int cv[GW];
#pragma unroll
                    for (int u = 0; u < GW; ++u) {
                        const bool ok = eb + u < e1;
                        av[u] = ok ? vs[eb + u] : 0.0;
                        cv[u] = ok ? cs[eb + u] : -1;
                    }
#pragma unroll
                    for (int u = 0; u < GW; ++u) {
                        if (cv[u] >= 0) {
                            double2 x01, x23;
                            const unsigned lc = (unsigned)(cv[u] - r0);
                            const bool here = fast || lc < (unsigned)nrows;
                            loc = loc && here;
                            if (here) {
                                const uint32_t a = xin_a + 16u * lc;
                                x01 = lds2(a);
                                x23 = lds2(a + pb);
                            } else {
                                const double2 *g = reinterpret_cast<const double2 *>(xsrc + (long long)cv[u] * KV);
                                x01 = g[0];
                                x23 = g[1];
                            }
                            sum[0] = fma(av[u], x01.x, sum[0]);
                            sum[1] = fma(av[u], x01.y, sum[1]);
                            sum[2] = fma(av[u], x23.x, sum[2]);
                            sum[3] = fma(av[u], x23.y, sum[3]);
                        }
                    }
                }
                const uint32_t ax = xin_a + 16u * (uint32_t)rl, aw = ws_a + 16u * (uint32_t)rl;
                const double2 xi01 = lds2(ax), xi23 = lds2(ax + pb);
                double wv[KV];
#pragma unroll
                for (int v = 0; v < KV; ++v) wv[v] = sum[v] * xscale[v];
                acc[0] = fma(xi01.x, wv[0], acc[0]);
                acc[1] = fma(xi01.y, wv[1], acc[1]);
                acc[2] = fma(xi23.x, wv[2], acc[2]);
                acc[3] = fma(xi23.y, wv[3], acc[3]);
                if (fold) {
                    const double2 wo01 = lds2(aw), wo23 = lds2(aw + pb);
                    wv[0] = fma(-foldc[0], wo01.x, wv[0]);
                    wv[1] = fma(-foldc[1], wo01.y, wv[1]);
                    wv[2] = fma(-foldc[2], wo23.x, wv[2]);
                    wv[3] = fma(-foldc[3], wo23.y, wv[3]);
                }
                sts2(aw, make_double2(wv[0], wv[1]));
                sts2(aw + pb, make_double2(wv[2], wv[3]));
            }
            if (learn && c < MAXCH2) {
                if (!__all_sync(0xffffffffu, loc) && lane == 0) S->chunk_local[c] = 0;
            }
            cx.release();
        }

        PT_MARK(blockIdx.x, j, 1);
        // ---- alpha_v (its block reduction is the barrier that completes the w slice)
        mv_block_sum(cx, acc, s4, 1);
        if (cx.warp == 0) {
#pragma unroll
            for (int v = 0; v < KV; ++v) ll_publish_warp(P, cx.team, tm, cx.seq + 1u, v, s4[v] * xscale[v], lane, false);
        }
        PT_MARK(blockIdx.x, j, 2);
        ll_collect(P, cx, tm, KV, S->hs, false);
        PT_MARK(blockIdx.x, j, 3);
        double coef[KV];
#pragma unroll
        for (int v = 0; v < KV; ++v) {
            const double alpha = S->hs[v];
            coef[v] = alpha * xscale[v];
            if (tm.rank == 0 && tid == v && !dead[v])
                P.Hd[(long long)(grp * KV + v) * P.H_stride + (long long)jc * ldh + jc] = alpha;
        }

        // ---- w -= alpha v_j (v_{j-1} was folded), squared norms, unnormalised w to the gather buffer
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double clo = coef[2 * h], chi = coef[2 * h + 1];
            double nlo = 0.0, nhi = 0.0;
            double2 *xo2 = reinterpret_cast<double2 *>(xout + (long long)r0 * KV) + h;
            for (int i = tid; i < nrows; i += NTC) {
                const uint32_t a = (uint32_t)h * pb + 16u * (uint32_t)i;
                double2 w2 = lds2(ws_a + a);
                const double2 x2 = lds2(xin_a + a);
                w2.x = fma(-clo, x2.x, w2.x);
                w2.y = fma(-chi, x2.y, w2.y);
                sts2(ws_a + a, w2);
                xo2[2 * i] = w2;
                nlo = fma(w2.x, w2.x, nlo);
                nhi = fma(w2.y, w2.y, nhi);
            }
            acc[2 * h] = nlo;
            acc[2 * h + 1] = nhi;
        }
        PT_MARK(blockIdx.x, j, 4);
        mv_block_sum(cx, acc, s4, 0);
        if (cx.warp == 0) {
#pragma unroll
            for (int v = 0; v < KV; ++v) ll_publish_warp(P, cx.team, tm, cx.seq + 1u, v, s4[v], lane, v == 0);
        }
        PT_MARK(blockIdx.x, j, 7);
        // ---- column jc of every live problem goes to V now, one step late (overlaps the packet flight time)
#pragma unroll
        for (int v = 0; v < KV; ++v) {
            if (run[v] && jc <= m_out[v]) {  // (a problem that broke down at step m_out still gets column m_out, then stops)
                double *col = Vp[v] + (long long)jc * ldv + r0;
                const double sc = vscale[v];
                const uint32_t a = xin_a + (uint32_t)(v >> 1) * pb + 8u * (uint32_t)(v & 1);
                for (int i = tid; i < nrows; i += NTC) col[i] = lds1(a + 16u * (uint32_t)i) * sc;
            }
        }
        PT_MARK(blockIdx.x, j, 8);
        ll_collect(P, cx, tm, KV, S->hs, true);
        PT_MARK(blockIdx.x, j, 5);
        any_alive = false;
#pragma unroll
        for (int v = 0; v < KV; ++v) {
            const double bt = sqrt(S->hs[v]);
            if (!dead[v]) {
                if (tm.rank == 0 && tid == v)
                    P.Hd[(long long)(grp * KV + v) * P.H_stride + (long long)jc * ldh + jc + 1] = bt;
                beta_prev[v] = bt;
                xscale_prev[v] = xscale[v];
                vscale[v] = 1.0 / bt;
                beta[v] = bt;
                if (bt < P.tol) {  // happy breakdown of this problem (arnoldi.jl:370-374)
                    m_out[v] = j;
                    brk[v] = 1;
                    dead[v] = true;
                    xscale[v] = 0.0;
                    beta_prev[v] = 0.0;
                    xscale_prev[v] = 0.0;
                } else {
                    xscale[v] = vscale[v];
                }
            } else {
                beta_prev[v] = 0.0;
                xscale_prev[v] = 0.0;
            }
            any_alive = any_alive || !dead[v];
        }
        consumer_sync();  // S->hs is rewritten by the next collect; the lazy stores read xin, which becomes ws
        {
            const uint32_t t = cx.ws_a;
            cx.ws_a = cx.xin_a;
            cx.xin_a = t;
        }
        xsrc = xout;
        PT_MARK(blockIdx.x, j, 6);
    }
    // ---- epilogue: column jlast of every problem whose last step was jlast (still running, or broke down in it);
    // true division: beta may be tiny on breakdown (arnoldi.jl:306 runs before the breakdown test)
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        if (run[v] && jlast > 0 && (dead[v] ? m_out[v] : jlast) == jlast) {
            double *col = Vp[v] + (long long)jlast * ldv + r0;
            const double bt = beta[v];
            const uint32_t a = cx.xin_a + (uint32_t)(v >> 1) * pb + 8u * (uint32_t)(v & 1);
            for (int i = tid; i < nrows; i += NTC) col[i] = lds1(a + 16u * (uint32_t)i) / bt;
        }
        if (tm.rank == 0 && tid == v && valid[v]) {
            const int prob = grp * KV + v;
            P.stat[prob * 4 + 0] = m_out[v];
            P.stat[prob * 4 + 1] = brk[v];
        }
    }
}

template <int GW>
__global__ void __launch_bounds__(NT2, 1) krylov_mv_kernel(const __grid_constant__ KrylovParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemTma *S = reinterpret_cast<SmemTma *>(smem_raw);
    const size_t buf_bytes = ((size_t)P.slice * KV * 8 + 127) & ~(size_t)127;
    unsigned char *bufs = smem_raw + sizeof(SmemTma);
    unsigned char *ring = bufs + 2 * buf_bytes;

    const int tid = threadIdx.x;
    const int team = blockIdx.x / P.team_size;
    Team tm;
    tm.rank = blockIdx.x % P.team_size;
    tm.C = P.team_size;
    tm.bar = nullptr;
    tm.target = 0;
    tm.seq = P.seq_base;
    TmaGeom G;
    G.r0 = min(P.n, tm.rank * P.slice);
    G.nrows = min(P.n, G.r0 + P.slice) - G.r0;
    G.TR = P.tile_rows;
    G.ntk = 0;
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
        S->cols_ready = 0;
        S->stop_seq = 0;
        mbar_fence_init();
    }
    __syncthreads();

    double *xb0 = P.peer_xbuf[0] + (long long)team * 2 * P.xlen * KV;
    double *xb1 = xb0 + P.xlen * KV;
    Cons cx;
    cx.S = S;
    cx.ws = nullptr;
    cx.xin = nullptr;
    cx.ws_a = smem_u32(bufs);
    cx.xin_a = cx.ws_a + (uint32_t)buf_bytes;
    cx.team = team;
    cx.tid = tid;
    cx.lane = tid & 31;
    cx.warp = tid >> 5;
    cx.seq = P.seq_base;

    const int ngroups = (P.nprob + KV - 1) / KV;
    const bool is_producer = tid >= NTC;
    int nlocal = -1;
    for (int grp = team; grp < ngroups; grp += P.nteams) {
        ++nlocal;
        if (is_producer) {
            if (tid == NTC) {
                Ring rg{ring, P.nslot, 0, 0u};
                unsigned issued = 0;
                mv_producer(P, S, rg, G, nlocal + 1, issued);
            }
            __syncwarp();
        } else {
            cx.rg = Ring{ring, P.nslot, 0, 0u};
            mv_consumer<GW>(P, cx, G, tm, grp, nlocal, xb0, xb1);
            consumer_sync();
            if (tid == 0) flag_set(&S->stop_seq, nlocal + 1);
        }
        __syncthreads();
        if (grp + P.nteams < ngroups) {
            if (tid == 0) {
                for (int s = 0; s < P.nslot; ++s) {
                    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&S->full[s])) : "memory");
                    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&S->empty[s])) : "memory");
                    mbar_init(&S->full[s], 1);
                    mbar_init(&S->empty[s], NW);
                }
                mbar_fence_init();
            }
            __syncthreads();
        }
    }
}

}  // namespace b200k
