// b200krylov.cu -- C ABI (include/b200krylov.h) of the B200 Krylov expmv/phiv engine.
// Host orchestration only: geometry, scratch, launches, the small dense phase on the host H, kiops.
#include "../../include/b200krylov.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <atomic>
#include <vector>

#include "aux_kernels.cuh"
#include "krylov_kernel.cuh"
#include "krylov_kernel_tma.cuh"
#include "krylov_kernel_mv.cuh"
#include "krylov_kernel_z.cuh"
#include "krylov_kernel_tma_z.cuh"
#include "smallexp_kernel.cuh"
#include "smallmat.hpp"

using namespace b200k;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostBuf {  // pinned
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes + 256);
        if (e == cudaSuccess) {
            cap = bytes + 256;
            std::memset(p, 0, cap);
        }
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

inline long long round_up(long long x, long long a) { return (x + a - 1) / a * a; }

}  // namespace

struct b200k_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    int max_ctas = 0;  // co-resident CTAs of the persistent kernel
    void *encode_tiled = nullptr;  // cuTensorMapEncodeTiled, fetched through the runtime (no libcuda link dependency)
    int l2hint = -1;        // B200K_FLAG_L2HINT: -1 automatic, 0 never, 1 always evict_first for operator chunks
    long long l2_bytes = 0;
    int no_lz1 = 0;         // B200K_FLAG_NO_LZ1: 0 = one-reduction Lanczos step for row-sharded operators only, 1 = never, 2 = always
    int sym_pade = 0;       // B200K_FLAG_SYM_PADE: device small exponential of a Lanczos H by Pade instead of Chebyshev
    int host_smallexp = 0;  // B200K_SMALLEXP=host: one-shot / batched expv do the small exponential on the host
    DevBuf tdev, errdev;
    HostBuf errh;
    int no_mv = 0;     // B200K_FLAG_NO_MV / B200K_MV: 2 = batched Lanczos uses the lock-step multi-vector kernel whenever it
                       // is possible; 0 (default) and 1 = per-problem teams (the kernel is opt-in, see b200k_expv_batched)
    int no_xl = 0;     // B200K_FLAG_NO_XL / B200K_XL=0: never use the short-window (XL) instance
    int last_xl = 0;   // the last persistent launch used the XL instance
    DevBuf llpkt;      // packet all-reduce inboxes of the XL instance [2][LLQ][CPAD dest][CPAD src] x 16 B, zeroed at allocation
    unsigned ll_seq = 0;  // packet sequence numbers consumed by earlier launches
    int force_ldg = 0; // B200K_KERNEL=ldg: use the LDG kernel even where the TMA-ring kernel applies
    int last_kernel = 0;  // 1 = LDG kernel, 2 = TMA-ring kernel, 3 = complex kernel, 4 = TMA-ring kernel, XL instance
    std::string err;
    int64_t launches = 0;
    // scratch (device)
    DevBuf xbuf, part, partn, bar, wglob, Hd, scal, stat, btail, Y, corr, mvec, betavec, tmp;
    // scratch (host, pinned)
    HostBuf Hh, scalh, stath;
    // pinned staging for host -> device coefficient copies: a ring of slots, each guarded by an event recorded after
    // its copy was queued, so a slot is never rewritten while an earlier cudaMemcpyAsync may still read it
    static constexpr int NSTAGE = 8;
    HostBuf stage[NSTAGE];
    cudaEvent_t stage_ev[NSTAGE] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool stage_busy[NSTAGE] = {false, false, false, false, false, false, false, false};
    int stage_next = 0, stage_cur = 0;
    // internal Krylov storage for the one-shot calls
    DevBuf V, bdev, wdev;
    DevBuf tsV, tsW, tsP, tsu;  // phiv_timestep workspace (basis, W, P, u)
    DevBuf zV, zH, zpart, zxbuf, zw, zy;  // complex path: internal basis, device H, partial sums, gather buffers
    HostBuf zHh;
    DevBuf kV, kB;  // kiops basis / flipped-u storage, kept across calls (cudaMalloc/cudaFree cost milliseconds)
    std::vector<double> H;
    smallmat::ExpWork expwork;
    // timing
    int timing = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float krylov_ms = 0.f, project_ms = 0.f;
    bool ev_k = false, ev_p = false;  // which event pairs have been recorded (elapsed time on unrecorded events errors)
};

struct b200k_comm {
    b200k_context *ctx = nullptr;
    int device = 0;
    int rank = 0, nranks = 1;
    long long xlen = 0;
    int cpad = 0;
    size_t bytes = 0;
    void *local = nullptr;
    void *peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool connected = false;
    unsigned seq_upper = 0;  // upper bound of the packet sequence numbers consumed so far (exhaustion check only; the
                             // exact barrier target / sequence number are carried on the device, see state())
    // {sequence number, barrier target} of the next launch, at byte 512 of the header (written by the kernels)
    unsigned *state() const { return reinterpret_cast<unsigned *>(reinterpret_cast<char *>(local) + 512); }
    // layout of each rank's buffer: [header: barrier counter @0, LL packet inbox @1024: 2 x (MAXCOL+1) x 8 x 16 B,
    // local packet inbox of the XL instance]
    // [part 2*MAXCOL*cpad][partn 4*cpad][xbuf 4*xlen] doubles (four gather buffers: the one-reduction Lanczos step)
    static constexpr size_t PKT_BYTES = (size_t)2 * (MAXCOL + 1) * 8 * 16;
    // + this GPU's own packet inbox of the short-window instance: [2 parities][LLQ quantities][CPAD source CTAs]
    static constexpr size_t LLLOC_BYTES = (size_t)2 * LLLOCQ * CPAD * 16;
    static constexpr size_t HDR = 1024 + PKT_BYTES + LLLOC_BYTES;
    uint4 *llloc() const { return reinterpret_cast<uint4 *>(reinterpret_cast<char *>(local) + 1024 + PKT_BYTES); }
    unsigned *bar_of(int r) const { return reinterpret_cast<unsigned *>(peer[r]); }
    uint4 *pkt_of(int r) const { return reinterpret_cast<uint4 *>(reinterpret_cast<char *>(peer[r]) + 1024); }
    double *part_of(int r) const { return reinterpret_cast<double *>(reinterpret_cast<char *>(peer[r]) + HDR); }
    double *partn_of(int r) const { return part_of(r) + (size_t)2 * MAXCOL * cpad; }
    double *xbuf_of(int r) const { return partn_of(r) + (size_t)4 * cpad; }
};

struct b200k_operator {
    b200k_comm *comm = nullptr;  // row-sharded operator
    long long nhalo = 0;
    DevBuf send_row, send_peer, send_pos, send_ofs;
    int send_ofs_C = -1, send_ofs_slice = -1;  // geometry the uploaded send_ofs table was built for
    std::vector<int> send_row_host;
    b200k_context *ctx = nullptr;  // creating handle (NOT dereferenced on destroy: it may already be gone)
    int device = 0;
    int kind = 0;  // 0 CSR, 1 dense
    int is_complex = 0;  // ComplexF64 entries: val / Ad hold interleaved (re, im) pairs
    long long n = 0, nnz = 0;
    DevBuf rowptr, colind, val;  // owned, 0-based, padded
    const double *Ad = nullptr;  // dense: alias (device input) or owned copy
    long long ncols = 0;         // dense row block of a row-sharded operator: global dimension (0: square)
    DevBuf Aown;
    long long lda = 0;
    int max_row_nnz = 0;
    int is_herm = 0;
    double opnorm_inf = 0.0;
};

namespace {

int fail(b200k_context *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    return code;
}

#define CK(h, call)                                                                                        \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return fail((h), B200K_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));            \
    } while (0)

// Pinned staging slot of at least `bytes` bytes that no queued copy still reads (nullptr on failure); after queueing
// the copies out of it on h->stream call stage_commit.
void *stage_acquire(b200k_context *h, size_t bytes) {
    const int s = h->stage_next;
    h->stage_next = (s + 1) % b200k_context::NSTAGE;
    if (h->stage_busy[s]) {
        if (cudaEventSynchronize(h->stage_ev[s]) != cudaSuccess) return nullptr;
        h->stage_busy[s] = false;
    }
    if (h->stage[s].ensure(bytes) != cudaSuccess) return nullptr;
    h->stage_cur = s;
    return h->stage[s].p;
}
cudaError_t stage_commit(b200k_context *h) {
    const int s = h->stage_cur;
    if (!h->stage_ev[s]) {
        cudaError_t e = cudaEventCreateWithFlags(&h->stage_ev[s], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(h->stage_ev[s], h->stream);
    if (e == cudaSuccess) h->stage_busy[s] = true;
    return e;
}

struct Geom {
    int C = 1, nteams = 1, slice = 16, w_in_smem = 1;
    size_t smem = 0;
};

constexpr size_t SMEM_LIMIT = 232448;  // 227 KB opt-in maximum per CTA

// Rows per CTA and shared-memory footprint for a team of C CTAs.
Geom make_geom(long long n, int C, int nteams) {
    Geom g;
    g.C = C;
    g.nteams = nteams;
    g.slice = (int)round_up((n + C - 1) / C, 16);
    if (g.slice < 16) g.slice = 16;
    const size_t need = sizeof(SmemFixed) + (size_t)g.slice * 8;
    g.w_in_smem = need <= SMEM_LIMIT ? 1 : 0;
    g.smem = g.w_in_smem ? need : sizeof(SmemFixed);
    return g;
}

Geom single_geom(b200k_context *h, long long n) {
    int C = (int)std::min<long long>(h->max_ctas, std::max<long long>(1, (n + 15) / 16));
    return make_geom(n, C, 1);
}

// Team size for a batch of nb problems: maximise problems in flight x SM use, w slice in shared memory.
// Short-window batches (Lanczos / IOP on the XL instance) are latency-bound, not bandwidth-bound: a step costs a
// fixed part (two packet all-reduces, ~7000 + 40 C cycles for a team of C CTAs -- measured at C = 37 and 148) plus
// ~3 cycles per row of the CTA's slice, so smaller teams (more problems in flight) win as long as the two slice
// buffers and two ring slots still fit in shared memory.  B200K_BATCH_TEAM=C forces the team size (experiments).
Geom batch_geom(b200k_context *h, long long n, int nb, bool short_window, int m) {
    if (const char *env = std::getenv("B200K_BATCH_TEAM")) {
        const int C = std::max(1, std::min(std::atoi(env), h->max_ctas));
        return make_geom(n, C, std::min(h->max_ctas / C, nb));
    }
    if (short_window && !h->no_xl) {
        Geom bestx;
        double best_cost = -1.0;
        for (int C = 1; C <= h->max_ctas; ++C) {
            const int nteams = std::min(h->max_ctas / C, nb);
            if (nteams < 1) break;
            Geom g = make_geom(n, C, nteams);
            if ((long long)g.slice * (C - 1) >= n && C > 1) continue;  // empty trailing CTAs
            const size_t need = sizeof(SmemTma) + 2 * (size_t)round_up((long long)g.slice * 8, 128) + 2 * (size_t)SLOT_BYTES;
            if (need > SMEM_LIMIT) continue;
            const int rounds = (nb + nteams - 1) / nteams;
            const double cost = (double)rounds * (7000.0 + 40.0 * C + 3.0 * g.slice);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                bestx = g;
            }
        }
        if (best_cost >= 0) return bestx;
    }
    Geom best;
    double best_score = -1.0;
    for (int C = 1; C <= h->max_ctas; ++C) {
        int nteams = h->max_ctas / C;
        if (nteams < 1) break;
        nteams = std::min(nteams, nb);
        Geom g = make_geom(n, C, nteams);
        if (!g.w_in_smem) continue;
        if ((long long)g.slice * (C - 1) >= n && C > 1) continue;  // empty trailing CTAs
        const int rounds = (nb + nteams - 1) / nteams;
        // problems per round x SM use, times two measured effects (C5 sweep over 16 team sizes, profiles/r2_c5_team_sweep.md):
        // throughput grows linearly with the bytes of a basis tile (every tile costs one mbarrier round trip: 16.8 KB
        // tiles 6.6 k, 32 KB tiles 8.8 k expv/s per fully used GPU), and it drops once the bases of the teams in flight
        // (nteams x (m + 1) columns) exceed about twice the L2.
        const int ntk0 = (g.slice + TILE_ROWS_MAX - 1) / TILE_ROWS_MAX;
        const double tile_bytes = 8.0 * (double)round_up((g.slice + ntk0 - 1) / ntk0, 16);
        const double ws_mb = (double)nteams * (double)(m + 1) * (double)n * 8.0 / 1e6;
        const double l2pen = std::max(0.5, 1.0 - 0.0007 * std::max(0.0, ws_mb - 2.0 * (double)h->l2_bytes / 1e6));
        const double score = (double)nb / ((double)rounds * nteams) * ((double)nteams * C / h->max_ctas) *
                             (tile_bytes + 30720.0) / (32768.0 + 30720.0) * l2pen;
        if (score > best_score + 1e-12) {
            best_score = score;
            best = g;
        }
    }
    if (best_score < 0) best = single_geom(h, n);
    return best;
}

struct KrylovCall {
    b200k_operator *op;
    const double *b;
    long long b_stride;
    int nprob;
    double *V;
    long long ldv, V_stride;
    int m, j0, iop, lanczos;
    double tol;
    int p;
    const double *B;
    long long ldb;
    const double *btail_host;
    Geom g;
};

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }


// Launch the persistent kernel; results land in h->Hd / h->scal / h->stat (device), ldhd = m + 1.
int launch_krylov(b200k_context *h, const KrylovCall &c) {
    b200k_operator *op = c.op;
    const long long n = op->n;
    KrylovParams P;
    std::memset(&P, 0, sizeof(P));
    P.n = (int)n;
    if (op->kind == 0) {
        P.rowptr = op->rowptr.as<int>();
        P.colind = op->colind.as<int>();
        P.val = op->val.as<double>();
        int ch_rows = op->max_row_nnz > 0 ? CH_NNZ / op->max_row_nnz : NT;
        ch_rows = std::min(ch_rows, NT) / 32 * 32;
        if (ch_rows >= 64) {
            P.op_kind = OP_CSR_STREAM;
            P.ch_rows = ch_rows;
        } else {
            P.op_kind = OP_CSR_WARP;
            P.ch_rows = 32;
        }
    } else {
        P.op_kind = OP_DENSE;
        P.Ad = op->Ad;
        P.lda = op->lda;
        P.ncols = (int)(op->ncols ? op->ncols : n);
    }
    P.p = c.p;
    P.Bm = c.B;
    P.ldb = c.ldb;
    P.team_size = c.g.C;
    P.nteams = c.g.nteams;
    P.nprob = c.nprob;
    P.slice = c.g.slice;
    P.b = c.b;
    P.b_stride = c.b_stride;
    P.V = c.V;
    P.ldv = c.ldv;
    P.V_stride = c.V_stride;
    P.m = c.m;
    P.j0 = c.j0;
    P.iop = c.iop;
    P.lanczos = c.lanczos;
    P.tol = c.tol;
    P.w_in_smem = c.g.w_in_smem;
    const int ldhd = c.m + 1;
    P.ldh = ldhd;
    P.H_stride = (long long)ldhd * (c.m + 1);
    P.xlen = round_up(n + c.p, 16);

    bool vec2 = (n % 2 == 0) && (c.ldv % 2 == 0) && (c.V_stride % 2 == 0) && aligned16(c.V) &&
                (c.j0 != 0 || (aligned16(c.b) && c.b_stride % 2 == 0));
    if (op->kind == 1) vec2 = vec2 && aligned16(op->Ad) && (op->lda % 2 == 0);
    P.vec2 = vec2 ? 1 : 0;

    const int nt = c.g.nteams;
    CK(h, h->xbuf.ensure((size_t)nt * 4 * P.xlen * 8));  // (four gather buffers per team: one-reduction Lanczos)
    CK(h, h->part.ensure((size_t)nt * 2 * MAXCOL * CPAD * 8));
    CK(h, h->partn.ensure((size_t)nt * 4 * CPAD * 8));
    CK(h, h->bar.ensure((size_t)nt * 4));
    if (!c.g.w_in_smem) CK(h, h->wglob.ensure((size_t)nt * n * 8));
    CK(h, h->Hd.ensure((size_t)c.nprob * P.H_stride * 8));
    CK(h, h->scal.ensure((size_t)c.nprob * 4 * 8));
    CK(h, h->stat.ensure((size_t)c.nprob * 4 * 4));
    CK(h, h->btail.ensure(MAXP * 8));
    P.xbuf = h->xbuf.as<double>();
    P.part = h->part.as<double>();
    P.partn = h->partn.as<double>();
    P.bar = h->bar.as<unsigned>();
    P.wglob = h->wglob.as<double>();
    P.Hd = h->Hd.as<double>();
    P.scal = h->scal.as<double>();
    P.stat = h->stat.as<int>();
    P.btail = h->btail.as<double>();

    // team spanning GPUs (row-sharded operator) or a single GPU (peer tables point at the local buffers)
    b200k_comm *cm = op->comm;
    P.nranks = 1;
    P.myrank = 0;
    P.nhalo = 0;
    P.cpad = CPAD;
    P.peer_part[0] = P.part;
    P.peer_partn[0] = P.partn;
    P.peer_bar[0] = P.bar;
    P.peer_xbuf[0] = P.xbuf;
    unsigned bar_base = 0;
    if (cm) {
        if (!cm->connected) return fail(h, B200K_ECOMM, "communicator is not connected");
        if (!vec2 || h->force_ldg || c.nprob != 1 || c.g.nteams != 1)
            return fail(h, B200K_EUNSUPPORTED, "row-sharded operators need even nloc/ldv, 16-byte aligned vectors, one problem");
        if (n + op->nhalo + c.p > cm->xlen) return fail(h, B200K_EDIM, "communicator gather buffer (xlen) too small");
        if (cm->seq_upper > 0xf0000000u)  // packet sequence numbers are never reused: ~1.6e7 factorisations of 30 steps
            return fail(h, B200K_ECOMM, "communicator sequence numbers exhausted: destroy and re-create the communicator");
        P.nranks = cm->nranks;
        P.myrank = cm->rank;
        P.nhalo = (int)op->nhalo;
        P.cpad = cm->cpad;
        P.xlen = cm->xlen;
        for (int r = 0; r < cm->nranks; ++r) {
            P.peer_part[r] = cm->part_of(r);
            P.peer_partn[r] = cm->partn_of(r);
            P.peer_bar[r] = cm->bar_of(r);
            P.peer_pkt[r] = cm->pkt_of(r);
            P.peer_xbuf[r] = cm->xbuf_of(r);
        }
        // halo push ranges per CTA slice (depend on the geometry only: uploaded once per (operator, team size, slice))
        if (op->send_ofs_C != c.g.C || op->send_ofs_slice != c.g.slice) {
            std::vector<int> ofs(c.g.C + 1, 0);
            const std::vector<int> &sr = op->send_row_host;
            for (int q = 0; q <= c.g.C; ++q) {
                const long long bound = std::min<long long>(n, (long long)q * c.g.slice);
                ofs[q] = (int)(std::lower_bound(sr.begin(), sr.end(), (int)bound) - sr.begin());
            }
            ofs[c.g.C] = (int)sr.size();
            CK(h, cudaStreamSynchronize(h->stream));  // (an earlier launch may still read the old table)
            CK(h, op->send_ofs.ensure((size_t)(c.g.C + 1) * 4));
            CK(h, cudaMemcpyAsync(op->send_ofs.p, ofs.data(), (size_t)(c.g.C + 1) * 4, cudaMemcpyHostToDevice, h->stream));
            CK(h, cudaStreamSynchronize(h->stream));  // ofs is a stack vector
            op->send_ofs_C = c.g.C;
            op->send_ofs_slice = c.g.slice;
        }
        P.send_row = op->send_row.as<int>();
        P.send_peer = op->send_peer.as<int>();
        P.send_pos = op->send_pos.as<int>();
        P.send_ofs = op->send_ofs.as<int>();
        P.comm_state = cm->state();
        P.llloc = cm->llloc();  // level-1 packet inbox of the two-level all-reduce (both instances)
        // (upper bound of the sequence numbers used so far: <= 2 reductions per step + 2 per re-orthogonalised step
        //  + prologue, for the fast and the SAFE instance)
        cm->seq_upper += 8u * (unsigned)(c.m + 2);
    } else {
        CK(h, cudaMemsetAsync(P.bar, 0, (size_t)nt * 4, h->stream));
    }
    P.bar_base = bar_base;
    CK(h, cudaMemsetAsync(P.Hd, 0, (size_t)c.nprob * P.H_stride * 8, h->stream));
    CK(h, cudaMemsetAsync(P.scal, 0, (size_t)c.nprob * 4 * 8, h->stream));
    if (c.p > 0 && c.j0 == 0)
        CK(h, cudaMemcpyAsync(h->btail.p, c.btail_host, (size_t)c.p * 8, cudaMemcpyHostToDevice, h->stream));

    // ---- kernel selection: TMA-ring kernel whenever the layout allows 16-byte-aligned bulk copies ----
    size_t smem = c.g.smem;
    CUtensorMap tmapA;
    std::memset(&tmapA, 0, sizeof(tmapA));
    const void *kern = (const void *)krylov_persistent_kernel;
    const void *kern_safe = nullptr;
    bool lz1 = false;
    int threads = NT;
    if (vec2 && !h->force_ldg) {
        const size_t fixed = sizeof(SmemTma);
        const size_t extra = op->kind == 1 ? (size_t)NTC * 2 * 8 : 0;
        size_t wsb = (size_t)round_up((long long)c.g.slice * 8, 128);
        int nslot = (int)((SMEM_LIMIT - fixed - extra - wsb) / SLOT_BYTES);
        P.w_in_smem = 1;
        if (SMEM_LIMIT < fixed + extra + wsb || nslot < 3) {  // slice too large: w lives in HBM/L2 scratch
            P.w_in_smem = 0;
            wsb = 0;
            nslot = (int)((SMEM_LIMIT - fixed - extra) / SLOT_BYTES);
        }
        nslot = std::min(nslot, MAXSLOT);
        P.nslot = nslot;
        const int ntk0 = (c.g.slice + TILE_ROWS_MAX - 1) / TILE_ROWS_MAX;
        P.tile_rows = (int)round_up((c.g.slice + ntk0 - 1) / ntk0, 16);
        if (op->kind == 0) {
            const int mrn = std::max(op->max_row_nnz, 1);
            int ch_rows = (int)((SLOT_BYTES - 12 * 8 - 16) / (12LL * mrn + 4));
            ch_rows = std::min(ch_rows, NTC) / 32 * 32;
            if (ch_rows >= 64) {
                P.op_kind = OP_CSR_STREAM;
                P.ch_rows = ch_rows;
                P.nnz_cap = (int)round_up((long long)ch_rows * mrn + 8, 4);
            } else {
                P.op_kind = OP_CSR_WARP;
                P.ch_rows = 32;
            }
        }
        if (op->kind == 1 && c.g.slice <= 1024 && h->encode_tiled) {
            // dense operator tiles through a 2-D tensor map: boxes of <= 256 rows x dense_cpt (<= 256) columns
            int nrb = (c.g.slice + 255) / 256;
            while (c.g.slice % nrb != 0 || (c.g.slice / nrb) % 2 != 0) ++nrb;
            const int box_rows = c.g.slice / nrb;
            int cpt = (int)std::min<size_t>(SLOT_BYTES / ((size_t)c.g.slice * 8), 256);
            typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                          const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                          CUtensorMapFloatOOBfill);
            const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)(op->ncols ? op->ncols : n)};
            const cuuint64_t gstr[1] = {(cuuint64_t)op->lda * 8};
            const cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)cpt};
            const cuuint32_t estr[2] = {1, 1};
            if (cpt >= 1 && box_rows <= 256 &&
                ((encode_fn)h->encode_tiled)(&tmapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)op->Ad, gdim, gstr, box,
                                             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
                P.dense_cpt = cpt;
                P.dense_box_rows = box_rows;
            }
        }
        // short-window instance (Lanczos / IOP-q): second slice buffer for the resident basis vector
        bool xl = false;
        if (P.op_kind == OP_CSR_STREAM && !h->no_xl && P.w_in_smem && (c.lanczos || c.iop > 0)) {
            const int ns2 = (int)((SMEM_LIMIT - fixed - 2 * wsb) / SLOT_BYTES);
            if (SMEM_LIMIT >= fixed + 2 * wsb && ns2 >= 2) {
                xl = true;
                nslot = std::min(ns2, MAXSLOT);
                P.nslot = nslot;
                wsb *= 2;
            }
        }
        P.xl = xl ? 1 : 0;
        if (xl) {
            const size_t pbytes = (size_t)2 * LLQ * CPAD * CPAD * sizeof(uint4);  // [parity][quantity][dest CTA][source rank]
            if (pbytes > h->llpkt.cap) {
                CK(h, h->llpkt.ensure(pbytes));
                CK(h, cudaMemsetAsync(h->llpkt.p, 0, h->llpkt.cap, h->stream));
                h->ll_seq = 0;
            }
            // sequence numbers are unique per team across launches; a team passes at most 2 reductions per step
            // plus one per firststep! of each of its problems
            const unsigned rounds = (unsigned)((c.nprob + nt - 1) / nt);
            const unsigned need = rounds * (2u * (unsigned)c.m + 4u) + 8u;
            if (h->ll_seq > 0xffffffffu - need - 16u) {  // wrap: clear the stale packets and start over
                CK(h, cudaMemsetAsync(h->llpkt.p, 0, h->llpkt.cap, h->stream));
                h->ll_seq = 0;
            }
            P.llpkt = h->llpkt.as<uint4>();
            if (cm) P.llloc = cm->llloc();
            if (!cm) {
                P.seq_base = h->ll_seq;
                h->ll_seq += need;
            }
        }
        // L2 policy for the operator stream (see producer_problem)
        P.l2hint = 0;
        P.hintA_cols = 1 << 30;
        if (op->kind == 0) {
            if (h->l2hint == 1) P.hintA_cols = 0;
            else if (h->l2hint < 0) {
                const double budget = 0.75 * (double)h->l2_bytes;
                const double S_A = 12.0 * (double)op->nnz + 4.0 * (double)(n + 1);
                const double per_col = 2.0 * 8.0 * (double)n * (c.nprob > 1 ? c.g.nteams : 1);
                const double cols = (budget - S_A) / per_col;
                P.hintA_cols = cols <= 0 ? 0 : (cols > 1e6 ? (1 << 30) : (int)std::ceil(cols));
            }
        }
        P.dscratch_off = (int)(fixed + wsb + (size_t)nslot * SLOT_BYTES);
        smem = fixed + wsb + (size_t)nslot * SLOT_BYTES + extra;
        if (!P.w_in_smem) CK(h, h->wglob.ensure((size_t)nt * n * 8));
        P.wglob = h->wglob.as<double>();
        {
            const bool aug = c.p > 0;
            // Arnoldi / IOP on the general instance: the fast kernel hands over to the SAFE instance (second
            // Gram-Schmidt pass) at the first step whose re-orthogonalisation test fires
#if !defined(B200K_NO_DGKS) && !defined(B200K_NO_SAFE_LAUNCH)  // (A/B builds only)
            if (!xl && !c.lanczos) {
                switch (P.op_kind) {
                    case OP_CSR_STREAM: kern_safe = aug ? (const void *)krylov_tma_kernel<OP_CSR_STREAM, true, false, 8, true> : (const void *)krylov_tma_kernel<OP_CSR_STREAM, false, false, 8, true>; break;
                    case OP_CSR_WARP: kern_safe = aug ? (const void *)krylov_tma_kernel<OP_CSR_WARP, true, false, 8, true> : (const void *)krylov_tma_kernel<OP_CSR_WARP, false, false, 8, true>; break;
                    default: kern_safe = aug ? (const void *)krylov_tma_kernel<OP_DENSE, true, false, 8, true> : (const void *)krylov_tma_kernel<OP_DENSE, false, false, 8, true>; break;
                }
            }
#endif
            switch (P.op_kind) {
                case OP_CSR_STREAM:
                    if (xl && c.lanczos && (h->no_lz1 == 2 || (h->no_lz1 == 0 && cm))) {  // one-reduction Lanczos step (row-sharded)
                        if (op->max_row_nnz <= 5) kern = aug ? (const void *)krylov_tma_kernel<OP_CSR_STREAM, true, true, 5, false, true> : (const void *)krylov_tma_kernel<OP_CSR_STREAM, false, true, 5, false, true>;
                        else kern = aug ? (const void *)krylov_tma_kernel<OP_CSR_STREAM, true, true, 8, false, true> : (const void *)krylov_tma_kernel<OP_CSR_STREAM, false, true, 8, false, true>;
                        lz1 = true;
                    } else if (xl && op->max_row_nnz <= 5) kern = aug ? (const void *)krylov_tma_kernel<OP_CSR_STREAM, true, true, 5> : (const void *)krylov_tma_kernel<OP_CSR_STREAM, false, true, 5>;
                    else if (xl) kern = aug ? (const void *)krylov_tma_kernel<OP_CSR_STREAM, true, true, 8> : (const void *)krylov_tma_kernel<OP_CSR_STREAM, false, true, 8>;
                    else kern = aug ? (const void *)krylov_tma_kernel<OP_CSR_STREAM, true, false> : (const void *)krylov_tma_kernel<OP_CSR_STREAM, false, false>;
                    break;
                case OP_CSR_WARP: kern = aug ? (const void *)krylov_tma_kernel<OP_CSR_WARP, true, false> : (const void *)krylov_tma_kernel<OP_CSR_WARP, false, false>; break;
                default: kern = aug ? (const void *)krylov_tma_kernel<OP_DENSE, true, false> : (const void *)krylov_tma_kernel<OP_DENSE, false, false>; break;
            }
        }
        threads = NT2;
        h->last_kernel = lz1 ? 6 : (xl ? 4 : 2);
        h->last_xl = xl ? 1 : 0;
    } else {
        h->last_kernel = 1;
    }

    if (h->timing) CK(h, cudaEventRecord(h->ev[0], h->stream));
    void *args[] = {(void *)&P, (void *)&tmapA};
    CK(h, cudaLaunchCooperativeKernel(kern, dim3(c.g.C * c.g.nteams), dim3(threads), args, smem, h->stream));
    h->launches += 1;
    if (kern_safe && cm) {
        // Row-sharded: same hand-over; barrier target / sequence number travel in the communicator buffer on the device
        KrylovParams P2 = P;
        P2.j0_from_stat = 1;
        void *args2[] = {(void *)&P2, (void *)&tmapA};
        CK(h, cudaLaunchCooperativeKernel(kern_safe, dim3(c.g.C * c.g.nteams), dim3(threads), args2, smem, h->stream));
        h->launches += 1;
    } else if (kern_safe) {
        // One GPU: the SAFE instance is always queued behind the fast one and reads per problem where to resume
        // (nowhere, normally: 148 CTAs read one word and exit -- a few microseconds, no host round trip).
        KrylovParams P2 = P;
        P2.j0_from_stat = 1;
        CK(h, cudaMemsetAsync(P.bar, 0, (size_t)nt * 4, h->stream));
        void *args2[] = {(void *)&P2, (void *)&tmapA};
        CK(h, cudaLaunchCooperativeKernel(kern_safe, dim3(c.g.C * c.g.nteams), dim3(threads), args2, smem, h->stream));
        h->launches += 1;
    }
    if (h->timing) CK(h, cudaEventRecord(h->ev[1], h->stream));
    if (h->timing) { h->ev_k = true; h->ev_p = false; }
    return B200K_OK;
}

// ---- lock-step multi-vector Lanczos for batches (krylov_kernel_mv.cuh) ---------------------------------------------
// Cost model (cycles per step of one team, measured on the XL instance): a fixed part for the two packet all-reduces
// plus a per-row part; the multi-vector step does four problems for ~5 cycles per row instead of 3 for one.
struct BatchPlan {
    Geom g;
    double cost = -1.0;  // cycles per Krylov step for the whole batch, < 0: not possible
    int nslot = 0;
};

BatchPlan plan_batch(b200k_context *h, long long n, int nb, int per_group, double fixed, double per_row, int bytes_per_row) {
    BatchPlan best;
    const int ngroups = (nb + per_group - 1) / per_group;
    for (int C = 1; C <= h->max_ctas; ++C) {
        const int nteams = std::min(h->max_ctas / C, ngroups);
        if (nteams < 1) break;
        Geom g = make_geom(n, C, nteams);
        if ((long long)g.slice * (C - 1) >= n && C > 1) continue;  // empty trailing CTAs
        const size_t bufs = 2 * (size_t)round_up((long long)g.slice * bytes_per_row, 128);
        if (sizeof(SmemTma) + bufs + 2 * (size_t)SLOT_BYTES > SMEM_LIMIT) continue;
        const int rounds = (ngroups + nteams - 1) / nteams;
        const double cost = (double)rounds * (fixed + 40.0 * C + per_row * g.slice);
        if (best.cost < 0 || cost < best.cost) {
            best.cost = cost;
            best.g = g;
            best.nslot = (int)std::min<size_t>((SMEM_LIMIT - sizeof(SmemTma) - bufs) / SLOT_BYTES, MAXSLOT);
        }
    }
    return best;
}

int launch_krylov_mv(b200k_context *h, const KrylovCall &c, const BatchPlan &plan) {
    b200k_operator *op = c.op;
    const long long n = op->n;
    KrylovParams P;
    std::memset(&P, 0, sizeof(P));
    P.n = (int)n;
    P.rowptr = op->rowptr.as<int>();
    P.colind = op->colind.as<int>();
    P.val = op->val.as<double>();
    P.op_kind = OP_CSR_STREAM;
    const int mrn = std::max(op->max_row_nnz, 1);
    int ch_rows = (int)((SLOT_BYTES - 12 * 8 - 16) / (12LL * mrn + 4));
    ch_rows = std::min(ch_rows, NTC) / 32 * 32;
    if (ch_rows < 64) return fail(h, B200K_EUNSUPPORTED, "multi-vector kernel needs short rows");
    P.ch_rows = ch_rows;
    P.nnz_cap = (int)round_up((long long)ch_rows * mrn + 8, 4);
    P.team_size = plan.g.C;
    P.nteams = plan.g.nteams;
    P.nprob = c.nprob;
    P.slice = plan.g.slice;
    P.b = c.b;
    P.b_stride = c.b_stride;
    P.V = c.V;
    P.ldv = c.ldv;
    P.V_stride = c.V_stride;
    P.m = c.m;
    P.lanczos = 1;
    P.tol = c.tol;
    P.w_in_smem = 1;
    P.nslot = plan.nslot;
    P.tile_rows = 16;
    const int ldhd = c.m + 1;
    P.ldh = ldhd;
    P.H_stride = (long long)ldhd * (c.m + 1);
    P.xlen = round_up(n, 16);
    P.nranks = 1;
    P.cpad = CPAD;
    const int nt = plan.g.nteams;
    CK(h, h->xbuf.ensure((size_t)nt * 2 * P.xlen * KV * 8));
    CK(h, h->Hd.ensure((size_t)c.nprob * P.H_stride * 8));
    CK(h, h->scal.ensure((size_t)c.nprob * 4 * 8));
    CK(h, h->stat.ensure((size_t)c.nprob * 4 * 4));
    P.peer_xbuf[0] = h->xbuf.as<double>();
    P.Hd = h->Hd.as<double>();
    P.scal = h->scal.as<double>();
    P.stat = h->stat.as<int>();
    CK(h, cudaMemsetAsync(P.Hd, 0, (size_t)c.nprob * P.H_stride * 8, h->stream));
    CK(h, cudaMemsetAsync(P.scal, 0, (size_t)c.nprob * 4 * 8, h->stream));
    const size_t pbytes = (size_t)2 * LLQ * CPAD * CPAD * sizeof(uint4);
    if (pbytes > h->llpkt.cap) {
        CK(h, h->llpkt.ensure(pbytes));
        CK(h, cudaMemsetAsync(h->llpkt.p, 0, h->llpkt.cap, h->stream));
        h->ll_seq = 0;
    }
    const int ngroups = (c.nprob + KV - 1) / KV;
    const unsigned rounds = (unsigned)((ngroups + nt - 1) / nt);
    const unsigned need = rounds * (2u * (unsigned)c.m + 4u) + 8u;
    if (h->ll_seq > 0xffffffffu - need - 16u) {
        CK(h, cudaMemsetAsync(h->llpkt.p, 0, h->llpkt.cap, h->stream));
        h->ll_seq = 0;
    }
    P.llpkt = h->llpkt.as<uint4>();
    P.seq_base = h->ll_seq;
    h->ll_seq += need;
    const size_t smem = sizeof(SmemTma) + 2 * (size_t)round_up((long long)P.slice * KV * 8, 128) + (size_t)P.nslot * SLOT_BYTES;
    const void *kern = mrn <= 5 ? (const void *)krylov_mv_kernel<5> : (const void *)krylov_mv_kernel<8>;
    if (h->timing) CK(h, cudaEventRecord(h->ev[0], h->stream));
    void *args[] = {(void *)&P};
    CK(h, cudaLaunchCooperativeKernel(kern, dim3(plan.g.C * plan.g.nteams), dim3(NT2), args, smem, h->stream));
    if (h->timing) CK(h, cudaEventRecord(h->ev[1], h->stream));
    if (h->timing) { h->ev_k = true; h->ev_p = false; }
    h->launches += 1;
    h->last_kernel = 5;
    h->last_xl = 0;
    return B200K_OK;
}

// Copy H / beta / (m, breakdown) of nprob problems to pinned host memory and wait.
int fetch_krylov(b200k_context *h, int nprob, int m) {
    const size_t hbytes = (size_t)nprob * (m + 1) * (m + 1) * 8;
    CK(h, h->Hh.ensure(hbytes));
    CK(h, h->scalh.ensure((size_t)nprob * 4 * 8));
    CK(h, h->stath.ensure((size_t)nprob * 4 * 4));
    CK(h, cudaMemcpyAsync(h->Hh.p, h->Hd.p, hbytes, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(h->scalh.p, h->scal.p, (size_t)nprob * 4 * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(h->stath.p, h->stat.p, (size_t)nprob * 4 * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->timing && h->ev_k) cudaEventElapsedTime(&h->krylov_ms, h->ev[0], h->ev[1]);
    return B200K_OK;
}

// Small dense phase of expv! on the host H (krylov_phiv.jl:223-244): y = exp(t H[1:m,1:m]) e1.
// Thread-safe core (own workspace): 0 ok, 1 eigensolver failure, 2 singular Pade denominator.
int expv_small_ws(double t, const double *H, int ldh, int m, double *y, smallmat::ExpWork &work) {
    if (smallmat::is_exactly_symmetric(m, H, ldh)) return smallmat::exp_symtridiag_e1(m, H, ldh, t, y) ? 0 : 1;
    std::vector<double> Hc((size_t)m * m);
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) Hc[(size_t)j * m + i] = t * H[(size_t)j * ldh + i];
    if (smallmat::expm_higham2005base(m, Hc.data(), work)) return 2;
    for (int i = 0; i < m; ++i) y[i] = Hc[i];
    return 0;
}

int expv_small(b200k_context *h, double t, const double *H, int ldh, int m, double *y) {
    const int st = expv_small_ws(t, H, ldh, m, y, h->expwork);
    if (st == 1) return fail(h, B200K_ESINGULAR, "symmetric tridiagonal eigensolver did not converge");
    if (st == 2) return fail(h, B200K_ESINGULAR, "SingularException(0): Pade denominator is singular");
    return B200K_OK;
}

// W (nrows x nc) = beta * V[:, 0:m] * Y (+ corr[c] * V[:, m]); Yhost is m x nc with leading dimension ldy.
int launch_project(b200k_context *h, const double *V, long long ldv, long long nrows, int m, double beta,
                   const double *Yhost, int ldy, int nc, double *W, long long ldw, const double *corr_host) {
    if (m > PROJ_MAXM) return fail(h, B200K_EUNSUPPORTED, "projection supports m <= 256");
    CK(h, h->Y.ensure((size_t)m * nc * 8));
    double *yh = reinterpret_cast<double *>(stage_acquire(h, (size_t)m * nc * 8 + (size_t)nc * 8));
    if (!yh) return fail(h, B200K_ENOMEM, "pinned staging buffer");
    for (int c = 0; c < nc; ++c)
        for (int i = 0; i < m; ++i) yh[(size_t)c * m + i] = Yhost[(size_t)c * ldy + i];
    CK(h, cudaMemcpyAsync(h->Y.p, yh, (size_t)m * nc * 8, cudaMemcpyHostToDevice, h->stream));
    ProjectParams P;
    std::memset(&P, 0, sizeof(P));
    if (corr_host) {
        CK(h, h->corr.ensure((size_t)nc * 8));
        double *ch = yh + (size_t)m * nc;
        for (int c = 0; c < nc; ++c) ch[c] = corr_host[c];
        CK(h, cudaMemcpyAsync(h->corr.p, ch, (size_t)nc * 8, cudaMemcpyHostToDevice, h->stream));
        P.corr = h->corr.as<double>();
    }
    CK(h, stage_commit(h));
    P.V = V;
    P.ldv = ldv;
    P.nrows = nrows;
    P.Y = h->Y.as<double>();
    P.ldy = m;
    P.m = m;
    P.beta = beta;
    P.nc = nc;
    P.W = W;
    P.ldw = ldw;
    P.vec2 = (nrows % 2 == 0) && (ldv % 2 == 0) && (ldw % 2 == 0) && aligned16(V) && aligned16(W);
    const long long units = P.vec2 ? nrows / 2 : nrows;
    int gx = (int)std::min<long long>((units + PROJ_NT - 1) / PROJ_NT, (long long)h->sm_count * 16);
    if (gx < 1) gx = 1;
    dim3 grid(gx, nc == 1 ? 1 : (nc + PROJ_NC - 1) / PROJ_NC, 1);
    if (h->timing) CK(h, cudaEventRecord(h->ev[2], h->stream));
    launch_project_kernel(P, grid, h->stream);
    CK(h, cudaGetLastError());
    if (h->timing) CK(h, cudaEventRecord(h->ev[3], h->stream));
    if (h->timing) h->ev_p = true;
    h->launches += 1;
    return B200K_OK;
}

// b_aug of the augmented firststep! (arnoldi.jl:259-266)
void make_btail(int p, double t, double mu, double *out) {
    for (int k = 1; k <= p; ++k) {
        if (k == p) {
            out[k - 1] = mu;
        } else {
            const int i = p - k;
            double fact = 1.0;
            for (int q = 2; q <= i; ++q) fact *= q;
            out[k - 1] = std::pow(t, i) / fact * mu;
        }
    }
}

// Core of arnoldi!/lanczos! shared by b200k_arnoldi, the one-shots and kiops.
int arnoldi_core(b200k_context *h, b200k_operator *op, const double *b, const b200k_krylov_opts *o, double *V,
                 long long ldv, int maxiter, double *H, int ldh, double *beta, int *m_out, int *breakdown) {
    if (op->is_complex) return fail(h, B200K_EARG, "complex operator: use the b200k_*_z entry points");
    const long long n = op->n;
    const int p = o->p;
    const int m = o->m;
    const int aug = p != 0 ? 1 : 0;
    if (m < 1) return fail(h, B200K_EARG, "m must be >= 1");
    if (m > maxiter) return fail(h, B200K_EDIM, "m exceeds Ks.maxiter: resize the KrylovSubspace first");
    if (m >= MAXCOL) return fail(h, B200K_EUNSUPPORTED, "Krylov dimension m must be < 256");
    if (p < 0 || p > MAXP) return fail(h, B200K_EUNSUPPORTED, "augmented rows p must be <= 16");
    if (ldv < n + p) return fail(h, B200K_EDIM, "DimensionMismatch: size(V,1) - p != size(A,1)");
    if (ldh < m + 1) return fail(h, B200K_EDIM, "H has fewer than m+1 rows");
    if (o->init < 0 || o->init > m) return fail(h, B200K_EARG, "init must be in 0..m");
    if (p > 0 && (o->B == nullptr || o->ldb < n)) return fail(h, B200K_EARG, "augmented operator needs B (n x p)");
    int herm = o->hermitian;
    if (herm < 0) herm = op->is_herm;
    *breakdown = 0;
    *m_out = m;
    if (o->init == 0) {  // fill!(H, 0) on the getH view (arnoldi.jl:232)
        for (int j = 0; j < m + aug; ++j)
            for (int i = 0; i < m + 1; ++i) H[(size_t)j * ldh + i] = 0.0;
    } else if (*beta == 0.0) {
        return B200K_OK;  // iszero(Ks.beta) && return Ks (arnoldi.jl:366)
    }
    KrylovCall c;
    c.op = op;
    c.b = b;
    c.b_stride = 0;
    c.nprob = 1;
    c.V = V;
    c.ldv = ldv;
    c.V_stride = 0;
    c.m = m;
    c.lanczos = herm ? 1 : 0;
    // lanczos! ignores `init` for its loop but still skips firststep! when init != 0 (arnoldi.jl:468-480)
    c.j0 = herm ? (o->init == 0 ? 0 : 1) : o->init;
    c.iop = o->iop;
    c.tol = o->tol;
    c.p = p;
    c.B = o->B;
    c.ldb = o->ldb;
    double btail[MAXP];
    if (p > 0) make_btail(p, o->t, o->mu, btail);
    c.btail_host = btail;
    c.g = single_geom(h, n);
    if (op->comm && c.g.C < 2 * LLQ)  // (the owner CTAs of the packet all-reduce; every GPU may use its own team size)
        return fail(h, B200K_EUNSUPPORTED, "row-sharded operators need at least 128 rows on every rank");
    int st = launch_krylov(h, c);
    if (st) return st;
    st = fetch_krylov(h, 1, m);
    if (st) return st;
    const double *Hh = h->Hh.as<double>();
    const int ldhd = m + 1;
    if (o->init == 0) *beta = h->scalh.as<double>()[0];
    if (*beta == 0.0) return B200K_OK;
    *m_out = h->stath.as<int>()[0];
    *breakdown = h->stath.as<int>()[1];
    const int jc0 = c.j0 == 0 ? 0 : c.j0 - 1;
    for (int jc = jc0; jc < *m_out; ++jc)
        for (int i = 0; i <= jc + 1; ++i) H[(size_t)jc * ldh + i] = Hh[(size_t)jc * ldhd + i];
    if (herm) {  // copyto!(@diagview(H, 1), v[1:end-1]) (arnoldi.jl:488)
        for (int i = 0; i + 1 < m; ++i) H[(size_t)(i + 1) * ldh + i] = H[(size_t)i * ldh + i + 1];
    }
    return B200K_OK;
}

int expv_ks_core(b200k_context *h, double t, const double *V, long long ldv, long long nrows, const double *H,
                 int ldh, int m, double beta, double *w) {
    if (m < 1) return fail(h, B200K_EARG, "m must be >= 1");
    if (ldv < nrows) return fail(h, B200K_EDIM, "Dimension mismatch");
    if (beta == 0.0) {  // w .= 0
        CK(h, cudaMemsetAsync(w, 0, (size_t)nrows * 8, h->stream));
        return B200K_OK;
    }
    std::vector<double> y(m);
    int st = expv_small(h, t, H, ldh, m, y.data());
    if (st) return st;
    return launch_project(h, V, ldv, nrows, m, beta, y.data(), m, 1, w, nrows, nullptr);
}

int phiv_ks_core(b200k_context *h, double t, const double *V, long long ldv, long long nrows, const double *H,
                 int ldh, int m, double beta, int k, int correct, double *W, long long ldw, double *errest) {
    if (m < 1 || k < 1) return fail(h, B200K_EARG, "m >= 1 and k >= 1 required");
    if (ldv < nrows || ldw < nrows) return fail(h, B200K_EDIM, "Dimension mismatch");
    std::vector<double> Hc((size_t)m * m), e(m, 0.0), C2((size_t)m * (k + 1));
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) Hc[(size_t)j * m + i] = t * H[(size_t)j * ldh + i];
    e[0] = 1.0;
    const int st = smallmat::phiv_dense(m, Hc.data(), m, e.data(), k, C2.data(), m, h->expwork);
    if (st) return fail(h, B200K_ESINGULAR, "SingularException(0): Pade denominator is singular");
    const double hlast = H[(size_t)(m - 1) * ldh + m];  // H[end, end] of the (m+1) x m view
    std::vector<double> corr(k + 1, 0.0);
    if (correct) {
        const double betah = beta * hlast * t;
        for (int i = 1; i <= k; ++i) corr[i - 1] = betah * C2[(size_t)i * m + (m - 1)];
    }
    if (errest) *errest = std::fabs(beta * hlast * t * C2[(size_t)k * m + (m - 1)]);
    if (beta == 0.0) {
        // the reference multiplies an uninitialised V by zero here; we return zeros
        for (int c = 0; c <= k; ++c) CK(h, cudaMemsetAsync(W + (size_t)c * ldw, 0, (size_t)nrows * 8, h->stream));
        return B200K_OK;
    }
    return launch_project(h, V, ldv, nrows, m, beta, C2.data(), m, k + 1, W, ldw, correct ? corr.data() : nullptr);
}

// Device-side small dense phase + projection for nprob problems whose factorisation (H, beta, m) is still in
// h->Hd / h->scal / h->stat on the device: no host round trip.  t_host: nprob times.
int launch_smallexp_project(b200k_context *h, int nprob, int m, int lanczos, const double *t_host, const double *V,
                            long long ldv, long long vstride, long long n, double *W, long long ldw,
                            long long wstride) {
    CK(h, h->Y.ensure((size_t)nprob * m * 8));
    CK(h, h->betavec.ensure((size_t)nprob * 8));
    CK(h, h->mvec.ensure((size_t)nprob * 4));
    CK(h, h->tdev.ensure((size_t)nprob * 8));
    CK(h, h->errdev.ensure(64));
    CK(h, h->errh.ensure(64));
    // a single time goes by value; a batch of times is copied from the caller's (pageable) array, which the
    // runtime stages before returning -- no pinned buffer of ours may be rewritten while a copy is still queued
    if (nprob > 1)
        CK(h, cudaMemcpyAsync(h->tdev.p, t_host, (size_t)nprob * 8, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemsetAsync(h->errdev.p, 0, 4, h->stream));
    SmallExpParams S;
    std::memset(&S, 0, sizeof(S));
    S.Hd = h->Hd.as<double>();
    S.ldh = m + 1;
    S.H_stride = (long long)(m + 1) * (m + 1);
    S.scal = h->scal.as<double>();
    S.stat = h->stat.as<int>();
    S.tvec = nprob > 1 ? h->tdev.as<double>() : nullptr;
    S.t = t_host[0];
    S.m = m;
    S.lanczos = lanczos;
    S.force_pade = h->sym_pade;
    S.Y = h->Y.as<double>();
    S.ldy = m;
    S.betavec = h->betavec.as<double>();
    S.mvec = h->mvec.as<int>();
    S.err = h->errdev.as<int>();
    if (h->timing) CK(h, cudaEventRecord(h->ev[2], h->stream));
    small_exp_kernel<<<nprob, SE_NT, (size_t)6 * m * m * 8, h->stream>>>(S);
    CK(h, cudaGetLastError());
    h->launches += 1;
    ProjectParams P;
    std::memset(&P, 0, sizeof(P));
    P.V = V;
    P.ldv = ldv;
    P.V_stride = vstride;
    P.nrows = n;
    P.Y = h->Y.as<double>();
    P.ldy = m;
    P.Y_stride = m;
    P.mvec = h->mvec.as<int>();
    P.betavec = h->betavec.as<double>();
    P.m = m;
    P.nc = 1;
    P.W = W;
    P.ldw = ldw;
    P.W_stride = wstride;
    P.vec2 = (n % 2 == 0) && (ldv % 2 == 0) && (vstride % 2 == 0) && (ldw % 2 == 0) && (wstride % 2 == 0) &&
             aligned16(V) && aligned16(W);
    const long long units = P.vec2 ? n / 2 : n;
    const long long want = (units + PROJ_NT - 1) / PROJ_NT;
    int gx = (int)std::min<long long>(want, nprob > 1 ? 64 : (long long)h->sm_count * 16);
    if (gx < 1) gx = 1;
    for (int base = 0; base < nprob; base += 32768) {  // gridDim.z limit
        const int cnt = std::min(nprob - base, 32768);
        ProjectParams Q = P;
        Q.V += (long long)base * vstride;
        Q.Y += (long long)base * m;
        Q.mvec += base;
        Q.betavec += base;
        Q.W += (long long)base * wstride;
        launch_project_kernel(Q, dim3(gx, 1, cnt), h->stream);
        h->launches += 1;
    }
    CK(h, cudaGetLastError());
    if (h->timing) CK(h, cudaEventRecord(h->ev[3], h->stream));
    if (h->timing) h->ev_p = true;
    // the (practically impossible) singular Pade denominator is reported by the next synchronising call
    CK(h, cudaMemcpyAsync(h->errh.p, h->errdev.p, 4, cudaMemcpyDeviceToHost, h->stream));
    return B200K_OK;
}

int ensure_internal_ks(b200k_context *h, long long nrows, int m, long long *ldv) {
    *ldv = round_up(nrows, 16);
    CK(h, h->V.ensure((size_t)(*ldv) * (m + 1) * 8));
    h->H.assign((size_t)(m + 2) * (m + 2), 0.0);
    return B200K_OK;
}

}  // namespace

// ====================================================================================================
extern "C" {

int b200k_version(void) { return B200K_VERSION; }

int b200k_sizeof(int which) {
    switch (which) {
        case B200K_STRUCT_KRYLOV_OPTS: return (int)sizeof(b200k_krylov_opts);
        case B200K_STRUCT_KIOPS_OPTS: return (int)sizeof(b200k_kiops_opts);
        case B200K_STRUCT_TIMESTEP_OPTS: return (int)sizeof(b200k_timestep_opts);
        default: return -1;
    }
}

const char *b200k_status_string(int s) {
    switch (s) {
        case B200K_OK: return "ok";
        case B200K_EDIM: return "dimension mismatch";
        case B200K_EARG: return "invalid argument";
        case B200K_ESINGULAR: return "singular matrix in the small dense phase";
        case B200K_ECUDA: return "CUDA error";
        case B200K_ECOMM: return "communicator error";
        case B200K_EUNSUPPORTED: return "unsupported configuration";
        case B200K_ENOMEM: return "out of memory";
        default: return "unknown status";
    }
}

void b200k_krylov_opts_default(b200k_krylov_opts *o) {
    std::memset(o, 0, sizeof(*o));
    o->m = 30;
    o->tol = 1.0e-7;
    o->iop = 0;
    o->hermitian = -1;
    o->init = 0;
    o->p = 0;
    o->B = nullptr;
    o->ldb = 0;
    o->t = NAN;
    o->mu = NAN;
}

void b200k_kiops_opts_default(b200k_kiops_opts *o) {
    o->mmin = 10;
    o->mmax = 128;
    o->m = 10;
    o->tol = 1.0e-7;
    o->iop = 2;
    o->hermitian = -1;
    o->task1 = 0;
    o->opnorm = NAN;
    o->normU = NAN;
}

int b200k_create(b200k_handle_t *out, int device, void *stream) {
    if (!out) return B200K_EARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return B200K_ECUDA;  // no CPU fallback
    if (device < 0 || device >= ndev) return B200K_EARG;
    if (cudaSetDevice(device) != cudaSuccess) return B200K_ECUDA;
    b200k_context *h = new b200k_context();
    h->device = device;
    h->stream = (cudaStream_t)stream;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete h;
        return B200K_ECUDA;
    }
    h->sm_count = prop.multiProcessorCount;
    h->l2_bytes = prop.l2CacheSize;
    if (!prop.cooperativeLaunch ||
        cudaFuncSetAttribute((const void *)krylov_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)SMEM_LIMIT) != cudaSuccess) {
        delete h;
        return B200K_ECUDA;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, krylov_persistent_kernel, NT, SMEM_LIMIT) !=
            cudaSuccess ||
        per_sm < 1) {
        delete h;
        return B200K_ECUDA;
    }
    const void *tma_kernels[] = {(const void *)krylov_tma_kernel<OP_CSR_STREAM, false, false>, (const void *)krylov_tma_kernel<OP_CSR_STREAM, true, false>,
                                 (const void *)krylov_tma_kernel<OP_CSR_STREAM, false, true, 8>, (const void *)krylov_tma_kernel<OP_CSR_STREAM, true, true, 8>,
                                 (const void *)krylov_tma_kernel<OP_CSR_STREAM, false, true, 5>, (const void *)krylov_tma_kernel<OP_CSR_STREAM, true, true, 5>,
                                 (const void *)krylov_tma_kernel<OP_CSR_WARP, false, false>, (const void *)krylov_tma_kernel<OP_CSR_WARP, true, false>,
                                 (const void *)krylov_tma_kernel<OP_DENSE, false, false>, (const void *)krylov_tma_kernel<OP_DENSE, true, false>,
                                 (const void *)krylov_tma_kernel<OP_CSR_STREAM, false, true, 8, false, true>, (const void *)krylov_tma_kernel<OP_CSR_STREAM, true, true, 8, false, true>,
                                 (const void *)krylov_tma_kernel<OP_CSR_STREAM, false, true, 5, false, true>, (const void *)krylov_tma_kernel<OP_CSR_STREAM, true, true, 5, false, true>,
                                 (const void *)krylov_tma_kernel<OP_CSR_STREAM, false, false, 8, true>, (const void *)krylov_tma_kernel<OP_CSR_STREAM, true, false, 8, true>,
                                 (const void *)krylov_tma_kernel<OP_CSR_WARP, false, false, 8, true>, (const void *)krylov_tma_kernel<OP_CSR_WARP, true, false, 8, true>,
                                 (const void *)krylov_tma_kernel<OP_DENSE, false, false, 8, true>, (const void *)krylov_tma_kernel<OP_DENSE, true, false, 8, true>};
    cudaFuncSetAttribute((const void *)krylov_mv_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    cudaFuncSetAttribute((const void *)krylov_mv_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    bool attr_ok = true;
    for (const void *k : tma_kernels)
        attr_ok = attr_ok && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT) == cudaSuccess;
    if (!attr_ok ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, krylov_tma_kernel<OP_CSR_STREAM, false, false>, NT2, SMEM_LIMIT) != cudaSuccess ||
        per_sm < 1) {
        delete h;
        return B200K_ECUDA;
    }
    {
        cudaDriverEntryPointQueryResult qres;
        void *fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            h->encode_tiled = fn;
    }
    if (const char *env = std::getenv("B200K_SMALLEXP")) h->host_smallexp = std::strcmp(env, "host") == 0 ? 1 : 0;
    cudaFuncSetAttribute((const void *)krylov_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    cudaFuncSetAttribute((const void *)krylov_tma_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    cudaFuncSetAttribute((const void *)small_exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         6 * SE_MAXM * SE_MAXM * 8);
    cudaFuncSetAttribute((const void *)small_exp_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         6 * SE_MAXM * SE_MAXM * 8);
    if (const char *env = std::getenv("B200K_MV")) h->no_mv = std::strcmp(env, "0") == 0 ? 1 : (std::strcmp(env, "2") == 0 ? 2 : 0);
    if (const char *env = std::getenv("B200K_XL")) h->no_xl = std::strcmp(env, "0") == 0 ? 1 : 0;
    if (const char *env = std::getenv("B200K_KERNEL")) h->force_ldg = std::strcmp(env, "ldg") == 0 ? 1 : 0;
    h->max_ctas = std::min(h->sm_count, CPAD);  // one CTA per SM
    for (int i = 0; i < 4; ++i) cudaEventCreate(&h->ev[i]);
    *out = h;
    return B200K_OK;
}

int b200k_destroy(b200k_handle_t h) {
    if (!h) return B200K_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf *bufs[] = {&h->zV, &h->zH, &h->zpart, &h->zxbuf, &h->zw, &h->zy, &h->tsV, &h->tsW, &h->tsP, &h->tsu, &h->tdev, &h->errdev, &h->llpkt, &h->kV, &h->kB, &h->xbuf, &h->part, &h->partn, &h->bar, &h->wglob, &h->Hd, &h->scal, &h->stat, &h->btail,
                      &h->Y, &h->corr, &h->mvec, &h->betavec, &h->tmp, &h->V, &h->bdev, &h->wdev};
    for (DevBuf *b : bufs) b->release();
    HostBuf *hb[] = {&h->zHh, &h->errh, &h->Hh, &h->scalh, &h->stath};
    for (HostBuf *b : hb) b->release();
    for (int i = 0; i < b200k_context::NSTAGE; ++i) {
        h->stage[i].release();
        if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
    }
    for (int i = 0; i < 4; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    delete h;
    return B200K_OK;
}

int b200k_set_stream(b200k_handle_t h, void *stream) {
    if (!h) return B200K_EARG;
    h->stream = (cudaStream_t)stream;
    return B200K_OK;
}

int b200k_synchronize(b200k_handle_t h) {
    if (!h) return B200K_EARG;
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->errh.p && *h->errh.as<int>()) {  // deferred report of the device-side small exponential
        *h->errh.as<int>() = 0;
        return fail(h, B200K_ESINGULAR, "SingularException(0): Pade denominator is singular");
    }
    return B200K_OK;
}

const char *b200k_last_error(b200k_handle_t h) { return h ? h->err.c_str() : "null handle"; }

int b200k_device_info(b200k_handle_t h, int *sm_count, int *max_team, int64_t *launches) {
    if (!h) return B200K_EARG;
    if (sm_count) *sm_count = h->sm_count;
    if (max_team) *max_team = h->max_ctas;
    if (launches) *launches = h->launches;
    return B200K_OK;
}

int b200k_set_timing(b200k_handle_t h, int enabled) {
    if (!h) return B200K_EARG;
    h->timing = enabled ? 1 : 0;
    return B200K_OK;
}

int b200k_set_flag(b200k_handle_t h, int flag, int value) {
    if (!h) return B200K_EARG;
    if (flag == B200K_FLAG_FORCE_LDG) h->force_ldg = value ? 1 : 0;
    else if (flag == B200K_FLAG_HOST_SMALLEXP) h->host_smallexp = value ? 1 : 0;
    else if (flag == B200K_FLAG_NO_XL) h->no_xl = value ? 1 : 0;
    else if (flag == B200K_FLAG_NO_MV) h->no_mv = value == 2 ? 2 : (value ? 1 : 0);
    else if (flag == B200K_FLAG_L2HINT) h->l2hint = value < 0 ? -1 : (value ? 1 : 0);
    else if (flag == B200K_FLAG_SYM_PADE) h->sym_pade = value ? 1 : 0;
    else if (flag == B200K_FLAG_NO_LZ1) h->no_lz1 = value == 2 ? 2 : (value ? 1 : 0);
    else return fail(h, B200K_EARG, "unknown flag");
    return B200K_OK;
}

int b200k_last_kernel(b200k_handle_t h, int *which) {
    if (!h || !which) return B200K_EARG;
    *which = h->last_kernel;
    return B200K_OK;
}

int b200k_last_timing(b200k_handle_t h, float *krylov_ms, float *project_ms) {
    if (!h) return B200K_EARG;
    if (h->timing && h->ev_k) {
        cudaEventSynchronize(h->ev[1]);
        cudaEventElapsedTime(&h->krylov_ms, h->ev[0], h->ev[1]);
        h->project_ms = 0.f;
        if (h->ev_p) {
            cudaEventSynchronize(h->ev[3]);
            cudaEventElapsedTime(&h->project_ms, h->ev[2], h->ev[3]);
        }
    }
    if (krylov_ms) *krylov_ms = h->krylov_ms;
    if (project_ms) *project_ms = h->project_ms;
    return B200K_OK;
}

// ---- operators ---------------------------------------------------------------------------------------
int b200k_op_csr_create(b200k_handle_t h, int64_t n, int64_t nnz, const int32_t *rowptr, const int32_t *colind,
                        const double *val, int index_base, int location, b200k_op_t *out) {
    if (!h || !out) return B200K_EARG;
    *out = nullptr;
    if (n < 1 || nnz < 0 || !rowptr || (nnz > 0 && (!colind || !val)))
        return fail(h, B200K_EARG, "invalid CSR description");
    if (n > 2000000000LL || nnz > 2000000000LL) return fail(h, B200K_EUNSUPPORTED, "int32 CSR indices only");
    if (index_base != 0 && index_base != 1) return fail(h, B200K_EARG, "index_base must be 0 or 1");
    CK(h, cudaSetDevice(h->device));
    b200k_operator *op = new b200k_operator();
    op->ctx = h;
    op->device = h->device;
    op->kind = 0;
    op->n = n;
    op->nnz = nnz;
    const size_t pad = 64;
    cudaError_t e = op->rowptr.ensure((size_t)(n + 1 + pad) * 4);
    if (e == cudaSuccess) e = op->colind.ensure((size_t)(nnz + pad) * 4);
    if (e == cudaSuccess) e = op->val.ensure((size_t)(nnz + pad) * 8);
    if (e != cudaSuccess) {
        delete op;
        return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
    }
    const cudaMemcpyKind kind = location == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    cudaMemsetAsync(op->colind.p, 0, op->colind.cap, h->stream);
    cudaMemsetAsync(op->val.p, 0, op->val.cap, h->stream);
    cudaMemcpyAsync(op->rowptr.p, rowptr, (size_t)(n + 1) * 4, kind, h->stream);
    if (nnz > 0) {
        cudaMemcpyAsync(op->colind.p, colind, (size_t)nnz * 4, kind, h->stream);
        cudaMemcpyAsync(op->val.p, val, (size_t)nnz * 8, kind, h->stream);
    }
    if (index_base == 1) {
        rebase_kernel<<<256, 256, 0, h->stream>>>(op->rowptr.as<int>(), op->rowptr.as<int>(), n + 1, 1);
        if (nnz > 0) rebase_kernel<<<1024, 256, 0, h->stream>>>(op->colind.as<int>(), op->colind.as<int>(), nnz, 1);
    }
    // row statistics, ishermitian(A), opnorm(A, Inf)
    e = h->tmp.ensure(64);
    if (e != cudaSuccess) {
        delete op;
        return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
    }
    cudaMemsetAsync(h->tmp.p, 0, 64, h->stream);
    int *d_max = h->tmp.as<int>();
    int *d_nonsym = d_max + 1;
    double *d_norm = reinterpret_cast<double *>(h->tmp.as<char>() + 16);
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 4096);
    csr_analyze_kernel<<<blocks, 256, 0, h->stream>>>((int)n, op->rowptr.as<int>(), op->colind.as<int>(),
                                                      op->val.as<double>(), d_max, d_nonsym, d_norm);
    struct {
        int maxnnz, nonsym;
        double pad0, norm;
    } res;
    char raw[32];
    e = cudaMemcpyAsync(raw, h->tmp.p, 32, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        delete op;
        return fail(h, B200K_ECUDA, cudaGetErrorString(e));
    }
    std::memcpy(&res.maxnnz, raw, 4);
    std::memcpy(&res.nonsym, raw + 4, 4);
    std::memcpy(&res.norm, raw + 16, 8);
    op->max_row_nnz = res.maxnnz;
    op->is_herm = res.nonsym ? 0 : 1;
    op->opnorm_inf = res.norm;
    h->launches += 1;
    *out = op;
    return B200K_OK;
}

int b200k_op_dense_create(b200k_handle_t h, int64_t n, const double *A, int64_t lda, int location,
                          b200k_op_t *out) {
    if (!h || !out) return B200K_EARG;
    *out = nullptr;
    if (n < 1 || !A || lda < n) return fail(h, B200K_EARG, "invalid dense description");
    if (n > 2000000000LL) return fail(h, B200K_EUNSUPPORTED, "n too large");
    CK(h, cudaSetDevice(h->device));
    b200k_operator *op = new b200k_operator();
    op->ctx = h;
    op->device = h->device;
    op->kind = 1;
    op->n = n;
    op->nnz = n * n;
    if (location == 1) {
        const long long ld = round_up(n, 2);
        cudaError_t e = op->Aown.ensure((size_t)ld * n * 8);
        if (e != cudaSuccess) {
            delete op;
            return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
        }
        e = cudaMemcpy2DAsync(op->Aown.p, (size_t)ld * 8, A, (size_t)lda * 8, (size_t)n * 8, (size_t)n,
                              cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) {
            delete op;
            return fail(h, B200K_ECUDA, cudaGetErrorString(e));
        }
        op->Ad = op->Aown.as<double>();
        op->lda = ld;
    } else {
        op->Ad = A;  // aliased: the caller keeps the matrix alive (it is 8 n^2 bytes)
        op->lda = lda;
    }
    cudaError_t e = h->tmp.ensure(64);
    if (e != cudaSuccess) {
        delete op;
        return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
    }
    cudaMemsetAsync(h->tmp.p, 0, 64, h->stream);
    int *d_nonsym = h->tmp.as<int>() + 1;
    double *d_norm = reinterpret_cast<double *>(h->tmp.as<char>() + 16);
    dense_analyze_kernel<<<(int)std::min<int64_t>((n + 127) / 128, 2048), 128, 0, h->stream>>>(
        (int)n, op->Ad, op->lda, d_nonsym, d_norm);
    char raw[32];
    e = cudaMemcpyAsync(raw, h->tmp.p, 32, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        delete op;
        return fail(h, B200K_ECUDA, cudaGetErrorString(e));
    }
    int nonsym;
    double norm;
    std::memcpy(&nonsym, raw + 4, 4);
    std::memcpy(&norm, raw + 16, 8);
    op->is_herm = nonsym ? 0 : 1;
    op->opnorm_inf = norm;
    h->launches += 1;
    *out = op;
    return B200K_OK;
}

int b200k_op_destroy(b200k_op_t op) {
    if (!op) return B200K_OK;
    cudaSetDevice(op->device);
    op->send_row.release();
    op->send_peer.release();
    op->send_pos.release();
    op->send_ofs.release();
    op->rowptr.release();
    op->colind.release();
    op->val.release();
    op->Aown.release();
    delete op;
    return B200K_OK;
}

int b200k_op_info(b200k_op_t op, int64_t *n, int64_t *nnz, int *kind, int *is_hermitian, double *opnorm_inf) {
    if (!op) return B200K_EARG;
    if (n) *n = op->n;
    if (nnz) *nnz = op->nnz;
    if (kind) *kind = op->kind;
    if (is_hermitian) *is_hermitian = op->is_herm;
    if (opnorm_inf) *opnorm_inf = op->opnorm_inf;
    return B200K_OK;
}

int b200k_op_apply(b200k_handle_t h, b200k_op_t op, const double *x, double *y) {
    if (!h || !op || !x || !y) return B200K_EARG;
    if (op->comm) return fail(h, B200K_EUNSUPPORTED, "mul! on a row-sharded operator (the halo exchange lives in the Krylov kernel)");
    if (op->kind == 0) {
        const int blocks = (int)std::min<long long>((op->n + 7) / 8, (long long)h->sm_count * 16);
        csr_apply_kernel<<<blocks, 256, 0, h->stream>>>((int)op->n, op->rowptr.as<int>(), op->colind.as<int>(),
                                                        op->val.as<double>(), x, y);
    } else {
        const int blocks = (int)std::min<long long>((op->n + 63) / 64, (long long)h->sm_count * 8);
        dense_apply_kernel<<<blocks, 256, 0, h->stream>>>((int)op->n, op->Ad, op->lda, x, y);
    }
    CK(h, cudaGetLastError());
    h->launches += 1;
    return B200K_OK;
}

// ---- arnoldi! / expv! / phiv! --------------------------------------------------------------------------
int b200k_arnoldi(b200k_handle_t h, b200k_op_t op, const double *b, const b200k_krylov_opts *opts, double *V,
                  int64_t ldv, int maxiter, double *H, int ldh, double *beta, int *m_out, int *breakdown) {
    if (!h || !op || !b || !opts || !V || !H || !beta || !m_out || !breakdown) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    return arnoldi_core(h, op, b, opts, V, ldv, maxiter, H, ldh, beta, m_out, breakdown);
}

int b200k_expv_ks(b200k_handle_t h, double t, const double *V, int64_t ldv, int64_t nrows, const double *H,
                  int ldh, int m, double beta, double *w) {
    if (!h || !V || !H || !w) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    return expv_ks_core(h, t, V, ldv, nrows, H, ldh, m, beta, w);
}

int b200k_phiv_ks(b200k_handle_t h, double t, const double *V, int64_t ldv, int64_t nrows, const double *H,
                  int ldh, int m, double beta, int k, int correct, double *W, int64_t ldw, double *errest) {
    if (!h || !V || !H || !W) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    return phiv_ks_core(h, t, V, ldv, nrows, H, ldh, m, beta, k, correct, W, ldw, errest);
}

int b200k_expv(b200k_handle_t h, b200k_op_t op, double t, const double *b, const b200k_krylov_opts *opts,
               double *w, int *m_out, int *breakdown, double *beta_out) {
    if (!h || !op || !b || !opts || !w) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    if (opts->p != 0 || opts->init != 0) return fail(h, B200K_EARG, "one-shot expv takes a plain operator, init = 0");
    b200k_krylov_opts o = *opts;
    o.m = (int)std::min<long long>(o.m, op->n);
    long long ldv;
    int st = ensure_internal_ks(h, op->n, o.m, &ldv);
    if (st) return st;
    double beta = 0.0;
    int mo = 0, bd = 0;
    const int ldh = o.m + 2;
    if (!h->host_smallexp && o.m <= SE_MAXM && o.m >= 1 && o.m < MAXCOL) {
        // fully device-side: Krylov kernel -> small_exp_kernel -> project_kernel, no host round trip unless the
        // caller asks for m / breakdown / beta
        int herm = o.hermitian;
        if (herm < 0) herm = op->is_herm;
        KrylovCall c;
        c.op = op;
        c.b = b;
        c.b_stride = 0;
        c.nprob = 1;
        c.V = h->V.as<double>();
        c.ldv = ldv;
        c.V_stride = 0;
        c.m = o.m;
        c.j0 = 0;
        c.iop = o.iop;
        c.lanczos = herm ? 1 : 0;
        c.tol = o.tol;
        c.p = 0;
        c.B = nullptr;
        c.ldb = 0;
        c.btail_host = nullptr;
        c.g = single_geom(h, op->n);
        if (op->comm && c.g.C < 2 * LLQ)
            return fail(h, B200K_EUNSUPPORTED, "row-sharded operators need at least 128 rows on every rank");
        st = launch_krylov(h, c);
        if (st) return st;
        st = launch_smallexp_project(h, 1, o.m, c.lanczos, &t, h->V.as<double>(), ldv, 0, op->n, w, op->n, 0);
        if (st) return st;
        if (m_out || breakdown || beta_out) {
            st = fetch_krylov(h, 1, o.m);
            if (st) return st;
            beta = h->scalh.as<double>()[0];
            mo = beta == 0.0 ? o.m : h->stath.as<int>()[0];
            bd = beta == 0.0 ? 0 : h->stath.as<int>()[1];
            if (*h->errh.as<int>()) return fail(h, B200K_ESINGULAR, "SingularException(0): Pade denominator is singular");
            if (m_out) *m_out = mo;
            if (breakdown) *breakdown = bd;
            if (beta_out) *beta_out = beta;
        }
        return B200K_OK;
    }
    st = arnoldi_core(h, op, b, &o, h->V.as<double>(), ldv, o.m, h->H.data(), ldh, &beta, &mo, &bd);
    if (st) return st;
    if (m_out) *m_out = mo;
    if (breakdown) *breakdown = bd;
    if (beta_out) *beta_out = beta;
    return expv_ks_core(h, t, h->V.as<double>(), ldv, op->n, h->H.data(), ldh, mo, beta, w);
}

int b200k_expv_ee(b200k_handle_t h, b200k_op_t op, double t, const double *b, int m, double atol, double rtol,
                  double *w, int *m_out) {
    if (!h || !op || !b || !w) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    if (!op->is_herm)
        return fail(h, B200K_EUNSUPPORTED, "Error estimation not yet available for non-Hermitian matrices.");
    m = (int)std::min<long long>(m, op->n);
    if (m < 1) return fail(h, B200K_EARG, "m must be >= 1");
    long long ldv;
    int st = ensure_internal_ks(h, op->n, m, &ldv);
    if (st) return st;
    b200k_krylov_opts o;
    b200k_krylov_opts_default(&o);
    o.m = m;
    o.tol = 0.0;  // lanczos_step! is called directly: no happy-breakdown test in this mode
    o.hermitian = 1;
    double beta = 0.0;
    int mo = 0, bd = 0;
    const int ldh = m + 2;
    st = arnoldi_core(h, op, b, &o, h->V.as<double>(), ldv, m, h->H.data(), ldh, &beta, &mo, &bd);
    if (st) return st;
    if (beta == 0.0) {  // Ks.m = 0; w .= 0
        if (m_out) *m_out = 0;
        CK(h, cudaMemsetAsync(w, 0, (size_t)op->n * 8, h->stream));
        return B200K_OK;
    }
    const double eps = atol + rtol * beta;
    const double *H = h->H.data();
    int jstop = m;
    std::vector<double> d, e, zf, zl;
    for (int j = 1; j <= m; ++j) {
        d.resize(j);
        e.resize(std::max(j - 1, 0));
        for (int i = 0; i < j; ++i) d[i] = H[(size_t)i * ldh + i];
        for (int i = 0; i + 1 < j; ++i) e[i] = H[(size_t)i * ldh + i + 1];
        if (!smallmat::symtridiag_eig_firstlast(j, d, e, zf, zl))
            return fail(h, B200K_ESINGULAR, "symmetric tridiagonal eigensolver did not converge");
        double vj = 0.0;
        for (int k = 0; k < j; ++k) vj += std::exp(t * d[k]) * zf[k] * zl[k];
        const double sigma = H[(size_t)(j - 1) * ldh + j] * beta * std::fabs(vj);
        if (sigma < eps) {
            jstop = j;
            break;
        }
    }
    if (m_out) *m_out = jstop;
    std::vector<double> y(jstop);
    if (!smallmat::exp_symtridiag_e1(jstop, H, ldh, t, y.data()))
        return fail(h, B200K_ESINGULAR, "symmetric tridiagonal eigensolver did not converge");
    return launch_project(h, h->V.as<double>(), ldv, op->n, jstop, beta, y.data(), jstop, 1, w, op->n, nullptr);
}

int b200k_expv_host_async(b200k_handle_t h, b200k_op_t op, double t, const double *b_host,
                          const b200k_krylov_opts *opts, double *w_host) {
    if (!h || !op || !b_host || !opts || !w_host) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    const size_t bytes = (size_t)op->n * 8;
    CK(h, h->bdev.ensure(bytes));
    CK(h, h->wdev.ensure(bytes));
    CK(h, cudaMemcpyAsync(h->bdev.p, b_host, bytes, cudaMemcpyHostToDevice, h->stream));
    const int st = b200k_expv(h, op, t, h->bdev.as<double>(), opts, h->wdev.as<double>(), nullptr, nullptr, nullptr);
    if (st) return st;
    CK(h, cudaMemcpyAsync(w_host, h->wdev.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    return B200K_OK;
}

int b200k_expv_host(b200k_handle_t h, b200k_op_t op, double t, const double *b_host,
                    const b200k_krylov_opts *opts, double *w_host, int *m_out, int *breakdown) {
    if (!h || !op || !b_host || !opts || !w_host) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    const size_t bytes = (size_t)op->n * 8;
    CK(h, h->bdev.ensure(bytes));
    CK(h, h->wdev.ensure(bytes));
    CK(h, cudaMemcpyAsync(h->bdev.p, b_host, bytes, cudaMemcpyHostToDevice, h->stream));
    const int st = b200k_expv(h, op, t, h->bdev.as<double>(), opts, h->wdev.as<double>(), m_out, breakdown, nullptr);
    if (st) return st;
    CK(h, cudaMemcpyAsync(w_host, h->wdev.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->errh.p && *h->errh.as<int>()) return fail(h, B200K_ESINGULAR, "SingularException(0): Pade denominator is singular");
    return B200K_OK;
}

int b200k_phiv(b200k_handle_t h, b200k_op_t op, double t, const double *b, int k, const b200k_krylov_opts *opts,
               int correct, double *W, int64_t ldw, double *errest, int *m_out, int *breakdown) {
    if (!h || !op || !b || !opts || !W) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    if (opts->p != 0 || opts->init != 0) return fail(h, B200K_EARG, "one-shot phiv takes a plain operator, init = 0");
    b200k_krylov_opts o = *opts;
    o.m = (int)std::min<long long>(o.m, op->n);
    long long ldv;
    int st = ensure_internal_ks(h, op->n, o.m, &ldv);
    if (st) return st;
    double beta = 0.0;
    int mo = 0, bd = 0;
    const int ldh = o.m + 2;
    st = arnoldi_core(h, op, b, &o, h->V.as<double>(), ldv, o.m, h->H.data(), ldh, &beta, &mo, &bd);
    if (st) return st;
    if (m_out) *m_out = mo;
    if (breakdown) *breakdown = bd;
    return phiv_ks_core(h, t, h->V.as<double>(), ldv, op->n, h->H.data(), ldh, mo, beta, k, correct, W, ldw, errest);
}

// ---- batched independent expv ----------------------------------------------------------------------------
int b200k_expv_batched(b200k_handle_t h, b200k_op_t op, int nb, const double *t, const double *Bv, int64_t ldbv,
                       const b200k_krylov_opts *opts, double *W, int64_t ldw, int *m_out, int *breakdown) {
    if (!h || !op || !t || !Bv || !opts || !W || nb < 1) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    if (opts->p != 0 || opts->init != 0) return fail(h, B200K_EARG, "batched expv takes a plain operator, init = 0");
    const long long n = op->n;
    if (ldbv < n || ldw < n) return fail(h, B200K_EDIM, "Dimension mismatch");
    const int m = (int)std::min<long long>(opts->m, n);
    if (m < 1 || m >= MAXCOL) return fail(h, B200K_EUNSUPPORTED, "Krylov dimension m must be in 1..255");
    int herm = opts->hermitian;
    if (herm < 0) herm = op->is_herm;
    const long long ldv = round_up(n, 16);
    const long long vstride = ldv * (m + 1);
    CK(h, h->V.ensure((size_t)vstride * nb * 8));
    KrylovCall c;
    c.op = op;
    c.b = Bv;
    c.b_stride = ldbv;
    c.nprob = nb;
    c.V = h->V.as<double>();
    c.ldv = ldv;
    c.V_stride = vstride;
    c.m = m;
    c.j0 = 0;
    c.iop = opts->iop;
    c.lanczos = herm ? 1 : 0;
    c.tol = opts->tol;
    c.p = 0;
    c.B = nullptr;
    c.ldb = 0;
    c.btail_host = nullptr;
    c.g = batch_geom(h, n, nb, op->kind == 0 && (c.lanczos || c.iop > 0), m);
    int st = B200K_OK;
    bool mv = false;
    if (c.lanczos && op->kind == 0 && !op->comm && h->no_mv != 1 && !h->no_xl && !h->force_ldg && nb >= 2 &&
        op->max_row_nnz > 0 && (SLOT_BYTES - 12 * 8 - 16) / (12LL * op->max_row_nnz + 4) >= 64 &&
        !std::getenv("B200K_BATCH_TEAM")) {
        // Four problems per team in lock step.  Measured at C5 (profiles/r1_s2_phase_c5_mv.json): a team-step costs
        // 30.6 k cycles for four problems (mat-vec 10.0 k = 4.9 cycles per row as modelled, but 20.6 k of fixed cost:
        // update with strided gather-buffer stores 4.6 k, release fence behind 64 KB of stores 3.9 k, four fp64
        // divisions + square roots per thread 1.9 k, reductions 7.3 k) against 32.9 k for one problem on the
        // per-problem teams with four times as many teams: 20.9 k vs 22.4 k expv/s/GPU end to end.  Until that fixed
        // part is trimmed the kernel is opt-in (B200K_FLAG_NO_MV = 2 / B200K_MV=2); it is parity-tested either way.
        const BatchPlan pm = plan_batch(h, n, nb, KV, 18000.0, 5.0, KV * 8);
        if (pm.cost > 0 && h->no_mv == 2) {
            st = launch_krylov_mv(h, c, pm);
            if (st) return st;
            mv = true;
        }
    }
    if (!mv) {
        st = launch_krylov(h, c);
        if (st) return st;
    }
    if (!h->host_smallexp && m <= SE_MAXM) {
        // nb exponentials on nb SMs + one batched projection, all on the device
        st = launch_smallexp_project(h, nb, m, c.lanczos, t, h->V.as<double>(), ldv, vstride, n, W, ldw, ldw);
        if (st) return st;
        if (m_out || breakdown) {
            st = fetch_krylov(h, nb, m);
            if (st) return st;
            for (int i = 0; i < nb; ++i) {
                const double beta = h->scalh.as<double>()[i * 4];
                if (m_out) m_out[i] = beta == 0.0 ? m : h->stath.as<int>()[i * 4];
                if (breakdown) breakdown[i] = beta == 0.0 ? 0 : h->stath.as<int>()[i * 4 + 1];
            }
            if (*h->errh.as<int>()) return fail(h, B200K_ESINGULAR, "SingularException(0) in the batch");
        }
        return B200K_OK;
    }
    st = fetch_krylov(h, nb, m);
    if (st) return st;
    // small dense phase per problem on the host, then one batched projection launch
    const int ldhd = m + 1;
    const size_t hstride = (size_t)ldhd * (m + 1);
    double *Yh = reinterpret_cast<double *>(stage_acquire(h, (size_t)nb * m * 8 + (size_t)nb * 16));
    if (!Yh) return fail(h, B200K_ENOMEM, "pinned staging buffer");
    double *betah = Yh + (size_t)nb * m;
    int *mh = reinterpret_cast<int *>(betah + nb);
    // the nb independent m x m exponentials are spread over host threads (they are ~50 us each)
    std::atomic<int> next(0), bad(0);
    auto worker = [&]() {
        smallmat::ExpWork work;
        std::vector<double> Hm((size_t)ldhd * (m + 1));
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= nb) break;
            const double beta = h->scalh.as<double>()[i * 4];
            int mo = h->stath.as<int>()[i * 4 + 0];
            const int bd = h->stath.as<int>()[i * 4 + 1];
            if (beta == 0.0) mo = m;
            if (m_out) m_out[i] = mo;
            if (breakdown) breakdown[i] = beta == 0.0 ? 0 : bd;
            betah[i] = beta;
            mh[i] = mo;
            double *y = Yh + (size_t)i * m;
            std::fill(y, y + m, 0.0);
            if (beta == 0.0) continue;
            std::memcpy(Hm.data(), h->Hh.as<double>() + i * hstride, hstride * 8);
            if (herm)
                for (int q = 0; q + 1 < m; ++q) Hm[(size_t)(q + 1) * ldhd + q] = Hm[(size_t)q * ldhd + q + 1];
            if (expv_small_ws(t[i], Hm.data(), ldhd, mo, y, work)) bad.store(1);
        }
    };
    {
        int nthr = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
        nthr = std::max(1, std::min(nthr, nb / 4));
        std::vector<std::thread> pool;
        for (int q = 1; q < nthr; ++q) pool.emplace_back(worker);
        worker();
        for (auto &th : pool) th.join();
    }
    if (bad.load()) return fail(h, B200K_ESINGULAR, "small dense exponential failed for a problem of the batch");
    CK(h, h->Y.ensure((size_t)nb * m * 8));
    CK(h, h->betavec.ensure((size_t)nb * 8));
    CK(h, h->mvec.ensure((size_t)nb * 4));
    CK(h, cudaMemcpyAsync(h->Y.p, Yh, (size_t)nb * m * 8, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->betavec.p, betah, (size_t)nb * 8, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->mvec.p, mh, (size_t)nb * 4, cudaMemcpyHostToDevice, h->stream));
    CK(h, stage_commit(h));
    ProjectParams P;
    std::memset(&P, 0, sizeof(P));
    P.V = h->V.as<double>();
    P.ldv = ldv;
    P.V_stride = vstride;
    P.nrows = n;
    P.Y = h->Y.as<double>();
    P.ldy = m;
    P.Y_stride = m;
    P.mvec = h->mvec.as<int>();
    P.betavec = h->betavec.as<double>();
    P.m = m;
    P.nc = 1;
    P.W = W;
    P.ldw = ldw;
    P.W_stride = ldw;
    P.vec2 = (n % 2 == 0) && (ldw % 2 == 0) && aligned16(W);
    const long long units = P.vec2 ? n / 2 : n;
    int gx = (int)std::min<long long>((units + PROJ_NT - 1) / PROJ_NT, 64);
    if (gx < 1) gx = 1;
    if (h->timing) CK(h, cudaEventRecord(h->ev[2], h->stream));
    for (int base = 0; base < nb; base += 32768) {  // gridDim.z limit is 65535
        const int cnt = std::min(nb - base, 32768);
        ProjectParams Q = P;
        Q.V += (long long)base * vstride;
        Q.Y += (long long)base * m;
        Q.mvec += base;
        Q.betavec += base;
        Q.W += (long long)base * ldw;
        launch_project_kernel(Q, dim3(gx, 1, cnt), h->stream);
        h->launches += 1;
    }
    CK(h, cudaGetLastError());
    if (h->timing) CK(h, cudaEventRecord(h->ev[3], h->stream));
    if (h->timing) h->ev_p = true;
    return B200K_OK;
}

// ---- kiops (src/kiops.jl:57-326) -------------------------------------------------------------------------
int b200k_kiops(b200k_handle_t h, b200k_op_t op, int ntau, const double *tau_out, int tau_is_row, const double *U,
                int64_t ldu, int ppo, const b200k_kiops_opts *ko, double *W, int64_t ldw, int64_t *stats) {
    if (!h || !op || !tau_out || !U || !ko || !W || ntau < 1 || ppo < 1) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    const long long n = op->n;
    if (ldu < n || ldw < n) return fail(h, B200K_EDIM, "Dimension mismatch");
    int p = ppo - 1;
    const bool pad_col = (p == 0);  // u = [u zero(u)]
    if (pad_col) p = 1;
    if (p > MAXP) return fail(h, B200K_EUNSUPPORTED, "kiops supports at most 16 phi columns");
    const int mmin = ko->mmin, mmax = ko->mmax;
    int m = ko->m > 0 ? ko->m : std::min(mmin, mmax);
    if (mmax >= MAXCOL - 1) return fail(h, B200K_EUNSUPPORTED, "mmax must be < 255");
    int herm = ko->hermitian;
    if (herm < 0) herm = op->is_herm;
    const int numSteps = tau_is_row ? ntau : 1;
    if (numSteps > 1)  // length(w) == size(A,1) fails in checkdims for a multi-column w (arnoldi.jl:217)
        return fail(h, B200K_EDIM, "DimensionMismatch: kiops with several output times (as in the reference)");
    const double tau_last = tau_out[ntau - 1];
    const double sgn = (tau_last > 0) - (tau_last < 0);
    double tau_now = 0.0;
    const double tau_end = std::fabs(tau_last);
    int j = 0;

    // Krylov storage: V (n+p) x (cap+1) grown like resize! (arnoldi.jl:80-93, contents preserved)
    const long long ldv = round_up(n + p, 16);
    int cap = m;
    {   // start with room for the dimensions the controller usually grows to, unless that is a lot of memory
        const int want = std::min(mmax, std::max(2 * m, 32));
        if ((size_t)ldv * (want + 1) * 8 <= ((size_t)8 << 30)) cap = std::max(cap, want);
    }
    DevBuf &Vbuf = h->kV, &Bbuf = h->kB;
    CK(h, Vbuf.ensure((size_t)ldv * (cap + 1) * 8));
    cap = std::max(cap, std::min(mmax, (int)(Vbuf.cap / ((size_t)ldv * 8)) - 1));  // use what is already there
    const int ldh = mmax + 2;
    std::vector<double> H((size_t)ldh * ldh, 0.0);
    int Ks_m = m;

    // w = zeros(n, numSteps); w[:, 1] = u[:, 1]
    CK(h, cudaMemcpyAsync(W, U, (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream));

    // normalisation factors (kiops.jl:94-102)
    double nu = 1.0, mu = 1.0;
    const long long ldbm = round_up(n, 2);
    cudaError_t e = Bbuf.ensure((size_t)ldbm * p * 8);
    if (e != cudaSuccess) return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
    auto cleanup = [&]() {};  // the buffers belong to the handle
    if (pad_col) {
        cudaMemsetAsync(Bbuf.p, 0, (size_t)ldbm * p * 8, h->stream);
    } else if (ko->normU == ko->normU) {  // supplied by the caller (row-sharded: the global 1-norm)
        if (ko->normU > 0) {
            const double ex = std::ceil(std::log2(ko->normU));
            nu = std::exp2(-ex);
            mu = std::exp2(ex);
        }
        flip_scale_kernel<<<1024, 256, 0, h->stream>>>(n, p, U, ldu, nu, Bbuf.as<double>(), ldbm);
        h->launches += 1;
    } else {
        const int nblk = 1024;
        e = h->tmp.ensure((size_t)nblk * 8);
        if (e != cudaSuccess) {
            cleanup();
            return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
        }
        abs_sum_kernel<<<nblk, 256, 0, h->stream>>>(n, p, U + ldu, ldu, h->tmp.as<double>());
        std::vector<double> part(nblk);
        cudaMemcpyAsync(part.data(), h->tmp.p, (size_t)nblk * 8, cudaMemcpyDeviceToHost, h->stream);
        e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) {
            cleanup();
            return fail(h, B200K_ECUDA, cudaGetErrorString(e));
        }
        double normU = 0.0;
        for (double v : part) normU += v;
        if (normU > 0) {
            const double ex = std::ceil(std::log2(normU));
            nu = std::exp2(-ex);
            mu = std::exp2(ex);
        }
        flip_scale_kernel<<<1024, 256, 0, h->stream>>>(n, p, U, ldu, nu, Bbuf.as<double>(), ldbm);
        h->launches += 2;
    }

    double tau = tau_end;
    double gamma, gamma_mmax;
    if (tau_end > 1) {
        gamma = 0.2;
        gamma_mmax = 0.1;
    } else {
        gamma = 0.9;
        gamma_mmax = 0.6;
    }
    const double delta = 1.4;
    int oldm = -1;
    double oldtau = NAN, omega = NAN;
    bool orderold = true, kestold = true;
    double order = 0.0, kest = 2.0;
    int l = 1;
    int64_t step = 0, reject = 0, ireject = 0, exps = 0;
    const int64_t krystep = 0;
    double beta = 0.0;
    std::vector<double> F, Hc;
    int status = B200K_OK;

    while (tau_now < tau_end) {
        const int oldj = Ks_m;
        if (m > cap) {  // resize!(Ks, m)
            DevBuf nv;
            e = nv.ensure((size_t)ldv * (m + 1) * 8);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(nv.p, Vbuf.p, (size_t)ldv * (cap + 1) * 8, cudaMemcpyDeviceToDevice, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) {
                nv.release();
                cleanup();
                return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
            }
            Vbuf.release();
            Vbuf.p = nv.p;
            Vbuf.cap = nv.cap;
            cap = m;
        }
        b200k_krylov_opts o;
        b200k_krylov_opts_default(&o);
        o.m = m;
        o.tol = 1.0e-7;  // kiops does not forward its tol to arnoldi! (kiops.jl:138-141)
        o.iop = ko->iop;
        o.hermitian = herm;
        o.init = j;
        o.p = p;
        o.B = Bbuf.as<double>();
        o.ldb = ldbm;
        o.t = tau_now;
        o.mu = mu;
        int mo = 0, bd = 0;
        status = arnoldi_core(h, op, W + (size_t)(l - 1) * ldw, &o, Vbuf.as<double>(), ldv, cap, H.data(), ldh,
                              &beta, &mo, &bd);
        if (status) break;
        Ks_m = (beta == 0.0) ? m : mo;
        j = Ks_m;
        bool happy = j < oldj;
        // phi_1 column for the error estimate; h_{j+1,j} removed while exponentiating (kiops.jl:149-160)
        H[(size_t)j * ldh + 0] = 1.0;
        const double nrm = H[(size_t)(j - 1) * ldh + j];
        H[(size_t)(j - 1) * ldh + j] = 0.0;
        const int N = j + 1;
        F.assign((size_t)N * N, 0.0);
        for (int c = 0; c < N; ++c)
            for (int r = 0; r < N; ++r) F[(size_t)c * N + r] = sgn * tau * H[(size_t)c * ldh + r];
        if (smallmat::expm_higham2005base(N, F.data(), h->expwork)) {
            status = fail(h, B200K_ESINGULAR, "SingularException(0) in kiops");
            break;
        }
        exps += 1;
        H[(size_t)(j - 1) * ldh + j] = nrm;
        double tau_new;
        int m_new;
        if (happy) {
            omega = 0;
            tau_new = std::min(tau_end - (tau_now + tau), tau);
            m_new = m;
            happy = false;
        } else {
            const double err = std::fabs(beta * nrm * F[(size_t)j * N + (j - 1)]);
            const double oldomega = omega;
            omega = tau_end * err / (tau * ko->tol);
            if (!(omega == omega)) {  // NaN (non-finite input): `ceil(Int, NaN)` is an InexactError in the reference
                status = fail(h, B200K_EARG, "InexactError: the kiops error estimate is NaN (non-finite operator or input?)");
                break;
            }
            if (m == oldm && tau != oldtau && ireject >= 1) {
                order = std::max(1.0, std::log(omega / oldomega) / std::log(tau / oldtau));
                orderold = false;
            } else if (orderold || ireject == 0) {
                orderold = true;
                order = j / 4.0;
            } else {
                orderold = true;
            }
            if (m != oldm && tau == oldtau && ireject >= 1) {
                kest = std::max(1.1, std::pow(omega / oldomega, 1.0 / (oldm - m)));
                kestold = false;
            } else if (kestold || ireject == 0) {
                kestold = true;
                kest = 2;
            } else {
                kestold = true;
            }
            const double remaining_time = omega > delta ? tau_end - tau_now : tau_end - (tau_now + tau);
            const double same_tau = std::min(remaining_time, tau);
            double tau_opt = tau * std::pow(gamma / omega, 1.0 / order);
            tau_opt = std::min(remaining_time, std::max(tau / 5, std::min(5 * tau, tau_opt)));
            // m_opt = ceil(Int, ...) throws InexactError in Julia for NaN/Inf; clamp instead
            double mo_d = std::ceil(j + std::log(omega / gamma) / std::log(kest));
            if (!(mo_d == mo_d)) mo_d = mmax;
            mo_d = std::max(-1.0e9, std::min(1.0e9, mo_d));
            int m_opt = (int)mo_d;
            // quirk kept: `3 ÷ 4 * m` == 0 and `cld(4, 3) * m` == 2m (kiops.jl:210)
            m_opt = std::max(mmin, std::min(mmax, std::max(0, std::min(m_opt, 2 * m))));
            if (j == mmax) {
                if (omega > delta) {
                    m_new = j;
                    tau_new = tau * std::pow(gamma_mmax / omega, 1.0 / order);
                    tau_new = std::min(tau_end - tau_now, std::max(tau / 5, tau_new));
                } else {
                    tau_new = tau_opt;
                    m_new = m;
                }
            } else {
                m_new = m_opt;
                tau_new = same_tau;
            }
        }
        if (omega <= delta) {  // kiops_update_solution! (kiops.jl:283-326)
            reject += ireject;
            step += 1;
            int blownTs = 0;
            const double nextT = tau_now + tau;
            for (int k = l; k <= numSteps; ++k)
                if (std::fabs(tau_out[k - 1]) < std::fabs(nextT)) blownTs += 1;
            if (blownTs != 0) {
                if (l + blownTs > numSteps) {  // w[:, l + blownTs] is a BoundsError in the reference (kiops.jl:303)
                    status = fail(h, B200K_EDIM, "BoundsError: kiops output time inside a step but w has no column for it "
                                                 "(pass several output times as a 1 x k row, as in the reference)");
                    break;
                }
                cudaMemcpyAsync(W + (size_t)(l + blownTs - 1) * ldw, W + (size_t)(l - 1) * ldw, (size_t)n * 8,
                                cudaMemcpyDeviceToDevice, h->stream);
                for (int k = 0; k < blownTs; ++k) {
                    const double tauPhantom = tau_out[l + k - 1] - tau_now;
                    Hc.assign((size_t)j * j, 0.0);
                    for (int c = 0; c < j; ++c)
                        for (int r = 0; r < j; ++r) Hc[(size_t)c * j + r] = sgn * tauPhantom * H[(size_t)c * ldh + r];
                    if (smallmat::expm_higham2005base(j, Hc.data(), h->expwork)) {
                        status = fail(h, B200K_ESINGULAR, "SingularException(0) in kiops");
                        break;
                    }
                    status = launch_project(h, Vbuf.as<double>(), ldv, n, j, beta, Hc.data(), j, 1,
                                            W + (size_t)(l + k - 1) * ldw, ldw, nullptr);
                    if (status) break;
                }
                if (status) break;
                l += blownTs;
            }
            status = launch_project(h, Vbuf.as<double>(), ldv, n, j, beta, F.data(), N, 1, W + (size_t)(l - 1) * ldw,
                                    ldw, nullptr);
            if (status) break;
            tau_now += tau;
            j = 0;
            ireject = 0;
        } else {
            ireject += 1;
            H[(size_t)j * ldh + 0] = 0.0;
        }
        oldtau = tau;
        tau = tau_new;
        oldm = m;
        m = m_new;
    }
    if (!status && tau_out[0] != 1 && ko->task1 && ntau == 1) {
        scale_kernel<<<1024, 256, 0, h->stream>>>(n, W + (size_t)(l - 1) * ldw, std::pow(1.0 / tau_out[l - 1], p));
        h->launches += 1;
    }
    cudaError_t es = cudaStreamSynchronize(h->stream);
    cleanup();
    if (status) return status;
    if (es != cudaSuccess) return fail(h, B200K_ECUDA, cudaGetErrorString(es));
    if (stats) {
        stats[0] = step;
        stats[1] = reject;
        stats[2] = krystep;
        stats[3] = exps;
        stats[4] = m;
    }
    return B200K_OK;
}

// ---- ComplexF64 path ----------------------------------------------------------------------------
namespace {
typedef smallmat::cplx cplx;

// y = exp(t H[0:m,0:m]) e1 for complex H and/or complex t; 0 ok, 1 eigensolver failure, 2 singular.
int expv_small_z_ws(cplx t, const cplx *H, int ldh, int m, cplx *y, int *branch) {
    // ishermitian(Hcopy) of a real-U (Lanczos) subspace: exactly real symmetric -> eigen!(SymTridiagonal(real(H)))
    bool realsym = true;
    for (int j = 0; j < m && realsym; ++j)
        for (int i = 0; i < m; ++i) {
            const cplx a = H[(size_t)j * ldh + i], bb = H[(size_t)i * ldh + j];
            if (a.imag() != 0.0 || a.real() != bb.real()) { realsym = false; break; }
        }
    if (branch) *branch = realsym ? 1 : 0;
    if (realsym) {
        std::vector<double> d(m), e(std::max(m - 1, 0)), Z;
        for (int i = 0; i < m; ++i) d[i] = H[(size_t)i * ldh + i].real();
        for (int i = 0; i + 1 < m; ++i) e[i] = H[(size_t)i * ldh + i + 1].real();
        if (!smallmat::symtridiag_eig(m, d, e, Z)) return 1;
        for (int i = 0; i < m; ++i) y[i] = cplx(0.0, 0.0);
        for (int k = 0; k < m; ++k) {
            const cplx wk = std::exp(t * d[k]) * Z[(size_t)k * m + 0];
            for (int i = 0; i < m; ++i) y[i] += Z[(size_t)k * m + i] * wk;
        }
        return 0;
    }
    std::vector<cplx> Hc((size_t)m * m);
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) Hc[(size_t)j * m + i] = t * H[(size_t)j * ldh + i];
    smallmat::ExpWorkT<cplx> work;
    if (smallmat::expm_higham2005base(m, Hc.data(), work)) return 2;
    for (int i = 0; i < m; ++i) y[i] = Hc[i];
    return 0;
}

int arnoldi_z_core(b200k_context *h, b200k_operator *op, const double *b, const b200k_krylov_opts *o, double *V,
                   long long ldv, int maxiter, cplx *H, int ldh, double *beta, int *m_out, int *breakdown) {
    if (!op->is_complex) return fail(h, B200K_EARG, "real operator: use the real entry points");
    const long long n = op->n;
    const int m = o->m;
    if (m < 1) return fail(h, B200K_EARG, "m must be >= 1");
    if (m > maxiter) return fail(h, B200K_EDIM, "m exceeds Ks.maxiter: resize the KrylovSubspace first");
    if (m >= MAXCOL) return fail(h, B200K_EUNSUPPORTED, "Krylov dimension m must be < 256");
    if (o->p != 0) return fail(h, B200K_EUNSUPPORTED, "the augmented operator is not available for complex types");
    if (ldv < n) return fail(h, B200K_EDIM, "DimensionMismatch: size(V,1) != size(A,1)");
    if (ldh < m + 1) return fail(h, B200K_EDIM, "H has fewer than m+1 rows");
    if (o->init < 0 || o->init > m) return fail(h, B200K_EARG, "init must be in 0..m");
    int herm = o->hermitian;
    if (herm < 0) herm = op->is_herm;
    *breakdown = 0;
    *m_out = m;
    if (o->init == 0) {
        for (int j = 0; j < m; ++j)
            for (int i = 0; i < m + 1; ++i) H[(size_t)j * ldh + i] = cplx(0.0, 0.0);
    } else if (*beta == 0.0) {
        return B200K_OK;
    }
    Geom g = single_geom(h, n);
    KrylovParamsZ P;
    std::memset(&P, 0, sizeof(P));
    P.n = (int)n;
    if (op->kind == 0) {
        P.op_kind = OP_CSR_WARP;
        P.rowptr = op->rowptr.as<int>();
        P.colind = op->colind.as<int>();
        P.val = op->val.as<double2>();
        P.lanes_per_row = op->max_row_nnz <= 8 ? 1 : (op->max_row_nnz <= 24 ? 4 : 32);
    } else {
        P.op_kind = OP_DENSE;
        P.Ad = reinterpret_cast<const double2 *>(op->Ad);
        P.lda = op->lda;
    }
    P.team_size = g.C;
    P.slice = g.slice;
    P.b = reinterpret_cast<const double2 *>(b);
    P.V = reinterpret_cast<double2 *>(V);
    P.ldv = ldv;
    P.m = m;
    P.lanczos = herm ? 1 : 0;
    P.j0 = herm ? (o->init == 0 ? 0 : 1) : o->init;
    P.iop = o->iop;
    P.tol = o->tol;
    const int ldhd = m + 1;
    P.ldh = ldhd;
    P.xlen = round_up(n, 16);
    size_t smem = sizeof(SmemZ) + (size_t)g.slice * 16;
    P.w_in_smem = smem <= SMEM_LIMIT ? 1 : 0;
    if (!P.w_in_smem) smem = sizeof(SmemZ);
    CK(h, h->zxbuf.ensure((size_t)2 * P.xlen * 16));
    CK(h, h->zpart.ensure((size_t)2 * MAXCOL * CPAD * 16));
    CK(h, h->partn.ensure((size_t)4 * CPAD * 8));
    CK(h, h->bar.ensure(64));
    CK(h, h->zH.ensure((size_t)ldhd * (m + 1) * 16));
    CK(h, h->scal.ensure(64));
    CK(h, h->stat.ensure(64));
    if (!P.w_in_smem) CK(h, h->zw.ensure((size_t)n * 16));
    P.xbuf = h->zxbuf.as<double2>();
    P.part = h->zpart.as<double2>();
    P.partn = h->partn.as<double>();
    P.bar = h->bar.as<unsigned>();
    P.Hd = h->zH.as<double2>();
    P.scal = h->scal.as<double>();
    P.stat = h->stat.as<int>();
    P.wglob = h->zw.as<double2>();
    CK(h, cudaMemsetAsync(P.bar, 0, 4, h->stream));
    CK(h, cudaMemsetAsync(P.Hd, 0, (size_t)ldhd * (m + 1) * 16, h->stream));
    CK(h, cudaMemsetAsync(P.scal, 0, 32, h->stream));
    if (h->timing) CK(h, cudaEventRecord(h->ev[0], h->stream));
    // ---- kernel selection: the TMA-ring instance for CSR operators with short rows whose w slice fits next to >= 3
    // ring slots (16-byte aligned bases; every complex element is one 16-byte unit, so no parity condition on n / ldv)
    bool tma = false;
    if (op->kind == 0 && !h->force_ldg && P.w_in_smem && ((uintptr_t)V & 15) == 0 && ((uintptr_t)b & 15) == 0) {
        const int mrn = std::max(op->max_row_nnz, 1);
        const size_t wsb = (size_t)round_up((long long)g.slice * 16, 128);
        const long long ring_room = (long long)SMEM_LIMIT - (long long)sizeof(SmemTmaZ) - (long long)wsb;
        // slot size = one basis tile; among the tilings of the slice pick the one with the most bytes in flight
        // (fewer, larger tiles on ties: every tile costs one mbarrier round trip)
        const int ntk0 = (g.slice + TILE_ROWS_Z - 1) / TILE_ROWS_Z;
        int best_ntk = 0, best_slot = 0, best_nslot = 0;
        long long best_inflight = 0;
        for (int ntk = ntk0; ntk <= ntk0 + 3; ++ntk) {
            const int tr = (int)round_up((g.slice + ntk - 1) / ntk, 16);
            // (a slot also holds one CSR chunk: room for the 256 rows the consumers can work on at once, up to 32 KB)
            const long long want_chunk = std::min<long long>(SLOT_BYTES, (long long)CHZ_ROWS_MAX * (20LL * mrn + 4) + 272);
            const int sb = (int)round_up(std::max<long long>((long long)tr * 16, want_chunk), 128);
            const int ns = (int)std::min<long long>(ring_room / sb, MAXSLOT);
            if (ns < 3) continue;
            const long long inflight = (long long)ns * tr * 16;
            if (inflight > best_inflight + best_inflight / 16) {
                best_inflight = inflight;
                best_ntk = ntk;
                best_slot = sb;
                best_nslot = ns;
            }
        }
        int ch_rows = best_slot > 0 ? (int)((best_slot - 20 * 12 - 32) / (20LL * mrn + 4)) : 0;
        ch_rows = std::min(ch_rows, CHZ_ROWS_MAX) / 32 * 32;
        if (ch_rows >= 64 && best_nslot >= 3) {
            tma = true;
            P.ch_rows = ch_rows;
            P.nnz_cap = (int)round_up((long long)ch_rows * mrn + 8, 4);
            P.nslot = best_nslot;
            P.slot_bytes = best_slot;
            P.tile_rows = (int)round_up((g.slice + best_ntk - 1) / best_ntk, 16);
            // L2 policy of the operator stream, as for the real kernel: evict_first once operator + two passes over the
            // orthogonalisation window no longer fit in 3/4 of L2
            P.hintA_cols = 1 << 30;
            if (h->l2hint == 1) P.hintA_cols = 0;
            else if (h->l2hint < 0) {
                const double budget = 0.75 * (double)h->l2_bytes;
                const double opb = 20.0 * (double)op->nnz + 4.0 * (double)(n + 1);
                const double colb = 16.0 * (double)n;
                const double cols = (budget - opb) / (2.0 * colb);
                P.hintA_cols = cols < 1.0 ? 1 : (cols > 1e9 ? (1 << 30) : (int)cols);
            }
            smem = sizeof(SmemTmaZ) + wsb + (size_t)best_nslot * best_slot;
        }
    }
    void *args[] = {(void *)&P};
    if (tma) {
        CK(h, cudaLaunchCooperativeKernel((const void *)krylov_tma_z_kernel, dim3(g.C), dim3(NT2), args, smem, h->stream));
        h->launches += 1;
        if (!herm) {
            // re-orthogonalisation check from the stored H (exits at once in the normal case), DESIGN 3.1e
            KrylovParamsZ P2 = P;
            P2.safe_scan = 1;
            P2.op_kind = OP_CSR_WARP;
            const size_t smem2 = sizeof(SmemZ) + (size_t)g.slice * 16;
            void *args2[] = {(void *)&P2};
            CK(h, cudaMemsetAsync(P.bar, 0, 4, h->stream));
            CK(h, cudaLaunchCooperativeKernel((const void *)krylov_z_kernel, dim3(g.C), dim3(NT), args2, smem2, h->stream));
            h->launches += 1;
        }
        h->last_kernel = 7;
    } else {
        CK(h, cudaLaunchCooperativeKernel((const void *)krylov_z_kernel, dim3(g.C), dim3(NT), args, smem, h->stream));
        h->launches += 1;
        h->last_kernel = 3;
    }
    if (h->timing) CK(h, cudaEventRecord(h->ev[1], h->stream));
    if (h->timing) { h->ev_k = true; h->ev_p = false; }
    const size_t hbytes = (size_t)ldhd * (m + 1) * 16;
    CK(h, h->zHh.ensure(hbytes + 64));
    CK(h, h->scalh.ensure(64));
    CK(h, h->stath.ensure(64));
    CK(h, cudaMemcpyAsync(h->zHh.p, h->zH.p, hbytes, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(h->scalh.p, h->scal.p, 32, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(h->stath.p, h->stat.p, 16, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (o->init == 0) *beta = h->scalh.as<double>()[0];
    if (*beta == 0.0) return B200K_OK;
    *m_out = h->stath.as<int>()[0];
    *breakdown = h->stath.as<int>()[1];
    const cplx *Hh = h->zHh.as<cplx>();
    const int jc0 = P.j0 == 0 ? 0 : P.j0 - 1;
    for (int jc = jc0; jc < *m_out; ++jc)
        for (int i = 0; i <= jc + 1; ++i) H[(size_t)jc * ldh + i] = Hh[(size_t)jc * ldhd + i];
    if (herm)
        for (int i = 0; i + 1 < m; ++i) H[(size_t)(i + 1) * ldh + i] = H[(size_t)i * ldh + i + 1];
    return B200K_OK;
}

int expv_ks_z_core(b200k_context *h, cplx t, const double *V, long long ldv, long long nrows, const cplx *H, int ldh,
                   int m, double beta, double *w) {
    if (m < 1) return fail(h, B200K_EARG, "m must be >= 1");
    if (m > MAXCOL) return fail(h, B200K_EUNSUPPORTED, "m must be <= 256");
    if (ldv < nrows) return fail(h, B200K_EDIM, "Dimension mismatch");
    if (beta == 0.0) {
        CK(h, cudaMemsetAsync(w, 0, (size_t)nrows * 16, h->stream));
        return B200K_OK;
    }
    std::vector<cplx> y(m);
    const int st = expv_small_z_ws(t, H, ldh, m, y.data(), nullptr);
    if (st == 1) return fail(h, B200K_ESINGULAR, "symmetric tridiagonal eigensolver did not converge");
    if (st == 2) return fail(h, B200K_ESINGULAR, "SingularException(0): Pade denominator is singular");
    CK(h, h->zy.ensure((size_t)m * 16));
    // (pageable source: the runtime stages it before returning, so the vector may go out of scope)
    CK(h, cudaMemcpyAsync(h->zy.p, y.data(), (size_t)m * 16, cudaMemcpyHostToDevice, h->stream));
    const int blocks = (int)std::min<long long>((nrows + 255) / 256, (long long)h->sm_count * 8);
    project_z_kernel<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<const double2 *>(V), ldv, nrows, m, beta,
                                                     h->zy.as<double2>(), reinterpret_cast<double2 *>(w));
    CK(h, cudaGetLastError());
    h->launches += 1;
    return B200K_OK;
}

int op_create_z_finish(b200k_context *h, b200k_operator *op) {
    cudaError_t e = h->tmp.ensure(64);
    if (e != cudaSuccess) return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
    cudaMemsetAsync(h->tmp.p, 0, 64, h->stream);
    int *d_max = h->tmp.as<int>();
    int *d_non = d_max + 1;
    double *d_norm = reinterpret_cast<double *>(h->tmp.as<char>() + 16);
    if (op->kind == 0) {
        const int blocks = (int)std::min<long long>((op->n + 255) / 256, 4096);
        csr_analyze_z_kernel<<<blocks, 256, 0, h->stream>>>((int)op->n, op->rowptr.as<int>(), op->colind.as<int>(),
                                                            op->val.as<double2>(), d_max, d_non, d_norm);
    } else {
        dense_analyze_z_kernel<<<(int)std::min<long long>((op->n + 127) / 128, 2048), 128, 0, h->stream>>>(
            (int)op->n, reinterpret_cast<const double2 *>(op->Ad), op->lda, d_non, d_norm);
    }
    char raw[32];
    e = cudaMemcpyAsync(raw, h->tmp.p, 32, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, B200K_ECUDA, cudaGetErrorString(e));
    int mx, non;
    double nrm;
    std::memcpy(&mx, raw, 4);
    std::memcpy(&non, raw + 4, 4);
    std::memcpy(&nrm, raw + 16, 8);
    op->max_row_nnz = mx;
    op->is_herm = non ? 0 : 1;
    op->opnorm_inf = nrm;
    h->launches += 1;
    return B200K_OK;
}
}  // namespace

int b200k_op_csr_create_z(b200k_handle_t h, int64_t n, int64_t nnz, const int32_t *rowptr, const int32_t *colind,
                          const double *val, int index_base, int location, b200k_op_t *out) {
    if (!h || !out) return B200K_EARG;
    *out = nullptr;
    if (n < 1 || nnz < 0 || !rowptr || (nnz > 0 && (!colind || !val))) return fail(h, B200K_EARG, "invalid CSR description");
    if (n > 2000000000LL || nnz > 2000000000LL) return fail(h, B200K_EUNSUPPORTED, "int32 CSR indices only");
    if (index_base != 0 && index_base != 1) return fail(h, B200K_EARG, "index_base must be 0 or 1");
    CK(h, cudaSetDevice(h->device));
    b200k_operator *op = new b200k_operator();
    op->ctx = h;
    op->device = h->device;
    op->kind = 0;
    op->is_complex = 1;
    op->n = n;
    op->nnz = nnz;
    const size_t pad = 64;
    cudaError_t e = op->rowptr.ensure((size_t)(n + 1 + pad) * 4);
    if (e == cudaSuccess) e = op->colind.ensure((size_t)(nnz + pad) * 4);
    if (e == cudaSuccess) e = op->val.ensure((size_t)(nnz + pad) * 16);
    if (e != cudaSuccess) {
        b200k_op_destroy(op);
        return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
    }
    const cudaMemcpyKind kind = location == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    cudaMemsetAsync(op->colind.p, 0, op->colind.cap, h->stream);
    cudaMemsetAsync(op->val.p, 0, op->val.cap, h->stream);
    cudaMemcpyAsync(op->rowptr.p, rowptr, (size_t)(n + 1) * 4, kind, h->stream);
    if (nnz > 0) {
        cudaMemcpyAsync(op->colind.p, colind, (size_t)nnz * 4, kind, h->stream);
        cudaMemcpyAsync(op->val.p, val, (size_t)nnz * 16, kind, h->stream);
    }
    if (index_base == 1) {
        rebase_kernel<<<256, 256, 0, h->stream>>>(op->rowptr.as<int>(), op->rowptr.as<int>(), n + 1, 1);
        if (nnz > 0) rebase_kernel<<<1024, 256, 0, h->stream>>>(op->colind.as<int>(), op->colind.as<int>(), nnz, 1);
    }
    const int st = op_create_z_finish(h, op);
    if (st) {
        b200k_op_destroy(op);
        return st;
    }
    *out = op;
    return B200K_OK;
}

int b200k_op_dense_create_z(b200k_handle_t h, int64_t n, const double *A, int64_t lda, int location, b200k_op_t *out) {
    if (!h || !out) return B200K_EARG;
    *out = nullptr;
    if (n < 1 || !A || lda < n) return fail(h, B200K_EARG, "invalid dense description");
    CK(h, cudaSetDevice(h->device));
    b200k_operator *op = new b200k_operator();
    op->ctx = h;
    op->device = h->device;
    op->kind = 1;
    op->is_complex = 1;
    op->n = n;
    op->nnz = n * n;
    if (location == 1) {
        cudaError_t e = op->Aown.ensure((size_t)n * n * 16);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(op->Aown.p, (size_t)n * 16, A, (size_t)lda * 16, (size_t)n * 16, (size_t)n,
                                  cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) {
            b200k_op_destroy(op);
            return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
        }
        op->Ad = op->Aown.as<double>();
        op->lda = n;
    } else {
        op->Ad = A;
        op->lda = lda;
    }
    const int st = op_create_z_finish(h, op);
    if (st) {
        b200k_op_destroy(op);
        return st;
    }
    *out = op;
    return B200K_OK;
}

int b200k_arnoldi_z(b200k_handle_t h, b200k_op_t op, const double *b, const b200k_krylov_opts *opts, double *V,
                    int64_t ldv, int maxiter, double *H, int ldh, double *beta, int *m_out, int *breakdown) {
    if (!h || !op || !b || !opts || !V || !H || !beta || !m_out || !breakdown) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    return arnoldi_z_core(h, op, b, opts, V, ldv, maxiter, reinterpret_cast<cplx *>(H), ldh, beta, m_out, breakdown);
}

int b200k_expv_ks_z(b200k_handle_t h, double t_re, double t_im, const double *V, int64_t ldv, int64_t nrows,
                    const double *H, int ldh, int m, double beta, double *w) {
    if (!h || !V || !H || !w) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    return expv_ks_z_core(h, cplx(t_re, t_im), V, ldv, nrows, reinterpret_cast<const cplx *>(H), ldh, m, beta, w);
}

int b200k_phiv_ks_z(b200k_handle_t h, double t_re, double t_im, const double *V, int64_t ldv, int64_t nrows,
                    const double *H, int ldh, int m, double beta, int k, int correct, double *W, int64_t ldw,
                    double *errest) {
    if (!h || !V || !H || !W) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    if (m < 1 || k < 1) return fail(h, B200K_EARG, "m >= 1 and k >= 1 required");
    if (m + 1 > MAXCOL) return fail(h, B200K_EUNSUPPORTED, "m must be < 256");
    if (ldv < nrows || ldw < nrows) return fail(h, B200K_EDIM, "Dimension mismatch");
    const cplx t(t_re, t_im);
    const cplx *Hz = reinterpret_cast<const cplx *>(H);
    std::vector<cplx> Hc((size_t)m * m), e(m, cplx(0.0, 0.0)), C2((size_t)m * (k + 1));
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) Hc[(size_t)j * m + i] = t * Hz[(size_t)j * ldh + i];
    e[0] = cplx(1.0, 0.0);
    smallmat::ExpWorkT<cplx> work;
    if (smallmat::phiv_dense(m, Hc.data(), m, e.data(), k, C2.data(), m, work))
        return fail(h, B200K_ESINGULAR, "SingularException(0): Pade denominator is singular");
    const cplx hlast = Hz[(size_t)(m - 1) * ldh + m];
    if (errest) *errest = std::abs(beta * hlast * t * C2[(size_t)k * m + (m - 1)]);
    if (beta == 0.0) {
        for (int c = 0; c <= k; ++c) CK(h, cudaMemsetAsync(W + (size_t)c * ldw * 2, 0, (size_t)nrows * 16, h->stream));
        return B200K_OK;
    }
    // column c: beta * V[:, 1:m] * C2[:, c] (+ betah * C2[m, c+1] * v_{m+1} for c < k when `correct`): the correction is
    // one more basis column with coefficient betah * C2[m, c+1] / beta
    const int mm = correct ? m + 1 : m;
    cplx *stage = reinterpret_cast<cplx *>(stage_acquire(h, (size_t)mm * (k + 1) * 16));
    if (!stage) return fail(h, B200K_ENOMEM, "pinned staging buffer");
    for (int c = 0; c <= k; ++c) {
        for (int i = 0; i < m; ++i) stage[(size_t)c * mm + i] = C2[(size_t)c * m + i];
        if (correct) stage[(size_t)c * mm + m] = c < k ? hlast * t * C2[(size_t)(c + 1) * m + (m - 1)] : cplx(0.0, 0.0);
    }
    CK(h, h->zy.ensure((size_t)mm * (k + 1) * 16));
    CK(h, cudaMemcpyAsync(h->zy.p, stage, (size_t)mm * (k + 1) * 16, cudaMemcpyHostToDevice, h->stream));
    CK(h, stage_commit(h));
    const int blocks = (int)std::min<long long>((nrows + 255) / 256, (long long)h->sm_count * 8);
    for (int c = 0; c <= k; ++c) {
        project_z_kernel<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<const double2 *>(V), ldv, nrows, mm, beta,
                                                         h->zy.as<double2>() + (size_t)c * mm,
                                                         reinterpret_cast<double2 *>(W) + (size_t)c * ldw);
        h->launches += 1;
    }
    CK(h, cudaGetLastError());
    return B200K_OK;
}

int b200k_expv_z(b200k_handle_t h, b200k_op_t op, double t_re, double t_im, const double *b,
                 const b200k_krylov_opts *opts, double *w, int *m_out, int *breakdown) {
    if (!h || !op || !b || !opts || !w) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    if (opts->p != 0 || opts->init != 0) return fail(h, B200K_EARG, "one-shot expv takes a plain operator, init = 0");
    b200k_krylov_opts o = *opts;
    o.m = (int)std::min<long long>(o.m, op->n);
    const long long ldv = round_up(op->n, 8);
    CK(h, h->zV.ensure((size_t)ldv * (o.m + 1) * 16));
    const int ldh = o.m + 2;
    std::vector<cplx> H((size_t)ldh * (o.m + 1), cplx(0.0, 0.0));
    double beta = 0.0;
    int mo = 0, bd = 0;
    int st = arnoldi_z_core(h, op, b, &o, h->zV.as<double>(), ldv, o.m, H.data(), ldh, &beta, &mo, &bd);
    if (st) return st;
    if (m_out) *m_out = mo;
    if (breakdown) *breakdown = bd;
    return expv_ks_z_core(h, cplx(t_re, t_im), h->zV.as<double>(), ldv, op->n, H.data(), ldh, mo, beta, w);
}

int b200k_expv_small_z(int m, const double *H, int ldh, double t_re, double t_im, double *y, int *branch) {
    if (m < 1 || !H || !y || ldh < m) return B200K_EARG;
    const int st = expv_small_z_ws(cplx(t_re, t_im), reinterpret_cast<const cplx *>(H), ldh, m,
                                   reinterpret_cast<cplx *>(y), branch);
    return st ? B200K_ESINGULAR : B200K_OK;
}

int b200k_project(b200k_handle_t h, const double *V, int64_t ldv, int64_t nrows, int m, double beta, const double *Y,
                  int ldy, int nc, double *W, int64_t ldw) {
    if (!h || !V || !Y || !W || m < 1 || nc < 1 || ldy < m) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    if (ldv < nrows || ldw < nrows) return fail(h, B200K_EDIM, "Dimension mismatch");
    return launch_project(h, V, ldv, nrows, m, beta, Y, ldy, nc, W, ldw, nullptr);
}

// ---- phiv_timestep! (src/krylov_phiv_adaptive.jl:260-501) --------------------------------------------------
namespace {
long long ts_flops(int m, double tau, long long n, int p, long long NA, int iop, double Hnorm, double maxtau) {
    const long long flops_W = 2LL * (p - 1) * (NA + n);
    const long long flops_u = (2LL * p + 1) * n;
    if (iop == 0) iop = m;
    const long long flops_matvec = 2LL * m * NA;
    long long flops_vecvec = 0;
    for (int i = 1; i <= m; ++i) flops_vecvec += 3 * std::min(i, iop);
    const double MH = 44.0 / 3.0 + 2.0 * std::ceil(std::max(0.0, std::log2(Hnorm / 5.37)));
    const long long flops_phiv = (long long)std::llround(MH * std::pow((double)(m + p), 3));
    return (flops_W + flops_u + flops_matvec + flops_vecvec + flops_phiv) * (long long)std::ceil(maxtau / tau);
}
double hnorm1(const double *H, int ldh, int rows, int cols) {
    double best = 0.0;
    for (int j = 0; j < cols; ++j) {
        double s = 0.0;
        for (int i = 0; i < rows; ++i) s += std::fabs(H[(size_t)j * ldh + i]);
        best = std::max(best, s);
    }
    return best;
}
}  // namespace

void b200k_timestep_opts_default(b200k_timestep_opts *o) {
    o->tau = 0.0;
    o->m = 10;
    o->tol = 1.0e-7;
    o->opnorm = NAN;
    o->iop = 0;
    o->correct = 0;
    o->adaptive = 0;
    o->delta = 1.2;
    o->hermitian = -1;
    o->gamma = 0.8;
    o->NA = 0;
}

int b200k_phiv_timestep(b200k_handle_t h, b200k_op_t op, int nts, double *ts, const double *B, int64_t ldb,
                        int ncoef, const b200k_timestep_opts *to, double *U, int64_t ldu, int *num_timesteps) {
    if (!h || !op || !ts || !B || !to || !U || nts < 1 || ncoef < 1) return B200K_EARG;
    CK(h, cudaSetDevice(h->device));
    const long long n = op->n;
    if (ldb < n || ldu < n) return fail(h, B200K_EDIM, "Dimension mismatch");
    if (op->comm) return fail(h, B200K_EUNSUPPORTED, "phiv_timestep on a row-sharded operator is not implemented");
    const int p = ncoef - 1;
    int m = (int)std::min<long long>(to->m, n);
    if (m < 1) return fail(h, B200K_EARG, "m must be >= 1");
    const double tol = to->tol, gamma = to->gamma, delta = to->delta;
    int iop = to->iop;
    double tau = to->tau;
    const bool arnoldi_scale = !(to->opnorm == to->opnorm);
    bool have_abstol = false;
    double abstol = 0.0, opn = 0.0;
    auto launch1d = [&](long long len) { return (int)std::min<long long>((len + 255) / 256, (long long)h->sm_count * 8); };
    auto absmax = [&](const double *x, double *out) -> int {
        const int nblk = 256;
        CK(h, h->tmp.ensure((size_t)nblk * 8));
        abs_max_kernel<<<nblk, 256, 0, h->stream>>>(n, x, h->tmp.as<double>());
        std::vector<double> part(nblk);
        CK(h, cudaMemcpyAsync(part.data(), h->tmp.p, (size_t)nblk * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        h->launches += 1;
        *out = *std::max_element(part.begin(), part.end());
        return B200K_OK;
    };
    auto tau_formula = [&](double opn_, double abstol_, double b0norm) {
        return 10.0 / opn_ * std::pow(abstol_ * std::pow((m + 1) / M_E, m + 1) * std::sqrt(2 * M_PI * (m + 1)) /
                                          (4 * opn_ * b0norm),
                                      1.0 / m);
    };
    double b0norm = -1.0;
    if (!arnoldi_scale) {
        opn = to->opnorm;
        abstol = tol * opn;
        have_abstol = true;
        if (tau == 0.0) {
            int st = absmax(B, &b0norm);
            if (st) return st;
            tau = tau_formula(opn, abstol, b0norm);
        }
    }
    std::sort(ts, ts + nts);
    const double tend = ts[nts - 1];
    const bool seed_arnoldi_tau = arnoldi_scale && tau == 0.0;
    if (seed_arnoldi_tau) tau = tend;
    int herm = to->hermitian;
    if (herm < 0) herm = op->is_herm;
    long long NA = to->NA;
    if (to->adaptive) {
        if (herm) iop = 2;  // "does not have an effect on arnoldi!, just for flops estimation" (:332-334)
        if (NA == 0) NA = op->nnz;
    }
    // workspace: u, W (n x (p+1)), P (n x (p+2)), V (n x (cap+1))
    const long long ld = round_up(n, 16);
    int cap = m;
    CK(h, h->tsu.ensure((size_t)ld * 8));
    CK(h, h->tsW.ensure((size_t)ld * (p + 1) * 8));
    CK(h, h->tsP.ensure((size_t)ld * (p + 2) * 8));
    CK(h, h->tsV.ensure((size_t)ld * (cap + 1) * 8));
    double *u = h->tsu.as<double>(), *W = h->tsW.as<double>(), *P = h->tsP.as<double>();
    std::vector<double> H;
    int ldh = 0;
    auto ensure_cap = [&](int mm) -> int {
        if (mm > cap || H.empty()) {
            cap = std::max(cap, mm);
            CK(h, h->tsV.ensure((size_t)ld * (cap + 1) * 8));
            ldh = cap + 2;
            H.assign((size_t)ldh * (cap + 1), 0.0);
        }
        return B200K_OK;
    };
    int st = ensure_cap(m);
    if (st) return st;
    CK(h, cudaMemcpyAsync(u, B, (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream));  // u(0) = b0
    std::vector<double> coeffs(std::max(p, 1), 1.0);
    double beta = 0.0;
    int mo = 0, bd = 0;
    auto do_arnoldi = [&]() -> int {  // arnoldi!(Ks, A, W[:, end]; tol, m, iop): dispatches on ishermitian(A)
        int s2 = ensure_cap(m);
        if (s2) return s2;
        b200k_krylov_opts o;
        b200k_krylov_opts_default(&o);
        o.m = m;
        o.tol = tol;
        o.iop = iop;  // as the reference: the (possibly overridden) iop; Lanczos ignores it
        o.hermitian = op->is_herm;
        beta = 0.0;
        return arnoldi_core(h, op, W + (size_t)p * ld, &o, h->tsV.as<double>(), ld, cap, H.data(), ldh, &beta, &mo, &bd);
    };
    auto do_phiv = [&](double tt, double *eps) -> int {
        return phiv_ks_core(h, tt, h->tsV.as<double>(), ld, n, H.data(), ldh, mo, beta, p + 1, to->correct, P, ld, eps);
    };
    auto combine = [&](double tt, double *dst) -> int {  // dst = tt^p * P[:, end-1] + sum_j coeffs_j(tt) W[:, j]
        scalecopy_kernel<<<launch1d(n), 256, 0, h->stream>>>(n, std::pow(tt, p), P + (size_t)p * ld, dst);
        for (int l = 1; l <= p - 1; ++l) coeffs[l] = coeffs[l - 1] * tt / l;
        for (int j = 0; j <= p - 1; ++j) axpy_kernel<<<launch1d(n), 256, 0, h->stream>>>(n, coeffs[j], W + (size_t)j * ld, dst);
        CK(h, cudaGetLastError());
        h->launches += 1 + p;
        return B200K_OK;
    };
    double t = 0.0;
    int snapshot = 1, nsteps = 0;
    while (t < tend) {
        if (t + tau > tend) tau = tend - t;
        CK(h, cudaMemcpyAsync(W, u, (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream));  // w0 = u(t)
        for (int l = 1; l <= p - 1; ++l) coeffs[l] = coeffs[l - 1] * t / l;
        for (int j = 1; j <= p; ++j) {
            st = b200k_op_apply(h, op, W + (size_t)(j - 1) * ld, W + (size_t)j * ld);
            if (st) return st;
            for (int l = 0; l <= p - j; ++l)
                axpy_kernel<<<launch1d(n), 256, 0, h->stream>>>(n, coeffs[l], B + (size_t)(j + l) * ldb, W + (size_t)j * ld);
            h->launches += p - j + 1;
        }
        st = do_arnoldi();
        if (st) return st;
        if (!have_abstol) {
            opn = hnorm1(H.data(), ldh, mo + 1, mo);
            abstol = tol * opn;
            have_abstol = true;
            if (seed_arnoldi_tau) {
                if (b0norm < 0) {
                    st = absmax(B, &b0norm);
                    if (st) return st;
                }
                tau = std::min(tend - t, gamma * tau_formula(opn, abstol, b0norm));
            }
        }
        if (bd) tau = tend - t;
        double epsilon = 0.0;
        st = do_phiv(tau, &epsilon);
        if (st) return st;
        if (to->adaptive) {
            double omega = (tend / tau) * (epsilon / abstol);
            double epsilon_old = epsilon, tau_old = tau, q = m / 4.0, kappa = 2.0;
            int m_old = m;
            const double maxtau = tend - t;
            int guard = 0;
            while (omega > delta && guard++ < 200) {  // inner loop of Algorithm 3
                const double Hn = hnorm1(H.data(), ldh, mo + 1, mo);
                if (tau_old > tau) q = std::log(tau / tau_old) / std::log(epsilon / epsilon_old) - 1;
                double tau_new = tau * std::pow(gamma / omega, 1.0 / (q + 1));
                tau_new = std::min(std::min(std::max(tau_new, tau / 5), 2 * tau), maxtau);
                if (m_old < m) kappa = std::pow(epsilon / epsilon_old, 1.0 / (m_old - m));
                double mn = m + std::ceil(std::log(omega / gamma) / std::log(kappa));
                if (!(mn == mn)) mn = m;
                mn = std::max(-1.0e9, std::min(1.0e9, mn));
                int m_new = std::min(std::max(std::max((int)mn, (3 * m) / 4), 1), (int)std::ceil(4.0 * m / 3.0));
                m_new = (int)std::min<long long>(m_new, std::min<long long>(n, MAXCOL - 1));
                const long long cost_tau = ts_flops(m, tau_new, n, p, NA, iop, Hn, maxtau);
                const long long cost_m = ts_flops(m_new, tau, n, p, NA, iop, Hn, maxtau);
                if (cost_tau < cost_m) m_new = m;
                else tau_new = tau;
                m_old = m;
                m = m_new;
                tau_old = tau;
                tau = tau_new;
                st = do_arnoldi();
                if (st) return st;
                double eps_new = 0.0;
                st = do_phiv(tau, &eps_new);
                if (st) return st;
                epsilon_old = epsilon;
                epsilon = eps_new;
                omega = (tend / tau) * (epsilon / abstol);
            }
        }
        st = combine(tau, u);
        if (st) return st;
        while (snapshot <= nts && t + tau >= ts[snapshot - 1]) {
            const double tau_s = ts[snapshot - 1] - t;
            st = do_phiv(tau_s, nullptr);
            if (st) return st;
            st = combine(tau_s, U + (size_t)(snapshot - 1) * ldu);
            if (st) return st;
            snapshot += 1;
        }
        t += tau;
        nsteps += 1;
    }
    if (num_timesteps) *num_timesteps = nsteps;
    return B200K_OK;
}

// ---- row sharding across GPUs -----------------------------------------------------------------------------
int b200k_comm_create(b200k_handle_t h, int rank, int nranks, int64_t xlen, unsigned char *handle_out,
                      b200k_comm_t *out) {
    if (!h || !handle_out || !out) return B200K_EARG;
    *out = nullptr;
    if (nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks) return fail(h, B200K_EARG, "1 <= nranks <= 8");
    if (xlen < 16) return fail(h, B200K_EARG, "xlen too small");
    CK(h, cudaSetDevice(h->device));
    b200k_comm *cm = new b200k_comm();
    cm->ctx = h;
    cm->device = h->device;
    cm->rank = rank;
    cm->nranks = nranks;
    cm->xlen = round_up(xlen, 16);
    cm->cpad = (int)round_up((long long)h->max_ctas * nranks, 32);
    cm->bytes = b200k_comm::HDR + sizeof(double) * ((size_t)2 * MAXCOL * cm->cpad + (size_t)4 * cm->cpad + (size_t)4 * cm->xlen);
    cudaError_t e = cudaMalloc(&cm->local, cm->bytes);
    if (e == cudaSuccess) e = cudaMemset(cm->local, 0, cm->bytes);
    cudaIpcMemHandle_t hd;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hd, cm->local);
    if (e != cudaSuccess) {
        if (cm->local) cudaFree(cm->local);
        delete cm;
        return fail(h, B200K_ECOMM, std::string("comm_create: ") + cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == B200K_IPC_HANDLE_BYTES, "IPC handle size");
    std::memcpy(handle_out, &hd, sizeof(hd));
    cm->peer[rank] = cm->local;
    *out = cm;
    return B200K_OK;
}

int b200k_comm_connect(b200k_comm_t cm, const unsigned char *all_handles) {
    if (!cm || !all_handles) return B200K_EARG;
    b200k_context *h = cm->ctx;
    CK(h, cudaSetDevice(h->device));
    for (int r = 0; r < cm->nranks; ++r) {
        if (r == cm->rank) continue;
        cudaIpcMemHandle_t hd;
        std::memcpy(&hd, all_handles + (size_t)r * B200K_IPC_HANDLE_BYTES, sizeof(hd));
        cudaError_t e = cudaIpcOpenMemHandle(&cm->peer[r], hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return fail(h, B200K_ECOMM, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " +
                                            cudaGetErrorString(e));
    }
    cm->connected = true;
    return B200K_OK;
}

int b200k_comm_destroy(b200k_comm_t cm) {
    if (!cm) return B200K_OK;
    cudaSetDevice(cm->device);
    cudaDeviceSynchronize();  // the handle (and its stream) may already have been destroyed
    for (int r = 0; r < cm->nranks; ++r)
        if (r != cm->rank && cm->peer[r]) cudaIpcCloseMemHandle(cm->peer[r]);
    if (cm->local) cudaFree(cm->local);
    delete cm;
    return B200K_OK;
}

int b200k_op_csr_create_sharded(b200k_handle_t h, b200k_comm_t cm, int64_t nloc, int64_t nhalo, int64_t nnz,
                                const int32_t *rowptr, const int32_t *colind, const double *val, int index_base,
                                int location, int is_hermitian, int64_t nsend, const int32_t *send_row,
                                const int32_t *send_peer, const int32_t *send_pos, b200k_op_t *out) {
    if (!h || !cm || !out) return B200K_EARG;
    if (nhalo < 0 || nsend < 0 || (nsend > 0 && (!send_row || !send_peer || !send_pos)))
        return fail(h, B200K_EARG, "invalid halo description");
    if (nloc % 2 != 0) return fail(h, B200K_EUNSUPPORTED, "row-sharded blocks need an even number of rows");
    int st = b200k_op_csr_create(h, nloc, nnz, rowptr, colind, val, index_base, location, out);
    if (st) return st;
    b200k_operator *op = *out;
    op->comm = cm;
    op->nhalo = nhalo;
    op->is_herm = is_hermitian ? 1 : 0;  // the local block cannot decide symmetry of the global operator
    for (int64_t i = 0; i < nsend; ++i) {
        if (send_row[i] < 0 || send_row[i] >= nloc || send_peer[i] < 0 || send_peer[i] >= cm->nranks ||
            (i > 0 && send_row[i] < send_row[i - 1])) {
            b200k_op_destroy(op);
            *out = nullptr;
            return fail(h, B200K_EARG, "send list must be sorted by row with valid peers");
        }
    }
    op->send_row_host.assign(send_row, send_row + nsend);
    const size_t bytes = (size_t)std::max<int64_t>(nsend, 1) * 4;
    cudaError_t e = op->send_row.ensure(bytes);
    if (e == cudaSuccess) e = op->send_peer.ensure(bytes);
    if (e == cudaSuccess) e = op->send_pos.ensure(bytes);
    if (e == cudaSuccess && nsend > 0) {
        cudaMemcpy(op->send_row.p, send_row, (size_t)nsend * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(op->send_peer.p, send_peer, (size_t)nsend * 4, cudaMemcpyHostToDevice);
        e = cudaMemcpy(op->send_pos.p, send_pos, (size_t)nsend * 4, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        b200k_op_destroy(op);
        *out = nullptr;
        return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
    }
    return B200K_OK;
}

// A row block of a DENSE operator (C3 beyond one GPU).  Every rank needs the whole of x each step: its
// gather buffer holds the nloc own entries first, then all other rows in ascending global order ("halo" = n - nloc),
// so the block's COLUMNS are permuted into that order while it is copied into library storage, and every own row is
// pushed to every peer (send list built here from the row partition).  The persistent kernel is unchanged: the dense
// mat-vec streams an nloc x n block instead of n x n and the halo push / all-reduces are those of the sparse case.
int b200k_op_dense_create_sharded(b200k_handle_t h, b200k_comm_t cm, int64_t n_global, const int64_t *row_starts,
                                  const double *A_block, int64_t lda, int location, int is_hermitian, b200k_op_t *out) {
    if (!h || !cm || !out || !row_starts || !A_block) return B200K_EARG;
    *out = nullptr;
    const int R = cm->nranks, me = cm->rank;
    if (row_starts[0] != 0 || row_starts[R] != n_global) return fail(h, B200K_EARG, "row_starts must run from 0 to n");
    for (int r = 0; r < R; ++r)
        if (row_starts[r + 1] <= row_starts[r]) return fail(h, B200K_EARG, "row_starts must be increasing");
    const int64_t row0 = row_starts[me], nloc = row_starts[me + 1] - row0;
    if (nloc % 2 != 0 || row0 % 2 != 0) return fail(h, B200K_EUNSUPPORTED, "row-sharded blocks need even row counts");
    if (lda < nloc) return fail(h, B200K_EARG, "lda < number of local rows");
    if (n_global > 2000000000LL) return fail(h, B200K_EUNSUPPORTED, "n too large");
    if (n_global + MAXP > cm->xlen) return fail(h, B200K_EDIM, "communicator gather buffer (xlen) too small: need n + 16");
    CK(h, cudaSetDevice(h->device));
    b200k_operator *op = new b200k_operator();
    op->ctx = h;
    op->device = h->device;
    op->kind = 1;
    op->n = nloc;
    op->ncols = n_global;
    op->nnz = nloc * n_global;
    op->comm = cm;
    op->nhalo = n_global - nloc;
    op->is_herm = is_hermitian ? 1 : 0;
    const long long ld = round_up(nloc, 2);
    cudaError_t e = op->Aown.ensure((size_t)ld * n_global * 8);
    if (e != cudaSuccess) {
        delete op;
        return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
    }
    const cudaMemcpyKind kind = location == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    double *dst = op->Aown.as<double>();
    // own columns -> [0, nloc); columns left of the block -> [nloc, nloc + row0); columns right of it -> the rest
    struct Piece { int64_t src_col, dst_col, ncol; } pieces[3] = {
        {row0, 0, nloc}, {0, nloc, row0}, {row0 + nloc, nloc + row0, n_global - row0 - nloc}};
    for (const Piece &pc : pieces) {
        if (pc.ncol <= 0) continue;
        e = cudaMemcpy2DAsync(dst + (size_t)pc.dst_col * ld, (size_t)ld * 8, A_block + (size_t)pc.src_col * lda,
                              (size_t)lda * 8, (size_t)nloc * 8, (size_t)pc.ncol, kind, h->stream);
        if (e != cudaSuccess) break;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
        delete op;
        return fail(h, B200K_ECUDA, cudaGetErrorString(e));
    }
    op->Ad = dst;
    op->lda = ld;
    // opnorm(A, Inf) of the block (the caller combines the ranks with a max if it needs the global value)
    op->opnorm_inf = 0.0;
    // send list: own row i goes to every peer q at the position of global row row0 + i in q's gather buffer
    const int64_t nsend = nloc * (R - 1);
    std::vector<int> srow((size_t)nsend), speer((size_t)nsend), spos((size_t)nsend);
    size_t k = 0;
    for (int64_t i = 0; i < nloc; ++i) {
        const int64_t g = row0 + i;
        for (int q = 0; q < R; ++q) {
            if (q == me) continue;
            const int64_t q0 = row_starts[q], qn = row_starts[q + 1] - q0;
            srow[k] = (int)i;
            speer[k] = q;
            spos[k] = (int)(g < q0 ? qn + g : g);  // (g >= q0 + qn: qn + (g - qn) = g)
            ++k;
        }
    }
    op->send_row_host = srow;
    const size_t bytes = (size_t)std::max<int64_t>(nsend, 1) * 4;
    e = op->send_row.ensure(bytes);
    if (e == cudaSuccess) e = op->send_peer.ensure(bytes);
    if (e == cudaSuccess) e = op->send_pos.ensure(bytes);
    if (e == cudaSuccess && nsend > 0) {
        cudaMemcpy(op->send_row.p, srow.data(), (size_t)nsend * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(op->send_peer.p, speer.data(), (size_t)nsend * 4, cudaMemcpyHostToDevice);
        e = cudaMemcpy(op->send_pos.p, spos.data(), (size_t)nsend * 4, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        b200k_op_destroy(op);
        return fail(h, B200K_ENOMEM, cudaGetErrorString(e));
    }
    *out = op;
    return B200K_OK;
}

// ---- small dense (host) ----------------------------------------------------------------------------------
int b200k_exponential(int n, double *A, int lda) {
    if (n < 0 || !A || lda < n) return B200K_EARG;
    std::vector<double> C((size_t)n * n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) C[(size_t)j * n + i] = A[(size_t)j * lda + i];
    smallmat::ExpWork w;
    const int st = smallmat::expm_higham2005base(n, C.data(), w);
    if (st) return B200K_ESINGULAR;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) A[(size_t)j * lda + i] = C[(size_t)j * n + i];
    return B200K_OK;
}

int b200k_exponential_batched(b200k_handle_t h, int nbatch, int n, double *A, int lda, int64_t stride) {
    if (!h || !A || nbatch < 1 || n < 1 || lda < n || stride < (int64_t)lda * (n - 1) + n) return B200K_EARG;
    if (n > SE_MAXM) return fail(h, B200K_EUNSUPPORTED, "device-side exponential supports n <= 48");
    CK(h, cudaSetDevice(h->device));
    CK(h, h->errdev.ensure(64));
    CK(h, cudaMemsetAsync(h->errdev.p, 0, 4, h->stream));
    for (int base = 0; base < nbatch; base += 65535) {
        const int cnt = std::min(nbatch - base, 65535);
        small_exp_batched_kernel<<<cnt, SE_NT, (size_t)6 * n * n * 8, h->stream>>>(n, A + (long long)base * stride, lda,
                                                                                  stride, h->errdev.as<int>());
        h->launches += 1;
    }
    CK(h, cudaGetLastError());
    int err = 0;
    CK(h, cudaMemcpyAsync(&err, h->errdev.p, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (err) return fail(h, B200K_ESINGULAR, "SingularException(0): Pade denominator is singular");
    return B200K_OK;
}

int b200k_expv_small(int m, const double *H, int ldh, double t, double *y, int *branch) {
    if (m < 1 || !H || !y || ldh < m) return B200K_EARG;
    if (branch) *branch = smallmat::is_exactly_symmetric(m, H, ldh) ? 1 : 0;
    b200k_context tmp;
    return expv_small(&tmp, t, H, ldh, m, y);
}

int b200k_phiv_dense(int m, const double *A, int lda, const double *v, int k, double *w, int ldw) {
    if (m < 1 || k < 1 || !A || !v || !w || lda < m || ldw < m) return B200K_EARG;
    smallmat::ExpWork work;
    const int st = smallmat::phiv_dense(m, A, lda, v, k, w, ldw, work);
    return st ? B200K_ESINGULAR : B200K_OK;
}

#ifdef B200K_PHASE_TIMING
// profiling builds only: copy the per-phase clock64 stamps of the last krylov_tma_kernel launch to the host
int b200k_debug_phase_ts(long long *out, long long count) {
    const long long total = (long long)PT_CTAS * PT_STEPS * PT_MARKS;
    if (count > total) count = total;
    return (int)cudaMemcpyFromSymbol(out, g_phase_ts, (size_t)count * 8);
}
int b200k_debug_se_ts(long long *out) { return (int)cudaMemcpyFromSymbol(out, g_se_ts, 16 * 8); }
#endif

}  // extern "C"
