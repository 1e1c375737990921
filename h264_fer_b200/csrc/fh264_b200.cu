// fh264_b200 — C-ABI shim (include/fh264_b200.h) over the sm_100a kernels. One translation unit.
// Session = one GPU + `batch` sequences advancing in lockstep; all work is enqueued on one CUDA stream.
// There is deliberately NO CPU fallback: without a usable sm_100 device every entry point fails loudly.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "phase_r.cuh"
#include "phase_a.cuh"
#include "stage3.cuh"
#include "phase_b.cuh"
#include "phase_bw.cuh"
#include "phase_c.cuh"
#include "cavlc.cuh"
#include "intra.cuh"
#include "intpeak.cuh"
static_assert(sizeof(CvInfo) == sizeof(fh264_cavlc_mb_info) && sizeof(CvInfo) == 32, "CvInfo is the public fh264_cavlc_mb_info");

static thread_local std::string g_err;
static int fail(int code, const char *what, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof buf, "%s", what);
    g_err = buf;
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(FH264_E_CUDA, #call, e_); } while (0)
#define CKL() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return fail(FH264_E_CUDA, "kernel launch", e_); } while (0)

struct PeerSync { uint32_t *p[FH_MAX_WORLD]; };

struct fh264_session;
static cudaError_t sync_streams(fh264_session *s);
static int enter_main(fh264_session *s);

// A coding lane: the streams and events one half of a call's sequences is enqueued on. Lane 2 (`main`) is the session stream
// itself and codes a whole call the plain way. Lanes 0 / 1 are the software pipeline of fh264_encode_p_stream and of
// fh264_encode_p_async with FH264_PIPE=1: the call's sequences are split into two halves, each half on its own stream, phase B
// (a latency-bound wavefront that leaves most of the SMs idle) on a high-priority side stream, and the second half's phase A is
// held back until the first half's is through — so one half's wavefront always runs under the other half's search kernels, within
// a call and across consecutive calls (nothing on a lane waits for the other lane's picture to finish).
struct Lane {
    cudaStream_t st, hi, cp;        // coding stream; phase B stream (null: phase B on st); copies home
    cudaEvent_t ev_a, ev_b;         // phase A (+ S) enqueued work done; phase B done
    cudaEvent_t ev_c_done, ev_copy_done, ev_done;
    bool copy_pending;              // records / slice data of the lane's last picture still travelling (phase C of the next must wait)
    bool busy;                      // work enqueued since the last join with the session stream
    uint32_t *d_ticket;             // phase B ticket counter
    bool timing;                    // records the per-phase timing events (fh264_last_timings)
    cudaEvent_t tr[8][5];           // FH264_TRACE=1: timing events of the last 8 pictures (phase A start / end, phase B end, phase C end, phase R end)
    int tr_n;
};

struct fh264_session {
    Geo g;
    int batch, device;
    cudaStream_t stream;
    bool own_stream;
    cudaStream_t copy_stream;       // result records go home on their own stream, overlapping phase R and the next picture
    Lane lane[3];                   // [2] = the session stream (plain calls); [0], [1] = pipeline halves
    int pipeline;                   // 1: fh264_encode_p_async splits its sequences over lanes 0 / 1 (FH264_PIPE, fh264_set_pipeline)
    cudaEvent_t ev_fork;
    bool main_dirty;                // work enqueued on the session stream since the lanes last forked from it
    // double-buffered source pictures: uploads run on their own stream into the buffer that is not being coded
    cudaStream_t up_stream;
    std::vector<cudaEvent_t> ev_up;                 // per sequence: upload complete
    std::vector<cudaEvent_t> ev_free[2];            // per sequence and buffer: last picture that read it is through phase C
    std::vector<char> up_pending, cur_set, free_valid[2];
    std::vector<SeqDev> h;          // host mirror of the device SeqDev array
    SeqDev *d_seqs;
    int *d_wf_order;
    uint32_t *d_ticket;
    uint32_t epoch;
    std::vector<char> has_ref;
    uint32_t *h_status;             // pinned: batch * ST_WORDS, snapshot after phase C of the last encode
    uint64_t *h_sad;                // pinned: batch scene SADs
    unsigned long long *d_sadout;
    cudaEvent_t ev[5];
    cudaEvent_t evk[4];             // after stage3, after interp, after features (per-kernel split of phases A and R)
    cudaEvent_t ev_spec;            // after stage 2, before phase S
    int use_spec;                   // phase S + fast path in phase B (FH264_SPEC=0 turns it off: every partition takes the full search)
    // TMA descriptors of the 16 interpolated planes of every sequence (qwin.cuh): box 16 bytes x tmap_rows rows x 16 planes
    CUtensorMap *d_tmaps, *d_tmaps16, *d_tmaps48;   // d_tmaps16: box of 16 rows (P_Skip trials of phase S); d_tmaps48: 48-byte rows (stage 3)
    int tmap_rows;
    int use_tma;                    // FH264_TMA=0: fill the pixel windows with ordinary loads (development knob)
    int use_bw;                     // phase B kernel: 1 warp-level (phase_bw.cuh), 0 block-level, -1 (default) warp-level on the pipeline lanes only (FH264_PBW)
    uint32_t band_epoch;            // band mode: epoch of the last picture announced at the picture barrier
    int sad_y0, sad_y1;             // luma rows the scene SAD covers: the picture, or this rank's band once the peers' bands are known
    int trace;                      // FH264_TRACE=1: per-lane timing events (fh264_debug_trace)
    int force_miss;                 // FH264_PBW_FORCE_MISS=1: the warp-level phase B treats every phase-S lookup as a miss (test knob)
    bool timed;
    std::vector<void *> allocs;
    // scratch for the stand-alone entry points
    uint8_t *d_scr[3]; int16_t *d_scr16[2]; size_t scr_mbs;
    // band mode
    uint32_t *d_sync;               // FH_MAX_WORLD arrival slots (written by the peers)
    PeerSync peer_sync;             // every rank's sync area (own included)
    std::vector<void *> ipc_opened;
    // device CAVLC (allocated on first use)
    std::vector<CvSeq> cvh;
    CvSeq *d_cvs;
    uint32_t *h_cvstat;             // pinned: batch * 2 (flags, total bits)
    // I pictures (allocated on first use)
    std::vector<IntraSeq> ih;
    IntraSeq *d_is;
    cudaEvent_t ev_i[2];            // around k_intra
    bool intra_timed;
    int *d_iwf_order;               // anti-diagonal order x + 2y: an I macroblock needs left, up-left, up and up-right complete
    int *d_prev_p;                  // per sequence: the previous picture was a P picture whose records are in `results`
    std::vector<int> prev_p;
    std::vector<char> last_i;       // per sequence: `results` holds the I records of the picture coded last
    // scene gate (fh264_encode_p_stream): whether a gated call really coded a P picture is known only once its status is home.
    // Until then: the state before the first unresolved gated call, how many such calls were issued, and the sequence's
    // ST_GATED_TOTAL at that point (resolved in sync_streams).
    std::vector<int> gate_calls, saved_prev_p;
    std::vector<char> saved_last_i;
    std::vector<uint32_t> gated_seen;
};

__global__ void k_begin_picture(SeqDev *seqs, int seq0, uint32_t *ticket)
{
    uint32_t *st = seqs[seq0 + threadIdx.x].status;
    st[ST_FLAGS] = st[ST_FLAGS_NEXT];
    st[ST_GATE] = 0; st[ST_SAD_LO] = 0; st[ST_SAD_HI] = 0;             // (the scene gate of fh264_encode_p_stream accumulates and decides after this)
    for (int i = 0; i < 5; i++) st[ST_COUNTS + i] = 0;
    st[ST_S2REDO] = 0; st[ST_SPEC_HIT] = 0; st[ST_SPEC_MISS] = 0;
    st[ST_NSLOW] = 0;
    if (threadIdx.x == 0) { ticket[0] = 0; ticket[1] = 0; }
}
__global__ void k_begin_ref(SeqDev *seqs, int seq0) { seqs[seq0 + threadIdx.x].status[ST_FLAGS_NEXT] = 0; }
__global__ void k_zero_sad(SeqDev *seqs, int seq0) { uint32_t *st = seqs[seq0 + threadIdx.x].status; st[ST_SAD_LO] = 0; st[ST_SAD_HI] = 0; }
__global__ void k_gather_sad(SeqDev *seqs, int seq0, unsigned long long *out)
{
    const uint32_t *st = seqs[seq0 + threadIdx.x].status;
    out[threadIdx.x] = (unsigned long long)st[ST_SAD_LO] | ((unsigned long long)st[ST_SAD_HI] << 32);
}
// begin_ref: 1 = also what k_begin_ref does (one launch instead of two where no picture barrier sits between them)
__global__ void k_swap_ref(SeqDev *seqs, int seq0, int begin_ref = 0)
{
    SeqDev &S = seqs[seq0 + threadIdx.x];
    if (begin_ref) S.status[ST_FLAGS_NEXT] = 0;
    for (int c = 0; c < 3; c++) { uint8_t *t = S.ref[c]; S.ref[c] = S.rec[c]; S.rec[c] = t; }
    for (int r = 0; r < FH_MAX_WORLD; r++)
        for (int c = 0; c < 3; c++) { uint8_t *t = S.peer_ref[r][c]; S.peer_ref[r][c] = S.peer_rec[r][c]; S.peer_rec[r][c] = t; }
}

// Band mode: all ranks' reconstructed bands must have landed in this rank's picture before phase R reads it.
// One thread: announce "my phase C of picture `epoch` is complete" in every rank's sync area, then wait (bounded) for all.
// A timeout (a peer is missing or far behind: this rank's reference picture lacks that peer's band) is raised in ST_FLAGS_NEXT of
// EVERY sequence of the call — the word phase R leaves alone after k_begin_ref, which therefore runs BEFORE this kernel — and
// becomes the status of the next picture coded from that reference (FH264_E_STATE).
// Waits (bounded) until every rank has announced picture `epoch` (band mode, before an I picture: it is coded whole by every rank and
// reads what ALL bands of the previous P picture left — the per-halo barrier of that picture only waited for the neighbours).
__global__ void k_band_wait_all(PeerSync ps, SeqDev *seqs, int seq0, int nseq, uint32_t epoch, int rank, int world)
{
    bool ok = true;
    for (int r = 0; r < world; r++) ok &= wait_progress(&ps.p[rank][r], epoch, true);
    if (!ok) for (int b = 0; b < nseq; b++) atomicOr(&seqs[seq0 + b].status[ST_FLAGS], FLAG_TIMEOUT);
}

// With the peers' bands known (fh264_band_peers) only the ranks whose bands touch this rank's halo are waited for (wait_mask): a rank
// whose halo is complete goes on with phase R and the next picture while the wavefront of this one is still running further down.
__global__ void k_band_barrier(PeerSync ps, SeqDev *seqs, int seq0, int nseq, uint32_t epoch, int rank, int world, uint32_t wait_mask)
{
    __threadfence_system();
    for (int r = 0; r < world; r++) st_release_sys_u32(&ps.p[r][rank], epoch);
    bool ok = true;
    for (int r = 0; r < world; r++) if ((wait_mask >> r) & 1u) ok &= wait_progress(&ps.p[rank][r], epoch, true);
    if (!ok) for (int b = 0; b < nseq; b++) atomicOr(&seqs[seq0 + b].status[ST_FLAGS_NEXT], FLAG_TIMEOUT);
}

static cudaError_t sync_streams(fh264_session *s)
{
    cudaError_t e = cudaStreamSynchronize(s->stream);
    for (int l = 0; l < 2 && e == cudaSuccess; l++) {
        Lane &L = s->lane[l];
        if (L.st) e = cudaStreamSynchronize(L.st);
        if (e == cudaSuccess && L.hi) e = cudaStreamSynchronize(L.hi);
        if (e == cudaSuccess && L.cp) e = cudaStreamSynchronize(L.cp);
        L.busy = false; L.copy_pending = false;
    }
    if (e == cudaSuccess && s->copy_stream) e = cudaStreamSynchronize(s->copy_stream);
    if (e == cudaSuccess && s->up_stream) e = cudaStreamSynchronize(s->up_stream);
    if (e == cudaSuccess) s->lane[2].copy_pending = false;
    if (e == cudaSuccess)
        for (int b = 0; b < s->batch; b++) {
            if (!s->gate_calls[b]) continue;
            const uint32_t *st = s->h_status + (size_t)b * ST_WORDS;
            // the records are P records unless every call since the last resolve was stopped by the gate
            if (st[ST_GATE_DONE] && st[ST_GATED_TOTAL] - s->gated_seen[b] == (uint32_t)s->gate_calls[b]) { s->prev_p[b] = s->saved_prev_p[b]; s->last_i[b] = s->saved_last_i[b]; }
            s->gated_seen[b] = st[ST_GATED_TOTAL];
            s->gate_calls[b] = 0;
        }
    return e;
}

// Wavefront visiting order x + slope * y. Any slope >= 2 keeps every dependency (left, up, up-right, up-left) on a smaller
// ticket. With phase S in front a macroblock is ~1.2 us of lookups and what bounds the wavefront is the serial chain along the first
// macroblock row (every macroblock there predicts from its left neighbour alone) followed by the chain down the right picture edge:
// the rows below can only trail the first one, so the best order is close to raster order — tickets are not spent on macroblocks
// that wait for the row above. Measured, 8 sequences: slope 2 / 3 / 4 / 6 / 8 / 12 -> phase B 1.27 / 1.14 / 1.09 / 1.04 / 1.02 / 1.03 ms
// (round 1's block-level search had its optimum at 3). FH264_WF_SLOPE overrides.
static void wf_slope(int &num, int &den)      // order key = den * x + num * y, slope = num / den ("8" or "5/2")
{
    num = 8; den = 1;
    if (const char *e = getenv("FH264_WF_SLOPE")) { int a = 0, b = 1; if (sscanf(e, "%d/%d", &a, &b) >= 1 && b >= 1 && a >= 2 * b) { num = a; den = b; } }
}

template <typename T>
static cudaError_t dalloc(fh264_session *s, T **p, size_t count)
{
    void *v = nullptr;
    cudaError_t e = cudaMalloc(&v, count * sizeof(T) + 256);
    if (e == cudaSuccess) { s->allocs.push_back(v); e = cudaMemset(v, 0, count * sizeof(T) + 256); }
    *p = (T *)v;
    return e;
}

extern "C" int fh264_abi_version(void) { return FH264_ABI_VERSION; }
extern "C" const char *fh264_last_error(void) { return g_err.c_str(); }

extern "C" int fh264_close(fh264_session *s)
{
    if (!s) return FH264_OK;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    for (void *p : s->ipc_opened) cudaIpcCloseMemHandle(p);
    for (void *p : s->allocs) cudaFree(p);
    for (int i = 0; i < 3; i++) if (s->d_scr[i]) cudaFree(s->d_scr[i]);
    for (int i = 0; i < 2; i++) if (s->d_scr16[i]) cudaFree(s->d_scr16[i]);
    if (s->h_status) cudaFreeHost(s->h_status);
    if (s->h_sad) cudaFreeHost(s->h_sad);
    if (s->h_cvstat) cudaFreeHost(s->h_cvstat);
    for (int i = 0; i < 5; i++) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    for (int i = 0; i < 4; i++) if (s->evk[i]) cudaEventDestroy(s->evk[i]);
    if (s->ev_spec) cudaEventDestroy(s->ev_spec);
    for (int i = 0; i < 2; i++) if (s->ev_i[i]) cudaEventDestroy(s->ev_i[i]);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->up_stream) cudaStreamDestroy(s->up_stream);
    for (auto e : s->ev_up) cudaEventDestroy(e);
    for (int k = 0; k < 2; k++) for (auto e : s->ev_free[k]) cudaEventDestroy(e);
    for (int l = 0; l < 3; l++) {
        Lane &L = s->lane[l];
        if (l < 2) { if (L.st) cudaStreamDestroy(L.st); if (L.hi) cudaStreamDestroy(L.hi); if (L.cp) cudaStreamDestroy(L.cp); }
        cudaEvent_t *evs[5] = { &L.ev_a, &L.ev_b, &L.ev_c_done, &L.ev_copy_done, &L.ev_done };
        for (auto e : evs) if (*e) cudaEventDestroy(*e);
    }
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    delete s;
    return FH264_OK;
}

extern "C" int fh264_open(int width, int height, int batch, int device, fh264_session **out)
{
    if (!out) return fail(FH264_E_ARG, "out is null");
    *out = nullptr;
    if (width <= 0 || height <= 0 || (width & 15) || (height & 15)) return fail(FH264_E_ARG, "width/height must be positive multiples of 16");
    if ((width >> 4) * (height >> 4) > 10000) return fail(FH264_E_ARG, "more than 10000 macroblocks (reference limit, h264_globals.cpp:180)");
    if (width > 65535 || height > 65535) return fail(FH264_E_ARG, "picture dimension exceeds 65535");
    if (batch < 1 || batch > 1024) return fail(FH264_E_ARG, "batch must be in 1..1024");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(FH264_E_NO_DEVICE, "no CUDA device: fh264_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(FH264_E_ARG, "device ordinal out of range");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(FH264_E_NO_DEVICE, "device is not sm_100 (B200): kernels are built for sm_100a only");
    CK(cudaSetDevice(device));

    fh264_session *s = new fh264_session();
    s->batch = batch; s->device = device; s->epoch = 0; s->timed = false; s->own_stream = true;
    s->d_sync = nullptr; memset(&s->peer_sync, 0, sizeof s->peer_sync);
    s->up_stream = nullptr;
    s->copy_stream = nullptr;
    memset(s->lane, 0, sizeof s->lane);
    s->ev_fork = nullptr; s->main_dirty = true;
    { const char *e = getenv("FH264_PIPE"); s->pipeline = (e && atoi(e) == 1) ? 1 : 0; }
    s->d_cvs = nullptr; s->h_cvstat = nullptr;
    s->d_is = nullptr; s->d_prev_p = nullptr; s->d_iwf_order = nullptr; s->ev_i[0] = s->ev_i[1] = nullptr; s->intra_timed = false;
    s->d_seqs = nullptr; s->h_status = nullptr; s->h_sad = nullptr; s->d_sadout = nullptr; s->scr_mbs = 0;
    for (int i = 0; i < 5; i++) s->ev[i] = nullptr;
    for (int i = 0; i < 4; i++) s->evk[i] = nullptr;
    s->ev_spec = nullptr;
    { const char *e = getenv("FH264_SPEC"); s->use_spec = !(e && atoi(e) == 0); }
    { const char *e = getenv("FH264_TMA"); s->use_tma = !(e && atoi(e) == 0); }
    { const char *e = getenv("FH264_PBW"); s->use_bw = e ? atoi(e) : -1; }       // -1: warp-level phase B on the pipeline lanes only
    { const char *e = getenv("FH264_TRACE"); s->trace = (e && atoi(e) == 1) ? 1 : 0; }
    { const char *e = getenv("FH264_PBW_FORCE_MISS"); s->force_miss = (e && atoi(e) == 1) ? 1 : 0; }
    s->d_tmaps = nullptr; s->d_tmaps16 = nullptr; s->d_tmaps48 = nullptr; s->tmap_rows = 0;
    for (int i = 0; i < 3; i++) s->d_scr[i] = nullptr;
    s->d_scr16[0] = s->d_scr16[1] = nullptr;
    Geo &g = s->g;
    g.W = width; g.H = height; g.Wmb = width >> 4; g.Hmb = height >> 4; g.nmb = g.Wmb * g.Hmb; g.nparts = g.nmb * 4;
    g.tilesx = (width + FH_TILE - 1) / FH_TILE; g.tilesy = (height + FH_TILE - 1) / FH_TILE; g.ntiles = g.tilesx * g.tilesy;
    g.WH = width * height;
    g.band_mb0 = 0; g.band_nmb = g.nmb; g.rank = 0; g.world = 1;
    g.wmb_magic = udiv_magic((uint32_t)g.Wmb);
    g.halo_y0 = 0; g.halo_y1 = height; g.wait_mask = 0xffffffffu; g.gather_on = 0;
    s->sad_y0 = 0; s->sad_y1 = height; s->band_epoch = 0;
    s->has_ref.assign(batch, 0);
    s->prev_p.assign(batch, 0);
    s->last_i.assign(batch, 0);
    s->gate_calls.assign(batch, 0); s->saved_prev_p.assign(batch, 0); s->saved_last_i.assign(batch, 0); s->gated_seen.assign(batch, 0);
    s->h.resize(batch);
#define OPEN_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(FH264_E_CUDA, #call, e_); fh264_close(s); return FH264_E_CUDA; } } while (0)
    OPEN_CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    OPEN_CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    OPEN_CK(cudaStreamCreateWithFlags(&s->up_stream, cudaStreamNonBlocking));
    s->ev_up.assign(batch, nullptr); s->up_pending.assign(batch, 0); s->cur_set.assign(batch, 0);
    for (int k = 0; k < 2; k++) { s->ev_free[k].assign(batch, nullptr); s->free_valid[k].assign(batch, 0); }
    for (int b = 0; b < batch; b++) {
        OPEN_CK(cudaEventCreateWithFlags(&s->ev_up[b], cudaEventDisableTiming));
        for (int k = 0; k < 2; k++) OPEN_CK(cudaEventCreateWithFlags(&s->ev_free[k][b], cudaEventDisableTiming));
    }
    {
        int prio_lo = 0, prio_hi = 0;
        OPEN_CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));           // (numerically lower = higher priority)
        for (int l = 0; l < 3; l++) {
            Lane &L = s->lane[l];
            if (l < 2) {
                OPEN_CK(cudaStreamCreateWithPriority(&L.st, cudaStreamNonBlocking, prio_lo));
                OPEN_CK(cudaStreamCreateWithPriority(&L.hi, cudaStreamNonBlocking, prio_hi));
                OPEN_CK(cudaStreamCreateWithFlags(&L.cp, cudaStreamNonBlocking));
            } else { L.st = s->stream; L.hi = nullptr; L.cp = s->copy_stream; }
            cudaEvent_t *evs[5] = { &L.ev_a, &L.ev_b, &L.ev_c_done, &L.ev_copy_done, &L.ev_done };
            for (auto e : evs) OPEN_CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
            L.copy_pending = false; L.busy = false; L.timing = l == 2;
            L.tr_n = 0;
            if (s->trace) for (int i = 0; i < 8; i++) for (int k = 0; k < 5; k++) OPEN_CK(cudaEventCreate(&L.tr[i][k]));
        }
        OPEN_CK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    }
    for (int i = 0; i < 5; i++) OPEN_CK(cudaEventCreate(&s->ev[i]));
    for (int i = 0; i < 4; i++) OPEN_CK(cudaEventCreate(&s->evk[i]));
    OPEN_CK(cudaEventCreate(&s->ev_spec));
    OPEN_CK(cudaHostAlloc((void **)&s->h_status, sizeof(uint32_t) * ST_WORDS * batch, cudaHostAllocDefault));
    OPEN_CK(cudaHostAlloc((void **)&s->h_sad, sizeof(uint64_t) * batch, cudaHostAllocDefault));
    OPEN_CK(dalloc(s, &s->d_sadout, (size_t)batch));
    memset(s->h_status, 0, sizeof(uint32_t) * ST_WORDS * batch);
    const size_t WH = (size_t)g.WH, CWH = WH / 4;
    fh264_mb_result *results = nullptr;
    OPEN_CK(dalloc(s, &results, (size_t)batch * g.nmb));
    uint32_t *status_block = nullptr;
    OPEN_CK(dalloc(s, &status_block, (size_t)batch * ST_WORDS));
    for (int b = 0; b < batch; b++) {
        SeqDev &S = s->h[b];
        // source pictures: Y | Cb | Cr in one block each (a picture arrives by one copy, fh264_upload_source_batch)
        OPEN_CK(dalloc(s, &S.cur[0], WH + 2 * CWH));
        OPEN_CK(dalloc(s, &S.cur_alt[0], WH + 2 * CWH));
        S.cur[1] = S.cur[0] + WH; S.cur[2] = S.cur[1] + CWH;
        S.cur_alt[1] = S.cur_alt[0] + WH; S.cur_alt[2] = S.cur_alt[1] + CWH;
        for (int c = 0; c < 3; c++) {
            OPEN_CK(dalloc(s, &S.ref[c], c ? CWH : WH));
            OPEN_CK(dalloc(s, &S.rec[c], c ? CWH : WH));
        }
        OPEN_CK(dalloc(s, &S.planes, 16 * WH));
        OPEN_CK(dalloc(s, &S.kar, WH));                  // plane 0 only (16 B per position)
        OPEN_CK(dalloc(s, &S.k0p, WH + 64));
        OPEN_CK(dalloc(s, &S.tent, (size_t)g.ntiles * FH_TILE * FH_TILE));
        OPEN_CK(dalloc(s, &S.tstart, (size_t)g.ntiles * FH_TSTART_PITCH));
        OPEN_CK(dalloc(s, &S.parta, (size_t)g.nparts));
        OPEN_CK(dalloc(s, &S.s3, (size_t)g.nparts * FH_S3_MAX));
        OPEN_CK(dalloc(s, &S.s2pool, (size_t)g.nparts * S2_SLICE));
        OPEN_CK(dalloc(s, &S.motion, (size_t)g.nmb));
        OPEN_CK(dalloc(s, &S.spec, (size_t)g.nparts));
        OPEN_CK(dalloc(s, &S.mbspec, (size_t)g.nmb));
        OPEN_CK(dalloc(s, &S.prev_gen16, (size_t)g.nmb));
        OPEN_CK(cudaMemset(S.prev_gen16, 0x7f, sizeof(uint32_t) * (size_t)g.nmb));
        OPEN_CK(dalloc(s, &S.proxy, (size_t)g.nparts));
        OPEN_CK(cudaMemset(S.proxy, 0x7f, sizeof(uint32_t) * (size_t)g.nparts));
        OPEN_CK(dalloc(s, &S.prev_gen, (size_t)g.nparts));
        OPEN_CK(cudaMemset(S.prev_gen, 0x7f, sizeof(uint32_t) * (size_t)g.nparts));       // SPEC_PREV_NONE: no previous P picture
        OPEN_CK(dalloc(s, &S.qmv, (size_t)g.nmb * 4));
        S.status = status_block + (size_t)b * ST_WORDS;     // one block: a range of sequences is snapshot by one copy
        S.results = results + (size_t)b * g.nmb;
        S.dbg = nullptr;
        memset(S.peer_ref, 0, sizeof S.peer_ref); memset(S.peer_rec, 0, sizeof S.peer_rec); memset(S.peer_motion, 0, sizeof S.peer_motion);
        S.gather[0] = S.gather[1] = nullptr;
        S.peer_qmv_next = nullptr;
    }
    OPEN_CK(dalloc(s, &s->d_seqs, (size_t)batch));
    OPEN_CK(cudaMemcpy(s->d_seqs, s->h.data(), sizeof(SeqDev) * batch, cudaMemcpyHostToDevice));
    // anti-diagonal (x + slope * y) visiting order of the wavefront
    std::vector<int> order(g.nmb);
    for (int i = 0; i < g.nmb; i++) order[i] = i;
    const int Wmb = g.Wmb;
    int sn, sd;
    wf_slope(sn, sd);
    std::stable_sort(order.begin(), order.end(), [Wmb, sn, sd](int a, int b) { return sd * (a % Wmb) + sn * (a / Wmb) < sd * (b % Wmb) + sn * (b / Wmb); });
    OPEN_CK(dalloc(s, &s->d_wf_order, (size_t)g.nmb));
    OPEN_CK(cudaMemcpy(s->d_wf_order, order.data(), sizeof(int) * g.nmb, cudaMemcpyHostToDevice));
    OPEN_CK(dalloc(s, &s->d_ticket, (size_t)8));               // [0..1] session stream, [2..3], [4..5] pipeline lanes, [6] I pictures
    s->lane[2].d_ticket = s->d_ticket; s->lane[0].d_ticket = s->d_ticket + 2; s->lane[1].d_ticket = s->d_ticket + 4;   // pairs: warp-level, block-level kernel
    OPEN_CK(dalloc(s, &s->d_tmaps, (size_t)batch));
    OPEN_CK(dalloc(s, &s->d_tmaps16, (size_t)batch));
    OPEN_CK(dalloc(s, &s->d_tmaps48, (size_t)batch));
    OPEN_CK(dalloc(s, &s->d_sync, (size_t)FH_MAX_WORLD));
    OPEN_CK(cudaFuncSetAttribute(k_tile_index, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_CELLS * 4));
    OPEN_CK(cudaFuncSetAttribute(k_stage3, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    OPEN_CK(cudaFuncSetAttribute(k_spec, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    OPEN_CK(cudaFuncSetAttribute(k_skipspec, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * SKIPWIN_BYTES + 64));
    OPEN_CK(cudaDeviceSynchronize());
    *out = s;
    return FH264_OK;
}

extern "C" int fh264_set_stream(fh264_session *s, void *cuda_stream)
{
    if (!s) return fail(FH264_E_ARG, "null session");
    CK(cudaSetDevice(s->device));
    CK(sync_streams(s));
    if (s->own_stream) { cudaStreamDestroy(s->stream); s->own_stream = false; }
    if (cuda_stream) s->stream = (cudaStream_t)cuda_stream;
    else { CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)); s->own_stream = true; }
    s->lane[2].st = s->stream;
    s->main_dirty = true;
    return FH264_OK;
}

extern "C" int fh264_sync(fh264_session *s)
{
    if (!s) return fail(FH264_E_ARG, "null session");
    CK(cudaSetDevice(s->device));
    CK(sync_streams(s));
    return FH264_OK;
}

extern "C" void *fh264_host_alloc(size_t bytes) { void *p = nullptr; return cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess ? p : nullptr; }
extern "C" void fh264_host_free(void *p) { if (p) cudaFreeHost(p); }

// band mode and every peer's sync area mapped (fh264_ipc_import done for all of them)
static bool band_linked(const fh264_session *s)
{
    if (s->g.world < 2) return false;
    for (int r = 0; r < s->g.world; r++) if (!s->peer_sync.p[r]) return false;
    return true;
}

static int check_seq(fh264_session *s, int seq0, int nseq)
{
    if (!s) return fail(FH264_E_ARG, "null session");
    if (seq0 < 0 || nseq < 1 || seq0 + nseq > s->batch) return fail(FH264_E_ARG, "sequence range outside the session batch");
    return FH264_OK;
}

// Source pictures are double buffered: an upload goes, on its own stream, into the buffer that is NOT being coded (after the
// last picture that read that buffer is through phase C), so the H2D copy of picture t+1 overlaps the coding of picture t.
// The buffers swap when the picture is first used (scene_sad / encode_p), which also makes the coding stream wait for the copy.
static int upload_planes(fh264_session *s, int seq, const void *y, const void *cb, const void *cr, cudaMemcpyKind kind)
{
    const size_t WH = (size_t)s->g.WH;
    const int target = 1 - s->cur_set[seq];
    if (s->free_valid[target][seq]) CK(cudaStreamWaitEvent(s->up_stream, s->ev_free[target][seq], 0));
    uint8_t **dst = s->h[seq].cur_alt;
    CK(cudaMemcpyAsync(dst[0], y, WH, kind, s->up_stream));
    CK(cudaMemcpyAsync(dst[1], cb, WH / 4, kind, s->up_stream));
    CK(cudaMemcpyAsync(dst[2], cr, WH / 4, kind, s->up_stream));
    CK(cudaEventRecord(s->ev_up[seq], s->up_stream));
    s->up_pending[seq] = 1;
    return FH264_OK;
}

__global__ void k_swap_cur(SeqDev *seqs, int seq0, unsigned long long mask)
{
    if (!((mask >> threadIdx.x) & 1ull)) return;
    SeqDev &S = seqs[seq0 + threadIdx.x];
    for (int c = 0; c < 3; c++) { uint8_t *t = S.cur[c]; S.cur[c] = S.cur_alt[c]; S.cur_alt[c] = t; }
}

// Makes the pictures uploaded since the last use current for sequences [seq0, seq0+nseq) (called by everything that reads `cur`).
static int adopt_uploads(fh264_session *s, cudaStream_t st, int seq0, int nseq)
{
    for (int b0 = seq0; b0 < seq0 + nseq; b0 += 64) {
        const int n = std::min(64, seq0 + nseq - b0);
        unsigned long long mask = 0;
        for (int i = 0; i < n; i++) {
            const int b = b0 + i;
            if (!s->up_pending[b]) continue;
            CK(cudaStreamWaitEvent(st, s->ev_up[b], 0));
            mask |= 1ull << i;
            for (int c = 0; c < 3; c++) std::swap(s->h[b].cur[c], s->h[b].cur_alt[c]);
            s->cur_set[b] ^= 1;
            s->up_pending[b] = 0;
        }
        if (mask) k_swap_cur<<<1, n, 0, st>>>(s->d_seqs, b0, mask);
    }
    return FH264_OK;
}

// One Y4M FRAME payload (Y in_w x in_h, then Cb, Cr at half size), centre-cropped to the coded size on the copy engine:
// ReadFromY4M (fileIO.cpp:286-337) — luma rows/columns from ((in - coded) >> 1), chroma from that offset >> 1.
extern "C" int fh264_upload_source_frame(fh264_session *s, int seq, const uint8_t *frame420, int in_w, int in_h)
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    if (!frame420) return fail(FH264_E_ARG, "null frame");
    const int W = s->g.W, H = s->g.H;
    if ((in_w & ~15) != W || (in_h & ~15) != H) return fail(FH264_E_ARG, "input size does not crop to the session's coded size (fileIO.cpp:242-243)");
    CK(cudaSetDevice(s->device));
    const int target = 1 - s->cur_set[seq];
    if (s->free_valid[target][seq]) CK(cudaStreamWaitEvent(s->up_stream, s->ev_free[target][seq], 0));
    uint8_t **dst = s->h[seq].cur_alt;
    const int ct = (in_h - H) >> 1, cl = (in_w - W) >> 1, in_wc = in_w >> 1, in_hc = in_h >> 1;
    const size_t luma = (size_t)in_w * in_h, chroma = luma >> 2;       // as the reference sizes them (:262-263)
    (void)in_hc;
    CK(cudaMemcpy2DAsync(dst[0], W, frame420 + (size_t)ct * in_w + cl, in_w, W, H, cudaMemcpyHostToDevice, s->up_stream));
    CK(cudaMemcpy2DAsync(dst[1], W / 2, frame420 + luma + (size_t)(ct >> 1) * in_wc + (cl >> 1), in_wc, W / 2, H / 2, cudaMemcpyHostToDevice, s->up_stream));
    CK(cudaMemcpy2DAsync(dst[2], W / 2, frame420 + luma + chroma + (size_t)(ct >> 1) * in_wc + (cl >> 1), in_wc, W / 2, H / 2, cudaMemcpyHostToDevice, s->up_stream));
    CK(cudaEventRecord(s->ev_up[seq], s->up_stream));
    s->up_pending[seq] = 1;
    return FH264_OK;
}

extern "C" int fh264_upload_source(fh264_session *s, int seq, const uint8_t *y, const uint8_t *cb, const uint8_t *cr)
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    if (!y || !cb || !cr) return fail(FH264_E_ARG, "null plane");
    CK(cudaSetDevice(s->device));
    return upload_planes(s, seq, y, cb, cr, cudaMemcpyHostToDevice);
}

extern "C" int fh264_upload_source_batch(fh264_session *s, int seq0, int nseq, const void *block, size_t stride, int device)
{
    int rc = check_seq(s, seq0, nseq); if (rc) return rc;
    const size_t pic = (size_t)s->g.WH + (size_t)s->g.WH / 2;
    if (!block || stride < pic) return fail(FH264_E_ARG, "null block or stride smaller than one 4:2:0 picture");
    CK(cudaSetDevice(s->device));
    for (int b = seq0; b < seq0 + nseq; b++) {
        const int target = 1 - s->cur_set[b];
        if (s->free_valid[target][b]) CK(cudaStreamWaitEvent(s->up_stream, s->ev_free[target][b], 0));
    }
    for (int b = seq0; b < seq0 + nseq; b++) {               // Y | Cb | Cr are contiguous on both sides: one copy per picture
        CK(cudaMemcpyAsync(s->h[b].cur_alt[0], (const uint8_t *)block + (size_t)(b - seq0) * stride, pic, device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s->up_stream));
        CK(cudaEventRecord(s->ev_up[b], s->up_stream));
        s->up_pending[b] = 1;
    }
    return FH264_OK;
}

// Phase R on the current dpb of sequences [seq0, seq0+nseq).
static int launch_phase_r(fh264_session *s, cudaStream_t st, bool timing, int seq0, int nseq, bool begin = true)
{
    const Geo &g = s->g;
    if (begin) k_begin_ref<<<1, nseq, 0, st>>>(s->d_seqs, seq0);
    // rows [y0, y1) of the reference picture (everything, or in band mode the band's halo); the planes reach 16 rows further down
    // (the features of the last rows sum 8 rows below them, SADs read 8 rows from a candidate's origin)
    const int y0 = g.halo_y0, y1 = g.halo_y1, yi1 = std::min(g.H, y1 + 16);
    dim3 gi((g.W + IT_W - 1) / IT_W, (yi1 - y0 + IT_H - 1) / IT_H, nseq);
    if (timing) CK(cudaEventRecord(s->evk[1], st));
    k_interp<<<gi, 256, 0, st>>>(s->d_seqs, seq0, g, y0);
    if (timing) CK(cudaEventRecord(s->evk[2], st));
    dim3 gf((g.W + FT_W - 1) / FT_W, (y1 - y0 + FT_H - 1) / FT_H, nseq);
    k_features<<<gf, 256, 0, st>>>(s->d_seqs, seq0, g, 0, nullptr, y0);
    if (timing) CK(cudaEventRecord(s->evk[3], st));
    const int ty0 = y0 / FH_TILE, ty1 = (y1 + FH_TILE - 1) / FH_TILE;
    dim3 gt((ty1 - ty0) * g.tilesx, nseq);
    k_tile_index<<<gt, 256, FH_CELLS * 4, st>>>(s->d_seqs, seq0, g, ty0 * g.tilesx);
    CKL();
    return FH264_OK;
}

// Tensor maps over the interpolated planes [16][H][W] of every sequence with a box of 32 bytes x `rows` rows x 16 planes
// (re-encoded when the window geometry changes; stream-ordered copy, so launches in flight keep the maps they were issued with).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int ensure_tmaps(fh264_session *s, int rows)
{
    if (!s->use_tma || s->tmap_rows == rows) return FH264_OK;
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn || qr != cudaDriverEntryPointSuccess) return fail(FH264_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        enc = (EncodeTiledFn)fn;
    }
    const Geo &g = s->g;
    CK(sync_streams(s));                                // window geometry changed: nothing in flight may still read the old maps
    s->main_dirty = true;
    std::vector<CUtensorMap> maps(s->batch);
    const int g1w = (rows - 9) / 2;
    for (int pass = s->tmap_rows == 0 ? 0 : 1; pass < 3; pass++) {          // first call: the 16-row maps too
        const int brows = pass == 0 ? 16 : (pass == 1 ? rows : s3_win_rows(g1w));
        for (int b = 0; b < s->batch; b++) {
            const cuuint64_t dims[3] = { (cuuint64_t)g.W, (cuuint64_t)g.H, 16 };
            const cuuint64_t strides[2] = { (cuuint64_t)g.W, (cuuint64_t)g.WH };
            const cuuint32_t box[3] = { (cuuint32_t)(pass == 2 ? S3_ROWB : QW_ROWB), (cuuint32_t)brows, 16 }, estr[3] = { 1, 1, 1 };
            const CUresult r = enc(&maps[b], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, s->h[b].planes, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(FH264_E_CUDA, "cuTensorMapEncodeTiled failed");
        }
        CK(cudaMemcpyAsync(pass == 0 ? s->d_tmaps16 : (pass == 1 ? s->d_tmaps : s->d_tmaps48), maps.data(), sizeof(CUtensorMap) * s->batch, cudaMemcpyHostToDevice, s->stream));
        CK(cudaStreamSynchronize(s->stream));          // `maps` is a local
    }
    s->tmap_rows = rows;
    return FH264_OK;
}

extern "C" int fh264_upload_recon(fh264_session *s, int seq, const uint8_t *y, const uint8_t *cb, const uint8_t *cr)
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    if (!y || !cb || !cr) return fail(FH264_E_ARG, "null plane");
    CK(cudaSetDevice(s->device));
    rc = enter_main(s); if (rc) return rc;
    const size_t WH = (size_t)s->g.WH;
    // band mode: the per-halo picture barrier lets this rank run ahead of ranks that are still storing their band of the previous
    // picture into the very buffer the upload replaces — all of them are waited for first
    if (band_linked(s)) k_band_wait_all<<<1, 1, 0, s->stream>>>(s->peer_sync, s->d_seqs, seq, 1, s->band_epoch, s->g.rank, s->g.world);
    CK(cudaMemcpyAsync(s->h[seq].ref[0], y, WH, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(s->h[seq].ref[1], cb, WH / 4, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(s->h[seq].ref[2], cr, WH / 4, cudaMemcpyHostToDevice, s->stream));
    rc = launch_phase_r(s, s->stream, true, seq, 1); if (rc) return rc;
    s->has_ref[seq] = 1;
    if (!s->prev_p.empty()) { s->prev_p[seq] = 0; s->last_i[seq] = 0; }     // a picture coded elsewhere: no P_Skip entries in mb_type_array
    return FH264_OK;
}

extern "C" int fh264_scene_sad_batch(fh264_session *s, int seq0, int nseq, uint64_t *sads)
{
    int rc = check_seq(s, seq0, nseq); if (rc) return rc;
    if (!sads) return fail(FH264_E_ARG, "null output");
    for (int b = seq0; b < seq0 + nseq; b++)
        if (!s->has_ref[b]) return fail(FH264_E_STATE, "scene_sad before any reference picture (dpb.L == NULL => IDR, ref_frames.cpp:191)");
    CK(cudaSetDevice(s->device));
    rc = enter_main(s); if (rc) return rc;
    rc = adopt_uploads(s, s->stream, seq0, nseq); if (rc) return rc;
    k_zero_sad<<<1, nseq, 0, s->stream>>>(s->d_seqs, seq0);
    dim3 gs(nseq >= 8 ? 74 : 296, nseq);
    k_scene_sad<<<gs, 256, 0, s->stream>>>(s->d_seqs, seq0, s->g, s->sad_y0, s->sad_y1);
    k_gather_sad<<<1, nseq, 0, s->stream>>>(s->d_seqs, seq0, s->d_sadout);
    CKL();
    CK(cudaMemcpyAsync(s->h_sad, s->d_sadout, sizeof(uint64_t) * nseq, cudaMemcpyDeviceToHost, s->stream));
    CK(sync_streams(s));
    memcpy(sads, s->h_sad, sizeof(uint64_t) * nseq);
    return FH264_OK;
}

extern "C" int fh264_scene_sad(fh264_session *s, int seq, uint64_t *sad) { return fh264_scene_sad_batch(s, seq, 1, sad); }

extern "C" int fh264_upload_source_device(fh264_session *s, int seq, const void *dy, const void *dcb, const void *dcr)
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    if (!dy || !dcb || !dcr) return fail(FH264_E_ARG, "null plane");
    CK(cudaSetDevice(s->device));
    return upload_planes(s, seq, dy, dcb, dcr, cudaMemcpyDeviceToDevice);
}

// ---- lanes --------------------------------------------------------------------------------------------------------------------
// Everything but the pipelined encode works on the session stream: it first waits for whatever the lanes still have in flight.
static int enter_main(fh264_session *s)
{
    for (int l = 0; l < 2; l++) {
        Lane &L = s->lane[l];
        if (!L.busy) continue;
        CK(cudaStreamWaitEvent(s->stream, L.ev_done, 0));
        if (L.copy_pending) { CK(cudaStreamWaitEvent(s->stream, L.ev_copy_done, 0)); L.copy_pending = false; }
        L.busy = false;
    }
    s->main_dirty = true;
    return FH264_OK;
}
// The lanes start from the state the session stream has reached (once per batch of work enqueued there).
static int fork_lanes(fh264_session *s)
{
    if (!s->main_dirty) return FH264_OK;
    CK(cudaEventRecord(s->ev_fork, s->stream));
    for (int l = 0; l < 2; l++) CK(cudaStreamWaitEvent(s->lane[l].st, s->ev_fork, 0));
    if (s->lane[2].copy_pending) for (int l = 0; l < 2; l++) CK(cudaStreamWaitEvent(s->lane[l].st, s->lane[2].ev_copy_done, 0));
    s->main_dirty = false;
    return FH264_OK;
}

// selectNALUnitType's scene-change rule on the device (ref_frames.cpp:210-224): sum |cur Y - dpb Y| > MBs << 12 -> the picture
// must be an IDR picture; the P kernels of the call then leave the sequence alone (ST_GATE).
__global__ void k_scene_gate(SeqDev *seqs, int seq0, unsigned long long thr)
{
    uint32_t *st = seqs[seq0 + threadIdx.x].status;
    const unsigned long long sad = (unsigned long long)st[ST_SAD_LO] | ((unsigned long long)st[ST_SAD_HI] << 32);
    st[ST_GATE] = sad > thr ? 1u : 0u;
}

struct StreamOut {                  // what one lane sends home after phase C (all optional, pinned host memory)
    fh264_mb_result *records;       // nseq * nmb
    int cavlc;                      // entropy-code the slice data on the device
    int first_bit;
    uint8_t *slice; size_t slice_stride, slice_copy;
    uint32_t *slice_stat;           // 2 words per sequence: flags (CV_FLAG_*), total bits
    fh264_cavlc_mb_info *mb_info;   // nseq * nmb
    uint32_t *status;               // ST_WORDS per sequence (scene SAD, gate, mode counts, flags)
};

static int ensure_cavlc(fh264_session *s);

// One P picture of sequences [seq0, seq0 + nseq) on lane L. gate: scene-change rule decided on the device first.
// hold: event the lane's phase A waits for (the pipeline's stagger), or null.
static int encode_lane(fh264_session *s, Lane &L, int seq0, int nseq, const fh264_params &prm, int gate, cudaEvent_t hold, const StreamOut &o)
{
    const Geo &g = s->g;
    cudaStream_t st = L.st;
    int rc = adopt_uploads(s, st, seq0, nseq); if (rc) return rc;
    if (L.timing) CK(cudaEventRecord(s->ev[0], st));
    k_begin_picture<<<1, nseq, 0, st>>>(s->d_seqs, seq0, L.d_ticket);
    if (gate) {
        k_scene_sad<<<dim3(nseq >= 8 ? 74 : 296, nseq), 256, 0, st>>>(s->d_seqs, seq0, g, s->sad_y0, s->sad_y1);
        if (gate == 1) k_scene_gate<<<1, nseq, 0, st>>>(s->d_seqs, seq0, (unsigned long long)g.nmb << 12);
    }
    if (hold) CK(cudaStreamWaitEvent(st, hold, 0));
    cudaEvent_t *tr = s->trace ? L.tr[L.tr_n++ & 7] : nullptr;
    if (tr) CK(cudaEventRecord(tr[0], st));
    if (prm.basic && L.timing) CK(cudaEventRecord(s->evk[0], st));
    const int g1 = prm.window / 16;
    const WinMagic wm = { udiv_magic((uint32_t)(2 * (prm.window / 2) + 1)), udiv_magic((uint32_t)(2 * g1 + 1)) };
    if (!prm.basic) {
        const size_t smem3 = (size_t)s3_win_bytes(g1) + 16 + 4 * sizeof(S3WarpV2);
        dim3 g3d(g.band_nmb, nseq), g2d(g.band_nmb * 2, nseq);          // 4 / 2 partitions per CTA
        k_stage3<<<g3d, 128, smem3, st>>>(s->d_seqs, seq0, g, prm, wm, s->use_tma ? s->d_tmaps48 : nullptr);
        if (L.timing) CK(cudaEventRecord(s->evk[0], st));
        k_stage2<2><<<g2d, 64, 0, st>>>(s->d_seqs, seq0, g, prm);
    }
    if (L.timing) CK(cudaEventRecord(s->ev_spec, st));
    if (s->use_spec) {
        // phase S: the search completed for the guessed integer predictors (spec.cuh)
        const size_t smems = 4 * (size_t)qwin_bytes(g1) + 4 * sizeof(SpecWarp) + 16;
        static const int spec_flags = 1 | ((getenv("FH264_SPEC_NG") && atoi(getenv("FH264_SPEC_NG")) == 1) ? 2 : 0);   // development knob
        k_spec<<<dim3(g.band_nmb, nseq), 128, smems, st>>>(s->d_seqs, seq0, g, prm, spec_flags, wm, s->use_tma ? s->d_tmaps : nullptr);
        k_skipspec<<<dim3((g.band_nmb + 3) / 4, nseq), 128, 4 * SKIPWIN_BYTES + 64, st>>>(s->d_seqs, seq0, g, prm, s->use_tma ? s->d_tmaps16 : nullptr);
        CKL();
    }
    if (L.timing) CK(cudaEventRecord(s->ev[1], st));
    CK(cudaEventRecord(L.ev_a, st));
    if (tr) CK(cudaEventRecord(tr[1], st));
    // persistent wavefront CTAs: two anti-diagonals' worth per sequence — one set working, one set that has already
    // drawn its ticket and prefetched (otherwise ticket + prefetch latency sits on the wavefront's critical path)
    unsigned pb_ctas = (unsigned)std::min<long long>((long long)g.band_nmb * nseq, (long long)nseq * (g.Wmb + 16));
    { static const char *e = getenv("FH264_PB_CTAS"); if (e && atoi(e) > 0) pb_ctas = (unsigned)std::min<long long>((long long)g.band_nmb * nseq, atoll(e)); }   // development knob
    cudaStream_t sb = L.hi ? L.hi : st;
    if (L.hi) CK(cudaStreamWaitEvent(sb, L.ev_a, 0));
    // the warp-level kernel (phase_bw.cuh) codes every picture phase S prepared; k_phase_b only those that need the block-level search
    const int bw = s->use_spec && (s->use_bw < 0 ? L.hi != nullptr : s->use_bw != 0);
    if (bw) k_phase_b_warp<<<pb_ctas, 32, pbw_smem_bytes(g1), sb>>>(s->d_seqs, seq0, nseq, g, prm, s->epoch, s->d_wf_order, L.d_ticket, wm, s->use_tma ? s->d_tmaps : nullptr, s->force_miss);
    k_phase_b<<<pb_ctas, PB_NT, 0, sb>>>(s->d_seqs, seq0, nseq, g, prm, s->epoch, s->d_wf_order, L.d_ticket + 1, s->use_spec, bw);
    if (tr) CK(cudaEventRecord(tr[2], sb));
    if (L.hi) { CK(cudaEventRecord(L.ev_b, sb)); CK(cudaStreamWaitEvent(st, L.ev_b, 0)); }
    if (L.timing) CK(cudaEventRecord(s->ev[2], st));
    dim3 gc((g.band_nmb + 3) / 4, nseq);
    if (L.copy_pending) { CK(cudaStreamWaitEvent(st, L.ev_copy_done, 0)); L.copy_pending = false; }   // records of the previous picture are home
    k_phase_c<<<gc, 128, 0, st>>>(s->d_seqs, seq0, g, prm, s->epoch);
    CKL();
    if (tr) CK(cudaEventRecord(tr[3], st));
    for (int b = seq0; b < seq0 + nseq; b++) { CK(cudaEventRecord(s->ev_free[(int)s->cur_set[b]][b], st)); s->free_valid[(int)s->cur_set[b]][b] = 1; }   // `cur` is free for the upload after next
    if (L.timing) CK(cudaEventRecord(s->ev[3], st));
    CK(cudaMemcpyAsync(s->h_status + (size_t)seq0 * ST_WORDS, s->h[seq0].status, sizeof(uint32_t) * ST_WORDS * nseq, cudaMemcpyDeviceToHost, st));
    // (the caller's snapshot too is taken in stream order: the next picture's k_begin_picture resets the counters)
    if (o.status) CK(cudaMemcpyAsync(o.status, s->h[seq0].status, sizeof(uint32_t) * ST_WORDS * nseq, cudaMemcpyDeviceToHost, st));
    if (o.records || o.cavlc)
    {
        // everything that goes home travels on the lane's copy stream while the coding stream goes on with the dpb swap, phase R
        // and the next picture; the entropy coder (cavlc.cuh) reads the records phase C left, so it runs there too
        CK(cudaEventRecord(L.ev_c_done, st));
        CK(cudaStreamWaitEvent(L.cp, L.ev_c_done, 0));
        if (o.records) {
            if (g.world == 1) CK(cudaMemcpyAsync(o.records, s->h[seq0].results, sizeof(fh264_mb_result) * (size_t)g.nmb * nseq, cudaMemcpyDeviceToHost, L.cp));
            else for (int b = 0; b < nseq; b++)      // band mode: only this rank's band of every sequence is valid
                CK(cudaMemcpyAsync(o.records + (size_t)b * g.nmb + g.band_mb0, s->h[seq0 + b].results + g.band_mb0, sizeof(fh264_mb_result) * (size_t)g.band_nmb, cudaMemcpyDeviceToHost, L.cp));
        }
        if (o.cavlc) {
            const int nmb = g.nmb, wmb = g.Wmb;
            for (int b = seq0; b < seq0 + nseq; b++) CK(cudaMemsetAsync(s->cvh[b].stream, 0, (size_t)CV_STREAM_BYTES + 64, L.cp));
            // band mode (rank 0): the slice is coded from the records every rank's phase C gathered here — once all of them are through
            const int par = g.world > 1 ? (int)(s->epoch & 1u) : -1;
            if (g.world > 1) k_band_wait_all<<<1, 1, 0, L.cp>>>(s->peer_sync, s->d_seqs, seq0, nseq, s->epoch, g.rank, g.world);
            k_cavlc_prep<<<dim3((nmb + 127) / 128, nseq), 128, 0, L.cp>>>(s->d_seqs, s->d_cvs, seq0, nmb, par);
            k_cavlc_code<<<dim3((nmb + 1 + 127) / 128, nseq), 128, 0, L.cp>>>(s->d_seqs, s->d_cvs, seq0, nmb, wmb, par);
            k_cavlc_scan<<<nseq, 1024, 0, L.cp>>>(s->d_cvs, seq0, nmb, o.first_bit);
            k_cavlc_pack<<<dim3((nmb + 1 + 3) / 4, nseq), 128, 0, L.cp>>>(s->d_cvs, seq0, nmb, o.first_bit);
            CKL();
            for (int b = 0; b < nseq; b++) {
                CK(cudaMemcpyAsync(o.slice_stat + 2 * b, s->cvh[seq0 + b].stat, sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost, L.cp));
                CK(cudaMemcpyAsync(o.slice + (size_t)b * o.slice_stride, s->cvh[seq0 + b].stream, o.slice_copy, cudaMemcpyDeviceToHost, L.cp));
                if (o.mb_info) CK(cudaMemcpyAsync(o.mb_info + (size_t)b * nmb, s->cvh[seq0 + b].info, sizeof(CvInfo) * (size_t)nmb, cudaMemcpyDeviceToHost, L.cp));
            }
        }
        CK(cudaEventRecord(L.ev_copy_done, L.cp));
        L.copy_pending = true;
    }
    // dpb := reconstruction (frameDeepCopy, ref_frames.cpp:17-35) by pointer swap, then phase R for the next picture
    if (g.world > 1) {
        k_begin_ref<<<1, nseq, 0, st>>>(s->d_seqs, seq0);      // before the barrier: a barrier timeout must survive into the next picture's status
        k_band_barrier<<<1, 1, 0, st>>>(s->peer_sync, s->d_seqs, seq0, nseq, s->epoch, g.rank, g.world, g.wait_mask);
        s->band_epoch = s->epoch;
    }
    k_swap_ref<<<1, nseq, 0, st>>>(s->d_seqs, seq0, g.world == 1);
    for (int b = seq0; b < seq0 + nseq; b++)
        {
            for (int c = 0; c < 3; c++) std::swap(s->h[b].ref[c], s->h[b].rec[c]);
            for (int r = 0; r < FH_MAX_WORLD; r++) for (int c = 0; c < 3; c++) std::swap(s->h[b].peer_ref[r][c], s->h[b].peer_rec[r][c]);
        }
    rc = launch_phase_r(s, st, L.timing, seq0, nseq, false); if (rc) return rc;
    if (L.timing) { CK(cudaEventRecord(s->ev[4], st)); s->timed = true; }
    if (tr) CK(cudaEventRecord(tr[4], st));
    CK(cudaEventRecord(L.ev_done, st));
    L.busy = true;
    for (int b = seq0; b < seq0 + nseq; b++) {
        if (gate == 1) { if (!s->gate_calls[b]) { s->saved_prev_p[b] = s->prev_p[b]; s->saved_last_i[b] = s->last_i[b]; } s->gate_calls[b]++; }
        else s->gate_calls[b] = 0;
        s->prev_p[b] = 1; s->last_i[b] = 0;
    }
    return FH264_OK;
}

static int check_encode_p(fh264_session *s, int seq0, int nseq, const fh264_params *p)
{
    int rc = check_seq(s, seq0, nseq); if (rc) return rc;
    if (!p) return fail(FH264_E_ARG, "null params");
    if (p->qp < 0 || p->qp > 51) return fail(FH264_E_ARG, "qp outside 0..51");
    if (p->window < 0 || p->window > FH_MAX_WINDOW) return fail(FH264_E_UNSUPPORTED, "WindowSize above 64 is not supported");
    if (p->maxdiff_set < -1) return fail(FH264_E_ARG, "maxdiff_set below -1");
    for (int b = seq0; b < seq0 + nseq; b++)
        if (!s->has_ref[b]) return fail(FH264_E_STATE, "encode_p before any reference picture (upload_recon first)");
    return FH264_OK;
}

// The call's sequences on the session stream (plain) or split over the two pipeline lanes.
static int encode_dispatch(fh264_session *s, int seq0, int nseq, const fh264_params *p, int gate, bool piped, const StreamOut &o)
{
    CK(cudaSetDevice(s->device));
    if (s->g.world > 1 && !band_linked(s)) return fail(FH264_E_STATE, "band mode: fh264_ipc_import every peer before coding");
    fh264_params prm = *p;
    prm.basic = prm.basic ? 1 : 0;
    s->epoch++;
    int rc = ensure_tmaps(s, 8 + 2 * (prm.window / 16) + 1); if (rc) return rc;
    if (o.cavlc) { rc = ensure_cavlc(s); if (rc) return rc; }
    if (!piped || nseq < 2 || s->g.world > 1) {
        rc = enter_main(s); if (rc) return rc;
        return encode_lane(s, s->lane[2], seq0, nseq, prm, gate, nullptr, o);
    }
    rc = fork_lanes(s); if (rc) return rc;
    const int n0 = (nseq + 1) / 2, n1 = nseq - n0, nmb = s->g.nmb;
    StreamOut o1 = o;
    if (o.records) o1.records = o.records + (size_t)n0 * nmb;
    if (o.slice) o1.slice = o.slice + (size_t)n0 * o.slice_stride;
    if (o.slice_stat) o1.slice_stat = o.slice_stat + 2 * n0;
    if (o.mb_info) o1.mb_info = o.mb_info + (size_t)n0 * nmb;
    if (o.status) o1.status = o.status + (size_t)n0 * ST_WORDS;
    rc = encode_lane(s, s->lane[0], seq0, n0, prm, gate, nullptr, o); if (rc) return rc;
    return encode_lane(s, s->lane[1], seq0 + n0, n1, prm, gate, s->lane[0].ev_a, o1);
}

extern "C" int fh264_encode_p_async(fh264_session *s, int seq0, int nseq, const fh264_params *p, fh264_mb_result *results)
{
    int rc = check_encode_p(s, seq0, nseq, p); if (rc) return rc;
    StreamOut o;
    memset(&o, 0, sizeof o);
    o.records = results;
    return encode_dispatch(s, seq0, nseq, p, 0, s->pipeline != 0, o);
}

extern "C" int fh264_encode_p(fh264_session *s, int seq0, int nseq, const fh264_params *p, fh264_mb_result *results)
{
    int rc = fh264_encode_p_async(s, seq0, nseq, p, results); if (rc) return rc;
    CK(sync_streams(s));
    for (int b = seq0; b < seq0 + nseq; b++) { rc = fh264_picture_status(s, b); if (rc) return rc; }
    return FH264_OK;
}

// Development tap (FH264_TRACE=1): out[lane][picture][5] = ms since the oldest traced phase-A start of lane 0, for the last 8
// pictures of the three lanes (0, 1: pipeline halves; 2: session stream); -1 where nothing was recorded. Synchronises.
extern "C" int fh264_debug_trace(fh264_session *s, float *out)
{
    if (!s || !out) return fail(FH264_E_ARG, "null argument");
    if (!s->trace) return fail(FH264_E_STATE, "FH264_TRACE=1 was not set when the session was opened");
    CK(cudaSetDevice(s->device));
    CK(sync_streams(s));
    for (int i = 0; i < 3 * 8 * 5; i++) out[i] = -1.0f;
    cudaEvent_t base = nullptr;
    for (int l = 0; l < 3 && !base; l++) { const Lane &L = s->lane[l]; if (L.tr_n > 0) base = L.tr[(L.tr_n >= 8 ? L.tr_n : 0) & 7][0]; }
    if (!base) return FH264_OK;
    for (int l = 0; l < 3; l++) {
        const Lane &L = s->lane[l];
        const int n = std::min(L.tr_n, 8), first = L.tr_n - n;
        for (int i = 0; i < n; i++)
            for (int k = 0; k < 5; k++) { float ms = 0; if (cudaEventElapsedTime(&ms, base, L.tr[(first + i) & 7][k]) == cudaSuccess) out[(l * 8 + i) * 5 + k] = ms; else cudaGetLastError(); }
    }
    return FH264_OK;
}

extern "C" int fh264_set_pipeline(fh264_session *s, int on)
{
    if (!s) return fail(FH264_E_ARG, "null session");
    s->pipeline = on ? 1 : 0;
    return FH264_OK;
}

// One step of a batch of sequences without a host round trip. See the header.
extern "C" int fh264_encode_p_stream(fh264_session *s, int seq0, int nseq, const fh264_params *p, int scene_gate, const fh264_stream_out *out)
{
    int rc = check_encode_p(s, seq0, nseq, p); if (rc) return rc;
    if (scene_gate == 1 && s->g.world > 1) return fail(FH264_E_UNSUPPORTED, "the device-side scene gate is not available in band mode (scene_gate = 2 measures only)");
    if (scene_gate < 0 || scene_gate > 2) return fail(FH264_E_ARG, "scene_gate must be 0, 1 or 2");
    StreamOut o;
    memset(&o, 0, sizeof o);
    if (out) {
        o.records = out->records; o.status = out->status;
        if (out->slice) {
            if (s->g.world > 1 && (s->g.rank != 0 || !s->g.gather_on)) return fail(FH264_E_UNSUPPORTED, "band mode: the slice is entropy-coded on rank 0, and only with the record gather on (fh264_band_gather)");
            if (!out->slice_stat || out->first_bit < 0 || out->first_bit > 7 || out->slice_copy_bytes > out->slice_stride || out->slice_copy_bytes > (size_t)CV_STREAM_BYTES)
                return fail(FH264_E_ARG, "slice output: slice_stat missing, first_bit outside 0..7 or copy size above the stride / 500000");
            o.cavlc = 1; o.first_bit = out->first_bit; o.slice = out->slice; o.slice_stride = out->slice_stride; o.slice_copy = out->slice_copy_bytes;
            o.slice_stat = out->slice_stat; o.mb_info = out->mb_info;
        }
    }
    return encode_dispatch(s, seq0, nseq, p, scene_gate, s->pipeline != 0, o);
}

// Decoder inverse path: reconstructs the P picture described by `records` (quadrant MVs, mb_type,
// levels — what the reference's decoder holds per macroblock after entropy decoding, rbsp_decoding.cpp:98-109,330-346) from the
// current reference picture, then makes it the reference (dpb swap + phase R) exactly as fh264_encode_p does.
extern "C" int fh264_decode_p(fh264_session *s, int seq0, int nseq, int qp, const fh264_mb_result *records)
{
    int rc = check_seq(s, seq0, nseq); if (rc) return rc;
    if (!records) return fail(FH264_E_ARG, "null records");
    if (qp < 0 || qp > 51) return fail(FH264_E_ARG, "qp outside 0..51");
    if (s->g.world > 1) return fail(FH264_E_UNSUPPORTED, "decode_p is not available in band mode");
    for (int b = seq0; b < seq0 + nseq; b++)
        if (!s->has_ref[b]) return fail(FH264_E_STATE, "decode_p before any reference picture (upload_recon first)");
    CK(cudaSetDevice(s->device));
    rc = enter_main(s); if (rc) return rc;
    const Geo &g = s->g;
    cudaStream_t st = s->stream;
    if (s->lane[2].copy_pending) { CK(cudaStreamWaitEvent(st, s->lane[2].ev_copy_done, 0)); s->lane[2].copy_pending = false; }
    CK(cudaMemcpyAsync(s->h[seq0].results, records, sizeof(fh264_mb_result) * (size_t)g.nmb * nseq, cudaMemcpyHostToDevice, st));
    k_decode_p<<<dim3((g.nmb + 3) / 4, nseq), 128, 0, st>>>(s->d_seqs, seq0, g, qp);
    CKL();
    k_swap_ref<<<1, nseq, 0, st>>>(s->d_seqs, seq0);
    for (int b = seq0; b < seq0 + nseq; b++)
        for (int c = 0; c < 3; c++) std::swap(s->h[b].ref[c], s->h[b].rec[c]);
    rc = launch_phase_r(s, st, true, seq0, nseq); if (rc) return rc;
    CK(sync_streams(s));
    if (!s->prev_p.empty()) for (int b = seq0; b < seq0 + nseq; b++) { s->prev_p[b] = 1; s->last_i[b] = 0; }
    return FH264_OK;
}

extern "C" int fh264_last_intra_ms(fh264_session *s, float *ms)
{
    if (!s || !ms) return fail(FH264_E_ARG, "null argument");
    if (!s->intra_timed) return fail(FH264_E_STATE, "no encode_i yet");
    CK(cudaSetDevice(s->device));
    CK(cudaEventSynchronize(s->ev_i[1]));
    CK(cudaEventElapsedTime(ms, s->ev_i[0], s->ev_i[1]));
    return FH264_OK;
}

// I pictures on the device (intra.cuh). See the header.
static int ensure_intra(fh264_session *s)
{
    if (s->d_is) return FH264_OK;
    const size_t nmb = (size_t)s->g.nmb;
    s->ih.assign(s->batch, IntraSeq());
    for (int b = 0; b < s->batch; b++) {
        CK(dalloc(s, &s->ih[b].info, nmb));
        CK(dalloc(s, &s->ih[b].done, nmb));
        CK(cudaMemset(s->ih[b].done, 0, sizeof(uint32_t) * nmb));
    }
    CK(dalloc(s, &s->d_prev_p, (size_t)s->batch));
    {
        // every dependency of a macroblock has a smaller key x + 2y (up-right: -1), so it holds a smaller ticket
        std::vector<int> order(nmb);
        for (size_t i = 0; i < nmb; i++) order[i] = (int)i;
        const int Wmb = s->g.Wmb;
        std::stable_sort(order.begin(), order.end(), [Wmb](int a, int b) { return (a % Wmb) + 2 * (a / Wmb) < (b % Wmb) + 2 * (b / Wmb); });
        CK(dalloc(s, &s->d_iwf_order, nmb));
        CK(cudaMemcpy(s->d_iwf_order, order.data(), sizeof(int) * nmb, cudaMemcpyHostToDevice));
    }
    for (int i = 0; i < 2; i++) CK(cudaEventCreate(&s->ev_i[i]));
    CK(dalloc(s, &s->d_is, (size_t)s->batch));
    CK(cudaMemcpy(s->d_is, s->ih.data(), sizeof(IntraSeq) * s->batch, cudaMemcpyHostToDevice));
    return FH264_OK;
}

extern "C" int fh264_encode_i(fh264_session *s, int seq0, int nseq, int qp, fh264_mb_result_i *results)
{
    int rc = check_seq(s, seq0, nseq); if (rc) return rc;
    if (qp < 0 || qp > 51) return fail(FH264_E_ARG, "qp outside 0..51");
    CK(cudaSetDevice(s->device));
    if (s->g.world > 1 && !band_linked(s)) return fail(FH264_E_STATE, "band mode: fh264_ipc_import every peer before coding");
    for (int b = seq0; b < seq0 + nseq; b++) if (s->gate_calls[b]) { CK(sync_streams(s)); break; }      // prev_p of a gated call is known once its status is home
    rc = enter_main(s); if (rc) return rc;
    rc = ensure_intra(s); if (rc) return rc;
    const Geo &g = s->g;
    cudaStream_t st = s->stream;
    rc = adopt_uploads(s, st, seq0, nseq); if (rc) return rc;
    if (s->lane[2].copy_pending) { CK(cudaStreamWaitEvent(st, s->lane[2].ev_copy_done, 0)); s->lane[2].copy_pending = false; }   // the I records reuse the result buffer
    s->epoch++;
    // band mode: an I picture is not split — every rank codes the whole picture (it holds the whole source and, once all ranks have
    // delivered the previous picture, everything an I picture reads across pictures); no exchange, identical results everywhere
    CK(cudaMemcpyAsync(s->d_prev_p + seq0, s->prev_p.data() + seq0, sizeof(int) * nseq, cudaMemcpyHostToDevice, st));
    k_begin_intra<<<1, nseq, 0, st>>>(s->d_ticket + 6, s->d_seqs, seq0);
    if (g.world > 1) k_band_wait_all<<<1, 1, 0, st>>>(s->peer_sync, s->d_seqs, seq0, nseq, s->band_epoch, g.rank, g.world);   // (a timeout fails this picture)
    int nl = 32;
    { const char *e = getenv("FH264_INTRA_LANES"); if (e && atoi(e) == 1) nl = 1; }   // development knob (read per call): everything on lane 0
    const unsigned ctas = (unsigned)std::min<long long>((long long)g.nmb * nseq, (long long)nseq * (g.Wmb + 16));
    CK(cudaEventRecord(s->ev_i[0], st));
    k_intra<<<ctas, 32, 0, st>>>(s->d_seqs, s->d_is, s->d_prev_p, seq0, nseq, g, qp, s->epoch, s->d_iwf_order, s->d_ticket + 6, nl);
    CKL();
    CK(cudaEventRecord(s->ev_i[1], st));
    s->intra_timed = true;
    for (int b = seq0; b < seq0 + nseq; b++) { CK(cudaEventRecord(s->ev_free[(int)s->cur_set[b]][b], st)); s->free_valid[(int)s->cur_set[b]][b] = 1; }
    CK(cudaMemcpyAsync(s->h_status + (size_t)seq0 * ST_WORDS, s->h[seq0].status, sizeof(uint32_t) * ST_WORDS * nseq, cudaMemcpyDeviceToHost, st));
    if (results) CK(cudaMemcpyAsync(results, s->h[seq0].results, sizeof(fh264_mb_result_i) * (size_t)g.nmb * nseq, cudaMemcpyDeviceToHost, st));
    // dpb := reconstruction, then phase R for the next picture (rbsp_encoding.cpp:317-322)
    if (g.world > 1) {
        // nobody may start the next P picture (whose phase C mirrors macroblock types into every rank) before every rank is through this one
        k_begin_ref<<<1, nseq, 0, st>>>(s->d_seqs, seq0);          // (before the barrier: its timeout must survive into the next picture's status)
        k_band_barrier<<<1, 1, 0, st>>>(s->peer_sync, s->d_seqs, seq0, nseq, s->epoch, g.rank, g.world, 0xffffffffu);
        s->band_epoch = s->epoch;
    }
    k_swap_ref<<<1, nseq, 0, st>>>(s->d_seqs, seq0);
    for (int b = seq0; b < seq0 + nseq; b++) {
        for (int c = 0; c < 3; c++) std::swap(s->h[b].ref[c], s->h[b].rec[c]);
        for (int r = 0; r < FH_MAX_WORLD; r++) for (int c = 0; c < 3; c++) std::swap(s->h[b].peer_ref[r][c], s->h[b].peer_rec[r][c]);
    }
    rc = launch_phase_r(s, st, true, seq0, nseq, g.world == 1); if (rc) return rc;
    for (int b = seq0; b < seq0 + nseq; b++) { s->has_ref[b] = 1; s->prev_p[b] = 0; s->last_i[b] = 1; }
    CK(sync_streams(s));
    for (int b = seq0; b < seq0 + nseq; b++)
        if (s->h_status[(size_t)b * ST_WORDS + ST_FLAGS] & FLAG_TIMEOUT) return fail(FH264_E_CUDA, "intra wavefront wait timed out");
    return FH264_OK;
}

extern "C" int fh264_picture_status(fh264_session *s, int seq)
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    CK(cudaSetDevice(s->device));
    CK(sync_streams(s));
    const uint32_t f = s->h_status[(size_t)seq * ST_WORDS + ST_FLAGS];
    if (f & FLAG_UB_INPUT) return fail(FH264_E_UB_INPUT, "reference picture has an 8x8 window sum of 0 or >= 16203: undefined in the reference (moestimation.cpp:153-158,477-480)");
    if (f & FLAG_TIMEOUT) return fail(FH264_E_STATE, "a wavefront / cross-GPU wait timed out (band mode: is every rank encoding the same picture?)");
    if (f & FLAG_CAPACITY) return fail(FH264_E_CAPACITY, "stage-2 candidate capacity exceeded: a partition has more than 1023 candidates up to j_stop (flat content)");
    return FH264_OK;
}

extern "C" int fh264_mode_counts(fh264_session *s, int seq, int32_t counts[5])
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    if (!counts) return fail(FH264_E_ARG, "null output");
    CK(cudaSetDevice(s->device));
    CK(sync_streams(s));
    // device counters: [0] skip, [1] 16x16, [2] 16x8, [3] 8x16, [4] 8x8 == brojTipova order
    for (int i = 0; i < 5; i++) counts[i] = (int32_t)s->h_status[(size_t)seq * ST_WORDS + ST_COUNTS + i];
    return FH264_OK;
}

// Slice data of the P picture(s) last coded by fh264_encode_p, entropy-coded on the device (cavlc.cuh). See the header.
static int ensure_cavlc(fh264_session *s)
{
    if (s->d_cvs) return FH264_OK;
    const size_t nmb = (size_t)s->g.nmb;
    s->cvh.assign(s->batch, CvSeq());
    for (int b = 0; b < s->batch; b++) {
        CvSeq &c = s->cvh[b];
        CK(dalloc(s, &c.info, nmb));
        CK(dalloc(s, &c.buf, (nmb + 1) * CV_MB_WORDS));
        CK(dalloc(s, &c.bits, nmb + 1));
        CK(dalloc(s, &c.off, nmb + 2));
        CK(dalloc(s, &c.stream, (size_t)CV_STREAM_BYTES / 4 + 16));
        CK(dalloc(s, &c.stat, (size_t)2));
    }
    CK(dalloc(s, &s->d_cvs, (size_t)s->batch));
    CK(cudaMemcpy(s->d_cvs, s->cvh.data(), sizeof(CvSeq) * s->batch, cudaMemcpyHostToDevice));
    CK(cudaHostAlloc((void **)&s->h_cvstat, sizeof(uint32_t) * 2 * s->batch, cudaHostAllocDefault));
    return FH264_OK;
}

// shared tail of fh264_cavlc_p / fh264_cavlc_i: offsets, packing, status and the copies home
static int cavlc_finish(fh264_session *s, int seq0, int nseq, int first_bit, uint8_t *out, size_t out_stride, uint32_t *nbits, fh264_cavlc_mb_info *mb_info)
{
    const int nmb = s->g.nmb;
    cudaStream_t st = s->stream;
    k_cavlc_scan<<<nseq, 1024, 0, st>>>(s->d_cvs, seq0, nmb, first_bit);
    k_cavlc_pack<<<dim3((nmb + 1 + 3) / 4, nseq), 128, 0, st>>>(s->d_cvs, seq0, nmb, first_bit);
    CKL();
    for (int b = seq0; b < seq0 + nseq; b++) CK(cudaMemcpyAsync(s->h_cvstat + 2 * b, s->cvh[b].stat, sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int b = seq0; b < seq0 + nseq; b++) {            // every sequence is validated before any copy into the caller's buffers is queued
        const uint32_t fl = s->h_cvstat[2 * b], total = s->h_cvstat[2 * b + 1];
        if (fl & CV_FLAG_LEVEL_RANGE) return fail(FH264_E_UNSUPPORTED, "a level is outside the reference's level table (level_prefix > 15, residual_tables.cpp:940-1008)");
        if (fl & (CV_FLAG_MB_OVERFLOW | CV_FLAG_STREAM_OVERFLOW)) return fail(FH264_E_CAPACITY, "slice data exceeds the 500000-byte RBSP buffer of the reference (fer_h264.cpp:93)");
        if (((size_t)total + 7) / 8 > out_stride) return fail(FH264_E_ARG, "out_stride smaller than the slice data");
    }
    for (int b = seq0; b < seq0 + nseq; b++) {
        const uint32_t total = s->h_cvstat[2 * b + 1];
        nbits[b - seq0] = total;
        CK(cudaMemcpyAsync(out + (size_t)(b - seq0) * out_stride, s->cvh[b].stream, ((size_t)total + 7) / 8, cudaMemcpyDeviceToHost, st));
        if (mb_info) CK(cudaMemcpyAsync(mb_info + (size_t)(b - seq0) * nmb, s->cvh[b].info, sizeof(CvInfo) * (size_t)nmb, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return FH264_OK;
}

extern "C" int fh264_cavlc_p(fh264_session *s, int seq0, int nseq, int first_bit, uint8_t *out, size_t out_stride, uint32_t *nbits,
                             fh264_cavlc_mb_info *mb_info)
{
    int rc = check_seq(s, seq0, nseq); if (rc) return rc;
    if (!out || !nbits || first_bit < 0 || first_bit > 7) return fail(FH264_E_ARG, "bad argument");
    if (s->g.world > 1 && (s->g.rank != 0 || !s->g.gather_on || !s->h[seq0].gather[0]))
        return fail(FH264_E_UNSUPPORTED, "band mode: the slice is entropy-coded on rank 0, and only with the record gather on (fh264_band_gather on every rank, after fh264_ipc_import)");
    if (s->epoch == 0) return fail(FH264_E_STATE, "cavlc_p before any encode_p");
    for (int b = seq0; b < seq0 + nseq; b++) if (s->gate_calls[b]) { CK(cudaSetDevice(s->device)); CK(sync_streams(s)); break; }
    for (int b = seq0; b < seq0 + nseq; b++)
        if (!s->prev_p[b] || s->last_i[b]) return fail(FH264_E_STATE, "cavlc_p: the last picture of the sequence was not coded by encode_p (its records are not P records)");
    CK(cudaSetDevice(s->device));
    rc = enter_main(s); if (rc) return rc;
    rc = ensure_cavlc(s); if (rc) return rc;
    const int nmb = s->g.nmb, wmb = s->g.Wmb;
    cudaStream_t st = s->stream;
    for (int b = seq0; b < seq0 + nseq; b++) CK(cudaMemsetAsync(s->cvh[b].stream, 0, (size_t)CV_STREAM_BYTES + 64, st));
    const int par = s->g.world > 1 ? (int)(s->epoch & 1u) : -1;          // band mode (rank 0): the gathered records of this picture
    if (s->g.world > 1) k_band_wait_all<<<1, 1, 0, st>>>(s->peer_sync, s->d_seqs, seq0, nseq, s->band_epoch, s->g.rank, s->g.world);
    k_cavlc_prep<<<dim3((nmb + 127) / 128, nseq), 128, 0, st>>>(s->d_seqs, s->d_cvs, seq0, nmb, par);
    k_cavlc_code<<<dim3((nmb + 1 + 127) / 128, nseq), 128, 0, st>>>(s->d_seqs, s->d_cvs, seq0, nmb, wmb, par);
    return cavlc_finish(s, seq0, nseq, first_bit, out, out_stride, nbits, mb_info);
}

// slice_data() of the I picture(s) last coded by fh264_encode_i (intra.cuh, k_cavlc_code_i). See the header.
extern "C" int fh264_cavlc_i(fh264_session *s, int seq0, int nseq, int first_bit, uint8_t *out, size_t out_stride, uint32_t *nbits)
{
    int rc = check_seq(s, seq0, nseq); if (rc) return rc;
    if (!out || !nbits || first_bit < 0 || first_bit > 7) return fail(FH264_E_ARG, "bad argument");
    for (int b = seq0; b < seq0 + nseq; b++) if (s->gate_calls[b]) { CK(cudaSetDevice(s->device)); CK(sync_streams(s)); break; }
    for (int b = seq0; b < seq0 + nseq; b++) if (!s->last_i[b]) return fail(FH264_E_STATE, "cavlc_i: the last picture of the sequence was not coded by encode_i");
    CK(cudaSetDevice(s->device));
    rc = enter_main(s); if (rc) return rc;
    rc = ensure_cavlc(s); if (rc) return rc;
    const int nmb = s->g.nmb, wmb = s->g.Wmb;
    cudaStream_t st = s->stream;
    for (int b = seq0; b < seq0 + nseq; b++) {
        CK(cudaMemsetAsync(s->cvh[b].stream, 0, (size_t)CV_STREAM_BYTES + 64, st));
        CK(cudaMemsetAsync(s->cvh[b].stat, 0, sizeof(uint32_t) * 2, st));
    }
    k_cavlc_code_i<<<dim3((nmb + 1 + 127) / 128, nseq), 128, 0, st>>>(s->d_seqs, s->d_cvs, seq0, nmb, wmb);
    return cavlc_finish(s, seq0, nseq, first_bit, out, out_stride, nbits, nullptr);
}

// Snapshot of the 16 status words of sequence seq taken after phase C of its last encode_p (tests / diagnostics):
// [0] flags, [1] stage-2 pool entries used, [2..6] mode counts, [12] partitions redone by the large stage-2 launch.
extern "C" int fh264_debug_status(fh264_session *s, int seq, uint32_t out[16])
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    if (!out) return fail(FH264_E_ARG, "null output");
    CK(cudaSetDevice(s->device));
    CK(sync_streams(s));
    memcpy(out, s->h_status + (size_t)seq * ST_WORDS, sizeof(uint32_t) * 16);
    return FH264_OK;
}

extern "C" int fh264_download_recon(fh264_session *s, int seq, uint8_t *y, uint8_t *cb, uint8_t *cr)
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    if (!y || !cb || !cr) return fail(FH264_E_ARG, "null plane");
    if (!s->has_ref[seq]) return fail(FH264_E_STATE, "no reference picture yet");
    CK(cudaSetDevice(s->device));
    rc = enter_main(s); if (rc) return rc;
    const size_t WH = (size_t)s->g.WH;
    CK(cudaMemcpyAsync(y, s->h[seq].ref[0], WH, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(cb, s->h[seq].ref[1], WH / 4, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(cr, s->h[seq].ref[2], WH / 4, cudaMemcpyDeviceToHost, s->stream));
    CK(sync_streams(s));
    return FH264_OK;
}

extern "C" int fh264_last_timings(fh264_session *s, float ms[10])
{
    if (!s || !ms) return fail(FH264_E_ARG, "null argument");
    if (!s->timed) return fail(FH264_E_STATE, "no encode_p issued yet");
    CK(cudaSetDevice(s->device));
    CK(cudaEventSynchronize(s->ev[4]));
    for (int i = 0; i < 4; i++) CK(cudaEventElapsedTime(&ms[i], s->ev[i], s->ev[i + 1]));
    CK(cudaEventElapsedTime(&ms[4], s->ev[0], s->ev[4]));
    CK(cudaEventElapsedTime(&ms[5], s->ev[0], s->evk[0]));      // k_stage3 (+ k_begin_picture)
    CK(cudaEventElapsedTime(&ms[6], s->evk[0], s->ev_spec));    // k_stage2
    CK(cudaEventElapsedTime(&ms[7], s->evk[1], s->evk[2]));     // k_interp
    CK(cudaEventElapsedTime(&ms[8], s->evk[2], s->evk[3]));     // k_features
    CK(cudaEventElapsedTime(&ms[9], s->evk[3], s->ev[4]));      // k_tile_index
    return FH264_OK;
}

// k_spec (phase S) of the last encode_p, in ms (0 when FH264_SPEC=0)
extern "C" int fh264_last_spec_ms(fh264_session *s, float *ms)
{
    if (!s || !ms) return fail(FH264_E_ARG, "null argument");
    if (!s->timed) return fail(FH264_E_STATE, "no encode_p issued yet");
    CK(cudaSetDevice(s->device));
    CK(cudaEventSynchronize(s->ev[4]));
    CK(cudaEventElapsedTime(ms, s->ev_spec, s->ev[1]));
    return FH264_OK;
}

// Sustained integer-pipe rate of this GPU, T lane-statements per second: [0] IMAD, [1] VIADDMNMX.S16x2, [2] VIADDMNMX,
// [3] VABSDIFF4.U8.ACC + VIADD pairs (two instructions per statement). About 40 ms of GPU time.
extern "C" int fh264_measure_int_peak(int device, double tops[4])
{
    if (!tops) return fail(FH264_E_ARG, "null output");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(FH264_E_NO_DEVICE, "no such CUDA device");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    CK(cudaSetDevice(device));
    unsigned *d = nullptr;
    CK(cudaMalloc(&d, 64));
    const int sms = prop.multiProcessorCount, iters = 512;
    tops[0] = int_peak_run<0>(d, sms, iters, 0);
    tops[1] = int_peak_run<1>(d, sms, iters, 0);
    tops[2] = int_peak_run<2>(d, sms, iters, 0);
    tops[3] = int_peak_run<3>(d, sms, iters, 0);
    CK(cudaDeviceSynchronize());
    cudaFree(d);
    return FH264_OK;
}

static int ensure_scratch(fh264_session *s, size_t nmbs)
{
    if (nmbs <= s->scr_mbs) return FH264_OK;
    CK(sync_streams(s));
    for (int i = 0; i < 3; i++) if (s->d_scr[i]) { cudaFree(s->d_scr[i]); s->d_scr[i] = nullptr; }
    for (int i = 0; i < 2; i++) if (s->d_scr16[i]) { cudaFree(s->d_scr16[i]); s->d_scr16[i] = nullptr; }
    for (int i = 0; i < 3; i++) CK(cudaMalloc((void **)&s->d_scr[i], nmbs * 384));
    for (int i = 0; i < 2; i++) CK(cudaMalloc((void **)&s->d_scr16[i], nmbs * 384 * 2));
    s->scr_mbs = nmbs;
    return FH264_OK;
}

extern "C" int fh264_tq_macroblocks(fh264_session *s, int n, const uint8_t *src384, const uint8_t *pred384, int qp, int16_t *levels384, uint8_t *recon384)
{
    if (!s || !src384 || !pred384 || !levels384 || !recon384 || n < 1) return fail(FH264_E_ARG, "bad argument");
    if (qp < 0 || qp > 51) return fail(FH264_E_ARG, "qp outside 0..51");
    CK(cudaSetDevice(s->device));
    int rc = ensure_scratch(s, (size_t)n); if (rc) return rc;
    cudaStream_t st = s->stream;
    CK(cudaMemcpyAsync(s->d_scr[0], src384, (size_t)n * 384, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->d_scr[1], pred384, (size_t)n * 384, cudaMemcpyHostToDevice, st));
    k_tq_only<<<(n + 3) / 4, 128, 0, st>>>(s->d_scr[0], s->d_scr[1], n, qp, s->d_scr16[0], s->d_scr[2]);
    CKL();
    CK(cudaMemcpyAsync(levels384, s->d_scr16[0], (size_t)n * 384 * 2, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(recon384, s->d_scr[2], (size_t)n * 384, cudaMemcpyDeviceToHost, st));
    CK(sync_streams(s));
    return FH264_OK;
}

extern "C" int fh264_tq_luma_intra16(fh264_session *s, int n, const uint8_t *src256, const uint8_t *pred256, int qp, int16_t *dc16, int16_t *ac16x15, uint8_t *recon256)
{
    if (!s || !src256 || !pred256 || !dc16 || !ac16x15 || !recon256 || n < 1) return fail(FH264_E_ARG, "bad argument");
    if (qp < 0 || qp > 51) return fail(FH264_E_ARG, "qp outside 0..51");
    CK(cudaSetDevice(s->device));
    int rc = ensure_scratch(s, (size_t)n); if (rc) return rc;
    cudaStream_t st = s->stream;
    CK(cudaMemcpyAsync(s->d_scr[0], src256, (size_t)n * 256, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->d_scr[1], pred256, (size_t)n * 256, cudaMemcpyHostToDevice, st));
    k_tq_intra16<<<(n + 3) / 4, 128, 0, st>>>(s->d_scr[0], s->d_scr[1], n, qp, s->d_scr16[0], s->d_scr16[1], s->d_scr[2]);
    CKL();
    CK(cudaMemcpyAsync(dc16, s->d_scr16[0], (size_t)n * 16 * 2, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ac16x15, s->d_scr16[1], (size_t)n * 240 * 2, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(recon256, s->d_scr[2], (size_t)n * 256, cudaMemcpyDeviceToHost, st));
    CK(sync_streams(s));
    return FH264_OK;
}

extern "C" int fh264_motion_compensate(fh264_session *s, int seq, const int16_t *qmv, uint8_t *pred384)
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    if (!qmv || !pred384) return fail(FH264_E_ARG, "null argument");
    if (!s->has_ref[seq]) return fail(FH264_E_STATE, "no reference picture yet");
    CK(cudaSetDevice(s->device));
    rc = enter_main(s); if (rc) return rc;
    const int n = s->g.nmb;
    rc = ensure_scratch(s, (size_t)n); if (rc) return rc;
    cudaStream_t st = s->stream;
    CK(cudaMemcpyAsync(s->d_scr16[0], qmv, (size_t)n * 8 * 2, cudaMemcpyHostToDevice, st));
    k_mc_only<<<(n + 3) / 4, 128, 0, st>>>(s->d_seqs, seq, s->g, s->d_scr16[0], s->d_scr[2]);
    CKL();
    CK(cudaMemcpyAsync(pred384, s->d_scr[2], (size_t)n * 384, cudaMemcpyDeviceToHost, st));
    CK(sync_streams(s));
    return FH264_OK;
}

extern "C" int fh264_debug_plane(fh264_session *s, int seq, int f, uint8_t *out)
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    if (!out || f < 0 || f > 15) return fail(FH264_E_ARG, "bad argument");
    CK(cudaSetDevice(s->device));
    rc = enter_main(s); if (rc) return rc;
    CK(cudaMemcpyAsync(out, s->h[seq].planes + (size_t)f * s->g.WH, (size_t)s->g.WH, cudaMemcpyDeviceToHost, s->stream));
    CK(sync_streams(s));
    return FH264_OK;
}

__global__ void k_unpack_feature(const uint4 *__restrict__ kar, int n, int k, uint16_t *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 v = kar[i];
    out[i] = (uint16_t)feat_raw(v, k);
}

extern "C" int fh264_debug_feature(fh264_session *s, int seq, int k, int f, uint16_t *out)
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    if (!out || f < 0 || f > 15 || k < 0 || k > 4) return fail(FH264_E_ARG, "bad argument");
    CK(cudaSetDevice(s->device));
    rc = enter_main(s); if (rc) return rc;
    const int n = s->g.WH;
    rc = ensure_scratch(s, (size_t)(n * 2 + 767) / 768); if (rc) return rc;
    // the path keeps plane 0 only; the tap evaluates the same kernel for plane f into a temporary
    uint4 *tmp = nullptr;
    CK(cudaMalloc(&tmp, (size_t)n * sizeof(uint4)));
    const Geo &g = s->g;
    dim3 gf((g.W + FT_W - 1) / FT_W, (g.H + FT_H - 1) / FT_H, 1);
    k_features<<<gf, 256, 0, s->stream>>>(s->d_seqs, seq, g, f, tmp, 0);
    k_unpack_feature<<<(n + 255) / 256, 256, 0, s->stream>>>(tmp, n, k, (uint16_t *)s->d_scr16[0]);
    CKL();
    CK(cudaMemcpyAsync(out, s->d_scr16[0], (size_t)n * 2, cudaMemcpyDeviceToHost, s->stream));
    CK(sync_streams(s));
    cudaFree(tmp);
    return FH264_OK;
}

// Debug: per-macroblock clock64() samples of the phase-B wavefront of sequence seq (12 int64 per MB):
// [0] CTA start, [1] prefetch issued, [2] dependencies satisfied, [3] neighbour MVs loaded, [4] P_Skip decided,
// [5..8] partitions 0..3 decided, [9] published. Enable with out == NULL (allocates), read back with out != NULL.
extern "C" int fh264_debug_timeline(fh264_session *s, int seq, long long *out)
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    CK(cudaSetDevice(s->device));
    CK(sync_streams(s));
    const size_t n = (size_t)s->g.nmb * 24;
    if (!s->h[seq].dbg) {
        long long *p = nullptr;
        CK(cudaMalloc((void **)&p, n * sizeof(long long)));
        CK(cudaMemset(p, 0, n * sizeof(long long)));
        s->allocs.push_back(p);
        s->h[seq].dbg = p;
        CK(cudaMemcpy(&s->d_seqs[seq], &s->h[seq], sizeof(SeqDev), cudaMemcpyHostToDevice));
    }
    if (out) CK(cudaMemcpy(out, s->h[seq].dbg, n * sizeof(long long), cudaMemcpyDeviceToHost));
    return FH264_OK;
}

// ---- band mode (BASELINE config 4): one picture split into macroblock-row bands over the ranks of a node.
// Every rank keeps the whole reference picture; phases A/B/C run on the band only. The wavefront crosses GPUs through
// mirrored progress flags (phase B writes the band's last MB row into the next rank's memory), the reconstruction exchange
// is fused into phase C (peer stores), k_band_barrier separates pictures. Buffers are shared with CUDA IPC.
extern "C" int fh264_band_config(fh264_session *s, int rank, int world, int mb_row0, int mb_row1)
{
    if (!s) return fail(FH264_E_ARG, "null session");
    if (world < 1 || world > FH_MAX_WORLD || rank < 0 || rank >= world) return fail(FH264_E_ARG, "rank/world out of range (at most 8 ranks)");
    if (mb_row0 < 0 || mb_row1 <= mb_row0 || mb_row1 > s->g.Hmb) return fail(FH264_E_ARG, "empty or out-of-range macroblock-row band");
    CK(cudaSetDevice(s->device));
    CK(sync_streams(s));
    Geo &g = s->g;
    g.rank = rank; g.world = world; g.band_mb0 = mb_row0 * g.Wmb; g.band_nmb = (mb_row1 - mb_row0) * g.Wmb;
    g.halo_y0 = 0; g.halo_y1 = g.H; g.wait_mask = 0xffffffffu;       // until fh264_band_peers says where the other bands are
    s->sad_y0 = 0; s->sad_y1 = g.H;
    // wavefront order restricted to the band
    std::vector<int> order(g.band_nmb);
    for (int i = 0; i < g.band_nmb; i++) order[i] = g.band_mb0 + i;
    const int Wmb = g.Wmb;
    int sn, sd;
    wf_slope(sn, sd);
    std::stable_sort(order.begin(), order.end(), [Wmb, sn, sd](int a, int b) { return sd * (a % Wmb) + sn * (a / Wmb) < sd * (b % Wmb) + sn * (b / Wmb); });
    CK(cudaMemcpy(s->d_wf_order, order.data(), sizeof(int) * g.band_nmb, cudaMemcpyHostToDevice));
    s->peer_sync.p[rank] = s->d_sync;
    for (int b = 0; b < s->batch; b++)
        for (int k = 0; k < 2; k++)
            if (!s->h[b].gather[k]) CK(dalloc(s, &s->h[b].gather[k], (size_t)g.nmb));      // (used on rank 0; every rank exports one so that the handle blob is uniform)
    for (int b = 0; b < s->batch; b++)
        { for (int c = 0; c < 3; c++) { s->h[b].peer_ref[rank][c] = s->h[b].ref[c]; s->h[b].peer_rec[rank][c] = s->h[b].rec[c]; } s->h[b].peer_motion[rank] = s->h[b].motion; }
    CK(cudaMemcpy(s->d_seqs, s->h.data(), sizeof(SeqDev) * s->batch, cudaMemcpyHostToDevice));
    return FH264_OK;
}

// The bands of ALL ranks (mb_rows[2r], mb_rows[2r+1] = first / end macroblock row of rank r). With them the picture barrier stops
// being global: this rank's phases A / S / B / C read the reference picture only within FH_BAND_HALO luma rows of its band (stage 2:
// |dx| + |dy| < 280, moestimation.cpp:481; + the 8x8 block, the 8-row box sums below a position and the 6-tap's 3 rows), so phase R
// covers only those rows (rounded to the 64-row index tiles) and the barrier waits only for the ranks whose bands they touch — rank r
// starts picture t+1 while the wavefront of picture t is still running through the bands further down. From here on the scene SAD of
// a rank covers its own band (the bands' sums add up to the picture's), and fh264_download_recon returns the whole picture only
// after every rank has finished it (synchronise the ranks on the host first).
#define FH_BAND_HALO 304
extern "C" int fh264_band_peers(fh264_session *s, int world, const int *mb_rows)
{
    if (!s || !mb_rows) return fail(FH264_E_ARG, "null argument");
    Geo &g = s->g;
    if (world != g.world || g.world < 2) return fail(FH264_E_STATE, "fh264_band_config first (same world size)");
    if (mb_rows[2 * g.rank] * g.Wmb != g.band_mb0 || (mb_rows[2 * g.rank + 1] - mb_rows[2 * g.rank]) * g.Wmb != g.band_nmb)
        return fail(FH264_E_ARG, "this rank's entry differs from its fh264_band_config");
    CK(cudaSetDevice(s->device));
    CK(sync_streams(s));
    // reads[a] = ranks whose bands rank a reads (its halo rows, 16 more for the planes below them, 3 for the 6-tap). Rank a must not
    // start picture t+1 before they have delivered picture t — and must not OVERWRITE (its phase C of picture t+1 stores into the
    // buffer that held picture t-1) what a rank that reads a's band may still be reading: the relation is made symmetric.
    uint32_t reads[FH_MAX_WORLD] = { 0 };
    int hy0[FH_MAX_WORLD], hy1[FH_MAX_WORLD];
    for (int a = 0; a < world; a++) {
        const int top = mb_rows[2 * a] * 16, bottom = mb_rows[2 * a + 1] * 16;
        hy0[a] = std::max(0, top - FH_BAND_HALO) / FH_TILE * FH_TILE;
        hy1[a] = std::min(g.H, (bottom + FH_BAND_HALO + FH_TILE - 1) / FH_TILE * FH_TILE);
        const int need0 = hy0[a] - 3, need1 = std::min(g.H, hy1[a] + 16) + 3;
        for (int r = 0; r < world; r++)
            if (mb_rows[2 * r] * 16 < need1 && mb_rows[2 * r + 1] * 16 > need0) reads[a] |= 1u << r;
        reads[a] |= 1u << a;
        if (a + 1 < world) reads[a] |= 1u << (a + 1);        // the rank below reads this band's last row of vectors (qmv mirror): never more than a picture apart
        if (a > 0) reads[a] |= 1u << (a - 1);
    }
    g.halo_y0 = hy0[g.rank]; g.halo_y1 = hy1[g.rank];
    g.wait_mask = reads[g.rank];
    for (int a = 0; a < world; a++) if ((reads[a] >> g.rank) & 1u) g.wait_mask |= 1u << a;
    s->sad_y0 = mb_rows[2 * g.rank] * 16; s->sad_y1 = mb_rows[2 * g.rank + 1] * 16;
    return FH264_OK;
}

// Band mode: from the next picture on, phase C of this rank also stores every macroblock's record into rank 0's gather buffer (6.7 MB per
// 1080p picture over NVLink in total), so that rank 0 can entropy-code the slice (fh264_cavlc_p). Every rank must make the same call.
extern "C" int fh264_band_gather(fh264_session *s, int on)
{
    if (!s) return fail(FH264_E_ARG, "null session");
    if (s->g.world < 2) return fail(FH264_E_STATE, "fh264_band_config first");
    CK(cudaSetDevice(s->device));
    CK(sync_streams(s));
    s->g.gather_on = on ? 1 : 0;
    return FH264_OK;
}

// handles: FH264_IPC_HANDLES x 64 bytes = ref[3], rec[3], motion, qmv, sync, gather[2]
extern "C" int fh264_ipc_export(fh264_session *s, int seq, uint8_t *handles)
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    if (!handles) return fail(FH264_E_ARG, "null output");
    CK(cudaSetDevice(s->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == FH264_IPC_HANDLE_BYTES, "IPC handle size");
    const SeqDev &S = s->h[seq];
    if (!S.gather[0]) return fail(FH264_E_STATE, "fh264_band_config first");
    void *ptrs[FH264_IPC_HANDLES] = { S.ref[0], S.ref[1], S.ref[2], S.rec[0], S.rec[1], S.rec[2], S.motion, S.qmv, s->d_sync, S.gather[0], S.gather[1] };
    for (int i = 0; i < FH264_IPC_HANDLES; i++) {
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, ptrs[i]));
        memcpy(handles + (size_t)i * FH264_IPC_HANDLE_BYTES, &h, FH264_IPC_HANDLE_BYTES);
    }
    return FH264_OK;
}

extern "C" int fh264_ipc_import(fh264_session *s, int seq, int peer_rank, const uint8_t *handles)
{
    int rc = check_seq(s, seq, 1); if (rc) return rc;
    if (!handles || peer_rank < 0 || peer_rank >= s->g.world || peer_rank == s->g.rank) return fail(FH264_E_ARG, "bad peer rank / handles");
    CK(cudaSetDevice(s->device));
    CK(sync_streams(s));
    void *ptrs[FH264_IPC_HANDLES];
    for (int i = 0; i < FH264_IPC_HANDLES; i++) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)i * FH264_IPC_HANDLE_BYTES, FH264_IPC_HANDLE_BYTES);
        ptrs[i] = nullptr;
        if (i == 8 && s->peer_sync.p[peer_rank]) { ptrs[i] = s->peer_sync.p[peer_rank]; continue; }   // sync area: once per peer
        if (i >= 9 && peer_rank != 0) continue;                                                     // only rank 0 gathers the records
        CK(cudaIpcOpenMemHandle(&ptrs[i], h, cudaIpcMemLazyEnablePeerAccess));
        s->ipc_opened.push_back(ptrs[i]);
    }
    SeqDev &S = s->h[seq];
    for (int c = 0; c < 3; c++) { S.peer_ref[peer_rank][c] = (uint8_t *)ptrs[c]; S.peer_rec[peer_rank][c] = (uint8_t *)ptrs[3 + c]; }
    S.peer_motion[peer_rank] = (MbMotion *)ptrs[6];
    if (peer_rank == 0) { S.gather[0] = (fh264_mb_result *)ptrs[9]; S.gather[1] = (fh264_mb_result *)ptrs[10]; }
    if (peer_rank == s->g.rank + 1) S.peer_qmv_next = (unsigned long long *)ptrs[7];
    s->peer_sync.p[peer_rank] = (uint32_t *)ptrs[8];
    CK(cudaMemcpy(&s->d_seqs[seq], &S, sizeof(SeqDev), cudaMemcpyHostToDevice));
    return FH264_OK;
}
