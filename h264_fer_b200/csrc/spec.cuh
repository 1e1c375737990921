// Phase S — speculative completion of the 8x8 search, fully parallel over partitions (one warp each).
//
// Everything in interEncoding's per-partition search (moestimation.cpp:430-528) that depends on the MV predictor depends on
// it in two steps: the three candidate lists depend only on gen = mvp >> 2 (window centre of stage 1 and the multiplier
// |dx-genx|+|dy-geny|+4 of stages 1 and 2, :433-437,458,492-494; stage 3 uses centre 0, :509-510), and the final choice
// argmin(SAD + |mv-mvp|_1) (:460-469,498-507,511-520) on mvp itself. The predictor is only known inside the wavefront
// (phase B), but gen is highly predictable: it equals the integer part of the partition's own best stage-3 vector, or the
// gen the same partition had one picture earlier, for ~99 % of the partitions of moving content. So for up to two GUESSED
// values of gen this kernel runs stage 1 (feature window around the guess, 17 best, SADs), ranks the stage-2 set with the
// guess's multiplier (33 best) and takes the stage-3 list, and reduces the <= 83 evaluated candidates to the few FINALISTS
// that can win for SOME mvp of the guessed cell [4*gen, 4*gen+3]^2:
//   hi(c) = SAD(c) + max over the cell of |mv(c)-mvp|_1,  lo(c) likewise with min;   U = min_c hi(c)
//   c can be the strict first minimum for some mvp of the cell only if lo(c) <= U; of candidates with equal MVs only the
//   first in (stage, list position) order can. (Exact: the candidate attaining U beats every c with lo(c) > U for every
//   mvp of the cell, strictly.)
// Phase B then only compares gen with the guesses and evaluates the finalists (<= SPEC_NF, 1.4 on average) with the true
// mvp; a wrong guess, an overflow or an oversized stage-2 set falls back to the full search there. Results never depend on
// the guesses — only the time does.
#pragma once
#include "common.cuh"
#include "warp_select.cuh"
#include "qfeat.cuh"
#include "phase_a.cuh"
#include "qwin.cuh"
#include "topk.cuh"
#include "stage3.cuh"

#ifndef FH_SPEC_MINB
#define FH_SPEC_MINB 5
#endif
#define SPEC_NF 7
#define SPEC_INVALID 0xffu
#define SPEC_NOGUESS 0x7fff
#define SPEC_PF_CAP 96           // stage 1 (17) + stage 2 (33) + stage 3 (33) evaluated candidates at most
#define SPEC_PREV_NONE 0x7f7f7f7fu     // cudaMemset(0x7f): no previous P picture

struct __align__(8) SpecFinal { int16_t mvx, mvy; uint16_t sad, order; };        // order = stage << 14 | list position
struct __align__(16) PartSpec {
    int16_t gx[2], gy[2];         // guessed gen per slot (SPEC_NOGUESS: none)
    uint8_t nf[2];                // finalists per slot (SPEC_INVALID: slot unusable)
    uint8_t pad[6];
    SpecFinal f[2][SPEC_NF];
};
static_assert(sizeof(PartSpec) == 128, "PartSpec size");

// |v - t| over t in [4g, 4g+3] with a = v - 4g: smallest and largest value
__device__ __forceinline__ int cell_min_(int a) { return a < 0 ? -a : (a > 3 ? a - 3 : 0); }
__device__ __forceinline__ int cell_max_(int a) { return max(iabs_(a), iabs_(a - 3)); }

struct __align__(128) SpecWarp {
    uint64_t bar;                            // mbarrier of the TMA-staged pixel window (qwin.cuh)
    TopKBufT<128> tk;                        // selection of the 17 stage-1 candidates (topk.cuh; appended after the bound is tight)
    uint32_t stage[QW_STAGE_WORDS];          // the window's costs until the bound is known (qwin_select)
    uint16_t members[FH_S1_MAX + 3];
    uint16_t msad[FH_S1_MAX + 3];
    S3Entry s3[FH_S3_MAX + 1];
    uint32_t pf_mv[SPEC_PF_CAP];             // evaluated candidates that can still win: mvx & 0xffff | mvy << 16
    uint32_t pf_so[SPEC_PF_CAP];             // sad | order << 16
};

__device__ __forceinline__ void pf_emit(bool have, int mvx, int mvy, int sad, int order, SpecWarp *sw, int &npf)
{
    const int lane = threadIdx.x & 31;
    const unsigned b = __ballot_sync(0xffffffffu, have);
    if (have) {
        const int p = npf + __popc(b & ((1u << lane) - 1u));
        if (p < SPEC_PF_CAP) { sw->pf_mv[p] = ((uint32_t)mvx & 0xffffu) | ((uint32_t)mvy << 16); sw->pf_so[p] = (uint32_t)sad | ((uint32_t)order << 16); }
    }
    npf += __popc(b);
}

// The complete search of one partition for gen = (Gx, Gy) (stage 3 list of phase A, stage 2 ranked with the multiplier of this
// gen and measured, stage 1 window around it): every evaluated candidate that can be the strict first minimum of
// SAD + |mv - mvp|_1 for SOME mvp of the cell [4*gen, 4*gen+3]^2 is left in sw->pf_mv / sw->pf_so (vector; SAD | order << 16).
// Returns their number; usable = false when the stage-2 set is not in the pool (S2_SLOW). Warp-uniform call.
__device__ __forceinline__ int spec_collect(const SeqDev &S, const Geo &g, const fh264_params &prm, int xP, int yP, const uint2 (&rows)[8],
                                            const FeatQ &fq, int n3, uint32_t n2w, uint32_t s2_off, int Gx, int Gy,
                                            SpecWarp *sw, uint8_t *win, const CUtensorMap *tmap, uint32_t &phase, const WinMagic &wm, bool &usable)
{
    const int lane = threadIdx.x & 31, W = g.W, H = g.H;
    const int g1 = prm.window / 16, w1 = 2 * g1 + 1;
    const uint32_t i1 = wm.i1;
    const int cxq = 4 * Gx, cyq = 4 * Gy;
    int npf = 0;
    usable = !(n2w & S2_SLOW);                             // oversized stage-2 set: phase B enumerates it itself
    // the pixel window of stage 1 starts to arrive now (TMA) and is consumed after stages 3 and 2
    const QWinGeo qg = qwin_geo(g1);
    int woff;
    const bool tma = qwin_fill(S, g, tmap, qg, xP + Gx - g1, yP + Gy - g1, win, &sw->bar, woff);
    // ---- stage 3 (list of phase A, SADs known): bound U, then the entries that can still win
    int U = 0x7fffffff;
    for (int i = lane; i < n3; i += 32) {
        const S3Entry e = sw->s3[i];
        U = min(U, (int)e.sad + cell_max_(e.mvx - cxq) + cell_max_(e.mvy - cyq));
    }
    U = __reduce_min_sync(0xffffffffu, U);
    for (int i0 = 0; i0 < n3; i0 += 32) {
        const int i = i0 + lane;
        bool have = false; int mvx = 0, mvy = 0, sad = 0;
        if (i < n3) {
            const S3Entry e = sw->s3[i];
            mvx = e.mvx; mvy = e.mvy; sad = e.sad;
            have = sad + cell_min_(mvx - cxq) + cell_min_(mvy - cyq) <= U;
        }
        pf_emit(have, mvx, mvy, sad, (2 << 14) | i, sw, npf);
    }
    // ---- stage 2 (:470-507): candidates of phase A with feature distance, a lower bound of the SAD and the arrival key. List
    //      membership = one of the 33 smallest keys (cost, arrival). Only candidates whose lower bound leaves them a chance have
    //      their rank counted, and only those on the list their SAD measured (8 lanes, one row each).
    if (usable) {
        const int n2 = (int)n2w;
        const uint4 *__restrict__ pool = S.s2pool + s2_off;
        const uint2 cr8 = pick_row(rows, lane & 7);
#pragma unroll 1
        for (int i0 = 0; i0 < n2; i0 += 32) {
            const int i = i0 + lane;
            bool pot = false; int dx = 0, dy = 0; uint32_t cst = 0, ak = 0;
            if (i < n2) {
                const uint4 v = __ldg(&pool[i]);
                dx = (int16_t)(v.x & 0xffff); dy = (int16_t)(v.x >> 16); ak = v.w;
                cst = (uint32_t)(iabs_(dx - Gx) + iabs_(dy - Gy) + 4) * v.y;
                // (the second listing of a bucket-s0 entry has the same vector and a later arrival: only the first can win)
                pot = cst < (uint32_t)FH_COST_EMPTY && !((ak >> 21) == 0u && (ak & (1u << 20))) &&
                      (int)v.z + cell_min_(4 * dx - cxq) + cell_min_(4 * dy - cyq) <= U;
            }
            unsigned pm = __ballot_sync(0xffffffffu, pot);
            while (pm) {
                const int src = __ffs(pm) - 1;
                pm &= pm - 1;
                const uint32_t ci = __shfl_sync(0xffffffffu, cst, src), ai = __shfl_sync(0xffffffffu, ak, src);
                int c = 0;
#pragma unroll 2
                for (int j = lane; j < n2; j += 32) {
                    const uint4 w = __ldg(&pool[j]);
                    const int jx = (int16_t)(w.x & 0xffff), jy = (int16_t)(w.x >> 16);
                    const uint32_t cj = (uint32_t)(iabs_(jx - Gx) + iabs_(jy - Gy) + 4) * w.y;
                    c += (cj < ci) || (cj == ci && w.w < ai);
                }
                c = __reduce_add_sync(0xffffffffu, c);
                bool have = false; int sad = 0;
                const int sx = __shfl_sync(0xffffffffu, dx, src), sy = __shfl_sync(0xffffffffu, dy, src);
                if (c < FH_S3_MAX) {                                   // on the list: measure it
                    sad = lane < 8 ? sad8(cr8, load_row8(S.planes, W, H, xP + sx, yP + sy + lane)) : 0;
                    sad += __shfl_xor_sync(0xffffffffu, sad, 1);
                    sad += __shfl_xor_sync(0xffffffffu, sad, 2);
                    sad += __shfl_xor_sync(0xffffffffu, sad, 4);
                    sad = __shfl_sync(0xffffffffu, sad, 0);
                    have = lane == 0 && sad + cell_min_(4 * sx - cxq) + cell_min_(4 * sy - cyq) <= U;
                }
                pf_emit(have, sx * 4, sy * 4, sad, (1 << 14) | c, sw, npf);
            }
        }
    }
    // ---- stage 1 (:458-469): feature window around the guess, 17 best by (cost, arrival), their SADs
    {
        qwin_wait(tma, &sw->bar, phase);
        const QWinView qv = { win, QW_ROWB, qg.rows * QW_ROWB, woff };
        const int nvalid = max(0, min(W - 1, xP + Gx + g1) - max(0, xP + Gx - g1) + 1) * max(0, min(H - 1, yP + Gy + g1) - max(0, yP + Gy - g1) + 1) * 16;
        TopK tk;
        tk_init(tk, min(FH_S1_MAX, nvalid), 128);
        qwin_select_w(w1, g, qv, xP, yP, Gx, Gy, fq, sw->tk.key, tk, sw->stage, 0u);
        const int nm = tk_finish(sw->tk.key, tk, nvalid, sw->members, true);
        const int r = lane & 7;
        const uint2 cr = pick_row(rows, r);
        for (int base = 0; base < nm; base += 4 * 5) {
            uint2 rr[5];
#pragma unroll
            for (int u = 0; u < 5; u++) {
                const int m = base + u * 4 + (lane >> 3);
                rr[u] = make_uint2(0, 0);
                if (m < nm) {
                    // the candidate's block lies inside the staged window (same clamping as satdLuma8x8MVs for an origin inside the picture)
                    const int i = (int)sw->members[m], f = i & 15, pos = i >> 4, cx = udiv_by(pos, i1), cy = pos - cx * w1;
                    rr[u] = qwin_row8(qv, f, cx, cy + r);
                }
            }
#pragma unroll
            for (int u = 0; u < 5; u++) {
                const int m = base + u * 4 + (lane >> 3);
                int sad = m < nm ? sad8(cr, rr[u]) : 0;
                sad += __shfl_xor_sync(0xffffffffu, sad, 1);
                sad += __shfl_xor_sync(0xffffffffu, sad, 2);
                sad += __shfl_xor_sync(0xffffffffu, sad, 4);
                if (m < nm && r == 0) sw->msad[m] = (uint16_t)sad;
            }
        }
        __syncwarp();
        {
            const int m = lane;
            bool have = false; int mvx = 0, mvy = 0, sad = 0;
            if (m < nm) {
                const int i = (int)sw->members[m], f = i & 15, pos = i >> 4, cx = udiv_by(pos, i1), dx = Gx + cx - g1, dy = Gy + pos - cx * w1 - g1;
                mvx = (dx << 2) | (f & 3); mvy = (dy << 2) | (f >> 2); sad = sw->msad[m];
                have = sad + cell_min_(mvx - cxq) + cell_min_(mvy - cyq) <= U;
            }
            pf_emit(have, mvx, mvy, sad, m, sw, npf);
        }
    }
    __syncwarp();
    if (npf > SPEC_PF_CAP) usable = false;                 // cannot happen (17 + 33 + 33 candidates at most)
    return min(npf, SPEC_PF_CAP);
}

// The finalists of one partition for one guessed gen = (Gx, Gy) into out->f[slot]. Warp-uniform call.
__device__ __forceinline__ void spec_slot(const SeqDev &S, const Geo &g, const fh264_params &prm, int part, int xP, int yP, const uint2 (&rows)[8],
                                          const FeatQ &fq, int n3, uint32_t n2w, uint32_t s2_off, int Gx, int Gy, int slot,
                                          SpecWarp *sw, uint8_t *win, PartSpec *out, const CUtensorMap *tmap, uint32_t &phase, const WinMagic &wm)
{
    const int lane = threadIdx.x & 31;
    const int cxq = 4 * Gx, cyq = 4 * Gy;
    bool usable;
    const int n = spec_collect(S, g, prm, xP, yP, rows, fq, n3, n2w, s2_off, Gx, Gy, sw, win, tmap, phase, wm, usable);
    // ---- finalists: bound over everything still in the race, duplicates of one MV dropped (the first in list order stays)
    int Uf = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
        const uint32_t mv = sw->pf_mv[i], so = sw->pf_so[i];
        Uf = min(Uf, (int)(so & 0xffffu) + cell_max_((int)(int16_t)(mv & 0xffffu) - cxq) + cell_max_((int)(int16_t)(mv >> 16) - cyq));
    }
    Uf = __reduce_min_sync(0xffffffffu, Uf);
    int nf = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        bool keep = false; uint32_t mv = 0, so = 0;
        if (i < n) {
            mv = sw->pf_mv[i]; so = sw->pf_so[i];
            keep = (int)(so & 0xffffu) + cell_min_((int)(int16_t)(mv & 0xffffu) - cxq) + cell_min_((int)(int16_t)(mv >> 16) - cyq) <= Uf;
            for (int j = 0; j < n && keep; j++) keep = !(sw->pf_mv[j] == mv && (sw->pf_so[j] >> 16) < (so >> 16));
        }
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int p = nf + __popc(b & ((1u << lane) - 1u));
            if (p < SPEC_NF) {
                SpecFinal sf;
                sf.mvx = (int16_t)(mv & 0xffffu); sf.mvy = (int16_t)(mv >> 16); sf.sad = (uint16_t)(so & 0xffffu); sf.order = (uint16_t)(so >> 16);
                out->f[slot][p] = sf;
            }
        }
        nf += __popc(b);
    }
    if (nf > SPEC_NF || nf == 0) usable = false;
    if (lane == 0) {
        out->gx[slot] = (int16_t)Gx; out->gy[slot] = (int16_t)Gy;
        out->nf[slot] = usable ? (uint8_t)nf : (uint8_t)SPEC_INVALID;
    }
    __syncwarp();
}

// Proxy of a quadrant's final vector: the best stage-3 vector by SAD of that 8x8 partition (k_stage3); zero if it has none.
__device__ __forceinline__ void proxy_mv(const SeqDev &S, int mb, int q, int &x, int &y)
{
    const uint32_t v = __ldg(&S.proxy[(size_t)mb * 4 + q]);
    x = v == SPEC_PREV_NONE ? 0 : (int)(int16_t)(v & 0xffffu);
    y = v == SPEC_PREV_NONE ? 0 : (int)(int16_t)(v >> 16);
}
// The predictor of 8x8 partition pi of macroblock mb (mode_pred.cpp:113-161,252-332) evaluated on the proxies
// instead of the final vectors: the neighbours are partitions of the same picture, whose best stage-3 vector is the final
// vector for ~96 % of them, and the median absorbs a single wrong one.
__device__ __forceinline__ void proxy_predictor(const SeqDev &S, const Geo &g, int mb, int pi, int &ox, int &oy)
{
    int mbx, mby;
    mb_xy(g, mb, mbx, mby);
    const int aL = mbx > 0, aU = mby > 0, aUR = mby > 0 && mbx < g.Wmb - 1, aUL = mby > 0 && mbx > 0;
    int ax = 0, ay = 0, bx = 0, by = 0, cx = 0, cy = 0, aA = 1, aB = 1, aC = 1;
    if (pi == 0) {
        aA = aL; aB = aU; aC = aU ? 1 : aUL;
        if (aL) proxy_mv(S, mb - 1, 1, ax, ay);
        if (aU) { proxy_mv(S, mb - g.Wmb, 2, bx, by); proxy_mv(S, mb - g.Wmb, 3, cx, cy); }
        else if (aUL) proxy_mv(S, mb - g.Wmb - 1, 3, cx, cy);
    } else if (pi == 1) {
        aB = aU; aC = aUR ? 1 : aU;
        proxy_mv(S, mb, 0, ax, ay);
        if (aU) proxy_mv(S, mb - g.Wmb, 3, bx, by);
        if (aUR) proxy_mv(S, mb - g.Wmb + 1, 2, cx, cy);
        else if (aU) proxy_mv(S, mb - g.Wmb, 2, cx, cy);
    } else if (pi == 2) {
        aA = aL;
        if (aL) proxy_mv(S, mb - 1, 3, ax, ay);
        proxy_mv(S, mb, 0, bx, by); proxy_mv(S, mb, 1, cx, cy);
    } else {
        proxy_mv(S, mb, 2, ax, ay); proxy_mv(S, mb, 1, bx, by); proxy_mv(S, mb, 0, cx, cy);
    }
    median_pred(aA, ax, ay, aB, bx, by, aC, cx, cy, ox, oy);
}

// One warp per partition, four per CTA (one macroblock). Guesses of gen = mvp >> 2: the predictor rule applied to the
// neighbours' proxies (98.5 % right on the bench content), then the partition's own proxy (together 99.7 %); without stage-3
// lists (BasicInterEncoding) the gen phase B used for this partition in the previous P picture, else zero.
__global__ void __launch_bounds__(128, FH_SPEC_MINB) k_spec(const SeqDev *__restrict__ seqs, int seq0, Geo g, fh264_params prm, int use_prev, WinMagic wm,
                                                 const CUtensorMap *__restrict__ tmaps)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // dynamic shared memory: 4 windows (128-byte aligned, qwin_bytes each) | 4 SpecWarp
    const int wbytes = qwin_bytes(prm.window / 16);
    uint8_t *win = smem_raw + (size_t)warp * wbytes;
    SpecWarp *sw = (SpecWarp *)(smem_raw + 4 * (size_t)wbytes) + warp;
    const SeqDev &S = seqs[seq0 + blockIdx.y];
    if (S.status[ST_GATE]) return;
    const CUtensorMap *tmap = tmaps ? tmaps + seq0 + blockIdx.y : nullptr;
    uint32_t phase = 0;
    if (lane == 0) mbar_init(&sw->bar, 1);
    __syncwarp();
    const int part = g.band_mb0 * 4 + blockIdx.x * 4 + warp;
    int xP, yP;
    part_origin(g, part, xP, yP);
    // every global load of the prologue is issued before anything waits for one (the warp is latency bound here: they used to be
    // four dependent round trips — source rows, partition record, its list, the neighbours' proxies)
    uint2 rows[8];
    load_cur8x8(S.cur[0], g, xP, yP, rows);
    PartSpec *out = &S.spec[part];
    int n3 = 0; uint32_t n2w = 0, s2_off = 0;
    PartA pa;
    S3Entry e0, e1;
    uint32_t own = SPEC_PREV_NONE;
    int pdx = 0, pdy = 0;
    if (!prm.basic) {
        pa = S.parta[part];
        e0 = S.s3[(size_t)part * FH_S3_MAX + lane];                                   // the list is read whole (33 slots); n3 says how much of it counts
        if (lane == 0) e1 = S.s3[(size_t)part * FH_S3_MAX + 32];
        own = __ldg(&S.proxy[part]);
        proxy_predictor(S, g, part >> 2, part & 3, pdx, pdy);
        n3 = pa.n3; n2w = pa.n2; s2_off = pa.s2_off;
        sw->s3[lane] = e0;
        if (lane == 0) sw->s3[32] = e1;
    }
    int s[5];
    block_sums(rows, s);
    const FeatQ fq = feat_query(s);
    __syncwarp();
    // guesses
    int g0x = 0, g0y = 0, g1x = 0, g1y = 0, ng = 0;
    if (!prm.basic) {
        const int dx = pdx, dy = pdy;
        g0x = dx >> 2; g0y = dy >> 2; ng = 1;
        if (own != SPEC_PREV_NONE) {
            const int ox = (int)(int16_t)(own & 0xffffu) >> 2, oy = (int)(int16_t)(own >> 16) >> 2;
            if (ox != g0x || oy != g0y) { g1x = ox; g1y = oy; ng = 2; }
        }
    }
    if ((use_prev & 1) && ng == 0) {
        const uint32_t pg = S.prev_gen[part];
        if (pg != SPEC_PREV_NONE) {
            const int px = (int16_t)(pg & 0xffffu), py = (int16_t)(pg >> 16);
            if (ng == 0) { g0x = px; g0y = py; ng = 1; }
            else if (px != g0x || py != g0y) { g1x = px; g1y = py; ng = 2; }
        }
    }
    if (ng == 0) ng = 1;                                   // no list and no history: guess gen = (0, 0)
    if (use_prev & 2) ng = 1;                              // (development knob FH264_SPEC_NG=1: first guess only)
    for (int slot = 0; slot < ng; slot++)
        spec_slot(S, g, prm, part, xP, yP, rows, fq, n3, n2w, s2_off, slot ? g1x : g0x, slot ? g1y : g0y, slot, sw, win, out, tmap, phase, wm);
    if (ng == 1 && lane == 0) { out->gx[1] = SPEC_NOGUESS; out->gy[1] = SPEC_NOGUESS; out->nf[1] = SPEC_INVALID; }
}

// ---- P_Skip trial, precomputed (moestimation.cpp:402-425; mode_pred.cpp:383-401) -------------------------------------------
// The skip vector is either zero or the 16x16 median predictor, known only inside the wavefront; whether the macroblock skips
// is "all 256 luma samples within MAXDIFF of the prediction at that vector" (:228-244). Per macroblock this kernel evaluates the
// test for the zero vector and for all 16 quarter-pel vectors of up to two guessed integer cells (the cell of partition 0's
// first guess, and the cell of the predictor the macroblock had one picture earlier): a 16-bit mask per cell. Phase B looks the
// answer up and only runs the test itself when the predictor falls outside the guessed cells. MAXDIFF (fixed, or the mean
// absolute deviation of the macroblock, :407-419) is a function of the source alone and is computed here as well.
struct __align__(16) MbSpec {
    int16_t cx[2], cy[2];        // guessed integer cell of the skip vector per slot (SPEC_NOGUESS: none / not evaluable from the planes)
    uint16_t mask[2];            // bit fy*4+fx: the macroblock skips with vector (4*cx+fx, 4*cy+fy)
    int16_t maxdiff;
    uint8_t zero_ok, pad;
};
static_assert(sizeof(MbSpec) == 16, "MbSpec size");
#define SKIPWIN_BYTES (16 * 16 * QW_ROWB)

// One warp per macroblock, four per CTA. tmaps16: tensor maps of the planes with a box of 32 bytes x 16 rows x 16 planes.
__global__ void __launch_bounds__(128) k_skipspec(const SeqDev *__restrict__ seqs, int seq0, Geo g, fh264_params prm, const CUtensorMap *__restrict__ tmaps16)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];      // 4 windows (128-byte aligned for the TMA) | 4 mbarriers
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *win = smem_raw + (size_t)warp * SKIPWIN_BYTES;
    uint64_t *bar = (uint64_t *)(smem_raw + 4 * SKIPWIN_BYTES) + warp;
    const SeqDev &S = seqs[seq0 + blockIdx.y];
    if (S.status[ST_GATE]) return;
    const int mb = g.band_mb0 + blockIdx.x * 4 + warp;
    if (mb >= g.band_mb0 + g.band_nmb) return;
    const CUtensorMap *tmap = tmaps16 ? tmaps16 + seq0 + blockIdx.y : nullptr;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    uint32_t phase = 0;
    int mbx, mby;
    mb_xy(g, mb, mbx, mby);
    const int W = g.W, H = g.H;
    const int r = lane >> 1, hf = lane & 1;                       // this lane's 8 samples: row r, columns 8*hf ..
    const size_t own = (size_t)(mby * 16 + r) * W + mbx * 16 + 8 * hf;
    const uint2 c = __ldg((const uint2 *)(S.cur[0] + own));
    int maxdiff = prm.maxdiff_set;
    if (prm.maxdiff_set == -1) {                                   // :407-419
        const int sum = (int)__reduce_add_sync(0xffffffffu, __vsadu4(c.x, 0u) + __vsadu4(c.y, 0u));
        const uint32_t mean4 = (uint32_t)(sum / 256) * 0x01010101u;
        const int dev = (int)__reduce_add_sync(0xffffffffu, __vsadu4(c.x, mean4) + __vsadu4(c.y, mean4));
        maxdiff = max(3, dev / 256);
    }
    const uint32_t md4 = (uint32_t)min(max(maxdiff, 0), 255) * 0x01010101u;
    auto within = [&](uint2 p) -> bool {                            // every |cur - pred| <= MAXDIFF, warp-wide (:228-244)
        const uint32_t bad = __vcmpgtu4(__vabsdiffu4(c.x, p.x), md4) | __vcmpgtu4(__vabsdiffu4(c.y, p.y), md4);
        return !__any_sync(0xffffffffu, bad != 0u);
    };
    MbSpec ms;
    ms.maxdiff = (int16_t)maxdiff; ms.pad = 0;
    ms.zero_ok = maxdiff < 0 ? 0 : (uint8_t)within(__ldg((const uint2 *)(S.planes + own)));        // plane 0 == the reference picture
    // guessed cells
    int gx[2] = { SPEC_NOGUESS, SPEC_NOGUESS }, gy[2] = { SPEC_NOGUESS, SPEC_NOGUESS };
    {
        int n = 0;
        if (!prm.basic) {
            // 16x16 predictor (A = left q1, B = up q2, C = up-right q2 else up-left q3) on the proxies
            const int aL = mbx > 0, aU = mby > 0, aUR = mby > 0 && mbx < g.Wmb - 1, aUL = mby > 0 && mbx > 0;
            int ax = 0, ay = 0, bx = 0, by = 0, cx = 0, cy = 0, px, py;
            if (aL) proxy_mv(S, mb - 1, 1, ax, ay);
            if (aU) proxy_mv(S, mb - g.Wmb, 2, bx, by);
            if (aUR) proxy_mv(S, mb - g.Wmb + 1, 2, cx, cy);
            else if (aUL) proxy_mv(S, mb - g.Wmb - 1, 3, cx, cy);
            median_pred(aL, ax, ay, aU, bx, by, aUR ? 1 : aUL, cx, cy, px, py);
            gx[0] = px >> 2; gy[0] = py >> 2; n = 1;
        }
        const uint32_t pg = S.prev_gen16[mb];
        if (pg != SPEC_PREV_NONE) {
            const int px = (int16_t)(pg & 0xffffu), py = (int16_t)(pg >> 16);
            if (n == 0) { gx[0] = px; gy[0] = py; }
            else if (px != gx[0] || py != gy[0]) { gx[1] = px; gy[1] = py; }
        }
    }
#pragma unroll
    for (int sl = 0; sl < 2; sl++) {
        ms.cx[sl] = SPEC_NOGUESS; ms.cy[sl] = SPEC_NOGUESS; ms.mask[sl] = 0;
        if (gx[sl] == SPEC_NOGUESS || maxdiff < 0) continue;
        const int X = mbx * 16 + gx[sl], Y = mby * 16 + gy[sl], xa = X & ~15, off = X - xa;
        // the prediction equals the interpolated planes only where every sample position lies inside the picture (phase_c.cuh)
        if (!tmap || X < 0 || xa + QW_ROWB > W || Y < 0 || Y + 16 > H) continue;
        __syncwarp();
        if (lane == 0) { mbar_expect_tx(bar, (uint32_t)SKIPWIN_BYTES); tma_load_window(tmap, win, bar, xa, Y); }
        mbar_wait(bar, phase); phase ^= 1u;
        uint32_t mask = 0;
        const uint32_t *wp = (const uint32_t *)(win + (size_t)r * QW_ROWB) + ((off + 8 * hf) >> 2);
        const uint32_t sh = (uint32_t)((off + 8 * hf) & 3) * 8;
#pragma unroll 4
        for (int f = 0; f < 16; f++) {
            const uint32_t *w = wp + (size_t)f * (16 * QW_ROWB / 4);
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            if (within(make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh)))) mask |= 1u << f;
        }
        ms.cx[sl] = (int16_t)gx[sl]; ms.cy[sl] = (int16_t)gy[sl]; ms.mask[sl] = (uint16_t)mask;
    }
    if (lane == 0) *(uint4 *)&S.mbspec[mb] = *(const uint4 *)&ms;
}
