// Stage 3 of the 8x8 search (moestimation.cpp:508-520), fully parallel over partitions — k_stage3.
//   MEstimation(g = window/2, fraction 0, centre 0) then MEstimation(g = window/16, all 16 fractions, centre 0) into one list:
//   the 33 best by (cost, arrival) and their SADs. Independent of the MV predictor (only the final + |mv - mvp|_1 is not).
// One CTA per macroblock, one warp per 8x8 partition. What round 1's kernel spent its time on, and what replaces it
// (ncu per-line profile, profiles/r02_*):
//   * the second call's pixel rows (768 L1 tag requests per partition) and row sums through shared memory (450 conflicting
//     wavefronts): ONE TMA box per macroblock (48 bytes x (R+9) rows x 16 planes) stages the quarter-pel window of all four
//     partitions; a lane owns one (plane, column) and slides the box sums down it in registers (qwin.cuh).
//   * 1089 feature evaluations of the first call: |s0 - K0| * multiplier is a lower bound of a candidate's cost
//     (the feature distance starts with |s0 - K0|, :267), so a candidate whose bound already exceeds the 33rd smallest cost so
//     far cannot enter the list. The bound needs only the 8x8 sum of the position (a 2-byte plane written by k_features):
//     the window is swept with 2-byte loads, the few survivors (a few per cent on textured content) are evaluated densely.
//   * selection over a 6 KB cost array: streaming selection (topk.cuh).
//   * SADs of quarter-pel list members: read from the staged window.
#pragma once
#include "common.cuh"
#include "phase_a.cuh"
#include "qwin.cuh"
#include "topk.cuh"

#ifndef FH_S3_MINB
#define FH_S3_MINB 5
#endif
#define S3_ROWB 48                       // bytes per staged row of the macroblock's window
#define S3_SURV_CAP 288                  // first-call survivors of one block of rows (8 rows x 33 columns at WindowSize 32)
__host__ __device__ __forceinline__ int s3_win_rows(int g1) { return 8 + 2 * g1 + 8 + 1; }          // R + 8, + 1: odd
__host__ __device__ __forceinline__ int s3_win_bytes(int g1) { return 16 * s3_win_rows(g1) * S3_ROWB; }

struct WinMagic { uint32_t i3, i1; };     // udiv_magic(2*(window/2)+1), udiv_magic(2*(window/16)+1), computed by the host
struct __align__(16) S3WarpV2 {
    TopKBufT<320> tk;
    uint32_t stage[QW_STAGE_WORDS];      // the quarter-pel window's costs until the bound is known (qwin_select)
    uint16_t members[FH_S3_MAX + 3];
    uint16_t msad[FH_S3_MAX + 3];
    uint16_t surv[S3_SURV_CAP];          // arrival indices of the first-call candidates that passed the sum bound
};

__global__ void __launch_bounds__(128, FH_S3_MINB) k_stage3(const SeqDev *__restrict__ seqs, int seq0, Geo g, fh264_params prm, WinMagic wm, const CUtensorMap *__restrict__ tmaps48)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];       // window (128-byte aligned) | mbarrier | 4 x S3WarpV2
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g3 = prm.window / 2, g1 = prm.window / 16;
    const int w3 = 2 * g3 + 1, w1 = 2 * g1 + 1;
    const int wbytes = s3_win_bytes(g1);
    uint8_t *win = smem_raw;
    uint64_t *bar = (uint64_t *)(smem_raw + wbytes);
    S3WarpV2 *sw = (S3WarpV2 *)(smem_raw + wbytes + 16) + warp;
    const SeqDev &S = seqs[seq0 + blockIdx.y];
    if (S.status[ST_GATE]) return;                                   // scene cut: this picture is not coded as P (CTA-uniform)
    const int mb = g.band_mb0 + blockIdx.x;
    const int part = mb * 4 + warp;
    const int W = g.W, H = g.H;
    int mbxM, mbyM;
    mb_xy(g, mb, mbxM, mbyM);
    const int xM = mbxM * 16, yM = mbyM * 16;
    // ---- the macroblock's quarter-pel window: pixels [xM - g1, xM - g1 + R + 8) x [yM - g1, yM - g1 + R + 8) of the 16 planes
    const int R = 8 + w1 - 1, rows = s3_win_rows(g1);
    const int wx0 = xM - g1, wy0 = yM - g1, wxa = wx0 & ~15;
    const bool interior = tmaps48 && wx0 >= 0 && wxa + S3_ROWB <= W && wy0 >= 0 && wy0 + rows <= H;
    int woff;
    if (interior) {
        if (threadIdx.x == 0) { mbar_init(bar, 1); }
        __syncthreads();
        if (threadIdx.x == 0) { mbar_expect_tx(bar, (uint32_t)wbytes); tma_load_window(tmaps48 + seq0 + blockIdx.y, win, bar, wxa, wy0); }
        woff = wx0 - wxa;
    } else {
        // picture border: clamped loads (replicate padding, :107-115); 24 pixels per row as 16 + 16 overlapping bytes
        for (int t = threadIdx.x; t < 16 * (R + 8); t += 128) {
            const int f = t / (R + 8), r = t - f * (R + 8);
            uint8_t *dst = win + (size_t)(f * rows + r) * S3_ROWB;
            *(uint4 *)dst = qf_load16(S.planes + (size_t)f * g.WH, W, H, wx0, wy0 + r);
            *(uint4 *)(dst + 16) = qf_load16(S.planes + (size_t)f * g.WH, W, H, wx0 + 16, wy0 + r);
        }
        woff = 0;
        __syncthreads();
    }
    int xP, yP;
    part_origin(g, part, xP, yP);
    uint2 crow[8];
    load_cur8x8(S.cur[0], g, xP, yP, crow);
    int s[5];
    block_sums(crow, s);
    const FeatQ fq = feat_query(s);
    const int n3a = w3 * w3;
    const uint32_t i3 = wm.i3, i1 = wm.i1;
    const int rlo = max(0, g3 - yP), rhi = min(w3, H - yP + g3), clo = max(0, g3 - xP), chi = min(w3, W - xP + g3);
    const int nva = max(0, chi - clo) * max(0, rhi - rlo);
    const int nvb = max(0, min(W - 1, xP + g1) - max(0, xP - g1) + 1) * max(0, min(H - 1, yP + g1) - max(0, yP - g1) + 1) * 16;
    TopK tk;
    tk_init(tk, min(FH_S3_MAX, nva + nvb), 320);
    // ---- second call first (its costs bound the first call's): MEstimation(window/16, 16 fractions, centre 0); arrival n3a + ...
    const QWinView qv = { win + (size_t)(warp >> 1) * 8 * S3_ROWB, S3_ROWB, rows * S3_ROWB, woff + (warp & 1) * 8 };
    if (interior) mbar_wait(bar, 0);
    qwin_select_w(w1, g, qv, xP, yP, 0, 0, fq, sw->tk.key, tk, sw->stage, (uint32_t)n3a);
    // ---- first call: MEstimation(window/2, fraction 0, centre 0); arrival (dx + g3) * w3 + (dy + g3). Rows [rlo, rhi) and columns
    //      [clo, chi) of the window have their block origin inside the picture (:265).
    const uint16_t *__restrict__ k0p = S.k0p;
    const uint4 *__restrict__ kar = S.kar;
    // Sum bound per block of RB rows: lane = column, one bit per row in a lane mask; the columns beyond the last full group of 32
    // (one column at WindowSize 32) are swept with lane = row. Survivors are compacted once per block, evaluated densely, then
    // the bound tightens for the next block.
    const int RB = max(1, min(8, S3_SURV_CAP / w3));
    const int ncf = clo + ((chi - clo) & ~31);                    // columns [clo, ncf) in full groups of 32
#pragma unroll 1
    for (int rb = rlo; rb < rhi; rb += RB) {
        const int re = min(rhi, rb + RB);
        const uint32_t tcost = (uint32_t)min(tk.bound >> 16, (u64)0xfffffffeu);
        int nsv = 0;
        for (int cb = clo; cb < ncf; cb += 32) {
            const int c = cb + lane, mx = iabs_(c - g3) + 4;
            const uint16_t *kp = k0p + (size_t)(yP - g3 + rb) * W + (xP - g3 + c);
            uint32_t mask = 0;
#pragma unroll 4
            for (int r = rb; r < re; r++, kp += W) {
                const uint32_t lb = (uint32_t)((mx + iabs_(r - g3)) * iabs_(s[0] - (int)__ldg(kp)));
                mask |= (uint32_t)(lb <= tcost) << (r - rb);
            }
            // exclusive prefix of the lanes' survivor counts, then every lane writes its own
            const int cnt = __popc(mask);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t2; }
            int pos = nsv + incl - cnt;
            while (mask) { const int b = __ffs(mask) - 1; mask &= mask - 1; sw->surv[FH_IDX(pos, S3_SURV_CAP)] = (uint16_t)(c * w3 + rb + b); pos++; }
            nsv += __shfl_sync(0xffffffffu, incl, 31);
        }
        for (int c = ncf; c < chi; c++) {                            // leftover columns (one at WindowSize 32): lane = row of the block
            const int r = rb + lane;
            bool pass = false;
            if (r < re) {
                const uint32_t lb = (uint32_t)((iabs_(c - g3) + iabs_(r - g3) + 4) * iabs_(s[0] - (int)__ldg(k0p + (size_t)(yP - g3 + r) * W + (xP - g3 + c))));
                pass = lb <= tcost;
            }
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if (pass) sw->surv[FH_IDX(nsv + __popc(m & ((1u << lane) - 1u)), S3_SURV_CAP)] = (uint16_t)(c * w3 + r);
            nsv += __popc(m);
        }
        __syncwarp();
#pragma unroll 1
        for (int i0 = 0; i0 < nsv; i0 += 32) {
            const int i = i0 + lane;
            uint32_t cst = TK_NONE, idx = 0;
            if (i < nsv) {
                idx = sw->surv[i];
                const int c = udiv_by((int)idx, i3), r = (int)idx - c * w3;
                const uint4 rec = __ldg(kar + (size_t)(yP - g3 + r) * W + (xP - g3 + c));
                cst = (uint32_t)((iabs_(c - g3) + iabs_(r - g3) + 4) * feat_of(fq, rec));
            }
            tk_offer(sw->tk.key, tk, cst, idx);
        }
        tk_tighten(tk);
        __syncwarp();
    }
    const int nm = tk_finish(sw->tk.key, tk, nva + nvb, sw->members);
    // ---- SADs of the members (satdLuma8x8MVs, :175-195): 8 lanes per member, one row each. Quarter-pel window members read the
    //      staged window (their block lies inside it; same clamping as the reference for an origin inside the picture).
    const int r8 = lane & 7;
    const uint2 cr = pick_row(crow, r8);
#pragma unroll 1
    for (int base = 0; base < nm; base += 4 * FH_S3_SADR) {
        uint2 rr[FH_S3_SADR];
#pragma unroll
        for (int u = 0; u < FH_S3_SADR; u++) {
            const int m = base + u * 4 + (lane >> 3);
            rr[u] = make_uint2(0, 0);
            if (m < nm) {
                const int i = (int)sw->members[m];
                if (i >= n3a) {
                    const int j = i - n3a, f = j & 15, pos = j >> 4, cx = udiv_by(pos, i1), cy = pos - cx * w1;
                    rr[u] = qwin_row8(qv, f, cx, cy + r8);
                } else {
                    const int c = udiv_by(i, i3), r = i - c * w3;
                    rr[u] = load_row8(S.planes, W, H, xP + c - g3, yP + r - g3 + r8);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < FH_S3_SADR; u++) {
            const int m = base + u * 4 + (lane >> 3);
            int sad = m < nm ? sad8(cr, rr[u]) : 0;
            sad += __shfl_xor_sync(0xffffffffu, sad, 1);
            sad += __shfl_xor_sync(0xffffffffu, sad, 2);
            sad += __shfl_xor_sync(0xffffffffu, sad, 4);
            if (m < nm && r8 == 0) sw->msad[m] = (uint16_t)sad;
        }
    }
    __syncwarp();
    for (int m = lane; m < nm; m += 32) {
        int dx, dy, f;
        s3_decode((int)sw->members[m], n3a, w3, g3, i3, w1, g1, i1, dx, dy, f);
        S3Entry e;
        e.mvx = (int16_t)((dx << 2) | (f & 3)); e.mvy = (int16_t)((dy << 2) | (f >> 2)); e.sad = sw->msad[m]; e.pad = 0;
        S.s3[(size_t)part * FH_S3_MAX + m] = e;
    }
    // the list's best vector by SAD (first in list order): phase S guesses the neighbours' final vectors with it (spec.cuh)
    {
        uint32_t k = 0xffffffffu;
        for (int m = lane; m < nm; m += 32) k = min(k, ((uint32_t)sw->msad[m] << 8) | (uint32_t)m);
        k = __reduce_min_sync(0xffffffffu, k);
        if (lane == 0) {
            uint32_t px = 0x7f7f7f7fu;
            if (nm > 0) {
                int dx, dy, f;
                s3_decode((int)sw->members[k & 255u], n3a, w3, g3, i3, w1, g1, i1, dx, dy, f);
                px = ((uint32_t)((dx << 2) | (f & 3)) & 0xffffu) | ((uint32_t)((dy << 2) | (f >> 2)) << 16);
            }
            S.proxy[part] = px;
            PartA *pa = &S.parta[part];
#pragma unroll
            for (int q = 0; q < 5; q++) pa->suma[q] = (uint16_t)s[q];
            pa->n3 = (uint16_t)nm;
        }
    }
}
