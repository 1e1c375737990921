// Phase B, warp-level kernel — the predictor-dependent part of interEncoding (moestimation.cpp:392-570) as a wavefront of
// single-warp CTAs.
//
// With phase S in front (spec.cuh) every decision of a macroblock is, for ~99.7 % of the partitions, a lookup in the phase-S
// products (the warp-level fast path that k_phase_b already had). What made k_phase_b expensive to keep resident — 128 threads x
// 128 registers and 24 KB of shared memory per CTA, all for the block-level full search of the remaining 0.3 % — is not needed
// for that: a partition whose predictor falls outside the guessed cells reruns phase S's own warp-level search (spec_collect)
// for the TRUE gen = mvp >> 2 right here and picks the winner among the candidates it leaves; a P_Skip vector outside the
// guessed cells is measured by the warp. A CTA is one warp (128 registers, ~11 KB of shared memory), so a picture's wavefront
// (a latency-bound chain through L2) occupies a fraction of an SM slot each and the search kernels of OTHER pictures run under it
// (the pipeline lanes of fh264_b200.cu). Pictures with partitions whose stage-2 set phase A could not store (S2_SLOW: flat
// content, enumerated by the whole block) and sessions with the timeline tap on stay with k_phase_b (`seq_is_heavy`), which
// is launched right behind this kernel and returns at once when no sequence of the call needs it.
// Dependencies, publication (tagged quadrant words), merge and mvd are the same as in k_phase_b (phase_b.cuh).
#pragma once
#include "phase_b.cuh"
#include "spec.cuh"

#ifndef PBW_MINB
#define PBW_MINB 16
#endif

struct __align__(128) PBWShared {
    SpecWarp sw;                         // spec_collect's work area (mbarrier of the window included)
    __align__(16) PartSpec spec[4];      // phase-S finalists of the four partitions
    __align__(16) MbSpec ms;             // phase-S P_Skip trials of the macroblock
};
__host__ __device__ __forceinline__ size_t pbw_smem_bytes(int g1) { return (size_t)((qwin_bytes(g1) + 127) & ~127) + sizeof(PBWShared); }

__global__ void __launch_bounds__(32, PBW_MINB) k_phase_b_warp(const SeqDev *__restrict__ seqs, int seq0, int nseq, Geo g, fh264_params prm, uint32_t epoch,
                                                               const int *__restrict__ wf_order, uint32_t *__restrict__ ticket, WinMagic wm,
                                                               const CUtensorMap *__restrict__ tmaps, int force_miss)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];       // window (128-byte aligned) | PBWShared
    const int lane = threadIdx.x;
    const int g1 = prm.window / 16;
    uint8_t *win = smem_raw;
    PBWShared &sh = *(PBWShared *)(smem_raw + ((qwin_bytes(g1) + 127) & ~127));
    {
        bool any = false;
        for (int b = 0; b < nseq; b++) any |= !seq_is_heavy(seqs[seq0 + b]) && !seqs[seq0 + b].status[ST_GATE];
        if (!any) return;
    }
    if (lane == 0) mbar_init(&sh.sw.bar, 1);
    __syncwarp();
    uint32_t phase = 0;
    const uint32_t total = (uint32_t)g.band_nmb * (uint32_t)nseq;
    const int W = g.W, H = g.H;
    for (;;) {
        __syncwarp();
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= total) return;
        const int sq = seq0 + (int)(t % (uint32_t)nseq);
        const SeqDev &S = seqs[sq];
        if (S.status[ST_GATE] || seq_is_heavy(S)) continue;
        const int mb = wf_order[t / (uint32_t)nseq];
        int mbx, mby;
        mb_xy(g, mb, mbx, mby);
        const CUtensorMap *tmap = tmaps ? tmaps + sq : nullptr;
        // ---- what does not depend on the neighbours is fetched before waiting on them
        ((uint4 *)sh.spec)[lane] = ((const uint4 *)&S.spec[(size_t)mb * 4])[lane];                 // 4 x 128 bytes
        if (lane == 0) *(uint4 *)&sh.ms = *(const uint4 *)&S.mbspec[mb];
        // ---- the row above: up q2 / q3, up-right q2, up-left q3, each polled by its own lane (A.7)
        const bool sysw = g.world > 1;
        const bool mirror = S.peer_qmv_next != nullptr && mby == (g.band_mb0 + g.band_nmb) / g.Wmb - 1;
        const int aL = mbx > 0, aU = mby > 0, aUR = mby > 0 && mbx < g.Wmb - 1, aUL = mby > 0 && mbx > 0;
        int vx = 0, vy = 0;
        {
            const int nmb = lane < 2 ? mb - g.Wmb : (lane == 2 ? mb - g.Wmb + 1 : mb - g.Wmb - 1);
            const int q = lane == 0 ? 2 : (lane == 1 ? 3 : (lane == 2 ? 2 : 3));
            const bool used = lane < 2 ? aU : (lane == 2 ? aUR : (lane == 3 ? aUL : false));
            if (used && !wait_qmv(&S.qmv[(size_t)nmb * 4 + q], epoch, sysw, vx, vy)) atomicOr(&S.status[ST_FLAGS], FLAG_TIMEOUT);
        }
        const int u2x = __shfl_sync(0xffffffffu, vx, 0), u2y = __shfl_sync(0xffffffffu, vy, 0);
        const int u3x = __shfl_sync(0xffffffffu, vx, 1), u3y = __shfl_sync(0xffffffffu, vy, 1);
        const int r2x = __shfl_sync(0xffffffffu, vx, 2), r2y = __shfl_sync(0xffffffffu, vy, 2);
        const int d3x = __shfl_sync(0xffffffffu, vx, 3), d3y = __shfl_sync(0xffffffffu, vy, 3);
        __syncwarp();                                                    // sh.spec / sh.ms complete

        int l1x = 0, l1y = 0, l3x = 0, l3y = 0;
        bool have1 = !aL, have3 = !aL;
        auto poll_left = [&](int q, int &ox, int &oy) {
            int x = 0, y = 0;
            if (lane == 0 && !wait_qmv(&S.qmv[(size_t)(mb - 1) * 4 + q], epoch, false, x, y)) atomicOr(&S.status[ST_FLAGS], FLAG_TIMEOUT);
            ox = __shfl_sync(0xffffffffu, x, 0); oy = __shfl_sync(0xffffffffu, y, 0);
        };
        auto publish = [&](int q, int px, int py) {
            if (lane == 0) {
                const unsigned long long wq = qmv_word(epoch, px, py);
                st_relaxed_u64(&S.qmv[(size_t)mb * 4 + q], wq, false);
                if (mirror) st_relaxed_u64(&S.peer_qmv_next[(size_t)mb * 4 + q], wq, true);
            }
        };
        const MbSpec &ms = sh.ms;
        const int maxdiff = ms.maxdiff;
        // P_Skip trial for vector (sx, sy): the phase-S masks, else measured here (:228-244; every |cur - pred| <= MAXDIFF over the
        // 256 luma samples, prediction with the per-sample clamp of the motion compensation). 0: skips.
        auto skip_nbad = [&](int sx, int sy) -> int {
            const int k = force_miss ? -1 : skip_lookup(ms, sx, sy);      // (force_miss: test knob — every lookup is treated as a miss)
            if (k >= 0) return k;
            const int r = lane >> 1, hf = lane & 1;
            const uint2 c = __ldg((const uint2 *)(S.cur[0] + (size_t)(mby * 16 + r) * W + mbx * 16 + 8 * hf));
            int p[8];
            luma_pred_block<8, 1>(S, g, mbx * 16 + 8 * hf + (sx >> 2), mby * 16 + r + (sy >> 2), sx & 3, sy & 3, p);
            bool bad = false;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int cv = (int)(((i < 4 ? c.x : c.y) >> (8 * (i & 3))) & 255u);
                bad |= iabs_(cv - p[i]) > maxdiff;
            }
            return __any_sync(0xffffffffu, bad) ? 1 : 0;
        };
        // ---- P_Skip (mode_pred.cpp:383-401, moestimation.cpp:402-425)
        int smx = 0, smy = 0, nbad = 1;
        if (!aL || mby == 0 || (u2x == 0 && u2y == 0)) nbad = skip_nbad(0, 0);
        else {
            const int cx16 = aUR ? r2x : d3x, cy16 = aUR ? r2y : d3y;
            if (u2x == cx16 && u2y == cy16) {
                // B == C: the skip vector is B, or zero if the left quadrant turns out to be zero — both are tried without waiting
                const int nbB = skip_nbad(u2x, u2y), nb0 = skip_nbad(0, 0);
                if (lane == 0) S.prev_gen16[mb] = ((uint32_t)(u2x >> 2) & 0xffffu) | ((uint32_t)(u2y >> 2) << 16);
                if (nbB != 0 && nb0 != 0) nbad = 1;
                else {
                    poll_left(1, l1x, l1y); have1 = true;
                    const bool lz = l1x == 0 && l1y == 0;
                    smx = lz ? 0 : u2x; smy = lz ? 0 : u2y; nbad = lz ? nb0 : nbB;
                }
            } else {
                poll_left(1, l1x, l1y); have1 = true;
                if (!(l1x == 0 && l1y == 0)) {
                    median_pred(1, l1x, l1y, 1, u2x, u2y, 1, cx16, cy16, smx, smy);      // A = left q1, B = up q2, C = up-right q2 else up-left q3: all available here
                    if (lane == 0) S.prev_gen16[mb] = ((uint32_t)(smx >> 2) & 0xffffu) | ((uint32_t)(smy >> 2) << 16);
                }
                nbad = skip_nbad(smx, smy);
            }
        }
        if (nbad == 0) {
            for (int q = 0; q < 4; q++) publish(q, smx, smy);
            if (lane == 0) {
                MbMotion mo;
                mo.maxdiff = (int16_t)maxdiff; mo.pad = 0; mo.mb_type = FH264_P_SKIP; mo.num_parts = 0;
                for (int i = 0; i < 4; i++) { mo.mv[i][0] = (int16_t)smx; mo.mv[i][1] = (int16_t)smy; mo.mvd[i][0] = mo.mvd[i][1] = 0; mo.sad[i] = 0; }
                uint4 *d = (uint4 *)&S.motion[mb];
                const uint4 *s4 = (const uint4 *)&mo;
                d[0] = s4[0]; d[1] = s4[1]; d[2] = s4[2];
                atomicAdd(&S.status[ST_COUNTS + 0], 1u);
            }
            continue;
        }
        // ---- the four 8x8 partitions in order (:430-528); every final quadrant vector is published at once
        int q0x = 0, q0y = 0, q1x = 0, q1y = 0, q2x = 0, q2y = 0, q3x = 0, q3y = 0, s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        int p0x = 0, p0y = 0, p1x = 0, p1y = 0, p2x = 0, p2y = 0, p3x = 0, p3y = 0;
        int nhit = 0;
#pragma unroll 1
        for (int pi = 0; pi < 4; pi++) {
            int px, py, bx = 0, by = 0, bs = 0;
            // partition 0 predicts from left q1 unless up q2 == up q3; partition 2 from left q3 unless own q0 == q1
            if (pi == 0) {
                if (!have1 && !(aU && u2x == u3x && u2y == u3y)) { poll_left(1, l1x, l1y); have1 = true; }
                median_pred(aL, l1x, l1y, aU, u2x, u2y, aU ? 1 : aUL, aU ? u3x : d3x, aU ? u3y : d3y, px, py);
            } else if (pi == 1) median_pred(1, q0x, q0y, aU, u3x, u3y, aUR ? 1 : aU, aUR ? r2x : u2x, aUR ? r2y : u2y, px, py);
            else if (pi == 2) {
                if (!have3 && !(q0x == q1x && q0y == q1y)) { poll_left(3, l3x, l3y); have3 = true; }
                median_pred(aL, l3x, l3y, 1, q0x, q0y, 1, q1x, q1y, px, py);
            } else median_pred(1, q2x, q2y, 1, q1x, q1y, 1, q0x, q0y, px, py);
            // the winner for predictor (px, py): lookup among the phase-S finalists, else the full search for the true gen
            const int genx = px >> 2, geny = py >> 2;
            if (!force_miss && spec_lookup_w(sh.spec[pi], genx, geny, px, py, bx, by, bs)) nhit++;
            else {
                const int part = mb * 4 + pi, xP = mbx * 16 + (pi & 1) * 8, yP = mby * 16 + (pi >> 1) * 8;
                uint2 rows[8];
                load_cur8x8(S.cur[0], g, xP, yP, rows);
                int s[5];
                block_sums(rows, s);
                const FeatQ fq = feat_query(s);
                int n3 = 0; uint32_t n2w = 0, s2_off = 0;
                if (!prm.basic) { const PartA pa = S.parta[part]; n3 = pa.n3; n2w = pa.n2; s2_off = pa.s2_off; }
                __syncwarp();
                for (int i = lane; i < n3; i += 32) sh.sw.s3[i] = S.s3[(size_t)part * FH_S3_MAX + i];
                __syncwarp();
                bool usable;
                const int n = spec_collect(S, g, prm, xP, yP, rows, fq, n3, n2w, s2_off, genx, geny, &sh.sw, win, tmap, phase, wm, usable);
                // first strict minimum of SAD + |mv - mvp|_1 in (stage, list position) order (:460-469,498-507,511-520)
                u64 k = KEY_NONE;
                for (int i = lane; i < n; i += 32) {
                    const uint32_t mv = sh.sw.pf_mv[i], so = sh.sw.pf_so[i];
                    const int cx = (int)(int16_t)(mv & 0xffffu), cy = (int)(int16_t)(mv >> 16);
                    k = min(k, ((u64)((int)(so & 0xffffu) + mv_cost(cx, cy, px, py)) << 32) | ((u64)(so >> 16) << 16) | (u64)i);
                }
                k = warp_min_u64(k);
                if (k != KEY_NONE) {
                    const uint32_t mv = sh.sw.pf_mv[(int)(k & 0xffffu)], so = sh.sw.pf_so[(int)(k & 0xffffu)];
                    bx = (int)(int16_t)(mv & 0xffffu); by = (int)(int16_t)(mv >> 16); bs = (int)(so & 0xffffu);
                } else {
                    // no candidate at all (:452): zero vector, SAD measured at the block's own position
                    int sad = lane < 8 ? sad_row8(pick_row(rows, lane & 7), S.planes, W, H, xP, yP + lane) : 0;
                    sad += __shfl_xor_sync(0xffffffffu, sad, 1);
                    sad += __shfl_xor_sync(0xffffffffu, sad, 2);
                    sad += __shfl_xor_sync(0xffffffffu, sad, 4);
                    bx = 0; by = 0; bs = __shfl_sync(0xffffffffu, sad, 0);
                }
                __syncwarp();
            }
            publish(pi, bx, by);
            if (pi == 0) { q0x = bx; q0y = by; s0 = bs; p0x = px; p0y = py; }
            else if (pi == 1) { q1x = bx; q1y = by; s1 = bs; p1x = px; p1y = py; }
            else if (pi == 2) { q2x = bx; q2y = by; s2 = bs; p2x = px; p2y = py; }
            else { q3x = bx; q3y = by; s3 = bs; p3x = px; p3y = py; }
        }
        // ---- merge (:529-551) and final mvd with the merged type's predictors (:552-564)
        const bool eq01 = q0x == q1x && q0y == q1y, eq23 = q2x == q3x && q2y == q3y;
        const bool eq02 = q0x == q2x && q0y == q2y, eq13 = q1x == q3x && q1y == q3y;
        if ((eq01 && eq23) || (eq02 && eq13)) {               // 16x16 / 16x8 / 8x16: the predictors read the left macroblock
            if (!have1) { poll_left(1, l1x, l1y); have1 = true; }
            if (!have3) { poll_left(3, l3x, l3y); have3 = true; }
        }
        if (lane == 0) {
            NbCache nc;
            nc.avail[0] = aL; nc.avail[1] = aU; nc.avail[2] = aUR; nc.avail[3] = aUL;
            for (int w = 0; w < 4; w++) for (int q = 0; q < 4; q++) { nc.mvx[w][q] = 0; nc.mvy[w][q] = 0; }
            nc.mvx[0][1] = l1x; nc.mvy[0][1] = l1y; nc.mvx[0][3] = l3x; nc.mvy[0][3] = l3y;
            nc.mvx[1][2] = u2x; nc.mvy[1][2] = u2y; nc.mvx[1][3] = u3x; nc.mvy[1][3] = u3y;
            nc.mvx[2][2] = r2x; nc.mvy[2][2] = r2y; nc.mvx[3][3] = d3x; nc.mvy[3][3] = d3y;
            const int mv[4][2] = { { q0x, q0y }, { q1x, q1y }, { q2x, q2y }, { q3x, q3y } };
            const int mvps[4][2] = { { p0x, p0y }, { p1x, p1y }, { p2x, p2y }, { p3x, p3y } };
            const int sadq[4] = { s0, s1, s2, s3 };
            int type = FH264_P_8x8ref0, nparts = 4, cnt = 4;
            if (eq01 && eq23 && eq02) { type = FH264_P_L0_16x16; nparts = 1; cnt = 1; }
            else if (eq01 && eq23) { type = FH264_P_L0_L0_16x8; nparts = 2; cnt = 2; }
            else if (eq02 && eq13) { type = FH264_P_L0_L0_8x16; nparts = 2; cnt = 3; }
            MbMotion mo;
            mo.maxdiff = (int16_t)maxdiff; mo.pad = 0;
            int fin[4][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };
            for (int i = 0; i < 4; i++) { mo.mvd[i][0] = mo.mvd[i][1] = 0; }
            for (int i = 0; i < nparts; i++) {
                int ppx = 0, ppy = 0, pw = 16, dir = 0, qsel = i, ox, oy;
                if (type == FH264_P_L0_L0_16x8) { ppy = i * 8; dir = i == 0 ? 1 : 2; qsel = i * 2; }
                else if (type == FH264_P_L0_L0_8x16) { ppx = i * 8; pw = 8; dir = i == 0 ? 2 : 3; }
                else if (type == FH264_P_8x8ref0) { ppx = (i & 1) * 8; ppy = (i >> 1) * 8; pw = 8; }
                if (type == FH264_P_8x8ref0) { ox = mvps[i][0]; oy = mvps[i][1]; }        // same neighbours as during the search
                else predict_mv_(nc, ppx, ppy, pw, dir, fin, ox, oy);
                mo.mvd[i][0] = (int16_t)(mv[qsel][0] - ox); mo.mvd[i][1] = (int16_t)(mv[qsel][1] - oy);
                for (int q = 0; q < 4; q++) {
                    const bool in = type == FH264_P_L0_16x16 || (type == FH264_P_L0_L0_16x8 && (q >> 1) == i) ||
                                    (type == FH264_P_L0_L0_8x16 && (q & 1) == i) || (type == FH264_P_8x8ref0 && q == i);
                    if (in) { fin[q][0] = mv[qsel][0]; fin[q][1] = mv[qsel][1]; }
                }
            }
            mo.mb_type = (int16_t)type; mo.num_parts = (int16_t)nparts;
            for (int q = 0; q < 4; q++) { mo.mv[q][0] = (int16_t)fin[q][0]; mo.mv[q][1] = (int16_t)fin[q][1]; mo.sad[q] = (uint16_t)sadq[q]; }
            uint4 *d = (uint4 *)&S.motion[mb];
            const uint4 *s4 = (const uint4 *)&mo;
            d[0] = s4[0]; d[1] = s4[1]; d[2] = s4[2];
            for (int i = 0; i < 4; i++) S.prev_gen[(size_t)mb * 4 + i] = ((uint32_t)(mvps[i][0] >> 2) & 0xffffu) | ((uint32_t)(mvps[i][1] >> 2) << 16);
            atomicAdd(&S.status[ST_COUNTS + cnt], 1u);
            if (nhit) atomicAdd(&S.status[ST_SPEC_HIT], (uint32_t)nhit);
            if (nhit < 4) atomicAdd(&S.status[ST_SPEC_MISS], (uint32_t)(4 - nhit));
        }
    }
}
