// Phase B — the predictor-dependent part of interEncoding (moestimation.cpp:392-570), a wavefront over
// macroblocks: every cost uses the median MV predictor of the already decided left / up / up-right / up-left
// neighbours (mode_pred.cpp:252-371). One CTA (128 threads) per macroblock; CTAs draw tickets in anti-diagonal
// order (x + 2y), interleaved over the sequences of the batch, and spin on the `done` epochs of the two
// neighbours that dominate the dependency set. A ticket's dependencies always hold smaller tickets, so the
// smallest unfinished ticket can always run: no co-residency assumption, no deadlock.
// Per MB: P_Skip test (mode_pred.cpp:383-401, moestimation.cpp:402-425), then per 8x8 partition: predictor,
// stage 1 (window/16 quarter-pel window around the predictor, 17 best by feature cost -> SAD), ranking of the
// phase-A stage-2 set with the now known multiplier (33 best -> SAD looked up), the phase-A stage-3 list;
// winner = first strict minimum of SAD + |mv - mvp|_1 in list order; then merge and final mvd (:529-564).
#pragma once
#include "common.cuh"
#include "select.cuh"
#include "phase_a.cuh"
#include "phase_c.cuh"

#define PB_NT 128
#define S1_KEY_CAP 1296     // (2*4+1)^2*16 at WindowSize 64

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

struct NbCache {            // quadrant MVs of the four neighbouring macroblocks (A.7): 0 left, 1 up, 2 up-right, 3 up-left
    int mvx[4][4], mvy[4][4];
    int avail[4];
};

struct Nb { int avail, mvx, mvy; };

// DeriveNeighbourLocation + get_neighbour_mv (mode_pred.cpp:48-97), quadrant formulation.
__device__ __forceinline__ Nb neighbour_mv(const NbCache &nc, int xN, int yN, const int cur[4][2])
{
    Nb n = { 0, 0, 0 };
    if ((xN > 15 && yN >= 0) || yN > 15) return n;
    if (xN >= 0 && xN < 16 && yN >= 0) { const int q = (yN >> 3) * 2 + (xN >> 3); n.avail = 1; n.mvx = cur[q][0]; n.mvy = cur[q][1]; return n; }
    int w;
    if (yN < 0) { w = xN > 15 ? 2 : (xN < 0 ? 3 : 1); if (xN > 15) xN -= 16; else if (xN < 0) xN += 16; yN += 16; }
    else { w = 0; xN += 16; }
    if (!nc.avail[w]) return n;
    const int q = (yN >> 3) * 2 + (xN >> 3);
    n.avail = 1; n.mvx = nc.mvx[w][q]; n.mvy = nc.mvy[w][q];
    return n;
}

__device__ __forceinline__ int median3_(int a, int b, int c) { return max(min(a, b), min(c, max(a, b))); }

// PredictMV_Luma (mode_pred.cpp:252-332). Every available neighbour of a P picture is inter with refIdx 0.
// dir: 0 median, 1 prefer B (16x8 top), 2 prefer A (16x8 bottom / 8x16 left), 3 prefer C (8x16 right).
__device__ __forceinline__ void predict_mv_(const NbCache &nc, int px, int py, int pw, int dir, const int cur[4][2], int &ox, int &oy)
{
    Nb A = neighbour_mv(nc, px - 1, py, cur), B = neighbour_mv(nc, px, py - 1, cur), C = neighbour_mv(nc, px + pw, py - 1, cur);
    if (!C.avail) C = neighbour_mv(nc, px - 1, py - 1, cur);
    if (dir == 1 && B.avail) { ox = B.mvx; oy = B.mvy; return; }
    if (dir == 2 && A.avail) { ox = A.mvx; oy = A.mvy; return; }
    if (dir == 3 && C.avail) { ox = C.mvx; oy = C.mvy; return; }
    int sa = A.avail, sb = B.avail, sc = C.avail;       // "same reference" flags
    if (!A.avail && !B.avail) { A.mvx = A.mvy = 0; sa = 1; A.avail = 1; }         // :299-302
    else if (!A.avail) { A.mvx = A.mvy = 0; sa = 0; A.avail = 1; }                // :303-306
    if (!B.avail) { B = A; sb = sa; }                                             // :307-310
    if (!C.avail) { C = A; sc = sa; }                                             // :311-314
    if (sa + sb + sc == 1) { const Nb &o = sa ? A : (sb ? B : C); ox = o.mvx; oy = o.mvy; return; }
    ox = median3_(A.mvx, B.mvx, C.mvx); oy = median3_(A.mvy, B.mvy, C.mvy);
}

__device__ __forceinline__ int mv_cost(int mvx, int mvy, int px, int py) { return iabs_(mvx - px) + iabs_(mvy - py); }

__global__ void __launch_bounds__(PB_NT) k_phase_b(const SeqDev *__restrict__ seqs, int seq0, int nseq, Geo g, fh264_params prm,
                                                   uint32_t epoch, const int *__restrict__ wf_order, uint32_t *__restrict__ ticket)
{
    __shared__ uint32_t keys1[S1_KEY_CAP];
    __shared__ unsigned long long keys2[1024];
    __shared__ __align__(16) uint8_t cur[16][16];
    __shared__ SelectScratch sc;
    __shared__ NbCache nc;
    __shared__ unsigned long long best;
    __shared__ uint32_t my_ticket;
    __shared__ int red[4];
    __shared__ int n_valid1;
    const int tid = threadIdx.x;
    if (tid == 0) my_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t t = my_ticket;
    const SeqDev &S = seqs[seq0 + (int)(t % (uint32_t)nseq)];
    const int mb = wf_order[t / (uint32_t)nseq];
    const int mbx = mb % g.Wmb, mby = mb / g.Wmb;
    const int W = g.W, H = g.H;

    // ---- wait for the dependencies (left; up-right, or up in the last column) -------------------------------
    if (tid == 0) {
        if (mbx > 0) while (ld_acquire_u32(&S.done[mb - 1]) != epoch) __nanosleep(40);
        if (mby > 0) {
            const int d = mbx < g.Wmb - 1 ? mb - g.Wmb + 1 : mb - g.Wmb;
            while (ld_acquire_u32(&S.done[d]) != epoch) __nanosleep(40);
        }
    }
    __syncthreads();
    if (tid < 16) {
        const int w = tid >> 2, q = tid & 3;
        const int nmb = w == 0 ? mb - 1 : (w == 1 ? mb - g.Wmb : (w == 2 ? mb - g.Wmb + 1 : mb - g.Wmb - 1));
        const bool av = w == 0 ? mbx > 0 : (w == 1 ? mby > 0 : (w == 2 ? (mby > 0 && mbx < g.Wmb - 1) : (mby > 0 && mbx > 0)));
        int vx = 0, vy = 0;
        if (av) { const int v = __ldcg((const int *)&S.motion[nmb].mv[q][0]); vx = (int16_t)(v & 0xffff); vy = v >> 16; }
        nc.mvx[w][q] = vx; nc.mvy[w][q] = vy;
        if (q == 0) nc.avail[w] = av;
    }
    if (tid < 16) *(uint4 *)&cur[tid][0] = *(const uint4 *)(S.cur[0] + (size_t)(mby * 16 + tid) * W + mbx * 16);
    __syncthreads();

    // ---- P_Skip trial ---------------------------------------------------------------------------------------
    int zero4[4][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };
    int smx = 0, smy = 0;
    if (mbx > 0 && mby > 0 && !(nc.mvx[1][2] == 0 && nc.mvy[1][2] == 0) && !(nc.mvx[0][1] == 0 && nc.mvy[0][1] == 0))
        predict_mv_(nc, 0, 0, 16, 0, zero4, smx, smy);                    // mode_pred.cpp:383-401
    const int py = tid >> 3, px = (tid & 7) * 2;                          // this thread's two luma samples
    int p2[2];
    luma_pred_block<2, 1>(S, g, mbx * 16 + px + (smx >> 2), mby * 16 + py + (smy >> 2), smx & 3, smy & 3, p2);
    const int c0 = cur[py][px], c1 = cur[py][px + 1];
    int maxdiff = prm.maxdiff_set;
    if (prm.maxdiff_set == -1) {                                          // moestimation.cpp:407-419
        int v = c0 + c1;
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        const int mean = (red[0] + red[1] + red[2] + red[3]) / 256;
        __syncthreads();
        v = iabs_(c0 - mean) + iabs_(c1 - mean);
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        maxdiff = max(3, (red[0] + red[1] + red[2] + red[3]) / 256);
    }
    const int nbad = __syncthreads_count(iabs_(c0 - p2[0]) > maxdiff || iabs_(c1 - p2[1]) > maxdiff);
    MbMotion mo;
    mo.maxdiff = (int16_t)maxdiff; mo.pad = 0;
    if (nbad == 0) {
        if (tid == 0) {
            mo.mb_type = FH264_P_SKIP; mo.num_parts = 0;
            for (int i = 0; i < 4; i++) { mo.mv[i][0] = (int16_t)smx; mo.mv[i][1] = (int16_t)smy; mo.mvd[i][0] = mo.mvd[i][1] = 0; mo.sad[i] = 0; }
            uint4 *d = (uint4 *)&S.motion[mb];
            const uint4 *s4 = (const uint4 *)&mo;
            d[0] = s4[0]; d[1] = s4[1]; d[2] = s4[2];
            atomicAdd(&S.status[ST_COUNTS + 0], 1u);
            __threadfence();
            st_release_u32(&S.done[mb], epoch);
        }
        return;
    }

    // ---- 8x8 search, partitions in order ----------------------------------------------------------------------
    int mv[4][2], sadq[4], curq[4][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };
    const int g1 = prm.window / 16, w1 = 2 * g1 + 1, n1 = w1 * w1 * 16;
    for (int pi = 0; pi < 4; pi++) {
        const int part = mb * 4 + pi;
        const int xP = mbx * 16 + (pi & 1) * 8, yP = mby * 16 + (pi >> 1) * 8;
        int mvpx, mvpy;
        predict_mv_(nc, (pi & 1) * 8, (pi >> 1) * 8, 8, 0, curq, mvpx, mvpy);
        const int genx = mvpx >> 2, geny = mvpy >> 2;
        PartA pa = S.parta[part];
        if (prm.basic) { pa.n2 = 0; pa.n3 = 0; pa.s2_off = 0; }
        int s[5];
        uint2 rows[8];
#pragma unroll
        for (int r = 0; r < 8; r++) rows[r] = *(const uint2 *)&cur[(pi >> 1) * 8 + r][(pi & 1) * 8];
        block_sums(rows, s);          // suma[0..4] (:440-451); phase A is skipped with BasicInterEncoding
        if (tid == 0) { best = ~0ull; n_valid1 = 0; }
        __syncthreads();

        // stage 1 (:458-469): cost key = cost << 11 | arrival index
        int valid = 0;
        for (int i = tid; i < n1; i += PB_NT) {
            const int f = i & 15, pos = i >> 4, ox = pos / w1 - g1, oy = pos % w1 - g1;
            const int rx = xP + genx + ox, ry = yP + geny + oy;
            uint32_t key = COST_INVALID;
            if (rx >= 0 && rx < W && ry >= 0 && ry < H) { key = ((uint32_t)((iabs_(ox) + iabs_(oy) + 4) * feat_at(S.kar, g, s, f, rx, ry)) << 11) | (uint32_t)i; valid++; }
            keys1[i] = key;
        }
        if (valid) atomicAdd(&n_valid1, valid);
        __syncthreads();
        const int K1 = min(FH_S1_MAX, n_valid1);
        if (K1 > 0) {
            int lt;
            const uint32_t T1 = block_kth_smallest<uint32_t, PB_NT>(n1, K1, 32, [&](int i) { return keys1[i]; }, &sc, &lt);
            // SAD of the K1 members: scan keys, every member is picked up by the thread that owns its slot;
            // 8 rows by the same thread (K1 <= 17 members spread over the block)
            for (int i = tid; i < n1; i += PB_NT) {
                const uint32_t key = keys1[i];
                if (key <= T1) {
                    const int f = i & 15, pos = i >> 4, dx = genx + pos / w1 - g1, dy = geny + pos % w1 - g1;
                    const int mvx = (dx << 2) | (f & 3), mvy = (dy << 2) | (f >> 2);
                    const uint8_t *pl = S.planes + (size_t)f * g.WH;
                    const int x0 = clampi_(xP + dx, 0, W - 1), y0 = clampi_(yP + dy, 0, H - 1);
                    int sad = 0;
#pragma unroll
                    for (int r = 0; r < 8; r++) sad += sad_row8(rows[r], pl, W, H, x0, y0 + r);
                    const unsigned long long fk = ((unsigned long long)(sad + mv_cost(mvx, mvy, mvpx, mvpy)) << 44) | (unsigned long long)key;
                    atomicMin(&best, fk);
                }
            }
        }
        if (!prm.basic) {
            // stage 2 (:470-507): rank the phase-A set with the predictor-dependent multiplier
            const int n2 = (int)pa.n2;
            const uint2 *pool = S.s2pool + pa.s2_off;
            for (int i = tid; i < n2; i += PB_NT) {
                const uint2 v = __ldg(&pool[i]);
                const int dx = (int16_t)(v.x & 0xffff), dy = (int16_t)(v.x >> 16);
                const uint32_t cost = (uint32_t)(iabs_(dx - genx) + iabs_(dy - geny) + 4) * (v.y & 0x3ffffu);
                keys2[i] = ((unsigned long long)cost << 10) | (unsigned long long)i;
            }
            __syncthreads();
            unsigned long long T2 = ~0ull;
            if (n2 > FH_S3_MAX) { int lt; T2 = block_kth_smallest<unsigned long long, PB_NT>(n2, FH_S3_MAX, 38, [&](int i) { return keys2[i]; }, &sc, &lt); }
            for (int i = tid; i < n2; i += PB_NT) {
                const unsigned long long key = keys2[i];
                if (key <= T2 && (key >> 10) < (unsigned long long)FH_COST_EMPTY) {
                    const uint2 v = __ldg(&pool[i]);
                    const int dx = (int16_t)(v.x & 0xffff), dy = (int16_t)(v.x >> 16), sad = (int)(v.y >> 18);
                    const unsigned long long fk = ((unsigned long long)(sad + mv_cost(dx << 2, dy << 2, mvpx, mvpy)) << 44) | (1ull << 42) | key;
                    atomicMin(&best, fk);
                }
            }
            // stage 3 (:508-520): phase-A list, already in list order
            if (tid < (int)pa.n3) {
                const S3Entry e = S.s3[(size_t)part * FH_S3_MAX + tid];
                const unsigned long long fk = ((unsigned long long)((int)e.sad + mv_cost(e.mvx, e.mvy, mvpx, mvpy)) << 44) | (2ull << 42) | (unsigned long long)tid;
                atomicMin(&best, fk);
            }
        }
        __syncthreads();
        // decode the winner from its key (:523-527); no candidate at all leaves bx = by = 0 (:452)
        const unsigned long long b = best;
        int bx = 0, by = 0, bs = 0;
        if (b != ~0ull) {
            const int stage = (int)((b >> 42) & 3), total = (int)(b >> 44);
            if (stage == 0) {
                const int i = (int)(b & 2047), f = i & 15, pos = i >> 4;
                bx = ((genx + pos / w1 - g1) << 2) | (f & 3); by = ((geny + pos % w1 - g1) << 2) | (f >> 2);
            } else if (stage == 1) {
                const uint2 v = __ldg(&S.s2pool[pa.s2_off + (uint32_t)(b & 1023)]);
                bx = ((int)(int16_t)(v.x & 0xffff)) << 2; by = ((int)(int16_t)(v.x >> 16)) << 2;
            } else {
                const S3Entry e = S.s3[(size_t)part * FH_S3_MAX + (int)(b & 63)];
                bx = e.mvx; by = e.mvy;
            }
            bs = total - mv_cost(bx, by, mvpx, mvpy);
        } else {
            // the reference then reports SAD of MV (0,0) nowhere; sad[] is defined as the SAD of the chosen MV
            const uint8_t *pl = S.planes;
            for (int r = 0; r < 8; r++) bs += sad_row8(rows[r], pl, W, H, xP, yP + r);
        }
        mv[pi][0] = bx; mv[pi][1] = by; sadq[pi] = bs;
        curq[pi][0] = bx; curq[pi][1] = by;
        __syncthreads();
    }

    // ---- merge (:529-551) and final mvd with the merged type's predictors (:552-564) ---------------------------
    if (tid == 0) {
        const bool eq01 = mv[0][0] == mv[1][0] && mv[0][1] == mv[1][1], eq23 = mv[2][0] == mv[3][0] && mv[2][1] == mv[3][1];
        const bool eq02 = mv[0][0] == mv[2][0] && mv[0][1] == mv[2][1], eq13 = mv[1][0] == mv[3][0] && mv[1][1] == mv[3][1];
        int type = FH264_P_8x8ref0, nparts = 4, cnt = 4;
        if (eq01 && eq23 && eq02) { type = FH264_P_L0_16x16; nparts = 1; cnt = 1; }
        else if (eq01 && eq23) { type = FH264_P_L0_L0_16x8; nparts = 2; cnt = 2; }
        else if (eq02 && eq13) { type = FH264_P_L0_L0_8x16; nparts = 2; cnt = 3; }
        int fin[4][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };
        for (int i = 0; i < 4; i++) { mo.mvd[i][0] = mo.mvd[i][1] = 0; }
        for (int i = 0; i < nparts; i++) {
            int ppx = 0, ppy = 0, pw = 16, dir = 0, qsel = i, ox, oy;
            if (type == FH264_P_L0_L0_16x8) { ppy = i * 8; dir = i == 0 ? 1 : 2; qsel = i * 2; }
            else if (type == FH264_P_L0_L0_8x16) { ppx = i * 8; pw = 8; dir = i == 0 ? 2 : 3; }
            else if (type == FH264_P_8x8ref0) { ppx = (i & 1) * 8; ppy = (i >> 1) * 8; pw = 8; }
            predict_mv_(nc, ppx, ppy, pw, dir, fin, ox, oy);
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
        atomicAdd(&S.status[ST_COUNTS + cnt], 1u);
        __threadfence();
        st_release_u32(&S.done[mb], epoch);
    }
}
