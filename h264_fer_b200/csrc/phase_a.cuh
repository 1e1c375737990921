// Phase A — per current picture, fully parallel over 8x8 partitions, independent of the MV predictors.
//   k_stage3  the complete stage-3 list (moestimation.cpp:508-520): feature costs over the +-window/2 integer
//             window and the +-window/16 quarter-pel window around (0,0), the 33 best by (cost, arrival), their SADs.
//   k_stage2  the stage-2 candidate SET (moestimation.cpp:470-497): positions whose 8x8 sum is within +-j_stop of
//             the block's, gated by Manhattan distance and the two half-sums, in the reference's arrival order,
//             with feature distance and SAD. Only the multiplier (|dx-genx|+|dy-geny|+4) is left to phase B.
// One WARP per partition (no block barriers).
#pragma once
#include "common.cuh"
#include "warp_select.cuh"
#include "qfeat.cuh"

#define COST_INVALID 0xffffffffu
// tuning constants (each one the winner of an A/B build on the B200, profiles/tools/ab_variants.sh)
#ifndef FH_S3_SADR
#define FH_S3_SADR 9      // stage-3 SAD: member rows in flight per lane (all 33 members in one round)
#endif
#ifndef FH_S2_UNR
#define FH_S2_UNR 4        // index entries in flight per lane in the stage-2 visit loop
#endif
#ifndef FH_S2_SADR
#define FH_S2_SADR 8       // stage-2 SAD: candidate rows in flight per lane
#endif
#ifndef FH_S3_UNR
#define FH_S3_UNR 4        // stage-3 first call: feature records in flight per lane
#endif
#ifndef FH_S2_MINB
#define FH_S2_MINB 12     // fast stage-2 launch: 80 registers, 12 CTAs per SM
#endif

__device__ __forceinline__ void part_origin(const Geo &g, int part, int &xP, int &yP)
{
    const int mb = part >> 2, pi = part & 3;
    xP = (mb % g.Wmb) * 16 + (pi & 1) * 8;
    yP = (mb / g.Wmb) * 16 + (pi >> 1) * 8;
}

// Loads the 8x8 source block of a partition (8 rows of two words) into every lane's registers.
__device__ __forceinline__ void load_cur8x8(const uint8_t *__restrict__ cur, const Geo &g, int xP, int yP, uint2 rows[8])
{
#pragma unroll
    for (int r = 0; r < 8; r++) rows[r] = __ldg((const uint2 *)(cur + (size_t)(yP + r) * g.W + xP));
}
__device__ __forceinline__ uint2 pick_row(const uint2 rows[8], int r)
{
    uint2 cr = rows[0];
#pragma unroll
    for (int q = 1; q < 8; q++) if (r == q) cr = rows[q];
    return cr;
}

// arrival index of the stage-3 list -> displacement and fraction (i3 / i1 = 2^32 / w + 1: exact quotients by one IMAD.HI)
__device__ __forceinline__ void s3_decode(int i, int n3a, int w3, int g3, uint32_t i3, int w1, int g1, uint32_t i1, int &dx, int &dy, int &f)
{
    const bool a = i < n3a;
    const int t = a ? i : (i - n3a) >> 4, w = a ? w3 : w1, gg = a ? g3 : g1;
    const int c = udiv_by(t, a ? i3 : i1);
    dx = c - gg; dy = t - c * w - gg; f = a ? 0 : (i - n3a) & 15;
}

struct S3Warp { WarpSelScratch ws; uint16_t members[FH_S3_MAX + 1]; uint16_t msad[FH_S3_MAX + 1]; };

// Selection of the K = min(k, nvalid) smallest (cost, arrival index) pairs of cost[0..n-1] (0xffffffff = no candidate) when
// every lane already holds the two smallest COSTS (m1 <= m2) of the elements it produced. The bound on the K-th cost
// is the one of warp_select_smallest; survivors are the elements with cost <= bound (ties included), ranked exactly by
// the 64-bit (cost, index) key. Same contract as warp_select_smallest.
__device__ __forceinline__ int warp_select_costs(const uint32_t *cost, int n, int k, int nvalid, uint32_t m1, uint32_t m2, WarpSelScratch *ws, uint16_t *members)
{
    const int lane = threadIdx.x & 31;
    const int K = min(k, nvalid);
    if (K == 0) return 0;
    const int L1 = __popc(__ballot_sync(0xffffffffu, m1 != COST_INVALID));
    const int L2 = __popc(__ballot_sync(0xffffffffu, m2 != COST_INVALID));
    uint32_t thr = COST_INVALID - 1;
    if (L1 >= K) thr = __reduce_max_sync(0xffffffffu, m1 != COST_INVALID ? m1 : 0u);
    else if (L1 + 1 >= K && L2 >= 1) thr = max(__reduce_max_sync(0xffffffffu, m1 != COST_INVALID ? m1 : 0u), __reduce_min_sync(0xffffffffu, m2));
    else if (2 * L2 >= K) thr = __reduce_max_sync(0xffffffffu, m2 != COST_INVALID ? m2 : 0u);
    if (L1 + L2 >= K) {
        // tighter: the K-th smallest of the (up to 64) per-lane minima themselves — K distinct candidates lie at or below it.
        // Bisection on the value with two ballots per step (about 20 steps); it typically halves the survivors to rank.
        uint32_t lo = __reduce_min_sync(0xffffffffu, m1), hi = thr;
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            const int c = __popc(__ballot_sync(0xffffffffu, m1 <= mid)) + __popc(__ballot_sync(0xffffffffu, m2 <= mid));
            if (c >= K) hi = mid; else lo = mid + 1;
        }
        thr = lo;
    }
    int ns = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const uint32_t c = i < n ? cost[i] : COST_INVALID;
        const bool sv = c <= thr;                              // COST_INVALID > thr always
        const unsigned b = __ballot_sync(0xffffffffu, sv);
        if (sv) {
            const int pos = ns + __popc(b & ((1u << lane) - 1u));
            if (pos < WSEL_CAP) { ws->skey[pos] = ((u64)c << 16) | (u64)i; ws->sidx[pos] = (uint16_t)i; }
        }
        ns += __popc(b);
    }
    __syncwarp();
    if (ns > WSEL_CAP) {
        // the bound was loose (the good candidates sat in few lanes): bisect the exact K-th smallest cost, then compact again
        uint32_t lo = 0, hi = thr;
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            int c = 0;
            for (int i = lane; i < n; i += 32) c += cost[i] <= mid;
            if (__reduce_add_sync(0xffffffffu, c) >= K) hi = mid; else lo = mid + 1;
        }
        thr = lo;
        ns = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const uint32_t c = i < n ? cost[i] : COST_INVALID;
            const bool sv = c <= thr;
            const unsigned b = __ballot_sync(0xffffffffu, sv);
            if (sv) {
                const int pos = ns + __popc(b & ((1u << lane) - 1u));
                if (pos < WSEL_CAP) { ws->skey[pos] = ((u64)c << 16) | (u64)i; ws->sidx[pos] = (uint16_t)i; }
            }
            ns += __popc(b);
        }
        __syncwarp();
    }
    if (ns <= WSEL_CAP) {
        for (int s = lane; s < ns; s += 64) {
            const u64 ka = ws->skey[s];
            const bool hb = s + 32 < ns;
            const u64 kb = hb ? ws->skey[s + 32] : 0ull;
            int ra = 0, rb = 0;
            for (int j = 0; j < ns; j++) { const u64 kj = ws->skey[j]; ra += kj < ka; rb += kj < kb; }
            if (ra < K) members[ra] = ws->sidx[s];
            if (hb && rb < K) members[rb] = ws->sidx[s + 32];
        }
    } else {
        // degenerate (flat content: hundreds of candidates tie at the K-th cost): rank against every candidate
        for (int i = lane; i < n; i += 32) {
            const uint32_t c = cost[i];
            if (c > thr) continue;
            const u64 key = ((u64)c << 16) | (u64)i;
            int rank = 0;
            for (int j = 0; j < n && rank < K; j++) rank += (((u64)cost[j] << 16) | (u64)j) < key;
            if (rank < K) members[rank] = (uint16_t)i;
        }
    }
    __syncwarp();
    return K;
}

__global__ void __launch_bounds__(128, 6) k_stage3(const SeqDev *__restrict__ seqs, int seq0, Geo g, fh264_params prm, int npad)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    S3Warp *sw = (S3Warp *)smem_raw + warp;
    uint32_t *cost = (uint32_t *)(smem_raw + 4 * sizeof(S3Warp)) + (size_t)warp * npad;
    const SeqDev &S = seqs[seq0 + blockIdx.y];
    const int part = g.band_mb0 * 4 + blockIdx.x * 4 + warp;
    int xP, yP;
    part_origin(g, part, xP, yP);
    uint2 rows[8];
    load_cur8x8(S.cur[0], g, xP, yP, rows);
    int s[5];
    block_sums(rows, s);
    const FeatQ fq = feat_query(s);
    const int W = g.W, H = g.H;
    const int g3 = prm.window / 2, g1 = prm.window / 16;
    const int w3 = 2 * g3 + 1, w1 = 2 * g1 + 1;
    const int n3a = w3 * w3, n3b = w1 * w1 * 16, N = n3a + n3b;
    const uint32_t i3 = udiv_magic((uint32_t)w3), i1 = udiv_magic((uint32_t)w1);
    const uint4 *__restrict__ K0p = S.kar;
    uint32_t m1 = COST_INVALID, m2 = COST_INVALID;     // this lane's two smallest costs so far (selection bound)
    // first call: MEstimation(g = window/2, frac 0); arrival index = (dx + g3) * w3 + (dy + g3).
    // lane = column (coalesced 16-byte loads), 4 rows in flight; rows [rlo, rhi) lie inside the picture.
    const int rlo = max(0, g3 - yP), rhi = min(w3, H - yP + g3);
    const int ncf = w3 & ~31;
    for (int cb = 0; cb < ncf; cb += 32) {
        const int c = cb + lane, dx = c - g3, rx = xP + dx;
        const bool xok = rx >= 0 && rx < W;
        const int adx = iabs_(dx) + 4;
        uint32_t *cc = cost + c * w3;
        for (int r = 0; r < rlo; r++) cc[r] = COST_INVALID;
        for (int r = rhi; r < w3; r++) cc[r] = COST_INVALID;
        const uint4 *kp = K0p + (size_t)(yP - g3 + rlo) * W + rx;
        auto ld4 = [&](const uint4 *q, int r0, uint4 (&v)[FH_S3_UNR]) {
#pragma unroll
            for (int u = 0; u < FH_S3_UNR; u++) { v[u] = make_uint4(0, 0, 0, 0); if (xok && r0 + u < rhi) v[u] = __ldg(q + (size_t)u * W); }
        };
        auto ev4 = [&](const uint4 (&v)[FH_S3_UNR], int r0) {
#pragma unroll
            for (int u = 0; u < FH_S3_UNR; u++) {
                const int r = r0 + u;
                if (r < rhi) {
                    const uint32_t cst = xok ? (uint32_t)((adx + iabs_(r - g3)) * feat_of(fq, v[u])) : COST_INVALID;
                    cc[r] = cst;
                    m2 = min(m2, max(m1, cst)); m1 = min(m1, cst);
                }
            }
        };
        uint4 va[FH_S3_UNR];
        for (int r0 = rlo; r0 < rhi; r0 += FH_S3_UNR) { ld4(kp, r0, va); ev4(va, r0); kp += (size_t)FH_S3_UNR * W; }
    }
    // leftover columns: lane = row
    for (int c = ncf; c < w3; c++) {
        const int dx = c - g3, rx = xP + dx;
        const bool xok = rx >= 0 && rx < W;
        for (int r = lane; r < w3; r += 32) {
            uint32_t cst = COST_INVALID;
            if (xok && r >= rlo && r < rhi) cst = (uint32_t)((iabs_(dx) + iabs_(r - g3) + 4) * feat_of(fq, __ldg(K0p + (size_t)(yP + r - g3) * W + rx)));
            cost[c * w3 + r] = cst;
            m2 = min(m2, max(m1, cst)); m1 = min(m1, cst);
        }
    }
    // second call: MEstimation(g = window/16, all 16 fractions); arrival = ((dx+g1)*w1 + (dy+g1))*16 + frac.
    // The features of the quarter-pel planes are computed from the planes themselves (qfeat.cuh), a few planes per batch,
    // in the selection scratch (free until the selection below).
    {
        const int npos = w1 * w1, R = 8 + w1 - 1, ps = R * w1, pb = w1 <= 5 ? 4 : 1;
        uint32_t *X = (uint32_t *)sw->ws.skey;
        uint16_t *RC = sw->ws.sidx;
        const uint32_t iR = udiv_magic((uint32_t)R), iN = udiv_magic((uint32_t)npos);
        for (int f0 = 0; f0 < 16; f0 += pb) {
            // step A: two rows per lane in flight
            for (int sg0 = 0; sg0 < pb * R; sg0 += 64) {
                uint4 wa = make_uint4(0, 0, 0, 0), wb = wa;
                const int sa = sg0 + lane, sb = sg0 + 32 + lane;
                const int fa = udiv_by(sa, iR), ra = sa - fa * R, fb = udiv_by(sb, iR), rb = sb - fb * R;
                if (sa < pb * R) wa = qf_load16(S.planes + (size_t)(f0 + fa) * g.WH, W, H, xP - g1, yP - g1 + ra);
                if (sb < pb * R) wb = qf_load16(S.planes + (size_t)(f0 + fb) * g.WH, W, H, xP - g1, yP - g1 + rb);
                if (sa < pb * R) qf_row_sums(wa, w1, X + fa * ps + ra * w1, RC + fa * ps + ra * w1);
                if (sb < pb * R) qf_row_sums(wb, w1, X + fb * ps + rb * w1, RC + fb * ps + rb * w1);
            }
            __syncwarp();
            // step B: one (plane, position) per lane and round
            // (the lane -> element map is rotated from batch to batch: every lane should meet many different window
            // positions, or the per-lane minima that bound the selection below stay loose)
            const int rot = (f0 * 13) % (pb * npos);
            for (int o0 = lane; o0 < pb * npos; o0 += 32) {
                const int o = o0 + rot < pb * npos ? o0 + rot : o0 + rot - pb * npos;
                const int fl = udiv_by(o, iN), pos = o - fl * npos, cx = udiv_by(pos, i1), cy = pos - cx * w1;
                const int rx = xP + cx - g1, ry = yP + cy - g1;
                uint32_t cst = COST_INVALID;
                if (rx >= 0 && rx < W && ry >= 0 && ry < H)
                    cst = (uint32_t)((iabs_(cx - g1) + iabs_(cy - g1) + 4) * feat_of(fq, qf_record(X, RC, fl, ps, w1, cx, cy)));
                cost[n3a + pos * 16 + f0 + fl] = cst;
                m2 = min(m2, max(m1, cst)); m1 = min(m1, cst);
            }
            __syncwarp();
        }
    }
    __syncwarp();
    // candidates = positions whose block origin lies inside the picture (:263-266)
    const int nva = max(0, min(W - 1, xP + g3) - max(0, xP - g3) + 1) * max(0, rhi - rlo);
    const int nvb = max(0, min(W - 1, xP + g1) - max(0, xP - g1) + 1) * max(0, min(H - 1, yP + g1) - max(0, yP - g1) + 1) * 16;
    // the 33 smallest by (cost, arrival index), in list order
    const int nm = warp_select_costs(cost, N, FH_S3_MAX, nva + nvb, m1, m2, &sw->ws, sw->members);
    // SADs of the members (satdLuma8x8MVs): 8 lanes per member, one row each; 3 rounds of loads in flight
    const int r = lane & 7;
    const uint2 cr = pick_row(rows, r);
    for (int base = 0; base < nm; base += 4 * FH_S3_SADR) {
        uint2 rr[FH_S3_SADR];                           // 33 members x 8 rows = 264 row loads
#pragma unroll
        for (int u = 0; u < FH_S3_SADR; u++) {
            const int m = base + u * 4 + (lane >> 3);
            rr[u] = make_uint2(0, 0);
            if (m < nm) {
                int dx, dy, f;
                s3_decode((int)sw->members[m], n3a, w3, g3, i3, w1, g1, i1, dx, dy, f);
                rr[u] = load_row8(S.planes + (size_t)f * g.WH, W, H, xP + dx, yP + dy + r);
            }
        }
#pragma unroll
        for (int u = 0; u < FH_S3_SADR; u++) {
            const int m = base + u * 4 + (lane >> 3);
            int sad = m < nm ? sad8(cr, rr[u]) : 0;
            sad += __shfl_xor_sync(0xffffffffu, sad, 1);
            sad += __shfl_xor_sync(0xffffffffu, sad, 2);
            sad += __shfl_xor_sync(0xffffffffu, sad, 4);
            if (m < nm && r == 0) sw->msad[m] = (uint16_t)sad;
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
        }
    }
    if (lane == 0) {
        PartA *pa = &S.parta[part];
#pragma unroll
        for (int k = 0; k < 5; k++) pa->suma[k] = (uint16_t)s[k];
        pa->n3 = (uint16_t)nm;
    }
}

#define S2_BINS 384           // (j, side) bins: 2*j + side, j <= 180
#define S2_CHUNK_CAP 512      // a round of 32 short ranges (<= 64 entries each) adds at most 256 chunks
#define S2_CAP_FAST 512       // gated survivors per partition held by the main launch
#define S2_CAP_BIG 4096       // ... and by the fallback launch (beyond: FH264_E_CAPACITY)
#define S2_REDO_MAX 1024      // partitions per picture the fallback launch can take over

template <int CAP>
struct S2Warp {
    uint32_t akey[CAP];              // arrival key: j<<21 | side<<20 | (dx+279)<<10 | (dy+279)
    uint32_t aval[CAP];              // index entry number, later feature distance (18 bits) | SAD << 18
    uint16_t order[CAP];             // survivor index by output slot
    uint32_t bins[S2_BINS];          // counts, then exclusive starts, then scatter cursors
    uint32_t chunk[S2_CHUNK_CAP];    // chunks of <= 8 consecutive index entries still to be visited
    int n_surv;
};

// CAP = survivor capacity per warp, NW = warps (= partitions) per CTA. The main launch (CAP = 512) keeps shared memory
// small for occupancy; partitions with more gated survivors are marked and redone by a second launch with CAP = 4096.
template <int CAP, int NW, bool REDO>
__global__ void __launch_bounds__(32 * NW, REDO ? 1 : FH_S2_MINB) k_stage2(const SeqDev *__restrict__ seqs, int seq0, Geo g, fh264_params prm)
{
    __shared__ S2Warp<CAP> sm[NW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    S2Warp<CAP> *w = &sm[warp];
    const SeqDev &S = seqs[seq0 + blockIdx.y];
    // REDO launch: CTA b takes the b-th partition the main launch listed as overflowing (usually none: exit at once)
    if (REDO && blockIdx.x >= min(S.status[ST_S2REDO], (uint32_t)S2_REDO_MAX)) return;
    const int part = REDO ? (int)S.s2redo[blockIdx.x] : g.band_mb0 * 4 + blockIdx.x * NW + warp;
    int xP, yP;
    part_origin(g, part, xP, yP);
    uint2 rows[8];
    load_cur8x8(S.cur[0], g, xP, yP, rows);
    int s[5];
    block_sums(rows, s);
    for (int i = lane; i < S2_BINS; i += 32) w->bins[i] = 0;
    if (lane == 0) w->n_surv = 0;
    __syncwarp();
    // tiles intersecting the bounding box of the diamond |dx|+|dy| < 280 (moestimation.cpp:481)
    const int tx0 = max(0, xP - 279) >> FH_TILE_SHIFT, tx1 = min(g.W - 1, xP + 279) >> FH_TILE_SHIFT;
    const int ty0 = max(0, yP - 279) >> FH_TILE_SHIFT, ty1 = min(g.H - 1, yP + 279) >> FH_TILE_SHIFT;
    const int ntx = tx1 - tx0 + 1, nty = ty1 - ty0 + 1;
    // index cells = (K1 >> 6, K2 >> 6): the two half-sum gates select the cell rectangle (at most 5 rows of K1), the K0 gate is
    // checked per entry. (K0 is strongly correlated with K1 + K2, so it prunes little as an index key: on the bench content this
    // choice visits 37 % fewer entries than (K0 >> 7, K1 >> 6) cells of the same table size.)
    const int qlo = max(0, s[1] - 99) >> 6, qhi = min(8191, s[1] + 99) >> 6, nq = qhi - qlo + 1;
    const int k1lo = max(0, s[2] - 99) >> 6, k1hi = min(8191, s[2] + 99) >> 6;       // column range (K2 cells)
    const int inv_nq = 65536 / nq + 1;
    const int ntiles = ntx * nty, inv_ntx = 65536 / ntx + 1;
    const uint4 *__restrict__ tent = (const uint4 *)S.tent;
    // gate of one index entry (:481). Gated entries are counted per (j, side); they are KEPT only while j <= jb, a running
    // upper bound of j_stop (counts only grow, so the first j whose running total exceeds 128 can only move down). The
    // feature distance is computed afterwards in a dense pass over the kept entries.
    int jb = 180;
    // the four gates in 16-bit lanes: (|dx|, |dy|) and (j, |dK1|) by max(a - b, b - a), the Manhattan sum by a dot product
    const uint32_t Pxy = (uint32_t)xP | ((uint32_t)yP << 16), S01 = (uint32_t)s[0] | ((uint32_t)s[1] << 16), LIM01 = 180u | (99u << 16);
    const int s2m = s[2] - 99;
    auto visit = [&](const uint4 v, uint32_t eidx) {
        const uint32_t ad = __vmaxs2(__vsub2(v.x, Pxy), __vsub2(Pxy, v.x)), ae = __vmaxs2(__vsub2(v.y, S01), __vsub2(S01, v.y));
        const bool ok = __vmaxu2(ae, LIM01) == LIM01 && __dp2a_lo(ad, 0x0101u, 0u) < 280u && (uint32_t)((int)(v.z & 0xffffu) - s2m) < 199u;
        if (ok) {
            const int j = (int)(ae & 0xffffu), side = (int)(v.y & 0xffffu) > s[0];
            atomicAdd(&w->bins[2 * j + side], 1u);
            if (j <= jb) {
                const int pos = atomicAdd(&w->n_surv, 1);
                if (pos < CAP) {
                    const int dx = (int)(v.x & 0xffffu) - xP, dy = (int)(v.x >> 16) - yP;
                    w->akey[pos] = ((uint32_t)j << 21) | ((uint32_t)side << 20) | ((uint32_t)(dx + 279) << 10) | (uint32_t)(dy + 279);
                    w->aval[pos] = eidx;
                }
            }
        }
    };
    // j_stop bound from the counts so far: first j whose running gated count (bucket s0 twice) exceeds 128 (:496)
    auto bound_from_bins = [&]() -> int {
        uint32_t local = 0, c6[6];
#pragma unroll
        for (int i = 0; i < 6; i++) { c6[i] = w->bins[lane * 12 + 2 * i] + w->bins[lane * 12 + 2 * i + 1]; if (lane == 0 && i == 0) c6[i] *= 2; local += c6[i]; }
        uint32_t incl = local;
        for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        uint32_t run = incl - local;
        int first = 180;
#pragma unroll
        for (int i = 0; i < 6; i++) { run += c6[i]; if (run > 128 && first == 180) first = min(180, lane * 6 + i); }
        return __reduce_min_sync(0xffffffffu, first);
    };
    // keep the kept-entry buffer from overflowing: tighten jb and drop entries beyond it (called at warp-uniform points)
    auto tighten = [&]() {
        __syncwarp();
        if (w->n_surv <= CAP - 256) return;
        const int nb = bound_from_bins();
        const int n = min(w->n_surv, CAP);
        int outn = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const uint32_t k = i < n ? w->akey[i] : 0xffffffffu, e = i < n ? w->aval[i] : 0u;
            const bool keep = i < n && (int)(k >> 21) <= nb;
            const unsigned b = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (keep) { const int p = outn + __popc(b & ((1u << lane) - 1u)); w->akey[p] = k; w->aval[p] = e; }
            outn += __popc(b);
        }
        __syncwarp();
        // entries lost to an overflow since the last call cannot be recovered: poison the count so the partition is redone
        if (lane == 0) w->n_surv = (w->n_surv > CAP) ? 0x40000000 : outn;
        jb = nb;
        __syncwarp();
    };
    int nchunk = 0;
    const int nitems = ntiles * nq;
    for (int it0 = 0; it0 < nitems; it0 += 32) {
        const int it = it0 + lane, tt = fdiv_(it, inv_nq), q = it - tt * nq;
        int len = 0; uint32_t gbase = 0;
        if (it < nitems) {
            const int tyy = fdiv_(tt, inv_ntx), tx = tx0 + tt - tyy * ntx, ty = ty0 + tyy;
            const int rx0 = tx << FH_TILE_SHIFT, ry0 = ty << FH_TILE_SHIFT;
            const int ddx = max(0, max(rx0 - xP, xP - (rx0 + FH_TILE - 1))), ddy = max(0, max(ry0 - yP, yP - (ry0 + FH_TILE - 1)));
            if (ddx + ddy < 280) {
                const int tile = ty * g.tilesx + tx;
                const uint16_t *ts = S.tstart + (size_t)tile * FH_TSTART_PITCH;
                const int e0 = __ldg(&ts[(qlo + q) * 128 + k1lo]), e1 = __ldg(&ts[(qlo + q) * 128 + k1hi + 1]);
                len = e1 - e0; gbase = (uint32_t)tile * (FH_TILE * FH_TILE) + (uint32_t)e0;
            }
        }
        // long ranges (flat content) are walked by the whole warp right away
        unsigned longm = __ballot_sync(0xffffffffu, len > 64);
        while (longm) {
            const int src = __ffs(longm) - 1;
            longm &= longm - 1;
            const uint32_t gb = __shfl_sync(0xffffffffu, gbase, src);
            const int ln = __shfl_sync(0xffffffffu, len, src);
            for (int e0 = 0; e0 < ln; e0 += 32) { if (e0 + lane < ln) visit(__ldg(tent + gb + e0 + lane), gb + e0 + lane); if ((e0 & 255) == 224) tighten(); }
            tighten();
            if (lane == src) len = 0;
        }
        // short ranges: chunks of <= 8 consecutive entries (one 128-byte line): first entry | (count - 1) << 28
        const int nc = (len + 7) >> 3;
        int incl = nc;
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        const int pos = nchunk + incl - nc;
        for (int c = 0; c < nc; c++) w->chunk[pos + c] = (gbase + 8u * c) | ((uint32_t)(min(8, len - 8 * c) - 1) << 28);
        nchunk += __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        if (nchunk > S2_CHUNK_CAP - 256 || it0 + 32 >= nitems) {
            // 32 chunks per round: 8 lanes share a chunk (one 128-byte line per 8 lanes: coalesced, 4 lines per load
            // instruction instead of 32), lane takes entry (lane & 7) of chunks c0 + (lane >> 3) + 4u; 8 loads in flight
            for (int c0 = 0; c0 < nchunk; c0 += 4 * FH_S2_UNR) {
                uint4 v[FH_S2_UNR];
                uint32_t eid[FH_S2_UNR];
#pragma unroll
                for (int u = 0; u < FH_S2_UNR; u++) {
                    const int cidx = c0 + 4 * u + (lane >> 3);
                    const uint32_t cw = cidx < nchunk ? w->chunk[cidx] : 0u;
                    const int cnt = cidx < nchunk ? (int)(cw >> 28) + 1 : 0;
                    eid[u] = (cw & 0x0fffffffu) + (uint32_t)(lane & 7);
                    v[u] = (lane & 7) < cnt ? __ldg(tent + eid[u]) : make_uint4(0, 0xffffu, 0x7fffu, 0);
                }
#pragma unroll
                for (int u = 0; u < FH_S2_UNR; u++) visit(v[u], eid[u]);
                tighten();                                  // at most 256 entries were added since the last check
            }
            nchunk = 0;
            __syncwarp();
        }
    }
    __syncwarp();
    const int ns = w->n_surv;
    if (ns > CAP) {
        // too many gated survivors for this launch's buffers: the fallback launch takes the partition over; when that cannot
        // hold it either (or its list is full) phase B enumerates the set itself. j_stop is final here (the counts saw every entry).
        const int jsb = bound_from_bins();
        // candidates up to j_stop (bucket s0 twice): more than a pool slice holds -> no point in the fallback launch either
        uint32_t upto = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            uint32_t c = w->bins[lane * 12 + 2 * i] + w->bins[lane * 12 + 2 * i + 1];
            if (lane == 0 && i == 0) c *= 2;
            if (lane * 6 + i <= jsb) upto += c;
        }
        const bool fits = __reduce_add_sync(0xffffffffu, upto) <= 1023u;
        if (lane == 0) {
            PartA *pa = &S.parta[part];
            pa->s2_off = 0;
            bool listed = false;
            if (!REDO && fits) { const uint32_t k = atomicAdd(&S.status[ST_S2REDO], 1u); if (k < S2_REDO_MAX) { S.s2redo[k] = (uint32_t)part; listed = true; } }
            pa->n2 = listed ? 0u : (S2_SLOW | (uint32_t)jsb);
        }
        return;
    }
    // dense pass over the kept entries: feature distance (:267-276)
    for (int i = lane; i < ns; i += 32) {
        const uint4 v = __ldg(tent + w->aval[i]);
        w->aval[i] = (uint32_t)feat_dist(s, (int)(v.y & 0xffff), (int)(v.y >> 16), (int)(v.z & 0xffff), (int)(v.z >> 16), (int)(v.w & 0xffff));
    }
    __syncwarp();
    // j_stop: first j at which the running gated count exceeds 128 (:496), else 180. Bucket s0 is visited by both
    // sides (:476,486), so its entries count twice. Lane l owns j = 6l .. 6l+5 (bins 12l .. 12l+11).
    uint32_t cb[12], cj[6], local = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) cb[i] = w->bins[lane * 12 + i];
#pragma unroll
    for (int i = 0; i < 6; i++) { cj[i] = cb[2 * i] + cb[2 * i + 1]; if (lane == 0 && i == 0) cj[i] *= 2; local += cj[i]; }
    uint32_t incl = local;
    for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    const uint32_t excl = incl - local;
    uint32_t run = excl;
    int first = 1 << 20;
#pragma unroll
    for (int i = 0; i < 6; i++) { run += cj[i]; if (run > 128 && first == (1 << 20)) first = lane * 6 + i; }
    first = __reduce_min_sync(0xffffffffu, first);
    const int js = min(first, 180);
    uint32_t upto = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) if (lane * 6 + i <= js) upto += cj[i];
    int n2 = (int)__reduce_add_sync(0xffffffffu, upto);
    const int cnt0 = (int)__shfl_sync(0xffffffffu, cb[0], 0);
    // every partition owns a fixed 1024-entry slice of the candidate pool (no allocation, and phase B can prefetch it
    // without first reading the partition header)
    uint32_t off = (uint32_t)part * 1024u;
    if (lane == 0) {
        PartA *pa = &S.parta[part];
        pa->s2_off = off;
        if (n2 > 1023) { pa->n2 = S2_SLOW | (uint32_t)js; n2 = 0; }      // more candidates than a pool slice: phase B's own enumeration
        else if (!REDO && n2 > CAP) {
            // the kept entries fit, but with bucket s0 listed twice the output slots would run past order[CAP] (half-flat content:
            // hundreds of positions share the block's exact sum): the fallback launch, whose buffers hold any pool slice, redoes it
            const uint32_t k = atomicAdd(&S.status[ST_S2REDO], 1u);
            const bool listed = k < S2_REDO_MAX;
            if (listed) S.s2redo[k] = (uint32_t)part;
            pa->s2_off = 0;
            pa->n2 = listed ? 0u : (S2_SLOW | (uint32_t)js);
            n2 = 0;
        }
        else pa->n2 = (uint32_t)n2;
    }
    n2 = __shfl_sync(0xffffffffu, n2, 0);
    if (n2 == 0) return;
    // Arrival order (:474-495): j ascending; minus side before plus side; x then y inside a bucket; bucket s0 twice.
    // Output slot of bin (j, side): start = (entries of smaller j, bucket s0 counted twice) + (side ? count of side 0 : 0);
    // the second visit of bucket s0 is a copy at cnt0. Starts replace the counts in w->bins.
    {
        uint32_t st = excl;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            w->bins[lane * 12 + 2 * i] = st;
            w->bins[lane * 12 + 2 * i + 1] = st + cb[2 * i];
            st += cj[i];
        }
    }
    __syncwarp();
    for (int i = lane; i < ns; i += 32) {
        const uint32_t k = w->akey[i];
        const int bin = (int)(k >> 20);
        if ((bin >> 1) > js) continue;
        w->order[atomicAdd(&w->bins[bin], 1u)] = (uint16_t)i;    // arbitrary order inside a bin, fixed below
    }
    __syncwarp();
    {
        // order inside each (j, side) bin by (x, y): insertion sort by the owning lane (bins hold 0-2 entries on textured content)
        uint32_t st = excl;
#pragma unroll
        for (int i = 0; i < 6; i++) {
#pragma unroll
            for (int sd = 0; sd < 2; sd++) {
                const int j = lane * 6 + i;
                const int b0 = (int)(sd ? st + cb[2 * i] : st), cnt = (int)cb[2 * i + sd];
                if (j <= js && cnt > 1) {
                    for (int a = 1; a < cnt; a++) {
                        const uint16_t ia = w->order[b0 + a];
                        const uint32_t ka = w->akey[ia];
                        int p = a - 1;
                        while (p >= 0 && w->akey[w->order[b0 + p]] > ka) { w->order[b0 + p + 1] = w->order[b0 + p]; p--; }
                        w->order[b0 + p + 1] = ia;
                    }
                }
            }
            st += cj[i];
        }
    }
    __syncwarp();
    // SAD at integer displacement (fraction 0 => plane 0) for the kept candidates: 8 lanes per candidate, one row
    // each, 8 rounds of loads in flight. Slots of the second visit of bucket s0 are copies.
    const uint8_t *pl = S.planes;
    const int r = lane & 7;
    const uint2 cr = pick_row(rows, r);
    const int nuniq = n2 - cnt0;          // distinct candidates: slots [0, cnt0) and [2*cnt0, n2)
    for (int base = 0; base < nuniq; base += 4 * FH_S2_SADR) {
        uint2 rr[FH_S2_SADR];
        int idx[FH_S2_SADR];
#pragma unroll
        for (int u = 0; u < FH_S2_SADR; u++) {
            const int m = base + u * 4 + (lane >> 3);
            rr[u] = make_uint2(0, 0); idx[u] = -1;
            if (m < nuniq) {
                const int slot = m < cnt0 ? m : m + cnt0;
                idx[u] = w->order[slot];
                const uint32_t k = w->akey[idx[u]];
                rr[u] = load_row8(pl, g.W, g.H, xP + (int)((k >> 10) & 1023) - 279, yP + (int)(k & 1023) - 279 + r);
            }
        }
#pragma unroll
        for (int u = 0; u < FH_S2_SADR; u++) {
            int sad = idx[u] >= 0 ? sad8(cr, rr[u]) : 0;
            sad += __shfl_xor_sync(0xffffffffu, sad, 1);
            sad += __shfl_xor_sync(0xffffffffu, sad, 2);
            sad += __shfl_xor_sync(0xffffffffu, sad, 4);
            if (idx[u] >= 0 && r == 0) w->aval[idx[u]] |= (uint32_t)sad << 18;
        }
    }
    __syncwarp();
    uint2 *pool = S.s2pool + off;
    for (int m = lane; m < nuniq; m += 32) {
        const int slot = m < cnt0 ? m : m + cnt0, i = w->order[slot];
        const uint32_t k = w->akey[i];
        const int dx = (int)((k >> 10) & 1023) - 279, dy = (int)(k & 1023) - 279;
        const uint2 v = make_uint2(((uint32_t)dx & 0xffffu) | ((uint32_t)dy << 16), w->aval[i]);
        pool[slot] = v;
        if (m < cnt0) pool[cnt0 + m] = v;
    }
}
