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
#define FH_S3_SADR 3      // stage-3 SAD: member rows in flight per lane
#endif
#ifndef FH_S2_UNR
#define FH_S2_UNR 4        // index entries in flight per lane in the stage-2 visit loop
#endif
#ifndef FH_S2_MINB
#define FH_S2_MINB 16     // 64 registers, 16 CTAs of two warps per SM (A/B on the B200: 2.81 -> 2.60 ms against 12)
#endif

__device__ __forceinline__ void part_origin(const Geo &g, int part, int &xP, int &yP)
{
    const int mb = part >> 2, pi = part & 3;
    int mbx, mby;
    mb_xy(g, mb, mbx, mby);
    xP = mbx * 16 + (pi & 1) * 8;
    yP = mby * 16 + (pi >> 1) * 8;
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

#define S2_BINS 384           // (j, side) bins: 2*j + side, j <= 180
#define S2_CHUNK_CAP 512      // a round of 32 short ranges (<= 64 entries each) adds at most 256 chunks
#define S2_CAP 512            // gated entries up to the running j_stop bound held per partition
#define S2_SLICE 512          // pool entries per partition (16 bytes each); more candidates than that: S2_SLOW

struct S2Warp {
    uint32_t akey[S2_CAP];           // arrival key: j<<21 | side<<20 | (dx+279)<<10 | (dy+279)
    uint32_t aval[S2_CAP];           // index entry number
    uint32_t bins[S2_BINS];          // per (j, side) counts of the kept entries (recounted when the bound is tightened)
    uint32_t chunk[S2_CHUNK_CAP];    // chunks of <= 8 consecutive index entries still to be visited
};

// Stage-2 candidate SET of every partition (moestimation.cpp:470-497), one warp per partition: positions whose 8x8 sum is within
// +-j_stop of the block's, gated by Manhattan distance and the two half-sums. Output: pool entries {dx | dy << 16, feature
// distance, lower bound of the SAD, arrival key} in no particular order (bucket s0 twice, :476,486) — ranking by cost needs the
// multiplier of the (guessed or true) predictor and happens in phase S / phase B, and so do the SADs of the few candidates that
// can matter (the lower bound — the largest of the feature distance's box-pair terms, each a sum of |differences| over
// complementary boxes — rules the others out).
// The walk keeps gated entries only while j <= jb, a running upper bound of j_stop: the first j whose gated count (bucket s0
// twice) exceeds 128 can only move down as more entries are seen. Entries are appended by ballot compaction; jb is recomputed
// from the kept entries whenever the buffer fills. Content where thousands of positions share one sum (flat) overflows the
// buffer: the walk then only counts, records the exact j_stop, and phase B enumerates the set itself (S2_SLOW).
template <int NW>
__global__ void __launch_bounds__(32 * NW, FH_S2_MINB) k_stage2(const SeqDev *__restrict__ seqs, int seq0, Geo g, fh264_params prm)
{
    __shared__ S2Warp sm[NW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    S2Warp *w = &sm[warp];
    const SeqDev &S = seqs[seq0 + blockIdx.y];
    if (S.status[ST_GATE]) return;
    const int part = g.band_mb0 * 4 + blockIdx.x * NW + warp;
    int xP, yP;
    part_origin(g, part, xP, yP);
    uint2 rows[8];
    load_cur8x8(S.cur[0], g, xP, yP, rows);
    int s[5];
    block_sums(rows, s);
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
    int jb = 180;                     // running upper bound of j_stop
    int ns = 0;                       // kept entries (warp-uniform)
    bool countonly = false;           // the buffer overflowed for good: only the (j, side) counts go on
    // the four gates in 16-bit lanes: (|dx|, |dy|) and (j, |dK1|) by max(a - b, b - a), the Manhattan sum by a dot product
    const uint32_t Pxy = (uint32_t)xP | ((uint32_t)yP << 16), S01 = (uint32_t)s[0] | ((uint32_t)s[1] << 16);
    uint32_t LIM01 = (uint32_t)jb | (99u << 16);
    const int s2m = s[2] - 99;
    auto visit = [&](const uint4 v, uint32_t eidx) {
        const uint32_t ad = __vmaxs2(__vsub2(v.x, Pxy), __vsub2(Pxy, v.x)), ae = __vmaxs2(__vsub2(v.y, S01), __vsub2(S01, v.y));
        const bool ok = __vmaxu2(ae, LIM01) == LIM01 && __dp2a_lo(ad, 0x0101u, 0u) < 280u && (uint32_t)((int)(v.z & 0xffffu) - s2m) < 199u;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m == 0u) return;
        if (ok) {
            const int j = (int)(ae & 0xffffu), side = (int)(v.y & 0xffffu) > s[0];
            if (countonly) atomicAdd(&w->bins[FH_IDX(2 * j + side, S2_BINS)], 1u);
            else {
                const int pos = ns + __popc(m & ((1u << lane) - 1u));
                const int dx = (int)(v.x & 0xffffu) - xP, dy = (int)(v.x >> 16) - yP;
                w->akey[FH_IDX(pos, S2_CAP)] = ((uint32_t)j << 21) | ((uint32_t)side << 20) | ((uint32_t)(dx + 279) << 10) | (uint32_t)(dy + 279);
                w->aval[pos] = eidx;
            }
        }
        if (!countonly) ns += __popc(m);
    };
    // first j whose running gated count (bucket s0 twice) exceeds 128 (:496), else 180, from the (j, side) counts. Lane l owns
    // j = 6l .. 6l+5 (bins 12l .. 12l+11).
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
    auto recount = [&]() -> int {     // (j, side) counts of the kept entries -> bound
        __syncwarp();
        for (int i = lane; i < S2_BINS; i += 32) w->bins[i] = 0;
        __syncwarp();
        for (int i = lane; i < ns; i += 32) atomicAdd(&w->bins[FH_IDX(w->akey[i] >> 20, S2_BINS)], 1u);
        __syncwarp();
        return bound_from_bins();
    };
    // room for `need` more entries: tighten jb from the kept entries and drop those beyond it (warp-uniform call sites)
    auto make_room = [&](int need) {
        if (countonly || ns + need <= S2_CAP) return;
        const int nb = recount();
        int outn = 0;
        for (int base = 0; base < ns; base += 32) {
            const int i = base + lane;
            const uint32_t k = i < ns ? w->akey[i] : 0xffffffffu, e = i < ns ? w->aval[i] : 0u;
            const bool keep = i < ns && (int)(k >> 21) <= nb;
            const unsigned b = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (keep) { const int p = outn + __popc(b & ((1u << lane) - 1u)); w->akey[p] = k; w->aval[p] = e; }
            outn += __popc(b);
            __syncwarp();
        }
        ns = outn; jb = nb;
        LIM01 = (uint32_t)jb | (99u << 16);
        if (ns + need > S2_CAP) {
            // more entries at or below the bound than the buffer holds (thousands of positions share one sum): count only from here on
            countonly = true;
            recount();                // bins = counts of what is kept (all with j <= jb); later entries add to them
        }
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
            for (int e0 = 0; e0 < ln; e0 += 32) {
                make_room(32);
                visit(e0 + lane < ln ? __ldg(tent + gb + e0 + lane) : make_uint4(0, 0xffffu, 0x7fffu, 0), gb + e0 + lane);
            }
            if (lane == src) len = 0;
        }
        // short ranges: chunks of <= 8 consecutive entries (one 128-byte line): first entry | (count - 1) << 28
        const int nc = (len + 7) >> 3;
        int incl = nc;
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        const int pos = nchunk + incl - nc;
        for (int c = 0; c < nc; c++) w->chunk[FH_IDX(pos + c, S2_CHUNK_CAP)] = (gbase + 8u * c) | ((uint32_t)(min(8, len - 8 * c) - 1) << 28);
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
                make_room(32 * FH_S2_UNR);
#pragma unroll
                for (int u = 0; u < FH_S2_UNR; u++) visit(v[u], eid[u]);
            }
            nchunk = 0;
            __syncwarp();
        }
    }
    __syncwarp();
    PartA *pa = &S.parta[part];
    if (countonly) {
        // j_stop is exact: the counts saw every gated entry at or below the bound that was current when it arrived
        const int jsb = bound_from_bins();
        if (lane == 0) { pa->s2_off = 0; pa->n2 = S2_SLOW | (uint32_t)jsb; atomicAdd(&S.status[ST_NSLOW], 1u); }
        return;
    }
    const int js = recount();         // every gated entry with j <= jb is kept and jb >= j_stop: the exact j_stop
    // candidates: kept entries with j <= j_stop, bucket s0 listed twice (the second visit gets side 1, :476,486)
    uint4 *pool = S.s2pool + (size_t)part * S2_SLICE;
    int n2 = 0;
    bool fits = true;
    for (int base = 0; base < ns; base += 32) {
        const int i = base + lane;
        const uint32_t k = i < ns ? w->akey[i] : 0xffffffffu;
        const bool in = i < ns && (int)(k >> 21) <= js;
        const int cnt = in ? ((k >> 21) == 0u ? 2 : 1) : 0;
        int incl = cnt;
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        if (n2 + tot > S2_SLICE - 1) { fits = false; break; }
        if (in) {
            const uint4 v = __ldg(tent + w->aval[i]);
            const int k0 = (int)(v.y & 0xffff), k1 = (int)(v.y >> 16), k2 = (int)(v.z & 0xffff), k3 = (int)(v.z >> 16), k4 = (int)(v.w & 0xffff);
            // feature distance (:267-276) and its largest box-pair term (a lower bound of the SAD)
            const int d0 = s[0] - k0, ad = iabs_(d0);
            const int p1 = max(ad, iabs_(2 * (s[1] - k1) - d0)), p2 = max(ad, iabs_(2 * (s[2] - k2) - d0));
            const int p3 = max(ad, iabs_(2 * (s[3] - k3) - d0)), p4 = max(ad, iabs_(2 * (s[4] - k4) - d0));
            const uint32_t feat = (uint32_t)(ad + p1 + p2 + p3 + p4), lb = (uint32_t)max(max(p1, p2), max(p3, p4));
            const int dx = (int)((k >> 10) & 1023) - 279, dy = (int)(k & 1023) - 279;
            const uint32_t xy = ((uint32_t)dx & 0xffffu) | ((uint32_t)dy << 16);
            const int o = n2 + incl - cnt;
            pool[FH_IDX(o, S2_SLICE)] = make_uint4(xy, feat, lb, k);
            if (cnt == 2) pool[FH_IDX(o + 1, S2_SLICE)] = make_uint4(xy, feat, lb, k | (1u << 20));
        }
        n2 += tot;
    }
    if (lane == 0) {
        pa->s2_off = (uint32_t)part * S2_SLICE;
        pa->n2 = fits ? (uint32_t)n2 : (S2_SLOW | (uint32_t)js);
        if (!fits) atomicAdd(&S.status[ST_NSLOW], 1u);
    }
}
