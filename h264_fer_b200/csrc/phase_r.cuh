// Phase R — per reference picture, fully parallel (reference: FillInterpolatedRefFrame, moestimation.cpp:74-173).
//   k_interp      16 quarter-pel luma planes             (:79-104 via mocomp.cpp:50-78)
//   k_features    5 box-sum features x 16 planes, packed 16 B per position (:105-139)
//   k_tile_index  plane-0 positions bucketed per 64x64 tile by (K0>>7, K1>>6); replaces the global counting
//                 sort sortedSuma0/koliko (:140-172) with an index from which stage 2 enumerates the same set
//   k_scene_sad   sum |frame - dpb| over luma            (ref_frames.cpp:210-224; h264_kernels.cl:1-5)
#pragma once
#include "common.cuh"

#define IT_W 64
#define IT_H 16
#define IT_PW (IT_W + 6)   // 70 columns: x0-2 .. x0+IT_W+3

__global__ void __launch_bounds__(256) k_interp(const SeqDev *__restrict__ seqs, int seq0, Geo g, int yoff)
{
    const SeqDev &S = seqs[seq0 + blockIdx.z];
    const uint8_t *__restrict__ ref = S.ref[0];
    __shared__ uint8_t pix[IT_H + 6][IT_PW + 2];   // E(x0-2+c, y0-2+r): edge-extended reference (mocomp.cpp:11-23)
    __shared__ uint8_t hv[IT_H][IT_PW + 2];        // column half-pel at (x0-2+c, y0+r)
    __shared__ uint8_t bh[IT_H + 1][IT_W];         // row half-pel at (x0+c, y0+r)
    const int x0 = blockIdx.x * IT_W, y0 = yoff + blockIdx.y * IT_H, tid = threadIdx.x;       // yoff: first row of the band's halo (band mode)
    const int W = g.W, H = g.H;

    for (int i = tid; i < (IT_H + 6) * IT_PW; i += 256) {
        int r = i / IT_PW, c = i - r * IT_PW;
        pix[r][c] = ref[(size_t)clampi_(y0 - 2 + r, 0, H - 1) * W + clampi_(x0 - 2 + c, 0, W - 1)];
    }
    __syncthreads();
    for (int i = tid; i < IT_H * IT_PW; i += 256) {
        int r = i / IT_PW, c = i - r * IT_PW;
        hv[r][c] = (uint8_t)tap6_(pix[r][c], pix[r + 1][c], pix[r + 2][c], pix[r + 3][c], pix[r + 4][c], pix[r + 5][c]);
    }
    for (int i = tid; i < (IT_H + 1) * IT_W; i += 256) {
        int r = i / IT_W, c = i - r * IT_W;
        const uint8_t *p = &pix[r + 2][c];
        bh[r][c] = (uint8_t)tap6_(p[0], p[1], p[2], p[3], p[4], p[5]);
    }
    __syncthreads();

    const int ty = tid >> 4, tx = tid & 15;
    const int y = y0 + ty, x = x0 + 4 * tx;
    if (y >= H || x >= W) return;
    uint32_t out[16];
#pragma unroll
    for (int f = 0; f < 16; f++) out[f] = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int c = 4 * tx + k;
        const int G = pix[ty + 2][c + 2], Gr = pix[ty + 2][c + 3], Gd = pix[ty + 3][c + 2];
        const int b = bh[ty][c], s = bh[ty + 1][c], h = hv[ty][c + 2], m = hv[ty][c + 3];
        // centre half-pel from ROUNDED column half-pels (mocomp.cpp:67-71)
        const int j = tap6_(hv[ty][c], hv[ty][c + 1], h, m, hv[ty][c + 4], hv[ty][c + 5]);
        int v[16];
        v[0] = G;          v[1] = mid_(G, b);  v[2] = b;          v[3] = mid_(b, Gr);
        v[4] = mid_(G, h); v[5] = mid_(b, h);  v[6] = mid_(b, j); v[7] = mid_(b, m);
        v[8] = h;          v[9] = mid_(h, j);  v[10] = j;         v[11] = mid_(j, m);
        v[12] = mid_(h, Gd); v[13] = mid_(h, s); v[14] = mid_(j, s); v[15] = mid_(s, m);
#pragma unroll
        for (int f = 0; f < 16; f++) out[f] |= (uint32_t)v[f] << (8 * k);
    }
    uint8_t *dst = S.planes + (size_t)y * W + x;
#pragma unroll
    for (int f = 0; f < 16; f++) *(uint32_t *)(dst + (size_t)f * g.WH) = out[f];
}

#define FT_W 64
#define FT_H 32
// Box-sum features of one plane tile (64 x 32 positions). Horizontal pass: a thread takes 4 consecutive positions of a
// row from three aligned words (12 bytes) and slides the 8-wide / 4-wide / columns{0,1,4,5} sums. Vertical pass: a thread
// walks down one column and slides the 8-row / 4-row / rows{0,1,4,5} sums: ~35 instructions per position instead of the
// ~140 of a direct evaluation, so the kernel is bound by its 16-byte-per-position HBM writes.
// Only plane `f` is materialised (phase R: f = 0 into S.kar, read by stage 3's integer window and the stage-2 index; the
// quarter-pel planes' features are computed where they are needed, qfeat.cuh). `out` overrides the destination (debug tap).
__global__ void __launch_bounds__(256) k_features(const SeqDev *__restrict__ seqs, int seq0, Geo g, int f, uint4 *__restrict__ out, int yoff)
{
    const SeqDev &S = seqs[seq0 + blockIdx.z];
    const uint8_t *__restrict__ pl = S.planes + (size_t)f * g.WH;
    __shared__ __align__(16) uint32_t t[FT_H + 8][(FT_W + 8) / 4];      // padded plane rows as words
    __shared__ uint16_t r8[FT_H + 8][FT_W + 2], r4[FT_H + 8][FT_W + 2], rc[FT_H + 8][FT_W + 2];
    const int x0 = blockIdx.x * FT_W, y0 = yoff + blockIdx.y * FT_H, tid = threadIdx.x;
    const int W = g.W, H = g.H;
    // padded plane: replicate the last column / row (moestimation.cpp:107-115). W is a multiple of 16, so a word is either
    // entirely inside the picture or entirely in the padding.
    for (int i = tid; i < (FT_H + 8) * ((FT_W + 8) / 4); i += 256) {
        const int r = i / ((FT_W + 8) / 4), cw = i - r * ((FT_W + 8) / 4), x = x0 + 4 * cw, y = min(y0 + r, H - 1);
        uint32_t w;
        if (x < W) w = __ldg((const uint32_t *)(pl + (size_t)y * W + x));
        else w = 0x01010101u * (uint32_t)pl[(size_t)y * W + W - 1];
        t[r][cw] = w;
    }
    __syncthreads();
    // horizontal sums, 4 positions per thread
    for (int i = tid; i < (FT_H + 8) * (FT_W / 4); i += 256) {
        const int r = i / (FT_W / 4), c4 = i - r * (FT_W / 4);
        const uint32_t w0 = t[r][c4], w1 = t[r][c4 + 1], w2 = t[r][c4 + 2];
        int b[12];
#pragma unroll
        for (int k = 0; k < 4; k++) { b[k] = (w0 >> (8 * k)) & 255; b[4 + k] = (w1 >> (8 * k)) & 255; b[8 + k] = (w2 >> (8 * k)) & 255; }
        int s8 = (int)__vsadu4(w0, 0) + (int)__vsadu4(w1, 0), s4 = (int)__vsadu4(w0, 0);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            r8[r][4 * c4 + k] = (uint16_t)s8;
            r4[r][4 * c4 + k] = (uint16_t)s4;
            rc[r][4 * c4 + k] = (uint16_t)(b[k] + b[k + 1] + b[k + 4] + b[k + 5]);
            s8 += b[k + 8] - b[k];
            s4 += b[k + 4] - b[k];
        }
    }
    __syncthreads();
    // vertical sliding sums: thread = (column, quarter of the rows)
    const int tx = tid & 63, tg = tid >> 6;
    const int x = x0 + tx;
    if (x >= W) return;
    uint4 *K = out ? out : S.kar;
    const int rs = tg * (FT_H / 4);
    int a[8], q4[8], qc[8];                // rows rs .. rs+7 of the three horizontal sums (ring of 8)
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = r8[rs + i][tx]; q4[i] = r4[rs + i][tx]; qc[i] = rc[rs + i][tx]; }
    int k0 = 0, k2 = 0, k4 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { k0 += a[i]; k2 += q4[i]; k4 += qc[i]; }
#pragma unroll
    for (int q = 0; q < FT_H / 4; q++) {
        const int y = y0 + rs + q;
        // ring position of row (rs+q+i) is (q+i)&7 — compile-time after unrolling
        const int k1 = a[q & 7] + a[(q + 1) & 7] + a[(q + 2) & 7] + a[(q + 3) & 7];
        const int k3 = a[q & 7] + a[(q + 1) & 7] + a[(q + 4) & 7] + a[(q + 5) & 7];
        // K0: 8x8 (:137) | K1: rows 0-3 (:136) ; K2: columns 0-3 (:135) | K3: rows 0,1,4,5 (:133-134) ; K4: columns 0,1,4,5 (:131-132)
        if (y < H) {
            K[(size_t)y * W + x] = feat_record(k0, k1, k2, k3, k4);
            if (!out) S.k0p[(size_t)y * W + x] = (uint16_t)k0;
        }
        if (q + 1 < FT_H / 4) {
            const int na = r8[rs + q + 8][tx], n4 = r4[rs + q + 8][tx], nc = rc[rs + q + 8][tx];
            k0 += na - a[q & 7]; k2 += n4 - q4[q & 7]; k4 += nc - qc[q & 7];
            a[q & 7] = na; q4[q & 7] = n4; qc[q & 7] = nc;
        }
    }
}

// One CTA per 64x64 tile of plane-0 positions. Cell = (K1>>6)*128 + (K2>>6) (the two half-sum gates of stage 2). Order inside a cell is arbitrary:
// stage 2 re-derives the reference's arrival order (sum bucket, side, x, y) from the entry itself.
__global__ void __launch_bounds__(256) k_tile_index(const SeqDev *__restrict__ seqs, int seq0, Geo g, int tile0)
{
    extern __shared__ uint32_t hist[];   // FH_CELLS counters
    __shared__ uint32_t wsum[8];
    const SeqDev &S = seqs[seq0 + blockIdx.y];
    const int tile = tile0 + blockIdx.x, tid = threadIdx.x;
    const int tx0 = (tile % g.tilesx) * FH_TILE, ty0 = (tile / g.tilesx) * FH_TILE;
    const int tw = min(FH_TILE, g.W - tx0), th = min(FH_TILE, g.H - ty0);
    const uint4 *__restrict__ K = S.kar;      // plane 0 (f = 0) feature records
    for (int i = tid; i < FH_CELLS; i += 256) hist[i] = 0;
    __syncthreads();
    bool ub = false;
    for (int i = tid; i < tw * th; i += 256) {
        int ly = i / tw, lx = i - ly * tw;
        size_t o = (size_t)(ty0 + ly) * g.W + tx0 + lx;
        const uint4 kv = K[o];
        const int k0 = feat_raw(kv, 0), k1 = feat_raw(kv, 1), k2 = feat_raw(kv, 2);
        ub |= (k0 == 0) | (k0 >= 16203);
        atomicAdd(&hist[((k1 >> 6) << 7) | (k2 >> 6)], 1u);
    }
    if (ub) atomicOr(&S.status[ST_FLAGS_NEXT], FLAG_UB_INPUT);
    __syncthreads();
    // exclusive scan of 16384 counters: each warp owns 2048 consecutive cells and walks them 32 at a time
    // (bank-conflict free), after a first pass that produces the per-warp totals
    const int lane = tid & 31, wid = tid >> 5;
    const int base = wid * 2048;
    uint32_t local = 0;
    for (int it = 0; it < 64; it++) local += hist[base + it * 32 + lane];
    for (int d = 16; d; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if (lane == 0) wsum[wid] = local;
    __syncthreads();
    uint32_t run = 0;
    for (int w = 0; w < wid; w++) run += wsum[w];
    uint16_t *ts = S.tstart + (size_t)tile * FH_TSTART_PITCH;
    for (int it = 0; it < 64; it++) {
        const int cidx = base + it * 32 + lane;
        const uint32_t v = hist[cidx];
        uint32_t incl = v;
        for (int d = 1; d < 32; d <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
        const uint32_t ex = run + incl - v;
        hist[cidx] = ex;
        ts[cidx] = (uint16_t)ex;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (tid == 255) ts[FH_CELLS] = (uint16_t)run;
    __syncthreads();
    TileEntry *te = S.tent + (size_t)tile * (FH_TILE * FH_TILE);
    for (int i = tid; i < tw * th; i += 256) {
        int ly = i / tw, lx = i - ly * tw;
        size_t o = (size_t)(ty0 + ly) * g.W + tx0 + lx;
        const uint4 kv = K[o];
        TileEntry e;
        e.x = (uint16_t)(tx0 + lx); e.y = (uint16_t)(ty0 + ly);
        e.k0 = (uint16_t)feat_raw(kv, 0); e.k1 = (uint16_t)feat_raw(kv, 1); e.k2 = (uint16_t)feat_raw(kv, 2);
        e.k3 = (uint16_t)feat_raw(kv, 3); e.k4 = (uint16_t)feat_raw(kv, 4); e.pad = 0;
        uint32_t pos = atomicAdd(&hist[((e.k1 >> 6) << 7) | (e.k2 >> 6)], 1u);
        *(uint4 *)&te[pos] = *(const uint4 *)&e;
    }
}

// Scene-change measure. grid.x blocks per sequence, grid.y = sequence. Result: 64-bit sum in status[ST_SAD_LO/HI].
// (band mode: rows [y0, y1) = this rank's band — the bands' sums add up to the picture's)
__global__ void __launch_bounds__(256) k_scene_sad(const SeqDev *__restrict__ seqs, int seq0, Geo g, int y0, int y1)
{
    const SeqDev &S = seqs[seq0 + blockIdx.y];
    const uint4 *__restrict__ a = (const uint4 *)(S.cur[0] + (size_t)y0 * g.W);
    const uint4 *__restrict__ b = (const uint4 *)(S.ref[0] + (size_t)y0 * g.W);
    const int n16 = ((y1 - y0) * g.W) >> 4;   // W is a multiple of 16
    uint32_t acc = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n16; i += gridDim.x * 256) {
        uint4 p = a[i], q = b[i];
        acc += __vsadu4(p.x, q.x) + __vsadu4(p.y, q.y) + __vsadu4(p.z, q.z) + __vsadu4(p.w, q.w);
    }
    for (int d = 16; d; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    __shared__ uint32_t ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; w++) t += ws[w];
        atomicAdd((unsigned long long *)&S.status[ST_SAD_LO], t);
    }
}
