// Box-sum features of a small quarter-pel window, computed from the interpolated planes instead of being read from a
// materialised feature array. The reference keeps refFrameKar[k][frac][y][x] for all 16 planes (moestimation.cpp:105-139,
// 80 int planes = 666 MB at 1080p); here only plane 0 is materialised (stage 3's integer window and the stage-2 index
// read it everywhere). The (2*g1+1)^2 x 16 candidates of MEstimation(window/16) (:458-469, :508-520) need, per plane, the
// pixels of an (8 + 2*g1)^2 window: 33 MB of planes (L2 resident) are read instead of 527 MB of features (HBM).
//   step A  one thread per (plane, window row): sliding horizontal sums of the row for the w1 positions
//           X = (8-wide sum) | (columns 0-3) << 16, RC = columns {0,1,4,5}
//   step B  one thread per (plane, position): vertical sums of 8 rows -> feature record (common.cuh feat_record)
// Replicate padding at the right / bottom picture edge as the reference's padded plane (:107-115).
#pragma once
#include "common.cuh"

#define QF_MAXW1 9        // window/16 = 4 at WindowSize 64

// 16 pixels of a plane row starting at x0 (any x0; columns clamped into the row), as four little-endian words.
// The clamped (picture edge) case and the row sums are kept out of line: they are used from unrolled code in two large
// kernels whose instruction-cache footprint matters more than a call.
__device__ __noinline__ uint4 qf_load16_edge(const uint8_t *__restrict__ row, int W, int x0)
{
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) v |= (uint32_t)row[clampi_(x0 + 4 * k + i, 0, W - 1)] << (8 * i);
        w[k] = v;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 qf_load16(const uint8_t *__restrict__ plane, int W, int H, int x0, int y)
{
    const uint8_t *row = plane + (size_t)clampi_(y, 0, H - 1) * W;
    if (x0 >= 0 && x0 + 16 <= W) {
        const uint2 a = load8_unaligned(row + x0), b = load8_unaligned(row + x0 + 8);
        return make_uint4(a.x, a.y, b.x, b.y);
    }
    return qf_load16_edge(row, W, x0);
}

// step A for one row: w1 <= 9 positions, pixels b[p .. p+7] of the 16 loaded. Position p's 8 pixels are two funnel-shifted
// words (lo = pixels p..p+3, hi = p+4..p+7): 8-wide sum = sad(lo) + sad(hi), columns 0-3 = sad(lo), columns {0,1,4,5} =
// sad of the low halves (VABSDIFF4.U8.ACC adds for free).
__device__ __forceinline__ void qf_row_sums(const uint4 v, int w1, uint32_t *__restrict__ X, uint16_t *__restrict__ RC)
{
    const uint32_t w[5] = { v.x, v.y, v.z, v.w, 0u };
#pragma unroll
    for (int p = 0; p < QF_MAXW1; p++) {
        if (p >= w1) break;
        const int k = p >> 2, sh = (p & 3) * 8;
        const uint32_t lo = __funnelshift_r(w[k], w[k + 1], sh), hi = __funnelshift_r(w[k + 1], w[k + 2 < 5 ? k + 2 : 4], sh);
        const uint32_t s4 = __vsadu4(lo, 0u), s8 = __vsadu4(hi, 0u) + s4;
        X[p] = s8 | (s4 << 16);
        RC[p] = (uint16_t)(__vsadu4(lo & 0xffffu, 0u) + __vsadu4(hi & 0xffffu, 0u));
    }
}

// step B: feature record of plane slot `fl` at window position (cx, cy); ps = plane stride in X / RC, w1 = row stride
__device__ __forceinline__ uint4 qf_record(const uint32_t *__restrict__ X, const uint16_t *__restrict__ RC, int fl, int ps, int w1, int cx, int cy)
{
    const int base = fl * ps + cy * w1 + cx;
    uint32_t all = 0, top = 0, alt = 0, k4 = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t x = X[base + j * w1];
        all += x;
        if (j < 4) top += x;
        if ((j & 3) < 2) alt += x;
        k4 += RC[base + j * w1];
    }
    // K0: 8x8 | K1: rows 0-3 | K2: columns 0-3 | K3: rows 0,1,4,5 | K4: columns 0,1,4,5 (moestimation.cpp:131-137)
    return feat_record((int)(all & 0xffffu), (int)(top & 0xffffu), (int)(all >> 16), (int)(alt & 0xffffu), (int)k4);
}
