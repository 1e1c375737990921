// Quarter-pel feature window, staged in shared memory by TMA.
//
// MEstimation(g = window/16, all 16 fractions) (moestimation.cpp:254-296, called at :458 around the predictor and at :510 around
// zero) evaluates w1 x w1 integer positions (w1 = 2*(window/16)+1) on each of the 16 interpolated planes. Every position needs
// the five box sums of an 8x8 block of that plane (:128-139), so a partition needs, per plane, an (8 + w1 - 1)-row window of
// 8 + w1 - 1 <= 16 pixels. The ncu profile of round 1's code showed the kernels that do this bound by the LSU data pipe (80 %
// of its wavefront rate): every lane fetched its own pixel row from its own 128-byte line (one L1 tag request per lane), the
// row sums went through shared memory with bank conflicts, and the SADs of the chosen candidates fetched the same rows again.
// Here ONE cp.async.bulk.tensor.3d per window (box 32 bytes x (R+1) rows x 16 planes, issued by one lane, completion on an
// mbarrier) puts the whole window into shared memory without touching the LSU (the innermost box coordinate of a TMA must be
// 16-byte aligned, so the box starts at x0 & ~15 and is 32 bytes wide; the window keeps its byte offset x0 & 15); a lane owns one (plane, column) of the window,
// computes the horizontal sums of its 8 pixels for each row and slides the vertical sums down the column in registers (no
// intermediate arrays); the SADs of the selected candidates read the same staged window. Windows that touch the outside of
// the picture are filled by ordinary clamped loads instead (replicate padding of the reference's padded plane, :107-115).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "qfeat.cuh"
#include "topk.cuh"

#define QW_ROWB 32                          // bytes per staged row
#define QW_MAX_ROWS 17                      // R + 1 box rows at WindowSize 64 (R = 8 + 2*4)
// per-warp window: [plane][row][32 bytes], plane stride (R+1)*32 bytes (R + 1 is odd: the planes fall into four bank classes)
__host__ __device__ __forceinline__ int qwin_bytes(int g1) { return 16 * (8 + 2 * g1 + 1) * QW_ROWB; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// box (32 bytes, rows, 16 planes) of the plane tensor at (x, y, plane 0), x a multiple of 16 -> dst; completes its bytes on bar
__device__ __forceinline__ void tma_load_window(const CUtensorMap *map, void *dst, uint64_t *bar, int x, int y)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // earlier generic-proxy accesses of dst are done
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(x), "r"(y), "r"(0), "r"(smem_u32(bar)) : "memory");
}

struct QWinGeo { int w1, R, rows; };        // positions per side, pixel rows needed, box rows (R + 1: odd => conflict-free plane stride)
__device__ __forceinline__ QWinGeo qwin_geo(int g1) { QWinGeo q; q.w1 = 2 * g1 + 1; q.R = 8 + q.w1 - 1; q.rows = q.R + 1; return q; }

// Starts the fill of this warp's window with the pixels [x0, x0+16) x [y0, y0+R) of all 16 planes; `off` = byte offset of
// pixel x0 inside the staged rows. Interior windows: one TMA (wait with qwin_wait); windows touching the outside of the
// picture: clamped loads by the whole warp (complete on return). Returns true when a TMA is in flight. Warp-uniform.
__device__ __forceinline__ bool qwin_fill(const SeqDev &S, const Geo &g, const CUtensorMap *map, const QWinGeo &q, int x0, int y0, uint8_t *win, uint64_t *bar, int &off)
{
    const int lane = threadIdx.x & 31;
    __syncwarp();
    const int xa = x0 & ~15;
    if (map && x0 >= 0 && xa + QW_ROWB <= g.W && y0 >= 0 && y0 + q.rows <= g.H) {
        if (lane == 0) { mbar_expect_tx(bar, (uint32_t)(16 * q.rows * QW_ROWB)); tma_load_window(map, win, bar, xa, y0); }
        off = x0 - xa;
        return true;
    }
    for (int t = lane; t < 16 * q.R; t += 32) {
        const int f = t / q.R, r = t - f * q.R;
        *(uint4 *)(win + (size_t)(f * q.rows + r) * QW_ROWB) = qf_load16(S.planes + (size_t)f * g.WH, g.W, g.H, x0, y0 + r);
    }
    off = 0;
    __syncwarp();
    return false;
}
__device__ __forceinline__ void qwin_wait(bool tma, uint64_t *bar, uint32_t &phase)
{
    if (tma) { mbar_wait(bar, phase); phase ^= 1u; }
}

// A staged window as one partition sees it: p = byte 0 of its first row in plane 0, rowb / planeb = row / plane stride in bytes,
// off = byte offset of the partition's first pixel column inside a row.
struct QWinView { const uint8_t *p; int rowb, planeb, off; };

// Costs of the w1 x w1 x 16 candidates of the window centred at (xP + Gx, yP + Gy): candidate ((cx*w1) + cy)*16 + f has cost
// (|cx-g1| + |cy-g1| + 4) * feature distance (:267-276); none where the block origin leaves the picture (:265). Every candidate
// is handed to emit(cost, index) — called by all 32 lanes together, cost QW_COST_NONE for "nothing".
// A lane owns (plane f, column cx): horizontal sums of its 8 pixels for every row (all 8 | columns 0-3 packed in one word,
// columns {0,1,4,5} in another), pair sums down the column, then the five box sums of each of its w1 positions by one or two
// adds each: rows 0-3 = Q[cy], all = Q[cy] + Q[cy+4], rows {0,1,4,5} = P[cy] + P[cy+4] with P[j] = row j + row j+1, Q[j] = P[j] + P[j+2].
#define QW_COST_NONE 0xffffffffu
template <int W1, typename Emit>
__device__ __forceinline__ void qwin_costs(const Geo &g, const QWinView &v, int xP, int yP, int Gx, int Gy, const FeatQ &fq, Emit emit)
{
    constexpr int R = 8 + W1 - 1, G1 = (W1 - 1) / 2;
    const int lane = threadIdx.x & 31;
    const int x0 = xP + Gx - G1, y0 = yP + Gy - G1;
    const int rw = v.rowb >> 2;
#pragma unroll 1
    for (int u0 = 0; u0 < 16 * W1; u0 += 32) {
        const bool active = u0 + lane < 16 * W1;
        const int u = active ? u0 + lane : 0;
        const int f = u / W1, cx = u - f * W1;
        const uint32_t *wrow = (const uint32_t *)(v.p + (size_t)f * v.planeb) + ((cx + v.off) >> 2);
        const uint32_t sh = (uint32_t)((cx + v.off) & 3) * 8;
        uint32_t A[R], B[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t w0 = wrow[r * rw], w1 = wrow[r * rw + 1], w2 = wrow[r * rw + 2];
            const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
            const uint32_t s4 = __vsadu4(lo, 0u), s8 = __vsadu4(hi, 0u) + s4;
            A[r] = s8 | (s4 << 16);
            B[r] = __vsadu4(__byte_perm(lo, hi, 0x5410), 0u);
        }
        uint32_t P[R - 1], Q[R - 3], PB[R - 1], QB[R - 3];
#pragma unroll
        for (int r = 0; r < R - 1; r++) { P[r] = A[r] + A[r + 1]; PB[r] = B[r] + B[r + 1]; }
#pragma unroll
        for (int r = 0; r < R - 3; r++) { Q[r] = P[r] + P[r + 2]; QB[r] = PB[r] + PB[r + 2]; }
        const int rx = x0 + cx;
        const bool xok = active && rx >= 0 && rx < g.W;
        const int adx = iabs_(cx - G1) + 4;
#pragma unroll
        for (int cy = 0; cy < W1; cy++) {
            const uint32_t all = Q[cy] + Q[cy + 4];                  // K0 | K2 << 16
            const int k0 = (int)(all & 0xffffu), k2 = (int)(all >> 16), k1 = (int)(Q[cy] & 0xffffu), k3 = (int)((P[cy] + P[cy + 4]) & 0xffffu);
            const int k4 = (int)(QB[cy] + QB[cy + 4]);
            const int ry = y0 + cy;
            uint32_t cst = QW_COST_NONE;
            if (xok && ry >= 0 && ry < g.H) cst = (uint32_t)((adx + iabs_(cy - G1)) * feat_of(fq, feat_record(k0, k1, k2, k3, k4)));
            emit(cst, (uint32_t)((cx * W1 + cy) * 16 + f));
        }
    }
}
template <typename Emit>
__device__ __forceinline__ void qwin_costs_w(int w1, const Geo &g, const QWinView &v, int xP, int yP, int Gx, int Gy, const FeatQ &fq, Emit emit)
{
    switch (w1) {
    case 1: qwin_costs<1>(g, v, xP, yP, Gx, Gy, fq, emit); break;
    case 3: qwin_costs<3>(g, v, xP, yP, Gx, Gy, fq, emit); break;
    case 5: qwin_costs<5>(g, v, xP, yP, Gx, Gy, fq, emit); break;
    case 7: qwin_costs<7>(g, v, xP, yP, Gx, Gy, fq, emit); break;
    default: qwin_costs<9>(g, v, xP, yP, Gx, Gy, fq, emit); break;
    }
}

// One row (8 pixels) of the candidate block (plane f, window column cx, window row wr) from the staged window.
__device__ __forceinline__ uint2 qwin_row8(const QWinView &v, int f, int cx, int wr)
{
    const uint32_t *w = (const uint32_t *)(v.p + (size_t)f * v.planeb + (size_t)wr * v.rowb) + ((cx + v.off) >> 2);
    const uint32_t sh = (uint32_t)((cx + v.off) & 3) * 8;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

// The same candidates with selection (topk.cuh). Small windows (w1 <= 5: at most 15 candidates per lane): every cost is tracked
// and parked in `stage` (a [15][32] word array per warp, one wavefront per access); the bound is tightened ONCE over all of them
// and only the candidates at or below it are appended to the selection buffer, with arrival index base + ((cx*w1) + cy)*16 + f.
// Larger windows stream through tk_offer. The loops stay rolled: the unrolled window code is already 6 KB of instructions.
#define QW_STAGE_WORDS (15 * 32)
template <int W1>
__device__ __forceinline__ void qwin_select(const Geo &g, const QWinView &v, int xP, int yP, int Gx, int Gy, const FeatQ &fq, u64 *buf, TopK &tk,
                                            uint32_t *stage, uint32_t base)
{
    const int lane = threadIdx.x & 31;
    if constexpr (W1 <= 5) {
        constexpr int NP = (16 * W1 + 31) / 32;
        int n = 0;
        qwin_costs<W1>(g, v, xP, yP, Gx, Gy, fq, [&](uint32_t cst, uint32_t) { stage[FH_IDX(n * 32 + lane, QW_STAGE_WORDS)] = cst; n++; tk_track(tk, cst); });
        tk_tighten_exact_minima(tk);
        __syncwarp();
#pragma unroll 1
        for (int p = 0; p < NP; p++) {
            const int u = p * 32 + lane, f = u / W1, cx = u - f * W1;
            const uint32_t ib = base + (uint32_t)(cx * W1 * 16 + f);
#pragma unroll 1
            for (int cy = 0; cy < W1; cy++) tk_append(buf, tk, stage[(p * W1 + cy) * 32 + lane], ib + (uint32_t)(cy * 16));
        }
    } else {
        qwin_costs<W1>(g, v, xP, yP, Gx, Gy, fq, [&](uint32_t cst, uint32_t idx) { tk_offer(buf, tk, cst, base + idx); });
        tk_tighten_exact_minima(tk);
        tk_filter(buf, tk, tk.bound);
    }
}
template <typename = void>
__device__ __forceinline__ void qwin_select_w(int w1, const Geo &g, const QWinView &v, int xP, int yP, int Gx, int Gy, const FeatQ &fq, u64 *buf, TopK &tk,
                                              uint32_t *stage, uint32_t base)
{
    switch (w1) {
    case 1: qwin_select<1>(g, v, xP, yP, Gx, Gy, fq, buf, tk, stage, base); break;
    case 3: qwin_select<3>(g, v, xP, yP, Gx, Gy, fq, buf, tk, stage, base); break;
    case 5: qwin_select<5>(g, v, xP, yP, Gx, Gy, fq, buf, tk, stage, base); break;
    case 7: qwin_select<7>(g, v, xP, yP, Gx, Gy, fq, buf, tk, stage, base); break;
    default: qwin_select<9>(g, v, xP, yP, Gx, Gy, fq, buf, tk, stage, base); break;
    }
}
