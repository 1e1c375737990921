// One macroblock of an I picture: what the reference's I-slice macroblock loop computes before it
// writes the macroblock (rbsp_encoding.cpp:196-215): intraPredictionEncoding (intra.cpp:949-1109) = Intra16x16 mode search,
// Intra4x4 mode search on the unreconstructed macroblock, block-by-block Intra4x4 coding, the two coded_mb_size() bit-cost
// trials (rbsp_encoding.cpp:330-487 with residual_block_cavlc_size, residual.cpp:673-957) and the decision; then
// quantizationTransform(..., true) (quantizationTransform.cpp:349-485) of the winner with its in-loop reconstruction.
// The semantics are the CPU (non-OpenCL) ones (intra.cpp:978-1049).
//
// The core is plain scalar C++ that compiles for the device (intra.cuh, the product path) and for the host (tests only: the
// same source is checked against the compiled reference's I pictures without a GPU, tests/intra_host.cpp).
#pragma once
#include "../../include/fh264_b200.h"
#include "cavlc_core.h"

// ---- lanes: on the device the macroblock is worked on by nl = 32 lanes of one warp (the two mode searches are spread over the
//      lanes, the serial remainder runs on lane 0) or by nl = 1 lane; on the host always by one -------------------------------
#ifdef __CUDACC__
FH_HD void ic_sync(int nl) { if (nl > 1) __syncwarp(); }
FH_HD int ic_red_add(int v, int nl) { return nl > 1 ? __reduce_add_sync(0xffffffffu, v) : v; }
FH_HD int ic_red_min(int v, int nl) { return nl > 1 ? __reduce_min_sync(0xffffffffu, v) : v; }
#elif !defined(IC_CUSTOM_LANES)     // (a host test may supply its own lane primitives: tests/intra_host_lanes.cpp runs 32 threads)
FH_HD void ic_sync(int) {}
FH_HD int ic_red_add(int v, int) { return v; }
FH_HD int ic_red_min(int v, int) { return v; }
#endif

// ---- tables ---------------------------------------------------------------------------------------------------------------
FH_TAB int16_t ic_LQ[6][3] = { { 205, 158, 128 }, { 186, 146, 114 }, { 158, 128, 102 }, { 146, 114, 89 }, { 128, 102, 82 }, { 114, 89, 71 } };   // LevelQuantize by (row & 1) + (col & 1), quantizationTransform.cpp:24-32
FH_TAB int16_t ic_LS[6][3] = { { 160, 208, 256 }, { 176, 224, 288 }, { 208, 256, 320 }, { 224, 288, 368 }, { 256, 320, 400 }, { 288, 368, 464 } };   // LevelScale, scaleTransform.cpp:32-40
FH_TAB uint8_t ic_ZZ[16] = { 0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15 };           // zigzag: scan position -> row * 4 + col (scaleTransform.cpp:43-47)
FH_TAB uint8_t ic_ZZI[16] = { 0, 1, 5, 6, 2, 4, 7, 12, 3, 8, 11, 13, 9, 10, 14, 15 };          // row * 4 + col -> scan position
FH_TAB uint8_t ic_QPC[52] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30,
                              31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39 };                              // inttransform.cpp:8-14
// coded_block_pattern -> codeNum for Intra macroblocks (Table 9-4, ChromaArrayType 1; h264_globals.cpp:155)
FH_TAB uint8_t ic_cbp_intra[48] = { 3, 29, 30, 17, 31, 18, 37, 8, 32, 38, 19, 9, 20, 10, 11, 2, 16, 33, 34, 21, 35, 22, 39, 4, 36, 40, 23, 5, 24, 6, 7, 1,
                                    41, 42, 43, 25, 44, 26, 46, 12, 45, 47, 27, 13, 28, 14, 15, 0 };
FH_TAB uint8_t ic_chroma_of_16[4] = { 2, 1, 0, 3 };                                          // intraToChromaPredMode, intra.cpp:16

// ---- per-macroblock state the neighbours read (48 bytes): final mb_type, CodedBlockPattern, TotalCoeff of the coded blocks
//      (0 elsewhere, which is what residual.cpp:473,493 substitute), Intra4x4PredMode ------------------------------------------
struct IcInfo {
    uint8_t mb_type, cbp_luma, cbp_chroma, is4x4;
    uint8_t tc_luma[16];
    uint8_t tc_chroma[2][4];
    uint8_t mode4[16];
    uint8_t pad[4];
};

// levels of one macroblock
struct IcLevels {
    int16_t luma[16][16];       // LumaLevel (Intra4x4)
    int16_t dc16[16];           // Intra16x16DCLevel
    int16_t ac16[16][15];       // Intra16x16ACLevel
    int16_t cdc[2][4];
    int16_t cac[2][4][15];
};

// Working state of one macroblock, shared by the lanes that work on it (shared memory on the device).
struct IcCtx {
    const uint8_t *src[3];      // `frame` as read from the file
    uint8_t *rec[3];            // `frame` as left by the macroblocks already coded (reconstruction)
    int W, H, xP, yP, qp, qpc;
    uint8_t S[256];             // source luma of this macroblock (originalMB, intra.cpp:1057)
    uint8_t L[256];             // `frame.L` inside this macroblock: source, overwritten block by block by the Intra4x4 reconstruction
    uint8_t SC[2][64];          // source chroma
    uint8_t RC[2][64];          // reconstructed chroma
    uint8_t nbT[24], nbL[16];   // luma samples around the macroblock, fetched once: row above x = -1 .. 19 (nbT[x + 1]), left column
    uint8_t nbC[2][20];         // chroma: [0] corner, [1..8] left column, [9..16] row above
    uint8_t pred16[256];        // Intra16x16 prediction of the chosen mode
    uint8_t predC[2][64];       // chroma prediction
    IcLevels lv;
    int DC[16], cq16[16];       // Intra16x16: DC of every block before / after the Hadamard + quantisation (row-major)
    int cdcraw[2][4];           // chroma: DC of every block before the 2x2 transform
    int rdc[16], crdc[2][4];    // reconstructed (scaled) DCs
    uint8_t tcl[2][16], tcc[2][2][4];   // TotalCoeff of the coded blocks, [0] Intra16x16 trial, [1] Intra4x4 trial
    uint8_t mode4[16], flag[16], rem[16];
    int cbpl16, cbpl4, cbpc, type16;
    int w0[16], w1[16], pb[16]; // one Intra4x4 block in flight: residual / coefficients, intermediate, prediction
};

FH_HD int ic_abs(int a) { return a < 0 ? -a : a; }
FH_HD int ic_clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
FH_HD int ic_blkx(int b) { return ((b & 1) << 2) | ((b & 4) << 1); }      // Intra4x4ScanOrder, h264_globals.cpp:209-214
FH_HD int ic_blky(int b) { return ((b & 2) << 1) | (b & 8); }
FH_HD int ic_ue_len(int v) { return 2 * cv_ilog2((uint32_t)v + 1u) + 1; }  // expgolomb_UC_codes[v][0] * 2 + 1 (expgolomb.cpp:8-40)

// `frame.L` at absolute (x, y) (inside the picture): inside the current macroblock its working copy, elsewhere the reconstruction
// of the macroblocks coded before — only the row above and the column to the left are ever asked for, and those were staged
// by ic_stage_neighbours
FH_HD int ic_px(const IcCtx &c, int x, int y)
{
    const int lx = x - c.xP, ly = y - c.yP;
    if (ly < 0) return c.nbT[lx + 1];
    if (lx < 0) return c.nbL[ly];
    return c.L[ly * 16 + lx];
}
// one read of every reconstructed sample the macroblock predicts from (spread over the lanes; the caller synchronises)
FH_HD void ic_stage_neighbours(IcCtx &c, int lane, int nl)
{
    const int W = c.W, CW = c.W >> 1, xP = c.xP, yP = c.yP, xM = xP >> 1, yM = yP >> 1;
    for (int i = lane; i < 21 + 16 + 34; i += nl) {
        if (i < 21) { const int x = xP - 1 + i; c.nbT[i] = (yP > 0 && x >= 0 && x < W) ? c.rec[0][(size_t)(yP - 1) * W + x] : 0; }
        else if (i < 37) c.nbL[i - 21] = xP > 0 ? c.rec[0][(size_t)(yP + i - 21) * W + xP - 1] : 0;
        else {
            const int k = (i - 37) / 17, j = (i - 37) % 17;
            const uint8_t *r = c.rec[1 + k];
            int v = 0;
            if (j == 0) { if (xM > 0 && yM > 0) v = r[(size_t)(yM - 1) * CW + xM - 1]; }
            else if (j <= 8) { if (xM > 0) v = r[(size_t)(yM + j - 1) * CW + xM - 1]; }
            else if (yM > 0) v = r[(size_t)(yM - 1) * CW + xM + j - 9];
            c.nbC[k][j] = (uint8_t)v;
        }
    }
}

// ---- transform / quantisation (quantizationTransform.cpp:41-100,183-223; scaleTransform.cpp:101-150,308-340) ----------------
FH_HD void ic_fwd4(int a, int b, int c, int d, int &o0, int &o1, int &o2, int &o3)
{
    o0 = ((a + b + c + d) * 256 + 512) >> 10;
    o1 = (416 * a + 208 * b - 208 * c - 416 * d + 512) >> 10;
    o2 = ((a - b - c + d) * 256 + 512) >> 10;
    o3 = (208 * a - 416 * b + 416 * c - 208 * d + 512) >> 10;
}
FH_HD void ic_forward4x4(const int r[16], int d[16])
{
    int h[16], f[16];
    for (int i = 0; i < 16; i++) h[i] = r[i] == 0 ? 0 : r[i] * 64 - 32;
    for (int j = 0; j < 4; j++) ic_fwd4(h[j], h[4 + j], h[8 + j], h[12 + j], f[j], f[4 + j], f[8 + j], f[12 + j]);
    for (int i = 0; i < 4; i++) ic_fwd4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3], d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
}
FH_HD int ic_quant1(int d, int i, int qP)                 // coefficient i = row * 4 + col
{
    const int per = qP / 6, lq = ic_LQ[qP % 6][((i >> 2) & 1) + (i & 1)];
    const int t = qP < 24 ? (d * (1 << (4 - per)) - (1 << (3 - per))) * lq : (d >> (per - 4)) * lq;
    return (t + 16384) >> 15;
}
FH_HD int ic_dequant1(int c, int i, int qP)
{
    const int per = qP / 6, ls = ic_LS[qP % 6][((i >> 2) & 1) + (i & 1)];
    return qP >= 24 ? (c * ls) * (1 << (per - 4)) : (c * ls + (1 << (3 - per))) >> (4 - per);
}
FH_HD void ic_quant4x4(const int d[16], int c[16], int qP, bool keep_dc)
{
    for (int i = 0; i < 16; i++) c[i] = ic_quant1(d[i], i, qP);
    if (keep_dc) c[0] = d[0];
}
FH_HD void ic_dequant4x4(const int c[16], int d[16], int qP, bool keep_dc)
{
    for (int i = 0; i < 16; i++) d[i] = ic_dequant1(c[i], i, qP);
    if (keep_dc) d[0] = c[0];
}
FH_HD void ic_inverse4x4(const int d[16], int r[16])
{
    int f[16], h[16];
    for (int i = 0; i < 4; i++) {
        const int *q = d + 4 * i;
        const int e0 = q[0] + q[2], e1 = q[0] - q[2], e2 = (q[1] >> 1) - q[3], e3 = q[1] + (q[3] >> 1);
        f[4 * i] = e0 + e3; f[4 * i + 1] = e1 + e2; f[4 * i + 2] = e1 - e2; f[4 * i + 3] = e0 - e3;
    }
    for (int j = 0; j < 4; j++) {
        const int g0 = f[j] + f[8 + j], g1 = f[j] - f[8 + j], g2 = (f[4 + j] >> 1) - f[12 + j], g3 = f[4 + j] + (f[12 + j] >> 1);
        h[j] = g0 + g3; h[4 + j] = g1 + g2; h[8 + j] = g1 - g2; h[12 + j] = g0 - g3;
    }
    for (int i = 0; i < 16; i++) r[i] = (h[i] + 32) >> 6;
}
// residual -> quantised coefficients of one 4x4 block (forwardResidual, quantizationTransform.cpp:284-291)
FH_HD void ic_forward_residual(const int diff[16], int c[16], int qP, bool keep_dc)
{
    int d[16];
    ic_forward4x4(diff, d);
    ic_quant4x4(d, c, qP, keep_dc);
}
// chroma DC: forward + quantisation (quantizationTransform.cpp:157-178,264-282), inverse + scaling (scaleTransform.cpp:247-262,408-420)
FH_HD void ic_chroma_dc_forward(const int dc[4], int qPc, int lvl[4])
{
    const int a = dc[0], b = dc[1], c = dc[2], d = dc[3];
    const int f[4] = { (a + b + c + d + 2) >> 2, (a - b + c - d + 2) >> 2, (a + b - c - d + 2) >> 2, (a - b - c + d + 2) >> 2 };
    for (int i = 0; i < 4; i++) lvl[i] = ((((f[i] * 32) >> (qPc / 6)) * ic_LQ[qPc % 6][0]) + 16384) >> 15;
}
FH_HD void ic_chroma_dc_inverse(const int lvl[4], int qPc, int dc[4])
{
    const int a = lvl[0], b = lvl[1], c = lvl[2], d = lvl[3];
    const int f[4] = { a + b + c + d, a - b + c - d, a + b - c - d, a - b - c + d };
    for (int i = 0; i < 4; i++) dc[i] = ((f[i] * ic_LS[qPc % 6][0]) * (1 << (qPc / 6))) >> 5;
}
// Intra16x16 luma DC: 4x4 Hadamard + quantisation (quantizationTransform.cpp:105-152,227-260); DC[row][col] -> levels in zigzag order
FH_HD void ic_luma_dc_forward(const int DC[16], int qp, int cq[16])
{
    int t[16], u[16];
    for (int j = 0; j < 4; j++) {
        const int g0 = DC[j] + DC[12 + j], g1 = DC[4 + j] + DC[8 + j], g2 = DC[4 + j] - DC[8 + j], g3 = DC[j] - DC[12 + j];
        t[j] = g0 + g1; t[4 + j] = g3 + g2; t[8 + j] = g0 - g1; t[12 + j] = g3 - g2;
    }
    for (int i = 0; i < 4; i++) {
        const int *q = t + 4 * i;
        const int d0 = q[0] + q[3], d1 = q[1] + q[2], d2 = q[1] - q[2], d3 = q[0] - q[3];
        u[4 * i] = (d0 + d1 + 8) >> 4; u[4 * i + 1] = (d3 + d2 + 8) >> 4; u[4 * i + 2] = (d0 - d1 + 8) >> 4; u[4 * i + 3] = (d3 - d2 + 8) >> 4;
    }
    const int per = qp / 6, lq = ic_LQ[qp % 6][0];
    for (int i = 0; i < 16; i++) {
        const int tt = qp >= 36 ? (u[i] >> (per - 6)) * lq : (u[i] * (1 << (6 - per)) - (1 << (5 - per))) * lq;
        cq[i] = (tt + 16384) >> 15;
    }
}
// inverse Hadamard + scaling (scaleTransform.cpp:154-189,344-376): quantised DC (row-major) -> DC coefficient of every block (row-major)
FH_HD void ic_luma_dc_inverse(const int cq[16], int qp, int DC[16])
{
    int t[16], u[16];
    for (int i = 0; i < 4; i++) {
        const int *q = cq + 4 * i;
        const int d0 = q[0] + q[2], d1 = q[0] - q[2], d2 = q[1] - q[3], d3 = q[1] + q[3];
        t[4 * i] = d0 + d3; t[4 * i + 1] = d1 + d2; t[4 * i + 2] = d1 - d2; t[4 * i + 3] = d0 - d3;
    }
    for (int j = 0; j < 4; j++) {
        const int g0 = t[j] + t[8 + j], g1 = t[j] - t[8 + j], g2 = t[4 + j] - t[12 + j], g3 = t[4 + j] + t[12 + j];
        u[j] = g0 + g3; u[4 + j] = g1 + g2; u[8 + j] = g1 - g2; u[12 + j] = g0 - g3;
    }
    const int per = qp / 6, ls = ic_LS[qp % 6][0];
    for (int i = 0; i < 16; i++) DC[i] = qp >= 36 ? (u[i] * ls) * (1 << (per - 6)) : (u[i] * ls + (1 << (5 - per))) >> (6 - per);
}

// ---- Intra4x4 prediction (intra.cpp:143-420) ---------------------------------------------------------------------------------
// p[0] = p[-1,-1], p[1..4] = p[-1,0..3], p[5..12] = p[0..7,-1]; -1 = not available
FH_HD void ic_fetch4(const IcCtx &c, int blk, int p[13])
{
    const int x = c.xP + ic_blkx(blk), y = c.yP + ic_blky(blk);
    p[0] = (x - 1 < 0 || y - 1 < 0) ? -1 : ic_px(c, x - 1, y - 1);
    for (int i = 0; i < 4; i++) p[1 + i] = x - 1 < 0 ? -1 : ic_px(c, x - 1, y + i);
    if (y - 1 < 0) { for (int i = 5; i < 13; i++) p[i] = -1; return; }
    for (int i = 0; i < 4; i++) p[5 + i] = ic_px(c, x + i, y - 1);
    // above-right: replaced by the last sample above when it lies outside the picture, in a macroblock coded later
    // (x0 == 12 below the first block row) or in a block of this macroblock coded later (blocks 3 and 11), intra.cpp:353-375
    const bool edge = (x + 4 >= c.W) || (ic_blkx(blk) == 12 && ic_blky(blk) > 0) || blk == 3 || blk == 11;
    for (int i = 0; i < 4; i++) p[9 + i] = edge ? p[8] : ic_px(c, x + 4 + i, y - 1);
}
#define IC_T(i) p[(i) + 5]      /* p[i, -1], i = 0 .. 7 */
#define IC_L(i) p[(i) + 1]      /* p[-1, i], i = 0 .. 3 */
#define IC_C p[0]               /* p[-1, -1] */
FH_HD int ic_f3(int a, int b, int c) { return (a + 2 * b + c + 2) >> 2; }
FH_HD int ic_f2(int a, int b) { return (a + b + 1) >> 1; }
// p(x, -1) for x = -1 .. 7 and p(-1, y) for y = -1 .. 3 with the corner shared
FH_HD int ic_top(const int p[13], int x) { return x < 0 ? p[0] : p[x + 5]; }
FH_HD int ic_left(const int p[13], int y) { return y < 0 ? p[0] : p[y + 1]; }
FH_HD int ic_pred4_px(int mode, const int p[13], int x, int y)
{
    int v;
    switch (mode) {
    case 0: v = IC_T(x); break;                                                                         // vertical
    case 1: v = IC_L(y); break;                                                                         // horizontal
    case 2:                                                                                            // DC (intra.cpp:168-186: the corner decides "both available")
        if (IC_C != -1) v = (IC_T(0) + IC_T(1) + IC_T(2) + IC_T(3) + IC_L(0) + IC_L(1) + IC_L(2) + IC_L(3) + 4) >> 3;
        else if (IC_L(0) != -1) v = (IC_L(0) + IC_L(1) + IC_L(2) + IC_L(3) + 2) >> 2;
        else if (IC_T(0) != -1) v = (IC_T(0) + IC_T(1) + IC_T(2) + IC_T(3) + 2) >> 2;
        else v = 128;
        break;
    case 3:                                                                                            // diagonal down-left
        v = (x == 3 && y == 3) ? (IC_T(6) + 3 * IC_T(7) + 2) >> 2 : ic_f3(IC_T(x + y), IC_T(x + y + 1), IC_T(x + y + 2));
        break;
    case 4:                                                                                            // diagonal down-right
        if (x > y) v = ic_f3(ic_top(p, x - y - 2), ic_top(p, x - y - 1), ic_top(p, x - y));
        else if (x < y) v = ic_f3(ic_left(p, y - x - 2), ic_left(p, y - x - 1), ic_left(p, y - x));
        else v = ic_f3(IC_T(0), IC_C, IC_L(0));
        break;
    case 5: {                                                                                          // vertical-right
        const int z = 2 * x - y, i = x - (y >> 1);
        if (z >= 0 && !(z & 1)) v = ic_f2(ic_top(p, i - 1), ic_top(p, i));
        else if (z >= 0) v = ic_f3(ic_top(p, i - 2), ic_top(p, i - 1), ic_top(p, i));
        else if (z == -1) v = ic_f3(IC_L(0), IC_C, IC_T(0));
        else v = ic_f3(ic_left(p, y - 1), ic_left(p, y - 2), ic_left(p, y - 3));
        break;
    }
    case 6: {                                                                                          // horizontal-down
        const int z = 2 * y - x, i = y - (x >> 1);
        if (z >= 0 && !(z & 1)) v = ic_f2(ic_left(p, i - 1), ic_left(p, i));
        else if (z >= 0) v = ic_f3(ic_left(p, i - 2), ic_left(p, i - 1), ic_left(p, i));
        else if (z == -1) v = ic_f3(IC_L(0), IC_C, IC_T(0));
        else v = ic_f3(ic_top(p, x - 1), ic_top(p, x - 2), ic_top(p, x - 3));
        break;
    }
    case 7: {                                                                                          // vertical-left
        const int i = x + (y >> 1);
        v = (y & 1) ? ic_f3(IC_T(i), IC_T(i + 1), IC_T(i + 2)) : ic_f2(IC_T(i), IC_T(i + 1));
        break;
    }
    default: {                                                                                         // horizontal-up
        const int z = x + 2 * y, i = y + (x >> 1);
        if (z > 5) v = IC_L(3);
        else if (z == 5) v = (IC_L(2) + 3 * IC_L(3) + 2) >> 2;
        else if (z & 1) v = ic_f3(IC_L(i), IC_L(i + 1), IC_L(i + 2));
        else v = ic_f2(IC_L(i), IC_L(i + 1));
        break;
    }
    }
    return v;
}
FH_HD void ic_pred4(int mode, const int p[13], int o[16])
{
    for (int i = 0; i < 16; i++) o[i] = ic_pred4_px(mode, p, i & 3, i >> 2);
}
// a mode is tried only when the samples it is defined on exist (intra.cpp:1022-1033)
FH_HD bool ic_mode4_allowed(int mode, const int p[13])
{
    switch (mode) {
    case 0: case 3: case 7: return p[5] != -1;
    case 1: case 8: return p[1] != -1;
    case 4: case 5: case 6: return p[0] != -1;
    default: return true;
    }
}

// sum of the absolute QUANTISED transform coefficients of (frame.L - pred) for one 4x4 block (satdLuma4x4, intra.cpp:820-852)
FH_HD int ic_satd4(const IcCtx &c, int blk, const int pred[16])
{
    const int x0 = ic_blkx(blk), y0 = ic_blky(blk);
    int diff[16], r[16];
    for (int i = 0; i < 16; i++) diff[i] = (int)c.L[(y0 + (i >> 2)) * 16 + x0 + (i & 3)] - pred[i];
    ic_forward_residual(diff, r, c.qp, false);
    int s = 0;
    for (int i = 0; i < 16; i++) s += ic_abs(r[i]);
    return s;
}

// ---- Intra16x16 prediction (intra.cpp:422-560): p[0] corner, p[1..16] left column, p[17..32] row above -----------------------
FH_HD void ic_fetch16(const IcCtx &c, int p[33])
{
    const int xP = c.xP, yP = c.yP;
    p[0] = (xP > 0 && yP > 0) ? c.nbT[0] : -1;
    for (int i = 0; i < 16; i++) p[1 + i] = xP > 0 ? c.nbL[i] : -1;
    for (int i = 0; i < 16; i++) p[17 + i] = yP > 0 ? c.nbT[1 + i] : -1;
}
// the prediction of one mode for the samples of one 4x4 block (Intra_16x16_Vertical / Horizontal / DC / Plane, intra.cpp:424-500)
FH_HD void ic_pred16_block(int mode, const int p[33], int blk, int o[16])
{
    const int x0 = ic_blkx(blk), y0 = ic_blky(blk);
    if (mode == 0) { for (int i = 0; i < 16; i++) o[i] = p[17 + x0 + (i & 3)]; return; }
    if (mode == 1) { for (int i = 0; i < 16; i++) o[i] = p[1 + y0 + (i >> 2)]; return; }
    if (mode == 2) {
        int sx = 0, sy = 0;
        for (int i = 0; i < 16; i++) { sx += p[17 + i]; sy += p[1 + i]; }
        int v = 128;
        if (p[0] != -1) v = (sx + sy + 16) >> 5;
        else if (p[1] != -1) v = (sy + 8) >> 4;
        else if (p[17] != -1) v = (sx + 8) >> 4;
        for (int i = 0; i < 16; i++) o[i] = v;
        return;
    }
    int Hh = 0, V = 0;
    for (int i = 0; i <= 7; i++) {
        Hh += (i + 1) * (p[17 + 8 + i] - (6 - i >= 0 ? p[17 + 6 - i] : p[0]));
        V += (i + 1) * (p[1 + 8 + i] - (6 - i >= 0 ? p[1 + 6 - i] : p[0]));
    }
    const int a = (p[16] + p[32]) << 4, b = (5 * Hh + 32) >> 6, cc = (5 * V + 32) >> 6;
    for (int i = 0; i < 16; i++) o[i] = ic_clip255((a + b * (x0 + (i & 3) - 7) + cc * (y0 + (i >> 2) - 7) + 16) >> 5);
}
FH_HD bool ic_mode16_allowed(int mode, const int p[33]) { return mode == 0 ? p[17] != -1 : (mode == 1 ? p[1] != -1 : (mode == 3 ? p[0] != -1 : true)); }

// ---- chroma prediction (intra.cpp:562-790): p[0] corner, p[1..8] left column, p[9..16] row above -----------------------------
FH_HD void ic_pred_chroma(const IcCtx &c, int comp, int mode, uint8_t o[64])
{
    const int xM = c.xP >> 1, yM = c.yP >> 1;
    int p[17];
    p[0] = (xM > 0 && yM > 0) ? c.nbC[comp][0] : -1;
    for (int i = 0; i < 8; i++) p[1 + i] = xM > 0 ? c.nbC[comp][1 + i] : -1;
    for (int i = 0; i < 8; i++) p[9 + i] = yM > 0 ? c.nbC[comp][9 + i] : -1;
    if (mode == 1) { for (int i = 0; i < 64; i++) o[i] = (uint8_t)p[1 + (i >> 3)]; return; }
    if (mode == 2) { for (int i = 0; i < 64; i++) o[i] = (uint8_t)p[9 + (i & 7)]; return; }
    if (mode == 0) {
        for (int b = 0; b < 4; b++) {
            const int x0 = (b & 1) << 2, y0 = (b >> 1) << 2;
            int sx = 0, sy = 0;
            for (int i = 0; i < 4; i++) { sx += p[9 + x0 + i]; sy += p[1 + y0 + i]; }
            const bool la = p[1 + y0] != -1, ta = p[9 + x0] != -1;
            int v = 128;
            if (x0 == y0) { if (la && ta) v = (sx + sy + 4) >> 3; else if (la) v = (sy + 2) >> 2; else if (ta) v = (sx + 2) >> 2; }
            else if (x0 > 0) { if (ta) v = (sx + 2) >> 2; else if (la) v = (sy + 2) >> 2; }
            else { if (la) v = (sy + 2) >> 2; else if (ta) v = (sx + 2) >> 2; }
            for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) o[(y0 + y) * 8 + x0 + x] = (uint8_t)v;
        }
        return;
    }
    int Hh = 0, V = 0;
    for (int i = 0; i <= 3; i++) {
        Hh += (i + 1) * (p[9 + 4 + i] - (2 - i >= 0 ? p[9 + 2 - i] : p[0]));
        V += (i + 1) * (p[1 + 4 + i] - (2 - i >= 0 ? p[1 + 2 - i] : p[0]));
    }
    const int a = (p[8] + p[16]) << 4, b = (34 * Hh + 32) >> 6, cc = (34 * V + 32) >> 6;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) o[y * 8 + x] = (uint8_t)ic_clip255((a + b * (x - 3) + cc * (y - 3) + 16) >> 5);
}

// ---- transform / quantisation of the macroblock, one 4x4 block per call so that the blocks can be spread over the lanes ---------
// unit v of the forward pass: v < 16 luma block v as Intra16x16 (quantizationTransform.cpp:381-412), v = 16 .. 23 chroma block
// (v - 16) & 3 of component (v - 16) >> 2 (:424-462). DC coefficients are left unquantised for the DC transforms.
FH_HD void ic_forward_unit(IcCtx &c, int v)
{
    int diff[16], r[16];
    if (v < 16) {
        const int x0 = ic_blkx(v), y0 = ic_blky(v);
        for (int i = 0; i < 16; i++) { const int o = (y0 + (i >> 2)) * 16 + x0 + (i & 3); diff[i] = (int)c.L[o] - (int)c.pred16[o]; }
        ic_forward_residual(diff, r, c.qp, true);
        c.DC[(y0 >> 2) * 4 + (x0 >> 2)] = r[0];
        for (int k = 1; k < 16; k++) c.lv.ac16[v][k - 1] = (int16_t)r[ic_ZZ[k]];
    } else {
        const int comp = (v - 16) >> 2, b = (v - 16) & 3, x0 = (b & 1) * 4, y0 = (b >> 1) * 4;
        for (int i = 0; i < 16; i++) { const int o = (y0 + (i >> 2)) * 8 + x0 + (i & 3); diff[i] = (int)c.SC[comp][o] - (int)c.predC[comp][o]; }
        ic_forward_residual(diff, r, c.qpc, true);
        c.cdcraw[comp][b] = r[0];
        for (int k = 1; k < 16; k++) c.lv.cac[comp][b][k - 1] = (int16_t)r[ic_ZZ[k]];
    }
}
// the DC transforms that follow (forwardDCLumaIntra :393-397, forwardDCChroma :464-478)
FH_HD void ic_forward_dcs(IcCtx &c)
{
    ic_luma_dc_forward(c.DC, c.qp, c.cq16);
    for (int k = 0; k < 16; k++) c.lv.dc16[k] = (int16_t)c.cq16[ic_ZZ[k]];
    for (int comp = 0; comp < 2; comp++) {
        int dl[4];
        ic_chroma_dc_forward(c.cdcraw[comp], c.qpc, dl);
        for (int i = 0; i < 4; i++) c.lv.cdc[comp][i] = (int16_t)dl[i];
    }
}
// reconstruction: DCs first (scaleTransform.cpp:154-189,344-376 luma; :247-262,408-420 chroma) ...
FH_HD void ic_inverse_dcs(IcCtx &c, bool luma16)
{
    if (luma16) ic_luma_dc_inverse(c.cq16, c.qp, c.rdc);
    for (int comp = 0; comp < 2; comp++) {
        int dl[4];
        for (int i = 0; i < 4; i++) dl[i] = c.lv.cdc[comp][i];
        ic_chroma_dc_inverse(dl, c.qpc, c.crdc[comp]);
    }
}
// ... then unit v: v < 16 luma block of an Intra16x16 macroblock (transformDecodingIntra_16x16Luma, inttransform.cpp:157-208) into L,
// v = 16 .. 23 chroma block (transformDecodingChroma, inttransform.cpp:237-321) into RC
FH_HD void ic_inverse_unit(IcCtx &c, int v)
{
    int cf[16], d[16], rr[16];
    if (v < 16) {
        const int x0 = ic_blkx(v), y0 = ic_blky(v);
        cf[0] = c.rdc[(y0 >> 2) * 4 + (x0 >> 2)];
        for (int k = 1; k < 16; k++) cf[ic_ZZ[k]] = c.lv.ac16[v][k - 1];
        ic_dequant4x4(cf, d, c.qp, true);
        ic_inverse4x4(d, rr);
        for (int i = 0; i < 16; i++) { const int o = (y0 + (i >> 2)) * 16 + x0 + (i & 3); c.L[o] = (uint8_t)ic_clip255((int)c.pred16[o] + rr[i]); }
    } else {
        const int comp = (v - 16) >> 2, b = (v - 16) & 3, x0 = (b & 1) * 4, y0 = (b >> 1) * 4;
        cf[0] = c.crdc[comp][b];
        for (int k = 1; k < 16; k++) cf[ic_ZZ[k]] = c.lv.cac[comp][b][k - 1];
        ic_dequant4x4(cf, d, c.qpc, true);
        ic_inverse4x4(d, rr);
        for (int i = 0; i < 16; i++) { const int o = (y0 + (i >> 2)) * 8 + x0 + (i & 3); c.RC[comp][o] = (uint8_t)ic_clip255((int)c.predC[comp][o] + rr[i]); }
    }
}

// setCodedBlockPattern (rbsp_encoding.cpp:21-105)
FH_HD void ic_cbp(bool is16, const IcLevels &lv, int &cbpl, int &cbpc)
{
    cbpl = 0;
    for (int i8 = 0; i8 < 4; i8++) {
        int any = 0;
        for (int i4 = 0; i4 < 4; i4++) any |= is16 ? cv_count(lv.ac16[i8 * 4 + i4], 15) : cv_count(lv.luma[i8 * 4 + i4], 16);
        if (any) cbpl |= 1 << i8;
    }
    if (is16 && cbpl) cbpl = 15;
    cbpc = 0;
    for (int i = 0; i < 4; i++) if (lv.cdc[0][i] != 0 || lv.cdc[1][i] != 0) cbpc = 1;
    for (int i4 = 0; i4 < 4; i4++) if (cv_count(lv.cac[0][i4], 15) || cv_count(lv.cac[1][i4], 15)) cbpc = 2;
}

// coded_mb_size (rbsp_encoding.cpp:330-487) of an I macroblock, the residual blocks spread over the lanes: the size of a block
// depends on its neighbours only through their TotalCoeff (nC), and those are plain non-zero counts, so they are taken first.
// self_skip: mb_type_array[CurrMbAddr] == P_Skip while the trial runs — during the Intra16x16 trial that entry still holds the
// PREVIOUS picture's type (intra.cpp:1012 clears it only afterwards), and residual.cpp:473,493 then take every neighbour block
// inside this macroblock as empty. tcl / tcc (shared by the lanes) receive the TotalCoeff of the blocks the trial codes, 0 elsewhere.
FH_HD int ic_mb_bits(const IcLevels &lv, bool is16, int mb_type, int chroma_mode, const uint8_t prev_flag[16], int cbpl, int cbpc, bool self_skip,
                     const IcInfo *left, const IcInfo *up, uint8_t tcl[16], uint8_t tcc[2][4], int lane, int nl)
{
    for (int v = lane; v < 24; v += nl) {
        if (v < 16) tcl[v] = (uint8_t)((cbpl >> (v >> 2)) & 1 ? (is16 ? cv_count(lv.ac16[v], 15) : cv_count(lv.luma[v], 16)) : 0);
        else tcc[(v - 16) >> 2][(v - 16) & 3] = (uint8_t)((cbpc & 2) ? cv_count(lv.cac[(v - 16) >> 2][(v - 16) & 3], 15) : 0);
    }
    ic_sync(nl);
    int head = ic_ue_len(mb_type);
    if (!is16) for (int k = 0; k < 16; k++) head += prev_flag[k] ? 1 : 4;
    head += ic_ue_len(chroma_mode);
    if (!is16) head += ic_ue_len(ic_cbp_intra[(cbpc << 4) | cbpl]);
    if (!(cbpl > 0 || cbpc > 0 || is16)) return head;
    head += 1;                                                // mb_qp_delta
    CvBits b;
    cv_init(b, nullptr, 0);                                   // counts only
    int bad = 0;
    for (int v = lane; v < 27; v += nl) {
        if (v == 0) {                                         // Intra16x16DCLevel: neighbours of block 0
            if (is16) cv_block(b, lv.dc16, 16, cv_nc(left ? left->tc_luma[5] : -1, up ? up->tc_luma[10] : -1), &bad);
        } else if (v <= 16) {
            const int blk = v - 1;
            if (!(cbpl & (1 << (blk >> 2)))) continue;
            const int bx = ((blk >> 2) & 1) * 2 + (blk & 1), by = (blk >> 3) * 2 + ((blk >> 1) & 1);
            int nA, nB;
            if (bx > 0) { const int a = (by >> 1) * 8 + ((bx - 1) >> 1) * 4 + (by & 1) * 2 + ((bx - 1) & 1); nA = self_skip ? 0 : tcl[a]; }
            else { const int a = (by >> 1) * 8 + 4 + (by & 1) * 2 + 1; nA = left ? left->tc_luma[a] : -1; }
            if (by > 0) { const int a = ((by - 1) >> 1) * 8 + (bx >> 1) * 4 + ((by - 1) & 1) * 2 + (bx & 1); nB = self_skip ? 0 : tcl[a]; }
            else { const int a = 8 + (bx >> 1) * 4 + 2 + (bx & 1); nB = up ? up->tc_luma[a] : -1; }
            if (is16) cv_block(b, lv.ac16[blk], 15, cv_nc(nA, nB), &bad); else cv_block(b, lv.luma[blk], 16, cv_nc(nA, nB), &bad);
        } else if (v <= 18) {
            if (cbpc & 3) cv_block(b, lv.cdc[v - 17], 4, -1, &bad);
        } else {
            if (!(cbpc & 2)) continue;
            const int c = (v - 19) >> 2, blk = (v - 19) & 3, bx = blk & 1, by = blk >> 1;
            const int nA = bx ? (self_skip ? 0 : tcc[c][blk - 1]) : (left ? left->tc_chroma[c][blk + 1] : -1);
            const int nB = by ? (self_skip ? 0 : tcc[c][blk - 2]) : (up ? up->tc_chroma[c][blk + 2] : -1);
            cv_block(b, lv.cac[c][blk], 15, cv_nc(nA, nB), &bad);
        }
    }
    return head + ic_red_add(cv_bits(b), nl);
}

// predIntra4x4PredMode of block blk (setIntra4x4PredMode, intra.cpp:877-941): min of the neighbours' modes, DC when a
// neighbour macroblock is missing or not Intra4x4. mine = the modes of this macroblock.
FH_HD int ic_pred_mode_of(int blk, const uint8_t mine[16], const IcInfo *left, const IcInfo *up)
{
    const int bx = ((blk >> 2) & 1) * 2 + (blk & 1), by = (blk >> 3) * 2 + ((blk >> 1) & 1);
    int mA, mB;
    if (bx > 0) mA = mine[(by >> 1) * 8 + ((bx - 1) >> 1) * 4 + (by & 1) * 2 + ((bx - 1) & 1)];
    else { if (!left) return 2; mA = left->is4x4 ? left->mode4[(by >> 1) * 8 + 4 + (by & 1) * 2 + 1] : 2; }
    if (by > 0) mB = mine[((by - 1) >> 1) * 8 + (bx >> 1) * 4 + ((by - 1) & 1) * 2 + (bx & 1)];
    else { if (!up) return 2; mB = up->is4x4 ? up->mode4[8 + (bx >> 1) * 4 + 2 + (bx & 1)] : 2; }
    return mA <= mB ? mA : mB;
}

// ---- the macroblock ------------------------------------------------------------------------------------------------------------
// c: picture pointers, W, H, xP, yP, qp set by the caller (shared by the lanes). prev_skip: this macroblock was P_Skip in the
// previous picture. left / up: state of the neighbouring macroblocks of THIS picture (null outside the picture); the macroblock
// above-right must be complete as well (its reconstruction feeds the Intra4x4 above-right samples). Writes the reconstruction
// into c.rec. Called by lanes 0 .. nl-1 of a warp (nl = 32), or by one lane with nl = 1. The blocks of the Intra4x4 coding
// depend on each other and are taken one after the other (16 lanes per block); the small DC transforms run on lane 0 alone.
FH_HD void ic_macroblock(IcCtx &c, bool prev_skip, const IcInfo *left, const IcInfo *up, fh264_mb_result_i &out, IcInfo &info, int lane, int nl)
{
    const int W = c.W, CW = c.W >> 1, xP = c.xP, yP = c.yP;
    if (lane == 0) c.qpc = ic_QPC[c.qp < 0 ? 0 : (c.qp > 51 ? 51 : c.qp)];
    for (int i = lane; i < 256; i += nl) c.S[i] = c.L[i] = c.src[0][(size_t)(yP + (i >> 4)) * W + xP + (i & 15)];
    for (int i = lane; i < 128; i += nl) c.SC[i >> 6][i & 63] = c.src[1 + (i >> 6)][(size_t)((yP >> 1) + ((i & 63) >> 3)) * CW + (xP >> 1) + (i & 7)];
    ic_stage_neighbours(c, lane, nl);
    ic_sync(nl);

    // Intra16x16 mode search (intra.cpp:980-1001): smallest sum of absolute quantised coefficients, first mode wins ties.
    // 4 modes x 16 blocks spread over the lanes.
    int p16[33];
    ic_fetch16(c, p16);
    int part[4] = { 0, 0, 0, 0 };
    for (int item = lane; item < 64; item += nl) {
        const int m = item >> 4, b = item & 15;
        if (!ic_mode16_allowed(m, p16)) continue;
        int pb[16];
        ic_pred16_block(m, p16, b, pb);
        part[m] += ic_satd4(c, b, pb);
    }
    int mode16 = 2, min16 = 0x7fffffff;
    for (int m = 0; m < 4; m++) {
        const int satd = ic_red_add(part[m], nl);
        if (ic_mode16_allowed(m, p16) && satd < min16) { min16 = satd; mode16 = m; }
    }
    const int chroma_mode = ic_chroma_of_16[mode16];
    // Intra4x4 mode search on the macroblock as it stands: neighbours inside it are still SOURCE samples (intra.cpp:1011-1049).
    // 16 blocks x 9 modes spread over the lanes; the smallest (cost, mode) pair is the first mode reaching the minimum.
    int best[16];
    for (int b = 0; b < 16; b++) best[b] = 0x7fffffff;
    for (int item = lane; item < 144; item += nl) {
        const int blk = item & 15, m = item >> 4;
        int p[13], pb[16];
        ic_fetch4(c, blk, p);
        if (!ic_mode4_allowed(m, p)) continue;
        ic_pred4(m, p, pb);
        const int key = (ic_satd4(c, blk, pb) << 4) | m;
        if (key < best[blk]) best[blk] = key;
    }
    for (int b = 0; b < 16; b++) best[b] = ic_red_min(best[b], nl);

    // predictions of the chosen Intra16x16 / chroma modes, then the first trial: the macroblock as Intra16x16 (intra.cpp:1003-1008)
    for (int v = lane; v < 18; v += nl) {
        if (v < 16) {
            int pb[16];
            ic_pred16_block(mode16, p16, v, pb);
            const int x0 = ic_blkx(v), y0 = ic_blky(v);
            for (int i = 0; i < 16; i++) c.pred16[(y0 + (i >> 2)) * 16 + x0 + (i & 3)] = (uint8_t)pb[i];
        } else ic_pred_chroma(c, v - 16, chroma_mode, c.predC[v - 16]);
    }
    ic_sync(nl);
    for (int v = lane; v < 24; v += nl) ic_forward_unit(c, v);
    ic_sync(nl);
    if (lane == 0) {
        ic_forward_dcs(c);
        ic_cbp(true, c.lv, c.cbpl16, c.cbpc);
        c.type16 = mode16 + 1 + (c.cbpc << 2) + (c.cbpl16 == 15 ? 12 : 0);
    }
    ic_sync(nl);
    const int bits16 = ic_mb_bits(c.lv, true, c.type16, chroma_mode, nullptr, c.cbpl16, c.cbpc, prev_skip, left, up, c.tcl[0], c.tcc[0], lane, nl);

    // code the blocks one by one with the modes found; each reconstruction feeds the next prediction (intra.cpp:1063-1086).
    // The blocks are serial, the 16 samples / coefficients of a block are not: one lane each, the 4x4 butterflies one row or
    // column per lane, the stages separated by warp barriers.
    for (int blk = lane; blk < 16; blk += nl) c.mode4[blk] = (uint8_t)(best[blk] & 15);
    ic_sync(nl);
    for (int blk = lane; blk < 16; blk += nl) {
        const int pm = ic_pred_mode_of(blk, c.mode4, left, up);
        c.flag[blk] = c.mode4[blk] == pm;
        c.rem[blk] = (uint8_t)(c.mode4[blk] < pm ? c.mode4[blk] : c.mode4[blk] - 1);
    }
    for (int blk = 0; blk < 16; blk++) {
        const int x0 = ic_blkx(blk), y0 = ic_blky(blk), mode = c.mode4[blk];
        int p[13];
        ic_fetch4(c, blk, p);
        for (int i = lane; i < 16; i += nl) {
            const int pv = ic_pred4_px(mode, p, i & 3, i >> 2), r = (int)c.L[(y0 + (i >> 2)) * 16 + x0 + (i & 3)] - pv;
            c.pb[i] = pv;
            c.w0[i] = r == 0 ? 0 : r * 64 - 32;                                   // forwardTransform4x4's input scaling
        }
        ic_sync(nl);
        for (int j = lane; j < 4; j += nl) ic_fwd4(c.w0[j], c.w0[4 + j], c.w0[8 + j], c.w0[12 + j], c.w1[j], c.w1[4 + j], c.w1[8 + j], c.w1[12 + j]);
        ic_sync(nl);
        for (int i = lane; i < 4; i += nl) ic_fwd4(c.w1[4 * i], c.w1[4 * i + 1], c.w1[4 * i + 2], c.w1[4 * i + 3], c.w0[4 * i], c.w0[4 * i + 1], c.w0[4 * i + 2], c.w0[4 * i + 3]);
        ic_sync(nl);
        for (int i = lane; i < 16; i += nl) {
            const int q = ic_quant1(c.w0[i], i, c.qp);
            c.lv.luma[blk][ic_ZZI[i]] = (int16_t)q;
            c.w1[i] = ic_dequant1(q, i, c.qp);
        }
        ic_sync(nl);
        for (int i = lane; i < 4; i += nl) {                                      // inverseTransform4x4: rows, then columns
            const int *q = c.w1 + 4 * i;
            const int e0 = q[0] + q[2], e1 = q[0] - q[2], e2 = (q[1] >> 1) - q[3], e3 = q[1] + (q[3] >> 1);
            c.w0[4 * i] = e0 + e3; c.w0[4 * i + 1] = e1 + e2; c.w0[4 * i + 2] = e1 - e2; c.w0[4 * i + 3] = e0 - e3;
        }
        ic_sync(nl);
        for (int j = lane; j < 4; j += nl) {
            const int g0 = c.w0[j] + c.w0[8 + j], g1 = c.w0[j] - c.w0[8 + j], g2 = (c.w0[4 + j] >> 1) - c.w0[12 + j], g3 = c.w0[4 + j] + (c.w0[12 + j] >> 1);
            const int h[4] = { g0 + g3, g1 + g2, g1 - g2, g0 - g3 };
            for (int y = 0; y < 4; y++) c.L[(y0 + y) * 16 + x0 + j] = (uint8_t)ic_clip255(c.pb[4 * y + j] + ((h[y] + 32) >> 6));
        }
        ic_sync(nl);
    }
    if (lane == 0) {
        int cbpc4;
        ic_cbp(false, c.lv, c.cbpl4, cbpc4);            // the chroma levels are the first trial's: same CodedBlockPatternChroma
    }
    ic_sync(nl);
    // second trial: Intra4x4 (intra.cpp:1088), then the decision
    const int bits4 = ic_mb_bits(c.lv, false, 0, chroma_mode, c.flag, c.cbpl4, c.cbpc, false, left, up, c.tcl[1], c.tcc[1], lane, nl);
    const bool use4 = bits4 < bits16;
    if (!use4) for (int i = lane; i < 256; i += nl) c.L[i] = c.S[i];       // restore the source, code as Intra16x16 (intra.cpp:1095-1106)
    if (lane == 0) ic_inverse_dcs(c, !use4);
    ic_sync(nl);
    for (int v = lane + (use4 ? 16 : 0); v < 24; v += nl) ic_inverse_unit(c, v);
    ic_sync(nl);

    // results
    int16_t *ol = &out.luma[0][0];
    if (use4) for (int i = lane; i < 256; i += nl) ol[i] = c.lv.luma[i >> 4][i & 15];
    else for (int i = lane; i < 256; i += nl) ol[i] = i < 16 ? c.lv.dc16[i] : c.lv.ac16[(i - 16) / 15][(i - 16) % 15];
    for (int i = lane; i < 128; i += nl) {
        if (i < 8) out.chroma_dc[i >> 2][i & 3] = c.lv.cdc[i >> 2][i & 3];
        else out.chroma_ac[(i - 8) / 60][((i - 8) % 60) / 15][(i - 8) % 15] = c.lv.cac[(i - 8) / 60][((i - 8) % 60) / 15][(i - 8) % 15];
    }
    if (lane == 0) {
        out.mb_type = (int16_t)(use4 ? 0 : c.type16);
        out.intra16x16_pred_mode = (int8_t)(use4 ? -1 : mode16);
        out.intra_chroma_pred_mode = (uint8_t)chroma_mode;
        out.cbp_luma = (uint8_t)(use4 ? c.cbpl4 : c.cbpl16);
        out.cbp_chroma = (uint8_t)c.cbpc;
        out.bits_intra16x16 = (uint16_t)bits16;
        out.bits_intra4x4 = (uint16_t)bits4;
        for (int i = 0; i < 16; i++) { out.intra4x4_pred_mode[i] = c.mode4[i]; out.prev_intra4x4_pred_mode_flag[i] = c.flag[i]; out.rem_intra4x4_pred_mode[i] = c.rem[i]; }
        out.reserved[0] = out.reserved[1] = out.reserved[2] = 0;
        info.mb_type = (uint8_t)out.mb_type; info.cbp_luma = out.cbp_luma; info.cbp_chroma = out.cbp_chroma; info.is4x4 = use4;
        const int t = use4 ? 1 : 0;
        for (int i = 0; i < 16; i++) { info.tc_luma[i] = c.tcl[t][i]; info.mode4[i] = c.mode4[i]; }
        for (int i = 0; i < 8; i++) info.tc_chroma[i >> 2][i & 3] = c.tcc[t][i >> 2][i & 3];
        info.pad[0] = info.pad[1] = info.pad[2] = info.pad[3] = 0;
    }
    for (int i = lane; i < 256; i += nl) c.rec[0][(size_t)(yP + (i >> 4)) * W + xP + (i & 15)] = c.L[i];
    for (int i = lane; i < 128; i += nl) c.rec[1 + (i >> 6)][(size_t)((yP >> 1) + ((i & 63) >> 3)) * CW + (xP >> 1) + (i & 7)] = c.RC[i >> 6][i & 63];
}

// ---- macroblock_layer() of one I macroblock as the reference writes it (rbsp_encoding.cpp:221-305, residual_write
//      residual.cpp:300-372): mb_type, the Intra4x4 mode syntax, intra_chroma_pred_mode, coded_block_pattern, mb_qp_delta and
//      the residual blocks — the same walk as the bit-cost trial above, with the bits kept --------------------------------------
// state of a coded macroblock from its record (what ic_macroblock leaves in IcInfo)
FH_HD void ic_info_from_record(const fh264_mb_result_i &r, IcInfo &o)
{
    const bool is16 = r.intra16x16_pred_mode >= 0;
    const int16_t *l = &r.luma[0][0];
    o.mb_type = (uint8_t)r.mb_type; o.cbp_luma = r.cbp_luma; o.cbp_chroma = r.cbp_chroma; o.is4x4 = !is16;
    for (int b = 0; b < 16; b++) {
        o.tc_luma[b] = (uint8_t)((r.cbp_luma >> (b >> 2)) & 1 ? (is16 ? cv_count(l + 16 + b * 15, 15) : cv_count(l + b * 16, 16)) : 0);
        o.mode4[b] = r.intra4x4_pred_mode[b];
    }
    for (int c = 0; c < 2; c++) for (int b = 0; b < 4; b++) o.tc_chroma[c][b] = (uint8_t)((r.cbp_chroma & 2) ? cv_count(r.chroma_ac[c][b], 15) : 0);
    o.pad[0] = o.pad[1] = o.pad[2] = o.pad[3] = 0;
}
FH_HD void ic_write_macroblock(CvBits &b, const fh264_mb_result_i &r, const IcInfo &me, const IcInfo *left, const IcInfo *up, int *bad)
{
    const bool is16 = r.intra16x16_pred_mode >= 0;
    const int16_t *l = &r.luma[0][0];
    const int cbpl = me.cbp_luma, cbpc = me.cbp_chroma;
    cv_ue(b, (uint32_t)r.mb_type);
    if (!is16)
        for (int k = 0; k < 16; k++) {
            cv_put(b, 1, r.prev_intra4x4_pred_mode_flag[k] ? 1u : 0u);
            if (!r.prev_intra4x4_pred_mode_flag[k]) cv_put(b, 3, r.rem_intra4x4_pred_mode[k]);
        }
    cv_ue(b, r.intra_chroma_pred_mode);
    if (!is16) cv_ue(b, ic_cbp_intra[(cbpc << 4) | cbpl]);
    if (!(cbpl > 0 || cbpc > 0 || is16)) return;
    cv_se(b, 0);                                              // mb_qp_delta
    if (is16) cv_block(b, l, 16, cv_nc(left ? left->tc_luma[5] : -1, up ? up->tc_luma[10] : -1), bad);
    for (int blk = 0; blk < 16; blk++) {
        if (!(cbpl & (1 << (blk >> 2)))) continue;
        const int bx = ((blk >> 2) & 1) * 2 + (blk & 1), by = (blk >> 3) * 2 + ((blk >> 1) & 1);
        int nA, nB;
        if (bx > 0) nA = me.tc_luma[(by >> 1) * 8 + ((bx - 1) >> 1) * 4 + (by & 1) * 2 + ((bx - 1) & 1)];
        else nA = left ? left->tc_luma[(by >> 1) * 8 + 4 + (by & 1) * 2 + 1] : -1;
        if (by > 0) nB = me.tc_luma[((by - 1) >> 1) * 8 + (bx >> 1) * 4 + ((by - 1) & 1) * 2 + (bx & 1)];
        else nB = up ? up->tc_luma[8 + (bx >> 1) * 4 + 2 + (bx & 1)] : -1;
        if (is16) cv_block(b, l + 16 + blk * 15, 15, cv_nc(nA, nB), bad); else cv_block(b, l + blk * 16, 16, cv_nc(nA, nB), bad);
    }
    if (cbpc & 3) for (int c = 0; c < 2; c++) cv_block(b, r.chroma_dc[c], 4, -1, bad);
    if (cbpc & 2)
        for (int c = 0; c < 2; c++)
            for (int blk = 0; blk < 4; blk++) {
                const int bx = blk & 1, by = blk >> 1;
                const int nA = bx ? me.tc_chroma[c][blk - 1] : (left ? left->tc_chroma[c][blk + 1] : -1);
                const int nB = by ? me.tc_chroma[c][blk - 2] : (up ? up->tc_chroma[c][blk + 2] : -1);
                cv_block(b, r.chroma_ac[c][blk], 15, cv_nc(nA, nB), bad);
            }
}
