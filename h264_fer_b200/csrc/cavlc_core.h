// CAVLC macroblock-layer coder of a P slice (H.264 clauses 7.3.5, 9.1, 9.2), restating what the reference's host code
// emits per macroblock: the P-slice branch of RBSP_encode (rbsp_encoding.cpp:175-305), setCodedBlockPattern (:21-105),
// residual_write / residual_luma_write / residual_block_cavlc_write (residual.cpp:300-666) and the ue/se writers
// (expgolomb.cpp:80-106). SURVEY.md §8(f) rank 1.
//
// The core is plain C++ that compiles for the device (cavlc.cuh, the product path) and for the host (tests only: the same
// source is checked against the reference's slice data without a GPU, tests/cavlc_host.cpp).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FH_HD __device__ __forceinline__
#define FH_TAB static __device__ const
#else
#define FH_HD static inline
#define FH_TAB static const
#endif

// ---- tables (ITU-T H.264 Tables 9-4, 9-5, 9-7, 9-9, 9-10); [len | code] --------------------------------------------------
// coeff_token: [table][TotalCoeff][TrailingOnes]; tables: 0 <= nC < 2, 2 <= nC < 4, 4 <= nC < 8, 8 <= nC, nC == -1 (chroma DC)
FH_TAB uint8_t cv_ct_len[5][17][4] = {
    { { 1, 0, 0, 0 }, { 6, 2, 0, 0 }, { 8, 6, 3, 0 }, { 9, 8, 7, 5 }, { 10, 9, 8, 6 }, { 11, 10, 9, 7 }, { 13, 11, 10, 8 }, { 13, 13, 11, 9 }, { 13, 13, 13, 10 },
      { 14, 14, 13, 11 }, { 14, 14, 14, 13 }, { 15, 15, 14, 14 }, { 15, 15, 15, 14 }, { 16, 15, 15, 15 }, { 16, 16, 16, 15 }, { 16, 16, 16, 16 }, { 16, 16, 16, 16 } },
    { { 2, 0, 0, 0 }, { 6, 2, 0, 0 }, { 6, 5, 3, 0 }, { 7, 6, 6, 4 }, { 8, 6, 6, 4 }, { 8, 7, 7, 5 }, { 9, 8, 8, 6 }, { 11, 9, 9, 6 }, { 11, 11, 11, 7 },
      { 12, 11, 11, 9 }, { 12, 12, 12, 11 }, { 12, 12, 12, 11 }, { 13, 13, 13, 12 }, { 13, 13, 13, 13 }, { 13, 14, 13, 13 }, { 14, 14, 14, 13 }, { 14, 14, 14, 14 } },
    { { 4, 0, 0, 0 }, { 6, 4, 0, 0 }, { 6, 5, 4, 0 }, { 6, 5, 5, 4 }, { 7, 5, 5, 4 }, { 7, 5, 5, 4 }, { 7, 6, 6, 4 }, { 7, 6, 6, 4 }, { 8, 7, 7, 5 },
      { 8, 8, 7, 6 }, { 9, 8, 8, 7 }, { 9, 9, 8, 8 }, { 9, 9, 9, 8 }, { 10, 9, 9, 9 }, { 10, 10, 10, 10 }, { 10, 10, 10, 10 }, { 10, 10, 10, 10 } },
    { { 6, 0, 0, 0 }, { 6, 6, 0, 0 }, { 6, 6, 6, 0 }, { 6, 6, 6, 6 }, { 6, 6, 6, 6 }, { 6, 6, 6, 6 }, { 6, 6, 6, 6 }, { 6, 6, 6, 6 }, { 6, 6, 6, 6 },
      { 6, 6, 6, 6 }, { 6, 6, 6, 6 }, { 6, 6, 6, 6 }, { 6, 6, 6, 6 }, { 6, 6, 6, 6 }, { 6, 6, 6, 6 }, { 6, 6, 6, 6 }, { 6, 6, 6, 6 } },
    { { 2, 0, 0, 0 }, { 6, 1, 0, 0 }, { 6, 6, 3, 0 }, { 6, 7, 7, 6 }, { 6, 8, 8, 7 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 },
      { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 } },
};
FH_TAB uint8_t cv_ct_code[5][17][4] = {
    { { 1, 0, 0, 0 }, { 5, 1, 0, 0 }, { 7, 4, 1, 0 }, { 7, 6, 5, 3 }, { 7, 6, 5, 3 }, { 7, 6, 5, 4 }, { 15, 6, 5, 4 }, { 11, 14, 5, 4 }, { 8, 10, 13, 4 },
      { 15, 14, 9, 4 }, { 11, 10, 13, 12 }, { 15, 14, 9, 12 }, { 11, 10, 13, 8 }, { 15, 1, 9, 12 }, { 11, 14, 13, 8 }, { 7, 10, 9, 12 }, { 4, 6, 5, 8 } },
    { { 3, 0, 0, 0 }, { 11, 2, 0, 0 }, { 7, 7, 3, 0 }, { 7, 10, 9, 5 }, { 7, 6, 5, 4 }, { 4, 6, 5, 6 }, { 7, 6, 5, 8 }, { 15, 6, 5, 4 }, { 11, 14, 13, 4 },
      { 15, 10, 9, 4 }, { 11, 14, 13, 12 }, { 8, 10, 9, 8 }, { 15, 14, 13, 12 }, { 11, 10, 9, 12 }, { 7, 11, 6, 8 }, { 9, 8, 10, 1 }, { 7, 6, 5, 4 } },
    { { 15, 0, 0, 0 }, { 15, 14, 0, 0 }, { 11, 15, 13, 0 }, { 8, 12, 14, 12 }, { 15, 10, 11, 11 }, { 11, 8, 9, 10 }, { 9, 14, 13, 9 }, { 8, 10, 9, 8 }, { 15, 14, 13, 13 },
      { 11, 14, 10, 12 }, { 15, 10, 13, 12 }, { 11, 14, 9, 12 }, { 8, 10, 13, 8 }, { 13, 7, 9, 12 }, { 9, 12, 11, 10 }, { 5, 8, 7, 6 }, { 1, 4, 3, 2 } },
    { { 3, 0, 0, 0 }, { 0, 1, 0, 0 }, { 4, 5, 6, 0 }, { 8, 9, 10, 11 }, { 12, 13, 14, 15 }, { 16, 17, 18, 19 }, { 20, 21, 22, 23 }, { 24, 25, 26, 27 }, { 28, 29, 30, 31 },
      { 32, 33, 34, 35 }, { 36, 37, 38, 39 }, { 40, 41, 42, 43 }, { 44, 45, 46, 47 }, { 48, 49, 50, 51 }, { 52, 53, 54, 55 }, { 56, 57, 58, 59 }, { 60, 61, 62, 63 } },
    { { 1, 0, 0, 0 }, { 7, 1, 0, 0 }, { 4, 6, 1, 0 }, { 3, 3, 2, 5 }, { 2, 3, 2, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 },
      { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 } },
};
// total_zeros for 4x4 blocks: [TotalCoeff - 1][total_zeros]
FH_TAB uint8_t cv_tz_len[15][16] = {
    { 1, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 9 }, { 3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 6, 6, 6, 6, 0 }, { 4, 3, 3, 3, 4, 4, 3, 3, 4, 5, 5, 6, 5, 6, 0, 0 },
    { 5, 3, 4, 4, 3, 3, 3, 4, 3, 4, 5, 5, 5, 0, 0, 0 }, { 4, 4, 4, 3, 3, 3, 3, 3, 4, 5, 4, 5, 0, 0, 0, 0 }, { 6, 5, 3, 3, 3, 3, 3, 3, 4, 3, 6, 0, 0, 0, 0, 0 },
    { 6, 5, 3, 3, 3, 2, 3, 4, 3, 6, 0, 0, 0, 0, 0, 0 }, { 6, 4, 5, 3, 2, 2, 3, 3, 6, 0, 0, 0, 0, 0, 0, 0 }, { 6, 6, 4, 2, 2, 3, 2, 5, 0, 0, 0, 0, 0, 0, 0, 0 },
    { 5, 5, 3, 2, 2, 2, 4, 0, 0, 0, 0, 0, 0, 0, 0, 0 }, { 4, 4, 3, 3, 1, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }, { 4, 4, 2, 1, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 },
    { 3, 3, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }, { 2, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }, { 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 },
};
FH_TAB uint8_t cv_tz_code[15][16] = {
    { 1, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 1 }, { 7, 6, 5, 4, 3, 5, 4, 3, 2, 3, 2, 3, 2, 1, 0, 0 }, { 5, 7, 6, 5, 4, 3, 4, 3, 2, 3, 2, 1, 1, 0, 0, 0 },
    { 3, 7, 5, 4, 6, 5, 4, 3, 3, 2, 2, 1, 0, 0, 0, 0 }, { 5, 4, 3, 7, 6, 5, 4, 3, 2, 1, 1, 0, 0, 0, 0, 0 }, { 1, 1, 7, 6, 5, 4, 3, 2, 1, 1, 0, 0, 0, 0, 0, 0 },
    { 1, 1, 5, 4, 3, 3, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0 }, { 1, 1, 1, 3, 3, 2, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0 }, { 1, 0, 1, 3, 2, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0 },
    { 1, 0, 1, 3, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0 }, { 0, 1, 1, 2, 1, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }, { 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 },
    { 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }, { 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }, { 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 },
};
// total_zeros for the 2x2 chroma DC block: [TotalCoeff - 1][total_zeros]
FH_TAB uint8_t cv_tzc_len[3][4] = { { 1, 2, 3, 3 }, { 1, 2, 2, 0 }, { 1, 1, 0, 0 } };
FH_TAB uint8_t cv_tzc_code[3][4] = { { 1, 1, 1, 0 }, { 1, 1, 0, 0 }, { 1, 0, 0, 0 } };
// run_before: [min(zerosLeft, 7) - 1][run_before] for zerosLeft <= 6; zerosLeft > 6 is the unary tail below
FH_TAB uint8_t cv_rb_len[6][7] = { { 1, 1, 0, 0, 0, 0, 0 }, { 1, 2, 2, 0, 0, 0, 0 }, { 2, 2, 2, 2, 0, 0, 0 }, { 2, 2, 2, 3, 3, 0, 0 }, { 2, 2, 3, 3, 3, 3, 0 }, { 2, 3, 3, 3, 3, 3, 3 } };
FH_TAB uint8_t cv_rb_code[6][7] = { { 1, 0, 0, 0, 0, 0, 0 }, { 1, 1, 0, 0, 0, 0, 0 }, { 3, 2, 1, 0, 0, 0, 0 }, { 3, 2, 1, 1, 0, 0, 0 }, { 3, 2, 3, 2, 1, 0, 0 }, { 3, 0, 1, 3, 2, 5, 4 } };
// coded_block_pattern -> codeNum for Inter macroblocks (Table 9-4, ChromaArrayType 1)
FH_TAB uint8_t cv_cbp_inter[48] = { 0, 2, 3, 7, 4, 8, 17, 13, 5, 18, 9, 14, 10, 15, 16, 11, 1, 32, 33, 36, 34, 37, 44, 40, 35, 45, 38, 41, 39, 42, 43, 19,
                                    6, 24, 25, 20, 26, 21, 46, 28, 27, 47, 22, 29, 23, 30, 31, 12 };

// ---- bit writer: MSB-first into 32-bit words (word i holds stream bits 32i .. 32i+31, bit 32i in the MSB) ------------------
struct CvBits {
    uint32_t *buf;
    int cap;                // words
    int nw;                 // words written
    uint64_t acc;
    int nacc;               // bits pending in acc (< 32)
    int ovf;
};
FH_HD void cv_init(CvBits &b, uint32_t *buf, int cap_words) { b.buf = buf; b.cap = cap_words; b.nw = 0; b.acc = 0; b.nacc = 0; b.ovf = 0; }
FH_HD void cv_put(CvBits &b, int n, uint32_t v)          // n <= 32, v < 2^n
{
    if (n == 0) return;
    b.acc = (b.acc << n) | v;
    b.nacc += n;
    if (b.nacc >= 32) {
        b.nacc -= 32;
        const uint32_t w = (uint32_t)(b.acc >> b.nacc);
        if (b.nw < b.cap) b.buf[b.nw] = w; else b.ovf = 1;
        b.nw++;
    }
}
FH_HD int cv_bits(const CvBits &b) { return b.nw * 32 + b.nacc; }
FH_HD void cv_flush(CvBits &b)                           // zero-pads the last word
{
    if (b.nacc) { const uint32_t w = (uint32_t)(b.acc << (32 - b.nacc)); if (b.nw < b.cap) b.buf[b.nw] = w; else b.ovf = 1; }
}
FH_HD int cv_ilog2(uint32_t v) { int n = 0; while (v >>= 1) n++; return n; }
FH_HD void cv_ue(CvBits &b, uint32_t v) { const int k = cv_ilog2(v + 1); cv_put(b, k, 0); cv_put(b, k + 1, v + 1); }       // expgolomb.cpp:80-92
FH_HD void cv_se(CvBits &b, int v) { cv_ue(b, v <= 0 ? (uint32_t)(-v) * 2u : (uint32_t)v * 2u - 1u); }                    // :94-106

// ---- one residual block (what residual_block_cavlc_write emits, residual.cpp:374-666). coef[0 .. maxc-1] in scan order; returns
//      TotalCoeff; *bad set when a level is outside what the reference's level table can code (level_prefix <= 15,
//      residual_tables.cpp:940-1008).
// The block is described by its NON-ZERO MASK (bit i = coef[i] != 0): TotalCoeff is its population count, the coefficients come
// in coding order (highest frequency first) by peeling the top set bit, TrailingOnes are the leading +-1 among the first three
// peeled, total_zeros = zeros below the top set bit, and run_before is the gap between consecutive set bits — no level / run
// arrays, no backward scans.
#if defined(__CUDA_ARCH__)
#define CV_POPC(x) __popc(x)
#define CV_TOPBIT(x) (31 - __clz((int)(x)))
#else
#define CV_POPC(x) __builtin_popcount(x)
#define CV_TOPBIT(x) (31 - __builtin_clz(x))
#endif
FH_HD int cv_block(CvBits &b, const int16_t *coef, int maxc, int nC, int *bad)
{
    uint32_t nz = 0;
    for (int i = 0; i < maxc; i++) nz |= (uint32_t)(coef[i] != 0) << i;
    const int tc = CV_POPC(nz);
    const int tab = nC < 0 ? 4 : (nC < 2 ? 0 : (nC < 4 ? 1 : (nC < 8 ? 2 : 3)));
    if (tc == 0) { cv_put(b, cv_ct_len[tab][0][0], cv_ct_code[tab][0][0]); return 0; }
    int t1 = 0;
    for (uint32_t m = nz; m && t1 < 3; t1++) {
        const int p = CV_TOPBIT(m);
        if (coef[p] != 1 && coef[p] != -1) break;
        m &= ~(1u << p);
    }
    cv_put(b, cv_ct_len[tab][tc][t1], cv_ct_code[tab][tc][t1]);
    int sl = (tc > 10 && t1 < 3) ? 1 : 0;
    uint32_t m = nz;
    for (int i = 0; i < tc; i++) {
        const int p = CV_TOPBIT(m);
        m &= ~(1u << p);
        const int lv = coef[p];
        if (i < t1) { cv_put(b, 1, (uint32_t)((1 - lv) >> 1)); continue; }
        int lc = lv < 0 ? -(lv << 1) - 1 : (lv << 1) - 2;
        if (i == t1 && t1 < 3) lc -= 2;
        // level_prefix / level_suffix (9.2.2.1 inverted): escape at prefix 14 (suffixLength 0) and 15
        int prefix, ssize;
        uint32_t suffix;
        if (sl == 0) {
            if (lc < 14) { prefix = lc; ssize = 0; suffix = 0; }
            else if (lc < 30) { prefix = 14; ssize = 4; suffix = (uint32_t)(lc - 14); }
            else { prefix = 15; ssize = 12; suffix = (uint32_t)(lc - 30); }
        } else if (lc < (15 << sl)) { prefix = lc >> sl; ssize = sl; suffix = (uint32_t)(lc & ((1 << sl) - 1)); }
        else { prefix = 15; ssize = 12; suffix = (uint32_t)(lc - (15 << sl)); }
        if (suffix >= 4096u) { *bad = 1; suffix &= 4095u; }
        cv_put(b, prefix, 0);
        cv_put(b, 1, 1);
        cv_put(b, ssize, suffix);
        if (sl == 0) sl = 1;
        const int al = lv < 0 ? -lv : lv;
        if (al > (3 << (sl - 1)) && sl < 6) sl++;
    }
    int zl = CV_TOPBIT(nz) + 1 - tc;                       // total_zeros
    if (tc < maxc) {
        if (nC >= 0) cv_put(b, cv_tz_len[tc - 1][zl], cv_tz_code[tc - 1][zl]);
        else cv_put(b, cv_tzc_len[tc - 1][zl], cv_tzc_code[tc - 1][zl]);
    } else zl = 0;
    m = nz;
    int p = CV_TOPBIT(m);
    m &= ~(1u << p);
    while (m && zl > 0) {                                  // run_before of every coefficient but the last, while zeros are left
        const int q = CV_TOPBIT(m);
        m &= ~(1u << q);
        const int run = p - q - 1;
        if (zl > 6) { if (run < 7) cv_put(b, 3, (uint32_t)(7 - run)); else { cv_put(b, run - 4, 0); cv_put(b, 1, 1); } }   // residual.cpp:73-84
        else cv_put(b, cv_rb_len[zl - 1][run], cv_rb_code[zl - 1][run]);
        zl -= run;
        p = q;
    }
    return tc;
}

// ---- per-macroblock side information shared between neighbours (the reference's CodedBlockPattern*Array, totalcoeff_array_*
//      and mb_type_array): 32 bytes. tc_* hold 0 for blocks that are not coded (P_Skip, or their CBP bit clear), which is what
//      residual.cpp:458-486 substitutes when it reads a neighbour ------------------------------------------------------------
struct CvInfo {
    uint8_t skip, cbp_luma, cbp_chroma, mb_type;
    uint8_t tc_luma[16];        // by luma4x4BlkIdx (z-order)
    uint8_t tc_chroma[2][4];    // by chroma4x4BlkIdx
    uint8_t pad2[4];
};

FH_HD int cv_count(const int16_t *c, int n) { int t = 0; for (int i = 0; i < n; i++) t += c[i] != 0; return t; }

// setCodedBlockPattern (rbsp_encoding.cpp:21-105) + the TotalCoeff of every block that will be coded
FH_HD void cv_prepare(int mb_type, const int16_t luma[16][16], const int16_t cdc[2][4], const int16_t cac[2][4][15], int p_skip_type, CvInfo &o)
{
    for (int i = 0; i < 16; i++) o.tc_luma[i] = 0;
    for (int i = 0; i < 8; i++) o.tc_chroma[i >> 2][i & 3] = 0;
    o.mb_type = (uint8_t)mb_type; o.pad2[0] = o.pad2[1] = o.pad2[2] = o.pad2[3] = 0;
    o.skip = mb_type == p_skip_type; o.cbp_luma = 0; o.cbp_chroma = 0;
    if (o.skip) return;
    int cl = 0;
    for (int i8 = 0; i8 < 4; i8++) {
        int any = 0;
        for (int i4 = 0; i4 < 4; i4++) any |= cv_count(luma[i8 * 4 + i4], 16);
        if (any) cl |= 1 << i8;
    }
    int cc = 0;
    for (int i = 0; i < 4; i++) if (cdc[0][i] != 0 || cdc[1][i] != 0) cc = 1;
    for (int i4 = 0; i4 < 4; i4++) if (cv_count(cac[0][i4], 15) || cv_count(cac[1][i4], 15)) cc = 2;       // 1 | 2 = 3 is folded to 2 (:94-98)
    o.cbp_luma = (uint8_t)cl; o.cbp_chroma = (uint8_t)cc;
    for (int i = 0; i < 16; i++) if (cl & (1 << (i >> 2))) o.tc_luma[i] = (uint8_t)cv_count(luma[i], 16);
    if (cc & 2) for (int c = 0; c < 2; c++) for (int i = 0; i < 4; i++) o.tc_chroma[c][i] = (uint8_t)cv_count(cac[c][i], 15);
}

// nC of a block from its left (A) / upper (B) neighbour counts; -1 = not available (residual.cpp:488-503)
FH_HD int cv_nc(int nA, int nB) { return (nA >= 0 && nB >= 0) ? (nA + nB + 1) >> 1 : (nA >= 0 ? nA : (nB >= 0 ? nB : 0)); }

// macroblock_layer() of one non-skipped P macroblock (rbsp_encoding.cpp:222-305), preceded by mb_skip_run (:186).
// left / up = side information of the neighbouring macroblocks or null when outside the picture.
FH_HD void cv_macroblock(CvBits &b, int skip_run, int mb_type, int nparts, const int16_t mvd[4][2], const int16_t luma[16][16], const int16_t cdc[2][4],
                         const int16_t cac[2][4][15], const CvInfo &me, const CvInfo *left, const CvInfo *up, int *bad)
{
    cv_ue(b, (uint32_t)skip_run);
    cv_ue(b, (uint32_t)mb_type);
    if (nparts == 4) for (int i = 0; i < 4; i++) cv_ue(b, 0);                 // sub_mb_type: always 0 (P_L0_8x8)
    for (int i = 0; i < nparts; i++) { cv_se(b, mvd[i][0]); cv_se(b, mvd[i][1]); }
    const int cbp = (me.cbp_chroma << 4) | me.cbp_luma;
    cv_ue(b, cv_cbp_inter[cbp]);
    if (cbp == 0) return;
    cv_se(b, 0);                                                              // mb_qp_delta
    for (int blk = 0; blk < 16; blk++) {
        if (!(me.cbp_luma & (1 << (blk >> 2)))) continue;
        // luma4x4BlkIdx -> 4x4 coordinates (z-order) and the neighbours A (left) / B (above), 6.4.10.4
        const int bx = ((blk >> 2) & 1) * 2 + (blk & 1), by = (blk >> 3) * 2 + ((blk >> 1) & 1);
        int nA, nB;
        if (bx > 0) { const int a = (by >> 1) * 8 + ((bx - 1) >> 1) * 4 + (by & 1) * 2 + ((bx - 1) & 1); nA = me.tc_luma[a]; }
        else { const int a = (by >> 1) * 8 + 4 + (by & 1) * 2 + 1; nA = left ? left->tc_luma[a] : -1; }
        if (by > 0) { const int a = ((by - 1) >> 1) * 8 + (bx >> 1) * 4 + ((by - 1) & 1) * 2 + (bx & 1); nB = me.tc_luma[a]; }
        else { const int a = 8 + (bx >> 1) * 4 + 2 + (bx & 1); nB = up ? up->tc_luma[a] : -1; }
        cv_block(b, luma[blk], 16, cv_nc(nA, nB), bad);
    }
    if (me.cbp_chroma & 3) for (int c = 0; c < 2; c++) cv_block(b, cdc[c], 4, -1, bad);
    if (me.cbp_chroma & 2)
        for (int c = 0; c < 2; c++)
            for (int blk = 0; blk < 4; blk++) {
                const int bx = blk & 1, by = blk >> 1;
                const int nA = bx ? me.tc_chroma[c][blk - 1] : (left ? left->tc_chroma[c][blk + 1] : -1);
                const int nB = by ? me.tc_chroma[c][blk - 2] : (up ? up->tc_chroma[c][blk + 2] : -1);
                cv_block(b, cac[c][blk], 15, cv_nc(nA, nB), bad);
            }
}
