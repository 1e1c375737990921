// Phase C — fully parallel over macroblocks: final motion compensation, pixel snapping, and the fused
// residual -> 4x4 transform -> quant -> zigzag -> dequant -> inverse -> clip -> reconstruction, all in registers.
// One warp per macroblock: lanes 0-15 own the luma 4x4 blocks (z-order, h264_globals.cpp:209-214), lanes 16-23
// the Cb/Cr 4x4 blocks; the 2x2 chroma-DC Hadamard is exchanged with warp shuffles.
// Reference: mocomp.cpp:152-208, moestimation.cpp:571-584, quantizationTransform.cpp:41-100,157-282,349-485,
// scaleTransform.cpp:101-150,247-262,308-340,408-420, inttransform.cpp:133-154,215-321.
#pragma once
#include "common.cuh"

__constant__ int c_LQ[6][3] = { {205, 158, 128}, {186, 146, 114}, {158, 128, 102}, {146, 114, 89}, {128, 102, 82}, {114, 89, 71} };
__constant__ int c_LS[6][3] = { {160, 208, 256}, {176, 224, 288}, {208, 256, 320}, {224, 288, 368}, {256, 320, 400}, {288, 368, 464} };
__constant__ int c_QPC[52] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27,
                               28, 29, 29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39 };
// zigzag position k -> raster index row*4+col (scaleTransform.cpp:43-47)
__constant__ int c_ZZ[16] = { 0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15 };

// ---- generic fractional samples (picture-edge blocks only; interior luma comes from the phase-R planes) ----
struct ImgRef { const uint8_t *p; int W, H; };
__device__ __forceinline__ int E_(const ImgRef &r, int x, int y) { return r.p[(size_t)clampi_(y, 0, r.H - 1) * r.W + clampi_(x, 0, r.W - 1)]; }
__device__ __forceinline__ int halfh_(const ImgRef &r, int x, int y) { return tap6_(E_(r, x - 2, y), E_(r, x - 1, y), E_(r, x, y), E_(r, x + 1, y), E_(r, x + 2, y), E_(r, x + 3, y)); }
__device__ __forceinline__ int halfv_(const ImgRef &r, int x, int y) { return tap6_(E_(r, x, y - 2), E_(r, x, y - 1), E_(r, x, y), E_(r, x, y + 1), E_(r, x, y + 2), E_(r, x, y + 3)); }

// mocomp.cpp:50-78 with the per-coordinate clamp of mocomp.cpp:11-23; (x,y) may be outside the picture.
__device__ __noinline__ int luma_frac_generic(ImgRef r, int x, int y, int fx, int fy)
{
    const int G = E_(r, x, y);
    if (fx == 0 && fy == 0) return G;
    const int b = halfh_(r, x, y);
    if (fy == 0) return fx == 1 ? mid_(G, b) : (fx == 2 ? b : mid_(b, E_(r, x + 1, y)));
    const int h = halfv_(r, x, y);
    if (fx == 0) return fy == 1 ? mid_(G, h) : (fy == 2 ? h : mid_(h, E_(r, x, y + 1)));
    if (fx == 1 && fy == 1) return mid_(b, h);
    const int m = halfv_(r, x + 1, y);
    if (fx == 3 && fy == 1) return mid_(b, m);
    const int s = halfh_(r, x, y + 1);
    if (fx == 1 && fy == 3) return mid_(h, s);
    if (fx == 3 && fy == 3) return mid_(s, m);
    const int j = tap6_(halfv_(r, x - 2, y), halfv_(r, x - 1, y), h, m, halfv_(r, x + 2, y), halfv_(r, x + 3, y));
    if (fx == 2 && fy == 2) return j;
    if (fx == 2 && fy == 1) return mid_(b, j);
    if (fx == 1 && fy == 2) return mid_(h, j);
    if (fx == 2 && fy == 3) return mid_(j, s);
    return mid_(j, m);
}

// mocomp.cpp:176-194, samples clamped per coordinate (mocomp.cpp:24-35)
__device__ __forceinline__ int chroma_frac_(const ImgRef &r, int x, int y, int xf, int yf)
{
    const int A = E_(r, x, y), B = E_(r, x + 1, y), C = E_(r, x, y + 1), D = E_(r, x + 1, y + 1);
    return ((8 - xf) * (8 - yf) * A + xf * (8 - yf) * B + (8 - xf) * yf * C + xf * yf * D + 32) >> 6;
}

// Luma prediction of a w x h block at picture position (X,Y) (already displaced by mv>>2) with fraction f.
// Interior blocks read the interpolated plane (identical values: phase R evaluates the same formula at every
// integer position); blocks touching the outside of the picture take the generic path.
template <int BW, int BH>
__device__ __forceinline__ void luma_pred_block(const SeqDev &S, const Geo &g, int X, int Y, int fx, int fy, int out[BW * BH])
{
    if (X >= 0 && Y >= 0 && X + BW <= g.W && Y + BH <= g.H) {
        const uint8_t *pl = S.planes + (size_t)(fy * 4 + fx) * g.WH + (size_t)Y * g.W + X;
#pragma unroll
        for (int r = 0; r < BH; r++)
#pragma unroll
            for (int c = 0; c < BW; c++) out[r * BW + c] = pl[(size_t)r * g.W + c];
    } else {
        ImgRef R = { S.ref[0], g.W, g.H };
        for (int r = 0; r < BH; r++)
            for (int c = 0; c < BW; c++) out[r * BW + c] = luma_frac_generic(R, X + c, Y + r, fx, fy);
    }
}

// ---- 4x4 transform pipeline in registers -----------------------------------------------------------------
__device__ __forceinline__ void fwd4_(int a, int b, int c, int d, int &o0, int &o1, int &o2, int &o3)
{   // quantizationTransform.cpp:58-77: 416 = 256+128+32, 208 = 128+64+16
    o0 = ((a + b + c + d) * 256 + 512) >> 10;
    o1 = (416 * a + 208 * b - 208 * c - 416 * d + 512) >> 10;
    o2 = ((a - b - c + d) * 256 + 512) >> 10;
    o3 = (208 * a - 416 * b + 416 * c - 208 * d + 512) >> 10;
}

__device__ __forceinline__ void forward4x4_(const int r[16], int d[16])
{
    int h[16], f[16];
#pragma unroll
    for (int i = 0; i < 16; i++) h[i] = r[i] == 0 ? 0 : r[i] * 64 - 32;        // :53
#pragma unroll
    for (int j = 0; j < 4; j++) fwd4_(h[j], h[4 + j], h[8 + j], h[12 + j], f[j], f[4 + j], f[8 + j], f[12 + j]);
#pragma unroll
    for (int i = 0; i < 4; i++) fwd4_(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3], d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
}

__device__ __forceinline__ void quant4x4_(const int d[16], int c[16], int qP, bool keep_dc)
{   // quantizationTransform.cpp:183-223
    const int per = qP / 6, rem = qP - per * 6;
    const int lq0 = c_LQ[rem][0], lq1 = c_LQ[rem][1], lq2 = c_LQ[rem][2];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int cls = ((i >> 2) & 1) + (i & 1);
        const int lq = cls == 0 ? lq0 : (cls == 1 ? lq1 : lq2);
        int t;
        if (qP < 24) t = (d[i] * (1 << (4 - per)) - (1 << (3 - per))) * lq;
        else t = (d[i] >> (per - 4)) * lq;
        c[i] = (t + 16384) >> 15;
    }
    if (keep_dc) c[0] = d[0];
}

__device__ __forceinline__ void dequant4x4_(const int c[16], int d[16], int qP, bool keep_dc)
{   // scaleTransform.cpp:308-340
    const int per = qP / 6, rem = qP - per * 6;
    const int ls0 = c_LS[rem][0], ls1 = c_LS[rem][1], ls2 = c_LS[rem][2];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int cls = ((i >> 2) & 1) + (i & 1);
        const int ls = cls == 0 ? ls0 : (cls == 1 ? ls1 : ls2);
        if (qP >= 24) d[i] = (c[i] * ls) * (1 << (per - 4));
        else d[i] = (c[i] * ls + (1 << (3 - per))) >> (4 - per);
    }
    if (keep_dc) d[0] = c[0];
}

__device__ __forceinline__ void inverse4x4_(const int d[16], int r[16])
{   // scaleTransform.cpp:101-150
    int f[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int e0 = d[4 * i] + d[4 * i + 2], e1 = d[4 * i] - d[4 * i + 2];
        const int e2 = (d[4 * i + 1] >> 1) - d[4 * i + 3], e3 = d[4 * i + 1] + (d[4 * i + 3] >> 1);
        f[4 * i] = e0 + e3; f[4 * i + 1] = e1 + e2; f[4 * i + 2] = e1 - e2; f[4 * i + 3] = e0 - e3;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int g0 = f[j] + f[8 + j], g1 = f[j] - f[8 + j];
        const int g2 = (f[4 + j] >> 1) - f[12 + j], g3 = f[4 + j] + (f[12 + j] >> 1);
        r[j] = (g0 + g3 + 32) >> 6; r[4 + j] = (g1 + g2 + 32) >> 6; r[8 + j] = (g1 - g2 + 32) >> 6; r[12 + j] = (g0 - g3 + 32) >> 6;
    }
}

__device__ __forceinline__ int blkx_(int b) { return ((b & 1) << 2) | ((b & 4) << 1); }
__device__ __forceinline__ int blky_(int b) { return ((b & 2) << 1) | (b & 8); }

// The per-lane body shared by the in-pipeline kernel and the stand-alone TQ entry point.
// lane 0-15: luma block `lane`; lane 16-23: chroma (comp = (lane-16)>>2, blk = (lane-16)&3); other lanes idle
// but must call (shuffles). src/pred: this lane's 4x4 samples. Writes levels into rec (shared memory) and
// leaves the reconstruction in recon[16].
__device__ __forceinline__ void tq_lane(int lane, int qp, const int src[16], const int pred[16], fh264_mb_result *rec, int recon[16])
{
    const bool luma = lane < 16, chroma = lane >= 16 && lane < 24;
    const int qpc = c_QPC[clampi_(qp, 0, 51)];
    const int q = luma ? qp : qpc;
    int r[16], d[16], c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = src[i] - pred[i];
    forward4x4_(r, d);
    quant4x4_(d, c, q, !luma);
    // chroma DC: 2x2 Hadamard over the four blocks of the component (quantizationTransform.cpp:157-178,264-282)
    const int grp = 16 + ((lane - 16) & 4);             // first lane of this component's group
    const int a0 = __shfl_sync(0xffffffffu, c[0], grp & 31), a1 = __shfl_sync(0xffffffffu, c[0], (grp + 1) & 31);
    const int a2 = __shfl_sync(0xffffffffu, c[0], (grp + 2) & 31), a3 = __shfl_sync(0xffffffffu, c[0], (grp + 3) & 31);
    if (chroma) {
        const int blk = (lane - 16) & 3, comp = (lane - 16) >> 2;
        const int per = qpc / 6, rem = qpc - per * 6;
        int f[4], lv[4], g[4];
        f[0] = (a0 + a1 + a2 + a3 + 2) >> 2; f[1] = (a0 - a1 + a2 - a3 + 2) >> 2;
        f[2] = (a0 + a1 - a2 - a3 + 2) >> 2; f[3] = (a0 - a1 - a2 + a3 + 2) >> 2;
#pragma unroll
        for (int i = 0; i < 4; i++) lv[i] = ((((f[i] * 32) >> per) * c_LQ[rem][0]) + 16384) >> 15;
        // inverse (scaleTransform.cpp:247-262,408-420)
        g[0] = lv[0] + lv[1] + lv[2] + lv[3]; g[1] = lv[0] - lv[1] + lv[2] - lv[3];
        g[2] = lv[0] + lv[1] - lv[2] - lv[3]; g[3] = lv[0] - lv[1] - lv[2] + lv[3];
        const int gm = blk == 0 ? g[0] : (blk == 1 ? g[1] : (blk == 2 ? g[2] : g[3]));
        rec->chroma_dc[comp][blk] = (int16_t)(blk == 0 ? lv[0] : (blk == 1 ? lv[1] : (blk == 2 ? lv[2] : lv[3])));
#pragma unroll
        for (int k = 1; k < 16; k++) rec->chroma_ac[comp][blk][k - 1] = (int16_t)c[c_ZZ[k]];
        c[0] = ((gm * c_LS[rem][0]) * (1 << per)) >> 5;
    } else if (luma) {
#pragma unroll
        for (int k = 0; k < 16; k++) rec->luma[lane][k] = (int16_t)c[c_ZZ[k]];
    }
    int x[16], rr[16];
    dequant4x4_(c, x, q, !luma);
    inverse4x4_(x, rr);
#pragma unroll
    for (int i = 0; i < 16; i++) recon[i] = clip255_(pred[i] + rr[i]);
}

__global__ void __launch_bounds__(128) k_phase_c(const SeqDev *__restrict__ seqs, int seq0, Geo g, fh264_params prm, uint32_t epoch)
{
    __shared__ __align__(16) fh264_mb_result recs[4];
    const SeqDev &S = seqs[seq0 + blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = g.band_mb0 + blockIdx.x * 4 + warp;
    const uint32_t gated = S.status[ST_GATE];
    if (blockIdx.x == 0 && threadIdx.x == 0) { S.status[ST_GATE_DONE] = gated; if (gated) S.status[ST_GATED_TOTAL] += 1u; }
    if (mb >= g.band_mb0 + g.band_nmb) return;
    if (gated) {
        // Scene cut (fh264_encode_p_stream): the picture is not coded as P. The reconstruction buffer takes a copy of the reference so
        // that the dpb swap and phase R that follow on the stream leave the sequence exactly where it was; records stay untouched.
        const int mbx = mb % g.Wmb, mby = mb / g.Wmb, CW = g.W >> 1;
        if (lane < 16) {
            const size_t o = (size_t)(mby * 16 + lane) * g.W + mbx * 16;
            *(uint4 *)(S.rec[0] + o) = *(const uint4 *)(S.ref[0] + o);
        } else {
            const int c = (lane - 16) >> 3, r = (lane - 16) & 7;
            const size_t o = (size_t)(mby * 8 + r) * CW + mbx * 8;
            *(uint2 *)(S.rec[1 + c] + o) = *(const uint2 *)(S.ref[1 + c] + o);
        }
        return;
    }
    fh264_mb_result *rec = &recs[warp];
    {   // zero the record (52 x 16 bytes)
        uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = lane; i < (int)(sizeof(fh264_mb_result) / 16); i += 32) ((uint4 *)rec)[i] = z;
    }
    __syncwarp();
    const MbMotion mo = S.motion[mb];
    const int mbx = mb % g.Wmb, mby = mb / g.Wmb;
    const bool skip = mo.mb_type == FH264_P_SKIP;
    if (lane == 0) {
        rec->mb_type = mo.mb_type; rec->num_parts = mo.num_parts;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            rec->mv[i][0] = mo.mv[i][0]; rec->mv[i][1] = mo.mv[i][1];
            rec->mvd[i][0] = mo.mvd[i][0]; rec->mvd[i][1] = mo.mvd[i][1];
            rec->sad[i] = mo.sad[i];
        }
    }
    const bool luma = lane < 16, chroma = lane >= 16 && lane < 24;
    int src[16], pred[16], recon[16];
    int comp = 0, bx = 0, by = 0, CW = g.W >> 1;
#pragma unroll
    for (int i = 0; i < 16; i++) src[i] = pred[i] = 0;
    if (luma) {
        bx = blkx_(lane); by = blky_(lane);
        const int qd = (by >> 3) * 2 + (bx >> 3);
        const int mvx = mo.mv[qd][0], mvy = mo.mv[qd][1];
        luma_pred_block<4, 4>(S, g, mbx * 16 + bx + (mvx >> 2), mby * 16 + by + (mvy >> 2), mvx & 3, mvy & 3, pred);
        const uint8_t *sp = S.cur[0] + (size_t)(mby * 16 + by) * g.W + mbx * 16 + bx;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            uint32_t w = *(const uint32_t *)(sp + (size_t)r * g.W);
#pragma unroll
            for (int c = 0; c < 4; c++) src[r * 4 + c] = (w >> (8 * c)) & 255;
        }
    } else if (chroma) {
        comp = (lane - 16) >> 2;
        const int blk = (lane - 16) & 3;
        bx = (blk & 1) * 4; by = (blk >> 1) * 4;
        const int mvx = mo.mv[blk][0], mvy = mo.mv[blk][1];
        ImgRef R = { S.ref[1 + comp], CW, g.H >> 1 };
        const int X = mbx * 8 + bx + (mvx >> 3), Y = mby * 8 + by + (mvy >> 3);
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) pred[r * 4 + c] = chroma_frac_(R, X + c, Y + r, mvx & 7, mvy & 7);
        const uint8_t *sp = S.cur[1 + comp] + (size_t)(mby * 8 + by) * CW + mbx * 8 + bx;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            uint32_t w = *(const uint32_t *)(sp + (size_t)r * CW);
#pragma unroll
            for (int c = 0; c < 4; c++) src[r * 4 + c] = (w >> (8 * c)) & 255;
        }
    }
    if (skip) {
        // P_Skip: all levels zero, reconstruction == prediction (inttransform.cpp:215-229)
#pragma unroll
        for (int i = 0; i < 16; i++) recon[i] = pred[i];
    } else {
        // pixel snapping of the source toward the prediction (moestimation.cpp:571-584): luma <, chroma <=
        const int md = mo.maxdiff;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int ad = iabs_(src[i] - pred[i]);
            if (luma ? (ad < md) : (ad <= md)) src[i] = pred[i];
        }
        tq_lane(lane, prm.qp, src, pred, rec, recon);
    }
    if (luma || chroma) {
        const int pitch = luma ? g.W : CW;
        const size_t off = luma ? (size_t)(mby * 16 + by) * g.W + mbx * 16 + bx : (size_t)(mby * 8 + by) * CW + mbx * 8 + bx;
        const int plane = luma ? 0 : 1 + comp;
        uint32_t rw[4];
#pragma unroll
        for (int r = 0; r < 4; r++) rw[r] = (uint32_t)recon[r * 4] | ((uint32_t)recon[r * 4 + 1] << 8) | ((uint32_t)recon[r * 4 + 2] << 16) | ((uint32_t)recon[r * 4 + 3] << 24);
        uint8_t *dp = S.rec[plane] + off;
#pragma unroll
        for (int r = 0; r < 4; r++) *(uint32_t *)(dp + (size_t)r * pitch) = rw[r];
        // band mode: the reconstruction exchange is fused here — the band's rows go straight into every other rank's
        // picture over NVLink peer stores (completed by the kernel boundary + k_band_barrier before anyone reads them)
        for (int pr = 0; pr < g.world; pr++) {
            if (pr == g.rank) continue;
            uint8_t *pp = S.peer_rec[pr][plane] + off;
#pragma unroll
            for (int r = 0; r < 4; r++) *(uint32_t *)(pp + (size_t)r * pitch) = rw[r];
            if (lane == 0) S.peer_motion[pr][mb].mb_type = mo.mb_type;
        }
    }
    __syncwarp();
    uint4 *dst = (uint4 *)&S.results[mb];
    for (int i = lane; i < (int)(sizeof(fh264_mb_result) / 16); i += 32) dst[i] = ((const uint4 *)rec)[i];
    if (g.world > 1 && g.gather_on && S.gather[0]) {                          // band mode: rank 0 collects the picture's records (device CAVLC of the slice)
        uint4 *gd = (uint4 *)&S.gather[epoch & 1u][mb];
        for (int i = lane; i < (int)(sizeof(fh264_mb_result) / 16); i += 32) gd[i] = ((const uint4 *)rec)[i];
    }
}

// Stand-alone fused TQ over n macroblocks given source and prediction (unit tests; I-picture helper).
// ---- decoder side: the inverse half of the same per-lane body --------------------------------
// Levels come from the record instead of the forward transform: inverse zigzag, chroma-DC inverse Hadamard + scaling
// (scaleTransform.cpp:247-262,408-420), dequantisation, inverse 4x4, Clip1(pred + r) — what rbsp_decoding.cpp:330-346 does per
// macroblock through transformDecoding4x4LumaResidual / transformDecodingChroma (inttransform.cpp:133-154,237-321).
__device__ __forceinline__ void dec_lane(int lane, int qp, const int pred[16], const fh264_mb_result *rec, int recon[16])
{
    const bool luma = lane < 16, chroma = lane >= 16 && lane < 24;
    const int qpc = c_QPC[clampi_(qp, 0, 51)];
    const int q = luma ? qp : qpc;
    int c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = 0;
    if (luma) {
#pragma unroll
        for (int k = 0; k < 16; k++) c[c_ZZ[k]] = rec->luma[lane][k];
    } else if (chroma) {
        const int blk = (lane - 16) & 3, comp = (lane - 16) >> 2;
        const int per = qpc / 6, rem = qpc - per * 6;
#pragma unroll
        for (int k = 1; k < 16; k++) c[c_ZZ[k]] = rec->chroma_ac[comp][blk][k - 1];
        const int l0 = rec->chroma_dc[comp][0], l1 = rec->chroma_dc[comp][1], l2 = rec->chroma_dc[comp][2], l3 = rec->chroma_dc[comp][3];
        const int gm = blk == 0 ? l0 + l1 + l2 + l3 : (blk == 1 ? l0 - l1 + l2 - l3 : (blk == 2 ? l0 + l1 - l2 - l3 : l0 - l1 - l2 + l3));
        c[0] = ((gm * c_LS[rem][0]) * (1 << per)) >> 5;
    }
    int x[16], rr[16];
    dequant4x4_(c, x, q, !luma);
    inverse4x4_(x, rr);
#pragma unroll
    for (int i = 0; i < 16; i++) recon[i] = clip255_(pred[i] + rr[i]);
}

// One warp per macroblock: motion compensation from the record's quadrant MVs (Decode, mocomp.cpp:200-208) + dec_lane;
// the reconstruction goes into S.rec like the encoder's (P_Skip: reconstruction = prediction, inttransform.cpp:215-229).
__global__ void __launch_bounds__(128) k_decode_p(const SeqDev *__restrict__ seqs, int seq0, Geo g, int qp)
{
    const SeqDev &S = seqs[seq0 + blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * 4 + warp;
    if (mb >= g.nmb) return;
    const fh264_mb_result *rec = &S.results[mb];
    const int mbx = mb % g.Wmb, mby = mb / g.Wmb;
    const bool skip = rec->mb_type == FH264_P_SKIP;
    const bool luma = lane < 16, chroma = lane >= 16 && lane < 24;
    int pred[16], recon[16];
    int comp = 0, bx = 0, by = 0;
    const int CW = g.W >> 1;
#pragma unroll
    for (int i = 0; i < 16; i++) pred[i] = 0;
    if (luma) {
        bx = blkx_(lane); by = blky_(lane);
        const int qd = (by >> 3) * 2 + (bx >> 3);
        const int mvx = rec->mv[qd][0], mvy = rec->mv[qd][1];
        luma_pred_block<4, 4>(S, g, mbx * 16 + bx + (mvx >> 2), mby * 16 + by + (mvy >> 2), mvx & 3, mvy & 3, pred);
    } else if (chroma) {
        comp = (lane - 16) >> 2;
        const int blk = (lane - 16) & 3;
        bx = (blk & 1) * 4; by = (blk >> 1) * 4;
        const int mvx = rec->mv[blk][0], mvy = rec->mv[blk][1];
        ImgRef R = { S.ref[1 + comp], CW, g.H >> 1 };
        const int X = mbx * 8 + bx + (mvx >> 3), Y = mby * 8 + by + (mvy >> 3);
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) pred[r * 4 + c] = chroma_frac_(R, X + c, Y + r, mvx & 7, mvy & 7);
    }
    if (skip) {
#pragma unroll
        for (int i = 0; i < 16; i++) recon[i] = pred[i];
    } else dec_lane(lane, qp, pred, rec, recon);
    if (luma || chroma) {
        const int pitch = luma ? g.W : CW;
        const size_t off = luma ? (size_t)(mby * 16 + by) * g.W + mbx * 16 + bx : (size_t)(mby * 8 + by) * CW + mbx * 8 + bx;
        uint8_t *dp = S.rec[luma ? 0 : 1 + comp] + off;
#pragma unroll
        for (int r = 0; r < 4; r++)
            *(uint32_t *)(dp + (size_t)r * pitch) = (uint32_t)recon[r * 4] | ((uint32_t)recon[r * 4 + 1] << 8) | ((uint32_t)recon[r * 4 + 2] << 16) | ((uint32_t)recon[r * 4 + 3] << 24);
    }
}

__global__ void __launch_bounds__(128) k_tq_only(const uint8_t *__restrict__ src384, const uint8_t *__restrict__ pred384, int n, int qp,
                                                 int16_t *__restrict__ levels384, uint8_t *__restrict__ recon384)
{
    __shared__ __align__(16) fh264_mb_result recs[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * 4 + warp;
    if (mb >= n) return;
    fh264_mb_result *rec = &recs[warp];
    const bool luma = lane < 16, chroma = lane >= 16 && lane < 24;
    int src[16], pred[16], recon[16];
#pragma unroll
    for (int i = 0; i < 16; i++) src[i] = pred[i] = 0;
    int off = 0, pitch = 16;
    if (luma) off = blky_(lane) * 16 + blkx_(lane);
    else if (chroma) { const int blk = (lane - 16) & 3; pitch = 8; off = 256 + ((lane - 16) >> 2) * 64 + (blk >> 1) * 32 + (blk & 1) * 4; }
    if (luma || chroma) {
        const uint8_t *sp = src384 + (size_t)mb * 384 + off, *pp = pred384 + (size_t)mb * 384 + off;
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) { src[r * 4 + c] = sp[r * pitch + c]; pred[r * 4 + c] = pp[r * pitch + c]; }
    }
    tq_lane(lane, qp, src, pred, rec, recon);
    __syncwarp();
    if (luma || chroma) {
        uint8_t *dp = recon384 + (size_t)mb * 384 + off;
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) dp[r * pitch + c] = (uint8_t)recon[r * 4 + c];
    }
    // levels: luma[16][16], chroma_dc[2][4], chroma_ac[2][4][15] are contiguous in the record = 384 int16
    const int16_t *lv = &rec->luma[0][0];
    for (int i = lane; i < 384; i += 32) levels384[(size_t)mb * 384 + i] = lv[i];
}

// Intra16x16 luma: per-block transform with the DC kept, 4x4 Hadamard of the 16 DCs, reconstruction.
// quantizationTransform.cpp:105-152,227-260,387-415; scaleTransform.cpp:154-189,344-376; inttransform.cpp:157-208.
__global__ void __launch_bounds__(128) k_tq_intra16(const uint8_t *__restrict__ src256, const uint8_t *__restrict__ pred256, int n, int qp,
                                                    int16_t *__restrict__ dc16, int16_t *__restrict__ ac240, uint8_t *__restrict__ recon256)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * 4 + warp;
    if (mb >= n) return;
    const int b = lane & 15;
    const int x0 = blkx_(b), y0 = blky_(b);
    int r[16], pred[16], d[16], c[16];
    const uint8_t *sp = src256 + (size_t)mb * 256 + y0 * 16 + x0, *pp = pred256 + (size_t)mb * 256 + y0 * 16 + x0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { pred[i * 4 + j] = pp[i * 16 + j]; r[i * 4 + j] = (int)sp[i * 16 + j] - pred[i * 4 + j]; }
    forward4x4_(r, d);
    quant4x4_(d, c, qp, true);
    // gather the DC matrix DC[y0/4][x0/4] from the 16 block lanes
    int DC[16];
#pragma unroll
    for (int bb = 0; bb < 16; bb++) {
        const int v = __shfl_sync(0xffffffffu, c[0], bb);
        DC[(blky_(bb) >> 2) * 4 + (blkx_(bb) >> 2)] = v;
    }
    int t[16], u[16], cq[16];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int g0 = DC[j] + DC[12 + j], g1 = DC[4 + j] + DC[8 + j], g2 = DC[4 + j] - DC[8 + j], g3 = DC[j] - DC[12 + j];
        t[j] = g0 + g1; t[4 + j] = g3 + g2; t[8 + j] = g0 - g1; t[12 + j] = g3 - g2;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int d0 = t[4 * i] + t[4 * i + 3], d1 = t[4 * i + 1] + t[4 * i + 2], d2 = t[4 * i + 1] - t[4 * i + 2], d3 = t[4 * i] - t[4 * i + 3];
        u[4 * i] = (d0 + d1 + 8) >> 4; u[4 * i + 1] = (d3 + d2 + 8) >> 4; u[4 * i + 2] = (d0 - d1 + 8) >> 4; u[4 * i + 3] = (d3 - d2 + 8) >> 4;
    }
    const int per = qp / 6, rem = qp - per * 6, lq = c_LQ[rem][0], ls = c_LS[rem][0];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int tt = qp >= 36 ? (u[i] >> (per - 6)) * lq : (u[i] * (1 << (6 - per)) - (1 << (5 - per))) * lq;
        cq[i] = (tt + 16384) >> 15;
    }
    if (lane < 16) {
        dc16[(size_t)mb * 16 + lane] = (int16_t)cq[c_ZZ[lane]];
#pragma unroll
        for (int k = 1; k < 16; k++) ac240[(size_t)mb * 240 + b * 15 + k - 1] = (int16_t)c[c_ZZ[k]];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int d0 = cq[4 * i] + cq[4 * i + 2], d1 = cq[4 * i] - cq[4 * i + 2], d2 = cq[4 * i + 1] - cq[4 * i + 3], d3 = cq[4 * i + 1] + cq[4 * i + 3];
        t[4 * i] = d0 + d3; t[4 * i + 1] = d1 + d2; t[4 * i + 2] = d1 - d2; t[4 * i + 3] = d0 - d3;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int g0 = t[j] + t[8 + j], g1 = t[j] - t[8 + j], g2 = t[4 + j] - t[12 + j], g3 = t[4 + j] + t[12 + j];
        u[j] = g0 + g3; u[4 + j] = g1 + g2; u[8 + j] = g1 - g2; u[12 + j] = g0 - g3;
    }
    const int mine = u[(y0 >> 2) * 4 + (x0 >> 2)];
    c[0] = qp >= 36 ? (mine * ls) * (1 << (per - 6)) : (mine * ls + (1 << (5 - per))) >> (6 - per);
    int x[16], rr[16];
    dequant4x4_(c, x, qp, true);
    inverse4x4_(x, rr);
    if (lane < 16) {
        uint8_t *dp = recon256 + (size_t)mb * 256 + y0 * 16 + x0;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dp[i * 16 + j] = (uint8_t)clip255_(pred[i * 4 + j] + rr[i * 4 + j]);
    }
}

// Whole-picture motion compensation from given quadrant MVs (unit tests of the MC rules in isolation).
__global__ void __launch_bounds__(128) k_mc_only(const SeqDev *__restrict__ seqs, int seq, Geo g, const int16_t *__restrict__ qmv, uint8_t *__restrict__ pred384)
{
    const SeqDev &S = seqs[seq];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * 4 + warp;
    if (mb >= g.nmb || lane >= 24) return;
    const int mbx = mb % g.Wmb, mby = mb / g.Wmb;
    int pred[16];
    if (lane < 16) {
        const int bx = blkx_(lane), by = blky_(lane), qd = (by >> 3) * 2 + (bx >> 3);
        const int mvx = qmv[mb * 8 + qd * 2], mvy = qmv[mb * 8 + qd * 2 + 1];
        luma_pred_block<4, 4>(S, g, mbx * 16 + bx + (mvx >> 2), mby * 16 + by + (mvy >> 2), mvx & 3, mvy & 3, pred);
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) pred384[(size_t)mb * 384 + (by + r) * 16 + bx + c] = (uint8_t)pred[r * 4 + c];
    } else {
        const int comp = (lane - 16) >> 2, blk = (lane - 16) & 3, bx = (blk & 1) * 4, by = (blk >> 1) * 4;
        const int mvx = qmv[mb * 8 + blk * 2], mvy = qmv[mb * 8 + blk * 2 + 1];
        ImgRef R = { S.ref[1 + comp], g.W >> 1, g.H >> 1 };
        const int X = mbx * 8 + bx + (mvx >> 3), Y = mby * 8 + by + (mvy >> 3);
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++)
            pred384[(size_t)mb * 384 + 256 + comp * 64 + (by + r) * 8 + bx + c] = (uint8_t)chroma_frac_(R, X + c, Y + r, mvx & 7, mvy & 7);
    }
}
