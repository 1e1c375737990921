/*
 * ORACLE / TEST INFRASTRUCTURE — never linked into, imported by, or called from the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Plain-C restatement of the P-picture hot path of zoltanmaric/h264-fer ("reference", paths relative to
 * fer_h264/fer_h264/): phase R (quarter-pel planes, box-sum features, sum-sorted index), the three-stage
 * motion search with its list semantics, MV prediction, motion compensation, P_Skip test, pixel snapping,
 * forward 4x4 transform / quantisation / zigzag, dequantisation / inverse transform / reconstruction,
 * chroma-DC 2x2 and Intra16x16 luma-DC 4x4 Hadamard paths, and the scene-change SAD.
 *
 * Parity status: PINNED against the reference itself — tests/test_oracle_vs_reference.py runs
 * oracle/_ref/ref_encoder (the unmodified reference compiled by oracle/Makefile) and compares per-MB
 * records and reconstructions bit-for-bit; tests/golden/ holds committed vectors made the same way.
 * (The reference repository ships no golden vectors of its own — SURVEY.md §4.)
 *
 * Every function cites the reference lines it restates. All arithmetic is 32-bit int, >> is arithmetic.
 *
 * Known reference undefined behaviour that is NOT emulated (inputs must avoid it; fo_phase_r reports it):
 *   - 8x8 window sums equal to 0 break the counting sort prefix (moestimation.cpp:153-158);
 *   - window sums >= 16203 make the bucket lookup read koliko[16384] (moestimation.cpp:477-480).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FO_REC_INTS 405 /* mb_type, mv[4][2], mvd[4][2], sad[4], luma[16][16], cdc[2][4], cac[2][4][15] */
#define P_L0_16x16 0
#define P_L0_L0_16x8 1
#define P_L0_L0_8x16 2
#define P_8x8ref0 4
#define P_SKIP 31

typedef struct {
    int W, H, Wmb, Hmb;
    uint8_t *plane[16];     /* refFrameInterpolated[f].L            (moestimation.cpp:21)  */
    uint16_t *kar[5][16];   /* refFrameKar[k][f][y][x], W*H each    (moestimation.cpp:23)  */
    int32_t *sorted[5];     /* sortedSuma0: K0, y, x, K1, K2        (moestimation.cpp:24)  */
    int32_t start[16385];   /* koliko after the final shift: bucket s = [start[s], start[s+1]) */
    int ub_inputs;          /* 1 if the reference would hit the UB listed in the header */
    /* per-picture MV store: quadrant MVs of already coded MBs (mode_pred.cpp:16, A.7) */
    int32_t *qmv;           /* [mb][4][2] */
    uint8_t *coded_inter;   /* P_and_SP_macroblock_modes[type][2] not in {0, NA} (mode_pred.cpp:50) */
    /* statistics of the last fo_encode_p: candidates seen (for bench bookkeeping only) */
    long long n_feat_evals, n_sads;
} fo_ctx;

static inline int iabs(int a) { return a < 0 ? -a : a; }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

fo_ctx *fo_create(int W, int H)
{
    fo_ctx *c = (fo_ctx *)calloc(1, sizeof(fo_ctx));
    size_t n = (size_t)W * H;
    c->W = W; c->H = H; c->Wmb = W >> 4; c->Hmb = H >> 4;
    for (int f = 0; f < 16; f++) {
        c->plane[f] = (uint8_t *)malloc(n);
        for (int k = 0; k < 5; k++) c->kar[k][f] = (uint16_t *)malloc(n * 2);
    }
    for (int a = 0; a < 5; a++) c->sorted[a] = (int32_t *)malloc(n * 4);
    c->qmv = (int32_t *)calloc((size_t)c->Wmb * c->Hmb * 8, 4);
    c->coded_inter = (uint8_t *)calloc((size_t)c->Wmb * c->Hmb, 1);
    return c;
}

void fo_destroy(fo_ctx *c)
{
    if (!c) return;
    for (int f = 0; f < 16; f++) { free(c->plane[f]); for (int k = 0; k < 5; k++) free(c->kar[k][f]); }
    for (int a = 0; a < 5; a++) free(c->sorted[a]);
    free(c->qmv); free(c->coded_inter); free(c);
}

const uint8_t *fo_plane(const fo_ctx *c, int f) { return c->plane[f]; }
const uint16_t *fo_kar(const fo_ctx *c, int k, int f) { return c->kar[k][f]; }
const int32_t *fo_sorted(const fo_ctx *c, int a) { return c->sorted[a]; }
const int32_t *fo_bucket_start(const fo_ctx *c) { return c->start; }
int fo_ub_inputs(const fo_ctx *c) { return c->ub_inputs; }

/* ------------------------------------------------------------------------------------------------
 * Luma fractional sample (mocomp.cpp:39-78 L_MC_frac_interpol, with the per-coordinate clamp of the 9x9
 * fetch, mocomp.cpp:11-23). (x,y) may lie outside the picture; E() is the edge-extended reference.
 * The centre half-pel j is filtered from ROUNDED column half-pels (mocomp.cpp:67-71), not the standard's
 * unrounded intermediates.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { const uint8_t *p; int W, H; } fo_img;
static inline int E(const fo_img *r, int x, int y) { return r->p[(size_t)clampi(y, 0, r->H - 1) * r->W + clampi(x, 0, r->W - 1)]; }
static inline int tap6(int a, int b, int c, int d, int e, int f) { return clip255((a - 5 * b + 20 * c + 20 * d - 5 * e + f + 16) >> 5); }
static inline int mid(int a, int b) { return (a + b + 1) >> 1; }
static int half_h(const fo_img *r, int x, int y) { return tap6(E(r, x - 2, y), E(r, x - 1, y), E(r, x, y), E(r, x + 1, y), E(r, x + 2, y), E(r, x + 3, y)); }
static int half_v(const fo_img *r, int x, int y) { return tap6(E(r, x, y - 2), E(r, x, y - 1), E(r, x, y), E(r, x, y + 1), E(r, x, y + 2), E(r, x, y + 3)); }

static int luma_frac(const fo_img *r, int x, int y, int fx, int fy)
{
    int G = E(r, x, y);
    if (fx == 0 && fy == 0) return G;
    int b = half_h(r, x, y);
    if (fy == 0) return fx == 1 ? mid(G, b) : fx == 2 ? b : mid(b, E(r, x + 1, y));
    int h = half_v(r, x, y);
    if (fx == 0) return fy == 1 ? mid(G, h) : fy == 2 ? h : mid(h, E(r, x, y + 1));
    if (fx == 1 && fy == 1) return mid(b, h);
    int m = half_v(r, x + 1, y);
    if (fx == 3 && fy == 1) return mid(b, m);
    int s = half_h(r, x, y + 1);
    if (fx == 1 && fy == 3) return mid(h, s);
    if (fx == 3 && fy == 3) return mid(s, m);
    int j = tap6(half_v(r, x - 2, y), half_v(r, x - 1, y), h, m, half_v(r, x + 2, y), half_v(r, x + 3, y));
    if (fx == 2 && fy == 2) return j;
    if (fx == 2 && fy == 1) return mid(b, j);
    if (fx == 1 && fy == 2) return mid(h, j);
    if (fx == 2 && fy == 3) return mid(j, s);
    return mid(j, m); /* (3,2) */
}

/* Chroma 1/8-pel bilinear sample (mocomp.cpp:24-35,176-194), coordinates clamped per sample. */
static int chroma_frac(const fo_img *r, int x, int y, int xf, int yf)
{
    int A = E(r, x, y), B = E(r, x + 1, y), C = E(r, x, y + 1), D = E(r, x + 1, y + 1);
    return ((8 - xf) * (8 - yf) * A + xf * (8 - yf) * B + (8 - xf) * yf * C + xf * yf * D + 32) >> 6;
}

/* ------------------------------------------------------------------------------------------------
 * Phase R: FillInterpolatedRefFrame (moestimation.cpp:74-173).
 * ---------------------------------------------------------------------------------------------- */
static void box_features(fo_ctx *c, int f)
{
    /* moestimation.cpp:105-139: replicate-pad the plane by 8 at the right/bottom, suffix-sum table,
       then five box sums per position. Computed here from a prefix-sum table of the padded plane. */
    const int W = c->W, H = c->H, PW = W + 8, PH = H + 8;
    int32_t *S = (int32_t *)calloc((size_t)(PW + 1) * (PH + 1), 4);
    const uint8_t *pl = c->plane[f];
#define SAT(x, y) S[(size_t)(y) * (PW + 1) + (x)]
    for (int y = 0; y < PH; y++)
        for (int x = 0; x < PW; x++) {
            int v = pl[(size_t)(y < H ? y : H - 1) * W + (x < W ? x : W - 1)];
            SAT(x + 1, y + 1) = v + SAT(x, y + 1) + SAT(x + 1, y) - SAT(x, y);
        }
#define BOX(x0, y0, w, h) (SAT((x0) + (w), (y0) + (h)) - SAT((x0), (y0) + (h)) - SAT((x0) + (w), (y0)) + SAT((x0), (y0)))
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            size_t i = (size_t)y * W + x;
            c->kar[0][f][i] = (uint16_t)BOX(x, y, 8, 8);                                   /* :137 */
            c->kar[1][f][i] = (uint16_t)BOX(x, y, 8, 4);                                   /* :136 rows 0-3 */
            c->kar[2][f][i] = (uint16_t)BOX(x, y, 4, 8);                                   /* :135 cols 0-3 */
            c->kar[3][f][i] = (uint16_t)(BOX(x, y, 8, 2) + BOX(x, y + 4, 8, 2));           /* :133-134 rows 0,1,4,5 */
            c->kar[4][f][i] = (uint16_t)(BOX(x, y, 2, 8) + BOX(x + 4, y, 2, 8));           /* :131-132 cols 0,1,4,5 */
        }
#undef BOX
#undef SAT
    free(S);
}

void fo_phase_r(fo_ctx *c, const uint8_t *refY)
{
    const int W = c->W, H = c->H;
    fo_img r = { refY, W, H };
    /* (i) moestimation.cpp:79-104 — 16 planes; the 4x4 blocking of the reference is only loop structure */
    for (int f = 0; f < 16; f++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) c->plane[f][(size_t)y * W + x] = (uint8_t)luma_frac(&r, x, y, f & 3, f >> 2);
    /* (ii) */
    for (int f = 0; f < 16; f++) box_features(c, f);
    /* (iii) moestimation.cpp:140-172 — stable counting sort of plane-0 positions, x outer / y inner */
    int32_t *cnt = (int32_t *)calloc(16385, 4);
    c->ub_inputs = 0;
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            int k0 = c->kar[0][0][(size_t)y * W + x];
            cnt[k0]++;
            if (k0 == 0 || k0 >= 16203) c->ub_inputs = 1;
        }
    c->start[0] = 0;
    for (int s = 0; s < 16384; s++) c->start[s + 1] = c->start[s] + cnt[s];
    memcpy(cnt, c->start, 16384 * 4);
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            size_t i = (size_t)y * W + x;
            int pos = cnt[c->kar[0][0][i]]++;
            c->sorted[0][pos] = c->kar[0][0][i];
            c->sorted[1][pos] = y;
            c->sorted[2][pos] = x;
            c->sorted[3][pos] = c->kar[1][0][i];
            c->sorted[4][pos] = c->kar[2][0][i];
        }
    free(cnt);
}

/* ------------------------------------------------------------------------------------------------
 * Candidate list: the 65-slot insertion list bmins/bxs/bys (moestimation.cpp:27,277-291).
 * ---------------------------------------------------------------------------------------------- */
#define EMPTY_COST 1000000000
#define EMPTY_MV 100000000
typedef struct { int cost[65], mvx[65], mvy[65]; } fo_list;

static void list_clear_all(fo_list *L) { for (int j = 0; j < 65; j++) { L->cost[j] = EMPTY_COST; L->mvx[j] = L->mvy[j] = EMPTY_MV; } } /* :453-457 */
static void list_reset_costs(fo_list *L) { for (int j = 0; j < 65; j++) L->cost[j] = EMPTY_COST; }                                   /* :473,508 */
static void list_insert(fo_list *L, int cost, int mvx, int mvy)
{
    if (L->cost[64] < cost) return;                    /* :277 */
    int j = 64;
    while (j > 0 && cost < L->cost[j - 1]) {           /* :282-291 strict: equal costs keep arrival order */
        L->cost[j] = L->cost[j - 1]; L->mvx[j] = L->mvx[j - 1]; L->mvy[j] = L->mvy[j - 1];
        j--;
    }
    L->cost[j] = cost; L->mvx[j] = mvx; L->mvy[j] = mvy;
}

/* Feature distance (moestimation.cpp:267-276), s = suma[0..4] of the current block. */
static inline int feat_dist(const fo_ctx *c, const int s[5], int f, int x, int y)
{
    size_t i = (size_t)y * c->W + x;
    int K0 = c->kar[0][f][i], d = iabs(s[0] - K0);
    for (int k = 1; k < 5; k++) {
        int Kk = c->kar[k][f][i];
        d += iabs(s[k] - Kk) + iabs(s[0] - s[k] - K0 + Kk);
    }
    return d;
}

/* MEstimation (moestimation.cpp:254-296): window of half-size g around (px,py), fractions stepped by fstep. */
static void feature_search(fo_ctx *c, fo_list *L, const int s[5], int sx, int sy, int g, int fstep, int genx, int geny, int px, int py)
{
    for (int dx = px - g; dx <= px + g; dx++)
        for (int dy = py - g; dy <= py + g; dy++)
            for (int f = 0; f < 16; f += fstep) {
                int rx = sx + dx, ry = sy + dy;
                if (ry < 0 || ry >= c->H || rx < 0 || rx >= c->W) continue;           /* :265 block ORIGIN inside */
                int cost = (iabs(dx - genx) + iabs(dy - geny) + 4) * feat_dist(c, s, f, rx, ry);
                c->n_feat_evals++;
                list_insert(L, cost, (dx << 2) | (f & 3), (dy << 2) | ((f >> 2) & 3)); /* :280-281 */
            }
}

/* satdLuma8x8MVs (moestimation.cpp:175-195): plain 8x8 SAD against the interpolated planes; block origin
   clamped to the picture, then each index clamped at the right/bottom edge only. */
static int sad8x8(fo_ctx *c, const uint8_t *curY, int xP, int yP, int mvx, int mvy)
{
    const int W = c->W, H = c->H;
    int x0 = clampi(xP + (mvx >> 2), 0, W - 1), y0 = clampi(yP + (mvy >> 2), 0, H - 1);
    const uint8_t *pl = c->plane[(mvx & 3) + (mvy & 3) * 4];
    int sad = 0;
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            int px = x0 + j < W ? x0 + j : W - 1, py = y0 + i < H ? y0 + i : H - 1;
            sad += iabs((int)curY[(size_t)(yP + i) * W + xP + j] - (int)pl[(size_t)py * W + px]);
        }
    c->n_sads++;
    return sad;
}

/* ------------------------------------------------------------------------------------------------
 * MV prediction (mode_pred.cpp:48-161,252-332,381-426), formulated on quadrant MVs.
 * cur[4][2] are the quadrant MVs of the current MB decided so far.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int avail, mvx, mvy, same_ref; } fo_nb;

static fo_nb neighbour(const fo_ctx *c, int mbx, int mby, int xN, int yN, int cur[4][2])
{
    fo_nb n = { 0, 0, 0, 0 };
    int mb;
    if ((xN > 15 && yN >= 0) || yN > 15) return n;                                   /* mode_pred.cpp:65-66 */
    if (xN >= 0 && xN < 16 && yN >= 0) {                                             /* :69 current MB */
        int q = (yN >> 3) * 2 + (xN >> 3);
        n.avail = 1; n.mvx = cur[q][0]; n.mvy = cur[q][1]; n.same_ref = 1;
        return n;
    }
    if (yN < 0) {
        if (mby == 0) return n;                                                      /* :73,80,89 */
        if (xN > 15) { if (mbx == c->Wmb - 1) return n; mb = (mby - 1) * c->Wmb + mbx + 1; xN -= 16; } /* :78-83 */
        else if (xN < 0) { if (mbx == 0) return n; mb = (mby - 1) * c->Wmb + mbx - 1; xN += 16; }      /* :87-92 */
        else mb = (mby - 1) * c->Wmb + mbx;                                                             /* :70-76 */
        yN += 16;
    } else {
        if (mbx == 0) return n;                                                      /* :94 */
        mb = mby * c->Wmb + mbx - 1; xN += 16;
    }
    n.avail = 1;
    if (!c->coded_inter[mb]) { n.mvx = n.mvy = 0; n.same_ref = 0; return n; }         /* :50-54 refIdx -1 */
    int q = (yN >> 3) * 2 + (xN >> 3);                                               /* :100-110 resolved per quadrant */
    n.mvx = c->qmv[mb * 8 + q * 2]; n.mvy = c->qmv[mb * 8 + q * 2 + 1]; n.same_ref = 1;
    return n;
}

static int median3(int a, int b, int c3)
{
    int mn = a < b ? a : b, mx = a > b ? a : b;
    int t = c3 < mx ? c3 : mx;
    return mn > t ? mn : t;                                                          /* mode_pred.cpp:14 */
}

/* PredictMV_Luma for a partition at (px,py) with prediction width pw; dir: 0 none, 1 take B, 2 take A, 3 take C
   first (the 16x8 / 8x16 shortcuts, mode_pred.cpp:279-298). */
static void predict_mv(const fo_ctx *c, int mbx, int mby, int px, int py, int pw, int dir, int cur[4][2], int *ox, int *oy)
{
    fo_nb A = neighbour(c, mbx, mby, px - 1, py, cur);
    fo_nb B = neighbour(c, mbx, mby, px, py - 1, cur);
    fo_nb C = neighbour(c, mbx, mby, px + pw, py - 1, cur);
    if (!C.avail) C = neighbour(c, mbx, mby, px - 1, py - 1, cur);                    /* :264-270 C -> D */
    if (dir == 1 && B.avail && B.same_ref) { *ox = B.mvx; *oy = B.mvy; return; }
    if (dir == 2 && A.avail && A.same_ref) { *ox = A.mvx; *oy = A.mvy; return; }
    if (dir == 3 && C.avail && C.same_ref) { *ox = C.mvx; *oy = C.mvy; return; }
    if (!A.avail && !B.avail) { A.mvx = A.mvy = 0; A.same_ref = 1; A.avail = 1; }      /* :299-302 */
    else if (!A.avail) { A.mvx = A.mvy = 0; A.same_ref = 0; A.avail = 1; }             /* :303-306 */
    if (!B.avail) B = A;                                                             /* :307-310 */
    if (!C.avail) C = A;                                                             /* :311-314 */
    int n = A.same_ref + B.same_ref + C.same_ref;
    if (n == 1) {                                                                    /* :315-329 */
        fo_nb *o = A.same_ref ? &A : (B.same_ref ? &B : &C);
        *ox = o->mvx; *oy = o->mvy; return;
    }
    *ox = median3(A.mvx, B.mvx, C.mvx); *oy = median3(A.mvy, B.mvy, C.mvy);           /* :331-332 */
}

/* P_Skip motion vector (mode_pred.cpp:383-401). */
static void predict_skip_mv(const fo_ctx *c, int mbx, int mby, int *ox, int *oy)
{
    int dummy[4][2] = { { 0 } };
    *ox = *oy = 0;
    if (mby == 0 || mbx == 0) return;
    int up = (mby - 1) * c->Wmb + mbx, left = mby * c->Wmb + mbx - 1;
    if (c->coded_inter[up] && c->qmv[up * 8 + 4] == 0 && c->qmv[up * 8 + 5] == 0) return;      /* quadrant 2 of the MB above */
    if (c->coded_inter[left] && c->qmv[left * 8 + 2] == 0 && c->qmv[left * 8 + 3] == 0) return; /* quadrant 1 of the MB to the left */
    predict_mv(c, mbx, mby, 0, 0, 16, 0, dummy, ox, oy);
}

/* ------------------------------------------------------------------------------------------------
 * Motion compensation of one MB from quadrant MVs (mocomp.cpp:152-208).
 * pred: 256 luma + 64 Cb + 64 Cr.
 * ---------------------------------------------------------------------------------------------- */
static void motion_compensate(const fo_ctx *c, const uint8_t *rY, const uint8_t *rCb, const uint8_t *rCr, int mbx, int mby, int q[4][2], uint8_t *pred)
{
    fo_img L = { rY, c->W, c->H }, U = { rCb, c->W / 2, c->H / 2 }, V = { rCr, c->W / 2, c->H / 2 };
    for (int y = 0; y < 16; y++)
        for (int x = 0; x < 16; x++) {
            const int *mv = q[(y >> 3) * 2 + (x >> 3)];
            pred[y * 16 + x] = (uint8_t)luma_frac(&L, mbx * 16 + x + (mv[0] >> 2), mby * 16 + y + (mv[1] >> 2), mv[0] & 3, mv[1] & 3);
        }
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            const int *mv = q[(y >> 2) * 2 + (x >> 2)];
            int cx = mbx * 8 + x + (mv[0] >> 3), cy = mby * 8 + y + (mv[1] >> 3);
            /* the reference anchors the 3x3 fetch at the 4x4 luma block's chroma origin (mocomp.cpp:165);
               every sample of the 2x2 group ends up at (chroma position + (mv>>3)) either way */
            pred[256 + y * 8 + x] = (uint8_t)chroma_frac(&U, cx, cy, mv[0] & 7, mv[1] & 7);
            pred[320 + y * 8 + x] = (uint8_t)chroma_frac(&V, cx, cy, mv[0] & 7, mv[1] & 7);
        }
}

/* ------------------------------------------------------------------------------------------------
 * Transform / quantisation / reconstruction (quantizationTransform.cpp, scaleTransform.cpp, inttransform.cpp)
 * ---------------------------------------------------------------------------------------------- */
static const int LQ[6][3] = { /* LevelQuantize[qP%6] at (even,even), (mixed), (odd,odd) — quantizationTransform.cpp:24-32 */
    { 205, 158, 128 }, { 186, 146, 114 }, { 158, 128, 102 }, { 146, 114, 89 }, { 128, 102, 82 }, { 114, 89, 71 } };
static const int LS[6][3] = { /* LevelScale[qP%6] — scaleTransform.cpp:32-40 */
    { 160, 208, 256 }, { 176, 224, 288 }, { 208, 256, 320 }, { 224, 288, 368 }, { 256, 320, 400 }, { 288, 368, 464 } };
static inline int pos_class(int i, int j) { return (i & 1) + (j & 1); }
static const int ZZ[16][2] = { /* {row, col}; scaleTransform.cpp:43-47 */
    { 0, 0 }, { 0, 1 }, { 1, 0 }, { 2, 0 }, { 1, 1 }, { 0, 2 }, { 0, 3 }, { 1, 2 }, { 2, 1 }, { 3, 0 }, { 3, 1 }, { 2, 2 }, { 1, 3 }, { 2, 3 }, { 3, 2 }, { 3, 3 } };
static const int QPC[52] = { /* inttransform.cpp:8-14 */
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30,
    31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39 };

/* one pass of the scaled forward kernel (quantizationTransform.cpp:58-77): weights 256 / 416,208 */
static inline void fwd4(int a, int b, int c, int d, int o[4])
{
    o[0] = ((a + b + c + d) * 256 + 512) >> 10;
    o[1] = (416 * a + 208 * b - 208 * c - 416 * d + 512) >> 10;
    o[2] = ((a - b - c + d) * 256 + 512) >> 10;
    o[3] = (208 * a - 416 * b + 416 * c - 208 * d + 512) >> 10;
}

static void forward4x4(int r[4][4], int d[4][4])                 /* quantizationTransform.cpp:41-100 */
{
    int h[4][4], f[4][4], o[4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) h[i][j] = r[i][j] == 0 ? 0 : r[i][j] * 64 - 32;
    for (int j = 0; j < 4; j++) { fwd4(h[0][j], h[1][j], h[2][j], h[3][j], o); for (int i = 0; i < 4; i++) f[i][j] = o[i]; }
    for (int i = 0; i < 4; i++) { fwd4(f[i][0], f[i][1], f[i][2], f[i][3], o); for (int j = 0; j < 4; j++) d[i][j] = o[j]; }
}

static void quant4x4(int d[4][4], int c[4][4], int qP, int keep_dc)   /* quantizationTransform.cpp:183-223 */
{
    int per = qP / 6, rem = qP % 6;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            int lq = LQ[rem][pos_class(i, j)], t;
            if (qP < 24) t = (d[i][j] * (1 << (4 - per)) - (1 << (3 - per))) * lq;
            else t = (d[i][j] >> (per - 4)) * lq;
            c[i][j] = (t + 16384) >> 15;
        }
    if (keep_dc) c[0][0] = d[0][0];
}

static void dequant4x4(int c[4][4], int d[4][4], int qP, int keep_dc) /* scaleTransform.cpp:308-340 */
{
    int per = qP / 6, rem = qP % 6;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            int ls = LS[rem][pos_class(i, j)];
            if (qP >= 24) d[i][j] = (c[i][j] * ls) * (1 << (per - 4));
            else d[i][j] = (c[i][j] * ls + (1 << (3 - per))) >> (4 - per);
        }
    if (keep_dc) d[0][0] = c[0][0];
}

static void inverse4x4(int d[4][4], int r[4][4])                 /* scaleTransform.cpp:101-150 */
{
    int f[4][4], h[4][4];
    for (int i = 0; i < 4; i++) {
        int e0 = d[i][0] + d[i][2], e1 = d[i][0] - d[i][2], e2 = (d[i][1] >> 1) - d[i][3], e3 = d[i][1] + (d[i][3] >> 1);
        f[i][0] = e0 + e3; f[i][1] = e1 + e2; f[i][2] = e1 - e2; f[i][3] = e0 - e3;
    }
    for (int j = 0; j < 4; j++) {
        int g0 = f[0][j] + f[2][j], g1 = f[0][j] - f[2][j], g2 = (f[1][j] >> 1) - f[3][j], g3 = f[1][j] + (f[3][j] >> 1);
        h[0][j] = g0 + g3; h[1][j] = g1 + g2; h[2][j] = g1 - g2; h[3][j] = g0 - g3;
    }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r[i][j] = (h[i][j] + 32) >> 6;
}

/* z-order 4x4 block origin inside the MB (h264_globals.cpp:209-214) */
static inline int blk_x(int b) { return ((b & 1) << 2) | ((b & 4) << 1); }
static inline int blk_y(int b) { return ((b & 2) << 1) | (b & 8); }

/* chroma DC forward + quant (quantizationTransform.cpp:157-178,264-282), inverse + scale (scaleTransform.cpp:247-262,408-420) */
static void chroma_dc_forward(const int dc[4], int qPc, int lvl[4])
{
    int a = dc[0], b = dc[1], c = dc[2], d = dc[3], f[4];
    f[0] = (a + b + c + d + 2) >> 2; f[1] = (a - b + c - d + 2) >> 2; f[2] = (a + b - c - d + 2) >> 2; f[3] = (a - b - c + d + 2) >> 2;
    for (int i = 0; i < 4; i++) lvl[i] = ((((f[i] * 32) >> (qPc / 6)) * LQ[qPc % 6][0]) + 16384) >> 15;
}
static void chroma_dc_inverse(const int lvl[4], int qPc, int dc[4])
{
    int a = lvl[0], b = lvl[1], c = lvl[2], d = lvl[3], f[4];
    f[0] = a + b + c + d; f[1] = a - b + c - d; f[2] = a + b - c - d; f[3] = a - b - c + d;
    for (int i = 0; i < 4; i++) dc[i] = ((f[i] * LS[qPc % 6][0]) * (1 << (qPc / 6))) >> 5;
}

/* Inter macroblock: quantizationTransform(...,reconstruct=true) for a P MB (quantizationTransform.cpp:349-485)
 * followed by its in-loop reconstruction (inttransform.cpp:133-154,237-321).
 * src/pred/recon: 256 Y + 64 Cb + 64 Cr. levels: luma[16][16] (z-order blocks, zigzag), cdc[2][4], cac[2][4][15]. */
void fo_tq_mb(const uint8_t *src, const uint8_t *pred, int qp, int32_t *levels, uint8_t *recon)
{
    int r[4][4], d[4][4], c[4][4], x[4][4], rr[4][4];
    for (int b = 0; b < 16; b++) {
        int x0 = blk_x(b), y0 = blk_y(b);
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r[i][j] = (int)src[(y0 + i) * 16 + x0 + j] - (int)pred[(y0 + i) * 16 + x0 + j];
        forward4x4(r, d); quant4x4(d, c, qp, 0);
        for (int k = 0; k < 16; k++) levels[b * 16 + k] = c[ZZ[k][0]][ZZ[k][1]];
        dequant4x4(c, x, qp, 0); inverse4x4(x, rr);
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) recon[(y0 + i) * 16 + x0 + j] = (uint8_t)clip255(pred[(y0 + i) * 16 + x0 + j] + rr[i][j]);
    }
    int qPc = QPC[clampi(qp, 0, 51)];
    for (int comp = 0; comp < 2; comp++) {
        const uint8_t *s = src + 256 + comp * 64, *p = pred + 256 + comp * 64;
        uint8_t *o = recon + 256 + comp * 64;
        int ac[4][4][4], dc[4], dclvl[4], dcrec[4];
        for (int b = 0; b < 4; b++) {
            int x0 = (b & 1) * 4, y0 = (b >> 1) * 4;
            for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r[i][j] = (int)s[(y0 + i) * 8 + x0 + j] - (int)p[(y0 + i) * 8 + x0 + j];
            forward4x4(r, d); quant4x4(d, ac[b], qPc, 1);
            dc[b] = ac[b][0][0];
            for (int k = 1; k < 16; k++) levels[264 + comp * 60 + b * 15 + (k - 1)] = ac[b][ZZ[k][0]][ZZ[k][1]];
        }
        chroma_dc_forward(dc, qPc, dclvl);
        for (int i = 0; i < 4; i++) levels[256 + comp * 4 + i] = dclvl[i];
        chroma_dc_inverse(dclvl, qPc, dcrec);
        for (int b = 0; b < 4; b++) {
            int x0 = (b & 1) * 4, y0 = (b >> 1) * 4;
            ac[b][0][0] = dcrec[b];
            dequant4x4(ac[b], x, qPc, 1); inverse4x4(x, rr);
            for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) o[(y0 + i) * 8 + x0 + j] = (uint8_t)clip255(p[(y0 + i) * 8 + x0 + j] + rr[i][j]);
        }
    }
}

/* Intra16x16 luma variant: per-block transform with DC kept, 4x4 Hadamard of the 16 DCs
 * (quantizationTransform.cpp:105-152 forward, :227-260 quant; scaleTransform.cpp:154-189,344-376 inverse;
 * inttransform.cpp:157-208 reconstruction). dc_levels[16] zigzag, ac_levels[16][15] z-order blocks. */
void fo_tq_luma_intra16(const uint8_t *src, const uint8_t *pred, int qp, int32_t *dc_levels, int32_t *ac_levels, uint8_t *recon)
{
    int r[4][4], d[4][4], blk[16][4][4], DC[4][4], t[4][4], u[4][4], cq[4][4];
    for (int b = 0; b < 16; b++) {
        int x0 = blk_x(b), y0 = blk_y(b);
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r[i][j] = (int)src[(y0 + i) * 16 + x0 + j] - (int)pred[(y0 + i) * 16 + x0 + j];
        forward4x4(r, d); quant4x4(d, blk[b], qp, 1);
        DC[y0 >> 2][x0 >> 2] = blk[b][0][0];
        for (int k = 1; k < 16; k++) ac_levels[b * 15 + k - 1] = blk[b][ZZ[k][0]][ZZ[k][1]];
    }
    /* forward Hadamard: columns then rows, (x+8)>>4 */
    for (int j = 0; j < 4; j++) {
        int g0 = DC[0][j] + DC[3][j], g1 = DC[1][j] + DC[2][j], g2 = DC[1][j] - DC[2][j], g3 = DC[0][j] - DC[3][j];
        t[0][j] = g0 + g1; t[1][j] = g3 + g2; t[2][j] = g0 - g1; t[3][j] = g3 - g2;
    }
    for (int i = 0; i < 4; i++) {
        int d0 = t[i][0] + t[i][3], d1 = t[i][1] + t[i][2], d2 = t[i][1] - t[i][2], d3 = t[i][0] - t[i][3];
        u[i][0] = (d0 + d1 + 8) >> 4; u[i][1] = (d3 + d2 + 8) >> 4; u[i][2] = (d0 - d1 + 8) >> 4; u[i][3] = (d3 - d2 + 8) >> 4;
    }
    int per = qp / 6, lq = LQ[qp % 6][0], ls = LS[qp % 6][0];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            int tt = qp >= 36 ? (u[i][j] >> (per - 6)) * lq : (u[i][j] * (1 << (6 - per)) - (1 << (5 - per))) * lq;
            cq[i][j] = (tt + 16384) >> 15;
        }
    for (int k = 0; k < 16; k++) dc_levels[k] = cq[ZZ[k][0]][ZZ[k][1]];
    /* inverse Hadamard (rows then columns, no shifts) + scaling */
    for (int i = 0; i < 4; i++) {
        int d0 = cq[i][0] + cq[i][2], d1 = cq[i][0] - cq[i][2], d2 = cq[i][1] - cq[i][3], d3 = cq[i][1] + cq[i][3];
        t[i][0] = d0 + d3; t[i][1] = d1 + d2; t[i][2] = d1 - d2; t[i][3] = d0 - d3;
    }
    for (int j = 0; j < 4; j++) {
        int g0 = t[0][j] + t[2][j], g1 = t[0][j] - t[2][j], g2 = t[1][j] - t[3][j], g3 = t[1][j] + t[3][j];
        u[0][j] = g0 + g3; u[1][j] = g1 + g2; u[2][j] = g1 - g2; u[3][j] = g0 - g3;
    }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            DC[i][j] = qp >= 36 ? (u[i][j] * ls) * (1 << (per - 6)) : (u[i][j] * ls + (1 << (5 - per))) >> (6 - per);
    for (int b = 0; b < 16; b++) {
        int x0 = blk_x(b), y0 = blk_y(b), x[4][4], rr[4][4];
        blk[b][0][0] = DC[y0 >> 2][x0 >> 2];
        dequant4x4(blk[b], x, qp, 1); inverse4x4(x, rr);
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) recon[(y0 + i) * 16 + x0 + j] = (uint8_t)clip255(pred[(y0 + i) * 16 + x0 + j] + rr[i][j]);
    }
}

/* selectNALUnitType's scene-change measure (ref_frames.cpp:210-224): sum |cur - dpb| over luma. */
uint64_t fo_scene_sad(const uint8_t *a, const uint8_t *b, size_t n)
{
    uint64_t s = 0;
    for (size_t i = 0; i < n; i++) s += (uint64_t)iabs((int)a[i] - (int)b[i]);
    return s;
}

/* ------------------------------------------------------------------------------------------------
 * interEncoding for one MB (moestimation.cpp:392-585) + the caller's TQ / skip reconstruction
 * (rbsp_encoding.cpp:175-192). cur planes are updated in place exactly like the reference's `frame`
 * (P_Skip overwrite :214-226, pixel snapping :571-584, reconstruction inttransform.cpp:62-126).
 * ---------------------------------------------------------------------------------------------- */
static void load_mb(const fo_ctx *c, const uint8_t *Y, const uint8_t *U, const uint8_t *V, int mbx, int mby, uint8_t *mb)
{
    for (int y = 0; y < 16; y++) memcpy(mb + y * 16, Y + (size_t)(mby * 16 + y) * c->W + mbx * 16, 16);
    for (int y = 0; y < 8; y++) {
        memcpy(mb + 256 + y * 8, U + (size_t)(mby * 8 + y) * (c->W / 2) + mbx * 8, 8);
        memcpy(mb + 320 + y * 8, V + (size_t)(mby * 8 + y) * (c->W / 2) + mbx * 8, 8);
    }
}
static void store_mb(const fo_ctx *c, uint8_t *Y, uint8_t *U, uint8_t *V, int mbx, int mby, const uint8_t *mb)
{
    for (int y = 0; y < 16; y++) memcpy(Y + (size_t)(mby * 16 + y) * c->W + mbx * 16, mb + y * 16, 16);
    for (int y = 0; y < 8; y++) {
        memcpy(U + (size_t)(mby * 8 + y) * (c->W / 2) + mbx * 8, mb + 256 + y * 8, 8);
        memcpy(V + (size_t)(mby * 8 + y) * (c->W / 2) + mbx * 8, mb + 320 + y * 8, 8);
    }
}

/* optional trace of the candidate lists of one partition (tests of the GPU's phase A/B intermediates) */
typedef struct { int n[3]; int cost[3][33], mvx[3][33], mvy[3][33], sad[3][33]; int mvpx, mvpy, s[5]; int n_stage2_set; } fo_trace;
static fo_trace *g_trace; static int g_trace_mb = -1, g_trace_part = -1;
void fo_set_trace(fo_trace *t, int mb, int part) { g_trace = t; g_trace_mb = mb; g_trace_part = part; }

static void eval_list(fo_ctx *c, fo_list *L, int upto, int need_cost, const uint8_t *curY, int xP, int yP, int mvpx, int mvpy, int *bmin, int *bx, int *by, fo_trace *tr, int stage)
{
    /* moestimation.cpp:460-469 (stage 1: only the MV sentinel is tested) and :498-507,:511-520 */
    for (int j = 0; j <= upto; j++) {
        if (need_cost && !(L->cost[j] < EMPTY_MV)) continue;
        if (!(L->mvx[j] < EMPTY_MV && L->mvy[j] < EMPTY_MV)) continue;
        int sad = sad8x8(c, curY, xP, yP, L->mvx[j], L->mvy[j]);
        if (tr) { int k = tr->n[stage]++; tr->cost[stage][k] = L->cost[j]; tr->mvx[stage][k] = L->mvx[j]; tr->mvy[stage][k] = L->mvy[j]; tr->sad[stage][k] = sad; }
        int tot = sad + iabs(L->mvx[j] - mvpx) + iabs(L->mvy[j] - mvpy);
        if (tot < *bmin) { *bmin = tot; *bx = L->mvx[j]; *by = L->mvy[j]; }
    }
}

static void search_partition(fo_ctx *c, const uint8_t *curY, int mbx, int mby, int part, int mvpx, int mvpy, int window, int basic, int *omx, int *omy)
{
    const int W = c->W;
    const int xP = mbx * 16 + (part & 1) * 8, yP = mby * 16 + (part >> 1) * 8;
    const int genx = mvpx >> 2, geny = mvpy >> 2;
    int s[5] = { 0, 0, 0, 0, 0 };
    for (int ty = 0; ty < 8; ty++)                                               /* :440-451 */
        for (int tx = 0; tx < 8; tx++) {
            int v = curY[(size_t)(yP + ty) * W + xP + tx];
            s[0] += v; if (ty < 4) s[1] += v; if (tx < 4) s[2] += v; if ((ty & 3) < 2) s[3] += v; if ((tx & 3) < 2) s[4] += v;
        }
    fo_trace *tr = (g_trace && g_trace_mb == mby * c->Wmb + mbx && g_trace_part == part) ? g_trace : 0;
    if (tr) { memset(tr, 0, sizeof *tr); tr->mvpx = mvpx; tr->mvpy = mvpy; memcpy(tr->s, s, sizeof s); }
    fo_list L;
    int bx = 0, by = 0, bmin = 2000000000;                                       /* :452,459 */
    list_clear_all(&L);
    feature_search(c, &L, s, xP, yP, window / 16, 1, genx, geny, genx, geny);      /* stage 1 :458 */
    eval_list(c, &L, 16, 0, curY, xP, yP, mvpx, mvpy, &bmin, &bx, &by, tr, 0);
    if (!basic) {
        int tren = 0;
        list_reset_costs(&L);
        for (int j = 0; j <= 180; j++) {                                         /* stage 2 :474-497 */
            for (int side = 0; side < 2; side++) {
                int a = side ? s[0] + j : s[0] - j;
                if (a < 0 || a >= 16384) continue;
                for (int k = c->start[a]; k < c->start[a + 1]; k++) {
                    int dx = c->sorted[2][k] - xP, dy = c->sorted[1][k] - yP;
                    if (iabs(dx) + iabs(dy) < 280 && iabs(c->sorted[3][k] - s[1]) < 100 && iabs(c->sorted[4][k] - s[2]) < 100) {
                        tren++;
                        feature_search(c, &L, s, xP, yP, 0, 16, genx, geny, dx, dy);
                    }
                }
            }
            if (tren > 128) break;
        }
        if (tr) tr->n_stage2_set = tren;
        eval_list(c, &L, 32, 1, curY, xP, yP, mvpx, mvpy, &bmin, &bx, &by, tr, 1);
        list_reset_costs(&L);
        feature_search(c, &L, s, xP, yP, window / 2, 16, 0, 0, 0, 0);              /* stage 3 :509-510 */
        feature_search(c, &L, s, xP, yP, window / 16, 1, 0, 0, 0, 0);
        eval_list(c, &L, 32, 1, curY, xP, yP, mvpx, mvpy, &bmin, &bx, &by, tr, 2);
    }
    *omx = bx; *omy = by;
}

static void encode_mb(fo_ctx *c, uint8_t *Y, uint8_t *U, uint8_t *V, const uint8_t *rY, const uint8_t *rU, const uint8_t *rV,
                      int mbx, int mby, int qp, int window, int maxdiff_set, int basic, int32_t *R)
{
    const int W = c->W, mb = mby * c->Wmb + mbx;
    uint8_t src[384], pred[384], recon[384];
    int q[4][2];
    memset(R, 0, sizeof(int32_t) * FO_REC_INTS);
    load_mb(c, Y, U, V, mbx, mby, src);

    /* P_Skip trial (:402-425) */
    int smx, smy;
    predict_skip_mv(c, mbx, mby, &smx, &smy);
    for (int i = 0; i < 4; i++) { q[i][0] = smx; q[i][1] = smy; }
    motion_compensate(c, rY, rU, rV, mbx, mby, q, pred);
    int maxdiff = maxdiff_set;
    if (maxdiff_set == -1) {                                                      /* :407-419 */
        int sum = 0, dev = 0;
        for (int i = 0; i < 256; i++) sum += src[i];
        int mean = sum / 256;
        for (int i = 0; i < 256; i++) dev += iabs((int)src[i] - mean);
        maxdiff = dev / 256; if (maxdiff < 3) maxdiff = 3;
    }
    int exact = 0;
    for (int i = 0; i < 256; i++) exact += iabs((int)src[i] - (int)pred[i]) <= maxdiff;
    if (exact == 256) {
        R[0] = P_SKIP;
        for (int i = 0; i < 4; i++) { R[1 + 2 * i] = smx; R[2 + 2 * i] = smy; c->qmv[mb * 8 + 2 * i] = smx; c->qmv[mb * 8 + 2 * i + 1] = smy; }
        c->coded_inter[mb] = 1;
        store_mb(c, Y, U, V, mbx, mby, pred);   /* luma := pred (:214-226); recon = pred for Y, Cb, Cr (inttransform.cpp:215-229) */
        return;
    }

    /* 8x8 search, partitions in order, predictor from already decided quadrants (:426-528) */
    int mv[4][2], cur[4][2] = { { 0 } };
    for (int i = 0; i < 4; i++) {
        int mvpx, mvpy;
        predict_mv(c, mbx, mby, (i & 1) * 8, (i >> 1) * 8, 8, 0, cur, &mvpx, &mvpy);
        search_partition(c, Y, mbx, mby, i, mvpx, mvpy, window, basic, &mv[i][0], &mv[i][1]);
        cur[i][0] = mv[i][0]; cur[i][1] = mv[i][1];
    }
    /* SADs of the winners against the unsnapped source (what the search measured) */
    for (int i = 0; i < 4; i++) R[17 + i] = sad8x8(c, Y, mbx * 16 + (i & 1) * 8, mby * 16 + (i >> 1) * 8, mv[i][0], mv[i][1]);

    /* merge (:529-551) and final mvd per partition with that type's predictor (:552-564) */
    int type = P_8x8ref0, nparts = 4;
    int eq01 = mv[0][0] == mv[1][0] && mv[0][1] == mv[1][1], eq23 = mv[2][0] == mv[3][0] && mv[2][1] == mv[3][1];
    int eq02 = mv[0][0] == mv[2][0] && mv[0][1] == mv[2][1], eq13 = mv[1][0] == mv[3][0] && mv[1][1] == mv[3][1];
    if (eq01 && eq23 && eq02) { type = P_L0_16x16; nparts = 1; }
    else if (eq01 && eq23) { type = P_L0_L0_16x8; nparts = 2; }
    else if (eq02 && eq13) { type = P_L0_L0_8x16; nparts = 2; }
    memset(cur, 0, sizeof cur);
    for (int i = 0; i < nparts; i++) {
        int px = 0, py = 0, pw = 16, dir = 0, qsel = i, px2, py2;
        if (type == P_L0_L0_16x8) { py = i * 8; dir = i == 0 ? 1 : 2; qsel = i * 2; }
        else if (type == P_L0_L0_8x16) { px = i * 8; pw = 8; dir = i == 0 ? 2 : 3; qsel = i; }
        else if (type == P_8x8ref0) { px = (i & 1) * 8; py = (i >> 1) * 8; pw = 8; }
        predict_mv(c, mbx, mby, px, py, pw, dir, cur, &px2, &py2);
        R[9 + 2 * i] = mv[qsel][0] - px2; R[10 + 2 * i] = mv[qsel][1] - py2;
        /* quadrants covered by this partition now hold its MV */
        for (int qq = 0; qq < 4; qq++) {
            int in = type == P_L0_16x16 || (type == P_L0_L0_16x8 && (qq >> 1) == i) || (type == P_L0_L0_8x16 && (qq & 1) == i) || (type == P_8x8ref0 && qq == i);
            if (in) { cur[qq][0] = mv[qsel][0]; cur[qq][1] = mv[qsel][1]; }
        }
    }
    R[0] = type;
    for (int i = 0; i < 4; i++) { R[1 + 2 * i] = cur[i][0]; R[2 + 2 * i] = cur[i][1]; c->qmv[mb * 8 + 2 * i] = cur[i][0]; c->qmv[mb * 8 + 2 * i + 1] = cur[i][1]; }
    c->coded_inter[mb] = 1;

    /* final prediction (:565-570) and pixel snapping of the source (:571-584) */
    motion_compensate(c, rY, rU, rV, mbx, mby, cur, pred);
    for (int i = 0; i < 256; i++) if (iabs((int)src[i] - (int)pred[i]) < maxdiff) src[i] = pred[i];
    for (int i = 256; i < 384; i++) if (iabs((int)src[i] - (int)pred[i]) <= maxdiff) src[i] = pred[i];
    fo_tq_mb(src, pred, qp, R + 21, recon);
    store_mb(c, Y, U, V, mbx, mby, recon);
    (void)W;
}

/* One P picture. Y/U/V: current source in, reconstruction out. rY/rU/rV: previous reconstruction (dpb).
 * fo_phase_r(c, rY) must have been run on the same reference. rec: Wmb*Hmb*FO_REC_INTS int32. */
void fo_encode_p(fo_ctx *c, uint8_t *Y, uint8_t *U, uint8_t *V, const uint8_t *rY, const uint8_t *rU, const uint8_t *rV,
                 int qp, int window, int maxdiff_set, int basic, int32_t *rec)
{
    memset(c->coded_inter, 0, (size_t)c->Wmb * c->Hmb);
    c->n_feat_evals = c->n_sads = 0;
    for (int mby = 0; mby < c->Hmb; mby++)
        for (int mbx = 0; mbx < c->Wmb; mbx++)
            encode_mb(c, Y, U, V, rY, rU, rV, mbx, mby, qp, window, maxdiff_set, basic, rec + (size_t)(mby * c->Wmb + mbx) * FO_REC_INTS);
}

/* MC of a whole picture from per-MB quadrant MVs (tests of the GPU MC kernel in isolation). */
void fo_mc_picture(fo_ctx *c, const uint8_t *rY, const uint8_t *rU, const uint8_t *rV, const int32_t *qmv, uint8_t *pred384)
{
    for (int mb = 0; mb < c->Wmb * c->Hmb; mb++) {
        int q[4][2];
        for (int i = 0; i < 4; i++) { q[i][0] = qmv[mb * 8 + 2 * i]; q[i][1] = qmv[mb * 8 + 2 * i + 1]; }
        motion_compensate(c, rY, rU, rV, mb % c->Wmb, mb / c->Wmb, q, pred384 + (size_t)mb * 384);
    }
}

long long fo_stat(const fo_ctx *c, int which) { return which == 0 ? c->n_feat_evals : c->n_sads; }
int fo_rec_ints(void) { return FO_REC_INTS; }
