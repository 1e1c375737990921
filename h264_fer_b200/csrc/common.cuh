// fh264_b200 — shared device-side definitions.
// Data layout in HBM (per sequence): planar u8 pictures (stride == width), 16 quarter-pel luma
// planes, 16 planes of packed box-sum features (16 B per position), a 64x64-tile index of plane-0 positions sorted by (K0>>7, K1>>6),
// per-partition phase-A lists, per-MB motion records and the ABI result records.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include "../../include/fh264_b200.h"

// -DFH_BOUNDS (debug build, `python -m h264_fer_b200.build --bounds`): every index into a shared-memory work array of the search
// kernels is checked and traps (compute-sanitizer is closed on this pool; the parity suite is run on this build once per round).
#ifdef FH_BOUNDS
__device__ __forceinline__ int fh_idx_(int i, int n, int line) { if ((unsigned)i >= (unsigned)n) { printf("FH_BOUNDS: index %d outside [0, %d) at line %d\n", i, n, line); __trap(); } return i; }
#define FH_IDX(i, n) fh_idx_((int)(i), (int)(n), __LINE__)
#else
#define FH_IDX(i, n) (i)
#endif

#define FH_TILE 64              // spatial tile of the stage-2 index
#define FH_TILE_SHIFT 6
#define FH_CELLS 16384          // (K0>>7) * 128 + (K1>>6)
#define FH_TSTART_PITCH 16386   // uint16 per tile: 16384 cell starts + total + pad
#define FH_S3_MAX 33            // list slots evaluated in stages 2/3 (moestimation.cpp:498,511)
#define FH_S1_MAX 17            // list slots evaluated in stage 1 (moestimation.cpp:460)
#define FH_MAX_WINDOW 64
#define FH_COST_EMPTY 100000000 // stages 2/3 ignore list slots with cost >= 1e8 (moestimation.cpp:499,512)

// status word indices (uint32 per sequence)
#define ST_FLAGS 0
#define ST_S2CURSOR 1
#define ST_COUNTS 2   // ..6
#define ST_SAD_LO 8
#define ST_SAD_HI 9
#define ST_TICKET 10
#define ST_FLAGS_NEXT 11   // flags raised by phase R for the NEXT picture's reference
#define ST_S2REDO 12       // number of partitions that overflowed the fast stage-2 launch
#define ST_SPEC_HIT 13     // partitions decided from the speculative finalists (spec.cuh)
#define ST_SPEC_MISS 14    // partitions that took the full search inside the wavefront
#define ST_GATE 16         // scene-change gate of the picture being coded (fh264_encode_p_stream): 1 = sum |cur - dpb| above the IDR threshold
                           // (ref_frames.cpp:210-224) -> every P kernel leaves this sequence alone; 0 otherwise
#define ST_GATE_DONE 17    // the gate as phase C saw it (read by the entropy coder and the status snapshot; the next picture's phase C rewrites it)
#define ST_GATED_TOTAL 18  // pictures of this sequence the gate has stopped since the session was opened
#define ST_NSLOW 19        // partitions of the picture whose stage-2 set phase A could not store (S2_SLOW): the picture takes the block-level phase B
#define ST_WORDS 24
#define FLAG_UB_INPUT 1u
#define FLAG_CAPACITY 2u
#define FLAG_TIMEOUT 4u      // a bounded wavefront / cross-GPU wait gave up (peer rank missing or far behind)

struct Geo {
    int W, H, Wmb, Hmb, nmb, nparts, tilesx, tilesy, ntiles;
    int WH;
    // macroblock-row band of this rank (band mode): MBs [band_mb0, band_mb0 + band_nmb); whole picture otherwise
    int band_mb0, band_nmb, rank, world;
    uint32_t wmb_magic;             // udiv_magic(Wmb): macroblock address -> (x, y) without a division routine in the kernels
    // band mode with the peers' bands known (fh264_band_peers): luma rows [halo_y0, halo_y1) (multiples of 64, or H) are everything
    // this rank's phases A / S / B / C read of the reference picture (stage 2 reaches 279 rows); phase R is restricted to them and the
    // picture barrier only waits for the ranks in wait_mask, whose bands they touch. Whole picture / all ranks otherwise.
    int halo_y0, halo_y1;
    uint32_t wait_mask;
    int gather_on;                  // band mode: phase C also stores every record into rank 0's gather buffer (fh264_band_gather)
};
#define FH_MAX_WORLD 8

struct __align__(16) TileEntry { uint16_t x, y, k0, k1, k2, k3, k4, pad; };

struct __align__(16) MbMotion {       // phase B output per macroblock (48 bytes)
    int16_t mb_type, num_parts;
    int16_t mv[4][2];
    int16_t mvd[4][2];
    uint16_t sad[4];
    int16_t maxdiff, pad;
};

struct S3Entry { int16_t mvx, mvy; uint16_t sad, pad; };           // stage-3 list slot, list order
struct PartA { uint16_t suma[5]; uint16_t n3; uint32_t s2_off; uint32_t n2; };   // per 8x8 partition
// n2 with this bit: the stage-2 candidate set up to j_stop (low 8 bits) does not fit the phase-A buffers (flat / low-contrast
// content: thousands of positions share one 8x8 sum); phase B enumerates it itself once the predictor is known (stage2_slow).
#define S2_SLOW 0x80000000u

struct PartSpec;                     // spec.cuh: finalists per partition for the guessed integer predictors
struct MbSpec;                       // spec.cuh: P_Skip trial results per macroblock for the guessed skip vectors

static_assert(sizeof(fh264_mb_result) == 832, "ABI record size");
static_assert(sizeof(MbMotion) == 48, "MbMotion size");

struct SeqDev {
    uint8_t *cur[3];        // `frame`: source picture (Y, Cb, Cr)
    uint8_t *cur_alt[3];    // the other source buffer: the NEXT picture is uploaded here while this one is being coded
    uint8_t *ref[3];        // `dpb`: previous reconstruction
    uint8_t *rec[3];        // reconstruction of the picture being coded (swapped with ref afterwards)
    uint8_t *planes;        // refFrameInterpolated[f].L, f-major, WH each (+16 bytes slack at the end)
    uint16_t *k0p;          // 8x8 sums K0 of plane 0 per position (the first word of the feature distance: lower bound for stage 3's sweep)
    uint4 *kar;             // refFrameKar[0..4][f] per position: [f][y][x] = {F1|F2<<16, F3|F4<<16, K0, 0}, Fk = K0 - 2*Kk (int16)
    TileEntry *tent;        // ntiles * 4096 entries
    uint16_t *tstart;       // ntiles * FH_TSTART_PITCH
    PartA *parta;           // nparts
    S3Entry *s3;            // nparts * 33
    uint4 *s2pool;          // stage-2 candidates, a slice of S2_SLICE entries per partition, unordered: {dx | dy<<16, feature distance, lower bound of the SAD, arrival key}
    MbMotion *motion;       // nmb
    unsigned long long *qmv; // nmb * 4 tagged quadrant words (phase B wavefront): epoch << 32 | mvy << 16 | (mvx & 0xffff)
    PartSpec *spec;         // nparts: speculative finalists of the 8x8 search (phase S, read by phase B)
    MbSpec *mbspec;         // nmb: precomputed P_Skip trials (phase S, read by phase B)
    uint32_t *prev_gen16;   // nmb: 16x16 predictor >> 2 of the macroblock's P_Skip trial in the previous P picture
    uint32_t *proxy;        // nparts: best stage-3 vector by SAD (mvx & 0xffff | mvy << 16; SPEC_PREV_NONE: none) — stands in for the final MV when phase S guesses predictors
    uint32_t *prev_gen;     // nparts: mvp >> 2 phase B used for the partition in the previous P picture (genx & 0xffff | geny << 16)
    fh264_mb_result *results;
    uint32_t *status;       // ST_WORDS
    // band mode: the same buffers of the other ranks, mapped through CUDA IPC (NVLink peer access)
    uint8_t *peer_ref[FH_MAX_WORLD][3], *peer_rec[FH_MAX_WORLD][3];
    unsigned long long *peer_qmv_next;   // rank + 1: mirror of the band's last MB row (tagged quadrant words)
    fh264_mb_result *gather[2];          // band mode: rank 0's copy of the WHOLE picture's records, one buffer per epoch parity — phase C of
                                         // every rank stores its macroblocks' records there over NVLink, so rank 0 can entropy-code the slice
    MbMotion *peer_motion[FH_MAX_WORLD]; // every rank's motion records: phase C mirrors each macroblock's mb_type there (an I picture, coded
                                         // whole by every rank, needs to know which macroblocks of the previous P picture were P_Skip)
    long long *dbg;         // optional: nmb * 12 clock samples of phase B (fh264_debug_timeline), else null
};

__device__ __forceinline__ int iabs_(int a) { return a < 0 ? -a : a; }
__device__ __forceinline__ int clampi_(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int clip255_(int v) { return min(max(v, 0), 255); }
// mocomp.cpp:39-40,47
__device__ __forceinline__ int tap6_(int a, int b, int c, int d, int e, int f) { return clip255_((a - 5 * b + 20 * c + 20 * d - 5 * e + f + 16) >> 5); }
__device__ __forceinline__ int mid_(int a, int b) { return (a + b + 1) >> 1; }

__device__ __forceinline__ int median3_(int a, int b, int c) { return max(min(a, b), min(c, max(a, b))); }

// Median prediction from three neighbour candidates with availability flags (mode_pred.cpp:299-332; every available
// neighbour of a P picture is inter with refIdx 0, so "same reference" == available, except the A := 0 substitutions).
__device__ __forceinline__ void median_pred(int aA, int ax, int ay, int aB, int bx, int by, int aC, int cx, int cy, int &ox, int &oy)
{
    int sa = aA, sb = aB, sc = aC;
    if (!aA && !aB) { ax = ay = 0; sa = 1; }
    else if (!aA) { ax = ay = 0; sa = 0; }
    if (!aB) { bx = ax; by = ay; sb = sa; }
    if (!aC) { cx = ax; cy = ay; sc = sa; }
    if (sa + sb + sc == 1) { ox = sa ? ax : (sb ? bx : cx); oy = sa ? ay : (sb ? by : cy); return; }
    ox = median3_(ax, bx, cx); oy = median3_(ay, by, cy);
}

// 8 consecutive bytes starting at an arbitrary address, as two little-endian words: two aligned 8-byte loads (two LSU
// wavefronts per lane instead of three 4-byte ones) and a funnel shift; reads up to 7 bytes past the block (the plane
// buffers carry 16 bytes of slack).
__device__ __forceinline__ uint2 load8_unaligned(const uint8_t *p)
{
    const uintptr_t a = (uintptr_t)p;
    const uint2 *q = (const uint2 *)(a & ~(uintptr_t)7);
    const uint32_t sh = (uint32_t)(a & 7) * 8;
    const uint2 lo = __ldg(q), hi = __ldg(q + 1);
    const bool b = sh >= 32;
    const uint32_t w0 = b ? lo.y : lo.x, w1 = b ? hi.x : lo.y, w2 = b ? hi.y : hi.x;
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));       // shift amount is taken mod 32
}

// satdLuma8x8MVs (moestimation.cpp:175-195), one row of the reference block with the reference's clamping rule
// (block origin clamped to the picture as a whole, then each index clamped at the right/bottom edge only).
// (the right-edge case is kept out of line: it sits in unrolled SAD loops whose instruction-cache footprint matters)
__device__ __noinline__ uint2 load_row8_edge(const uint8_t *row, int W, int x0)
{
    uint32_t b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = row[min(x0 + i, W - 1)];
    return make_uint2(b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24), b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24));
}
__device__ __forceinline__ uint2 load_row8(const uint8_t *plane, int W, int H, int x0, int y)
{
    y = min(y, H - 1);
    if (x0 + 8 <= W) return load8_unaligned(plane + (size_t)y * W + x0);
    return load_row8_edge(plane + (size_t)y * W, W, x0);
}
__device__ __forceinline__ int sad8(uint2 a, uint2 b) { return __vsadu4(a.x, b.x) + __vsadu4(a.y, b.y); }
__device__ __forceinline__ int sad_row8(uint2 cur, const uint8_t *plane, int W, int H, int x0, int y) { return sad8(cur, load_row8(plane, W, H, x0, y)); }

// Exact n / d by one IMAD.HI for n * d < 2^32: magic = 2^32 / d + 1; d == 1 has no 32-bit magic and is marked by 0.
__host__ __device__ __forceinline__ uint32_t udiv_magic(uint32_t d) { return d > 1u ? 0xffffffffu / d + 1u : 0u; }
__device__ __forceinline__ int udiv_by(int n, uint32_t magic) { return magic ? (int)__umulhi((uint32_t)n, magic) : n; }

__device__ __forceinline__ void mb_xy(const Geo &g, int mb, int &mbx, int &mby) { mby = udiv_by(mb, g.wmb_magic); mbx = mb - mby * g.Wmb; }

// n / d for 0 <= n < 65536 / d with inv = 65536 / d + 1 (window geometry: d <= 129)
__device__ __forceinline__ int fdiv_(int n, int inv) { return (int)(((unsigned)n * (unsigned)inv) >> 16); }

// Feature distance (moestimation.cpp:267-276):
//   |s0-K0| + sum_k ( |sk-Kk| + |(s0-sk)-(K0-Kk)| ),  k = 1..4.
// With d = s0-K0 and a = sk-Kk the k-th pair is |a| + |d-a| = max(|d|, |2a-d|) (the identity |x|+|y| = max(|x+y|,|x-y|)),
// an exact integer rewrite that needs one abs + one max per pair instead of two abs: 2a-d = (2sk-s0) + K0 - 2Kk.
__device__ __forceinline__ int feat_dist(const int s[5], int K0, int K1, int K2, int K3, int K4)
{
    const int ad = iabs_(s[0] - K0);
    int d = ad;
    d += max(ad, iabs_((2 * s[1] - s[0]) + K0 - 2 * K1));
    d += max(ad, iabs_((2 * s[2] - s[0]) + K0 - 2 * K2));
    d += max(ad, iabs_((2 * s[3] - s[0]) + K0 - 2 * K3));
    d += max(ad, iabs_((2 * s[4] - s[0]) + K0 - 2 * K4));
    return d;
}

// The same distance on the packed feature record {F1|F2<<16, F3|F4<<16, K0, -} with Fk = K0 - 2*Kk (|Fk| <= 8160):
// the k-th pair is max(|d|, |ck + Fk|) with ck = 2sk - s0, evaluated two at a time in 16-bit lanes
// (VIADDMNMX.S16x2: max(a + b, c); -(c + F) = ~F + (1 - c) per lane). All lane values stay within +-16320.
struct FeatQ { uint32_t c12, c34, n12, n34; int s0; };
__device__ __forceinline__ FeatQ feat_query(const int s[5])
{
    FeatQ q;
    const int c1 = 2 * s[1] - s[0], c2 = 2 * s[2] - s[0], c3 = 2 * s[3] - s[0], c4 = 2 * s[4] - s[0];
    q.c12 = ((uint32_t)c1 & 0xffffu) | ((uint32_t)c2 << 16); q.c34 = ((uint32_t)c3 & 0xffffu) | ((uint32_t)c4 << 16);
    q.n12 = ((uint32_t)(1 - c1) & 0xffffu) | ((uint32_t)(1 - c2) << 16); q.n34 = ((uint32_t)(1 - c3) & 0xffffu) | ((uint32_t)(1 - c4) << 16);
    q.s0 = s[0];
    return q;
}
__device__ __forceinline__ int feat_of(const FeatQ &q, const uint4 v)
{
    const int ad = iabs_(q.s0 - (int)v.z);
    const uint32_t dd = (uint32_t)ad * 0x10001u;
    uint32_t m12 = __vmaxs2(__vadd2(v.x, q.c12), dd), m34 = __vmaxs2(__vadd2(v.y, q.c34), dd);
    m12 = __vmaxs2(__vadd2(~v.x, q.n12), m12); m34 = __vmaxs2(__vadd2(~v.y, q.n34), m34);
    return (int)__dp2a_lo(__vadd2(m12, m34), 0x0101u, (uint32_t)ad);
}
__device__ __forceinline__ uint4 feat_record(int k0, int k1, int k2, int k3, int k4)
{
    return make_uint4(((uint32_t)(k0 - 2 * k1) & 0xffffu) | ((uint32_t)(k0 - 2 * k2) << 16),
                      ((uint32_t)(k0 - 2 * k3) & 0xffffu) | ((uint32_t)(k0 - 2 * k4) << 16), (uint32_t)k0, 0u);
}
// raw box sums K1..K4 back from a record (index build, debug taps)
__device__ __forceinline__ int feat_raw(const uint4 v, int k)
{
    const int k0 = (int)v.z;
    if (k == 0) return k0;
    const uint32_t w = k <= 2 ? v.x : v.y;
    const int F = (k & 1) ? (int)(int16_t)(w & 0xffffu) : (int)(int16_t)(w >> 16);
    return (k0 - F) >> 1;
}

// suma[0..4] of one 8x8 source block held as 8 rows of two words (moestimation.cpp:440-451):
// all, rows 0-3, columns 0-3, rows {0,1,4,5}, columns {0,1,4,5}.
__device__ __forceinline__ void block_sums(const uint2 rows[8], int s[5])
{
    s[0] = s[1] = s[2] = s[3] = s[4] = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        int lo = __vsadu4(rows[r].x, 0), hi = __vsadu4(rows[r].y, 0);
        int c01 = __vsadu4(rows[r].x & 0x0000ffffu, 0) + __vsadu4(rows[r].y & 0x0000ffffu, 0);
        s[0] += lo + hi;
        if (r < 4) s[1] += lo + hi;
        s[2] += lo;
        if ((r & 3) < 2) s[3] += lo + hi;
        s[4] += c01;
    }
}
