// Phase B — the predictor-dependent part of interEncoding (moestimation.cpp:392-570), a wavefront over
// macroblocks: every cost uses the median MV predictor of the already decided left / up / up-right / up-left
// neighbours (mode_pred.cpp:252-371). One CTA (128 threads) per macroblock; persistent CTAs draw tickets in anti-diagonal
// order (x + 3y), interleaved over the sequences of the batch, and poll the tagged quadrant words of exactly the
// neighbours their predictors read. A ticket's dependencies always hold smaller tickets, so the smallest unfinished
// ticket can always run: no co-residency assumption, no deadlock; every wait is bounded.
// Per MB: P_Skip test (mode_pred.cpp:383-401, moestimation.cpp:402-425), then per 8x8 partition, by the whole block:
// predictor; stage 1 (window/16 quarter-pel window around the predictor, features computed from the planes, 17 best by
// feature cost -> SAD); the phase-A stage-2 set ranked with the now known multiplier (lazily: the best candidate is verified
// to be among the 33 smallest keys; oversized sets are enumerated here, stage2_slow); the phase-A stage-3 list;
// winner = first strict minimum of SAD + |mv - mvp|_1 in list order; then merge and final mvd (:529-564).
#pragma once
#include "common.cuh"
#include "warp_select.cuh"
#include "qfeat.cuh"
#include "phase_a.cuh"
#include "phase_c.cuh"
#include "spec.cuh"

#define PB_NT 128
#ifndef PB_MINB
#define PB_MINB 4            // resident CTAs per SM the register budget is set for (4 -> 128 registers)
#endif
#define S1_KEY_CAP 1296     // (2*4+1)^2*16 at WindowSize 64
#define PB_POOL_PREF 192    // stage-2 candidates per partition staged in shared memory (the rest is read from HBM)

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// Bounded wait until *p >= target (flags written by another CTA, or in band mode by another GPU). Returns false on timeout.
#define FH_WAIT_NS 2000000000ll          // every cross-CTA / cross-GPU wait gives up after 2 s of %globaltimer
__device__ __forceinline__ long long gtime_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ bool wait_progress(const uint32_t *p, uint32_t target, bool sys)
{
    long long t0 = 0;
    for (unsigned it = 0;; it++) {
        const uint32_t v = sys ? ld_acquire_sys_u32(p) : ld_acquire_u32(p);
        if (v >= target) return true;
        if ((it & 1023u) == 1023u) { const long long t = gtime_ns(); if (t0 == 0) t0 = t; else if (t - t0 > FH_WAIT_NS) return false; }
        __nanosleep(20);
    }
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// Wavefront hand-off. A quadrant's final MV travels in ONE naturally aligned 64-bit word together with the epoch of the
// picture it belongs to (epoch << 32 | mvy << 16 | mvx & 0xffff). An aligned 8-byte access is single-copy atomic, so the
// reader that sees the tag has the value: no release fence at the producer, no second round trip at the consumer.
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p, bool sys)
{
    unsigned long long v;
    if (sys) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v, bool sys)
{
    if (sys) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    else asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long qmv_word(uint32_t epoch, int mvx, int mvy) { return ((unsigned long long)epoch << 32) | (uint32_t)((mvx & 0xffff) | (mvy << 16)); }
// Bounded wait for a quadrant word of this picture (written by another CTA, or in band mode by another GPU).
__device__ __forceinline__ bool wait_qmv(const unsigned long long *p, uint32_t epoch, bool sys, int &mvx, int &mvy)
{
    long long t0 = 0;
    for (unsigned it = 0;; it++) {
        const unsigned long long v = ld_relaxed_u64(p, sys);
        if ((uint32_t)(v >> 32) == epoch) { mvx = (int)(int16_t)(v & 0xffffu); mvy = (int)(int16_t)((v >> 16) & 0xffffu); return true; }
        if ((it & 1023u) == 1023u) { const long long t = gtime_ns(); if (t0 == 0) t0 = t; else if (t - t0 > FH_WAIT_NS) break; }
        __nanosleep(20);
    }
    mvx = mvy = 0;
    return false;
}

struct NbCache {            // quadrant MVs of the four neighbouring macroblocks (A.7): 0 left, 1 up, 2 up-right, 3 up-left
    int mvx[4][4], mvy[4][4];
    int avail[4];
};

struct Nb { int avail, mvx, mvy; };

// DeriveNeighbourLocation + get_neighbour_mv (mode_pred.cpp:48-97), quadrant formulation.
__device__ __forceinline__ Nb neighbour_mv(const NbCache &nc, int xN, int yN, const int cur[4][2])
{
    Nb n = { 0, 0, 0 };
    if ((xN > 15 && yN >= 0) || yN > 15) return n;
    if (xN >= 0 && xN < 16 && yN >= 0) { const int q = (yN >> 3) * 2 + (xN >> 3); n.avail = 1; n.mvx = cur[q][0]; n.mvy = cur[q][1]; return n; }
    int w;
    if (yN < 0) { w = xN > 15 ? 2 : (xN < 0 ? 3 : 1); if (xN > 15) xN -= 16; else if (xN < 0) xN += 16; yN += 16; }
    else { w = 0; xN += 16; }
    if (!nc.avail[w]) return n;
    const int q = (yN >> 3) * 2 + (xN >> 3);
    n.avail = 1; n.mvx = nc.mvx[w][q]; n.mvy = nc.mvy[w][q];
    return n;
}


// PredictMV_Luma (mode_pred.cpp:252-332). Every available neighbour of a P picture is inter with refIdx 0.
// dir: 0 median, 1 prefer B (16x8 top), 2 prefer A (16x8 bottom / 8x16 left), 3 prefer C (8x16 right).
__device__ __forceinline__ void predict_mv_(const NbCache &nc, int px, int py, int pw, int dir, const int cur[4][2], int &ox, int &oy)
{
    Nb A = neighbour_mv(nc, px - 1, py, cur), B = neighbour_mv(nc, px, py - 1, cur), C = neighbour_mv(nc, px + pw, py - 1, cur);
    if (!C.avail) C = neighbour_mv(nc, px - 1, py - 1, cur);
    if (dir == 1 && B.avail) { ox = B.mvx; oy = B.mvy; return; }
    if (dir == 2 && A.avail) { ox = A.mvx; oy = A.mvy; return; }
    if (dir == 3 && C.avail) { ox = C.mvx; oy = C.mvy; return; }
    int sa = A.avail, sb = B.avail, sc = C.avail;       // "same reference" flags
    if (!A.avail && !B.avail) { A.mvx = A.mvy = 0; sa = 1; A.avail = 1; }         // :299-302
    else if (!A.avail) { A.mvx = A.mvy = 0; sa = 0; A.avail = 1; }                // :303-306
    if (!B.avail) { B = A; sb = sa; }                                             // :307-310
    if (!C.avail) { C = A; sc = sa; }                                             // :311-314
    if (sa + sb + sc == 1) { const Nb &o = sa ? A : (sb ? B : C); ox = o.mvx; oy = o.mvy; return; }
    ox = median3_(A.mvx, B.mvx, C.mvx); oy = median3_(A.mvy, B.mvy, C.mvy);
}

__device__ __forceinline__ int mv_cost(int mvx, int mvy, int px, int py) { return iabs_(mvx - px) + iabs_(mvy - py); }

// Which phase-B kernel codes the current picture of a sequence: the block-level one (this file) when a partition's stage-2 set
// has to be enumerated here (S2_SLOW) or the timeline tap is on, the warp-level one (phase_bw.cuh) otherwise. Both kernels are
// launched for every call and evaluate this the same way.
__device__ __forceinline__ bool seq_is_heavy(const SeqDev &S) { return S.status[ST_NSLOW] != 0u || S.dbg != nullptr; }


// ---- the whole macroblock from the phase-S products, by ONE warp, in registers (no block barrier, no shared-memory state) -------
// Every decision interEncoding makes for a macroblock (:402-564) once the predictors are known is a lookup: the P_Skip trial in
// the phase-S masks, each partition's winner among its finalists, merge and mvd from the four winners. The left macroblock's
// quadrants are polled by lane 0 exactly where a predictor really depends on them (same rules as the block-level path below).
// Returns false — without having published anything that the block-level path would not publish identically — as soon as a
// lookup cannot answer (predictor outside the guessed cells, oversized finalist set); the caller then runs the full search.
struct FastNb { int aL, aU, aUR, aUL, u2x, u2y, u3x, u3y, r2x, r2y, d3x, d3y; };
__device__ __forceinline__ int spec_lookup(const PartSpec &sp, int genx, int geny, int mvpx, int mvpy, int &bx, int &by, int &bs)
{
    int slot = -1;
    if (sp.nf[0] != SPEC_INVALID && sp.gx[0] == genx && sp.gy[0] == geny) slot = 0;
    else if (sp.nf[1] != SPEC_INVALID && sp.gx[1] == genx && sp.gy[1] == geny) slot = 1;
    if (slot < 0) return 0;
    const int nf = sp.nf[slot];
    u64 bk = KEY_NONE;
    for (int k = 0; k < nf; k++) {
        const SpecFinal f = sp.f[slot][k];
        const u64 key = ((u64)((int)f.sad + mv_cost(f.mvx, f.mvy, mvpx, mvpy)) << 32) | ((u64)f.order << 16) | (u64)k;
        bk = min(bk, key);
    }
    const SpecFinal f = sp.f[slot][(int)(bk & 15)];
    bx = f.mvx; by = f.mvy; bs = f.sad;
    return 1;
}
// The same lookup by a whole warp (all 32 lanes call it with the same arguments): lane k weighs finalist k, the winner is one
// 32-bit warp minimum of total << 11 | stage << 9 | list position << 3 | k (total < 2^15: SAD <= 16320 plus two vector distances;
// list positions < 64) — the order of (total, order, k). The wavefront's critical path is a chain of these lookups: as scalar
// code by one lane it cost 0.3-0.7 us per partition (profiles/tools/phase_b_waits.py).
__device__ __forceinline__ int spec_lookup_w(const PartSpec &sp, int genx, int geny, int mvpx, int mvpy, int &bx, int &by, int &bs)
{
    const int lane = threadIdx.x & 31;
    const uint4 h = *(const uint4 *)&sp;                   // gx[2] | gy[2] | nf[2] | pad
    const int gx0 = (int)(int16_t)(h.x & 0xffffu), gx1 = (int)(int16_t)(h.x >> 16), gy0 = (int)(int16_t)(h.y & 0xffffu), gy1 = (int)(int16_t)(h.y >> 16);
    const int nf0 = (int)(h.z & 0xffu), nf1 = (int)((h.z >> 8) & 0xffu);
    int slot = -1;
    if (nf0 != SPEC_INVALID && gx0 == genx && gy0 == geny) slot = 0;
    else if (nf1 != SPEC_INVALID && gx1 == genx && gy1 == geny) slot = 1;
    if (slot < 0) return 0;
    const int nf = slot ? nf1 : nf0;
    uint32_t key = 0xffffffffu;
    if (lane < nf) {
        const SpecFinal f = sp.f[slot][lane];
        const uint32_t total = (uint32_t)((int)f.sad + mv_cost(f.mvx, f.mvy, mvpx, mvpy));
        key = (total << 11) | ((uint32_t)(f.order >> 14) << 9) | ((uint32_t)(f.order & 63u) << 3) | (uint32_t)lane;
    }
    key = __reduce_min_sync(0xffffffffu, key);
    const SpecFinal f = sp.f[slot][(int)(key & 7u)];
    bx = f.mvx; by = f.mvy; bs = f.sad;
    return 1;
}
__device__ __forceinline__ int skip_lookup(const MbSpec &ms, int vx, int vy)       // 0: skips, 1: does not, -1: not precomputed
{
    if (vx == 0 && vy == 0) return ms.zero_ok ? 0 : 1;
    const int cxx = vx >> 2, cyy = vy >> 2, f = (vy & 3) * 4 + (vx & 3);
    if (ms.cx[0] == cxx && ms.cy[0] == cyy) return ((ms.mask[0] >> f) & 1) ? 0 : 1;
    if (ms.cx[1] == cxx && ms.cy[1] == cyy) return ((ms.mask[1] >> f) & 1) ? 0 : 1;
    return -1;
}

struct BlockSel { u64 skey[256]; uint16_t sidx[256]; int ns[2]; int phase; };

struct PBShared {
    uint32_t keys1[S1_KEY_CAP];          // stage-1 keys  cost << 11 | arrival index
    u64 keys2[1024];                     // stage-2 keys  cost << 10 | arrival rank
    BlockSel bs;
    uint16_t mem1[FH_S1_MAX + 3], mem2[FH_S3_MAX + 3];
    __align__(16) uint8_t cur[16][16];
    NbCache nc;
    u64 best[2][4];                      // per-warp minima, double-buffered across partitions
    u64 sel_min[4]; int sel_cnt[4];      // lazy stage-2 verification
    __align__(16) PartSpec spec[4];      // phase-S finalists of the four partitions (spec.cuh)
    __align__(16) MbSpec ms;             // phase-S P_Skip trials of the macroblock
    PartA pa[4];                         // phase-A products of the four partitions, fetched BEFORE the dependency wait
    S3Entry s3[4][FH_S3_MAX + 1];
    uint32_t qx[16 * 61];                // qfeat.cuh step A: row sums of the stage-1 window (16 planes x 12 rows x 5 positions at WindowSize 32)
    uint16_t qrc[16 * 61];
    uint32_t my_ticket;
    int fast_done;                       // the warp-level fast path decided the whole macroblock
    int red[4];
    // stage2_slow
    int slow_ns, slow_total, slow_arg[12];
    u64 slow_thr;
    u64 slow_key[FH_S3_MAX + 1];
    uint32_t slow_pos[FH_S3_MAX + 1];        // (dx + 512) | (dy + 512) << 16 of the selected candidates, list order
};

// Block-cooperative selection: the K = min(k, #valid) smallest of
// keyfn(0..n-1) in ascending order into members[]. EVERY warp scans all keys (n/32 per lane), so all warps derive the
// same tight upper bound on the k-th key without exchanging anything; the survivors are then compacted and ranked by
// all PB_NT threads. `call` alternates the survivor counter so that no reset barrier is needed. Ends with a barrier.
template <typename KeyFn>
__device__ __forceinline__ int block_select_smallest(int n, int k, KeyFn keyfn, BlockSel *bs, uint16_t *members, int call)
{
    const int tid = threadIdx.x, lane = tid & 31;
    u64 m1 = KEY_NONE, m2 = KEY_NONE;
    int nv = 0;
    for (int i = lane; i < n; i += 32) {
        const u64 key = keyfn(i);
        if (key != KEY_NONE) {
            nv++;
            const u64 lo = key < m1 ? key : m1, hi = key < m1 ? m1 : key;
            m1 = lo; m2 = hi < m2 ? hi : m2;
        }
    }
    const int K = min(k, __reduce_add_sync(0xffffffffu, nv));
    if (K == 0) return 0;
    const int L1 = __popc(__ballot_sync(0xffffffffu, m1 != KEY_NONE));
    const int L2 = __popc(__ballot_sync(0xffffffffu, m2 != KEY_NONE));
    u64 thr = KEY_NONE - 1;
    if (L1 >= K) thr = warp_max_u64(m1 != KEY_NONE ? m1 : 0ull);
    else if (L1 + 1 >= K && L2 >= 1) { const u64 a = warp_max_u64(m1 != KEY_NONE ? m1 : 0ull), b = warp_min_u64(m2); thr = a > b ? a : b; }
    else if (2 * L2 >= K) thr = warp_max_u64(m2 != KEY_NONE ? m2 : 0ull);
    int *nsp = &bs->ns[call & 1];
    for (int i = tid; i < n; i += PB_NT) {
        const u64 key = keyfn(i);
        if (key <= thr) {
            const int pos = atomicAdd(nsp, 1);
            if (pos < 256) { bs->skey[FH_IDX(pos, 256)] = key; bs->sidx[pos] = (uint16_t)i; }
        }
    }
    __syncthreads();
    const int ns = *nsp;
    if (tid == 0) bs->ns[(call & 1) ^ 1] = 0;          // the other counter is idle until the next call
    if (ns <= 64) {
        // G threads per survivor split the comparisons, partial ranks are added by shuffles
        const int G = ns <= 32 ? 4 : 2, sidx = tid / G, part = tid - sidx * G;
        int rank = 0;
        u64 key = 0;
        if (sidx < ns) { key = bs->skey[sidx]; for (int j = part; j < ns; j += G) rank += bs->skey[j] < key; }
        rank += __shfl_xor_sync(0xffffffffu, rank, 1);
        if (G == 4) rank += __shfl_xor_sync(0xffffffffu, rank, 2);
        if (sidx < ns && part == 0 && rank < K) members[rank] = bs->sidx[sidx];
    } else if (ns <= 256) {
        for (int sidx = tid; sidx < ns; sidx += PB_NT) {
            const u64 key = bs->skey[sidx];
            int rank = 0;
            for (int j = 0; j < ns; j++) rank += bs->skey[j] < key;
            if (rank < K) members[rank] = bs->sidx[sidx];
        }
    } else {
        for (int i = tid; i < n; i += PB_NT) {
            const u64 key = keyfn(i);
            if (key > thr) continue;
            int rank = 0;
            for (int j = 0; j < n && rank < K; j++) rank += keyfn(j) < key;
            if (rank < K) members[rank] = (uint16_t)i;
        }
    }
    __syncthreads();
    return K;
}

// 32-bit-key variant for stage 1 (keys = cost << 11 | arrival index, unique; 0xffffffff = no candidate): single-instruction
// warp reductions and half the compare cost of the 64-bit path. Same contract as block_select_smallest.
__device__ __forceinline__ int block_select_smallest_u32(const uint32_t *keys, int n, int k, BlockSel *bs, uint16_t *members, int call)
{
    const int tid = threadIdx.x, lane = tid & 31;
    uint32_t m1 = 0xffffffffu, m2 = 0xffffffffu;
    int nv = 0;
    for (int i = lane; i < n; i += 32) {
        const uint32_t key = keys[i];
        nv += key != 0xffffffffu;
        m2 = min(m2, max(m1, key));
        m1 = min(m1, key);
    }
    const int K = min(k, __reduce_add_sync(0xffffffffu, nv));
    if (K == 0) return 0;
    const int L1 = __popc(__ballot_sync(0xffffffffu, m1 != 0xffffffffu));
    const int L2 = __popc(__ballot_sync(0xffffffffu, m2 != 0xffffffffu));
    uint32_t thr = 0xfffffffeu;
    if (L1 >= K) thr = __reduce_max_sync(0xffffffffu, m1 != 0xffffffffu ? m1 : 0u);
    else if (L1 + 1 >= K && L2 >= 1) thr = max(__reduce_max_sync(0xffffffffu, m1 != 0xffffffffu ? m1 : 0u), __reduce_min_sync(0xffffffffu, m2));
    else if (2 * L2 >= K) thr = __reduce_max_sync(0xffffffffu, m2 != 0xffffffffu ? m2 : 0u);
    int *nsp = &bs->ns[call & 1];
    uint32_t *skey = (uint32_t *)bs->skey;
    for (int i = tid; i < n; i += PB_NT) {
        const uint32_t key = keys[i];
        if (key <= thr) {
            const int pos = atomicAdd(nsp, 1);
            if (pos < 256) { skey[pos] = key; bs->sidx[pos] = (uint16_t)i; }
        }
    }
    __syncthreads();
    const int ns = *nsp;
    if (tid == 0) bs->ns[(call & 1) ^ 1] = 0;
    if (ns <= 64) {
        const int G = ns <= 32 ? 4 : 2, sidx = tid / G, part = tid - sidx * G;
        int rank = 0;
        if (sidx < ns) { const uint32_t key = skey[sidx]; for (int j = part; j < ns; j += G) rank += skey[j] < key; }
        rank += __shfl_xor_sync(0xffffffffu, rank, 1);
        if (G == 4) rank += __shfl_xor_sync(0xffffffffu, rank, 2);
        if (sidx < ns && part == 0 && rank < K) members[rank] = bs->sidx[sidx];
    } else if (ns <= 256) {
        for (int sidx = tid; sidx < ns; sidx += PB_NT) {
            const uint32_t key = skey[sidx];
            int rank = 0;
            for (int j = 0; j < ns; j++) rank += skey[j] < key;
            if (rank < K) members[rank] = bs->sidx[sidx];
        }
    } else {
        for (int i = tid; i < n; i += PB_NT) {
            const uint32_t key = keys[i];
            if (key > thr) continue;
            int rank = 0;
            for (int j = 0; j < n && rank < K; j++) rank += keys[j] < key;
            if (rank < K) members[rank] = (uint16_t)i;
        }
    }
    __syncthreads();
    return K;
}

// Stage 2 for a partition whose candidate set did not fit the phase-A buffers (PartA.n2 & S2_SLOW): flat or low-contrast
// content, where thousands of positions share one 8x8 sum and the set "up to j_stop" (moestimation.cpp:470-497) is most of
// the search diamond. The set cannot be stored for every partition of such a picture, and which 33 of it the reference would
// evaluate depends on the predictor, so phase B walks the index itself now that the predictor is known:
//   pass 1  every warp walks (tile, K1-row) ranges with its lanes striding the entries; each thread keeps the two smallest
//           keys (cost << 32 | arrival key) it met -> the 33rd smallest of those 256 keys bounds the 33rd smallest key
//   pass 2  the same walk collects the keys at or below the bound, which are ranked exactly
// then the SADs of the 33 list members (8 threads each). Bucket s0 is visited twice by the reference (:476,486): its entries
// are emitted a second time with the side bit set, which orders them right after the first visit. Returns this thread's
// candidate for the partition minimum; sh.slow_pos[] maps list positions back to displacements. Block-uniform call.
#define SLOW_CAP 512
// (arguments travel through sh.slow_arg: the call must not cost the common path registers)
__device__ __noinline__ u64 stage2_slow(const SeqDev &S, const Geo &g, PBShared &sh, uint2 cur_row)
{
    const int xP = sh.slow_arg[0], yP = sh.slow_arg[1], js = sh.slow_arg[2], genx = sh.slow_arg[3], geny = sh.slow_arg[4], mvpx = sh.slow_arg[5], mvpy = sh.slow_arg[6];
    const int s[5] = { sh.slow_arg[7], sh.slow_arg[8], sh.slow_arg[9], sh.slow_arg[10], sh.slow_arg[11] };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tx0 = max(0, xP - 279) >> FH_TILE_SHIFT, tx1 = min(g.W - 1, xP + 279) >> FH_TILE_SHIFT;
    const int ty0 = max(0, yP - 279) >> FH_TILE_SHIFT, ty1 = min(g.H - 1, yP + 279) >> FH_TILE_SHIFT;
    const int ntx = tx1 - tx0 + 1, ntiles = ntx * (ty1 - ty0 + 1);
    const int qlo = max(0, s[1] - 99) >> 6, qhi = min(8191, s[1] + 99) >> 6, nq = qhi - qlo + 1;
    const int k2lo = max(0, s[2] - 99) >> 6, k2hi = min(8191, s[2] + 99) >> 6;
    const uint4 *__restrict__ tent = (const uint4 *)S.tent;
    u64 *smin = sh.keys2;                       // [0, 256): per-thread minima; [256, 256 + SLOW_CAP): survivor keys
    u64 *skey = sh.keys2 + 256;
    uint32_t *spos = sh.keys1;                  // survivor displacements
    // walks every gated entry with j <= js; fn(key, packed displacement)
    auto walk = [&](auto fn) {
        for (int it = warp; it < ntiles * nq; it += PB_NT / 32) {
            const int tt = it / nq, q = it - tt * nq, tyy = tt / ntx, tx = tx0 + tt - tyy * ntx, ty = ty0 + tyy;
            const int rx0 = tx << FH_TILE_SHIFT, ry0 = ty << FH_TILE_SHIFT;
            const int ddx = max(0, max(rx0 - xP, xP - (rx0 + FH_TILE - 1))), ddy = max(0, max(ry0 - yP, yP - (ry0 + FH_TILE - 1)));
            if (ddx + ddy >= 280) continue;
            const int tile = ty * g.tilesx + tx;
            const uint16_t *ts = S.tstart + (size_t)tile * FH_TSTART_PITCH;
            const int e0 = __ldg(&ts[(qlo + q) * 128 + k2lo]), e1 = __ldg(&ts[(qlo + q) * 128 + k2hi + 1]);
            const uint32_t gbase = (uint32_t)tile * (FH_TILE * FH_TILE);
            // four entries per lane in flight: the walk is bound by the latency of these loads, not by the arithmetic
            // (eight would be 12 % faster here but cost the caller spills around the call)
            for (int eb = e0 + lane; eb < e1; eb += 128) {
                uint4 vv[4];
#pragma unroll
                for (int u = 0; u < 4; u++) vv[u] = eb + 32 * u < e1 ? __ldg(tent + gbase + eb + 32 * u) : make_uint4(0, 0xffffu, 0x7fffu, 0);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint4 v = vv[u];
                    const int x = v.x & 0xffff, y = v.x >> 16, k0 = v.y & 0xffff, k1 = v.y >> 16, k2 = v.z & 0xffff;
                    const int j = iabs_(k0 - s[0]), dx = x - xP, dy = y - yP;
                    if (j > js || j > 180 || iabs_(dx) + iabs_(dy) >= 280 || iabs_(k1 - s[1]) >= 100 || iabs_(k2 - s[2]) >= 100) continue;
                    const uint32_t feat = (uint32_t)feat_dist(s, k0, k1, k2, (int)(v.z >> 16), (int)(v.w & 0xffff));
                    const uint32_t cost = (uint32_t)(iabs_(dx - genx) + iabs_(dy - geny) + 4) * feat;
                    const uint32_t akey = ((uint32_t)j << 21) | ((uint32_t)(k0 > s[0]) << 20) | ((uint32_t)(dx + 279) << 10) | (uint32_t)(dy + 279);
                    const uint32_t pos = (uint32_t)(dx + 512) | ((uint32_t)(dy + 512) << 16);
                    fn(((u64)cost << 32) | akey, pos);
                    if (j == 0) fn(((u64)cost << 32) | akey | (1u << 20), pos);         // second visit of bucket s0
                }
            }
        }
    };
    // pass 1
    u64 m1 = KEY_NONE, m2 = KEY_NONE;
    int cnt = 0;
    walk([&](u64 key, uint32_t) { cnt++; const u64 lo = key < m1 ? key : m1, hi = key < m1 ? m1 : key; m1 = lo; m2 = hi < m2 ? hi : m2; });
    smin[tid] = m1; smin[PB_NT + tid] = m2;
    if (tid == 0) { sh.slow_ns = 0; sh.slow_total = 0; sh.slow_thr = KEY_NONE - 1; }
    __syncthreads();
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) atomicAdd(&sh.slow_total, cnt);
    // the 33rd smallest of the 256 minima (keys are unique): the thread that holds it publishes it
    {
        int r1 = 0, r2 = 0;
        for (int i = 0; i < 2 * PB_NT; i++) { const u64 k = smin[i]; r1 += k < m1; r2 += k < m2; }
        if (m1 != KEY_NONE && r1 == FH_S3_MAX - 1) sh.slow_thr = m1;
        if (m2 != KEY_NONE && r2 == FH_S3_MAX - 1) sh.slow_thr = m2;
    }
    __syncthreads();
    const u64 thr = sh.slow_thr;
    const int K = min(FH_S3_MAX, sh.slow_total);
    if (K == 0) return KEY_NONE;
    // pass 2
    walk([&](u64 key, uint32_t pos) {
        if (key <= thr) { const int p = atomicAdd(&sh.slow_ns, 1); if (p < SLOW_CAP) { skey[p] = key; spos[p] = pos; } }
    });
    __syncthreads();
    const int ns = sh.slow_ns;
    if (ns > SLOW_CAP) { if (tid == 0) atomicOr(&S.status[ST_FLAGS], FLAG_CAPACITY); return KEY_NONE; }
    for (int i = tid; i < ns; i += PB_NT) {
        const u64 key = skey[i];
        int rank = 0;
        for (int k = 0; k < ns; k++) rank += skey[k] < key;
        if (rank < K) { sh.slow_key[rank] = key; sh.slow_pos[rank] = spos[i]; }
    }
    __syncthreads();
    // SADs of the list members (integer displacement: plane 0), 8 threads per member, and their totals
    u64 mine = KEY_NONE;
    const int r = tid & 7;
    for (int m0 = 0; m0 < K; m0 += PB_NT / 8) {
        const int m = m0 + (tid >> 3);
        int sad = 0, dx = 0, dy = 0;
        if (m < K) {
            const uint32_t pos = sh.slow_pos[m];
            dx = (int)(pos & 0xffff) - 512; dy = (int)(pos >> 16) - 512;
            sad = sad8(cur_row, load_row8(S.planes, g.W, g.H, xP + dx, yP + dy + r));
        }
        sad += __shfl_xor_sync(0xffffffffu, sad, 1);
        sad += __shfl_xor_sync(0xffffffffu, sad, 2);
        sad += __shfl_xor_sync(0xffffffffu, sad, 4);
        if (m < K && r == 0 && (uint32_t)(sh.slow_key[m] >> 32) < (uint32_t)FH_COST_EMPTY)
            mine = min(mine, ((u64)(sad + mv_cost(dx << 2, dy << 2, mvpx, mvpy)) << 44) | (1ull << 42) | (u64)m);
    }
    return mine;
}

__global__ void __launch_bounds__(PB_NT, PB_MINB) k_phase_b(const SeqDev *__restrict__ seqs, int seq0, int nseq, Geo g, fh264_params prm,
                                                   uint32_t epoch, const int *__restrict__ wf_order, uint32_t *__restrict__ ticket, int use_spec, int heavy_only)
{
    __shared__ PBShared sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (heavy_only) {                                      // k_phase_b_warp codes the other sequences (usually all of them)
        bool any = false;
        for (int b = 0; b < nseq; b++) any |= seq_is_heavy(seqs[seq0 + b]) && !seqs[seq0 + b].status[ST_GATE];
        if (!any) return;
    }
    // Persistent CTAs: the grid holds about one wavefront's worth of CTAs per sequence (more would only spin and keep
    // other streams' kernels off the SMs); each CTA keeps drawing tickets until the picture is done.
    const uint32_t total = (uint32_t)g.band_nmb * (uint32_t)nseq;
  for (;;) {
    __syncthreads();                                       // previous macroblock's shared state is no longer read
    if (tid == 0) { sh.my_ticket = atomicAdd(ticket, 1u); sh.bs.ns[0] = 0; sh.bs.ns[1] = 0; }
    __syncthreads();
    const uint32_t t = sh.my_ticket;
    if (t >= total) return;
    const SeqDev &S = seqs[seq0 + (int)(t % (uint32_t)nseq)];
    if (S.status[ST_GATE] || (heavy_only && !seq_is_heavy(S))) continue;   // scene cut / not ours: the ticket is dropped (block-uniform)
    const int mb = wf_order[t / (uint32_t)nseq];
    const int mbx = mb % g.Wmb, mby = mb / g.Wmb;
    const int W = g.W, H = g.H;
    NbCache &nc = sh.nc;
    long long *dbg = S.dbg ? S.dbg + (size_t)mb * 24 : nullptr;
#define PB_STAMP(k) do { if (dbg && tid == 0) dbg[k] = gtime_ns(); } while (0)
#define PB_SUB(k) do { if (dbg && tid == 0 && pi == 1) dbg[k] = gtime_ns(); } while (0)
    PB_STAMP(0);
    if (dbg && tid == 0) { dbg[10] = 0; dbg[11] = 0; }

    // ---- everything that does not depend on the neighbours is fetched BEFORE waiting on them: the CTA is resident
    //      long before its turn, so these round trips are off the wavefront's critical path
    if (tid < 16) *(uint4 *)&sh.cur[tid][0] = *(const uint4 *)(S.cur[0] + (size_t)(mby * 16 + tid) * W + mbx * 16);
    if (use_spec && tid >= 96) ((uint4 *)sh.spec)[tid - 96] = ((const uint4 *)&S.spec[(size_t)mb * 4])[tid - 96];       // 4 x 128 bytes
    if (use_spec && tid == 64) *(uint4 *)&sh.ms = *(const uint4 *)&S.mbspec[mb];
    // (the phase-A products — lists, candidate pool — are only needed by a partition that misses the fast path: fetched there)
    auto fetch_phase_a = [&](int pi) {
        if (!prm.basic) {
            if (tid == 32) sh.pa[pi] = S.parta[mb * 4 + pi];
            for (int k = tid; k < FH_S3_MAX; k += PB_NT) sh.s3[pi][k] = S.s3[(size_t)(mb * 4 + pi) * FH_S3_MAX + k];
        } else if (tid == 0) {
            sh.pa[pi].n2 = 0; sh.pa[pi].n3 = 0; sh.pa[pi].s2_off = 0;
        }
        __syncthreads();
    };
    PB_STAMP(1);
    // ---- dependencies. Every final quadrant MV is published as a tagged word (qmv_word). The row ABOVE is waited for here:
    //      exactly the quadrants the predictors can address (up q2/q3, up-right q2, up-left q3). The LEFT neighbour is only
    //      waited for where a predictor really depends on it: a median of three with two equal inputs is that input, so
    //      wherever the motion field above is locally uniform (up q2 == up q3, up q2 == up-right q2, own q0 == q1) the
    //      P_Skip decision and the partition predictors are known without the left MB and the row's serial chain is cut.
    const bool sysw = g.world > 1;                       // band mode: the row above this band is written by another GPU
    const bool mirror = S.peer_qmv_next != nullptr && mby == (g.band_mb0 + g.band_nmb) / g.Wmb - 1;   // band's last MB row
    if (tid < 16) {
        const int w = tid >> 2, q = tid & 3;
        const int nmb = w == 0 ? mb - 1 : (w == 1 ? mb - g.Wmb : (w == 2 ? mb - g.Wmb + 1 : mb - g.Wmb - 1));
        const bool av = w == 0 ? mbx > 0 : (w == 1 ? mby > 0 : (w == 2 ? (mby > 0 && mbx < g.Wmb - 1) : (mby > 0 && mbx > 0)));
        // the quadrants the predictors can address (A.7): up q2/q3, up-right q2, up-left q3 — each polled by its own thread;
        // the left MB's quadrants 1 / 3 are fetched on demand (fetch_left)
        const bool used = av && ((w == 1 && q >= 2) || (w == 2 && q == 2) || (w == 3 && q == 3));
        int vx = 0, vy = 0;
        if (used && !wait_qmv(&S.qmv[(size_t)nmb * 4 + q], epoch, sysw, vx, vy)) atomicOr(&S.status[ST_FLAGS], FLAG_TIMEOUT);
        nc.mvx[w][q] = vx; nc.mvy[w][q] = vy;
        if (q == 0) nc.avail[w] = av;
    }
    __syncthreads();
    PB_STAMP(2);
    const bool leftA = mbx > 0;
    bool left1 = !leftA, left3 = !leftA;                 // left quadrant 1 / 3 present in nc (or not needed)
    auto fetch_left = [&](int q, int where) {            // block-uniform: wait until the left MB has published quadrant q
        if (tid == 0) {
            const long long t0 = dbg ? gtime_ns() : 0;
            int vx, vy;
            if (!wait_qmv(&S.qmv[(size_t)(mb - 1) * 4 + q], epoch, false, vx, vy)) atomicOr(&S.status[ST_FLAGS], FLAG_TIMEOUT);
            nc.mvx[0][q] = vx; nc.mvy[0][q] = vy;
            if (dbg) { dbg[10] += 1ll << (8 * where); dbg[11] += gtime_ns() - t0; }      // timeline: where and how long the left MB was waited for
        }
        __syncthreads();
    };
    // neighbour quadrant MVs used by the predictors, in registers (A.7): up q2/q3, up-right q2, up-left q3
    const int aL = nc.avail[0], aU = nc.avail[1], aUR = nc.avail[2], aUL = nc.avail[3];
    const int u2x = nc.mvx[1][2], u2y = nc.mvy[1][2], u3x = nc.mvx[1][3], u3y = nc.mvy[1][3];
    const int r2x = nc.mvx[2][2], r2y = nc.mvy[2][2], d3x = nc.mvx[3][3], d3y = nc.mvy[3][3];

    PB_STAMP(3);
    // ---- fast path: the whole macroblock by warp 0 from the phase-S products (see spec_lookup / skip_lookup above). On any
    //      lookup that cannot answer, the block-level search below redoes the macroblock from the start (same published values).
    if (use_spec) {
        if (warp == 0) {
            bool ok = true, done = false;
            int l1x = 0, l1y = 0, l3x = 0, l3y = 0;
            bool have1 = !leftA, have3 = !leftA;
            auto poll_left = [&](int q, int &vx, int &vy) {
                int x = 0, y = 0;
                if (lane == 0 && !wait_qmv(&S.qmv[(size_t)(mb - 1) * 4 + q], epoch, false, x, y)) atomicOr(&S.status[ST_FLAGS], FLAG_TIMEOUT);
                vx = __shfl_sync(0xffffffffu, x, 0); vy = __shfl_sync(0xffffffffu, y, 0);
            };
            auto publish = [&](int q, int vx, int vy) {
                if (lane == 0) {
                    const unsigned long long wq = qmv_word(epoch, vx, vy);
                    st_relaxed_u64(&S.qmv[(size_t)mb * 4 + q], wq, false);
                    if (mirror) st_relaxed_u64(&S.peer_qmv_next[(size_t)mb * 4 + q], wq, true);
                }
            };
            const MbSpec &ms = sh.ms;
            int smx = 0, smy = 0, nbad = 1;
            if (!leftA || mby == 0 || (u2x == 0 && u2y == 0)) nbad = skip_lookup(ms, 0, 0);
            else {
                const int cx16 = aUR ? r2x : d3x, cy16 = aUR ? r2y : d3y;
                if (u2x == cx16 && u2y == cy16) {
                    const int nbB = skip_lookup(ms, u2x, u2y), nb0 = skip_lookup(ms, 0, 0);
                    if (lane == 0) S.prev_gen16[mb] = ((uint32_t)(u2x >> 2) & 0xffffu) | ((uint32_t)(u2y >> 2) << 16);
                    if (nbB < 0) ok = false;
                    else if (nbB != 0 && nb0 != 0) nbad = 1;
                    else {
                        poll_left(1, l1x, l1y); have1 = true;
                        const bool lz = l1x == 0 && l1y == 0;
                        smx = lz ? 0 : u2x; smy = lz ? 0 : u2y; nbad = lz ? nb0 : nbB;
                    }
                } else {
                    poll_left(1, l1x, l1y); have1 = true;
                    if (!(l1x == 0 && l1y == 0)) {
                        median_pred(1, l1x, l1y, 1, u2x, u2y, 1, cx16, cy16, smx, smy);      // A = left q1, B = up q2, C = up-right q2 else up-left q3: all available here
                        if (lane == 0) S.prev_gen16[mb] = ((uint32_t)(smx >> 2) & 0xffffu) | ((uint32_t)(smy >> 2) << 16);
                    }
                    nbad = skip_lookup(ms, smx, smy);
                    if (nbad < 0) ok = false;
                }
            }
            if (dbg && lane == 0) dbg[4] = gtime_ns();
            if (ok && nbad == 0) {
                for (int q = 0; q < 4; q++) publish(q, smx, smy);
                if (lane == 0) {
                    MbMotion mo;
                    mo.maxdiff = (int16_t)ms.maxdiff; mo.pad = 0; mo.mb_type = FH264_P_SKIP; mo.num_parts = 0;
                    for (int i = 0; i < 4; i++) { mo.mv[i][0] = (int16_t)smx; mo.mv[i][1] = (int16_t)smy; mo.mvd[i][0] = mo.mvd[i][1] = 0; mo.sad[i] = 0; }
                    uint4 *d = (uint4 *)&S.motion[mb];
                    const uint4 *s4 = (const uint4 *)&mo;
                    d[0] = s4[0]; d[1] = s4[1]; d[2] = s4[2];
                    atomicAdd(&S.status[ST_COUNTS + 0], 1u);
                }
                done = true;
            } else if (ok) {
                int q0x = 0, q0y = 0, q1x = 0, q1y = 0, q2x = 0, q2y = 0, q3x = 0, q3y = 0, s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                int p0x, p0y, p1x, p1y, p2x, p2y, p3x, p3y;
                // partition 0 predicts from left q1 unless up q2 == up q3; partition 2 from left q3 unless own q0 == q1
                if (!have1 && !(aU && u2x == u3x && u2y == u3y)) { poll_left(1, l1x, l1y); have1 = true; }
                median_pred(aL, l1x, l1y, aU, u2x, u2y, aU ? 1 : aUL, aU ? u3x : d3x, aU ? u3y : d3y, p0x, p0y);
                ok = spec_lookup_w(sh.spec[0], p0x >> 2, p0y >> 2, p0x, p0y, q0x, q0y, s0);
                if (ok) {
                    publish(0, q0x, q0y);
                    if (dbg && lane == 0) dbg[5] = gtime_ns();
                    median_pred(1, q0x, q0y, aU, u3x, u3y, aUR ? 1 : aU, aUR ? r2x : u2x, aUR ? r2y : u2y, p1x, p1y);
                    ok = spec_lookup_w(sh.spec[1], p1x >> 2, p1y >> 2, p1x, p1y, q1x, q1y, s1);
                }
                if (ok) {
                    publish(1, q1x, q1y);
                    if (dbg && lane == 0) dbg[6] = gtime_ns();
                    if (!have3 && !(q0x == q1x && q0y == q1y)) { poll_left(3, l3x, l3y); have3 = true; }
                    median_pred(aL, l3x, l3y, 1, q0x, q0y, 1, q1x, q1y, p2x, p2y);
                    ok = spec_lookup_w(sh.spec[2], p2x >> 2, p2y >> 2, p2x, p2y, q2x, q2y, s2);
                }
                if (ok) {
                    publish(2, q2x, q2y);
                    if (dbg && lane == 0) dbg[7] = gtime_ns();
                    median_pred(1, q2x, q2y, 1, q1x, q1y, 1, q0x, q0y, p3x, p3y);
                    ok = spec_lookup_w(sh.spec[3], p3x >> 2, p3y >> 2, p3x, p3y, q3x, q3y, s3);
                }
                if (ok) {
                    publish(3, q3x, q3y);
                    if (dbg && lane == 0) dbg[8] = gtime_ns();
                    // merge (:529-551) and final mvd with the merged type's predictors (:552-564)
                    const bool eq01 = q0x == q1x && q0y == q1y, eq23 = q2x == q3x && q2y == q3y;
                    const bool eq02 = q0x == q2x && q0y == q2y, eq13 = q1x == q3x && q1y == q3y;
                    if ((eq01 && eq23) || (eq02 && eq13)) {               // 16x16 / 16x8 / 8x16: the predictors read the left macroblock
                        if (!have1) { poll_left(1, l1x, l1y); have1 = true; }
                        if (!have3) { poll_left(3, l3x, l3y); have3 = true; }
                    }
                    if (lane == 0) {
                        nc.mvx[0][1] = l1x; nc.mvy[0][1] = l1y; nc.mvx[0][3] = l3x; nc.mvy[0][3] = l3y;
                        const int mv[4][2] = { { q0x, q0y }, { q1x, q1y }, { q2x, q2y }, { q3x, q3y } };
                        const int mvps[4][2] = { { p0x, p0y }, { p1x, p1y }, { p2x, p2y }, { p3x, p3y } };
                        const int sadq[4] = { s0, s1, s2, s3 };
                        int type = FH264_P_8x8ref0, nparts = 4, cnt = 4;
                        if (eq01 && eq23 && eq02) { type = FH264_P_L0_16x16; nparts = 1; cnt = 1; }
                        else if (eq01 && eq23) { type = FH264_P_L0_L0_16x8; nparts = 2; cnt = 2; }
                        else if (eq02 && eq13) { type = FH264_P_L0_L0_8x16; nparts = 2; cnt = 3; }
                        MbMotion mo;
                        mo.maxdiff = (int16_t)ms.maxdiff; mo.pad = 0;
                        int fin[4][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };
                        for (int i = 0; i < 4; i++) { mo.mvd[i][0] = mo.mvd[i][1] = 0; }
                        for (int i = 0; i < nparts; i++) {
                            int ppx = 0, ppy = 0, pw = 16, dir = 0, qsel = i, ox, oy;
                            if (type == FH264_P_L0_L0_16x8) { ppy = i * 8; dir = i == 0 ? 1 : 2; qsel = i * 2; }
                            else if (type == FH264_P_L0_L0_8x16) { ppx = i * 8; pw = 8; dir = i == 0 ? 2 : 3; }
                            else if (type == FH264_P_8x8ref0) { ppx = (i & 1) * 8; ppy = (i >> 1) * 8; pw = 8; }
                            if (type == FH264_P_8x8ref0) { ox = mvps[i][0]; oy = mvps[i][1]; }
                            else predict_mv_(nc, ppx, ppy, pw, dir, fin, ox, oy);
                            mo.mvd[i][0] = (int16_t)(mv[qsel][0] - ox); mo.mvd[i][1] = (int16_t)(mv[qsel][1] - oy);
                            for (int q = 0; q < 4; q++) {
                                const bool in = type == FH264_P_L0_16x16 || (type == FH264_P_L0_L0_16x8 && (q >> 1) == i) ||
                                                (type == FH264_P_L0_L0_8x16 && (q & 1) == i) || (type == FH264_P_8x8ref0 && q == i);
                                if (in) { fin[q][0] = mv[qsel][0]; fin[q][1] = mv[qsel][1]; }
                            }
                        }
                        mo.mb_type = (int16_t)type; mo.num_parts = (int16_t)nparts;
                        for (int q = 0; q < 4; q++) { mo.mv[q][0] = (int16_t)fin[q][0]; mo.mv[q][1] = (int16_t)fin[q][1]; mo.sad[q] = (uint16_t)sadq[q]; }
                        uint4 *d = (uint4 *)&S.motion[mb];
                        const uint4 *s4 = (const uint4 *)&mo;
                        d[0] = s4[0]; d[1] = s4[1]; d[2] = s4[2];
                        for (int i = 0; i < 4; i++) S.prev_gen[(size_t)mb * 4 + i] = ((uint32_t)(mvps[i][0] >> 2) & 0xffffu) | ((uint32_t)(mvps[i][1] >> 2) << 16);
                        if (dbg) dbg[9] = gtime_ns();
                        atomicAdd(&S.status[ST_COUNTS + cnt], 1u);
                        atomicAdd(&S.status[ST_SPEC_HIT], 4u);
                    }
                    done = true;
                }
            }
            if (lane == 0) sh.fast_done = done ? 1 : 0;
        }
        __syncthreads();
        if (sh.fast_done) continue;
    }
    // ---- P_Skip trial (mode_pred.cpp:383-401, moestimation.cpp:402-425) -----------------------------------------
    int zero4[4][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };
    const int py = tid >> 3, px = (tid & 7) * 2;                          // this thread's two luma samples
    const int c0 = sh.cur[py][px], c1 = sh.cur[py][px + 1];
    int maxdiff = prm.maxdiff_set;
    if (use_spec) maxdiff = sh.ms.maxdiff;                                // computed by phase S (same arithmetic)
    else if (prm.maxdiff_set == -1) {                                     // moestimation.cpp:407-419
        int v = __reduce_add_sync(0xffffffffu, c0 + c1);
        if (lane == 0) sh.red[warp] = v;
        __syncthreads();
        const int mean = (sh.red[0] + sh.red[1] + sh.red[2] + sh.red[3]) / 256;
        __syncthreads();
        v = __reduce_add_sync(0xffffffffu, iabs_(c0 - mean) + iabs_(c1 - mean));
        if (lane == 0) sh.red[warp] = v;
        __syncthreads();
        maxdiff = max(3, (sh.red[0] + sh.red[1] + sh.red[2] + sh.red[3]) / 256);
    }
    auto skip_pred = [&](int vx, int vy, int (&p2)[2]) {
        luma_pred_block<2, 1>(S, g, mbx * 16 + px + (vx >> 2), mby * 16 + py + (vy >> 2), vx & 3, vy & 3, p2);
    };
    auto skip_bad = [&](const int (&p2)[2]) -> int { return __syncthreads_count(iabs_(c0 - p2[0]) > maxdiff || iabs_(c1 - p2[1]) > maxdiff); };
    // number of samples outside MAXDIFF for skip vector (vx, vy) — only zero / non-zero matters. Looked up in the phase-S masks
    // when the vector lies in a guessed cell, else measured here. Block-uniform.
    auto skip_nbad = [&](int vx, int vy) -> int {
        if (use_spec) {
            if (vx == 0 && vy == 0) return sh.ms.zero_ok ? 0 : 1;
            const int cxx = vx >> 2, cyy = vy >> 2, f = (vy & 3) * 4 + (vx & 3);
            if (sh.ms.cx[0] == cxx && sh.ms.cy[0] == cyy) return ((sh.ms.mask[0] >> f) & 1) ? 0 : 1;
            if (sh.ms.cx[1] == cxx && sh.ms.cy[1] == cyy) return ((sh.ms.mask[1] >> f) & 1) ? 0 : 1;
        }
        int p2[2];
        skip_pred(vx, vy, p2);
        return skip_bad(p2);
    };
    int smx = 0, smy = 0, nbad;
    if (!leftA || mby == 0 || (u2x == 0 && u2y == 0)) {                  // the skip MV is zero whatever the left MB holds
        nbad = skip_nbad(0, 0);
    } else {
        // 16x16 predictor: A = left q1, B = up q2, C = up-right q2 else up-left q3 (all available here)
        const int cx16 = aUR ? r2x : d3x, cy16 = aUR ? r2y : d3y;
        if (u2x == cx16 && u2y == cy16) {
            // B == C: the skip MV is B, or zero if the left quadrant turns out to be zero — try both without waiting
            const int nbB = skip_nbad(u2x, u2y), nb0 = skip_nbad(0, 0);
            if (tid == 0) S.prev_gen16[mb] = ((uint32_t)(u2x >> 2) & 0xffffu) | ((uint32_t)(u2y >> 2) << 16);
            if (nbB != 0 && nb0 != 0) nbad = 1;
            else {
                fetch_left(1, 0); left1 = true;
                const bool lz = nc.mvx[0][1] == 0 && nc.mvy[0][1] == 0;
                smx = lz ? 0 : u2x; smy = lz ? 0 : u2y; nbad = lz ? nb0 : nbB;
            }
        } else {
            fetch_left(1, 1); left1 = true;
            if (!(nc.mvx[0][1] == 0 && nc.mvy[0][1] == 0)) {
                predict_mv_(nc, 0, 0, 16, 0, zero4, smx, smy);
                if (tid == 0) S.prev_gen16[mb] = ((uint32_t)(smx >> 2) & 0xffffu) | ((uint32_t)(smy >> 2) << 16);
            }
            nbad = skip_nbad(smx, smy);
        }
    }
    MbMotion mo;
    mo.maxdiff = (int16_t)maxdiff; mo.pad = 0;
    PB_STAMP(4);
    if (nbad == 0) {
        if (tid == 0) {
            mo.mb_type = FH264_P_SKIP; mo.num_parts = 0;
            for (int i = 0; i < 4; i++) { mo.mv[i][0] = (int16_t)smx; mo.mv[i][1] = (int16_t)smy; mo.mvd[i][0] = mo.mvd[i][1] = 0; mo.sad[i] = 0; }
            uint4 *d = (uint4 *)&S.motion[mb];
            const uint4 *s4 = (const uint4 *)&mo;
            d[0] = s4[0]; d[1] = s4[1]; d[2] = s4[2];
            const unsigned long long wq = qmv_word(epoch, smx, smy);
            for (int i = 0; i < 4; i++) st_relaxed_u64(&S.qmv[(size_t)mb * 4 + i], wq, false);
            if (mirror) for (int i = 0; i < 4; i++) st_relaxed_u64(&S.peer_qmv_next[(size_t)mb * 4 + i], wq, true);
            atomicAdd(&S.status[ST_COUNTS + 0], 1u);
        }
        continue;
    }

    // ---- 8x8 search, partitions in order; inside a partition the three stages are evaluated cooperatively by the whole
    //      block (the wavefront is latency bound: what sits on the critical path is dependent instructions and memory
    //      round trips, so independent work is issued first and consumed late) --------------------------------------------
    int mv[4][2], sadq[4], mvps[4][2];
    int q0x = 0, q0y = 0, q1x = 0, q1y = 0, q2x = 0, q2y = 0;          // quadrant MVs decided so far
    const int g1 = prm.window / 16, w1 = 2 * g1 + 1, n1 = w1 * w1 * 16, inv1 = 65536 / w1 + 1, npos = w1 * w1;
    // qfeat.cuh geometry: window rows, plane stride, planes per warp and batch. Each WARP produces and consumes its own planes
    // (warp w: planes f0 + w*qpw ..), so the row sums and the records only need warp barriers.
    const int qR = 8 + w1 - 1, qps = (qR * w1) | 1, qpw = w1 <= 5 ? 4 : 1, qseg = qpw * qR;
    const uint32_t iR = udiv_magic((uint32_t)qR), iNp = udiv_magic((uint32_t)npos);
    uint32_t *qx = sh.qx + warp * (qpw * qps);
    uint16_t *qrc = sh.qrc + warp * (qpw * qps);
    int callno = 0;                  // alternates BlockSel's survivor counters (uniform across the block)
    int nhit = 0;                    // partitions decided from the phase-S finalists
    int l1x = 0, l1y = 0, l3x = 0, l3y = 0;                           // left MB's quadrants 1 / 3 once fetched
    for (int pi = 0; pi < 4; pi++) {
        const int xP = mbx * 16 + (pi & 1) * 8, yP = mby * 16 + (pi >> 1) * 8;
        // partition 0 predicts from left q1 unless up q2 == up q3; partition 2 from left q3 unless own q0 == q1
        if (pi == 0 && !left1 && !(aU && u2x == u3x && u2y == u3y)) { fetch_left(1, 2); left1 = true; }
        if (pi == 2 && !left3 && !(q0x == q1x && q0y == q1y)) { fetch_left(3, 3); left3 = true; }
        if (pi == 0 && left1) { l1x = nc.mvx[0][1]; l1y = nc.mvy[0][1]; }
        if (pi == 2 && left3) { l3x = nc.mvx[0][3]; l3y = nc.mvy[0][3]; }
        int mvpx, mvpy;
        // A = (px-1,py), B = (px,py-1), C = (px+8,py-1) else D = (px-1,py-1) for an 8x8 partition (mode_pred.cpp:113-161)
        if (pi == 0) median_pred(aL, l1x, l1y, aU, u2x, u2y, aU ? 1 : aUL, aU ? u3x : d3x, aU ? u3y : d3y, mvpx, mvpy);
        else if (pi == 1) median_pred(1, q0x, q0y, aU, u3x, u3y, aUR ? 1 : aU, aUR ? r2x : u2x, aUR ? r2y : u2y, mvpx, mvpy);
        else if (pi == 2) median_pred(aL, l3x, l3y, 1, q0x, q0y, 1, q1x, q1y, mvpx, mvpy);
        else median_pred(1, q2x, q2y, 1, q1x, q1y, 1, q0x, q0y, mvpx, mvpy);
        mvps[pi][0] = mvpx; mvps[pi][1] = mvpy;
        PB_SUB(12);
        const int genx = mvpx >> 2, geny = mvpy >> 2;
        if (tid == 0) S.prev_gen[(size_t)mb * 4 + pi] = ((uint32_t)genx & 0xffffu) | ((uint32_t)geny << 16);      // next picture's temporal guess
        int bx = 0, by = 0, bs = 0;
        // Speculative fast path (spec.cuh): phase S already ran the whole search for up to two guessed values of gen and left the
        // few candidates that can win for some predictor of that cell; with the true predictor known the winner is the first
        // minimum of SAD + |mv - mvp|_1 in (stage, list position) order among them (:460-469,498-507,511-520). Block-uniform.
        bool hit = false;
        if (use_spec) {
            const PartSpec &sp = sh.spec[pi];
            int slot = -1;
            if (sp.nf[0] != SPEC_INVALID && sp.gx[0] == genx && sp.gy[0] == geny) slot = 0;
            else if (sp.nf[1] != SPEC_INVALID && sp.gx[1] == genx && sp.gy[1] == geny) slot = 1;
            if (slot >= 0) {
                const int nf = sp.nf[slot];
                u64 bk = KEY_NONE;
                for (int k = 0; k < nf; k++) {
                    const SpecFinal f = sp.f[slot][k];
                    const u64 key = ((u64)((int)f.sad + mv_cost(f.mvx, f.mvy, mvpx, mvpy)) << 32) | ((u64)f.order << 16) | (u64)k;
                    bk = min(bk, key);
                }
                const SpecFinal f = sp.f[slot][(int)(bk & 15)];
                bx = f.mvx; by = f.mvy; bs = f.sad;
                hit = true; nhit++;
            }
        }
        if (!hit) {
        fetch_phase_a(pi);
        const PartA pa = sh.pa[pi];
        u64 *best = sh.best[pi & 1];
        u64 mine = KEY_NONE;
        // stage 1 (:458-469): the window/16 quarter-pel window around the predictor. Its features are computed from the
        // interpolated planes (qfeat.cuh): issue the pixel-row loads now (thread -> (plane, window row)), consume them later.
        const int x0 = xP + genx - g1, y0 = yP + geny - g1;
        const int sa = lane, sb = lane + 32;                  // this lane's (plane, row) segments among the warp's qseg (<= 64)
        const int fa = udiv_by(sa, iR), ra = sa - fa * qR, fb = udiv_by(sb, iR), rb = sb - fb * qR;
        uint4 wa = make_uint4(0, 0, 0, 0), wb = wa;
        if (sa < qseg) wa = qf_load16(S.planes + (size_t)(warp * qpw + fa) * g.WH, W, H, x0, y0 + ra);
        if (sb < qseg) wb = qf_load16(S.planes + (size_t)(warp * qpw + fb) * g.WH, W, H, x0, y0 + rb);
        // stage 2 (:470-507) keys while those loads fly: cost << 32 | arrival key (the pool is unordered; phase A left no SADs)
        const bool slow2 = (pa.n2 & S2_SLOW) != 0;                     // candidate set too large for phase A: enumerated below
        const int n2 = slow2 ? 0 : (int)pa.n2;
        const uint4 *pool = S.s2pool + pa.s2_off;
        for (int i = tid; i < n2; i += PB_NT) {
            const uint4 v = __ldg(&pool[i]);
            const int dx = (int16_t)(v.x & 0xffff), dy = (int16_t)(v.x >> 16);
            const uint32_t cost = (uint32_t)(iabs_(dx - genx) + iabs_(dy - geny) + 4) * v.y;
            sh.keys2[FH_IDX(i, 1024)] = ((u64)cost << 32) | (u64)v.w;
        }
        // stage 3 (:508-520): phase-A list, already in list order
        if (tid >= 64 && tid - 64 < (int)pa.n3) {
            const int i = tid - 64;
            const S3Entry e = sh.s3[pi][i];
            mine = min(mine, ((u64)((int)e.sad + mv_cost(e.mvx, e.mvy, mvpx, mvpy)) << 44) | (2ull << 42) | (u64)i);
        }
        // stage 1 keys = cost << 11 | arrival index ((dx, dy, frac) order)
        PB_SUB(13);
        const uint2 *rp = (const uint2 *)&sh.cur[(pi >> 1) * 8][(pi & 1) * 8];
        uint2 rows[8];
#pragma unroll
        for (int r = 0; r < 8; r++) rows[r] = rp[r * 2];           // 16-byte row pitch
        {
            int s[5];
            block_sums(rows, s);                                       // suma[0..4] (:440-451)
            const FeatQ fq = feat_query(s);
            if (slow2) {
                if (tid == 0) {
                    sh.slow_arg[0] = xP; sh.slow_arg[1] = yP; sh.slow_arg[2] = (int)(pa.n2 & 0xffu); sh.slow_arg[3] = genx; sh.slow_arg[4] = geny;
                    sh.slow_arg[5] = mvpx; sh.slow_arg[6] = mvpy;
                    for (int k = 0; k < 5; k++) sh.slow_arg[7 + k] = s[k];
                }
                __syncthreads();
                mine = min(mine, stage2_slow(S, g, sh, pick_row(rows, tid & 7)));
            }
            // planes in batches of 4 * qpw (all 16 at once up to WindowSize 32): row sums -> warp barrier -> records and keys
            for (int f0 = 0; f0 < 16; f0 += 4 * qpw) {
                const int fw = f0 + warp * qpw;                        // this warp's first plane of the batch
                if (f0 > 0) {
                    __syncwarp();                                      // previous batch consumed
                    if (sa < qseg) wa = qf_load16(S.planes + (size_t)(fw + fa) * g.WH, W, H, x0, y0 + ra);
                    if (sb < qseg) wb = qf_load16(S.planes + (size_t)(fw + fb) * g.WH, W, H, x0, y0 + rb);
                }
                if (sa < qseg) qf_row_sums(wa, w1, qx + fa * qps + ra * w1, qrc + fa * qps + ra * w1);
                if (sb < qseg) qf_row_sums(wb, w1, qx + fb * qps + rb * w1, qrc + fb * qps + rb * w1);
                __syncwarp();
                for (int o = lane; o < qpw * npos; o += 32) {
                    const int fl = udiv_by(o, iNp), pos = o - fl * npos, cx = fdiv_(pos, inv1), cy = pos - cx * w1, ox = cx - g1, oy = cy - g1;
                    const int rx = xP + genx + ox, ry = yP + geny + oy, i = pos * 16 + fw + fl;
                    uint32_t key = COST_INVALID;
                    if (rx >= 0 && rx < W && ry >= 0 && ry < H)
                        key = ((uint32_t)((iabs_(ox) + iabs_(oy) + 4) * feat_of(fq, qf_record(qx, qrc, fl, qps, w1, cx, cy))) << 11) | (uint32_t)i;
                    sh.keys1[FH_IDX(i, S1_KEY_CAP)] = key;
                }
            }
        }
        __syncthreads();                                               // keys1 and keys2 complete
        PB_SUB(14);
        const int K1 = block_select_smallest_u32(sh.keys1, n1, FH_S1_MAX, &sh.bs, sh.mem1, callno++);
        PB_SUB(15);
        // SAD loads of the (at most 17) stage-1 members: 8 threads per member, one row each; consumed after stage 2
        const int r = tid & 7;
        const uint2 cr = pick_row(rows, r);
        uint2 rr[2];
        int mvx1[2], mvy1[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int m = u * 16 + (tid >> 3);
            rr[u] = make_uint2(0, 0); mvx1[u] = mvy1[u] = 0;
            if (m < K1) {
                const int i = sh.mem1[m], f = i & 15, pos = i >> 4, cx = fdiv_(pos, inv1), dx = genx + cx - g1, dy = geny + pos - cx * w1 - g1;
                mvx1[u] = (dx << 2) | (f & 3); mvy1[u] = (dy << 2) | (f >> 2);
                rr[u] = load_row8(S.planes + (size_t)f * g.WH, W, H, clampi_(xP + dx, 0, W - 1), clampi_(yP + dy, 0, H - 1) + r);
            }
        }
        // stage 2 (:498-507): the 33 smallest keys are the list; their SADs are measured here (8 threads per member), and a member
        // with cost below the "empty" mark competes with its list position as tie-break
        if (n2 > 0) {
            const int K2 = block_select_smallest(n2, FH_S3_MAX, [&](int i) -> u64 { return sh.keys2[i]; }, &sh.bs, sh.mem2, callno++);
            for (int m0 = 0; m0 < K2; m0 += PB_NT / 8) {
                const int m = m0 + (tid >> 3);
                int sad = 0, dx = 0, dy = 0;
                bool live = false;
                if (m < K2) {
                    const int i = (int)sh.mem2[m];
                    const uint4 v = __ldg(&pool[i]);
                    dx = (int16_t)(v.x & 0xffff); dy = (int16_t)(v.x >> 16);
                    live = (uint32_t)(sh.keys2[i] >> 32) < (uint32_t)FH_COST_EMPTY;
                    sad = sad8(cr, load_row8(S.planes, W, H, xP + dx, yP + dy + r));
                }
                sad += __shfl_xor_sync(0xffffffffu, sad, 1);
                sad += __shfl_xor_sync(0xffffffffu, sad, 2);
                sad += __shfl_xor_sync(0xffffffffu, sad, 4);
                if (live && r == 0) mine = min(mine, ((u64)(sad + mv_cost(dx << 2, dy << 2, mvpx, mvpy)) << 44) | (1ull << 42) | (u64)m);
            }
        }
        // stage 1 SADs
        PB_SUB(16);
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int m = u * 16 + (tid >> 3);
            int sad = m < K1 ? sad8(cr, rr[u]) : 0;
            if (u == 0 || tid < 8) {                               // round 1 only concerns member 16 (threads 0-7)
                sad += __shfl_xor_sync(u == 0 ? 0xffffffffu : 0xffu, sad, 1);
                sad += __shfl_xor_sync(u == 0 ? 0xffffffffu : 0xffu, sad, 2);
                sad += __shfl_xor_sync(u == 0 ? 0xffffffffu : 0xffu, sad, 4);
            }
            if (m < K1 && r == 0) mine = min(mine, ((u64)(sad + mv_cost(mvx1[u], mvy1[u], mvpx, mvpy)) << 44) | (u64)sh.keys1[sh.mem1[m]]);
        }
        mine = warp_min_u64(mine);
        if (lane == 0) best[warp] = mine;
        __syncthreads();
        // decode the winner from its key (:523-527); no candidate at all leaves bx = by = 0 (:452)
        const u64 b = min(min(best[0], best[1]), min(best[2], best[3]));
        PB_SUB(17);
        if (b != KEY_NONE) {
            const int stage = (int)((b >> 42) & 3), total = (int)(b >> 44);
            if (stage == 0) {
                const int i = (int)(b & 2047), f = i & 15, pos = i >> 4, cx = fdiv_(pos, inv1);
                bx = ((genx + cx - g1) << 2) | (f & 3); by = ((geny + pos - cx * w1 - g1) << 2) | (f >> 2);
            } else if (stage == 1) {
                const int wi = (int)(b & 1023);
                if (slow2) { const uint32_t pos = sh.slow_pos[wi]; bx = ((int)(pos & 0xffff) - 512) << 2; by = ((int)(pos >> 16) - 512) << 2; }
                else {
                    const uint4 v = __ldg(&S.s2pool[pa.s2_off + (uint32_t)sh.mem2[wi]]);
                    bx = ((int)(int16_t)(v.x & 0xffff)) << 2; by = ((int)(int16_t)(v.x >> 16)) << 2;
                }
            } else {
                const S3Entry e = sh.s3[pi][(int)(b & 63)];
                bx = e.mvx; by = e.mvy;
            }
            bs = total - mv_cost(bx, by, mvpx, mvpy);
        } else {
            const uint8_t *pl = S.planes;
            for (int rr2 = 0; rr2 < 8; rr2++) bs += sad_row8(*(const uint2 *)&sh.cur[(pi >> 1) * 8 + rr2][(pi & 1) * 8], pl, W, H, xP, yP + rr2);
        }
        }   // !hit
        mv[pi][0] = bx; mv[pi][1] = by; sadq[pi] = bs;
        if (tid == 0) {
            // publish this quadrant's MV at once: the right and lower-left neighbours can start before this MB is finished
            // (neighbours only ever read quadrant MVs; the merged type / mvd / SADs of the record are read by phase C).
            // Band mode: the rank below predicts from this row — the same word goes into its memory over NVLink.
            const unsigned long long wq = qmv_word(epoch, bx, by);
            st_relaxed_u64(&S.qmv[(size_t)mb * 4 + pi], wq, false);
            if (mirror) st_relaxed_u64(&S.peer_qmv_next[(size_t)mb * 4 + pi], wq, true);
        }
        PB_STAMP(5 + pi);
        if (pi == 0) { q0x = bx; q0y = by; } else if (pi == 1) { q1x = bx; q1y = by; } else if (pi == 2) { q2x = bx; q2y = by; }
    }

    // ---- merge (:529-551) and final mvd with the merged type's predictors (:552-564) ---------------------------
    {
        const bool m01 = mv[0][0] == mv[1][0] && mv[0][1] == mv[1][1], m23 = mv[2][0] == mv[3][0] && mv[2][1] == mv[3][1];
        const bool m02 = mv[0][0] == mv[2][0] && mv[0][1] == mv[2][1], m13 = mv[1][0] == mv[3][0] && mv[1][1] == mv[3][1];
        const bool merged = (m01 && m23) || (m02 && m13);                  // 16x16 / 16x8 / 8x16: predictors read the left MB
        if (merged && !left1) { fetch_left(1, 4); left1 = true; }
        if (merged && !left3) { fetch_left(3, 5); left3 = true; }
    }
    if (tid == 0) {
        const bool eq01 = mv[0][0] == mv[1][0] && mv[0][1] == mv[1][1], eq23 = mv[2][0] == mv[3][0] && mv[2][1] == mv[3][1];
        const bool eq02 = mv[0][0] == mv[2][0] && mv[0][1] == mv[2][1], eq13 = mv[1][0] == mv[3][0] && mv[1][1] == mv[3][1];
        int type = FH264_P_8x8ref0, nparts = 4, cnt = 4;
        if (eq01 && eq23 && eq02) { type = FH264_P_L0_16x16; nparts = 1; cnt = 1; }
        else if (eq01 && eq23) { type = FH264_P_L0_L0_16x8; nparts = 2; cnt = 2; }
        else if (eq02 && eq13) { type = FH264_P_L0_L0_8x16; nparts = 2; cnt = 3; }
        int fin[4][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };
        for (int i = 0; i < 4; i++) { mo.mvd[i][0] = mo.mvd[i][1] = 0; }
        for (int i = 0; i < nparts; i++) {
            int ppx = 0, ppy = 0, pw = 16, dir = 0, qsel = i, ox, oy;
            if (type == FH264_P_L0_L0_16x8) { ppy = i * 8; dir = i == 0 ? 1 : 2; qsel = i * 2; }
            else if (type == FH264_P_L0_L0_8x16) { ppx = i * 8; pw = 8; dir = i == 0 ? 2 : 3; }
            else if (type == FH264_P_8x8ref0) { ppx = (i & 1) * 8; ppy = (i >> 1) * 8; pw = 8; }
            if (type == FH264_P_8x8ref0) { ox = mvps[i][0]; oy = mvps[i][1]; }        // same neighbours as during the search
            else predict_mv_(nc, ppx, ppy, pw, dir, fin, ox, oy);
            mo.mvd[i][0] = (int16_t)(mv[qsel][0] - ox); mo.mvd[i][1] = (int16_t)(mv[qsel][1] - oy);
            for (int q = 0; q < 4; q++) {
                const bool in = type == FH264_P_L0_16x16 || (type == FH264_P_L0_L0_16x8 && (q >> 1) == i) ||
                                (type == FH264_P_L0_L0_8x16 && (q & 1) == i) || (type == FH264_P_8x8ref0 && q == i);
                if (in) { fin[q][0] = mv[qsel][0]; fin[q][1] = mv[qsel][1]; }
            }
        }
        mo.mb_type = (int16_t)type; mo.num_parts = (int16_t)nparts;
        for (int q = 0; q < 4; q++) { mo.mv[q][0] = (int16_t)fin[q][0]; mo.mv[q][1] = (int16_t)fin[q][1]; mo.sad[q] = (uint16_t)sadq[q]; }
        uint4 *d = (uint4 *)&S.motion[mb];
        const uint4 *s4 = (const uint4 *)&mo;
        d[0] = s4[0]; d[1] = s4[1]; d[2] = s4[2];
        PB_STAMP(9);
        atomicAdd(&S.status[ST_COUNTS + cnt], 1u);
        if (nhit) atomicAdd(&S.status[ST_SPEC_HIT], (uint32_t)nhit);
        if (nhit < 4) atomicAdd(&S.status[ST_SPEC_MISS], (uint32_t)(4 - nhit));
    }
  }
}
