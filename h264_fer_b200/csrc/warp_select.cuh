// Warp-synchronous "k smallest, in order" selection. The reference keeps insertion-sorted lists
// (moestimation.cpp:277-291); downstream only list MEMBERSHIP (the k smallest by (cost, arrival)) and the order
// among members matter, so each list is rebuilt by: (1) a cheap upper bound on the k-th key from the per-lane
// smallest keys, (2) compaction of the survivors, (3) exact ranking of the survivors by counting.
// Keys are 64-bit and UNIQUE (cost in the high bits, arrival index in the low bits); ~0 marks "no candidate".
#pragma once
#include "common.cuh"

#define KEY_NONE 0xffffffffffffffffull
typedef unsigned long long u64;

__device__ __forceinline__ u64 warp_min_u64(u64 v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) { u64 o = __shfl_xor_sync(0xffffffffu, v, d); v = o < v ? o : v; }
    return v;
}
__device__ __forceinline__ u64 warp_max_u64(u64 v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) { u64 o = __shfl_xor_sync(0xffffffffu, v, d); v = o > v ? o : v; }
    return v;
}

struct WarpSelScratch {           // per warp
    u64 skey[256];
    uint16_t sidx[256];
};
#define WSEL_CAP 256

// Selects the K = min(k, #valid) smallest keys among keyfn(0..n-1), k <= 64. On return members[r] (r < K) is the
// element index with rank r (ascending key). Returns K. All 32 lanes must call; ends with __syncwarp().
//
// Upper bound on the K-th smallest key: with m1/m2 the smallest / second smallest key of each lane,
//   L1 lanes hold a key        => at least L1 keys are <= max(m1)
//   and if exactly L1 are, the next one is min(m2)  => at least L1+1 keys are <= max(max(m1), min(m2))
//   L2 lanes hold two keys     => at least 2*L2 keys are <= max(m2)
// The tightest applicable bound is used; otherwise every valid key survives.
template <typename KeyFn>
__device__ __forceinline__ int warp_select_smallest(int n, int k, KeyFn keyfn, WarpSelScratch *ws, uint16_t *members)
{
    const int lane = threadIdx.x & 31;
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
    const int nvalid = __reduce_add_sync(0xffffffffu, nv);
    const int K = min(k, nvalid);
    if (K == 0) return 0;
    const int L1 = __popc(__ballot_sync(0xffffffffu, m1 != KEY_NONE));
    const int L2 = __popc(__ballot_sync(0xffffffffu, m2 != KEY_NONE));
    u64 thr = KEY_NONE - 1;
    if (L1 >= K) thr = warp_max_u64(m1 != KEY_NONE ? m1 : 0ull);
    else if (L1 + 1 >= K && L2 >= 1) { const u64 a = warp_max_u64(m1 != KEY_NONE ? m1 : 0ull), b = warp_min_u64(m2); thr = a > b ? a : b; }
    else if (2 * L2 >= K) thr = warp_max_u64(m2 != KEY_NONE ? m2 : 0ull);
    // compaction of the survivors (key <= thr)
    int ns = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const u64 key = i < n ? keyfn(i) : KEY_NONE;
        const bool sv = key <= thr;                            // KEY_NONE > thr always
        const unsigned b = __ballot_sync(0xffffffffu, sv);
        if (sv) {
            const int pos = ns + __popc(b & ((1u << lane) - 1u));
            if (pos < WSEL_CAP) { ws->skey[pos] = key; ws->sidx[pos] = (uint16_t)i; }
        }
        ns += __popc(b);
    }
    __syncwarp();
    if (ns <= WSEL_CAP) {
        // exact rank among the survivors (two survivors per lane and pass: the broadcast loads are shared)
        for (int s = lane; s < ns; s += 64) {
            const u64 ka = ws->skey[s];
            const bool hb = s + 32 < ns;
            const u64 kb = hb ? ws->skey[s + 32] : 0ull;
            int ra = 0, rb = 0;
            for (int j = 0; j < ns; j++) { const u64 kj = ws->skey[j]; ra += kj < ka; rb += kj < kb; }
            if (ra < K) members[ra] = ws->sidx[s];
            if (hb && rb < K) members[rb] = ws->sidx[s + 32];
        }
    } else {
        // degenerate (flat content / thousands of equal costs): rank against every valid key
        for (int i = lane; i < n; i += 32) {
            const u64 key = keyfn(i);
            if (key > thr) continue;
            int rank = 0;
            for (int j = 0; j < n && rank < K; j++) rank += keyfn(j) < key;
            if (rank < K) members[rank] = (uint16_t)i;
        }
    }
    __syncwarp();
    return K;
}
