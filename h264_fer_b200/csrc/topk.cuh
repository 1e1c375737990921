// Warp-synchronous STREAMING "k smallest, in order" selection.
//
// The reference keeps insertion-sorted lists (moestimation.cpp:277-291); downstream only list MEMBERSHIP (the k smallest by
// (cost, arrival)) and the order among members matter. Round 1 stored every cost in shared memory and selected afterwards
// (~1500 warp instructions and 6 KB per partition in stage 3). Here candidates are offered as they are
// produced: a candidate is kept only if its key (cost << 16 | arrival index) is at or below a running bound, kept candidates go
// to a small buffer by a ballot-compacted append, and the bound tightens from the two smallest costs every lane has produced
// (the k-th smallest of the per-lane minima bounds the k-th smallest cost).
// Exact for any arrival order and any number of ties: when the buffer would overflow it is reduced to exactly its k smallest
// keys (bisection on the key), which makes the k-th of them the new bound.
#pragma once
#include "common.cuh"
#include "warp_select.cuh"

#define TK_NONE 0xffffffffu        // cost of "no candidate"

template <int CAP> struct TopKBufT { u64 key[CAP]; };   // buffered candidates of one warp (CAP >= 64 + k)
struct TopKBuf { u64 *key; };      // view
struct TopK {
    uint32_t m1, m2;               // this lane's two smallest costs so far
    u64 bound;                     // keys above this cannot be among the k smallest
    int ns;                        // buffered candidates (warp-uniform)
    int k, cap;
};

__device__ __forceinline__ void tk_init(TopK &t, int k, int cap) { t.m1 = t.m2 = TK_NONE; t.bound = ((u64)(TK_NONE - 1u) << 16) | 0xffffu; t.ns = 0; t.k = k; t.cap = cap; }

// ---- out-of-line helpers (kept out of the unrolled producers: instruction-cache footprint matters more than a call) ----
// number of buffered keys <= x (warp-uniform result)
__device__ __forceinline__ int tk_count_le(const u64 *b, int ns, u64 x)
{
    const int lane = threadIdx.x & 31;
    int c = 0;
#pragma unroll 1
    for (int i = lane; i < ns; i += 32) c += b[i] <= x;
    return __reduce_add_sync(0xffffffffu, c);
}
// keeps the buffered keys <= x (only the set is preserved); returns the new count
__device__ __noinline__ int tk_filter_ni(u64 *b, int ns, u64 x)
{
    const int lane = threadIdx.x & 31;
    int outn = 0;
#pragma unroll 1
    for (int base = 0; base < ns; base += 32) {
        const int i = base + lane;
        const u64 kv = i < ns ? b[i] : ~0ull;
        const bool keep = i < ns && kv <= x;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) b[outn + __popc(m & ((1u << lane) - 1u))] = kv;      // outn + rank <= i: never overtakes the reads
        outn += __popc(m);
        __syncwarp();
    }
    return outn;
}
// the k-th smallest buffered key (keys are unique; ns > k, every buffered key <= bound)
__device__ __noinline__ u64 tk_kth_ni(const u64 *b, int ns, int k, u64 bound)
{
    u64 lo = 0, hi = bound;
    while (lo < hi) {                                               // smallest x with count(keys <= x) >= k
        const u64 mid = lo + ((hi - lo) >> 1);
        if (tk_count_le(b, ns, mid) >= k) hi = mid; else lo = mid + 1;
    }
    return lo;
}
// Cost bound from the lanes' minima: with L1 lanes holding one cost and L2 holding two, at least L1 (L1+1, 2*L2) candidates lie
// at or below max m1 (max(max m1, min m2), max m2); with `exact`, the K-th smallest of the up to 64 minima themselves (bisection
// on the value, ~20 steps of two ballots). Returns TK_NONE when the minima do not bound anything yet. The bound only ever counts
// candidates that were really offered.
__device__ __noinline__ uint32_t tk_minima_bound_ni(uint32_t m1, uint32_t m2, int K, uint32_t cur, bool exact)
{
    const int L1 = __popc(__ballot_sync(0xffffffffu, m1 != TK_NONE));
    const int L2 = __popc(__ballot_sync(0xffffffffu, m2 != TK_NONE));
    uint32_t thr = TK_NONE;
    if (L1 >= K) thr = __reduce_max_sync(0xffffffffu, m1 != TK_NONE ? m1 : 0u);
    else if (L1 + 1 >= K && L2 >= 1) thr = max(__reduce_max_sync(0xffffffffu, m1 != TK_NONE ? m1 : 0u), __reduce_min_sync(0xffffffffu, m2));
    else if (2 * L2 >= K) thr = __reduce_max_sync(0xffffffffu, m2 != TK_NONE ? m2 : 0u);
    thr = min(thr, cur);
    if (exact && L1 + L2 >= K && thr != TK_NONE) {
        uint32_t lo = __reduce_min_sync(0xffffffffu, m1), hi = thr;
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            const int c = __popc(__ballot_sync(0xffffffffu, m1 <= mid)) + __popc(__ballot_sync(0xffffffffu, m2 <= mid));
            if (c >= K) hi = mid; else lo = mid + 1;
        }
        thr = lo;
    }
    return thr;
}
__device__ __forceinline__ void tk_apply_cost_bound(TopK &t, uint32_t thr)
{
    if (thr != TK_NONE) { const u64 nb = ((u64)thr << 16) | 0xffffu; if (nb < t.bound) t.bound = nb; }
}
__device__ __forceinline__ uint32_t tk_cost_bound(const TopK &t) { return (uint32_t)min(t.bound >> 16, (u64)TK_NONE); }
__device__ __forceinline__ void tk_tighten(TopK &t) { tk_apply_cost_bound(t, tk_minima_bound_ni(t.m1, t.m2, t.k, tk_cost_bound(t), false)); }
__device__ __forceinline__ void tk_tighten_exact_minima(TopK &t) { tk_apply_cost_bound(t, tk_minima_bound_ni(t.m1, t.m2, t.k, tk_cost_bound(t), true)); }
__device__ __forceinline__ void tk_filter(u64 *b, TopK &t, u64 x) { t.ns = tk_filter_ni(b, t.ns, x); }
// Reduces the buffer to exactly its min(k, ns) smallest keys and makes the largest of them the bound.
__device__ __forceinline__ void tk_exact(u64 *b, TopK &t)
{
    if (t.ns <= t.k) return;
    t.bound = tk_kth_ni(b, t.ns, t.k, t.bound);
    t.ns = tk_filter_ni(b, t.ns, t.bound);
}
// room for one more round of 32 appends
__device__ __forceinline__ void tk_make_room(u64 *b, TopK &t)
{
    tk_tighten(t);
    tk_filter(b, t, t.bound);
    if (t.ns > t.cap - 32) tk_exact(b, t);                         // ties / adversarial order: exact reduction (k < t.cap - 32)
}

// The two halves of offering a candidate. tk_track: the lane has produced this cost (bound bookkeeping only; once per candidate).
// tk_append: keep the candidate if its key is at or below the bound. A caller that holds its candidates in registers tracks them
// all, tightens, and only then appends — nothing that cannot be a member ever touches shared memory. All 32 lanes must call
// tk_append together (cost == TK_NONE: nothing); idx < 65536 is the arrival index.
__device__ __forceinline__ void tk_track(TopK &t, uint32_t cost) { t.m2 = min(t.m2, max(t.m1, cost)); t.m1 = min(t.m1, cost); }
__device__ __forceinline__ void tk_append(u64 *b, TopK &t, uint32_t cost, uint32_t idx)
{
    const int lane = threadIdx.x & 31;
    if (t.ns > t.cap - 32) tk_make_room(b, t);                      // (warp-uniform, rare)
    const u64 key = ((u64)cost << 16) | (u64)idx;
    const bool sv = cost != TK_NONE && key <= t.bound;
    const unsigned m = __ballot_sync(0xffffffffu, sv);
    if (m) {
        if (sv) b[FH_IDX(t.ns + __popc(m & ((1u << lane) - 1u)), t.cap)] = key;
        t.ns += __popc(m);
    }
}
__device__ __forceinline__ void tk_offer(u64 *b, TopK &t, uint32_t cost, uint32_t idx) { tk_append(b, t, cost, idx); tk_track(t, cost); }

// Ends the selection: members[r] (r < K) = arrival index of the candidate of rank r; returns K = min(k, nvalid), nvalid = number
// of real candidates offered. Ends with __syncwarp().
__device__ __forceinline__ int tk_finish(u64 *b, TopK &t, int nvalid, uint16_t *members, bool tightened = false)
{
    const int lane = threadIdx.x & 31;
    const int K = min(t.k, nvalid);
    if (K == 0) return 0;
    t.k = K;
    __syncwarp();
    if (!tightened) {                                               // (tightened: everything buffered is already at or below the final bound)
        tk_tighten(t);
        tk_tighten_exact_minima(t);
        tk_filter(b, t, t.bound);
    }
    if (t.ns > 96) tk_exact(b, t);
    const int ns = t.ns;
    // exact rank among the survivors (two survivors per lane and pass: the broadcast loads are shared)
    for (int s = lane; s < ns; s += 64) {
        const u64 ka = b[s];
        const bool hb = s + 32 < ns;
        const u64 kb = hb ? b[s + 32] : 0ull;
        int ra = 0, rb = 0;
#pragma unroll 2
        for (int j = 0; j < ns; j++) { const u64 kj = b[j]; ra += kj < ka; rb += kj < kb; }
        if (ra < K) members[FH_IDX(ra, FH_S3_MAX + 1)] = (uint16_t)(ka & 0xffffu);
        if (hb && rb < K) members[FH_IDX(rb, FH_S3_MAX + 1)] = (uint16_t)(kb & 0xffffu);
    }
    __syncwarp();
    return K;
}
