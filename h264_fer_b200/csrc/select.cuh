// Block-wide k-th smallest by most-significant-digit radix selection (8-bit digits, shared-memory histogram).
// The reference keeps sorted insertion lists (moestimation.cpp:277-291); only list MEMBERSHIP (the k smallest by
// (cost, arrival order)) and the order among members matter downstream, so the lists are rebuilt from a
// threshold instead of being maintained by insertion.
#pragma once
#include "common.cuh"

struct SelectScratch { uint32_t hist[256]; uint32_t digit, below; };

// keyfn(i) -> key of element i (only the low `bits` bits may be set), or ~0 to exclude the element.
// Returns T with  #{key < T} < k <= #{key <= T}  over the included elements (k >= 1 and k <= #included).
// *count_lt receives #{key < T}. All NT threads must call. Ends with a barrier.
template <typename KeyT, int NT, typename KeyFn>
__device__ KeyT block_kth_smallest(int n, int k, int bits, KeyFn keyfn, SelectScratch *sc, int *count_lt)
{
    const int tid = threadIdx.x;
    KeyT prefix = 0, mask = 0;
    int lt = 0;
    for (int shift = ((bits + 7) / 8 - 1) * 8; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += NT) sc->hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += NT) {
            const KeyT key = keyfn(i);
            if (key != (KeyT)~(KeyT)0 && (key & mask) == prefix) atomicAdd(&sc->hist[(uint32_t)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            uint32_t c[8], local = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) { c[j] = sc->hist[tid * 8 + j]; local += c[j]; }
            uint32_t incl = local;
            for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (tid >= d) incl += v; }
            uint32_t ex = incl - local;
            if ((uint32_t)k > ex && (uint32_t)k <= incl) {
                uint32_t run = ex;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if ((uint32_t)k > run && (uint32_t)k <= run + c[j]) { sc->digit = tid * 8 + j; sc->below = run; }
                    run += c[j];
                }
            }
        }
        __syncthreads();
        prefix |= (KeyT)sc->digit << shift;
        mask |= (KeyT)255 << shift;
        k -= (int)sc->below;
        lt += (int)sc->below;
        __syncthreads();
    }
    *count_lt = lt;
    return prefix;
}
