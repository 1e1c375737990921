// 64-bit selection keys and their warp reductions. The reference keeps insertion-sorted lists (moestimation.cpp:277-291);
// downstream only list MEMBERSHIP (the k smallest by (cost, arrival)) and the order among members matter, so lists are rebuilt
// from UNIQUE keys (cost in the high bits, arrival index in the low bits; ~0 marks "no candidate"): streaming in topk.cuh,
// block-cooperative in phase_b.cuh (block_select_smallest).
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
