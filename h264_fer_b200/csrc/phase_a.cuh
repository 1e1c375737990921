// Phase A — per current picture, fully parallel over 8x8 partitions, independent of the MV predictors.
//   k_stage3  the complete stage-3 list (moestimation.cpp:508-520): feature costs over the +-window/2 integer
//             window and the +-window/16 quarter-pel window around (0,0), the 33 best by (cost, arrival), their SADs.
//   k_stage2  the stage-2 candidate SET (moestimation.cpp:470-497): positions whose 8x8 sum is within +-j_stop of
//             the block's, gated by Manhattan distance and the two half-sums, in the reference's arrival order,
//             with feature distance and SAD. Only the multiplier (|dx-genx|+|dy-geny|+4) is left to phase B.
// One CTA (128 threads) per partition.
#pragma once
#include "common.cuh"
#include "select.cuh"

#define PA_NT 128
#define S3_COST_CAP 5632      // (64+1)^2 + (2*4+1)^2*16 = 5521 candidates at WindowSize 64
#define COST_INVALID 0xffffffffu

__device__ __forceinline__ void part_origin(const Geo &g, int part, int &xP, int &yP)
{
    const int mb = part >> 2, pi = part & 3;
    xP = (mb % g.Wmb) * 16 + (pi & 1) * 8;
    yP = (mb / g.Wmb) * 16 + (pi >> 1) * 8;
}

__device__ __forceinline__ int feat_at(const uint16_t *__restrict__ kar, const Geo &g, const int s[5], int f, int x, int y)
{
    const uint16_t *K = kar + (size_t)f * 5 * g.WH + (size_t)y * g.W + x;
    return feat_dist(s, __ldg(K), __ldg(K + g.WH), __ldg(K + 2 * (size_t)g.WH), __ldg(K + 3 * (size_t)g.WH), __ldg(K + 4 * (size_t)g.WH));
}

__global__ void __launch_bounds__(PA_NT) k_stage3(const SeqDev *__restrict__ seqs, int seq0, Geo g, fh264_params prm)
{
    __shared__ uint32_t cost[S3_COST_CAP];
    __shared__ uint2 currow[8];
    __shared__ SelectScratch sc;
    __shared__ uint32_t mem_idx[FH_S3_MAX], mem_sad[FH_S3_MAX];
    __shared__ int n_mem, n_valid;
    const SeqDev &S = seqs[seq0 + blockIdx.y];
    const int part = blockIdx.x, tid = threadIdx.x;
    int xP, yP;
    part_origin(g, part, xP, yP);
    if (tid < 8) currow[tid] = *(const uint2 *)(S.cur[0] + (size_t)(yP + tid) * g.W + xP);
    if (tid == 0) { n_mem = 0; n_valid = 0; }
    __syncthreads();
    int s[5];
    { uint2 rows[8];
#pragma unroll
      for (int r = 0; r < 8; r++) rows[r] = currow[r];
      block_sums(rows, s); }
    const int g3 = prm.window / 2, g1 = prm.window / 16;
    const int w3 = 2 * g3 + 1, w1 = 2 * g1 + 1;
    const int n3a = w3 * w3, n3b = w1 * w1 * 16, N = n3a + n3b;
    int valid = 0;
    // first call: MEstimation(g = window/2, frac 0) — threads walk row-major (coalesced), arrival index is x-major
    for (int t = tid; t < n3a; t += PA_NT) {
        const int r = t / w3, c = t - r * w3, dx = c - g3, dy = r - g3;
        const int rx = xP + dx, ry = yP + dy;
        uint32_t cst = COST_INVALID;
        if (rx >= 0 && rx < g.W && ry >= 0 && ry < g.H) { cst = (uint32_t)((iabs_(dx) + iabs_(dy) + 4) * feat_at(S.kar, g, s, 0, rx, ry)); valid++; }
        cost[c * w3 + r] = cst;
    }
    // second call: MEstimation(g = window/16, all 16 fractions); arrival = (dx, dy, frac)
    for (int t = tid; t < n3b; t += PA_NT) {
        const int f = t & 15, pos = t >> 4, dx = pos / w1 - g1, dy = pos % w1 - g1;
        const int rx = xP + dx, ry = yP + dy;
        uint32_t cst = COST_INVALID;
        if (rx >= 0 && rx < g.W && ry >= 0 && ry < g.H) { cst = (uint32_t)((iabs_(dx) + iabs_(dy) + 4) * feat_at(S.kar, g, s, f, rx, ry)); valid++; }
        cost[n3a + t] = cst;
    }
    if (valid) atomicAdd(&n_valid, valid);
    __syncthreads();
    const int K = min(FH_S3_MAX, n_valid);
    if (K > 0) {
        int lt, lt2;
        const uint32_t T = block_kth_smallest<uint32_t, PA_NT>(N, K, 24, [&](int i) { return cost[i]; }, &sc, &lt);
        // ties at the threshold enter in arrival order: the (K - lt) smallest indices among cost == T
        const uint32_t Ti = block_kth_smallest<uint32_t, PA_NT>(N, K - lt, 13, [&](int i) { return cost[i] == T ? (uint32_t)i : COST_INVALID; }, &sc, &lt2);
        for (int i = tid; i < N; i += PA_NT) {
            const uint32_t c = cost[i];
            if (c < T || (c == T && (uint32_t)i <= Ti)) mem_idx[atomicAdd(&n_mem, 1)] = (uint32_t)i;
        }
    }
    __syncthreads();
    const int nm = n_mem;
    // SADs of the members (satdLuma8x8MVs): 4 threads per member, 2 rows each
    for (int base = 0; base < nm; base += PA_NT / 4) {
        const int m = base + (tid >> 2);
        int sad = 0;
        if (m < nm) {
            const int i = (int)mem_idx[m];
            int dx, dy, f;
            if (i < n3a) { dx = i / w3 - g3; dy = i % w3 - g3; f = 0; }
            else { const int t = i - n3a; f = t & 15; dx = (t >> 4) / w1 - g1; dy = (t >> 4) % w1 - g1; }
            const uint8_t *pl = S.planes + (size_t)f * g.WH;
            const int r0 = (tid & 3) * 2;
            sad = sad_row8(currow[r0], pl, g.W, g.H, xP + dx, yP + dy + r0) + sad_row8(currow[r0 + 1], pl, g.W, g.H, xP + dx, yP + dy + r0 + 1);
        }
        sad += __shfl_xor_sync(0xffffffffu, sad, 1);
        sad += __shfl_xor_sync(0xffffffffu, sad, 2);
        if (m < nm && (tid & 3) == 0) mem_sad[m] = (uint32_t)sad;
    }
    __syncthreads();
    // list order = (cost, arrival index); rank by counting among <= 33 members
    if (tid < nm) {
        const uint32_t i = mem_idx[tid], c = cost[i];
        int rank = 0;
        for (int k = 0; k < nm; k++) { const uint32_t ik = mem_idx[k], ck = cost[ik]; rank += (ck < c) || (ck == c && ik < i); }
        int dx, dy, f;
        if ((int)i < n3a) { dx = (int)i / w3 - g3; dy = (int)i % w3 - g3; f = 0; }
        else { const int t = (int)i - n3a; f = t & 15; dx = (t >> 4) / w1 - g1; dy = (t >> 4) % w1 - g1; }
        S3Entry e;
        e.mvx = (int16_t)((dx << 2) | (f & 3)); e.mvy = (int16_t)((dy << 2) | (f >> 2)); e.sad = (uint16_t)mem_sad[tid]; e.pad = 0;
        S.s3[(size_t)part * FH_S3_MAX + rank] = e;
    }
    if (tid == 0) {
        PartA *pa = &S.parta[part];
#pragma unroll
        for (int k = 0; k < 5; k++) pa->suma[k] = (uint16_t)s[k];
        pa->n3 = (uint16_t)nm;
    }
}

__global__ void __launch_bounds__(PA_NT) k_stage2(const SeqDev *__restrict__ seqs, int seq0, Geo g, fh264_params prm)
{
    __shared__ uint32_t akey[FH_S2_SMEM_CAP];    // arrival key: j<<21 | side<<20 | (dx+279)<<10 | (dy+279)
    __shared__ uint32_t afeat[FH_S2_SMEM_CAP];
    __shared__ uint16_t asad[FH_S2_SMEM_CAP];
    __shared__ uint32_t jcount[192];
    __shared__ uint2 currow[8];
    __shared__ int n_surv, j_stop_s, n2_s, cnt0_s;
    __shared__ uint32_t pool_off;
    const SeqDev &S = seqs[seq0 + blockIdx.y];
    const int part = blockIdx.x, tid = threadIdx.x;
    int xP, yP;
    part_origin(g, part, xP, yP);
    if (tid < 8) currow[tid] = *(const uint2 *)(S.cur[0] + (size_t)(yP + tid) * g.W + xP);
    for (int i = tid; i < 192; i += PA_NT) jcount[i] = 0;
    if (tid == 0) n_surv = 0;
    __syncthreads();
    int s[5];
    { uint2 rows[8];
#pragma unroll
      for (int r = 0; r < 8; r++) rows[r] = currow[r];
      block_sums(rows, s); }
    // tiles intersecting the bounding box of the diamond |dx|+|dy| < 280 (moestimation.cpp:481)
    const int tx0 = max(0, xP - 279) >> FH_TILE_SHIFT, tx1 = min(g.W - 1, xP + 279) >> FH_TILE_SHIFT;
    const int ty0 = max(0, yP - 279) >> FH_TILE_SHIFT, ty1 = min(g.H - 1, yP + 279) >> FH_TILE_SHIFT;
    const int ntx = tx1 - tx0 + 1, nty = ty1 - ty0 + 1;
    const int qlo = max(0, s[0] - 180) >> 7, qhi = min(16383, s[0] + 180) >> 7, nq = qhi - qlo + 1;
    const int k1lo = max(0, s[1] - 99) >> 6, k1hi = min(8191, s[1] + 99) >> 6;
    const int items = ntx * nty * nq;
    for (int it = tid; it < items; it += PA_NT) {
        const int q = qlo + it % nq, tt = it / nq;
        const int tx = tx0 + tt % ntx, ty = ty0 + tt / ntx;
        // closest point of the tile rectangle to the block origin
        const int rx0 = tx << FH_TILE_SHIFT, ry0 = ty << FH_TILE_SHIFT;
        const int ddx = max(0, max(rx0 - xP, xP - (rx0 + FH_TILE - 1))), ddy = max(0, max(ry0 - yP, yP - (ry0 + FH_TILE - 1)));
        if (ddx + ddy >= 280) continue;
        const int tile = ty * g.tilesx + tx;
        const uint16_t *ts = S.tstart + (size_t)tile * FH_TSTART_PITCH;
        const int e0 = __ldg(&ts[q * 128 + k1lo]), e1 = __ldg(&ts[q * 128 + k1hi + 1]);
        const uint4 *te = (const uint4 *)(S.tent + (size_t)tile * (FH_TILE * FH_TILE));
        for (int e = e0; e < e1; e++) {
            const uint4 v = __ldg(&te[e]);
            const int x = v.x & 0xffff, y = v.x >> 16, k0 = v.y & 0xffff, k1 = v.y >> 16, k2 = v.z & 0xffff, k3 = v.z >> 16, k4 = v.w & 0xffff;
            const int j = iabs_(k0 - s[0]), dx = x - xP, dy = y - yP;
            if (j <= 180 && iabs_(dx) + iabs_(dy) < 280 && iabs_(k1 - s[1]) < 100 && iabs_(k2 - s[2]) < 100) {
                atomicAdd(&jcount[j], j == 0 ? 2u : 1u);       // both sides visit bucket s0 when j == 0 (:476,486)
                const int pos = atomicAdd(&n_surv, 1);
                if (pos < FH_S2_SMEM_CAP) {
                    akey[pos] = ((uint32_t)j << 21) | ((uint32_t)(k0 > s[0]) << 20) | ((uint32_t)(dx + 279) << 10) | (uint32_t)(dy + 279);
                    afeat[pos] = (uint32_t)feat_dist(s, k0, k1, k2, k3, k4);
                }
            }
        }
    }
    __syncthreads();
    // j_stop: first j at which the running gated count exceeds 128 (:496), else 180
    if (tid < 32) {
        uint32_t c[6], local = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) { c[i] = jcount[tid * 6 + i]; local += c[i]; }
        uint32_t incl = local;
        for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (tid >= d) incl += v; }
        uint32_t run = incl - local;
        int first = 1 << 20;
#pragma unroll
        for (int i = 0; i < 6; i++) { run += c[i]; if (run > 128 && first == (1 << 20)) first = tid * 6 + i; }
        for (int d = 16; d; d >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, d));
        const int js = min(first, 180);
        // candidates kept: all gated entries with j <= j_stop
        uint32_t upto = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) if (tid * 6 + i <= js) upto += c[i];
        for (int d = 16; d; d >>= 1) upto += __shfl_xor_sync(0xffffffffu, upto, d);
        if (tid == 0) { j_stop_s = js; n2_s = (int)upto; cnt0_s = (int)(jcount[0] >> 1); }
    }
    __syncthreads();
    const int ns = n_surv, js = j_stop_s, cnt0 = cnt0_s;
    int n2 = n2_s;
    if (tid == 0) {
        uint32_t off = 0;
        if (ns > FH_S2_SMEM_CAP || n2 > 1023) { atomicOr(&S.status[ST_FLAGS], FLAG_CAPACITY); n2 = 0; }
        else if (n2 > 0) {
            off = atomicAdd(&S.status[ST_S2CURSOR], (uint32_t)n2);
            if (off + (uint32_t)n2 > S.s2pool_size) { atomicOr(&S.status[ST_FLAGS], FLAG_CAPACITY); n2 = 0; }
        }
        pool_off = off; n2_s = n2;
        PartA *pa = &S.parta[part];
        pa->s2_off = off; pa->n2 = (uint32_t)n2;
    }
    __syncthreads();
    n2 = n2_s;
    if (n2 == 0) return;
    // SAD at integer displacement (fraction 0 => plane 0): 4 threads per survivor
    const uint8_t *pl = S.planes;
    for (int base = 0; base < ns; base += PA_NT / 4) {
        const int m = base + (tid >> 2);
        int sad = 0;
        const bool live = m < ns && (int)(akey[m < ns ? m : 0] >> 21) <= js;
        if (live) {
            const uint32_t k = akey[m];
            const int dx = (int)((k >> 10) & 1023) - 279, dy = (int)(k & 1023) - 279, r0 = (tid & 3) * 2;
            sad = sad_row8(currow[r0], pl, g.W, g.H, xP + dx, yP + dy + r0) + sad_row8(currow[r0 + 1], pl, g.W, g.H, xP + dx, yP + dy + r0 + 1);
        }
        sad += __shfl_xor_sync(0xffffffffu, sad, 1);
        sad += __shfl_xor_sync(0xffffffffu, sad, 2);
        if (live && (tid & 3) == 0) asad[m] = (uint16_t)sad;
    }
    __syncthreads();
    // arrival order (:474-495): j ascending; minus side before plus side; x then y inside a bucket; bucket s0 twice
    uint2 *pool = S.s2pool + pool_off;
    for (int i = tid; i < ns; i += PA_NT) {
        const uint32_t k = akey[i];
        if ((int)(k >> 21) > js) continue;
        int rank = 0;
        for (int m = 0; m < ns; m++) rank += akey[m] < k;     // entries beyond j_stop have larger keys: never counted
        const int dx = (int)((k >> 10) & 1023) - 279, dy = (int)(k & 1023) - 279;
        const uint2 v = make_uint2(((uint32_t)dx & 0xffffu) | ((uint32_t)dy << 16), afeat[i] | ((uint32_t)asad[i] << 18));
        if ((k >> 21) == 0) { pool[rank] = v; pool[cnt0 + rank] = v; }
        else pool[cnt0 + rank] = v;
    }
}
