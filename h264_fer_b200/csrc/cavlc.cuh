// Device CAVLC of a P slice's slice_data. The reference writes the macroblock layer serially on
// the host (rbsp_encoding.cpp:175-313, residual.cpp:300-666); the only cross-macroblock state is (a) the nC context = the
// TotalCoeff of the left / upper 4x4 blocks (residual.cpp:421-503) and (b) mb_skip_run. Both depend on the quantised
// levels only, so the slice is coded in four parallel steps over the result records phase C left in HBM:
//   k_cavlc_prep   per MB: CodedBlockPattern + TotalCoeff of every coded block (CvInfo, 32 B)
//   k_cavlc_code   per MB: mb_skip_run, mb_type, mvds, cbp, residual blocks -> private bit buffer + bit length
//   k_cavlc_scan   per sequence: exclusive prefix sum of the bit lengths (+ the trailing mb_skip_run)
//   k_cavlc_pack   per MB (one warp): funnel the private bits into the slice at their bit offset (atomicOr on the words)
// The bit-exact coder itself is cavlc_core.h (also compiled for the host by the CPU tests).
#pragma once
#include "common.cuh"
#include "cavlc_core.h"

#define CV_MB_WORDS 448                 // private buffer per macroblock: 14336 bits (384 levels of the largest codable size fit)
#define CV_STREAM_BYTES 500000          // the reference's RBSP buffer (fer_h264.cpp:93): a slice that does not fit is an error there too
#define CV_FLAG_MB_OVERFLOW 1u
#define CV_FLAG_LEVEL_RANGE 2u
#define CV_FLAG_STREAM_OVERFLOW 4u

struct CvSeq {                          // per sequence, allocated on the first fh264_cavlc_p
    CvInfo *info;                       // nmb
    uint32_t *buf;                      // (nmb + 1) * CV_MB_WORDS   (slot nmb = trailing mb_skip_run)
    uint32_t *bits;                     // nmb + 1 bit lengths, then exclusive offsets in off[]
    uint32_t *off;                      // nmb + 2
    uint32_t *stream;                   // CV_STREAM_BYTES / 4 words (+ slack), memory byte order = stream order
    uint32_t *stat;                     // [0] flags, [1] total bits (first_bit included)
};

// par >= 0 (band mode, rank 0): the records are the gathered ones of that epoch parity instead of the sequence's own result buffer
__device__ __forceinline__ const fh264_mb_result *cv_records(const SeqDev &S, int par) { return par >= 0 ? S.gather[par] : S.results; }

__global__ void __launch_bounds__(128) k_cavlc_prep(const SeqDev *__restrict__ seqs, const CvSeq *__restrict__ cvs, int seq0, int nmb, int par)
{
    const int mb = blockIdx.x * 128 + threadIdx.x;
    if (mb >= nmb) return;
    if (seqs[seq0 + blockIdx.y].status[ST_GATE_DONE]) {            // scene cut: no P picture was coded (empty slice data)
        if (mb == 0) { cvs[seq0 + blockIdx.y].stat[0] = 0; cvs[seq0 + blockIdx.y].stat[1] = 0; }
        return;
    }
    const fh264_mb_result &r = cv_records(seqs[seq0 + blockIdx.y], par)[mb];
    CvInfo o;
    cv_prepare(r.mb_type, r.luma, r.chroma_dc, r.chroma_ac, FH264_P_SKIP, o);
    cvs[seq0 + blockIdx.y].info[mb] = o;
    if (mb == 0) { cvs[seq0 + blockIdx.y].stat[0] = 0; cvs[seq0 + blockIdx.y].stat[1] = 0; }
}

__global__ void __launch_bounds__(128) k_cavlc_code(const SeqDev *__restrict__ seqs, const CvSeq *__restrict__ cvs, int seq0, int nmb, int wmb, int par)
{
    const int mb = blockIdx.x * 128 + threadIdx.x;
    if (mb > nmb) return;
    const CvSeq &cv = cvs[seq0 + blockIdx.y];
    if (seqs[seq0 + blockIdx.y].status[ST_GATE_DONE]) { cv.bits[mb] = 0; return; }
    // mb_skip_run: the skipped macroblocks right before this one (rbsp_encoding.cpp:181-188); slot nmb = the run that ends the slice (:310)
    // (only a macroblock that writes something walks back, so every skipped macroblock is visited once per picture)
    int run = 0;
    if (mb == nmb || !cv.info[mb].skip)
        for (int m = mb - 1; m >= 0 && cv.info[m].skip; m--) run++;
    CvBits b;
    cv_init(b, cv.buf + (size_t)mb * CV_MB_WORDS, CV_MB_WORDS);
    int bad = 0;
    if (mb == nmb) { if (run > 0) cv_ue(b, (uint32_t)run); }
    else if (!cv.info[mb].skip) {
        const fh264_mb_result &r = cv_records(seqs[seq0 + blockIdx.y], par)[mb];
        const CvInfo me = cv.info[mb];
        CvInfo left, up;
        const bool hl = (mb % wmb) != 0, hu = mb >= wmb;
        if (hl) left = cv.info[mb - 1];
        if (hu) up = cv.info[mb - wmb];
        cv_macroblock(b, run, r.mb_type, r.num_parts, r.mvd, r.luma, r.chroma_dc, r.chroma_ac, me, hl ? &left : nullptr, hu ? &up : nullptr, &bad);
    }
    cv_flush(b);
    cv.bits[mb] = (uint32_t)cv_bits(b);
    const uint32_t fl = (b.ovf ? CV_FLAG_MB_OVERFLOW : 0u) | (bad ? CV_FLAG_LEVEL_RANGE : 0u);
    if (fl) atomicOr(&cv.stat[0], fl);
}

// one CTA per sequence: exclusive scan of nmb + 1 lengths
__global__ void __launch_bounds__(1024) k_cavlc_scan(const CvSeq *__restrict__ cvs, int seq0, int nmb, int first_bit)
{
    const CvSeq &cv = cvs[seq0 + blockIdx.x];
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base <= nmb; base += 1024) {
        const int i = base + tid;
        const uint32_t v = i <= nmb ? cv.bits[i] : 0u;
        uint32_t incl = v;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = wsum[lane], wi = w;
            for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += u; }
            wsum[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t ex = carry + wsum[warp] + incl - v;
        if (i <= nmb) cv.off[i] = ex;
        __syncthreads();
        if (tid == 1023) carry = ex + v;
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t total = (uint32_t)first_bit + carry;
        cv.stat[1] = total;
        if (total > (uint32_t)CV_STREAM_BYTES * 8u) atomicOr(&cv.stat[0], CV_FLAG_STREAM_OVERFLOW);
    }
}

// one warp per macroblock slot: OR its bits into the slice. Word i of the private buffer holds bits 32i.. MSB first; the slice
// words are stored byte-swapped so that the bytes in memory are in stream order.
__global__ void __launch_bounds__(128) k_cavlc_pack(const CvSeq *__restrict__ cvs, int seq0, int nmb, int first_bit)
{
    const int mb = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (mb > nmb) return;
    const CvSeq &cv = cvs[seq0 + blockIdx.y];
    const uint32_t nb = cv.bits[mb];
    if (nb == 0 || (cv.stat[0] & (CV_FLAG_STREAM_OVERFLOW | CV_FLAG_MB_OVERFLOW))) return;
    const uint32_t base = (uint32_t)first_bit + cv.off[mb];
    const uint32_t *src = cv.buf + (size_t)mb * CV_MB_WORDS;
    for (uint32_t i = lane; i * 32u < nb; i += 32) {
        const uint32_t w = src[i], p = base + 32u * i, wi = p >> 5, sh = p & 31u;
        atomicOr(&cv.stream[wi], __byte_perm(w >> sh, 0, 0x0123));
        if (sh) atomicOr(&cv.stream[wi + 1], __byte_perm(w << (32u - sh), 0, 0x0123));
    }
}
