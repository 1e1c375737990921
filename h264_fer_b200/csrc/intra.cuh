// I pictures on the device: the reference's I-slice macroblock loop (rbsp_encoding.cpp:196-215,
// intraPredictionEncoding intra.cpp:949-1109) as a wavefront over the macroblocks of a picture. A macroblock predicts from the
// reconstruction of its left, upper, upper-left and upper-right neighbours and its bit-cost trials read their CAVLC state, so
// the picture is swept in the same anti-diagonal ticket order as phase B (a ticket's dependencies always hold smaller tickets:
// no co-residency assumption). One warp per macroblock: the two mode searches are spread over the lanes, the serial remainder
// (trials, block-by-block Intra4x4 coding, final transform) runs on lane 0. The per-macroblock arithmetic is intra_core.h.
#pragma once
#include "common.cuh"
#include "intra_core.h"
#include "cavlc.cuh"

struct IntraSeq {
    IcInfo *info;           // nmb: state of the macroblocks of the picture being coded
    uint32_t *done;         // nmb: epoch of the last I picture in which the macroblock was completed
};

static_assert(sizeof(fh264_mb_result_i) == sizeof(fh264_mb_result), "I records share the result buffer");
static_assert(sizeof(IcInfo) == 48, "IcInfo size");

// (also clears a wavefront timeout left by an earlier picture: intra_wait gives up early while it is set)
__global__ void k_begin_intra(uint32_t *ticket, SeqDev *seqs, int seq0)
{
    seqs[seq0 + threadIdx.x].status[ST_FLAGS] &= ~FLAG_TIMEOUT;
    if (threadIdx.x == 0) *ticket = 0;
}

// Bounded wait for a neighbour's completion flag; gives up at once when another wait of this picture has already timed out
// (the picture is lost anyway and must drain quickly).
__device__ __forceinline__ bool intra_wait(const uint32_t *p, uint32_t epoch, const uint32_t *status)
{
    for (unsigned it = 0; it < (1u << 23); it++) {
        if (ld_acquire_u32(p) >= epoch) return true;
        if ((it & 1023u) == 1023u && (ld_acquire_u32(status + ST_FLAGS) & FLAG_TIMEOUT)) return false;
        __nanosleep(40);
    }
    return false;
}

__global__ void __launch_bounds__(32) k_intra(const SeqDev *__restrict__ seqs, const IntraSeq *__restrict__ iseqs, const int *__restrict__ prev_p,
                                              int seq0, int nseq, Geo g, int qp, uint32_t epoch, const int *__restrict__ wf_order,
                                              uint32_t *__restrict__ ticket, int nl)
{
    __shared__ IcCtx c;
    __shared__ uint32_t sh_ticket;
    const int lane = threadIdx.x;
    const uint32_t total = (uint32_t)g.nmb * (uint32_t)nseq;
    for (;;) {
        __syncwarp();
        if (lane == 0) sh_ticket = atomicAdd(ticket, 1u);
        __syncwarp();
        const uint32_t t = sh_ticket;
        if (t >= total) return;
        const int b = seq0 + (int)(t % (uint32_t)nseq);
        const SeqDev &S = seqs[b];
        const IntraSeq &I = iseqs[b];
        const int mb = wf_order[t / (uint32_t)nseq];
        const int mbx = mb % g.Wmb, mby = mb / g.Wmb;
        // mb_type_array[CurrMbAddr] of the previous picture (read before this macroblock's record replaces it)
        // (band mode: the records of the other bands live on other GPUs; phase C mirrored every macroblock's type into `motion`)
        const bool prev_skip = prev_p[b] && (g.world > 1 ? S.motion[mb].mb_type : S.results[mb].mb_type) == FH264_P_SKIP;
        if (lane == 0) {
            for (int k = 0; k < 3; k++) { c.src[k] = S.cur[k]; c.rec[k] = S.rec[k]; }
            c.W = g.W; c.H = g.H; c.xP = mbx * 16; c.yP = mby * 16; c.qp = qp;
        }
        // neighbours complete? (every lane acquires, so every lane may read what they wrote)
        bool ok = true;
        if (mbx > 0) ok &= intra_wait(&I.done[mb - 1], epoch, S.status);
        if (mby > 0) {
            ok &= intra_wait(&I.done[mb - g.Wmb], epoch, S.status);
            if (mbx > 0) ok &= intra_wait(&I.done[mb - g.Wmb - 1], epoch, S.status);
            if (mbx + 1 < g.Wmb) ok &= intra_wait(&I.done[mb - g.Wmb + 1], epoch, S.status);
        }
        if (!ok && lane == 0) atomicOr(&S.status[ST_FLAGS], FLAG_TIMEOUT);
        __syncwarp();
        if (lane < nl)
            ic_macroblock(c, prev_skip, mbx > 0 ? &I.info[mb - 1] : nullptr, mby > 0 ? &I.info[mb - g.Wmb] : nullptr,
                          *(fh264_mb_result_i *)&S.results[mb], I.info[mb], lane, nl);
        __threadfence();
        __syncwarp();
        if (lane == 0) { __threadfence(); st_release_u32(&I.done[mb], epoch); }
    }
}

// slice_data() of an I picture from the records fh264_encode_i left (SURVEY.md §8(f) ranks 1 + 2): every macroblock codes itself
// into its private bit buffer (ic_write_macroblock; the nC contexts come from the neighbours' records), k_cavlc_scan and
// k_cavlc_pack of the P path then concatenate the buffers. Slot nmb stays empty (no mb_skip_run in an I slice).
__global__ void __launch_bounds__(128) k_cavlc_code_i(const SeqDev *__restrict__ seqs, const CvSeq *__restrict__ cvs, int seq0, int nmb, int wmb)
{
    const int mb = blockIdx.x * 128 + threadIdx.x;
    if (mb > nmb) return;
    const CvSeq &cv = cvs[seq0 + blockIdx.y];
    if (mb == nmb) { cv.bits[mb] = 0; return; }
    const fh264_mb_result_i *rec = (const fh264_mb_result_i *)seqs[seq0 + blockIdx.y].results;
    IcInfo me, left, up;
    const bool hl = (mb % wmb) != 0, hu = mb >= wmb;
    ic_info_from_record(rec[mb], me);
    if (hl) ic_info_from_record(rec[mb - 1], left);
    if (hu) ic_info_from_record(rec[mb - wmb], up);
    CvBits b;
    cv_init(b, cv.buf + (size_t)mb * CV_MB_WORDS, CV_MB_WORDS);
    int bad = 0;
    ic_write_macroblock(b, rec[mb], me, hl ? &left : nullptr, hu ? &up : nullptr, &bad);
    cv_flush(b);
    cv.bits[mb] = (uint32_t)cv_bits(b);
    const uint32_t fl = (b.ovf ? CV_FLAG_MB_OVERFLOW : 0u) | (bad ? CV_FLAG_LEVEL_RANGE : 0u);
    if (fl) atomicOr(&cv.stat[0], fl);
}
