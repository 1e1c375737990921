/*
 * fh264_b200 — C ABI of the B200 (sm_100a) implementation of the encoder hot path of zoltanmaric/h264-fer (P pictures; I pictures
 * through fh264_encode_i).
 *
 * This is the drop-in boundary: plain C, opaque handle, plain pointers and sizes, int status codes.
 * Each entry point names the reference interface it stands in for (paths relative to the reference's
 * fer_h264/fer_h264/). The reference communicates through process globals (h264_globals.h:99-193,
 * residual.h:6-15, mode_pred.h:19-22); INTEGRATION.md shows the C++ shim that copies fh264_mb_result
 * records into those globals so the unchanged host CAVLC/NAL code emits the bitstream.
 *
 * A session owns one GPU and `batch` independent sequences that advance in lockstep (one picture of every
 * sequence per call), because the reference itself is one-sequence-per-process (function-local statics,
 * fer_h264.cpp:55-79) and a single 1080p picture cannot fill a B200's wavefront-limited phase.
 *
 * Per picture and sequence, in the order the reference's RBSP_encode()/encode() issue them:
 *   fh264_upload_source   <- ReadFromY4M() filling `frame`                         (fileIO.cpp:258-345)
 *   fh264_scene_sad       <- selectNALUnitType()'s luma |frame-dpb| sum             (ref_frames.cpp:185-234)
 *   fh264_encode_p        <- the P-slice MB loop: interEncoding() + quantizationTransform()/
 *                            transformDecodingP_Skip() for every MB                (rbsp_encoding.cpp:175-192),
 *                            then modificationProcess()/frameDeepCopy() and
 *                            FillInterpolatedRefFrame() for the next picture       (rbsp_encoding.cpp:317-322)
 *   fh264_upload_recon    <- after a host-coded I picture: frame -> dpb, then FillInterpolatedRefFrame()
 *   fh264_encode_i        <- the I-slice MB loop: intraPredictionEncoding() + quantizationTransform() for every MB
 *                            (rbsp_encoding.cpp:196-215), then the same dpb copy and reference preparation
 * All arithmetic is integer; results are bit-exact with the reference CPU path.
 */
#ifndef FH264_B200_H
#define FH264_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FH264_ABI_VERSION 1

/* status codes (the reference has none: void functions + assert, openCL_functions.cpp:60-139) */
enum {
    FH264_OK = 0,
    FH264_E_ARG = -1,          /* bad argument (null pointer, size not a multiple of 16, > 10000 MBs, ...) */
    FH264_E_CUDA = -2,         /* CUDA runtime error, see fh264_last_error() */
    FH264_E_NO_DEVICE = -3,    /* no usable sm_100 device: the library never falls back to a CPU path */
    FH264_E_STATE = -4,        /* call order violated (e.g. encode_p before any reference picture) */
    FH264_E_UB_INPUT = -5,     /* reference-undefined input: an 8x8 reference window sums to 0 or >= 16203
                                  (moestimation.cpp:153-158,477-480) */
    FH264_E_CAPACITY = -6,     /* a fixed-size buffer cannot hold the picture's data */
    FH264_E_UNSUPPORTED = -7   /* parameter outside the supported range (WindowSize > 64) */
};

/* mb_type values as the reference stores them in mb_type_array[] (h264_globals.h:24-28,58) */
enum { FH264_P_L0_16x16 = 0, FH264_P_L0_L0_16x8 = 1, FH264_P_L0_L0_8x16 = 2, FH264_P_8x8ref0 = 4, FH264_P_SKIP = 31 };

/* Encoder parameters = Starter::PostaviParametre (fer_h264.cpp:169-178), the subset the P path reads. */
typedef struct fh264_params {
    int32_t qp;           /* _qParameter == QPy (headers_and_parameter_sets.cpp:186-189), 0..51 */
    int32_t window;       /* WindowSize: stage 3 scans +-window/2, stages 1/3b +-window/16 (moestimation.cpp:458,509-510) */
    int32_t maxdiff_set;  /* MAXDIFF_SET; -1 = adaptive per macroblock (moestimation.cpp:407-419) */
    int32_t basic;        /* BasicInterEncoding: 1 = stage 1 only (moestimation.cpp:470) */
} fh264_params;

/* One macroblock's outputs = the globals interEncoding()/quantizationTransform() leave behind:
 * mb_type, mvL0x/mvL0y[CurrMbAddr][q][0] (quadrant MVs, quarter-pel), mvd_l0[part][0][], the 8x8 SADs the
 * search measured for the chosen MVs (satdLuma8x8MVs, moestimation.cpp:175-195), LumaLevel[16][16]
 * (z-order 4x4 blocks, zigzag), ChromaDCLevel[2][4], ChromaACLevel[2][4][0..14]. 832 bytes (13 x 64). */
typedef struct fh264_mb_result {
    int16_t mb_type;
    int16_t num_parts;          /* NumMbPart(mb_type): 1, 2 or 4; 0 for P_Skip */
    int16_t mv[4][2];           /* per 8x8 quadrant: {x, y} */
    int16_t mvd[4][2];          /* per partition (first num_parts valid, rest 0) */
    uint16_t sad[4];            /* per 8x8 quadrant; 0 for P_Skip */
    int16_t luma[16][16];
    int16_t chroma_dc[2][4];
    int16_t chroma_ac[2][4][15];
    int16_t reserved[10];
} fh264_mb_result;

typedef struct fh264_session fh264_session;

/* Lifecycle seam. Replaces InitCL()/AllocateFrameBuffersCL()/InitializeInterpolatedRefFrame()/AllocateMemory()
 * (openCL_functions.cpp:52,142; moestimation.cpp:29; mode_pred.cpp:22). width/height are the CODED luma size
 * (multiples of 16, <= 10000 MBs as h264_globals.cpp:180). device = CUDA ordinal. */
int fh264_open(int width, int height, int batch, int device, fh264_session **out);
int fh264_close(fh264_session *s);                                   /* CloseCL() openCL_functions.cpp:162 */
const char *fh264_last_error(void);
int fh264_abi_version(void);

/* Run on a caller-owned CUDA stream (cudaStream_t passed as void*), or NULL for the session's own stream. */
int fh264_set_stream(fh264_session *s, void *cuda_stream);
int fh264_sync(fh264_session *s);

/* Pinned host staging for the asynchronous calls (cudaHostAlloc / cudaFreeHost). */
void *fh264_host_alloc(size_t bytes);
void fh264_host_free(void *p);

/* `frame` := source picture of sequence `seq` (planar 4:2:0, row-major, stride == width). Asynchronous; the host buffers
 * must stay valid until the next fh264_sync()/synchronous call. Source pictures are double buffered:
 * the copy runs on the session's upload stream into the buffer that is not being coded, so the picture for t+1 may be handed
 * over right after fh264_encode_p_async(t) and its H2D overlaps the coding of t; the first call that reads `frame`
 * (fh264_scene_sad*, fh264_encode_p*) waits for the copy and makes it current. */
int fh264_upload_source(fh264_session *s, int seq, const uint8_t *y, const uint8_t *cb, const uint8_t *cr);

/* Same for one raw Y4M FRAME payload (Y in_w x in_h, then Cb and Cr at half size) of ANY input size that crops to the session's
 * coded size: the centre crop of ReadFromY4M (fileIO.cpp:286-337: rows / columns from (in - coded) >> 1, chroma from that
 * offset >> 1) is done by the copy engine (strided H2D), no host-side repacking. */
int fh264_upload_source_frame(fh264_session *s, int seq, const uint8_t *frame420, int in_w, int in_h);

/* Same, from planes already resident in device memory (device pointers; device-to-device copy on the stream). */
int fh264_upload_source_device(fh264_session *s, int seq, const void *dy, const void *dcb, const void *dcr);

/* dpb := reconstruction of a picture coded elsewhere (host I picture), then phase R on it
 * (frameDeepCopy ref_frames.cpp:17-35 + FillInterpolatedRefFrame moestimation.cpp:74-173). */
int fh264_upload_recon(fh264_session *s, int seq, const uint8_t *y, const uint8_t *cb, const uint8_t *cr);

/* Sum over luma of |frame - dpb| for sequence seq (selectNALUnitType ref_frames.cpp:210-224; the reference's
 * OpenCL AbsDiff kernel h264_kernels.cl:1-5 + host sum). Synchronous (returns the value). */
int fh264_scene_sad(fh264_session *s, int seq, uint64_t *sad);
/* The same for sequences [seq0, seq0+nseq) with one launch and one synchronisation. */
int fh264_scene_sad_batch(fh264_session *s, int seq0, int nseq, uint64_t *sads);

/* Code one P picture of sequences [seq0, seq0+nseq): motion search, mode decision, motion compensation,
 * pixel snapping, transform/quant/reconstruction for every MB; then dpb := reconstruction and phase R for the
 * next picture. results: nseq * (width/16*height/16) records (host memory, pinned preferred), MB raster order
 * per sequence; may be NULL (results stay on the device; see fh264_download_results). Synchronous. */
int fh264_encode_p(fh264_session *s, int seq0, int nseq, const fh264_params *p, fh264_mb_result *results);

/* Asynchronous variant: enqueues everything (including the D2H of results if non-NULL) and returns. */
int fh264_encode_p_async(fh264_session *s, int seq0, int nseq, const fh264_params *p, fh264_mb_result *results);

/* Status of the last encode_p of sequence seq (FH264_OK, FH264_E_UB_INPUT, FH264_E_CAPACITY); synchronous. */
int fh264_picture_status(fh264_session *s, int seq);

/* Current dpb (== `frame` after RBSP_encode) of sequence seq. Tests / decoder-side use. Synchronous. */
int fh264_download_recon(fh264_session *s, int seq, uint8_t *y, uint8_t *cb, uint8_t *cr);

/* brojTipova[5] of the last coded P picture (moestimation.cpp:422,529-551): #P_Skip,#16x16,#16x8,#8x16,#8x8. */
int fh264_mode_counts(fh264_session *s, int seq, int32_t counts[5]);

/* ---- stand-alone building blocks (unit tests, and the I-picture helper named by the north star) -------- */

/* Fused per-macroblock residual -> 4x4 transform -> quant -> zigzag -> dequant -> inverse -> clip for n inter
 * macroblocks given source and prediction (384 bytes each: 16x16 Y, 8x8 Cb, 8x8 Cr). Host pointers.
 * quantizationTransform(...,true) quantizationTransform.cpp:349-485 + inttransform.cpp:133-154,237-321. */
int fh264_tq_macroblocks(fh264_session *s, int n, const uint8_t *src384, const uint8_t *pred384, int qp,
                         int16_t *levels384, uint8_t *recon384);

/* Intra16x16 luma variant with the 4x4 Hadamard of the 16 DCs (quantizationTransform.cpp:105-152,227-260;
 * scaleTransform.cpp:154-189,344-376; inttransform.cpp:157-208). n macroblocks, 256 bytes each. */
int fh264_tq_luma_intra16(fh264_session *s, int n, const uint8_t *src256, const uint8_t *pred256, int qp,
                          int16_t *dc16, int16_t *ac16x15, uint8_t *recon256);

/* Motion compensation of a whole picture of sequence seq from per-MB quadrant MVs (mocomp.cpp:152-208):
 * qmv = nmb*4*2 int16 (host), pred384 = nmb*384 bytes (host). */
int fh264_motion_compensate(fh264_session *s, int seq, const int16_t *qmv, uint8_t *pred384);

/* Phase-R products of sequence seq's current dpb (host copies; tests): plane f (W*H bytes), feature k of plane f
 * (W*H uint16). */
int fh264_debug_plane(fh264_session *s, int seq, int f, uint8_t *out);
int fh264_debug_feature(fh264_session *s, int seq, int k, int f, uint16_t *out);

/* Device time of the last fh264_encode_p call, milliseconds, from CUDA events on the session stream:
 * [0] phase A (predictor-independent search), [1] phase B (wavefront), [2] phase C (MC + TQ + reconstruction),
 * [3] result copies + phase R (reference preparation for the next picture), [4] total; per kernel:
 * [5] k_stage3, [6] k_stage2, [7] k_interp, [8] k_features, [9] k_tile_index. */
int fh264_last_timings(fh264_session *s, float ms[10]);
/* Phase S of the last fh264_encode_p (k_spec: the search completed in parallel for the guessed integer predictors, so that the
 * wavefront only evaluates a handful of finalists per 8x8 partition), milliseconds; it is part of [0] above. */
int fh264_last_spec_ms(fh264_session *s, float *ms);
/* Measurement aid: sustained rate of the integer pipe the motion-search instructions issue on, measured on
 * `device` now (about 40 ms), in T lane-statements per second: [0] IMAD, [1] VIADDMNMX.S16x2, [2] VIADDMNMX,
 * [3] VABSDIFF4.U8.ACC + VIADD pairs. bench.py divides by it in the same run that it times the kernels. */
int fh264_measure_int_peak(int device, double tops[4]);

/* ---- device entropy coding of a P slice ---------------------------------------------------------
 * slice_data() of the P picture last coded by fh264_encode_p for sequences [seq0, seq0 + nseq): what the P-slice macroblock
 * loop of RBSP_encode writes between shd_write() and RBSP_trailing_bits() (rbsp_encoding.cpp:175-313): mb_skip_run, mb_type,
 * sub_mb_type, mvd_l0, coded_block_pattern (setCodedBlockPattern :21-105), mb_qp_delta and the CAVLC residual
 * (residual_write / residual_block_cavlc_write, residual.cpp:300-666), including the mb_skip_run that ends the slice (:310).
 * The caller has written the slice header with the reference's own shd_write(); `first_bit` (0..7) is the bit position inside
 * its last, partially filled byte, so the returned bytes can be OR-ed / appended without shifting: sequence b's bytes start at
 * out + b * out_stride, bits [0, first_bit) are zero, slice data occupies bits [first_bit, nbits[b]), the rest of the last
 * byte is zero (RBSP_trailing_bits stays with the caller). Errors: FH264_E_UNSUPPORTED if a level needs level_prefix > 15
 * (outside the reference's level table, residual_tables.cpp:940-1008); FH264_E_CAPACITY if the slice data
 * exceeds the reference's 500000-byte RBSP buffer (fer_h264.cpp:93). */
typedef struct fh264_cavlc_mb_info {   /* per macroblock, 32 bytes: what the reference's macroblock loop leaves in its host arrays */
    uint8_t skip;                      /* mb_type == P_Skip */
    uint8_t cbp_luma, cbp_chroma;      /* CodedBlockPatternLumaArray / ChromaArray (rbsp_encoding.cpp:103-104); 0 for P_Skip */
    uint8_t mb_type;                   /* mb_type_array (:180) */
    uint8_t total_coeff_luma[16];      /* totalcoeff_array_luma by luma4x4BlkIdx (residual.cpp:508); 0 where the block is not coded */
    uint8_t total_coeff_chroma[2][4];  /* totalcoeff_array_chroma */
    uint8_t reserved[4];
} fh264_cavlc_mb_info;
/* mb_info (nullable, nseq * MBs entries): side information a host that keeps coding I pictures with the reference's own code
 * needs to keep those arrays as the reference would (its intra bit-cost trials read them across pictures). */
/* Band mode: after fh264_band_gather(s, 1) on every rank, phase C of every rank also stores its macroblocks' records into rank 0's
 * memory (NVLink peer stores, one buffer per picture parity), so the slice — whose contexts cross the bands — is entropy-coded on
 * rank 0 once all ranks have delivered the picture; the other ranks (and rank 0 without the gather) get FH264_E_UNSUPPORTED. */
int fh264_cavlc_p(fh264_session *s, int seq0, int nseq, int first_bit, uint8_t *out, size_t out_stride, uint32_t *nbits,
                  fh264_cavlc_mb_info *mb_info);

/* ---- streaming step (no host round trip per picture) ------------------------------------------------------------------------
 * The reference decides P vs IDR on the host before every picture (selectNALUnitType, ref_frames.cpp:185-234) and then walks the
 * slice (RBSP_encode, rbsp_encoding.cpp:139-323); a caller that mirrors that with fh264_scene_sad + fh264_encode_p + fh264_cavlc_p
 * synchronises with the GPU three times per picture. fh264_encode_p_stream enqueues the whole step and returns:
 *   - scene_gate == 2: sum |frame.L - dpb.L| is measured on the device and reported in the status words, nothing else changes
 *     (the only form available in band mode, where every rank holds the whole picture and measures the same value).
 *   - scene_gate == 1: the sum is also compared with MBs << 12 (:210-224). A sequence above the
 *     threshold is NOT coded: its reference picture, records and entropy-coder state stay as they were, status word
 *     FH264_ST_GATE of its snapshot is 1, and the caller codes that picture with fh264_encode_i (the source picture is still
 *     current). The first-picture and IntraEvery rules (:191) need no pixels and stay with the caller.
 *   - records / slice data / side information / status go home on a copy stream, overlapping the dpb swap, phase R and the next
 *     step. All output pointers are optional and must be pinned host memory (fh264_host_alloc); they are valid after fh264_sync().
 *   - slice (device CAVLC, as fh264_cavlc_p): sequence b's bytes start at slice + b * slice_stride; only the first
 *     slice_copy_bytes of every sequence are copied home (bound chosen by the caller: a 1080p P slice at QP 28 is ~20 KB; the
 *     reference's own limit is 500000). slice_stat[2b] = error flags (1: a macroblock's private buffer overflowed, 2: level outside
 *     the reference's table, 4: slice above 500000 bytes), slice_stat[2b+1] = slice_data bits (first_bit included). A slice longer
 *     than slice_copy_bytes is complete on the device: fetch it with fh264_cavlc_p.
 *   - status: FH264_STATUS_WORDS uint32 per sequence, the snapshot after phase C: [0] flags, [2..6] mode counts (brojTipova order),
 *     [8],[9] scene SAD low / high word (when scene_gate), [FH264_ST_GATE] the gate, [FH264_ST_GATED_TOTAL] pictures stopped so far.
 * fh264_set_pipeline(s, 1) (or FH264_PIPE=1 in the environment at fh264_open) software-pipelines calls with two or more sequences:
 * the sequences are split into two halves on two streams and the latency-bound mode-decision wavefront of one half (warp-level
 * kernel, high-priority stream) runs under the search kernels of the other half, within a call and across consecutive calls.
 * Results are identical. Off by default: on the B200 the search kernels lose as much to the resident wavefront as the overlap
 * gains (profiles/r02_pipeline.md), and the per-kernel timings (fh264_last_timings) are only meaningful unpipelined. */
#define FH264_STATUS_WORDS 24
#define FH264_ST_SAD_LO 8
#define FH264_ST_SAD_HI 9
#define FH264_ST_GATE 17
#define FH264_ST_GATED_TOTAL 18
typedef struct fh264_stream_out {
    fh264_mb_result *records;          /* nseq * MBs, or NULL */
    uint8_t *slice;                    /* or NULL: no entropy coding */
    size_t slice_stride, slice_copy_bytes;
    int first_bit;                     /* 0..7, as fh264_cavlc_p */
    uint32_t *slice_stat;              /* 2 * nseq (required with slice) */
    fh264_cavlc_mb_info *mb_info;      /* nseq * MBs, or NULL */
    uint32_t *status;                  /* FH264_STATUS_WORDS * nseq, or NULL */
} fh264_stream_out;
int fh264_encode_p_stream(fh264_session *s, int seq0, int nseq, const fh264_params *p, int scene_gate, const fh264_stream_out *out);
int fh264_set_pipeline(fh264_session *s, int on);

/* `frame` := the source pictures of sequences [seq0, seq0 + nseq) from ONE host block (sequence b's Y, Cb, Cr planes contiguous at
 * block + b * stride): one call per step instead of one per sequence (ReadFromY4M per process, fileIO.cpp:286-337). Asynchronous,
 * double buffered like fh264_upload_source. device != 0: the block is device memory (device-to-device copies). */
int fh264_upload_source_batch(fh264_session *s, int seq0, int nseq, const void *block, size_t stride, int device);

/* ---- decoder inverse path ------------------------------------------------------------------------
 * Reconstructs the P picture described by `records` ([nseq][MBs]: mb_type, quadrant MVs and levels, i.e. what RBSP_decode holds
 * per macroblock after entropy decoding, rbsp_decoding.cpp:98-109,330-346) from the current reference picture with the same
 * motion-compensation and inverse-transform code the encoder reconstructs with (Decode, mocomp.cpp:200-208;
 * transformDecoding4x4LumaResidual / transformDecodingChroma / transformDecodingP_Skip, inttransform.cpp:133-321), then makes
 * it the reference picture like fh264_encode_p does. Read it back with fh264_download_recon. Synchronous. */
int fh264_decode_p(fh264_session *s, int seq0, int nseq, int qp, const fh264_mb_result *records);

/* ---- I pictures on the device ---------------------------------------------------------------------
 * One macroblock of an I picture = what intraPredictionEncoding() (intra.cpp:949-1109, CPU semantics) and the following
 * quantizationTransform(..., true) (rbsp_encoding.cpp:196-215) leave in the reference's globals. 832 bytes like fh264_mb_result. */
typedef struct fh264_mb_result_i {
    int16_t mb_type;                            /* as stored in mb_type_array[]: 0 = I_4x4, 1..24 = I_16x16_<pred>_<cbp chroma>_<cbp luma> */
    int8_t intra16x16_pred_mode;                /* return value of intraPredictionEncoding(): 0..3, -1 = Intra4x4 chosen */
    uint8_t intra_chroma_pred_mode;
    uint8_t cbp_luma, cbp_chroma;               /* CodedBlockPatternLuma / Chroma (setCodedBlockPattern, rbsp_encoding.cpp:21-105) */
    uint16_t bits_intra16x16, bits_intra4x4;    /* coded_mb_size() of the two trials (intra.cpp:1008,1088); Intra4x4 wins when smaller */
    uint8_t intra4x4_pred_mode[16];             /* Intra4x4PredMode[(CurrMbAddr << 4) + luma4x4BlkIdx] (the search result, kept either way) */
    uint8_t prev_intra4x4_pred_mode_flag[16];
    uint8_t rem_intra4x4_pred_mode[16];         /* meaningful where the flag is 0 */
    int16_t luma[16][16];                       /* Intra4x4: LumaLevel[blk][k]; Intra16x16: flat [0..15] = Intra16x16DCLevel, [16 + blk*15 + k] = Intra16x16ACLevel[blk][k] */
    int16_t chroma_dc[2][4];
    int16_t chroma_ac[2][4][15];
    int16_t reserved[3];
} fh264_mb_result_i;

/* Code the current source picture of sequences [seq0, seq0+nseq) as an I picture: for every macroblock the Intra16x16 and
 * Intra4x4 mode searches, the two CAVLC bit-cost trials, the decision, transform/quantisation and in-loop reconstruction
 * (intraPredictionEncoding + quantizationTransform, rbsp_encoding.cpp:196-215); then dpb := reconstruction and phase R, as after
 * fh264_encode_p. The bit-cost trial of a macroblock reads whether the SAME macroblock was P_Skip in the previous picture
 * (mb_type_array is only cleared after the first trial, intra.cpp:1008-1012); the session remembers that from its last
 * fh264_encode_p / fh264_decode_p. results: nseq * MBs records (host) or NULL. Synchronous. Band mode: the picture is not split —
 * every rank codes the whole picture (every rank must make the call; all ranks are waited for before and after it). */
int fh264_encode_i(fh264_session *s, int seq0, int nseq, int qp, fh264_mb_result_i *results);
/* slice_data() of the I picture last coded by fh264_encode_i for sequences [seq0, seq0 + nseq): what the I-slice macroblock loop of
 * RBSP_encode writes between shd_write() and RBSP_trailing_bits() (rbsp_encoding.cpp:221-305: mb_type, prev_intra4x4_pred_mode_flag /
 * rem_intra4x4_pred_mode, intra_chroma_pred_mode, coded_block_pattern, mb_qp_delta, residual_write residual.cpp:300-372).
 * Arguments, bit alignment and errors as fh264_cavlc_p. */
int fh264_cavlc_i(fh264_session *s, int seq0, int nseq, int first_bit, uint8_t *out, size_t out_stride, uint32_t *nbits);
/* Device time of the intra wavefront kernel of the last fh264_encode_i call, milliseconds (CUDA events on the session stream). */
int fh264_last_intra_ms(fh264_session *s, float *ms);

/* ---- band mode: one picture split into macroblock-row bands over the GPUs of a node (BASELINE config 4) --------------
 * One process and one session per GPU, every rank encodes the same pictures in the same order. Each rank keeps the whole
 * reference picture (upload_source / upload_recon take full pictures on every rank) and codes MB rows [mb_row0, mb_row1);
 * fh264_encode_p fills only those records of `results`. The phase-B wavefront crosses GPUs through progress flags mirrored
 * into the next rank's memory, the reconstructed bands are exchanged by peer stores inside phase C (NVLink, CUDA IPC), a
 * device-side barrier separates pictures. Setup: band_config on every rank, then exchange the export blobs (e.g. with
 * torch.distributed all_gather_object) and import every other rank's blob. */
#define FH264_IPC_HANDLE_BYTES 64
#define FH264_IPC_HANDLES 11
int fh264_band_config(fh264_session *s, int rank, int world, int mb_row0, int mb_row1);
int fh264_ipc_export(fh264_session *s, int seq, uint8_t *handles /* FH264_IPC_HANDLES * FH264_IPC_HANDLE_BYTES */);
/* Optional, after fh264_band_config on every rank: the bands of ALL ranks (mb_rows[2r], mb_rows[2r+1] = first / end MB row of rank r).
 * The picture barrier then only waits for the ranks whose bands lie within this rank's halo (304 luma rows: stage 2 reaches 279,
 * moestimation.cpp:481) and phase R only covers those rows, so a rank starts the next picture while the wavefront of the current one
 * is still running through the bands further down (pictures are pipelined across the GPUs). From then on the scene SAD of a rank
 * covers its own band (the ranks' sums add up to the picture's, ref_frames.cpp:210-224) and fh264_download_recon returns the whole
 * picture only once every rank has finished it: synchronise the ranks on the host first. FH264_E_UB_INPUT is raised by the ranks
 * whose halo contains the offending window: treat any rank's status as the picture's. */
int fh264_band_peers(fh264_session *s, int world, const int *mb_rows);
/* Band mode, optional: gather every rank's macroblock records on rank 0 (see fh264_cavlc_p). Same call on every rank. */
int fh264_band_gather(fh264_session *s, int on);
int fh264_ipc_import(fh264_session *s, int seq, int peer_rank, const uint8_t *handles);

/* Snapshot of the 16 status words of sequence seq after phase C of its last encode_p: [0] flags, [1] stage-2 pool
 * entries used, [2..6] mode counts, [12] partitions redone by the large-buffer stage-2 launch. */
int fh264_debug_status(fh264_session *s, int seq, uint32_t out[16]);
/* FH264_TRACE=1 (environment, read by fh264_open): out[3][8][5] = ms of phase A start / end, phase B end, phase C end, phase R end of the
 * last 8 pictures of the two pipeline lanes and the session stream, relative to the oldest of them; -1 where nothing was recorded. */
int fh264_debug_trace(fh264_session *s, float *out);

/* Debug: clock64() samples of the phase-B wavefront, 12 int64 per macroblock of sequence seq ([0] CTA start,
 * [1] prefetch issued, [2] dependencies satisfied, [3] neighbour MVs loaded, [4] P_Skip decided, [5..8] partitions
 * decided, [9] published). Call with out == NULL to enable sampling, with a buffer to read the last picture back. */
int fh264_debug_timeline(fh264_session *s, int seq, long long *out);

#ifdef __cplusplus
}
#endif
#endif /* FH264_B200_H */
