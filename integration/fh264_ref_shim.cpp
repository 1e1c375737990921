// Reference-side binding of fh264_b200 (the stub a maintainer of zoltanmaric/h264-fer would add).
//
// It re-implements, with the reference's own names and signatures, exactly the entry points the picture coder
// RBSP_encode() calls on the P path (rbsp_encoding.cpp:175-192,317-322) plus selectNALUnitType() (fer_h264.cpp:65,121):
//
//   interEncoding(predL, predCr, predCb)            moestimation.h:4     -> per-MB copy-out of the GPU's results
//   quantizationTransform(predL, predCb, predCr, r) quantizationTransform.h:10 -> levels into LumaLevel/ChromaDCLevel/ChromaACLevel
//   transformDecodingP_Skip(...)                    inttransform.h:5     -> no-op (the GPU reconstructed the picture)
//   FillInterpolatedRefFrame()                      moestimation.h:8     -> keeps host `frame`/`dpb` and the device in step
//   selectNALUnitType()                             ref_frames.h:20      -> same rule, scene SAD from the device
//   InitCL / AllocateFrameBuffersCL / CloseCL        openCL_functions.h:4-19 -> the accelerator lifecycle seam the reference already
//                                                    has (fer_h264.cpp:90,74,129; rbsp_encoding.cpp:129): session open / close
//
// following the pattern the reference itself uses for its OpenCL offload: launch whole-picture work when
// CurrMbAddr == 0 and consume it per macroblock (IntraCL() at rbsp_encoding.cpp:144, WaitIntraCL at intra.cpp:963-966).
// The unmodified reference sources are compiled with -Dname=ref_name for these five symbols (integration/Makefile), so
// I pictures and intra bit-cost trials still run the reference's own code; everything else (CAVLC, NAL, headers,
// intra prediction, Y4M input) is the untouched reference host code. Two further variants move more onto the device:
// -DFH264_SHIM_DEVICE_CAVLC (P-slice entropy coding) and -DFH264_SHIM_DEVICE_INTRA (I pictures: intraPredictionEncoding()).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "h264_globals.h"
#include "headers_and_parameter_sets.h"
#include "residual.h"
#include "mode_pred.h"
#include "ref_frames.h"
#include "moestimation.h"
#include "../include/fh264_b200.h"
#ifdef FH264_SHIM_DEVICE_CAVLC
#include "nal.h"
#include "rbsp_IO.h"
#include "rbsp_encoding.h"
#endif

void ref_interEncoding(int predL[16][16], int predCr[8][8], int predCb[8][8]);
void ref_FillInterpolatedRefFrame();
void ref_quantizationTransform(int predL[16][16], int predCb[8][8], int predCr[8][8], bool reconstruct);
void ref_transformDecodingP_Skip(int predL[16][16], int predCb[8][8], int predCr[8][8], int QPy);
int ref_selectNALUnitType();
extern frame_type dpb;

static fh264_session *g_sess = 0;
static std::vector<fh264_mb_result> g_res;
static bool g_last_was_p = false;
static bool g_last_was_device_i = false;     // FH264_SHIM_DEVICE_INTRA: the I picture was coded (and made the reference) on the device

static void die(const char *what, int rc)
{
    fprintf(stderr, "fh264 shim: %s failed (%d): %s\n", what, rc, fh264_last_error());
    exit(3);                       // fail loudly: there is no CPU fallback behind this shim
}

static void open_session()
{
    if (g_sess) return;
    const char *dev = getenv("FH264_DEVICE");
    int rc = fh264_open(frame.Lwidth, frame.Lheight, 1, dev ? atoi(dev) : 0, &g_sess);
    if (rc) die("fh264_open", rc);
    g_res.resize((size_t)(frame.Lwidth >> 4) * (frame.Lheight >> 4));
}
// (a host that never calls AllocateFrameBuffersCL — none in the reference — still gets a session on first use)
static void ensure_session() { open_session(); }

// ---- the reference's accelerator lifecycle seam (openCL_functions.h:4-19), taken over as it stands -------------------------
// OpenCLEnabled stays false: the reference's own OpenCL intra path yields different bitstreams (intra.cpp:961-977 vs :978-1049),
// and the symbols its host code references must exist. InitCL() runs before the picture size is known (fer_h264.cpp:90), so the
// session opens in AllocateFrameBuffersCL(), which RBSP_encode() calls right after the SPS fixed the size (rbsp_encoding.cpp:129)
// — the same place the reference allocates its device frame buffers — and closes in CloseCL() (fer_h264.cpp:74,129).
#include <CL/cl.h>
bool OpenCLEnabled = false;
int *predModes16x16 = 0, *predModes4x4 = 0;
cl_mem frame_mem, dpb_mem, ans_mem;
cl_command_queue cmd_queue;
cl_context context;
cl_kernel kernel[2];
void InitCL() {}
void AllocateFrameBuffersCL() { open_session(); }
void CloseCL()
{
    if (!g_sess) return;
    int rc = fh264_close(g_sess);
    g_sess = 0;
    if (rc) die("fh264_close", rc);
}
void IntraCL() {}                               // (only reached with OpenCLEnabled)
void WaitIntraCL(int) {}
void subtractFramesCL(unsigned char *, unsigned char *)
{
    // the reference's selectNALUnitType() would take this path with OpenCLEnabled (ref_frames.cpp:196-208); the shim replaces
    // selectNALUnitType() itself (below: fh264_scene_sad returns the 64-bit sum instead of W*H differences), so nobody calls it
    fprintf(stderr, "fh264 shim: subtractFramesCL is not part of the binding (selectNALUnitType uses fh264_scene_sad)\n");
    exit(3);
}

int selectNALUnitType()
{
    // ref_frames.cpp:185-234 with the luma |frame - dpb| sum taken on the device
    if (dpb.L == NULL || currFrameCount % IntraEvery == 0) return NAL_UNIT_TYPE_IDR;
    ensure_session();
    int rc = fh264_upload_source(g_sess, 0, frame.L, frame.C[0], frame.C[1]);
    if (rc) die("fh264_upload_source", rc);
    uint64_t sad = 0;
    rc = fh264_scene_sad(g_sess, 0, &sad);
    if (rc) die("fh264_scene_sad", rc);
    const unsigned long picSizeInMBs = (unsigned long)PicWidthInMbs * PicHeightInMbs;
    return sad > (picSizeInMBs << 12) ? NAL_UNIT_TYPE_IDR : NAL_UNIT_TYPE_NOT_IDR;
}

void interEncoding(int predL[16][16], int predCr[8][8], int predCb[8][8])
{
    (void)predL; (void)predCr; (void)predCb;
    if (CurrMbAddr == 0) {
        // whole-picture device pipeline; `frame` was uploaded by selectNALUnitType() for this picture
        ensure_session();
        fh264_params p;
        p.qp = QPy; p.window = WindowSize; p.maxdiff_set = MAXDIFF_SET; p.basic = BasicInterEncoding ? 1 : 0;
        int rc = fh264_encode_p(g_sess, 0, 1, &p, g_res.data());
        if (rc) die("fh264_encode_p", rc);
        int32_t c[5];
        if (fh264_mode_counts(g_sess, 0, c) == FH264_OK) for (int i = 0; i < 5; i++) brojTipova[i] = c[i];
        g_last_was_p = true;
    }
    const fh264_mb_result &r = g_res[CurrMbAddr];
    mb_type = r.mb_type;
    mb_type_array[CurrMbAddr] = r.mb_type;
    ClearMVD();
    for (int i = 0; i < r.num_parts; i++) { mvd_l0[i][0][0] = r.mvd[i][0]; mvd_l0[i][0][1] = r.mvd[i][1]; }
    for (int q = 0; q < 4; q++)
        for (int j = 0; j < 4; j++) { mvL0x[CurrMbAddr][q][j] = r.mv[q][0]; mvL0y[CurrMbAddr][q][j] = r.mv[q][1]; }
    refIdxL0[CurrMbAddr] = 0;
}

#ifdef FH264_SHIM_DEVICE_INTRA
// ---- variant with the I pictures on the device too (integration/_build/fh264_encoder_b200_intra) ---------
// intra.cpp is compiled with -DintraPredictionEncoding=ref_intraPredictionEncoding. The I-slice macroblock loop of RBSP_encode
// (rbsp_encoding.cpp:196-215) stays the reference's: it calls intraPredictionEncoding() and quantizationTransform(..., true),
// derives mb_type / CodedBlockPattern from the levels it finds in the globals and writes the macroblock with its own CAVLC.
// Here intraPredictionEncoding() runs the whole picture on the device when CurrMbAddr == 0 (the pattern of IntraCL() /
// WaitIntraCL(), rbsp_encoding.cpp:144, intra.cpp:963-966) and then hands out, per macroblock, what the reference function
// leaves in the globals; quantizationTransform() copies the levels and the reconstructed samples.
static std::vector<fh264_mb_result_i> g_ires;
static std::vector<unsigned char> g_irec[3];

int intraPredictionEncoding(int predL[16][16], int predCr[8][8], int predCb[8][8])
{
    (void)predL; (void)predCr; (void)predCb;      // the prediction samples are only inputs of quantizationTransform(), replaced below
    if (CurrMbAddr == 0) {
        ensure_session();
        g_ires.resize(g_res.size());
        const size_t WH = (size_t)frame.Lwidth * frame.Lheight;
        g_irec[0].resize(WH); g_irec[1].resize(WH / 4); g_irec[2].resize(WH / 4);
        int rc = fh264_upload_source(g_sess, 0, frame.L, frame.C[0], frame.C[1]);
        if (rc) die("fh264_upload_source", rc);
        rc = fh264_encode_i(g_sess, 0, 1, QPy, g_ires.data());
        if (rc) die("fh264_encode_i", rc);
        rc = fh264_download_recon(g_sess, 0, g_irec[0].data(), g_irec[1].data(), g_irec[2].data());
        if (rc) die("fh264_download_recon", rc);
        g_last_was_device_i = true;
    }
    const fh264_mb_result_i &r = g_ires[CurrMbAddr];
    for (int b = 0; b < 16; b++) {
        Intra4x4PredMode[(CurrMbAddr << 4) + b] = r.intra4x4_pred_mode[b];
        prev_intra4x4_pred_mode_flag[b] = r.prev_intra4x4_pred_mode_flag[b] != 0;
        rem_intra4x4_pred_mode[b] = r.rem_intra4x4_pred_mode[b];
    }
    intra_chroma_pred_mode = r.intra_chroma_pred_mode;
    mb_type_array[CurrMbAddr] = 0;                // intra.cpp:1012,1058; the macroblock loop stores the final type afterwards
    return r.intra16x16_pred_mode;
}

static void intra_levels_and_reconstruction()
{
    const fh264_mb_result_i &r = g_ires[CurrMbAddr];
    const int16_t *l = &r.luma[0][0];
    if (r.intra16x16_pred_mode < 0) for (int b = 0; b < 16; b++) for (int k = 0; k < 16; k++) LumaLevel[b][k] = l[b * 16 + k];
    else {
        for (int k = 0; k < 16; k++) Intra16x16DCLevel[k] = l[k];
        for (int b = 0; b < 16; b++) for (int k = 0; k < 15; k++) Intra16x16ACLevel[b][k] = l[16 + b * 15 + k];
    }
    for (int c = 0; c < 2; c++) for (int k = 0; k < 4; k++) ChromaDCLevel[c][k] = r.chroma_dc[c][k];
    for (int c = 0; c < 2; c++) for (int b = 0; b < 4; b++) for (int k = 0; k < 15; k++) ChromaACLevel[c][b][k] = r.chroma_ac[c][b][k];
    const int W = frame.Lwidth, CW = frame.Cwidth, xP = (CurrMbAddr % PicWidthInMbs) << 4, yP = (CurrMbAddr / PicWidthInMbs) << 4;
    for (int y = 0; y < 16; y++) memcpy(frame.L + (size_t)(yP + y) * W + xP, g_irec[0].data() + (size_t)(yP + y) * W + xP, 16);
    for (int c = 0; c < 2; c++)
        for (int y = 0; y < 8; y++) memcpy(frame.C[c] + (size_t)(yP / 2 + y) * CW + xP / 2, g_irec[1 + c].data() + (size_t)(yP / 2 + y) * CW + xP / 2, 8);
}
#endif

void quantizationTransform(int predL[16][16], int predCb[8][8], int predCr[8][8], bool reconstruct)
{
#ifdef FH264_SHIM_DEVICE_INTRA
    if ((shd.slice_type % 5) != P_SLICE && reconstruct) { intra_levels_and_reconstruction(); return; }
#endif
    if ((shd.slice_type % 5) != P_SLICE) { ref_quantizationTransform(predL, predCb, predCr, reconstruct); return; }
    const fh264_mb_result &r = g_res[CurrMbAddr];
    for (int b = 0; b < 16; b++) for (int k = 0; k < 16; k++) LumaLevel[b][k] = r.luma[b][k];
    for (int c = 0; c < 2; c++) for (int k = 0; k < 4; k++) ChromaDCLevel[c][k] = r.chroma_dc[c][k];
    for (int c = 0; c < 2; c++) for (int b = 0; b < 4; b++) for (int k = 0; k < 15; k++) ChromaACLevel[c][b][k] = r.chroma_ac[c][b][k];
}

void transformDecodingP_Skip(int predL[16][16], int predCb[8][8], int predCr[8][8], int qpy)
{
    (void)predL; (void)predCb; (void)predCr; (void)qpy;   // the device already reconstructed the picture
}

void FillInterpolatedRefFrame()
{
    // Called once per picture after modificationProcess() copied `frame` into `dpb` (rbsp_encoding.cpp:317-322).
    ensure_session();
    if (g_last_was_p) {
        // P picture: the reconstruction lives on the device (dpb swap + phase R already done inside fh264_encode_p);
        // bring it back so the host's `frame` / `dpb` stay what the reference would hold.
        int rc = fh264_download_recon(g_sess, 0, frame.L, frame.C[0], frame.C[1]);
        if (rc) die("fh264_download_recon", rc);
        memcpy(dpb.L, frame.L, (size_t)frame.Lwidth * frame.Lheight);
        memcpy(dpb.C[0], frame.C[0], (size_t)frame.Cwidth * frame.Cheight);
        memcpy(dpb.C[1], frame.C[1], (size_t)frame.Cwidth * frame.Cheight);
    } else if (!g_last_was_device_i) {
        // I picture coded by the reference host path: its reconstruction becomes the device's reference picture
        int rc = fh264_upload_recon(g_sess, 0, frame.L, frame.C[0], frame.C[1]);
        if (rc) die("fh264_upload_recon", rc);
    }   // else: I picture coded by fh264_encode_i, which already made its reconstruction the reference picture (host `frame` holds it too)
    g_last_was_p = false;
    g_last_was_device_i = false;
}

#ifdef FH264_SHIM_DEVICE_CAVLC
// ---- variant with the P-slice entropy coding on the device (integration/_build/fh264_encoder_b200_cavlc) --
// rbsp_encoding.cpp is compiled with -DRBSP_encode=ref_RBSP_encode: SPS, PPS and IDR slices still run the reference's own
// RBSP_encode; for a P slice this restates its frame (rbsp_encoding.cpp:119-127,160-170,308-325) around ONE device call
// chain: slice header by the reference's shd_write(), slice_data() from fh264_cavlc_p(), RBSP_trailing_bits() by the
// reference. No per-macroblock record crosses PCIe any more (fh264_encode_p with results == NULL).
void ref_RBSP_encode(NALunit &nal_unit);
void RBSP_trailing_bits();

#ifdef FH264_SHIM_DEVICE_INTRA
// ---- both variants together (integration/_build/fh264_encoder_b200_all): the IDR slice is coded AND entropy-coded on the device.
// Restates the frame of RBSP_encode around an IDR slice (rbsp_encoding.cpp:119-123,141-170,308-325) with ONE device call chain in
// place of the macroblock loop: slice header by the reference's shd_write(), slice_data() from fh264_encode_i + fh264_cavlc_i,
// RBSP_trailing_bits() by the reference. Neither the 832-byte records nor the prediction modes cross PCIe.
static void encode_idr_on_device(NALunit &nal_unit)
{
    initRawWriter(nal_unit.rbsp_byte, 500000);
    static bool firstFrame = true;                       // rbsp_encoding.cpp:146-162
    shd.slice_type = I_SLICE;
    if (firstFrame) { firstFrame = false; shd.idr_pic_id = 0; }
    else if (shd.frame_num == 0) shd.idr_pic_id++;
    else shd.idr_pic_id = 0;
    shd.frame_num = 0;
    shd_write(nal_unit);
    flushWriteBuffer();

    ensure_session();
    int rc = fh264_upload_source(g_sess, 0, frame.L, frame.C[0], frame.C[1]);
    if (rc) die("fh264_upload_source", rc);
    rc = fh264_encode_i(g_sess, 0, 1, QPy, NULL);
    if (rc) die("fh264_encode_i", rc);
    static std::vector<unsigned char> sl(500064);
    uint32_t nbits = 0;
    const int first_bit = (int)RBSP_write_current_bit;
    rc = fh264_cavlc_i(g_sess, 0, 1, first_bit, sl.data(), sl.size(), &nbits);
    if (rc) die("fh264_cavlc_i", rc);
    unsigned char *dst = RBSP_write_data + RBSP_write_current_byte;
    const size_t nbytes = ((size_t)nbits + 7) / 8;
    if (nbytes) {
        if (first_bit) dst[0] |= sl[0]; else dst[0] = sl[0];
        memcpy(dst + 1, sl.data() + 1, nbytes - 1);
    }
    RBSP_write_current_byte += nbits >> 3;
    RBSP_write_current_bit = nbits & 7;
    RBSP_trailing_bits();
    nal_unit.NumBytesInRBSP = RBSP_write_current_byte;
    // host `frame` := reconstruction, as it is after the reference's own macroblock loop (modificationProcess copies it into dpb)
    rc = fh264_download_recon(g_sess, 0, frame.L, frame.C[0], frame.C[1]);
    if (rc) die("fh264_download_recon", rc);
    g_last_was_device_i = true;
    initialisationProcess();                             // rbsp_encoding.cpp:314-317
    modificationProcess();
    FillInterpolatedRefFrame();
    flushWriteBuffer();
}
#endif

void RBSP_encode(NALunit &nal_unit)
{
#ifdef FH264_SHIM_DEVICE_INTRA
    if (nal_unit.nal_unit_type == NAL_UNIT_TYPE_IDR) { encode_idr_on_device(nal_unit); return; }
#endif
    if (nal_unit.nal_unit_type != NAL_UNIT_TYPE_NOT_IDR) { ref_RBSP_encode(nal_unit); return; }
    initRawWriter(nal_unit.rbsp_byte, 500000);
    shd.slice_type = P_SLICE;
    shd.frame_num++;
    shd_write(nal_unit);
    flushWriteBuffer();                                  // header bits out of the 64-bit staging word: byte / bit position is exact now

    ensure_session();
    fh264_params p;
    p.qp = QPy; p.window = WindowSize; p.maxdiff_set = MAXDIFF_SET; p.basic = BasicInterEncoding ? 1 : 0;
    int rc = fh264_encode_p(g_sess, 0, 1, &p, NULL);     // `frame` was uploaded by selectNALUnitType() for this picture
    if (rc) die("fh264_encode_p", rc);
    int32_t c[5];
    if (fh264_mode_counts(g_sess, 0, c) == FH264_OK) for (int i = 0; i < 5; i++) brojTipova[i] = c[i];
    g_last_was_p = true;

    static std::vector<unsigned char> sl(500064);
    uint32_t nbits = 0;
    const int first_bit = (int)RBSP_write_current_bit;
    static std::vector<fh264_cavlc_mb_info> info;
    info.resize(g_res.size());
    rc = fh264_cavlc_p(g_sess, 0, 1, first_bit, sl.data(), sl.size(), &nbits, info.data());
    if (rc) die("fh264_cavlc_p", rc);
    // The reference's intra bit-cost trials (coded_mb_size -> residual_block_cavlc_size) read mb_type_array, the
    // CodedBlockPattern arrays and totalcoeff_array_* of the PREVIOUS picture for blocks not yet coded in the current one, so
    // the next I picture only comes out identical if the P loop's leftovers are: rbsp_encoding.cpp:180,103-104, residual.cpp:508-517.
    for (size_t m = 0; m < info.size(); m++) {
        const fh264_cavlc_mb_info &f = info[m];
        mb_type_array[m] = f.mb_type;
        if (f.skip) continue;
        CodedBlockPatternLumaArray[m] = f.cbp_luma;
        CodedBlockPatternChromaArray[m] = f.cbp_chroma;
        for (int b = 0; b < 16; b++) if (f.cbp_luma & (1 << (b >> 2))) totalcoeff_array_luma[m][b] = f.total_coeff_luma[b];
        if (f.cbp_chroma & 2) for (int c = 0; c < 2; c++) for (int b = 0; b < 4; b++) totalcoeff_array_chroma[c][m][b] = f.total_coeff_chroma[c][b];
    }
    unsigned char *dst = RBSP_write_data + RBSP_write_current_byte;
    const size_t nbytes = ((size_t)nbits + 7) / 8;
    if (nbytes) {
        if (first_bit) dst[0] |= sl[0]; else dst[0] = sl[0];
        memcpy(dst + 1, sl.data() + 1, nbytes - 1);
    }
    RBSP_write_current_byte += nbits >> 3;
    RBSP_write_current_bit = nbits & 7;
    if (RBSP_write_current_bit == 0 && nbytes) { /* byte aligned: the next byte is cleared by the writer itself */ }
    RBSP_trailing_bits();
    nal_unit.NumBytesInRBSP = RBSP_write_current_byte;
    modificationProcess();
    FillInterpolatedRefFrame();
    flushWriteBuffer();
}
#endif
