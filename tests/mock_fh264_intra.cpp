// TEST INFRASTRUCTURE — NOT a CPU fallback of the product and never linked into libfh264_b200.so or any shipped binary.
//
// A mock of the few C-ABI entry points (include/fh264_b200.h) that the reference-side binding integration/fh264_ref_shim.cpp uses
// for I pictures, implemented with the HOST build of the I-picture core (h264_fer_b200/csrc/intra_core.h, the source the device
// runs). Its only purpose: integration/Makefile links the all-on-device shim variant against it
// (integration/_build/shimtest_all_on_cpu_mock), so that the HOST logic of that binding — slice header, appending the device's
// slice data without shifting, trailing bits, frame / dpb bookkeeping — is checked byte for byte against the reference's
// bitstream on an all-IDR clip in the CPU suite (tests/test_intra_host.py), where no GPU exists. Every P-path entry point aborts.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "h264_fer_b200/csrc/intra_core.h"

struct fh264_session {
    int W, H, wmb, nmb;
    std::vector<unsigned char> src[3], rec[3];
    std::vector<fh264_mb_result_i> recs;
    bool have_i;
};

static void p_path(const char *what) { fprintf(stderr, "mock_fh264_intra: %s is not available in the mock (I pictures only)\n", what); abort(); }

extern "C" {
const char *fh264_last_error(void) { return "mock"; }
int fh264_open(int width, int height, int batch, int device, fh264_session **out)
{
    (void)device;
    if (batch != 1 || (width & 15) || (height & 15)) return FH264_E_ARG;
    fh264_session *s = new fh264_session();
    s->W = width; s->H = height; s->wmb = width / 16; s->nmb = s->wmb * (height / 16); s->have_i = false;
    for (int c = 0; c < 3; c++) { s->src[c].assign((size_t)width * height / (c ? 4 : 1), 0); s->rec[c] = s->src[c]; }
    s->recs.resize(s->nmb);
    *out = s;
    return FH264_OK;
}
int fh264_close(fh264_session *s) { delete s; return FH264_OK; }
int fh264_upload_source(fh264_session *s, int seq, const uint8_t *y, const uint8_t *cb, const uint8_t *cr)
{
    if (seq != 0) return FH264_E_ARG;
    memcpy(s->src[0].data(), y, s->src[0].size()); memcpy(s->src[1].data(), cb, s->src[1].size()); memcpy(s->src[2].data(), cr, s->src[2].size());
    return FH264_OK;
}
int fh264_encode_i(fh264_session *s, int seq0, int nseq, int qp, fh264_mb_result_i *results)
{
    if (seq0 != 0 || nseq != 1) return FH264_E_ARG;
    std::vector<IcInfo> info(s->nmb);
    for (int m = 0; m < s->nmb; m++) {
        IcCtx c;
        for (int k = 0; k < 3; k++) { c.src[k] = s->src[k].data(); c.rec[k] = s->rec[k].data(); }
        c.W = s->W; c.H = s->H; c.xP = (m % s->wmb) * 16; c.yP = (m / s->wmb) * 16; c.qp = qp;
        ic_macroblock(c, false, (m % s->wmb) ? &info[m - 1] : nullptr, m >= s->wmb ? &info[m - s->wmb] : nullptr, s->recs[m], info[m], 0, 1);
    }
    if (results) memcpy(results, s->recs.data(), sizeof(fh264_mb_result_i) * s->nmb);
    s->have_i = true;
    return FH264_OK;
}
int fh264_cavlc_i(fh264_session *s, int seq0, int nseq, int first_bit, uint8_t *out, size_t out_stride, uint32_t *nbits)
{
    if (seq0 != 0 || nseq != 1 || !s->have_i) return FH264_E_STATE;
    std::vector<IcInfo> info(s->nmb);
    for (int m = 0; m < s->nmb; m++) ic_info_from_record(s->recs[m], info[m]);
    std::vector<uint32_t> words(out_stride / 4 + 2, 0u);
    CvBits b;
    cv_init(b, words.data(), (int)words.size());
    cv_put(b, first_bit, 0);
    int bad = 0;
    for (int m = 0; m < s->nmb; m++)
        ic_write_macroblock(b, s->recs[m], info[m], (m % s->wmb) ? &info[m - 1] : nullptr, m >= s->wmb ? &info[m - s->wmb] : nullptr, &bad);
    *nbits = (uint32_t)cv_bits(b);
    cv_flush(b);
    if (b.ovf || bad) return FH264_E_CAPACITY;
    for (size_t i = 0; i < ((size_t)*nbits + 7) / 8; i++) out[i] = (uint8_t)(words[i >> 2] >> (24 - 8 * (i & 3)));
    return FH264_OK;
}
int fh264_download_recon(fh264_session *s, int seq, uint8_t *y, uint8_t *cb, uint8_t *cr)
{
    if (seq != 0) return FH264_E_ARG;
    memcpy(y, s->rec[0].data(), s->rec[0].size()); memcpy(cb, s->rec[1].data(), s->rec[1].size()); memcpy(cr, s->rec[2].data(), s->rec[2].size());
    return FH264_OK;
}
// ---- P path: not in the mock
int fh264_upload_recon(fh264_session *, int, const uint8_t *, const uint8_t *, const uint8_t *) { p_path("fh264_upload_recon"); return FH264_E_UNSUPPORTED; }
int fh264_scene_sad(fh264_session *, int, uint64_t *) { p_path("fh264_scene_sad"); return FH264_E_UNSUPPORTED; }
int fh264_encode_p(fh264_session *, int, int, const fh264_params *, fh264_mb_result *) { p_path("fh264_encode_p"); return FH264_E_UNSUPPORTED; }
int fh264_mode_counts(fh264_session *, int, int32_t *) { p_path("fh264_mode_counts"); return FH264_E_UNSUPPORTED; }
int fh264_cavlc_p(fh264_session *, int, int, int, uint8_t *, size_t, uint32_t *, fh264_cavlc_mb_info *) { p_path("fh264_cavlc_p"); return FH264_E_UNSUPPORTED; }
}
