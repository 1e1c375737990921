// TEST INFRASTRUCTURE — host build of the device I-picture core (h264_fer_b200/csrc/intra_core.h), so that it is checked on CPU
// against I pictures of the compiled reference (oracle/_ref/ref_encoder, chunk IMBR) and the golden fixtures made from it.
// Built by tests/test_intra_host.py with g++; never part of the product path (the product runs the same core in intra.cuh).
#include <cstring>
#include <vector>
#include "../h264_fer_b200/csrc/intra_core.h"

static_assert(sizeof(fh264_mb_result_i) == 832, "ABI record size");

// prev_types: mb_type_array of the previous picture (int32 per MB) or null. src / rec: planar 4:2:0, rec receives the reconstruction.
extern "C" int intra_host_picture(const unsigned char *sy, const unsigned char *su, const unsigned char *sv, unsigned char *ry, unsigned char *ru,
                                  unsigned char *rv, int W, int H, int qp, const int *prev_types, fh264_mb_result_i *out)
{
    const int wmb = W / 16, nmb = wmb * (H / 16);
    std::vector<IcInfo> info(nmb);
    for (int m = 0; m < nmb; m++) {
        IcCtx c;
        c.src[0] = sy; c.src[1] = su; c.src[2] = sv; c.rec[0] = ry; c.rec[1] = ru; c.rec[2] = rv;
        c.W = W; c.H = H; c.xP = (m % wmb) * 16; c.yP = (m / wmb) * 16; c.qp = qp;
        ic_macroblock(c, prev_types && prev_types[m] == 31, (m % wmb) ? &info[m - 1] : nullptr, m >= wmb ? &info[m - wmb] : nullptr, out[m], info[m], 0, 1);
    }
    return 0;
}

// slice_data() of an I picture from its records: bits [first_bit, *nbits) of out (MSB first), like cavlc_host_slice for P pictures
extern "C" int intra_host_slice(const fh264_mb_result_i *rec, int nmb, int wmb, int first_bit, unsigned char *out, int cap_bytes, int *nbits, int *bad)
{
    std::vector<IcInfo> info(nmb);
    for (int m = 0; m < nmb; m++) ic_info_from_record(rec[m], info[m]);
    std::vector<uint32_t> words(cap_bytes / 4 + 2, 0u);
    CvBits b;
    cv_init(b, words.data(), (int)words.size());
    cv_put(b, first_bit, 0);
    *bad = 0;
    for (int m = 0; m < nmb; m++)
        ic_write_macroblock(b, rec[m], info[m], (m % wmb) ? &info[m - 1] : nullptr, m >= wmb ? &info[m - wmb] : nullptr, bad);
    *nbits = cv_bits(b);
    cv_flush(b);
    if (b.ovf) return -1;
    const int nbytes = (*nbits + 7) / 8;
    for (int i = 0; i < nbytes; i++) out[i] = (unsigned char)(words[i >> 2] >> (24 - 8 * (i & 3)));
    return 0;
}
