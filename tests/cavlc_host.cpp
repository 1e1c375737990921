// TEST INFRASTRUCTURE — host build of the device CAVLC core (h264_fer_b200/csrc/cavlc_core.h), so that the coder and its
// tables are checked against the reference's slice data (tests/golden/*.npz rbsp_*) and tables (cavlc_tables.npz) on CPU.
// Built by tests/test_cavlc_host.py with g++; never part of the product path (the product runs the same core in cavlc.cuh).
#include <cstring>
#include <vector>
#include "../include/fh264_b200.h"
#include "../h264_fer_b200/csrc/cavlc_core.h"

extern "C" int cavlc_host_slice(const fh264_mb_result *rec, int nmb, int wmb, int first_bit, unsigned char *out, int cap_bytes, int *nbits, int *bad)
{
    std::vector<CvInfo> info(nmb);
    for (int m = 0; m < nmb; m++) cv_prepare(rec[m].mb_type, rec[m].luma, rec[m].chroma_dc, rec[m].chroma_ac, FH264_P_SKIP, info[m]);
    std::vector<uint32_t> words(cap_bytes / 4 + 2, 0u);
    CvBits b;
    cv_init(b, words.data(), (int)words.size());
    cv_put(b, first_bit, 0);
    int run = 0;
    *bad = 0;
    for (int m = 0; m < nmb; m++) {
        if (info[m].skip) { run++; continue; }
        const CvInfo *left = (m % wmb) ? &info[m - 1] : nullptr, *up = m >= wmb ? &info[m - wmb] : nullptr;
        cv_macroblock(b, run, rec[m].mb_type, rec[m].num_parts, rec[m].mvd, rec[m].luma, rec[m].chroma_dc, rec[m].chroma_ac, info[m], left, up, bad);
        run = 0;
    }
    if (run > 0) cv_ue(b, (uint32_t)run);                       // rbsp_encoding.cpp:310-313
    *nbits = cv_bits(b);
    cv_flush(b);
    if (b.ovf) return -1;
    const int nbytes = (*nbits + 7) / 8;
    for (int i = 0; i < nbytes; i++) out[i] = (unsigned char)(words[i >> 2] >> (24 - 8 * (i & 3)));
    return 0;
}

// the tables in the order of the reference dump (oracle/ref_harness/driver.cpp, chunk CVTB): (length, code) pairs
extern "C" int cavlc_host_tables(int *t)
{
    int n = 0;
    for (int tab = 0; tab < 5; tab++) for (int i = 0; i < 17; i++) for (int j = 0; j < 4; j++) { t[n++] = cv_ct_len[tab][i][j]; t[n++] = cv_ct_code[tab][i][j]; }
    for (int i = 0; i < 15; i++) for (int j = 0; j < 16; j++) { t[n++] = cv_tz_len[i][j]; t[n++] = cv_tz_code[i][j]; }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) { t[n++] = cv_tzc_len[i][j]; t[n++] = cv_tzc_code[i][j]; }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 7; j++) { t[n++] = cv_rb_len[i][j]; t[n++] = cv_rb_code[i][j]; }
    for (int i = 0; i < 48; i++) t[n++] = cv_cbp_inter[i];
    return n;
}
