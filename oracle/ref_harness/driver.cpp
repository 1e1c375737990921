// ORACLE / TEST INFRASTRUCTURE — not part of the product path.
//
// Command-line harness around the UNMODIFIED reference encoder, compiled in place from /root/reference
// by oracle/Makefile (outputs only under oracle/_ref/). It replaces the C++/CLI facade
// fer_h264.cpp:55-134 (encode()/NastaviEncode()) and Starter::PostaviParametre (fer_h264.cpp:169-178),
// which cannot be built outside MSVC /clr, with a plain loop issuing the same calls in the same order.
//
// The hot-path entry points are tapped without editing reference files: moestimation.cpp,
// quantizationTransform.cpp and inttransform.cpp are compiled with -Dname=ref_name, and the same-named
// wrappers below forward to ref_name, time the call and dump the globals the reference communicates
// through (h264_globals.h:99-193, residual.h:6-15, mode_pred.h:19-22).
//
// usage: ref_encoder in.y4m out.264 dump.bin|- frames qp basic window maxdiff intraEvery [dumpmask [planes_pic]]
//   dumpmask bits: 1 per-MB records (P pictures)   2 reconstruction per picture   4 cropped source per picture
//                  8 phase-R data after picture `planes_pic`   16 per-MB TQ input (snapped source + prediction)
//                  32 Intra16x16 luma records of I pictures (source, prediction, DC/AC levels, reconstruction)
//                  64 slice RBSP of P pictures, with 256 also of I pictures (SLDT: bit position of the first slice_data bit, then the RBSP bytes)
//                  128 the CAVLC coder tables once (CVTB; fixture for the table check of the device coder)
//                  256 per-MB records of I pictures (IMBR: final mb_type, prediction modes, both bit-cost trials, CBP, levels)
// stdout: one JSON line with per-picture types/bytes and timings.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "nal.h"
#include "fileIO.h"
#include "rbsp_IO.h"
#include "h264_globals.h"
#include "headers_and_parameter_sets.h"
#include "residual_tables.h"
#include "residual.h"
#include "ref_frames.h"
#include "expgolomb.h"
#include "rbsp_encoding.h"
#include "openCL_functions.h"
#include "mode_pred.h"
#include "moestimation.h"

// renamed reference entry points (see Makefile -D flags)
void ref_interEncoding(int predL[16][16], int predCr[8][8], int predCb[8][8]);
void ref_FillInterpolatedRefFrame();
void ref_quantizationTransform(int predL[16][16], int predCb[8][8], int predCr[8][8], bool reconstruct);
void ref_transformDecodingP_Skip(int predL[16][16], int predCb[8][8], int predCr[8][8], int QPy);
int ref_intraPredictionEncoding(int predL[16][16], int predCr[8][8], int predCb[8][8]);
unsigned int ref_coded_mb_size(int intra16x16PredMode, int predL[16][16], int predCb[8][8], int predCr[8][8]);
// non-static helpers of moestimation.cpp used by the taps
int satdLuma8x8MVs(int mvx, int mvy, int luma8x8BlkIdx);
extern int **refFrameKar[6][16];
extern int *sortedSuma0[5];
extern int koliko[16384];

typedef std::chrono::steady_clock clk;
static double t_inter = 0, t_tq = 0, t_skip = 0, t_fill = 0;
static inline double since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }

static FILE *dumpf = 0;
static int dumpmask = 0, planes_pic = -1, pic_index = 0;

static void chunk(const char tag[4], const void *p, size_t n)
{
	if (!dumpf) return;
	unsigned int hdr[3] = {0, (unsigned)pic_index, (unsigned)n};
	memcpy(&hdr[0], tag, 4);
	fwrite(hdr, 4, 3, dumpf);
	fwrite(p, 1, n, dumpf);
}

// ---- per-MB record: mb_type, mv[4][2], mvd[4][2], sad[4], luma[16][16], cdc[2][4], cac[2][4][15]
enum { REC_INTS = 1 + 8 + 8 + 4 + 256 + 8 + 120 };
static std::vector<int> mbrec;      // PicSizeInMbs * REC_INTS
static std::vector<unsigned char> tqio;  // per MB: snapped source 384 + prediction 384
static unsigned char savedL[256];
static int slice_data_bit0 = 0;           // writer position when the first macroblock of the slice starts
static std::vector<short> i16rec;        // per Intra16x16 MB: 256 src, 256 pred, 16 dc, 240 ac, 256 recon (as int16)

// ---- per-MB record of an I picture: mb_type (final), Intra16x16PredMode (-1 = Intra4x4), intra_chroma_pred_mode, bits of the
//      Intra16x16 trial, bits of the Intra4x4 trial, CBP luma, CBP chroma, Intra4x4PredMode[16], prev_intra4x4_pred_mode_flag[16],
//      rem_intra4x4_pred_mode[16], 256 luma levels (Intra4x4: LumaLevel[16][16]; Intra16x16: DC[16] then AC[16][15]), cdc[2][4], cac[2][4][15]
enum { IREC_INTS = 7 + 48 + 256 + 8 + 120 };
static std::vector<int> imbrec;

#ifndef FH264_NO_TAPS   // the integration build (integration/) supplies these entry points itself
static int *rec(int mb) { return &mbrec[(size_t)mb * REC_INTS]; }

void interEncoding(int predL[16][16], int predCr[8][8], int predCb[8][8])
{
	const int W = frame.Lwidth;
	const int xp = (CurrMbAddr % PicWidthInMbs) << 4, yp = (CurrMbAddr / PicWidthInMbs) << 4;
	if (CurrMbAddr == 0) slice_data_bit0 = (int)(RBSP_write_current_byte * 8 + RBSP_write_current_bit + RBSP_write_buffer_bit);
	const bool tap = (dumpmask & 1) != 0;
	if (tap)
		for (int r = 0; r < 16; r++) memcpy(savedL + r * 16, frame.L + (yp + r) * W + xp, 16);
	clk::time_point t0 = clk::now();
	ref_interEncoding(predL, predCr, predCb);
	t_inter += since(t0);
	if (!tap) return;
	int *R = rec(CurrMbAddr);
	memset(R, 0, sizeof(int) * REC_INTS);
	R[0] = mb_type;
	for (int q = 0; q < 4; q++) {
		R[1 + q * 2] = mvL0x[CurrMbAddr][q][0];
		R[2 + q * 2] = mvL0y[CurrMbAddr][q][0];
	}
	if (mb_type != P_Skip) {
		int parts = NumMbPart(mb_type);
		for (int i = 0; i < parts; i++) {
			R[9 + i * 2] = mvd_l0[i][0][0];
			R[10 + i * 2] = mvd_l0[i][0][1];
		}
		// SAD of the chosen quadrant MVs against the ORIGINAL source (before pixel snapping,
		// moestimation.cpp:571-584): restore, measure with the reference's own SAD, put back.
		unsigned char snapped[256];
		for (int r = 0; r < 16; r++) {
			memcpy(snapped + r * 16, frame.L + (yp + r) * W + xp, 16);
			memcpy(frame.L + (yp + r) * W + xp, savedL + r * 16, 16);
		}
		for (int q = 0; q < 4; q++) R[17 + q] = satdLuma8x8MVs(R[1 + q * 2], R[2 + q * 2], q);
		for (int r = 0; r < 16; r++) memcpy(frame.L + (yp + r) * W + xp, snapped + r * 16, 16);
	}
}

int intraPredictionEncoding(int predL[16][16], int predCr[8][8], int predCb[8][8])
{
	if (CurrMbAddr == 0) slice_data_bit0 = (int)(RBSP_write_current_byte * 8 + RBSP_write_current_bit + RBSP_write_buffer_bit);
	const int m = ref_intraPredictionEncoding(predL, predCr, predCb);
	if (dumpmask & 256) {
		int *R = &imbrec[(size_t)CurrMbAddr * IREC_INTS];
		R[1] = m;
		R[2] = intra_chroma_pred_mode;
		for (int b = 0; b < 16; b++) {
			R[7 + b] = Intra4x4PredMode[(CurrMbAddr << 4) + b];
			R[23 + b] = prev_intra4x4_pred_mode_flag[b] ? 1 : 0;
			R[39 + b] = rem_intra4x4_pred_mode[b];
		}
	}
	return m;
}

unsigned int coded_mb_size(int intra16x16PredMode, int predL[16][16], int predCb[8][8], int predCr[8][8])
{
	const unsigned int n = ref_coded_mb_size(intra16x16PredMode, predL, predCb, predCr);
	if (dumpmask & 256) imbrec[(size_t)CurrMbAddr * IREC_INTS + (intra16x16PredMode == -1 ? 4 : 3)] = (int)n;
	return n;
}

static void tap_tq_input(int predL[16][16], int predCb[8][8], int predCr[8][8])
{
	if (!(dumpmask & 16)) return;
	const int W = frame.Lwidth, CW = frame.Cwidth;
	const int xp = (CurrMbAddr % PicWidthInMbs) << 4, yp = (CurrMbAddr / PicWidthInMbs) << 4;
	unsigned char *o = &tqio[(size_t)CurrMbAddr * 768];
	for (int r = 0; r < 16; r++) for (int c = 0; c < 16; c++) *o++ = frame.L[(yp + r) * W + xp + c];
	for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) *o++ = frame.C[0][(yp / 2 + r) * CW + xp / 2 + c];
	for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) *o++ = frame.C[1][(yp / 2 + r) * CW + xp / 2 + c];
	for (int r = 0; r < 16; r++) for (int c = 0; c < 16; c++) *o++ = (unsigned char)predL[r][c];
	for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) *o++ = (unsigned char)predCb[r][c];
	for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) *o++ = (unsigned char)predCr[r][c];
}

void quantizationTransform(int predL[16][16], int predCb[8][8], int predCr[8][8], bool reconstruct)
{
	const bool p_pic = (shd.slice_type % 5) == P_SLICE;
	if (p_pic) tap_tq_input(predL, predCb, predCr);
	const bool i16 = !p_pic && reconstruct && (dumpmask & 32) && MbPartPredMode(mb_type, 0) == Intra_16x16;
	const int W = frame.Lwidth, xp = (CurrMbAddr % PicWidthInMbs) << 4, yp = (CurrMbAddr / PicWidthInMbs) << 4;
	size_t i16base = 0;
	if (i16) {
		i16base = i16rec.size();
		i16rec.resize(i16base + 1024);
		short *o = &i16rec[i16base];
		for (int r = 0; r < 16; r++) for (int c = 0; c < 16; c++) { o[r * 16 + c] = frame.L[(yp + r) * W + xp + c]; o[256 + r * 16 + c] = (short)predL[r][c]; }
	}
	clk::time_point t0 = clk::now();
	ref_quantizationTransform(predL, predCb, predCr, reconstruct);
	if (p_pic) t_tq += since(t0);
	if (i16) {
		short *o = &i16rec[i16base];
		for (int k = 0; k < 16; k++) o[512 + k] = (short)Intra16x16DCLevel[k];
		for (int b = 0; b < 16; b++) for (int k = 0; k < 15; k++) o[528 + b * 15 + k] = (short)Intra16x16ACLevel[b][k];
		for (int r = 0; r < 16; r++) for (int c = 0; c < 16; c++) o[768 + r * 16 + c] = frame.L[(yp + r) * W + xp + c];
	}
	if (!p_pic && reconstruct && (dumpmask & 256)) {
		int *R = &imbrec[(size_t)CurrMbAddr * IREC_INTS + 55];
		if (MbPartPredMode(mb_type, 0) == Intra_16x16) {
			for (int k = 0; k < 16; k++) *R++ = Intra16x16DCLevel[k];
			for (int b = 0; b < 16; b++) for (int k = 0; k < 15; k++) *R++ = Intra16x16ACLevel[b][k];
		} else
			for (int b = 0; b < 16; b++) for (int k = 0; k < 16; k++) *R++ = LumaLevel[b][k];
		for (int c = 0; c < 2; c++) for (int k = 0; k < 4; k++) *R++ = ChromaDCLevel[c][k];
		for (int c = 0; c < 2; c++) for (int b = 0; b < 4; b++) for (int k = 0; k < 15; k++) *R++ = ChromaACLevel[c][b][k];
	}
	if (p_pic && (dumpmask & 1)) {
		int *R = rec(CurrMbAddr) + 21;
		for (int b = 0; b < 16; b++) for (int k = 0; k < 16; k++) *R++ = LumaLevel[b][k];
		for (int c = 0; c < 2; c++) for (int k = 0; k < 4; k++) *R++ = ChromaDCLevel[c][k];
		for (int c = 0; c < 2; c++) for (int b = 0; b < 4; b++) for (int k = 0; k < 15; k++) *R++ = ChromaACLevel[c][b][k];
	}
}

void transformDecodingP_Skip(int predL[16][16], int predCb[8][8], int predCr[8][8], int qpy)
{
	tap_tq_input(predL, predCb, predCr);
	clk::time_point t0 = clk::now();
	ref_transformDecodingP_Skip(predL, predCb, predCr, qpy);
	t_skip += since(t0);
}

void FillInterpolatedRefFrame()
{
	clk::time_point t0 = clk::now();
	ref_FillInterpolatedRefFrame();
	t_fill += since(t0);
	if ((dumpmask & 8) && pic_index == planes_pic) {
		const int W = frame.Lwidth, H = frame.Lheight;
		for (int f = 0; f < 16; f++) chunk("PLNE", refFrameInterpolated[f].L, (size_t)W * H);
		std::vector<unsigned short> k((size_t)W * H);
		for (int f = 0; f < 16; f++)
			for (int kar = 0; kar < 5; kar++) {
				for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) k[(size_t)y * W + x] = (unsigned short)refFrameKar[kar][f][y][x];
				chunk("KARF", k.data(), k.size() * 2);
			}
		for (int a = 0; a < 5; a++) chunk("SORT", sortedSuma0[a], sizeof(int) * (size_t)W * H);
		chunk("KOLI", koliko, sizeof(koliko));
	}
}

#endif  // FH264_NO_TAPS

int main(int argc, char **argv)
{
	if (argc < 10) {
		fprintf(stderr, "usage: %s in.y4m out.264 dump.bin|- frames qp basic window maxdiff intraEvery [dumpmask [planes_pic]]\n", argv[0]);
		return 2;
	}
	const char *in = argv[1], *out = argv[2], *dump = argv[3];
	const int frames = atoi(argv[4]);
	_qParameter = atoi(argv[5]);
	BasicInterEncoding = atoi(argv[6]);
	WindowSize = atoi(argv[7]);
	MAXDIFF_SET = atoi(argv[8]);
	IntraEvery = atoi(argv[9]);
	dumpmask = argc > 10 ? atoi(argv[10]) : 0;
	planes_pic = argc > 11 ? atoi(argv[11]) : -1;
	startFrame = 1;
	endFrame = frames;

	stream = fopen(out, "wb");
	yuvinput = fopen(in, "rb");
	if (!stream || !yuvinput) { fprintf(stderr, "cannot open files\n"); return 2; }
	if (strcmp(dump, "-") != 0) dumpf = fopen(dump, "wb");

	generate_residual_level_tables();
	init_expgolomb_UC_codes();
	InitNAL();
	InitCL();

	NALunit nu;
	frameCount = 0;
	currFrameCount = 0;
	nu.rbsp_byte = new unsigned char[500000];
	nu.forbidden_zero_bit = 0;
	LoadY4MHeader();

	nu.nal_ref_idc = 1;
	nu.nal_unit_type = NAL_UNIT_TYPE_SPS;
	RBSP_encode(nu);
	writeNAL(nu);
	nu.nal_unit_type = NAL_UNIT_TYPE_PPS;
	RBSP_encode(nu);
	writeNAL(nu);

	const int W = frame.Lwidth, H = frame.Lheight, nmb = (W * H) >> 8;
	if (dumpmask & 128) {
		// coder tables as (length, code) int pairs, in a fixed order: coeff_token nC 0-1 / 2-3 / 4-7 / 8+ [17][4], chroma DC [17][4],
		// total_zeros 4x4 [15][16], total_zeros chroma DC [3][4], run_before [6][7], coded_block_pattern inter map [48]
		std::vector<int> t;
		auto add = [&](const int *len, const unsigned int *code, int n) { for (int i = 0; i < n; i++) { t.push_back(len[i]); t.push_back((int)code[i]); } };
		add(&CoeffTokenCodesCoder_nC_0_to_2_length[0][0], &CoeffTokenCodesCoder_nC_0_to_2_data_int[0][0], 68);
		add(&CoeffTokenCodesCoder_nC_2_to_4_length[0][0], &CoeffTokenCodesCoder_nC_2_to_4_data_int[0][0], 68);
		add(&CoeffTokenCodesCoder_nC_4_to_8_length[0][0], &CoeffTokenCodesCoder_nC_4_to_8_data_int[0][0], 68);
		add(&CoeffTokenCodesCoder_nC_8_to_max_length[0][0], &CoeffTokenCodesCoder_nC_8_to_max_data_int[0][0], 68);
		add(&CoeffTokenCodeTableCoder_ChromaDC_length[0][0], &CoeffTokenCodeTableCoder_ChromaDC_data_int[0][0], 68);
		add(&TotalZerosCodeTableCoder_4x4_length[0][0], &TotalZerosCodeTableCoder_4x4_data_int[0][0], 240);
		add(&TotalZerosCodeTableCoder_ChromaDC_length[0][0], &TotalZerosCodeTableCoder_ChromaDC_data_int[0][0], 12);
		add(&RunBeforeCodeTableCoder_length[0][0], &RunBeforeCodeTableCoder_data_int[0][0], 42);
		for (int i = 0; i < 48; i++) t.push_back(coded_block_pattern_to_codeNum_inter[i]);
		chunk("CVTB", t.data(), t.size() * sizeof(int));
	}
	mbrec.assign((size_t)nmb * REC_INTS, 0);
	tqio.assign((size_t)nmb * 768, 0);
	imbrec.assign((size_t)nmb * IREC_INTS, 0);

	std::string types, bytes, tpic, tsel, tin, ttq, tfl;
	double total = 0, total_p = 0, hot_p = 0;
	int npics = 0, np = 0;
	for (int n = 0; n < frames; n++) {
		if (ReadFromY4M() == -1) break;
		frameCount++;
		if (n > 0) currFrameCount++;
		for (int i = 0; i < 5; i++) brojTipova[i] = 0;
		pic_index = n;
		if (dumpmask & 4) {
			chunk("SRCY", frame.L, (size_t)W * H);
			chunk("SRCU", frame.C[0], (size_t)W * H / 4);
			chunk("SRCV", frame.C[1], (size_t)W * H / 4);
		}
		t_inter = t_tq = t_skip = t_fill = 0;
		clk::time_point t0 = clk::now();
		nu.nal_unit_type = selectNALUnitType();
		double ts = since(t0);
		RBSP_encode(nu);
		double tp = since(t0);
		writeNAL(nu);
		const bool isP = nu.nal_unit_type == NAL_UNIT_TYPE_NOT_IDR;
		int hdr[10] = {(int)nu.nal_unit_type, (int)nu.NumBytesInRBSP, W, H, brojTipova[0], brojTipova[1], brojTipova[2], brojTipova[3], brojTipova[4], QPy};
		chunk("PICH", hdr, sizeof hdr);
		if (isP && (dumpmask & 1)) chunk("MBRC", mbrec.data(), mbrec.size() * sizeof(int));
		if (isP && (dumpmask & 16)) chunk("TQIO", tqio.data(), tqio.size());
		if ((isP || (dumpmask & 256)) && (dumpmask & 64)) {
			std::vector<unsigned char> sl(4 + nu.NumBytesInRBSP);
			memcpy(sl.data(), &slice_data_bit0, 4);
			memcpy(sl.data() + 4, nu.rbsp_byte, nu.NumBytesInRBSP);
			chunk("SLDT", sl.data(), sl.size());
		}
		if (!isP && (dumpmask & 256)) {
			for (int m = 0; m < nmb; m++) {
				int *R = &imbrec[(size_t)m * IREC_INTS];
				R[0] = mb_type_array[m]; R[5] = CodedBlockPatternLumaArray[m]; R[6] = CodedBlockPatternChromaArray[m];
			}
			chunk("IMBR", imbrec.data(), imbrec.size() * sizeof(int));
		}
		if (!isP && (dumpmask & 32) && !i16rec.empty()) chunk("I16M", i16rec.data(), i16rec.size() * sizeof(short));
		i16rec.clear();
		if (dumpmask & 2) {
			chunk("RECY", frame.L, (size_t)W * H);
			chunk("RECU", frame.C[0], (size_t)W * H / 4);
			chunk("RECV", frame.C[1], (size_t)W * H / 4);
		}
		char b[64];
		types += isP ? 'P' : 'I';
		snprintf(b, sizeof b, "%s%u", npics ? "," : "", nu.NumBytesInRBSP); bytes += b;
		snprintf(b, sizeof b, "%s%.6f", npics ? "," : "", tp); tpic += b;
		snprintf(b, sizeof b, "%s%.6f", npics ? "," : "", ts); tsel += b;
		snprintf(b, sizeof b, "%s%.6f", npics ? "," : "", t_inter); tin += b;
		snprintf(b, sizeof b, "%s%.6f", npics ? "," : "", t_tq + t_skip); ttq += b;
		snprintf(b, sizeof b, "%s%.6f", npics ? "," : "", t_fill); tfl += b;
		total += tp;
		if (isP) { total_p += tp; hot_p += ts + t_inter + t_tq + t_skip + t_fill; np++; }
		npics++;
	}
	CloseCL();
	CloseNAL();
	fclose(stream);
	fclose(yuvinput);
	if (dumpf) fclose(dumpf);
	printf("{\"width\": %d, \"height\": %d, \"pictures\": %d, \"p_pictures\": %d, \"types\": \"%s\", \"bytes\": [%s], "
	       "\"t_picture\": [%s], \"t_select\": [%s], \"t_inter\": [%s], \"t_tq\": [%s], \"t_fill\": [%s], "
	       "\"total_s\": %.6f, \"total_p_s\": %.6f, \"hot_p_s\": %.6f}\n",
	       W, H, npics, np, types.c_str(), bytes.c_str(), tpic.c_str(), tsel.c_str(), tin.c_str(), ttq.c_str(), tfl.c_str(),
	       total, total_p, hot_p);
	return 0;
}
