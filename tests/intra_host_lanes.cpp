// TEST INFRASTRUCTURE — the device I-picture core (h264_fer_b200/csrc/intra_core.h) run on the host the way a warp runs it:
// 32 threads per macroblock share one IcCtx, ic_sync is a barrier, the reductions go through a shared array. Checks the
// cross-lane structure of ic_macroblock (what is shared, what each lane computes, where the barriers are) without a GPU; built
// with -fsanitize=thread it also reports unsynchronised accesses to the shared state. Never part of the product path.
#include <barrier>
#include <thread>
#include <vector>

static thread_local int t_lane = 0;
static std::barrier<> *g_bar = nullptr;
static int g_red[32];

#define IC_CUSTOM_LANES
static inline void ic_sync(int nl) { if (nl > 1) g_bar->arrive_and_wait(); }
static inline int ic_red_add(int v, int nl)
{
    if (nl <= 1) return v;
    g_red[t_lane] = v;
    g_bar->arrive_and_wait();
    int s = 0;
    for (int i = 0; i < nl; i++) s += g_red[i];
    g_bar->arrive_and_wait();
    return s;
}
static inline int ic_red_min(int v, int nl)
{
    if (nl <= 1) return v;
    g_red[t_lane] = v;
    g_bar->arrive_and_wait();
    int s = g_red[0];
    for (int i = 1; i < nl; i++) s = g_red[i] < s ? g_red[i] : s;
    g_bar->arrive_and_wait();
    return s;
}

#include "../h264_fer_b200/csrc/intra_core.h"

extern "C" int intra_host_picture_lanes(const unsigned char *sy, const unsigned char *su, const unsigned char *sv, unsigned char *ry, unsigned char *ru,
                                        unsigned char *rv, int W, int H, int qp, const int *prev_types, fh264_mb_result_i *out)
{
    const int wmb = W / 16, nmb = wmb * (H / 16), NL = 32;
    std::vector<IcInfo> info(nmb);
    std::barrier<> bar(NL);
    g_bar = &bar;
    for (int m = 0; m < nmb; m++) {
        IcCtx c;
        c.src[0] = sy; c.src[1] = su; c.src[2] = sv; c.rec[0] = ry; c.rec[1] = ru; c.rec[2] = rv;
        c.W = W; c.H = H; c.xP = (m % wmb) * 16; c.yP = (m / wmb) * 16; c.qp = qp;
        const bool prev_skip = prev_types && prev_types[m] == 31;
        const IcInfo *left = (m % wmb) ? &info[m - 1] : nullptr, *up = m >= wmb ? &info[m - wmb] : nullptr;
        std::vector<std::thread> th;
        for (int l = 0; l < NL; l++)
            th.emplace_back([&, l]() { t_lane = l; ic_macroblock(c, prev_skip, left, up, out[m], info[m], l, NL); });
        for (auto &t : th) t.join();
    }
    g_bar = nullptr;
    return 0;
}
