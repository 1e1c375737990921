// Integer-pipe micro-benchmark behind bench.py's roofline denominator ("INT32 peak must be measured by a
// vabsdiff4 / IADD3 micro-benchmark on the box and recorded next to it"). Sustained per-lane rate of the instructions the ME
// kernels are made of, on all SMs: 2048 threads per SM, 8 independent dependency chains per thread, 128 statements per loop
// iteration. The SASS of the loop bodies is the named instruction (cuobjdump; profiles/r02_sass_excerpts.md).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

template <int OP>
__global__ void __launch_bounds__(1024) k_int_peak(unsigned *out, int iters, unsigned x, unsigned y)
{
    unsigned r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = threadIdx.x * 0x9e3779b9u + blockIdx.x + j * 0x01010101u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (OP == 0) r[j] = r[j] * x + y;                     // IMAD
                if (OP == 1) r[j] = __vmaxs2(__vadd2(r[j], x), y);    // VIADDMNMX.S16x2 (packed add + max), as in the feature distance
                if (OP == 2) r[j] = max((int)(r[j] + x), (int)y);     // VIADDMNMX (scalar add + max)
                if (OP == 3) r[j] = __vsadu4(r[j], x) + y;            // VABSDIFF4.U8.ACC + VIADD (two instructions: SAD + add)
            }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= r[j];
    if (s == 0x12345678u) out[0] = s;      // keeps the chains alive, practically never stores
}

template <int OP>
static double int_peak_run(unsigned *d, int sms, int iters, cudaStream_t st)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_int_peak<OP><<<sms * 2, 1024, 0, st>>>(d, 64, 0x00030201u, 0x00010001u);       // warm-up
    cudaEventRecord(a, st);
    k_int_peak<OP><<<sms * 2, 1024, 0, st>>>(d, iters, 0x00030201u, 0x00010001u);
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return (double)sms * 2 * 1024 * (double)iters * 16 * 8 / (ms * 1e-3) / 1e12;     // T statements per second over all lanes
}
