// Integer-pipe micro-benchmark for the ME roofline ("MEASURED_PEAKS.json has no integer-throughput
// entry — add one before quoting an ME roofline fraction"): sustained per-lane rate of the instructions the ME kernels are made
// of, on all SMs, 8 independent dependency chains per thread, 2048 threads per SM. Prints one JSON line.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu ; SASS of the loop bodies checked with cuobjdump.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024) k(unsigned *out, int iters, unsigned x, unsigned y)
{
    unsigned r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = threadIdx.x * 0x9e3779b9u + blockIdx.x + j * 0x01010101u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (OP == 0) r[j] = r[j] + x + (r[j] >> 31);           // LEA.HI + VIADD
                if (OP == 1) r[j] = __vsadu4(r[j], x) + y;            // VABSDIFF4.U8.ACC (sum of 4 byte differences) + VIADD
                if (OP == 2) r[j] = r[j] * x + y;                     // IMAD
                if (OP == 3) r[j] = (r[j] & x) ^ (r[j] >> 1);         // SHF + LOP3
                if (OP == 4) r[j] = __vmaxs2(__vadd2(r[j], x), y);    // VIADDMNMX.S16x2 (packed add + max), as in the feature distance
                if (OP == 5) r[j] = max((int)(r[j] + x), (int)y);     // VIADDMNMX (scalar add + max)
            }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= r[j];
    if (s == 0x12345678u) out[0] = s;      // keeps the chains alive, practically never stores
}

template <int OP>
double run(unsigned *d, int sms, int iters)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<sms * 2, 1024>>>(d, 64, 0x00030201u, 0x00010001u);       // warm-up
    cudaEventRecord(a);
    k<OP><<<sms * 2, 1024>>>(d, iters, 0x00030201u, 0x00010001u);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return (double)sms * 2 * 1024 * (double)iters * 16 * 8 / (ms * 1e-3) / 1e12;     // T statements per second (per lane)
}

int main()
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
    unsigned *d;
    cudaMalloc(&d, 4);
    const int sms = p.multiProcessorCount, iters = 512;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double t0 = run<0>(d, sms, iters), t1 = run<1>(d, sms, iters), t2 = run<2>(d, sms, iters), t3 = run<3>(d, sms, iters), t4 = run<4>(d, sms, iters), t5 = run<5>(d, sms, iters);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %d, \"unit\": \"T statements/s over all lanes (one statement = the instructions named)\", "
           "\"lea_hi+viadd (2 instr)\": %.3f, \"vabsdiff4.u8.acc+viadd (2 instr)\": %.3f, \"imad (1 instr)\": %.3f, \"shf+lop3 (2 instr)\": %.3f, \"viaddmnmx.s16x2 (1 instr)\": %.3f, \"viaddmnmx (1 instr)\": %.3f, "
           "\"nominal_lane_rate\": %.3f}\n",
           p.name, sms, clk / 1000, t0, t1, t2, t3, t4, t5, sms * 128.0 * clk * 1e3 / 1e12);
    return 0;
}
