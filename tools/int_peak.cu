// int_peak.cu -- issue rate of the integer instructions the kernels are made of, measured on this GPU: warp instructions per
// clock per SM for independent chains of one opcode (or a mix), at full occupancy.  Context for the instruction-issue roofline
// in DESIGN.md: the theoretical ceiling is 4 warp instructions / clock / SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/int_peak.cu -o build/int_peak && build/int_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CH = 8;

template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    if (OP == 0) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));                 // IMAD
    if (OP == 1) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));             // LOP3
    if (OP == 2) asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));               // IDP.4A
    if (OP == 3) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));            // IDP.2A
    if (OP == 4) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));  // VABSDIFF4
    if (OP == 5) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));             // SHF
    if (OP == 6) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));    // I2IP
    if (OP == 7) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(c));                                // IADD
    if (OP == 8) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c & 0x7777));         // PRMT
    return d;
}

// MIX == 0: CH independent chains of OP.  MIX == 1: half the chains OP, half OP2 (pipe pairing).
template <int OP, int OP2>
__global__ void __launch_bounds__(256) k_peak(uint32_t *out, uint32_t seed)
{
    uint32_t v[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = seed + threadIdx.x * 17 + i;
    const uint32_t b = seed | 1;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) v[i] = (i & 1) ? op<OP2>(v[i], b, v[i]) : op<OP>(v[i], b, v[i]);
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= v[i];
    if (s == 0x12345678u) out[0] = s;
}

template <int OP, int OP2> static void run(const char *name, int sms, double mhz)
{
    uint32_t *d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 8;
    k_peak<OP, OP2><<<grid, 256>>>(d, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_peak<OP, OP2><<<grid, 256>>>(d, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inst = (double)grid * 8 * ITERS * CH;          // 8 warps per CTA
    const double per_clk_sm = warp_inst / (ms * 1e-3) / (mhz * 1e6) / sms;
    printf("%-28s %7.3f warp-inst/clk/SM  (%.1f G warp-inst/s)\n", name, per_clk_sm, warp_inst / (ms * 1e-3) / 1e9);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, %.0f MHz (nominal; rates assume the GPU runs at it)\n", p.name, p.multiProcessorCount, mhz);
    const int sms = p.multiProcessorCount;
    run<0, 0>("IMAD", sms, mhz);
    run<1, 1>("LOP3", sms, mhz);
    run<7, 7>("IADD", sms, mhz);
    run<5, 5>("SHF.R.W", sms, mhz);
    run<8, 8>("PRMT", sms, mhz);
    run<2, 2>("IDP.4A (dp4a)", sms, mhz);
    run<3, 3>("IDP.2A (dp2a)", sms, mhz);
    run<4, 4>("VABSDIFF4.ACC", sms, mhz);
    run<6, 6>("I2IP (cvt.pack.sat)", sms, mhz);
    run<0, 1>("IMAD + LOP3 (1:1)", sms, mhz);
    run<2, 1>("IDP.4A + LOP3 (1:1)", sms, mhz);
    run<3, 5>("IDP.2A + SHF (1:1)", sms, mhz);
    run<4, 0>("VABSDIFF4 + IMAD (1:1)", sms, mhz);
    run<4, 1>("VABSDIFF4 + LOP3 (1:1)", sms, mhz);
    return 0;
}
