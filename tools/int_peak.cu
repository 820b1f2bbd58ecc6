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
    if (OP == 9) asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(d) : "r"(a ^ threadIdx.x));         // LOP3 + REDUX.SUM
    if (OP == 10) asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, %2; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "=r"(d) : "r"(a), "r"(c ^ b));   // ISETP + VOTE
    if (OP == 11) asm volatile("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(c ^ b));                         // VIMNMX.U16x2
    if (OP == 12) asm volatile("match.any.sync.b32 %0, %1, 0xffffffff;" : "=r"(d) : "r"((a ^ threadIdx.x) & 7u)); // LOP3 + MATCH.ANY
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

// shared-memory histogram updates, the forms a statistics kernel can choose from: a private column per lane (plain load / add / store,
// bank = lane), an atomic on a per-warp table with the lanes spread over `spread` bins (same-address lanes serialise)
template <int MODE>
__global__ void __launch_bounds__(256) k_smem(uint32_t *out, uint32_t seed, int spread)
{
    __shared__ uint32_t s[8][40][32];
    for (int i = threadIdx.x; i < 8 * 40 * 32; i += 256) (&s[0][0][0])[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t x = seed + threadIdx.x * 2654435761u;
    for (int it = 0; it < ITERS; it++) {
        x = x * 1664525u + 1013904223u;
        const uint32_t bin = (x >> 16) % static_cast<uint32_t>(spread);
        if (MODE == 0) s[warp][bin][lane] += 0x10001u;
        else atomicAdd(&s[warp][0][bin], 0x10001u);
    }
    __syncthreads();
    uint32_t t = 0;
    for (int i = threadIdx.x; i < 8 * 40 * 32; i += 256) t += (&s[0][0][0])[i];
    if (t == 0x12345678u) out[0] = t;
}

static void run_smem(int sms, double mhz)
{
    uint32_t *d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 3;                                        // 40 KB of shared memory per CTA
    for (int mode = 0; mode < 2; mode++)
        for (int spread : { 1, 2, 5, 32, 40 }) {
            if (mode == 0 && spread != 40) continue;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) k_smem<0><<<grid, 256>>>(d, 3, spread); else k_smem<1><<<grid, 256>>>(d, 3, spread);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double upd = (double)grid * 8 * ITERS;                 // warp-wide updates
            printf("%-44s %7.3f warp-updates/clk/SM\n", mode == 0 ? "private column load/add/store (+ LCG, modulo)" : (spread == 1 ? "atomicAdd, 1 bin" : spread == 2 ? "atomicAdd, 2 bins" : spread == 5 ? "atomicAdd, 5 bins" : spread == 32 ? "atomicAdd, 32 bins" : "atomicAdd, 40 bins"),
                   upd / (ms * 1e-3) / (mhz * 1e6) / sms);
        }
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
    run<9, 9>("LOP3 + REDUX.SUM (pairs)", sms, mhz);
    run<10, 10>("ISETP + VOTE.BALLOT (pairs)", sms, mhz);
    run<11, 11>("VIMNMX.U16x2", sms, mhz);
    run<12, 12>("LOP3 + MATCH.ANY (pairs)", sms, mhz);
    run_smem(sms, mhz);
    return 0;
}
