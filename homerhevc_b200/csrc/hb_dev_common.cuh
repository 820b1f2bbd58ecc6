// hb_dev_common.cuh -- device helpers shared by all kernels (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define HB_FULL_MASK 0xffffffffu

// unaligned 4-byte read of 8-bit samples: two aligned words + funnel shift (read-only path)
// three consecutive 32-bit words from the 4-byte aligned byte offset `off` of `base` (8-byte aligned), as TWO 64-bit loads: one L1
// request less than three word loads -- the search kernels are bound by L1 wavefronts (l1tex 70-90 % of peak), not by ALU work.
// Reads up to 4 bytes past the third word (inside the padded planes).
__device__ __forceinline__ void hb_ld_words3(const uint8_t *base, uint32_t off, uint32_t &w0, uint32_t &w1, uint32_t &w2)
{
    const uint2 a = __ldg(reinterpret_cast<const uint2 *>(base + (off & ~7u)));
    const uint2 b = __ldg(reinterpret_cast<const uint2 *>(base + (off & ~7u) + 8));
    const bool odd = (off & 4u) != 0;
    w0 = odd ? a.y : a.x; w1 = odd ? b.x : a.y; w2 = odd ? b.y : b.x;
}

__device__ __forceinline__ uint32_t hb_ld_u8x4(const uint8_t *p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    const uint32_t lo = __ldg(q), hi = __ldg(q + 1);
    return __funnelshift_r(lo, hi, static_cast<uint32_t>(a & 3) * 8u);
}

// acc + sum of |a.b[i] - b.b[i]| over the four packed bytes, one instruction (VABSDIFF4.U8.ACC)
__device__ __forceinline__ uint32_t hb_sad4_acc(uint32_t a, uint32_t b, uint32_t acc)
{
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}

// c + sum of four u8 (a) x s8 (b) products
__device__ __forceinline__ int hb_dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// four ints clipped to 0..255 and packed, v0 in the low byte (two I2IP)
__device__ __forceinline__ uint32_t hb_pack_sat_u8x4(int v0, int v1, int v2, int v3)
{
    uint32_t hi, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(v3), "r"(v2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v1), "r"(v0), "r"(hi));
    return d;
}

__device__ __forceinline__ int hb_clip255(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ int hb_sat16(int v) { return min(max(v, -32768), 32767); }

__device__ __forceinline__ uint32_t hb_pack4(int a, int b, int c, int d)
{
    return static_cast<uint32_t>(a) | (static_cast<uint32_t>(b) << 8) | (static_cast<uint32_t>(c) << 16) | (static_cast<uint32_t>(d) << 24);
}

// HEVC interpolation taps (hmr_motion_inter.c:240-258) as compile-time immediates
template <int F> __device__ __forceinline__ int hb_luma8(int a0, int a1, int a2, int a3, int a4, int a5, int a6, int a7)
{
    if (F == 1) return -a0 + 4 * a1 - 10 * a2 + 58 * a3 + 17 * a4 - 5 * a5 + a6;
    if (F == 2) return -a0 + 4 * a1 - 11 * a2 + 40 * a3 + 40 * a4 - 11 * a5 + 4 * a6 - a7;
    if (F == 3) return a1 - 5 * a2 + 17 * a3 + 58 * a4 - 10 * a5 + 4 * a6 - a7;
    return 64 * a3;
}
__device__ __forceinline__ int hb_luma8_dyn(int f, int a0, int a1, int a2, int a3, int a4, int a5, int a6, int a7)
{
    switch (f) {
    case 1: return hb_luma8<1>(a0, a1, a2, a3, a4, a5, a6, a7);
    case 2: return hb_luma8<2>(a0, a1, a2, a3, a4, a5, a6, a7);
    case 3: return hb_luma8<3>(a0, a1, a2, a3, a4, a5, a6, a7);
    default: return 64 * a3;
    }
}
__device__ __forceinline__ int hb_chroma4_dyn(int f, int a0, int a1, int a2, int a3)
{
    switch (f) {
    case 1: return -2 * a0 + 58 * a1 + 10 * a2 - 2 * a3;
    case 2: return -4 * a0 + 54 * a1 + 16 * a2 - 2 * a3;
    case 3: return -6 * a0 + 46 * a1 + 28 * a2 - 4 * a3;
    case 4: return -4 * a0 + 36 * a1 + 36 * a2 - 4 * a3;
    case 5: return -4 * a0 + 28 * a1 + 46 * a2 - 6 * a3;
    case 6: return -2 * a0 + 16 * a1 + 54 * a2 - 4 * a3;
    case 7: return -2 * a0 + 10 * a1 + 58 * a2 - 2 * a3;
    default: return 64 * a1;
    }
}

// ---- packed luma filter taps (shared by the search and the compensation kernels).  Horizontal: u8 samples x s8 taps through dp4a, taps k..k+3 of fraction f in one word.
__host__ __device__ constexpr int luma_tap(int f, int k)
{
    constexpr int t[4][8] = { {0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1} };
    return t[f][k];
}
__host__ __device__ constexpr uint32_t pack_s8(int b0, int b1, int b2, int b3)
{
    return static_cast<uint32_t>(b0 & 255) | (static_cast<uint32_t>(b1 & 255) << 8) | (static_cast<uint32_t>(b2 & 255) << 16) | (static_cast<uint32_t>(b3 & 255) << 24);
}
__host__ __device__ constexpr uint32_t htap4(int f, int half)
{
    return pack_s8(luma_tap(f, 4 * half), luma_tap(f, 4 * half + 1), luma_tap(f, 4 * half + 2), luma_tap(f, 4 * half + 3));
}
