// hb_tq_core.cuh -- warp-level transform / quantisation building blocks.
//
// One warp owns a stack of 32 rows = 32/N transform units of N x N int16 samples in shared memory
// (row stride N+2 so that every row, column and element-wise access pattern below is bank-conflict free).
// A lane runs a whole N-point 1-D transform in registers; the HEVC matrix never touches memory: the
// even/odd (partial butterfly) recursion is fully unrolled and every matrix entry is folded into an
// integer multiply-add immediate.  All arithmetic is exact int32, so the results equal the matrix
// products of the reference's partialButterfly* (hmr_transform.c:172-507) bit for bit.
#pragma once
#include "hb_dev_common.cuh"

// 64*sqrt(2)*cos(m*pi/64) as rounded by HEVC: the 32 magnitudes of the core transform (m = 0..31; m = 32 is 0)
__device__ __forceinline__ constexpr int hb_dct_mag(int m)
{
    switch (m) {
    case 0: return 64; case 1: return 90; case 2: return 90; case 3: return 90; case 4: return 89; case 5: return 88;
    case 6: return 87; case 7: return 85; case 8: return 83; case 9: return 82; case 10: return 80; case 11: return 78;
    case 12: return 75; case 13: return 73; case 14: return 70; case 15: return 67; case 16: return 64; case 17: return 61;
    case 18: return 57; case 19: return 54; case 20: return 50; case 21: return 46; case 22: return 43; case 23: return 38;
    case 24: return 36; case 25: return 31; case 26: return 25; case 27: return 22; case 28: return 18; case 29: return 13;
    case 30: return 9; case 31: return 4; default: return 0;
    }
}
// entry [k][j] of the n-point matrix
__device__ __forceinline__ constexpr int hb_dct_coef(int n, int k, int j)
{
    int m = (k * (32 / n) * (2 * j + 1)) & 127;
    if (m > 64) m = 128 - m;
    return m <= 32 ? hb_dct_mag(m) : -hb_dct_mag(64 - m);
}

// y = T_N * x (no rounding).  Even rows recurse on the folded half, odd rows are a dense N/2 x N/2 product.
template <int N> __device__ __forceinline__ void hb_fwd1d(const int (&x)[N], int (&y)[N])
{
    if constexpr (N == 1) {
        y[0] = 64 * x[0];
    } else {
        int e[N / 2], o[N / 2], ye[N / 2];
#pragma unroll
        for (int j = 0; j < N / 2; j++) { e[j] = x[j] + x[N - 1 - j]; o[j] = x[j] - x[N - 1 - j]; }
        hb_fwd1d<N / 2>(e, ye);
#pragma unroll
        for (int i = 0; i < N / 2; i++) {
            int s = 0;
#pragma unroll
            for (int j = 0; j < N / 2; j++) s += hb_dct_coef(N, 2 * i + 1, j) * o[j];
            y[2 * i] = ye[i];
            y[2 * i + 1] = s;
        }
    }
}
// x = T_N^T * y (no rounding)
template <int N> __device__ __forceinline__ void hb_inv1d(const int (&y)[N], int (&x)[N])
{
    if constexpr (N == 1) {
        x[0] = 64 * y[0];
    } else {
        int ye[N / 2], ee[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; i++) ye[i] = y[2 * i];
        hb_inv1d<N / 2>(ye, ee);
#pragma unroll
        for (int j = 0; j < N / 2; j++) {
            int s = 0;
#pragma unroll
            for (int i = 0; i < N / 2; i++) s += hb_dct_coef(N, 2 * i + 1, j) * y[2 * i + 1];
            x[j] = ee[j] + s;
            x[N - 1 - j] = ee[j] - s;
        }
    }
}
// 4-point DST-VII (hmr_transform.c:133/:152)
__device__ __forceinline__ void hb_dst4_fwd(const int (&x)[4], int (&y)[4])
{
    y[0] = 29 * x[0] + 55 * x[1] + 74 * x[2] + 84 * x[3];
    y[1] = 74 * (x[0] + x[1] - x[3]);
    y[2] = 84 * x[0] - 29 * x[1] - 74 * x[2] + 55 * x[3];
    y[3] = 55 * x[0] - 84 * x[1] + 74 * x[2] - 29 * x[3];
}
__device__ __forceinline__ void hb_dst4_inv(const int (&y)[4], int (&x)[4])
{
    x[0] = 29 * y[0] + 74 * y[1] + 84 * y[2] + 55 * y[3];
    x[1] = 55 * y[0] + 74 * y[1] - 29 * y[2] - 84 * y[3];
    x[2] = 74 * (y[0] - y[2] + y[3]);
    x[3] = 84 * y[0] - 74 * y[1] + 55 * y[2] - 29 * y[3];
}

template <int N> struct HbTq {
    static constexpr int S = N + 2;            // row stride in int16 (odd number of 32-bit words)
    static constexpr int ELEMS = 32 * S;       // one warp's stack of 32 rows
    static constexpr int TPW = 32 / N;         // transform units per warp
    static constexpr int LOG2 = (N == 4) ? 2 : (N == 8) ? 3 : (N == 16) ? 4 : 5;

    // One forward stage: lane reads row `lane` of `in`, writes column `lane%N` of its unit in `out`
    // (out[unit*N + k][lane%N]); value = (sum + add) >> shift truncated to int16 (hmr_transform.c:172).
    template <bool DST> static __device__ __forceinline__ void fwd_stage(const int16_t *in, int16_t *out, int lane, int shift)
    {
        int x[N], y[N];
        const int16_t *row = in + lane * S;
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            const uint32_t w = *reinterpret_cast<const uint32_t *>(row + j);
            x[j] = static_cast<int16_t>(w & 0xffff);
            x[j + 1] = static_cast<int32_t>(w) >> 16;
        }
        if constexpr (DST && N == 4) hb_dst4_fwd(reinterpret_cast<const int(&)[4]>(x), reinterpret_cast<int(&)[4]>(y));
        else hb_fwd1d<N>(x, y);
        const int add = 1 << (shift - 1);
        int16_t *col = out + (lane / N) * N * S + (lane % N);
#pragma unroll
        for (int k = 0; k < N; k++) col[k * S] = static_cast<int16_t>((y[k] + add) >> shift);
    }
    // The same stage on a row that is already in registers (the residual of the fused chain)
    static __device__ __forceinline__ void fwd_stage_regs(const int (&x)[N], int16_t *out, int lane, int shift)
    {
        int y[N];
        hb_fwd1d<N>(x, y);
        const int add = 1 << (shift - 1);
        int16_t *col = out + (lane / N) * N * S + (lane % N);
#pragma unroll
        for (int k = 0; k < N; k++) col[k * S] = static_cast<int16_t>((y[k] + add) >> shift);
    }
    // One inverse stage: lane reads column `lane%N` of its unit in `in`, writes row `lane` of `out`,
    // value = clip16((sum + add) >> shift) (hmr_transform.c:195).
    template <bool DST> static __device__ __forceinline__ void inv_stage(const int16_t *in, int16_t *out, int lane, int shift)
    {
        int y[N], x[N];
        const int16_t *col = in + (lane / N) * N * S + (lane % N);
#pragma unroll
        for (int k = 0; k < N; k++) y[k] = col[k * S];
        if constexpr (DST && N == 4) hb_dst4_inv(reinterpret_cast<const int(&)[4]>(y), reinterpret_cast<int(&)[4]>(x));
        else hb_inv1d<N>(y, x);
        const int add = 1 << (shift - 1);
        int16_t *row = out + lane * S;
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            const int a = hb_sat16((x[j] + add) >> shift), b = hb_sat16((x[j + 1] + add) >> shift);
            *reinterpret_cast<uint32_t *>(row + j) = (static_cast<uint32_t>(a) & 0xffffu) | (static_cast<uint32_t>(b) << 16);
        }
    }

    // First inverse stage with the dequantisation (hmr_sse42_functions_quant.c:135, iq_shift = log2N + 3) folded into its column
    // loads: lane reads column lane%N of the LEVELS of its unit and of the dequant table (consecutive lanes, consecutive words)
    static __device__ __forceinline__ void inv_stage1_dequant(const int16_t *Lv, const int32_t *__restrict__ dqtab, int per, int16_t *out, int lane)
    {
        int y[N], x[N];
        const int c = lane % N;
        const int16_t *col = Lv + (lane / N) * N * S + c;
        constexpr int iq_shift = LOG2 + 3;
#pragma unroll
        for (int k = 0; k < N; k++) {
            const uint32_t prod = static_cast<uint32_t>(static_cast<int>(col[k * S])) * static_cast<uint32_t>(__ldg(dqtab + k * N + c));
            int v;
            if (iq_shift > per) v = static_cast<int32_t>(prod + (1u << (iq_shift - per - 1))) >> (iq_shift - per);
            else v = static_cast<int32_t>(prod << (per - iq_shift));
            y[k] = hb_sat16(v);
        }
        hb_inv1d<N>(y, x);
        int16_t *row = out + lane * S;
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            const int a = hb_sat16((x[j] + 64) >> 7), b = hb_sat16((x[j + 1] + 64) >> 7);
            *reinterpret_cast<uint32_t *>(row + j) = (static_cast<uint32_t>(a) & 0xffffu) | (static_cast<uint32_t>(b) << 16);
        }
    }
    // Second inverse stage (shift 12) that leaves its row in registers
    static __device__ __forceinline__ void inv_stage2_regs(const int16_t *in, int (&x)[N], int lane)
    {
        int y[N];
        const int16_t *col = in + (lane / N) * N * S + (lane % N);
#pragma unroll
        for (int k = 0; k < N; k++) y[k] = col[k * S];
        hb_inv1d<N>(y, x);
#pragma unroll
        for (int j = 0; j < N; j++) x[j] = hb_sat16((x[j] + 2048) >> 12);
    }

    // 2-D forward transform of the warp's units: X -> C (T is scratch).  8-bit video: shifts log2N-1 and log2N+6.
    template <bool DST> static __device__ __forceinline__ void forward(const int16_t *X, int16_t *T, int16_t *C, int lane)
    {
        fwd_stage<DST>(X, T, lane, LOG2 - 1);
        __syncwarp();
        fwd_stage<DST>(T, C, lane, LOG2 + 6);
        __syncwarp();
    }
    // 2-D inverse: D -> R (T scratch), shifts 7 and 12
    template <bool DST> static __device__ __forceinline__ void inverse(const int16_t *D, int16_t *T, int16_t *R, int lane)
    {
        inv_stage<DST>(D, T, lane, 7);
        __syncwarp();
        inv_stage<DST>(T, R, lane, 12);
        __syncwarp();
    }

    // ---- element-wise sweeps handle FOUR consecutive samples of a row per lane and iteration
    static constexpr int GPU4 = N * N / 4;                        // groups of four per unit
    static constexpr int ITERS4 = N / 4;                          // 8N groups in the stack / 32 lanes
    static constexpr int UPI = GPU4 >= 32 ? 1 : 32 / GPU4;        // units covered by one iteration (N=4: 8, N=8: 2)
    static constexpr int IPU = GPU4 >= 32 ? GPU4 / 32 : 1;        // iterations per unit

    struct G4 { int row, col, off, pos4, unit; };                 // stack row, first column, smem offset, position inside the unit
    static __device__ __forceinline__ G4 group4(int it, int lane)
    {
        G4 g;
        const int i = it * 32 + lane;
        g.row = i / (N / 4); g.col = (i % (N / 4)) * 4; g.off = g.row * S + g.col;
        g.pos4 = (i % GPU4) * 4; g.unit = i / GPU4;
        return g;
    }
    static __device__ __forceinline__ void ld4(const int16_t *p, int (&v)[4])
    {
        const uint32_t a = *reinterpret_cast<const uint32_t *>(p), b = *reinterpret_cast<const uint32_t *>(p + 2);
        v[0] = static_cast<int16_t>(a & 0xffff); v[1] = static_cast<int32_t>(a) >> 16;
        v[2] = static_cast<int16_t>(b & 0xffff); v[3] = static_cast<int32_t>(b) >> 16;
    }
    static __device__ __forceinline__ void st4(int16_t *p, const int (&v)[4])
    {
        *reinterpret_cast<uint32_t *>(p) = (static_cast<uint32_t>(v[0]) & 0xffffu) | (static_cast<uint32_t>(v[1]) << 16);
        *reinterpret_cast<uint32_t *>(p + 2) = (static_cast<uint32_t>(v[2]) & 0xffffu) | (static_cast<uint32_t>(v[3]) << 16);
    }
    // add this iteration's per-unit totals of v into acc[] (identical in every lane)
    template <class T> static __device__ __forceinline__ void unit_add(T v, int it, T (&acc)[TPW])
    {
        if constexpr (UPI == 1) {
            v = __reduce_add_sync(HB_FULL_MASK, v);
#pragma unroll
            for (int u = 0; u < TPW; u++) if (u == it / IPU) acc[u] += v;
        } else {
            constexpr int W = 32 / UPI;                           // lanes per unit
#pragma unroll
            for (int d = W / 2; d > 0; d >>= 1) v += __shfl_xor_sync(HB_FULL_MASK, v, d);
#pragma unroll
            for (int k = 0; k < UPI; k++) {
                const T t = __shfl_sync(HB_FULL_MASK, v, k * W);
#pragma unroll
                for (int u = 0; u < TPW; u++) if (u == it * UPI + k) acc[u] += t;
            }
        }
    }

    // Quantise C -> L with deltaU -> U (hmr_sse42_functions_quant.c:52-118).  qtab is the N*N int32 table of this
    // (list, qp%6).  unit_sum[u] (u < TPW) receives the sum of levels of unit u, identical in every lane.
    static __device__ __forceinline__ void quantise(const int16_t *C, int16_t *L, int16_t *U, const int32_t *__restrict__ qtab,
                                                    int qbits, int add, int lane, int (&unit_sum)[TPW])
    {
#pragma unroll
        for (int u = 0; u < TPW; u++) unit_sum[u] = 0;
#pragma unroll
        for (int it = 0; it < ITERS4; it++) {
            const G4 g = group4(it, lane);
            int c[4], lv[4], du[4];
            ld4(C + g.off, c);
            const int4 q = __ldg(reinterpret_cast<const int4 *>(qtab + g.pos4));
            const int qq[4] = { q.x, q.y, q.z, q.w };
            int sum = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t a = static_cast<uint32_t>(c[k] < 0 ? -c[k] : c[k]);
                const uint32_t prod = a * static_cast<uint32_t>(qq[k]);
                const int level = static_cast<int32_t>(prod + static_cast<uint32_t>(add)) >> qbits;
                const int delta = static_cast<int32_t>(prod - (static_cast<uint32_t>(level) << qbits)) >> (qbits - 8);
                const int sat = hb_sat16(level);
                lv[k] = c[k] > 0 ? sat : (c[k] < 0 ? -sat : 0);
                du[k] = hb_sat16(delta);
                sum += level;
            }
            st4(L + g.off, lv);
            st4(U + g.off, du);
            unit_add<int>(sum, it, unit_sum);
        }
        __syncwarp();
    }

    // Sign-data hiding (hmr_quant.c:61) on the units whose level sum is >= 2.  One lane per 4x4 coefficient group.
    // scan: diagonal (or hor/ver) scan of the N x N unit as raster positions.
    static __device__ __forceinline__ void sign_hide(int16_t *L, const int16_t *C, const int16_t *U, const uint16_t *__restrict__ scan,
                                                     int lane, const int (&unit_sum)[TPW])
    {
        constexpr int CG_PER_UNIT = N * N / 16;
        constexpr int CGS = TPW * CG_PER_UNIT;         // 2N groups in the stack
        constexpr int ROUNDS = (CGS + 31) / 32;
        bool any_unit = false;
#pragma unroll
        for (int u = 0; u < TPW; u++) any_unit |= unit_sum[u] >= 2;
        if (!any_unit) return;                            // warp-uniform: nothing to hide anywhere in the stack
        // non-zero map of the coefficient groups, so that each group knows whether it is the last significant one of its unit
        uint32_t nz[ROUNDS];
        int first[ROUNDS], last[ROUNDS], asum[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; r++) {
            const int cg = r * 32 + lane;
            first[r] = 16; last[r] = -1; asum[r] = 0;
            int my_sum = 0;
#pragma unroll
            for (int u = 0; u < TPW; u++) if (u == cg / CG_PER_UNIT) my_sum = unit_sum[u];
            if (cg < CGS && my_sum >= 2) {
                const int unit = cg / CG_PER_UNIT, sub = cg % CG_PER_UNIT;
                const int16_t *Lu = L + unit * N * S;
                const uint16_t *sc = scan + 16 * sub;
                for (int i = 0; i < 16; i++) {
                    const int p = sc[i];
                    const int v = Lu[(p / N) * S + (p % N)];
                    if (v) { if (first[r] == 16) first[r] = i; last[r] = i; }
                }
                for (int i = first[r]; i <= last[r]; i++) { const int p = sc[i]; asum[r] += Lu[(p / N) * S + (p % N)]; }
            }
            nz[r] = __ballot_sync(HB_FULL_MASK, last[r] >= 0);
        }
#pragma unroll
        for (int r = 0; r < ROUNDS; r++) {
            const int cg = r * 32 + lane;
            if (cg >= CGS) continue;
            const int unit = cg / CG_PER_UNIT, sub = cg % CG_PER_UNIT;
            int usum = 0;
#pragma unroll
            for (int u = 0; u < TPW; u++) if (u == unit) usum = unit_sum[u];
            if (usum < 2 || last[r] - first[r] < 4) continue;
            // is there a significant group after this one inside the same unit?
            bool later = false;
            if constexpr (CG_PER_UNIT > 1) {
                if constexpr (ROUNDS == 2) {                      // N = 32: one unit, 64 groups
                    const uint64_t m = (static_cast<uint64_t>(nz[1]) << 32) | nz[0];
                    later = (m >> cg) >> 1 != 0;
                } else {
                    const uint32_t unit_mask = (CG_PER_UNIT == 32) ? 0xffffffffu : (((1u << CG_PER_UNIT) - 1u) << (unit * CG_PER_UNIT));
                    const uint32_t m = nz[0] & unit_mask;
                    later = sub + 1 < CG_PER_UNIT && (m >> (cg + 1)) != 0;
                }
            }
            int16_t *Lu = L + unit * N * S;
            const int16_t *Cu = C + unit * N * S, *Uu = U + unit * N * S;
            const uint16_t *sc = scan + 16 * sub;
            const int pf = sc[first[r]];
            const unsigned sign_bit = Lu[(pf / N) * S + (pf % N)] > 0 ? 0u : 1u;
            if (sign_bit == static_cast<unsigned>(asum[r] & 1)) continue;
            int min_cost = 0x7fffffff, min_off = -1, final_change = 0, cur_cost = 0x7fffffff, cur_change = 0;
            for (int i = later ? 15 : last[r]; i >= 0; i--) {
                const int p = sc[i];
                const int off = (p / N) * S + (p % N);
                const int lv = Lu[off], du = Uu[off];
                if (lv != 0) {
                    if (du > 0) { cur_cost = -du; cur_change = 1; }
                    else if (i == first[r] && (lv == 1 || lv == -1)) cur_cost = 0x7fffffff;
                    else { cur_cost = du; cur_change = -1; }
                } else if (i < first[r]) {
                    const unsigned this_sign = Cu[off] >= 0 ? 0u : 1u;
                    if (this_sign != sign_bit) cur_cost = 0x7fffffff;
                    else { cur_cost = -du; cur_change = 1; }
                } else {
                    cur_cost = -du; cur_change = 1;
                }
                if (cur_cost < min_cost) { min_cost = cur_cost; final_change = cur_change; min_off = off; }
            }
            const int lv = Lu[min_off];
            if (lv == 32767 || lv == -32768) final_change = -1;
            Lu[min_off] = static_cast<int16_t>(Cu[min_off] >= 0 ? lv + final_change : lv - final_change);
        }
        __syncwarp();
    }

    // L -> D (hmr_sse42_functions_quant.c:135): iq_shift = log2N + 3
    static __device__ __forceinline__ void dequantise(const int16_t *L, int16_t *D, const int32_t *__restrict__ dqtab, int per, int lane)
    {
        const int iq_shift = LOG2 + 3;
#pragma unroll
        for (int it = 0; it < ITERS4; it++) {
            const G4 g = group4(it, lane);
            int lv[4], d[4];
            ld4(L + g.off, lv);
            const int4 q = __ldg(reinterpret_cast<const int4 *>(dqtab + g.pos4));
            const int qq[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t prod = static_cast<uint32_t>(lv[k]) * static_cast<uint32_t>(qq[k]);
                int v;
                if (iq_shift > per) v = static_cast<int32_t>(prod + (1u << (iq_shift - per - 1))) >> (iq_shift - per);
                else v = static_cast<int32_t>(prod << (per - iq_shift));
                d[k] = hb_sat16(v);
            }
            st4(D + g.off, d);
        }
        __syncwarp();
    }
};
