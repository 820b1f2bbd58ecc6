// hb_kernels_tq.cu -- fused inter T/Q chain (encode_inter_cu / encode_inter_cu_chroma, hmr_motion_inter.c:40/:133)
// and the per-call transform / quant kernels built from the same warp routines.
//
// k_tq<N> (N = 8, 16, 32): one warp = 32/N transform units, 4 warps per CTA, no block-level synchronisation at all.  A lane
// owns one row of its unit from the first load to the last store: residual = cur - pred (vector loads, kept as packed
// words) -> first forward stage in registers -> second stage -> quant (+ sign hiding) -> if any level: first inverse stage with
// the dequantisation folded into its column loads -> second inverse stage into registers -> the lane's part of SSD(resid,
// decoded resid) and SSD(resid, 0), butterfly over the unit's lanes, the reference's zero-out test in IEEE double without
// contraction -> reconstructed row (vector store), levels (int16) and {sum, ssd, ssd_zero, zeroed}.
// k_tq4: 4x4 units, one THREAD per unit, everything in registers.  k_tq_intra<N>: the intra chain (DST for 4x4 luma).
// HBM traffic per sample: 2 B read (cur, pred) + 1 B recon + 2 B levels; everything else stays on chip.
#include "hb_shim.h"
#include "hb_tq_core.cuh"
#include "hb_intra_core.cuh"

namespace {

constexpr int kWarpsPerCta = 4;

template <int N>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_tq(const hbd_tq_args a)
{
    using Q = HbTq<N>;
    constexpr int TPW = Q::TPW, NW = N / 4;
    __shared__ __align__(16) int16_t smem[kWarpsPerCta][4][Q::ELEMS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first_job = (blockIdx.x * kWarpsPerCta + warp) * TPW;
    if (first_job >= a.n_jobs) return;
    const double thr_k = a.dyn ? a.dyn->thr_k : a.thr_k;
    int16_t *T = smem[warp][0], *C = smem[warp][1], *L = smem[warp][2], *U = smem[warp][3];

    // ---- a lane owns stack row `lane` = row lane%N of unit lane/N from the first load to the last store: its current and
    // predicted samples stay in registers as packed words (the residual never goes through shared memory), it computes its
    // part of both SSDs from the decoded row the last inverse stage leaves in registers, and writes its reconstructed row.
    const int unit = lane / N, row = lane % N;
    const bool valid = first_job + unit < a.n_jobs;
    const int2 xy = __ldg(reinterpret_cast<const int2 *>(a.jobs_xy) + min(first_job + unit, a.n_jobs - 1));
    uint32_t ow[NW], pw[NW];
    {
        const uint8_t *co = a.cur.org + (xy.y + row) * a.cur.pitch + xy.x, *po = a.pred.org + (xy.y + row) * a.pred.pitch + xy.x;
        if constexpr (N == 8) {
            const uint2 o = *reinterpret_cast<const uint2 *>(co), p = *reinterpret_cast<const uint2 *>(po);
            ow[0] = o.x; ow[1] = o.y; pw[0] = p.x; pw[1] = p.y;
        } else {
#pragma unroll
            for (int q = 0; q < NW / 4; q++) {
                const uint4 o = reinterpret_cast<const uint4 *>(co)[q], p = reinterpret_cast<const uint4 *>(po)[q];
                ow[4 * q] = o.x; ow[4 * q + 1] = o.y; ow[4 * q + 2] = o.z; ow[4 * q + 3] = o.w;
                pw[4 * q] = p.x; pw[4 * q + 1] = p.y; pw[4 * q + 2] = p.z; pw[4 * q + 3] = p.w;
            }
        }
    }
    auto resid = [&](int j) -> int { return static_cast<int>((ow[j >> 2] >> (8 * (j & 3))) & 255u) - static_cast<int>((pw[j >> 2] >> (8 * (j & 3))) & 255u); };
    {
        int x[N];
#pragma unroll
        for (int j = 0; j < N; j++) x[j] = resid(j);
        Q::fwd_stage_regs(x, T, lane, Q::LOG2 - 1);
    }
    __syncwarp();
    Q::template fwd_stage<false>(T, C, lane, Q::LOG2 + 6);
    __syncwarp();

    int unit_sum[TPW];
    Q::quantise(C, L, U, a.qtab, a.qbits, a.add, lane, unit_sum);
    if (a.sign_hiding) Q::sign_hide(L, C, U, a.scan, lane, unit_sum);

    bool any = false;
    int my_sum = 0;
#pragma unroll
    for (int u = 0; u < TPW; u++) { any |= unit_sum[u] > 0; if (u == unit) my_sum = unit_sum[u]; }
    int dec[N];                                           // decoded residual of this lane's row
#pragma unroll
    for (int j = 0; j < N; j++) dec[j] = 0;
    if (any) {                                            // warp-uniform
        Q::inv_stage1_dequant(L, a.dqtab, a.per, T, lane);     // dequantisation folded into the column loads
        __syncwarp();
        Q::inv_stage2_regs(T, dec, lane);
    }

    // ---- per-unit SSDs: row sums, then a butterfly over the N lanes of the unit
    uint32_t z = 0, sd = 0;
#pragma unroll
    for (int j = 0; j < N; j++) {
        const int r = resid(j), e = r - dec[j];
        z += static_cast<uint32_t>(r) * static_cast<uint32_t>(r);
        sd += static_cast<uint32_t>(e) * static_cast<uint32_t>(e);
    }
    if constexpr (N == 32) { z = __reduce_add_sync(HB_FULL_MASK, z); sd = __reduce_add_sync(HB_FULL_MASK, sd); }
    else {
#pragma unroll
        for (int d = N / 2; d > 0; d >>= 1) { z += __shfl_xor_sync(HB_FULL_MASK, z, d); sd += __shfl_xor_sync(HB_FULL_MASK, sd, d); }
    }

    // ---- decision (hmr_motion_inter.c:90-121 / :186-224), identical in every lane of a unit
    bool keep = false;
    {
        hb_tu_result r;
        r.sum = my_sum; r.zeroed = 0; r.ssd_zero = 0;
        uint32_t zw = z, dw = sd;
        if (!a.is_luma) {
            zw = __double2uint_rz(__dmul_rn(a.weight, static_cast<double>(zw)));
            dw = __double2uint_rz(__dmul_rn(a.weight, static_cast<double>(dw)));
        }
        if (my_sum > 0) {
            const double lhs = static_cast<double>(zw);
            const double base = a.is_luma ? static_cast<double>(static_cast<int32_t>(dw)) : static_cast<double>(dw);
            const double rhs = __dadd_rn(base, __dmul_rn(thr_k, static_cast<double>(my_sum)));
            r.ssd = dw; r.ssd_zero = zw;
            if (lhs <= rhs) { r.zeroed = 1; r.sum = 0; }
            else keep = true;
        } else {
            r.ssd = zw;                                   // ssd16b(residual, zeros)
        }
        if (valid && row == 0) a.res_out[first_job + unit] = r;
    }
    const uint32_t keep_mask = __ballot_sync(HB_FULL_MASK, keep);          // bit unit*N: verdict of that unit

    // ---- reconstruction of this lane's row
    if (valid) {
        uint32_t out[NW];
#pragma unroll
        for (int q = 0; q < NW; q++) {
            out[q] = pw[q];
            if (keep) out[q] = hb_pack_sat_u8x4(static_cast<int>(pw[q] & 255u) + dec[4 * q], static_cast<int>((pw[q] >> 8) & 255u) + dec[4 * q + 1],
                                                static_cast<int>((pw[q] >> 16) & 255u) + dec[4 * q + 2], static_cast<int>(pw[q] >> 24) + dec[4 * q + 3]);
        }
        uint8_t *ro = a.rec.org + (xy.y + row) * a.rec.pitch + xy.x;
        if constexpr (N == 8) *reinterpret_cast<uint2 *>(ro) = make_uint2(out[0], out[1]);
        else {
#pragma unroll
            for (int q = 0; q < NW / 4; q++) reinterpret_cast<uint4 *>(ro)[q] = make_uint4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
        }
    }
    // ---- levels (int16x4 stores), zeros for a unit that was zeroed out
#pragma unroll
    for (int it = 0; it < Q::ITERS4; it++) {
        const typename Q::G4 g = Q::group4(it, lane);
        if (first_job + g.unit >= a.n_jobs) continue;
        int lv[4] = { 0, 0, 0, 0 };
        if ((keep_mask >> (g.unit * N)) & 1u) Q::ld4(L + g.off, lv);
        uint2 w;
        w.x = (static_cast<uint32_t>(lv[0]) & 0xffffu) | (static_cast<uint32_t>(lv[1]) << 16);
        w.y = (static_cast<uint32_t>(lv[2]) & 0xffffu) | (static_cast<uint32_t>(lv[3]) << 16);
        *reinterpret_cast<uint2 *>(a.coeff_out + static_cast<size_t>(first_job + g.unit) * (N * N) + g.pos4) = w;
    }
}

// ------------------------------------------------------------------ 4x4 units, inter chain: one THREAD per unit
// A 4x4 unit is 16 samples: the whole chain fits in registers, so nothing is shared, shuffled or synchronised, and every lane
// of the warp carries a unit (the warp-stack kernel above spends most of its instructions on per-phase overhead at this size).
// Same arithmetic, same results: stage shifts 1 / 8 forward, 7 / 12 inverse, truncation / saturation as in HbTq<4>.
// Only sign-data hiding indexes coefficients by a run-time scan position; it reads them from a small element-major shared
// array (s_l / s_u, [position][thread]) that is written just before it -- the port of HbTq::sign_hide for a single group.
constexpr int kTq4Threads = 128;

__global__ void __launch_bounds__(kTq4Threads) k_tq4(const hbd_tq_args a)
{
    __shared__ int16_t s_l[16][kTq4Threads], s_u[16][kTq4Threads];
    const int tid = threadIdx.x;
    const int job = blockIdx.x * kTq4Threads + tid;
    if (job >= a.n_jobs) return;
    const int2 xy = __ldg(reinterpret_cast<const int2 *>(a.jobs_xy) + job);
    const double thr_k = a.dyn ? a.dyn->thr_k : a.thr_k;

    // ---- residual
    int r[16];
    uint32_t pw[4];
#pragma unroll
    for (int row = 0; row < 4; row++) {
        const uint32_t o = *reinterpret_cast<const uint32_t *>(a.cur.org + (xy.y + row) * a.cur.pitch + xy.x);
        pw[row] = *reinterpret_cast<const uint32_t *>(a.pred.org + (xy.y + row) * a.pred.pitch + xy.x);
#pragma unroll
        for (int k = 0; k < 4; k++) r[4 * row + k] = static_cast<int>((o >> (8 * k)) & 255u) - static_cast<int>((pw[row] >> (8 * k)) & 255u);
    }
    // ---- forward: t[k][row] = (T x_row)[k], c[k'][k] = (T t_k)[k']
    int t[16], c[16];
#pragma unroll
    for (int row = 0; row < 4; row++) {
        const int x[4] = { r[4 * row], r[4 * row + 1], r[4 * row + 2], r[4 * row + 3] };
        int y[4];
        hb_fwd1d<4>(x, y);
#pragma unroll
        for (int k = 0; k < 4; k++) t[4 * k + row] = static_cast<int16_t>((y[k] + 1) >> 1);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x[4] = { t[4 * k], t[4 * k + 1], t[4 * k + 2], t[4 * k + 3] };
        int y[4];
        hb_fwd1d<4>(x, y);
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) c[4 * k2 + k] = static_cast<int16_t>((y[k2] + 128) >> 8);
    }
    // ---- quantise (hmr_sse42_functions_quant.c:52-118)
    int lv[16], du[16], sum = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const int4 q = __ldg(reinterpret_cast<const int4 *>(a.qtab) + g);
        const int qq[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int p = 4 * g + k;
            const uint32_t ab = static_cast<uint32_t>(c[p] < 0 ? -c[p] : c[p]);
            const uint32_t prod = ab * static_cast<uint32_t>(qq[k]);
            const int level = static_cast<int32_t>(prod + static_cast<uint32_t>(a.add)) >> a.qbits;
            const int delta = static_cast<int32_t>(prod - (static_cast<uint32_t>(level) << a.qbits)) >> (a.qbits - 8);
            const int sat = hb_sat16(level);
            lv[p] = c[p] > 0 ? sat : (c[p] < 0 ? -sat : 0);
            du[p] = hb_sat16(delta);
            sum += level;
        }
    }
    // ---- sign-data hiding (hmr_quant.c:61), one coefficient group
    if (a.sign_hiding && sum >= 2) {
        uint32_t neg_c = 0;                                 // bit p: coefficient p is negative
#pragma unroll
        for (int p = 0; p < 16; p++) { s_l[p][tid] = static_cast<int16_t>(lv[p]); s_u[p][tid] = static_cast<int16_t>(du[p]); neg_c |= c[p] < 0 ? (1u << p) : 0u; }
        const uint16_t *sc = a.scan;
        int first = 16, last = -1, asum = 0;
        for (int i = 0; i < 16; i++) {
            const int v = s_l[__ldg(sc + i)][tid];
            if (v) { if (first == 16) first = i; last = i; }
        }
        if (last - first >= 4) {
            for (int i = first; i <= last; i++) asum += s_l[__ldg(sc + i)][tid];
            const unsigned sign_bit = s_l[__ldg(sc + first)][tid] > 0 ? 0u : 1u;
            if (sign_bit != static_cast<unsigned>(asum & 1)) {
                int min_cost = 0x7fffffff, min_pos = -1, final_change = 0, cur_cost = 0x7fffffff, cur_change = 0;
                for (int i = last; i >= 0; i--) {
                    const int p = __ldg(sc + i);
                    const int l = s_l[p][tid], d = s_u[p][tid];
                    if (l != 0) {
                        if (d > 0) { cur_cost = -d; cur_change = 1; }
                        else if (i == first && (l == 1 || l == -1)) cur_cost = 0x7fffffff;
                        else { cur_cost = d; cur_change = -1; }
                    } else if (i < first) {
                        const unsigned this_sign = (neg_c >> p) & 1u;
                        if (this_sign != sign_bit) cur_cost = 0x7fffffff;
                        else { cur_cost = -d; cur_change = 1; }
                    } else {
                        cur_cost = -d; cur_change = 1;
                    }
                    if (cur_cost < min_cost) { min_cost = cur_cost; final_change = cur_change; min_pos = p; }
                }
                const int l = s_l[min_pos][tid];
                if (l == 32767 || l == -32768) final_change = -1;
                const int nv = static_cast<int16_t>(((neg_c >> min_pos) & 1u) ? l - final_change : l + final_change);
#pragma unroll
                for (int p = 0; p < 16; p++) if (p == min_pos) lv[p] = nv;
            }
        }
    }
    // ---- dequantise + inverse when anything is left
    int dec[16];
#pragma unroll
    for (int p = 0; p < 16; p++) dec[p] = 0;
    if (sum > 0) {
        int d[16];
        const int iq_shift = 2 + 3;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int4 q = __ldg(reinterpret_cast<const int4 *>(a.dqtab) + g);
            const int qq[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t prod = static_cast<uint32_t>(lv[4 * g + k]) * static_cast<uint32_t>(qq[k]);
                int v;
                if (iq_shift > a.per) v = static_cast<int32_t>(prod + (1u << (iq_shift - a.per - 1))) >> (iq_shift - a.per);
                else v = static_cast<int32_t>(prod << (a.per - iq_shift));
                d[4 * g + k] = hb_sat16(v);
            }
        }
        int t2[16];                                         // t2[col][j] = sum_k coef[k][j] d[k][col]
#pragma unroll
        for (int col = 0; col < 4; col++) {
            const int y[4] = { d[col], d[4 + col], d[8 + col], d[12 + col] };
            int x[4];
            hb_inv1d<4>(y, x);
#pragma unroll
            for (int j = 0; j < 4; j++) t2[4 * col + j] = hb_sat16((x[j] + 64) >> 7);
        }
#pragma unroll
        for (int row = 0; row < 4; row++) {
            const int y[4] = { t2[row], t2[4 + row], t2[8 + row], t2[12 + row] };
            int x[4];
            hb_inv1d<4>(y, x);
#pragma unroll
            for (int j = 0; j < 4; j++) dec[4 * row + j] = hb_sat16((x[j] + 2048) >> 12);
        }
    }
    // ---- SSDs and the zero-out decision (hmr_motion_inter.c:90-121 / :186-224)
    uint32_t z = 0, sd = 0;
#pragma unroll
    for (int p = 0; p < 16; p++) {
        const int e = r[p] - dec[p];
        z += static_cast<uint32_t>(r[p]) * static_cast<uint32_t>(r[p]);
        sd += static_cast<uint32_t>(e) * static_cast<uint32_t>(e);
    }
    hb_tu_result res;
    res.sum = sum; res.zeroed = 0; res.ssd_zero = 0;
    uint32_t zw = z, dw = sd;
    if (!a.is_luma) {
        zw = __double2uint_rz(__dmul_rn(a.weight, static_cast<double>(zw)));
        dw = __double2uint_rz(__dmul_rn(a.weight, static_cast<double>(dw)));
    }
    bool keep = false;
    if (sum > 0) {
        const double lhs = static_cast<double>(zw);
        const double base = a.is_luma ? static_cast<double>(static_cast<int32_t>(dw)) : static_cast<double>(dw);
        const double rhs = __dadd_rn(base, __dmul_rn(thr_k, static_cast<double>(sum)));
        res.ssd = dw; res.ssd_zero = zw;
        if (lhs <= rhs) { res.zeroed = 1; res.sum = 0; }
        else keep = true;
    } else {
        res.ssd = zw;
    }
    a.res_out[job] = res;
    // ---- outputs
    uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
    if (keep) {
        auto pk = [&](int i) { return (static_cast<uint32_t>(lv[i]) & 0xffffu) | (static_cast<uint32_t>(lv[i + 1]) << 16); };
        w0 = make_uint4(pk(0), pk(2), pk(4), pk(6)); w1 = make_uint4(pk(8), pk(10), pk(12), pk(14));
    }
    uint4 *co = reinterpret_cast<uint4 *>(a.coeff_out + static_cast<size_t>(job) * 16);
    co[0] = w0; co[1] = w1;
#pragma unroll
    for (int row = 0; row < 4; row++) {
        uint32_t out = pw[row];
        if (keep) out = hb_pack_sat_u8x4(static_cast<int>(pw[row] & 255u) + dec[4 * row], static_cast<int>((pw[row] >> 8) & 255u) + dec[4 * row + 1],
                                         static_cast<int>((pw[row] >> 16) & 255u) + dec[4 * row + 2], static_cast<int>(pw[row] >> 24) + dec[4 * row + 3]);
        *reinterpret_cast<uint32_t *>(a.rec.org + (xy.y + row) * a.rec.pitch + xy.x) = out;
    }
}

// ------------------------------------------------------------------ intra T/Q chain after prediction
// encode_intra_cu (hmr_motion_intra.c:1023-1069) and the chroma loop of hmr_motion_intra_chroma.c:340-365: residual ->
// DST-VII (4x4 luma) or DCT -> quant (intra lists, scan from the intra mode, sign hiding) -> if any level: dequant -> inverse
// -> reconstruction; distortion = SSD(original, reconstruction) (chroma weighted).  No zero-out heuristic on this path.
// one warp's stack of 32 / N units starting at job `first_job`; sm: 5 * HbTq<N>::ELEMS int16 of the warp's shared memory
template <int N>
__device__ __forceinline__ void tq_intra_warp(const hbd_tq_args &a, const int first_job, int16_t *sm, const int lane)
{
    using Q = HbTq<N>;
    constexpr int TPW = Q::TPW;
    int16_t *X = sm, *T = sm + Q::ELEMS, *C = sm + 2 * Q::ELEMS, *L = sm + 3 * Q::ELEMS, *U = sm + 4 * Q::ELEMS;
    const bool dst = (N == 4) && a.is_luma;
    int jx[TPW], jy[TPW];
#pragma unroll
    for (int u = 0; u < TPW; u++) {
        const int j = min(first_job + u, a.n_jobs - 1);
        jx[u] = __ldg(a.jobs_xy + 2 * j);
        jy[u] = __ldg(a.jobs_xy + 2 * j + 1);
    }
    auto unit_xy = [&](int unit, int &x, int &y) {
        x = jx[0]; y = jy[0];
#pragma unroll
        for (int u = 1; u < TPW; u++) if (u == unit) { x = jx[u]; y = jy[u]; }
    };
#pragma unroll
    for (int it = 0; it < Q::ITERS4; it++) {
        const typename Q::G4 g = Q::group4(it, lane);
        int x, y;
        unit_xy(g.unit, x, y);
        const int r = g.row % N;
        const uint32_t o = *reinterpret_cast<const uint32_t *>(a.cur.org + (y + r) * a.cur.pitch + x + g.col);
        const uint32_t p = *reinterpret_cast<const uint32_t *>(a.pred.org + (y + r) * a.pred.pitch + x + g.col);
        int d[4];
#pragma unroll
        for (int k = 0; k < 4; k++) d[k] = static_cast<int>((o >> (8 * k)) & 255u) - static_cast<int>((p >> (8 * k)) & 255u);
        Q::st4(X + g.off, d);
    }
    __syncwarp();
    if (dst) Q::template forward<true>(X, T, C, lane); else Q::template forward<false>(X, T, C, lane);
    int unit_sum[TPW];
    Q::quantise(C, L, U, a.qtab, a.qbits, a.add, lane, unit_sum);
    if (a.sign_hiding) Q::sign_hide(L, C, U, a.scan, lane, unit_sum);
    bool any = false;
#pragma unroll
    for (int u = 0; u < TPW; u++) any |= unit_sum[u] != 0;
    if (any) {
        Q::dequantise(L, T, a.dqtab, a.per, lane);
        if (dst) Q::template inverse<true>(T, U, C, lane); else Q::template inverse<false>(T, U, C, lane);
    }
    uint32_t ssd[TPW];
#pragma unroll
    for (int u = 0; u < TPW; u++) ssd[u] = 0;
#pragma unroll
    for (int it = 0; it < Q::ITERS4; it++) {
        const typename Q::G4 g = Q::group4(it, lane);
        int x, y;
        unit_xy(g.unit, x, y);
        int usum = 0;
#pragma unroll
        for (int u = 0; u < TPW; u++) if (u == g.unit) usum = unit_sum[u];
        const bool valid = first_job + g.unit < a.n_jobs;
        const int r = g.row % N;
        int lv[4], dr[4] = { 0, 0, 0, 0 };
        Q::ld4(L + g.off, lv);
        if (usum != 0) Q::ld4(C + g.off, dr);
        const uint32_t o = *reinterpret_cast<const uint32_t *>(a.cur.org + (y + r) * a.cur.pitch + x + g.col);
        const uint32_t p = *reinterpret_cast<const uint32_t *>(a.pred.org + (y + r) * a.pred.pitch + x + g.col);
        uint32_t out = 0, s = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int rec = hb_clip255(static_cast<int>((p >> (8 * q)) & 255u) + dr[q]);
            const int e = static_cast<int>((o >> (8 * q)) & 255u) - rec;
            s += static_cast<uint32_t>(e * e);
            out |= static_cast<uint32_t>(rec) << (8 * q);
        }
        Q::template unit_add<uint32_t>(s, it, ssd);
        if (valid) {
            uint2 w;
            w.x = (static_cast<uint32_t>(lv[0]) & 0xffffu) | (static_cast<uint32_t>(lv[1]) << 16);
            w.y = (static_cast<uint32_t>(lv[2]) & 0xffffu) | (static_cast<uint32_t>(lv[3]) << 16);
            *reinterpret_cast<uint2 *>(a.coeff_out + static_cast<size_t>(first_job + g.unit) * (N * N) + g.pos4) = w;
            *reinterpret_cast<uint32_t *>(a.rec.org + (y + r) * a.rec.pitch + x + g.col) = out;
        }
    }
    uint32_t my_ssd = 0; int my_sum = 0;
#pragma unroll
    for (int u = 0; u < TPW; u++) if (u == lane) { my_ssd = ssd[u]; my_sum = unit_sum[u]; }
    if (lane < TPW && first_job + lane < a.n_jobs) {
        hb_tu_result r;
        r.sum = my_sum; r.zeroed = 0; r.ssd_zero = 0;
        r.ssd = a.is_luma ? my_ssd : static_cast<uint32_t>(__double2int_rz(__dmul_rn(a.weight, static_cast<double>(my_ssd))));
        a.res_out[first_job + lane] = r;
    }
}

template <int N>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_tq_intra(const hbd_tq_args a)
{
    using Q = HbTq<N>;
    __shared__ __align__(16) int16_t smem[kWarpsPerCta][5 * Q::ELEMS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first_job = (blockIdx.x * kWarpsPerCta + warp) * Q::TPW;
    if (first_job >= a.n_jobs) return;
    tq_intra_warp<N>(a, first_job, smem[warp], lane);
}

// ------------------------------------------------------------------ a whole intra picture in ONE launch
// hb_intra_reconstruct: the units of a picture sorted by dependency level and, inside a level, by (plane, size, QP, scan); a TASK is up to
// 32 / N consecutive units of one such group -- what one warp stacks into a T/Q pass.  The kernel is persistent: every warp draws the next
// task from a global counter (tasks are numbered level by level), waits until all units of the earlier levels are finished (one global
// count of finished units against the task's prefix; the units it waits for belong to tasks drawn before its own, i.e. to warps that are
// running, so nobody waits for work that is not in flight), gathers the reference samples of its units from the reconstructed picture,
// writes their predictions, runs the intra T/Q chain on them, and adds its units to the count.  One launch instead of ~7 per level.
template <int N>
__device__ __forceinline__ void intra_wave_task(const hbd_wave_args &w, const hbd_wave_task &t, int16_t *sm, int16_t *adi, const int lane)
{
    const hbd_wave_group g = w.groups[t.group];
    for (int u = 0; u < t.n_units; u++) {
        const hbd_wave_unit un = w.units[t.first_unit + u];
        int16_t *raw = adi, *flt = adi + 132;
        intra_gather_adi(hbd_pick_plane(w.rec, g.comp), un.x, un.y, N, un.flags, un.lbs, un.trs, raw, lane);
        __syncwarp();
        intra_predict_block(hbd_pick_plane(w.pred, g.comp), g.comp, un.x, un.y, N, un.mode, g.comp ? 0 : -1, raw, flt, lane);
        __syncwarp();
    }
    __threadfence_block();
    __syncwarp();                                            // the warp's own predictions are visible to all its lanes
    hbd_tq_args a;
    a.cur = hbd_pick_plane(w.cur, g.comp); a.pred = hbd_pick_plane(w.pred, g.comp); a.rec = hbd_pick_plane(w.rec, g.comp);
    a.jobs_xy = w.xy + 2 * t.first_unit; a.n_jobs = t.n_units; a.n = N;
    a.qtab = g.qtab; a.dqtab = g.dqtab; a.scan = g.scan; a.qbits = g.qbits; a.add = g.add; a.per = g.per;
    a.sign_hiding = w.sign_hiding; a.is_luma = g.comp == 0; a.intra = 1; a.thr_k = 0.; a.weight = g.weight; a.dyn = nullptr;
    a.coeff_out = w.coeff + t.coeff_off; a.res_out = w.res + t.first_unit;
    tq_intra_warp<N>(a, 0, sm, lane);
}

__global__ void __launch_bounds__(kWarpsPerCta * 32) k_intra_wave(const hbd_wave_args w)
{
    __shared__ __align__(16) int16_t smem[kWarpsPerCta][5 * HbTq<32>::ELEMS];
    __shared__ int16_t s_adi[kWarpsPerCta][2 * 132];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (;;) {
        unsigned ti = 0;
        if (lane == 0) ti = atomicAdd(&w.counters[0], 1u);
        ti = __shfl_sync(HB_FULL_MASK, ti, 0);
        if (ti >= static_cast<unsigned>(w.n_tasks)) return;
        const hbd_wave_task t = w.tasks[ti];
        if (lane == 0) {
            const volatile unsigned *done = w.counters + 1;
            while (*done < static_cast<unsigned>(t.units_before)) __nanosleep(64);
        }
        __syncwarp();
        __threadfence();                                     // what the earlier levels wrote is read after this point
        switch (t.size) {
        case 4: intra_wave_task<4>(w, t, smem[warp], s_adi[warp], lane); break;
        case 8: intra_wave_task<8>(w, t, smem[warp], s_adi[warp], lane); break;
        case 16: intra_wave_task<16>(w, t, smem[warp], s_adi[warp], lane); break;
        default: intra_wave_task<32>(w, t, smem[warp], s_adi[warp], lane); break;
        }
        __threadfence();                                     // reconstruction, levels and records before the count moves
        __syncwarp();
        if (lane == 0) atomicAdd(&w.counters[1], static_cast<unsigned>(t.n_units));
    }
}

// ------------------------------------------------------------------ per-call kernels (one warp, one unit)
template <int N, bool DST>
__global__ void __launch_bounds__(32) k_pc_transform(const int16_t *block, int bs, int16_t *coeff)
{
    using Q = HbTq<N>;
    __shared__ __align__(16) int16_t sm[3][Q::ELEMS];
    const int lane = threadIdx.x;
    for (int i = lane; i < Q::ELEMS; i += 32) sm[0][i] = 0;
    __syncwarp();
    for (int e = lane; e < N * N; e += 32) sm[0][(e / N) * Q::S + e % N] = block[(e / N) * bs + e % N];
    __syncwarp();
    Q::template forward<DST>(sm[0], sm[1], sm[2], lane);
    for (int e = lane; e < N * N; e += 32) coeff[e] = sm[2][(e / N) * Q::S + e % N];
}

template <int N, bool DST>
__global__ void __launch_bounds__(32) k_pc_itransform(int16_t *block, int bs, const int16_t *coeff)
{
    using Q = HbTq<N>;
    __shared__ __align__(16) int16_t sm[3][Q::ELEMS];
    const int lane = threadIdx.x;
    for (int i = lane; i < Q::ELEMS; i += 32) sm[0][i] = 0;
    __syncwarp();
    for (int e = lane; e < N * N; e += 32) sm[0][(e / N) * Q::S + e % N] = coeff[e];
    __syncwarp();
    Q::template inverse<DST>(sm[0], sm[1], sm[2], lane);
    for (int e = lane; e < N * N; e += 32) block[(e / N) * bs + e % N] = sm[2][(e / N) * Q::S + e % N];
}

template <int N>
__global__ void __launch_bounds__(32) k_pc_quant(const int16_t *src, int16_t *dst, int16_t *delta_u, int32_t *sum,
                                                 const int32_t *qtab, const uint16_t *scan, int qbits, int add, int sign_hiding)
{
    using Q = HbTq<N>;
    __shared__ __align__(16) int16_t sm[3][Q::ELEMS];
    const int lane = threadIdx.x;
    for (int i = lane; i < Q::ELEMS; i += 32) sm[0][i] = 0;
    __syncwarp();
    for (int e = lane; e < N * N; e += 32) sm[0][(e / N) * Q::S + e % N] = src[e];
    __syncwarp();
    int unit_sum[Q::TPW];
    Q::quantise(sm[0], sm[1], sm[2], qtab, qbits, add, lane, unit_sum);
    if (sign_hiding) Q::sign_hide(sm[1], sm[0], sm[2], scan, lane, unit_sum);
    for (int e = lane; e < N * N; e += 32) {
        dst[e] = sm[1][(e / N) * Q::S + e % N];
        if (delta_u) delta_u[e] = sm[2][(e / N) * Q::S + e % N];
    }
    if (lane == 0) *sum = unit_sum[0];
}

template <int N>
__global__ void __launch_bounds__(32) k_pc_inv_quant(const int16_t *src, int16_t *dst, const int32_t *dqtab, int per)
{
    using Q = HbTq<N>;
    __shared__ __align__(16) int16_t sm[2][Q::ELEMS];
    const int lane = threadIdx.x;
    for (int i = lane; i < Q::ELEMS; i += 32) sm[0][i] = 0;
    __syncwarp();
    for (int e = lane; e < N * N; e += 32) sm[0][(e / N) * Q::S + e % N] = src[e];
    __syncwarp();
    Q::dequantise(sm[0], sm[1], dqtab, per, lane);
    for (int e = lane; e < N * N; e += 32) dst[e] = sm[1][(e / N) * Q::S + e % N];
}

template <int N> int launch_tq(const hbd_tq_args *a, cudaStream_t s)
{
    const int per_cta = kWarpsPerCta * HbTq<N>::TPW;
    const int grid = (a->n_jobs + per_cta - 1) / per_cta;
    if (a->intra) k_tq_intra<N><<<grid, kWarpsPerCta * 32, 0, s>>>(*a);
    else if (N == 4) k_tq4<<<(a->n_jobs + kTq4Threads - 1) / kTq4Threads, kTq4Threads, 0, s>>>(*a);
    else k_tq<N><<<grid, kWarpsPerCta * 32, 0, s>>>(*a);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace

extern "C" int hbk_intra_wave(const hbd_wave_args *w, int ctas, void *stream)
{
    if (w->n_tasks <= 0) return 0;
    k_intra_wave<<<ctas, kWarpsPerCta * 32, 0, static_cast<cudaStream_t>(stream)>>>(*w);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_tq_encode(const hbd_tq_args *a, void *stream)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (a->n_jobs <= 0) return 0;
    switch (a->n) {
    case 4: return launch_tq<4>(a, s);
    case 8: return launch_tq<8>(a, s);
    case 16: return launch_tq<16>(a, s);
    case 32: return launch_tq<32>(a, s);
    default: return static_cast<int>(cudaErrorInvalidValue);
    }
}

#define HB_DISPATCH_N(n, ...)                                 \
    switch (n) {                                              \
    case 4: { constexpr int N = 4; __VA_ARGS__; } break;      \
    case 8: { constexpr int N = 8; __VA_ARGS__; } break;      \
    case 16: { constexpr int N = 16; __VA_ARGS__; } break;    \
    case 32: { constexpr int N = 32; __VA_ARGS__; } break;    \
    default: return static_cast<int>(cudaErrorInvalidValue); }

extern "C" int hbk_pc_transform(const int16_t *block, int bs, int16_t *coeff, int n, int dst4, void *stream)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n == 4 && dst4) k_pc_transform<4, true><<<1, 32, 0, s>>>(block, bs, coeff);
    else HB_DISPATCH_N(n, k_pc_transform<N, false><<<1, 32, 0, s>>>(block, bs, coeff))
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_pc_itransform(int16_t *block, int bs, const int16_t *coeff, int n, int dst4, void *stream)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n == 4 && dst4) k_pc_itransform<4, true><<<1, 32, 0, s>>>(block, bs, coeff);
    else HB_DISPATCH_N(n, k_pc_itransform<N, false><<<1, 32, 0, s>>>(block, bs, coeff))
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_pc_quant(const int16_t *src, int16_t *dst, int16_t *delta_u, int32_t *sum, int n, const int32_t *qtab,
                            const uint16_t *scan, int qbits, int add, int sign_hiding, void *stream)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HB_DISPATCH_N(n, k_pc_quant<N><<<1, 32, 0, s>>>(src, dst, delta_u, sum, qtab, scan, qbits, add, sign_hiding))
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_pc_inv_quant(const int16_t *src, int16_t *dst, int n, const int32_t *dqtab, int per, void *stream)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HB_DISPATCH_N(n, k_pc_inv_quant<N><<<1, 32, 0, s>>>(src, dst, dqtab, per))
    return static_cast<int>(cudaGetLastError());
}
