// hb_kernels_intra.cu -- intra prediction: planar (hmr_motion_intra.c:408), DC and the 33 angular
// modes with their edge filters (:482), reference-sample smoothing (adi_filter, :189) and the filtered-or-not rule of the
// mode search (:1122), as closed forms per sample -- no intermediate refMain/refSide arrays, no transposition pass:
//   planar   ((l << s) + N + (x+1)(tr - l) + (t << s) + (y+1)(lb - t)) >> (s+1)
//   angular  pos = (j+1)*angle, k = i + (pos >> 5) + 1, f = pos & 31, sample = f ? ((32-f) R(k) + f R(k+1) + 16) >> 5 : R(k)
//            with R(k) = main[k] for k >= 0 and side[(128 - k*inv_angle) >> 8] for the projected part (k < 0);
//            vertical modes read (j,i) = (y,x), horizontal modes (x,y).
// One warp per job.  mode >= 0: the prediction is written to the prediction plane (input of the intra T/Q chain);
// mode < 0: the SADs of all 35 luma modes against the current block (what the host's mode search probes).
#include "hb_shim.h"
#include "hb_dev_common.cuh"
#include "hb_intra_core.cuh"

namespace {

constexpr int kIntraWarps = 4;

__global__ void __launch_bounds__(kIntraWarps * 32) k_intra(const hbd_intra_args a)
{
    __shared__ int16_t s_adi[kIntraWarps][2][132];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ji = blockIdx.x * kIntraWarps + warp;
    if (ji >= a.n_jobs) return;
    const hbd_intra_job job = a.jobs[ji];
    if (job.mode < 0) return;                          // SAD form: k_intra_sads
    const int n = job.size;
    int lg = 2;
    while ((1 << lg) < n) lg++;
    int16_t *raw = s_adi[warp][0], *flt = s_adi[warp][1];
    for (int i = lane; i < 4 * n + 1; i += 32) raw[i] = a.adi[job.adi_off + i];
    __syncwarp();
    const bool is_luma = job.comp == 0;
    if (is_luma) intra_filter_adi(raw, flt, n, lg, lane);
    __syncwarp();
    const bool edge = is_luma && n <= 16;
    if (job.mode >= 0) {
        // ---- one prediction into the plane.  The caller decides which reference array a chroma block uses (always the raw
        // one in the reference); luma follows the search rule unless the job says otherwise
        const bool use_flt = is_luma && (job.filtered < 0 ? intra_uses_filtered(lg, job.mode) : job.filtered != 0);
        const int16_t *mid = (use_flt ? flt : raw) + 2 * n;
        const IntraMode m = intra_mode_info(job.mode);
        const int dc = m.kind == 1 ? intra_dc(mid, n, lane) : 0;
        const hbd_plane p = a.pred.p[job.comp];
        for (int e = lane; e < n * n; e += 32) {
            const int x = e % n, y = e / n;
            p.org[(job.y + y) * p.pitch + job.x + x] = static_cast<uint8_t>(intra_sample(mid, n, lg, m, dc, edge, x, y));
        }
        return;
    }
}

// ---- SAD form, one launch per block size: the SADs of all 35 luma modes against the current block (what the mode search of
// homer_loop1_motion_intra, hmr_motion_intra.c:1084, probes one call at a time).  `idx` lists the jobs of this size; a 4x4
// block takes 16 lanes, so two of them share a warp.
// The reference samples are laid out once per job as four arrays indexed -N..2N+1 from the corner: {top, left} x {raw, smoothed}.
// A mode's main array is the top one (vertical modes) or the left one (horizontal modes, evaluated on the transposed block, which
// turns them into the vertical formula); for negative angles its slots -1..(N*angle)>>5 are refilled per mode with the projected
// samples of the other array, so the inner loop is two neighbouring loads and one interpolation per sample with no branches.
template <int N> struct IntraSadCfg {
    static constexpr int LJ = (N * N < 32) ? N * N : 32;      // lanes per job
    static constexpr int JPW = 32 / LJ;                       // jobs per warp
    static constexpr int SPL = N * N / LJ;                    // samples per lane
    static constexpr int LG = (N == 4) ? 2 : (N == 8) ? 3 : (N == 16) ? 4 : 5;
    static constexpr int EXT = 3 * N + 4;                     // -N .. 2N+1, padded to an even count
};

// modes of one orientation on the lane's samples `cur` (block transposed when HOR); main/side arrays point at their corner element
template <int N, bool HOR>
__device__ __forceinline__ void intra_sads_pass(int16_t (*arr)[IntraSadCfg<N>::EXT], const int (&cur)[IntraSadCfg<N>::SPL], int sl, int sub, bool valid,
                                                uint32_t *out)
{
    using C = IntraSadCfg<N>;
    const int first = HOR ? 2 : 18, last = HOR ? 17 : 34;
    for (int mode = first; mode <= last; mode++) {
        const int a = HOR ? 10 - mode : mode - 26;
        const bool flt = intra_uses_filtered(C::LG, mode);
        int16_t *mainp = arr[(HOR ? 2 : 0) + flt] + N;
        const int16_t *side = arr[(HOR ? 0 : 2) + flt] + N;
        uint32_t acc = 0;
        if (a == 0) {                                          // pure vertical / horizontal with the first line smoothed on small blocks
            const int corner = mainp[0];
#pragma unroll
            for (int q = 0; q < C::SPL; q++) {
                const int e = sl + q * C::LJ, i = e % N, j = e / N;
                int v = mainp[i + 1];
                if (N <= 16 && i == 0) v = hb_clip255(v + ((side[j + 1] - corner) >> 1));
                acc = __sad(v, cur[q], acc);
            }
        } else {
            const int aa = abs(a);
            const int angle = a < 0 ? -c_ang[aa] : c_ang[aa];
            if (a < 0) {
                const int inv = c_inv_ang[aa], kmin = (N * angle) >> 5;
                __syncwarp();
                for (int kk = -1 - sl; kk >= kmin; kk -= C::LJ) mainp[kk] = side[(128 - kk * inv) >> 8];
                __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < C::SPL; q++) {
                const int e = sl + q * C::LJ, i = e % N, j = e / N;
                const int pos = (j + 1) * angle, f = pos & 31;
                const int16_t *p = mainp + i + (pos >> 5) + 1;
                const int r0 = p[0], r1 = p[1];
                acc = __sad((32 * r0 + f * (r1 - r0) + 16) >> 5, cur[q], acc);
            }
        }
        if constexpr (C::JPW == 1) {
            acc = __reduce_add_sync(HB_FULL_MASK, acc);
            if (sl == 0) out[mode] = acc;
        } else {
            const uint32_t t0 = __reduce_add_sync(HB_FULL_MASK, sub == 0 ? acc : 0u), t1 = __reduce_add_sync(HB_FULL_MASK, sub == 1 ? acc : 0u);
            if (sl == 0 && valid) out[mode] = sub ? t1 : t0;
        }
    }
}

template <int N>
__global__ void __launch_bounds__(kIntraWarps * 32) k_intra_sads(const hbd_intra_args a, const int32_t *idx, int n_idx)
{
    using C = IntraSadCfg<N>;
    constexpr int LJ = C::LJ, JPW = C::JPW, SPL = C::SPL, LG = C::LG;
    __shared__ int16_t s_lin[kIntraWarps][JPW][4 * N + 4];
    __shared__ int16_t s_arr[kIntraWarps][JPW][4][C::EXT];              // top raw, top smoothed, left raw, left smoothed
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane / LJ, sl = lane % LJ;
    const int k = (blockIdx.x * kIntraWarps + warp) * JPW + sub;
    if ((blockIdx.x * kIntraWarps + warp) * JPW >= n_idx) return;          // whole warp
    const bool valid = k < n_idx;
    const int ji = idx[min(k, n_idx - 1)];
    const hbd_intra_job job = a.jobs[ji];
    int16_t *raw = s_lin[warp][sub];
    int16_t (*arr)[C::EXT] = s_arr[warp][sub];
    for (int i = sl; i < 4 * N + 1; i += LJ) raw[i] = a.adi[job.adi_off + i];
    __syncwarp();
    {   // smoothing (as intra_filter_adi) fused with the split into the four corner-relative arrays
        constexpr int size = 4 * N + 1;
        const int lb = raw[0], lt = raw[2 * N], tr = raw[size - 1];
        const bool strong = N >= 32 && abs(lb + lt - 2 * raw[N]) < 8 && abs(lt + tr - 2 * raw[3 * N]) < 8;
        for (int i = sl; i < size; i += LJ) {
            int v;
            if (i == 0 || i == size - 1) v = raw[i];
            else if (strong) {
                if (i == 2 * N) v = raw[i];
                else if (i < 2 * N) v = ((2 * N - i) * lb + i * lt + N) >> (LG + 1);
                else v = ((4 * N - i) * lt + (i - 2 * N) * tr + N) >> (LG + 1);
            } else v = (raw[i - 1] + 2 * raw[i] + raw[i + 1] + 2) >> 2;
            const int kk = i - 2 * N;
            if (kk >= 0) { arr[0][N + kk] = raw[i]; arr[1][N + kk] = static_cast<int16_t>(v); }
            if (kk <= 0) { arr[2][N - kk] = raw[i]; arr[3][N - kk] = static_cast<int16_t>(v); }
        }
        if (sl < 4) arr[sl][3 * N + 1] = 0;                    // read with a zero weight only
    }
    __syncwarp();
    uint32_t *out = a.sads + static_cast<size_t>(ji) * 35;
    int cur[SPL];
#pragma unroll
    for (int q = 0; q < SPL; q++) {
        const int e = sl + q * LJ;
        cur[q] = a.cur.org[(job.y + e / N) * a.cur.pitch + job.x + e % N];
    }
    {   // planar and DC (symmetric in the two reference arrays): x = i, y = j
        const bool pf = intra_uses_filtered(LG, 0);
        const int16_t *top = arr[pf] + N, *left = arr[2 + pf] + N;
        const int lbv = left[N + 1], trv = top[N + 1];
        uint32_t acc0 = 0;
#pragma unroll
        for (int q = 0; q < SPL; q++) {
            const int e = sl + q * LJ, x = e % N, y = e / N;
            const int l = left[y + 1], t = top[x + 1];
            acc0 = __sad(((l << LG) + N + (x + 1) * (trv - l) + (t << LG) + (y + 1) * (lbv - t)) >> (LG + 1), cur[q], acc0);
        }
        const int16_t *rt = arr[0] + N, *rl = arr[2] + N;       // DC always reads the raw samples
        int dcs = 0;
        for (int i = 1 + sl; i <= N; i += LJ) dcs += rt[i] + rl[i];
#pragma unroll
        for (int d = LJ / 2; d > 0; d >>= 1) dcs += __shfl_xor_sync(HB_FULL_MASK, dcs, d);
        const int dc = (dcs + N) >> (LG + 1);
        uint32_t acc1 = 0;
#pragma unroll
        for (int q = 0; q < SPL; q++) {
            const int e = sl + q * LJ, x = e % N, y = e / N;
            int v = dc;
            if (N <= 16) {
                if (x == 0 && y == 0) v = (rl[1] + rt[1] + 2 * dc + 2) >> 2;
                else if (y == 0) v = (rt[1 + x] + 3 * dc + 2) >> 2;
                else if (x == 0) v = (rl[1 + y] + 3 * dc + 2) >> 2;
            }
            acc1 = __sad(v, cur[q], acc1);
        }
        if constexpr (JPW == 1) {
            acc0 = __reduce_add_sync(HB_FULL_MASK, acc0); acc1 = __reduce_add_sync(HB_FULL_MASK, acc1);
            if (sl == 0) { out[0] = acc0; out[1] = acc1; }
        } else {
            const uint32_t p0 = __reduce_add_sync(HB_FULL_MASK, sub == 0 ? acc0 : 0u), p1 = __reduce_add_sync(HB_FULL_MASK, sub == 1 ? acc0 : 0u);
            const uint32_t d0 = __reduce_add_sync(HB_FULL_MASK, sub == 0 ? acc1 : 0u), d1 = __reduce_add_sync(HB_FULL_MASK, sub == 1 ? acc1 : 0u);
            if (sl == 0 && valid) { out[0] = sub ? p1 : p0; out[1] = sub ? d1 : d0; }
        }
    }
    intra_sads_pass<N, false>(arr, cur, sl, sub, valid, out);
    // horizontal modes: the same formula on the transposed block with the left array as the main one
#pragma unroll
    for (int q = 0; q < SPL; q++) {
        const int e = sl + q * LJ;
        cur[q] = a.cur.org[(job.y + e % N) * a.cur.pitch + job.x + e / N];
    }
    intra_sads_pass<N, true>(arr, cur, sl, sub, valid, out);
}

// per-call form: prediction of one block as int16 into a caller buffer (the table members create_intra_*_prediction)
__global__ void __launch_bounds__(32) k_pc_intra(const int16_t *adi, int n, int mode, int is_luma, int16_t *pred, int stride)
{
    const int lane = threadIdx.x;
    int lg = 2;
    while ((1 << lg) < n) lg++;
    const int16_t *mid = adi + 2 * n;
    const IntraMode m = intra_mode_info(mode);
    const int dc = m.kind == 1 ? intra_dc(mid, n, lane) : 0;
    const bool edge = is_luma && n <= 16;
    for (int e = lane; e < n * n; e += 32) pred[(e / n) * stride + e % n] = static_cast<int16_t>(intra_sample(mid, n, lg, m, dc, edge, e % n, e / n));
}

}  // namespace

extern "C" int hbk_intra(const hbd_intra_args *a, void *stream)
{
    if (a->n_jobs <= 0) return 0;
    k_intra<<<(a->n_jobs + kIntraWarps - 1) / kIntraWarps, kIntraWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(*a);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_intra_sads(const hbd_intra_args *a, int size, const int32_t *idx, int n_idx, void *stream)
{
    if (n_idx <= 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int per_cta = kIntraWarps * (size == 4 ? 2 : 1);
    const int grid = (n_idx + per_cta - 1) / per_cta;
    switch (size) {
    case 4: k_intra_sads<4><<<grid, kIntraWarps * 32, 0, s>>>(*a, idx, n_idx); break;
    case 8: k_intra_sads<8><<<grid, kIntraWarps * 32, 0, s>>>(*a, idx, n_idx); break;
    case 16: k_intra_sads<16><<<grid, kIntraWarps * 32, 0, s>>>(*a, idx, n_idx); break;
    case 32: k_intra_sads<32><<<grid, kIntraWarps * 32, 0, s>>>(*a, idx, n_idx); break;
    default: return static_cast<int>(cudaErrorInvalidValue);
    }
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_pc_intra(const int16_t *adi, int n, int mode, int is_luma, int16_t *pred, int stride, void *stream)
{
    k_pc_intra<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(adi, n, mode, is_luma, pred, stride);
    return static_cast<int>(cudaGetLastError());
}


// ---- reference samples of a batch of intra units from the reconstructed picture (fill_reference_samples, hmr_motion_intra.c:246-406, in the
// closed form of hb_intra_core.cuh): one warp per unit.  adi: index 2n = corner, 2n + 1 + i above, 2n - 1 - r the left column at row y + r.
namespace {
__global__ void __launch_bounds__(128) k_intra_adi(const hbd_frame rec, const hbd_adi_job *jobs, int n_jobs, int16_t *adi_all)
{
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n_jobs) return;
    const hbd_adi_job j = jobs[i];
    intra_gather_adi(hbd_pick_plane(rec, j.comp), j.x, j.y, j.n, j.flags, j.lbs, j.trs, adi_all + j.adi_off, threadIdx.x & 31);
}
}  // namespace

extern "C" int hbk_intra_adi(const hbd_frame *rec, const hbd_adi_job *jobs, int n_jobs, int16_t *adi, void *stream)
{
    if (n_jobs <= 0) return 0;
    k_intra_adi<<<(n_jobs + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(*rec, jobs, n_jobs, adi);
    return static_cast<int>(cudaGetLastError());
}
