// hb_intra_core.cuh -- intra prediction as closed forms per sample (planar hmr_motion_intra.c:408, DC / angular with their edge filters :482,
// reference-sample smoothing adi_filter :189, the filtered-or-not rule :1122 / :1011), shared by the prediction kernels (hb_kernels_intra.cu)
// and the persistent intra reconstruction kernel (hb_kernels_tq.cu).  Everything here works on one warp and a 4n+1 sample array.
#pragma once
#include "hb_dev_common.cuh"

namespace {

__constant__ int c_ang[9] = { 0, 2, 5, 9, 13, 17, 21, 26, 32 };
__constant__ int c_inv_ang[9] = { 0, 4096, 1638, 910, 630, 482, 390, 315, 256 };
__constant__ int c_flt_thr[4] = { 10, 7, 1, 0 };

struct IntraMode { int kind; int hor; int angle; int inv; };      // kind 0 planar, 1 DC, 2 pure H/V, 3 angular

__device__ __forceinline__ IntraMode intra_mode_info(int mode)
{
    IntraMode m;
    m.hor = mode < 18; m.angle = 0; m.inv = 0;
    if (mode == 0) { m.kind = 0; return m; }
    if (mode == 1) { m.kind = 1; return m; }
    const int a = m.hor ? -(mode - 10) : mode - 26;
    const int aa = abs(a);
    m.angle = (a < 0 ? -1 : 1) * c_ang[aa];
    m.inv = c_inv_ang[aa];
    m.kind = a == 0 ? 2 : 3;
    return m;
}

// mid points at the top-left corner sample of the 4n+1 array
__device__ __forceinline__ int intra_sample(const int16_t *mid, int n, int lg, const IntraMode &m, int dc, bool edge, int x, int y)
{
    if (m.kind == 0) {
        const int l = mid[-(y + 1)], t = mid[x + 1], lb = mid[-(n + 1)], tr = mid[n + 1];
        return ((l << lg) + n + (x + 1) * (tr - l) + (t << lg) + (y + 1) * (lb - t)) >> (lg + 1);
    }
    if (m.kind == 1) {
        if (edge) {
            if (x == 0 && y == 0) return (mid[-1] + mid[1] + 2 * dc + 2) >> 2;
            if (y == 0) return (mid[1 + x] + 3 * dc + 2) >> 2;
            if (x == 0) return (mid[-1 - y] + 3 * dc + 2) >> 2;
        }
        return dc;
    }
    const int j = m.hor ? x : y, i = m.hor ? y : x;
    const int sm = m.hor ? -1 : 1;                          // main[k] = mid[sm*k], side[k] = mid[-sm*k]
    if (m.kind == 2) {
        int v = mid[sm * (i + 1)];
        if (edge && i == 0) v = hb_clip255(v + ((mid[-sm * (j + 1)] - mid[0]) >> 1));   // first sample of every line along the direction
        return v;
    }
    const int pos = (j + 1) * m.angle, d = pos >> 5, f = pos & 31;
    const int k = i + d + 1;
    auto ref = [&](int kk) -> int { return kk >= 0 ? mid[sm * kk] : mid[-sm * ((128 - kk * m.inv) >> 8)]; };
    if (!f) return ref(k);
    return ((32 - f) * ref(k) + f * ref(k + 1) + 16) >> 5;
}

// [1,2,1] or strong bilinear smoothing of the 4n+1 reference samples (hmr_motion_intra.c:189, strong_intra_smooth on)
__device__ __forceinline__ void intra_filter_adi(const int16_t *adi, int16_t *flt, int n, int lg, int lane)
{
    const int size = 4 * n + 1;
    const int lb = adi[0], lt = adi[2 * n], tr = adi[size - 1];
    const bool strong = n >= 32 && abs(lb + lt - 2 * adi[n]) < 8 && abs(lt + tr - 2 * adi[3 * n]) < 8;
    for (int i = lane; i < size; i += 32) {
        int v;
        if (i == 0 || i == size - 1) v = adi[i];
        else if (strong) {
            if (i == 2 * n) v = adi[i];
            else if (i < 2 * n) v = ((2 * n - i) * lb + i * lt + n) >> (lg + 1);
            else v = ((4 * n - i) * lt + (i - 2 * n) * tr + n) >> (lg + 1);
        } else v = (adi[i - 1] + 2 * adi[i] + adi[i + 1] + 2) >> 2;
        flt[i] = static_cast<int16_t>(v);
    }
}

__device__ __forceinline__ bool intra_uses_filtered(int lg, int mode)
{
    const int d = min(abs(mode - 10), abs(mode - 26));
    return mode != 1 && d > c_flt_thr[lg - 2];
}

__device__ __forceinline__ int intra_dc(const int16_t *mid, int n, int lane)
{
    int s = 0;
    for (int i = 1 + lane; i <= n; i += 32) s += mid[i] + mid[-i];
    s = __reduce_add_sync(HB_FULL_MASK, s);
    return (s + n) / (2 * n);
}

// The 4n+1 reference samples of an n x n block at (x, y) of plane p straight from the reconstructed picture: fill_reference_samples
// (hmr_motion_intra.c:246-406) in closed form.  With the reference's flags a left-bottom run implies a left column and a top-right run a row
// above (:625-657), and its two padding runs repeat the last sample it copied, so: index 2n - 1 - r = the left column at row y + min(r, rows
// available - 1), index 2n + 1 + c = the row above at column x + min(c, columns available - 1); a missing side repeats the other side's
// nearest sample (no side at all: 128); the corner exists only with both sides.  flags: bit 0 left, 1 top, 2 left-bottom, 3 top-right.
__device__ __forceinline__ void intra_gather_adi(const hbd_plane &p, int x, int y, int n, int flags, int lbs, int trs, int16_t *adi, int lane)
{
    const bool l = flags & 1, t = flags & 2;
    const int rows = n + ((flags & 4) ? lbs : 0), cols = n + ((flags & 8) ? trs : 0);
    const uint8_t *org = p.org;
    const int pitch = p.pitch;
    for (int k = lane; k < 4 * n + 1; k += 32) {
        int v;
        // ld.global.cg: inside the persistent kernel the samples were written by other SMs moments ago and L1 is not coherent
        if (!l && !t) v = 128;
        else if (k < 2 * n) v = l ? __ldcg(org + (y + min(2 * n - 1 - k, rows - 1)) * pitch + x - 1) : __ldcg(org + (y - 1) * pitch + x);
        else if (k == 2 * n) v = (l && t) ? __ldcg(org + (y - 1) * pitch + x - 1) : l ? __ldcg(org + y * pitch + x - 1) : __ldcg(org + (y - 1) * pitch + x);
        else v = t ? __ldcg(org + (y - 1) * pitch + x + min(k - 2 * n - 1, cols - 1)) : __ldcg(org + y * pitch + x - 1);
        adi[k] = static_cast<int16_t>(v);
    }
}

// one prediction of a block into the prediction plane from its raw reference samples (smoothed copy made here when the rule asks for it)
__device__ __forceinline__ void intra_predict_block(const hbd_plane &pred, int comp, int x, int y, int n, int mode, int filtered, const int16_t *raw, int16_t *flt, int lane)
{
    int lg = 2;
    while ((1 << lg) < n) lg++;
    const bool is_luma = comp == 0;
    const bool use_flt = is_luma && (filtered < 0 ? intra_uses_filtered(lg, mode) : filtered != 0);
    if (use_flt) { intra_filter_adi(raw, flt, n, lg, lane); __syncwarp(); }
    const int16_t *mid = (use_flt ? flt : raw) + 2 * n;
    const IntraMode m = intra_mode_info(mode);
    const int dc = m.kind == 1 ? intra_dc(mid, n, lane) : 0;
    const bool edge = is_luma && n <= 16;
    for (int e = lane; e < n * n; e += 32) {
        const int px = e % n, py = e / n;
        pred.org[(y + py) * pred.pitch + x + px] = static_cast<uint8_t>(intra_sample(mid, n, lg, m, dc, edge, px, py));
    }
}

}  // namespace
