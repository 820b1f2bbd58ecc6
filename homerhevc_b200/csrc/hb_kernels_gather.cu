// hb_kernels_gather.cu -- after the host has chosen a partition depth per CTU from the cost tables, pack what the
// host-side entropy coder and in-loop filters need of that choice: the reconstruction of the chosen depth (tight
// 8-bit planes) and the levels of the CODED transform units only, as one contiguous stream per CTU.
//
// One CTA per CTU.  Stream of a CTU (int16 units, starting at ctu_off[ctu]): for plane Y, U, V in turn, for every coded
// TU in raster order inside the CTU: { header_lo, header_hi, N*N levels }, header = plane << 28 | tu_size << 16 | TU
// raster position inside the CTU.  The host computed ctu_off from the sums it already fetched, so the launch needs no
// global scan and the total size is known before the copy is queued.
#include "hb_shim.h"
#include "hb_dev_common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_gather(const hbd_gather_args a)
{
    __shared__ int s_scan[256];
    __shared__ int s_base;
    const int ctu = blockIdx.x, tid = threadIdx.x;
    const int cx = ctu % a.ctu_cols, cy = ctu / a.ctu_cols;
    const int sel = a.sel[ctu];
    if (tid == 0) s_base = 0;

    // ---- reconstruction of the chosen depth
    for (int c = 0; c < 3; c++) {
        const int pass = c ? min(sel, 3) : sel;
        const hbd_plane src = a.recon[pass].p[c];
        const int cs = c ? 32 : 64, x0 = cx * cs, y0 = cy * cs;
        uint8_t *dst = a.out_recon[c];
        const int w = src.w, h = src.h;
        for (int e = tid; e < cs * cs / 4; e += 256) {
            const int r = e / (cs / 4), q = (e % (cs / 4)) * 4;
            if (y0 + r < h && x0 + q < w)
                *reinterpret_cast<uint32_t *>(dst + static_cast<size_t>(y0 + r) * a.out_pitch[c] + x0 + q) =
                    *reinterpret_cast<const uint32_t *>(src.org + (y0 + r) * src.pitch + x0 + q);
        }
    }
    __syncthreads();

    // ---- levels of the coded TUs
    int16_t *out = a.out_levels + a.ctu_off[ctu];
    for (int c = 0; c < 3; c++) {
        const int pass = c ? min(sel, 3) : sel;
        const hbd_gather_pc pc = a.pc[pass][c];
        const int cs = c ? 32 : 64, tpr = cs / pc.tu, n = tpr * tpr, nn = pc.tu * pc.tu;
        int idx = -1, len = 0;
        if (tid < n) {
            const int tx = cx * tpr + tid % tpr, ty = cy * tpr + tid / tpr;
            if (tx < pc.grid_w && ty < pc.grid_h) idx = pc.tu_index[ty * pc.grid_w + tx];
            if (idx >= 0 && pc.res[idx].sum > 0) len = 2 + nn; else idx = -1;
        }
        // exclusive scan of len over the 256 threads
        s_scan[tid] = len;
        __syncthreads();
        for (int d = 1; d < 256; d <<= 1) {
            const int v = tid >= d ? s_scan[tid - d] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int base = s_base;
        const int my_off = base + s_scan[tid] - len;
        if (idx >= 0) {
            const uint32_t hdr = (static_cast<uint32_t>(c) << 28) | (static_cast<uint32_t>(pc.tu) << 16) | static_cast<uint32_t>(tid);
            out[my_off] = static_cast<int16_t>(hdr & 0xffffu);
            out[my_off + 1] = static_cast<int16_t>(hdr >> 16);
        }
        // publish (offset, idx) so that warps can copy cooperatively
        const int total = s_scan[255];
        __syncthreads();                                   // everyone has read the scan before it is overwritten
        s_scan[tid] = idx >= 0 ? my_off : -1;
        __shared__ int s_idx[256];
        s_idx[tid] = idx;
        __syncthreads();
        const int warp = tid >> 5, lane = tid & 31;
        for (int t = warp; t < n; t += 8) {
            const int o = s_scan[t];
            if (o < 0) continue;
            const int16_t *src = pc.coeff + static_cast<size_t>(s_idx[t]) * nn;
            for (int e = lane; e < nn; e += 32) out[o + 2 + e] = src[e];
        }
        __syncthreads();
        if (tid == 0) s_base = base + total;
        __syncthreads();
    }
}

// ---- what deblocking needs to know per 4x4 unit about the host's choice, straight from the pre-pass tables in HBM (no host
// round trip): pass p of a CTU means CUs of 64 >> min(p, 3) with one luma TU each (four for p = 0 and p = 4)
__global__ void __launch_bounds__(256) k_units_from_selection(const hbd_units_args a)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= a.uw * a.uh) return;
    const int ux = i % a.uw, uy = i / a.uw;
    const int pass = a.sel[(uy >> 4) * a.ctu_cols + (ux >> 4)];
    const int d = min(pass, 3), cu = 64 >> d;
    const hb_me_result m = a.me[d][((uy * 4) / cu) * a.me_grid_w[d] + (ux * 4) / cu];
    const hbd_gather_pc pc = a.luma[pass];
    const int idx = pc.tu_index[((uy * 4) / pc.tu) * pc.grid_w + (ux * 4) / pc.tu];
    const int tu_depth = (pass == 0 || pass == 4) ? 1 : 0;
    hb_unit_info u;
    u.cu_depth = static_cast<uint8_t>(d); u.tu_depth = static_cast<uint8_t>(tu_depth); u.intra = 0;
    u.cbf_luma = static_cast<uint8_t>((idx >= 0 && pc.res[idx].sum > 0) ? (1 << tu_depth) : 0);
    u.ref_idx = 0; u.qp = static_cast<uint8_t>(a.qp);
    u.mvx = static_cast<int16_t>(m.mv.x); u.mvy = static_cast<int16_t>(m.mv.y);
    a.units[uy * a.units_w + ux] = u;
}

}  // namespace

// ---- SAO statistics (sao_get_ctu_stats, hmr_sao.c:75-330 / sse_sao_get_ctu_stats, hmr_sse42_sao.c:35):
// per CTU and component, for the four edge directions the count and the sum of (source - reconstruction) of every sample by
// edge class sign(c - a) + sign(c - b), and the same by band (c >> 3).  The reference walks rows with running sign buffers;
// per sample this is a closed form inside a rectangle that depends on the neighbouring CTUs and on the columns / rows the
// deblocking filter has not finished (5/3 columns, 4/2 rows).  One CTA per (CTU, component): edge classes accumulate in
// registers and are reduced once, bands go through shared atomics.
namespace {
__device__ __forceinline__ int sgn3(int v) { return (v > 0) - (v < 0); }

__global__ void __launch_bounds__(256) k_sao_stats(hbd_frame org, hbd_frame rec, int ctu_cols, hb_sao_stats *out)
{
    __shared__ int s_acc[104];                           // eo_diff[4][5], eo_count[4][5], bo_diff[32], bo_count[32]
    const int comp = blockIdx.y, ctu = blockIdx.x;
    const hbd_plane &pr = rec.p[comp], &po = org.p[comp];
    const int cs = comp ? 32 : 64;
    const int x0 = (ctu % ctu_cols) * cs, y0 = (ctu / ctu_cols) * cs;
    const int w = min(cs, pr.w - x0), h = min(cs, pr.h - y0);
    const bool l = x0 > 0, t = y0 > 0, r = x0 + cs < pr.w, b = y0 + cs < pr.h;
    const int skr = comp ? 3 : 5, skb = comp ? 2 : 4;
    for (int i = threadIdx.x; i < 104; i += 256) s_acc[i] = 0;
    __syncthreads();
    // the rectangles of the five types (hmr_sao.c:123-127, :154-157, :201-205, :262-264, :312-314)
    const int ex_e = r ? w - skr : w - 1, ex_f = r ? w - skr : w;          // EO_0/135/45 stop one short of a picture edge; EO_90 and BO do not
    const int sx_e = l ? 0 : 1, sy_v = t ? 0 : 1;
    const int ey_all = b ? h - skb : h, ey_v = b ? h - skb : h - 1;
    int cnt[4][5], dif[4][5];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int c = 0; c < 5; c++) { cnt[k][c] = 0; dif[k][c] = 0; }
    for (int i = threadIdx.x; i < w * h; i += 256) {
        const int x = i % w, y = i / w;
        const uint8_t *p = pr.org + (y0 + y) * pr.pitch + x0 + x;
        const int c = p[0], d = static_cast<int>(po.org[(y0 + y) * po.pitch + x0 + x]) - c;
        const bool in_e = x >= sx_e && x < ex_e, in_f = x < ex_f;
        const bool v0 = in_e && y < ey_all, v1 = in_f && y >= sy_v && y < ey_v, v23 = in_e && y >= sy_v && y < ey_v, vb = in_f && y < ey_all;
        // samples outside the picture are never classified (the rectangles exclude them); the border keeps the loads legal
        const int cls[4] = { 2 + sgn3(c - p[-1]) + sgn3(c - p[1]), 2 + sgn3(c - p[-pr.pitch]) + sgn3(c - p[pr.pitch]),
                             2 + sgn3(c - p[-pr.pitch - 1]) + sgn3(c - p[pr.pitch + 1]), 2 + sgn3(c - p[-pr.pitch + 1]) + sgn3(c - p[pr.pitch - 1]) };
        const bool val[4] = { v0, v1, v23, v23 };
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int q = 0; q < 5; q++) { const bool hit = val[k] && cls[k] == q; cnt[k][q] += hit; dif[k][q] += hit ? d : 0; }
        if (vb) { atomicAdd(&s_acc[40 + (c >> 3)], d); atomicAdd(&s_acc[72 + (c >> 3)], 1); }
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const int sd = __reduce_add_sync(HB_FULL_MASK, dif[k][q]), sc = __reduce_add_sync(HB_FULL_MASK, cnt[k][q]);
            if ((threadIdx.x & 31) == 0 && sc) { atomicAdd(&s_acc[k * 5 + q], sd); atomicAdd(&s_acc[20 + k * 5 + q], sc); }
        }
    __syncthreads();
    int *o = reinterpret_cast<int *>(out + (static_cast<size_t>(ctu) * 3 + comp));
    for (int i = threadIdx.x; i < 104; i += 256) o[i] = s_acc[i];
}
}  // namespace

// ---- deblocking, pixel stage (deblock_filter_luma / _chroma, filter_luma, filter_chroma, use_strong_filter,
// hmr_deblocking_filter.c:264-627; picture order of hmr_deblock_filter :827: one launch for every vertical edge, one for every
// horizontal edge).  One thread per 4x4 luma unit whose left (top) side lies on the 8x8 grid and carries a non-zero boundary
// strength: it decides on lines 0 and 3 and filters the four luma lines of its segment in place, plus -- strength 2, 8x8 chroma
// grid -- two lines of U and V.  Segments of one direction never touch the same samples.  Strengths and QPs are inputs.
namespace {
__constant__ uint8_t c_dbk_tc[54] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24 };
__constant__ uint8_t c_dbk_beta[52] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64 };
__constant__ uint8_t c_dbk_chroma_qp[58] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,29,30,31,32,33,33,34,34,35,35,36,36,37,37,38,39,40,41,42,43,44,45,46,47,48,49,50,51 };

struct DbkArgs {
    hbd_frame f;
    const uint8_t *bs, *qp;      // strengths of this direction, CU QP; per 4x4 unit, picture raster
    int units_w, dir;
    int cb_off, cr_off, beta_off2, tc_off2;
};

__device__ __forceinline__ int clamp3(int v, int lo, int hi) { return min(max(v, lo), hi); }

// one luma line across the edge: m[0..3] = p3 p2 p1 p0, m[4..7] = q0 q1 q2 q3
__device__ __forceinline__ void dbk_luma_line(int (&m)[8], int tc, bool sw, int thr_cut, bool second_p, bool second_q)
{
    const int m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4], m5 = m[5], m6 = m[6], m7 = m[7];
    if (sw) {
        m[3] = clamp3((m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3, m3 - 2 * tc, m3 + 2 * tc);
        m[4] = clamp3((m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3, m4 - 2 * tc, m4 + 2 * tc);
        m[2] = clamp3((m1 + m2 + m3 + m4 + 2) >> 2, m2 - 2 * tc, m2 + 2 * tc);
        m[5] = clamp3((m3 + m4 + m5 + m6 + 2) >> 2, m5 - 2 * tc, m5 + 2 * tc);
        m[1] = clamp3((2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3, m1 - 2 * tc, m1 + 2 * tc);
        m[6] = clamp3((m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3, m6 - 2 * tc, m6 + 2 * tc);
    } else {
        int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
        if (abs(delta) < thr_cut) {
            const int tc2 = tc >> 1;
            delta = clamp3(delta, -tc, tc);
            m[3] = hb_clip255(m3 + delta);
            m[4] = hb_clip255(m4 - delta);
            if (second_p) m[2] = hb_clip255(m2 + clamp3((((m1 + m3 + 1) >> 1) - m2 + delta) >> 1, -tc2, tc2));
            if (second_q) m[5] = hb_clip255(m5 + clamp3((((m6 + m4 + 1) >> 1) - m5 - delta) >> 1, -tc2, tc2));
        }
    }
}

__global__ void __launch_bounds__(256) k_deblock(const DbkArgs a)
{
    const hbd_plane &py = a.f.p[0];
    const int uw = py.w >> 2, uh = py.h >> 2;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= uw * uh) return;
    const int ux = i % uw, uy = i / uw, along = a.dir ? uy : ux;
    const int bs = a.bs[uy * a.units_w + ux];
    if (!bs || along == 0 || (along & 1)) return;
    const int qpq = a.qp[uy * a.units_w + ux], qpp = a.dir ? a.qp[(uy - 1) * a.units_w + ux] : a.qp[uy * a.units_w + ux - 1];
    const int q = (qpp + qpq + 1) >> 1;
    {
        const int tc = c_dbk_tc[clamp3(q + 2 * (bs - 1) + 2 * a.tc_off2, 0, 53)], beta = c_dbk_beta[clamp3(q + 2 * a.beta_off2, 0, 51)];
        const int side_thr = (beta + (beta >> 1)) >> 3, thr_cut = tc * 10;
        uint8_t *e = py.org + (4 * uy) * py.pitch + 4 * ux;          // q0 of line 0
        int m[4][8];
        if (a.dir == 0) {                                            // vertical edge: a line is a row, samples x-4 .. x+3
#pragma unroll
            for (int l = 0; l < 4; l++) {
                const uint32_t *rw = reinterpret_cast<const uint32_t *>(e + l * py.pitch - 4);      // 4-byte aligned, not 8
                const uint32_t wp = rw[0], wq = rw[1];
#pragma unroll
                for (int k = 0; k < 4; k++) { m[l][k] = (wp >> (8 * k)) & 255; m[l][4 + k] = (wq >> (8 * k)) & 255; }
            }
        } else {                                                     // horizontal edge: a line is a column, rows y-4 .. y+3
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t wv = *reinterpret_cast<const uint32_t *>(e + (k - 4) * py.pitch);
#pragma unroll
                for (int l = 0; l < 4; l++) m[l][k] = (wv >> (8 * l)) & 255;
            }
        }
        auto dp = [&](int l) { return abs(m[l][1] - 2 * m[l][2] + m[l][3]); };
        auto dq = [&](int l) { return abs(m[l][4] - 2 * m[l][5] + m[l][6]); };
        const int dp0 = dp(0), dq0 = dq(0), dp3 = dp(3), dq3 = dq(3);
        const int d0 = dp0 + dq0, d3 = dp3 + dq3;
        if (d0 + d3 < beta) {
            auto strong = [&](int l, int d) { return abs(m[l][0] - m[l][3]) + abs(m[l][7] - m[l][4]) < (beta >> 3) && d < (beta >> 2) && abs(m[l][3] - m[l][4]) < ((tc * 5 + 1) >> 1); };
            const bool sw = strong(0, 2 * d0) && strong(3, 2 * d3);
#pragma unroll
            for (int l = 0; l < 4; l++) dbk_luma_line(m[l], tc, sw, thr_cut, dp0 + dp3 < side_thr, dq0 + dq3 < side_thr);
            if (a.dir == 0) {
#pragma unroll
                for (int l = 0; l < 4; l++)
                {
                    uint32_t *rw = reinterpret_cast<uint32_t *>(e + l * py.pitch - 4);
                    rw[0] = hb_pack4(m[l][0], m[l][1], m[l][2], m[l][3]); rw[1] = hb_pack4(m[l][4], m[l][5], m[l][6], m[l][7]);
                }
            } else {
#pragma unroll
                for (int k = 1; k < 7; k++)
                    *reinterpret_cast<uint32_t *>(e + (k - 4) * py.pitch) = hb_pack4(m[0][k], m[1][k], m[2][k], m[3][k]);
            }
        }
    }
    if (bs > 1 && (along & 3) == 0) {
#pragma unroll
        for (int c = 1; c < 3; c++) {
            const hbd_plane &pc = a.f.p[c];
            const int qc = c_dbk_chroma_qp[clamp3(q + (c == 1 ? a.cb_off : a.cr_off), 0, 57)];
            const int tc = c_dbk_tc[clamp3(qc + 2 * (bs - 1) + 2 * a.tc_off2, 0, 53)];
            const int o = a.dir ? pc.pitch : 1, step = a.dir ? 1 : pc.pitch;
            uint8_t *e = pc.org + (2 * uy) * pc.pitch + 2 * ux;
#pragma unroll
            for (int l = 0; l < 2; l++) {
                uint8_t *s2 = e + l * step;
                const int m2 = s2[-2 * o], m3 = s2[-o], m4 = s2[0], m5 = s2[o];
                const int delta = clamp3((((m4 - m3) << 2) + m2 - m5 + 4) >> 3, -tc, tc);
                s2[-o] = static_cast<uint8_t>(hb_clip255(m3 + delta));
                s2[0] = static_cast<uint8_t>(hb_clip255(m4 - delta));
            }
        }
    }
}
}  // namespace

namespace {
// ---- boundary strengths of a P picture from per-unit mode data (hmr_deblock_filter_cu :737 edge marking, set_edge_filter_pu :692,
// get_boundary_strength_single :138): one thread per 4x4 unit writes the strength of its left and top side and its QP
__global__ void __launch_bounds__(256) k_deblock_strengths(const hb_unit_info *units, int units_w, int uw, int uh, uint8_t *bs_ver, uint8_t *bs_hor, uint8_t *qp)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= uw * uh) return;
    const int ux = i % uw, uy = i / uw, qi = uy * units_w + ux;
    const hb_unit_info q = units[qi];
    qp[qi] = q.qp;
    const int ts = max(64 >> (q.cu_depth + q.tu_depth), 8);
#pragma unroll
    for (int dir = 0; dir < 2; dir++) {
        const int pos = 4 * (dir ? uy : ux);
        int bs = 0;
        if (pos != 0 && pos % ts == 0) {
            const hb_unit_info p = units[dir ? qi - units_w : qi - 1];
            if (p.intra || q.intra) bs = 2;
            else if (((q.cbf_luma >> q.tu_depth) & 1) || ((p.cbf_luma >> p.tu_depth) & 1)) bs = 1;
            else bs = (p.ref_idx != q.ref_idx) || abs(q.mvx - p.mvx) >= 4 || abs(q.mvy - p.mvy) >= 4;
        }
        (dir ? bs_hor : bs_ver)[qi] = static_cast<uint8_t>(bs);
    }
}
}  // namespace

extern "C" int hbk_deblock_strengths(const hb_unit_info *units, int units_w, int w, int h, uint8_t *bs_ver, uint8_t *bs_hor, uint8_t *qp, void *stream)
{
    const int uw = w >> 2, uh = h >> 2;
    k_deblock_strengths<<<(uw * uh + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(units, units_w, uw, uh, bs_ver, bs_hor, qp);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_deblock(const hbd_frame *f, const uint8_t *bs_ver, const uint8_t *bs_hor, const uint8_t *qp, int units_w,
                           int cb_off, int cr_off, int beta_off2, int tc_off2, void *stream)
{
    DbkArgs a;
    a.f = *f; a.qp = qp; a.units_w = units_w; a.cb_off = cb_off; a.cr_off = cr_off; a.beta_off2 = beta_off2; a.tc_off2 = tc_off2;
    const int n = (f->p[0].w >> 2) * (f->p[0].h >> 2);
    a.bs = bs_ver; a.dir = 0;
    k_deblock<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    a.bs = bs_hor; a.dir = 1;
    k_deblock<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return static_cast<int>(cudaGetLastError());
}

namespace {
// ---- SAO offset pass (offset_block, hmr_sao.c:960; sao_offset_ctu :1210): dst = clip(src + offset[class]) inside the rectangle of
// the CTU's type, a plain copy elsewhere, so dst is a complete picture.  src is the deblocked picture; classes always come
// from src, as the reference reads them from its untouched copy (sao_aux_wnd).  Four samples per thread along a row.
__global__ void __launch_bounds__(256) k_sao_apply(hbd_frame src, hbd_frame dst, int ctu_cols, const hb_sao_param *prm)
{
    __shared__ int s_off[32];
    const int comp = blockIdx.y, ctu = blockIdx.x;
    const hbd_plane &ps = src.p[comp], &pd = dst.p[comp];
    const int cs = comp ? 32 : 64;
    const int x0 = (ctu % ctu_cols) * cs, y0 = (ctu / ctu_cols) * cs;
    const int w = min(cs, ps.w - x0), h = min(cs, ps.h - y0);
    const bool l = x0 > 0, t = y0 > 0, r = x0 + cs < ps.w, b = y0 + cs < ps.h;
    const int type = prm[ctu].type[comp];
    if (threadIdx.x < 32) s_off[threadIdx.x] = prm[ctu].offset[comp][threadIdx.x];
    __syncthreads();
    int sx = 0, ex = w, sy = 0, ey = h;
    if (type == 0 || type == 2 || type == 3) { sx = l ? 0 : 1; ex = r ? w : w - 1; }
    if (type == 1 || type == 2 || type == 3) { sy = t ? 0 : 1; ey = b ? h : h - 1; }
    if (type < 0 || type > 4) { ex = 0; ey = 0; }
    const int ddx = type == 3 ? -1 : (type == 1 ? 0 : 1), ddy = type == 0 ? 0 : 1;       // second neighbour; the first is its opposite
    const int nb = ddy * ps.pitch + ddx;
    for (int i = threadIdx.x; i < (w >> 2) * h; i += 256) {
        const int x4 = (i % (w >> 2)) * 4, y = i / (w >> 2);
        const uint8_t *p = ps.org + (y0 + y) * ps.pitch + x0 + x4;
        uint32_t outw = *reinterpret_cast<const uint32_t *>(p);
        if (y >= sy && y < ey) {
            int o[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int c = p[k];
                int v = c;
                if (x4 + k >= sx && x4 + k < ex) {
                    const int cls = type == 4 ? (c >> 3) : 2 + sgn3(c - p[k - nb]) + sgn3(c - p[k + nb]);
                    v = c + s_off[cls];
                }
                o[k] = v;
            }
            outw = hb_pack_sat_u8x4(o[0], o[1], o[2], o[3]);
        }
        *reinterpret_cast<uint32_t *>(pd.org + (y0 + y) * pd.pitch + x0 + x4) = outw;
    }
}
}  // namespace

extern "C" int hbk_sao_apply(const hbd_frame *src, const hbd_frame *dst, int ctu_cols, int n_ctus, const hb_sao_param *prm, void *stream)
{
    if (n_ctus <= 0) return 0;
    k_sao_apply<<<dim3(n_ctus, 3), 256, 0, static_cast<cudaStream_t>(stream)>>>(*src, *dst, ctu_cols, prm);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_sao_stats(const hbd_frame *org, const hbd_frame *rec, int ctu_cols, int n_ctus, hb_sao_stats *out, void *stream)
{
    if (n_ctus <= 0) return 0;
    k_sao_stats<<<dim3(n_ctus, 3), 256, 0, static_cast<cudaStream_t>(stream)>>>(*org, *rec, ctu_cols, out);
    return static_cast<int>(cudaGetLastError());
}

// ---- compact wire format of the result tables: 12-byte records instead of 24 / 16 (hb_prepass_cfg.compact_tables)
namespace {
__global__ void __launch_bounds__(256) k_pack_tables(const hb_me_result *me, const hb_tu_result *tu, hb_me_result_c *me_c, hb_tu_result_c *tu_c, int n_me, int n_tu)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n_me) {
        const hb_me_result r = me[i];
        hb_me_result_c c;
        c.mvx = static_cast<int16_t>(r.mv.x); c.mvy = static_cast<int16_t>(r.mv.y); c.sad = r.sad;
        c.n_probes = static_cast<uint16_t>(min(r.n_probes, 65535u)); c.subx = static_cast<int8_t>(r.subpix.x); c.suby = static_cast<int8_t>(r.subpix.y);
        me_c[i] = c;
    } else if (i < n_me + n_tu) {
        const hb_tu_result r = tu[i - n_me];
        hb_tu_result_c c;
        c.ssd = r.ssd; c.ssd_zero = r.ssd_zero; c.sum_zeroed = (static_cast<uint32_t>(r.sum) & 0x7fffffffu) | (r.zeroed ? 0x80000000u : 0u);
        tu_c[i - n_me] = c;
    }
}
}  // namespace

extern "C" int hbk_pack_tables(const void *full, void *compact, int n_me, int n_tu, void *stream)
{
    const int n = n_me + n_tu;
    if (n <= 0) return 0;
    const hb_me_result *me = static_cast<const hb_me_result *>(full);
    hb_me_result_c *me_c = static_cast<hb_me_result_c *>(compact);
    k_pack_tables<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(me, reinterpret_cast<const hb_tu_result *>(me + n_me), me_c,
                                                                                 reinterpret_cast<hb_tu_result_c *>(me_c + n_me), n_me, n_tu);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_units_from_selection(const hbd_units_args *a, void *stream)
{
    const int n = a->uw * a->uh;
    if (n <= 0) return 0;
    k_units_from_selection<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_gather(const hbd_gather_args *a, int n_ctus, void *stream)
{
    if (n_ctus <= 0) return 0;
    k_gather<<<n_ctus, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
    return static_cast<int>(cudaGetLastError());
}
