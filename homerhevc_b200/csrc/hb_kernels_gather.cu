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
    __shared__ int s_warp[8];                            // coded TUs per warp (scan)
    __shared__ int s_list[256];                          // compacted coded TUs of the plane: index among the plane's coded TUs
    __shared__ int s_pos[256];                           // ... and raster position inside the CTU
    const int ctu = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cx = ctu % a.ctu_cols, cy = ctu / a.ctu_cols;
    const int sel = a.sel[ctu];

    // ---- reconstruction of the chosen pass: all loads of a thread are issued before its stores
    for (int c = 0; c < 3; c++) {
        const int pass = c ? min(sel, 3) : sel;
        const hbd_plane src = a.recon[pass].p[c];
        const int cs = c ? 32 : 64, x0 = cx * cs, y0 = cy * cs;
        uint8_t *__restrict__ dst = a.out_recon[c];
        const int pitch = a.out_pitch[c], wq = cs / 4, per = cs * cs / 4 / 256;     // 4 words per thread for luma, 1 for chroma
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int e = tid + k * 256, r = e / wq, q = (e % wq) * 4;
            const bool ok = k < per && y0 + r < src.h && x0 + q < src.w;
            v[k] = ok ? __ldg(reinterpret_cast<const uint32_t *>(src.org + (y0 + r) * src.pitch + x0 + q)) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int e = tid + k * 256, r = e / wq, q = (e % wq) * 4;
            if (k < per && y0 + r < src.h && x0 + q < src.w) *reinterpret_cast<uint32_t *>(dst + static_cast<size_t>(y0 + r) * pitch + x0 + q) = v[k];
        }
    }

    // ---- levels of the coded TUs: { hdr_lo, hdr_hi, N*N levels } per TU in raster order.  Every coded TU of a plane has the same
    // length, so a ballot scan yields its rank and the copy runs over (TU, 16-byte chunk) pairs with independent loads
    int16_t *out = a.out_levels + a.ctu_off[ctu];
    int base = 0;
    for (int c = 0; c < 3; c++) {
        const int pass = c ? min(sel, 3) : sel;
        const hbd_gather_pc pc = a.pc[pass][c];
        const int cs = c ? 32 : 64, tpr = cs / pc.tu, n = tpr * tpr, nn = pc.tu * pc.tu;
        int idx = -1;
        if (tid < n) {
            const int tx = cx * tpr + tid % tpr, ty = cy * tpr + tid / tpr;
            if (tx < pc.grid_w && ty < pc.grid_h) idx = pc.tu_index[ty * pc.grid_w + tx];
            if (idx >= 0 && pc.res[idx].sum <= 0) idx = -1;
        }
        const uint32_t m = __ballot_sync(HB_FULL_MASK, idx >= 0);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { const int v = s_warp[k]; before += k < warp ? v : 0; total += v; }
        if (idx >= 0) {
            const int rank = before + __popc(m & ((1u << lane) - 1u));
            s_list[rank] = idx; s_pos[rank] = tid;
        }
        __syncthreads();
        const int rec_len = 2 + nn, cpt = nn / 8;         // int16 per record, 16-byte chunks per TU
        for (int k = tid; k < total; k += 256) {
            const uint32_t hdr = (static_cast<uint32_t>(c) << 28) | (static_cast<uint32_t>(pc.tu) << 16) | static_cast<uint32_t>(s_pos[k]);
            *reinterpret_cast<uint32_t *>(out + base + k * rec_len) = hdr;              // low half first: hdr_lo, hdr_hi (the stream is 4-byte aligned)
        }
        for (int q = tid; q < total * cpt; q += 256) {
            const int k = q / cpt, j = q % cpt;
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(pc.coeff + static_cast<size_t>(s_list[k]) * nn) + j);
            uint32_t *o = reinterpret_cast<uint32_t *>(out + base + k * rec_len + 2 + j * 8);
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        }
        base += total * rec_len;
        __syncthreads();                                   // the lists are rewritten by the next plane
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
// deblocking filter has not finished (5/3 columns, 4/2 rows).  One CTA per (CTU, component).
namespace {
__device__ __forceinline__ int sgn3(int v) { return (v > 0) - (v < 0); }

// Accumulation: every lane owns a private column of 52 bins (20 edge-class bins, 32 bands) in shared memory -- address
// (warp, bin, lane), so bank = lane and plain load / add / store never conflicts and needs no atomics.  A bin word packs the
// sample count (bits 16..) and the sum of the biased differences d + 256 (bits 0..15): a lane sees at most 32 samples per bin.
constexpr int kSaoWarps = 4, kSaoBins = 52;

__global__ void __launch_bounds__(kSaoWarps * 32) k_sao_stats(hbd_frame org, hbd_frame rec, int ctu_cols, hb_sao_stats *out)
{
    __shared__ uint32_t s_bin[kSaoWarps][kSaoBins][32];
    const int comp = blockIdx.y, ctu = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const hbd_plane pr = hbd_pick_plane(rec, comp), po = hbd_pick_plane(org, comp);
    const int cs = comp ? 32 : 64, lcs = comp ? 5 : 6;
    const int x0 = (ctu % ctu_cols) * cs, y0 = (ctu / ctu_cols) * cs;
    const int w = min(cs, pr.w - x0), h = min(cs, pr.h - y0);
    const bool l = x0 > 0, t = y0 > 0, r = x0 + cs < pr.w, b = y0 + cs < pr.h;
    const int skr = comp ? 3 : 5, skb = comp ? 2 : 4;
    for (int i = threadIdx.x; i < kSaoWarps * kSaoBins * 32; i += kSaoWarps * 32) (&s_bin[0][0][0])[i] = 0;
    __syncthreads();
    // the rectangles of the five types (hmr_sao.c:123-127, :154-157, :201-205, :262-264, :312-314)
    const int ex_e = r ? w - skr : w - 1, ex_f = r ? w - skr : w;          // EO_0/135/45 stop one short of a picture edge; EO_90 and BO do not
    const int sx_e = l ? 0 : 1, sy_v = t ? 0 : 1;
    const int ey_all = b ? h - skb : h, ey_v = b ? h - skb : h - 1;
    uint32_t *mine = &s_bin[warp][0][lane];
    const int pitch = pr.pitch;
    // four samples of a row per step: the row and its two neighbours as aligned words (9 loads for 4 samples), the four edge classes of all
    // four samples by packed compares (class = 2 + sgn(c - a) + sgn(c - b) per byte), then the bin updates sample by sample
    const int pq = pitch >> 2, cs4 = cs >> 2;
    for (int i = threadIdx.x; i < cs4 * h; i += kSaoWarps * 32) {
        const int x = (i % cs4) * 4, y = i / cs4;
        if (x >= w) continue;
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(pr.org + (y0 + y) * pitch + x0 + x);
        const uint32_t c4 = pw[0], u0 = pw[-pq], d0 = pw[pq];
        const uint32_t l4 = __funnelshift_r(pw[-1], c4, 24), r4 = __funnelshift_r(c4, pw[1], 8);
        const uint32_t ul = __funnelshift_r(pw[-pq - 1], u0, 24), ur = __funnelshift_r(u0, pw[-pq + 1], 8);
        const uint32_t dl = __funnelshift_r(pw[pq - 1], d0, 24), dr = __funnelshift_r(d0, pw[pq + 1], 8);
        auto cls4 = [&](uint32_t a, uint32_t b) -> uint32_t {      // no byte ever borrows: 2 + (0..2) - (0..2), and a sample cannot be both above and below
            return 0x02020202u + (__vcmpgtu4(c4, a) & 0x01010101u) + (__vcmpgtu4(c4, b) & 0x01010101u)
                               - (__vcmpltu4(c4, a) & 0x01010101u) - (__vcmpltu4(c4, b) & 0x01010101u);
        };
        const uint32_t e0 = cls4(l4, r4), e1 = cls4(u0, d0), e2 = cls4(ul, dr), e3 = cls4(ur, dl);
        const uint32_t o4 = *reinterpret_cast<const uint32_t *>(po.org + (y0 + y) * po.pitch + x0 + x);
        const bool row_all = y < ey_all, row_v = y >= sy_v && y < ey_v;
        // samples outside the picture are never classified (the rectangles exclude them); the border keeps the loads legal
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int xx = x + k, c = (c4 >> (8 * k)) & 255;
            const uint32_t inc = 0x10000u + 256u + ((o4 >> (8 * k)) & 255u) - static_cast<uint32_t>(c);
            const bool in_e = xx >= sx_e && xx < ex_e, in_f = xx < ex_f;
            if (in_e && row_all) mine[(0 + ((e0 >> (8 * k)) & 7)) * 32] += inc;
            if (in_f && row_v) mine[(5 + ((e1 >> (8 * k)) & 7)) * 32] += inc;
            if (in_e && row_v) {
                mine[(10 + ((e2 >> (8 * k)) & 7)) * 32] += inc;
                mine[(15 + ((e3 >> (8 * k)) & 7)) * 32] += inc;
            }
            if (in_f && row_all) mine[(20 + (c >> 3)) * 32] += inc;
        }
    }
    __syncthreads();
    int *o = reinterpret_cast<int *>(out + (static_cast<size_t>(ctu) * 3 + comp));     // eo_diff[4][5], eo_count[4][5], bo_diff[32], bo_count[32]
    for (int bin = warp; bin < kSaoBins; bin += kSaoWarps) {
        int cnt = 0, sum = 0;
#pragma unroll
        for (int ww = 0; ww < kSaoWarps; ww++) { const uint32_t v = s_bin[ww][bin][lane]; cnt += v >> 16; sum += v & 0xffffu; }
        cnt = __reduce_add_sync(HB_FULL_MASK, cnt); sum = __reduce_add_sync(HB_FULL_MASK, sum);
        if (lane == 0) {
            const int dif = sum - 256 * cnt;
            if (bin < 20) { o[bin] = dif; o[20 + bin] = cnt; }
            else { o[40 + bin - 20] = dif; o[72 + bin - 20] = cnt; }
        }
    }
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
struct DbkPics { int32_t l0[16], l1[16]; };               // the pictures the reference indices of the two lists name (B pictures)
__global__ void __launch_bounds__(256) k_deblock_strengths(const hb_unit_info *units, const hb_unit_l1 *units1, const DbkPics pics, int units_w, int uw, int uh,
                                                           uint8_t *bs_ver, uint8_t *bs_hor, uint8_t *qp)
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
            else if (units1 == nullptr) bs = (p.ref_idx != q.ref_idx) || abs(q.mvx - p.mvx) >= 4 || abs(q.mvy - p.mvy) >= 4;
            else {
                // B picture (hmr_deblocking_filter.c:173-229): two (picture, vector) pairs a side; an unused list has no picture and a zero vector
                const hb_unit_l1 p1 = units1[dir ? qi - units_w : qi - 1], q1 = units1[qi];
                auto pic = [&](int r, bool list1) { return r < 0 ? -1 : (list1 ? pics.l1[r & 15] : pics.l0[r & 15]); };
                const int r0 = pic(p.ref_idx, false), r1 = pic(p1.ref_idx, true), c0 = pic(q.ref_idx, false), c1 = pic(q1.ref_idx, true);
                const int p0x = r0 < 0 ? 0 : p.mvx, p0y = r0 < 0 ? 0 : p.mvy, p1x = r1 < 0 ? 0 : p1.mvx, p1y = r1 < 0 ? 0 : p1.mvy;
                const int q0x = c0 < 0 ? 0 : q.mvx, q0y = c0 < 0 ? 0 : q.mvy, q1x = c1 < 0 ? 0 : q1.mvx, q1y = c1 < 0 ? 0 : q1.mvy;
                auto far = [](int ax, int ay, int bx, int by) { return abs(ax - bx) >= 4 || abs(ay - by) >= 4; };
                if ((r0 == c0 && r1 == c1) || (r0 == c1 && r1 == c0)) {
                    if (r0 != r1) bs = (r0 == c0) ? (far(q0x, q0y, p0x, p0y) || far(q1x, q1y, p1x, p1y)) : (far(q1x, q1y, p0x, p0y) || far(q0x, q0y, p1x, p1y));
                    else bs = (far(q0x, q0y, p0x, p0y) || far(q1x, q1y, p1x, p1y)) && (far(q1x, q1y, p0x, p0y) || far(q0x, q0y, p1x, p1y));
                } else bs = 1;
            }
        }
        (dir ? bs_hor : bs_ver)[qi] = static_cast<uint8_t>(bs);
    }
}
}  // namespace

namespace {
// ---- AMVP candidates of 2Nx2N PUs from the per-unit motion field (get_amvp_candidates hmr_motion_inter.c:2342; neighbour look-ups
// hmr_arithmetic_encoding.c:229-354; neighbour flags of the quadtree hmr_motion_intra.c:625-657, CTU level :676-683): one thread per PU
__device__ __forceinline__ int zscan16(int ux, int uy)
{
    int a = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) a |= ((ux >> b) & 1) << (2 * b) | ((uy >> b) & 1) << (2 * b + 1);
    return a;
}
// the five spatial neighbours of a PU (A0 left-bottom, A1 left, B0 top-right, B1 top, B2 top-left): unit index in the maps, vector,
// and the bit mask of those that exist, were coded before the PU and are inter predicted
__device__ __forceinline__ int pu_neighbours(const hb_unit_info *units, int units_w, int w, int h, int x, int y, int size, hb_mv (&cand)[5])
{
    const int ctu_x = x & ~63, ctu_y = y & ~63, px = x - ctu_x, py = y - ctu_y;
    const int cols = (w + 63) / 64;
    const bool has_left = ctu_x > 0, has_top = ctu_y > 0, has_top_right = ctu_y > 0 && ctu_x / 64 + 1 < cols, has_top_left = ctu_x > 0 && ctu_y > 0;
    // the partition's left_bottom / top_right flags, handed down from the CTU (whose left-bottom neighbour never exists)
    bool l = has_left, t = has_top, lb = false, tr = has_top_right;
    {
        const int valid_lines = min(64, h - ctu_y), valid_cols = min(64, w - ctu_x);
        int par_x = 0, par_y = 0;
        for (int s = 32; s >= size; s >>= 1) {
            const int cx = par_x + ((px - par_x) >= s ? s : 0), cy = par_y + ((py - par_y) >= s ? s : 0);
            const bool nlb = (lb && cx == par_x) || (l && cx == par_x && cy == par_y && valid_lines > cy + s);
            const bool ntr = (tr && cy == par_y) || (t && cx == par_x && cy == par_y && valid_cols > cx + s) || (cx == par_x && cy != par_y && valid_cols > cx + s);
            l = l || cx; t = t || cy; lb = nlb; tr = ntr; par_x = cx; par_y = cy;
        }
    }
    const int gx0 = ctu_x / 4, gy0 = ctu_y / 4;
    int cu[5], okm = 0;
    {
        const int ux = px / 4, uy = (py + size) / 4 - 1;   // bottom-left unit of the PU
        bool a0;
        if (!lb) a0 = false;
        else if (ux == 0 && uy == 15) a0 = false;
        else if (ux == 0) a0 = has_left;
        else if (uy == 15) a0 = false;
        else a0 = zscan16(ux, uy) > zscan16(ux - 1, uy + 1);
        cu[0] = (gy0 + uy + 1) * units_w + gx0 + ux - 1;
        cu[1] = (gy0 + uy) * units_w + gx0 + ux - 1;
        okm |= (a0 ? 1 : 0) | ((ux == 0 ? has_left : true) ? 2 : 0);
    }
    {
        const int ux = (px + size) / 4 - 1, uy = py / 4;   // top-right unit
        bool b0;
        if (!tr) b0 = false;
        else if (ux == 15 && uy == 0) b0 = has_top_right;
        else if (uy == 0) b0 = has_top;
        else if (ux == 15) b0 = false;
        else b0 = zscan16(ux, uy) > zscan16(ux + 1, uy - 1);
        cu[2] = (gy0 + uy - 1) * units_w + gx0 + ux + 1;
        cu[3] = (gy0 + uy - 1) * units_w + gx0 + ux;
        okm |= (b0 ? 4 : 0) | ((uy == 0 ? has_top : true) ? 8 : 0);
    }
    {
        const int ux = px / 4, uy = py / 4;                // top-left unit
        cu[4] = (gy0 + uy - 1) * units_w + gx0 + ux - 1;
        okm |= ((ux == 0 && uy == 0) ? has_top_left : uy == 0 ? has_top : ux == 0 ? has_left : true) ? 16 : 0;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
        cand[k].x = 0; cand[k].y = 0;
        if ((okm >> k) & 1) {
            const hb_unit_info u = units[cu[k]];
            if (u.intra || u.ref_idx < 0) okm &= ~(1 << k);
            else { cand[k].x = u.mvx; cand[k].y = u.mvy; }
        }
    }
    return okm;
}

__device__ __forceinline__ hb_amvp_list amvp_of(const hb_unit_info *units, int units_w, int w, int h, int x, int y, int size)
{
    hb_mv cand[5];
    const int okm = pu_neighbours(units, units_w, w, h, x, y, size, cand);
    hb_mv list[3];
    int n = 0;
    const bool smvp = (okm & 3) != 0;
    if (okm & 1) list[n++] = cand[0]; else if (okm & 2) list[n++] = cand[1];
    const int b = (okm & 4) ? 2 : (okm & 8) ? 3 : (okm & 16) ? 4 : -1;
    if (b >= 0) { list[n] = b == 2 ? cand[2] : b == 3 ? cand[3] : cand[4]; n++; }
    if (!smvp && b >= 0) { list[n] = list[n - 1]; n++; }           // the above group walked a second time (:2405-2420): the same vector again
    if (n == 2 && list[0].x == list[1].x && list[0].y == list[1].y) n = 1;
    if (n > 2) n = 2;
    hb_amvp_list r;
    r.mv[0].x = n > 0 ? list[0].x : 0; r.mv[0].y = n > 0 ? list[0].y : 0;
    r.mv[1].x = n > 1 ? list[1].x : 0; r.mv[1].y = n > 1 ? list[1].y : 0;
    return r;
}
__global__ void __launch_bounds__(128) k_amvp(const hb_unit_info *units, int units_w, int w, int h, const hb_amvp_job *jobs, int n_jobs, hb_amvp_list *out)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n_jobs) return;
    out[i] = amvp_of(units, units_w, w, h, jobs[i].x, jobs[i].y, jobs[i].size);
}
// the same, written straight into the predictor fields of queued search jobs (hb_me_search_field: no trip to the host in between)
__global__ void __launch_bounds__(128) k_amvp_fill(const hb_unit_info *units, int units_w, int w, int h, hbd_me_job *jobs, int n_jobs, int size)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n_jobs) return;
    const hb_amvp_list r = amvp_of(units, units_w, w, h, jobs[i].x, jobs[i].y, size);
    jobs[i].n_amvp = 2;
    jobs[i].amvp[0] = r.mv[0].x; jobs[i].amvp[1] = r.mv[0].y; jobs[i].amvp[2] = r.mv[1].x; jobs[i].amvp[3] = r.mv[1].y;
}

// merge candidates (get_merge_mvp_candidates hmr_motion_inter.c:1937, P picture, one reference picture): A1, B1, B0, A0, then B2 while fewer
// than four, each pruned against the neighbours the reference compares it with (equal_motion :1915); closed at max_cands, zero filled
__global__ void __launch_bounds__(128) k_merge_cands(const hb_unit_info *units, int units_w, int w, int h, const hb_amvp_job *jobs, int n_jobs, int max_cands, hb_mv *out)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n_jobs) return;
    hb_mv c[5];
    const int okm = pu_neighbours(units, units_w, w, h, jobs[i].x, jobs[i].y, jobs[i].size, c);
    const bool a0 = okm & 1, a1 = okm & 2, b0 = okm & 4, b1 = okm & 8, b2 = okm & 16;
    auto same = [&](int p, int q) { return c[p].x == c[q].x && c[p].y == c[q].y; };
    hb_mv *o = out + static_cast<size_t>(i) * max_cands;
    int n = 0;
    if (a1) o[n++] = c[1];
    if (n < max_cands && b1 && (!a1 || !same(1, 3))) o[n++] = c[3];
    if (n < max_cands && b0 && (!b1 || !same(3, 2))) o[n++] = c[2];
    if (n < max_cands && a0 && (!a1 || !same(1, 0))) o[n++] = c[0];
    if (n < max_cands && n < 4 && b2 && (!a1 || !same(1, 4)) && (!b1 || !same(3, 4))) o[n++] = c[4];
    for (; n < max_cands; n++) { o[n].x = 0; o[n].y = 0; }
}
}  // namespace

extern "C" int hbk_amvp(const hb_unit_info *units, int units_w, int w, int h, const hb_amvp_job *jobs, int n_jobs, hb_amvp_list *out, void *stream)
{
    if (n_jobs <= 0) return 0;
    k_amvp<<<(n_jobs + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(units, units_w, w, h, jobs, n_jobs, out);
    return static_cast<int>(cudaGetLastError());
}
extern "C" int hbk_amvp_fill(const hb_unit_info *units, int units_w, int w, int h, hbd_me_job *jobs, int n_jobs, int size, void *stream)
{
    if (n_jobs <= 0) return 0;
    k_amvp_fill<<<(n_jobs + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(units, units_w, w, h, jobs, n_jobs, size);
    return static_cast<int>(cudaGetLastError());
}
extern "C" int hbk_merge_cands(const hb_unit_info *units, int units_w, int w, int h, const hb_amvp_job *jobs, int n_jobs, int max_cands, hb_mv *out, void *stream)
{
    if (n_jobs <= 0) return 0;
    k_merge_cands<<<(n_jobs + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(units, units_w, w, h, jobs, n_jobs, max_cands, out);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_deblock_strengths(const hb_unit_info *units, int units_w, int w, int h, uint8_t *bs_ver, uint8_t *bs_hor, uint8_t *qp, void *stream)
{
    const int uw = w >> 2, uh = h >> 2;
    DbkPics pics;
    memset(&pics, 0, sizeof pics);
    k_deblock_strengths<<<(uw * uh + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(units, nullptr, pics, units_w, uw, uh, bs_ver, bs_hor, qp);
    return static_cast<int>(cudaGetLastError());
}
extern "C" int hbk_deblock_strengths_b(const hb_unit_info *units, const hb_unit_l1 *units1, const int32_t *pic_l0, int n_l0, const int32_t *pic_l1, int n_l1,
                                       int units_w, int w, int h, uint8_t *bs_ver, uint8_t *bs_hor, uint8_t *qp, void *stream)
{
    const int uw = w >> 2, uh = h >> 2;
    DbkPics pics;
    for (int i = 0; i < 16; i++) { pics.l0[i] = i < n_l0 ? pic_l0[i] : -1; pics.l1[i] = i < n_l1 ? pic_l1[i] : -1; }
    k_deblock_strengths<<<(uw * uh + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(units, units1, pics, units_w, uw, uh, bs_ver, bs_hor, qp);
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
    const hbd_plane ps = hbd_pick_plane(src, comp), pd = hbd_pick_plane(dst, comp);
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

namespace {
// ---- the arithmetic half of the SAO decision on the device (sao_derive_offsets hmr_sao.c:480 with est_iter_offset :445,
// sao_get_distortion :620, 8-bit video): one warp per (CTU, component).  Lanes 0..19 own an (edge type, class) pair, every lane
// owns a band; the costs are IEEE doubles combined in the reference's order with explicit round-to-nearest operations (no fused
// multiply-add), so the choices are the host function's bit for bit.
struct SaoIter { int off; long long dist; double cost; };

__device__ __forceinline__ SaoIter sao_class_offset(int cnt, int dif, double lambda, bool is_bo, int sign_rule)
{
    SaoIter r; r.off = 0; r.dist = 0; r.cost = lambda;
    if (cnt == 0) return r;
    const double x = __ddiv_rn(static_cast<double>(dif), static_cast<double>(cnt));
    int o = x >= 0 ? static_cast<int>(__dadd_rn(x, 0.5)) : static_cast<int>(__dadd_rn(x, -0.5));
    o = min(max(o, -7), 7);
    if (sign_rule > 0 && o < 0) o = 0;                   // valleys are only raised
    if (sign_rule < 0 && o > 0) o = 0;                   // peaks only lowered
    double best = lambda;
    for (int it = o; it != 0; it += (it > 0) ? -1 : 1) {
        int bits = abs(it) + (is_bo ? 2 : 1);
        if (abs(it) == 7) bits--;
        const long long d = static_cast<long long>(cnt) * it * it - 2ll * dif * it;
        const double c = __dadd_rn(static_cast<double>(d), __dmul_rn(lambda, static_cast<double>(bits)));
        if (c < best) { best = c; r.off = it; r.dist = d; r.cost = c; }
    }
    return r;
}

__global__ void __launch_bounds__(128) k_sao_derive(const hb_sao_stats *stats, int n_units, double lam_y, double lam_u, double lam_v, hb_sao_candidate *out)
{
    const int unit = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (unit >= n_units) return;
    const hb_sao_stats &st = stats[unit];
    const int comp = unit % 3;
    const double lambda = comp == 0 ? lam_y : (comp == 1 ? lam_u : lam_v);
    hb_sao_candidate *o = out + static_cast<size_t>(unit) * 5;
    // ---- edge types: lane = type * 5 + class
    {
        const int type = min(lane / 5, 3), cls = lane % 5;
        const bool act = lane < 20 && cls != 2;
        const int cnt = act ? st.eo_count[type][cls] : 0, dif = act ? st.eo_diff[type][cls] : 0;
        const SaoIter r = sao_class_offset(cnt, dif, lambda, false, cls < 2 ? 1 : -1);
        const long long cd = static_cast<long long>(cnt) * r.off * r.off - 2ll * dif * r.off;
        long long sum = cd;
        int offs[5];
#pragma unroll
        for (int k = 0; k < 5; k++) offs[k] = __shfl_sync(HB_FULL_MASK, r.off, (lane / 5) * 5 + k);
#pragma unroll
        for (int k = 1; k < 5; k++) sum += __shfl_down_sync(HB_FULL_MASK, cd, k);
        if (lane < 20 && cls == 0) {
            hb_sao_candidate c;
            c.dist = sum; c.offset[0] = static_cast<int8_t>(offs[0]); c.offset[1] = static_cast<int8_t>(offs[1]);
            c.offset[2] = static_cast<int8_t>(offs[3]); c.offset[3] = static_cast<int8_t>(offs[4]);
            c.band = 0; c.reserved[0] = c.reserved[1] = c.reserved[2] = 0;
            o[type] = c;
        }
    }
    // ---- band type: lane = band
    {
        const int cnt = st.bo_count[lane], dif = st.bo_diff[lane];
        const SaoIter r = sao_class_offset(cnt, dif, lambda, true, 0);
        double c = r.cost;                               // the cost of four consecutive bands, summed in the reference's order
        c = __dadd_rn(c, __shfl_down_sync(HB_FULL_MASK, r.cost, 1));
        c = __dadd_rn(c, __shfl_down_sync(HB_FULL_MASK, r.cost, 2));
        c = __dadd_rn(c, __shfl_down_sync(HB_FULL_MASK, r.cost, 3));
        if (lane > 28) c = 1.0e300;
        int band = lane;                                 // first minimum wins
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double oc = __shfl_xor_sync(HB_FULL_MASK, c, d);
            const int ob = __shfl_xor_sync(HB_FULL_MASK, band, d);
            if (oc < c || (oc == c && ob < band)) { c = oc; band = ob; }
        }
        const long long cd = static_cast<long long>(cnt) * r.off * r.off - 2ll * dif * r.off;
        long long sum = 0;
        int offs[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { sum += __shfl_sync(HB_FULL_MASK, cd, band + k); offs[k] = __shfl_sync(HB_FULL_MASK, r.off, band + k); }
        if (lane == 0) {
            hb_sao_candidate cnd;
            cnd.dist = sum;
#pragma unroll
            for (int k = 0; k < 4; k++) cnd.offset[k] = static_cast<int8_t>(offs[k]);
            cnd.band = static_cast<int8_t>(band); cnd.reserved[0] = cnd.reserved[1] = cnd.reserved[2] = 0;
            o[4] = cnd;
        }
    }
}
}  // namespace

extern "C" int hbk_sao_derive(const hb_sao_stats *stats, int n_units, const double lambda[3], hb_sao_candidate *out, void *stream)
{
    if (n_units <= 0) return 0;
    k_sao_derive<<<(n_units + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(stats, n_units, lambda[0], lambda[1], lambda[2], out);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_sao_stats(const hbd_frame *org, const hbd_frame *rec, int ctu_cols, int n_ctus, hb_sao_stats *out, void *stream)
{
    if (n_ctus <= 0) return 0;
    k_sao_stats<<<dim3(n_ctus, 3), kSaoWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(*org, *rec, ctu_cols, out);
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

namespace {
// ---- one cost record per coding unit and pass (hb_cu_cost): the sums a choice between partition depths needs, so that the
// per-TU tables need not cross the link
__global__ void __launch_bounds__(256) k_pack_cu_costs(const hbd_cu_pack_args a)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= a.first[5]) return;
    int p = 0;
    while (i >= a.first[p + 1]) p++;
    const int d = min(p, 3), cu = 64 >> d, k = i - a.first[p];
    const int cx = k % a.grid_w[d], cy = k / a.grid_w[d];
    // two rounds of independent loads (all TU indices, then all records) instead of a dependent chain per TU
    int idx[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const hbd_gather_pc &pc = a.pc[c ? d : p][c];
        const int per = (c ? cu / 2 : cu) / pc.tu;                      // 1 or 2 TUs per CU side
#pragma unroll
        for (int t = 0; t < 4; t++)
            idx[c][t] = t < per * per ? __ldg(pc.tu_index + (cy * per + t / per) * pc.grid_w + cx * per + t % per) : -1;
    }
    int2 rec[3][4];                                      // {sum, ssd}: the first half of hb_tu_result (the table block guarantees 8-byte alignment only)
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int t = 0; t < 4; t++)
            rec[c][t] = idx[c][t] >= 0 ? __ldg(reinterpret_cast<const int2 *>(a.pc[c ? d : p][c].res + idx[c][t])) : make_int2(0, 0);
    uint32_t ssd = 0, sum = 0, cbf = 0;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int t = 0; t < 4; t++) {
            ssd += static_cast<uint32_t>(rec[c][t].y); sum += static_cast<uint32_t>(rec[c][t].x);
            if (rec[c][t].x > 0) cbf |= 1u << (4 * c + t);
        }
    hb_cu_cost o;
    o.ssd = ssd; o.sum = sum; o.cbf = static_cast<uint16_t>(cbf); o.reserved = 0;
    a.out[i] = o;
}
}  // namespace

extern "C" int hbk_pack_cu_costs(const hbd_cu_pack_args *a, void *stream)
{
    if (a->first[5] <= 0) return 0;
    k_pack_cu_costs<<<(a->first[5] + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_units_from_selection(const hbd_units_args *a, void *stream)
{
    const int n = a->uw * a->uh;
    if (n <= 0) return 0;
    k_units_from_selection<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
    return static_cast<int>(cudaGetLastError());
}

namespace {
// ---- the levels of the chosen passes in the layout the reference's entropy coder reads (ctu->coeff_wnd, hmr_encoder_lib.c:2945,
// encode_residual hmr_arithmetic_encoding.c:1087): per CTU three 1-D windows (64*64 luma, 32*32 U, 32*32 V int16); a transform unit's
// N*N levels lie row-major at offset abs_index << 4 (luma) / (abs_index << 4) >> 2 (chroma), abs_index = z-order number of the
// unit's first 4x4 luma block inside the CTU (hmr_motion_inter.c:73, :174).  Units without levels are zero.  One CTA per CTU.
__device__ __forceinline__ int zorder4(int ux, int uy)
{
    int a = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) a |= ((ux >> b) & 1) << (2 * b) | ((uy >> b) & 1) << (2 * b + 1);
    return a;
}
__global__ void __launch_bounds__(256) k_coeff_wnd(const hbd_gather_args a, int16_t *out)
{
    const int ctu = blockIdx.x, tid = threadIdx.x;
    const int cx = ctu % a.ctu_cols, cy = ctu / a.ctu_cols;
    const int sel = a.sel[ctu];
    int16_t *dst_ctu = out + static_cast<size_t>(ctu) * (64 * 64 + 2 * 32 * 32);
    for (int c = 0; c < 3; c++) {
        const int pass = c ? min(sel, 3) : sel;
        const hbd_gather_pc pc = a.pc[pass][c];
        const int cs = c ? 32 : 64, tpr = cs / pc.tu, nn = pc.tu * pc.tu, lum = c ? pc.tu * 2 : pc.tu;      // lum: the unit's side in luma samples
        int16_t *dst = dst_ctu + (c == 0 ? 0 : c == 1 ? 64 * 64 : 64 * 64 + 32 * 32);
        // (unit, 8-level chunk) pairs over the whole window: every level of the window is written, coded or zero
        for (int q = tid; q < cs * cs / 8; q += 256) {
            const int unit = q / (nn / 8), chunk = q % (nn / 8);
            const int tx = cx * tpr + unit % tpr, ty = cy * tpr + unit / tpr;
            int idx = -1;
            if (tx < pc.grid_w && ty < pc.grid_h) idx = pc.tu_index[ty * pc.grid_w + tx];
            if (idx >= 0 && pc.res[idx].sum <= 0) idx = -1;
            const int abs_index = zorder4((unit % tpr) * lum / 4, (unit / tpr) * lum / 4);
            const int off = c ? (abs_index << 4) >> 2 : abs_index << 4;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (idx >= 0) v = __ldg(reinterpret_cast<const uint4 *>(pc.coeff + static_cast<size_t>(idx) * nn) + chunk);
            reinterpret_cast<uint4 *>(dst + off)[chunk] = v;
        }
    }
}
}  // namespace

extern "C" int hbk_coeff_wnd(const hbd_gather_args *a, int n_ctus, int16_t *out, void *stream)
{
    if (n_ctus <= 0) return 0;
    k_coeff_wnd<<<n_ctus, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a, out);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_gather(const hbd_gather_args *a, int n_ctus, void *stream)
{
    if (n_ctus <= 0) return 0;
    k_gather<<<n_ctus, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
    return static_cast<int>(cudaGetLastError());
}
