// hb_kernels_mc.cu -- motion compensation (hmr_motion_compensation_luma / _chroma, hmr_motion_inter.c:1779/:1860,
// uni- and bi-prediction) on resident u8 planes, plus frame maintenance (ingest, border replication, int16 -> u8 narrowing).
//
// k_mc_chroma<RS, BI> / k_mc_luma<RS, BI>: one THREAD predicts a 4-wide, RS-tall strip of a plane entirely in registers: the 4-tap
// (luma: 8-tap, two dp4a per sum) horizontal filter of four neighbours is dp4a on funnel-shifted words of the resident reference,
// the vertical filter runs over a rotating 4-row (luma: 8-row) window, four clipped samples leave as one 32-bit store.  BI: two
// lists and their weighted average in the same pass.  One arithmetic path serves the three
// branches of the reference: with taps {0,64,0,0} for a zero fraction the two-stage result equals the single-stage one
// ((64 s + 2048) >> 12 == (s + 32) >> 6), and the 14-bit offset cancels because the taps sum to 64.
#include "hb_shim.h"
#include "hb_dev_common.cuh"

namespace {

// ---- chroma (hmr_motion_compensation_chroma, hmr_motion_inter.c:1860; vectors in eighth-sample units :1863-1867)
__constant__ int8_t c_ctap[8][4] = { {0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4},
                                     {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2} };

struct McCArgs {
    hbd_plane ref[2], pred[2];
    hbd_plane ref1[2];     // second list (bi-prediction)
    const hbd_mc_pu *pus;
    const hb_me_result *mvsrc, *mvsrc1;
    int total;             // work items: n_pus * 2 planes * segments * strips
    int lg_strips, lg_segs;
};

// BI: the block is the average of two 14-bit predictions (weighted_average_motion, hmr_motion_inter.c:2903).  With X the sum
// of vertical taps x horizontal sums, a list's 14-bit value is (X >> 6) - 8192 in all three branches of the reference
// (is_last = 0), so the average (a + b + 64 + 2*8192) >> 7 becomes ((X0 >> 6) + (X1 >> 6) + 64) >> 7.
template <int RS, bool BI>
__global__ void __launch_bounds__(256) k_mc_chroma(const McCArgs a)
{
    const int item = blockIdx.x * 256 + threadIdx.x;
    if (item >= a.total) return;
    // adjacent threads: adjacent strips of a row segment, then the segments, then the two planes of a PU
    const int strip = item & ((1 << a.lg_strips) - 1);
    int t = item >> a.lg_strips;
    const int seg = t & ((1 << a.lg_segs) - 1);
    t >>= a.lg_segs;
    const int plane = t & 1;
    const hbd_mc_pu pu = a.pus[t >> 1];
    const int x = (pu.x >> 1) + strip * 4, y = (pu.y >> 1) + seg * RS;
    const hbd_plane &pp = a.pred[plane];
    constexpr int NL = BI ? 2 : 1;
    const uint32_t *q[NL];
    uint32_t sh[NL], htap[NL];
    int qstep[NL], tv[NL][4];
#pragma unroll
    for (int l = 0; l < NL; l++) {
        const hb_mv mv = (l ? a.mvsrc1 : a.mvsrc)[pu.mv_idx].mv;
        const hbd_plane &rp = l ? a.ref1[plane] : a.ref[plane];
        const uint8_t *src = rp.org + (y + (mv.y >> 3) - 1) * rp.pitch + x + (mv.x >> 3) - 1;
        sh[l] = (static_cast<uint32_t>(reinterpret_cast<uintptr_t>(src)) & 3u) * 8u;
        q[l] = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(src) & ~uintptr_t(3));
        qstep[l] = rp.pitch >> 2;
        const int8_t *th = c_ctap[mv.x & 7], *tvp = c_ctap[mv.y & 7];
        htap[l] = hb_pack4(th[0] & 255, th[1] & 255, th[2] & 255, th[3] & 255);
#pragma unroll
        for (int k = 0; k < 4; k++) tv[l][k] = tvp[k];
    }
    uint8_t *dst = pp.org + y * pp.pitch + x;
    int h[NL][4][4];                                        // rotating windows of horizontal sums: [list][row & 3][column]
#pragma unroll
    for (int r = 0; r < RS + 3; r++) {
#pragma unroll
        for (int l = 0; l < NL; l++) {
            const uint32_t w0 = __ldg(q[l]), w1 = __ldg(q[l] + 1), w2 = __ldg(q[l] + 2);
            q[l] += qstep[l];
            const uint32_t x0 = __funnelshift_r(w0, w1, sh[l]), x1 = __funnelshift_r(w1, w2, sh[l]);     // samples x-1 .. x+6
            h[l][r & 3][0] = hb_dp4a_us(x0, htap[l], 0);
#pragma unroll
            for (int c = 1; c < 4; c++) h[l][r & 3][c] = hb_dp4a_us(__funnelshift_r(x0, x1, 8 * c), htap[l], 0);
        }
        if (r >= 3) {
            int o[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int s[NL];
#pragma unroll
                for (int l = 0; l < NL; l++)
                    s[l] = tv[l][0] * h[l][(r - 3) & 3][c] + tv[l][1] * h[l][(r - 2) & 3][c] + tv[l][2] * h[l][(r - 1) & 3][c] + tv[l][3] * h[l][r & 3][c];
                if constexpr (BI) o[c] = ((s[0] >> 6) + (s[NL - 1] >> 6) + 64) >> 7;
                else o[c] = (s[0] + 2048) >> 12;
            }
            *reinterpret_cast<uint32_t *>(dst) = hb_pack_sat_u8x4(o[0], o[1], o[2], o[3]);
            dst += pp.pitch;
        }
    }
}

// ---- luma (hmr_motion_compensation_luma, hmr_motion_inter.c:1779): the same thread-per-strip form with the 8-tap filters -- two
// dp4a per horizontal sum, an 8-row rotating window per list.  BI: is_bi_predict = 1 for both lists + weighted_average_motion.
__constant__ uint32_t c_ltap4[4][2] = { { htap4(0, 0), htap4(0, 1) }, { htap4(1, 0), htap4(1, 1) }, { htap4(2, 0), htap4(2, 1) }, { htap4(3, 0), htap4(3, 1) } };
__constant__ int8_t c_ltap[4][8] = { {0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1} };

struct McLArgs {
    hbd_plane ref0, ref1, pred;
    const hbd_mc_pu *pus;
    const hb_me_result *mvsrc0, *mvsrc1;
    int total, lg_strips, lg_segs;
};

template <int RS, bool BI>
__global__ void __launch_bounds__(128) k_mc_luma(const McLArgs a)
{
    constexpr int NL = BI ? 2 : 1;
    const int item = blockIdx.x * 128 + threadIdx.x;
    if (item >= a.total) return;
    const int strip = item & ((1 << a.lg_strips) - 1);
    int t = item >> a.lg_strips;
    const int seg = t & ((1 << a.lg_segs) - 1);
    const hbd_mc_pu pu = a.pus[t >> a.lg_segs];
    const int x = pu.x + strip * 4, y = pu.y + seg * RS;
    const uint32_t *q[NL];
    uint32_t sh[NL], tlo[NL], thi[NL];
    int qstep[NL], tv[NL][8];
#pragma unroll
    for (int l = 0; l < NL; l++) {
        const hb_mv mv = (l ? a.mvsrc1 : a.mvsrc0)[pu.mv_idx].mv;
        const hbd_plane &rp = l ? a.ref1 : a.ref0;
        const uint8_t *src = rp.org + (y + (mv.y >> 2) - 3) * rp.pitch + x + (mv.x >> 2) - 3;
        sh[l] = (static_cast<uint32_t>(reinterpret_cast<uintptr_t>(src)) & 3u) * 8u;
        q[l] = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(src) & ~uintptr_t(3));
        qstep[l] = rp.pitch >> 2;
        tlo[l] = c_ltap4[mv.x & 3][0]; thi[l] = c_ltap4[mv.x & 3][1];
#pragma unroll
        for (int k = 0; k < 8; k++) tv[l][k] = c_ltap[mv.y & 3][k];
    }
    uint8_t *dst = a.pred.org + y * a.pred.pitch + x;
    int h[NL][8][4];                                        // rotating windows: [list][row & 7][column]
#pragma unroll
    for (int r = 0; r < RS + 7; r++) {
#pragma unroll
        for (int l = 0; l < NL; l++) {
            const uint32_t w0 = __ldg(q[l]), w1 = __ldg(q[l] + 1), w2 = __ldg(q[l] + 2), w3 = __ldg(q[l] + 3);
            q[l] += qstep[l];
            const uint32_t x0 = __funnelshift_r(w0, w1, sh[l]), x1 = __funnelshift_r(w1, w2, sh[l]), x2 = __funnelshift_r(w2, w3, sh[l]);   // samples x-3 .. x+8
            h[l][r & 7][0] = hb_dp4a_us(x1, thi[l], hb_dp4a_us(x0, tlo[l], 0));
#pragma unroll
            for (int c = 1; c < 4; c++)
                h[l][r & 7][c] = hb_dp4a_us(__funnelshift_r(x1, x2, 8 * c), thi[l], hb_dp4a_us(__funnelshift_r(x0, x1, 8 * c), tlo[l], 0));
        }
        if (r >= 7) {
            int o[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int s[NL];
#pragma unroll
                for (int l = 0; l < NL; l++) {
                    s[l] = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++) s[l] += tv[l][k] * h[l][(r - 7 + k) & 7][c];
                }
                if constexpr (BI) o[c] = ((s[0] >> 6) + (s[NL - 1] >> 6) + 64) >> 7;
                else o[c] = (s[0] + 2048) >> 12;           // two-stage == single-stage when a fraction is zero (taps {..,64,..})
            }
            *reinterpret_cast<uint32_t *>(dst) = hb_pack_sat_u8x4(o[0], o[1], o[2], o[3]);
            dst += a.pred.pitch;
        }
    }
}

// ---- SSD between the current picture and a prediction over square blocks of the three planes (ssd16b of the "no residual" branch of
// the merge check, hmr_motion_inter.c:3686-3688): one warp per (block, plane), out[block * 3 + plane]
__global__ void __launch_bounds__(256) k_block_ssd(hbd_frame cur, hbd_frame pred, const hbd_mc_pu *pus, int n_pus, int size, uint32_t *out)
{
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= n_pus * 3) return;
    const int plane = w % 3;
    const hbd_mc_pu pu = pus[w / 3];
    const int n = plane ? size >> 1 : size, x = plane ? pu.x >> 1 : pu.x, y = plane ? pu.y >> 1 : pu.y;
    const hbd_plane pc = hbd_pick_plane(cur, plane), pp = hbd_pick_plane(pred, plane);
    uint32_t acc = 0;
    for (int i = lane; i < (n >> 2) * n; i += 32) {
        const int r = i / (n >> 2), c = (i % (n >> 2)) * 4;
        const uint32_t a = *reinterpret_cast<const uint32_t *>(pc.org + (y + r) * pc.pitch + x + c), b = *reinterpret_cast<const uint32_t *>(pp.org + (y + r) * pp.pitch + x + c);
#pragma unroll
        for (int k = 0; k < 4; k++) { const int d = static_cast<int>((a >> (8 * k)) & 255u) - static_cast<int>((b >> (8 * k)) & 255u); acc += d * d; }
    }
    acc = __reduce_add_sync(HB_FULL_MASK, acc);
    if (lane == 0) out[w] = acc;
}

// ---- border replication of one plane (reference_picture_border_padding_ctu, hmr_encoder_lib.c:1723): every sample
// outside the picture takes the nearest picture sample.
__global__ void k_pad_frame(hbd_frame f)
{
    const hbd_plane p = hbd_pick_plane(f, blockIdx.y);  // one grid row per plane: a single launch pads Y, U and V
    const int pw = p.pad >> 2, Ww = (p.w + 2 * p.pad) >> 2;     // widths and borders are multiples of 4, rows 4-byte aligned: words
    const int n_side = 2 * pw * p.h;                    // left+right strips of the picture rows
    const int n_tb = 2 * p.pad * Ww;                    // top+bottom bands, full padded width
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_side + n_tb; i += gridDim.x * blockDim.x) {
        int px, py;                                     // padded coordinates of the word's first sample
        if (i < n_side) {
            const int r = i / (2 * pw), c = i % (2 * pw);
            py = p.pad + r;
            px = c < pw ? 4 * c : p.w + 4 * c;
        } else {
            const int j = i - n_side, r = j / Ww;
            px = 4 * (j % Ww);
            py = r < p.pad ? r : p.h + r;
        }
        const int x = px - p.pad, sy = min(max(py - p.pad, 0), p.h - 1);
        const uint8_t *row = p.org + sy * p.pitch;
        uint32_t v;
        if (x >= 0 && x < p.w) v = *reinterpret_cast<const uint32_t *>(row + x);
        else v = static_cast<uint32_t>(x < 0 ? row[0] : row[p.w - 1]) * 0x01010101u;
        *reinterpret_cast<uint32_t *>(p.org + (py - p.pad) * p.pitch + x) = v;
    }
}

// ---- dense host layout -> padded planes.  `stage` holds the three planes back to back with tight pitches (what one plain
// 1-D copy from the host delivers at full link speed); every 4-byte group of the padded destination is either inside the
// picture or a replicated edge sample (widths and borders are multiples of 4).  border = 0 writes the picture only.
__global__ void k_ingest_frame(hbd_frame f, const uint8_t *stage, int border)
{
    const hbd_plane p = hbd_pick_plane(f, blockIdx.y);
    const uint8_t *src = stage + (blockIdx.y == 0 ? 0 : f.p[0].w * f.p[0].h + (blockIdx.y == 2 ? f.p[1].w * f.p[1].h : 0));
    const int pad = border ? p.pad : 0;
    const int ww = (p.w + 2 * pad) >> 2, rows = p.h + 2 * pad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ww * rows; i += gridDim.x * blockDim.x) {
        const int r = i / ww, x = (i % ww) * 4 - pad;
        const int y = min(max(r - pad, 0), p.h - 1);
        const uint8_t *row = src + y * p.w;
        uint32_t v;
        if (x >= 0 && x < p.w) v = *reinterpret_cast<const uint32_t *>(row + x);
        else v = static_cast<uint32_t>(x < 0 ? row[0] : row[p.w - 1]) * 0x01010101u;
        *reinterpret_cast<uint32_t *>(p.org + (r - pad) * p.pitch + x) = v;
    }
}

// the same, 16 bytes per thread in padded coordinates (rows start 128-byte aligned): every 8-byte half of a chunk is either inside
// the picture (one 8-byte load from the dense plane; needs width % 16 == 0 so that chroma rows stay 8-byte aligned) or a
// replicated edge sample.  border = 0 visits only the chunks that touch the picture.
__global__ void k_ingest_frame16(hbd_frame f, const uint8_t *stage, int border)
{
    const hbd_plane p = hbd_pick_plane(f, blockIdx.y);
    const uint8_t *src = stage + (blockIdx.y == 0 ? 0 : f.p[0].w * f.p[0].h + (blockIdx.y == 2 ? f.p[1].w * f.p[1].h : 0));
    const int W = p.w + 2 * p.pad;
    const int j0 = border ? 0 : p.pad >> 4, j1 = border ? (W + 15) >> 4 : (p.pad + p.w + 15) >> 4;
    const int r0 = border ? 0 : p.pad, rows = border ? p.h + 2 * p.pad : p.h, cols = j1 - j0;
    uint8_t *row0 = p.org - p.pad * p.pitch - p.pad;                    // first padded row, 128-byte aligned
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cols * rows; i += gridDim.x * blockDim.x) {
        const int r = r0 + i / cols, j = j0 + i % cols;
        const int y = min(max(r - p.pad, 0), p.h - 1);
        const uint8_t *srow = src + y * p.w;
        uint2 h[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int x = 16 * j + 8 * k - p.pad;
            if (x >= 0 && x < p.w) h[k] = *reinterpret_cast<const uint2 *>(srow + x);
            else { const uint32_t v = static_cast<uint32_t>(x < 0 ? srow[0] : srow[p.w - 1]) * 0x01010101u; h[k] = make_uint2(v, v); }
        }
        *reinterpret_cast<uint4 *>(row0 + static_cast<size_t>(r) * p.pitch + 16 * j) = make_uint4(h[0].x, h[0].y, h[1].x, h[1].y);
    }
}

__global__ void k_narrow_plane(const int16_t *src, int src_stride, hbd_plane dst, uint32_t *range_flag)
{
    bool bad = false;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dst.w * dst.h; i += gridDim.x * blockDim.x) {
        const int r = i / dst.w, c = i % dst.w;
        const int v = src[r * src_stride + c];
        bad |= (v < 0 || v > 255);
        dst.org[r * dst.pitch + c] = static_cast<uint8_t>(v);
    }
    if (__any_sync(HB_FULL_MASK, bad) && (threadIdx.x & 31) == 0) atomicOr(range_flag, 1u);
}

}  // namespace

extern "C" int hbk_mc_predict(const hbd_frame *ref, const hbd_frame *pred, int size, const hbd_mc_pu *pus, int n_pus,
                              const hb_me_result *mvsrc, int planes, void *stream)
{
    if (n_pus <= 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (size != 8 && size != 16 && size != 32 && size != 64) return static_cast<int>(cudaErrorInvalidValue);
    if (planes & 1) {
        McLArgs l;
        const int rs = 8;
        l.ref0 = ref->p[0]; l.ref1 = ref->p[0]; l.pred = pred->p[0]; l.pus = pus; l.mvsrc0 = mvsrc; l.mvsrc1 = mvsrc;
        l.lg_strips = 0; while ((4 << l.lg_strips) < size) l.lg_strips++;
        l.lg_segs = 0; while ((rs << l.lg_segs) < size) l.lg_segs++;
        l.total = n_pus << (l.lg_strips + l.lg_segs);
        k_mc_luma<8, false><<<(l.total + 127) / 128, 128, 0, s>>>(l);
    }
    if (planes & 2) {
        McCArgs c;
        const int cs = size / 2, rs = cs < 8 ? 4 : 8;      // chroma block size, rows per work item
        c.ref[0] = ref->p[1]; c.ref[1] = ref->p[2]; c.pred[0] = pred->p[1]; c.pred[1] = pred->p[2];
        c.pus = pus; c.mvsrc = mvsrc;
        c.lg_strips = 0; while ((4 << c.lg_strips) < cs) c.lg_strips++;
        c.lg_segs = 0; while ((rs << c.lg_segs) < cs) c.lg_segs++;
        c.total = n_pus * 2 << (c.lg_strips + c.lg_segs);
        const int grid = (c.total + 255) / 256;
        c.mvsrc1 = nullptr;
        if (rs == 4) k_mc_chroma<4, false><<<grid, 256, 0, s>>>(c);
        else k_mc_chroma<8, false><<<grid, 256, 0, s>>>(c);
    }
    return static_cast<int>(cudaGetLastError());
}

// bi-prediction: pred = average of the 14-bit predictions from (ref0, mvsrc0) and (ref1, mvsrc1), luma and chroma
extern "C" int hbk_mc_predict_bi(const hbd_frame *ref0, const hbd_frame *ref1, const hbd_frame *pred, int size, const hbd_mc_pu *pus, int n_pus,
                                 const hb_me_result *mvsrc0, const hb_me_result *mvsrc1, void *stream)
{
    if (n_pus <= 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (size != 8 && size != 16 && size != 32 && size != 64) return static_cast<int>(cudaErrorInvalidValue);
    {
        McLArgs l;
        const int rs = 8;
        l.ref0 = ref0->p[0]; l.ref1 = ref1->p[0]; l.pred = pred->p[0]; l.pus = pus; l.mvsrc0 = mvsrc0; l.mvsrc1 = mvsrc1;
        l.lg_strips = 0; while ((4 << l.lg_strips) < size) l.lg_strips++;
        l.lg_segs = 0; while ((rs << l.lg_segs) < size) l.lg_segs++;
        l.total = n_pus << (l.lg_strips + l.lg_segs);
        k_mc_luma<8, true><<<(l.total + 127) / 128, 128, 0, s>>>(l);
    }
    {
        McCArgs c;
        const int cs = size / 2, rs = cs < 8 ? 4 : 8;
        c.ref[0] = ref0->p[1]; c.ref[1] = ref0->p[2]; c.ref1[0] = ref1->p[1]; c.ref1[1] = ref1->p[2]; c.pred[0] = pred->p[1]; c.pred[1] = pred->p[2];
        c.pus = pus; c.mvsrc = mvsrc0; c.mvsrc1 = mvsrc1;
        c.lg_strips = 0; while ((4 << c.lg_strips) < cs) c.lg_strips++;
        c.lg_segs = 0; while ((rs << c.lg_segs) < cs) c.lg_segs++;
        c.total = n_pus * 2 << (c.lg_strips + c.lg_segs);
        const int grid = (c.total + 255) / 256;
        if (rs == 4) k_mc_chroma<4, true><<<grid, 256, 0, s>>>(c);
        else k_mc_chroma<8, true><<<grid, 256, 0, s>>>(c);
    }
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_block_ssd(const hbd_frame *cur, const hbd_frame *pred, const hbd_mc_pu *pus, int n_pus, int size, uint32_t *out, void *stream)
{
    if (n_pus <= 0) return 0;
    k_block_ssd<<<(n_pus * 3 + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(*cur, *pred, pus, n_pus, size, out);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_pad_frame(const hbd_frame *f, void *stream)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const hbd_plane &p = f->p[0];
    const int n = (2 * p.pad * p.h + 2 * p.pad * (p.w + 2 * p.pad)) / 4;
    k_pad_frame<<<dim3((n + 255) / 256, 3), 256, 0, s>>>(*f);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_ingest_frame(const hbd_frame *f, const uint8_t *stage, int border, void *stream)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const hbd_plane &p = f->p[0];
    const int pad = border ? p.pad : 0;
    if (p.w % 16 == 0 && p.pad % 16 == 0 && f->p[1].pad % 8 == 0 && p.pitch % 16 == 0 && f->p[1].pitch % 16 == 0) {
        const int n16 = ((p.w + 2 * p.pad) / 16 + 1) * (p.h + 2 * pad);
        k_ingest_frame16<<<dim3(min((n16 + 255) / 256, 148 * 8), 3), 256, 0, s>>>(*f, stage, border);
        return static_cast<int>(cudaGetLastError());
    }
    const int n = ((p.w + 2 * pad) / 4) * (p.h + 2 * pad);
    k_ingest_frame<<<dim3(min((n + 255) / 256, 148 * 8), 3), 256, 0, s>>>(*f, stage, border);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbk_narrow_plane(const int16_t *src, int src_stride, hbd_plane dst, uint32_t *range_flag, void *stream)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n = dst.w * dst.h;
    k_narrow_plane<<<min((n + 255) / 256, 148 * 8), 256, 0, s>>>(src, src_stride, dst, range_flag);
    return static_cast<int>(cudaGetLastError());
}
