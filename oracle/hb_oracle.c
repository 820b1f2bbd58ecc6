/*
 * hb_oracle.c -- CPU restatement of the HomerHEVC ME + interpolation + T/Q hot path.
 * TEST INFRASTRUCTURE ONLY (see hb_oracle.h).  Plain C99, scalar, single-threaded.
 *
 * Citations are to /root/reference/src/homer_lib/<file>:<line>.
 */
#include "hb_oracle.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------------------------
 * Pixel primitives.  hmr_motion_intra.c:51-186.  The SSE4.2 twins (hmr_sse42_functions_pixel.c:462,
 * :728, :817, :919) give the same values on the reachable sample ranges.
 * ------------------------------------------------------------------------------------------ */
uint32_t orc_sad(const int16_t *src, int src_stride, const int16_t *pred, int pred_stride, int size)
{
    int acc = 0;
    for (int y = 0; y < size; y++, src += src_stride, pred += pred_stride)
        for (int x = 0; x < size; x++)
            acc += abs(src[x] - pred[x]);
    return (uint32_t)acc;
}

uint32_t orc_ssd16b(const int16_t *src, int src_stride, const int16_t *pred, int pred_stride, int size)
{
    uint32_t acc = 0;
    for (int y = 0; y < size; y++, src += src_stride, pred += pred_stride)
        for (int x = 0; x < size; x++) {
            int d = src[x] - pred[x];
            acc += (uint32_t)d * (uint32_t)d;
        }
    return acc;
}

/* The table members as the SSE4.2 build computes them on ANY int16 operands (hmr_sse42_functions_pixel.c:330-460 sad, :619-745
 * ssd16b): equal to the plain sums above on 8-bit video, but the encoder also calls them on the wrapped 16-bit predictions of its
 * 64x64 intra mode search (hmr_motion_intra.c:1130), where the SIMD lane arithmetic decides the value.  Differences wrap at 16 bits
 * (psubw) and |-32768| stays 32768 (pabsw); 4/8/16: eight wrapping 16-bit lanes folded 8 -> 4 -> 2, the last two added in 32 bits
 * (sse_128_hacc_i16_ :42); 32: sixteen lanes with unsigned saturation over all rows (:374-410); 64: eight lanes saturating over the
 * eight column groups of each row (:413-449); ssd16b: squares of the wrapped difference modulo 2^32 (pmaddwd / paddd). */
static uint32_t absdiff16(int a, int b)
{
    const int t = (int16_t)(a - b);
    return (uint32_t)(t < 0 ? -t : t);
}
uint32_t orc_sad_sse(const int16_t *src, int src_stride, const int16_t *pred, int pred_stride, int size)
{
    uint32_t lane[16] = { 0 }, total = 0;
    for (int y = 0; y < size; y++) {
        uint32_t row[8] = { 0 };
        for (int x = 0; x < size; x++) {
            const uint32_t d = absdiff16(src[y * src_stride + x], pred[y * pred_stride + x]);
            if (size == 64) row[x & 7] += d;
            else lane[size == 4 ? ((y & 1) * 4 + x) : size == 32 ? ((x >> 4) * 8 + (x & 7)) : (x & 7)] += d;
        }
        if (size == 64) for (int j = 0; j < 8; j++) total += row[j] > 65535u ? 65535u : row[j];
    }
    if (size == 64) return total;
    if (size == 32) { for (int l = 0; l < 16; l++) total += lane[l] > 65535u ? 65535u : lane[l]; return total; }
    return ((lane[0] + lane[4] + lane[2] + lane[6]) & 0xffffu) + ((lane[1] + lane[5] + lane[3] + lane[7]) & 0xffffu);
}
uint32_t orc_ssd16b_sse(const int16_t *src, int src_stride, const int16_t *pred, int pred_stride, int size)
{
    uint32_t acc = 0;
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++) {
            const int t = (int16_t)(src[y * src_stride + x] - pred[y * pred_stride + x]);
            acc += (uint32_t)(t * t);
        }
    return acc;
}

void orc_predict(const int16_t *orig, int orig_stride, const int16_t *pred, int pred_stride,
                 int16_t *resid, int resid_stride, int size)
{
    for (int y = 0; y < size; y++, orig += orig_stride, pred += pred_stride, resid += resid_stride)
        for (int x = 0; x < size; x++)
            resid[x] = (int16_t)(orig[x] - pred[x]);
}

void orc_reconst(const int16_t *pred, int pred_stride, const int16_t *resid, int resid_stride,
                 int16_t *dec, int dec_stride, int size)
{
    for (int y = 0; y < size; y++, pred += pred_stride, resid += resid_stride, dec += dec_stride)
        for (int x = 0; x < size; x++)
            dec[x] = (int16_t)clampi(resid[x] + pred[x], 0, 255);
}

/* ------------------------------------------------------------------------------------------
 * Interpolation.  Taps: hmr_motion_inter.c:240-258.  One pass = one call with (is_first,is_last):
 *   (1,0) 8-bit -> 14-bit minus 8192      (1,1) 8-bit -> 8-bit, rounded and clipped
 *   (0,1) 14-bit -> 8-bit, clipped        (0,0) 14-bit -> 14-bit
 * Every result goes through an int16 before the clip, as in the reference ("short val").
 * ------------------------------------------------------------------------------------------ */
static const int8_t k_luma_taps[4][8] = {
    { 0, 0,   0, 64,  0,   0, 0,  0 },
    {-1, 4, -10, 58, 17,  -5, 1,  0 },
    {-1, 4, -11, 40, 40, -11, 4, -1 },
    { 0, 1,  -5, 17, 58, -10, 4, -1 },
};
static const int8_t k_chroma_taps[8][4] = {
    { 0, 64,  0,  0 }, {-2, 58, 10, -2 }, {-4, 54, 16, -2 }, {-6, 46, 28, -4 },
    {-4, 36, 36, -4 }, {-4, 28, 46, -6 }, {-2, 16, 54, -4 }, {-2, 10, 58, -2 },
};

/* fraction 0 of the luma path: hmr_motion_inter.c:261 (filter_copy) */
static void copy_pass(const int16_t *src, int src_stride, int16_t *dst, int dst_stride,
                      int width, int height, int is_first, int is_last)
{
    const int shift = 14 - 8;
    for (int y = 0; y < height; y++, src += src_stride, dst += dst_stride) {
        for (int x = 0; x < width; x++) {
            if (is_first == is_last) {
                dst[x] = src[x];
            } else if (is_first) {
                int16_t v = (int16_t)(src[x] << shift);
                dst[x] = (int16_t)(v - (int16_t)8192);
            } else {
                int16_t v = src[x];
                v = (int16_t)((v + (int16_t)(8192 + (1 << (shift - 1)))) >> shift);
                dst[x] = (int16_t)clampi(v, 0, 255);
            }
        }
    }
}

/* hmr_motion_inter.c:312 (8 taps) and :878 (4 taps) share this shape */
static void filter_pass(const int16_t *src, int src_stride, int16_t *dst, int dst_stride,
                        const int8_t *taps, int ntaps, int width, int height,
                        int is_vertical, int is_first, int is_last)
{
    const int head_room = 14 - 8;
    const int step = is_vertical ? src_stride : 1;
    int shift = 6, offset;
    if (is_last) {
        shift += is_first ? 0 : head_room;
        offset = 1 << (shift - 1);
        offset += is_first ? 0 : (8192 << 6);
    } else {
        shift -= is_first ? head_room : 0;
        offset = is_first ? -(8192 << shift) : 0;
    }
    src -= (ntaps / 2 - 1) * step;
    for (int y = 0; y < height; y++, src += src_stride, dst += dst_stride) {
        for (int x = 0; x < width; x++) {
            int sum = 0;
            for (int k = 0; k < ntaps; k++)
                sum += src[x + k * step] * taps[k];
            int16_t v = (int16_t)((sum + offset) >> shift);
            if (is_last)
                v = (int16_t)clampi(v, 0, 255);
            dst[x] = v;
        }
    }
}

void orc_interpolate_luma(const int16_t *src, int src_stride, int16_t *dst, int dst_stride,
                          int fraction, int width, int height, int is_vertical, int is_first, int is_last)
{
    if (fraction == 0)
        copy_pass(src, src_stride, dst, dst_stride, width, height, is_first, is_last);
    else
        filter_pass(src, src_stride, dst, dst_stride, k_luma_taps[fraction], 8, width, height,
                    is_vertical, is_first, is_last);
}

void orc_interpolate_chroma(const int16_t *src, int src_stride, int16_t *dst, int dst_stride,
                            int fraction, int width, int height, int is_vertical, int is_first, int is_last)
{
    /* the plain-C reference runs the {0,64,0,0} filter for fraction 0 (hmr_motion_inter.c:878) */
    filter_pass(src, src_stride, dst, dst_stride, k_chroma_taps[fraction], 4, width, height,
                is_vertical, is_first, is_last);
}

/* ------------------------------------------------------------------------------------------
 * Transforms.  hmr_transform.c:514/:553 are HM's partial butterflies; a butterfly is an exact
 * refactoring of the integer matrix product, so the restatement is the matrix product itself with
 * the same stage shifts, int16 truncation (forward) and int16 clipping (inverse).
 * The HEVC core matrix is generated from its 32 distinct magnitudes (first column of the 32-point
 * matrix, hmr_transform.c:91-128); smaller sizes are its row-subsampled top-left corners.
 * ------------------------------------------------------------------------------------------ */
static const int8_t k_dct_mag[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67,
                                      64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13,  9,  4, 0 };
static const int8_t k_dst4[4][4] = { { 29, 55, 74, 84 }, { 74, 74, 0, -74 }, { 84, -29, -74, 55 }, { 55, -84, 74, -29 } };

static int dct_coef(int n, int k, int j)          /* T_n[k][j] */
{
    int m = (k * (32 / n) * (2 * j + 1)) & 127;    /* angle index: cos(m*pi/64) */
    if (m > 64) m = 128 - m;
    return m <= 32 ? k_dct_mag[m] : -k_dct_mag[64 - m];
}
static int tr_coef(int n, int is_dst, int k, int j) { return is_dst ? k_dst4[k][j] : dct_coef(n, k, j); }

static int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

void orc_transform(int bit_depth, const int16_t *block, int block_stride, int16_t *coeff, int n, int is_dst)
{
    int16_t tmp[32 * 32];
    const int lg = ilog2(n);
    const int shift1 = lg - 2 + 1 + bit_depth - 8, shift2 = lg - 2 + 8;
    /* stage 1: rows of the residual -> tmp[k][row] */
    for (int r = 0; r < n; r++)
        for (int k = 0; k < n; k++) {
            int s = 0;
            for (int j = 0; j < n; j++) s += tr_coef(n, is_dst, k, j) * block[r * block_stride + j];
            tmp[k * n + r] = (int16_t)((s + (1 << (shift1 - 1))) >> shift1);
        }
    /* stage 2: rows of tmp -> coeff[k][row] */
    for (int r = 0; r < n; r++)
        for (int k = 0; k < n; k++) {
            int s = 0;
            for (int j = 0; j < n; j++) s += tr_coef(n, is_dst, k, j) * tmp[r * n + j];
            coeff[k * n + r] = (int16_t)((s + (1 << (shift2 - 1))) >> shift2);
        }
}

void orc_itransform(int bit_depth, int16_t *block, int block_stride, const int16_t *coeff, int n, int is_dst)
{
    int16_t tmp[32 * 32];
    const int shift1 = 7, shift2 = 12 - (bit_depth - 8);
    /* stage 1: columns of coeff -> rows of tmp */
    for (int c = 0; c < n; c++)
        for (int j = 0; j < n; j++) {
            int s = 0;
            for (int k = 0; k < n; k++) s += tr_coef(n, is_dst, k, j) * coeff[k * n + c];
            tmp[c * n + j] = (int16_t)clampi((s + (1 << (shift1 - 1))) >> shift1, -32768, 32767);
        }
    for (int c = 0; c < n; c++)
        for (int j = 0; j < n; j++) {
            int s = 0;
            for (int k = 0; k < n; k++) s += tr_coef(n, is_dst, k, j) * tmp[k * n + c];
            block[c * block_stride + j] = (int16_t)clampi((s + (1 << (shift2 - 1))) >> shift2, -32768, 32767);
        }
}

/* ------------------------------------------------------------------------------------------
 * Tables.  Scans: hmr_tables.c:62-196.  Quant pyramids from the default HEVC scaling lists:
 * hmr_tables.c:199-250, hmr_tables.h:53-85, wiring hmr_encoder_lib.c:103-140.
 * ------------------------------------------------------------------------------------------ */
static const uint8_t k_sl_intra8[64] = {
    16, 16, 16, 16, 17, 18, 21, 24, 16, 16, 16, 16, 17, 19, 22, 25, 16, 16, 17, 18, 20, 22, 25, 29,
    16, 16, 18, 21, 24, 27, 31, 36, 17, 17, 20, 24, 30, 35, 41, 47, 18, 19, 22, 27, 35, 44, 54, 65,
    21, 22, 25, 31, 41, 54, 70, 88, 24, 25, 29, 36, 47, 65, 88, 115 };
static const uint8_t k_sl_inter8[64] = {
    16, 16, 16, 16, 17, 18, 20, 24, 16, 16, 16, 17, 18, 20, 24, 25, 16, 16, 17, 18, 20, 24, 25, 28,
    16, 17, 18, 20, 24, 25, 28, 33, 17, 18, 20, 24, 25, 28, 33, 41, 18, 20, 24, 25, 28, 33, 41, 54,
    20, 24, 25, 28, 33, 41, 54, 71, 24, 25, 28, 33, 41, 54, 71, 91 };
static const int k_qscale[6] = { 26214, 23302, 20560, 18396, 16384, 14564 };
static const int k_iqscale[6] = { 40, 45, 51, 57, 64, 72 };

const uint8_t orc_chroma_qp_table[58] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
    29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51 };

/* up-right diagonal order of a w x w grid, entries are row*w+col */
static void simple_diag(uint32_t *out, int w)
{
    int pos = 0;
    for (int line = 0; pos < w * w; line++) {
        int row = line, col = 0;
        while (row >= w) { row--; col++; }
        for (; row >= 0 && col < w; row--, col++)
            out[pos++] = (uint32_t)(row * w + col);
    }
}

static void build_scans(orc_tables *t)
{
    uint32_t diag8[64];
    simple_diag(diag8, 8);
    for (int lv = 0; lv < 5; lv++) {                 /* sizes 2,4,8,16,32 */
        const int w = 2 << lv, cgs = w >> 2;
        for (int m = 0; m < 4; m++) t->scan[m][lv] = (uint32_t *)calloc((size_t)w * w, sizeof(uint32_t));
        uint32_t *hor = t->scan[1][lv], *ver = t->scan[2][lv], *diag = t->scan[3][lv];
        if (w <= 4) {
            simple_diag(diag, w);
        } else {
            uint32_t cg_order[64], in_cg[16];
            if (w == 32) memcpy(cg_order, diag8, sizeof diag8); else simple_diag(cg_order, cgs);
            simple_diag(in_cg, 4);
            for (int b = 0; b < cgs * cgs; b++) {
                const int by = cg_order[b] / cgs, bx = cg_order[b] % cgs;
                for (int i = 0; i < 16; i++)
                    diag[16 * b + i] = (uint32_t)((4 * by + in_cg[i] / 4) * w + 4 * bx + in_cg[i] % 4);
            }
        }
        if (w > 2) {
            int n = 0;
            for (int by = 0; by < cgs; by++) for (int bx = 0; bx < cgs; bx++)
                for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
                    hor[n++] = (uint32_t)((4 * by + y) * w + 4 * bx + x);
            n = 0;
            for (int bx = 0; bx < cgs; bx++) for (int by = 0; by < cgs; by++)
                for (int x = 0; x < 4; x++) for (int y = 0; y < 4; y++)
                    ver[n++] = (uint32_t)((4 * by + y) * w + 4 * bx + x);
        } else {
            int n = 0;
            for (int y = 0; y < w; y++) for (int x = 0; x < w; x++) hor[n++] = (uint32_t)(y * w + x);
            n = 0;
            for (int x = 0; x < w; x++) for (int y = 0; y < w; y++) ver[n++] = (uint32_t)(y * w + x);
        }
    }
}

static void build_quant(orc_tables *t)
{
    static const int n_lists[4] = { 6, 6, 6, 2 };
    for (int sm = 0; sm < 4; sm++) {
        const int w = 4 << sm, side = w < 8 ? w : 8, ratio = w / side;
        for (int list = 0; list < n_lists[sm]; list++) {
            const int inter = (sm == 3) ? (list >= 1) : (list >= 3);
            for (int rem = 0; rem < 6; rem++) {
                int32_t *q = (int32_t *)malloc(sizeof(int32_t) * (size_t)w * w);
                int32_t *dq = (int32_t *)malloc(sizeof(int32_t) * (size_t)w * w);
                for (int y = 0; y < w; y++)
                    for (int x = 0; x < w; x++) {
                        int m = 16;
                        if (sm > 0) m = (inter ? k_sl_inter8 : k_sl_intra8)[8 * (y / ratio) + x / ratio];
                        q[y * w + x] = (k_qscale[rem] << 4) / m;
                        dq[y * w + x] = k_iqscale[rem] * m;
                    }
                if (ratio > 1) { q[0] = (k_qscale[rem] << 4) / 16; dq[0] = k_iqscale[rem] * 16; }
                t->quant[sm][list][rem] = q;
                t->dequant[sm][list][rem] = dq;
            }
        }
    }
    for (int rem = 0; rem < 6; rem++) {              /* hmr_encoder_lib.c:135-140: 32x32 list 3 aliases list 1 */
        t->quant[3][3][rem] = t->quant[3][1][rem];
        t->dequant[3][3][rem] = t->dequant[3][1][rem];
    }
}

orc_tables *orc_tables_create(void)
{
    orc_tables *t = (orc_tables *)calloc(1, sizeof *t);
    build_scans(t);
    build_quant(t);
    return t;
}

void orc_tables_destroy(orc_tables *t)
{
    if (!t) return;
    for (int m = 0; m < 4; m++) for (int lv = 0; lv < 7; lv++) free(t->scan[m][lv]);
    for (int sm = 0; sm < 4; sm++) for (int l = 0; l < 6; l++) for (int r = 0; r < 6; r++) {
        if (sm == 3 && l == 3) continue;
        free(t->quant[sm][l][r]); free(t->dequant[sm][l][r]);
    }
    free(t);
}

const uint32_t *orc_tables_scan(const orc_tables *t, int mode, int log2n) { return t->scan[mode][log2n - 1]; }
const int32_t *orc_tables_quant(const orc_tables *t, int log2n, int list, int rem) { return t->quant[log2n - 2][list][rem]; }
const int32_t *orc_tables_dequant(const orc_tables *t, int log2n, int list, int rem) { return t->dequant[log2n - 2][list][rem]; }

/* ------------------------------------------------------------------------------------------
 * Quantisation.  Target = the SSE4.2 function the reference selects on x86
 * (hmr_sse42_functions_quant.c:34-133): rounding offset 171 on I slices and 85 otherwise (:47),
 * |level| saturated to int16 by a signed pack before the sign is put back (:69), deltaU saturated
 * likewise (:70), ac_sum = sum of unsaturated levels, then sign-data hiding when enabled and sum >= 2.
 * ------------------------------------------------------------------------------------------ */
static inline int16_t sat16(int32_t v) { return (int16_t)clampi(v, -32768, 32767); }

void orc_quant(const orc_tables *t, const int16_t *src, int16_t *dst, int16_t *delta_u,
               int scan_mode, int log2n, int comp, int is_intra, int is_islice, int sign_hiding,
               int per, int rem, int *ac_sum)
{
    const int n = 1 << log2n;
    const int32_t *q = orc_tables_quant(t, log2n, (is_intra ? 0 : 3) + comp, rem);
    const int qbits = 14 + per + (15 - 8 - log2n);
    const int32_t add = (int32_t)((uint32_t)(is_islice ? 171 : 85) << (qbits - 9));
    int32_t sum = 0;
    for (int i = 0; i < n * n; i++) {
        const int32_t a = src[i] < 0 ? -(int32_t)src[i] : src[i];
        const int32_t prod = (int32_t)((uint32_t)a * (uint32_t)q[i]);
        const int32_t level = (int32_t)((uint32_t)prod + (uint32_t)add) >> qbits;
        const int32_t delta = (int32_t)((uint32_t)prod - ((uint32_t)level << qbits)) >> (qbits - 8);
        const int sgn = src[i] > 0 ? 1 : (src[i] < 0 ? -1 : 0);
        sum += level;
        dst[i] = (int16_t)(sgn * sat16(level));
        delta_u[i] = sat16(delta);
    }
    *ac_sum = sum;
    if (sign_hiding && sum >= 2)
        orc_sign_bit_hiding(dst, src, orc_tables_scan(t, scan_mode, log2n), delta_u, n);
}

/* hmr_quant.c:61-169.  Coefficient groups of 16 in scan order, from the last to the first.  The first group
 * (from the end) that holds a non-zero level is "the last CG": only there the adjustment search starts at the
 * last non-zero position instead of position 15. */
void orc_sign_bit_hiding(int16_t *dst, const int16_t *src, const uint32_t *scan, const int16_t *delta_u, int n)
{
    int seen_last_cg = 0;
    for (int cg = (n * n - 1) >> 4; cg >= 0; cg--) {
        const uint32_t *sc = scan + 16 * cg;
        int first = 16, last = -1, abs_sum = 0;
        for (int i = 15; i >= 0; i--) if (dst[sc[i]]) { last = i; break; }
        for (int i = 0; i < 16; i++) if (dst[sc[i]]) { first = i; break; }
        for (int i = first; i <= last; i++) abs_sum += dst[sc[i]];
        const int is_last_cg = (last >= 0 && !seen_last_cg);
        if (last >= 0) seen_last_cg = 1;
        if (last - first < 4) continue;
        const unsigned sign_bit = dst[sc[first]] > 0 ? 0u : 1u;
        if (sign_bit == (unsigned)(abs_sum & 1)) continue;

        int min_cost = INT_MAX, min_pos = -1, final_change = 0, cur_cost = INT_MAX, cur_change = 0;
        for (int i = is_last_cg ? last : 15; i >= 0; i--) {
            const uint32_t p = sc[i];
            if (dst[p] != 0) {
                if (delta_u[p] > 0) { cur_cost = -delta_u[p]; cur_change = 1; }
                else if (i == first && abs(dst[p]) == 1) cur_cost = INT_MAX;
                else { cur_cost = delta_u[p]; cur_change = -1; }
            } else if (i < first) {
                const unsigned this_sign = src[p] >= 0 ? 0u : 1u;
                if (this_sign != sign_bit) cur_cost = INT_MAX;
                else { cur_cost = -delta_u[p]; cur_change = 1; }
            } else {
                cur_cost = -delta_u[p]; cur_change = 1;
            }
            if (cur_cost < min_cost) { min_cost = cur_cost; final_change = cur_change; min_pos = (int)p; }
        }
        if (dst[min_pos] == 32767 || dst[min_pos] == -32768) final_change = -1;
        if (src[min_pos] >= 0) dst[min_pos] = (int16_t)(dst[min_pos] + final_change);
        else                   dst[min_pos] = (int16_t)(dst[min_pos] - final_change);
    }
}

/* hmr_sse42_functions_quant.c:135-245 (== hmr_quant.c:224 on the default tables).  The SSE version picks
 * the table with `is_intra?0:3+comp` (:138), i.e. list 0 for every intra block. */
void orc_inv_quant(const orc_tables *t, const int16_t *src, int16_t *dst,
                   int log2n, int comp, int is_intra, int per, int rem)
{
    const int n = 1 << log2n;
    const int32_t *dq = orc_tables_dequant(t, log2n, is_intra ? 0 : 3 + comp, rem);
    const int iq_shift = 20 - 14 - (15 - 8 - log2n) + 4;
    if (iq_shift > per) {
        const int sh = iq_shift - per;
        const int32_t add = 1 << (sh - 1);
        for (int i = 0; i < n * n; i++)
            dst[i] = sat16((int32_t)((uint32_t)(int32_t)src[i] * (uint32_t)dq[i] + (uint32_t)add) >> sh);
    } else {
        const int sh = per - iq_shift;
        for (int i = 0; i < n * n; i++)
            dst[i] = sat16((int32_t)(((uint32_t)(int32_t)src[i] * (uint32_t)dq[i]) << sh));
    }
}

/* ------------------------------------------------------------------------------------------
 * Motion search.  hmr_motion_inter.c:1404-1774.
 * ------------------------------------------------------------------------------------------ */
/* select_mv_candidate_fast, hmr_motion_inter.c:1004; calc_mv_correction hmr_common.h:53 */
uint32_t orc_mv_cost(const orc_mv *cands, int n, int qp, double avg_dist, int mvx, int mvy, int *best_idx)
{
    uint32_t best = INT_MAX;
    int bi = 0;
    for (int i = 0; i < n; i++) {
        double w = avg_dist / 2000.;
        w = w < .15 ? .15 : (w > 1.4 ? 1.4 : w);
        const double corr = (uint32_t)qp * w;
        const double cx = corr * ((float)abs(cands[i].x - mvx));
        const double cy = corr * ((float)abs(cands[i].y - mvy));
        const uint32_t c = (uint32_t)(cx + cy + .5);
        if (best > c) { best = c; bi = i; }
    }
    if (best_idx) *best_idx = bi;
    return best;
}

typedef struct me_state {
    const orc_me_in *in;
    int xlo, xhi, ylo, yhi;
    int bx, by;                 /* running best ("curr_best") */
    uint32_t bsad, brd;
    uint32_t n_sads;
} me_state;

static int me_inside(const me_state *s, int x, int y) { return x >= s->xlo && x <= s->xhi && y >= s->ylo && y <= s->yhi; }

/* test aid (tools/me_reuse_study.py): when a buffer is set, every integer probe is appended as (x, y), then (0x7fff, 0x7fff) and the
 * half-pel winner (hx, hy) */
static __thread int16_t *g_trace; static __thread int g_trace_cap, g_trace_n;
void orc_me_trace(int16_t *buf, int cap) { g_trace = buf; g_trace_cap = cap; g_trace_n = 0; }
int orc_me_trace_count(void) { return g_trace_n; }
static void trace_put(int x, int y) { if (g_trace && g_trace_n + 2 <= g_trace_cap) { g_trace[g_trace_n++] = (int16_t)x; g_trace[g_trace_n++] = (int16_t)y; } }

static uint32_t me_sad_at(me_state *s, int x, int y)
{
    const orc_me_in *in = s->in;
    s->n_sads++;
    trace_put(x, y);
    return orc_sad(in->orig, in->orig_stride, in->ref + y * in->ref_stride + x, in->ref_stride, in->size);
}

/* probe one integer position; returns 1 when it became the running best */
static int me_probe(me_state *s, int x, int y)
{
    const orc_me_in *in = s->in;
    if (!me_inside(s, x, y)) return 0;
    const uint32_t sad = me_sad_at(s, x, y);
    const uint32_t rd = sad + orc_mv_cost(in->amvp, in->n_amvp, in->qp, in->avg_dist, x << 2, y << 2, NULL);
    if (rd < s->brd) { s->bsad = sad; s->brd = rd; s->bx = x; s->by = y; return 1; }
    return 0;
}

static const int8_t k_small[4][2] = { {-1, 0}, {0, -1}, {1, 0}, {0, 1} };                                   /* :1079 */
static const int8_t k_big[8][2] = { {-2, 0}, {-1, -1}, {0, -2}, {1, -1}, {2, 0}, {1, 1}, {0, 2}, {-1, 1} }; /* :1080 */
static const int8_t k_half_order[9][2] = { {0,0}, {0,-1}, {0,1}, {-1,0}, {1,0}, {-1,-1}, {1,-1}, {-1,1}, {1,1} };    /* :1035 */
static const int8_t k_quarter_order[9][2] = { {0,0}, {0,-1}, {0,1}, {-1,-1}, {1,-1}, {-1,0}, {1,0}, {-1,1}, {1,1} }; /* :1062 */

#define PL_STRIDE 80
#define PL_ROWS   76
typedef struct subpel_planes {
    int16_t tmp[4][PL_STRIDE * PL_ROWS];        /* H-pass results, 14 bit, by x fraction */
    int16_t pl[4][4][PL_STRIDE * PL_ROWS];      /* [y fraction][x fraction] 8-bit planes */
} subpel_planes;

/* hmr_motion_inter.c:395 */
static void build_half_planes(subpel_planes *p, const int16_t *ref, int rs, int n)
{
    const int16_t *src = ref - 4 * rs - 1;
    orc_interpolate_luma(src, rs, p->tmp[0], PL_STRIDE, 0, n + 1, n + 8, 0, 1, 0);
    orc_interpolate_luma(src, rs, p->tmp[2], PL_STRIDE, 2, n + 1, n + 8, 0, 1, 0);
    orc_interpolate_luma(p->tmp[0] + 4 * PL_STRIDE + 1, PL_STRIDE, p->pl[0][0], PL_STRIDE, 0, n, n, 1, 0, 1);
    orc_interpolate_luma(p->tmp[0] + 3 * PL_STRIDE + 1, PL_STRIDE, p->pl[2][0], PL_STRIDE, 2, n, n + 1, 1, 0, 1);
    orc_interpolate_luma(p->tmp[2] + 4 * PL_STRIDE, PL_STRIDE, p->pl[0][2], PL_STRIDE, 0, n + 1, n, 1, 0, 1);
    orc_interpolate_luma(p->tmp[2] + 3 * PL_STRIDE, PL_STRIDE, p->pl[2][2], PL_STRIDE, 2, n + 1, n + 1, 1, 0, 1);
}

/* hmr_motion_inter.c:442; (hx,hy) is the half-pel winner in half-pel steps, each in {-1,0,1} */
static void build_quarter_planes(subpel_planes *p, const int16_t *ref, int rs, int n, int hx, int hy)
{
    const int ext_h = (hy == 0) ? n + 8 : n + 7;
    const int T = PL_STRIDE;
    const int16_t *base = ref - 4 * rs - 1 + (hy > 0 ? rs : 0);
    orc_interpolate_luma(base + (hx >= 0 ? 1 : 0), rs, p->tmp[1], T, 1, n, ext_h, 0, 1, 0);
    orc_interpolate_luma(base + (hx > 0 ? 1 : 0), rs, p->tmp[3], T, 3, n, ext_h, 0, 1, 0);

    const int row0 = 3 * T;                                  /* (half_filter_size-1) rows down */
    const int vz = (hy == 0) ? T : 0;
    orc_interpolate_luma(p->tmp[1] + row0 + vz, T, p->pl[1][1], T, 1, n, n, 1, 0, 1);
    orc_interpolate_luma(p->tmp[1] + row0, T, p->pl[3][1], T, 3, n, n, 1, 0, 1);
    if (hy != 0) {
        orc_interpolate_luma(p->tmp[1] + row0, T, p->pl[2][1], T, 2, n, n, 1, 0, 1);
        orc_interpolate_luma(p->tmp[3] + row0, T, p->pl[2][3], T, 2, n, n, 1, 0, 1);
    } else {
        orc_interpolate_luma(p->tmp[1] + 4 * T, T, p->pl[0][1], T, 0, n, n, 1, 0, 1);
        orc_interpolate_luma(p->tmp[3] + 4 * T, T, p->pl[0][3], T, 0, n, n, 1, 0, 1);
    }
    if (hx != 0) {
        const int16_t *s2 = p->tmp[2] + row0 + (hx > 0 ? 1 : 0);
        orc_interpolate_luma(s2 + (hy >= 0 ? T : 0), T, p->pl[1][2], T, 1, n, n, 1, 0, 1);
        orc_interpolate_luma(s2 + (hy > 0 ? T : 0), T, p->pl[3][2], T, 3, n, n, 1, 0, 1);
    } else {
        const int16_t *s0 = p->tmp[0] + row0 + 1;
        orc_interpolate_luma(s0 + (hy >= 0 ? T : 0), T, p->pl[1][0], T, 1, n, n, 1, 0, 1);
        orc_interpolate_luma(s0 + (hy > 0 ? T : 0), T, p->pl[3][0], T, 3, n, n, 1, 0, 1);
    }
    orc_interpolate_luma(p->tmp[3] + row0 + vz, T, p->pl[1][3], T, 1, n, n, 1, 0, 1);
    orc_interpolate_luma(p->tmp[3] + row0, T, p->pl[3][3], T, 3, n, n, 1, 0, 1);
}

/* plane + offset selection shared by both sub-pel stages, hmr_motion_inter.c:1700-1710 / :1745-1755 */
static const int16_t *subpel_plane_at(const subpel_planes *p, int cx, int cy)
{
    const int16_t *s = p->pl[cy & 3][cx & 3];
    if (cx == 2 && (cy & 1) == 0) s += 1;
    if ((cx & 1) == 0 && cy == 2) s += PL_STRIDE;
    return s;
}

void orc_motion_estimation(const orc_me_in *in, orc_me_out *out)
{
    me_state s;
    memset(&s, 0, sizeof s);
    s.in = in;
    const int n = in->size;
    s.xlo = (in->gx - in->range_x < 0) ? -in->gx : -in->range_x;
    s.xhi = (in->gx + in->range_x > in->frame_w - n) ? in->frame_w - in->gx - n : in->range_x;
    s.ylo = (in->gy - in->range_y < 0) ? -in->gy : -in->range_y;
    s.yhi = (in->gy + in->range_y > in->frame_h - n) ? in->frame_h - in->gy - n : in->range_y;

    orc_mv mv = { 0, 0 }, sub = { 0, 0 };
    uint32_t best_sad = UINT_MAX / 8;
    int cx0 = 0, cy0 = 0;             /* "best_x/best_y": centre of the pattern being scanned */

    if (in->action & 1) {
        s.bx = clampi(0, s.xlo, s.xhi);
        s.by = clampi(0, s.ylo, s.yhi);
        s.bsad = me_sad_at(&s, s.bx, s.by);
        s.brd = s.bsad + orc_mv_cost(in->amvp, in->n_amvp, in->qp, in->avg_dist, s.bx << 2, s.by << 2, NULL);
        best_sad = s.bsad; cx0 = s.bx; cy0 = s.by;
        if (best_sad <= 0) goto refine;

        for (int i = 0; i < in->n_start; i++) {                      /* :1465-1490 */
            const int x = in->start[i].x >> 2, y = in->start[i].y >> 2;
            if (x == 0 && y == 0) continue;
            me_probe(&s, x, y);
        }
        best_sad = s.bsad; cx0 = s.bx; cy0 = s.by;
        if (best_sad <= 0) goto refine;

        for (int i = 0; i < 4; i++)                                  /* :1501-1523, centre stays put */
            me_probe(&s, cx0 + k_small[i][0], cy0 + k_small[i][1]);
        if (best_sad <= 0) goto refine;

        {                                                            /* :1528-1599, only l == 0 runs */
            int dist = 2;
            const int end = (cx0 != 0 && cy0 != 0) ? 4 : 8;
            int next_start = 0, span = 8;
            cx0 = s.bx; cy0 = s.by;
            while (dist < end) {
                for (int i = next_start; i < next_start + span; i++) {
                    const int idx = i % 8;
                    if (me_probe(&s, cx0 + k_big[idx][0] * dist, cy0 + k_big[idx][1] * dist)) {
                        next_start = (idx - 2 + 8) % 8;
                        span = 8 - 3;
                    }
                }
                dist *= 2;
            }
        }
refine:                                                              /* :1601-1663 */
        cx0 = s.bx; cy0 = s.by;
        {
            int next_start = 0, span = 4;
            for (;;) {
                for (int i = next_start; i < next_start + span; i++) {
                    const int idx = i % 4;
                    if (me_probe(&s, cx0 + k_small[idx][0], cy0 + k_small[idx][1])) {
                        next_start = (idx - 1 + 4) % 4;
                        span = 4 - 1;
                    }
                }
                if (cx0 == s.bx && cy0 == s.by) break;
                cx0 = s.bx; cy0 = s.by;
            }
        }
        best_sad = s.bsad;
        mv.x = s.bx << 2; mv.y = s.by << 2;
    }

    if (in->action & 2) {                                            /* :1675-1767 */
        subpel_planes *p = (subpel_planes *)malloc(sizeof *p);
        const int ix = mv.x >> 2, iy = mv.y >> 2;
        const int16_t *ref = in->ref + iy * in->ref_stride + ix;
        uint32_t cur = s.bsad;
        if (!(in->action & 1))
            cur = orc_sad(in->orig, in->orig_stride, ref, in->ref_stride, n);
        build_half_planes(p, ref, in->ref_stride, n);
        int bx = 0, by = 0, bidx = 0;
        for (int i = 0; i < 9; i++) {
            const int cx = k_half_order[i][0] * 2, cy = k_half_order[i][1] * 2;
            const uint32_t v = orc_sad(in->orig, in->orig_stride, subpel_plane_at(p, cx, cy), PL_STRIDE, n);
            if (v < cur) { cur = v; bx = cx; by = cy; bidx = i; }
        }
        mv.x = (ix << 2) + bx; mv.y = (iy << 2) + by;
        sub.x = bx; sub.y = by;
        best_sad = cur;
        if (in->action & 4) {
            const int hx = k_half_order[bidx][0], hy = k_half_order[bidx][1];
            trace_put(0x7fff, 0x7fff); trace_put(hx, hy);
            build_quarter_planes(p, ref, in->ref_stride, n, hx, hy);
            bx = hx * 2; by = hy * 2;
            for (int i = 0; i < 9; i++) {
                const int cx = hx * 2 + k_quarter_order[i][0], cy = hy * 2 + k_quarter_order[i][1];
                const uint32_t v = orc_sad(in->orig, in->orig_stride, subpel_plane_at(p, cx, cy), PL_STRIDE, n);
                if (v < cur) { cur = v; bx = cx; by = cy; }
            }
            best_sad = cur;
            mv.x = (ix << 2) + bx; mv.y = (iy << 2) + by;
            sub.x = bx; sub.y = by;
        }
        free(p);
    }
    out->mv = mv; out->subpix = sub; out->sad = best_sad; out->n_int_sads = s.n_sads;
}

/* ------------------------------------------------------------------------------------------
 * Motion compensation (uni-prediction).  hmr_motion_inter.c:1779 and :1860.
 * ------------------------------------------------------------------------------------------ */
/* is_bi != 0: the 14-bit prediction of one list of a bi-predicted block (is_last = !is_bi_predict, :1797-1812), to be
 * averaged by orc_weighted_average */
void orc_mc_luma_ex(const int16_t *ref, int ref_stride, int16_t *pred, int pred_stride, int size, orc_mv mv, int is_bi)
{
    const int fx = mv.x & 3, fy = mv.y & 3;
    const int16_t *src = ref + (mv.y >> 2) * ref_stride + (mv.x >> 2);
    if (fx == 0) {
        orc_interpolate_luma(src, ref_stride, pred, pred_stride, fy, size, size, 1, 1, !is_bi);
    } else if (fy == 0) {
        orc_interpolate_luma(src, ref_stride, pred, pred_stride, fx, size, size, 0, 1, !is_bi);
    } else {
        int16_t tmp[PL_STRIDE * PL_ROWS];
        orc_interpolate_luma(src - 3 * ref_stride, ref_stride, tmp, PL_STRIDE, fx, size, size + 7, 0, 1, 0);
        orc_interpolate_luma(tmp + 3 * PL_STRIDE, PL_STRIDE, pred, pred_stride, fy, size, size, 1, 0, !is_bi);
    }
}
void orc_mc_luma(const int16_t *ref, int ref_stride, int16_t *pred, int pred_stride, int size, orc_mv mv)
{
    orc_mc_luma_ex(ref, ref_stride, pred, pred_stride, size, mv, 0);
}

void orc_mc_chroma_ex(const int16_t *ref, int ref_stride, int16_t *pred, int pred_stride, int size, orc_mv mv, int is_bi)
{
    const int fx = mv.x & 7, fy = mv.y & 7;
    const int16_t *src = ref + (mv.y >> 3) * ref_stride + (mv.x >> 3);
    if (fx == 0) {
        orc_interpolate_chroma(src, ref_stride, pred, pred_stride, fy, size, size, 1, 1, !is_bi);
    } else if (fy == 0) {
        orc_interpolate_chroma(src, ref_stride, pred, pred_stride, fx, size, size, 0, 1, !is_bi);
    } else {
        int16_t tmp[PL_STRIDE * PL_ROWS];
        orc_interpolate_chroma(src - ref_stride, ref_stride, tmp, PL_STRIDE, fx, size, size + 4 + 1, 0, 1, 0);
        orc_interpolate_chroma(tmp + PL_STRIDE, PL_STRIDE, pred, pred_stride, fy, size, size, 1, 0, !is_bi);
    }
}
void orc_mc_chroma(const int16_t *ref, int ref_stride, int16_t *pred, int pred_stride, int size, orc_mv mv)
{
    orc_mc_chroma_ex(ref, ref_stride, pred, pred_stride, size, mv, 0);
}

/* ------------------------------------------------------------------------------------------
 * T/Q chain of one inter TU.  hmr_motion_inter.c:40-131 (luma) and :133-228 (chroma):
 * predict -> transform -> quant -> [ssd(resid,0), inv_quant, itransform, ssd(resid,dec_resid), zero-out test] -> reconst.
 * The value handed back as `ssd` is the one the reference returns: it is NOT replaced by ssd_zero when the TU is
 * zeroed out (:106-113).
 * ------------------------------------------------------------------------------------------ */
void orc_encode_inter_tu(const orc_tables *t, const int16_t *orig, int orig_stride,
                         const int16_t *pred, int pred_stride,
                         int16_t *coeff, int16_t *dec, int dec_stride,
                         int n, int comp, int qp, int is_islice, int sign_hiding,
                         double avg_dist, double weight, orc_tu_out *out)
{
    int16_t resid[32 * 32], tc[32 * 32], dq[32 * 32], rdec[32 * 32], du[32 * 32];
    static const int16_t zeros[32] = { 0 };
    const int lg = ilog2(n), per = qp / 6, rem = qp % 6;
    int sum = 0;
    orc_predict(orig, orig_stride, pred, pred_stride, resid, n, n);
    orc_transform(8, resid, n, tc, n, 0);
    orc_quant(t, tc, coeff, du, 3 /* DIAG_SCAN */, lg, comp, 0, is_islice, sign_hiding, per, rem, &sum);
    out->zeroed = 0; out->ssd_zero = 0;
    if (sum > 0) {
        const uint32_t ssd_zero = (uint32_t)(weight * orc_ssd16b(resid, n, zeros, 0, n));
        orc_inv_quant(t, coeff, dq, lg, comp, 0, per, rem);
        orc_itransform(8, rdec, n, dq, n, 0);
        const uint32_t ssd = (uint32_t)(weight * orc_ssd16b(resid, n, rdec, n, n));
        double k = avg_dist / 2.5 - 5.;
        k = k < 1. ? 1. : (k > 20000. ? 20000. : k);
        out->ssd_zero = ssd_zero; out->ssd = ssd;
        if (comp == 0 ? ((double)ssd_zero <= (double)(int)ssd + k * sum) : ((double)ssd_zero <= (double)ssd + k * sum)) {
            memset(coeff, 0, sizeof(int16_t) * (size_t)n * n);
            sum = 0; out->zeroed = 1;
            orc_reconst(pred, pred_stride, zeros, 0, dec, dec_stride, n);
        } else {
            orc_reconst(pred, pred_stride, rdec, n, dec, dec_stride, n);
        }
    } else {
        out->ssd = (uint32_t)(weight * orc_ssd16b(resid, n, zeros, 0, n));
        orc_reconst(pred, pred_stride, zeros, 0, dec, dec_stride, n);
    }
    out->sum = sum;
}

/* ------------------------------------------------------------------------------------------
 * T/Q chain of one INTRA TU once its prediction exists.  hmr_motion_intra.c:1023-1069 (luma: DST-VII for 4x4 because the
 * intra mode is handed to transform() as uiMode != REG_DCT, scan chosen from the mode, quant/inv_quant with is_intra = 1,
 * distortion = ssd16b(orig, decoded)) and hmr_motion_intra_chroma.c:340-365 (chroma: DCT, distortion weighted).
 * ------------------------------------------------------------------------------------------ */
void orc_encode_intra_tu(const orc_tables *t, const int16_t *orig, int orig_stride, const int16_t *pred, int pred_stride,
                         int16_t *coeff, int16_t *dec, int dec_stride, int n, int comp, int qp, int scan_mode,
                         int is_islice, int sign_hiding, double weight, orc_tu_out *out)
{
    int16_t resid[32 * 32], tc[32 * 32], dq[32 * 32], du[32 * 32];
    static const int16_t zeros[32] = { 0 };
    const int lg = ilog2(n), per = qp / 6, rem = qp % 6, is_dst = (comp == 0 && n == 4);
    int sum = 0;
    orc_predict(orig, orig_stride, pred, pred_stride, resid, n, n);
    orc_transform(8, resid, n, tc, n, is_dst);
    orc_quant(t, tc, coeff, du, scan_mode, lg, comp, 1, is_islice, sign_hiding, per, rem, &sum);
    if (sum) {
        orc_inv_quant(t, coeff, dq, lg, comp, 1, per, rem);
        orc_itransform(8, resid, n, dq, n, is_dst);
        orc_reconst(pred, pred_stride, resid, n, dec, dec_stride, n);
    } else {
        orc_reconst(pred, pred_stride, zeros, 0, dec, dec_stride, n);
    }
    const uint32_t ssd = orc_ssd16b(orig, orig_stride, dec, dec_stride, n);
    out->sum = sum; out->zeroed = 0; out->ssd_zero = 0;
    out->ssd = comp == 0 ? ssd : (uint32_t)(int)(weight * ssd);
}

/* ------------------------------------------------------------------------------------------
 * Intra prediction.  Reference samples ("ADI") are one array of 4N+1 values: index 2N is the
 * top-left corner, 2N+i (i = 1..2N) the row above from left to right, 2N-i the column on the left from top to bottom.
 *   adi_filter                       hmr_motion_intra.c:189   ([1,2,1] or the strong bilinear smoothing for N >= 32)
 *   create_intra_planar_prediction   hmr_motion_intra.c:408
 *   create_intra_angular_prediction  hmr_motion_intra.c:482   (modes 1 = DC, 2..34 angular; DC/H/V edge filters for luma N <= 16)
 *   filtered-or-not per mode         hmr_motion_intra.c:1122-1123 (intra_filter[] thresholds :148)
 * ------------------------------------------------------------------------------------------ */
void orc_adi_filter(const int16_t *adi, int16_t *flt, int n, int strong_enabled)
{
    const int size = 4 * n + 1;
    const int lb = adi[0], lt = adi[2 * n], tr = adi[size - 1];
    const int thr = 1 << (8 - 5);
    const int bil_l = abs(lb + lt - 2 * adi[n]) < thr, bil_a = abs(lt + tr - 2 * adi[3 * n]) < thr;
    if (strong_enabled && n >= 32 && bil_l && bil_a) {
        const int sh = ilog2(n) + 1;
        flt[0] = adi[0]; flt[2 * n] = adi[2 * n]; flt[size - 1] = adi[size - 1];
        for (int i = 1; i < 2 * n; i++) flt[i] = (int16_t)(((2 * n - i) * lb + i * lt + n) >> sh);
        for (int i = 1; i < 2 * n; i++) flt[2 * n + i] = (int16_t)(((2 * n - i) * lt + i * tr + n) >> sh);
    } else {
        flt[0] = adi[0];
        for (int i = 1; i < size - 1; i++) flt[i] = (int16_t)((adi[i - 1] + 2 * adi[i] + adi[i + 1] + 2) >> 2);
        flt[size - 1] = adi[size - 1];
    }
}

static const int k_ang[9] = { 0, 2, 5, 9, 13, 17, 21, 26, 32 };            /* hmr_encoder_lib.c:35 */
static const int k_inv_ang[9] = { 0, 4096, 1638, 910, 630, 482, 390, 315, 256 };

void orc_intra_predict(const int16_t *adi, int n, int mode, int is_luma, int16_t *pred, int stride)
{
    const int16_t *mid = adi + 2 * n;
    const int lg = ilog2(n);
    if (mode == 0) {                                                          /* planar */
        int top_row[64], left_col[64], bottom[64], right[64];
        const int lb = mid[-(n + 1)], tr = mid[n + 1];
        for (int i = 0; i < n; i++) {
            const int l = mid[-(i + 1)], t = mid[i + 1];
            bottom[i] = lb - t; right[i] = tr - l; top_row[i] = t << lg; left_col[i] = l << lg;
        }
        for (int j = 0; j < n; j++) {
            int hor = left_col[j] + n;
            for (int i = 0; i < n; i++) {
                hor += right[j];
                top_row[i] += bottom[i];
                pred[j * stride + i] = (int16_t)((hor + top_row[i]) >> (lg + 1));
            }
        }
        return;
    }
    if (mode == 1) {                                                          /* DC, both neighbours always "available" (:257-258) */
        int sum = 0;
        for (int i = 1; i <= n; i++) sum += mid[i] + mid[-i];
        const int dc = (uint16_t)((sum + n) / (2 * n));
        for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) pred[j * stride + i] = (uint8_t)dc;
        if (n <= 16 && is_luma) {
            pred[0] = (int16_t)((mid[-1] + mid[1] + 2 * pred[0] + 2) >> 2);
            for (int i = 1; i < n; i++) pred[i] = (int16_t)((mid[1 + i] + 3 * pred[i] + 2) >> 2);
            for (int j = 1; j < n; j++) pred[j * stride] = (int16_t)((mid[-1 - j] + 3 * pred[j * stride] + 2) >> 2);
        }
        return;
    }
    const int hor_mode = mode < 18;
    int angle = hor_mode ? -(mode - 10) : mode - 26;
    const int a = abs(angle), sgn = angle < 0 ? -1 : 1;
    const int inv = k_inv_ang[a];
    angle = sgn * k_ang[a];
    int16_t above_buf[3 * 64 + 4], left_buf[3 * 64 + 4];
    int16_t *const above = above_buf + 2, *const left = left_buf + 2;     /* two spare samples in front keep gcc's bounds analysis quiet */
    int16_t *ref_main, *ref_side;
    if (angle < 0) {
        for (int i = 0; i < n + 1; i++) { above[i + n - 1] = mid[i]; left[i + n - 1] = mid[-i]; }
        ref_main = (hor_mode ? left : above) + (n - 1);
        ref_side = (hor_mode ? above : left) + (n - 1);
        int inv_sum = 128;
        for (int i = -1; i > ((n * angle) >> 5); i--) { inv_sum += inv; ref_main[i] = ref_side[inv_sum >> 8]; }
    } else {
        for (int i = 0; i < 2 * n + 1; i++) { above[i] = mid[i]; left[i] = mid[-i]; }
        ref_main = hor_mode ? left : above;
        ref_side = hor_mode ? above : left;
    }
    const int s1 = hor_mode ? 1 : stride, s2 = hor_mode ? stride : 1;       /* horizontal modes write transposed */
    if (angle == 0) {
        for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) pred[j * s1 + i * s2] = (uint8_t)ref_main[i + 1];
        if (is_luma && n <= 16)
            for (int i = 0; i < n; i++) pred[i * s1] = (int16_t)clampi(pred[i * s1] + ((ref_side[i + 1] - ref_side[0]) >> 1), 0, 255);
        return;
    }
    int pos = 0;
    for (int j = 0; j < n; j++) {
        pos += angle;
        const int d = pos >> 5, f = pos & 31;
        for (int i = 0; i < n; i++) {
            const int k = i + d + 1;
            pred[j * s1 + i * s2] = f ? (uint8_t)(((32 - f) * ref_main[k] + f * ref_main[k + 1] + 16) >> 5) : (uint8_t)ref_main[k];
        }
    }
}

/* which reference array a luma mode reads (hmr_motion_intra.c:1122-1123) */
int orc_intra_uses_filtered(int n, int mode)
{
    static const int thr[5] = { 10, 7, 1, 0, 10 };
    const int d1 = abs(mode - 10), d2 = abs(mode - 26);
    return mode != 1 && ((d1 < d2 ? d1 : d2) > thr[ilog2(n) - 2]);
}

/* SAD of every luma mode 0..34 against the original block: what the search loop of homer_loop1_motion_intra (:1084) probes */
void orc_intra_mode_sads(const int16_t *orig, int orig_stride, const int16_t *adi, int n, uint32_t sads[35])
{
    int16_t flt[4 * 64 + 1], pred[64 * 64];
    orc_adi_filter(adi, flt, n, 1);
    for (int m = 0; m < 35; m++) {
        orc_intra_predict(orc_intra_uses_filtered(n, m) ? flt : adi, n, m, 1, pred, n);
        sads[m] = orc_sad(orig, orig_stride, pred, n, n);
    }
}

/* Bi-prediction average of two 14-bit predictions (weighted_average_motion, hmr_motion_inter.c:2903; the SSE4.2 form,
 * hmr_sse42_functions_inter_prediction.c:848-945, adds in 32 bits, packs with saturation and clips: the same values) */
void orc_weighted_average(const int16_t *src0, int s0, const int16_t *src1, int s1, int16_t *dst, int ds, int height, int width)
{
    const int shift = 14 + 1 - 8, offset = (1 << (shift - 1)) + 2 * 8192;
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) dst[y * ds + x] = (int16_t)clampi((src0[y * s0 + x] + src1[y * s1 + x] + offset) >> shift, 0, 255);
}

/* ------------------------------------------------------------------------------------------
 * SAO statistics of one CTU and component: sao_get_ctu_stats, hmr_sao.c:75-330, with
 * calculate_preblock_stats = 0 (the only value the encoder sets, hmr_encoder_lib.c:684).  The reference walks each edge class
 * with running sign buffers; every sample it counts ends up classified by sign(c - a) + sign(c - b) of its two neighbours
 * along the class direction, inside a rectangle that depends on which neighbouring CTUs exist and on the columns / rows the
 * deblocking filter has not finished yet (skiped_lines_r = {5,3,3}, skiped_lines_b = {4,2,2}, :60-61).
 * rec / org: the component's planes (deblocked reconstruction, source), (x0, y0) the CTU origin in that plane.
 * ------------------------------------------------------------------------------------------ */
static int sgn3(int v) { return v == 0 ? 0 : (v < 0 ? -1 : 1); }
void orc_sao_ctu_stats(const int16_t *rec, int rec_stride, const int16_t *org, int org_stride, int comp, int x0, int y0,
                       int pic_w, int pic_h, int ctu_size, orc_sao_stats *st)
{
    static const int skip_r[3] = { 5, 3, 3 }, skip_b[3] = { 4, 2, 2 };
    static const int dx[4] = { 1, 0, 1, -1 }, dy[4] = { 0, 1, 1, 1 };      /* EO_0, EO_90, EO_135, EO_45: second neighbour; the first is its opposite */
    const int w = x0 + ctu_size > pic_w ? pic_w - x0 : ctu_size, h = y0 + ctu_size > pic_h ? pic_h - y0 : ctu_size;
    const int l = x0 > 0, t = y0 > 0, r = x0 + ctu_size < pic_w, b = y0 + ctu_size < pic_h;
    memset(st, 0, sizeof *st);
    for (int type = 0; type < 5; type++) {
        int sx, ex, sy, ey;
        if (type == 0)      { sx = l ? 0 : 1; ex = r ? w - skip_r[comp] : w - 1; sy = 0; ey = b ? h - skip_b[comp] : h; }
        else if (type == 1) { sx = 0; ex = r ? w - skip_r[comp] : w; sy = t ? 0 : 1; ey = b ? h - skip_b[comp] : h - 1; }
        else if (type < 4)  { sx = l ? 0 : 1; ex = r ? w - skip_r[comp] : w - 1; sy = t ? 0 : 1; ey = b ? h - skip_b[comp] : h - 1; }
        else                { sx = 0; ex = r ? w - skip_r[comp] : w; sy = 0; ey = b ? h - skip_b[comp] : h; }
        for (int y = sy; y < ey; y++)
            for (int x = sx; x < ex; x++) {
                const int c = rec[(y0 + y) * rec_stride + x0 + x], d = org[(y0 + y) * org_stride + x0 + x] - c;
                if (type == 4) { st->bo_diff[c >> 3] += d; st->bo_count[c >> 3]++; continue; }
                const int a = rec[(y0 + y - dy[type]) * rec_stride + x0 + x - dx[type]], n = rec[(y0 + y + dy[type]) * rec_stride + x0 + x + dx[type]];
                const int k = 2 + sgn3(c - a) + sgn3(c - n);
                st->eo_diff[type][k] += d; st->eo_count[type][k]++;
            }
    }
}

/* ------------------------------------------------------------------------------------------
 * SAO offset pass of one CTU and component: offset_block, hmr_sao.c:960-1208 (called by sao_offset_ctu :1210).  src: the
 * deblocked picture (the reference keeps a copy in sao_aux_wnd), dst: the picture being finalised.  type 0..3 edge classes
 * (offset[0..4] for edge types -2..2), 4 band offset (offset[0..31] by band), anything else: untouched.  With the
 * availability of the diagonal neighbours being the AND of the two sides (:1224-1231) the first / last line rules collapse
 * into one rectangle per type.
 * ------------------------------------------------------------------------------------------ */
void orc_sao_offset_ctu(const int16_t *src, int src_stride, int16_t *dst, int dst_stride, int x0, int y0, int pic_w, int pic_h,
                        int ctu_size, int type, const int *offset)
{
    static const int dx[4] = { 1, 0, 1, -1 }, dy[4] = { 0, 1, 1, 1 };
    const int w = x0 + ctu_size > pic_w ? pic_w - x0 : ctu_size, h = y0 + ctu_size > pic_h ? pic_h - y0 : ctu_size;
    const int l = x0 > 0, t = y0 > 0, r = x0 + ctu_size < pic_w, b = y0 + ctu_size < pic_h;
    if (type < 0 || type > 4) return;
    int sx = 0, ex = w, sy = 0, ey = h;
    if (type == 0 || type == 2 || type == 3) { sx = l ? 0 : 1; ex = r ? w : w - 1; }
    if (type == 1 || type == 2 || type == 3) { sy = t ? 0 : 1; ey = b ? h : h - 1; }
    for (int y = sy; y < ey; y++)
        for (int x = sx; x < ex; x++) {
            const int c = src[(y0 + y) * src_stride + x0 + x];
            int k;
            if (type == 4) k = c >> 3;
            else {
                const int a = src[(y0 + y - dy[type]) * src_stride + x0 + x - dx[type]], n = src[(y0 + y + dy[type]) * src_stride + x0 + x + dx[type]];
                k = 2 + sgn3(c - a) + sgn3(c - n);
            }
            dst[(y0 + y) * dst_stride + x0 + x] = (int16_t)clampi(c + offset[k], 0, 255);
        }
}

/* ------------------------------------------------------------------------------------------
 * Deblocking of a whole picture, pixel stage: deblock_filter_luma / _chroma, filter_luma, filter_chroma,
 * use_strong_filter, hmr_deblocking_filter.c:264-627, in the picture order of hmr_deblock_filter (:827): every vertical edge,
 * then every horizontal edge.  The boundary strengths are an INPUT (per 4x4 luma unit, picture raster, units_w per row; the
 * strength of the edge on the unit's left / top side): deriving them from modes, cbf and vectors is host data
 * (get_boundary_strength_single :138).  qp: the CU's QP per unit.  Edges lie on the 8x8 luma grid (chroma: 8x8 chroma grid,
 * strength 2 only); a unit's entry off that grid is ignored.  planes: int16 pictures, modified in place.
 * ------------------------------------------------------------------------------------------ */
static const uint8_t k_tc_table[54] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24 };
static const uint8_t k_beta_table[52] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64 };

static int dbk_strong(const int16_t *s, int o, int d, int beta, int tc)
{
    const int d_strong = abs(s[-4 * o] - s[-o]) + abs(s[3 * o] - s[0]);
    return d_strong < (beta >> 3) && d < (beta >> 2) && abs(s[-o] - s[0]) < ((tc * 5 + 1) >> 1);
}
static void dbk_luma_line(int16_t *s, int o, int tc, int sw, int thr_cut, int second_p, int second_q)
{
    const int m0 = s[-4 * o], m1 = s[-3 * o], m2 = s[-2 * o], m3 = s[-o], m4 = s[0], m5 = s[o], m6 = s[2 * o], m7 = s[3 * o];
    if (sw) {
        s[-o] = (int16_t)clampi((m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3, m3 - 2 * tc, m3 + 2 * tc);
        s[0] = (int16_t)clampi((m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3, m4 - 2 * tc, m4 + 2 * tc);
        s[-2 * o] = (int16_t)clampi((m1 + m2 + m3 + m4 + 2) >> 2, m2 - 2 * tc, m2 + 2 * tc);
        s[o] = (int16_t)clampi((m3 + m4 + m5 + m6 + 2) >> 2, m5 - 2 * tc, m5 + 2 * tc);
        s[-3 * o] = (int16_t)clampi((2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3, m1 - 2 * tc, m1 + 2 * tc);
        s[2 * o] = (int16_t)clampi((m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3, m6 - 2 * tc, m6 + 2 * tc);
    } else {
        int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
        if (abs(delta) < thr_cut) {
            const int tc2 = tc >> 1;
            delta = clampi(delta, -tc, tc);
            s[-o] = (int16_t)clampi(m3 + delta, 0, 255);
            s[0] = (int16_t)clampi(m4 - delta, 0, 255);
            if (second_p) s[-2 * o] = (int16_t)clampi(m2 + clampi((((m1 + m3 + 1) >> 1) - m2 + delta) >> 1, -tc2, tc2), 0, 255);
            if (second_q) s[o] = (int16_t)clampi(m5 + clampi((((m6 + m4 + 1) >> 1) - m5 - delta) >> 1, -tc2, tc2), 0, 255);
        }
    }
}
void orc_deblock_picture(int16_t *const planes[3], const int strides[3], int w, int h, const uint8_t *bs_ver, const uint8_t *bs_hor,
                         const uint8_t *qp, int units_w, int cb_qp_offset, int cr_qp_offset, int beta_offset_div2, int tc_offset_div2)
{
    for (int dir = 0; dir < 2; dir++) {
        const uint8_t *bsm = dir ? bs_hor : bs_ver;
        for (int uy = 0; uy < h / 4; uy++)
            for (int ux = 0; ux < w / 4; ux++) {
                const int along = dir ? uy : ux;                       /* unit coordinate across the edge */
                const int bs = bsm[uy * units_w + ux];
                if (!bs || along == 0 || (along & 1)) continue;
                const int qpq = qp[uy * units_w + ux], qpp = dir ? qp[(uy - 1) * units_w + ux] : qp[uy * units_w + ux - 1];
                const int q = (qpp + qpq + 1) >> 1;
                {   /* luma: four lines of the 8x8-grid edge */
                    const int tc = k_tc_table[clampi(q + 2 * (bs - 1) + (tc_offset_div2 << 1), 0, 53)];
                    const int beta = k_beta_table[clampi(q + (beta_offset_div2 << 1), 0, 51)];
                    const int side_thr = (beta + (beta >> 1)) >> 3, thr_cut = tc * 10;
                    const int o = dir ? strides[0] : 1, step = dir ? 1 : strides[0];
                    int16_t *e = planes[0] + (4 * uy) * strides[0] + 4 * ux;
                    const int16_t *l0 = e, *l3 = e + 3 * step;
                    const int dp0 = abs(l0[-3 * o] - 2 * l0[-2 * o] + l0[-o]), dq0 = abs(l0[0] - 2 * l0[o] + l0[2 * o]);
                    const int dp3 = abs(l3[-3 * o] - 2 * l3[-2 * o] + l3[-o]), dq3 = abs(l3[0] - 2 * l3[o] + l3[2 * o]);
                    const int d0 = dp0 + dq0, d3 = dp3 + dq3, d = d0 + d3;
                    if (d < beta) {
                        const int sw = dbk_strong(l0, o, 2 * d0, beta, tc) && dbk_strong(l3, o, 2 * d3, beta, tc);
                        for (int i = 0; i < 4; i++) dbk_luma_line(e + i * step, o, tc, sw, thr_cut, dp0 + dp3 < side_thr, dq0 + dq3 < side_thr);
                    }
                }
                if (bs > 1 && (along & 3) == 0)                         /* chroma: 8x8 chroma grid, two lines per luma unit */
                    for (int c = 1; c < 3; c++) {
                        const int qc = orc_chroma_qp_table[clampi(q + (c == 1 ? cb_qp_offset : cr_qp_offset), 0, 57)];
                        const int tc = k_tc_table[clampi(qc + 2 * (bs - 1) + (tc_offset_div2 << 1), 0, 53)];
                        const int o = dir ? strides[c] : 1, step = dir ? 1 : strides[c];
                        int16_t *e = planes[c] + (2 * uy) * strides[c] + 2 * ux;
                        for (int i = 0; i < 2; i++) {
                            int16_t *s = e + i * step;
                            const int m2 = s[-2 * o], m3 = s[-o], m4 = s[0], m5 = s[o];
                            const int delta = clampi((((m4 - m3) << 2) + m2 - m5 + 4) >> 3, -tc, tc);
                            s[-o] = (int16_t)clampi(m3 + delta, 0, 255);
                            s[0] = (int16_t)clampi(m4 - delta, 0, 255);
                        }
                    }
            }
    }
}

/* ------------------------------------------------------------------------------------------
 * Boundary strengths of a whole P picture (2Nx2N prediction units): which unit sides are transform / coding unit edges
 * (hmr_deblock_filter_cu :737 marks the left column and top row of every transform leaf of at least 8x8, set_edge_filter_pu :692
 * clears the picture border) and get_boundary_strength_single :138 -- 2 when either side is intra, else 1 when either side's
 * transform unit has luma coefficients, else 1 when the list-0 references differ or a vector component differs by 4 or more.
 * Per 4x4 unit, picture raster: cu_depth (0..3), tu_depth below the CU, intra flag, luma cbf byte (bit tu_depth), list-0
 * reference index (< 0 none) and vector.
 * ------------------------------------------------------------------------------------------ */
void orc_deblock_strengths(const uint8_t *cu_depth, const uint8_t *tu_depth, const uint8_t *intra, const uint8_t *cbf, const int8_t *ref_idx,
                           const int16_t *mv, int units_w, int w, int h, uint8_t *bs_ver, uint8_t *bs_hor)
{
    for (int uy = 0; uy < h / 4; uy++)
        for (int ux = 0; ux < w / 4; ux++) {
            const int q = uy * units_w + ux;
            int ts = 64 >> (cu_depth[q] + tu_depth[q]);
            if (ts < 8) ts = 8;
            for (int dir = 0; dir < 2; dir++) {
                const int pos = 4 * (dir ? uy : ux);
                uint8_t *out = (dir ? bs_hor : bs_ver) + q;
                *out = 0;
                if (pos == 0 || pos % ts) continue;
                const int p = dir ? q - units_w : q - 1;
                int bs;
                if (intra[p] || intra[q]) bs = 2;
                else if (((cbf[q] >> tu_depth[q]) & 1) || ((cbf[p] >> tu_depth[p]) & 1)) bs = 1;
                else bs = (ref_idx[p] != ref_idx[q]) || abs(mv[2 * q] - mv[2 * p]) >= 4 || abs(mv[2 * q + 1] - mv[2 * p + 1]) >= 4;
                *out = (uint8_t)bs;
            }
        }
}


/* the same for a B picture (:173-229): the motion of either side is two (picture, vector) pairs -- list 0 and list 1, a list that is not
 * used has no picture and a zero vector; pic_l0 / pic_l1 map a reference index to the picture it names (the reference compares picture
 * pointers).  Same set of pictures on both sides: no edge only if the vectors of matching pictures differ by less than 4 (both pairings
 * count when the two lists name one picture); different sets: strength 1. */
void orc_deblock_strengths_b(const uint8_t *cu_depth, const uint8_t *tu_depth, const uint8_t *intra, const uint8_t *cbf, const int8_t *ref0, const int16_t *mv0,
                             const int8_t *ref1, const int16_t *mv1, const int32_t *pic_l0, const int32_t *pic_l1, int units_w, int w, int h,
                             uint8_t *bs_ver, uint8_t *bs_hor)
{
    for (int uy = 0; uy < h / 4; uy++)
        for (int ux = 0; ux < w / 4; ux++) {
            const int q = uy * units_w + ux;
            int ts = 64 >> (cu_depth[q] + tu_depth[q]);
            if (ts < 8) ts = 8;
            for (int dir = 0; dir < 2; dir++) {
                const int pos = 4 * (dir ? uy : ux);
                uint8_t *out = (dir ? bs_hor : bs_ver) + q;
                *out = 0;
                if (pos == 0 || pos % ts) continue;
                const int p = dir ? q - units_w : q - 1;
                int bs;
                if (intra[p] || intra[q]) bs = 2;
                else if (((cbf[q] >> tu_depth[q]) & 1) || ((cbf[p] >> tu_depth[p]) & 1)) bs = 1;
                else {
                    #define PIC(r, tab) ((r) < 0 ? -1 : (tab)[r])
                    const int r0 = PIC(ref0[p], pic_l0), r1 = PIC(ref1[p], pic_l1), c0 = PIC(ref0[q], pic_l0), c1 = PIC(ref1[q], pic_l1);
                    #undef PIC
                    const int p0x = r0 < 0 ? 0 : mv0[2 * p], p0y = r0 < 0 ? 0 : mv0[2 * p + 1], p1x = r1 < 0 ? 0 : mv1[2 * p], p1y = r1 < 0 ? 0 : mv1[2 * p + 1];
                    const int q0x = c0 < 0 ? 0 : mv0[2 * q], q0y = c0 < 0 ? 0 : mv0[2 * q + 1], q1x = c1 < 0 ? 0 : mv1[2 * q], q1y = c1 < 0 ? 0 : mv1[2 * q + 1];
                    #define FAR(ax, ay, bx, by) (abs((ax) - (bx)) >= 4 || abs((ay) - (by)) >= 4)
                    if ((r0 == c0 && r1 == c1) || (r0 == c1 && r1 == c0)) {
                        if (r0 != r1) {
                            if (r0 == c0) bs = FAR(q0x, q0y, p0x, p0y) || FAR(q1x, q1y, p1x, p1y);
                            else bs = FAR(q1x, q1y, p0x, p0y) || FAR(q0x, q0y, p1x, p1y);
                        } else {
                            bs = (FAR(q0x, q0y, p0x, p0y) || FAR(q1x, q1y, p1x, p1y)) && (FAR(q1x, q1y, p0x, p0y) || FAR(q0x, q0y, p1x, p1y));
                        }
                    } else bs = 1;
                    #undef FAR
                }
                *out = (uint8_t)bs;
            }
        }
}


/* ------------------------------------------------------------------------------------------
 * AMVP candidates (P picture, one reference picture).  hmr_motion_inter.c:2342-2460.
 * ------------------------------------------------------------------------------------------ */
static int zscan16(int ux, int uy)                       /* raster2abs_table of a 16 x 16 unit CTU: x bits even, y bits odd */
{
    int a = 0;
    for (int b = 0; b < 4; b++) a |= ((ux >> b) & 1) << (2 * b) | ((uy >> b) & 1) << (2 * b + 1);
    return a;
}
/* left_bottom_neighbour / top_right_neighbour of the partition of `size` at (px, py) inside the CTU at (ctu_x, ctu_y): the CTU's
 * own flags (hmr_motion_intra.c:676-683; ctu_left_bottom is never set) handed down the quadtree by cu_partition_get_neighbours (:625-657) */
static void amvp_flags(int ctu_x, int ctu_y, int w, int h, int px, int py, int size, int *left_bottom, int *top_right)
{
    const int cols = (w + 63) / 64;
    int l = ctu_x > 0, t = ctu_y > 0, lb = 0, tr = ctu_y > 0 && ctu_x / 64 + 1 < cols;
    const int valid_lines = h - ctu_y < 64 ? h - ctu_y : 64, valid_cols = w - ctu_x < 64 ? w - ctu_x : 64;
    int par_x = 0, par_y = 0;
    for (int s = 32; s >= size; s >>= 1) {
        const int cx = par_x + ((px - par_x) >= s ? s : 0), cy = par_y + ((py - par_y) >= s ? s : 0);
        const int nlb = (lb && cx == par_x) || (l && cx == par_x && cy == par_y && valid_lines > cy + s);
        const int ntr = (tr && cy == par_y) || (t && cx == par_x && cy == par_y && valid_cols > cx + s) || (cx == par_x && cy != par_y && valid_cols > cx + s);
        l = l || cx; t = t || cy; lb = nlb; tr = ntr; par_x = cx; par_y = cy;
    }
    *left_bottom = lb; *top_right = tr;
}
/* the five spatial neighbours of the PU, as unit indices into the maps, and whether the reference would look at them at all
 * (A0 left-bottom, A1 left, B0 top-right, B1 top, B2 top-left) */
static void pu_neighbours(int units_w, int w, int h, int x, int y, int size, int cand[5], int ok[5])
{
    const int ctu_x = x & ~63, ctu_y = y & ~63, px = x - ctu_x, py = y - ctu_y;
    const int cols = (w + 63) / 64;
    const int has_left = ctu_x > 0, has_top = ctu_y > 0, has_top_right = ctu_y > 0 && ctu_x / 64 + 1 < cols, has_top_left = ctu_x > 0 && ctu_y > 0;
    int lb_flag, tr_flag;
    amvp_flags(ctu_x, ctu_y, w, h, px, py, size, &lb_flag, &tr_flag);
    const int gx0 = ctu_x / 4, gy0 = ctu_y / 4;                           /* the CTU's first unit, picture unit coordinates */
    {   /* bottom-left 4x4 unit of the PU */
        const int ux = px / 4, uy = (py + size) / 4 - 1;
        /* get_pu_left_bottom :245 */
        cand[0] = (gy0 + uy + 1) * units_w + gx0 + ux - 1;
        if (!lb_flag) ok[0] = 0;
        else if (ux == 0 && uy == 15) ok[0] = 0;                           /* ctu_left_bottom: NULL */
        else if (ux == 0) ok[0] = has_left;
        else if (uy == 15) ok[0] = 0;
        else ok[0] = zscan16(ux, uy) > zscan16(ux - 1, uy + 1);
        /* get_pu_left :229 */
        cand[1] = (gy0 + uy) * units_w + gx0 + ux - 1;
        ok[1] = ux == 0 ? has_left : 1;
    }
    {   /* top-right unit */
        const int ux = (px + size) / 4 - 1, uy = py / 4;
        /* get_pu_top_right :301 */
        cand[2] = (gy0 + uy - 1) * units_w + gx0 + ux + 1;
        if (!tr_flag) ok[2] = 0;
        else if (ux == 15 && uy == 0) ok[2] = has_top_right;
        else if (uy == 0) ok[2] = has_top;
        else if (ux == 15) ok[2] = 0;
        else ok[2] = zscan16(ux, uy) > zscan16(ux + 1, uy - 1);
        /* get_pu_top :282 */
        cand[3] = (gy0 + uy - 1) * units_w + gx0 + ux;
        ok[3] = uy == 0 ? has_top : 1;
    }
    {   /* top-left unit: get_pu_top_left :335 */
        const int ux = px / 4, uy = py / 4;
        cand[4] = (gy0 + uy - 1) * units_w + gx0 + ux - 1;
        ok[4] = (ux == 0 && uy == 0) ? has_top_left : uy == 0 ? has_top : ux == 0 ? has_left : 1;
    }
}
void orc_amvp_candidates(const uint8_t *inter, const int16_t *mv, int units_w, int w, int h, int x, int y, int size, int32_t out[4])
{
    int cand[5], ok[5];                                                    /* A0, A1, B0, B1, B2 */
    pu_neighbours(units_w, w, h, x, y, size, cand, ok);
    int32_t list[3][2];
    int n = 0;
    for (int k = 0; k < 5; k++) ok[k] = ok[k] && inter[cand[k]];          /* add_amvp_cand :2198: mv_ref_idx >= 0, same picture */
    const int smvp = ok[0] || ok[1];
    const int a = ok[0] ? 0 : ok[1] ? 1 : -1;
    if (a >= 0) { list[n][0] = mv[2 * cand[a]]; list[n][1] = mv[2 * cand[a] + 1]; n++; }
    const int b = ok[2] ? 2 : ok[3] ? 3 : ok[4] ? 4 : -1;
    if (b >= 0) { list[n][0] = mv[2 * cand[b]]; list[n][1] = mv[2 * cand[b] + 1]; n++; }
    /* no left candidate: the above group is walked a second time with add_amvp_cand_order (:2405-2420), which takes the same neighbour's
     * vector again (same picture distance: no scaling) */
    if (!smvp && b >= 0) { list[n][0] = list[n - 1][0]; list[n][1] = list[n - 1][1]; n++; }
    if (n == 2 && list[0][0] == list[1][0] && list[0][1] == list[1][1]) n = 1;
    if (n > 2) n = 2;
    for (; n < 2; n++) { list[n][0] = 0; list[n][1] = 0; }
    out[0] = list[0][0]; out[1] = list[0][1]; out[2] = list[1][0]; out[3] = list[1][1];
}

/* ------------------------------------------------------------------------------------------
 * Merge candidates (P picture, one reference picture).  hmr_motion_inter.c:1937-2196: A1, B1, B0, A0, then B2 while fewer than four,
 * each pruned against the neighbours the reference compares it with (equal_motion :1915), the list closed at max_cands and filled
 * up with zero vectors (:2163-2192).  out = max_cands x { x, y }; every candidate refers to picture 0 of list 0.
 * ------------------------------------------------------------------------------------------ */
void orc_merge_candidates(const uint8_t *inter, const int16_t *mv, int units_w, int w, int h, int x, int y, int size, int max_cands, int32_t *out)
{
    int cand[5], ok[5];
    pu_neighbours(units_w, w, h, x, y, size, cand, ok);
    for (int k = 0; k < 5; k++) ok[k] = ok[k] && inter[cand[k]];
    #define SAME(a, b) (mv[2 * cand[a]] == mv[2 * cand[b]] && mv[2 * cand[a] + 1] == mv[2 * cand[b] + 1])
    int n = 0;
    #define TAKE(k) do { out[2 * n] = mv[2 * cand[k]]; out[2 * n + 1] = mv[2 * cand[k] + 1]; n++; } while (0)
    if (ok[1]) TAKE(1);                                                    /* A1 */
    if (n < max_cands && ok[3] && (!ok[1] || !SAME(1, 3))) TAKE(3);        /* B1 vs A1 */
    if (n < max_cands && ok[2] && (!ok[3] || !SAME(3, 2))) TAKE(2);        /* B0 vs B1 */
    if (n < max_cands && ok[0] && (!ok[1] || !SAME(1, 0))) TAKE(0);        /* A0 vs A1 */
    if (n < max_cands && n < 4 && ok[4] && (!ok[1] || !SAME(1, 4)) && (!ok[3] || !SAME(3, 4))) TAKE(4);   /* B2 vs A1 and B1 */
    for (; n < max_cands; n++) { out[2 * n] = 0; out[2 * n + 1] = 0; }
    #undef TAKE
    #undef SAME
}


/* ------------------------------------------------------------------------------------------
 * Reconstruction of intra transform units from the reconstructed picture around them.
 * fill_reference_samples hmr_motion_intra.c:246-406, the neighbour flags of the quadtree :625-657 / :676-683, then the chain of
 * encode_intra_cu :1017-1069 (luma) / hmr_motion_intra_chroma.c:325-365 (chroma).
 * ------------------------------------------------------------------------------------------ */
/* the four neighbour flags of the quadtree node of `size` luma samples at luma position (x, y), and how many samples of its left-bottom /
 * top-right runs lie inside the picture, in samples of the plane the unit is coded in (n = size for luma, size / 2 for chroma; :291, :339) */
void orc_intra_neighbours(int w, int h, int x, int y, int size, int chroma, int32_t out[6])
{
    const int ctu_x = x & ~63, ctu_y = y & ~63, px = x - ctu_x, py = y - ctu_y;
    const int cols = (w + 63) / 64;
    int l = ctu_x > 0, t = ctu_y > 0, lb = 0, tr = ctu_y > 0 && ctu_x / 64 + 1 < cols;
    const int valid_lines = h - ctu_y < 64 ? h - ctu_y : 64, valid_cols = w - ctu_x < 64 ? w - ctu_x : 64;
    int par_x = 0, par_y = 0;
    for (int s = 32; s >= size; s >>= 1) {
        const int cx = par_x + ((px - par_x) >= s ? s : 0), cy = par_y + ((py - par_y) >= s ? s : 0);
        const int nlb = (lb && cx == par_x) || (l && cx == par_x && cy == par_y && valid_lines > cy + s);
        const int ntr = (tr && cy == par_y) || (t && cx == par_x && cy == par_y && valid_cols > cx + s) || (cx == par_x && cy != par_y && valid_cols > cx + s);
        l = l || cx; t = t || cy; lb = nlb; tr = ntr; par_x = cx; par_y = cy;
    }
    const int n = chroma ? size / 2 : size, ph = chroma ? h / 2 : h, pw = chroma ? w / 2 : w, gx = chroma ? x / 2 : x, gy = chroma ? y / 2 : y;
    int lbs = ph - (gy + n), trs = pw - (gx + n);
    if (lbs > n) lbs = n;
    if (trs > n) trs = n;
    out[0] = l; out[1] = t; out[2] = lb; out[3] = tr; out[4] = lbs; out[5] = trs;
}

/* rec: the plane's sample (0, 0); the unit's n x n block sits at (x, y).  adi: 4n + 1 samples, index 2n = corner, 2n + 1 + i above,
 * 2n - 1 - r the left column at row y + r.  A restatement of :246-406 statement for statement: the copies, then the two padding runs. */
void orc_intra_fill_reference(const int16_t *rec, int stride, int x, int y, int n, const int32_t nb[6], int16_t *adi)
{
    const int left = nb[0], top = nb[1], left_bottom = nb[2], top_right = nb[3], lbs = nb[4], trs = nb[5];
    if (!left && !top) { for (int i = 0; i < 4 * n + 1; i++) adi[i] = 128; return; }
    const int16_t *corner = rec + (y - 1) * stride + (x - 1);          /* decoded_buff - stride - 1 */
    int16_t first = 0, last = 0;
    int16_t *pad_left = NULL, *pad_top = NULL;
    int pad_left_n = 0, pad_top_n = 0;
    int16_t *ptr = adi + n;
    if (left) {
        for (int i = 0; i < n; i++) *ptr++ = corner[(n - i) * stride];
        first = ptr[-n]; last = ptr[-1];
    } else { pad_left = ptr; pad_left_n = n; }
    ptr = adi + n - 1;
    if (left_bottom) {
        for (int i = 0; i < lbs; i++) *ptr-- = corner[(n + 1 + i) * stride];
        first = ptr[1];
        if (lbs != n) { pad_left = adi; pad_left_n = n - lbs; }
    } else {
        pad_left = adi;
        if (left) pad_left_n = n; else pad_left_n += n;
    }
    ptr = adi + 2 * n + 1;
    const int16_t *rp = corner + 1;
    if (top) {
        for (int i = 0; i < n; i++) *ptr++ = *rp++;
        if (!left) first = ptr[-n];
        last = ptr[-1];
    } else { pad_top = ptr; pad_top_n = n; }
    if (top_right) {
        for (int i = 0; i < trs; i++) *ptr++ = *rp++;
        last = ptr[-1];
        if (trs != n) { pad_top = ptr; pad_top_n = n - trs; }
    } else {
        if (top) { pad_top = ptr; pad_top_n = n; } else pad_top_n += n;
    }
    if (left && top) adi[2 * n] = corner[0];
    else if (left) { pad_top--; pad_top_n++; }
    else pad_left_n++;
    for (int i = 0; i < pad_left_n; i++) *pad_left++ = first;
    for (int i = 0; i < pad_top_n; i++) *pad_top++ = last;
}

/* the units in coding order; every unit predicts from what the earlier ones left in rec.  tu: n x { comp, x, y, size, mode, qp (of the
 * component), scan_mode, luma x, luma y, luma size of the quadtree node whose neighbour flags apply }.  coeff: the levels back to back. */
void orc_intra_recon_tus(const orc_tables *tab, const int16_t *const orig[3], const int orig_stride[3], int16_t *const rec[3], const int rec_stride[3],
                         int w, int h, const int32_t *tu, int n_tus, int is_islice, int sign_hiding, double chroma_weight, int16_t *coeff, orc_tu_out *res)
{
    int16_t adi[4 * 32 + 1], flt[4 * 32 + 1], pred[32 * 32];
    for (int i = 0; i < n_tus; i++) {
        const int32_t *u = tu + 10 * i;
        const int comp = u[0], x = u[1], y = u[2], n = u[3], mode = u[4], qp = u[5], scan = u[6];
        int32_t nb[6];
        orc_intra_neighbours(w, h, u[7], u[8], u[9], comp != 0, nb);
        orc_intra_fill_reference(rec[comp], rec_stride[comp], x, y, n, nb, adi);
        const int16_t *use = adi;
        if (comp == 0 && orc_intra_uses_filtered(n, mode)) { orc_adi_filter(adi, flt, n, 1); use = flt; }
        orc_intra_predict(use, n, mode, comp == 0, pred, n);
        orc_encode_intra_tu(tab, orig[comp] + y * orig_stride[comp] + x, orig_stride[comp], pred, n, coeff, rec[comp] + y * rec_stride[comp] + x, rec_stride[comp],
                            n, comp, qp, scan, is_islice, sign_hiding, comp ? chroma_weight : 1.0, &res[i]);
        coeff += n * n;
    }
}
