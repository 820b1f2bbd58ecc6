/*
 * ref_driver.c -- thin harness around the UNMODIFIED reference (HomerHEVC), compiled against the headers
 * where they lie under /root/reference/src/homer_lib and linked with oracle/_ref/libhomer_ref.so.
 * TEST INFRASTRUCTURE ONLY: it exists so that tests (and bench.py --impl reference) can call the reference's
 * own functions that need a live henc_thread_t (quant, inv_quant, motion estimation, motion compensation,
 * the inter T/Q chain) and so that a whole-encode golden can be produced in lock step.
 * The simple kernels (sse_aligned_sad, sse_transform, ...) are called straight from libhomer_ref.so via ctypes.
 *
 * Build: see oracle/Makefile (outputs only into oracle/_ref/).
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include <time.h>

#include "hmr_private.h"
#include "hmr_common.h"
#include "hmr_sse42_functions.h"

/* reference functions without a prototype in its headers */
int encode_inter_cu(henc_thread_t *et, ctu_info_t *ctu, cu_partition_info_t *curr_cu_info, int depth, PartSize part_size_type, int *curr_sum, int gcnt);
int encode_inter_cu_chroma(henc_thread_t *et, ctu_info_t *ctu, cu_partition_info_t *curr_cu_info, int component, int depth, PartSize part_size_type, int *curr_sum, int gcnt);
uint32_t hmr_motion_estimation(henc_thread_t *et, ctu_info_t *ctu, cu_partition_info_t *curr_cu_info, int16_t *orig_buff, int orig_buff_stride, int16_t *reference_buff, int reference_buff_stride, int curr_part_global_x,
                               int curr_part_global_y, int init_x, int init_y, int curr_part_size, int curr_part_size_shift, int search_range_x, int search_range_y, int frame_size_x, int frame_size_y, motion_vector_t *mv, motion_vector_t *subpix_mv, mv_candiate_list_t *amvp_candidate_list, uint32_t threshold, unsigned int action);
void hmr_motion_compensation_luma(henc_thread_t *et, cu_partition_info_t *curr_cu_info, int16_t *reference_buff, int reference_buff_stride, int16_t *pred_buff, int pred_buff_stride, int width, int height, int curr_part_size_shift, motion_vector_t *mv, int is_bi_predict);
void hmr_motion_compensation_chroma(henc_thread_t *et, int16_t *reference_buff, int reference_buff_stride, int16_t *pred_buff, int pred_buff_stride, int curr_part_size, int curr_part_size_shift, motion_vector_t *mv, int is_bi_predict);

typedef struct refdrv {
    void *handle;
    hvenc_enc_t *enc;
    hvenc_engine_t *eng;
    henc_thread_t *et;
    HVENC_Cfg cfg;
} refdrv;

static int g_quiet_fd = -1;
static void hush(void)   { fflush(stdout); g_quiet_fd = dup(1); int n = open("/dev/null", 1); dup2(n, 1); close(n); }
static void unhush(void) { fflush(stdout); if (g_quiet_fd >= 0) { dup2(g_quiet_fd, 1); close(g_quiet_fd); g_quiet_fd = -1; } }

static void default_cfg(HVENC_Cfg *cfg, int width, int height, int qp, int sign_hiding)
{
    memset(cfg, 0, sizeof *cfg);
    cfg->size = sizeof *cfg;
    cfg->width = width; cfg->height = height;
    cfg->profile = 1;
    cfg->intra_period = 100; cfg->gop_size = 1; cfg->num_b = 0;
    cfg->motion_estimation_precision = QUARTER_PEL;
    cfg->qp = qp; cfg->frame_rate = 25; cfg->num_ref_frames = 1; cfg->cu_size = 64;
    cfg->max_pred_partition_depth = 4; cfg->max_intra_tr_depth = 2; cfg->max_inter_tr_depth = 1;
    cfg->num_enc_engines = 1; cfg->wfpp_enable = 0; cfg->wfpp_num_threads = 1;
    cfg->sign_hiding = sign_hiding; cfg->sample_adaptive_offset = 1;
    cfg->rd_mode = RD_FAST; cfg->bitrate_mode = BR_FIXED_QP; cfg->bitrate = 1250;
    cfg->vbv_size = 1250; cfg->vbv_init = 437; cfg->chroma_qp_offset = 2;
    cfg->reinit_gop_on_scene_change = 1; cfg->performance_mode = PERF_FASTER_COMPUTATION;
}

refdrv *refdrv_open(int width, int height, int qp, int sign_hiding)
{
    refdrv *d = (refdrv *)calloc(1, sizeof *d);
    hush();
    d->handle = HOMER_enc_init();
    default_cfg(&d->cfg, width, height, qp, sign_hiding);
    int ok = HOMER_enc_control(d->handle, HOMER_SETCFG, &d->cfg);
    unhush();
    if (!ok) { free(d); return NULL; }
    d->enc = (hvenc_enc_t *)d->handle;
    d->eng = d->enc->encoder_engines[0];
    d->et = d->eng->thread[0];
    return d;
}

void refdrv_close(refdrv *d)
{
    if (!d) return;
    hush();
    HOMER_enc_close(d->handle);
    unhush();
    free(d);
}

/* 1 when the CPUID branch picked the SSE4.2 table (hmr_encoder_lib.c:155-186) */
int refdrv_sse_selected(refdrv *d) { return d->enc->funcs.sad == sse_aligned_sad; }

const uint32_t *refdrv_scan(refdrv *d, int mode, int log2n) { return d->enc->scan_pyramid[mode][log2n - 1]; }
const int32_t *refdrv_quant_table(refdrv *d, int log2n, int list, int rem) { return d->enc->quant_pyramid[log2n - 2][list][rem]; }
const int32_t *refdrv_dequant_table(refdrv *d, int log2n, int list, int rem) { return d->enc->dequant_pyramid[log2n - 2][list][rem]; }

static void set_slice(refdrv *d, int is_islice, int sign_hiding)
{
    d->eng->current_pict.slice.slice_type = is_islice ? I_SLICE : P_SLICE;
    d->et->pps->sign_data_hiding_flag = sign_hiding;
}

/* src/dst must be 16-byte aligned (aligned SSE loads, hmr_sse42_functions_quant.c:59-82) */
void refdrv_quant(refdrv *d, int16_t *src, int16_t *dst, int16_t *delta_u_out, int scan_mode, int log2n, int comp,
                  int is_intra, int is_islice, int sign_hiding, int per, int rem, int *ac_sum)
{
    const int n = 1 << log2n;
    const int depth = d->et->max_cu_size_shift - log2n - (comp != Y_COMP);
    set_slice(d, is_islice, sign_hiding);
    d->enc->funcs.quant(d->et, src, dst, scan_mode, depth, comp, REG_DCT, is_intra, ac_sum, n, per, rem);
    if (delta_u_out) memcpy(delta_u_out, d->et->aux_buff, sizeof(int16_t) * n * n);
}

void refdrv_inv_quant(refdrv *d, int16_t *src, int16_t *dst, int log2n, int comp, int is_intra, int per, int rem)
{
    const int n = 1 << log2n;
    const int depth = d->et->max_cu_size_shift - log2n - (comp != Y_COMP);
    d->enc->funcs.inv_quant(d->et, src, dst, depth, comp, is_intra, n, per, rem);
}

/* the plain-C twins, to document where C and SSE4.2 differ */
void refdrv_quant_plainc(refdrv *d, int16_t *src, int16_t *dst, int scan_mode, int log2n, int comp,
                         int is_intra, int is_islice, int sign_hiding, int per, int rem, int *ac_sum)
{
    const int n = 1 << log2n;
    const int depth = d->et->max_cu_size_shift - log2n - (comp != Y_COMP);
    set_slice(d, is_islice, sign_hiding);
    quant(d->et, src, dst, scan_mode, depth, comp, REG_DCT, is_intra, ac_sum, n, per, rem);
}

/* out[0..3] = mv.x, mv.y, subpix.x, subpix.y; returns best SAD.  cands are (x,y) pairs in quarter-pel units. */
uint32_t refdrv_motion_estimation(refdrv *d, int16_t *orig, int orig_stride, int16_t *ref, int ref_stride,
                                  int gx, int gy, int size, int frame_w, int frame_h,
                                  int n_amvp, const int32_t *amvp, int n_start, const int32_t *start,
                                  int qp, double avg_dist, unsigned action, int32_t *out)
{
    henc_thread_t *et = d->et;
    cu_partition_info_t cu;
    ctu_info_t ctu;
    mv_candiate_list_t amvp_list;
    motion_vector_t mv = { 0, 0 }, sub = { 0, 0 };
    int shift = 0;
    memset(&cu, 0, sizeof cu); memset(&ctu, 0, sizeof ctu); memset(&amvp_list, 0, sizeof amvp_list);
    while ((1 << shift) < size) shift++;
    cu.size = (uint16_t)size; cu.qp = (uint32_t)qp;
    cu.x_position = (uint16_t)(gx & 63); cu.y_position = (uint16_t)(gy & 63);
    if (cu.x_position + size > 64) cu.x_position = 0;
    if (cu.y_position + size > 64) cu.y_position = 0;
    amvp_list.num_mv_candidates = n_amvp;
    for (int i = 0; i < n_amvp; i++) { amvp_list.mv_candidates[i].mv.hor_vector = amvp[2 * i]; amvp_list.mv_candidates[i].mv.ver_vector = amvp[2 * i + 1]; }
    et->mv_search_candidates.num_mv_candidates = n_start;
    for (int i = 0; i < n_start; i++) { et->mv_search_candidates.mv_candidates[i].mv.hor_vector = start[2 * i]; et->mv_search_candidates.mv_candidates[i].mv.ver_vector = start[2 * i + 1]; }
    d->eng->avg_dist = avg_dist;
    uint32_t best = hmr_motion_estimation(et, &ctu, &cu, orig, orig_stride, ref, ref_stride, gx, gy, 0, 0, size, shift,
                                          MOTION_SEARCH_RANGE_X, MOTION_SEARCH_RANGE_Y, frame_w, frame_h, &mv, &sub, &amvp_list, 0, action);
    out[0] = mv.hor_vector; out[1] = mv.ver_vector; out[2] = sub.hor_vector; out[3] = sub.ver_vector;
    return best;
}

void refdrv_mc_luma(refdrv *d, int16_t *ref, int ref_stride, int16_t *pred, int pred_stride, int size, int mvx, int mvy)
{
    motion_vector_t mv = { mvx, mvy };
    cu_partition_info_t cu;
    int shift = 0;
    memset(&cu, 0, sizeof cu);
    while ((1 << shift) < size) shift++;
    hmr_motion_compensation_luma(d->et, &cu, ref, ref_stride, pred, pred_stride, size, size, shift, &mv, 0);
}

void refdrv_mc_chroma(refdrv *d, int16_t *ref, int ref_stride, int16_t *pred, int pred_stride, int size, int mvx, int mvy)
{
    motion_vector_t mv = { mvx, mvy };
    int shift = 0;
    while ((1 << shift) < size) shift++;
    hmr_motion_compensation_chroma(d->et, ref, ref_stride, pred, pred_stride, size, shift, &mv, 0);
}

/* One inter TU through the reference's own chain (hmr_motion_inter.c:40 / :133).
 * orig/pred: n x n samples, row stride n.  depth: CU depth (1 -> 32x32 ... 4 -> 4x4 luma); part: index inside that depth.
 * comp 0 luma, 1 U, 2 V (chroma TU is n/2, taken from the same CU; a 4x4 luma CU has no chroma TU of its own).
 * Returns the function's return value (ssd); coeff_out n*n levels, dec_out n*n samples. */
int refdrv_encode_inter_tu(refdrv *d, const int16_t *orig, const int16_t *pred, int depth, int part, int comp, int qp,
                           int is_islice, int sign_hiding, double avg_dist, int16_t *coeff_out, int16_t *dec_out, int *sum_out)
{
    henc_thread_t *et = d->et;
    ctu_info_t *ctu = &d->eng->ctu_info[0];
    cu_partition_info_t *cu = &ctu->partition_list[et->partition_depth_start[depth]] + part;
    const int n = comp == Y_COMP ? cu->size : cu->size_chroma;
    const int px = comp == Y_COMP ? cu->x_position : cu->x_position_chroma;
    const int py = comp == Y_COMP ? cu->y_position : cu->y_position_chroma;
    int sum = 0, ret;
    wnd_t *quant_wnd = et->transform_quant_wnd[depth + 1];
    wnd_t *dec_wnd = et->decoded_mbs_wnd[depth + 1];

    set_slice(d, is_islice, sign_hiding);
    d->eng->current_pict.slice.qp = qp;
    d->eng->avg_dist = avg_dist;
    cu->qp = (uint32_t)qp;

    int16_t *o = WND_POSITION_2D(int16_t *, et->curr_mbs_wnd, comp, px, py, 0, et->ctu_width);
    int16_t *p = WND_POSITION_2D(int16_t *, et->prediction_wnd[0], comp, px, py, 0, et->ctu_width);
    int16_t *r = WND_POSITION_2D(int16_t *, et->residual_wnd, comp, px, py, 0, et->ctu_width);
    const int os = WND_STRIDE_2D(et->curr_mbs_wnd, comp), ps = WND_STRIDE_2D(et->prediction_wnd[0], comp), rs = WND_STRIDE_2D(et->residual_wnd, comp);
    for (int y = 0; y < n; y++) {
        memcpy(o + y * os, orig + y * n, sizeof(int16_t) * n);
        memcpy(p + y * ps, pred + y * n, sizeof(int16_t) * n);
    }
    et->funcs->predict(o, os, p, ps, r, rs, n);            /* predict_inter, hmr_motion_inter.c:3059-3061 */
    if (comp == Y_COMP)
        ret = encode_inter_cu(et, ctu, cu, depth, SIZE_2Nx2N, &sum, 0);
    else
        ret = encode_inter_cu_chroma(et, ctu, cu, comp, depth, SIZE_2Nx2N, &sum, 0);

    int16_t *q = comp == Y_COMP
        ? WND_POSITION_1D(int16_t *, *quant_wnd, comp, 0, et->ctu_width, (cu->abs_index << et->num_partitions_in_cu_shift))
        : WND_POSITION_1D(int16_t *, *quant_wnd, comp, 0, et->ctu_width, (cu->abs_index << et->num_partitions_in_cu_shift) >> 2);
    memcpy(coeff_out, q, sizeof(int16_t) * n * n);
    int16_t *dd = WND_POSITION_2D(int16_t *, *dec_wnd, comp, px, py, 0, et->ctu_width);
    const int ds = WND_STRIDE_2D(*dec_wnd, comp);
    for (int y = 0; y < n; y++) memcpy(dec_out + y * n, dd + y * ds, sizeof(int16_t) * n);
    *sum_out = sum;
    return ret;
}

int refdrv_chroma_qp(refdrv *d, int qp)
{
    extern const uint8_t chroma_scale_conversion_table[];
    int v = qp + d->eng->chroma_qp_offset;
    return chroma_scale_conversion_table[v < 0 ? 0 : (v > 57 ? 57 : v)];
}

/* ------------------------------------------------------------------------------------------------------------
 * Lock-step whole-encode driver: feed one frame, spin until its NAL units arrive, append them,
 * repeat; then HOMER_END and drain.  yuv: planar 4:2:0 8-bit frames back to back.  Returns bytes written to
 * `bitstream` (capacity cap) or -1.  recon (optional) receives the reconstructed frames, same layout as yuv.
 * force_intra != 0 sets image_type = IMAGE_I on every frame (the only way to get all-intra, hmr_encoder_lib.c:311).
 * hook (optional) is called once after SETCFG with the address of the encoder's low_level_funcs_t so that a
 * replacement table can be installed (this is the drop-in boundary, hmr_private.h:1063 / :1443).
 * ------------------------------------------------------------------------------------------------------------ */
typedef void (*refdrv_table_hook)(void *funcs_table, void *user);

long refdrv_encode_lockstep(int width, int height, int n_frames, const uint8_t *yuv, int qp, int sign_hiding,
                            int force_intra, int performance_mode, uint8_t *bitstream, long cap, uint8_t *recon,
                            refdrv_table_hook hook, void *hook_user, double *seconds_out)
{
    HVENC_Cfg cfg;
    encoder_in_out_t in, out_frame, out_stream;
    nalu_t *nalus[8];
    uint32_t n_nalus = 0;
    long written = 0;
    const long ysz = (long)width * height, csz = ysz >> 2, fsz = ysz + 2 * csz;
    int got = 0;

    hush();
    void *h = HOMER_enc_init();
    default_cfg(&cfg, width, height, qp, sign_hiding);
    if (performance_mode >= 0) cfg.performance_mode = performance_mode;
    if (!HOMER_enc_control(h, HOMER_SETCFG, &cfg)) { unhush(); return -1; }
    if (hook) hook(&((hvenc_enc_t *)h)->funcs, hook_user);

    memset(&in, 0, sizeof in); memset(&out_frame, 0, sizeof out_frame); memset(&out_stream, 0, sizeof out_stream);
    out_stream.stream.streams[0] = (uint8_t *)malloc(0x2000000);
    if (recon) for (int c = 0; c < 3; c++) out_frame.stream.streams[c] = (uint8_t *)malloc(c ? csz : ysz);

    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int f = 0; f <= n_frames; f++) {
        if (f < n_frames) {
            const uint8_t *fr = yuv + f * fsz;
            in.stream.streams[0] = (uint8_t *)fr; in.stream.streams[1] = (uint8_t *)fr + ysz; in.stream.streams[2] = (uint8_t *)fr + ysz + csz;
            in.stream.data_stride[0] = width; in.stream.data_stride[1] = in.stream.data_stride[2] = width / 2;
            in.pts = f; in.image_type = force_intra ? IMAGE_I : IMAGE_AUTO;
            HOMER_enc_encode(h, &in);
        } else {
            HOMER_enc_control(h, HOMER_END, NULL);
        }
        /* spin until this frame's output set is there */
        for (long spins = 0; got < (f < n_frames ? f + 1 : n_frames) && spins < 200000000L; spins++) {
            n_nalus = 8;
            HOMER_enc_get_coded_frame(h, &out_frame, nalus, &n_nalus);
            if (n_nalus > 0) {
                HOMER_enc_write_annex_b_output(nalus, n_nalus, &out_stream);
                const long sz = out_stream.stream.data_size[0];
                if (written + sz > cap) { unhush(); return -1; }
                memcpy(bitstream + written, out_stream.stream.streams[0], sz);
                written += sz;
                if (recon) {
                    uint8_t *r = recon + got * fsz;
                    memcpy(r, out_frame.stream.streams[0], ysz);
                    memcpy(r + ysz, out_frame.stream.streams[1], csz);
                    memcpy(r + ysz + csz, out_frame.stream.streams[2], csz);
                }
                got++;
            } else {
                usleep(400);          /* poll gently: the encoder thread owns the output queue */
            }
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds_out) *seconds_out = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    HOMER_enc_close(h);
    unhush();
    free(out_stream.stream.streams[0]);
    if (recon) for (int c = 0; c < 3; c++) free(out_frame.stream.streams[c]);
    return got == n_frames ? written : -1;
}

/* ------------------------------------------------------------------------------------------------------------
 * The frame-level pre-pass of include/homer_b200.h section D computed with the reference's OWN functions
 * (hmr_motion_estimation, hmr_motion_compensation_luma/_chroma, predict, encode_inter_cu/_chroma) on host threads:
 * the CPU arm of bench.py (--impl reference, cpu_baseline) and the strongest parity check of the GPU pre-pass.
 * One encoder instance per thread (each owns its henc_thread_t scratch); CTUs are dealt round-robin.
 * Output layouts equal the GPU library's: ME tables in PU raster order per depth, TU tables / levels in raster
 * order of the coded TUs per (pass, plane), reconstructions and predictions as tight 8-bit planes.
 * ------------------------------------------------------------------------------------------------------------ */
#include <pthread.h>

typedef struct rp_me { int32_t mvx, mvy, subx, suby; uint32_t sad, n_probes; } rp_me;
typedef struct rp_tu { int32_t sum; uint32_t ssd, ssd_zero; int32_t zeroed; } rp_tu;
typedef struct refdrv_prepass_out {
    rp_me   *me[4];
    rp_tu   *tu[5][3];
    int16_t *coeff[5][3];
    uint8_t *recon[5][3];
    uint8_t *pred[4][3];
} refdrv_prepass_out;

typedef struct rp_shared {
    int w, h, qp, ctu_cols, ctu_rows, row0, rows, n_threads;
    double avg_dist;
    int16_t *cur[3], *ref[3];           /* padded int16 planes (origin pointers) */
    int stride[3];
    const uint8_t *cur8[3], *ref8[3];   /* the caller's 8-bit planes */
    pthread_barrier_t planes_ready;
    int *tu_index[5][3];                /* raster position -> index among coded TUs, or -1 */
    int tu_grid_w[5][3];
    refdrv_prepass_out *out;
} rp_shared;
typedef struct rp_worker { rp_shared *sh; refdrv *drv; int tid; pthread_t th; } rp_worker;

static const int rp_pass_tu[5] = { 32, 32, 16, 8, 4 };
static int rp_pass_depth(int p) { return p < 4 ? p : 3; }
static int rp_pu_valid(const rp_shared *sh, int x, int y, int s)
{
    const int row = y / 64;
    return x + s <= sh->w && y + s <= sh->h && row >= sh->row0 && row < sh->row0 + sh->rows;
}

/* the reference's frames are int16 planes with a replicated border (put_frame_to_encode widens the 8-bit input, hmr_encoder_lib.c:293-305);
 * the workers fill disjoint row slices of the six planes before the CTU loop starts (rp_fill_rows), so that the conversion does
 * not serialise the multi-threaded run */
static int16_t *rp_alloc_plane(int w, int h, int pad, int *stride_out, int16_t **alloc_out)
{
    const int stride = w + 2 * pad;
    int16_t *a = (int16_t *)malloc(sizeof(int16_t) * (size_t)stride * (h + 2 * pad));
    *stride_out = stride; *alloc_out = a;
    return a + (size_t)pad * stride + pad;
}
static void rp_fill_rows(int16_t *org, int stride, const uint8_t *src, int w, int h, int pad, int part, int parts)
{
    const int rows = h + 2 * pad, lo = (int)((long)rows * part / parts) - pad, hi = (int)((long)rows * (part + 1) / parts) - pad;
    for (int y = lo; y < hi; y++) {
        const int sy = y < 0 ? 0 : (y >= h ? h - 1 : y);
        const uint8_t *s = src + (size_t)sy * w;
        int16_t *row = org + (ptrdiff_t)y * stride;
        for (int x = -pad; x < 0; x++) row[x] = s[0];
        for (int x = 0; x < w; x++) row[x] = s[x];
        for (int x = w; x < w + pad; x++) row[x] = s[w - 1];
    }
}

static cu_partition_info_t *rp_find_cu(henc_thread_t *et, ctu_info_t *ctu, int depth, int x, int y)
{
    const int n = 1 << (2 * depth);
    cu_partition_info_t *c = &ctu->partition_list[et->partition_depth_start[depth]];
    for (int i = 0; i < n; i++) if (c[i].x_position == x && c[i].y_position == y) return &c[i];
    return NULL;
}

static void *rp_thread(void *arg)
{
    rp_worker *wk = (rp_worker *)arg;
    rp_shared *sh = wk->sh;
    refdrv *d = wk->drv;
    henc_thread_t *et = d->et;
    ctu_info_t *ctu = &d->eng->ctu_info[0];
    refdrv_prepass_out *out = sh->out;
    d->eng->avg_dist = sh->avg_dist;
    d->eng->current_pict.slice.slice_type = P_SLICE;
    d->eng->current_pict.slice.qp = sh->qp;

    for (int c = 0; c < 3; c++) {
        const int pw = c ? sh->w / 2 : sh->w, ph = c ? sh->h / 2 : sh->h, pad = c ? 72 : 144;
        rp_fill_rows(sh->cur[c], sh->stride[c], sh->cur8[c], pw, ph, pad, wk->tid, sh->n_threads);
        rp_fill_rows(sh->ref[c], sh->stride[c], sh->ref8[c], pw, ph, pad, wk->tid, sh->n_threads);
    }
    pthread_barrier_wait(&sh->planes_ready);

    for (int ci = sh->row0 * sh->ctu_cols + wk->tid; ci < (sh->row0 + sh->rows) * sh->ctu_cols; ci += sh->n_threads) {
        const int x0 = (ci % sh->ctu_cols) * 64, y0 = (ci / sh->ctu_cols) * 64;
        /* frame -> CTU window (mem_transfer_move_curr_ctu_group, hmr_mem_transfer.c:284) */
        for (int c = 0; c < 3; c++) {
            const int cs = c ? 32 : 64, cx = c ? x0 / 2 : x0, cy = c ? y0 / 2 : y0;
            int16_t *dst = WND_DATA_PTR(int16_t *, et->curr_mbs_wnd, c);
            const int ds = WND_STRIDE_2D(et->curr_mbs_wnd, c);
            for (int r = 0; r < cs; r++) memcpy(dst + r * ds, sh->cur[c] + (size_t)(cy + r) * sh->stride[c] + cx, sizeof(int16_t) * cs);
        }
        motion_vector_t mvs[4][64];
        int valid[4][64];
        for (int dp = 0; dp < 4; dp++) {
            const int s = 64 >> dp, per = 64 / s, gw = sh->ctu_cols * per;
            /* ---- motion search of every PU of this depth (hmr_cu_motion_estimation, hmr_motion_inter.c:2471) */
            for (int py = 0; py < per; py++) for (int px = 0; px < per; px++) {
                const int gx = x0 + px * s, gy = y0 + py * s, li = py * per + px;
                valid[dp][li] = rp_pu_valid(sh, gx, gy, s);
                if (!valid[dp][li]) continue;
                cu_partition_info_t cu;
                mv_candiate_list_t amvp;
                motion_vector_t mv = { 0, 0 }, sub = { 0, 0 };
                memset(&cu, 0, sizeof cu); memset(&amvp, 0, sizeof amvp);
                cu.size = (uint16_t)s; cu.qp = (uint32_t)sh->qp; cu.x_position = (uint16_t)(px * s); cu.y_position = (uint16_t)(py * s);
                amvp.num_mv_candidates = 2;
                et->mv_search_candidates.num_mv_candidates = 0;
                if (dp > 0) {
                    const int pl = (py / 2) * (per / 2) + px / 2;
                    if (valid[dp - 1][pl] && mvs[dp - 1][pl].hor_vector != 0 && mvs[dp - 1][pl].ver_vector != 0)
                        et->mv_search_candidates.mv_candidates[et->mv_search_candidates.num_mv_candidates++].mv = mvs[dp - 1][pl];
                }
                int16_t *orig = WND_POSITION_2D(int16_t *, et->curr_mbs_wnd, Y_COMP, px * s, py * s, 0, et->ctu_width);
                int16_t *rf = sh->ref[0] + (size_t)gy * sh->stride[0] + gx;
                const uint32_t sad_ = hmr_motion_estimation(et, ctu, &cu, orig, WND_STRIDE_2D(et->curr_mbs_wnd, Y_COMP), rf, sh->stride[0], gx, gy, 0, 0, s, 6 - dp,
                                                            MOTION_SEARCH_RANGE_X, MOTION_SEARCH_RANGE_Y, sh->w, sh->h, &mv, &sub, &amvp, 0,
                                                            MOTION_PEL_MASK | MOTION_HALF_PEL_MASK | MOTION_QUARTER_PEL_MASK);
                mvs[dp][li] = mv;
                if (out->me[dp]) {
                    rp_me *m = &out->me[dp][(gy / s) * gw + gx / s];
                    m->mvx = mv.hor_vector; m->mvy = mv.ver_vector; m->subx = sub.hor_vector; m->suby = sub.ver_vector; m->sad = sad_; m->n_probes = 0;
                }
                /* ---- motion compensation into the CTU prediction window (predict_inter, hmr_motion_inter.c:3047-3049) */
                hmr_motion_compensation_luma(et, &cu, rf, sh->stride[0], WND_POSITION_2D(int16_t *, et->prediction_wnd[0], Y_COMP, px * s, py * s, 0, et->ctu_width),
                                             WND_STRIDE_2D(et->prediction_wnd[0], Y_COMP), s, s, 6 - dp, &mv, 0);
                for (int c = 1; c < 3; c++)
                    hmr_motion_compensation_chroma(et, sh->ref[c] + (size_t)(gy / 2) * sh->stride[c] + gx / 2, sh->stride[c],
                                                   WND_POSITION_2D(int16_t *, et->prediction_wnd[0], c, px * s / 2, py * s / 2, 0, et->ctu_width),
                                                   WND_STRIDE_2D(et->prediction_wnd[0], c), s / 2, 5 - dp, &mv, 0);
                if (out->pred[dp][0]) for (int c = 0; c < 3; c++) {
                    const int n = c ? s / 2 : s, ox = c ? gx / 2 : gx, oy = c ? gy / 2 : gy, pw = c ? sh->w / 2 : sh->w;
                    int16_t *p = WND_POSITION_2D(int16_t *, et->prediction_wnd[0], c, c ? px * s / 2 : px * s, c ? py * s / 2 : py * s, 0, et->ctu_width);
                    const int ps = WND_STRIDE_2D(et->prediction_wnd[0], c);
                    for (int r = 0; r < n; r++) for (int q = 0; q < n; q++) out->pred[dp][c][(size_t)(oy + r) * pw + ox + q] = (uint8_t)p[r * ps + q];
                }
            }
            /* ---- T/Q passes that use this depth's prediction */
            for (int p = 0; p < 5; p++) {
                if (rp_pass_depth(p) != dp) continue;
                const int node_depth = p == 0 ? 1 : (p == 4 ? 4 : dp);      /* 64x64 CUs are coded as four 32x32 TUs */
                const int ns = 64 >> node_depth;                              /* luma size of the coded node */
                for (int c = 0; c < 3; c++) {
                    if (c > 0 && p == 4) continue;
                    const int tu = c ? rp_pass_tu[p] / 2 : rp_pass_tu[p];
                    const int pw = c ? sh->w / 2 : sh->w;
                    for (int ny = 0; ny < 64 / ns; ny++) for (int nx = 0; nx < 64 / ns; nx++) {
                        const int lx = nx * ns, ly = ny * ns;                 /* luma position inside the CTU */
                        if (!rp_pu_valid(sh, x0 + (lx / s) * s, y0 + (ly / s) * s, s)) continue;
                        if (x0 + lx + ns > sh->w || y0 + ly + ns > sh->h) continue;
                        cu_partition_info_t *cu = rp_find_cu(et, ctu, node_depth, lx, ly);
                        const int tx = c ? lx / 2 : lx, ty = c ? ly / 2 : ly;
                        int sum = 0, ssd;
                        cu->qp = (uint32_t)sh->qp;
                        et->funcs->predict(WND_POSITION_2D(int16_t *, et->curr_mbs_wnd, c, tx, ty, 0, et->ctu_width), WND_STRIDE_2D(et->curr_mbs_wnd, c),
                                           WND_POSITION_2D(int16_t *, et->prediction_wnd[0], c, tx, ty, 0, et->ctu_width), WND_STRIDE_2D(et->prediction_wnd[0], c),
                                           WND_POSITION_2D(int16_t *, et->residual_wnd, c, tx, ty, 0, et->ctu_width), WND_STRIDE_2D(et->residual_wnd, c), tu);
                        if (c == 0) ssd = encode_inter_cu(et, ctu, cu, node_depth, SIZE_2Nx2N, &sum, 0);
                        else ssd = encode_inter_cu_chroma(et, ctu, cu, c, node_depth, SIZE_2Nx2N, &sum, 0);
                        const int gxp = (c ? x0 / 2 : x0) + tx, gyp = (c ? y0 / 2 : y0) + ty;
                        const int ti = sh->tu_index[p][c][(gyp / tu) * sh->tu_grid_w[p][c] + gxp / tu];
                        if (out->tu[p][c]) { rp_tu *t = &out->tu[p][c][ti]; t->sum = sum; t->ssd = (uint32_t)ssd; t->ssd_zero = 0; t->zeroed = 0; }
                        if (out->coeff[p][c]) {
                            wnd_t *qw = et->transform_quant_wnd[node_depth + 1];
                            int16_t *q = c == 0 ? WND_POSITION_1D(int16_t *, *qw, c, 0, et->ctu_width, (cu->abs_index << et->num_partitions_in_cu_shift))
                                                : WND_POSITION_1D(int16_t *, *qw, c, 0, et->ctu_width, (cu->abs_index << et->num_partitions_in_cu_shift) >> 2);
                            memcpy(out->coeff[p][c] + (size_t)ti * tu * tu, q, sizeof(int16_t) * tu * tu);
                        }
                        if (out->recon[p][c]) {
                            wnd_t *dw = et->decoded_mbs_wnd[node_depth + 1];
                            int16_t *dd = WND_POSITION_2D(int16_t *, *dw, c, tx, ty, 0, et->ctu_width);
                            const int ds = WND_STRIDE_2D(*dw, c);
                            for (int r = 0; r < tu; r++) for (int q = 0; q < tu; q++) out->recon[p][c][(size_t)(gyp + r) * pw + gxp + q] = (uint8_t)dd[r * ds + q];
                        }
                    }
                }
            }
        }
    }
    return NULL;
}

/* number of coded TUs of (pass, plane) for a frame / band: the sizes the caller must allocate */
int refdrv_prepass_num_tus(int w, int h, int ctu_row0, int ctu_rows, int pass, int comp)
{
    rp_shared sh; memset(&sh, 0, sizeof sh);
    sh.w = w; sh.h = h; sh.ctu_cols = (w + 63) / 64; sh.ctu_rows = (h + 63) / 64;
    sh.row0 = ctu_rows > 0 ? ctu_row0 : 0; sh.rows = ctu_rows > 0 ? ctu_rows : sh.ctu_rows;
    if (comp > 0 && pass == 4) return 0;
    const int s = 64 >> rp_pass_depth(pass), tu = comp ? rp_pass_tu[pass] / 2 : rp_pass_tu[pass], sc = comp ? s / 2 : s;
    const int pw = comp ? w / 2 : w, ph = comp ? h / 2 : h;
    const int tw = sh.ctu_cols * (comp ? 32 : 64) / tu, th = sh.ctu_rows * (comp ? 32 : 64) / tu;
    int n = 0;
    for (int ty = 0; ty < th; ty++) for (int tx = 0; tx < tw; tx++) {
        const int x = tx * tu, y = ty * tu;
        if (x + tu > pw || y + tu > ph) continue;
        if (!rp_pu_valid(&sh, (x / sc) * s, (y / sc) * s, s)) continue;
        n++;
    }
    return n;
}

/* returns the seconds spent (conversion of the 8-bit planes to the reference's int16 frames included, done by the workers) or < 0 */
double refdrv_prepass(refdrv **drv, int n_threads, const uint8_t *const cur[3], const uint8_t *const ref[3], int w, int h, int qp,
                      double avg_dist, int ctu_row0, int ctu_rows, refdrv_prepass_out *out)
{
    rp_shared sh;
    rp_worker *wk = (rp_worker *)calloc((size_t)n_threads, sizeof *wk);
    int16_t *alloc[6];
    struct timespec t0, t1;
    memset(&sh, 0, sizeof sh);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    sh.w = w; sh.h = h; sh.qp = qp; sh.avg_dist = avg_dist; sh.n_threads = n_threads; sh.out = out;
    sh.ctu_cols = (w + 63) / 64; sh.ctu_rows = (h + 63) / 64;
    sh.row0 = ctu_rows > 0 ? ctu_row0 : 0; sh.rows = ctu_rows > 0 ? ctu_rows : sh.ctu_rows;
    for (int c = 0; c < 3; c++) {
        const int pw = c ? w / 2 : w, ph = c ? h / 2 : h, pad = c ? 72 : 144;      /* room for partial CTUs at the bottom/right edge */
        (void)ph;
        sh.cur[c] = rp_alloc_plane(pw, c ? h / 2 : h, pad, &sh.stride[c], &alloc[c]);
        sh.ref[c] = rp_alloc_plane(pw, c ? h / 2 : h, pad, &sh.stride[c], &alloc[3 + c]);
        sh.cur8[c] = cur[c]; sh.ref8[c] = ref[c];
    }
    pthread_barrier_init(&sh.planes_ready, NULL, (unsigned)n_threads);
    for (int p = 0; p < 5; p++) for (int c = 0; c < 3; c++) {
        if (c > 0 && p == 4) continue;
        const int s = 64 >> rp_pass_depth(p), tu = c ? rp_pass_tu[p] / 2 : rp_pass_tu[p], sc = c ? s / 2 : s;
        const int pw = c ? w / 2 : w, ph = c ? h / 2 : h;
        const int tw = sh.ctu_cols * (c ? 32 : 64) / tu, th = sh.ctu_rows * (c ? 32 : 64) / tu;
        int n = 0;
        sh.tu_grid_w[p][c] = tw;
        sh.tu_index[p][c] = (int *)malloc(sizeof(int) * (size_t)tw * th);
        for (int ty = 0; ty < th; ty++) for (int tx = 0; tx < tw; tx++) {
            const int x = tx * tu, y = ty * tu;
            int ok = !(x + tu > pw || y + tu > ph) && rp_pu_valid(&sh, (x / sc) * s, (y / sc) * s, s);
            sh.tu_index[p][c][ty * tw + tx] = ok ? n++ : -1;
        }
    }
    for (int t = 0; t < n_threads; t++) { wk[t].sh = &sh; wk[t].drv = drv[t]; wk[t].tid = t; pthread_create(&wk[t].th, NULL, rp_thread, &wk[t]); }
    for (int t = 0; t < n_threads; t++) pthread_join(wk[t].th, NULL);
    pthread_barrier_destroy(&sh.planes_ready);
    for (int i = 0; i < 6; i++) free(alloc[i]);
    for (int p = 0; p < 5; p++) for (int c = 0; c < 3; c++) free(sh.tu_index[p][c]);
    free(wk);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------------------------------
 * The INTEGRATION.md section 1 adapter, compiled: installs libhomer_b200's per-call drop-ins into the reference
 * encoder's own function table.  Used by the whole-encode parity test: the UNMODIFIED reference host code
 * (mode decision, CABAC, deblocking, SAO, rate control) runs on the CPU and every SAD / SSD / predict / reconst /
 * interpolation / transform / quant / inv_quant call it makes through the table is computed on the GPU.
 * `lib` is a dlopen() handle of libhomer_b200.so.  Usable as the `hook` of refdrv_encode_lockstep.
 * ------------------------------------------------------------------------------------------------------------ */
#include <dlfcn.h>

typedef struct gpu_quant_env { int32_t is_islice, sign_hiding, max_cu_size_shift, bit_depth; int16_t *delta_u; } gpu_quant_env;
static void (*p_hb_quant)(const gpu_quant_env *, int16_t *, int16_t *, int, int, int, int, int, int *, int, int, int);
static void (*p_hb_inv_quant)(const gpu_quant_env *, int16_t *, int16_t *, int, int, int, int, int, int);
static long g_gpu_calls;

static void gpu_quant(henc_thread_t *et, int16_t *src, int16_t *dst, int scan_mode, int depth, int comp, int cu_mode,
                      int is_intra, int *ac_sum, int cu_size, int per, int rem)
{
    gpu_quant_env env = { et->enc_engine->current_pict.slice.slice_type == I_SLICE, (int32_t)et->pps->sign_data_hiding_flag,
                          et->max_cu_size_shift, et->bit_depth, et->aux_buff };
    g_gpu_calls++;
    p_hb_quant(&env, src, dst, scan_mode, depth, comp, cu_mode, is_intra, ac_sum, cu_size, per, rem);
}
static void gpu_inv_quant(henc_thread_t *et, short *src, short *dst, int depth, int comp, int is_intra, int cu_size, int per, int rem)
{
    gpu_quant_env env = { 0, 0, et->max_cu_size_shift, et->bit_depth, NULL };
    g_gpu_calls++;
    p_hb_inv_quant(&env, src, dst, depth, comp, is_intra, cu_size, per, rem);
}

/* which = bit mask of what to route to the GPU: 1 sad/ssd, 2 predict/reconst, 4 interpolation, 8 transforms, 16 quant */
void refdrv_install_gpu_table(void *funcs_table, void *user)
{
    struct { void *lib; int which; } *u = user;
    low_level_funcs_t *f = (low_level_funcs_t *)funcs_table;
    void *lib = u->lib;
    if (u->which & 1) { f->sad = dlsym(lib, "hb_sad"); f->ssd16b = dlsym(lib, "hb_ssd16b"); }
    if (u->which & 2) { f->predict = dlsym(lib, "hb_predict"); f->reconst = dlsym(lib, "hb_reconst"); }
    if (u->which & 4) {
        f->interpolate_luma_m_compensation = dlsym(lib, "hb_interpolate_luma");
        f->interpolate_chroma_m_compensation = dlsym(lib, "hb_interpolate_chroma");
        f->interpolate_luma_m_estimation = dlsym(lib, "hb_interpolate_luma");
    }
    if (u->which & 8) { f->transform = dlsym(lib, "hb_transform"); f->itransform = dlsym(lib, "hb_itransform"); }
    if (u->which & 16) {
        p_hb_quant = dlsym(lib, "hb_quant"); p_hb_inv_quant = dlsym(lib, "hb_inv_quant");
        f->quant = gpu_quant; f->inv_quant = gpu_inv_quant;
    }
}
long refdrv_gpu_quant_calls(void) { return g_gpu_calls; }
void *refdrv_install_gpu_table_addr(void) { return (void *)refdrv_install_gpu_table; }

/* intra prediction through the reference's table (create_intra_planar_prediction / create_intra_angular_prediction);
 * adi: 4n+1 samples, pred: n*n out (stride n) */
void refdrv_intra_predict(refdrv *d, int16_t *adi, int n, int mode, int is_luma, int16_t *pred)
{
    henc_thread_t *et = d->et;
    ctu_info_t ctu;
    int shift = 0;
    memset(&ctu, 0, sizeof ctu);
    ctu.top = 1; ctu.left = 1;                                   /* fill_reference_samples always sets both (hmr_motion_intra.c:257) */
    while ((1 << shift) < n) shift++;
    if (mode == PLANAR_IDX) d->enc->funcs.create_intra_planar_prediction(et, pred, n, adi, 4 * n + 1, n, shift);
    else d->enc->funcs.create_intra_angular_prediction(et, &ctu, pred, n, adi, 4 * n + 1, n, mode, is_luma);
}
void refdrv_adi_filter(refdrv *d, int16_t *adi, int16_t *flt, int n)
{
    int shift = 0;
    while ((1 << shift) < n) shift++;
    adi_filter(adi, flt, d->et->max_cu_size_shift - shift, 4 * n + 1, n, d->et->max_cu_size_shift, d->et->sps->strong_intra_smooth_enabled_flag, d->et->bit_depth);
}

/* ------------------------------------------------------------------------------------------------------------
 * Intra mode pre-search with the reference's own functions (the loop of homer_loop1_motion_intra,
 * hmr_motion_intra.c:1084-1140, without its bit-cost terms): per block smooth the reference samples (adi_filter), and for
 * each of the 35 modes build the prediction through the function table from the raw or smoothed samples (:1122) and take
 * its SAD against the original block through the table.  jobs: n_jobs x {x, y, size}; adi: the 4*size+1 reference samples
 * of every job back to back (adi_off[j] = first sample of job j); sads: n_jobs x 35.  Jobs are split over n_threads.
 * Returns the seconds spent.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct ip_shared { const uint8_t *luma; int w, h; const int32_t *jobs; int n_jobs; const int16_t *adi; const int32_t *adi_off; uint32_t *sads; int n_threads; } ip_shared;
typedef struct ip_worker { ip_shared *sh; refdrv *drv; int tid; pthread_t th; } ip_worker;

static void *ip_thread(void *arg)
{
    static const int thr[5] = { 10, 7, 1, 0, 10 };
    ip_worker *wk = (ip_worker *)arg;
    ip_shared *sh = wk->sh;
    refdrv *d = wk->drv;
    int16_t *orig = (int16_t *)aligned_alloc(64, 64 * 64 * 2), *pred = (int16_t *)aligned_alloc(64, 64 * 64 * 2);
    int16_t *raw = (int16_t *)aligned_alloc(64, 272 * 2), *flt = (int16_t *)aligned_alloc(64, 272 * 2);
    const int lo = (int)((long)sh->n_jobs * wk->tid / sh->n_threads), hi = (int)((long)sh->n_jobs * (wk->tid + 1) / sh->n_threads);
    for (int j = lo; j < hi; j++) {
        const int x = sh->jobs[3 * j], y = sh->jobs[3 * j + 1], n = sh->jobs[3 * j + 2];
        int lg = 0;
        while ((1 << lg) < n) lg++;
        for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) orig[r * n + c] = sh->luma[(size_t)(y + r) * sh->w + x + c];
        memcpy(raw, sh->adi + sh->adi_off[j], sizeof(int16_t) * (size_t)(4 * n + 1));
        refdrv_adi_filter(d, raw, flt, n);
        for (int m = 0; m < 35; m++) {
            const int d1 = abs(m - 10), d2 = abs(m - 26);
            const int use_flt = m != 1 && ((d1 < d2 ? d1 : d2) > thr[lg - 2]);
            refdrv_intra_predict(d, use_flt ? flt : raw, n, m, 1, pred);
            sh->sads[(size_t)j * 35 + m] = d->enc->funcs.sad(orig, n, pred, n, n);
        }
    }
    free(orig); free(pred); free(raw); free(flt);
    return NULL;
}

double refdrv_intra_presearch(refdrv **drv, int n_threads, const uint8_t *luma, int w, int h, const int32_t *jobs, int n_jobs,
                              const int16_t *adi, const int32_t *adi_off, uint32_t *sads)
{
    ip_shared sh = { luma, w, h, jobs, n_jobs, adi, adi_off, sads, n_threads };
    ip_worker *wk = (ip_worker *)calloc((size_t)n_threads, sizeof *wk);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < n_threads; t++) { wk[t].sh = &sh; wk[t].drv = drv[t]; wk[t].tid = t; pthread_create(&wk[t].th, NULL, ip_thread, &wk[t]); }
    for (int t = 0; t < n_threads; t++) pthread_join(wk[t].th, NULL);
    free(wk);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* the table's weighted_average_motion (bi-prediction average) */
void refdrv_weighted_average(refdrv *d, int16_t *src0, int s0, int16_t *src1, int s1, int16_t *dst, int ds, int height, int width)
{
    d->enc->funcs.weighted_average_motion(src0, s0, src1, s1, dst, ds, height, width, 8);
}

/* one list of a bi-predicted block: the reference's motion compensation with is_bi_predict = 1 (14-bit output) */
void refdrv_mc_luma_bi(refdrv *d, int16_t *ref, int ref_stride, int16_t *pred, int pred_stride, int size, int mvx, int mvy)
{
    motion_vector_t mv = { mvx, mvy };
    cu_partition_info_t cu;
    int shift = 0;
    memset(&cu, 0, sizeof cu);
    while ((1 << shift) < size) shift++;
    hmr_motion_compensation_luma(d->et, &cu, ref, ref_stride, pred, pred_stride, size, size, shift, &mv, 1);
}
void refdrv_mc_chroma_bi(refdrv *d, int16_t *ref, int ref_stride, int16_t *pred, int pred_stride, int size, int mvx, int mvy)
{
    motion_vector_t mv = { mvx, mvy };
    int shift = 0;
    while ((1 << shift) < size) shift++;
    hmr_motion_compensation_chroma(d->et, ref, ref_stride, pred, pred_stride, size, shift, &mv, 1);
}

/* ------------------------------------------------------------------------------------------------------------
 * SAO statistics through the table's get_sao_stats (sse_sao_get_ctu_stats on x86): the encoder structures the function reads
 * are filled in by hand -- the deblocked reconstruction as curr_reference_frame, the source as img2encode, the picture size in
 * the thread, one ctu_info_t per call.  rec / org: 8-bit planes (Y, U, V) of a w x h picture.  out: per CTU (raster) and
 * component 5 types x {diff[32], count[32]} as int64, the reference's own layout (sao_stat_data_t).  Returns the CTU count.
 * ------------------------------------------------------------------------------------------------------------ */
int refdrv_sao_stats(refdrv *d, const uint8_t *const rec[3], const uint8_t *const org[3], int w, int h, int64_t *out)
{
    henc_thread_t *et = d->et;
    hvenc_engine_t *eng = et->enc_engine;
    video_frame_t fr_rec, fr_org;
    video_frame_t *save_ref = eng->curr_reference_frame, *save_in = eng->current_pict.img2encode;
    int save_w[3], save_h[3], n = 0;
    const int save_pre = eng->calculate_preblock_stats;
    memset(&fr_rec, 0, sizeof fr_rec); memset(&fr_org, 0, sizeof fr_org);
    wnd_alloc(&fr_rec.img, w, h, 80, 80, sizeof(int16_t));
    wnd_alloc(&fr_org.img, w, h, 80, 80, sizeof(int16_t));
    for (int c = 0; c < 3; c++) {
        const int pw = c ? w / 2 : w, ph = c ? h / 2 : h;
        int16_t *pr = (int16_t *)fr_rec.img.pwnd[c], *po = (int16_t *)fr_org.img.pwnd[c];
        for (int y = 0; y < ph; y++)
            for (int x = 0; x < pw; x++) {
                pr[y * fr_rec.img.window_size_x[c] + x] = rec[c][y * pw + x];
                po[y * fr_org.img.window_size_x[c] + x] = org[c][y * pw + x];
            }
        save_w[c] = et->pict_width[c]; save_h[c] = et->pict_height[c];
        et->pict_width[c] = pw; et->pict_height[c] = ph;
    }
    eng->curr_reference_frame = &fr_rec; eng->current_pict.img2encode = &fr_org; eng->calculate_preblock_stats = 0;
    for (int cy = 0; cy < h; cy += 64)
        for (int cx = 0; cx < w; cx += 64) {
            ctu_info_t ctu;
            sao_stat_data_t stats[NUM_PICT_COMPONENTS][NUM_SAO_NEW_TYPES];
            memset(&ctu, 0, sizeof ctu); memset(stats, 0, sizeof stats);
            ctu.size = 64; ctu.x[0] = cx; ctu.y[0] = cy; ctu.x[1] = ctu.x[2] = cx / 2; ctu.y[1] = ctu.y[2] = cy / 2;
            d->enc->funcs.get_sao_stats(et, &eng->current_pict.slice, &ctu, stats);
            memcpy(out + (size_t)n * NUM_PICT_COMPONENTS * NUM_SAO_NEW_TYPES * 64, stats, sizeof stats);
            n++;
        }
    eng->curr_reference_frame = save_ref; eng->current_pict.img2encode = save_in; eng->calculate_preblock_stats = save_pre;
    for (int c = 0; c < 3; c++) { et->pict_width[c] = save_w[c]; et->pict_height[c] = save_h[c]; }
    wnd_delete(&fr_rec.img); wnd_delete(&fr_org.img);
    return n;
}

/* SAO offset pass through the reference's sao_offset_ctu (hmr_sao.c:1210) for every CTU of a picture.  src: 8-bit planes of the
 * deblocked picture; types[ctu*3+comp] in -1 (off), 0..3 (EO), 4 (BO); offsets[(ctu*3+comp)*32 ..]: per-class offsets exactly as
 * sao_offset_t.offset holds them (EO: indices 0..4 for edge types -2..2, BO: one per band).  out: the finalised 8-bit planes. */
void sao_offset_ctu(henc_thread_t *wpp_thread, ctu_info_t *ctu, sao_blk_param_t *sao_blk_param);
int refdrv_sao_apply(refdrv *d, const uint8_t *const src[3], int w, int h, const int8_t *types, const int32_t *offsets, uint8_t *const out[3])
{
    henc_thread_t *et = d->et;
    hvenc_engine_t *eng = et->enc_engine;
    video_frame_t fr_rec;
    video_frame_t *save_ref = eng->curr_reference_frame;
    wnd_t save_aux = eng->sao_aux_wnd, aux;
    int save_w[3], save_h[3], n = 0;
    const int save_cu = et->max_cu_size;
    memset(&fr_rec, 0, sizeof fr_rec); memset(&aux, 0, sizeof aux);
    wnd_alloc(&fr_rec.img, w, h, 80, 80, sizeof(int16_t));
    wnd_alloc(&aux, w, h, 80, 80, sizeof(int16_t));
    for (int c = 0; c < 3; c++) {
        const int pw = c ? w / 2 : w, ph = c ? h / 2 : h;
        int16_t *pr = (int16_t *)fr_rec.img.pwnd[c], *pa = (int16_t *)aux.pwnd[c];
        for (int y = 0; y < ph; y++)
            for (int x = 0; x < pw; x++) pr[y * fr_rec.img.window_size_x[c] + x] = pa[y * aux.window_size_x[c] + x] = src[c][y * pw + x];
        save_w[c] = et->pict_width[c]; save_h[c] = et->pict_height[c];
        et->pict_width[c] = pw; et->pict_height[c] = ph;
    }
    eng->curr_reference_frame = &fr_rec; eng->sao_aux_wnd = aux; et->max_cu_size = 64;
    for (int cy = 0; cy < h; cy += 64)
        for (int cx = 0; cx < w; cx += 64) {
            ctu_info_t ctu;
            sao_blk_param_t prm;
            memset(&ctu, 0, sizeof ctu); memset(&prm, 0, sizeof prm);
            ctu.size = 64; ctu.ctu_number = n; ctu.x[0] = cx; ctu.y[0] = cy; ctu.x[1] = ctu.x[2] = cx / 2; ctu.y[1] = ctu.y[2] = cy / 2;
            for (int c = 0; c < 3; c++) {
                const int ty = types[n * 3 + c];
                prm.offsetParam[c].modeIdc = ty < 0 ? SAO_MODE_OFF : SAO_MODE_NEW;
                prm.offsetParam[c].typeIdc = ty < 0 ? 0 : ty;
                for (int k = 0; k < 32; k++) prm.offsetParam[c].offset[k] = offsets[(n * 3 + c) * 32 + k];
            }
            sao_offset_ctu(et, &ctu, &prm);
            n++;
        }
    for (int c = 0; c < 3; c++) {
        const int pw = c ? w / 2 : w, ph = c ? h / 2 : h;
        const int16_t *pr = (const int16_t *)fr_rec.img.pwnd[c];
        for (int y = 0; y < ph; y++)
            for (int x = 0; x < pw; x++) out[c][y * pw + x] = (uint8_t)pr[y * fr_rec.img.window_size_x[c] + x];
        et->pict_width[c] = save_w[c]; et->pict_height[c] = save_h[c];
    }
    eng->curr_reference_frame = save_ref; eng->sao_aux_wnd = save_aux; et->max_cu_size = save_cu;
    wnd_delete(&fr_rec.img); wnd_delete(&aux);
    return n;
}

/* ------------------------------------------------------------------------------------------------------------
 * Deblocking through the reference's own per-CTU function (hmr_deblock_filter_cu, hmr_deblocking_filter.c:737), driven the way
 * hmr_deblock_filter (:827) drives it: all vertical edges of the picture, then all horizontal ones.  `d` must have been opened
 * with the picture size.  The CTU descriptions the function reads are filled in from per-4x4-unit maps in PICTURE raster order
 * (units_w = 16 * CTU columns per row): CU depth, TU depth below the CU, intra flag, luma cbf byte, QP, list-0 vector.  The
 * boundary strengths the reference derives are handed back (per unit, picture raster) so that the pixel stage of a
 * re-implementation can be pinned on exactly the same edges.  in / out: 8-bit planes.
 * ------------------------------------------------------------------------------------------------------------ */
void hmr_deblock_filter_cu(henc_thread_t *et, slice_t *currslice, ctu_info_t *ctu, int dir);
void create_partition_ctu_neighbours(henc_thread_t *et, ctu_info_t *ctu, cu_partition_info_t *curr_partition_info);
int refdrv_deblock(refdrv *d, const uint8_t *const in[3], int w, int h, const uint8_t *cu_depth, const uint8_t *tu_depth, const uint8_t *intra,
                   const uint8_t *cbf, const uint8_t *qp, const int16_t *mv, uint8_t *bs_ver, uint8_t *bs_hor, uint8_t *const out[3])
{
    henc_thread_t *et = d->et;
    hvenc_engine_t *eng = et->enc_engine;
    slice_t *slice = &eng->current_pict.slice;
    video_frame_t fr, dummy_ref;
    video_frame_t *save_ref = eng->curr_reference_frame;
    const int cols = (w + 63) / 64, rows = (h + 63) / 64, units_w = cols * 16;
    if (eng->pict_total_ctu != cols * rows || et->pict_width[0] != w || et->pict_height[0] != h) return -1;
    memset(&fr, 0, sizeof fr); memset(&dummy_ref, 0, sizeof dummy_ref);
    wnd_alloc(&fr.img, w, h, 80, 80, sizeof(int16_t));
    for (int c = 0; c < 3; c++) {
        const int pw = c ? w / 2 : w, ph = c ? h / 2 : h;
        int16_t *p = (int16_t *)fr.img.pwnd[c];
        for (int y = 0; y < ph; y++) for (int x = 0; x < pw; x++) p[y * fr.img.window_size_x[c] + x] = in[c][y * pw + x];
    }
    eng->curr_reference_frame = &fr;
    slice->slice_type = P_SLICE; slice->sps = &d->enc->sps; slice->pps = &d->enc->pps;
    slice->deblocking_filter_disabled_flag = 0; slice->slice_beta_offset_div2 = 0; slice->slice_tc_offset_div2 = 0;
    slice->ref_pic_list[REF_PIC_LIST_0][0] = &dummy_ref;
    for (int n = 0; n < cols * rows; n++) {
        ctu_info_t *ctu = &eng->ctu_info[n];
        const int cx = n % cols, cy = n / cols;
        ctu->ctu_number = n; ctu->size = 64;
        ctu->x[0] = cx * 64; ctu->y[0] = cy * 64; ctu->x[1] = ctu->x[2] = cx * 32; ctu->y[1] = ctu->y[2] = cy * 32;
        ctu->ctu_left = cx ? &eng->ctu_info[n - 1] : NULL;
        ctu->ctu_top = cy ? &eng->ctu_info[n - cols] : NULL;
        ctu->ctu_top_left = (cx && cy) ? &eng->ctu_info[n - cols - 1] : NULL;
        ctu->ctu_top_right = (cy && cx + 1 < cols) ? &eng->ctu_info[n - cols + 1] : NULL;
        ctu->ctu_left_bottom = NULL;
        for (int r = 0; r < 256; r++) {
            const int a = eng->raster2abs_table[r];
            const int u = (cy * 16 + r / 16) * units_w + cx * 16 + r % 16;
            ctu->pred_depth[a] = cu_depth[u]; ctu->tr_idx[a] = tu_depth[u];
            ctu->pred_mode[a] = intra[u] ? INTRA_MODE : INTER_MODE;
            ctu->cbf[Y_COMP][a] = cbf[u]; ctu->cbf[U_COMP][a] = 0; ctu->cbf[V_COMP][a] = 0;
            ctu->qp[a] = qp[u];
            ctu->part_size_type[a] = SIZE_2Nx2N;
            ctu->mv_ref[REF_PIC_LIST_0][a].hor_vector = mv[2 * u]; ctu->mv_ref[REF_PIC_LIST_0][a].ver_vector = mv[2 * u + 1];
            ctu->mv_ref_idx[REF_PIC_LIST_0][a] = intra[u] ? -1 : 0;
        }
    }
    memset(bs_ver, 0, (size_t)units_w * rows * 16); memset(bs_hor, 0, (size_t)units_w * rows * 16);
    hush();
    for (int dir = EDGE_VER; dir <= EDGE_HOR; dir++)
        for (int n = 0; n < cols * rows; n++) {
            ctu_info_t *ctu = &eng->ctu_info[n];
            const int cx = n % cols, cy = n / cols;
            create_partition_ctu_neighbours(et, ctu, ctu->partition_list);
            hmr_deblock_filter_cu(et, slice, ctu, dir);
            for (int r = 0; r < 256; r++) {
                const int u = (cy * 16 + r / 16) * units_w + cx * 16 + r % 16;
                (dir == EDGE_VER ? bs_ver : bs_hor)[u] = et->deblock_filter_strength_bs[dir][eng->raster2abs_table[r]];
            }
        }
    unhush();
    for (int c = 0; c < 3; c++) {
        const int pw = c ? w / 2 : w, ph = c ? h / 2 : h;
        const int16_t *p = (const int16_t *)fr.img.pwnd[c];
        for (int y = 0; y < ph; y++) for (int x = 0; x < pw; x++) out[c][y * pw + x] = (uint8_t)p[y * fr.img.window_size_x[c] + x];
    }
    eng->curr_reference_frame = save_ref;
    wnd_delete(&fr.img);
    return cols * rows;
}
void refdrv_pps_qp_offsets(refdrv *d, int *cb, int *cr) { *cb = d->enc->pps.cb_qp_offset; *cr = d->enc->pps.cr_qp_offset; }

/* ------------------------------------------------------------------------------------------------------------
 * The arithmetic half of the SAO decision through the reference's own sao_derive_offsets (hmr_sao.c:480),
 * sao_invert_quant_offsets (:592) and sao_get_distortion (:620) on one component's statistics of one type.
 * diff / count: 32 entries each (EO: classes 0..4).  offsets: 32 reconstructed offsets; *band: typeAuxInfo.  Returns the distortion.
 * ------------------------------------------------------------------------------------------------------------ */
void sao_init(int bit_depth);
void sao_derive_offsets(henc_thread_t *wpp_thread, int component, int type_idc, sao_stat_data_t *stats, int *quant_offsets, int *type_aux_info);
void sao_invert_quant_offsets(int component, int type_idc, int typeAuxInfo, int *dstOffsets, int *srcOffsets);
int64_t sao_get_distortion(int typeIdc, int typeAuxInfo, int *invQuantOffset, sao_stat_data_t *stats, int bit_depth);
int64_t refdrv_sao_derive(refdrv *d, const int64_t *diff, const int64_t *count, int comp, int type, double lambda, int32_t *offsets, int32_t *band)
{
    henc_thread_t *et = d->et;
    sao_stat_data_t st;
    int q[MAX_NUM_SAO_CLASSES], inv[MAX_NUM_SAO_CLASSES], aux = 0;
    const double save = et->enc_engine->sao_lambdas[comp];
    memcpy(st.diff, diff, sizeof st.diff); memcpy(st.count, count, sizeof st.count);
    sao_init(et->bit_depth);
    et->enc_engine->sao_lambdas[comp] = lambda;
    sao_derive_offsets(et, comp, type, &st, q, &aux);
    sao_invert_quant_offsets(comp, type, aux, inv, q);
    const int64_t dist = sao_get_distortion(type, aux, inv, &st, et->bit_depth);
    et->enc_engine->sao_lambdas[comp] = save;
    for (int k = 0; k < MAX_NUM_SAO_CLASSES; k++) offsets[k] = inv[k];
    *band = aux;
    return dist;
}


/* ------------------------------------------------------------------------------------------------------------
 * The reference's own get_amvp_candidates (hmr_motion_inter.c:2342) on CTU descriptions filled in from per-4x4-unit maps
 * (picture raster over whole CTUs, as refdrv_deblock): inter flag and list-0 vector.  jobs: n x { x, y, size } of 2Nx2N PUs
 * inside the picture; out: n x { mv0.x, mv0.y, mv1.x, mv1.y }.  `d` must have been opened with the picture size.
 * ------------------------------------------------------------------------------------------------------------ */
void get_amvp_candidates(henc_thread_t *et, slice_t *currslice, ctu_info_t *ctu, cu_partition_info_t *curr_cu_info, mv_candiate_list_t *search_candidate_list,
                         int ref_pic_list, int ref_idx, PartSize part_size_type);
void get_merge_mvp_candidates(henc_thread_t *et, slice_t *currslice, ctu_info_t *ctu, cu_partition_info_t *curr_cu_info, PartSize part_size_type, uint8_t *inter_mode_neighbours);
/* merge_max > 0: out receives merge_max x { x, y } per job from get_merge_mvp_candidates (:1937) instead of the two AMVP predictors */
int refdrv_amvp_or_merge(refdrv *d, int w, int h, const uint8_t *inter, const int16_t *mv, const int32_t *jobs, int n_jobs, int merge_max, int32_t *out);
int refdrv_amvp(refdrv *d, int w, int h, const uint8_t *inter, const int16_t *mv, const int32_t *jobs, int n_jobs, int32_t *out)
{
    return refdrv_amvp_or_merge(d, w, h, inter, mv, jobs, n_jobs, 0, out);
}
int refdrv_amvp_or_merge(refdrv *d, int w, int h, const uint8_t *inter, const int16_t *mv, const int32_t *jobs, int n_jobs, int merge_max, int32_t *out)
{
    henc_thread_t *et = d->et;
    hvenc_engine_t *eng = et->enc_engine;
    slice_t *slice = &eng->current_pict.slice;
    static video_frame_t ref_frame;
    const int cols = (w + 63) / 64, rows = (h + 63) / 64, units_w = cols * 16;
    if (eng->pict_total_ctu != cols * rows || et->pict_width[0] != w || et->pict_height[0] != h) return -1;
    memset(&ref_frame, 0, sizeof ref_frame);
    ref_frame.temp_info.poc = 7;
    slice->slice_type = P_SLICE; slice->poc = 8;
    slice->ref_pic_list[REF_PIC_LIST_0][0] = &ref_frame; slice->ref_poc_list[REF_PIC_LIST_0][0] = 7;
    for (int n = 0; n < cols * rows; n++) {
        ctu_info_t *ctu = &eng->ctu_info[n];
        const int cx = n % cols, cy = n / cols;
        ctu->ctu_number = n; ctu->size = 64;
        ctu->x[0] = cx * 64; ctu->y[0] = cy * 64; ctu->x[1] = ctu->x[2] = cx * 32; ctu->y[1] = ctu->y[2] = cy * 32;
        ctu->ctu_left = cx ? &eng->ctu_info[n - 1] : NULL;
        ctu->ctu_top = cy ? &eng->ctu_info[n - cols] : NULL;
        ctu->ctu_top_left = (cx && cy) ? &eng->ctu_info[n - cols - 1] : NULL;
        ctu->ctu_top_right = (cy && cx + 1 < cols) ? &eng->ctu_info[n - cols + 1] : NULL;
        ctu->ctu_left_bottom = NULL;
        for (int r = 0; r < 256; r++) {
            const int a = eng->raster2abs_table[r];
            const int u = (cy * 16 + r / 16) * units_w + cx * 16 + r % 16;
            ctu->pred_mode[a] = inter[u] ? INTER_MODE : INTRA_MODE;
            ctu->mv_ref[REF_PIC_LIST_0][a].hor_vector = mv[2 * u]; ctu->mv_ref[REF_PIC_LIST_0][a].ver_vector = mv[2 * u + 1];
            ctu->mv_ref_idx[REF_PIC_LIST_0][a] = inter[u] ? 0 : -1;
            ctu->inter_mode[a] = inter[u] ? 1 : 0;
            ctu->mv_ref[REF_PIC_LIST_1][a].hor_vector = 0; ctu->mv_ref[REF_PIC_LIST_1][a].ver_vector = 0;
            ctu->mv_ref_idx[REF_PIC_LIST_1][a] = -1;
        }
        create_partition_ctu_neighbours(et, ctu, ctu->partition_list);
    }
    for (int i = 0; i < n_jobs; i++) {
        const int x = jobs[3 * i], y = jobs[3 * i + 1], size = jobs[3 * i + 2];
        ctu_info_t *ctu = &eng->ctu_info[(y / 64) * cols + x / 64];
        int depth = 0;
        for (int s = 64; s > size; s >>= 1) depth++;
        /* the partition of that size at that place: abs_index = z-order of its first unit, position inside its depth = abs_index / units per partition */
        const int ux = (x & 63) / 4, uy = (y & 63) / 4;
        const int abs_index = eng->raster2abs_table[uy * 16 + ux];
        cu_partition_info_t *cu = &ctu->partition_list[et->partition_depth_start[depth]] + abs_index / ((size / 4) * (size / 4));
        if (cu->size != size || cu->x_position != (x & 63) || cu->y_position != (y & 63)) return -2 - i;
        create_partition_ctu_neighbours(et, ctu, ctu->partition_list);       /* the function overwrites flags of the corner units: start clean */
        if (merge_max > 0) {
            uint8_t modes[MERGE_MVP_MAX_NUM_CANDS];
            slice->max_num_merge_candidates = merge_max; slice->num_ref_idx[REF_PIC_LIST_0] = 1;
            get_merge_mvp_candidates(et, slice, ctu, cu, SIZE_2Nx2N, modes);
            const mv_candiate_list_t *l0 = &et->merge_mvp_candidates[REF_PIC_LIST_0];
            if (l0->num_mv_candidates != merge_max) return -200000 - i;
            for (int k = 0; k < merge_max; k++) {
                if (l0->mv_candidates[k].ref_idx != 0) return -300000 - i;
                out[2 * merge_max * i + 2 * k] = l0->mv_candidates[k].mv.hor_vector; out[2 * merge_max * i + 2 * k + 1] = l0->mv_candidates[k].mv.ver_vector;
            }
            continue;
        }
        mv_candiate_list_t list;
        memset(&list, 0, sizeof list);
        get_amvp_candidates(et, slice, ctu, cu, &list, REF_PIC_LIST_0, 0, SIZE_2Nx2N);
        if (list.num_mv_candidates != 2) return -100000 - i;
        for (int k = 0; k < 2; k++) { out[4 * i + 2 * k] = list.mv_candidates[k].mv.hor_vector; out[4 * i + 2 * k + 1] = list.mv_candidates[k].mv.ver_vector; }
    }
    return n_jobs;
}


/* ------------------------------------------------------------------------------------------------------------
 * Boundary strengths of a B picture through the reference's own hmr_deblock_filter_cu (its get_boundary_strength_single, two-list
 * branch :173-229): as refdrv_deblock, with per-unit reference indices and vectors of both lists and the pictures the indices name
 * (pic_l0 / pic_l1: small non-negative picture numbers; equal numbers = the same picture).  Only the strengths are handed back.
 * ------------------------------------------------------------------------------------------------------------ */
int refdrv_deblock_strengths_b(refdrv *d, int w, int h, const uint8_t *cu_depth, const uint8_t *tu_depth, const uint8_t *intra, const uint8_t *cbf,
                               const int8_t *ref0, const int16_t *mv0, const int8_t *ref1, const int16_t *mv1, const int32_t *pic_l0, int n_l0,
                               const int32_t *pic_l1, int n_l1, uint8_t *bs_ver, uint8_t *bs_hor)
{
    henc_thread_t *et = d->et;
    hvenc_engine_t *eng = et->enc_engine;
    slice_t *slice = &eng->current_pict.slice;
    static video_frame_t pics[32];
    video_frame_t fr;
    video_frame_t *save_ref = eng->curr_reference_frame;
    const int cols = (w + 63) / 64, rows = (h + 63) / 64, units_w = cols * 16;
    if (eng->pict_total_ctu != cols * rows || et->pict_width[0] != w || et->pict_height[0] != h) return -1;
    memset(&fr, 0, sizeof fr);
    wnd_alloc(&fr.img, w, h, 80, 80, sizeof(int16_t));
    for (int c = 0; c < 3; c++) {
        const int pw = c ? w / 2 : w, ph = c ? h / 2 : h;
        int16_t *p = (int16_t *)fr.img.pwnd[c];
        for (int y = 0; y < ph; y++) for (int x = 0; x < pw; x++) p[y * fr.img.window_size_x[c] + x] = 128;
    }
    eng->curr_reference_frame = &fr;
    slice->slice_type = B_SLICE; slice->sps = &d->enc->sps; slice->pps = &d->enc->pps;
    slice->deblocking_filter_disabled_flag = 0; slice->slice_beta_offset_div2 = 0; slice->slice_tc_offset_div2 = 0;
    for (int i = 0; i < n_l0; i++) slice->ref_pic_list[REF_PIC_LIST_0][i] = &pics[pic_l0[i] & 31];
    for (int i = 0; i < n_l1; i++) slice->ref_pic_list[REF_PIC_LIST_1][i] = &pics[pic_l1[i] & 31];
    for (int n = 0; n < cols * rows; n++) {
        ctu_info_t *ctu = &eng->ctu_info[n];
        const int cx = n % cols, cy = n / cols;
        ctu->ctu_number = n; ctu->size = 64;
        ctu->x[0] = cx * 64; ctu->y[0] = cy * 64; ctu->x[1] = ctu->x[2] = cx * 32; ctu->y[1] = ctu->y[2] = cy * 32;
        ctu->ctu_left = cx ? &eng->ctu_info[n - 1] : NULL;
        ctu->ctu_top = cy ? &eng->ctu_info[n - cols] : NULL;
        ctu->ctu_top_left = (cx && cy) ? &eng->ctu_info[n - cols - 1] : NULL;
        ctu->ctu_top_right = (cy && cx + 1 < cols) ? &eng->ctu_info[n - cols + 1] : NULL;
        ctu->ctu_left_bottom = NULL;
        for (int r = 0; r < 256; r++) {
            const int a = eng->raster2abs_table[r];
            const int u = (cy * 16 + r / 16) * units_w + cx * 16 + r % 16;
            ctu->pred_depth[a] = cu_depth[u]; ctu->tr_idx[a] = tu_depth[u];
            ctu->pred_mode[a] = intra[u] ? INTRA_MODE : INTER_MODE;
            ctu->cbf[Y_COMP][a] = cbf[u]; ctu->cbf[U_COMP][a] = 0; ctu->cbf[V_COMP][a] = 0;
            ctu->qp[a] = 30;
            ctu->part_size_type[a] = SIZE_2Nx2N;
            ctu->mv_ref[REF_PIC_LIST_0][a].hor_vector = mv0[2 * u]; ctu->mv_ref[REF_PIC_LIST_0][a].ver_vector = mv0[2 * u + 1];
            ctu->mv_ref[REF_PIC_LIST_1][a].hor_vector = mv1[2 * u]; ctu->mv_ref[REF_PIC_LIST_1][a].ver_vector = mv1[2 * u + 1];
            ctu->mv_ref_idx[REF_PIC_LIST_0][a] = intra[u] ? -1 : ref0[u];
            ctu->mv_ref_idx[REF_PIC_LIST_1][a] = intra[u] ? -1 : ref1[u];
        }
    }
    memset(bs_ver, 0, (size_t)units_w * rows * 16); memset(bs_hor, 0, (size_t)units_w * rows * 16);
    hush();
    for (int dir = EDGE_VER; dir <= EDGE_HOR; dir++)
        for (int n = 0; n < cols * rows; n++) {
            ctu_info_t *ctu = &eng->ctu_info[n];
            const int cx = n % cols, cy = n / cols;
            create_partition_ctu_neighbours(et, ctu, ctu->partition_list);
            hmr_deblock_filter_cu(et, slice, ctu, dir);
            for (int r = 0; r < 256; r++) {
                const int u = (cy * 16 + r / 16) * units_w + cx * 16 + r % 16;
                (dir == EDGE_VER ? bs_ver : bs_hor)[u] = et->deblock_filter_strength_bs[dir][eng->raster2abs_table[r]];
            }
        }
    unhush();
    slice->slice_type = P_SLICE;
    eng->curr_reference_frame = save_ref;
    wnd_delete(&fr.img);
    return cols * rows;
}
