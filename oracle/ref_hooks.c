/*
 * ref_hooks.c -- the reference-side binding of include/homer_b200.h section E, compiled: what a maintainer would add to
 * hmr_motion_inter.c so that the encoder's OWN loop (mode decision, AMVP / merge candidates, TU tree, CABAC, deblocking, SAO --
 * all unmodified, on the host) computes its motion searches, motion compensations and inter T/Q chains through the batched GPU
 * API.  TEST INFRASTRUCTURE (needs the reference's private headers, lives under oracle/, built into oracle/_ref/librefdrv.so).
 *
 * Mechanism: the reference is built as a shared library with -fPIC, so its internal calls to
 *     hmr_motion_estimation          (hmr_motion_inter.c:1404, called from hmr_cu_motion_estimation :2625)
 *     hmr_motion_compensation_luma   (:1779, called from predict_inter :3047, check_rd_cost_merge_2nx2n :3655)
 *     hmr_motion_compensation_chroma (:1860, :3048-3049, :3656-3657)
 *     encode_inter_cu                (:40,   called from encode_inter :3165)
 *     encode_inter_cu_chroma         (:133,  :3169-3170)
 *     hmr_rd_init                    (hmr_tables.c:315, called once per picture at hmr_encoder_lib.c:3201: the frame_begin hook)
 * go through the PLT and bind to the definitions below when this library precedes libhomer_ref.so in the lookup order (load
 * librefdrv.so first, RTLD_GLOBAL: tests/_oracle.py does).  Inactive hooks forward to the reference's own functions.
 *
 * Each hook maps the call's arguments one to one onto a job of the batched API and writes the results where the reference's
 * function would have left them (vectors, CTU windows, cu_partition_info_t fields, return value).  Nothing is decided here.
 * Two back ends implement the same five calls:
 *   - the product: hb_enc_* of libhomer_b200.so, resolved with dlsym (GPU);
 *   - an emulation of the same session semantics on the CPU with the oracle's restatement (oracle/hb_oracle.c), which lets
 *     the plumbing of this file be validated in a container without a GPU (tests/test_ref_hooks_cpu.py).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hmr_private.h"
#include "hmr_common.h"

#include "../include/homer_b200.h"
#include "hb_oracle.h"

extern const uint8_t chroma_scale_conversion_table[];

/* ------------------------------------------------------------------ the reference's own functions (forwarding targets) */
typedef uint32_t (*fn_me)(henc_thread_t *, ctu_info_t *, cu_partition_info_t *, int16_t *, int, int16_t *, int, int, int, int, int, int, int, int, int, int, int,
                          motion_vector_t *, motion_vector_t *, mv_candiate_list_t *, uint32_t, unsigned int);
typedef void (*fn_mc_luma)(henc_thread_t *, cu_partition_info_t *, int16_t *, int, int16_t *, int, int, int, int, motion_vector_t *, int);
typedef void (*fn_mc_chroma)(henc_thread_t *, int16_t *, int, int16_t *, int, int, int, motion_vector_t *, int);
typedef int (*fn_tq_luma)(henc_thread_t *, ctu_info_t *, cu_partition_info_t *, int, PartSize, int *, int);
typedef int (*fn_tq_chroma)(henc_thread_t *, ctu_info_t *, cu_partition_info_t *, int, int, PartSize, int *, int);
typedef void (*fn_rd_init)(hvenc_engine_t *, slice_t *);

static struct { fn_me me; fn_mc_luma mc_luma; fn_mc_chroma mc_chroma; fn_tq_luma tq_luma; fn_tq_chroma tq_chroma; fn_rd_init rd_init; } R;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void resolve_reference(void)
{
    /* the already loaded reference library by its soname (rpath $ORIGIN); RTLD_NEXT as a second try */
    void *ref = dlopen("libhomer_ref.so", RTLD_LAZY | RTLD_NOLOAD);
#define RESOLVE(field, name) do { R.field = ref ? dlsym(ref, name) : NULL; if (!R.field) R.field = dlsym(RTLD_NEXT, name); \
        if (!R.field) { fprintf(stderr, "ref_hooks: cannot find the reference's %s\n", name); abort(); } } while (0)
    RESOLVE(me, "hmr_motion_estimation");
    RESOLVE(mc_luma, "hmr_motion_compensation_luma");
    RESOLVE(mc_chroma, "hmr_motion_compensation_chroma");
    RESOLVE(tq_luma, "encode_inter_cu");
    RESOLVE(tq_chroma, "encode_inter_cu_chroma");
    RESOLVE(rd_init, "hmr_rd_init");
#undef RESOLVE
}

/* ------------------------------------------------------------------ back ends: the five calls of section E */
typedef struct enc_backend {
    void *session;
    int (*upload_i16)(void *session, int which, const int16_t *y, int ys, const int16_t *u, int us, const int16_t *v, int vs);
    int (*me)(void *session, const hb_me_job *jobs, int n, double avg_dist, int action, hb_me_result *results);
    int (*predict)(void *session, const hb_mc_job *jobs, int n, int16_t *blocks);
    int (*tq)(void *session, const hb_tu_job *jobs, int n, const hb_tq_params *params, int16_t *coeffs, int16_t *decoded, hb_tu_result *results);
    void (*destroy)(void *session);
} enc_backend;

/* --- the product (GPU) */
static struct {
    void *lib;
    int (*enc_create)(hb_ctx *, int, int, hb_enc **);
    void (*enc_destroy)(hb_enc *);
    hb_frame *(*enc_frame)(hb_enc *, int);
    int (*upload_i16)(hb_ctx *, hb_frame *, const int16_t *, int, const int16_t *, int, const int16_t *, int);
    hb_ctx *(*default_ctx)(void);
    const char *(*last_error)(void);
    int (*enc_me)(hb_enc *, const hb_me_job *, int, double, int, hb_me_result *);
    int (*enc_predict)(hb_enc *, const hb_mc_job *, int, int16_t *);
    int (*enc_tq)(hb_enc *, const hb_tu_job *, int, const hb_tq_params *, int16_t *, int16_t *, hb_tu_result *);
} G;

static int gpu_upload(void *s, int which, const int16_t *y, int ys, const int16_t *u, int us, const int16_t *v, int vs)
{
    return G.upload_i16(G.default_ctx(), G.enc_frame((hb_enc *)s, which), y, ys, u, us, v, vs);
}
static int gpu_me(void *s, const hb_me_job *j, int n, double avg, int action, hb_me_result *r) { return G.enc_me((hb_enc *)s, j, n, avg, action, r); }
static int gpu_predict(void *s, const hb_mc_job *j, int n, int16_t *b) { return G.enc_predict((hb_enc *)s, j, n, b); }
static int gpu_tq(void *s, const hb_tu_job *j, int n, const hb_tq_params *p, int16_t *c, int16_t *d, hb_tu_result *r) { return G.enc_tq((hb_enc *)s, j, n, p, c, d, r); }
static void gpu_destroy(void *s) { G.enc_destroy((hb_enc *)s); }

static int gpu_backend(void *lib, int w, int h, enc_backend *be)
{
    hb_enc *e = NULL;
    G.lib = lib;
#define SYM(field, name) do { G.field = dlsym(lib, name); if (!G.field) { fprintf(stderr, "ref_hooks: libhomer_b200 lacks %s\n", name); return -1; } } while (0)
    SYM(enc_create, "hb_enc_create"); SYM(enc_destroy, "hb_enc_destroy"); SYM(enc_frame, "hb_enc_frame"); SYM(upload_i16, "hb_frame_upload_i16");
    SYM(default_ctx, "hb_default_ctx"); SYM(last_error, "hb_last_error"); SYM(enc_me, "hb_enc_me"); SYM(enc_predict, "hb_enc_predict"); SYM(enc_tq, "hb_enc_tq");
#undef SYM
    if (G.enc_create(G.default_ctx(), w, h, &e) != HB_OK) { fprintf(stderr, "ref_hooks: hb_enc_create: %s\n", G.last_error()); return -1; }
    be->session = e; be->upload_i16 = gpu_upload; be->me = gpu_me; be->predict = gpu_predict; be->tq = gpu_tq; be->destroy = gpu_destroy;
    return 0;
}

/* --- the same session semantics on the CPU with the oracle's restatement: four int16 pictures with a replicated border */
#define EMU_PAD 96
typedef struct emu_session { int w, h; int16_t *alloc[4][3], *org[4][3]; int stride[3]; orc_tables *tab; } emu_session;

static void emu_destroy(void *sv)
{
    emu_session *s = (emu_session *)sv;
    if (!s) return;
    for (int f = 0; f < 4; f++) for (int c = 0; c < 3; c++) free(s->alloc[f][c]);
    if (s->tab) orc_tables_destroy(s->tab);
    free(s);
}
static int emu_upload(void *sv, int which, const int16_t *y, int ys, const int16_t *u, int us, const int16_t *v, int vs)
{
    emu_session *s = (emu_session *)sv;
    const int16_t *src[3] = { y, u, v };
    const int st[3] = { ys, us, vs };
    for (int c = 0; c < 3; c++) {
        const int pw = c ? s->w / 2 : s->w, ph = c ? s->h / 2 : s->h, pad = c ? EMU_PAD / 2 : EMU_PAD;
        for (int yy = -pad; yy < ph + pad; yy++) {
            const int sy = yy < 0 ? 0 : (yy >= ph ? ph - 1 : yy);
            int16_t *row = s->org[which][c] + (ptrdiff_t)yy * s->stride[c];
            for (int xx = -pad; xx < pw + pad; xx++) row[xx] = src[c][(size_t)sy * st[c] + (xx < 0 ? 0 : (xx >= pw ? pw - 1 : xx))];
        }
    }
    return 0;
}
static int emu_me(void *sv, const hb_me_job *jobs, int n, double avg_dist, int action, hb_me_result *res)
{
    emu_session *s = (emu_session *)sv;
    for (int i = 0; i < n; i++) {
        const hb_me_job *j = &jobs[i];
        orc_me_in in; orc_me_out out;
        memset(&in, 0, sizeof in);
        in.orig = s->org[HB_ENC_CUR][0] + (ptrdiff_t)j->y * s->stride[0] + j->x; in.orig_stride = s->stride[0];
        in.ref = s->org[HB_ENC_REF][0] + (ptrdiff_t)j->y * s->stride[0] + j->x; in.ref_stride = s->stride[0];
        in.gx = j->x; in.gy = j->y; in.size = j->size; in.frame_w = s->w; in.frame_h = s->h; in.range_x = 128; in.range_y = 64;
        in.n_amvp = j->n_amvp;
        for (int k = 0; k < 2; k++) { in.amvp[k].x = j->amvp[k].x; in.amvp[k].y = j->amvp[k].y; }
        in.n_start = j->n_start;
        for (int k = 0; k < 3; k++) { in.start[k].x = j->start[k].x; in.start[k].y = j->start[k].y; }
        in.qp = j->qp; in.avg_dist = avg_dist; in.action = action;
        orc_motion_estimation(&in, &out);
        res[i].mv.x = out.mv.x; res[i].mv.y = out.mv.y; res[i].subpix.x = out.subpix.x; res[i].subpix.y = out.subpix.y;
        res[i].sad = out.sad; res[i].n_probes = out.n_int_sads;
    }
    return 0;
}
static int emu_predict(void *sv, const hb_mc_job *jobs, int n, int16_t *blocks)
{
    emu_session *s = (emu_session *)sv;
    for (int i = 0; i < n; i++) {
        const hb_mc_job *j = &jobs[i];
        const orc_mv mv = { j->mv.x, j->mv.y };
        for (int c = 0; c < 3; c++) {
            const int x = c ? j->x / 2 : j->x, y = c ? j->y / 2 : j->y, sz = c ? j->size / 2 : j->size;
            int16_t *dst = s->org[HB_ENC_PRED][c] + (ptrdiff_t)y * s->stride[c] + x;
            const int16_t *ref = s->org[HB_ENC_REF][c] + (ptrdiff_t)y * s->stride[c] + x;
            if (c == 0) orc_mc_luma(ref, s->stride[c], dst, s->stride[c], sz, mv);
            else orc_mc_chroma(ref, s->stride[c], dst, s->stride[c], sz, mv);
            for (int r = 0; r < sz; r++) memcpy(blocks + (size_t)r * sz, dst + (ptrdiff_t)r * s->stride[c], sizeof(int16_t) * (size_t)sz);
            blocks += (size_t)sz * sz;
        }
    }
    return 0;
}
static int emu_tq(void *sv, const hb_tu_job *jobs, int n, const hb_tq_params *p, int16_t *coeffs, int16_t *decoded, hb_tu_result *res)
{
    emu_session *s = (emu_session *)sv;
    for (int i = 0; i < n; i++) {
        const hb_tu_job *j = &jobs[i];
        const int c = j->comp, st = s->stride[c];
        const ptrdiff_t at = (ptrdiff_t)j->y * st + j->x;
        orc_tu_out o;
        orc_encode_inter_tu(s->tab, s->org[HB_ENC_CUR][c] + at, st, s->org[HB_ENC_PRED][c] + at, st, coeffs, s->org[HB_ENC_RECON][c] + at, st,
                            j->size, c, j->qp, p->is_islice, p->sign_hiding, p->avg_dist, c ? p->chroma_weight : 1.0, &o);
        for (int r = 0; r < j->size; r++) memcpy(decoded + (size_t)r * j->size, s->org[HB_ENC_RECON][c] + at + (ptrdiff_t)r * st, sizeof(int16_t) * (size_t)j->size);
        res[i].sum = o.sum; res[i].ssd = o.ssd; res[i].ssd_zero = o.ssd_zero; res[i].zeroed = o.zeroed;
        coeffs += (size_t)j->size * j->size; decoded += (size_t)j->size * j->size;
    }
    return 0;
}
static int emu_backend(int w, int h, enc_backend *be)
{
    emu_session *s = (emu_session *)calloc(1, sizeof *s);
    s->w = w; s->h = h; s->tab = orc_tables_create();
    for (int c = 0; c < 3; c++) {
        const int pw = c ? w / 2 : w, ph = c ? h / 2 : h, pad = c ? EMU_PAD / 2 : EMU_PAD;
        s->stride[c] = pw + 2 * pad;
        for (int f = 0; f < 4; f++) {
            s->alloc[f][c] = (int16_t *)calloc((size_t)s->stride[c] * (ph + 2 * pad), sizeof(int16_t));
            s->org[f][c] = s->alloc[f][c] + (size_t)pad * s->stride[c] + pad;
        }
    }
    be->session = s; be->upload_i16 = emu_upload; be->me = emu_me; be->predict = emu_predict; be->tq = emu_tq; be->destroy = emu_destroy;
    return 0;
}

/* ------------------------------------------------------------------ hook state: one hooked encoder (one engine) at a time */
enum { C_FRAMES, C_P_FRAMES, C_ME, C_ME_FWD, C_MC, C_MC_CACHED, C_MC_FWD, C_TQ, C_TQ_CACHED, C_TQ_FWD, C_TQ_STALE, C_ERRORS, C_N };
static struct {
    volatile int active;
    enc_backend be;
    hvenc_engine_t *eng;
    int w, h;
    int is_p;                                 /* this picture's calls are served: P slice, one reference picture, uploads done */
    int16_t *ref_org[3]; int ref_stride[3];   /* the planes of ref_pic_list[0][0] of this picture */
    uint8_t *mirror[3];                       /* host copy of what the session's prediction picture holds (tight planes) */
    int batch_tq;                             /* luma call also computes the unit's two chroma jobs (one round trip per TU instead of three) */
    long cnt[C_N];
    pthread_mutex_t cnt_lock;
} H = { .cnt_lock = PTHREAD_MUTEX_INITIALIZER };

static void count(int i) { __atomic_add_fetch(&H.cnt[i], 1, __ATOMIC_RELAXED); }
static int hook_on(const henc_thread_t *et) { return H.active && et->enc_engine == H.eng; }

/* per calling thread: the chroma blocks of the last motion compensation and the chroma results of the last luma T/Q call */
static __thread struct {
    int valid; hb_mc_job job; long seq;
    int16_t blocks[64 * 64 * 3 / 2];
} t_mc;
static __thread struct {
    int valid[3]; cu_partition_info_t *cu; ctu_info_t *ctu; long mc_seq;
    hb_tu_job job[3]; hb_tq_params prm; hb_tu_result res[3];
    int16_t coeff[3][32 * 32], dec[3][32 * 32];
} t_tq;
static __thread long t_mc_seq;

/* position of a pointer inside plane c of this picture's reference, or 0 */
static int locate(const int16_t *p, int c, int *x, int *y)
{
    const int pw = c ? H.w / 2 : H.w, ph = c ? H.h / 2 : H.h;
    const ptrdiff_t off = p - H.ref_org[c];
    if (!H.ref_org[c] || off < 0) return 0;
    *y = (int)(off / H.ref_stride[c]); *x = (int)(off % H.ref_stride[c]);
    return *x < pw && *y < ph;
}

/* ------------------------------------------------------------------ frame_begin: hmr_rd_init, hmr_encoder_lib.c:3201 */
void hmr_rd_init(hvenc_engine_t *enc_engine, slice_t *currslice)
{
    pthread_once(&g_once, resolve_reference);
    R.rd_init(enc_engine, currslice);
    if (!H.active || enc_engine != H.eng) return;
    H.is_p = 0;
    count(C_FRAMES);
    if (currslice->slice_type != P_SLICE || currslice->num_ref_idx[REF_PIC_LIST_0] != 1 || !currslice->ref_pic_list[REF_PIC_LIST_0][0]) return;
    wnd_t *cur = &enc_engine->current_pict.img2encode->img, *ref = &currslice->ref_pic_list[REF_PIC_LIST_0][0]->img;
    if (H.be.upload_i16(H.be.session, HB_ENC_CUR, WND_DATA_PTR(int16_t *, *cur, Y_COMP), WND_STRIDE_2D(*cur, Y_COMP), WND_DATA_PTR(int16_t *, *cur, U_COMP),
                        WND_STRIDE_2D(*cur, U_COMP), WND_DATA_PTR(int16_t *, *cur, V_COMP), WND_STRIDE_2D(*cur, V_COMP)) ||
        H.be.upload_i16(H.be.session, HB_ENC_REF, WND_DATA_PTR(int16_t *, *ref, Y_COMP), WND_STRIDE_2D(*ref, Y_COMP), WND_DATA_PTR(int16_t *, *ref, U_COMP),
                        WND_STRIDE_2D(*ref, U_COMP), WND_DATA_PTR(int16_t *, *ref, V_COMP), WND_STRIDE_2D(*ref, V_COMP))) { count(C_ERRORS); return; }
    for (int c = 0; c < 3; c++) { H.ref_org[c] = WND_DATA_PTR(int16_t *, *ref, c); H.ref_stride[c] = WND_STRIDE_2D(*ref, c); }
    /* the session's prediction picture starts out unknown: nothing may match the mirror until a motion compensation wrote it */
    for (int c = 0; c < 3; c++) memset(H.mirror[c], 0, (size_t)(c ? H.w / 2 : H.w) * (c ? H.h / 2 : H.h));
    H.is_p = 1;
    count(C_P_FRAMES);
}

/* ------------------------------------------------------------------ hmr_motion_estimation, call site hmr_motion_inter.c:2625 */
uint32_t hmr_motion_estimation(henc_thread_t *et, ctu_info_t *ctu, cu_partition_info_t *curr_cu_info, int16_t *orig_buff, int orig_buff_stride,
                               int16_t *reference_buff, int reference_buff_stride, int curr_part_global_x, int curr_part_global_y, int init_x, int init_y,
                               int curr_part_size, int curr_part_size_shift, int search_range_x, int search_range_y, int frame_size_x, int frame_size_y,
                               motion_vector_t *mv, motion_vector_t *subpix_mv, mv_candiate_list_t *amvp_candidate_list, uint32_t threshold, unsigned int action)
{
    pthread_once(&g_once, resolve_reference);
    int x, y;
    const mv_candiate_list_t *starts = &et->mv_search_candidates;
    if (hook_on(et) && H.is_p && locate(reference_buff, 0, &x, &y) && x == curr_part_global_x && y == curr_part_global_y && init_x == 0 && init_y == 0 &&
        search_range_x == MOTION_SEARCH_RANGE_X && search_range_y == MOTION_SEARCH_RANGE_Y && frame_size_x == H.w && frame_size_y == H.h &&
        (action & MOTION_PEL_MASK) && (curr_part_size == 8 || curr_part_size == 16 || curr_part_size == 32 || curr_part_size == 64) &&
        curr_part_size == curr_cu_info->size && x + curr_part_size <= H.w && y + curr_part_size <= H.h &&
        amvp_candidate_list->num_mv_candidates >= 0 && amvp_candidate_list->num_mv_candidates <= 2 && starts->num_mv_candidates >= 0 && starts->num_mv_candidates <= 3) {
        hb_me_job j;
        hb_me_result r;
        memset(&j, 0, sizeof j);
        j.x = x; j.y = y; j.size = curr_part_size; j.qp = (int32_t)curr_cu_info->qp; j.parent = -1;
        j.n_amvp = amvp_candidate_list->num_mv_candidates;
        for (int i = 0; i < j.n_amvp; i++) { j.amvp[i].x = amvp_candidate_list->mv_candidates[i].mv.hor_vector; j.amvp[i].y = amvp_candidate_list->mv_candidates[i].mv.ver_vector; }
        j.n_start = starts->num_mv_candidates;
        for (int i = 0; i < j.n_start; i++) { j.start[i].x = starts->mv_candidates[i].mv.hor_vector; j.start[i].y = starts->mv_candidates[i].mv.ver_vector; }
        if (H.be.me(H.be.session, &j, 1, et->enc_engine->avg_dist, (int)(action & 7), &r) == 0) {
            mv->hor_vector = r.mv.x; mv->ver_vector = r.mv.y;
            subpix_mv->hor_vector = r.subpix.x; subpix_mv->ver_vector = r.subpix.y;
            count(C_ME);
            return r.sad;
        }
        count(C_ERRORS);
    }
    if (hook_on(et)) count(C_ME_FWD);
    return R.me(et, ctu, curr_cu_info, orig_buff, orig_buff_stride, reference_buff, reference_buff_stride, curr_part_global_x, curr_part_global_y, init_x, init_y,
                curr_part_size, curr_part_size_shift, search_range_x, search_range_y, frame_size_x, frame_size_y, mv, subpix_mv, amvp_candidate_list, threshold, action);
}

/* ------------------------------------------------------------------ motion compensation, call sites :3047-3049, :3655-3657 */
static void store_block(int16_t *dst, int dst_stride, const int16_t *src, int n)
{
    for (int r = 0; r < n; r++) memcpy(dst + (ptrdiff_t)r * dst_stride, src + (size_t)r * n, sizeof(int16_t) * (size_t)n);
}
static void mirror_block(int c, int x, int y, const int16_t *src, int n)
{
    const int pw = c ? H.w / 2 : H.w;
    for (int r = 0; r < n; r++) { uint8_t *m = H.mirror[c] + (size_t)(y + r) * pw + x; for (int q = 0; q < n; q++) m[q] = (uint8_t)src[r * n + q]; }
}

void hmr_motion_compensation_luma(henc_thread_t *et, cu_partition_info_t *curr_cu_info, int16_t *reference_buff, int reference_buff_stride, int16_t *pred_buff,
                                  int pred_buff_stride, int width, int height, int curr_part_size_shift, motion_vector_t *mv, int is_bi_predict)
{
    pthread_once(&g_once, resolve_reference);
    int x, y;
    if (hook_on(et) && H.is_p && !is_bi_predict && width == height && (width == 8 || width == 16 || width == 32 || width == 64) &&
        locate(reference_buff, 0, &x, &y) && !((x | y) & 7) && x + width <= H.w && y + width <= H.h) {
        hb_mc_job j = { x, y, width, { mv->hor_vector, mv->ver_vector } };
        t_mc.valid = 0;
        t_mc_seq++;
        if (H.be.predict(H.be.session, &j, 1, t_mc.blocks) == 0) {
            const int n = width, nc = width / 2;
            store_block(pred_buff, pred_buff_stride, t_mc.blocks, n);
            mirror_block(0, x, y, t_mc.blocks, n);
            mirror_block(1, x / 2, y / 2, t_mc.blocks + n * n, nc);
            mirror_block(2, x / 2, y / 2, t_mc.blocks + n * n + nc * nc, nc);
            t_mc.valid = 1; t_mc.job = j; t_mc.seq = t_mc_seq;
            count(C_MC);
            return;
        }
        count(C_ERRORS);       /* e.g. a vector further outside the picture than the resident border reaches: the host computes it */
    }
    if (hook_on(et)) count(C_MC_FWD);
    R.mc_luma(et, curr_cu_info, reference_buff, reference_buff_stride, pred_buff, pred_buff_stride, width, height, curr_part_size_shift, mv, is_bi_predict);
}

void hmr_motion_compensation_chroma(henc_thread_t *et, int16_t *reference_buff, int reference_buff_stride, int16_t *pred_buff, int pred_buff_stride,
                                    int curr_part_size, int curr_part_size_shift, motion_vector_t *mv, int is_bi_predict)
{
    pthread_once(&g_once, resolve_reference);
    int x, y, c;
    if (hook_on(et) && H.is_p && !is_bi_predict && t_mc.valid) {
        /* the luma call of this unit already computed both chroma blocks (same vector, same place): hand them out */
        for (c = 1; c < 3; c++)
            if (locate(reference_buff, c, &x, &y) && 2 * x == t_mc.job.x && 2 * y == t_mc.job.y && 2 * curr_part_size == t_mc.job.size &&
                mv->hor_vector == t_mc.job.mv.x && mv->ver_vector == t_mc.job.mv.y) {
                const int n = t_mc.job.size, nc = n / 2;
                store_block(pred_buff, pred_buff_stride, t_mc.blocks + n * n + (c == 2 ? nc * nc : 0), nc);
                count(C_MC_CACHED);
                return;
            }
    }
    if (hook_on(et)) count(C_MC_FWD);
    R.mc_chroma(et, reference_buff, reference_buff_stride, pred_buff, pred_buff_stride, curr_part_size, curr_part_size_shift, mv, is_bi_predict);
}

/* ------------------------------------------------------------------ the inter T/Q chain, call sites :3165-3170 */
/* is the session's prediction picture (as mirrored) what the host holds in et->prediction_wnd[0] for this block? */
static int pred_in_sync(const henc_thread_t *et, int c, int px, int py, int gx, int gy, int n)
{
    const int pw = c ? H.w / 2 : H.w;
    const int16_t *p = WND_POSITION_2D(int16_t *, et->prediction_wnd[0], c, px, py, 0, et->ctu_width);
    const int ps = WND_STRIDE_2D(et->prediction_wnd[0], c);
    for (int r = 0; r < n; r++) {
        const uint8_t *m = H.mirror[c] + (size_t)(gy + r) * pw + gx;
        for (int q = 0; q < n; q++) if (p[r * ps + q] != m[q]) return 0;
    }
    return 1;
}

static void tq_params(const henc_thread_t *et, hb_tq_params *p)
{
    const slice_t *sl = &et->enc_engine->current_pict.slice;
    const int off = et->enc_engine->chroma_qp_offset;
    p->is_islice = sl->slice_type == I_SLICE;
    p->sign_hiding = (int32_t)et->pps->sign_data_hiding_flag;
    p->avg_dist = et->enc_engine->avg_dist;
    p->chroma_weight = pow(2.0, (sl->qp - chroma_scale_conversion_table[clip(sl->qp + off, 0, 57)]) / 3.0);      /* :155 */
}

/* geometry of the unit a T/Q call works on, exactly as the reference derives it (:40-73 luma, :133-176 chroma) */
typedef struct tq_geom { int comp, px, py, gx, gy, n, qp, quant_off, wnd; } tq_geom;
static int tq_geometry(const henc_thread_t *et, const ctu_info_t *ctu, const cu_partition_info_t *cu, int comp, PartSize part, tq_geom *g)
{
    const cu_partition_info_t *proc = (comp == Y_COMP || cu->size_chroma != 2) ? cu : cu->parent;
    g->comp = comp;
    g->wnd = cu->depth + 1 + (part != SIZE_2Nx2N);
    if (comp == Y_COMP) {
        g->px = cu->x_position; g->py = cu->y_position; g->n = cu->size; g->qp = (int)cu->qp;
        g->quant_off = cu->abs_index << et->num_partitions_in_cu_shift;
    } else {
        g->px = proc->x_position_chroma; g->py = proc->y_position_chroma; g->n = proc->size_chroma;
        g->qp = chroma_scale_conversion_table[clip((int)cu->qp + et->enc_engine->chroma_qp_offset, 0, 57)];
        g->quant_off = (proc->abs_index << et->num_partitions_in_cu_shift) >> 2;
    }
    g->gx = ctu->x[comp] + g->px; g->gy = ctu->y[comp] + g->py;
    const int pw = comp ? H.w / 2 : H.w, ph = comp ? H.h / 2 : H.h;
    if (g->n != 4 && g->n != 8 && g->n != 16 && !(g->n == 32 && comp == Y_COMP)) return 0;
    if (g->qp < 0 || g->qp > 51 || g->gx < 0 || g->gy < 0 || g->gx + g->n > pw || g->gy + g->n > ph || (g->gx & ((g->n < 16 ? g->n : 16) - 1))) return 0;
    return g->wnd < NUM_QUANT_WNDS && g->wnd < NUM_DECODED_WNDS;
}

/* leave a unit's results where encode_inter_cu / _chroma leave them: levels (1-D) and decoded samples (2-D) in the CTU windows */
static void tq_store(henc_thread_t *et, const tq_geom *g, const int16_t *coeff, const int16_t *dec)
{
    int16_t *quant_buff = WND_POSITION_1D(int16_t *, *et->transform_quant_wnd[g->wnd], g->comp, 0, et->ctu_width, g->quant_off);
    int16_t *decoded_buff = WND_POSITION_2D(int16_t *, *et->decoded_mbs_wnd[g->wnd], g->comp, g->px, g->py, 0, et->ctu_width);
    memcpy(quant_buff, coeff, sizeof(int16_t) * (size_t)g->n * g->n);
    store_block(decoded_buff, WND_STRIDE_2D(*et->decoded_mbs_wnd[g->wnd], g->comp), dec, g->n);
}

static int same_params(const hb_tq_params *a, const hb_tq_params *b)
{
    return a->is_islice == b->is_islice && a->sign_hiding == b->sign_hiding && a->avg_dist == b->avg_dist && a->chroma_weight == b->chroma_weight;
}

/* result of component `comp` of this unit: from the luma call's look-ahead when it is still valid, else one job now */
static int tq_run(henc_thread_t *et, ctu_info_t *ctu, cu_partition_info_t *cu, PartSize part, int comp, const tq_geom *g, hb_tu_result *res,
                  const int16_t **coeff, const int16_t **dec)
{
    hb_tq_params prm;
    tq_params(et, &prm);
    const hb_tu_job job = { comp, g->gx, g->gy, g->n, g->qp };
    if (comp != Y_COMP && t_tq.valid[comp] && t_tq.cu == cu && t_tq.ctu == ctu && t_tq.mc_seq == t_mc_seq && !memcmp(&t_tq.job[comp], &job, sizeof job) &&
        same_params(&t_tq.prm, &prm)) {
        t_tq.valid[comp] = 0;
        *res = t_tq.res[comp]; *coeff = t_tq.coeff[comp]; *dec = t_tq.dec[comp];
        count(C_TQ_CACHED);
        return 1;
    }
    t_tq.valid[0] = t_tq.valid[1] = t_tq.valid[2] = 0;
    hb_tu_job jobs[3];
    tq_geom gc[3];
    int n = 1;
    jobs[0] = job;
    if (comp == Y_COMP && H.batch_tq && cu->size_chroma != 2) {
        /* the two chroma calls of this unit follow at once (:3169-3170) and read inputs that exist already: same round trip */
        for (int c = 1; c < 3; c++) {
            if (!tq_geometry(et, ctu, cu, c, part, &gc[c]) || !pred_in_sync(et, c, gc[c].px, gc[c].py, gc[c].gx, gc[c].gy, gc[c].n)) { n = 1; break; }
            const hb_tu_job jc = { c, gc[c].gx, gc[c].gy, gc[c].n, gc[c].qp };
            jobs[n++] = jc;
        }
    }
    static __thread int16_t co[3 * 32 * 32], de[3 * 32 * 32];
    hb_tu_result rs[3];
    if (H.be.tq(H.be.session, jobs, n, &prm, co, de, rs)) { count(C_ERRORS); return 0; }
    size_t off = 0;
    for (int k = 0; k < n; k++) {
        const int c = jobs[k].comp;
        const size_t nn = (size_t)jobs[k].size * jobs[k].size;
        memcpy(t_tq.coeff[c], co + off, sizeof(int16_t) * nn); memcpy(t_tq.dec[c], de + off, sizeof(int16_t) * nn);
        t_tq.res[c] = rs[k]; t_tq.job[c] = jobs[k];
        if (k > 0) t_tq.valid[c] = 1;
        off += nn;
    }
    t_tq.cu = cu; t_tq.ctu = ctu; t_tq.mc_seq = t_mc_seq; t_tq.prm = prm;
    *res = t_tq.res[comp]; *coeff = t_tq.coeff[comp]; *dec = t_tq.dec[comp];
    count(C_TQ);
    return 1;
}

int encode_inter_cu(henc_thread_t *et, ctu_info_t *ctu, cu_partition_info_t *curr_cu_info, int depth, PartSize part_size_type, int *curr_sum, int gcnt)
{
    pthread_once(&g_once, resolve_reference);
    tq_geom g;
    if (hook_on(et) && H.is_p && gcnt == 0 && tq_geometry(et, ctu, curr_cu_info, Y_COMP, part_size_type, &g)) {
        if (!pred_in_sync(et, Y_COMP, g.px, g.py, g.gx, g.gy, g.n)) count(C_TQ_STALE);
        else {
            hb_tu_result r; const int16_t *coeff, *dec;
            if (tq_run(et, ctu, curr_cu_info, part_size_type, Y_COMP, &g, &r, &coeff, &dec)) {
                tq_store(et, &g, coeff, dec);
                curr_cu_info->inter_cbf[Y_COMP] = (r.sum ? 1 : 0) << (curr_cu_info->depth - depth);       /* :86, cleared again by the zero-out :110 */
                curr_cu_info->inter_tr_idx = curr_cu_info->depth - depth;                                 /* :87 */
                *curr_sum = r.sum;
                curr_cu_info->sum = (uint32_t)r.sum;                                                       /* :124 */
                return (int)r.ssd;
            }
        }
    }
    if (hook_on(et)) count(C_TQ_FWD);
    return R.tq_luma(et, ctu, curr_cu_info, depth, part_size_type, curr_sum, gcnt);
}

int encode_inter_cu_chroma(henc_thread_t *et, ctu_info_t *ctu, cu_partition_info_t *curr_cu_info, int component, int depth, PartSize part_size_type,
                           int *curr_sum, int gcnt)
{
    pthread_once(&g_once, resolve_reference);
    tq_geom g;
    if (hook_on(et) && H.is_p && gcnt == 0 && (component == U_COMP || component == V_COMP) && tq_geometry(et, ctu, curr_cu_info, component, part_size_type, &g)) {
        if (!pred_in_sync(et, component, g.px, g.py, g.gx, g.gy, g.n)) count(C_TQ_STALE);
        else {
            hb_tu_result r; const int16_t *coeff, *dec;
            if (tq_run(et, ctu, curr_cu_info, part_size_type, component, &g, &r, &coeff, &dec)) {
                tq_store(et, &g, coeff, dec);
                curr_cu_info->inter_cbf[component] = (r.sum ? 1 : 0) << (curr_cu_info->depth - depth);    /* :190, :214 */
                *curr_sum = r.sum;
                curr_cu_info->sum += (uint32_t)r.sum;                                                      /* :228 */
                return (int)r.ssd;
            }
        }
    }
    if (hook_on(et)) count(C_TQ_FWD);
    return R.tq_chroma(et, ctu, curr_cu_info, component, depth, part_size_type, curr_sum, gcnt);
}

/* ------------------------------------------------------------------ installation (a refdrv_table_hook of refdrv_encode_lockstep) */
void refdrv_install_gpu_table(void *funcs_table, void *user);      /* ref_driver.c: the per-call table of INTEGRATION.md section 1 */

/* user: { lib (dlopen handle of libhomer_b200.so, or NULL for the CPU emulation), which (per-call table mask for everything the CU
 * hooks do not cover, 0 = keep the reference's own table), batch_tq } */
typedef struct cu_hook_cfg { void *lib; int which; int batch_tq; } cu_hook_cfg;

void refdrv_install_cu_hooks(void *funcs_table, void *user)
{
    const cu_hook_cfg *cfg = (const cu_hook_cfg *)user;
    hvenc_enc_t *enc = (hvenc_enc_t *)((char *)funcs_table - offsetof(hvenc_enc_t, funcs));
    pthread_once(&g_once, resolve_reference);
    if (H.active) { fprintf(stderr, "ref_hooks: hooks are already installed on another encoder\n"); abort(); }
    if (enc->num_encoder_engines != 1) { fprintf(stderr, "ref_hooks: the CU hooks need num_enc_engines = 1 (the reference picture must be complete at frame begin)\n"); abort(); }
    memset(H.cnt, 0, sizeof H.cnt);
    H.eng = enc->encoder_engines[0];
    H.w = enc->pict_width[Y_COMP]; H.h = enc->pict_height[Y_COMP];
    H.is_p = 0; H.batch_tq = cfg->batch_tq;
    if ((cfg->lib ? gpu_backend(cfg->lib, H.w, H.h, &H.be) : emu_backend(H.w, H.h, &H.be)) != 0) abort();
    for (int c = 0; c < 3; c++) H.mirror[c] = (uint8_t *)calloc((size_t)(c ? H.w / 2 : H.w) * (c ? H.h / 2 : H.h), 1);
    if (cfg->lib && cfg->which) {
        struct { void *lib; int which; } u = { cfg->lib, cfg->which };
        refdrv_install_gpu_table(funcs_table, &u);
    }
    H.active = 1;
}
void *refdrv_install_cu_hooks_addr(void) { return (void *)refdrv_install_cu_hooks; }

/* switch the hooks off, release the session, report the counters (C_N longs) */
int refdrv_cu_hooks_off(long *counters)
{
    if (counters) memcpy(counters, H.cnt, sizeof H.cnt);
    memset(H.cnt, 0, sizeof H.cnt);
    if (!H.active) return 0;
    H.active = 0;
    H.be.destroy(H.be.session);
    memset(&H.be, 0, sizeof H.be);
    for (int c = 0; c < 3; c++) { free(H.mirror[c]); H.mirror[c] = NULL; H.ref_org[c] = NULL; }
    H.eng = NULL;
    return C_N;
}


/* ------------------------------------------------------------------ capture of a picture's decisions, CTU by CTU
 * hmr_deblock_sao_pad_sync_ctu (hmr_encoder_lib.c:2386) is called right after a CTU's reconstruction went into the picture
 * (mem_transfer_decoded_blocks :2942) and its levels into ctu->coeff_wnd (:2945), before any in-loop filter touches them: the hook
 * copies the CTU's part of the UNFILTERED reconstruction, its levels and the per-4x4-unit decision arrays, then forwards.  This is what
 * tests/test_intra_recon*.py rebuild an intra picture from. */
typedef void (*fn_sync_ctu)(henc_thread_t *, slice_t *, ctu_info_t *);
typedef struct refdrv_capture {
    int32_t width, height, frame, n_seen;      /* capture the picture with num_encoded_frames == frame; n_seen counts its CTUs */
    uint8_t *recon[3];                         /* tight planes, unfiltered reconstruction */
    uint8_t *pred_depth, *part_size, *mode_y, *mode_c, *tr_idx, *qp, *pred_mode, *cbf[3];   /* per 4x4 unit, picture raster over whole CTUs */
    int16_t *coeff;                            /* per CTU 64*64 + 2*32*32 levels in the reference's own layout */
    int32_t slice_type, slice_qp;
} refdrv_capture;
static refdrv_capture *g_cap;
void refdrv_capture_set(refdrv_capture *c) { g_cap = c; if (c) c->n_seen = 0; }

void hmr_deblock_sao_pad_sync_ctu(henc_thread_t *et, slice_t *currslice, ctu_info_t *ctu)
{
    static fn_sync_ctu real;
    if (!real) {
        void *ref = dlopen("libhomer_ref.so", RTLD_LAZY | RTLD_NOLOAD);
        real = (fn_sync_ctu)(ref ? dlsym(ref, "hmr_deblock_sao_pad_sync_ctu") : NULL);
        if (!real) real = (fn_sync_ctu)dlsym(RTLD_NEXT, "hmr_deblock_sao_pad_sync_ctu");
        if (!real) { fprintf(stderr, "ref_hooks: cannot find the reference's hmr_deblock_sao_pad_sync_ctu\n"); abort(); }
    }
    refdrv_capture *c = g_cap;
    if (c && et->enc_engine->num_encoded_frames == c->frame && c->width == et->pict_width[Y_COMP] && c->height == et->pict_height[Y_COMP]) {
        const wnd_t *img = &et->enc_engine->curr_reference_frame->img;
        const int cols = (c->width + 63) / 64, units_w = cols * 16;
        const int cx = ctu->x[Y_COMP] / 64, cy = ctu->y[Y_COMP] / 64;
        for (int comp = 0; comp < 3; comp++) {
            const int pw = comp ? c->width / 2 : c->width, ph = comp ? c->height / 2 : c->height, cs = comp ? 32 : 64;
            const int16_t *src = (const int16_t *)img->pwnd[comp];
            for (int y = cy * cs; y < cy * cs + cs && y < ph; y++)
                for (int x = cx * cs; x < cx * cs + cs && x < pw; x++) c->recon[comp][(size_t)y * pw + x] = (uint8_t)src[(size_t)y * img->window_size_x[comp] + x];
        }
        for (int r = 0; r < 256; r++) {
            const int a = et->enc_engine->raster2abs_table[r];
            const size_t u = (size_t)(cy * 16 + r / 16) * units_w + cx * 16 + r % 16;
            c->pred_depth[u] = ctu->pred_depth[a]; c->part_size[u] = ctu->part_size_type[a]; c->mode_y[u] = ctu->intra_mode[Y_COMP][a];
            c->mode_c[u] = ctu->intra_mode[CHR_COMP][a]; c->tr_idx[u] = ctu->tr_idx[a]; c->qp[u] = ctu->qp[a]; c->pred_mode[u] = ctu->pred_mode[a];
            for (int comp = 0; comp < 3; comp++) c->cbf[comp][u] = ctu->cbf[comp][a];
        }
        int16_t *dst = c->coeff + (size_t)(cy * cols + cx) * (64 * 64 + 2 * 32 * 32);
        memcpy(dst, ctu->coeff_wnd->pwnd[Y_COMP], sizeof(int16_t) * 64 * 64);
        memcpy(dst + 64 * 64, ctu->coeff_wnd->pwnd[U_COMP], sizeof(int16_t) * 32 * 32);
        memcpy(dst + 64 * 64 + 32 * 32, ctu->coeff_wnd->pwnd[V_COMP], sizeof(int16_t) * 32 * 32);
        c->slice_type = currslice->slice_type; c->slice_qp = currslice->qp;
        c->n_seen++;
    }
    real(et, currslice, ctu);
}
