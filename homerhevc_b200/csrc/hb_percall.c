/*
 * hb_percall.c -- the members of low_level_funcs_t (hmr_private.h:1063-1092) as GPU calls.
 *
 * Same prototypes, same caller-owned HOST buffers as the reference's sse_* / plain-C functions.  Every call:
 *   1. packs the operands into the calling thread's pinned, device-mapped staging area,
 *   2. launches one kernel on the calling thread's own stream (kernels read/write the staging area directly over PCIe),
 *   3. waits for the stream and unpacks the result.
 * One stream + staging area per thread (thread-local), so the table can be shared by all encoder threads without
 * locks, as the reference does (hmr_encoder_lib.c:1262).  No CPU path: a missing device aborts with a message.
 */
#define _POSIX_C_SOURCE 200809L
#include "hb_host.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define STAGE_BYTES (256 * 1024)

typedef struct pc_slot {
    void *stream;
    char *host;        /* pinned + mapped */
    char *dev;         /* device alias of host */
} pc_slot;

static hb_ctx *g_ctx;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static __thread pc_slot t_slot;
static pthread_key_t g_slot_key;          /* its destructor releases a thread's stream and staging area when the thread exits */

static void slot_release(void *p)
{
    pc_slot *s = (pc_slot *)p;
    if (!s) return;
    if (g_ctx) hbc_set_device(g_ctx->device);
    if (s->stream) { hbc_stream_sync(s->stream); hbc_stream_destroy(s->stream); }
    if (s->host) hbc_host_free(s->host);
    free(s);
}

static void die(const char *what)
{
    fprintf(stderr, "libhomer_b200: %s: %s -- no CPU fallback, aborting\n", what, hb_last_error());
    abort();
}

static void default_ctx_init(void)
{
    const char *e = getenv("HB_DEVICE");
    if (hb_ctx_create(&g_ctx, e ? atoi(e) : 0) != HB_OK) die("default context");
    if (pthread_key_create(&g_slot_key, slot_release)) die("per-thread key");
}

hb_ctx *hb_default_ctx(void)
{
    pthread_once(&g_once, default_ctx_init);
    return g_ctx;
}

static pc_slot *slot(void)
{
    static int poison = -1;                              /* $HB_POISON_STAGING: fill the staging with a pattern before every call (tests) */
    if (poison < 0) { const char *e = getenv("HB_POISON_STAGING"); poison = e && *e == '1'; }
    if (!t_slot.stream) {
        hb_ctx *ctx = hb_default_ctx();
        int rc;
        hbc_set_device(ctx->device);
        if ((rc = hbc_stream_create(&t_slot.stream)) || (rc = hbc_host_alloc((void **)&t_slot.host, STAGE_BYTES)) ||
            (rc = hbc_host_devptr(t_slot.host, (void **)&t_slot.dev))) { hbi_cuda_fail(rc, "per-thread slot"); die("per-call setup"); }
        memset(t_slot.host, 0, STAGE_BYTES);             /* pinned pages may be recycled ones */
        /* an encoder that spawns its workers per frame or GOP must not leak a stream and 256 KiB of pinned memory per thread:
         * a heap copy of the handles rides on the thread-specific key and is released by its destructor at thread exit */
        pc_slot *own = (pc_slot *)malloc(sizeof *own);
        if (!own) die("per-call setup: out of memory");
        *own = t_slot;
        if (pthread_setspecific(g_slot_key, own)) die("per-call setup: pthread_setspecific");
    }
    if (poison) memset(t_slot.host, 0xA5, STAGE_BYTES);
    return &t_slot;
}

static void finish(pc_slot *s, int launch_rc, const char *what)
{
    int rc = launch_rc;
    if (!rc) rc = hbc_stream_sync(s->stream);
    if (rc) { hbi_cuda_fail(rc, what); die(what); }
    __atomic_add_fetch(&g_ctx->launches, 1, __ATOMIC_RELAXED);
}

/* staging is carved in int16 units; D() gives the device alias of a host staging pointer */
#define D(s, p) ((void *)((s)->dev + ((char *)(p) - (s)->host)))

static void pack(int16_t *dst, int dst_stride, const int16_t *src, int src_stride, int w, int h)
{
    for (int r = 0; r < h; r++) memcpy(dst + (size_t)r * dst_stride, src + (ptrdiff_t)r * src_stride, sizeof(int16_t) * (size_t)w);
}

static int size_ok(int n) { return n == 4 || n == 8 || n == 16 || n == 32 || n == 64; }

static uint32_t sad_like(int16_t *src, uint32_t src_stride, int16_t *pred, uint32_t pred_stride, int size, int squared)
{
    pc_slot *s = slot();
    if (!size_ok(size)) return 0;                       /* the reference falls through silently on other sizes */
    int16_t *a = (int16_t *)s->host, *b = a + 64 * 64;
    uint32_t *out = (uint32_t *)(b + 64 * 64);
    pack(a, size, src, (int)src_stride, size, size);
    pack(b, size, pred, (int)pred_stride, size, size);  /* stride 0 = one row against every row (hmr_motion_inter.c:94) */
    finish(s, hbk_pc_sad(D(s, a), size, D(s, b), size, size, squared, D(s, out), s->stream), squared ? "ssd16b" : "sad");
    return *out;
}

uint32_t hb_sad(int16_t *src, uint32_t src_stride, int16_t *pred, uint32_t pred_stride, int size)
{
    return sad_like(src, src_stride, pred, pred_stride, size, 0);
}
uint32_t hb_ssd16b(int16_t *src, uint32_t src_stride, int16_t *pred, uint32_t pred_stride, int size)
{
    return sad_like(src, src_stride, pred, pred_stride, size, 1);
}

void hb_predict(int16_t *orig, int orig_stride, int16_t *pred, int pred_stride, int16_t *residual, int residual_stride, int size)
{
    pc_slot *s = slot();
    if (!size_ok(size)) return;
    int16_t *a = (int16_t *)s->host, *b = a + 64 * 64, *c = b + 64 * 64;
    pack(a, size, orig, orig_stride, size, size);
    pack(b, size, pred, pred_stride, size, size);
    finish(s, hbk_pc_predict(D(s, a), size, D(s, b), size, D(s, c), size, size, s->stream), "predict");
    pack(residual, residual_stride, c, size, size, size);
}

void hb_reconst(int16_t *pred, int pred_stride, int16_t *residual, int residual_stride, int16_t *decoded, int decoded_stride, int size)
{
    pc_slot *s = slot();
    if (!size_ok(size)) return;
    int16_t *a = (int16_t *)s->host, *b = a + 64 * 64, *c = b + 64 * 64;
    pack(a, size, pred, pred_stride, size, size);
    pack(b, size, residual, residual_stride, size, size);   /* stride 0: the all-zero row of the "no residual" calls */
    finish(s, hbk_pc_reconst(D(s, a), size, D(s, b), size, D(s, c), size, size, s->stream), "reconst");
    pack(decoded, decoded_stride, c, size, size, size);
}

/* weighted_average_motion of the table (hmr_motion_inter.c:2903): the two 14-bit predictions of a bi-predicted block -> 8-bit samples */
void hb_weighted_average_motion(int16_t *src0, int src0_stride, int16_t *src1, int src1_stride, int16_t *dst, int dst_stride, int height, int width, int bit_depth)
{
    pc_slot *s = slot();
    (void)bit_depth;                                        /* 8-bit video only, like the rest of the library */
    if (width <= 0 || height <= 0 || width > 64 || height > 64) return;
    int16_t *a = (int16_t *)s->host, *b = a + 64 * 64, *c = b + 64 * 64;
    pack(a, width, src0, src0_stride, width, height);
    pack(b, width, src1, src1_stride, width, height);
    finish(s, hbk_pc_wavg(D(s, a), D(s, b), D(s, c), width, height, s->stream), "weighted_average_motion");
    pack(dst, dst_stride, c, width, width, height);
}

static void interpolate(int chroma, int16_t *ref, int ref_stride, int16_t *dst, int dst_stride, int fraction, int width, int height,
                        int is_vertical, int is_first, int is_last)
{
    pc_slot *s = slot();
    if (width <= 0 || height <= 0 || width > 80 || height > 80 || fraction < 0 || fraction > (chroma ? 7 : 3)) return;
    /* taps reach 3 before / 4 after the sample (luma) or 1 / 2 (chroma) along the filtered axis; fraction 0 touches nothing
     * but the sample itself, so no margin is read from the caller's buffer in that case (staging margins stay zero) */
    const int before = fraction ? (chroma ? 1 : 3) : 0, after = fraction ? (chroma ? 2 : 4) : 0;
    const int mb = 3, ma = 4;                               /* staging margins are always laid out for the widest filter */
    const int sw = is_vertical ? width : width + mb + ma, sh = is_vertical ? height + mb + ma : height;
    int16_t *in = (int16_t *)s->host;
    int16_t *out = in + (size_t)sw * sh;
    memset(in, 0, sizeof(int16_t) * (size_t)sw * sh);
    const int org_off = is_vertical ? mb * sw : mb;
    if (is_vertical) pack(in + (mb - before) * sw, sw, ref - (ptrdiff_t)before * ref_stride, ref_stride, width, height + before + after);
    else pack(in + (mb - before), sw, ref - before, ref_stride, width + before + after, height);
    finish(s, hbk_pc_interp(D(s, in), sw, org_off, D(s, out), width, chroma, fraction, width, height, is_vertical, is_first, is_last, s->stream),
           chroma ? "interpolate_chroma" : "interpolate_luma");
    pack(dst, dst_stride, out, width, width, height);
}

void hb_interpolate_luma(int16_t *reference_buff, int reference_buff_stride, int16_t *pred_buff, int pred_buff_stride,
                         int fraction, int width, int height, int is_vertical, int is_first, int is_last)
{
    interpolate(0, reference_buff, reference_buff_stride, pred_buff, pred_buff_stride, fraction, width, height, is_vertical, is_first, is_last);
}
void hb_interpolate_chroma(int16_t *reference_buff, int reference_buff_stride, int16_t *pred_buff, int pred_buff_stride,
                           int fraction, int width, int height, int is_vertical, int is_first, int is_last)
{
    interpolate(1, reference_buff, reference_buff_stride, pred_buff, pred_buff_stride, fraction, width, height, is_vertical, is_first, is_last);
}

static int tu_ok(int w, int h) { return w == h && (w == 4 || w == 8 || w == 16 || w == 32); }

void hb_transform(int bit_depth, int16_t *block, int16_t *coeff, int block_size, int iWidth, int iHeight,
                  int width_shift, int height_shift, uint16_t uiMode, int16_t *aux)
{
    pc_slot *s = slot();
    (void)width_shift; (void)height_shift; (void)aux;
    if (!tu_ok(iWidth, iHeight)) return;                    /* as the reference: other shapes do nothing (hmr_transform.c:519-546) */
    if (bit_depth != 8) { hbi_fail(HB_ERR_ARG, "bit depth %d (8-bit video only)", bit_depth); die("transform"); }
    const int n = iWidth;
    int16_t *a = (int16_t *)s->host, *c = a + 32 * 32;
    pack(a, n, block, block_size, n, n);
    finish(s, hbk_pc_transform(D(s, a), n, D(s, c), n, n == 4 && uiMode != HB_REG_DCT, s->stream), "transform");
    memcpy(coeff, c, sizeof(int16_t) * (size_t)n * n);
}

void hb_itransform(int bit_depth, int16_t *block, int16_t *coeff, int block_size, int iWidth, int iHeight,
                   unsigned int uiMode, int16_t *aux)
{
    pc_slot *s = slot();
    (void)aux;
    if (!tu_ok(iWidth, iHeight)) return;
    if (bit_depth != 8) { hbi_fail(HB_ERR_ARG, "bit depth %d (8-bit video only)", bit_depth); die("itransform"); }
    const int n = iWidth;
    int16_t *c = (int16_t *)s->host, *b = c + 32 * 32;
    memcpy(c, coeff, sizeof(int16_t) * (size_t)n * n);
    finish(s, hbk_pc_itransform(D(s, b), n, D(s, c), n, n == 4 && uiMode != HB_REG_DCT, s->stream), "itransform");
    pack(block, block_size, b, n, n, n);
}

void hb_quant(const hb_quant_env *env, int16_t *src, int16_t *dst, int scan_mode, int depth, int comp, int cu_mode,
              int is_intra, int *ac_sum, int cu_size, int per, int rem)
{
    pc_slot *s = slot();
    hb_ctx *ctx = hb_default_ctx();
    (void)cu_mode;
    const int lg = env->max_cu_size_shift - (depth + (comp != 0));      /* inv_depth, hmr_sse42_functions_quant.c:39 */
    if (lg < 2 || lg > 5 || cu_size != (1 << lg) || scan_mode < HB_SCAN_HOR || scan_mode > HB_SCAN_DIAG || rem < 0 || rem > 5 ||
        env->bit_depth != 8) { hbi_fail(HB_ERR_ARG, "quant: unsupported shape (size %d, depth %d, comp %d, scan %d)", cu_size, depth, comp, scan_mode); die("quant"); }
    const int n = cu_size, list = (is_intra ? 0 : 3) + comp;
    const int qbits = 14 + per + (15 - env->bit_depth - lg);
    const int add = (int)((uint32_t)(env->is_islice ? 171 : 85) << (qbits - 9));
    int16_t *a = (int16_t *)s->host, *l = a + 32 * 32, *u = l + 32 * 32;
    int32_t *sum = (int32_t *)(u + 32 * 32);
    memcpy(a, src, sizeof(int16_t) * (size_t)n * n);
    finish(s, hbk_pc_quant(D(s, a), D(s, l), D(s, u), D(s, sum), n, ctx->d_q + hbi_tab_q_off(lg, list, rem),
                           ctx->d_scan + hbi_tab_scan_off(scan_mode, lg), qbits, add, env->sign_hiding, s->stream), "quant");
    memcpy(dst, l, sizeof(int16_t) * (size_t)n * n);
    if (env->delta_u) memcpy(env->delta_u, u, sizeof(int16_t) * (size_t)n * n);
    *ac_sum = *sum;
}

void hb_inv_quant(const hb_quant_env *env, int16_t *src, int16_t *dst, int depth, int comp, int is_intra, int cu_size, int per, int rem)
{
    pc_slot *s = slot();
    hb_ctx *ctx = hb_default_ctx();
    const int lg = env->max_cu_size_shift - (depth + (comp != 0));
    if (lg < 2 || lg > 5 || cu_size != (1 << lg) || rem < 0 || rem > 5 || env->bit_depth != 8) {
        hbi_fail(HB_ERR_ARG, "inv_quant: unsupported shape (size %d, depth %d, comp %d)", cu_size, depth, comp); die("inv_quant");
    }
    const int n = cu_size, list = is_intra ? 0 : 3 + comp;  /* the SSE4.2 precedence quirk, hmr_sse42_functions_quant.c:138 */
    int16_t *a = (int16_t *)s->host, *d = a + 32 * 32;
    memcpy(a, src, sizeof(int16_t) * (size_t)n * n);
    finish(s, hbk_pc_inv_quant(D(s, a), D(s, d), n, ctx->d_dq + hbi_tab_q_off(lg, list, rem), per, s->stream), "inv_quant");
    memcpy(dst, d, sizeof(int16_t) * (size_t)n * n);
}

/* create_intra_planar_prediction / create_intra_angular_prediction (hmr_private.h:1076-1077) without the henc_thread_t* / ctu_info_t*
 * the reference only uses for scratch rows and the (always set) DC neighbour flags */
static void intra_predict(int16_t *pred, int pred_stride, int16_t *adi, int adi_size, int cu_size, int mode, int is_luma)
{
    pc_slot *s = slot();
    if ((cu_size != 4 && cu_size != 8 && cu_size != 16 && cu_size != 32) || adi_size != 4 * cu_size + 1 || mode < 0 || mode > 34) {
        hbi_fail(HB_ERR_ARG, "intra prediction: unsupported shape (size %d, adi_size %d, mode %d)", cu_size, adi_size, mode); die("intra prediction");
    }
    int16_t *a = (int16_t *)s->host, *p = a + 160;
    memcpy(a, adi, sizeof(int16_t) * (size_t)adi_size);
    finish(s, hbk_pc_intra(D(s, a), cu_size, mode, is_luma, D(s, p), cu_size, s->stream), "intra prediction");
    pack(pred, pred_stride, p, cu_size, cu_size, cu_size);
}
void hb_create_intra_planar_prediction(int16_t *prediction, int pred_stride, int16_t *adi_pred_buff, int adi_size, int cu_size, int cu_size_shift)
{
    (void)cu_size_shift;
    intra_predict(prediction, pred_stride, adi_pred_buff, adi_size, cu_size, 0, 1);
}
void hb_create_intra_angular_prediction(int16_t *prediction, int pred_stride, int16_t *adi_pred_buff, int adi_size, int cu_size, int cu_mode, int is_luma)
{
    intra_predict(prediction, pred_stride, adi_pred_buff, adi_size, cu_size, cu_mode, is_luma);
}

void hb_fill_low_level_funcs(hb_low_level_funcs *t)
{
    t->sad = hb_sad;
    t->ssd16b = hb_ssd16b;
    t->predict = hb_predict;
    t->reconst = hb_reconst;
    t->weighted_average_motion = hb_weighted_average_motion;
    t->interpolate_luma_m_compensation = hb_interpolate_luma;
    t->interpolate_chroma_m_compensation = hb_interpolate_chroma;
    t->interpolate_luma_m_estimation = hb_interpolate_luma;
    t->transform = hb_transform;
    t->itransform = hb_itransform;
    /* quant / inv_quant: installed by the adapter that turns henc_thread_t* into hb_quant_env */
}
