/*
 * hb_host.c -- C99 host layer of libhomer_b200: contexts, HEVC tables, resident frames, the per-call drop-ins of
 * low_level_funcs_t and the batched job API.  All GPU work goes through the thin C shim in hb_shim.h; there is no
 * CPU implementation of any operator in this file -- without a CUDA device every entry point fails.
 *
 * Reference citations are to /root/reference/src/homer_lib/<file>:<line>.
 */
#define _POSIX_C_SOURCE 200809L
#include "hb_host.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ errors */
static __thread char t_err[256];

const char *hb_last_error(void) { return t_err; }

int hbi_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
    return code;
}

int hbi_cuda_fail(int cuda_code, const char *what)
{
    return hbi_fail(HB_ERR_CUDA, "%s: %s", what, hbc_error_string(cuda_code));
}

/* ------------------------------------------------------------------ HEVC tables (host build, device resident)
 * Scans: hmr_tables.c:62-196.  Quant/dequant pyramids from the default scaling lists: hmr_tables.c:199-250,
 * hmr_tables.h:53-85, wiring hmr_encoder_lib.c:103-140. */
static const uint8_t k_list_intra[64] = {
    16, 16, 16, 16, 17, 18, 21, 24,  16, 16, 16, 16, 17, 19, 22, 25,  16, 16, 17, 18, 20, 22, 25, 29,  16, 16, 18, 21, 24, 27, 31, 36,
    17, 17, 20, 24, 30, 35, 41, 47,  18, 19, 22, 27, 35, 44, 54, 65,  21, 22, 25, 31, 41, 54, 70, 88,  24, 25, 29, 36, 47, 65, 88, 115 };
static const uint8_t k_list_inter[64] = {
    16, 16, 16, 16, 17, 18, 20, 24,  16, 16, 16, 17, 18, 20, 24, 25,  16, 16, 17, 18, 20, 24, 25, 28,  16, 17, 18, 20, 24, 25, 28, 33,
    17, 18, 20, 24, 25, 28, 33, 41,  18, 20, 24, 25, 28, 33, 41, 54,  20, 24, 25, 28, 33, 41, 54, 71,  24, 25, 28, 33, 41, 54, 71, 91 };
static const int32_t k_fwd_scale[6] = { 26214, 23302, 20560, 18396, 16384, 14564 };
static const int32_t k_inv_scale[6] = { 40, 45, 51, 57, 64, 72 };
static const uint8_t k_chroma_qp[58] = {            /* hmr_encoder_lib.c:2245 */
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
    29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51 };

int hbi_chroma_qp(int qp, int offset)
{
    int v = qp + offset;
    v = v < 0 ? 0 : (v > 57 ? 57 : v);
    return k_chroma_qp[v];
}

/* anti-diagonal walk of a w x w grid starting bottom-left of each diagonal; emits row*w+col */
static void diag_walk(uint16_t *out, int w)
{
    int n = 0;
    for (int d = 0; d <= 2 * (w - 1); d++)
        for (int col = (d < w ? 0 : d - w + 1); col < w && col <= d; col++)
            out[n++] = (uint16_t)((d - col) * w + col);
}

static void build_scan(uint16_t *out, int mode, int lg)
{
    const int w = 1 << lg, g = w >> 2;
    int n = 0;
    if (mode == HB_SCAN_DIAG) {
        if (w == 4) { diag_walk(out, 4); return; }
        uint16_t group_order[64], inner[16];
        diag_walk(group_order, g);
        diag_walk(inner, 4);
        for (int b = 0; b < g * g; b++) {
            const int gy = group_order[b] / g, gx = group_order[b] % g;
            for (int i = 0; i < 16; i++)
                out[n++] = (uint16_t)((4 * gy + inner[i] / 4) * w + 4 * gx + inner[i] % 4);
        }
    } else if (mode == HB_SCAN_HOR) {
        for (int gy = 0; gy < g; gy++) for (int gx = 0; gx < g; gx++)
            for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
                out[n++] = (uint16_t)((4 * gy + y) * w + 4 * gx + x);
    } else {
        for (int gx = 0; gx < g; gx++) for (int gy = 0; gy < g; gy++)
            for (int x = 0; x < 4; x++) for (int y = 0; y < 4; y++)
                out[n++] = (uint16_t)((4 * gy + y) * w + 4 * gx + x);
    }
}

static void build_qtables(int32_t *q, int32_t *dq, int lg, int list, int rem)
{
    const int w = 1 << lg, up = w > 8 ? w / 8 : 1;
    const int inter = (lg == 5) ? (list >= 1) : (list >= 3);
    const uint8_t *m8 = inter ? k_list_inter : k_list_intra;
    for (int y = 0; y < w; y++)
        for (int x = 0; x < w; x++) {
            const int m = (lg == 2) ? 16 : m8[8 * (y / up) + x / up];
            q[y * w + x] = (k_fwd_scale[rem] << 4) / m;
            dq[y * w + x] = k_inv_scale[rem] * m;
        }
    if (up > 1) { q[0] = (k_fwd_scale[rem] << 4) / 16; dq[0] = k_inv_scale[rem] * 16; }   /* DC of 16x16 / 32x32 lists */
}

size_t hbi_tab_scan_off(int mode, int lg)      /* in uint16 elements; mode 1..3, lg 2..5 */
{
    size_t off = 0;
    for (int m = 1; m <= 3; m++)
        for (int l = 2; l <= 5; l++) {
            if (m == mode && l == lg) return off;
            off += (size_t)1 << (2 * l);
        }
    return 0;
}
size_t hbi_tab_q_off(int lg, int list, int rem) /* in int32 elements */
{
    size_t off = 0;
    for (int l = 2; l <= 5; l++)
        for (int li = 0; li < 6; li++)
            for (int r = 0; r < 6; r++) {
                if (l == lg && li == list && r == rem) return off;
                off += (size_t)1 << (2 * l);
            }
    return 0;
}

static int tables_upload(hb_ctx *ctx)
{
    const size_t n_scan = 3 * (16 + 64 + 256 + 1024);
    const size_t n_q = 36 * (16 + 64 + 256 + 1024);
    uint16_t *scan = (uint16_t *)malloc(n_scan * sizeof *scan);
    int32_t *q = (int32_t *)malloc(n_q * sizeof *q), *dq = (int32_t *)malloc(n_q * sizeof *dq);
    int rc;
    if (!scan || !q || !dq) { free(scan); free(q); free(dq); return hbi_fail(HB_ERR_NOMEM, "tables: out of host memory"); }
    for (int m = 1; m <= 3; m++) for (int l = 2; l <= 5; l++) build_scan(scan + hbi_tab_scan_off(m, l), m, l);
    for (int l = 2; l <= 5; l++) for (int li = 0; li < 6; li++) for (int r = 0; r < 6; r++) {
        /* 32x32 has the two lists 0 (intra) and 1 (inter); index 3 aliases 1 (hmr_encoder_lib.c:135-140), the rest is never used */
        const int src_list = (l == 5 && li >= 1) ? 1 : li;
        build_qtables(q + hbi_tab_q_off(l, li, r), dq + hbi_tab_q_off(l, li, r), l, src_list, r);
    }
    if ((rc = hbc_malloc((void **)&ctx->d_scan, n_scan * sizeof *scan)) || (rc = hbc_malloc((void **)&ctx->d_q, n_q * sizeof *q)) ||
        (rc = hbc_malloc((void **)&ctx->d_dq, n_q * sizeof *dq))) { free(scan); free(q); free(dq); return hbi_cuda_fail(rc, "tables: cudaMalloc"); }
    rc = hbc_h2d_async(ctx->d_scan, scan, n_scan * sizeof *scan, ctx->stream);
    if (!rc) rc = hbc_h2d_async(ctx->d_q, q, n_q * sizeof *q, ctx->stream);
    if (!rc) rc = hbc_h2d_async(ctx->d_dq, dq, n_q * sizeof *dq, ctx->stream);
    if (!rc) rc = hbc_stream_sync(ctx->stream);
    free(scan); free(q); free(dq);
    return rc ? hbi_cuda_fail(rc, "tables: upload") : HB_OK;
}

/* ------------------------------------------------------------------ contexts */
int hb_device_count(void) { return hbc_device_count(); }

int hb_ctx_create(hb_ctx **out, int device)
{
    int rc;
    if (!out) return hbi_fail(HB_ERR_ARG, "hb_ctx_create: out is NULL");
    *out = NULL;
    if (hbc_device_count() <= 0) return hbi_fail(HB_ERR_CUDA, "hb_ctx_create: no CUDA device visible (there is no CPU fallback)");
    if (device < 0 || device >= hbc_device_count()) return hbi_fail(HB_ERR_ARG, "hb_ctx_create: device %d out of range", device);
    hb_ctx *ctx = (hb_ctx *)calloc(1, sizeof *ctx);
    if (!ctx) return hbi_fail(HB_ERR_NOMEM, "hb_ctx_create: out of memory");
    ctx->device = device;
    if ((rc = hbc_set_device(device)) || (rc = hbc_stream_create(&ctx->stream)) ||
        (rc = hbc_event_create(&ctx->ev[0])) || (rc = hbc_event_create(&ctx->ev[1]))) {
        free(ctx);
        return hbi_cuda_fail(rc, "hb_ctx_create");
    }
    pthread_mutex_init(&ctx->lock, NULL);
    if ((rc = hbk_me_configure())) { hb_ctx_destroy(ctx); return hbi_cuda_fail(rc, "hb_ctx_create: kernel attributes"); }
    if ((rc = tables_upload(ctx)) != HB_OK) { hb_ctx_destroy(ctx); return rc; }
    if ((rc = hbc_malloc((void **)&ctx->d_flag, 256))) { hb_ctx_destroy(ctx); return hbi_cuda_fail(rc, "hb_ctx_create: flag"); }
    hbc_memset_async(ctx->d_flag, 0, 256, ctx->stream);
    *out = ctx;
    return HB_OK;
}

void hb_ctx_destroy(hb_ctx *ctx)
{
    if (!ctx) return;
    hbc_set_device(ctx->device);
    if (ctx->stream) hbc_stream_sync(ctx->stream);
    for (int i = 0; i < HB_N_SCRATCH; i++) { if (ctx->d_scratch[i]) hbc_free(ctx->d_scratch[i]); if (ctx->h_scratch[i]) hbc_host_free(ctx->h_scratch[i]); }
    if (ctx->d_scan) hbc_free(ctx->d_scan);
    if (ctx->d_q) hbc_free(ctx->d_q);
    if (ctx->d_dq) hbc_free(ctx->d_dq);
    if (ctx->d_flag) hbc_free(ctx->d_flag);
    if (ctx->ev[0]) hbc_event_destroy(ctx->ev[0]);
    if (ctx->ev[1]) hbc_event_destroy(ctx->ev[1]);
    if (ctx->ev_sync) hbc_event_destroy(ctx->ev_sync);
    if (ctx->stream) hbc_stream_destroy(ctx->stream);
    pthread_mutex_destroy(&ctx->lock);
    free(ctx);
}

int hb_ctx_sync(hb_ctx *ctx)
{
    if (!ctx) return hbi_fail(HB_ERR_ARG, "hb_ctx_sync: NULL context");
    const int rc = hbc_stream_sync(ctx->stream);
    return rc ? hbi_cuda_fail(rc, "hb_ctx_sync") : HB_OK;
}
void *hb_ctx_stream(hb_ctx *ctx) { return ctx ? ctx->stream : NULL; }

/* everything queued on `waiter` after this call starts only when everything queued so far on `signaler` has finished */
int hb_ctx_wait(hb_ctx *waiter, hb_ctx *signaler)
{
    int rc;
    if (!waiter || !signaler) return hbi_fail(HB_ERR_ARG, "hb_ctx_wait: NULL context");
    if (!signaler->ev_sync && (rc = hbc_event_create_notiming(&signaler->ev_sync))) return hbi_cuda_fail(rc, "hb_ctx_wait: event");
    if ((rc = hbc_event_record(signaler->ev_sync, signaler->stream)) || (rc = hbc_stream_wait_event(waiter->stream, signaler->ev_sync)))
        return hbi_cuda_fail(rc, "hb_ctx_wait");
    return HB_OK;
}
uint64_t hb_ctx_launch_count(hb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int hb_timer_begin(hb_ctx *ctx)
{
    if (!ctx) return hbi_fail(HB_ERR_ARG, "hb_timer_begin: NULL context");
    const int rc = hbc_event_record(ctx->ev[0], ctx->stream);
    return rc ? hbi_cuda_fail(rc, "hb_timer_begin") : HB_OK;
}
int hb_timer_end(hb_ctx *ctx, float *ms_out)
{
    if (!ctx || !ms_out) return hbi_fail(HB_ERR_ARG, "hb_timer_end: NULL argument");
    int rc = hbc_event_record(ctx->ev[1], ctx->stream);
    if (!rc) rc = hbc_event_elapsed(ctx->ev[0], ctx->ev[1], ms_out);
    return rc ? hbi_cuda_fail(rc, "hb_timer_end") : HB_OK;
}

void *hb_pinned_alloc(size_t bytes)
{
    void *p = NULL;
    const int rc = hbc_host_alloc(&p, bytes);
    if (rc) { hbi_cuda_fail(rc, "hb_pinned_alloc"); return NULL; }
    return p;
}
void hb_pinned_free(void *p) { if (p) hbc_host_free(p); }

/* grow-only scratch: device buffer i and its pinned host twin */
int hbi_scratch(hb_ctx *ctx, int i, size_t bytes, void **dev, void **host)
{
    int rc;
    if (ctx->scratch_bytes[i] < bytes) {
        size_t cap = ctx->scratch_bytes[i] ? ctx->scratch_bytes[i] : 4096;
        while (cap < bytes) cap *= 2;
        if ((rc = hbc_stream_sync(ctx->stream))) return hbi_cuda_fail(rc, "scratch: sync");
        if (ctx->d_scratch[i]) hbc_free(ctx->d_scratch[i]);
        if (ctx->h_scratch[i]) hbc_host_free(ctx->h_scratch[i]);
        ctx->d_scratch[i] = ctx->h_scratch[i] = NULL; ctx->scratch_bytes[i] = 0;
        if ((rc = hbc_malloc(&ctx->d_scratch[i], cap))) return hbi_cuda_fail(rc, "scratch: cudaMalloc");
        if ((rc = hbc_host_alloc(&ctx->h_scratch[i], cap))) return hbi_cuda_fail(rc, "scratch: cudaHostAlloc");
        ctx->scratch_bytes[i] = cap;
    }
    if (dev) *dev = ctx->d_scratch[i];
    if (host) *host = ctx->h_scratch[i];
    return HB_OK;
}

/* ------------------------------------------------------------------ resident frames */
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

/* layout of plane c of a width x height picture (also what a view of another process's picture assumes) */
static void plane_geometry(hbd_plane *p, int c, int width, int height)
{
    p->w = c ? width / 2 : width; p->h = c ? height / 2 : height;
    p->pad = c ? HB_PAD_LUMA / 2 : HB_PAD_LUMA;
    p->pitch = (int32_t)align_up((size_t)p->w + 2 * (size_t)p->pad, 128);
}
int hb_frame_create(hb_ctx *ctx, int width, int height, hb_frame **out)
{
    if (!ctx || !out) return hbi_fail(HB_ERR_ARG, "hb_frame_create: NULL argument");
    if (width < 16 || height < 16 || (width & 7) || (height & 7)) return hbi_fail(HB_ERR_ARG, "hb_frame_create: %dx%d must be multiples of 8, at least 16", width, height);
    hb_frame *f = (hb_frame *)calloc(1, sizeof *f);
    if (!f) return hbi_fail(HB_ERR_NOMEM, "hb_frame_create: out of memory");
    f->ctx = ctx; f->w = width; f->h = height;
    hbc_set_device(ctx->device);
    for (int c = 0; c < 3; c++) {
        hbd_plane *p = &f->d.p[c];
        plane_geometry(p, c, width, height);
        const size_t bytes = (size_t)p->pitch * ((size_t)p->h + 2 * (size_t)p->pad) + 256;
        const int rc = hbc_malloc((void **)&p->base, bytes);
        if (rc) { hb_frame_destroy(f); return hbi_cuda_fail(rc, "hb_frame_create: cudaMalloc"); }
        hbc_memset_async(p->base, 0, bytes, ctx->stream);
        p->org = p->base + (size_t)p->pad * p->pitch + p->pad;
    }
    *out = f;
    return HB_OK;
}

void hb_frame_destroy(hb_frame *f)
{
    if (!f) return;
    hbc_set_device(f->ctx->device);
    hbc_stream_sync(f->ctx->stream);
    for (int c = 0; c < 3; c++) if (f->d.p[c].base) { if (f->ipc_view) hbc_ipc_close_mem(f->d.p[c].base); else hbc_free(f->d.p[c].base); }
    if (f->stage) hbc_free(f->stage);
    if (f->sp.base) hbc_free(f->sp.base);
    free(f);
}

int hbi_frame_subpel_alloc(hb_frame *f)
{
    if (f->sp.base) return HB_OK;
    hbd_subpel sp;
    memset(&sp, 0, sizeof sp);
    sp.w = f->w + 2 * HB_SUBPEL_OFF; sp.h = f->h + 2 * HB_SUBPEL_OFF;
    sp.pitch = (int32_t)align_up((size_t)sp.w, 128);
    sp.plane_bytes = (uint64_t)sp.pitch * (uint64_t)sp.h;
    hbc_set_device(f->ctx->device);
    const int rc = hbc_malloc((void **)&sp.base, (size_t)(15 * sp.plane_bytes) + 256);
    if (rc) return hbi_cuda_fail(rc, "sub-pel planes: cudaMalloc");
    f->sp = sp;
    return HB_OK;
}
int hb_frame_width(const hb_frame *f) { return f ? f->w : 0; }
int hb_frame_height(const hb_frame *f) { return f ? f->h : 0; }

int hb_frame_upload_u8_ex(hb_ctx *ctx, hb_frame *f, const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs, int flags);
int hb_frame_upload_u8(hb_ctx *ctx, hb_frame *f, const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs)
{
    return hb_frame_upload_u8_ex(ctx, f, y, ys, u, us, v, vs, 0);
}

/* flags: HB_UPLOAD_NO_BORDER skips the border replication -- enough for a CURRENT frame, whose samples are only ever read
 * inside the picture (the search reads the border of the REFERENCE frame only) */
int hb_frame_upload_u8_ex(hb_ctx *ctx, hb_frame *f, const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs, int flags)
{
    const uint8_t *src[3] = { y, u, v };
    const int st[3] = { ys, us, vs };
    int rc = 0;
    if (!ctx || !f || !y || !u || !v) return hbi_fail(HB_ERR_ARG, "hb_frame_upload_u8: NULL argument");
    hbc_set_device(ctx->device);
    /* One plain copy per plane (a single one when the caller's planes are one buffer) into a dense staging buffer -- 1-D copies
     * run at link speed, pitched ones at about half of it -- then one kernel spreads the samples into the padded planes and
     * replicates the borders in the same pass. */
    const size_t luma = (size_t)f->w * f->h;
    if (!f->stage && (rc = hbc_malloc((void **)&f->stage, luma + luma / 2 + 16))) return hbi_cuda_fail(rc, "hb_frame_upload_u8: cudaMalloc");
    /* one copy when the three planes are one tight buffer -- which adjacent addresses alone do not prove: three separately pinned planes
     * may touch, and the runtime refuses a copy that spans allocations (hbc_host_span_one_copy asks the driver) -- else plane by plane */
    if (ys == f->w && us == f->w / 2 && vs == f->w / 2 && u == y + luma && v == u + luma / 4 && hbc_host_span_one_copy(y, luma + luma / 2))
        rc = hbc_h2d_async(f->stage, y, luma + luma / 2, ctx->stream);
    else {
        size_t off = 0;
        for (int c = 0; c < 3 && !rc; c++) {
            const hbd_plane *p = &f->d.p[c];
            if (st[c] == p->w) rc = hbc_h2d_async(f->stage + off, src[c], (size_t)p->w * p->h, ctx->stream);
            else rc = hbc_h2d_2d_async(f->stage + off, (size_t)p->w, src[c], (size_t)st[c], (size_t)p->w, (size_t)p->h, ctx->stream);
            off += (size_t)p->w * p->h;
        }
    }
    if (!rc) { rc = hbk_ingest_frame(&f->d, f->stage, !(flags & HB_UPLOAD_NO_BORDER), ctx->stream); ctx->launches += 1; }
    return rc ? hbi_cuda_fail(rc, "hb_frame_upload_u8") : HB_OK;
}

int hb_frame_upload_i16(hb_ctx *ctx, hb_frame *f, const int16_t *y, int ys, const int16_t *u, int us, const int16_t *v, int vs)
{
    const int16_t *src[3] = { y, u, v };
    const int st[3] = { ys, us, vs };
    int rc = 0;
    void *dev = NULL, *host = NULL;
    if (!ctx || !f || !y || !u || !v) return hbi_fail(HB_ERR_ARG, "hb_frame_upload_i16: NULL argument");
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    const size_t luma = (size_t)f->w * f->h;
    if ((rc = hbi_scratch(ctx, 0, 2 * (luma + luma / 2) + 64, &dev, &host)) != HB_OK) { pthread_mutex_unlock(&ctx->lock); return rc; }
    size_t off = 0;
    uint32_t *flag_h = (uint32_t *)((char *)host);        /* reuse the first word of the pinned twin for the read-back */
    for (int c = 0; c < 3 && !rc; c++) {
        const hbd_plane *p = &f->d.p[c];
        int16_t *d = (int16_t *)dev + off;
        rc = hbc_h2d_2d_async(d, (size_t)p->w * 2, src[c], (size_t)st[c] * 2, (size_t)p->w * 2, (size_t)p->h, ctx->stream);
        if (!rc) { rc = hbk_narrow_plane(d, p->w, *p, ctx->d_flag, ctx->stream); ctx->launches++; }
        off += (size_t)p->w * p->h;
    }
    if (!rc) { rc = hbk_pad_frame(&f->d, ctx->stream); ctx->launches += 1; }
    if (!rc) rc = hbc_d2h_async(flag_h, ctx->d_flag, 4, ctx->stream);
    if (!rc) rc = hbc_memset_async(ctx->d_flag, 0, 4, ctx->stream);
    if (!rc) rc = hbc_stream_sync(ctx->stream);
    const uint32_t flag = rc ? 0 : *flag_h;
    pthread_mutex_unlock(&ctx->lock);
    if (rc) return hbi_cuda_fail(rc, "hb_frame_upload_i16");
    if (flag) return hbi_fail(HB_ERR_ARG, "hb_frame_upload_i16: samples outside 0..255 (8-bit video only)");
    return HB_OK;
}

int hb_frame_download_u8(hb_ctx *ctx, const hb_frame *f, uint8_t *y, int ys, uint8_t *u, int us, uint8_t *v, int vs)
{
    uint8_t *dst[3] = { y, u, v };
    const int st[3] = { ys, us, vs };
    int rc = 0;
    if (!ctx || !f || !y || !u || !v) return hbi_fail(HB_ERR_ARG, "hb_frame_download_u8: NULL argument");
    hbc_set_device(ctx->device);
    for (int c = 0; c < 3 && !rc; c++) {
        const hbd_plane *p = &f->d.p[c];
        rc = hbc_d2h_2d_async(dst[c], (size_t)st[c], p->org, (size_t)p->pitch, (size_t)p->w, (size_t)p->h, ctx->stream);
    }
    if (!rc) rc = hbc_stream_sync(ctx->stream);
    return rc ? hbi_cuda_fail(rc, "hb_frame_download_u8") : HB_OK;
}

/* Rows [row0, row0+n_rows) of one plane <-> a tight device buffer (width bytes per row), on the context's stream.  This is
 * what a CTU-row band exchanges with its neighbours: the caller moves the tight buffer between GPUs (NCCL send/recv or a peer
 * copy) and imports it on the other side; hb_frame_pad then refreshes the replicated border. */
int hb_frame_export_rows(hb_ctx *ctx, const hb_frame *f, int plane, int row0, int n_rows, void *dev_dst)
{
    if (!ctx || !f || !dev_dst || plane < 0 || plane > 2) return hbi_fail(HB_ERR_ARG, "hb_frame_export_rows: bad argument");
    const hbd_plane *p = &f->d.p[plane];
    if (row0 < 0 || n_rows < 0 || row0 + n_rows > p->h) return hbi_fail(HB_ERR_ARG, "hb_frame_export_rows: rows %d..%d outside the plane", row0, row0 + n_rows);
    if (!n_rows) return HB_OK;
    hbc_set_device(ctx->device);
    const int rc = hbc_d2d_2d_async(dev_dst, (size_t)p->w, p->org + (size_t)row0 * p->pitch, (size_t)p->pitch, (size_t)p->w, (size_t)n_rows, ctx->stream);
    return rc ? hbi_cuda_fail(rc, "hb_frame_export_rows") : HB_OK;
}
int hb_frame_import_rows(hb_ctx *ctx, hb_frame *f, int plane, int row0, int n_rows, const void *dev_src)
{
    if (!ctx || !f || !dev_src || plane < 0 || plane > 2) return hbi_fail(HB_ERR_ARG, "hb_frame_import_rows: bad argument");
    const hbd_plane *p = &f->d.p[plane];
    if (row0 < 0 || n_rows < 0 || row0 + n_rows > p->h) return hbi_fail(HB_ERR_ARG, "hb_frame_import_rows: rows %d..%d outside the plane", row0, row0 + n_rows);
    if (!n_rows) return HB_OK;
    hbc_set_device(ctx->device);
    const int rc = hbc_d2d_2d_async(p->org + (size_t)row0 * p->pitch, (size_t)p->pitch, dev_src, (size_t)p->w, (size_t)p->w, (size_t)n_rows, ctx->stream);
    return rc ? hbi_cuda_fail(rc, "hb_frame_import_rows") : HB_OK;
}
int hb_frame_pad(hb_ctx *ctx, hb_frame *f)
{
    if (!ctx || !f) return hbi_fail(HB_ERR_ARG, "hb_frame_pad: NULL argument");
    hbc_set_device(ctx->device);
    const int rc = hbk_pad_frame(&f->d, ctx->stream);
    ctx->launches += 1;
    return rc ? hbi_cuda_fail(rc, "hb_frame_pad") : HB_OK;
}

/* ---- peer pictures: CUDA IPC views and the halo pull (include/homer_b200.h) */
int hb_frame_ipc_export(const hb_frame *f, hb_frame_ipc *out)
{
    if (!f || !out) return hbi_fail(HB_ERR_ARG, "hb_frame_ipc_export: NULL argument");
    if (f->ipc_view) return hbi_fail(HB_ERR_ARG, "hb_frame_ipc_export: a view cannot be exported again");
    memset(out, 0, sizeof *out);
    hbc_set_device(f->ctx->device);
    for (int c = 0; c < 3; c++) {
        const int rc = hbc_ipc_get_mem(f->d.p[c].base, out->mem[c]);
        if (rc) return hbi_cuda_fail(rc, "hb_frame_ipc_export: cudaIpcGetMemHandle");
    }
    out->width = f->w; out->height = f->h;
    return HB_OK;
}
int hb_frame_ipc_open(hb_ctx *ctx, const hb_frame_ipc *in, hb_frame **view)
{
    if (!ctx || !in || !view) return hbi_fail(HB_ERR_ARG, "hb_frame_ipc_open: NULL argument");
    *view = NULL;
    if (in->width < 16 || in->height < 16 || (in->width & 7) || (in->height & 7)) return hbi_fail(HB_ERR_ARG, "hb_frame_ipc_open: bad geometry %dx%d", in->width, in->height);
    hb_frame *f = (hb_frame *)calloc(1, sizeof *f);
    if (!f) return hbi_fail(HB_ERR_NOMEM, "hb_frame_ipc_open: out of memory");
    f->ctx = ctx; f->w = in->width; f->h = in->height; f->ipc_view = 1;
    hbc_set_device(ctx->device);
    for (int c = 0; c < 3; c++) {
        hbd_plane *p = &f->d.p[c];
        plane_geometry(p, c, in->width, in->height);          /* the owner laid its planes out with the same rule (hb_frame_create) */
        const int rc = hbc_ipc_open_mem(in->mem[c], (void **)&p->base);
        if (rc) { p->base = NULL; hb_frame_destroy(f); return hbi_cuda_fail(rc, "hb_frame_ipc_open: cudaIpcOpenMemHandle"); }
        p->org = p->base + (size_t)p->pad * p->pitch + p->pad;
    }
    *view = f;
    return HB_OK;
}
int hb_frame_pull_rows(hb_ctx *ctx, hb_frame *dst, const hb_frame *const *srcs, int n_srcs, const hb_row_span *spans, int n_spans, int refresh_border)
{
    hbd_pull_span ps[HB_MAX_ROW_SPANS];
    if (!ctx || !dst || (n_spans > 0 && (!srcs || !spans)) || n_spans < 0 || n_spans > HB_MAX_ROW_SPANS) return hbi_fail(HB_ERR_ARG, "hb_frame_pull_rows: bad argument");
    if (dst->ipc_view) return hbi_fail(HB_ERR_ARG, "hb_frame_pull_rows: the destination is a view of another process's picture");
    for (int i = 0; i < n_spans; i++) {
        const hb_row_span *s = &spans[i];
        if (s->src < 0 || s->src >= n_srcs || !srcs[s->src] || s->plane < 0 || s->plane > 2) return hbi_fail(HB_ERR_ARG, "hb_frame_pull_rows: span %d: bad source or plane", i);
        const hb_frame *sf = srcs[s->src];
        if (sf->w != dst->w || sf->h != dst->h) return hbi_fail(HB_ERR_ARG, "hb_frame_pull_rows: span %d: picture sizes differ", i);
        const hbd_plane *sp = &sf->d.p[s->plane], *dp = &dst->d.p[s->plane];
        if (s->row0 < 0 || s->n_rows < 0 || s->row0 + s->n_rows > dp->h) return hbi_fail(HB_ERR_ARG, "hb_frame_pull_rows: span %d: rows %d..%d outside the plane", i, s->row0, s->row0 + s->n_rows);
        ps[i].src = sp->org + (size_t)s->row0 * sp->pitch; ps[i].dst = dp->org + (size_t)s->row0 * dp->pitch;
        ps[i].src_pitch = sp->pitch; ps[i].dst_pitch = dp->pitch; ps[i].width = dp->w; ps[i].rows = s->n_rows;
    }
    hbc_set_device(ctx->device);
    int rc = hbk_pull_rows(ps, n_spans, ctx->stream);
    if (!rc && n_spans) ctx->launches += 1;
    if (!rc && refresh_border) { rc = hbk_pad_frame(&dst->d, ctx->stream); ctx->launches += 1; }
    return rc ? hbi_cuda_fail(rc, "hb_frame_pull_rows") : HB_OK;
}
struct hb_ipc_event { void *ev; int device; };
int hb_ipc_event_create(hb_ctx *ctx, hb_ipc_event **ev, uint8_t handle[HB_IPC_HANDLE_BYTES])
{
    if (!ctx || !ev || !handle) return hbi_fail(HB_ERR_ARG, "hb_ipc_event_create: NULL argument");
    hb_ipc_event *e = (hb_ipc_event *)calloc(1, sizeof *e);
    if (!e) return hbi_fail(HB_ERR_NOMEM, "hb_ipc_event_create: out of memory");
    hbc_set_device(ctx->device);
    const int rc = hbc_ipc_event_create(&e->ev, handle);
    if (rc) { free(e); return hbi_cuda_fail(rc, "hb_ipc_event_create"); }
    e->device = ctx->device;
    *ev = e;
    return HB_OK;
}
int hb_ipc_event_open(hb_ctx *ctx, const uint8_t handle[HB_IPC_HANDLE_BYTES], hb_ipc_event **ev)
{
    if (!ctx || !ev || !handle) return hbi_fail(HB_ERR_ARG, "hb_ipc_event_open: NULL argument");
    hb_ipc_event *e = (hb_ipc_event *)calloc(1, sizeof *e);
    if (!e) return hbi_fail(HB_ERR_NOMEM, "hb_ipc_event_open: out of memory");
    hbc_set_device(ctx->device);
    const int rc = hbc_ipc_event_open(handle, &e->ev);
    if (rc) { free(e); return hbi_cuda_fail(rc, "hb_ipc_event_open"); }
    e->device = ctx->device;
    *ev = e;
    return HB_OK;
}
int hb_ipc_event_record(hb_ctx *ctx, hb_ipc_event *ev)
{
    if (!ctx || !ev) return hbi_fail(HB_ERR_ARG, "hb_ipc_event_record: NULL argument");
    hbc_set_device(ctx->device);
    const int rc = hbc_event_record(ev->ev, ctx->stream);
    return rc ? hbi_cuda_fail(rc, "hb_ipc_event_record") : HB_OK;
}
int hb_ipc_event_wait(hb_ctx *ctx, hb_ipc_event *ev)
{
    if (!ctx || !ev) return hbi_fail(HB_ERR_ARG, "hb_ipc_event_wait: NULL argument");
    hbc_set_device(ctx->device);
    const int rc = hbc_stream_wait_event(ctx->stream, ev->ev);
    return rc ? hbi_cuda_fail(rc, "hb_ipc_event_wait") : HB_OK;
}
void hb_ipc_event_destroy(hb_ipc_event *ev)
{
    if (!ev) return;
    hbc_set_device(ev->device);
    hbc_event_destroy(ev->ev);
    free(ev);
}

/* ------------------------------------------------------------------ batched jobs (host arrays in, host arrays out) */
static double mv_cost_weight(int qp, double avg_dist)     /* calc_mv_correction, hmr_common.h:53 */
{
    double w = avg_dist / 2000.;
    w = w < .15 ? .15 : (w > 1.4 ? 1.4 : w);
    return (uint32_t)qp * w;
}
double hbi_zero_out_k(double avg_dist)                     /* hmr_motion_inter.c:106 */
{
    double k = avg_dist / 2.5 - 5.;
    return k < 1. ? 1. : (k > 20000. ? 20000. : k);
}

static int me_search(hb_ctx *ctx, const hb_frame *cur, const hb_frame *ref, const hb_me_job *jobs, int n_jobs, const hb_me_result *parent_results, int n_parent,
                     double avg_dist, int action, hb_me_result *results, const hb_unit_info *units, int units_w);
int hb_me_search(hb_ctx *ctx, const hb_frame *cur, const hb_frame *ref, const hb_me_job *jobs, int n_jobs,
                 const hb_me_result *parent_results, int n_parent, double avg_dist, int action, hb_me_result *results)
{
    return me_search(ctx, cur, ref, jobs, n_jobs, parent_results, n_parent, avg_dist, action, results, NULL, 0);
}
/* the same with the AMVP predictors of every job derived ON THE DEVICE from the per-unit motion field (hb_amvp_candidates) right before its
 * search: jobs[i].amvp / n_amvp are ignored, the PUs must sit on their size grid */
int hb_me_search_field(hb_ctx *ctx, const hb_frame *cur, const hb_frame *ref, const hb_unit_info *units, int units_w, const hb_me_job *jobs, int n_jobs,
                       const hb_me_result *parent_results, int n_parent, double avg_dist, int action, hb_me_result *results)
{
    if (!units) return hbi_fail(HB_ERR_ARG, "hb_me_search_field: NULL unit field");
    if (cur && units_w < ((cur->w + 63) / 64) * 16) return hbi_fail(HB_ERR_ARG, "hb_me_search_field: units_w %d does not cover whole CTUs", units_w);
    for (int i = 0; jobs && i < n_jobs; i++)
        if (jobs[i].size > 0 && ((jobs[i].x % jobs[i].size) || (jobs[i].y % jobs[i].size))) return hbi_fail(HB_ERR_ARG, "hb_me_search_field: job %d is off its size grid", i);
    return me_search(ctx, cur, ref, jobs, n_jobs, parent_results, n_parent, avg_dist, action, results, units, units_w);
}
static int me_search(hb_ctx *ctx, const hb_frame *cur, const hb_frame *ref, const hb_me_job *jobs, int n_jobs, const hb_me_result *parent_results, int n_parent,
                     double avg_dist, int action, hb_me_result *results, const hb_unit_info *units, int units_w)
{
    static const int sizes[4] = { 64, 32, 16, 8 };
    int rc = HB_OK, crc = 0;
    void *d_jobs, *h_jobs, *d_res, *h_res, *d_par = NULL, *h_par = NULL, *d_units = NULL, *h_units = NULL;
    if (!ctx || !cur || !ref || !jobs || !results || n_jobs < 0) return hbi_fail(HB_ERR_ARG, "hb_me_search: bad argument");
    if (cur->w != ref->w || cur->h != ref->h) return hbi_fail(HB_ERR_ARG, "hb_me_search: frame sizes differ");
    if (n_jobs == 0) return HB_OK;
    for (int i = 0; i < n_jobs; i++) {
        const hb_me_job *j = &jobs[i];
        if ((j->size != 8 && j->size != 16 && j->size != 32 && j->size != 64) || j->x < 0 || j->y < 0 || j->x + j->size > cur->w ||
            j->y + j->size > cur->h || j->n_amvp < 0 || j->n_amvp > 2 || j->n_start < 0 || j->n_start > 3 || j->parent >= n_parent)
            return hbi_fail(HB_ERR_ARG, "hb_me_search: job %d is invalid", i);
    }
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    if ((rc = hbi_scratch(ctx, 0, sizeof(hbd_me_job) * (size_t)n_jobs, &d_jobs, &h_jobs)) != HB_OK) goto done;
    if ((rc = hbi_scratch(ctx, 1, sizeof(hb_me_result) * (size_t)n_jobs, &d_res, &h_res)) != HB_OK) goto done;
    if (n_parent > 0 && parent_results) {
        if ((rc = hbi_scratch(ctx, 2, sizeof(hb_me_result) * (size_t)n_parent, &d_par, &h_par)) != HB_OK) goto done;
        memcpy(h_par, parent_results, sizeof(hb_me_result) * (size_t)n_parent);
        if ((crc = hbc_h2d_async(d_par, h_par, sizeof(hb_me_result) * (size_t)n_parent, ctx->stream))) goto done;
    }
    /* group by size, keep the caller's index in `out` */
    hbd_me_job *hj = (hbd_me_job *)h_jobs;
    int n = 0, start_of[5];
    for (int s = 0; s < 4; s++) {
        start_of[s] = n;
        for (int i = 0; i < n_jobs; i++) {
            const hb_me_job *j = &jobs[i];
            if (j->size != sizes[s]) continue;
            hbd_me_job *d = &hj[n++];
            memset(d, 0, sizeof *d);
            d->x = j->x; d->y = j->y; d->n_amvp = j->n_amvp; d->n_start = j->n_start;
            for (int k = 0; k < 2; k++) { d->amvp[2 * k] = j->amvp[k].x; d->amvp[2 * k + 1] = j->amvp[k].y; }
            for (int k = 0; k < 3; k++) { d->start[2 * k] = j->start[k].x; d->start[2 * k + 1] = j->start[k].y; }
            d->parent = (d_par && j->parent >= 0) ? j->parent : -1;
            d->out = i;
            d->corr = mv_cost_weight(j->qp, avg_dist);
        }
    }
    start_of[4] = n;
    if ((crc = hbc_h2d_async(d_jobs, h_jobs, sizeof(hbd_me_job) * (size_t)n_jobs, ctx->stream))) goto done;
    if (units) {
        const size_t plane = (size_t)units_w * (size_t)((cur->h + 63) / 64) * 16;
        if ((rc = hbi_scratch(ctx, 3, sizeof(hb_unit_info) * plane, &d_units, &h_units)) != HB_OK) goto done;
        memcpy(h_units, units, sizeof(hb_unit_info) * plane);
        if ((crc = hbc_h2d_async(d_units, h_units, sizeof(hb_unit_info) * plane, ctx->stream))) goto done;
    }
    for (int s = 0; s < 4 && !crc; s++) {
        const int cnt = start_of[s + 1] - start_of[s];
        if (!cnt) continue;
        if (units) {
            crc = hbk_amvp_fill((const hb_unit_info *)d_units, units_w, cur->w, cur->h, (hbd_me_job *)d_jobs + start_of[s], cnt, sizes[s], ctx->stream);
            ctx->launches++;
            if (crc) break;
        }
        crc = hbk_me_search(&cur->d, &ref->d, sizes[s], (const hbd_me_job *)d_jobs + start_of[s], cnt, (const hb_me_result *)d_par,
                            (hb_me_result *)d_res, action, NULL, NULL, NULL, 0, ctx->stream);
        ctx->launches++;
    }
    if (!crc) crc = hbc_d2h_async(h_res, d_res, sizeof(hb_me_result) * (size_t)n_jobs, ctx->stream);
    if (!crc) crc = hbc_stream_sync(ctx->stream);
    if (!crc) memcpy(results, h_res, sizeof(hb_me_result) * (size_t)n_jobs);
done:
    pthread_mutex_unlock(&ctx->lock);
    if (crc) return hbi_cuda_fail(crc, "hb_me_search");
    return rc;
}

/* validate, pack and launch the motion compensation of `jobs`; the caller holds ctx->lock and waits for the stream */
int hbi_mc_predict_queue(hb_ctx *ctx, const hb_frame *ref, hb_frame *pred, const hb_mc_job *jobs, int n_jobs, const char *what)
{
    static const int sizes[4] = { 64, 32, 16, 8 };
    int rc = HB_OK, crc = 0;
    void *d_pus, *h_pus, *d_mv, *h_mv;
    const int reach = HB_PAD_LUMA - 16;                    /* how far outside the picture a predicted block may reach */
    for (int i = 0; i < n_jobs; i++) {
        const hb_mc_job *j = &jobs[i];
        const int x0 = j->x + (j->mv.x >> 2), y0 = j->y + (j->mv.y >> 2);
        if ((j->size != 8 && j->size != 16 && j->size != 32 && j->size != 64) || j->x < 0 || j->y < 0 || ((j->x | j->y) & 7) || j->x + j->size > ref->w ||
            j->y + j->size > ref->h || x0 < -reach || y0 < -reach || x0 + j->size > ref->w + reach || y0 + j->size > ref->h + reach)
            return hbi_fail(HB_ERR_ARG, "%s: job %d is invalid or points further than %d samples outside the picture", what, i, reach);
    }
    if ((rc = hbi_scratch(ctx, 0, sizeof(hbd_mc_pu) * (size_t)n_jobs, &d_pus, &h_pus)) != HB_OK) return rc;
    if ((rc = hbi_scratch(ctx, 1, sizeof(hb_me_result) * (size_t)n_jobs, &d_mv, &h_mv)) != HB_OK) return rc;
    hbd_mc_pu *hp = (hbd_mc_pu *)h_pus;
    hb_me_result *hm = (hb_me_result *)h_mv;
    int n = 0, start_of[5];
    memset(hm, 0, sizeof(hb_me_result) * (size_t)n_jobs);
    for (int s = 0; s < 4; s++) {
        start_of[s] = n;
        for (int i = 0; i < n_jobs; i++) {
            if (jobs[i].size != sizes[s]) continue;
            hp[n].x = jobs[i].x; hp[n].y = jobs[i].y; hp[n].mv_idx = i;
            hm[i].mv = jobs[i].mv;
            n++;
        }
    }
    start_of[4] = n;
    crc = hbc_h2d_async(d_pus, h_pus, sizeof(hbd_mc_pu) * (size_t)n_jobs, ctx->stream);
    if (!crc) crc = hbc_h2d_async(d_mv, h_mv, sizeof(hb_me_result) * (size_t)n_jobs, ctx->stream);
    for (int s = 0; s < 4 && !crc; s++) {
        const int cnt = start_of[s + 1] - start_of[s];
        if (!cnt) continue;
        crc = hbk_mc_predict(&ref->d, &pred->d, sizes[s], (const hbd_mc_pu *)d_pus + start_of[s], cnt, (const hb_me_result *)d_mv, 3, ctx->stream);
        ctx->launches += 2;                                /* one luma, one chroma kernel */
    }
    return crc ? hbi_cuda_fail(crc, what) : HB_OK;
}

int hb_mc_predict(hb_ctx *ctx, const hb_frame *ref, hb_frame *pred, const hb_mc_job *jobs, int n_jobs)
{
    int rc, crc;
    if (!ctx || !ref || !pred || !jobs || n_jobs < 0) return hbi_fail(HB_ERR_ARG, "hb_mc_predict: bad argument");
    if (pred->w != ref->w || pred->h != ref->h) return hbi_fail(HB_ERR_ARG, "hb_mc_predict: frame sizes differ");
    if (n_jobs == 0) return HB_OK;
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    rc = hbi_mc_predict_queue(ctx, ref, pred, jobs, n_jobs, "hb_mc_predict");
    crc = rc == HB_OK ? hbc_stream_sync(ctx->stream) : 0;
    pthread_mutex_unlock(&ctx->lock);
    if (crc) return hbi_cuda_fail(crc, "hb_mc_predict");
    return rc;
}

/* bi-prediction (B slices): both lists' hmr_motion_compensation_luma/_chroma with is_bi_predict = 1 and the weighted_average_motion
 * of the two 14-bit predictions (predict_inter, hmr_motion_inter.c:3047-3056), fused per block */
int hb_mc_predict_bi(hb_ctx *ctx, const hb_frame *ref0, const hb_frame *ref1, hb_frame *pred, const hb_mc_bi_job *jobs, int n_jobs)
{
    static const int sizes[4] = { 64, 32, 16, 8 };
    int rc = HB_OK, crc = 0;
    void *d_pus, *h_pus, *d_mv, *h_mv;
    if (!ctx || !ref0 || !ref1 || !pred || !jobs || n_jobs < 0) return hbi_fail(HB_ERR_ARG, "hb_mc_predict_bi: bad argument");
    if (pred->w != ref0->w || pred->h != ref0->h || ref1->w != ref0->w || ref1->h != ref0->h) return hbi_fail(HB_ERR_ARG, "hb_mc_predict_bi: frame sizes differ");
    if (n_jobs == 0) return HB_OK;
    const int reach = HB_PAD_LUMA - 16;
    for (int i = 0; i < n_jobs; i++) {
        const hb_mc_bi_job *j = &jobs[i];
        int bad = (j->size != 8 && j->size != 16 && j->size != 32 && j->size != 64) || j->x < 0 || j->y < 0 || ((j->x | j->y) & 7) ||
                  j->x + j->size > ref0->w || j->y + j->size > ref0->h;
        for (int l = 0; l < 2 && !bad; l++) {
            const hb_mv mv = l ? j->mv1 : j->mv0;
            const int x0 = j->x + (mv.x >> 2), y0 = j->y + (mv.y >> 2);
            bad = x0 < -reach || y0 < -reach || x0 + j->size > ref0->w + reach || y0 + j->size > ref0->h + reach;
        }
        if (bad) return hbi_fail(HB_ERR_ARG, "hb_mc_predict_bi: job %d is invalid or points further than %d samples outside the picture", i, reach);
    }
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    if ((rc = hbi_scratch(ctx, 0, sizeof(hbd_mc_pu) * (size_t)n_jobs, &d_pus, &h_pus)) != HB_OK) goto done;
    if ((rc = hbi_scratch(ctx, 1, sizeof(hb_me_result) * 2 * (size_t)n_jobs, &d_mv, &h_mv)) != HB_OK) goto done;
    hbd_mc_pu *hp = (hbd_mc_pu *)h_pus;
    hb_me_result *hm = (hb_me_result *)h_mv;                /* list 0 vectors, then list 1 vectors */
    int n = 0, start_of[5];
    memset(hm, 0, sizeof(hb_me_result) * 2 * (size_t)n_jobs);
    for (int s = 0; s < 4; s++) {
        start_of[s] = n;
        for (int i = 0; i < n_jobs; i++) {
            if (jobs[i].size != sizes[s]) continue;
            hp[n].x = jobs[i].x; hp[n].y = jobs[i].y; hp[n].mv_idx = i;
            hm[i].mv = jobs[i].mv0; hm[n_jobs + i].mv = jobs[i].mv1;
            n++;
        }
    }
    start_of[4] = n;
    crc = hbc_h2d_async(d_pus, h_pus, sizeof(hbd_mc_pu) * (size_t)n_jobs, ctx->stream);
    if (!crc) crc = hbc_h2d_async(d_mv, h_mv, sizeof(hb_me_result) * 2 * (size_t)n_jobs, ctx->stream);
    for (int s = 0; s < 4 && !crc; s++) {
        const int cnt = start_of[s + 1] - start_of[s];
        if (!cnt) continue;
        crc = hbk_mc_predict_bi(&ref0->d, &ref1->d, &pred->d, sizes[s], (const hbd_mc_pu *)d_pus + start_of[s], cnt,
                                (const hb_me_result *)d_mv, (const hb_me_result *)d_mv + n_jobs, ctx->stream);
        ctx->launches += 2;
    }
    if (!crc) crc = hbc_stream_sync(ctx->stream);
done:
    pthread_mutex_unlock(&ctx->lock);
    if (crc) return hbi_cuda_fail(crc, "hb_mc_predict_bi");
    return rc;
}

/* merge / skip candidates: motion compensation, the no-residual block SSDs and the inter T/Q chain of every candidate's transform units
 * are QUEUED back to back on the context's stream (the same kernels hb_mc_predict / hb_tq_encode launch) and waited for ONCE (round 1:
 * three blocking calls); the reduction per candidate runs on the host afterwards */
int hb_merge_eval(hb_ctx *ctx, const hb_frame *cur, const hb_frame *ref, hb_frame *pred, hb_frame *recon, const hb_mc_job *cands, int n_cands,
                  int qp, int chroma_qp_offset, const hb_tq_params *params, hb_merge_result *out)
{
    static const int sizes[4] = { 64, 32, 16, 8 };
    int rc = HB_OK, crc = 0;
    if (!ctx || !cur || !ref || !pred || !recon || !cands || !params || !out || n_cands < 0) return hbi_fail(HB_ERR_ARG, "hb_merge_eval: bad argument");
    if (qp < 0 || qp > 51) return hbi_fail(HB_ERR_ARG, "hb_merge_eval: qp %d", qp);
    if (pred->w != ref->w || pred->h != ref->h || cur->w != ref->w || cur->h != ref->h) return hbi_fail(HB_ERR_ARG, "hb_merge_eval: frame sizes differ");
    if (n_cands == 0) return HB_OK;
    for (int i = 0; i < n_cands; i++)        /* coding units sit on their own grid; their transform units inherit the alignment the T/Q kernels need */
        if (cands[i].size > 0 && ((cands[i].x | cands[i].y) & (cands[i].size - 1)))
            return hbi_fail(HB_ERR_ARG, "hb_merge_eval: candidate %d is not aligned to its size", i);
    const int qp_c = hbi_chroma_qp(qp, chroma_qp_offset);
    hb_tq_params prm = *params;
    prm.chroma_weight = pow(2.0, (qp - qp_c) / 3.0);                                        /* hmr_motion_inter.c:3525 */
    /* transform units: luma size (64x64: four 32x32), chroma half of that (encode_inter walks a 64x64 unit as four 32x32 ones, :3094) */
    size_t n_tu = 0, n_co = 0;
    for (int i = 0; i < n_cands; i++) {
        const int s = cands[i].size, nl = s == 64 ? 4 : 1, ls = s == 64 ? 32 : s;
        n_tu += 3 * (size_t)nl; n_co += (size_t)nl * (ls * ls + 2 * (ls / 2) * (ls / 2));
    }
    hb_tu_job *tj = (hb_tu_job *)malloc(sizeof *tj * (n_tu ? n_tu : 1));
    hb_tu_result *tr = (hb_tu_result *)malloc(sizeof *tr * (n_tu ? n_tu : 1));
    int16_t *co = (int16_t *)malloc(sizeof *co * (n_co ? n_co : 1));
    int *order = (int *)malloc(sizeof(int) * (size_t)n_cands);
    uint32_t *ssd3 = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (size_t)n_cands);
    if (!tj || !tr || !co || !order || !ssd3) { rc = hbi_fail(HB_ERR_NOMEM, "hb_merge_eval: out of memory"); goto out; }
    size_t k = 0;
    for (int i = 0; i < n_cands; i++) {
        const int s = cands[i].size, ls = s == 64 ? 32 : s;
        for (int c = 0; c < 3; c++) {
            const int ts = c ? ls / 2 : ls, bx = c ? cands[i].x / 2 : cands[i].x, by = c ? cands[i].y / 2 : cands[i].y, bs = c ? s / 2 : s;
            for (int yy = 0; yy < bs; yy += ts)
                for (int xx = 0; xx < bs; xx += ts) { tj[k].comp = c; tj[k].x = bx + xx; tj[k].y = by + yy; tj[k].size = ts; tj[k].qp = c ? qp_c : qp; k++; }
        }
    }
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    {
        hbi_tq_pack pk;
        void *d_pus, *h_pus, *d_out, *h_out;
        int have_pk = 0;
        rc = hbi_mc_predict_queue(ctx, ref, pred, cands, n_cands, "hb_merge_eval");          /* validates the candidates; scratch 0, 1 */
        /* no-residual distortion: ssd16b(orig, pred) of the whole block per plane (scratch 3, 4: the T/Q queue below takes 0..2) */
        if (rc == HB_OK) rc = hbi_scratch(ctx, 3, sizeof(hbd_mc_pu) * (size_t)n_cands, &d_pus, &h_pus);
        if (rc == HB_OK) rc = hbi_scratch(ctx, 4, sizeof(uint32_t) * 3 * (size_t)n_cands, &d_out, &h_out);
        if (rc == HB_OK) {
            hbd_mc_pu *hp = (hbd_mc_pu *)h_pus;
            int n = 0, start_of[5];
            for (int s = 0; s < 4; s++) {
                start_of[s] = n;
                for (int i = 0; i < n_cands; i++) if (cands[i].size == sizes[s]) { hp[n].x = cands[i].x; hp[n].y = cands[i].y; hp[n].mv_idx = i; order[n] = i; n++; }
            }
            start_of[4] = n;
            crc = hbc_h2d_async(d_pus, h_pus, sizeof(hbd_mc_pu) * (size_t)n_cands, ctx->stream);
            for (int s = 0; s < 4 && !crc; s++)
                if (start_of[s + 1] > start_of[s]) {
                    crc = hbk_block_ssd(&cur->d, &pred->d, (const hbd_mc_pu *)d_pus + start_of[s], start_of[s + 1] - start_of[s], sizes[s],
                                        (uint32_t *)d_out + 3 * start_of[s], ctx->stream);
                    ctx->launches++;
                }
            if (!crc) crc = hbc_d2h_async(h_out, d_out, sizeof(uint32_t) * 3 * (size_t)n_cands, ctx->stream);
            if (crc) rc = hbi_cuda_fail(crc, "hb_merge_eval");
        }
        /* the kernels queued so far still read scratch 0, 1, 3, 4: the T/Q queue takes 5..7, and ONE wait ends the call */
        if (rc == HB_OK) { rc = hbi_tq_encode_queue_at(ctx, cur, pred, recon, tj, (int)n_tu, &prm, &pk, "hb_merge_eval", 5); have_pk = rc == HB_OK; }
        if (rc == HB_OK) {
            crc = hbc_stream_sync(ctx->stream);
            if (crc) rc = hbi_cuda_fail(crc, "hb_merge_eval");
            else {
                hbi_tq_collect(&pk, tj, (int)n_tu, co, tr);
                for (int q = 0; q < n_cands; q++) memcpy(ssd3 + 3 * order[q], (uint32_t *)h_out + 3 * q, sizeof(uint32_t) * 3);
            }
        }
        if (have_pk) hbi_tq_pack_free(&pk);
    }
    pthread_mutex_unlock(&ctx->lock);
    if (rc == HB_OK) {
        k = 0;
        for (int i = 0; i < n_cands; i++) {
            const int n = 3 * (cands[i].size == 64 ? 4 : 1);
            hb_merge_result r = { 0, 0, 0, 0 };
            for (int t = 0; t < n; t++, k++) {
                r.dist_coded += tr[k].ssd; r.sum += tr[k].sum;
                if (tr[k].sum > 0) r.cbf |= 1 << tj[k].comp;
            }
            r.dist_skip = ssd3[3 * i] + (uint32_t)(prm.chroma_weight * ssd3[3 * i + 1]) + (uint32_t)(prm.chroma_weight * ssd3[3 * i + 2]);
            out[i] = r;
        }
    }
out:
    free(ssd3); free(order); free(tj); free(tr); free(co);
    return rc;
}

/* deblocking of a whole picture in place (pixel stage; strengths and QPs from the host) */
int hb_deblock_frame(hb_ctx *ctx, hb_frame *frame, const uint8_t *bs_ver, const uint8_t *bs_hor, const uint8_t *qp, int units_w,
                     const hb_deblock_params *params)
{
    int rc = HB_OK, crc = 0;
    void *d_maps, *h_maps;
    if (!ctx || !frame || !bs_ver || !bs_hor || !qp || !params) return hbi_fail(HB_ERR_ARG, "hb_deblock_frame: NULL argument");
    if (units_w < frame->w / 4) return hbi_fail(HB_ERR_ARG, "hb_deblock_frame: units_w %d is smaller than width/4", units_w);
    const size_t plane = (size_t)units_w * (size_t)(frame->h / 4);
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    if ((rc = hbi_scratch(ctx, 0, 3 * plane, &d_maps, &h_maps)) != HB_OK) goto done;
    memcpy(h_maps, bs_ver, plane); memcpy((char *)h_maps + plane, bs_hor, plane); memcpy((char *)h_maps + 2 * plane, qp, plane);
    crc = hbc_h2d_async(d_maps, h_maps, 3 * plane, ctx->stream);
    if (!crc) {
        crc = hbk_deblock(&frame->d, (const uint8_t *)d_maps, (const uint8_t *)d_maps + plane, (const uint8_t *)d_maps + 2 * plane, units_w,
                          params->cb_qp_offset, params->cr_qp_offset, params->beta_offset_div2, params->tc_offset_div2, ctx->stream);
        ctx->launches += 2;
    }
    if (!crc) { crc = hbk_pad_frame(&frame->d, ctx->stream); ctx->launches++; }
    if (!crc) crc = hbc_stream_sync(ctx->stream);
done:
    pthread_mutex_unlock(&ctx->lock);
    if (crc) return hbi_cuda_fail(crc, "hb_deblock_frame");
    return rc;
}

/* deblocking of a whole picture in place, strengths derived on the device from per-unit mode data */
static int deblock_units(hb_ctx *ctx, hb_frame *frame, const hb_unit_info *units, const hb_unit_l1 *units1, int units_w, const int32_t *pic_l0, int n_l0,
                         const int32_t *pic_l1, int n_l1, const hb_deblock_params *params, uint8_t *bs_ver_out, uint8_t *bs_hor_out, const char *what)
{
    int rc = HB_OK, crc = 0;
    void *d_units, *h_units, *d_maps, *h_maps, *d_u1 = NULL, *h_u1 = NULL;
    if (!ctx || !frame || !units || !params) return hbi_fail(HB_ERR_ARG, "%s: NULL argument", what);
    if (units_w < frame->w / 4) return hbi_fail(HB_ERR_ARG, "%s: units_w %d is smaller than width/4", what, units_w);
    if (units1 && (!pic_l0 || !pic_l1 || n_l0 < 0 || n_l0 > 16 || n_l1 < 0 || n_l1 > 16)) return hbi_fail(HB_ERR_ARG, "%s: bad reference picture lists", what);
    const size_t plane = (size_t)units_w * (size_t)(frame->h / 4);
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    if ((rc = hbi_scratch(ctx, 0, sizeof(hb_unit_info) * plane, &d_units, &h_units)) != HB_OK) goto done;
    if ((rc = hbi_scratch(ctx, 1, 3 * plane, &d_maps, &h_maps)) != HB_OK) goto done;
    if (units1 && (rc = hbi_scratch(ctx, 2, sizeof(hb_unit_l1) * plane, &d_u1, &h_u1)) != HB_OK) goto done;
    memcpy(h_units, units, sizeof(hb_unit_info) * plane);
    crc = hbc_h2d_async(d_units, h_units, sizeof(hb_unit_info) * plane, ctx->stream);
    if (!crc && units1) { memcpy(h_u1, units1, sizeof(hb_unit_l1) * plane); crc = hbc_h2d_async(d_u1, h_u1, sizeof(hb_unit_l1) * plane, ctx->stream); }
    if (!crc) crc = hbc_memset_async(d_maps, 0, 3 * plane, ctx->stream);
    uint8_t *bv = (uint8_t *)d_maps, *bh = bv + plane, *qp = bh + plane;
    if (!crc) {
        if (units1) crc = hbk_deblock_strengths_b((const hb_unit_info *)d_units, (const hb_unit_l1 *)d_u1, pic_l0, n_l0, pic_l1, n_l1, units_w, frame->w, frame->h, bv, bh, qp, ctx->stream);
        else crc = hbk_deblock_strengths((const hb_unit_info *)d_units, units_w, frame->w, frame->h, bv, bh, qp, ctx->stream);
        ctx->launches++;
    }
    if (!crc) {
        crc = hbk_deblock(&frame->d, bv, bh, qp, units_w, params->cb_qp_offset, params->cr_qp_offset, params->beta_offset_div2, params->tc_offset_div2, ctx->stream);
        ctx->launches += 2;
    }
    if (!crc) { crc = hbk_pad_frame(&frame->d, ctx->stream); ctx->launches++; }
    if (!crc && (bs_ver_out || bs_hor_out)) crc = hbc_d2h_async(h_maps, d_maps, 2 * plane, ctx->stream);
    if (!crc) crc = hbc_stream_sync(ctx->stream);
    if (!crc && bs_ver_out) memcpy(bs_ver_out, h_maps, plane);
    if (!crc && bs_hor_out) memcpy(bs_hor_out, (char *)h_maps + plane, plane);
done:
    pthread_mutex_unlock(&ctx->lock);
    if (crc) return hbi_cuda_fail(crc, what);
    return rc;
}
int hb_deblock_frame_units(hb_ctx *ctx, hb_frame *frame, const hb_unit_info *units, int units_w, const hb_deblock_params *params,
                           uint8_t *bs_ver_out, uint8_t *bs_hor_out)
{
    return deblock_units(ctx, frame, units, NULL, units_w, NULL, 0, NULL, 0, params, bs_ver_out, bs_hor_out, "hb_deblock_frame_units");
}
int hb_deblock_frame_units_b(hb_ctx *ctx, hb_frame *frame, const hb_unit_info *units, const hb_unit_l1 *units_l1, int units_w,
                             const int32_t *pic_l0, int n_l0, const int32_t *pic_l1, int n_l1, const hb_deblock_params *params,
                             uint8_t *bs_ver_out, uint8_t *bs_hor_out)
{
    if (!units_l1) return hbi_fail(HB_ERR_ARG, "hb_deblock_frame_units_b: NULL list-1 units");
    return deblock_units(ctx, frame, units, units_l1, units_w, pic_l0, n_l0, pic_l1, n_l1, params, bs_ver_out, bs_hor_out, "hb_deblock_frame_units_b");
}

/* ---- wavefront-batched reconstruction of intra transform units (include/homer_b200.h: hb_intra_reconstruct) */
static int tq_row_align(int size);
/* neighbour flags of the quadtree node of `size` luma samples at luma (x, y) -- the CTU's own (hmr_motion_intra.c:676-683; its left-bottom
 * neighbour never exists) handed down by cu_partition_get_neighbours (:625-657) -- and the lengths of its left-bottom / top-right runs inside
 * the picture, in samples of the plane the unit is coded in (:291, :339).  flags: bit 0 left, 1 top, 2 left-bottom, 3 top-right */
static void intra_neighbours(int w, int h, int x, int y, int size, int chroma, int *flags, int *lbs_out, int *trs_out)
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
    if (lbs < 0) lbs = 0;
    if (trs < 0) trs = 0;
    *flags = (l ? 1 : 0) | (t ? 2 : 0) | (lb ? 4 : 0) | (tr ? 8 : 0); *lbs_out = lbs; *trs_out = trs;
}

typedef struct intra_sorted { int level, comp, size, qp, scan, idx; } intra_sorted;
static int intra_sorted_cmp(const void *pa, const void *pb)
{
    const intra_sorted *a = (const intra_sorted *)pa, *b = (const intra_sorted *)pb;
    if (a->level != b->level) return a->level - b->level;
    if (a->comp != b->comp) return a->comp - b->comp;
    if (a->size != b->size) return a->size - b->size;
    if (a->qp != b->qp) return a->qp - b->qp;
    if (a->scan != b->scan) return a->scan - b->scan;
    return a->idx - b->idx;
}

int hb_intra_reconstruct(hb_ctx *ctx, const hb_frame *cur, hb_frame *pred, hb_frame *recon, const hb_intra_unit *units, int n_units,
                         int is_islice, int sign_hiding, double chroma_weight, int16_t *coeffs, hb_tu_result *results, int32_t *n_levels_out)
{
    return hb_intra_reconstruct_ex(ctx, cur, pred, recon, units, n_units, is_islice, sign_hiding, chroma_weight, 0, coeffs, results, n_levels_out);
}
int hb_intra_reconstruct_ex(hb_ctx *ctx, const hb_frame *cur, hb_frame *pred, hb_frame *recon, const hb_intra_unit *units, int n_units,
                            int is_islice, int sign_hiding, double chroma_weight, int flags, int16_t *coeffs, hb_tu_result *results, int32_t *n_levels_out)
{
    int rc = HB_OK, crc = 0;
    if (!ctx || !cur || !pred || !recon || !units || !coeffs || !results || n_units < 0) return hbi_fail(HB_ERR_ARG, "hb_intra_reconstruct: bad argument");
    if (cur->w != recon->w || cur->h != recon->h || pred->w != cur->w || pred->h != cur->h) return hbi_fail(HB_ERR_ARG, "hb_intra_reconstruct: picture sizes differ");
    if (n_levels_out) *n_levels_out = 0;
    if (n_units == 0) return HB_OK;
    const int w = cur->w, h = cur->h;
    size_t total = 0, n_adi = 0;
    for (int i = 0; i < n_units; i++) {
        const hb_intra_unit *u = &units[i];
        const int lum = u->comp ? 2 : 1;
        if (u->comp < 0 || u->comp > 2 || (u->size != 4 && u->size != 8 && u->size != 16 && u->size != 32) || (u->comp && u->size == 32) || u->qp < 0 || u->qp > 51 ||
            u->mode < 0 || u->mode > 34 || u->x < 0 || u->y < 0 || (u->x & (tq_row_align(u->size) - 1)) || (u->y & 3) ||
            u->x + u->size > cur->d.p[u->comp].w || u->y + u->size > cur->d.p[u->comp].h ||
            u->scan_mode < HB_SCAN_HOR || u->scan_mode > HB_SCAN_DIAG || (u->size > 8 && u->scan_mode != HB_SCAN_DIAG) ||
            (u->node_size != 4 && u->node_size != 8 && u->node_size != 16 && u->node_size != 32 && u->node_size != 64) ||
            u->node_x < 0 || u->node_y < 0 || (u->node_x % u->node_size) || (u->node_y % u->node_size) || u->node_x >= w || u->node_y >= h ||
            u->node_x != u->x * lum || u->node_y != u->y * lum || u->node_size < u->size * lum)
            return hbi_fail(HB_ERR_ARG, "hb_intra_reconstruct: unit %d is invalid", i);
        total += (size_t)u->size * u->size; n_adi += 4 * (size_t)u->size + 1;
    }
    /* ---- levels of the dependency: per plane a map of the level that produced every 4x4 block of samples */
    intra_sorted *srt = (intra_sorted *)malloc(sizeof *srt * (size_t)n_units);
    hbd_adi_job *aj = (hbd_adi_job *)malloc(sizeof *aj * (size_t)n_units);
    size_t *coeff_off = (size_t *)malloc(sizeof(size_t) * (size_t)n_units);
    int32_t *map[3] = { NULL, NULL, NULL };
    int mw[3], mh[3], n_levels = 0;
    for (int c = 0; c < 3; c++) {
        mw[c] = (c ? w / 2 : w) / 4; mh[c] = (c ? h / 2 : h) / 4;
        map[c] = (int32_t *)calloc((size_t)mw[c] * mh[c], sizeof(int32_t));       /* 0: nothing of this call wrote here */
    }
    if (!srt || !aj || !coeff_off || !map[0] || !map[1] || !map[2]) { rc = hbi_fail(HB_ERR_NOMEM, "hb_intra_reconstruct: out of memory"); goto cleanup; }
    {
        size_t co = 0, ao = 0;
        for (int i = 0; i < n_units; i++) {
            const hb_intra_unit *u = &units[i];
            const int c = u->comp, n = u->size, bx = u->x / 4, by = u->y / 4, nb = n / 4;
            int nflags, lbs, trs, level = 0;
            intra_neighbours(w, h, u->node_x, u->node_y, u->node_size, c != 0, &nflags, &lbs, &trs);
            #define AT(xx, yy) (((xx) >= 0 && (yy) >= 0 && (xx) < mw[c] && (yy) < mh[c]) ? map[c][(size_t)(yy) * mw[c] + (xx)] : 0)
            #define UP(v) do { const int v_ = (v); if (v_ > level) level = v_; } while (0)
            if (nflags & 1) for (int k = 0; k < nb; k++) UP(AT(bx - 1, by + k));
            if (nflags & 4) for (int k = 0; k < (lbs + 3) / 4; k++) UP(AT(bx - 1, by + nb + k));
            if (nflags & 2) for (int k = 0; k < nb; k++) UP(AT(bx + k, by - 1));
            if (nflags & 8) for (int k = 0; k < (trs + 3) / 4; k++) UP(AT(bx + nb + k, by - 1));
            if ((nflags & 3) == 3) UP(AT(bx - 1, by - 1));
            #undef UP
            #undef AT
            level += 1;
            for (int yy = by; yy < by + nb; yy++) for (int xx = bx; xx < bx + nb; xx++) map[c][(size_t)yy * mw[c] + xx] = level;
            if (level > n_levels) n_levels = level;
            srt[i].level = level; srt[i].comp = c; srt[i].size = n; srt[i].qp = u->qp; srt[i].scan = u->scan_mode; srt[i].idx = i;
            coeff_off[i] = co; co += (size_t)n * n;
            aj[i].comp = c; aj[i].x = u->x; aj[i].y = u->y; aj[i].n = n; aj[i].flags = nflags; aj[i].lbs = lbs; aj[i].trs = trs; aj[i].adi_off = (int32_t)ao;
            ao += 4 * (size_t)n + 1;
        }
    }
    qsort(srt, (size_t)n_units, sizeof *srt, intra_sorted_cmp);
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    {
        void *d_a, *h_a, *d_j, *h_j, *d_xy, *h_xy, *d_adi, *h_adi, *d_co, *h_co, *d_rs, *h_rs;
        if ((rc = hbi_scratch(ctx, 0, sizeof(hbd_adi_job) * (size_t)n_units, &d_a, &h_a)) != HB_OK) goto unlock;
        if ((rc = hbi_scratch(ctx, 1, sizeof(hbd_intra_job) * (size_t)n_units, &d_j, &h_j)) != HB_OK) goto unlock;
        if ((rc = hbi_scratch(ctx, 2, sizeof(int32_t) * 2 * (size_t)n_units, &d_xy, &h_xy)) != HB_OK) goto unlock;
        if ((rc = hbi_scratch(ctx, 3, sizeof(int16_t) * n_adi, &d_adi, &h_adi)) != HB_OK) goto unlock;
        if ((rc = hbi_scratch(ctx, 4, sizeof(int16_t) * total, &d_co, &h_co)) != HB_OK) goto unlock;
        if ((rc = hbi_scratch(ctx, 5, sizeof(hb_tu_result) * (size_t)n_units, &d_rs, &h_rs)) != HB_OK) goto unlock;
        /* records in launch order (level, plane, size, qp, scan) */
        size_t packed_coeff = 0;
        size_t *launch_coeff = (size_t *)malloc(sizeof(size_t) * (size_t)n_units);
        if (!launch_coeff) { rc = hbi_fail(HB_ERR_NOMEM, "hb_intra_reconstruct: out of memory"); goto unlock; }
        for (int k = 0; k < n_units; k++) {
            const int i = srt[k].idx;
            const hb_intra_unit *u = &units[i];
            ((hbd_adi_job *)h_a)[k] = aj[i];
            hbd_intra_job *pj = &((hbd_intra_job *)h_j)[k];
            pj->comp = u->comp; pj->x = u->x; pj->y = u->y; pj->size = u->size; pj->mode = u->mode;
            pj->filtered = u->comp ? 0 : -1;             /* luma: the reference's rule (:1011); chroma: never (hmr_motion_intra_chroma.c:337) */
            pj->adi_off = aj[i].adi_off; pj->pad_ = 0;
            ((int32_t *)h_xy)[2 * k] = u->x; ((int32_t *)h_xy)[2 * k + 1] = u->y;
            launch_coeff[k] = packed_coeff; packed_coeff += (size_t)u->size * u->size;
        }
        if (!(flags & HB_INTRA_PER_LEVEL_LAUNCHES)) {
            /* ---- one persistent launch: units, tasks (32 / size consecutive units of one group and level) and groups */
            void *d_w, *h_w, *d_t, *h_t;
            const size_t units_bytes = (sizeof(hbd_wave_unit) * (size_t)n_units + 15) & ~(size_t)15;
            const size_t tasks_cap = (size_t)n_units, groups_cap = (size_t)n_units;
            const size_t tasks_bytes = (sizeof(hbd_wave_task) * tasks_cap + 15) & ~(size_t)15;
            if ((rc = hbi_scratch(ctx, 0, units_bytes + 64, &d_w, &h_w)) != HB_OK) { free(launch_coeff); goto unlock; }
            if ((rc = hbi_scratch(ctx, 1, tasks_bytes + sizeof(hbd_wave_group) * groups_cap, &d_t, &h_t)) != HB_OK) { free(launch_coeff); goto unlock; }
            hbd_wave_unit *wu = (hbd_wave_unit *)h_w;
            hbd_wave_task *wt = (hbd_wave_task *)h_t;
            hbd_wave_group *wg = (hbd_wave_group *)((char *)h_t + tasks_bytes);
            const hbd_wave_group *d_groups = (const hbd_wave_group *)((char *)d_t + tasks_bytes);
            int n_tasks = 0, n_groups = 0, level_first = 0;
            for (int k = 0; k < n_units; k++) {
                const int i = srt[k].idx;
                wu[k].x = units[i].x; wu[k].y = units[i].y; wu[k].mode = units[i].mode; wu[k].flags = aj[i].flags; wu[k].lbs = aj[i].lbs; wu[k].trs = aj[i].trs;
            }
            for (int k0 = 0; k0 < n_units; ) {
                if (k0 == 0 || srt[k0].level != srt[k0 - 1].level) level_first = k0;
                int k1 = k0;
                while (k1 < n_units && srt[k1].level == srt[k0].level && srt[k1].comp == srt[k0].comp && srt[k1].size == srt[k0].size && srt[k1].qp == srt[k0].qp &&
                       srt[k1].scan == srt[k0].scan) k1++;
                /* the group's tables (one record per distinct plane / size / qp / scan) */
                const int comp = srt[k0].comp, size = srt[k0].size, qp = srt[k0].qp, scan = srt[k0].scan;
                hbd_tq_args a;
                memset(&a, 0, sizeof a);
                hbi_tq_setup(ctx, &a, comp, size, qp, is_islice, sign_hiding);
                int lg = 2;
                while ((1 << lg) < size) lg++;
                hbd_wave_group g;
                memset(&g, 0, sizeof g);
                g.comp = comp; g.qbits = a.qbits; g.add = a.add; g.per = a.per;
                g.qtab = ctx->d_q + hbi_tab_q_off(lg, comp, qp % 6);           /* intra lists: (is_intra ? 0 : 3) + comp */
                g.dqtab = ctx->d_dq + hbi_tab_q_off(lg, 0, qp % 6);            /* SSE4.2 inv_quant: is_intra -> list 0 (hmr_sse42_functions_quant.c:138) */
                g.scan = ctx->d_scan + hbi_tab_scan_off(scan, lg);
                g.weight = comp ? chroma_weight : 1.0;
                int gi = -1;
                for (int q = 0; q < n_groups; q++) if (!memcmp(&wg[q], &g, sizeof g)) { gi = q; break; }
                if (gi < 0) { gi = n_groups; wg[n_groups++] = g; }
                const int tpw = 32 / size;
                for (int f = k0; f < k1; f += tpw) {
                    hbd_wave_task *t = &wt[n_tasks++];
                    t->first_unit = f; t->n_units = k1 - f < tpw ? k1 - f : tpw; t->group = gi; t->size = size;
                    t->units_before = level_first; t->pad_ = 0; t->coeff_off = (int64_t)launch_coeff[f];
                }
                k0 = k1;
            }
            unsigned *d_cnt = (unsigned *)((char *)d_w + units_bytes);
            crc = hbc_h2d_async(d_w, h_w, units_bytes, ctx->stream);
            if (!crc) crc = hbc_memset_async(d_cnt, 0, 64, ctx->stream);
            if (!crc) crc = hbc_h2d_async(d_t, h_t, tasks_bytes + sizeof(hbd_wave_group) * (size_t)n_groups, ctx->stream);
            if (!crc) crc = hbc_h2d_async(d_xy, h_xy, sizeof(int32_t) * 2 * (size_t)n_units, ctx->stream);
            hbd_wave_args wa;
            memset(&wa, 0, sizeof wa);
            wa.cur = cur->d; wa.pred = pred->d; wa.rec = recon->d;
            wa.units = (const hbd_wave_unit *)d_w; wa.xy = (const int32_t *)d_xy; wa.tasks = (const hbd_wave_task *)d_t; wa.groups = d_groups;
            wa.n_tasks = n_tasks; wa.sign_hiding = sign_hiding; wa.coeff = (int16_t *)d_co; wa.res = (hb_tu_result *)d_rs; wa.counters = d_cnt;
            /* enough warps for the widest levels, few enough to leave SMs to the other pictures in flight */
            int ctas = (n_tasks + 15) / 16;
            ctas = ctas < 8 ? 8 : (ctas > 48 ? 48 : ctas);
            if (!crc) { crc = hbk_intra_wave(&wa, ctas, ctx->stream); ctx->launches++; }
            goto collect;
        }
        crc = hbc_h2d_async(d_a, h_a, sizeof(hbd_adi_job) * (size_t)n_units, ctx->stream);
        if (!crc) crc = hbc_h2d_async(d_j, h_j, sizeof(hbd_intra_job) * (size_t)n_units, ctx->stream);
        if (!crc) crc = hbc_h2d_async(d_xy, h_xy, sizeof(int32_t) * 2 * (size_t)n_units, ctx->stream);
        /* ---- level by level: reference samples from `recon`, predictions into `pred`, the T/Q chain of every (plane, size, qp, scan) group */
        for (int k0 = 0; k0 < n_units && !crc; ) {
            int k1 = k0;
            while (k1 < n_units && srt[k1].level == srt[k0].level) k1++;
            crc = hbk_intra_adi(&recon->d, (const hbd_adi_job *)d_a + k0, k1 - k0, (int16_t *)d_adi, ctx->stream);
            ctx->launches++;
            if (crc) break;
            hbd_intra_args ia;
            memset(&ia, 0, sizeof ia);
            ia.pred = pred->d; ia.jobs = (const hbd_intra_job *)d_j + k0; ia.n_jobs = k1 - k0; ia.adi = (const int16_t *)d_adi;
            crc = hbk_intra(&ia, ctx->stream);
            ctx->launches++;
            for (int g0 = k0; g0 < k1 && !crc; ) {
                int g1 = g0;
                while (g1 < k1 && srt[g1].comp == srt[g0].comp && srt[g1].size == srt[g0].size && srt[g1].qp == srt[g0].qp && srt[g1].scan == srt[g0].scan) g1++;
                const int comp = srt[g0].comp, size = srt[g0].size, qp = srt[g0].qp;
                hbd_tq_args a;
                memset(&a, 0, sizeof a);
                hbi_tq_setup(ctx, &a, comp, size, qp, is_islice, sign_hiding);
                int lg = 2;
                while ((1 << lg) < size) lg++;
                a.qtab = ctx->d_q + hbi_tab_q_off(lg, comp, qp % 6);          /* intra lists: (is_intra ? 0 : 3) + comp */
                a.dqtab = ctx->d_dq + hbi_tab_q_off(lg, 0, qp % 6);           /* SSE4.2 inv_quant: is_intra -> list 0 (hmr_sse42_functions_quant.c:138) */
                a.scan = ctx->d_scan + hbi_tab_scan_off(srt[g0].scan, lg);
                a.intra = 1;
                a.cur = cur->d.p[comp]; a.pred = pred->d.p[comp]; a.rec = recon->d.p[comp];
                a.jobs_xy = (const int32_t *)d_xy + 2 * g0; a.n_jobs = g1 - g0;
                a.weight = comp ? chroma_weight : 1.0;
                a.coeff_out = (int16_t *)d_co + launch_coeff[g0];
                a.res_out = (hb_tu_result *)d_rs + g0;
                crc = hbk_tq_encode(&a, ctx->stream);
                ctx->launches++;
                g0 = g1;
            }
            k0 = k1;
        }
collect:
        if (!crc) { crc = hbk_pad_frame(&recon->d, ctx->stream); ctx->launches++; }
        if (!crc) crc = hbc_d2h_async(h_co, d_co, sizeof(int16_t) * total, ctx->stream);
        if (!crc) crc = hbc_d2h_async(h_rs, d_rs, sizeof(hb_tu_result) * (size_t)n_units, ctx->stream);
        if (!crc) crc = hbc_stream_sync(ctx->stream);
        if (!crc)
            for (int k = 0; k < n_units; k++) {
                const int i = srt[k].idx;
                memcpy(coeffs + coeff_off[i], (int16_t *)h_co + launch_coeff[k], sizeof(int16_t) * (size_t)units[i].size * units[i].size);
                results[i] = ((hb_tu_result *)h_rs)[k];
            }
        free(launch_coeff);
    }
unlock:
    pthread_mutex_unlock(&ctx->lock);
    if (n_levels_out) *n_levels_out = n_levels;
cleanup:
    free(srt); free(aj); free(coeff_off);
    for (int c = 0; c < 3; c++) free(map[c]);
    if (crc) return hbi_cuda_fail(crc, "hb_intra_reconstruct");
    return rc;
}

/* AMVP / merge candidates of a batch of PUs from the per-unit motion field: one launch, one copy back.  max_cands = 0: AMVP */
static int neighbour_candidates(hb_ctx *ctx, const hb_unit_info *units, int units_w, int width, int height, const hb_amvp_job *jobs, int n_jobs, int max_cands,
                                void *out, const char *what)
{
    int rc = HB_OK, crc = 0;
    void *d_units, *h_units, *d_jobs, *h_jobs, *d_out, *h_out;
    if (!ctx || !units || !jobs || !out || n_jobs < 0) return hbi_fail(HB_ERR_ARG, "%s: bad argument", what);
    if (width < 16 || height < 16 || (width & 7) || (height & 7)) return hbi_fail(HB_ERR_ARG, "%s: %dx%d", what, width, height);
    const int cols = (width + 63) / 64, rows = (height + 63) / 64;
    if (units_w < cols * 16) return hbi_fail(HB_ERR_ARG, "%s: units_w %d does not cover %d whole CTUs", what, units_w, cols);
    for (int i = 0; i < n_jobs; i++) {
        const hb_amvp_job *j = &jobs[i];
        if ((j->size != 64 && j->size != 32 && j->size != 16 && j->size != 8) || j->x < 0 || j->y < 0 || (j->x % j->size) || (j->y % j->size) ||
            j->x + j->size > width || j->y + j->size > height)
            return hbi_fail(HB_ERR_ARG, "%s: job %d: %dx%d at (%d,%d)", what, i, j->size, j->size, j->x, j->y);
    }
    if (n_jobs == 0) return HB_OK;
    const size_t plane = (size_t)units_w * (size_t)rows * 16;
    const size_t out_bytes = (max_cands ? sizeof(hb_mv) * (size_t)max_cands : sizeof(hb_amvp_list)) * (size_t)n_jobs;
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    if ((rc = hbi_scratch(ctx, 0, sizeof(hb_unit_info) * plane, &d_units, &h_units)) != HB_OK) goto done;
    if ((rc = hbi_scratch(ctx, 1, sizeof(hb_amvp_job) * (size_t)n_jobs, &d_jobs, &h_jobs)) != HB_OK) goto done;
    if ((rc = hbi_scratch(ctx, 2, out_bytes, &d_out, &h_out)) != HB_OK) goto done;
    memcpy(h_units, units, sizeof(hb_unit_info) * plane);
    memcpy(h_jobs, jobs, sizeof(hb_amvp_job) * (size_t)n_jobs);
    crc = hbc_h2d_async(d_units, h_units, sizeof(hb_unit_info) * plane, ctx->stream);
    if (!crc) crc = hbc_h2d_async(d_jobs, h_jobs, sizeof(hb_amvp_job) * (size_t)n_jobs, ctx->stream);
    if (!crc) {
        if (max_cands) crc = hbk_merge_cands((const hb_unit_info *)d_units, units_w, width, height, (const hb_amvp_job *)d_jobs, n_jobs, max_cands, (hb_mv *)d_out, ctx->stream);
        else crc = hbk_amvp((const hb_unit_info *)d_units, units_w, width, height, (const hb_amvp_job *)d_jobs, n_jobs, (hb_amvp_list *)d_out, ctx->stream);
        ctx->launches++;
    }
    if (!crc) crc = hbc_d2h_async(h_out, d_out, out_bytes, ctx->stream);
    if (!crc) crc = hbc_stream_sync(ctx->stream);
    if (!crc) memcpy(out, h_out, out_bytes);
done:
    pthread_mutex_unlock(&ctx->lock);
    if (crc) return hbi_cuda_fail(crc, what);
    return rc;
}
int hb_amvp_candidates(hb_ctx *ctx, const hb_unit_info *units, int units_w, int width, int height, const hb_amvp_job *jobs, int n_jobs, hb_amvp_list *out)
{
    return neighbour_candidates(ctx, units, units_w, width, height, jobs, n_jobs, 0, out, "hb_amvp_candidates");
}
int hb_merge_candidates(hb_ctx *ctx, const hb_unit_info *units, int units_w, int width, int height, const hb_amvp_job *jobs, int n_jobs, int max_cands, hb_mv *out)
{
    if (max_cands < 1 || max_cands > 5) return hbi_fail(HB_ERR_ARG, "hb_merge_candidates: max_cands %d", max_cands);
    return neighbour_candidates(ctx, units, units_w, width, height, jobs, n_jobs, max_cands, out, "hb_merge_candidates");
}

/* SAO statistics of a whole picture: one launch, one copy back */
int hb_sao_stats_frame(hb_ctx *ctx, const hb_frame *orig, const hb_frame *rec, hb_sao_stats *out)
{
    int rc = HB_OK, crc = 0;
    void *d_out, *h_out;
    if (!ctx || !orig || !rec || !out) return hbi_fail(HB_ERR_ARG, "hb_sao_stats_frame: NULL argument");
    if (orig->w != rec->w || orig->h != rec->h) return hbi_fail(HB_ERR_ARG, "hb_sao_stats_frame: frame sizes differ");
    const int cols = (rec->w + 63) / 64, n_ctus = cols * ((rec->h + 63) / 64);
    const size_t bytes = sizeof(hb_sao_stats) * 3 * (size_t)n_ctus;
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    if ((rc = hbi_scratch(ctx, 0, bytes, &d_out, &h_out)) != HB_OK) goto done;
    crc = hbk_sao_stats(&orig->d, &rec->d, cols, n_ctus, (hb_sao_stats *)d_out, ctx->stream);
    ctx->launches++;
    if (!crc) crc = hbc_d2h_async(h_out, d_out, bytes, ctx->stream);
    if (!crc) crc = hbc_stream_sync(ctx->stream);
    if (!crc) memcpy(out, h_out, bytes);
done:
    pthread_mutex_unlock(&ctx->lock);
    if (crc) return hbi_cuda_fail(crc, "hb_sao_stats_frame");
    return rc;
}

/* SAO statistics + the per-type offsets / band positions / distortion estimates derived from them on the device */
int hb_sao_candidates_frame(hb_ctx *ctx, const hb_frame *orig, const hb_frame *rec, const double lambda[3], hb_sao_candidate *cand, hb_sao_stats *stats)
{
    int rc = HB_OK, crc = 0;
    void *d_st, *h_st, *d_cand, *h_cand;
    if (!ctx || !orig || !rec || !lambda || !cand) return hbi_fail(HB_ERR_ARG, "hb_sao_candidates_frame: NULL argument");
    if (orig->w != rec->w || orig->h != rec->h) return hbi_fail(HB_ERR_ARG, "hb_sao_candidates_frame: frame sizes differ");
    const int cols = (rec->w + 63) / 64, n_ctus = cols * ((rec->h + 63) / 64);
    const size_t st_bytes = sizeof(hb_sao_stats) * 3 * (size_t)n_ctus, cand_bytes = sizeof(hb_sao_candidate) * 15 * (size_t)n_ctus;
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    if ((rc = hbi_scratch(ctx, 0, st_bytes, &d_st, &h_st)) != HB_OK) goto done;
    if ((rc = hbi_scratch(ctx, 1, cand_bytes, &d_cand, &h_cand)) != HB_OK) goto done;
    crc = hbk_sao_stats(&orig->d, &rec->d, cols, n_ctus, (hb_sao_stats *)d_st, ctx->stream);
    ctx->launches++;
    if (!crc) { crc = hbk_sao_derive((const hb_sao_stats *)d_st, 3 * n_ctus, lambda, (hb_sao_candidate *)d_cand, ctx->stream); ctx->launches++; }
    if (!crc) crc = hbc_d2h_async(h_cand, d_cand, cand_bytes, ctx->stream);
    if (!crc && stats) crc = hbc_d2h_async(h_st, d_st, st_bytes, ctx->stream);
    if (!crc) crc = hbc_stream_sync(ctx->stream);
    if (!crc) memcpy(cand, h_cand, cand_bytes);
    if (!crc && stats) memcpy(stats, h_st, st_bytes);
done:
    pthread_mutex_unlock(&ctx->lock);
    if (crc) return hbi_cuda_fail(crc, "hb_sao_candidates_frame");
    return rc;
}

/* SAO offset pass of a whole picture; refreshes dst's replicated border afterwards (dst is the next reference picture) */
int hb_sao_apply_frame(hb_ctx *ctx, const hb_frame *src, hb_frame *dst, const hb_sao_param *params)
{
    int rc = HB_OK, crc = 0;
    void *d_prm, *h_prm;
    if (!ctx || !src || !dst || !params) return hbi_fail(HB_ERR_ARG, "hb_sao_apply_frame: NULL argument");
    if (src == dst) return hbi_fail(HB_ERR_ARG, "hb_sao_apply_frame: dst must not be src (classes are taken from the untouched picture)");
    if (src->w != dst->w || src->h != dst->h) return hbi_fail(HB_ERR_ARG, "hb_sao_apply_frame: frame sizes differ");
    const int cols = (src->w + 63) / 64, n_ctus = cols * ((src->h + 63) / 64);
    for (int i = 0; i < n_ctus; i++)
        for (int c = 0; c < 3; c++)
            if (params[i].type[c] < -1 || params[i].type[c] > 4) return hbi_fail(HB_ERR_ARG, "hb_sao_apply_frame: CTU %d component %d has type %d", i, c, params[i].type[c]);
    const size_t bytes = sizeof(hb_sao_param) * (size_t)n_ctus;
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    if ((rc = hbi_scratch(ctx, 0, bytes, &d_prm, &h_prm)) != HB_OK) goto done;
    memcpy(h_prm, params, bytes);
    crc = hbc_h2d_async(d_prm, h_prm, bytes, ctx->stream);
    if (!crc) { crc = hbk_sao_apply(&src->d, &dst->d, cols, n_ctus, (const hb_sao_param *)d_prm, ctx->stream); ctx->launches++; }
    if (!crc) { crc = hbk_pad_frame(&dst->d, ctx->stream); ctx->launches++; }
    if (!crc) crc = hbc_stream_sync(ctx->stream);
done:
    pthread_mutex_unlock(&ctx->lock);
    if (crc) return hbi_cuda_fail(crc, "hb_sao_apply_frame");
    return rc;
}

/* fill the launch-invariant part of a T/Q launch: tables and shifts of (component, size, qp) -- inter lists 3+comp */
void hbi_tq_setup(hb_ctx *ctx, hbd_tq_args *a, int comp, int n, int qp, int is_islice, int sign_hiding)
{
    int lg = 2;
    while ((1 << lg) < n) lg++;
    const int per = qp / 6, rem = qp % 6;
    a->n = n;
    a->qtab = ctx->d_q + hbi_tab_q_off(lg, 3 + comp, rem);
    a->dqtab = ctx->d_dq + hbi_tab_q_off(lg, 3 + comp, rem);
    a->scan = ctx->d_scan + hbi_tab_scan_off(HB_SCAN_DIAG, lg);
    a->qbits = 14 + per + (15 - 8 - lg);
    a->add = (int32_t)((uint32_t)(is_islice ? 171 : 85) << (a->qbits - 9));    /* hmr_sse42_functions_quant.c:47 */
    a->per = per;
    a->sign_hiding = sign_hiding;
    a->is_luma = comp == 0;
}

/* the T/Q kernels read and write a unit's rows with one vector access per lane: 4 bytes (4x4), 8 (8x8) or 16 (16x16, 32x32),
 * so a unit's x must be a multiple of min(size, 16) -- which every transform unit of the quadtree is */
static int tq_row_align(int size) { return size < 16 ? size : 16; }

void hbi_tq_pack_free(hbi_tq_pack *pk) { free(pk->order); free(pk->coeff_off); pk->order = NULL; pk->coeff_off = NULL; }

/* validate, group by (component, size, qp), launch, and queue the copies of levels and records into the pinned twins; the caller
 * holds ctx->lock, waits for the stream and calls hbi_tq_collect + hbi_tq_pack_free */
int hbi_tq_encode_queue(hb_ctx *ctx, const hb_frame *cur, const hb_frame *pred, hb_frame *recon, const hb_tu_job *jobs, int n_jobs,
                        const hb_tq_params *params, hbi_tq_pack *pk, const char *what)
{
    return hbi_tq_encode_queue_at(ctx, cur, pred, recon, jobs, n_jobs, params, pk, what, 0);
}

/* sb: first of the three scratch buffers to use (a caller that still has work queued on buffers 0.. picks others) */
int hbi_tq_encode_queue_at(hb_ctx *ctx, const hb_frame *cur, const hb_frame *pred, hb_frame *recon, const hb_tu_job *jobs, int n_jobs,
                           const hb_tq_params *params, hbi_tq_pack *pk, const char *what, int sb)
{
    int rc = HB_OK, crc = 0;
    size_t total = 0;
    memset(pk, 0, sizeof *pk);
    for (int i = 0; i < n_jobs; i++) {
        const hb_tu_job *j = &jobs[i];
        if (j->comp < 0 || j->comp > 2 || (j->size != 4 && j->size != 8 && j->size != 16 && j->size != 32) || j->qp < 0 || j->qp > 51 ||
            j->x < 0 || j->y < 0 || (j->x & (tq_row_align(j->size) - 1)) || j->x + j->size > cur->d.p[j->comp].w || j->y + j->size > cur->d.p[j->comp].h ||
            (j->comp && j->size == 32))
            return hbi_fail(HB_ERR_ARG, "%s: job %d is invalid (x must be a multiple of min(size, 16))", what, i);
        total += (size_t)j->size * j->size;
    }
    void *d_xy, *h_xy, *d_co, *h_co, *d_rs, *h_rs;
    char *done_flag = (char *)calloc((size_t)n_jobs, 1);
    pk->order = (int *)malloc(sizeof(int) * (size_t)n_jobs);
    pk->coeff_off = (size_t *)malloc(sizeof(size_t) * (size_t)n_jobs);
    pk->total = total;
    if (!pk->order || !done_flag || !pk->coeff_off) { rc = hbi_fail(HB_ERR_NOMEM, "%s: out of memory", what); goto done; }
    if ((rc = hbi_scratch(ctx, sb, sizeof(int32_t) * 2 * (size_t)n_jobs, &d_xy, &h_xy)) != HB_OK) goto done;
    if ((rc = hbi_scratch(ctx, sb + 1, sizeof(int16_t) * total, &d_co, &h_co)) != HB_OK) goto done;
    if ((rc = hbi_scratch(ctx, sb + 2, sizeof(hb_tu_result) * (size_t)n_jobs, &d_rs, &h_rs)) != HB_OK) goto done;
    pk->h_co = h_co; pk->h_rs = h_rs;
    { size_t o = 0; for (int i = 0; i < n_jobs; i++) { pk->coeff_off[i] = o; o += (size_t)jobs[i].size * jobs[i].size; } }
    /* launch one group per distinct (comp, size, qp); jobs of a group are packed in caller order */
    int packed = 0;
    size_t packed_coeff = 0;
    for (int i = 0; i < n_jobs && !crc; i++) {
        if (done_flag[i]) continue;
        const hb_tu_job key = jobs[i];
        const int g0 = packed;
        const size_t c0 = packed_coeff;
        int32_t *xy = (int32_t *)h_xy;
        for (int k = i; k < n_jobs; k++) {
            if (done_flag[k] || jobs[k].comp != key.comp || jobs[k].size != key.size || jobs[k].qp != key.qp) continue;
            done_flag[k] = 1; pk->order[packed] = k;
            xy[2 * packed] = jobs[k].x; xy[2 * packed + 1] = jobs[k].y;
            packed++; packed_coeff += (size_t)key.size * key.size;
        }
        const int cnt = packed - g0;
        crc = hbc_h2d_async((int32_t *)d_xy + 2 * g0, xy + 2 * g0, sizeof(int32_t) * 2 * (size_t)cnt, ctx->stream);
        if (crc) break;
        hbd_tq_args a;
        memset(&a, 0, sizeof a);
        hbi_tq_setup(ctx, &a, key.comp, key.size, key.qp, params->is_islice, params->sign_hiding);
        a.cur = cur->d.p[key.comp]; a.pred = pred->d.p[key.comp]; a.rec = recon->d.p[key.comp];
        a.jobs_xy = (const int32_t *)d_xy + 2 * g0; a.n_jobs = cnt;
        a.thr_k = hbi_zero_out_k(params->avg_dist);
        a.weight = key.comp ? params->chroma_weight : 1.0;
        a.dyn = NULL;
        a.coeff_out = (int16_t *)d_co + c0;
        a.res_out = (hb_tu_result *)d_rs + g0;
        crc = hbk_tq_encode(&a, ctx->stream);
        ctx->launches++;
    }
    if (!crc) crc = hbc_d2h_async(h_co, d_co, sizeof(int16_t) * total, ctx->stream);
    if (!crc) crc = hbc_d2h_async(h_rs, d_rs, sizeof(hb_tu_result) * (size_t)n_jobs, ctx->stream);
    if (crc) rc = hbi_cuda_fail(crc, what);
done:
    free(done_flag);
    if (rc != HB_OK) hbi_tq_pack_free(pk);
    return rc;
}

/* after the stream has been waited for: packed launch order -> the caller's job order */
void hbi_tq_collect(const hbi_tq_pack *pk, const hb_tu_job *jobs, int n_jobs, int16_t *coeffs, hb_tu_result *results)
{
    size_t o = 0;
    for (int p = 0; p < n_jobs; p++) {
        const int k = pk->order[p];
        const size_t nn = (size_t)jobs[k].size * jobs[k].size;
        memcpy(coeffs + pk->coeff_off[k], (int16_t *)pk->h_co + o, sizeof(int16_t) * nn);
        results[k] = ((hb_tu_result *)pk->h_rs)[p];
        o += nn;
    }
}

int hb_tq_encode(hb_ctx *ctx, const hb_frame *cur, const hb_frame *pred, hb_frame *recon, const hb_tu_job *jobs, int n_jobs,
                 const hb_tq_params *params, int16_t *coeffs, hb_tu_result *results)
{
    int rc, crc = 0;
    hbi_tq_pack pk;
    if (!ctx || !cur || !pred || !recon || !jobs || !params || !coeffs || !results || n_jobs < 0) return hbi_fail(HB_ERR_ARG, "hb_tq_encode: bad argument");
    if (n_jobs == 0) return HB_OK;
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    rc = hbi_tq_encode_queue(ctx, cur, pred, recon, jobs, n_jobs, params, &pk, "hb_tq_encode");
    if (rc == HB_OK) {
        crc = hbc_stream_sync(ctx->stream);
        if (!crc) hbi_tq_collect(&pk, jobs, n_jobs, coeffs, results);
        hbi_tq_pack_free(&pk);
    }
    pthread_mutex_unlock(&ctx->lock);
    if (crc) return hbi_cuda_fail(crc, "hb_tq_encode");
    return rc;
}

/* Intra TUs once their prediction exists: encode_intra_cu after the prediction step (hmr_motion_intra.c:1023-1069) and its
 * chroma counterpart (hmr_motion_intra_chroma.c:340-365).  `pred` holds the intra prediction (from the host, whose intra
 * predictors walk reconstructed neighbours); everything after it runs here. */
int hb_tq_encode_intra(hb_ctx *ctx, const hb_frame *cur, const hb_frame *pred, hb_frame *recon, const hb_intra_tu_job *jobs, int n_jobs,
                       int is_islice, int sign_hiding, double chroma_weight, int16_t *coeffs, hb_tu_result *results)
{
    int rc = HB_OK, crc = 0;
    if (!ctx || !cur || !pred || !recon || !jobs || !coeffs || !results || n_jobs < 0) return hbi_fail(HB_ERR_ARG, "hb_tq_encode_intra: bad argument");
    if (n_jobs == 0) return HB_OK;
    size_t total = 0;
    for (int i = 0; i < n_jobs; i++) {
        const hb_intra_tu_job *j = &jobs[i];
        if (j->comp < 0 || j->comp > 2 || (j->size != 4 && j->size != 8 && j->size != 16 && j->size != 32) || j->qp < 0 || j->qp > 51 ||
            j->x < 0 || j->y < 0 || (j->x & (tq_row_align(j->size) - 1)) || j->x + j->size > cur->d.p[j->comp].w || j->y + j->size > cur->d.p[j->comp].h ||
            (j->comp && j->size == 32) || j->scan_mode < HB_SCAN_HOR || j->scan_mode > HB_SCAN_DIAG || (j->size > 8 && j->scan_mode != HB_SCAN_DIAG))
            return hbi_fail(HB_ERR_ARG, "hb_tq_encode_intra: job %d is invalid", i);
        total += (size_t)j->size * j->size;
    }
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    void *d_xy, *h_xy, *d_co, *h_co, *d_rs, *h_rs;
    int *order = (int *)malloc(sizeof(int) * (size_t)n_jobs);
    char *done_flag = (char *)calloc((size_t)n_jobs, 1);
    size_t *coeff_off = (size_t *)malloc(sizeof(size_t) * (size_t)n_jobs);
    if (!order || !done_flag || !coeff_off) { rc = hbi_fail(HB_ERR_NOMEM, "hb_tq_encode_intra: out of memory"); goto done; }
    if ((rc = hbi_scratch(ctx, 0, sizeof(int32_t) * 2 * (size_t)n_jobs, &d_xy, &h_xy)) != HB_OK) goto done;
    if ((rc = hbi_scratch(ctx, 1, sizeof(int16_t) * total, &d_co, &h_co)) != HB_OK) goto done;
    if ((rc = hbi_scratch(ctx, 2, sizeof(hb_tu_result) * (size_t)n_jobs, &d_rs, &h_rs)) != HB_OK) goto done;
    { size_t o = 0; for (int i = 0; i < n_jobs; i++) { coeff_off[i] = o; o += (size_t)jobs[i].size * jobs[i].size; } }
    int packed = 0;
    size_t packed_coeff = 0;
    for (int i = 0; i < n_jobs && !crc; i++) {
        if (done_flag[i]) continue;
        const hb_intra_tu_job key = jobs[i];
        const int g0 = packed;
        const size_t c0 = packed_coeff;
        int32_t *xy = (int32_t *)h_xy;
        for (int k = i; k < n_jobs; k++) {
            if (done_flag[k] || jobs[k].comp != key.comp || jobs[k].size != key.size || jobs[k].qp != key.qp || jobs[k].scan_mode != key.scan_mode) continue;
            done_flag[k] = 1; order[packed] = k;
            xy[2 * packed] = jobs[k].x; xy[2 * packed + 1] = jobs[k].y;
            packed++; packed_coeff += (size_t)key.size * key.size;
        }
        const int cnt = packed - g0;
        crc = hbc_h2d_async((int32_t *)d_xy + 2 * g0, xy + 2 * g0, sizeof(int32_t) * 2 * (size_t)cnt, ctx->stream);
        if (crc) break;
        hbd_tq_args a;
        memset(&a, 0, sizeof a);
        hbi_tq_setup(ctx, &a, key.comp, key.size, key.qp, is_islice, sign_hiding);
        int lg = 2;
        while ((1 << lg) < key.size) lg++;
        a.qtab = ctx->d_q + hbi_tab_q_off(lg, key.comp, key.qp % 6);        /* intra lists: (is_intra ? 0 : 3) + comp */
        a.dqtab = ctx->d_dq + hbi_tab_q_off(lg, 0, key.qp % 6);             /* SSE4.2 inv_quant: is_intra -> list 0 (:138) */
        a.scan = ctx->d_scan + hbi_tab_scan_off(key.scan_mode, lg);
        a.intra = 1;
        a.cur = cur->d.p[key.comp]; a.pred = pred->d.p[key.comp]; a.rec = recon->d.p[key.comp];
        a.jobs_xy = (const int32_t *)d_xy + 2 * g0; a.n_jobs = cnt;
        a.weight = key.comp ? chroma_weight : 1.0;
        a.coeff_out = (int16_t *)d_co + c0;
        a.res_out = (hb_tu_result *)d_rs + g0;
        crc = hbk_tq_encode(&a, ctx->stream);
        ctx->launches++;
    }
    if (!crc) crc = hbc_d2h_async(h_co, d_co, sizeof(int16_t) * total, ctx->stream);
    if (!crc) crc = hbc_d2h_async(h_rs, d_rs, sizeof(hb_tu_result) * (size_t)n_jobs, ctx->stream);
    if (!crc) crc = hbc_stream_sync(ctx->stream);
    if (!crc) {
        size_t o = 0;
        for (int p = 0; p < n_jobs; p++) {
            const int k = order[p];
            const size_t nn = (size_t)jobs[k].size * jobs[k].size;
            memcpy(coeffs + coeff_off[k], (int16_t *)h_co + o, sizeof(int16_t) * nn);
            results[k] = ((hb_tu_result *)h_rs)[p];
            o += nn;
        }
    }
done:
    free(order); free(done_flag); free(coeff_off);
    pthread_mutex_unlock(&ctx->lock);
    if (crc) return hbi_cuda_fail(crc, "hb_tq_encode_intra");
    return rc;
}

/* ------------------------------------------------------------------ intra prediction
 * adi: per job 4*size+1 reference samples (index 2*size = top-left corner, +i above, -i left), jobs back to back.
 * mode >= 0: the prediction goes to `pred` (then hb_tq_encode_intra); mode < 0: sads[35*i + m] = SAD of luma mode m. */
int hb_intra_run(hb_ctx *ctx, const hb_frame *cur, hb_frame *pred, const hb_intra_job *jobs, int n_jobs, const int16_t *adi, uint32_t *sads)
{
    int rc = HB_OK, crc = 0;
    if (!ctx || !jobs || !adi || n_jobs < 0) return hbi_fail(HB_ERR_ARG, "hb_intra_run: bad argument");
    if (n_jobs == 0) return HB_OK;
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    void *d_jobs, *h_jobs, *d_adi, *h_adi, *d_sad = NULL, *h_sad = NULL;
    /* scratch 0: the job records, then one index list per block size for the SAD-form jobs */
    const size_t jobs_bytes = (sizeof(hbd_intra_job) * (size_t)n_jobs + 15) & ~(size_t)15;
    if ((rc = hbi_scratch(ctx, 0, jobs_bytes + sizeof(int32_t) * (size_t)n_jobs, &d_jobs, &h_jobs)) != HB_OK) goto done;
    hbd_intra_job *hj = (hbd_intra_job *)h_jobs;
    int32_t *hidx = (int32_t *)((char *)h_jobs + jobs_bytes);
    const int32_t *didx = (const int32_t *)((char *)d_jobs + jobs_bytes);
    size_t n_adi = 0;
    int n_pred = 0, first[5] = { 0, 0, 0, 0, 0 }, fill[4];
    for (int i = 0; i < n_jobs; i++) {                       /* one pass: validate, convert, count */
        const hb_intra_job *j = &jobs[i];
        const hb_frame *f = j->mode < 0 ? cur : pred;
        const int sz = j->size == 4 ? 0 : j->size == 8 ? 1 : j->size == 16 ? 2 : j->size == 32 ? 3 : -1;
        if (!f || sz < 0 || j->comp < 0 || j->comp > 2 || j->mode > 34 || (j->mode < 0 && j->comp != 0) || j->x < 0 || j->y < 0 ||
            j->x + j->size > f->d.p[j->comp].w || j->y + j->size > f->d.p[j->comp].h) {
            rc = hbi_fail(HB_ERR_ARG, "hb_intra_run: job %d is invalid", i);
            goto done;
        }
        hj[i].comp = j->comp; hj[i].x = j->x; hj[i].y = j->y; hj[i].size = j->size; hj[i].mode = j->mode;
        hj[i].filtered = j->filtered; hj[i].adi_off = (int32_t)n_adi; hj[i].pad_ = 0;
        n_adi += 4 * (size_t)j->size + 1;
        if (j->mode >= 0) n_pred++; else first[sz + 1]++;
    }
    const int n_sad = n_jobs - n_pred;
    if (n_sad && !sads) { rc = hbi_fail(HB_ERR_ARG, "hb_intra_run: sads is NULL"); goto done; }
    for (int s = 0; s < 4; s++) { first[s + 1] += first[s]; fill[s] = first[s]; }
    if (n_sad) for (int i = 0; i < n_jobs; i++)
        if (jobs[i].mode < 0) hidx[fill[jobs[i].size == 4 ? 0 : jobs[i].size == 8 ? 1 : jobs[i].size == 16 ? 2 : 3]++] = i;
    if ((rc = hbi_scratch(ctx, 1, sizeof(int16_t) * n_adi, &d_adi, &h_adi)) != HB_OK) goto done;
    if (n_sad && (rc = hbi_scratch(ctx, 2, sizeof(uint32_t) * 35 * (size_t)n_jobs, &d_sad, &h_sad)) != HB_OK) goto done;
    crc = hbc_h2d_async(d_jobs, h_jobs, jobs_bytes + sizeof(int32_t) * (size_t)first[4], ctx->stream);
    /* the caller's samples go up without an extra host copy (pinned: one DMA; pageable: staged by the runtime before the call returns) */
    if (!crc) crc = hbc_h2d_async(d_adi, adi, sizeof(int16_t) * n_adi, ctx->stream);
    hbd_intra_args a;
    memset(&a, 0, sizeof a);
    if (cur) a.cur = cur->d.p[0];
    if (pred) a.pred = pred->d;
    a.jobs = (const hbd_intra_job *)d_jobs; a.n_jobs = n_jobs; a.adi = (const int16_t *)d_adi; a.sads = (uint32_t *)d_sad;
    if (!crc && n_pred) { crc = hbk_intra(&a, ctx->stream); ctx->launches++; }
    for (int s = 0; s < 4 && !crc; s++)
        if (first[s + 1] > first[s]) { crc = hbk_intra_sads(&a, 4 << s, didx + first[s], first[s + 1] - first[s], ctx->stream); ctx->launches++; }
    /* rows of prediction-form jobs are written by no kernel: the caller's table is only touched for SAD-form jobs.  All SAD form
     * (the mode search): straight into the caller's table, no staging */
    if (!crc && n_sad) crc = n_pred ? hbc_d2h_async(h_sad, d_sad, sizeof(uint32_t) * 35 * (size_t)n_jobs, ctx->stream)
                                    : hbc_d2h_async(sads, d_sad, sizeof(uint32_t) * 35 * (size_t)n_jobs, ctx->stream);
    if (!crc) crc = hbc_stream_sync(ctx->stream);
    if (!crc && n_sad && n_pred)
        for (int i = 0; i < n_jobs; i++) if (jobs[i].mode < 0) memcpy(sads + 35 * (size_t)i, (uint32_t *)h_sad + 35 * (size_t)i, sizeof(uint32_t) * 35);
done:
    pthread_mutex_unlock(&ctx->lock);
    if (crc) return hbi_cuda_fail(crc, "hb_intra_run");
    return rc;
}
