/*
 * e2e_threads.c -- the end-to-end flow bench.py times (`e2e`), driven by C host threads instead of Python ones: T threads, each with
 * two frame streams in flight (it begins frame n+1 on its second stream before it blocks in the finish call of frame n, as an encoder
 * engine thread would).  Per frame and stream: the source goes up from pinned host memory, the pre-pass runs against the finished
 * previous picture in HBM, the cost tables come down, the host picks a depth per CTU, gather + deblocking + SAO produce the next
 * reference picture on the device, the coded levels and the SAO decision come down.  Prints frames/s over the whole run.
 *
 *   gcc -std=c99 -O2 -pthread -Iinclude examples/e2e_threads.c -o build/e2e_threads -Lhomerhevc_b200 -lhomer_b200 -Wl,-rpath,'$ORIGIN/../homerhevc_b200' -lm
 *   build/e2e_threads [width height threads frames_per_stream]
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "homer_b200.h"

#define N_SRC 8                       /* distinct source pictures, cycled */

typedef struct stream_state {
    hb_ctx *ctx;
    hb_frame *cur, *rec, *refs[2];
    hb_prepass *pp;
    void *tables, *levels;
    size_t tables_bytes, levels_cap;
    uint8_t *sel;
    int32_t *off;
    hb_sao_param *sao;
    int which, count;
} stream_state;

typedef struct worker {
    pthread_t th;
    stream_state s[2];
    int frames_per_stream, failed;
    uint64_t bytes_up, bytes_down;
} worker;

static int g_w, g_h;
static uint8_t *g_src[N_SRC];
static size_t g_luma;

static void make_frame(uint8_t *y, uint8_t *u, uint8_t *v, int w, int h, int n)
{
    unsigned s = 12345u + 977u * (unsigned)n;
    for (int r = 0; r < h; r++)
        for (int c = 0; c < w; c++) {
            const double x = c + 2.25 * n, yy = r + 1.5 * n;
            double t = 128 + 60 * sin(x * 0.071) * cos(yy * 0.053) + 30 * sin((x + yy) * 0.19);
            s = s * 1664525u + 1013904223u;
            t += (double)((s >> 24) & 7) - 3.5;
            y[r * w + c] = (uint8_t)(t < 0 ? 0 : t > 255 ? 255 : t);
        }
    for (int r = 0; r < h / 2; r++)
        for (int c = 0; c < w / 2; c++) {
            u[r * (w / 2) + c] = (uint8_t)(128 + 40 * sin((c + 1.125 * n) * 0.11));
            v[r * (w / 2) + c] = (uint8_t)(128 + 40 * cos((r + 0.75 * n) * 0.09));
        }
}

static int stream_open(stream_state *s)
{
    hb_prepass_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    memset(s, 0, sizeof *s);
    cfg.qp = 32; cfg.chroma_qp_offset = 2; cfg.sign_hiding = 1; cfg.me_action = HB_ME_PEL | HB_ME_HALF | HB_ME_QUARTER; cfg.use_graph = 1; cfg.compact_tables = 2;
    if (hb_ctx_create(&s->ctx, 0) || hb_frame_create(s->ctx, g_w, g_h, &s->cur) || hb_frame_create(s->ctx, g_w, g_h, &s->rec) ||
        hb_frame_create(s->ctx, g_w, g_h, &s->refs[0]) || hb_frame_create(s->ctx, g_w, g_h, &s->refs[1]) || hb_prepass_create(s->ctx, g_w, g_h, &cfg, &s->pp)) return 1;
    const int n_ctus = hb_prepass_num_ctus(s->pp);
    s->tables_bytes = hb_prepass_tables_bytes(s->pp);
    s->levels_cap = 4 * g_luma;
    s->tables = hb_pinned_alloc(s->tables_bytes); s->levels = hb_pinned_alloc(s->levels_cap);
    s->sel = (uint8_t *)malloc((size_t)n_ctus); s->off = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n_ctus + 1));
    s->sao = (hb_sao_param *)malloc(sizeof(hb_sao_param) * (size_t)n_ctus);
    if (!s->tables || !s->levels || !s->sel || !s->off || !s->sao) return 1;
    /* picture 0 stands for the intra picture the stream starts from */
    if (hb_frame_upload_u8(s->ctx, s->refs[0], g_src[0], g_w, g_src[0] + g_luma, g_w / 2, g_src[0] + g_luma + g_luma / 4, g_w / 2)) return 1;
    return hb_ctx_sync(s->ctx);
}

static int frame_begin(stream_state *s)
{
    const uint8_t *p = g_src[1 + s->count % (N_SRC - 1)];
    const uint8_t *cp[3] = { p, p + g_luma, p + g_luma + g_luma / 4 };
    s->count++;
    return hb_prepass_frame_begin_resident(s->pp, s->cur, s->refs[s->which], cp, 650.0, s->tables, s->tables_bytes);
}

static int frame_finish(stream_state *s, uint64_t *down)
{
    static const hb_deblock_params dbk = { 2, 2, 0, 0 };
    static const double sao_lambda[3] = { 60.0, 60.0 / 1.26, 60.0 / 1.26 };
    size_t level_bytes = 0;
    const int rc = hb_prepass_frame_finish_resident(s->pp, s->cur, 60, s->tables, s->sel, s->off, s->rec, s->refs[1 - s->which], &dbk, sao_lambda,
                                                    s->levels, s->levels_cap, &level_bytes, s->sao);
    s->which = 1 - s->which;
    *down += level_bytes + s->tables_bytes;
    return rc;
}

static void *work(void *arg)
{
    worker *wk = (worker *)arg;
    stream_state *pending = NULL;
    for (int i = 0; i < 2 * wk->frames_per_stream && !wk->failed; i++) {
        stream_state *s = &wk->s[i & 1];
        if (frame_begin(s)) { wk->failed = 1; break; }
        wk->bytes_up += g_luma * 3 / 2;
        if (pending && frame_finish(pending, &wk->bytes_down)) { wk->failed = 1; break; }
        pending = s;
    }
    if (pending && !wk->failed && frame_finish(pending, &wk->bytes_down)) wk->failed = 1;
    if (wk->failed) fprintf(stderr, "worker failed: %s\n", hb_last_error());
    return NULL;
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

int main(int argc, char **argv)
{
    g_w = argc > 2 ? atoi(argv[1]) : 1920; g_h = argc > 2 ? atoi(argv[2]) : 1080;
    const int threads = argc > 3 ? atoi(argv[3]) : 8, fps_ = argc > 4 ? atoi(argv[4]) : 200;
    g_luma = (size_t)g_w * g_h;
    for (int n = 0; n < N_SRC; n++) {
        g_src[n] = (uint8_t *)hb_pinned_alloc(g_luma * 3 / 2);
        if (!g_src[n]) { fprintf(stderr, "out of pinned memory\n"); return 1; }
        make_frame(g_src[n], g_src[n] + g_luma, g_src[n] + g_luma + g_luma / 4, g_w, g_h, n);
    }
    worker *wk = (worker *)calloc((size_t)threads, sizeof *wk);
    if (!wk) return 1;
    for (int t = 0; t < threads; t++)
        for (int k = 0; k < 2; k++)
            if (stream_open(&wk[t].s[k])) { fprintf(stderr, "stream_open failed: %s\n", hb_last_error()); return 1; }
    for (int pass = 0; pass < 2; pass++) {            /* pass 0: warm-up (graph captures, lazy allocations), pass 1: timed */
        const int per = pass ? fps_ : 6;
        for (int t = 0; t < threads; t++) { wk[t].frames_per_stream = per; wk[t].bytes_up = wk[t].bytes_down = 0; }
        const double t0 = now_s();
        for (int t = 0; t < threads; t++) if (pthread_create(&wk[t].th, NULL, work, &wk[t])) { fprintf(stderr, "pthread_create failed\n"); return 1; }
        for (int t = 0; t < threads; t++) pthread_join(wk[t].th, NULL);
        for (int t = 0; t < threads; t++) for (int k = 0; k < 2; k++) if (hb_ctx_sync(wk[t].s[k].ctx)) wk[t].failed = 1;
        const double secs = now_s() - t0;
        uint64_t up = 0, down = 0;
        for (int t = 0; t < threads; t++) { if (wk[t].failed) return 1; up += wk[t].bytes_up; down += wk[t].bytes_down; }
        if (pass) {
            const int frames = 2 * threads * per;
            printf("e2e_threads: %dx%d, %d C host threads x 2 streams, %d frames in %.3f s = %.1f frames/s (%.2f MB up, %.2f MB down per frame)\n",
                   g_w, g_h, threads, frames, secs, frames / secs, 1e-6 * (double)up / frames, 1e-6 * (double)down / frames);
        }
    }
    puts("e2e_threads ok");
    return 0;
}
