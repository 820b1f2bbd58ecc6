/*
 * prepass_demo.c -- the C ABI of libhomer_b200.so used from plain C99, the way an encoder host thread would:
 * two synthetic 8-bit 4:2:0 frames -> resident frames -> frame-level pre-pass (motion search at PU 64/32/16/8, chroma MC,
 * inter T/Q at TU 32/32/16/8/4) -> cost tables -> the stand-in depth decision -> gather of the chosen reconstruction and levels.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/prepass_demo.c -o build/prepass_demo -Lhomerhevc_b200 -lhomer_b200 -Wl,-rpath,'$ORIGIN/../homerhevc_b200' -lm
 *   build/prepass_demo [width height frames]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "homer_b200.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != HB_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, hb_last_error()); return 1; } } while (0)

/* a smooth texture that pans by (2.25, 1.5) samples per frame plus a little noise */
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

int main(int argc, char **argv)
{
    const int w = argc > 2 ? atoi(argv[1]) : 1280, h = argc > 2 ? atoi(argv[2]) : 720, frames = argc > 3 ? atoi(argv[3]) : 8;
    const size_t luma = (size_t)w * h, fb = luma * 3 / 2;
    hb_ctx *ctx;
    hb_frame *cur, *ref;
    hb_prepass *pp;
    hb_prepass_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.qp = 32; cfg.chroma_qp_offset = 2; cfg.sign_hiding = 1; cfg.me_action = HB_ME_PEL | HB_ME_HALF | HB_ME_QUARTER; cfg.use_graph = 1; cfg.compact_tables = 1;

    CHECK(hb_ctx_create(&ctx, 0));
    CHECK(hb_frame_create(ctx, w, h, &cur));
    CHECK(hb_frame_create(ctx, w, h, &ref));
    CHECK(hb_prepass_create(ctx, w, h, &cfg, &pp));

    /* pinned host memory: two input frames, the cost tables, the gathered output */
    uint8_t *in = (uint8_t *)hb_pinned_alloc(2 * fb);
    void *tables = hb_pinned_alloc(hb_prepass_tables_bytes(pp));
    const size_t out_cap = fb + 4 * luma;
    void *out = hb_pinned_alloc(out_cap);
    const int n_ctus = hb_prepass_num_ctus(pp);
    uint8_t *sel = (uint8_t *)malloc((size_t)n_ctus);
    int32_t *off = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n_ctus + 1));
    if (!in || !tables || !out || !sel || !off) { fprintf(stderr, "out of memory\n"); return 1; }

    make_frame(in, in + luma, in + luma + luma / 4, w, h, 0);
    double avg_dist = 650.0;
    for (int n = 1; n <= frames; n++) {
        uint8_t *prev = in + ((n - 1) & 1) * fb, *now = in + (n & 1) * fb;
        make_frame(now, now + luma, now + luma + luma / 4, w, h, n);
        const uint8_t *cp[3] = { now, now + luma, now + luma + luma / 4 }, *rp[3] = { prev, prev + luma, prev + luma + luma / 4 };
        size_t out_bytes = 0;
        float ms = 0;
        CHECK(hb_timer_begin(ctx));
        CHECK(hb_prepass_process_frame(pp, cur, ref, cp, rp, avg_dist, 60, tables, hb_prepass_tables_bytes(pp), sel, off, out, out_cap, &out_bytes));
        CHECK(hb_timer_end(ctx, &ms));
        int hist[5] = { 0, 0, 0, 0, 0 };
        for (int i = 0; i < n_ctus; i++) hist[sel[i]]++;
        /* the first depth-0 record of the compact tables: the 64x64 PU at the origin */
        const hb_me_result_c *me = (const hb_me_result_c *)tables;
        printf("frame %d: %.3f ms on the device, %zu bytes back; CTU choices 64/32/16/8/8+4x4 = %d/%d/%d/%d/%d; PU(0,0,64) mv (%d,%d)/4 sad %u\n",
               n, ms, out_bytes, hist[0], hist[1], hist[2], hist[3], hist[4], me[0].mvx, me[0].mvy, me[0].sad);
        if (me[0].sad == 0xffffffffu || out_bytes < fb) { fprintf(stderr, "unexpected result\n"); return 1; }
    }
    /* ---- the same stream of frames with the reference picture kept on the device: only the source goes up; the finished
     * picture of frame n (chosen units gathered, deblocked, SAO applied) is the reference of frame n + 1 */
    {
        hb_frame *rec, *refs[2] = { ref, NULL };
        hb_sao_param *sao = (hb_sao_param *)malloc(sizeof(hb_sao_param) * (size_t)n_ctus);
        const hb_deblock_params dbk = { 2, 2, 0, 0 };
        const double sao_lambda[3] = { 60.0, 48.0, 48.0 };
        CHECK(hb_frame_create(ctx, w, h, &rec));
        CHECK(hb_frame_create(ctx, w, h, &refs[1]));
        if (!sao) { fprintf(stderr, "out of memory\n"); return 1; }
        make_frame(in, in + luma, in + luma + luma / 4, w, h, 0);
        CHECK(hb_frame_upload_u8(ctx, refs[0], in, w, in + luma, w / 2, in + luma + luma / 4, w / 2));     /* frame 0 stands for an intra picture */
        for (int n = 1; n <= frames; n++) {
            uint8_t *now = in + (n & 1) * fb;
            make_frame(now, now + luma, now + luma + luma / 4, w, h, n);
            const uint8_t *cp[3] = { now, now + luma, now + luma + luma / 4 };
            size_t level_bytes = 0;
            CHECK(hb_prepass_frame_begin_resident(pp, cur, refs[(n - 1) & 1], cp, avg_dist, tables, hb_prepass_tables_bytes(pp)));
            CHECK(hb_prepass_frame_finish_resident(pp, cur, 60, tables, sel, off, rec, refs[n & 1], &dbk, sao_lambda, out, out_cap, &level_bytes, sao));
            int with_sao = 0;
            for (int i = 0; i < n_ctus; i++) with_sao += sao[i].type[0] >= 0;
            printf("resident frame %d: %zu bytes of levels back, SAO on in %d of %d CTUs (luma)\n", n, level_bytes, with_sao, n_ctus);
        }
        CHECK(hb_ctx_sync(ctx));
        free(sao);
        hb_frame_destroy(rec); hb_frame_destroy(refs[1]);
    }
    printf("%llu kernel launches\n", (unsigned long long)hb_ctx_launch_count(ctx));
    free(sel); free(off);
    hb_pinned_free(in); hb_pinned_free(tables); hb_pinned_free(out);
    hb_prepass_destroy(pp); hb_frame_destroy(cur); hb_frame_destroy(ref); hb_ctx_destroy(ctx);
    puts("prepass_demo ok");
    return 0;
}
