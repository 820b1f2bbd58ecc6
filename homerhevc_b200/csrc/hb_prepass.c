/*
 * hb_prepass.c -- frame-level batched pre-pass (section D of include/homer_b200.h).
 *
 * One plan per (frame size, QP, band).  Static job lists for every PU size 64/32/16/8 and every TU pass live in
 * HBM; a frame is one replay of: ME(d) -> MC(d) for d = 0..3 (children read their parent's vector from the result
 * table of depth d-1, hmr_motion_inter.c:2613), then the inter T/Q chain per pass for Y, U, V.  With use_graph the
 * whole sequence is captured once per (cur, ref) frame pair and replayed as a single CUDA graph launch; the only
 * per-frame scalars (MV-cost weight, zero-out threshold -- both functions of avg_dist) are read by the kernels
 * from a small device block refreshed before each replay.
 */
#define _POSIX_C_SOURCE 200809L
#include "hb_host.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define N_DEPTH HB_PREPASS_DEPTHS
#define N_PASS  HB_PREPASS_TQ_PASSES
#define MAX_GRAPHS 64

typedef struct pass_comp {
    int n_tus, tu;                 /* TU count and size (0 = component not coded in this pass) */
    int32_t *d_xy;                 /* job coordinates */
    int16_t *d_coeff;
    hb_tu_result *d_res;
    int grid_w, grid_h;            /* TU grid of the frame (whole CTUs) */
    int32_t *d_index;              /* grid position -> index among coded TUs or -1 (device) */
    int32_t *h_ctu;                /* coded TU -> CTU number (host) */
} pass_comp;

struct hb_prepass {
    hb_ctx *ctx;
    hb_prepass_cfg cfg;
    int w, h, ctu_cols, ctu_rows, row0, rows;
    int qp_c;
    double weight_c;
    /* motion search / compensation */
    int grid_w[N_DEPTH], grid_h[N_DEPTH];      /* PU raster grid per depth (covers whole CTUs) */
    int n_valid[N_DEPTH];
    hbd_me_job *d_jobs[N_DEPTH];
    hbd_me_job *d_jobs_strip[N_DEPTH]; int n_strip[N_DEPTH];     /* the same jobs in strip order for the windowed search kernels */
    hbd_mc_pu *d_pus[N_DEPTH];
    hb_me_result *d_me[N_DEPTH];
    char *d_tables;                                 /* one block: ME results of every depth, then TU results of every (pass, comp) --
                                                     * exactly the layout hb_prepass_fetch_tables delivers, so the fetch is ONE copy */
    size_t tables_bytes;
    char *d_tables_c;                               /* compact copy (cfg.compact_tables), packed right before each fetch */
    int n_me_total, n_tu_total, n_cu_total;
    hb_frame *pred[N_DEPTH];
    /* T/Q */
    pass_comp pc[N_PASS][3];
    hb_frame *recon[N_PASS];
    /* per-frame scalars */
    hbd_dyn_params *d_dyn;
    /* captured replays */
    /* keyed on the DEVICE addresses the captured kernels hold (the six planes of cur and ref), not on the host hb_frame
     * structs: a destroyed frame's struct address is readily handed out again by calloc for a frame with other planes */
    struct { const uint8_t *plane[7]; void *exec; } graphs[MAX_GRAPHS];
    int n_graphs;
    int launches_per_frame;
    /* gather of the host's selection */
    uint8_t *d_sel; int32_t *d_ctu_off; uint8_t *d_sel_recon; int16_t *d_sel_levels; size_t sel_levels_cap;
    hb_unit_info *d_units; uint8_t *d_dbk_maps;     /* hb_prepass_finalise: unit data, then strengths (2 planes) + QP map */
    char *d_sao, *h_sao;                            /* resident flow: statistics | candidates | parameters (device); candidates | parameters (pinned) */
    /* side streams per depth: [0] MC then luma T/Q, [1] chroma T/Q, [2] the 4x4 luma pass of depth 3.  They overlap the search of depth d+1 */
    void *side[N_DEPTH][3];
    void *ev_fork[N_DEPTH], *ev_mc[N_DEPTH], *ev_join[N_DEPTH][3];
    void *prof_ev[HB_PREPASS_MAX_KERNELS + 1];
    char prof_name[HB_PREPASS_MAX_KERNELS][16];
};

static int pass_depth(int pass) { return pass < N_DEPTH ? pass : N_DEPTH - 1; }
static int pass_luma_tu(int pass) { static const int t[N_PASS] = { 32, 32, 16, 8, 4 }; return t[pass]; }

static int pu_valid(const hb_prepass *pp, int x, int y, int s)
{
    const int ctu_row = y / 64;
    return x + s <= pp->w && y + s <= pp->h && ctu_row >= pp->row0 && ctu_row < pp->row0 + pp->rows;
}

static int upload(hb_ctx *ctx, void **dev, const void *host, size_t bytes)
{
    int rc = hbc_malloc(dev, bytes ? bytes : 16);
    if (rc) return hbi_cuda_fail(rc, "prepass: cudaMalloc");
    if (bytes) {
        rc = hbc_h2d_async(*dev, host, bytes, ctx->stream);
        if (!rc) rc = hbc_stream_sync(ctx->stream);      /* host staging is pageable and freed by the caller */
        if (rc) return hbi_cuda_fail(rc, "prepass: upload");
    }
    return HB_OK;
}

int hb_prepass_create(hb_ctx *ctx, int width, int height, const hb_prepass_cfg *cfg, hb_prepass **out)
{
    int rc = HB_OK;
    if (!ctx || !cfg || !out) return hbi_fail(HB_ERR_ARG, "hb_prepass_create: NULL argument");
    if (cfg->qp < 0 || cfg->qp > 51) return hbi_fail(HB_ERR_ARG, "hb_prepass_create: qp %d", cfg->qp);
    *out = NULL;
    hb_prepass *pp = (hb_prepass *)calloc(1, sizeof *pp);
    if (!pp) return hbi_fail(HB_ERR_NOMEM, "hb_prepass_create: out of memory");
    pp->ctx = ctx; pp->cfg = *cfg; pp->w = width; pp->h = height;
    pp->ctu_cols = (width + 63) / 64; pp->ctu_rows = (height + 63) / 64;
    pp->row0 = cfg->band_ctu_rows > 0 ? cfg->band_ctu_row0 : 0;
    pp->rows = cfg->band_ctu_rows > 0 ? cfg->band_ctu_rows : pp->ctu_rows;
    if (pp->row0 < 0 || pp->row0 + pp->rows > pp->ctu_rows) { free(pp); return hbi_fail(HB_ERR_ARG, "hb_prepass_create: band outside the frame"); }
    pp->qp_c = hbi_chroma_qp(cfg->qp, cfg->chroma_qp_offset);
    pp->weight_c = pow(2.0, (cfg->qp - pp->qp_c) / 3.0);           /* hmr_motion_inter.c:155 */
    hbc_set_device(ctx->device);

    /* ---- PU job lists */
    for (int d = 0; d < N_DEPTH && rc == HB_OK; d++) {
        const int s = 64 >> d;
        const int gw = pp->ctu_cols * (64 / s), gh = pp->ctu_rows * (64 / s);
        pp->grid_w[d] = gw; pp->grid_h[d] = gh;
        hbd_me_job *jobs = (hbd_me_job *)calloc((size_t)gw * gh, sizeof *jobs);
        hbd_mc_pu *pus = (hbd_mc_pu *)calloc((size_t)gw * gh, sizeof *pus);
        if (!jobs || !pus) { free(jobs); free(pus); rc = hbi_fail(HB_ERR_NOMEM, "hb_prepass_create: out of memory"); break; }
        int n = 0;
        for (int py = 0; py < gh; py++)
            for (int px = 0; px < gw; px++) {
                const int idx = py * gw + px;
                if (!pu_valid(pp, px * s, py * s, s)) continue;
                hbd_me_job *j = &jobs[n];
                j->x = px * s; j->y = py * s;
                j->n_amvp = 2;                               /* zero predictors: the AMVP list always holds two entries */
                j->n_start = 0;
                j->parent = -1;
                if (d > 0 && pu_valid(pp, (px / 2) * 2 * s, (py / 2) * 2 * s, 2 * s)) j->parent = (py / 2) * pp->grid_w[d - 1] + px / 2;
                j->out = idx;
                j->corr = 0.;                                /* taken from the per-frame block */
                pus[n].x = j->x; pus[n].y = j->y; pus[n].mv_idx = idx;
                n++;
            }
        pp->n_valid[d] = n;
        /* strip order: per PU row, runs of hbk_me_strip_pus(s) neighbours starting at a multiple of that count; runs without a valid PU are
         * dropped, the rest of a partly valid run is padding (x = -1).  A run is what one CTA of the windowed search kernel works on. */
        {
            const int run = hbk_me_strip_pus(s);
            hbd_me_job *strip = (hbd_me_job *)calloc((size_t)gh * ((size_t)(gw + run - 1) / run) * run, sizeof *strip);
            if (!strip) { free(jobs); free(pus); rc = hbi_fail(HB_ERR_NOMEM, "hb_prepass_create: out of memory"); break; }
            int ns = 0, k = 0;
            for (int py = 0; py < gh; py++)
                for (int px0 = 0; px0 < gw; px0 += run) {
                    if (!pu_valid(pp, px0 * s, py * s, s)) continue;
                    for (int px = px0; px < px0 + run; px++) {
                        if (px < gw && pu_valid(pp, px * s, py * s, s)) strip[ns++] = jobs[k++];       /* `jobs` is in raster order of the valid PUs */
                        else { memset(&strip[ns], 0, sizeof strip[ns]); strip[ns].x = -1; strip[ns].y = -1; strip[ns].parent = -1; ns++; }
                    }
                }
            pp->n_strip[d] = ns;
            if (k != n) rc = hbi_fail(HB_ERR_ARG, "hb_prepass_create: strip order lost PUs (%d of %d)", k, n);
            if (rc == HB_OK) rc = upload(ctx, (void **)&pp->d_jobs_strip[d], strip, sizeof *strip * (size_t)(ns ? ns : 1));
            free(strip);
            if (rc != HB_OK) { free(jobs); free(pus); break; }
        }
        rc = upload(ctx, (void **)&pp->d_jobs[d], jobs, sizeof *jobs * (size_t)n);
        if (rc == HB_OK) rc = upload(ctx, (void **)&pp->d_pus[d], pus, sizeof *pus * (size_t)n);
        free(jobs); free(pus);
        if (rc == HB_OK) rc = hb_frame_create(ctx, width, height, &pp->pred[d]);
    }
    /* ---- TU job lists: a TU is coded when the PU it belongs to is valid */
    for (int p = 0; p < N_PASS && rc == HB_OK; p++) {
        const int d = pass_depth(p), s = 64 >> d;
        rc = hb_frame_create(ctx, width, height, &pp->recon[p]);
        for (int c = 0; c < 3 && rc == HB_OK; c++) {
            pass_comp *pc = &pp->pc[p][c];
            const int tu = c == 0 ? pass_luma_tu(p) : pass_luma_tu(p) / 2;
            if (c > 0 && p == N_PASS - 1) { pc->tu = 0; continue; }   /* 8x8 CUs code chroma as one 4x4 (pass 3) */
            pc->tu = tu;
            const int pw = c ? width / 2 : width, ph = c ? height / 2 : height, sc = c ? s / 2 : s;
            const int tw = (pp->ctu_cols * (c ? 32 : 64)) / tu, th = (pp->ctu_rows * (c ? 32 : 64)) / tu;
            int32_t *xy = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)tw * th);
            int32_t *index = (int32_t *)malloc(sizeof(int32_t) * (size_t)tw * th);
            pc->h_ctu = (int32_t *)malloc(sizeof(int32_t) * (size_t)tw * th);
            pc->grid_w = tw; pc->grid_h = th;
            if (!xy || !index || !pc->h_ctu) { free(xy); free(index); rc = hbi_fail(HB_ERR_NOMEM, "hb_prepass_create: out of memory"); break; }
            int n = 0;
            for (int ty = 0; ty < th; ty++)
                for (int tx = 0; tx < tw; tx++) {
                    const int x = tx * tu, y = ty * tu;
                    index[ty * tw + tx] = -1;
                    if (x + tu > pw || y + tu > ph) continue;
                    const int pux = (x / sc) * s, puy = (y / sc) * s;     /* owning PU in luma samples */
                    if (!pu_valid(pp, pux, puy, s)) continue;
                    index[ty * tw + tx] = n;
                    pc->h_ctu[n] = (puy / 64) * pp->ctu_cols + pux / 64;
                    xy[2 * n] = x; xy[2 * n + 1] = y; n++;
                }
            pc->n_tus = n;
            rc = upload(ctx, (void **)&pc->d_xy, xy, sizeof(int32_t) * 2 * (size_t)n);
            if (rc == HB_OK) rc = upload(ctx, (void **)&pc->d_index, index, sizeof(int32_t) * (size_t)tw * th);
            free(xy); free(index);
            int crc = 0;
            if (rc == HB_OK && (crc = hbc_malloc((void **)&pc->d_coeff, sizeof(int16_t) * (size_t)(n ? n : 1) * tu * tu))) rc = hbi_cuda_fail(crc, "prepass: coeff");
        }
    }
    /* ---- result tables: one device block in the layout of hb_prepass_fetch_tables */
    if (rc == HB_OK) {
        size_t me_bytes = 0, total;
        for (int d = 0; d < N_DEPTH; d++) me_bytes += sizeof(hb_me_result) * (size_t)pp->grid_w[d] * pp->grid_h[d];
        total = me_bytes;
        for (int p = 0; p < N_PASS; p++) for (int c = 0; c < 3; c++) total += sizeof(hb_tu_result) * (size_t)pp->pc[p][c].n_tus;
        hb_me_result *init = (hb_me_result *)calloc(me_bytes / sizeof(hb_me_result) + 1, sizeof *init);
        if (!init) rc = hbi_fail(HB_ERR_NOMEM, "hb_prepass_create: out of memory");
        else {
            for (size_t i = 0; i < me_bytes / sizeof *init; i++) init[i].sad = 0xffffffffu;       /* PUs outside the band / picture */
            int crc = hbc_malloc((void **)&pp->d_tables, total + 16);
            if (!crc) crc = hbc_memset_async(pp->d_tables, 0, total + 16, ctx->stream);
            if (!crc) crc = hbc_h2d_async(pp->d_tables, init, me_bytes, ctx->stream);
            if (!crc) crc = hbc_stream_sync(ctx->stream);
            free(init);
            if (crc) rc = hbi_cuda_fail(crc, "prepass: result tables");
            else {
                char *o = pp->d_tables;
                pp->tables_bytes = total;
                pp->n_me_total = (int)(me_bytes / sizeof(hb_me_result));
                pp->n_tu_total = (int)((total - me_bytes) / sizeof(hb_tu_result));
                pp->n_cu_total = 0;
                for (int p = 0; p < N_PASS; p++) pp->n_cu_total += pp->grid_w[pass_depth(p)] * pp->grid_h[pass_depth(p)];
                if (cfg->compact_tables) {
                    const int c2 = hbc_malloc((void **)&pp->d_tables_c, sizeof(hb_me_result_c) * (size_t)pp->n_me_total + sizeof(hb_tu_result_c) * (size_t)pp->n_tu_total + 16);
                    if (c2) rc = hbi_cuda_fail(c2, "prepass: compact tables");
                }
                for (int d = 0; d < N_DEPTH; d++) { pp->d_me[d] = (hb_me_result *)o; o += sizeof(hb_me_result) * (size_t)pp->grid_w[d] * pp->grid_h[d]; }
                for (int p = 0; p < N_PASS; p++) for (int c = 0; c < 3; c++) { pp->pc[p][c].d_res = (hb_tu_result *)o; o += sizeof(hb_tu_result) * (size_t)pp->pc[p][c].n_tus; }
            }
        }
    }
    for (int d = 0; d < N_DEPTH && rc == HB_OK; d++) {
        int crc = hbc_event_create_notiming(&pp->ev_fork[d]);
        if (!crc) crc = hbc_event_create_notiming(&pp->ev_mc[d]);
        for (int k = 0; k < 3 && !crc; k++) {
            crc = hbc_stream_create(&pp->side[d][k]);
            if (!crc) crc = hbc_event_create_notiming(&pp->ev_join[d][k]);
        }
        if (crc) rc = hbi_cuda_fail(crc, "prepass: side streams");
    }
    if (rc == HB_OK) {
        const int crc = hbc_malloc((void **)&pp->d_dyn, sizeof *pp->d_dyn);
        if (crc) rc = hbi_cuda_fail(crc, "prepass: dyn block");
    }
    if (rc != HB_OK) { hb_prepass_destroy(pp); return rc; }
    *out = pp;
    return HB_OK;
}

void hb_prepass_destroy(hb_prepass *pp)
{
    if (!pp) return;
    hbc_set_device(pp->ctx->device);
    hbc_stream_sync(pp->ctx->stream);
    for (int i = 0; i < pp->n_graphs; i++) hbc_graph_destroy(pp->graphs[i].exec);
    for (int d = 0; d < N_DEPTH; d++) {
        if (pp->d_jobs[d]) hbc_free(pp->d_jobs[d]);
        if (pp->d_jobs_strip[d]) hbc_free(pp->d_jobs_strip[d]);
        if (pp->d_pus[d]) hbc_free(pp->d_pus[d]);
        hb_frame_destroy(pp->pred[d]);
    }
    for (int p = 0; p < N_PASS; p++) {
        for (int c = 0; c < 3; c++) {
            if (pp->pc[p][c].d_xy) hbc_free(pp->pc[p][c].d_xy);
            if (pp->pc[p][c].d_coeff) hbc_free(pp->pc[p][c].d_coeff);
            if (pp->pc[p][c].d_index) hbc_free(pp->pc[p][c].d_index);
            free(pp->pc[p][c].h_ctu);
        }
        hb_frame_destroy(pp->recon[p]);
    }
    if (pp->d_tables) hbc_free(pp->d_tables);
    if (pp->d_tables_c) hbc_free(pp->d_tables_c);
    if (pp->d_dyn) hbc_free(pp->d_dyn);
    if (pp->d_sel) hbc_free(pp->d_sel);
    if (pp->d_ctu_off) hbc_free(pp->d_ctu_off);
    if (pp->d_sel_recon) hbc_free(pp->d_sel_recon);
    if (pp->d_sel_levels) hbc_free(pp->d_sel_levels);
    if (pp->d_units) hbc_free(pp->d_units);
    if (pp->d_dbk_maps) hbc_free(pp->d_dbk_maps);
    if (pp->d_sao) hbc_free(pp->d_sao);
    if (pp->h_sao) hbc_host_free(pp->h_sao);
    for (int d = 0; d < N_DEPTH; d++) {
        for (int k = 0; k < 3; k++) {
            if (pp->side[d][k]) { hbc_stream_sync(pp->side[d][k]); hbc_stream_destroy(pp->side[d][k]); }
            if (pp->ev_join[d][k]) hbc_event_destroy(pp->ev_join[d][k]);
        }
        if (pp->ev_fork[d]) hbc_event_destroy(pp->ev_fork[d]);
        if (pp->ev_mc[d]) hbc_event_destroy(pp->ev_mc[d]);
    }
    for (int i = 0; i <= HB_PREPASS_MAX_KERNELS; i++) if (pp->prof_ev[i]) hbc_event_destroy(pp->prof_ev[i]);
    free(pp);
}

/* queue the kernels of one frame on the context's stream; returns a cudaError_t value */
#define PROF_MARK(...) do { if (prof && !crc) { snprintf(pp->prof_name[n], sizeof pp->prof_name[n], __VA_ARGS__); crc = hbc_event_record(pp->prof_ev[n], st); } } while (0)
/* motion compensation of depth d: chroma only when the search kernel already left the luma prediction */
static int enqueue_mc(hb_prepass *pp, const hb_frame *ref, int d, int fused, void *st)
{
    return hbk_mc_predict(&ref->d, &pp->pred[d]->d, 64 >> d, pp->d_pus[d], pp->n_valid[d], pp->d_me[d], fused ? 2 : 3, st);
}

/* T/Q launches of one pass: planes c0..c1 on stream st */
static int enqueue_tq(hb_prepass *pp, const hb_frame *cur, int p, int c0, int c1, void *st, int *n_io, int prof)
{
    hb_ctx *ctx = pp->ctx;
    int crc = 0, n = *n_io;
    const hb_frame *pred = pp->pred[pass_depth(p)];
    for (int c = c0; c <= c1 && !crc; c++) {
        pass_comp *pc = &pp->pc[p][c];
        if (!pc->tu || !pc->n_tus) continue;
        hbd_tq_args a;
        memset(&a, 0, sizeof a);
        hbi_tq_setup(ctx, &a, c, pc->tu, c ? pp->qp_c : pp->cfg.qp, pp->cfg.is_islice, pp->cfg.sign_hiding);
        a.cur = cur->d.p[c]; a.pred = pred->d.p[c]; a.rec = pp->recon[p]->d.p[c];
        a.jobs_xy = pc->d_xy; a.n_jobs = pc->n_tus;
        a.thr_k = 1.; a.weight = c ? pp->weight_c : 1.; a.dyn = pp->d_dyn;
        a.coeff_out = pc->d_coeff; a.res_out = pc->d_res;
        PROF_MARK("tq%d%c%d", p, "yuv"[c], pc->tu);
        if (!crc) crc = hbk_tq_encode(&a, st);
        n++;
    }
    *n_io = n;
    return crc;
}

/* queue the kernels of one frame; returns a cudaError_t value.  prof != 0: everything on the context's stream with an
 * event before each launch.  Otherwise the search chain ME(64) -> ME(32) -> ME(16) -> ME(8) runs on the context's stream
 * and, after each ME(d), side streams take MC(d) and then, in parallel, the luma and chroma T/Q passes of that depth
 * (and the 4x4 luma pass after depth 3), so that they overlap the next search and each other. */
static int enqueue(hb_prepass *pp, const hb_frame *cur, const hb_frame *ref, int *n_launches, int prof)
{
    hb_ctx *ctx = pp->ctx;
    void *main_st = ctx->stream;
    int crc = 0, n = 0;
    /* with sub-pel refinement on, the search kernel leaves the luma prediction of its winner itself; MC then does chroma only */
    const int fused = (pp->cfg.me_action & HB_ME_HALF) != 0;
    /* the reference picture's quarter-pel planes, once per picture: every depth's sub-pel probes and luma predictions read them */
    const hbd_subpel *sp = (fused && ref->sp.base && !pp->cfg.subpel_per_pu) ? &ref->sp : NULL;
    if (sp) {
        void *st = main_st;
        PROF_MARK("sp");
        /* a band needs the planes of its own rows and of the rows its searches reach: +-64 integer, one more above for the negative
         * quarter offsets, the block height below (72 rows on either side cover it) */
        if (!crc) crc = hbk_subpel_planes(&ref->d, sp, pp->row0 * 64 - 72, (pp->row0 + pp->rows) * 64 + 72, main_st);
        n++;
    }
    /* the search: one launch for the whole picture (a CTA per CTU, PU sizes 64 -> 8 in turn), or one launch per PU size */
    const int merged = sp && !pp->cfg.me_staged_window && !pp->cfg.me_per_depth;
    if (merged) {
        void *st = main_st;
        hb_me_result *outs[N_DEPTH];
        const hbd_frame *preds[N_DEPTH];
        for (int d = 0; d < N_DEPTH; d++) { outs[d] = pp->d_me[d]; preds[d] = &pp->pred[d]->d; }
        PROF_MARK("me");
        if (!crc) crc = hbk_me_search_ctus(&cur->d, &ref->d, sp, outs, preds, pp->cfg.me_action, pp->d_dyn, pp->ctu_cols, pp->row0, pp->rows, pp->grid_w, main_st);
        n++;
    }
    for (int d = 0; d < N_DEPTH && !crc; d++) {
        if (!pp->n_valid[d]) continue;
        void *st = main_st;                      /* name used by PROF_MARK */
        if (!merged) {
            PROF_MARK("me%d", 64 >> d);
            const int win = sp && pp->cfg.me_staged_window;       /* windowed kernels read the strip-ordered list */
            if (!crc) crc = hbk_me_search(&cur->d, &ref->d, 64 >> d, win ? pp->d_jobs_strip[d] : pp->d_jobs[d], win ? pp->n_strip[d] : pp->n_valid[d],
                                          d ? pp->d_me[d - 1] : NULL, pp->d_me[d], pp->cfg.me_action, pp->d_dyn, fused ? &pp->pred[d]->d : NULL, sp, win, main_st);
            n++;
        }
        if (prof) {
            PROF_MARK("mc%d", 64 >> d);
            if (!crc) { crc = enqueue_mc(pp, ref, d, fused, st); n++; }
            for (int p = 0; p < N_PASS && !crc; p++)
                if (pass_depth(p) == d) crc = enqueue_tq(pp, cur, p, 0, 2, st, &n, prof);
            continue;
        }
        void *s0 = pp->side[d][0], *s1 = pp->side[d][1], *s2 = pp->side[d][2];
        crc = hbc_event_record(pp->ev_fork[d], main_st);
        if (!crc) crc = hbc_stream_wait_event(s0, pp->ev_fork[d]);
        if (!crc) { crc = enqueue_mc(pp, ref, d, fused, s0); n += fused ? 1 : 2; }   /* chroma kernel (+ luma kernel when not fused) */
        if (!crc) crc = hbc_event_record(pp->ev_mc[d], s0);
        if (!crc) crc = hbc_stream_wait_event(s1, pp->ev_mc[d]);
        if (!crc) crc = enqueue_tq(pp, cur, d, 0, 0, s0, &n, 0);
        if (!crc) crc = enqueue_tq(pp, cur, d, 1, 2, s1, &n, 0);
        if (d == N_DEPTH - 1 && !crc) {
            crc = hbc_stream_wait_event(s2, pp->ev_mc[d]);
            if (!crc) crc = enqueue_tq(pp, cur, N_PASS - 1, 0, 0, s2, &n, 0);
            if (!crc) crc = hbc_event_record(pp->ev_join[d][2], s2);
        }
        if (!crc) crc = hbc_event_record(pp->ev_join[d][0], s0);
        if (!crc) crc = hbc_event_record(pp->ev_join[d][1], s1);
    }
    if (!prof)
        for (int d = 0; d < N_DEPTH && !crc; d++) {
            if (!pp->n_valid[d]) continue;
            crc = hbc_stream_wait_event(main_st, pp->ev_join[d][0]);
            if (!crc) crc = hbc_stream_wait_event(main_st, pp->ev_join[d][1]);
            if (!crc && d == N_DEPTH - 1) crc = hbc_stream_wait_event(main_st, pp->ev_join[d][2]);
        }
    {
        void *st = main_st;
        if (prof && !crc) crc = hbc_event_record(pp->prof_ev[n], st);
    }
    *n_launches = n;
    return crc;
}

int hb_prepass_run(hb_prepass *pp, const hb_frame *cur, const hb_frame *ref, double avg_dist)
{
    int crc = 0, n = 0;
    if (!pp || !cur || !ref) return hbi_fail(HB_ERR_ARG, "hb_prepass_run: NULL argument");
    if (cur->w != pp->w || cur->h != pp->h || ref->w != pp->w || ref->h != pp->h) return hbi_fail(HB_ERR_ARG, "hb_prepass_run: frame size differs from the plan");
    if ((pp->cfg.me_action & HB_ME_HALF) && !pp->cfg.subpel_per_pu && (crc = hbi_frame_subpel_alloc((hb_frame *)ref)) != HB_OK) return crc;   /* a cache inside the frame, outside any capture */
    hb_ctx *ctx = pp->ctx;
    hbc_set_device(ctx->device);
    /* per-frame scalars: same expressions, same host libm as the reference (hmr_common.h:53, hmr_motion_inter.c:106) */
    double w = avg_dist / 2000.;
    w = w < .15 ? .15 : (w > 1.4 ? 1.4 : w);
    /* pageable source on purpose: the runtime stages it before returning, so the block can be rebuilt every frame
     * without waiting for the previous replay */
    hbd_dyn_params dyn;
    dyn.corr = (uint32_t)pp->cfg.qp * w;
    dyn.thr_k = hbi_zero_out_k(avg_dist);
    if ((crc = hbc_h2d_async(pp->d_dyn, &dyn, sizeof dyn, ctx->stream))) return hbi_cuda_fail(crc, "hb_prepass_run: params");

    if (!pp->cfg.use_graph) {
        crc = enqueue(pp, cur, ref, &n, 0);
    } else {
        void *exec = NULL;
        const uint8_t *key[7];
        for (int c = 0; c < 3; c++) { key[c] = cur->d.p[c].base; key[3 + c] = ref->d.p[c].base; }
        key[6] = ref->sp.base;
        for (int i = 0; i < pp->n_graphs; i++) if (!memcmp(pp->graphs[i].plane, key, sizeof key)) exec = pp->graphs[i].exec;
        if (!exec) {
            if (pp->n_graphs == MAX_GRAPHS) { hbc_graph_destroy(pp->graphs[0].exec); memmove(&pp->graphs[0], &pp->graphs[1], sizeof pp->graphs[0] * (MAX_GRAPHS - 1)); pp->n_graphs--; }
            if ((crc = hbc_graph_begin(ctx->stream))) return hbi_cuda_fail(crc, "hb_prepass_run: begin capture");
            crc = enqueue(pp, cur, ref, &n, 0);
            const int erc = hbc_graph_end(ctx->stream, &exec);
            if (crc || erc) return hbi_cuda_fail(crc ? crc : erc, "hb_prepass_run: capture");
            memcpy(pp->graphs[pp->n_graphs].plane, key, sizeof key); pp->graphs[pp->n_graphs].exec = exec;
            pp->n_graphs++;
            pp->launches_per_frame = n;
        }
        n = pp->launches_per_frame;
        crc = hbc_graph_launch(exec, ctx->stream);
    }
    if (crc) return hbi_cuda_fail(crc, "hb_prepass_run");
    pp->launches_per_frame = n;
    ctx->launches += (uint64_t)n;
    return HB_OK;
}

/* same work as hb_prepass_run without the graph, with a CUDA event between consecutive launches: ms[i] is the device
 * time of kernel i (names via hb_prepass_kernel_name).  Synchronises.  Returns the number of kernels or < 0. */
int hb_prepass_run_profiled(hb_prepass *pp, const hb_frame *cur, const hb_frame *ref, double avg_dist, float *ms, int cap)
{
    int crc = 0, n = 0;
    if (!pp || !cur || !ref || !ms) return hbi_fail(HB_ERR_ARG, "hb_prepass_run_profiled: NULL argument");
    hb_ctx *ctx = pp->ctx;
    hbc_set_device(ctx->device);
    for (int i = 0; i <= HB_PREPASS_MAX_KERNELS && !crc; i++) if (!pp->prof_ev[i]) crc = hbc_event_create(&pp->prof_ev[i]);
    if (crc) return hbi_cuda_fail(crc, "hb_prepass_run_profiled: events");
    if ((pp->cfg.me_action & HB_ME_HALF) && !pp->cfg.subpel_per_pu && (crc = hbi_frame_subpel_alloc((hb_frame *)ref)) != HB_OK) return crc;
    double w = avg_dist / 2000.;
    w = w < .15 ? .15 : (w > 1.4 ? 1.4 : w);
    hbd_dyn_params dyn;
    dyn.corr = (uint32_t)pp->cfg.qp * w;
    dyn.thr_k = hbi_zero_out_k(avg_dist);
    if ((crc = hbc_h2d_async(pp->d_dyn, &dyn, sizeof dyn, ctx->stream))) return hbi_cuda_fail(crc, "hb_prepass_run_profiled: params");
    crc = enqueue(pp, cur, ref, &n, 1);
    if (crc) return hbi_cuda_fail(crc, "hb_prepass_run_profiled");
    ctx->launches += (uint64_t)n;
    if (n > cap) return hbi_fail(HB_ERR_ARG, "hb_prepass_run_profiled: %d kernels, room for %d", n, cap);
    for (int i = 0; i < n && !crc; i++) crc = hbc_event_elapsed(pp->prof_ev[i], pp->prof_ev[i + 1], &ms[i]);
    if (crc) return hbi_cuda_fail(crc, "hb_prepass_run_profiled: elapsed");
    return n;
}
const char *hb_prepass_kernel_name(const hb_prepass *pp, int i) { return (pp && i >= 0 && i < HB_PREPASS_MAX_KERNELS) ? pp->prof_name[i] : ""; }

int hb_prepass_num_pus(const hb_prepass *pp, int depth) { return (pp && depth >= 0 && depth < N_DEPTH) ? pp->grid_w[depth] * pp->grid_h[depth] : 0; }
int hb_prepass_num_tus(const hb_prepass *pp, int pass, int comp)
{
    return (pp && pass >= 0 && pass < N_PASS && comp >= 0 && comp < 3) ? pp->pc[pass][comp].n_tus : 0;
}
const hb_frame *hb_prepass_pred(const hb_prepass *pp, int depth) { return (pp && depth >= 0 && depth < N_DEPTH) ? pp->pred[depth] : NULL; }
const hb_frame *hb_prepass_recon(const hb_prepass *pp, int pass) { return (pp && pass >= 0 && pass < N_PASS) ? pp->recon[pass] : NULL; }

static int fetch(hb_prepass *pp, void *dst, const void *dev, size_t bytes, const char *what)
{
    void *h = NULL;
    int rc, crc;
    hb_ctx *ctx = pp->ctx;
    if (!bytes) return HB_OK;
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    if ((rc = hbi_scratch(ctx, 3, bytes, NULL, &h)) != HB_OK) { pthread_mutex_unlock(&ctx->lock); return rc; }
    crc = hbc_d2h_async(h, dev, bytes, ctx->stream);
    if (!crc) crc = hbc_stream_sync(ctx->stream);
    if (!crc) memcpy(dst, h, bytes);
    pthread_mutex_unlock(&ctx->lock);
    return crc ? hbi_cuda_fail(crc, what) : HB_OK;
}

int hb_prepass_fetch_me(hb_prepass *pp, int depth, hb_me_result *out)
{
    if (!pp || !out || depth < 0 || depth >= N_DEPTH) return hbi_fail(HB_ERR_ARG, "hb_prepass_fetch_me: bad argument");
    return fetch(pp, out, pp->d_me[depth], sizeof(hb_me_result) * (size_t)hb_prepass_num_pus(pp, depth), "hb_prepass_fetch_me");
}
int hb_prepass_fetch_tu(hb_prepass *pp, int pass, int comp, hb_tu_result *out)
{
    if (!pp || !out || pass < 0 || pass >= N_PASS || comp < 0 || comp > 2) return hbi_fail(HB_ERR_ARG, "hb_prepass_fetch_tu: bad argument");
    return fetch(pp, out, pp->pc[pass][comp].d_res, sizeof(hb_tu_result) * (size_t)pp->pc[pass][comp].n_tus, "hb_prepass_fetch_tu");
}
int hb_prepass_fetch_coeffs(hb_prepass *pp, int pass, int comp, int16_t *out)
{
    if (!pp || !out || pass < 0 || pass >= N_PASS || comp < 0 || comp > 2) return hbi_fail(HB_ERR_ARG, "hb_prepass_fetch_coeffs: bad argument");
    const pass_comp *pc = &pp->pc[pass][comp];
    return fetch(pp, out, pc->d_coeff, sizeof(int16_t) * (size_t)pc->n_tus * pc->tu * pc->tu, "hb_prepass_fetch_coeffs");
}
int hb_prepass_tu_xy(hb_prepass *pp, int pass, int comp, int32_t *xy_out)
{
    if (!pp || !xy_out || pass < 0 || pass >= N_PASS || comp < 0 || comp > 2) return hbi_fail(HB_ERR_ARG, "hb_prepass_tu_xy: bad argument");
    return fetch(pp, xy_out, pp->pc[pass][comp].d_xy, sizeof(int32_t) * 2 * (size_t)pp->pc[pass][comp].n_tus, "hb_prepass_tu_xy");
}
int hb_prepass_tu_size(const hb_prepass *pp, int pass, int comp)
{
    return (pp && pass >= 0 && pass < N_PASS && comp >= 0 && comp < 3) ? pp->pc[pass][comp].tu : 0;
}
int hb_prepass_fetch_recon(hb_prepass *pp, int pass, uint8_t *y, int ys, uint8_t *u, int us, uint8_t *v, int vs)
{
    if (!pp || pass < 0 || pass >= N_PASS) return hbi_fail(HB_ERR_ARG, "hb_prepass_fetch_recon: bad argument");
    return hb_frame_download_u8(pp->ctx, pp->recon[pass], y, ys, u, us, v, vs);
}

size_t hb_prepass_output_bytes(const hb_prepass *pp)
{
    size_t n = 0;
    if (!pp) return 0;
    for (int d = 0; d < N_DEPTH; d++) n += sizeof(hb_me_result) * (size_t)hb_prepass_num_pus(pp, d);
    for (int p = 0; p < N_PASS; p++) {
        for (int c = 0; c < 3; c++) {
            const pass_comp *pc = &pp->pc[p][c];
            n += sizeof(hb_tu_result) * (size_t)pc->n_tus + sizeof(int16_t) * (size_t)pc->n_tus * pc->tu * pc->tu;
        }
        n += (size_t)pp->w * pp->h * 3 / 2;
    }
    return n;
}

/* everything the host side consumes, packed in the order of hb_prepass_output_bytes: ME tables d0..d3, then per pass
 * {TU results Y,U,V; levels Y,U,V; reconstruction Y,U,V (tight pitch)}.  dst should be pinned. */
int hb_prepass_fetch_all(hb_prepass *pp, void *pinned_dst, size_t cap, size_t *bytes_out)
{
    if (!pp || !pinned_dst) return hbi_fail(HB_ERR_ARG, "hb_prepass_fetch_all: NULL argument");
    const size_t need = hb_prepass_output_bytes(pp);
    if (cap < need) return hbi_fail(HB_ERR_ARG, "hb_prepass_fetch_all: need %zu bytes, got %zu", need, cap);
    hb_ctx *ctx = pp->ctx;
    char *o = (char *)pinned_dst;
    int crc = 0;
    hbc_set_device(ctx->device);
    for (int d = 0; d < N_DEPTH && !crc; d++) {
        const size_t b = sizeof(hb_me_result) * (size_t)hb_prepass_num_pus(pp, d);
        crc = hbc_d2h_async(o, pp->d_me[d], b, ctx->stream); o += b;
    }
    for (int p = 0; p < N_PASS && !crc; p++) {
        for (int c = 0; c < 3 && !crc; c++) {
            const pass_comp *pc = &pp->pc[p][c];
            const size_t b = sizeof(hb_tu_result) * (size_t)pc->n_tus;
            if (b) { crc = hbc_d2h_async(o, pc->d_res, b, ctx->stream); o += b; }
        }
        for (int c = 0; c < 3 && !crc; c++) {
            const pass_comp *pc = &pp->pc[p][c];
            const size_t b = sizeof(int16_t) * (size_t)pc->n_tus * pc->tu * pc->tu;
            if (b) { crc = hbc_d2h_async(o, pc->d_coeff, b, ctx->stream); o += b; }
        }
        for (int c = 0; c < 3 && !crc; c++) {
            const hbd_plane *pl = &pp->recon[p]->d.p[c];
            crc = hbc_d2h_2d_async(o, (size_t)pl->w, pl->org, (size_t)pl->pitch, (size_t)pl->w, (size_t)pl->h, ctx->stream);
            o += (size_t)pl->w * pl->h;
        }
    }
    if (!crc) crc = hbc_stream_sync(ctx->stream);
    if (crc) return hbi_cuda_fail(crc, "hb_prepass_fetch_all");
    if (bytes_out) *bytes_out = (size_t)(o - (char *)pinned_dst);
    return HB_OK;
}

/* ------------------------------------------------------------------ cost tables -> host decision -> gather
 * The host's mode decision reads the cost tables (ME results and per-TU {sum, ssd}), picks a partition depth per CTU,
 * and only then asks for what entropy coding and the in-loop filters need of that choice. */
size_t hb_prepass_tables_bytes(const hb_prepass *pp)
{
    if (!pp) return 0;
    if (pp->cfg.compact_tables == 2) return sizeof(hb_me_result_c) * (size_t)pp->n_me_total + sizeof(hb_cu_cost) * (size_t)pp->n_cu_total;
    if (pp->cfg.compact_tables) return sizeof(hb_me_result_c) * (size_t)pp->n_me_total + sizeof(hb_tu_result_c) * (size_t)pp->n_tu_total;
    return pp->tables_bytes;
}

/* ME tables d0..d3, then TU tables pass 0..4 x (Y,U,V), packed; dst should be pinned.  Asynchronous: hb_ctx_sync before reading.
 * With cfg.compact_tables the same records in their 12-byte forms (one small packing kernel, then one copy). */
int hb_prepass_fetch_tables(hb_prepass *pp, void *pinned_dst, size_t cap)
{
    if (!pp || !pinned_dst) return hbi_fail(HB_ERR_ARG, "hb_prepass_fetch_tables: NULL argument");
    if (cap < hb_prepass_tables_bytes(pp)) return hbi_fail(HB_ERR_ARG, "hb_prepass_fetch_tables: buffer too small");
    hb_ctx *ctx = pp->ctx;
    hbc_set_device(ctx->device);
    int crc;
    if (pp->cfg.compact_tables == 2) {
        hbd_cu_pack_args a;
        memset(&a, 0, sizeof a);
        for (int p = 0; p < N_PASS; p++) {
            for (int c = 0; c < 3; c++) {
                const pass_comp *pc = &pp->pc[p][c];
                a.pc[p][c].tu_index = pc->d_index; a.pc[p][c].grid_w = pc->grid_w; a.pc[p][c].grid_h = pc->grid_h; a.pc[p][c].tu = pc->tu; a.pc[p][c].res = pc->d_res;
            }
            a.first[p + 1] = a.first[p] + pp->grid_w[pass_depth(p)] * pp->grid_h[pass_depth(p)];
        }
        for (int d = 0; d < N_DEPTH; d++) a.grid_w[d] = pp->grid_w[d];
        a.out = (hb_cu_cost *)(pp->d_tables_c + sizeof(hb_me_result_c) * (size_t)pp->n_me_total);
        crc = hbk_pack_tables(pp->d_tables, pp->d_tables_c, pp->n_me_total, 0, ctx->stream);
        if (!crc) crc = hbk_pack_cu_costs(&a, ctx->stream);
        ctx->launches += 2;
        if (!crc) crc = hbc_d2h_async(pinned_dst, pp->d_tables_c, hb_prepass_tables_bytes(pp), ctx->stream);
    } else if (pp->cfg.compact_tables) {
        crc = hbk_pack_tables(pp->d_tables, pp->d_tables_c, pp->n_me_total, pp->n_tu_total, ctx->stream);
        ctx->launches++;
        if (!crc) crc = hbc_d2h_async(pinned_dst, pp->d_tables_c, hb_prepass_tables_bytes(pp), ctx->stream);
    } else
        crc = hbc_d2h_async(pinned_dst, pp->d_tables, pp->tables_bytes, ctx->stream);
    return crc ? hbi_cuda_fail(crc, "hb_prepass_fetch_tables") : HB_OK;
}

int hb_prepass_num_ctus(const hb_prepass *pp) { return pp ? pp->ctu_cols * pp->ctu_rows : 0; }

/* A stand-in for the host's mode decision (which stays on the host): per CTU the pass p in 0..4
 * (PU 64/32/16/8, and 8 with 4x4 luma TUs) that minimises sum(ssd) + lambda * sum(|levels|) over its luma TUs and the chroma
 * TUs of pass min(p,3).  Also lays out the gather stream: ctu_off[i] = start of CTU i's levels (int16 units), ctu_off[n] = total.
 * `tables` is what hb_prepass_fetch_tables delivered.  Pure host code. */
static int popc4(unsigned v) { v &= 15u; return (int)((v & 1u) + ((v >> 1) & 1u) + ((v >> 2) & 1u) + (v >> 3)); }

/* the same choice from the per-CU records (compact_tables = 2): identical costs, the stream lengths come from the coded flags */
static int select_from_cu_costs(const hb_prepass *pp, const void *tables, int lambda, uint8_t *sel, int32_t *ctu_off)
{
    const int n_ctus = hb_prepass_num_ctus(pp);
    const hb_cu_cost *rec = (const hb_cu_cost *)((const char *)tables + sizeof(hb_me_result_c) * (size_t)pp->n_me_total);
    const hb_cu_cost *first[N_PASS];
    for (int p = 0; p < N_PASS; p++) { first[p] = rec; rec += (size_t)pp->grid_w[pass_depth(p)] * pp->grid_h[pass_depth(p)]; }
    int32_t off = 0;
    for (int i = 0; i < n_ctus; i++) {
        uint64_t best = ~(uint64_t)0; int bp = N_DEPTH - 1; int32_t best_len = 0;
        const int cx = i % pp->ctu_cols, cy = i / pp->ctu_cols;
        const int ew = pp->w - cx * 64, eh = pp->h - cy * 64;
        for (int p = 0; p < N_PASS; p++) {
            const int d = pass_depth(p), cu = 64 >> d, per = 64 / cu, gw = pp->grid_w[d];
            const int len_y = 2 + pass_luma_tu(p) * pass_luma_tu(p), tc = pp->pc[d][1].tu, len_c = 2 + tc * tc;
            uint64_t v = 0; int32_t len = 0;
            for (int y = 0; y < per; y++) {
                const hb_cu_cost *row = first[p] + (size_t)(cy * per + y) * gw + cx * per;
                for (int x = 0; x < per; x++) {
                    v += (uint64_t)row[x].ssd + (uint64_t)((int64_t)lambda * (int64_t)row[x].sum);
                    len += popc4(row[x].cbf) * len_y + (popc4(row[x].cbf >> 4) + popc4(row[x].cbf >> 8)) * len_c;
                }
            }
            if ((ew < 64 && ew % cu) || (eh < 64 && eh % cu)) continue;
            if (v < best) { best = v; bp = p; best_len = len; }
        }
        if (best == ~(uint64_t)0) {                      /* cannot happen (8x8 units always tile), keep the layout consistent anyway */
            const int d = N_DEPTH - 1, gw = pp->grid_w[d];
            best_len = 0;
            for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
                const hb_cu_cost *r = first[d] + (size_t)(cy * 8 + y) * gw + cx * 8 + x;
                best_len += popc4(r->cbf) * (2 + 64) + (popc4(r->cbf >> 4) + popc4(r->cbf >> 8)) * (2 + 16);
            }
        }
        sel[i] = (uint8_t)bp;
        ctu_off[i] = off;
        off += best_len;
    }
    ctu_off[n_ctus] = off;
    return HB_OK;
}

int hb_prepass_select(const hb_prepass *pp, const void *tables, int lambda, uint8_t *sel, int32_t *ctu_off)
{
    if (!pp || !tables || !sel || !ctu_off) return hbi_fail(HB_ERR_ARG, "hb_prepass_select: NULL argument");
    const int n_ctus = hb_prepass_num_ctus(pp);
    if (pp->cfg.compact_tables == 2) return select_from_cu_costs(pp, tables, lambda, sel, ctu_off);
    const int compact = pp->cfg.compact_tables != 0;
    const char *t = (const char *)tables + (compact ? sizeof(hb_me_result_c) : sizeof(hb_me_result)) * (size_t)pp->n_me_total;
    const size_t rec = compact ? sizeof(hb_tu_result_c) : sizeof(hb_tu_result);
    const char *res[N_PASS][3];
    for (int p = 0; p < N_PASS; p++) for (int c = 0; c < 3; c++) { res[p][c] = t; t += rec * (size_t)pp->pc[p][c].n_tus; }
    uint64_t *cost = (uint64_t *)calloc((size_t)n_ctus * 8, sizeof *cost);       /* [ctu][0..3 luma+chroma of pass p, 4 luma of pass 4] */
    int32_t *len = (int32_t *)calloc((size_t)n_ctus * 8, sizeof *len);            /* stream length of the same pieces */
    if (!cost || !len) { free(cost); free(len); return hbi_fail(HB_ERR_NOMEM, "hb_prepass_select: out of memory"); }
    for (int p = 0; p < N_PASS; p++)
        for (int c = 0; c < 3; c++) {
            const pass_comp *pc = &pp->pc[p][c];
            const int slot = (p == 4) ? 4 : p, piece = (p == 3 && c > 0) ? 5 : slot;   /* chroma of pass 3 is shared by choices 3 and 4 */
            const int rec_len = 2 + pc->tu * pc->tu;
            const hb_tu_result *rf = (const hb_tu_result *)res[p][c];
            const hb_tu_result_c *rc = (const hb_tu_result_c *)res[p][c];
            const int32_t *ctu_of = pc->h_ctu;
            /* TUs arrive in raster order: runs of consecutive TUs share a CTU, so accumulate in registers and flush per run */
            int i = 0;
            while (i < pc->n_tus) {
                const int ctu = ctu_of[i];
                uint64_t acc = 0; int32_t l = 0;
                for (; i < pc->n_tus && ctu_of[i] == ctu; i++) {
                    const uint32_t ssd = compact ? rc[i].ssd : rf[i].ssd;
                    const int32_t sum = compact ? (int32_t)(rc[i].sum_zeroed & 0x7fffffffu) : rf[i].sum;
                    acc += (uint64_t)ssd + (uint64_t)((int64_t)lambda * sum);
                    l += sum > 0 ? rec_len : 0;
                }
                cost[ctu * 8 + piece] += acc;
                len[ctu * 8 + piece] += l;
            }
        }
    int32_t off = 0;
    for (int i = 0; i < n_ctus; i++) {
        uint64_t best = ~(uint64_t)0; int bp = N_DEPTH - 1;
        /* the part of the CTU inside the picture must be tiled by the pass's units (the reference forces the split at the border) */
        const int ew = pp->w - (i % pp->ctu_cols) * 64, eh = pp->h - (i / pp->ctu_cols) * 64;
        for (int p = 0; p < N_PASS; p++) {
            const int cu = 64 >> pass_depth(p);
            if ((ew < 64 && ew % cu) || (eh < 64 && eh % cu)) continue;
            uint64_t v = cost[i * 8 + (p == 4 ? 4 : p)];
            if (p >= 3) v += cost[i * 8 + 5];
            if (v < best) { best = v; bp = p; }
        }
        sel[i] = (uint8_t)bp;
        ctu_off[i] = off;
        off += len[i * 8 + (bp == 4 ? 4 : bp)] + (bp >= 3 ? len[i * 8 + 5] : 0);
    }
    ctu_off[n_ctus] = off;
    free(cost); free(len);
    return HB_OK;
}

size_t hb_prepass_gather_bytes(const hb_prepass *pp, const int32_t *ctu_off)
{
    if (!pp || !ctu_off) return 0;
    return (size_t)pp->w * pp->h * 3 / 2 + sizeof(int16_t) * (size_t)ctu_off[hb_prepass_num_ctus(pp)];
}

/* queue the gather kernel of the host's choice: reconstruction into out_recon (NULL: the tight fetch buffer), levels into d_sel_levels */
static int gather_queue(hb_prepass *pp, const uint8_t *sel, const int32_t *ctu_off, uint8_t *const out_recon[3], const int out_pitch[3], const char *what)
{
    hb_ctx *ctx = pp->ctx;
    const int n_ctus = hb_prepass_num_ctus(pp);
    const size_t recon_bytes = (size_t)pp->w * pp->h * 3 / 2, lev = (size_t)ctu_off[n_ctus];
    int crc = 0;
    for (int i = 0; i < n_ctus; i++) {
        if (sel[i] > 4) return hbi_fail(HB_ERR_ARG, "%s: sel[%d] = %d", what, i, sel[i]);
        /* the in-picture part of a partial CTU must be tiled by the chosen pass's coding units (hb_prepass_select's rule): anything
         * else would gather samples no transform unit wrote and search records of units that were never searched */
        const int ew = pp->w - (i % pp->ctu_cols) * 64, eh = pp->h - (i / pp->ctu_cols) * 64, cu = 64 >> pass_depth(sel[i]);
        if ((ew < 64 && ew % cu) || (eh < 64 && eh % cu))
            return hbi_fail(HB_ERR_ARG, "%s: sel[%d] = %d, but %dx%d units do not tile the %dx%d picture part of that CTU", what, i, sel[i], cu, cu,
                            ew < 64 ? ew : 64, eh < 64 ? eh : 64);
    }
    hbc_set_device(ctx->device);
    if (!pp->d_sel && (crc = hbc_malloc((void **)&pp->d_sel, (size_t)n_ctus))) { pp->d_sel = NULL; return hbi_cuda_fail(crc, what); }
    if (!pp->d_ctu_off && (crc = hbc_malloc((void **)&pp->d_ctu_off, sizeof(int32_t) * ((size_t)n_ctus + 1)))) { pp->d_ctu_off = NULL; return hbi_cuda_fail(crc, what); }
    if (!pp->d_sel_recon && (crc = hbc_malloc((void **)&pp->d_sel_recon, recon_bytes))) { pp->d_sel_recon = NULL; return hbi_cuda_fail(crc, what); }
    if (!pp->d_sel_levels) {
        /* worst case once (every unit coded with the smallest transform: 18 int16 per 16 samples): growing the buffer later would
         * mean cudaFree / cudaMalloc, which wait for every stream of the device and drain the frames in flight on other contexts */
        pp->sel_levels_cap = (size_t)pp->ctu_cols * pp->ctu_rows * (64 * 64 * 3 / 2) * 18 / 16 + 64;
        if ((crc = hbc_malloc((void **)&pp->d_sel_levels, sizeof(int16_t) * pp->sel_levels_cap))) { pp->sel_levels_cap = 0; pp->d_sel_levels = NULL; return hbi_cuda_fail(crc, what); }
    }
    if (lev > pp->sel_levels_cap) return hbi_fail(HB_ERR_ARG, "%s: ctu_off claims %zu levels, more than the frame can hold", what, lev);
    /* sel / ctu_off are pageable: staged by the runtime before the call returns */
    crc = hbc_h2d_async(pp->d_sel, sel, (size_t)n_ctus, ctx->stream);
    if (!crc) crc = hbc_h2d_async(pp->d_ctu_off, ctu_off, sizeof(int32_t) * ((size_t)n_ctus + 1), ctx->stream);
    hbd_gather_args a;
    memset(&a, 0, sizeof a);
    a.sel = pp->d_sel; a.ctu_off = pp->d_ctu_off; a.ctu_cols = pp->ctu_cols;
    for (int p = 0; p < N_PASS; p++) {
        a.recon[p] = pp->recon[p]->d;
        for (int c = 0; c < 3; c++) {
            const pass_comp *pc = &pp->pc[p][c];
            a.pc[p][c].tu_index = pc->d_index; a.pc[p][c].grid_w = pc->grid_w; a.pc[p][c].grid_h = pc->grid_h; a.pc[p][c].tu = pc->tu;
            a.pc[p][c].res = pc->d_res; a.pc[p][c].coeff = pc->d_coeff;
        }
    }
    if (out_recon) for (int c = 0; c < 3; c++) { a.out_recon[c] = out_recon[c]; a.out_pitch[c] = out_pitch[c]; }
    else {
        a.out_recon[0] = pp->d_sel_recon; a.out_recon[1] = pp->d_sel_recon + (size_t)pp->w * pp->h; a.out_recon[2] = a.out_recon[1] + (size_t)pp->w * pp->h / 4;
        a.out_pitch[0] = pp->w; a.out_pitch[1] = a.out_pitch[2] = pp->w / 2;
    }
    a.out_levels = pp->d_sel_levels;
    if (!crc) { crc = hbk_gather(&a, n_ctus, ctx->stream); ctx->launches++; }
    return crc ? hbi_cuda_fail(crc, what) : HB_OK;
}

/* The levels of the host's choice in the reference's own hand-off layout (ctu->coeff_wnd): per CTU 64*64 luma, 32*32 U, 32*32 V int16,
 * a transform unit's levels row-major at abs_index << 4 (chroma >> 2).  Blocking. */
int hb_prepass_fetch_coeff_wnd(hb_prepass *pp, const uint8_t *sel, int16_t *out)
{
    if (!pp || !sel || !out) return hbi_fail(HB_ERR_ARG, "hb_prepass_fetch_coeff_wnd: NULL argument");
    hb_ctx *ctx = pp->ctx;
    const int n_ctus = hb_prepass_num_ctus(pp);
    const size_t bytes = sizeof(int16_t) * HB_COEFF_WND_PER_CTU * (size_t)n_ctus;
    int crc = 0, rc;
    void *d_out, *h_out;
    for (int i = 0; i < n_ctus; i++) if (sel[i] > 4) return hbi_fail(HB_ERR_ARG, "hb_prepass_fetch_coeff_wnd: sel[%d] = %d", i, sel[i]);
    hbc_set_device(ctx->device);
    if (!pp->d_sel && (crc = hbc_malloc((void **)&pp->d_sel, (size_t)n_ctus))) { pp->d_sel = NULL; return hbi_cuda_fail(crc, "hb_prepass_fetch_coeff_wnd"); }
    pthread_mutex_lock(&ctx->lock);
    if ((rc = hbi_scratch(ctx, 3, bytes, &d_out, &h_out)) != HB_OK) { pthread_mutex_unlock(&ctx->lock); return rc; }
    crc = hbc_h2d_async(pp->d_sel, sel, (size_t)n_ctus, ctx->stream);
    hbd_gather_args a;
    memset(&a, 0, sizeof a);
    a.sel = pp->d_sel; a.ctu_cols = pp->ctu_cols;
    for (int p = 0; p < N_PASS; p++)
        for (int c = 0; c < 3; c++) {
            const pass_comp *pc = &pp->pc[p][c];
            a.pc[p][c].tu_index = pc->d_index; a.pc[p][c].grid_w = pc->grid_w; a.pc[p][c].grid_h = pc->grid_h; a.pc[p][c].tu = pc->tu;
            a.pc[p][c].res = pc->d_res; a.pc[p][c].coeff = pc->d_coeff;
        }
    if (!crc) { crc = hbk_coeff_wnd(&a, n_ctus, (int16_t *)d_out, ctx->stream); ctx->launches++; }
    if (!crc) crc = hbc_d2h_async(h_out, d_out, bytes, ctx->stream);
    if (!crc) crc = hbc_stream_sync(ctx->stream);
    if (!crc) memcpy(out, h_out, bytes);
    pthread_mutex_unlock(&ctx->lock);
    return crc ? hbi_cuda_fail(crc, "hb_prepass_fetch_coeff_wnd") : HB_OK;
}

/* Queue the gather of the host's choice and its copy to pinned_dst: reconstruction Y,U,V (tight planes) followed by the level
 * streams of all CTUs (layout in hb_kernels_gather.cu).  Asynchronous: hb_ctx_sync before reading. */
int hb_prepass_gather(hb_prepass *pp, const uint8_t *sel, const int32_t *ctu_off, void *pinned_dst, size_t cap, size_t *bytes_out)
{
    if (!pp || !sel || !ctu_off || !pinned_dst) return hbi_fail(HB_ERR_ARG, "hb_prepass_gather: NULL argument");
    hb_ctx *ctx = pp->ctx;
    const int n_ctus = hb_prepass_num_ctus(pp);
    const size_t recon_bytes = (size_t)pp->w * pp->h * 3 / 2, lev = (size_t)ctu_off[n_ctus];
    const size_t need = recon_bytes + sizeof(int16_t) * lev;
    int rc, crc;
    if (cap < need) return hbi_fail(HB_ERR_ARG, "hb_prepass_gather: need %zu bytes, got %zu", need, cap);
    if ((rc = gather_queue(pp, sel, ctu_off, NULL, NULL, "hb_prepass_gather")) != HB_OK) return rc;
    crc = hbc_d2h_async(pinned_dst, pp->d_sel_recon, recon_bytes, ctx->stream);
    if (!crc && lev) crc = hbc_d2h_async((char *)pinned_dst + recon_bytes, pp->d_sel_levels, sizeof(int16_t) * lev, ctx->stream);
    if (crc) return hbi_cuda_fail(crc, "hb_prepass_gather");
    if (bytes_out) *bytes_out = need;
    return HB_OK;
}

/* The device-resident continuation of the host's choice: the reconstruction of the chosen passes goes
 * straight into `rec` (no trip to the host), the level streams of the coded TUs go to pinned_levels as in hb_prepass_gather, and
 * `rec` is deblocked in place with strengths derived on the device from the pre-pass's own tables (CU / TU sizes of the pass,
 * vectors, coded flags), then border-padded.  Asynchronous: hb_ctx_sync before reading pinned_levels.  What remains for the next
 * reference picture is SAO: hb_sao_stats_frame(cur, rec) -> the host's decision -> hb_sao_apply_frame(rec, next reference). */
static int finalise_queue(hb_prepass *pp, const uint8_t *sel, const int32_t *ctu_off, hb_frame *rec, const hb_deblock_params *dbk,
                          void *pinned_levels, size_t cap, size_t *bytes_out, int pad)
{
    if (!pp || !sel || !ctu_off || !rec || !dbk || !pinned_levels) return hbi_fail(HB_ERR_ARG, "hb_prepass_finalise: NULL argument");
    if (rec->w != pp->w || rec->h != pp->h) return hbi_fail(HB_ERR_ARG, "hb_prepass_finalise: frame size differs from the plan's");
    if (pp->rows != pp->ctu_rows) return hbi_fail(HB_ERR_ARG, "hb_prepass_finalise: not available for a CTU-row band");
    hb_ctx *ctx = pp->ctx;
    const int n_ctus = hb_prepass_num_ctus(pp);
    const size_t lev = (size_t)ctu_off[n_ctus], need = sizeof(int16_t) * lev;
    const int uw = pp->w / 4, uh = pp->h / 4;
    const size_t plane = (size_t)uw * uh;
    int rc, crc = 0;
    if (cap < need) return hbi_fail(HB_ERR_ARG, "hb_prepass_finalise: need %zu bytes, got %zu", need, cap);
    hbc_set_device(ctx->device);
    if (!pp->d_units && (crc = hbc_malloc((void **)&pp->d_units, sizeof(hb_unit_info) * plane))) { pp->d_units = NULL; return hbi_cuda_fail(crc, "hb_prepass_finalise: cudaMalloc"); }
    if (!pp->d_dbk_maps && (crc = hbc_malloc((void **)&pp->d_dbk_maps, 3 * plane))) { pp->d_dbk_maps = NULL; return hbi_cuda_fail(crc, "hb_prepass_finalise: cudaMalloc"); }
    uint8_t *planes[3] = { rec->d.p[0].org, rec->d.p[1].org, rec->d.p[2].org };
    const int pitch[3] = { rec->d.p[0].pitch, rec->d.p[1].pitch, rec->d.p[2].pitch };
    if ((rc = gather_queue(pp, sel, ctu_off, planes, pitch, "hb_prepass_finalise")) != HB_OK) return rc;
    if (lev) crc = hbc_d2h_async(pinned_levels, pp->d_sel_levels, need, ctx->stream);
    hbd_units_args u;
    memset(&u, 0, sizeof u);
    u.sel = pp->d_sel; u.ctu_cols = pp->ctu_cols; u.uw = uw; u.uh = uh; u.units_w = uw; u.qp = pp->cfg.qp;
    for (int d = 0; d < N_DEPTH; d++) { u.me[d] = pp->d_me[d]; u.me_grid_w[d] = pp->grid_w[d]; }
    for (int p = 0; p < N_PASS; p++) {
        const pass_comp *pc = &pp->pc[p][0];
        u.luma[p].tu_index = pc->d_index; u.luma[p].grid_w = pc->grid_w; u.luma[p].grid_h = pc->grid_h; u.luma[p].tu = pc->tu; u.luma[p].res = pc->d_res;
    }
    u.units = pp->d_units;
    uint8_t *bv = pp->d_dbk_maps, *bh = bv + plane, *qp = bh + plane;
    if (!crc) { crc = hbk_units_from_selection(&u, ctx->stream); ctx->launches++; }
    if (!crc) { crc = hbk_deblock_strengths(pp->d_units, uw, pp->w, pp->h, bv, bh, qp, ctx->stream); ctx->launches++; }
    if (!crc) { crc = hbk_deblock(&rec->d, bv, bh, qp, uw, dbk->cb_qp_offset, dbk->cr_qp_offset, dbk->beta_offset_div2, dbk->tc_offset_div2, ctx->stream); ctx->launches += 2; }
    if (!crc && pad) { crc = hbk_pad_frame(&rec->d, ctx->stream); ctx->launches++; }
    if (crc) return hbi_cuda_fail(crc, "hb_prepass_finalise");
    if (bytes_out) *bytes_out = need;
    return HB_OK;
}

int hb_prepass_finalise(hb_prepass *pp, const uint8_t *sel, const int32_t *ctu_off, hb_frame *rec, const hb_deblock_params *dbk,
                        void *pinned_levels, size_t cap, size_t *bytes_out)
{
    return finalise_queue(pp, sel, ctu_off, rec, dbk, pinned_levels, cap, bytes_out, 1);
}

/* debugging / testing aid: the unit data and strengths the last hb_prepass_finalise used (each may be NULL).  Blocking. */
int hb_prepass_fetch_units(hb_prepass *pp, hb_unit_info *units, uint8_t *bs_ver, uint8_t *bs_hor)
{
    if (!pp || !pp->d_units) return hbi_fail(HB_ERR_ARG, "hb_prepass_fetch_units: hb_prepass_finalise has not run");
    hb_ctx *ctx = pp->ctx;
    const size_t plane = (size_t)(pp->w / 4) * (pp->h / 4);
    int crc = 0;
    hbc_set_device(ctx->device);
    if (units) crc = hbc_d2h_async(units, pp->d_units, sizeof(hb_unit_info) * plane, ctx->stream);
    if (!crc && bs_ver) crc = hbc_d2h_async(bs_ver, pp->d_dbk_maps, plane, ctx->stream);
    if (!crc && bs_hor) crc = hbc_d2h_async(bs_hor, pp->d_dbk_maps + plane, plane, ctx->stream);
    if (!crc) crc = hbc_stream_sync(ctx->stream);
    return crc ? hbi_cuda_fail(crc, "hb_prepass_fetch_units") : HB_OK;
}

/* The whole per-frame host flow in one blocking call (what one encoder thread does per frame): upload cur and ref from
 * (pinned) host planes, run the pre-pass, fetch the cost tables, let the stand-in decision pick a depth per CTU, gather and
 * fetch the reconstruction + coded levels of that choice.  tables / out must be pinned; sel has num_ctus bytes, ctu_off
 * num_ctus + 1 entries.  Tight plane pitches (width, width/2). */
/* The same flow in two halves, so that ONE host thread can keep two frames in flight (begin frame n+1 on a second plan /
 * context before finishing frame n): begin queues the uploads, the pre-pass and the table fetch and returns at once;
 * finish waits for the tables, decides, gathers and waits for the results. */
int hb_prepass_frame_begin(hb_prepass *pp, hb_frame *cur, hb_frame *ref, const uint8_t *const cur_planes[3], const uint8_t *const ref_planes[3],
                           double avg_dist, void *tables, size_t tables_cap)
{
    int rc;
    if (!pp || !cur || !ref || !cur_planes || !ref_planes) return hbi_fail(HB_ERR_ARG, "hb_prepass_frame_begin: NULL argument");
    hb_ctx *ctx = pp->ctx;
    const int w = pp->w;
    if ((rc = hb_frame_upload_u8_ex(ctx, cur, cur_planes[0], w, cur_planes[1], w / 2, cur_planes[2], w / 2, HB_UPLOAD_NO_BORDER)) != HB_OK) return rc;
    if ((rc = hb_frame_upload_u8(ctx, ref, ref_planes[0], w, ref_planes[1], w / 2, ref_planes[2], w / 2)) != HB_OK) return rc;
    if ((rc = hb_prepass_run(pp, cur, ref, avg_dist)) != HB_OK) return rc;
    return hb_prepass_fetch_tables(pp, tables, tables_cap);
}

int hb_prepass_frame_finish(hb_prepass *pp, int lambda, const void *tables, uint8_t *sel, int32_t *ctu_off, void *out, size_t out_cap, size_t *out_bytes)
{
    int rc;
    if (!pp) return hbi_fail(HB_ERR_ARG, "hb_prepass_frame_finish: NULL argument");
    hb_ctx *ctx = pp->ctx;
    if ((rc = hb_ctx_sync(ctx)) != HB_OK) return rc;
    if ((rc = hb_prepass_select(pp, tables, lambda, sel, ctu_off)) != HB_OK) return rc;
    if ((rc = hb_prepass_gather(pp, sel, ctu_off, out, out_cap, out_bytes)) != HB_OK) return rc;
    return hb_ctx_sync(ctx);
}

int hb_prepass_process_frame(hb_prepass *pp, hb_frame *cur, hb_frame *ref, const uint8_t *const cur_planes[3], const uint8_t *const ref_planes[3],
                             double avg_dist, int lambda, void *tables, size_t tables_cap, uint8_t *sel, int32_t *ctu_off,
                             void *out, size_t out_cap, size_t *out_bytes)
{
    const int rc = hb_prepass_frame_begin(pp, cur, ref, cur_planes, ref_planes, avg_dist, tables, tables_cap);
    if (rc != HB_OK) return rc;
    return hb_prepass_frame_finish(pp, lambda, tables, sel, ctu_off, out, out_cap, out_bytes);
}

/* The per-frame flow with the reference picture kept on the device (SURVEY.md 8f item 4 closed into a loop): only the source
 * picture goes up and only the cost tables, the level streams and the SAO statistics come down.  begin queues the upload of
 * `cur`, the pre-pass against `ref` (a frame a previous finish produced, or an uploaded one) and the table fetch, and returns. */
int hb_prepass_frame_begin_resident(hb_prepass *pp, hb_frame *cur, hb_frame *ref, const uint8_t *const cur_planes[3], double avg_dist,
                                    void *tables, size_t tables_cap)
{
    int rc;
    if (!pp || !cur || !ref || !cur_planes) return hbi_fail(HB_ERR_ARG, "hb_prepass_frame_begin_resident: NULL argument");
    hb_ctx *ctx = pp->ctx;
    const int w = pp->w;
    if ((rc = hb_frame_upload_u8_ex(ctx, cur, cur_planes[0], w, cur_planes[1], w / 2, cur_planes[2], w / 2, HB_UPLOAD_NO_BORDER)) != HB_OK) return rc;
    if ((rc = hb_prepass_run(pp, cur, ref, avg_dist)) != HB_OK) return rc;
    return hb_prepass_fetch_tables(pp, tables, tables_cap);
}

/* finish waits for the tables, lets the stand-in decision pick a pass per CTU, gathers that choice into `rec` and deblocks it
 * (hb_prepass_finalise), derives the SAO candidates of `rec` against `cur` on the device, fetches them (one wait), runs the
 * stand-in SAO decision and QUEUES the offset pass into `next_ref`, border included -- it does not wait for it: work queued on
 * the same context afterwards (the next frame's begin) is ordered behind it; hb_ctx_sync before reading next_ref from elsewhere.
 * levels must be pinned and are complete on return; params_out (optional, num_ctus records) receives the SAO decision. */
int hb_prepass_frame_finish_resident(hb_prepass *pp, const hb_frame *cur, int lambda, const void *tables, uint8_t *sel, int32_t *ctu_off,
                                     hb_frame *rec, hb_frame *next_ref, const hb_deblock_params *dbk, const double sao_lambda[3],
                                     void *levels, size_t levels_cap, size_t *levels_bytes, hb_sao_param *params_out)
{
    int rc, crc = 0;
    if (!pp || !cur || !rec || !next_ref || !sao_lambda) return hbi_fail(HB_ERR_ARG, "hb_prepass_frame_finish_resident: NULL argument");
    if (next_ref == rec || next_ref->w != pp->w || next_ref->h != pp->h || cur->w != pp->w || cur->h != pp->h)
        return hbi_fail(HB_ERR_ARG, "hb_prepass_frame_finish_resident: next_ref must be another frame of the plan's size");
    hb_ctx *ctx = pp->ctx;
    const int n_ctus = hb_prepass_num_ctus(pp);
    const size_t st_bytes = sizeof(hb_sao_stats) * 3 * (size_t)n_ctus, cand_bytes = sizeof(hb_sao_candidate) * 15 * (size_t)n_ctus;
    const size_t prm_bytes = sizeof(hb_sao_param) * (size_t)n_ctus;
    hbc_set_device(ctx->device);
    if (!pp->d_sao && (crc = hbc_malloc((void **)&pp->d_sao, st_bytes + cand_bytes + prm_bytes))) { pp->d_sao = NULL; return hbi_cuda_fail(crc, "hb_prepass_frame_finish_resident: cudaMalloc"); }
    if (!pp->h_sao && (crc = hbc_host_alloc((void **)&pp->h_sao, cand_bytes + prm_bytes))) { pp->h_sao = NULL; return hbi_cuda_fail(crc, "hb_prepass_frame_finish_resident: cudaHostAlloc"); }
    hb_sao_stats *d_st = (hb_sao_stats *)pp->d_sao;
    hb_sao_candidate *d_cand = (hb_sao_candidate *)(pp->d_sao + st_bytes), *h_cand = (hb_sao_candidate *)pp->h_sao;
    hb_sao_param *d_prm = (hb_sao_param *)(pp->d_sao + st_bytes + cand_bytes), *h_prm = (hb_sao_param *)(pp->h_sao + cand_bytes);
    if ((rc = hb_ctx_sync(ctx)) != HB_OK) return rc;
    if ((rc = hb_prepass_select(pp, tables, lambda, sel, ctu_off)) != HB_OK) return rc;
    /* no border for `rec`: the SAO kernels never classify a sample whose neighbours lie outside the picture */
    if ((rc = finalise_queue(pp, sel, ctu_off, rec, dbk, levels, levels_cap, levels_bytes, 0)) != HB_OK) return rc;
    crc = hbk_sao_stats(&cur->d, &rec->d, pp->ctu_cols, n_ctus, d_st, ctx->stream); ctx->launches++;
    if (!crc) { crc = hbk_sao_derive(d_st, 3 * n_ctus, sao_lambda, d_cand, ctx->stream); ctx->launches++; }
    if (!crc) crc = hbc_d2h_async(h_cand, d_cand, cand_bytes, ctx->stream);
    if (!crc) crc = hbc_stream_sync(ctx->stream);
    if (crc) return hbi_cuda_fail(crc, "hb_prepass_frame_finish_resident");
    if ((rc = hb_sao_decide_from_candidates(h_cand, n_ctus, sao_lambda, h_prm)) != HB_OK) return rc;
    crc = hbc_h2d_async(d_prm, h_prm, prm_bytes, ctx->stream);
    if (!crc) { crc = hbk_sao_apply(&rec->d, &next_ref->d, pp->ctu_cols, n_ctus, d_prm, ctx->stream); ctx->launches++; }
    if (!crc) { crc = hbk_pad_frame(&next_ref->d, ctx->stream); ctx->launches++; }
    if (crc) return hbi_cuda_fail(crc, "hb_prepass_frame_finish_resident");
    if (params_out) memcpy(params_out, h_prm, prm_bytes);
    return HB_OK;
}
