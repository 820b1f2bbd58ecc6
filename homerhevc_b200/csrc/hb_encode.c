/*
 * hb_encode.c -- section E of include/homer_b200.h: the batched job API at the granularity of the reference's own call sites,
 * for an encoder loop that keeps every decision on the host (hmr_motion_inter.c: hmr_cu_motion_estimation :2625,
 * predict_inter :3047-3049, check_rd_cost_merge_2nx2n :3655-3657, encode_inter :3165-3170).
 *
 * A session owns four resident pictures -- source, reference, prediction, reconstruction -- and every call is ONE round trip:
 * jobs up, kernels, the blocks the host loop goes on reading (the reference keeps them in int16 CTU windows) down, one wait.
 * The prediction a unit's T/Q jobs subtract is whatever hb_enc_predict last left at that place of the session's prediction
 * picture, exactly as encode_inter_cu reads what hmr_motion_compensation_* last left in et->prediction_wnd[0].
 * C99 over the shim; no CPU path.
 */
#define _POSIX_C_SOURCE 200809L
#include "hb_host.h"

#include <stdlib.h>
#include <string.h>

struct hb_enc {
    hb_ctx *ctx;
    int w, h;
    hb_frame *f[4];         /* HB_ENC_CUR, _REF, _PRED, _RECON */
};

int hb_enc_create(hb_ctx *ctx, int width, int height, hb_enc **out)
{
    int rc = HB_OK;
    if (!ctx || !out) return hbi_fail(HB_ERR_ARG, "hb_enc_create: NULL argument");
    *out = NULL;
    hb_enc *e = (hb_enc *)calloc(1, sizeof *e);
    if (!e) return hbi_fail(HB_ERR_NOMEM, "hb_enc_create: out of memory");
    e->ctx = ctx; e->w = width; e->h = height;
    for (int i = 0; i < 4 && rc == HB_OK; i++) rc = hb_frame_create(ctx, width, height, &e->f[i]);
    if (rc != HB_OK) { hb_enc_destroy(e); return rc; }
    *out = e;
    return HB_OK;
}

void hb_enc_destroy(hb_enc *e)
{
    if (!e) return;
    for (int i = 0; i < 4; i++) hb_frame_destroy(e->f[i]);
    free(e);
}

hb_frame *hb_enc_frame(hb_enc *e, int which) { return (e && which >= 0 && which < 4) ? e->f[which] : NULL; }

int hb_enc_me(hb_enc *e, const hb_me_job *jobs, int n_jobs, double avg_dist, int action, hb_me_result *results)
{
    if (!e) return hbi_fail(HB_ERR_ARG, "hb_enc_me: NULL session");
    return hb_me_search(e->ctx, e->f[HB_ENC_CUR], e->f[HB_ENC_REF], jobs, n_jobs, NULL, 0, avg_dist, action, results);
}

/* queue the read-back of `n` blocks of frame f (descriptors in fj, pinned) into scratch 5 and its pinned twin */
static int fetch_queue(hb_ctx *ctx, const hb_frame *f, const hbd_fetch_job *fj, int n, size_t total, void **h_out, const char *what)
{
    void *d_fj, *h_fj, *d_out;
    int rc, crc;
    if ((rc = hbi_scratch(ctx, 4, sizeof(hbd_fetch_job) * (size_t)n, &d_fj, &h_fj)) != HB_OK) return rc;
    if ((rc = hbi_scratch(ctx, 5, sizeof(int16_t) * total, &d_out, h_out)) != HB_OK) return rc;
    memcpy(h_fj, fj, sizeof(hbd_fetch_job) * (size_t)n);
    crc = hbc_h2d_async(d_fj, h_fj, sizeof(hbd_fetch_job) * (size_t)n, ctx->stream);
    if (!crc) { crc = hbk_fetch_blocks(&f->d, (const hbd_fetch_job *)d_fj, n, (int16_t *)d_out, ctx->stream); ctx->launches++; }
    if (!crc) crc = hbc_d2h_async(*h_out, d_out, sizeof(int16_t) * total, ctx->stream);
    return crc ? hbi_cuda_fail(crc, what) : HB_OK;
}

/* hmr_motion_compensation_luma + 2 x _chroma per job (uni-prediction) into the session's prediction picture, and the same
 * samples back as int16: per job size^2 luma, (size/2)^2 U, (size/2)^2 V, jobs back to back */
int hb_enc_predict(hb_enc *e, const hb_mc_job *jobs, int n_jobs, int16_t *blocks)
{
    int rc, crc = 0;
    if (!e || !jobs || !blocks || n_jobs < 0) return hbi_fail(HB_ERR_ARG, "hb_enc_predict: bad argument");
    if (n_jobs == 0) return HB_OK;
    hb_ctx *ctx = e->ctx;
    hbd_fetch_job *fj = (hbd_fetch_job *)malloc(sizeof *fj * 3 * (size_t)n_jobs);
    if (!fj) return hbi_fail(HB_ERR_NOMEM, "hb_enc_predict: out of memory");
    size_t total = 0;
    for (int i = 0; i < n_jobs; i++)
        for (int c = 0; c < 3; c++) {
            hbd_fetch_job *q = &fj[3 * i + c];
            q->comp = c; q->x = c ? jobs[i].x / 2 : jobs[i].x; q->y = c ? jobs[i].y / 2 : jobs[i].y; q->size = c ? jobs[i].size / 2 : jobs[i].size;
            q->off = (int32_t)total; q->pad_ = 0;
            total += (size_t)(q->size > 0 ? q->size * q->size : 0);
        }
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    void *h_out = NULL;
    rc = hbi_mc_predict_queue(ctx, e->f[HB_ENC_REF], e->f[HB_ENC_PRED], jobs, n_jobs, "hb_enc_predict");   /* validates the jobs */
    if (rc == HB_OK) rc = fetch_queue(ctx, e->f[HB_ENC_PRED], fj, 3 * n_jobs, total, &h_out, "hb_enc_predict");
    if (rc == HB_OK) {
        crc = hbc_stream_sync(ctx->stream);
        if (!crc) memcpy(blocks, h_out, sizeof(int16_t) * total);
    }
    pthread_mutex_unlock(&ctx->lock);
    free(fj);
    if (crc) return hbi_cuda_fail(crc, "hb_enc_predict");
    return rc;
}

/* encode_inter_cu (comp 0) / encode_inter_cu_chroma (comp 1, 2) per job on the session's source and prediction pictures;
 * levels and the decoded samples (int16, size^2 per job, same offsets as the levels) come back with the records */
int hb_enc_tq(hb_enc *e, const hb_tu_job *jobs, int n_jobs, const hb_tq_params *params, int16_t *coeffs, int16_t *decoded, hb_tu_result *results)
{
    int rc, crc = 0;
    hbi_tq_pack pk;
    if (!e || !jobs || !params || !coeffs || !decoded || !results || n_jobs < 0) return hbi_fail(HB_ERR_ARG, "hb_enc_tq: bad argument");
    if (n_jobs == 0) return HB_OK;
    hb_ctx *ctx = e->ctx;
    hbd_fetch_job *fj = (hbd_fetch_job *)malloc(sizeof *fj * (size_t)n_jobs);
    if (!fj) return hbi_fail(HB_ERR_NOMEM, "hb_enc_tq: out of memory");
    size_t total = 0;
    for (int i = 0; i < n_jobs; i++) {
        fj[i].comp = jobs[i].comp; fj[i].x = jobs[i].x; fj[i].y = jobs[i].y; fj[i].size = jobs[i].size; fj[i].off = (int32_t)total; fj[i].pad_ = 0;
        total += (size_t)(jobs[i].size > 0 ? jobs[i].size * jobs[i].size : 0);
    }
    hbc_set_device(ctx->device);
    pthread_mutex_lock(&ctx->lock);
    void *h_out = NULL;
    rc = hbi_tq_encode_queue(ctx, e->f[HB_ENC_CUR], e->f[HB_ENC_PRED], e->f[HB_ENC_RECON], jobs, n_jobs, params, &pk, "hb_enc_tq");
    if (rc == HB_OK) {
        rc = fetch_queue(ctx, e->f[HB_ENC_RECON], fj, n_jobs, total, &h_out, "hb_enc_tq");
        if (rc == HB_OK) {
            crc = hbc_stream_sync(ctx->stream);
            if (!crc) { hbi_tq_collect(&pk, jobs, n_jobs, coeffs, results); memcpy(decoded, h_out, sizeof(int16_t) * total); }
        }
        hbi_tq_pack_free(&pk);
    }
    pthread_mutex_unlock(&ctx->lock);
    free(fj);
    if (crc) return hbi_cuda_fail(crc, "hb_enc_tq");
    return rc;
}
