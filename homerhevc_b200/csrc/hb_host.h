/* hb_host.h -- internal definitions of the C99 host layer (hb_host.c, hb_percall.c, hb_prepass.c). */
#ifndef HB_HOST_H
#define HB_HOST_H

#include <pthread.h>
#include "hb_shim.h"

#define HB_SCAN_HOR  1      /* hmr_private.h:92-96 */
#define HB_SCAN_VER  2
#define HB_SCAN_DIAG 3
#define HB_N_SCRATCH 8

struct hb_ctx {
    int device;
    void *stream;
    void *ev[2];
    void *ev_sync;                    /* hb_ctx_wait */
    pthread_mutex_t lock;             /* serialises users of the scratch buffers */
    uint16_t *d_scan;                 /* device tables, see hb_tab_*_off */
    int32_t *d_q, *d_dq;
    uint32_t *d_flag;
    void *d_scratch[HB_N_SCRATCH], *h_scratch[HB_N_SCRATCH];
    size_t scratch_bytes[HB_N_SCRATCH];
    uint64_t launches;
};

struct hb_frame {
    hb_ctx *ctx;
    int w, h;
    hbd_frame d;
    uint8_t *stage;            /* dense device copy of the last uploaded planes (allocated on first upload) */
    hbd_subpel sp;             /* quarter-pel planes of the luma (allocated the first time the frame is a pre-pass reference) */
    int ipc_view;              /* the planes belong to another process (hb_frame_ipc_open): closed, not freed */
};

hb_ctx *hb_default_ctx(void);
int hbi_fail(int code, const char *fmt, ...);
int hbi_cuda_fail(int cuda_code, const char *what);
int hbi_scratch(hb_ctx *ctx, int i, size_t bytes, void **dev, void **host);
size_t hbi_tab_scan_off(int mode, int lg);
size_t hbi_tab_q_off(int lg, int list, int rem);
int hbi_chroma_qp(int qp, int offset);
double hbi_zero_out_k(double avg_dist);
void hbi_tq_setup(hb_ctx *ctx, hbd_tq_args *a, int comp, int n, int qp, int is_islice, int sign_hiding);
/* the queueing halves of hb_mc_predict / hb_tq_encode: validate, pack, launch; the caller holds ctx->lock, synchronises and
 * (hbi_tq_collect) spreads the packed results.  hb_encode.c appends its block read-back before the one wait. */
int hbi_mc_predict_queue(hb_ctx *ctx, const hb_frame *ref, hb_frame *pred, const hb_mc_job *jobs, int n_jobs, const char *what);
typedef struct hbi_tq_pack { int *order; size_t *coeff_off; size_t total; void *h_co, *h_rs; } hbi_tq_pack;
int hbi_tq_encode_queue(hb_ctx *ctx, const hb_frame *cur, const hb_frame *pred, hb_frame *recon, const hb_tu_job *jobs, int n_jobs,
                        const hb_tq_params *params, hbi_tq_pack *pk, const char *what);
int hbi_tq_encode_queue_at(hb_ctx *ctx, const hb_frame *cur, const hb_frame *pred, hb_frame *recon, const hb_tu_job *jobs, int n_jobs,
                           const hb_tq_params *params, hbi_tq_pack *pk, const char *what, int scratch_base);
void hbi_tq_collect(const hbi_tq_pack *pk, const hb_tu_job *jobs, int n_jobs, int16_t *coeffs, hb_tu_result *results);
void hbi_tq_pack_free(hbi_tq_pack *pk);
/* storage for the quarter-pel planes of a frame (never inside a stream capture: it allocates) */
int hbi_frame_subpel_alloc(hb_frame *f);

#endif
