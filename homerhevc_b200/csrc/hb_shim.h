/*
 * hb_shim.h -- the thin C ABI between the C99 host layer (hb_host.c) and CUDA (hb_*.cu).
 * Internal: not installed, not part of include/.  Everything here is plain C.
 * Every launcher queues work on `stream` and returns 0 or a cudaError_t value.
 */
#ifndef HB_SHIM_H
#define HB_SHIM_H

#include <stddef.h>
#include <stdint.h>
#include "../../include/homer_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- device-side picture: three 8-bit planes, each with its own replicated border */
typedef struct hbd_plane {
    uint8_t *base;      /* allocation start */
    uint8_t *org;       /* first real sample */
    int32_t pitch;      /* bytes per row */
    int32_t w, h;       /* real samples */
    int32_t pad;        /* border on each side */
} hbd_plane;
typedef struct hbd_frame { hbd_plane p[3]; } hbd_frame;
#ifdef __CUDACC__
/* plane c of a by-value kernel argument through selects (indexing the parameter with a run-time value makes a local copy) */
__device__ __forceinline__ hbd_plane hbd_pick_plane(const hbd_frame &f, int c) { return c == 0 ? f.p[0] : (c == 1 ? f.p[1] : f.p[2]); }
#endif

/* ---- CUDA runtime wrappers */
int  hbc_device_count(void);
int  hbc_set_device(int dev);
int  hbc_stream_create(void **stream);
int  hbc_stream_destroy(void *stream);
int  hbc_stream_sync(void *stream);
int  hbc_malloc(void **p, size_t bytes);
int  hbc_free(void *p);
int  hbc_host_alloc(void **p, size_t bytes);          /* pinned + mapped */
int  hbc_host_free(void *p);
int  hbc_host_devptr(void *host, void **dev);         /* device alias of mapped pinned memory */
int  hbc_memset_async(void *p, int v, size_t bytes, void *stream);
int  hbc_h2d_async(void *dst, const void *src, size_t bytes, void *stream);
int  hbc_d2h_async(void *dst, const void *src, size_t bytes, void *stream);
int  hbc_d2d_2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, void *stream);
int  hbc_h2d_2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, void *stream);
int  hbc_d2h_2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, void *stream);
int  hbc_event_create(void **ev);
int  hbc_event_destroy(void *ev);
int  hbc_event_record(void *ev, void *stream);
int  hbc_event_elapsed(void *a, void *b, float *ms);   /* synchronises on b */
int  hbc_event_create_notiming(void **ev);
int  hbc_stream_wait_event(void *stream, void *ev);
int  hbc_graph_begin(void *stream);
int  hbc_graph_end(void *stream, void **exec);
int  hbc_graph_launch(void *exec, void *stream);
int  hbc_graph_destroy(void *exec);
const char *hbc_error_string(int code);

/* ---- frame maintenance */
int hbk_pad_frame(const hbd_frame *f, void *stream);                                   /* replicate borders of all planes */
int hbk_ingest_frame(const hbd_frame *f, const uint8_t *stage, int border, void *stream);          /* dense planes -> padded planes (+ borders) */
int hbk_narrow_plane(const int16_t *src, int src_stride, hbd_plane dst, uint32_t *range_flag, void *stream);

/* ---- per-call kernels: operands live in mapped pinned staging, strides in samples */
int hbk_pc_sad(const int16_t *a, int as, const int16_t *b, int bs, int n, int squared, uint32_t *out, void *stream);
int hbk_pc_predict(const int16_t *orig, int os, const int16_t *pred, int ps, int16_t *res, int rs, int n, void *stream);
int hbk_pc_wavg(const int16_t *s0, const int16_t *s1, int16_t *dst, int w, int h, void *stream);   /* tight w-sample rows */
int hbk_pc_reconst(const int16_t *pred, int ps, const int16_t *res, int rs, int16_t *dec, int ds, int n, void *stream);
/* src points at the first needed sample (margins included), org_off = offset of sample (0,0) */
int hbk_pc_interp(const int16_t *src, int ss, int org_off, int16_t *dst, int ds, int chroma, int fraction, int w, int h,
                  int vertical, int first, int last, void *stream);
int hbk_pc_transform(const int16_t *block, int bs, int16_t *coeff, int n, int dst4, void *stream);
int hbk_pc_itransform(int16_t *block, int bs, const int16_t *coeff, int n, int dst4, void *stream);
int hbk_pc_quant(const int16_t *src, int16_t *dst, int16_t *delta_u, int32_t *sum, int n, const int32_t *qtab,
                 const uint16_t *scan, int qbits, int add, int sign_hiding, void *stream);
int hbk_pc_inv_quant(const int16_t *src, int16_t *dst, int n, const int32_t *dqtab, int per, void *stream);

/* ---- batched jobs on resident frames */
/* per-frame scalars that change between replays of a captured launch sequence; kernels read them from device
 * memory when the pointer is non-NULL, otherwise the per-job / per-launch values are used */
typedef struct hbd_dyn_params {
    double corr;                  /* qp * clip(avg_dist/2000, .15, 1.4) */
    double thr_k;                 /* clip(avg_dist/2.5-5, 1, 20000) */
} hbd_dyn_params;
typedef struct hbd_me_job {       /* one PU, device layout */
    int32_t x, y;
    int32_t n_amvp; int32_t amvp[4];
    int32_t n_start; int32_t start[6];
    int32_t parent;               /* index into parent results or -1 */
    int32_t out;                  /* index into results */
    int32_t pad_;
    double  corr;                 /* qp * clip(avg_dist/2000, .15, 1.4), hmr_common.h:53 */
} hbd_me_job;
/* the fifteen quarter-pel planes of a reference picture's luma (hb_kernels_subpel.cu): plane q = fy * 4 + fx (1..15) starts at
 * base + (q - 1) * plane_bytes; the sample at picture position (X + fx/4, Y + fy/4) sits at row Y + HB_SUBPEL_OFF, column X + HB_SUBPEL_OFF;
 * w, h = picture size + 2 * HB_SUBPEL_OFF */
#define HB_SUBPEL_OFF 4
typedef struct hbd_subpel { uint8_t *base; int32_t pitch; int32_t w, h; int32_t pad_; uint64_t plane_bytes; } hbd_subpel;
int hbk_subpel_planes(const hbd_frame *ref, const hbd_subpel *sp, int y_lo, int y_hi /* picture rows wanted */, void *stream);
int hbk_subpel_uses_tma(void);
int hbc_clear_error(void);
int hbc_host_span_one_copy(const void *p, size_t bytes);   /* 1: [p, p + bytes) of host memory may be copied in one piece */     /* cudaGetLastError: returns and resets the last (non-sticky) error */
/* CUDA IPC (peer pictures of the CTU-row bands) */
int hbc_ipc_get_mem(void *dev, unsigned char out[64]);
int hbc_ipc_open_mem(const unsigned char in[64], void **dev);
int hbc_ipc_close_mem(void *dev);
int hbc_ipc_event_create(void **ev, unsigned char out[64]);
int hbc_ipc_event_open(const unsigned char in[64], void **ev);
typedef struct hbd_pull_span { const uint8_t *src; uint8_t *dst; int32_t src_pitch, dst_pitch, width, rows; } hbd_pull_span;
int hbk_pull_rows(const hbd_pull_span *spans, int n_spans, void *stream);       /* n_spans <= 12, widths multiples of 4 */
int hbk_me_configure(void);   /* once per device, before the first search launch */
int hbk_me_search(const hbd_frame *cur, const hbd_frame *ref, int size, const hbd_me_job *jobs, int n_jobs,
                  const hb_me_result *parent, hb_me_result *out, int action, const hbd_dyn_params *dyn,
                  const hbd_frame *pred_out /* NULL, or receives the luma prediction of every winner (needs HB_ME_HALF) */,
                  const hbd_subpel *sp /* NULL: sub-pel planes are built per PU in shared memory; else the picture's planes are read */,
                  int window /* != 0 (needs sp): `jobs` is strip ordered -- every hbk_me_strip_pus(size) consecutive entries are neighbours of one PU
                                row, the first one real, x < 0 marks padding -- and each CTA stages its strip's search window in shared memory */,
                  void *stream);
int hbk_me_strip_pus(int size);
/* the whole search of a picture (or a band of CTU rows) in one launch: a CTA per CTU, PU sizes 64 -> 8 in turn, zero predictors, the parent's
 * vector as extra start point; out[d] / pred_out[d] = result table (PU raster of the picture) and prediction picture of depth d */
int hbk_me_search_ctus(const hbd_frame *cur, const hbd_frame *ref, const hbd_subpel *sp, hb_me_result *const out[4], const hbd_frame *const pred_out[4],
                       int action, const hbd_dyn_params *dyn, int ctu_cols, int ctu_row0, int ctu_rows, const int grid_w[4], void *stream);

typedef struct hbd_mc_pu { int32_t x, y; int32_t mv_idx; } hbd_mc_pu;   /* luma position; mv = mvsrc[mv_idx].mv */
int hbk_mc_predict(const hbd_frame *ref, const hbd_frame *pred, int size, const hbd_mc_pu *pus, int n_pus,
                   const hb_me_result *mvsrc, int planes /* bit 0 luma, bit 1 chroma */, void *stream);
/* SSD(cur, pred) of square blocks of `size` (chroma size/2) at pus[i], all three planes: out[i * 3 + plane] */
int hbk_block_ssd(const hbd_frame *cur, const hbd_frame *pred, const hbd_mc_pu *pus, int n_pus, int size, uint32_t *out, void *stream);
/* bi-prediction: the average of the two lists' 14-bit predictions (one luma + one chroma kernel) */
int hbk_mc_predict_bi(const hbd_frame *ref0, const hbd_frame *ref1, const hbd_frame *pred, int size, const hbd_mc_pu *pus, int n_pus,
                      const hb_me_result *mvsrc0, const hb_me_result *mvsrc1, void *stream);

typedef struct hbd_tq_args {
    hbd_plane cur, pred, rec;     /* planes of the component being coded */
    const int32_t *jobs_xy;       /* n_jobs pairs (x,y) in samples of that plane */
    int32_t n_jobs;
    int32_t n;                    /* TU size 4..32 */
    const int32_t *qtab, *dqtab;  /* N*N tables of (list, qp%6) */
    const uint16_t *scan;         /* diagonal scan, N*N */
    int32_t qbits, add, per;
    int32_t sign_hiding;
    int32_t is_luma;
    int32_t intra;                /* 1: the intra chain (DST for 4x4 luma, no zero-out test, ssd against the reconstruction) */
    double  thr_k;                /* clip(avg_dist/2.5-5, 1, 20000) */
    double  weight;               /* chroma SSD weight, 1 for luma */
    const hbd_dyn_params *dyn;    /* overrides thr_k when non-NULL */
    int16_t *coeff_out;           /* n_jobs * N*N */
    hb_tu_result *res_out;        /* n_jobs */
} hbd_tq_args;
int hbk_tq_encode(const hbd_tq_args *a, void *stream);

/* ---- a whole intra picture in one persistent launch (k_intra_wave, hb_kernels_tq.cu): units sorted by (dependency level, plane, size, qp, scan) */
typedef struct hbd_wave_unit { int32_t x, y, mode; int32_t flags, lbs, trs; } hbd_wave_unit;          /* flags: bit 0 left, 1 top, 2 left-bottom, 3 top-right */
typedef struct hbd_wave_group { int32_t comp, qbits, add, per; const int32_t *qtab, *dqtab; const uint16_t *scan; double weight; } hbd_wave_group;
typedef struct hbd_wave_task { int32_t first_unit, n_units, group, size; int32_t units_before /* units of all earlier levels */, pad_; int64_t coeff_off; } hbd_wave_task;
typedef struct hbd_wave_args {
    hbd_frame cur, pred, rec;
    const hbd_wave_unit *units; const int32_t *xy; const hbd_wave_task *tasks; const hbd_wave_group *groups;
    int32_t n_tasks, sign_hiding;
    int16_t *coeff; hb_tu_result *res;
    unsigned *counters;           /* [0] next task, [1] finished units; zeroed before the launch */
} hbd_wave_args;
int hbk_intra_wave(const hbd_wave_args *w, int ctas, void *stream);

/* ---- block read-back (hb_kernels_enc.cu): size x size samples of plane `comp` at (x, y), x a multiple of 4, widened to int16 at
 * out + off (off in int16 units, a multiple of 4) */
typedef struct hbd_fetch_job { int32_t comp, x, y, size, off, pad_; } hbd_fetch_job;
int hbk_fetch_blocks(const hbd_frame *f, const hbd_fetch_job *jobs, int n_jobs, int16_t *out, void *stream);

/* ---- intra prediction (hb_kernels_intra.cu) */
typedef struct hbd_intra_job {
    int32_t comp, x, y, size;     /* plane, position and size (4..32) in samples of that plane */
    int32_t mode;                 /* 0 planar, 1 DC, 2..34 angular: write the prediction; < 0: SADs of all 35 luma modes */
    int32_t filtered;             /* luma only: 1 / 0 force the smoothed / raw reference samples, < 0 apply the search rule */
    int32_t adi_off;              /* first of the 4*size+1 reference samples of this job inside `adi` */
    int32_t pad_;
} hbd_intra_job;
typedef struct hbd_intra_args {
    hbd_plane cur;                /* luma plane of the current frame (SAD form) */
    hbd_frame pred;               /* receives predictions (prediction form) */
    const hbd_intra_job *jobs;
    int32_t n_jobs;
    const int16_t *adi;
    uint32_t *sads;               /* n_jobs x 35 */
} hbd_intra_args;
int hbk_intra(const hbd_intra_args *a, void *stream);                                  /* prediction-form jobs (mode >= 0) */
int hbk_intra_sads(const hbd_intra_args *a, int size, const int32_t *idx, int n_idx, void *stream);   /* SAD-form jobs of one size */
/* reference samples of intra units from the reconstructed picture (fill_reference_samples hmr_motion_intra.c:246): 4n+1 int16 per job at adi + adi_off */
typedef struct hbd_adi_job { int32_t comp, x, y, n; int32_t flags /* bit 0 left, 1 top, 2 left-bottom, 3 top-right */, lbs, trs, adi_off; } hbd_adi_job;
int hbk_intra_adi(const hbd_frame *rec, const hbd_adi_job *jobs, int n_jobs, int16_t *adi, void *stream);
int hbk_pc_intra(const int16_t *adi, int n, int mode, int is_luma, int16_t *pred, int stride, void *stream);

/* ---- gather of the host's selection (hb_kernels_gather.cu) */
typedef struct hbd_gather_pc {
    const int32_t *tu_index;      /* TU raster position (frame grid) -> index among the coded TUs of this (pass, plane), or -1 */
    int32_t grid_w, grid_h;       /* TU grid of the frame */
    int32_t tu;                   /* TU side, 0 when the plane is not coded in this pass */
    const hb_tu_result *res;
    const int16_t *coeff;
} hbd_gather_pc;
typedef struct hbd_gather_args {
    const uint8_t *sel;           /* chosen pass (0..4) per CTU */
    const int32_t *ctu_off;       /* start of each CTU's level stream, int16 units */
    int32_t ctu_cols;
    hbd_frame recon[5];
    hbd_gather_pc pc[5][3];
    uint8_t *out_recon[3];        /* destination planes: the tight fetch buffer (pitch = plane width) or a frame's planes */
    int32_t out_pitch[3];
    int16_t *out_levels;
} hbd_gather_args;
int hbk_gather(const hbd_gather_args *a, int n_ctus, void *stream);
int hbk_coeff_wnd(const hbd_gather_args *a, int n_ctus, int16_t *out /* n_ctus x 6144 */, void *stream);
/* deblocking unit data of the chosen passes, from what the pre-pass left on the device: per 4x4 unit the CU / TU depth of the CTU's
 * pass, the vector of the PU and the coded flag of the luma TU that cover it */
typedef struct hbd_units_args {
    const uint8_t *sel;
    int32_t ctu_cols, uw, uh, units_w, qp;
    const hb_me_result *me[4];
    int32_t me_grid_w[4];
    hbd_gather_pc luma[5];
    hb_unit_info *units;
} hbd_units_args;
int hbk_units_from_selection(const hbd_units_args *a, void *stream);
/* per-CU cost records of every pass (hb_cu_cost, compact_tables = 2) */
typedef struct hbd_cu_pack_args {
    hbd_gather_pc pc[5][3];
    int32_t grid_w[4];
    int32_t first[6];             /* first record of pass p; first[5] = total */
    hb_cu_cost *out;
} hbd_cu_pack_args;
int hbk_pack_cu_costs(const hbd_cu_pack_args *a, void *stream);
/* SAO statistics of every CTU and component of a frame: out[ctu * 3 + comp] */
int hbk_sao_stats(const hbd_frame *org, const hbd_frame *rec, int ctu_cols, int n_ctus, hb_sao_stats *out, void *stream);
/* offsets, band position and distortion estimate of all five types per (CTU, component): out[unit * 5 + type], device memory */
int hbk_sao_derive(const hb_sao_stats *stats, int n_units, const double lambda[3], hb_sao_candidate *out, void *stream);
/* boundary strengths + QP map of a P picture from per-unit mode data (device pointers) */
int hbk_amvp(const hb_unit_info *units, int units_w, int w, int h, const hb_amvp_job *jobs, int n_jobs, hb_amvp_list *out, void *stream);
int hbk_amvp_fill(const hb_unit_info *units, int units_w, int w, int h, hbd_me_job *jobs, int n_jobs, int size, void *stream);
int hbk_merge_cands(const hb_unit_info *units, int units_w, int w, int h, const hb_amvp_job *jobs, int n_jobs, int max_cands, hb_mv *out, void *stream);
int hbk_deblock_strengths_b(const hb_unit_info *units, const hb_unit_l1 *units1, const int32_t *pic_l0, int n_l0, const int32_t *pic_l1, int n_l1,
                            int units_w, int w, int h, uint8_t *bs_ver, uint8_t *bs_hor, uint8_t *qp, void *stream);
int hbk_deblock_strengths(const hb_unit_info *units, int units_w, int w, int h, uint8_t *bs_ver, uint8_t *bs_hor, uint8_t *qp, void *stream);
/* deblocking, pixel stage, in place: all vertical edges, then all horizontal ones (two launches); maps in device memory */
int hbk_deblock(const hbd_frame *f, const uint8_t *bs_ver, const uint8_t *bs_hor, const uint8_t *qp, int units_w,
                int cb_off, int cr_off, int beta_off2, int tc_off2, void *stream);
/* SAO offset pass: dst = src with the per-CTU offsets applied (prm[ctu], device memory) */
int hbk_sao_apply(const hbd_frame *src, const hbd_frame *dst, int ctu_cols, int n_ctus, const hb_sao_param *prm, void *stream);
/* full result tables (n_me hb_me_result, then n_tu hb_tu_result) -> compact records (hb_me_result_c, hb_tu_result_c) */
int hbk_pack_tables(const void *full, void *compact, int n_me, int n_tu, void *stream);

#ifdef __cplusplus
}
#endif
#endif
