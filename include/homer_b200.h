/*
 * homer_b200.h -- C ABI of the B200 (sm_100a) implementation of HomerHEVC's data-parallel hot path:
 * SAD/SSD motion search, luma/chroma interpolation with sub-pel refinement, motion compensation,
 * forward/inverse DCT/DST with quantisation (incl. sign-data hiding), dequantisation and reconstruction.
 *
 * Plain C: pointers, sizes and PODs only.  The library is libhomer_b200.so (homerhevc_b200/csrc).
 * There is NO CPU fallback: every entry point computes on the GPU and fails loudly when CUDA is unavailable
 * (status-returning calls give HB_ERR_CUDA; the void/value-returning drop-ins of section A print and abort()).
 *
 * Reference interfaces replaced (paths under /root/reference/src/homer_lib/):
 *   section A -- the members of struct low_level_funcs_t, hmr_private.h:1063-1092, same prototypes;
 *   section B -- frame residency: the int16 wnd_t frames of hmr_private.h:660 / hmr_encoder_lib.c:1512-1516
 *                and reference_picture_border_padding_ctu hmr_encoder_lib.c:1723;
 *   section C -- batched forms of hmr_motion_estimation hmr_motion_inter.c:1404,
 *                hmr_motion_compensation_luma/_chroma :1779/:1860, encode_inter_cu/_chroma :40/:133;
 *   section D -- the frame-level pre-pass the north star asks for (no counterpart in the reference; it is the
 *                batched composition of section C over every CTU and every partition size).
 */
#ifndef HOMER_B200_H
#define HOMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_OK         0
#define HB_ERR_CUDA  -1     /* no device / CUDA runtime error (message in hb_last_error) */
#define HB_ERR_ARG   -2
#define HB_ERR_NOMEM -3

typedef struct hb_ctx hb_ctx;         /* one per GPU: stream, scratch, tables */
typedef struct hb_frame hb_frame;     /* one YUV 4:2:0 8-bit picture resident in HBM, border-padded */
typedef struct hb_prepass hb_prepass; /* plan + resident outputs of the frame-level pre-pass */

/* ------------------------------------------------------------------ lifecycle ------------------------------ */
int         hb_device_count(void);
int         hb_ctx_create(hb_ctx **out, int device);
void        hb_ctx_destroy(hb_ctx *ctx);
const char *hb_last_error(void);                       /* thread-local text of the last failure */
int         hb_ctx_sync(hb_ctx *ctx);                  /* wait for everything queued on the context's stream */
void       *hb_ctx_stream(hb_ctx *ctx);                /* the cudaStream_t work is queued on (for interop/timing) */
int         hb_ctx_wait(hb_ctx *waiter, hb_ctx *signaler); /* order two contexts' streams on the device (no host wait) */
uint64_t    hb_ctx_launch_count(hb_ctx *ctx);          /* kernels launched so far through this context */
/* device-side stopwatch on the context's stream (CUDA events): begin, ..., end -> milliseconds */
int         hb_timer_begin(hb_ctx *ctx);
int         hb_timer_end(hb_ctx *ctx, float *ms_out);  /* synchronises */
/* pinned host memory for async uploads/downloads */
void       *hb_pinned_alloc(size_t bytes);
void        hb_pinned_free(void *p);

/* ------------------------------------------------------------------ A. per-call drop-ins -------------------
 * Exactly the prototypes of low_level_funcs_t (hmr_private.h:1069-1088).  Buffers are HOST memory owned by the
 * caller, as in the reference; each call stages them through pinned memory and runs one kernel on the calling
 * thread's stream of the default context (device $HB_DEVICE, default 0).  Thread-safe and re-entrant.
 * They exist so the library is a literal drop-in and so per-call parity can be tested; throughput comes from
 * sections C and D.  Semantics: the reference's plain-C definitions (== its SSE4.2 functions on 8-bit video),
 * quant with the SSE4.2 rounding rule (hmr_sse42_functions_quant.c:47). */
hb_ctx  *hb_default_ctx(void);          /* the lazily created context (device $HB_DEVICE) the drop-ins run on */
uint32_t hb_sad(int16_t *src, uint32_t src_stride, int16_t *pred, uint32_t pred_stride, int size);
uint32_t hb_ssd16b(int16_t *src, uint32_t src_stride, int16_t *pred, uint32_t pred_stride, int size);
void hb_predict(int16_t *orig, int orig_stride, int16_t *pred, int pred_stride, int16_t *residual, int residual_stride, int size);
void hb_reconst(int16_t *pred, int pred_stride, int16_t *residual, int residual_stride, int16_t *decoded, int decoded_stride, int size);
/* weighted_average_motion, hmr_private.h:1084 / hmr_motion_inter.c:2903: average of the two 14-bit predictions of a bi-predicted block */
void hb_weighted_average_motion(int16_t *src0, int src0_stride, int16_t *src1, int src1_stride, int16_t *dst, int dst_stride, int height, int width, int bit_depth);
void hb_interpolate_luma(int16_t *reference_buff, int reference_buff_stride, int16_t *pred_buff, int pred_buff_stride,
                         int fraction, int width, int height, int is_vertical, int is_first, int is_last);
void hb_interpolate_chroma(int16_t *reference_buff, int reference_buff_stride, int16_t *pred_buff, int pred_buff_stride,
                           int fraction, int width, int height, int is_vertical, int is_first, int is_last);
void hb_transform(int bit_depth, int16_t *block, int16_t *coeff, int block_size, int iWidth, int iHeight,
                  int width_shift, int height_shift, uint16_t uiMode, int16_t *aux);
void hb_itransform(int bit_depth, int16_t *block, int16_t *coeff, int block_size, int iWidth, int iHeight,
                   unsigned int uiMode, int16_t *aux);
#define HB_REG_DCT 65535              /* hmr_private.h:217; any other uiMode selects the 4x4 DST */

/* quant / inv_quant take a henc_thread_t* in the reference only to reach these five values
 * (hmr_sse42_functions_quant.c:36-48, :137-142); the drop-in takes them directly.  See INTEGRATION.md for the
 * two-line adapter on the reference side. */
typedef struct hb_quant_env {
    int32_t is_islice;            /* currslice->slice_type == I_SLICE */
    int32_t sign_hiding;          /* et->pps->sign_data_hiding_flag */
    int32_t max_cu_size_shift;    /* et->max_cu_size_shift (6) */
    int32_t bit_depth;            /* et->bit_depth (8) */
    int16_t *delta_u;             /* et->aux_buff, receives deltaU (may be NULL) */
} hb_quant_env;
void hb_quant(const hb_quant_env *env, int16_t *src, int16_t *dst, int scan_mode, int depth, int comp, int cu_mode,
              int is_intra, int *ac_sum, int cu_size, int per, int rem);
void hb_inv_quant(const hb_quant_env *env, int16_t *src, int16_t *dst, int depth, int comp, int is_intra,
                  int cu_size, int per, int rem);

/* The table itself, laid out member for member like low_level_funcs_t (19 pointers, hmr_private.h:1063-1092).
 * hb_fill_low_level_funcs overwrites the 10 members with table-compatible prototypes that this library implements and leaves the
 * rest untouched: copies, variance, SAO stats (out of scope), quant / inv_quant and the two intra predictors
 * (implemented, but their table prototypes take henc_thread_t* / ctu_info_t*: INTEGRATION.md shows the adapters). */
typedef struct hb_low_level_funcs {
    void *sse_copy_16_16, *sse_copy_16_8, *sse_copy_8_16;
    uint32_t (*sad)(int16_t *, uint32_t, int16_t *, uint32_t, int);
    uint32_t (*ssd16b)(int16_t *, uint32_t, int16_t *, uint32_t, int);
    void (*predict)(int16_t *, int, int16_t *, int, int16_t *, int, int);
    void (*reconst)(int16_t *, int, int16_t *, int, int16_t *, int, int);
    void *modified_variance, *create_intra_planar_prediction, *create_intra_angular_prediction;
    void (*interpolate_luma_m_compensation)(int16_t *, int, int16_t *, int, int, int, int, int, int, int);
    void (*interpolate_chroma_m_compensation)(int16_t *, int, int16_t *, int, int, int, int, int, int, int);
    void (*interpolate_luma_m_estimation)(int16_t *, int, int16_t *, int, int, int, int, int, int, int);
    void (*weighted_average_motion)(int16_t *, int, int16_t *, int, int16_t *, int, int, int, int);
    void *quant, *inv_quant;      /* need the henc_thread_t adapter, see INTEGRATION.md */
    void (*transform)(int, int16_t *, int16_t *, int, int, int, int, int, uint16_t, int16_t *);
    void (*itransform)(int, int16_t *, int16_t *, int, int, int, unsigned int, int16_t *);
    void *get_sao_stats;
} hb_low_level_funcs;
void hb_fill_low_level_funcs(hb_low_level_funcs *table);

/* ------------------------------------------------------------------ B. resident frames ---------------------
 * 8-bit planes with a replicated border of HB_PAD_LUMA (luma) / HB_PAD_LUMA/2 (chroma) samples on every side.
 * The reference keeps int16 frames with an 80-sample border; all its samples are 0..255, so u8 is lossless and
 * halves HBM/PCIe traffic.  hb_frame_upload_i16 accepts the reference's own wnd_t planes (int16) and narrows on
 * the device; it reports HB_ERR_ARG (after the copy) if a sample was outside 0..255. */
#define HB_PAD_LUMA 96
int  hb_frame_create(hb_ctx *ctx, int width, int height, hb_frame **out);
void hb_frame_destroy(hb_frame *f);
int  hb_frame_width(const hb_frame *f);
int  hb_frame_height(const hb_frame *f);
/* async on the context's stream; host planes should be pinned (hb_pinned_alloc) for true async copies.
 * Uploads finish with the border replication (hmr_encoder_lib.c:1723). */
int  hb_frame_upload_u8(hb_ctx *ctx, hb_frame *f, const uint8_t *y, int y_stride, const uint8_t *u, int u_stride,
                        const uint8_t *v, int v_stride);
#define HB_UPLOAD_NO_BORDER 1   /* a current frame is only read inside the picture: skip the border replication */
int  hb_frame_upload_u8_ex(hb_ctx *ctx, hb_frame *f, const uint8_t *y, int y_stride, const uint8_t *u, int u_stride,
                           const uint8_t *v, int v_stride, int flags);
int  hb_frame_upload_i16(hb_ctx *ctx, hb_frame *f, const int16_t *y, int y_stride, const int16_t *u, int u_stride,
                         const int16_t *v, int v_stride);
int  hb_frame_download_u8(hb_ctx *ctx, const hb_frame *f, uint8_t *y, int y_stride, uint8_t *u, int u_stride,
                          uint8_t *v, int v_stride);

/* CTU-row bands on several GPUs: rows [row0, row0+n_rows) of a plane <-> a tight device buffer (plane width bytes per row).
 * The caller moves that buffer to the neighbour GPU (NCCL send/recv, peer copy) and imports it there; hb_frame_pad
 * refreshes the replicated border afterwards.  A band needs 64 (search) + 4 (8-tap) luma rows and 36 chroma rows per side. */
#define HB_HALO_LUMA 68
#define HB_HALO_CHROMA 36
int  hb_frame_export_rows(hb_ctx *ctx, const hb_frame *f, int plane, int row0, int n_rows, void *dev_dst);
int  hb_frame_import_rows(hb_ctx *ctx, hb_frame *f, int plane, int row0, int n_rows, const void *dev_src);
int  hb_frame_pad(hb_ctx *ctx, hb_frame *f);

/* The same exchange without staging buffers or a collective (one process per GPU, NVLink / NVSwitch peer memory): the owner of a
 * picture exports its three device allocations once (CUDA IPC), a neighbour opens them as a read-only VIEW and pulls the halo rows
 * it needs straight out of the owner's HBM with ONE copy kernel per picture on its own stream (+ the border refresh); nothing for the
 * host to wait on.  Ordering across the two processes is an inter-process event: the owner records it on its stream when the picture
 * is finished, the neighbour makes its stream wait for it before it pulls. */
#define HB_IPC_HANDLE_BYTES 64
typedef struct hb_frame_ipc { uint8_t mem[3][HB_IPC_HANDLE_BYTES]; int32_t width, height; } hb_frame_ipc;
int  hb_frame_ipc_export(const hb_frame *f, hb_frame_ipc *out);
int  hb_frame_ipc_open(hb_ctx *ctx, const hb_frame_ipc *in, hb_frame **view);      /* another process's picture; hb_frame_destroy closes the view */
typedef struct hb_row_span { int32_t src, plane, row0, n_rows; } hb_row_span;      /* rows [row0, row0 + n_rows) of `plane` from srcs[src] */
#define HB_MAX_ROW_SPANS 12
int  hb_frame_pull_rows(hb_ctx *ctx, hb_frame *dst, const hb_frame *const *srcs, int n_srcs, const hb_row_span *spans, int n_spans, int refresh_border);
typedef struct hb_ipc_event hb_ipc_event;
int  hb_ipc_event_create(hb_ctx *ctx, hb_ipc_event **ev, uint8_t handle[HB_IPC_HANDLE_BYTES]);
int  hb_ipc_event_open(hb_ctx *ctx, const uint8_t handle[HB_IPC_HANDLE_BYTES], hb_ipc_event **ev);
int  hb_ipc_event_record(hb_ctx *ctx, hb_ipc_event *ev);     /* on the context's stream */
int  hb_ipc_event_wait(hb_ctx *ctx, hb_ipc_event *ev);       /* the context's stream waits; the host does not */
void hb_ipc_event_destroy(hb_ipc_event *ev);

/* ------------------------------------------------------------------ C. batched jobs on resident frames ----- */
typedef struct hb_mv { int32_t x, y; } hb_mv;                 /* quarter-pel units (motion_vector_t, hmr_private.h:712) */

/* one hmr_motion_estimation call (hmr_motion_inter.c:1404) */
typedef struct hb_me_job {
    int32_t x, y;               /* luma position of the PU in the frame (curr_part_global_x/y) */
    int32_t size;               /* 8, 16, 32 or 64; the PU must lie inside the frame */
    int32_t qp;                 /* curr_cu_info->qp, weights the MV cost (hmr_common.h:53) */
    int32_t n_amvp;             /* 0..2 AMVP candidates the MV cost is measured against */
    hb_mv   amvp[2];
    int32_t n_start;            /* 0..3 extra start points (et->mv_search_candidates) */
    hb_mv   start[3];
    int32_t parent;             /* >= 0: index into `parent_results`; that MV is appended to the start points when
                                   both its components are non-zero (hmr_motion_inter.c:2613).  -1: none */
} hb_me_job;
typedef struct hb_me_result {
    hb_mv    mv;                /* best MV, quarter-pel */
    hb_mv    subpix;            /* its sub-pel part as the reference reports it */
    uint32_t sad;               /* SAD at mv (return value of hmr_motion_estimation) */
    uint32_t n_probes;          /* integer positions evaluated (diagnostic) */
} hb_me_result;
#define HB_ME_PEL 1
#define HB_ME_HALF 2
#define HB_ME_QUARTER 4
/* jobs/results/parent_results are HOST arrays; search range is the reference's +-128 x +-64 (hmr_private.h:76). */
int hb_me_search(hb_ctx *ctx, const hb_frame *cur, const hb_frame *ref, const hb_me_job *jobs, int n_jobs,
                 const hb_me_result *parent_results, int n_parent, double avg_dist, int action, hb_me_result *results);

/* one hmr_motion_compensation_luma + two _chroma calls (uni-prediction) */
typedef struct hb_mc_job { int32_t x, y, size; hb_mv mv; } hb_mc_job;     /* size: luma 8..64 (chroma size/2) */
int hb_mc_predict(hb_ctx *ctx, const hb_frame *ref, hb_frame *pred, const hb_mc_job *jobs, int n_jobs);
/* bi-prediction (B slices): both lists with is_bi_predict = 1 (14-bit predictions) + weighted_average_motion, hmr_motion_inter.c:3047-3056 */
typedef struct hb_mc_bi_job { int32_t x, y, size; hb_mv mv0, mv1; } hb_mc_bi_job;     /* mv0 into ref0 (list 0), mv1 into ref1 (list 1) */
int hb_mc_predict_bi(hb_ctx *ctx, const hb_frame *ref0, const hb_frame *ref1, hb_frame *pred, const hb_mc_bi_job *jobs, int n_jobs);

/* one encode_inter_cu (comp 0) or encode_inter_cu_chroma (comp 1,2) call: T -> Q -> [IQ -> IT -> SSD -> zero-out] -> recon */
typedef struct hb_tu_job { int32_t comp; int32_t x, y; int32_t size; int32_t qp; } hb_tu_job; /* x,y,size in samples of that plane; qp already chroma-mapped */
typedef struct hb_tu_result { int32_t sum; uint32_t ssd; uint32_t ssd_zero; int32_t zeroed; } hb_tu_result;
typedef struct hb_tq_params {
    int32_t is_islice;          /* rounding 171 vs 85 (hmr_sse42_functions_quant.c:47) */
    int32_t sign_hiding;
    double  avg_dist;           /* zero-out threshold input (hmr_motion_inter.c:106) */
    double  chroma_weight;      /* pow(2,(qp-qp_c)/3) of hmr_motion_inter.c:155; luma uses 1 */
} hb_tq_params;
/* coeffs: host array receiving size*size levels per job, jobs back to back in job order. recon gets the decoded samples. */
int hb_tq_encode(hb_ctx *ctx, const hb_frame *cur, const hb_frame *pred, hb_frame *recon, const hb_tu_job *jobs, int n_jobs,
                 const hb_tq_params *params, int16_t *coeffs, hb_tu_result *results);

/* the intra chain once the prediction exists: encode_intra_cu after its prediction step (hmr_motion_intra.c:1023-1069) and the
 * chroma loop of hmr_motion_intra_chroma.c:340-365.  `pred` holds the intra prediction (built by the host's predictors, which
 * walk reconstructed neighbours).  4x4 luma uses the DST; scan_mode is find_scan_mode()'s value for the intra mode
 * (1 horizontal, 2 vertical, 3 diagonal; sizes above 8 are always diagonal); result.ssd = ssd16b(original, reconstruction),
 * chroma weighted by chroma_weight and truncated like the reference's (int) cast. */
typedef struct hb_intra_tu_job { int32_t comp; int32_t x, y; int32_t size; int32_t qp; int32_t scan_mode; } hb_intra_tu_job;
int hb_tq_encode_intra(hb_ctx *ctx, const hb_frame *cur, const hb_frame *pred, hb_frame *recon, const hb_intra_tu_job *jobs, int n_jobs,
                       int is_islice, int sign_hiding, double chroma_weight, int16_t *coeffs, hb_tu_result *results);

/* intra prediction: planar hmr_motion_intra.c:408, DC/angular + edge filters :482, reference-sample smoothing
 * adi_filter :189, filtered-or-not rule of the mode search :1122.  Every job brings its 4*size+1 reference samples (index 2*size =
 * top-left corner, +i the row above, -i the left column), which the host takes from reconstructed neighbours.
 * mode >= 0: the prediction is written into `pred`; mode < 0 (luma): sads[35*i + m] = SAD(current block, mode m) for all modes. */
typedef struct hb_intra_job {
    int32_t comp, x, y, size;     /* size 4..32 */
    int32_t mode;                 /* 0 planar, 1 DC, 2..34 angular, < 0 = all-mode SADs */
    int32_t filtered;             /* luma: 1 / 0 force smoothed / raw samples, < 0 = the reference's rule (intra_filter[] :148) */
} hb_intra_job;
int hb_intra_run(hb_ctx *ctx, const hb_frame *cur, hb_frame *pred, const hb_intra_job *jobs, int n_jobs, const int16_t *adi, uint32_t *sads);
/* per-call forms of the two table members (adapter drops et / ctu) */
void hb_create_intra_planar_prediction(int16_t *prediction, int pred_stride, int16_t *adi_pred_buff, int adi_size, int cu_size, int cu_size_shift);
void hb_create_intra_angular_prediction(int16_t *prediction, int pred_stride, int16_t *adi_pred_buff, int adi_size, int cu_size, int cu_mode, int is_luma);

/* Reconstruction of the intra transform units of a picture (BASELINE.json configs[0] / [4]; encode_intra_cu hmr_motion_intra.c:973-1069, chroma
 * hmr_motion_intra_chroma.c:293-365) from the host's decisions, WAVEFRONT-batched: a unit predicts from reconstructed samples of its
 * neighbours (fill_reference_samples :246 with the reference's own availability flags :625 / :676 and padding), so the units of a picture are
 * levelled by that dependency -- level k holds the units all of whose neighbours lie in levels < k -- and every level runs as a batch:
 * reference samples gathered from the reconstructed picture on the device, smoothing rule (:1011) and predictor of the unit's mode, residual,
 * DST / DCT, quantisation with sign hiding, dequantisation, inverse transform, reconstruction into `recon`, which the next level reads.
 * units: in coding order (any order in which a unit comes after the units it predicts from).  comp / x / y / size: plane and block in samples
 * of that plane; mode 0..34 (the derived chroma mode resolved by the host); qp of the component; scan_mode as find_scan_mode gives it
 * (hmr_tables.c:376); node_x / node_y / node_size: the quadtree node, in luma samples, whose neighbour flags apply (the unit itself; for the
 * 4x4 chroma unit of an 8x8 coding unit that node).  `pred` receives the predictions (scratch picture of the same size).
 * coeffs: size^2 levels per unit, back to back in unit order; results[i].sum / ssd as for hb_tq_encode_intra. */
typedef struct hb_intra_unit { int32_t comp, x, y, size, mode, qp, scan_mode, node_x, node_y, node_size; } hb_intra_unit;
int hb_intra_reconstruct(hb_ctx *ctx, const hb_frame *cur, hb_frame *pred, hb_frame *recon, const hb_intra_unit *units, int n_units,
                         int is_islice, int sign_hiding, double chroma_weight, int16_t *coeffs, hb_tu_result *results, int32_t *n_levels_out);
/* flags = 0: ONE persistent launch per picture -- warps draw tasks (32 / size units of one level and group) from a global counter and wait on
 * a global count of finished units for the earlier levels; HB_INTRA_PER_LEVEL_LAUNCHES: a batch of launches per level (gather, predict, one
 * T/Q launch per plane / size / QP / scan group), the first form, kept for comparison.  Same results. */
#define HB_INTRA_PER_LEVEL_LAUNCHES 1
int hb_intra_reconstruct_ex(hb_ctx *ctx, const hb_frame *cur, hb_frame *pred, hb_frame *recon, const hb_intra_unit *units, int n_units,
                            int is_islice, int sign_hiding, double chroma_weight, int flags, int16_t *coeffs, hb_tu_result *results, int32_t *n_levels_out);

/* Merge / skip candidate evaluation (the compute of check_rd_cost_merge_2nx2n, hmr_motion_inter.c:3493, with
 * one transform depth): for every candidate {CU, list-0 vector} motion compensation (luma + chroma) into `pred`, then the inter
 * T/Q chain of its transform units (luma: the block, a 64x64 one as four 32x32; chroma: half the luma unit) into `recon`.  Per candidate: dist_coded = sum of the
 * units' ssd (what encode_inter returns, :3071), sum = sum of their level sums, dist_skip = ssd16b(orig, pred) of the whole luma block plus
 * the truncated weighted ones of the two chroma blocks (:3686-3688, the "no residual" branch), cbf bit c set when component c keeps coefficients.  The blocks
 * of ONE call must not overlap (one call per candidate index of the merge list). */
typedef struct hb_merge_result { uint32_t dist_coded; int32_t sum; uint32_t dist_skip; int32_t cbf; } hb_merge_result;
int hb_merge_eval(hb_ctx *ctx, const hb_frame *cur, const hb_frame *ref, hb_frame *pred, hb_frame *recon, const hb_mc_job *cands, int n_cands,
                  int qp, int chroma_qp_offset, const hb_tq_params *params, hb_merge_result *out);

/* Deblocking of a whole picture, pixel stage, in place (deblock_filter_luma / _chroma hmr_deblocking_filter.c:351/:504 with
 * filter_luma :287, filter_chroma :478, use_strong_filter :275, in the picture order of hmr_deblock_filter :827: every vertical
 * edge, then every horizontal one).  The boundary strengths are INPUTS: bs_ver / bs_hor hold, per 4x4 luma unit in picture
 * raster (units_w entries per row, at least width/4), the strength 0..2 of the edge on the unit's left / top side -- the
 * values get_boundary_strength_single (:138) leaves in deblock_filter_strength_bs; entries off the 8x8 grid are ignored.
 * qp: the QP of the CU each unit belongs to.  The replicated border is refreshed afterwards. */
typedef struct hb_deblock_params { int32_t cb_qp_offset, cr_qp_offset, beta_offset_div2, tc_offset_div2; } hb_deblock_params;
int hb_deblock_frame(hb_ctx *ctx, hb_frame *frame, const uint8_t *bs_ver, const uint8_t *bs_hor, const uint8_t *qp, int units_w,
                     const hb_deblock_params *params);

/* The same with the strengths derived on the device from what the host's decisions left per 4x4 unit (P pictures, 2Nx2N units):
 * transform / coding unit edges as hmr_deblock_filter_cu :737 marks them (transform leaves of at least 8x8, picture border
 * excluded :692) and get_boundary_strength_single :138 -- 2 if either side is intra, else 1 if either side's transform unit has
 * luma coefficients (bit tu_depth of cbf_luma), else 1 if the list-0 references differ or a vector component differs by >= 4.
 * bs_ver_out / bs_hor_out (optional, units_w * height/4 bytes each) receive the strengths that were used. */
typedef struct hb_unit_info { uint8_t cu_depth, tu_depth, intra, cbf_luma; int8_t ref_idx; uint8_t qp; int16_t mvx, mvy; } hb_unit_info;   /* 10 bytes */
int hb_deblock_frame_units(hb_ctx *ctx, hb_frame *frame, const hb_unit_info *units, int units_w, const hb_deblock_params *params,
                           uint8_t *bs_ver_out, uint8_t *bs_hor_out);
/* B pictures (the two-list branch of get_boundary_strength_single, hmr_deblocking_filter.c:173-229): units carries list 0, units_l1 list 1
 * (ref_idx < 0 = list not used by that unit); pic_l0 / pic_l1 name the picture behind every reference index (any numbering; equal numbers
 * = the same picture -- the reference compares picture pointers).  Prediction units smaller than their CU (NxN) need nothing special: the
 * motion is per 4x4 unit and only the 8x8 edge grid is filtered. */
typedef struct hb_unit_l1 { int8_t ref_idx, reserved; int16_t mvx, mvy; } hb_unit_l1;   /* 6 bytes */
int hb_deblock_frame_units_b(hb_ctx *ctx, hb_frame *frame, const hb_unit_info *units, const hb_unit_l1 *units_l1, int units_w,
                             const int32_t *pic_l0, int n_l0, const int32_t *pic_l1, int n_l1, const hb_deblock_params *params,
                             uint8_t *bs_ver_out, uint8_t *bs_hor_out);

/* AMVP candidates (get_amvp_candidates, hmr_motion_inter.c:2342, P pictures with one reference picture) of a
 * batch of 2Nx2N PUs from the motion field the host's decisions left per 4x4 unit (the same hb_unit_info maps the deblocking reads;
 * on the device they can come straight from hb_prepass_finalise): left-bottom / left, top-right / top / top-left neighbours with the
 * reference's availability rules (z-order inside the CTU, neighbour CTUs, the quadtree's left_bottom / top_right flags :625), the
 * above group taken twice when there is no left candidate, duplicate removal, zero fill.  out[i] = the two predictors of jobs[i]
 * (what hb_me_job.amvp takes).  units must cover whole CTUs: units_w >= 16 * CTU columns, 16 * CTU rows rows. */
typedef struct hb_amvp_job { int32_t x, y, size; } hb_amvp_job;              /* size 64 / 32 / 16 / 8, inside the picture */
typedef struct hb_amvp_list { hb_mv mv[2]; } hb_amvp_list;
int hb_amvp_candidates(hb_ctx *ctx, const hb_unit_info *units, int units_w, int width, int height, const hb_amvp_job *jobs, int n_jobs, hb_amvp_list *out);
/* hb_me_search with the predictors of every job derived on the device from the unit field right before its search (k_amvp_fill -> k_me, no
 * host round trip in between): jobs[i].amvp / n_amvp are ignored.  == hb_amvp_candidates followed by hb_me_search. */
int hb_me_search_field(hb_ctx *ctx, const hb_frame *cur, const hb_frame *ref, const hb_unit_info *units, int units_w, const hb_me_job *jobs, int n_jobs,
                       const hb_me_result *parent_results, int n_parent, double avg_dist, int action, hb_me_result *results);
/* Merge candidates of the same PUs (get_merge_mvp_candidates, hmr_motion_inter.c:1937; P pictures, one reference picture): the neighbours in
 * the order A1, B1, B0, A0, then B2 while fewer than four, pruned pairwise as equal_motion :1915 does, closed at max_cands (1..5, the
 * slice's max_num_merge_candidates) and filled with zero vectors.  out[i * max_cands + k] = candidate k of jobs[i]; what hb_merge_eval takes. */
int hb_merge_candidates(hb_ctx *ctx, const hb_unit_info *units, int units_w, int width, int height, const hb_amvp_job *jobs, int n_jobs, int max_cands, hb_mv *out);

/* SAO statistics (get_sao_stats of the function table, hmr_private.h:1091; sao_get_ctu_stats hmr_sao.c:75 /
 * sse_sao_get_ctu_stats hmr_sse42_sao.c:35, calculate_preblock_stats = 0) for every CTU and component of a picture in one
 * launch: `rec` is the deblocked reconstruction (before SAO), `orig` the source.  out[ctu * 3 + comp], CTUs in raster order.
 * Edge classes 0..4 stand for the reference's edge types -2..2; eo_*[k]: k = 0 EO_0 (horizontal), 1 EO_90, 2 EO_135, 3 EO_45. */
typedef struct hb_sao_stats { int32_t eo_diff[4][5], eo_count[4][5], bo_diff[32], bo_count[32]; } hb_sao_stats;   /* 416 bytes */
int hb_sao_stats_frame(hb_ctx *ctx, const hb_frame *orig, const hb_frame *rec, hb_sao_stats *out);
/* SAO offset pass of a whole picture (sao_offset_ctu hmr_sao.c:1210 / offset_block :960 for every CTU): dst = src (the deblocked
 * picture) with the CTU's offsets applied.  Per CTU (raster) and component: type -1 off, 0..3 EO_0/90/135/45 (offset[0..4] for edge
 * types -2..2), 4 band offset (offset[band], 32 entries) -- the values of sao_offset_t.offset.  dst must be another frame. */
typedef struct hb_sao_param { int8_t type[3]; int8_t reserved; int16_t offset[3][32]; } hb_sao_param;   /* 196 bytes */
int hb_sao_apply_frame(hb_ctx *ctx, const hb_frame *src, hb_frame *dst, const hb_sao_param *params);
/* The arithmetic half of the SAO decision (host code, no device work): sao_derive_offsets hmr_sao.c:480 (with est_iter_offset
 * :445), sao_invert_quant_offsets :592 and sao_get_distortion :620 for 8-bit video, on one CTU component's statistics.
 * type 0..3 EO_0/90/135/45, 4 band offset; lambda = enc_engine->sao_lambdas[component].  offset[]: the 32 entries of
 * sao_offset_t.offset (what hb_sao_param carries), *band: typeAuxInfo, *dist: the estimated distortion change. */
int hb_sao_derive_offsets(const hb_sao_stats *stats, int type, double lambda, int16_t offset[32], int32_t *band, int64_t *dist);
/* The same arithmetic on the device, for every CTU, component and type of a picture, right after the statistics (one more launch):
 * cand[(ctu * 3 + comp) * 5 + type].  offset[]: edge types -- classes 0, 1, 3, 4 (the plain class never has one); band type -- bands
 * band .. band + 3.  stats (optional) also receives the statistics.  Identical to hb_sao_derive_offsets on those statistics. */
typedef struct hb_sao_candidate { int64_t dist; int8_t offset[4]; int8_t band; int8_t reserved[3]; } hb_sao_candidate;   /* 16 bytes */
int hb_sao_candidates_frame(hb_ctx *ctx, const hb_frame *orig, const hb_frame *rec, const double lambda[3], hb_sao_candidate *cand, hb_sao_stats *stats);
/* A stand-in for sao_decide_blk_params (hmr_sao.c:1295), whose real form prices the syntax with the CABAC state and stays with
 * the encoder: per CTU, luma alone and both chroma planes jointly, the type minimising dist + lambda * bits against "off", with
 * the constant prices of the reference's COMPUTE_AS_HM branch (8 / 11 bits, off = 2.5 lambda) and no merge candidates. */
int hb_sao_decide_standin(const hb_sao_stats *stats, int n_ctus, const double lambda[3], hb_sao_param *params);
int hb_sao_decide_from_candidates(const hb_sao_candidate *cand, int n_ctus, const double lambda[3], hb_sao_param *params);   /* the same from hb_sao_candidates_frame's output */

/* ------------------------------------------------------------------ D. frame-level pre-pass ----------------
 * For every CTU and every inter PU size 64/32/16/8 at once: motion search chained parent -> child exactly as
 * hmr_cu_motion_estimation does (zero AMVP predictors, parent MV as extra start), motion compensation of
 * luma + chroma with the MVs found, and the inter T/Q chain over TU sizes 32/32/16/8 (+4x4 for the 8x8 level),
 * chroma at half size.  Everything stays in HBM; the host fetches cost tables, coefficients and reconstructions. */
typedef struct hb_prepass_cfg {
    int32_t qp;                 /* fixed slice QP */
    int32_t chroma_qp_offset;   /* HVENC_Cfg.chroma_qp_offset (reference default 2) */
    int32_t sign_hiding;
    int32_t is_islice;          /* 0 for the P-slice pre-pass */
    int32_t me_action;          /* HB_ME_PEL|HB_ME_HALF|HB_ME_QUARTER */
    int32_t use_graph;          /* replay the launch sequence as one CUDA graph */
    /* CTU-row band of the frame this GPU works on: rows [band_ctu_row0, band_ctu_row0+band_ctu_rows); 0,0 = all */
    int32_t band_ctu_row0, band_ctu_rows;
    /* != 0: hb_prepass_fetch_tables delivers (and hb_prepass_select expects) 12-byte records -- hb_me_result_c, hb_tu_result_c --
     * instead of the 24 / 16-byte hb_me_result / hb_tu_result: 30 % fewer bytes on the way to the host's decision */
    int32_t compact_tables;     /* 0 full records, 1 compact ME + TU records, 2 compact ME + per-CU records (hb_cu_cost) */
    /* 0 (default): the fifteen quarter-pel planes of the reference picture are built once per picture (one launch, TMA-staged tiles)
     * and every depth's sub-pel probes and luma predictions read them; != 0: each PU builds its own planes in shared memory around
     * its integer winner, as the reference does (hmr_motion_inter.c:395 / :442) -- same results, kept for comparison */
    int32_t subpel_per_pu;
    /* != 0 (needs the per-picture planes): a search CTA first stages the reference area its PUs' walks can reach (their strip of the
     * picture widened by +-128 x +-64) in shared memory -- one bulk asynchronous copy per window row on an mbarrier -- and probes it
     * there; 0 (default): every probe gathers its words from global memory through L1.  Same results.  Measured on a B200 at 1080p
     * (profiles/ncu_r02.txt): shared memory shares the L1TEX data pipe with the gathers it replaces (l1tex 79-86 % of peak -> 62-74 %) and the
     * staging latency is paid once per CTA at three CTAs per SM: 146 us of search per frame against 142 us, 6 170 frames/s against 6 430 */
    int32_t me_staged_window;
    /* 0 (default, needs the per-picture planes and no staged window): the search of a picture is ONE launch -- a CTA owns a CTU and walks
     * its 64x64, 32x32, 16x16 and 8x8 PUs in turn, the parent's vector handed down in shared memory; != 0: one launch per PU size, each
     * reading the previous size's result table (the round-1 structure; same results).  One dependent frame costs 4 x ~35 us of search
     * launches in the latter form (each an under-filled wave of a latency-bound chain) against one launch of all four chains */
    int32_t me_per_depth;
} hb_prepass_cfg;
/* the compact wire records (same order and counts as the full tables) */
typedef struct hb_me_result_c { int16_t mvx, mvy; uint32_t sad; uint16_t n_probes; int8_t subx, suby; } hb_me_result_c;   /* 12 bytes */
typedef struct hb_tu_result_c { uint32_t ssd, ssd_zero, sum_zeroed; } hb_tu_result_c;   /* sum in bits 0..30, zeroed in bit 31 */
/* compact_tables = 2: after the compact ME records, ONE record per coding unit and pass instead of one per transform unit -- what a
 * choice between partition depths needs.  Pass p = 0..4 in turn, the CUs of 64 >> min(p, 3) in raster order over whole CTUs (the
 * order of the ME table of that depth).  ssd / sum: totals over the CU's luma TUs of pass p and chroma TUs of pass min(p, 3);
 * cbf: coded flags of those TUs in raster order inside the CU, bits 0..3 luma, 4..7 U, 8..11 V. */
typedef struct hb_cu_cost { uint32_t ssd, sum; uint16_t cbf, reserved; } hb_cu_cost;   /* 12 bytes */
#define HB_PREPASS_DEPTHS 4     /* PU 64,32,16,8 */
#define HB_PREPASS_TQ_PASSES 5  /* luma TU 32(d0),32(d1),16(d2),8(d3),4(d3) */
int  hb_prepass_create(hb_ctx *ctx, int width, int height, const hb_prepass_cfg *cfg, hb_prepass **out);
void hb_prepass_destroy(hb_prepass *pp);
/* queue one frame (async); results become valid after hb_ctx_sync or a fetch */
int  hb_prepass_run(hb_prepass *pp, const hb_frame *cur, const hb_frame *ref, double avg_dist);
/* the same frame without the graph and with a CUDA event between consecutive launches: ms[i] = device time of kernel i.
 * Synchronises; returns the number of kernels (<= HB_PREPASS_MAX_KERNELS) or a negative error. */
#define HB_PREPASS_MAX_KERNELS 24
int  hb_prepass_run_profiled(hb_prepass *pp, const hb_frame *cur, const hb_frame *ref, double avg_dist, float *ms, int cap);
const char *hb_prepass_kernel_name(const hb_prepass *pp, int i);  /* "me64", "mc16", "tq2y16", ... valid after a profiled run */
int  hb_prepass_num_pus(const hb_prepass *pp, int depth);            /* PUs per frame at that depth (raster order) */
int  hb_prepass_num_tus(const hb_prepass *pp, int pass, int comp);   /* coded TUs of that pass and plane (raster order of the coded ones) */
int  hb_prepass_tu_size(const hb_prepass *pp, int pass, int comp);   /* TU side in samples of that plane, 0 = plane not coded in this pass */
int  hb_prepass_tu_xy(hb_prepass *pp, int pass, int comp, int32_t *xy_out); /* num_tus (x,y) pairs, samples of that plane */
int  hb_prepass_fetch_me(hb_prepass *pp, int depth, hb_me_result *out);                 /* PU raster order; sad = UINT32_MAX outside the frame/band */
int  hb_prepass_fetch_tu(hb_prepass *pp, int pass, int comp, hb_tu_result *out);        /* TU raster order */
int  hb_prepass_fetch_coeffs(hb_prepass *pp, int pass, int comp, int16_t *out);         /* num_tus * n*n levels, TU raster order, row-major inside */
int  hb_prepass_fetch_recon(hb_prepass *pp, int pass, uint8_t *y, int y_stride, uint8_t *u, int u_stride, uint8_t *v, int v_stride);
int  hb_prepass_fetch_all(hb_prepass *pp, void *pinned_dst, size_t cap, size_t *bytes_out); /* everything above, packed, one async burst + sync */
size_t hb_prepass_output_bytes(const hb_prepass *pp);
/* The host-decision flow (mode decision stays on the host and reads GPU cost tables): fetch the small cost tables, let the
 * host choose a depth per CTU, then gather only the reconstruction of that choice and the levels of its CODED TUs.
 * fetch_tables / gather are asynchronous on the context's stream (hb_ctx_sync before reading the pinned buffer). */
size_t hb_prepass_tables_bytes(const hb_prepass *pp);
int  hb_prepass_fetch_tables(hb_prepass *pp, void *pinned_dst, size_t cap);   /* ME tables d0..3, then TU tables pass 0..4 x Y,U,V */
int  hb_prepass_num_ctus(const hb_prepass *pp);
/* stand-in for the host's decision: per CTU the pass (0..4) minimising sum(ssd) + lambda*sum(|level|); also lays out the stream */
int  hb_prepass_select(const hb_prepass *pp, const void *tables, int lambda, uint8_t *sel, int32_t *ctu_off /* num_ctus + 1 */);
size_t hb_prepass_gather_bytes(const hb_prepass *pp, const int32_t *ctu_off);
/* The levels of the chosen passes in the reference's OWN hand-off layout to its entropy coder (ctu->coeff_wnd, hmr_encoder_lib.c:2945; read by
 * encode_residual, hmr_arithmetic_encoding.c:1087): per CTU HB_COEFF_WND_PER_CTU int16 = 64*64 luma, then 32*32 U, then 32*32 V; the N*N
 * levels of a transform unit lie row-major at offset abs_index << 4 in the luma window and (abs_index << 4) >> 2 in a chroma window,
 * abs_index = z-order number of the unit's first 4x4 luma block inside the CTU (hmr_motion_inter.c:73, :174); units without levels are
 * zero.  sel: chosen pass per CTU as for hb_prepass_gather.  Blocking; out: hb_prepass_num_ctus(pp) * HB_COEFF_WND_PER_CTU int16. */
#define HB_COEFF_WND_PER_CTU (64 * 64 + 2 * 32 * 32)
int hb_prepass_fetch_coeff_wnd(hb_prepass *pp, const uint8_t *sel, int16_t *out);

/* out: recon Y,U,V tight planes, then per CTU (from ctu_off[i], int16 units) for Y,U,V the coded TUs in raster order:
 * { hdr_lo, hdr_hi, N*N levels }, hdr = plane << 28 | N << 16 | TU raster position inside the CTU */
int  hb_prepass_gather(hb_prepass *pp, const uint8_t *sel, const int32_t *ctu_off, void *pinned_dst, size_t cap, size_t *bytes_out);
/* The device-resident continuation of the choice: the reconstruction of the chosen passes goes straight
 * into `rec`, which is then deblocked in place (hmr_deblocking_filter.c, strengths derived on the device from the plan's own
 * tables: CU / TU sizes of the pass, vectors, coded flags; fixed QP) and border-padded; the level streams go to pinned_levels
 * as in hb_prepass_gather (without the reconstruction in front).  Asynchronous.  SAO follows through hb_sao_stats_frame(cur, rec),
 * the host's decision and hb_sao_apply_frame(rec, next reference). */
int  hb_prepass_finalise(hb_prepass *pp, const uint8_t *sel, const int32_t *ctu_off, hb_frame *rec, const hb_deblock_params *dbk,
                         void *pinned_levels, size_t cap, size_t *bytes_out);
/* the unit data and strengths the last hb_prepass_finalise used (width/4 x height/4 each; any may be NULL).  Blocking. */
int  hb_prepass_fetch_units(hb_prepass *pp, hb_unit_info *units, uint8_t *bs_ver, uint8_t *bs_hor);
/* the whole per-frame host flow as one blocking call: upload cur + ref (tight, pinned host planes) -> pre-pass -> cost tables ->
 * hb_prepass_select -> gather -> results in `out` (pinned).  One encoder thread per in-flight stream calls this per frame. */
int  hb_prepass_process_frame(hb_prepass *pp, hb_frame *cur, hb_frame *ref, const uint8_t *const cur_planes[3], const uint8_t *const ref_planes[3],
                              double avg_dist, int lambda, void *tables, size_t tables_cap, uint8_t *sel, int32_t *ctu_off,
                              void *out, size_t out_cap, size_t *out_bytes);
/* the two halves of hb_prepass_process_frame: begin only queues work (uploads, pre-pass, table fetch) and returns; finish
 * blocks (tables -> decision -> gather -> results).  One host thread can begin frame n+1 on a second plan before it
 * finishes frame n. */
int  hb_prepass_frame_begin(hb_prepass *pp, hb_frame *cur, hb_frame *ref, const uint8_t *const cur_planes[3], const uint8_t *const ref_planes[3],
                            double avg_dist, void *tables, size_t tables_cap);
int  hb_prepass_frame_finish(hb_prepass *pp, int lambda, const void *tables, uint8_t *sel, int32_t *ctu_off, void *out, size_t out_cap, size_t *out_bytes);
/* The same per-frame flow with the reference picture kept on the device: only the source goes up; the cost tables, the level
 * streams and the SAO candidates (hb_sao_candidates_frame's records) come down; the finished picture (deblocked, SAO applied
 * with hb_sao_decide_from_candidates, border padded) is queued into `next_ref` for the following frame -- finish does not wait
 * for that last step (work queued on the same context afterwards is ordered behind it; hb_ctx_sync otherwise).
 * `rec` is a scratch frame; levels pinned, complete on return; params_out (optional): the SAO decision, num_ctus records. */
int  hb_prepass_frame_begin_resident(hb_prepass *pp, hb_frame *cur, hb_frame *ref, const uint8_t *const cur_planes[3], double avg_dist,
                                     void *tables, size_t tables_cap);
int  hb_prepass_frame_finish_resident(hb_prepass *pp, const hb_frame *cur, int lambda, const void *tables, uint8_t *sel, int32_t *ctu_off,
                                      hb_frame *rec, hb_frame *next_ref, const hb_deblock_params *dbk, const double sao_lambda[3],
                                      void *levels, size_t levels_cap, size_t *levels_bytes, hb_sao_param *params_out);
const hb_frame *hb_prepass_pred(const hb_prepass *pp, int depth);     /* resident prediction of that depth */
const hb_frame *hb_prepass_recon(const hb_prepass *pp, int pass);

/* ------------------------------------------------------------------ E. CU-granularity calls inside the encoder's own loop ------
 * The batched jobs of section C at the granularity of the reference's call sites, for a host loop that keeps every decision
 * (AMVP / merge candidates, TU tree, mode choice, CABAC): one round trip per call -- jobs up, kernels, the blocks the loop goes
 * on reading down, one wait.  A session owns four resident pictures; the prediction a T/Q job subtracts is whatever
 * hb_enc_predict last left at that place, as encode_inter_cu reads what hmr_motion_compensation_* last left in
 * et->prediction_wnd[0].  INTEGRATION.md section 2 shows the reference-side binding; tests/test_gpu_encode_hooks.py encodes whole
 * streams through it (oracle/ref_hooks.c) and compares them byte for byte with the unmodified reference.
 *   per picture (hook after hmr_rd_init, hmr_encoder_lib.c:3201): hb_frame_upload_i16(hb_enc_frame(e, HB_ENC_CUR), source wnd_t),
 *                                                            hb_frame_upload_i16(hb_enc_frame(e, HB_ENC_REF), reference wnd_t)
 *   hmr_motion_estimation, hmr_motion_inter.c:2625         -> hb_enc_me (the real AMVP list and start points in the jobs)
 *   hmr_motion_compensation_luma/_chroma, :3047-3049, :3655 -> hb_enc_predict
 *   encode_inter_cu / encode_inter_cu_chroma, :3165-3170    -> hb_enc_tq */
typedef struct hb_enc hb_enc;
#define HB_ENC_CUR 0
#define HB_ENC_REF 1
#define HB_ENC_PRED 2
#define HB_ENC_RECON 3
int       hb_enc_create(hb_ctx *ctx, int width, int height, hb_enc **out);
void      hb_enc_destroy(hb_enc *e);
hb_frame *hb_enc_frame(hb_enc *e, int which);
/* = hb_me_search on the session's source and reference pictures (no parent table: the caller's start list carries the parent's vector) */
int hb_enc_me(hb_enc *e, const hb_me_job *jobs, int n_jobs, double avg_dist, int action, hb_me_result *results);
/* uni-prediction of every job into the session's prediction picture; blocks: per job size^2 luma, (size/2)^2 U, (size/2)^2 V samples as
 * int16 (the sample type of the reference's CTU windows), jobs back to back */
int hb_enc_predict(hb_enc *e, const hb_mc_job *jobs, int n_jobs, int16_t *blocks);
/* the inter T/Q chain of every job (hb_tq_encode) on the session's source / prediction pictures into its reconstruction picture;
 * coeffs and decoded: size^2 int16 per job, jobs back to back; results[i].ssd is what encode_inter_cu returns, .sum its *curr_sum */
int hb_enc_tq(hb_enc *e, const hb_tu_job *jobs, int n_jobs, const hb_tq_params *params, int16_t *coeffs, int16_t *decoded, hb_tu_result *results);

#ifdef __cplusplus
}
#endif
#endif
