/*
 * ref_shadow.c -- TEST INFRASTRUCTURE: a function table for the reference encoder in which every member runs TWICE, once as the
 * reference's own SSE4.2 function (whose result the encoder goes on with) and once as libhomer_b200's per-call drop-in on
 * private copies of the same operands; the two results are compared on the spot.  A whole encode through this table pins the first
 * call -- function, arguments, operands -- on which the GPU path and the CPU path ever disagree, which a comparison of finished
 * bitstreams cannot (tools/flake_hunt.py loops it to look for rare, timing-dependent differences).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hmr_private.h"
#include "hmr_common.h"
#include "hmr_sse42_functions.h"

typedef struct gpu_quant_env { int32_t is_islice, sign_hiding, max_cu_size_shift, bit_depth; int16_t *delta_u; } gpu_quant_env;

static struct {
    uint32_t (*sad)(int16_t *, uint32_t, int16_t *, uint32_t, int);
    uint32_t (*ssd16b)(int16_t *, uint32_t, int16_t *, uint32_t, int);
    void (*predict)(int16_t *, int, int16_t *, int, int16_t *, int, int);
    void (*reconst)(int16_t *, int, int16_t *, int, int16_t *, int, int);
    void (*interp_luma)(int16_t *, int, int16_t *, int, int, int, int, int, int, int);
    void (*interp_chroma)(int16_t *, int, int16_t *, int, int, int, int, int, int, int);
    void (*transform)(int, int16_t *, int16_t *, int, int, int, int, int, uint16_t, int16_t *);
    void (*itransform)(int, int16_t *, int16_t *, int, int, int, unsigned int, int16_t *);
    void (*quant)(const gpu_quant_env *, int16_t *, int16_t *, int, int, int, int, int, int *, int, int, int);
    void (*inv_quant)(const gpu_quant_env *, int16_t *, int16_t *, int, int, int, int, int, int);
} GPU;

#define SH_MAX_DUMP (80 * 80)
typedef struct shadow_report {
    long calls, mismatches;
    int first_fn;                 /* 1 sad 2 ssd16b 3 predict 4 reconst 5 interp luma 6 interp chroma 7 transform 8 itransform 9 quant 10 inv_quant */
    int args[12];
    long first_call;              /* ordinal of the first mismatching call */
    int first_at;                 /* index of the first differing output element */
    int32_t cpu_val, gpu_val;
    long per_fn[12];              /* mismatching calls per function (index = fn) */
} shadow_report;
static shadow_report S;
static __thread int16_t t_a[SH_MAX_DUMP], t_b[SH_MAX_DUMP];

static void report(int fn, const int *args, int n_args, int at, int cpu, int gpu)
{
    __atomic_add_fetch(&S.per_fn[fn], 1, __ATOMIC_RELAXED);
    if (__atomic_add_fetch(&S.mismatches, 1, __ATOMIC_RELAXED) != 1) return;
    S.first_fn = fn; S.first_call = S.calls; S.first_at = at; S.cpu_val = cpu; S.gpu_val = gpu;
    memset(S.args, 0, sizeof S.args);
    for (int i = 0; i < n_args && i < 12; i++) S.args[i] = args[i];
}
#define CALL() __atomic_add_fetch(&S.calls, 1, __ATOMIC_RELAXED)

static int cmp_block(const int16_t *a, int as, const int16_t *b, int bs, int w, int h, int *cv, int *gv)
{
    for (int r = 0; r < h; r++) for (int c = 0; c < w; c++)
        if (a[r * as + c] != b[r * bs + c]) { *cv = a[r * as + c]; *gv = b[r * bs + c]; return r * w + c; }
    return -1;
}

static uint32_t sh_sad(int16_t *src, uint32_t ss, int16_t *pred, uint32_t ps, int size)
{
    CALL();
    const uint32_t c = sse_aligned_sad(src, ss, pred, ps, size), g = GPU.sad(src, ss, pred, ps, size);
    if (c != g) { const int a[3] = { (int)ss, (int)ps, size }; report(1, a, 3, 0, (int)c, (int)g); }
    return c;
}
static uint32_t sh_ssd(int16_t *src, uint32_t ss, int16_t *pred, uint32_t ps, int size)
{
    CALL();
    const uint32_t c = sse_aligned_ssd16b(src, ss, pred, ps, size), g = GPU.ssd16b(src, ss, pred, ps, size);
    if (c != g) { const int a[3] = { (int)ss, (int)ps, size }; report(2, a, 3, 0, (int)c, (int)g); }
    return c;
}
static void sh_predict(int16_t *o, int os, int16_t *p, int ps, int16_t *r, int rs, int size)
{
    int cv, gv, at;
    CALL();
    GPU.predict(o, os, p, ps, t_a, size, size);
    sse_aligned_predict(o, os, p, ps, r, rs, size);
    if ((at = cmp_block(r, rs, t_a, size, size, size, &cv, &gv)) >= 0) { const int a[4] = { os, ps, rs, size }; report(3, a, 4, at, cv, gv); }
}
static void sh_reconst(int16_t *p, int ps, int16_t *r, int rs, int16_t *d, int ds, int size)
{
    int cv, gv, at;
    CALL();
    GPU.reconst(p, ps, r, rs, t_a, size, size);
    sse_aligned_reconst(p, ps, r, rs, d, ds, size);
    if ((at = cmp_block(d, ds, t_a, size, size, size, &cv, &gv)) >= 0) { const int a[4] = { ps, rs, ds, size }; report(4, a, 4, at, cv, gv); }
}
static void sh_interp(int chroma, int16_t *ref, int rs, int16_t *dst, int ds, int frac, int w, int h, int vert, int first, int last)
{
    int cv, gv, at;
    CALL();
    if (w * h <= SH_MAX_DUMP) (chroma ? GPU.interp_chroma : GPU.interp_luma)(ref, rs, t_a, w, frac, w, h, vert, first, last);
    (chroma ? sse_interpolate_chroma : sse_interpolate_luma)(ref, rs, dst, ds, frac, w, h, vert, first, last);
    if (w * h <= SH_MAX_DUMP && (at = cmp_block(dst, ds, t_a, w, w, h, &cv, &gv)) >= 0) {
        const int a[8] = { rs, ds, frac, w, h, vert, first, last };
        report(chroma ? 6 : 5, a, 8, at, cv, gv);
    }
}
static void sh_interp_luma(int16_t *ref, int rs, int16_t *dst, int ds, int frac, int w, int h, int vert, int first, int last) { sh_interp(0, ref, rs, dst, ds, frac, w, h, vert, first, last); }
static void sh_interp_chroma(int16_t *ref, int rs, int16_t *dst, int ds, int frac, int w, int h, int vert, int first, int last) { sh_interp(1, ref, rs, dst, ds, frac, w, h, vert, first, last); }
static void sh_transform(int bd, int16_t *block, int16_t *coeff, int bs, int w, int h, int wsh, int hsh, uint16_t mode, int16_t *aux)
{
    int cv, gv, at;
    CALL();
    GPU.transform(bd, block, t_a, bs, w, h, wsh, hsh, mode, t_b);
    sse_transform(bd, block, coeff, bs, w, h, wsh, hsh, mode, aux);
    if ((at = cmp_block(coeff, w, t_a, w, w, h, &cv, &gv)) >= 0) { const int a[4] = { bs, w, h, mode }; report(7, a, 4, at, cv, gv); }
}
static void sh_itransform(int bd, int16_t *block, int16_t *coeff, int bs, int w, int h, unsigned int mode, int16_t *aux)
{
    int cv, gv, at;
    CALL();
    GPU.itransform(bd, t_a, coeff, w, w, h, mode, t_b);
    sse_itransform(bd, block, coeff, bs, w, h, mode, aux);
    if ((at = cmp_block(block, bs, t_a, w, w, h, &cv, &gv)) >= 0) { const int a[4] = { bs, w, h, (int)mode }; report(8, a, 4, at, cv, gv); }
}
static void sh_quant(henc_thread_t *et, int16_t *src, int16_t *dst, int scan_mode, int depth, int comp, int cu_mode, int is_intra, int *ac_sum, int cu_size, int per, int rem)
{
    int cv, gv, at, gsum = 0;
    gpu_quant_env env = { et->enc_engine->current_pict.slice.slice_type == I_SLICE, (int32_t)et->pps->sign_data_hiding_flag, et->max_cu_size_shift, et->bit_depth, t_b };
    CALL();
    GPU.quant(&env, src, t_a, scan_mode, depth, comp, cu_mode, is_intra, &gsum, cu_size, per, rem);
    sse_aligned_quant(et, src, dst, scan_mode, depth, comp, cu_mode, is_intra, ac_sum, cu_size, per, rem);
    at = cmp_block(dst, cu_size, t_a, cu_size, cu_size, cu_size, &cv, &gv);
    if (at < 0 && gsum != *ac_sum) { at = cu_size * cu_size; cv = *ac_sum; gv = gsum; }
    if (at >= 0) { const int a[9] = { scan_mode, depth, comp, cu_mode, is_intra, cu_size, per, rem, env.is_islice }; report(9, a, 9, at, cv, gv); }
}
static void sh_inv_quant(henc_thread_t *et, short *src, short *dst, int depth, int comp, int is_intra, int cu_size, int per, int rem)
{
    int cv, gv, at;
    gpu_quant_env env = { 0, 0, et->max_cu_size_shift, et->bit_depth, NULL };
    CALL();
    GPU.inv_quant(&env, src, t_a, depth, comp, is_intra, cu_size, per, rem);
    sse_aligned_inv_quant(et, src, dst, depth, comp, is_intra, cu_size, per, rem);
    if ((at = cmp_block(dst, cu_size, t_a, cu_size, cu_size, cu_size, &cv, &gv)) >= 0) { const int a[6] = { depth, comp, is_intra, cu_size, per, rem }; report(10, a, 6, at, cv, gv); }
}

/* a refdrv_table_hook: user = { lib (dlopen handle of libhomer_b200.so), which (ignored) } */
void refdrv_install_shadow_table(void *funcs_table, void *user)
{
    struct { void *lib; int which; } *u = user;
    low_level_funcs_t *f = (low_level_funcs_t *)funcs_table;
    void *lib = u->lib;
    memset(&S, 0, sizeof S);
    GPU.sad = dlsym(lib, "hb_sad"); GPU.ssd16b = dlsym(lib, "hb_ssd16b"); GPU.predict = dlsym(lib, "hb_predict"); GPU.reconst = dlsym(lib, "hb_reconst");
    GPU.interp_luma = dlsym(lib, "hb_interpolate_luma"); GPU.interp_chroma = dlsym(lib, "hb_interpolate_chroma");
    GPU.transform = dlsym(lib, "hb_transform"); GPU.itransform = dlsym(lib, "hb_itransform"); GPU.quant = dlsym(lib, "hb_quant"); GPU.inv_quant = dlsym(lib, "hb_inv_quant");
    f->sad = sh_sad; f->ssd16b = sh_ssd; f->predict = sh_predict; f->reconst = sh_reconst;
    f->interpolate_luma_m_compensation = sh_interp_luma; f->interpolate_luma_m_estimation = sh_interp_luma; f->interpolate_chroma_m_compensation = sh_interp_chroma;
    f->transform = sh_transform; f->itransform = sh_itransform; f->quant = sh_quant; f->inv_quant = sh_inv_quant;
}
void *refdrv_install_shadow_table_addr(void) { return (void *)refdrv_install_shadow_table; }
void refdrv_shadow_report(shadow_report *out) { *out = S; }

/* ------------------------------------------------------------------------------------------------------------
 * Range audit (CPU only): which of the reference's own table calls hand sad / ssd16b operands that are NOT 8-bit video?  There
 * the SSE4.2 functions (16-bit lanes, saturating adds, hmr_sse42_functions_pixel.c:353-435) and the arithmetic definition
 * (hmr_motion_intra.c:51) give different numbers, so a bit-exact replacement is only defined on video-range operands
 * (SURVEY.md 8a a1).  The audit forwards every call to the SSE4.2 function, counts the calls with an operand outside
 * [-255, 510] (the widest range real data reaches: bi-prediction's 2*orig - pred) and keeps the call stack of the first one.
 * ------------------------------------------------------------------------------------------------------------ */
#include <execinfo.h>
typedef struct audit_report { long sad_calls, sad_out_of_range, ssd_calls, ssd_out_of_range; int first_size, first_min, first_max; long first_call; char first_stack[1024]; } audit_report;
static audit_report A;

static int block_range(const int16_t *p, int stride, int n, int *mn, int *mx)
{
    int lo = 32767, hi = -32768;
    for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) { const int v = p[r * stride + c]; if (v < lo) lo = v; if (v > hi) hi = v; }
    if (lo < *mn) *mn = lo;
    if (hi > *mx) *mx = hi;
    return lo < -255 || hi > 510;
}
static void audit_note(int size, int mn, int mx, long ordinal)
{
    void *bt[16];
    if (A.first_stack[0]) return;
    A.first_size = size; A.first_min = mn; A.first_max = mx; A.first_call = ordinal;
    const int n = backtrace(bt, 16);
    size_t used = 0;
    for (int i = 1; i < n && used + 64 < sizeof A.first_stack; i++) {
        Dl_info info;
        const char *name = (dladdr(bt[i], &info) && info.dli_sname) ? info.dli_sname : "?";
        used += (size_t)snprintf(A.first_stack + used, sizeof A.first_stack - used, "%s%s", i > 1 ? " < " : "", name);
    }
}
static uint32_t au_sad(int16_t *src, uint32_t ss, int16_t *pred, uint32_t ps, int size)
{
    int mn = 32767, mx = -32768;
    A.sad_calls++;
    if (A.sad_calls == 1 && getenv("HB_AUDIT_FIRST_CALL")) { block_range(src, (int)ss, size, &mn, &mx); block_range(pred, (int)ps, size, &mn, &mx); audit_note(size, mn, mx, 1); }
    if (block_range(src, (int)ss, size, &mn, &mx) | block_range(pred, (int)ps, size, &mn, &mx)) { A.sad_out_of_range++; audit_note(size, mn, mx, A.sad_calls); }
    return sse_aligned_sad(src, ss, pred, ps, size);
}
static uint32_t au_ssd(int16_t *src, uint32_t ss, int16_t *pred, uint32_t ps, int size)
{
    int mn = 32767, mx = -32768;
    A.ssd_calls++;
    if (block_range(src, (int)ss, size, &mn, &mx) | (ps ? block_range(pred, (int)ps, size, &mn, &mx) : 0)) A.ssd_out_of_range++;
    return sse_aligned_ssd16b(src, ss, pred, ps, size);
}
void refdrv_install_audit_table(void *funcs_table, void *user)
{
    low_level_funcs_t *f = (low_level_funcs_t *)funcs_table;
    (void)user;
    memset(&A, 0, sizeof A);
    f->sad = au_sad; f->ssd16b = au_ssd;
}
void *refdrv_install_audit_table_addr(void) { return (void *)refdrv_install_audit_table; }
void refdrv_audit_report(audit_report *out) { *out = A; }

/* ------------------------------------------------------------------------------------------------------------
 * Differential trace of sad / ssd16b: pass 1 (record) runs an encode on the reference's own functions and writes one record per
 * call -- function, size, strides, result, hashes of both operands -- to a file; pass 2 (compare) runs the same encode on the GPU
 * drop-ins alone and checks every call against the file on the fly.  The first record that differs says whether the GPU function
 * returned another value for the SAME operands, or whether the encode had already taken another path (other operands / sizes).
 * user = { lib (NULL: record with the SSE4.2 functions; else compare with the library's), path of the trace file }
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct trace_rec { int32_t fn, size; uint32_t ss, ps, result, hsrc, hpred; } trace_rec;
typedef struct trace_report { long calls; long first_diff; trace_rec want, got; char stack[1024]; } trace_report;
static struct { FILE *f; int compare; trace_report rep; uint32_t (*sad)(int16_t *, uint32_t, int16_t *, uint32_t, int); uint32_t (*ssd)(int16_t *, uint32_t, int16_t *, uint32_t, int); } T;

static uint32_t hash_block(const int16_t *p, int stride, int n)
{
    uint32_t h = 2166136261u;
    for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) { h ^= (uint16_t)p[(stride ? r * stride : 0) + c]; h *= 16777619u; }
    return h;
}
static uint32_t trace_call(int fn, int16_t *src, uint32_t ss, int16_t *pred, uint32_t ps, int size)
{
    trace_rec r = { fn, size, ss, ps, 0, 0, 0 };
    const int n = (size == 4 || size == 8 || size == 16 || size == 32 || size == 64) ? size : 0;
    r.hsrc = hash_block(src, (int)ss, n); r.hpred = hash_block(pred, (int)ps, n);
    r.result = (fn == 1 ? T.sad : T.ssd)(src, ss, pred, ps, size);
    T.rep.calls++;
    if (!T.compare) { fwrite(&r, sizeof r, 1, T.f); return r.result; }
    trace_rec w;
    /* call 1 is skipped: the 64x64 prediction buffer of the very first mode probe holds whatever memory held before (it differs
     * between two runs of the unmodified encoder as well) */
    if (T.rep.first_diff == 0 && fread(&w, sizeof w, 1, T.f) == 1 && T.rep.calls > 1 && memcmp(&w, &r, sizeof r)) {
        void *bt[16];
        T.rep.first_diff = T.rep.calls; T.rep.want = w; T.rep.got = r;
        const int k = backtrace(bt, 16);
        size_t used = 0;
        for (int i = 1; i < k && used + 64 < sizeof T.rep.stack; i++) {
            Dl_info info;
            used += (size_t)snprintf(T.rep.stack + used, sizeof T.rep.stack - used, "%s%s", i > 1 ? " < " : "", (dladdr(bt[i], &info) && info.dli_sname) ? info.dli_sname : "?");
        }
    }
    return r.result;
}
static uint32_t tr_sad(int16_t *s, uint32_t ss, int16_t *p, uint32_t ps, int n) { return trace_call(1, s, ss, p, ps, n); }
static uint32_t tr_ssd(int16_t *s, uint32_t ss, int16_t *p, uint32_t ps, int n) { return trace_call(2, s, ss, p, ps, n); }
void refdrv_install_trace_table(void *funcs_table, void *user)
{
    struct { void *lib; const char *path; } *u = user;
    low_level_funcs_t *f = (low_level_funcs_t *)funcs_table;
    if (T.f) fclose(T.f);
    memset(&T, 0, sizeof T);
    T.compare = u->lib != NULL;
    T.f = fopen(u->path, T.compare ? "rb" : "wb");
    T.sad = T.compare ? (uint32_t (*)(int16_t *, uint32_t, int16_t *, uint32_t, int))dlsym(u->lib, "hb_sad") : sse_aligned_sad;
    T.ssd = T.compare ? (uint32_t (*)(int16_t *, uint32_t, int16_t *, uint32_t, int))dlsym(u->lib, "hb_ssd16b") : sse_aligned_ssd16b;
    f->sad = tr_sad; f->ssd16b = tr_ssd;
}
void *refdrv_install_trace_table_addr(void) { return (void *)refdrv_install_trace_table; }
void refdrv_trace_report(trace_report *out) { if (T.f) { fclose(T.f); T.f = NULL; } *out = T.rep; }
