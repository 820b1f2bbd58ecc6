/* hb_sao_host.c -- the arithmetic half of the SAO decision, on the host (the decision itself prices the syntax with the CABAC
 * state and stays with the encoder).  hb_sao_derive_offsets restates, for 8-bit video,
 *   sao_derive_offsets        hmr_sao.c:480   initial offset = round(diff / count) clipped to +-7, sign rule per edge class,
 *                                             then the rate-distortion walk towards zero (est_iter_offset :445)
 *   sao_invert_quant_offsets  hmr_sao.c:592   (identity at 8 bits: the offset step is 1)
 *   sao_get_distortion        hmr_sao.c:620   sum of count*o*o - 2*diff*o over the classes in use
 * from the statistics hb_sao_stats_frame delivers.  Pure host code, no device work. */
#include "hb_host.h"
#include <string.h>
#include <stdlib.h>

#define SAO_MAX_OFFSET 7          /* (1 << (min(bit_depth, 10) - 5)) - 1, hmr_sao.c:70 */

static int64_t sao_dist(int64_t count, int64_t off, int64_t diff) { return count * off * off - 2 * diff * off; }

/* walk from the initial offset towards zero, keeping the cheapest dist + lambda * bits (est_iter_offset).  Starts from the cost
 * of sending a zero (lambda): returns 0 when nothing beats it, and then leaves *dist / *cost alone as the reference does. */
static int sao_iter_offset(int is_bo, double lambda, int start, int64_t count, int64_t diff, int64_t *dist, double *cost)
{
    double best = lambda;
    int out = 0;
    for (int o = start; o != 0; o += (o > 0) ? -1 : 1) {
        int bits = abs(o) + (is_bo ? 2 : 1);
        if (abs(o) == SAO_MAX_OFFSET) bits--;
        const int64_t d = sao_dist(count, o, diff);
        const double c = (double)d + lambda * (double)bits;
        if (c < best) { best = c; out = o; *dist = d; *cost = c; }
    }
    return out;
}

static int sao_initial_offset(int64_t diff, int64_t count)
{
    if (count == 0) return 0;
    const double x = (double)diff / (double)count;
    int o = x >= 0 ? (int)(x + 0.5) : (int)(x - 0.5);
    if (o < -SAO_MAX_OFFSET) o = -SAO_MAX_OFFSET;
    if (o > SAO_MAX_OFFSET) o = SAO_MAX_OFFSET;
    return o;
}

int hb_sao_derive_offsets(const hb_sao_stats *st, int type, double lambda, int16_t offset[32], int32_t *band, int64_t *dist)
{
    if (!st || !offset || !band || !dist) return hbi_fail(HB_ERR_ARG, "hb_sao_derive_offsets: NULL argument");
    if (type < 0 || type > 4) return hbi_fail(HB_ERR_ARG, "hb_sao_derive_offsets: type %d", type);
    memset(offset, 0, 32 * sizeof offset[0]);
    *band = 0; *dist = 0;
    if (type < 4) {
        for (int k = 0; k < 5; k++) {
            if (k == 2) continue;                                         /* the plain class never gets an offset */
            const int64_t cnt = st->eo_count[type][k], dif = st->eo_diff[type][k];
            int o = sao_initial_offset(dif, cnt);
            if (k < 2 && o < 0) o = 0;                                    /* valleys are only raised, peaks only lowered */
            if (k > 2 && o > 0) o = 0;
            if (o) { int64_t d = 0; double c = 0; o = sao_iter_offset(0, lambda, o, cnt, dif, &d, &c); }
            offset[k] = (int16_t)o;
            *dist += sao_dist(cnt, o, dif);
        }
        return HB_OK;
    }
    int q[32];
    double cost[32];
    for (int k = 0; k < 32; k++) {
        int64_t d = 0;
        cost[k] = lambda;
        q[k] = sao_initial_offset(st->bo_diff[k], st->bo_count[k]);
        if (q[k]) q[k] = sao_iter_offset(1, lambda, q[k], st->bo_count[k], st->bo_diff[k], &d, &cost[k]);
    }
    double best = (double)(0xffffffffu / 8);                              /* MAX_COST, hmr_private.h:54 */
    for (int b = 0; b <= 32 - 4; b++) {
        double c = cost[b];                                               /* same summation order as the reference */
        c += cost[b + 1]; c += cost[b + 2]; c += cost[b + 3];
        if (c < best) { best = c; *band = b; }
    }
    for (int k = *band; k < *band + 4; k++) {
        offset[k] = (int16_t)q[k];
        *dist += sao_dist(st->bo_count[k], q[k], st->bo_diff[k]);
    }
    return HB_OK;
}

/* Stand-in for sao_decide_blk_params (hmr_sao.c:1295): per CTU, luma alone and the two chroma planes jointly, the type that
 * minimises dist + lambda * bits against "off", with the constant syntax prices of the reference's COMPUTE_AS_HM branch
 * (8 bits for an edge type, 11 for the band type, off = 2.5 lambda; hmr_sao.c:690, :733, :776, :808) instead of the CABAC
 * estimate, and without the merge candidates.  stats: hb_sao_stats_frame's output; params: what hb_sao_apply_frame takes. */
int hb_sao_decide_standin(const hb_sao_stats *stats, int n_ctus, const double lambda[3], hb_sao_param *params)
{
    if (!stats || !lambda || !params || n_ctus < 0) return hbi_fail(HB_ERR_ARG, "hb_sao_decide_standin: bad argument");
    for (int i = 0; i < n_ctus; i++) {
        hb_sao_param *p = &params[i];
        memset(p, 0, sizeof *p);
        p->type[0] = p->type[1] = p->type[2] = -1;
        int16_t off[2][32];
        int32_t band;
        int64_t d[2];
        double best = 2.5 * lambda[0];
        for (int t = 0; t < 5; t++) {
            hb_sao_derive_offsets(&stats[i * 3], t, lambda[0], off[0], &band, &d[0]);
            const double c = (double)d[0] + lambda[0] * (t == 4 ? 11 : 8);
            if (c < best) { best = c; p->type[0] = (int8_t)t; memcpy(p->offset[0], off[0], sizeof off[0]); }
        }
        best = 2.5 * lambda[1];
        for (int t = 0; t < 5; t++) {
            double c = 0;
            for (int k = 0; k < 2; k++) {
                hb_sao_derive_offsets(&stats[i * 3 + 1 + k], t, lambda[1 + k], off[k], &band, &d[k]);
                c += (double)d[k];
                c += lambda[1 + k] * (t == 4 ? 11 : 8);
            }
            if (c < best) {
                best = c;
                for (int k = 0; k < 2; k++) { p->type[1 + k] = (int8_t)t; memcpy(p->offset[1 + k], off[k], sizeof off[k]); }
            }
        }
    }
    return HB_OK;
}

static void sao_candidate_to_param(const hb_sao_candidate *c, int type, int16_t offset[32])
{
    memset(offset, 0, 32 * sizeof offset[0]);
    if (type < 4) { offset[0] = c->offset[0]; offset[1] = c->offset[1]; offset[3] = c->offset[2]; offset[4] = c->offset[3]; }
    else for (int k = 0; k < 4; k++) offset[c->band + k] = c->offset[k];
}

/* the same stand-in from the candidates the device derived (hb_sao_candidates_frame): only the comparison of the five types is left */
int hb_sao_decide_from_candidates(const hb_sao_candidate *cand, int n_ctus, const double lambda[3], hb_sao_param *params)
{
    if (!cand || !lambda || !params || n_ctus < 0) return hbi_fail(HB_ERR_ARG, "hb_sao_decide_from_candidates: bad argument");
    for (int i = 0; i < n_ctus; i++) {
        hb_sao_param *p = &params[i];
        const hb_sao_candidate *c = cand + (size_t)i * 15;
        memset(p, 0, sizeof *p);
        p->type[0] = p->type[1] = p->type[2] = -1;
        double best = 2.5 * lambda[0];
        for (int t = 0; t < 5; t++) {
            const double v = (double)c[t].dist + lambda[0] * (t == 4 ? 11 : 8);
            if (v < best) { best = v; p->type[0] = (int8_t)t; }
        }
        if (p->type[0] >= 0) sao_candidate_to_param(&c[p->type[0]], p->type[0], p->offset[0]);
        best = 2.5 * lambda[1];
        int bt = -1;
        for (int t = 0; t < 5; t++) {
            double v = 0;
            for (int k = 0; k < 2; k++) { v += (double)c[5 * (1 + k) + t].dist; v += lambda[1 + k] * (t == 4 ? 11 : 8); }
            if (v < best) { best = v; bt = t; }
        }
        for (int k = 0; k < 2 && bt >= 0; k++) { p->type[1 + k] = (int8_t)bt; sao_candidate_to_param(&c[5 * (1 + k) + bt], bt, p->offset[1 + k]); }
    }
    return HB_OK;
}
