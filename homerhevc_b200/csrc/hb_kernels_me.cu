// hb_kernels_me.cu -- motion search of one PU per thread group, replaying hmr_motion_estimation
// (hmr_motion_inter.c:1404-1774) exactly: integer walk (start points, small diamond, rotating big diamond,
// iterated small diamond) with the reference's double-precision MV cost, then the 8+8 half/quarter-pel probes.
//
// Mapping: CTA = 256 threads; a PU is searched by GW warps (64x64: 8, 32x32: 2, 16x16 and 8x8: 1), so the
// CTA holds 8/GW PUs.  The current block lives in registers as packed u8x4 words; a candidate SAD is
// __vsadu4 over unaligned 4-sample reads of the resident reference plane (L2/L1 resident, funnel-shifted),
// reduced with redux.sync and, across warps of a group, one shared-memory exchange + named barrier.
// Sub-pel: the (N+8)x(N+12) reference patch around the integer winner is staged in shared memory once,
// the three horizontal 14-bit planes (fractions 1,2,3) are built from it, and each candidate's prediction
// is the vertical pass over those planes -- the same two-stage arithmetic as the reference's plane builders
// (:395, :442), sample for sample.
#include "hb_shim.h"
#include "hb_dev_common.cuh"

namespace {

__constant__ int8_t c_small[4][2] = { {-1, 0}, {0, -1}, {1, 0}, {0, 1} };
__constant__ int8_t c_big[8][2] = { {-2, 0}, {-1, -1}, {0, -2}, {1, -1}, {2, 0}, {1, 1}, {0, 2}, {-1, 1} };
__constant__ int8_t c_half[9][2] = { {0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1} };
__constant__ int8_t c_quarter[9][2] = { {0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1} };

struct MeArgs {
    hbd_plane cur, ref;
    const hbd_me_job *jobs;
    int n_jobs;
    const hb_me_result *parent;
    hb_me_result *out;
    int action;
    const hbd_dyn_params *dyn;
};

template <int N, int GW> struct MeCfg {
    static constexpr int GT = GW * 32;                 // threads per PU
    static constexpr int PUS = 8 / GW;                 // PUs per CTA
    static constexpr int WPR = N / 4;                  // packed words per row
    static constexpr int NW = N * N / 4;               // packed words per PU
    static constexpr int WPL = (NW + GT - 1) / GT;     // words per lane
    static constexpr int PROWS = N + 8;                // patch rows: iy-4 .. iy+N+3
    static constexpr int PS = N + 12;                  // patch row stride in bytes: ix-4 .. ix+N+7
    static constexpr int TS = N + 4;                   // plane row stride in int16: column j <-> x = ix-1+j (+f/4)
    static constexpr int PATCH_BYTES = PROWS * PS;
    static constexpr int PLANE_ELEMS = PROWS * TS;
    static constexpr int SMEM_PER_PU = PATCH_BYTES + 3 * PLANE_ELEMS * 2;
};

template <int GW> __device__ __forceinline__ void group_barrier(int group)
{
    if (GW == 1) __syncwarp();
    else if (GW == 8) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(GW * 32) : "memory");
}

template <int N, int GW>
__global__ void __launch_bounds__(256) k_me(const MeArgs a)
{
    using Cfg = MeCfg<N, GW>;
    constexpr int GT = Cfg::GT, PUS = Cfg::PUS, WPR = Cfg::WPR, NW = Cfg::NW, WPL = Cfg::WPL;
    constexpr int PS = Cfg::PS, TS = Cfg::TS, PROWS = Cfg::PROWS;

    __shared__ __align__(16) uint8_t s_raw[PUS * Cfg::SMEM_PER_PU];
    __shared__ uint32_t s_red[PUS][2][GW];

    const int group = threadIdx.x / GT, gl = threadIdx.x % GT, lane = threadIdx.x & 31, gwarp = gl >> 5;
    const int job_idx = blockIdx.x * PUS + group;
    if (job_idx >= a.n_jobs) return;                   // whole group leaves together (GW == 8 -> whole CTA)
    hbd_me_job job = a.jobs[job_idx];
    if (a.dyn) job.corr = a.dyn->corr;

    uint8_t *s_patch = s_raw + group * Cfg::SMEM_PER_PU;
    int16_t *s_plane = reinterpret_cast<int16_t *>(s_patch + Cfg::PATCH_BYTES);   // [3][PROWS][TS], fraction f -> plane f-1
    int red_phase = 0;

    // ---- current block -> registers
    uint32_t cur[WPL];
    int wrow[WPL], wcol[WPL];
#pragma unroll
    for (int k = 0; k < WPL; k++) {
        const int w = gl + k * GT;
        wrow[k] = w / WPR; wcol[k] = (w % WPR) * 4;
        cur[k] = 0;
        if (w < NW) cur[k] = hb_ld_u8x4(a.cur.org + (job.y + wrow[k]) * a.cur.pitch + job.x + wcol[k]);
    }
    const uint8_t *ref_pu = a.ref.org + job.y * a.ref.pitch + job.x;
    const int rpitch = a.ref.pitch;

    auto group_sum = [&](uint32_t v) -> uint32_t {
        v = __reduce_add_sync(HB_FULL_MASK, v);
        if (GW == 1) return v;
        if (lane == 0) s_red[group][red_phase][gwarp] = v;
        group_barrier<GW>(group);
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < GW; i++) s += s_red[group][red_phase][i];
        red_phase ^= 1;
        return s;
    };
    auto sad_at = [&](int dx, int dy) -> uint32_t {
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < WPL; k++) {
            if (gl + k * GT < NW) {
                const uint32_t r = hb_ld_u8x4(ref_pu + (dy + wrow[k]) * rpitch + dx + wcol[k]);
                acc = __vsadu4(cur[k], r) + acc;
            }
        }
        return group_sum(acc);
    };
    // select_mv_candidate_fast (hmr_motion_inter.c:1004): IEEE double, products and sums rounded separately
    auto mv_cost = [&](int mvx, int mvy) -> uint32_t {
        uint32_t best = 0x7fffffffu;
        for (int i = 0; i < job.n_amvp; i++) {
            const double cx = __dmul_rn(job.corr, static_cast<double>(static_cast<float>(abs(job.amvp[2 * i] - mvx))));
            const double cy = __dmul_rn(job.corr, static_cast<double>(static_cast<float>(abs(job.amvp[2 * i + 1] - mvy))));
            const uint32_t c = __double2uint_rz(__dadd_rn(__dadd_rn(cx, cy), 0.5));
            if (best > c) best = c;
        }
        return best;
    };

    const int fw = a.cur.w, fh = a.cur.h;
    const int xlo = (job.x - 128 < 0) ? -job.x : -128;
    const int xhi = (job.x + 128 > fw - N) ? fw - job.x - N : 128;
    const int ylo = (job.y - 64 < 0) ? -job.y : -64;
    const int yhi = (job.y + 64 > fh - N) ? fh - job.y - N : 64;

    int bx = 0, by = 0;
    uint32_t bsad = 0, brd = 0, n_probes = 0;
    auto probe = [&](int x, int y) -> bool {
        if (x < xlo || x > xhi || y < ylo || y > yhi) return false;
        const uint32_t sad = sad_at(x, y);
        const uint32_t rd = sad + mv_cost(x << 2, y << 2);
        n_probes++;
        if (rd < brd) { bsad = sad; brd = rd; bx = x; by = y; return true; }
        return false;
    };

    int mvx = 0, mvy = 0, subx = 0, suby = 0;
    uint32_t best_sad = 0xffffffffu / 8;

    if (a.action & HB_ME_PEL) {
        bx = min(max(0, xlo), xhi); by = min(max(0, ylo), yhi);
        bsad = sad_at(bx, by); n_probes++;
        brd = bsad + mv_cost(bx << 2, by << 2);
        int cx0 = bx, cy0 = by;
        bool skip = bsad == 0;
        if (!skip) {
            // extra start points: caller's list, then the parent PU's vector when both components are non-zero
            for (int i = 0; i < job.n_start; i++) {
                const int x = job.start[2 * i] >> 2, y = job.start[2 * i + 1] >> 2;
                if (x == 0 && y == 0) continue;
                probe(x, y);
            }
            if (job.parent >= 0) {
                const hb_mv pmv = a.parent[job.parent].mv;
                if (pmv.x != 0 && pmv.y != 0) {
                    const int x = pmv.x >> 2, y = pmv.y >> 2;
                    if (!(x == 0 && y == 0)) probe(x, y);
                }
            }
            cx0 = bx; cy0 = by;
            skip = bsad == 0;
        }
        if (!skip) {
            for (int i = 0; i < 4; i++) probe(cx0 + c_small[i][0], cy0 + c_small[i][1]);
            int dist = 2;
            const int end = (cx0 != 0 && cy0 != 0) ? 4 : 8;
            int next_start = 0, span = 8;
            cx0 = bx; cy0 = by;
            while (dist < end) {
                for (int i = next_start; i < next_start + span; i++) {
                    const int idx = i & 7;
                    if (probe(cx0 + c_big[idx][0] * dist, cy0 + c_big[idx][1] * dist)) {
                        next_start = (idx - 2 + 8) & 7;
                        span = 5;
                    }
                }
                dist *= 2;
            }
        }
        cx0 = bx; cy0 = by;
        {
            int next_start = 0, span = 4;
            for (;;) {
                for (int i = next_start; i < next_start + span; i++) {
                    const int idx = i & 3;
                    if (probe(cx0 + c_small[idx][0], cy0 + c_small[idx][1])) {
                        next_start = (idx - 1 + 4) & 3;
                        span = 3;
                    }
                }
                if (cx0 == bx && cy0 == by) break;
                cx0 = bx; cy0 = by;
            }
        }
        best_sad = bsad;
        mvx = bx << 2; mvy = by << 2;
    }

    if (a.action & HB_ME_HALF) {
        const int ix = mvx >> 2, iy = mvy >> 2;
        const uint8_t *ref_i = ref_pu + iy * rpitch + ix;
        uint32_t cur_best = (a.action & HB_ME_PEL) ? bsad : sad_at(ix, iy);

        // ---- stage the patch: rows iy-4.., columns ix-4..
        for (int w = gl; w < PROWS * (PS / 4); w += GT) {
            const int r = w / (PS / 4), c = (w % (PS / 4)) * 4;
            *reinterpret_cast<uint32_t *>(s_patch + r * PS + c) = hb_ld_u8x4(ref_i + (r - 4) * rpitch + (c - 4));
        }
        group_barrier<GW>(group);
        // ---- horizontal 14-bit planes for fractions 1,2,3 : T_f[r][j] = sum taps_f[k] * P[r][j+k] - 8192
        for (int w = gl; w < PROWS * (TS / 4); w += GT) {
            const int r = w / (TS / 4), j0 = (w % (TS / 4)) * 4;
            int p[11];
#pragma unroll
            for (int k = 0; k < 11; k++) p[k] = s_patch[r * PS + j0 + k];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                s_plane[0 * Cfg::PLANE_ELEMS + r * TS + j0 + q] = static_cast<int16_t>(hb_luma8<1>(p[q], p[q + 1], p[q + 2], p[q + 3], p[q + 4], p[q + 5], p[q + 6], p[q + 7]) - 8192);
                s_plane[1 * Cfg::PLANE_ELEMS + r * TS + j0 + q] = static_cast<int16_t>(hb_luma8<2>(p[q], p[q + 1], p[q + 2], p[q + 3], p[q + 4], p[q + 5], p[q + 6], p[q + 7]) - 8192);
                s_plane[2 * Cfg::PLANE_ELEMS + r * TS + j0 + q] = static_cast<int16_t>(hb_luma8<3>(p[q], p[q + 1], p[q + 2], p[q + 3], p[q + 4], p[q + 5], p[q + 6], p[q + 7]) - 8192);
            }
        }
        group_barrier<GW>(group);

        // SAD of the current block against the prediction at quarter-pel offset (cx,cy) in [-3,3]^2 from (ix,iy)
        auto subpel_sad = [&](int cx, int cy) -> uint32_t {
            const int fx = cx & 3, fy = cy & 3;
            const int cb = cx >> 2, rb = cy >> 2;              // -1 or 0
            uint32_t acc = 0;
#pragma unroll
            for (int k = 0; k < WPL; k++) {
                if (gl + k * GT < NW) {
                    const int j = wcol[k] + cb + 1;            // plane column of the first of 4 samples
                    const int r0 = wrow[k] + rb + 4;           // plane row of the sample itself
                    int px[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        int t[8];
                        if (fy == 0) {
                            const int v = fx ? static_cast<int>(s_plane[(fx - 1) * Cfg::PLANE_ELEMS + r0 * TS + j + q])
                                             : (static_cast<int>(s_patch[r0 * PS + j + q + 3]) << 6) - 8192;
                            px[q] = hb_clip255((v + 8192 + 32) >> 6);
                        } else {
#pragma unroll
                            for (int m = 0; m < 8; m++) {
                                const int rr = r0 - 3 + m;
                                t[m] = fx ? static_cast<int>(s_plane[(fx - 1) * Cfg::PLANE_ELEMS + rr * TS + j + q])
                                          : (static_cast<int>(s_patch[rr * PS + j + q + 3]) << 6) - 8192;
                            }
                            const int s = hb_luma8_dyn(fy, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7]);
                            px[q] = hb_clip255((s + 2048 + (8192 << 6)) >> 12);
                        }
                    }
                    acc = __vsadu4(cur[k], hb_pack4(px[0], px[1], px[2], px[3])) + acc;
                }
            }
            return group_sum(acc);
        };

        int sbx = 0, sby = 0, bidx = 0;
        for (int i = 1; i < 9; i++) {                      // candidate 0 is the integer position itself: never smaller
            const int cx = c_half[i][0] * 2, cy = c_half[i][1] * 2;
            const uint32_t v = subpel_sad(cx, cy);
            if (v < cur_best) { cur_best = v; sbx = cx; sby = cy; bidx = i; }
        }
        if (a.action & HB_ME_QUARTER) {
            const int hx = c_half[bidx][0], hy = c_half[bidx][1];
            for (int i = 1; i < 9; i++) {                  // candidate 0 repeats the half-pel winner
                const int cx = hx * 2 + c_quarter[i][0], cy = hy * 2 + c_quarter[i][1];
                const uint32_t v = subpel_sad(cx, cy);
                if (v < cur_best) { cur_best = v; sbx = cx; sby = cy; }
            }
        }
        best_sad = cur_best;
        mvx = (ix << 2) + sbx; mvy = (iy << 2) + sby;
        subx = sbx; suby = sby;
    }

    if (gl == 0) {
        hb_me_result r;
        r.mv.x = mvx; r.mv.y = mvy; r.subpix.x = subx; r.subpix.y = suby; r.sad = best_sad; r.n_probes = n_probes;
        a.out[job.out] = r;
    }
}

template <int N, int GW> int launch_me(const MeArgs &a, cudaStream_t s)
{
    const int grid = (a.n_jobs + MeCfg<N, GW>::PUS - 1) / MeCfg<N, GW>::PUS;
    k_me<N, GW><<<grid, 256, 0, s>>>(a);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace

extern "C" int hbk_me_search(const hbd_frame *cur, const hbd_frame *ref, int size, const hbd_me_job *jobs, int n_jobs,
                             const hb_me_result *parent, hb_me_result *out, int action, const hbd_dyn_params *dyn, void *stream)
{
    if (n_jobs <= 0) return 0;
    MeArgs a;
    a.cur = cur->p[0]; a.ref = ref->p[0]; a.jobs = jobs; a.n_jobs = n_jobs; a.parent = parent; a.out = out; a.action = action; a.dyn = dyn;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (size) {
    case 64: return launch_me<64, 8>(a, s);
    case 32: return launch_me<32, 2>(a, s);
    case 16: return launch_me<16, 1>(a, s);
    case 8: return launch_me<8, 1>(a, s);
    default: return static_cast<int>(cudaErrorInvalidValue);
    }
}
