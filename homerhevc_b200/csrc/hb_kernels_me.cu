// hb_kernels_me.cu -- motion search of one PU per thread group, replaying hmr_motion_estimation
// (hmr_motion_inter.c:1404-1774) exactly: integer walk (start points, small diamond, rotating big diamond,
// iterated small diamond) with the reference's double-precision MV cost, then the 8+8 half/quarter-pel probes.
//
// The reference probes one position after the other, but every probe of a stage lies on a pattern around a centre
// that is fixed for that stage, and a probe's SAD/cost does not depend on the order.  So each stage is computed as
// ROUNDS OF FOUR CANDIDATES IN PARALLEL (all lanes busy, one reduction + one cost evaluation per round) and the
// reference's sequential compare/rotate logic is then replayed on the four (or eight) results in registers --
// including which probes it would have skipped, so even the probe count matches.
//
// Mapping: CTA = 256 threads; a PU is searched by G lanes (8x8: 8 lanes, 16x16: 16 lanes -- four / two PUs share a warp and
// run their data-dependent loops in lock step, so every warp primitive names the full warp; 32x32: two warps; 64x64: the CTA),
// split into four candidate slots of L = G/4 lanes.  A lane keeps 8-sample pairs of the current block in registers; a
// candidate SAD reads the resident reference plane as three aligned words per pair (one misalignment shift per candidate,
// 32-bit offsets from one base) into VABSDIFF4.ACC on two chains, is summed inside the slot (redux / shuffle butterfly) and
// exchanged by shuffles (G <= 32) or through shared memory.  In the pattern stages a lane derives and range-checks only its
// own slot's position; whether the other slots were valid comes back with their cost.
// Sub-pel: the (N+8)x(N+12) patch around the integer winner is staged in shared memory by row runs; the four horizontal
// 14-bit planes (fractions 0..3) are dp4a products of u8 samples and s8 taps, stored PAIR-INTERLEAVED (one word = rows 2q,
// 2q+1 of a column) so that every vertical filter is dp2a on one shared load per two taps.  Half-pel: a lane walks one plane
// column and feeds each filtered sample to all candidates that share it; quarter-pel: four candidates per round; four clipped
// samples are packed (I2IP) and compared in one VABSDIFF4.  Same two-stage arithmetic as the reference's plane builders
// (:395, :442), sample for sample; the winner's luma prediction is written from the same planes.
#include <cstring>
#include "hb_shim.h"
#include "hb_dev_common.cuh"

namespace {

__constant__ int8_t c_small[4][2] = { {-1, 0}, {0, -1}, {1, 0}, {0, 1} };
__constant__ int8_t c_big[8][2] = { {-2, 0}, {-1, -1}, {0, -2}, {1, -1}, {2, 0}, {1, 1}, {0, 2}, {-1, 1} };
__constant__ int8_t c_half[9][2] = { {0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1} };
__constant__ int8_t c_quarter[9][2] = { {0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1} };
// the same table for compile-time indices (folds to immediates)
__host__ __device__ constexpr int quarter_off(int i, int c)
{
    constexpr int t[9][2] = { {0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1} };
    return t[i][c];
}

struct MeArgs {
    hbd_plane cur, ref;
    const hbd_me_job *jobs;
    int n_jobs;
    const hb_me_result *parent;
    hb_me_result *out;
    int action;
    const hbd_dyn_params *dyn;
    hbd_plane pred;             // luma plane receiving the prediction of the winning vector (org == nullptr: not wanted)
    hbd_subpel sp;              // PL kernels: the reference picture's quarter-pel planes (hb_kernels_subpel.cu)
};

template <int N> struct MeCfg {
    static constexpr int G = (N == 64) ? 256 : (N == 32) ? 64 : N;    // lanes per PU (8x8: 8, 16x16: 16 -> several PUs per warp)
    static constexpr int L = G / 4;                    // lanes per candidate slot
    static constexpr int SEG = L < 32 ? L : 32;        // lanes that reduce together with one redux
    static constexpr int NSEG = G / SEG;               // partial sums exchanged per round
    static constexpr int SEG_PER_SLOT = L / SEG;
    static constexpr int PUS = 256 / G;                // PUs per CTA
    static constexpr int WPR = N / 4;                  // packed words per row
    static constexpr int NW = N * N / 4;               // packed words per PU
    static constexpr int WPL = NW / L;                 // words per lane (of its slot's candidate)
    static constexpr int CPL = N / L;                  // sub-pel: columns per lane
    static constexpr int PROWS = N + 8;                // patch rows: iy-4 .. iy+N+3
    static constexpr int PS = N + 12;                  // patch row stride in bytes: ix-4 .. ix+N+7
    static constexpr int TS = N + 4;                   // plane row stride in int16: column j <-> x = ix-1+j (+f/4)
    static constexpr int PATCH_BYTES = PROWS * PS;
    static constexpr int PLANE_ELEMS = PROWS * TS;
    static constexpr int QROWS = PROWS / 2;            // plane rows are stored in pairs: one 32-bit word = rows 2q, 2q+1 of a column
    static constexpr int PLANE_WORDS = QROWS * TS;
    static constexpr int CS = N + 4;                   // column stride (bytes) of the transposed current block
    static constexpr int SMEM_PER_PU = PATCH_BYTES + 4 * PLANE_ELEMS * 2;
    static constexpr int SMEM_TOTAL = PUS * SMEM_PER_PU;
    // WIN kernels: the CTA's PUs are PUS neighbours of one PU row (a strip of PUS * N samples); every position their walks can probe
    // lies in the strip widened by the search range, +-128 x +-64 (hmr_private.h:76-77), clipped to the picture
    static constexpr int WIN_ROWS = N + 128;
    static constexpr int WIN_P = PUS * N + 256 + 16;       // bytes per window row: a multiple of 16, 4 or 20 words mod 32 (row-to-row bank shift)
    static constexpr int WIN_BYTES = WIN_ROWS * WIN_P;
};

// Vertical: the planes hold two rows per word, so an 8-tap column sum over rows rho0..rho0+7 is four dp2a when rho0 is even
// (tap pairs E_i = (t2i, t2i+1)) and five when it is odd (O_i = (t(2i-1), t2i), with t(-1) = t8 = 0).  A lane produces TWO
// consecutive output rows from the same five words W[m..m+4]: for par = 0 (first tap row even) these are E | O, for par = 1
// O | (0, E).  vtap(f, par, i) packs the pair of the first row in the low two bytes (dp2a.lo) and of the second in the high two.
__host__ __device__ constexpr int tap_or0(int f, int k) { return (k < 0 || k > 7) ? 0 : luma_tap(f, k); }
__host__ __device__ constexpr uint32_t vtap(int f, int par, int i)
{
    // pair of row A (first tap row par + 2m): taps start at half `par` of word m; row B: one row further down
    const int a0 = 2 * i - par, b0 = 2 * i - par - 1;
    return pack_s8(tap_or0(f, a0), tap_or0(f, a0 + 1), tap_or0(f, b0), tap_or0(f, b0 + 1));
}
#define HB_VTAP5(f, p) { vtap(f, p, 0), vtap(f, p, 1), vtap(f, p, 2), vtap(f, p, 3), vtap(f, p, 4) }
__constant__ uint32_t c_vtab[4][2][5] = { { HB_VTAP5(0, 0), HB_VTAP5(0, 1) }, { HB_VTAP5(1, 0), HB_VTAP5(1, 1) },
                                          { HB_VTAP5(2, 0), HB_VTAP5(2, 1) }, { HB_VTAP5(3, 0), HB_VTAP5(3, 1) } };
constexpr int kVRound = 2048 + (8192 << 6);            // second-pass rounding + the 14-bit offset of the first pass

// one column of the half-pel stage: the fraction-2 filter down the N+1 positions rho - 1/2 (rho = 0..N) of plane column `pl`
// (pair-interleaved words, row stride TS).  v[rho] is compared with current row rho ("minus" candidates, y = -2) and with
// current row rho-1 ("plus", y = +2) of the left (block column j-1) and, for T_2, right (column j) neighbours; T_2 also gives
// the horizontal-only half-pel sample u[rho] = round(T_2[rho+4]).  Four rows per packed compare.
template <int N, int TS, bool T2>
__device__ __forceinline__ void half_strip(const uint32_t *pl, const uint8_t *cur_l, const uint8_t *cur_r, uint32_t (&acc)[6])
{
    uint32_t win[5];
#pragma unroll
    for (int k = 0; k < 4; k++) win[k] = pl[k * TS];
    uint32_t prev_v = 0, prev_l = 0, prev_r = 0;
#pragma unroll
    for (int k4 = 0; k4 <= N / 4; k4++) {
        const bool last = k4 == N / 4;                 // only row N: feeds the plus candidates
        int v[4] = { 0, 0, 0, 0 }, u[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int m = 2 * k4 + h;
            if (last && h == 1) break;
            if (!last) win[(m + 4) % 5] = pl[(m + 4) * TS];
            int se = kVRound;
#pragma unroll
            for (int i = 0; i < 4; i++) se = __dp2a_lo(static_cast<int>(win[(m + i) % 5]), static_cast<int>(vtap(2, 0, i)), se);
            v[2 * h] = se >> 12;
            if (!last) {
                int so = kVRound;
#pragma unroll
                for (int i = 0; i < 5; i++) so = __dp2a_hi(static_cast<int>(win[(m + i) % 5]), static_cast<int>(vtap(2, 0, i)), so);
                v[2 * h + 1] = so >> 12;
                if (T2) {
                    const int w = static_cast<int>(win[(m + 2) % 5]);          // rows rho+4 of the two outputs
                    u[2 * h] = __dp2a_lo(w, 0x0001, 8192 + 32) >> 6;
                    u[2 * h + 1] = __dp2a_lo(w, 0x0100, 8192 + 32) >> 6;
                }
            }
        }
        const uint32_t pv = hb_pack_sat_u8x4(v[0], v[1], v[2], v[3]);
        if (k4 >= 1) {
            const uint32_t vs = __funnelshift_r(prev_v, pv, 8);               // rows 4(k4-1)+1 .. 4(k4-1)+4
            acc[2] = hb_sad4_acc(vs, prev_l, acc[2]);
            if (T2) acc[3] = hb_sad4_acc(vs, prev_r, acc[3]);
        }
        if (!last) {
            const uint32_t l4 = *reinterpret_cast<const uint32_t *>(cur_l + 4 * k4);
            acc[0] = hb_sad4_acc(pv, l4, acc[0]);
            prev_l = l4;
            if (T2) {
                const uint32_t r4 = *reinterpret_cast<const uint32_t *>(cur_r + 4 * k4);
                const uint32_t pu = hb_pack_sat_u8x4(u[0], u[1], u[2], u[3]);
                acc[1] = hb_sad4_acc(pv, r4, acc[1]);
                acc[4] = hb_sad4_acc(pu, l4, acc[4]);
                acc[5] = hb_sad4_acc(pu, r4, acc[5]);
                prev_r = r4;
            }
        }
        prev_v = pv;
    }
}

// barrier over the lanes that search one PU.  Sub-warp groups (several PUs per warp) run in lock step, so the whole warp
// synchronises; two-warp groups use a named barrier each, the 64x64 PU is the CTA.
template <int G> __device__ __forceinline__ void group_barrier(int group, uint32_t gmask)
{
    if (G <= 32) __syncwarp();
    else if (G == 256) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(G) : "memory");
}

template <int K> __device__ __forceinline__ uint32_t pick(const uint32_t (&a)[K], int i)
{
    uint32_t v = a[0];
#pragma unroll
    for (int k = 1; k < K; k++) v = (i == k) ? a[k] : v;
    return v;
}

// does any valid candidate beat the best cost so far?  (static register indexing, unlike the replay loop's pick)
template <int K> __device__ __forceinline__ bool beats(const uint32_t (&rd)[K], uint32_t vmask, uint32_t brd)
{
    uint32_t mn = 0xffffffffu;
#pragma unroll
    for (int s = 0; s < K; s++) mn = min(mn, ((vmask >> s) & 1u) ? rd[s] : 0xffffffffu);
    return mn < brd;
}
// number of valid positions among the `span` consecutive ones (mod K) that the replay loop visits from next_start
template <int K> __device__ __forceinline__ uint32_t visited(uint32_t vmask, int next_start, int span)
{
    const uint32_t m = (1u << span) - 1u;
    return __popc(vmask & ((m << next_start) | (m >> (K - next_start))) & ((1u << K) - 1u));
}

// one PU's inputs, in registers.  out < 0: a lock-step filler (the PU is searched, nothing is written)
struct MePu {
    int x, y;
    int n_amvp, a0x, a0y, a1x, a1y;
    int n_start, start[6];
    bool has_parent; int pmvx, pmvy;     // the parent PU's winning vector (quarter-pel)
    int out;
    double corr;
};

// PL = false: the sub-pel stage builds its planes per PU in shared memory (the reference's own scheme); PL = true: it reads the
// reference picture's fifteen quarter-pel planes, built once per picture, and needs no shared memory beyond the exchange slots
// WIN = true (pre-pass, PL only): the reference area the integer walk can reach is staged in shared memory first -- one bulk
// asynchronous copy (cp.async.bulk, the TMA engine) per window row, all completing on one mbarrier -- and every probe of the walk
// reads it from there instead of gathering 32-bit words through L1 (which was the limiter: l1tex 80-86 % of peak, ~8-10 sectors
// per request).  The CTA's jobs then are PUS neighbours of one PU row; entries with x < 0 pad the last strip of a row.
// PRE = true: the pre-pass form known at compile time -- two zero AMVP predictors, no caller start points.
// s_x: exchange slots of the groups wider than a warp, [group][phase][segment]{partial SAD, cost}; s_half: [group][8] (PL = false only)
// ---- SAD pyramid of a CTU (single-launch search).  Every position the 64x64 PU probes -- integer or sub-pel, keyed by its quarter-pel
// displacement -- leaves the SADs of the CTU's sixty-four 8x8 blocks at that displacement (the partial sums its own SAD is made of) in
// shared memory; a smaller PU that probes the same displacement later (it mostly does: same origin, parent's vector, same small
// patterns -- tools/me_reuse_study.py: 80-99 % of the probes of the children on the benchmark clip, 60-70 % under sigma-30 noise)
// adds up its blocks instead of reading the reference again.  Same samples, same sums: results cannot differ.
struct MeMemo {
    static constexpr int ENTRIES = 64, SLOTS = 256;
    uint32_t sum[ENTRIES][32];      // word (j, bx) = blocks (by = 2j, bx) | (by = 2j + 1, bx) << 16, j = 0..3... see block()
    uint32_t tag[ENTRIES];          // quarter-pel displacement: (x & 0xffff) | y << 16
    uint32_t idx[SLOTS / 4];        // hash slot -> entry, one byte each, 0xff = empty; claimed by compare-and-swap on the word
    static __device__ __forceinline__ uint32_t key(int qx, int qy) { return (static_cast<uint32_t>(qx) & 0xffffu) | (static_cast<uint32_t>(qy) << 16); }
    static __device__ __forceinline__ uint32_t hash(uint32_t k) { return (k * 2654435761u) >> 24; }
};
// MEMO: 0 no pyramid, 1 this PU is the CTU (N = 64) and fills it, 2 this PU looks its probes up (blk = its raster index inside the CTU)
template <int N, bool PL, bool WIN, bool PRE, int MEMO = 0>
__device__ __forceinline__ void me_pu(const MeArgs &a, const MePu &j, const int group, const int gl, uint8_t *s_raw, uint32_t *s_x, uint32_t *s_half,
                                      const int wx0, const int wy0, int &win_mvx, int &win_mvy, MeMemo *memo = nullptr, const int blk = 0)
{
    static_assert(MEMO != 1 || N == 64, "the pyramid is filled by the 64x64 PU");
    using Cfg = MeCfg<N>;
    constexpr int G = Cfg::G, L = Cfg::L, SEG = Cfg::SEG, NSEG = Cfg::NSEG, SPS = Cfg::SEG_PER_SLOT;
    constexpr int WPR = Cfg::WPR, WPL = Cfg::WPL, CPL = Cfg::CPL, PS = Cfg::PS, TS = Cfg::TS, PROWS = Cfg::PROWS;
    constexpr int QROWS = Cfg::QROWS, PLANE_WORDS = Cfg::PLANE_WORDS, CS = Cfg::CS;

    const int lane = threadIdx.x & 31;
    const int slot = gl / L, l = gl % L, seg = gl / SEG;
    const int jx = j.x, jy = j.y;
    const double corr = a.dyn ? a.dyn->corr : j.corr;
    const int n_amvp = PRE ? 2 : j.n_amvp;
    const int a0x = PRE ? 0 : j.a0x, a0y = PRE ? 0 : j.a0y, a1x = PRE ? 0 : j.a1x, a1y = PRE ? 0 : j.a1y;
    const bool two_costs = !PRE && n_amvp > 1 && (a0x != a1x || a0y != a1y);
    auto sx_at = [&](int ph, int sg, int k) -> uint32_t & { return s_x[((group * 2 + ph) * NSEG + sg) * 2 + k]; };
    // three consecutive words of a reference / plane row.  64x64 PUs (eight lanes side by side on a row) take them as two 64-bit loads,
    // 10 % faster there; for the smaller PUs, whose lanes scatter over rows, the wider requests cost more than the saved one
    // (measured: 16x16 34.8 -> 42.9 us, 8x8 36.9 -> 41.9 us per launch)
    auto ld3 = [&](const uint8_t *base, uint32_t off, uint32_t &w0, uint32_t &w1, uint32_t &w2) {
        if constexpr (N == 64) hb_ld_words3(base, off, w0, w1, w2);
        else { const uint32_t *q = reinterpret_cast<const uint32_t *>(base + off); w0 = __ldg(q); w1 = __ldg(q + 1); w2 = __ldg(q + 2); }
    };

    // MEMO == 2: the SAD of this PU at quarter-pel displacement (qx, qy) from the pyramid, if the CTU probed it
    auto memo_get = [&](int qx, int qy, uint32_t &val) -> bool {
        const uint32_t k = MeMemo::key(qx, qy), h = MeMemo::hash(k);
        uint32_t id = (memo->idx[h >> 2] >> ((h & 3u) * 8u)) & 0xffu;
        if (id == 0xffu) return false;
        if (memo->tag[id] != k) {                      // second home of a key: the neighbouring slot
            id = (memo->idx[(h ^ 1u) >> 2] >> (((h ^ 1u) & 3u) * 8u)) & 0xffu;
            if (id == 0xffu || memo->tag[id] != k) return false;
        }
        const uint32_t *e = memo->sum[id];
        constexpr int SIDE = 64 / N;
        const int bx = blk % SIDE, by = blk / SIDE;
        if constexpr (N == 8) {
            const uint32_t w = e[(by >> 1) * 8 + bx];
            val = (by & 1) ? (w >> 16) : (w & 0xffffu);
        } else if constexpr (N == 16) {
            const uint32_t w0 = e[by * 8 + 2 * bx], w1 = e[by * 8 + 2 * bx + 1];
            val = (w0 & 0xffffu) + (w0 >> 16) + (w1 & 0xffffu) + (w1 >> 16);
        } else {
            uint32_t t = 0;
#pragma unroll
            for (int jj = 0; jj < 2; jj++)
#pragma unroll
                for (int c = 0; c < 4; c++) { const uint32_t w = e[(2 * by + jj) * 8 + 4 * bx + c]; t += (w & 0xffffu) + (w >> 16); }
            val = t;
        }
        return true;
    };
    // MEMO == 1: lane l of a slot holds, per pair k, the 8 samples of row (l / 8) + 8k in block column l % 8, i.e. one row of block
    // (by = k, bx = l % 8).  p[k] = its SAD over them; the eight lanes l % 8 == bx (rows 0..7 of the blocks) add up by a reduce-scatter
    // (xor 8, xor 16: three shuffles) and the two warps of the slot by shared-memory atomics into entry `id`.
    int memo_n = 0;                                    // entries handed out so far (four per round, one per slot)
    int phase = 0;                                     // double buffer of the exchange slots (and of the pyramid's pending entries)
    auto memo_put = [&](const uint32_t (&pk)[4], int qx, int qy, bool valid) {
        const int id = memo_n + slot;
        if (id < MeMemo::ENTRIES && valid) {           // uniform per slot (two warps)
            const bool b3 = (l & 8) != 0, b4 = (l & 16) != 0;
            const uint32_t s0 = __shfl_xor_sync(HB_FULL_MASK, b3 ? pk[0] : pk[2], 8), s1 = __shfl_xor_sync(HB_FULL_MASK, b3 ? pk[1] : pk[3], 8);
            const uint32_t k0 = (b3 ? pk[2] : pk[0]) + s0, k1 = (b3 ? pk[3] : pk[1]) + s1;       // words j = 2 b3, 2 b3 + 1
            const uint32_t s2 = __shfl_xor_sync(HB_FULL_MASK, b4 ? k0 : k1, 16);
            const uint32_t kk = (b4 ? k1 : k0) + s2;                                              // word j = 2 b3 + b4
            atomicAdd(&memo->sum[id][((b3 ? 2 : 0) + (b4 ? 1 : 0)) * 8 + (l & 7)], kk);
        }
        // the slot's first lane files the entry: its tag, then one of the key's two hash slots claimed by compare-and-swap (the four slots of a
        // round file concurrently).  A displacement probed again in a later round finds its first entry (filed before an earlier barrier) and
        // keeps it; two slots of ONE round never probe the same position.
        if (l == 0 && id < MeMemo::ENTRIES && valid) {
            const uint32_t k = MeMemo::key(qx, qy), h = MeMemo::hash(k);
            memo->tag[id] = k;
            bool filed = false;
#pragma unroll
            for (int alt = 0; alt < 2 && !filed; alt++) {
                const uint32_t hh = h ^ static_cast<uint32_t>(alt), sh = (hh & 3u) * 8u;
                uint32_t *wp = &memo->idx[hh >> 2];
                uint32_t cur = atomicOr(wp, 0u);
                for (;;) {
                    const uint32_t e = (cur >> sh) & 0xffu;
                    if (e != 0xffu) { filed = static_cast<int>(e) < memo_n && memo->tag[e] == k; break; }      // taken: by this displacement's earlier entry?
                    const uint32_t prev = atomicCAS(wp, cur, (cur & ~(0xffu << sh)) | (static_cast<uint32_t>(id) << sh));
                    if (prev == cur) { filed = true; break; }
                    cur = prev;
                }
            }
        }
    };

    uint8_t *s_patch = s_raw + group * Cfg::SMEM_PER_PU;
    uint32_t *s_plane = reinterpret_cast<uint32_t *>(s_patch + Cfg::PATCH_BYTES);  // [4][QROWS][TS] by x fraction, rows in pairs
    // G < 32: several PUs share a warp.  Their data-dependent loops are run in LOCK STEP (a PU that is finished idles through the
    // rounds the others still need -- the hardware would serialise them anyway), so every warp primitive names the full warp
    // and compiles to a single instruction; groups are addressed through the width argument of the shuffles.
    constexpr bool LOCKSTEP = G < 32;
    auto any_pu = [&](bool p) -> bool { if constexpr (LOCKSTEP) return __any_sync(HB_FULL_MASK, p); else return p; };
    constexpr uint32_t gmask = HB_FULL_MASK;

    // ---- current block -> registers: lane l of every slot holds the 8-sample pairs l, l+L, ... (two packed words each).
    // Pair k sits (L/PPR) rows below pair k-1 in the same columns, so reference addresses advance by one constant.
    constexpr int PPR = WPR / 2, PPL = WPL / 2;        // pairs per row / per lane
    static_assert(L % PPR == 0, "pairs of a lane share their columns");
    uint32_t cur[WPL];
    const uint32_t rpitch = static_cast<uint32_t>(a.ref.pitch);
    const int prow0 = l / PPR, pcol = (l % PPR) * 8;
    // byte offset of this lane's first pair from the start of the reference allocation (unsigned: one uniform base + 32-bit offsets)
    const uint32_t ref_lane = static_cast<uint32_t>(jy + a.ref.pad + prow0) * rpitch + static_cast<uint32_t>(jx + a.ref.pad + pcol);
    const uint32_t ref_step = (L / PPR) * rpitch;
#pragma unroll
    for (int k = 0; k < PPL; k++) {
        const uint8_t *c = a.cur.org + (jy + prow0 + k * (L / PPR)) * a.cur.pitch + jx + pcol;
        cur[2 * k] = hb_ld_u8x4(c);
        cur[2 * k + 1] = hb_ld_u8x4(c + 4);
    }

    // select_mv_candidate_fast (hmr_motion_inter.c:1004): IEEE double, products and sums rounded separately
    auto mv_cost = [&](int mvx, int mvy) -> uint32_t {
        if (n_amvp <= 0) return 0x7fffffffu;
        const double c0 = __dadd_rn(__dadd_rn(__dmul_rn(corr, static_cast<double>(static_cast<float>(abs(a0x - mvx)))),
                                              __dmul_rn(corr, static_cast<double>(static_cast<float>(abs(a0y - mvy))))), 0.5);
        uint32_t best = min(0x7fffffffu, __double2uint_rz(c0));
        if (two_costs) {
            const double c1 = __dadd_rn(__dadd_rn(__dmul_rn(corr, static_cast<double>(static_cast<float>(abs(a1x - mvx)))),
                                                  __dmul_rn(corr, static_cast<double>(static_cast<float>(abs(a1y - mvy))))), 0.5);
            best = min(best, __double2uint_rz(c1));
        }
        return best;
    };

    const int fw = a.cur.w, fh = a.cur.h;
    const int xlo = (jx - 128 < 0) ? -jx : -128;
    const int xhi = (jx + 128 > fw - N) ? fw - jx - N : 128;
    const int ylo = (jy - 64 < 0) ? -jy : -64;
    const int yhi = (jy + 64 > fh - N) ? fh - jy - N : 64;
    auto inside = [&](int x, int y) { return x >= xlo && x <= xhi && y >= ylo && y <= yhi; };

    // exchange this lane's partial value: every lane gets the four slot totals (and the four slot costs)
    auto exchange = [&](uint32_t part, uint32_t cost, uint32_t (&tot)[4], uint32_t (&cst)[4]) {
        if constexpr (SEG == 32) {
            part = __reduce_add_sync(HB_FULL_MASK, part);
        } else {                                           // redux.sync with a partial mask is emulated: butterfly instead
#pragma unroll
            for (int d = SEG / 2; d > 0; d >>= 1) part += __shfl_xor_sync(HB_FULL_MASK, part, d);
        }
        if constexpr (G <= 32) {
#pragma unroll
            for (int s = 0; s < 4; s++) {
                tot[s] = __shfl_sync(HB_FULL_MASK, part, s * L, G);
                cst[s] = __shfl_sync(HB_FULL_MASK, cost, s * L, G);
            }
        } else {
            if ((gl & (SEG - 1)) == 0) { sx_at(phase, seg, 0) = part; sx_at(phase, seg, 1) = cost; }
            group_barrier<G>(group, gmask);
#pragma unroll
            for (int s = 0; s < 4; s++) {
                uint32_t t = 0;
#pragma unroll
                for (int q = 0; q < SPS; q++) t += sx_at(phase, s * SPS + q, 0);
                tot[s] = t;
                cst[s] = sx_at(phase, s * SPS, 1);
            }
            phase ^= 1;
        }
    };
    // SAD of this lane's words at integer displacement (mx, my)
    // this lane's first pair inside the staged window (WIN)
    const uint32_t win_lane = static_cast<uint32_t>((jy - wy0 + prow0) * Cfg::WIN_P + (jx - wx0 + pcol));
    // SAD of this lane's pairs against the rows at base + off (row step `step`); MEMO == 1: the per-pair sums packed two to a word in pk
    auto sad_rows = [&](const uint8_t *base, uint32_t off, const uint32_t step, uint32_t (&pk)[4]) -> uint32_t {
        // the pitch is a multiple of 4, so every word of this candidate has the same misalignment: shift once
        const uint32_t sh = (off & 3u) * 8u;
        off &= ~3u;
        if constexpr (MEMO == 1) {
            uint32_t tot = 0;
#pragma unroll
            for (int k = 0; k < PPL; k++) {
                uint32_t w0, w1, w2;
                ld3(base, off, w0, w1, w2);
                const uint32_t pr = hb_sad4_acc(cur[2 * k + 1], __funnelshift_r(w1, w2, sh), hb_sad4_acc(cur[2 * k], __funnelshift_r(w0, w1, sh), 0u));
                if (k & 1) pk[k / 2] |= pr << 16; else pk[k / 2] = pr;
                tot += pr;
                off += step;
            }
            return tot;
        } else {
            uint32_t acc = 0, acc1 = 0;
#pragma unroll
            for (int k = 0; k < PPL; k++) {
                uint32_t w0, w1, w2;
                ld3(base, off, w0, w1, w2);
                acc = hb_sad4_acc(cur[2 * k], __funnelshift_r(w0, w1, sh), acc);
                acc1 = hb_sad4_acc(cur[2 * k + 1], __funnelshift_r(w1, w2, sh), acc1);       // two chains for ILP
                off += step;
            }
            return acc + acc1;
        }
    };
    auto sad_at = [&](int mx, int my, uint32_t (&pk)[4]) -> uint32_t {
        if constexpr (WIN) {
            uint32_t acc = 0, acc1 = 0;
            uint32_t off = win_lane + static_cast<uint32_t>(my * Cfg::WIN_P + mx);
            const uint32_t sh = (off & 3u) * 8u;
            off &= ~3u;
#pragma unroll
            for (int k = 0; k < PPL; k++) {
                const uint32_t *q = reinterpret_cast<const uint32_t *>(s_raw + off);
                const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
                acc = hb_sad4_acc(cur[2 * k], __funnelshift_r(w0, w1, sh), acc);
                acc1 = hb_sad4_acc(cur[2 * k + 1], __funnelshift_r(w1, w2, sh), acc1);
                off += (L / PPR) * Cfg::WIN_P;
            }
            return acc + acc1;
        }
        return sad_rows(a.ref.base, ref_lane + static_cast<uint32_t>(my * static_cast<int>(rpitch) + mx), ref_step, pk);
    };
    // one probe of this lane's slot at integer (mx, my): from the CTU's pyramid when it is there (MEMO == 2: lane 0 of the slot carries
    // the sum), else from the reference picture; MEMO == 1 leaves the block sums in the pyramid
    auto probe_int = [&](int mx, int my, bool mv) -> uint32_t {
        uint32_t pk[4] = { 0, 0, 0, 0 }, acc = 0;
        if (mv) {
            if constexpr (MEMO == 2) {
                uint32_t v;
                if (memo_get(mx << 2, my << 2, v)) acc = (l == 0) ? v : 0u;
                else acc = sad_at(mx, my, pk);
            } else acc = sad_at(mx, my, pk);
        }
        if constexpr (MEMO == 1) { memo_put(pk, mx << 2, my << 2, mv); memo_n += 4; }
        return acc;
    };
    // one round: SAD + cost of up to four integer positions (cx[s], cy[s]); invalid slots give garbage that is never read
    auto round4 = [&](const int (&cx)[4], const int (&cy)[4], const bool (&cv)[4], uint32_t (&sad)[4], uint32_t (&rd)[4]) {
        int mx = cx[0], my = cy[0]; bool mv = cv[0];
#pragma unroll
        for (int s = 1; s < 4; s++) if (slot == s) { mx = cx[s]; my = cy[s]; mv = cv[s]; }
        const uint32_t acc = probe_int(mx, my, mv);
        uint32_t cst[4];
        exchange(acc, mv_cost(mx << 2, my << 2), sad, cst);
#pragma unroll
        for (int s = 0; s < 4; s++) rd[s] = sad[s] + cst[s];
    };
    // the same for the pattern stages: every lane derives ITS slot's position (mx, my) itself and tests it alone; whether the
    // other three slots were inside the search area comes back with their cost (no cost is ever 0xffffffff)
    auto round_own = [&](int mx, int my, bool mv, uint32_t (&sad)[4], uint32_t (&rd)[4], uint32_t &vmask) {
        const uint32_t acc = probe_int(mx, my, mv);
        uint32_t cst[4];
        exchange(acc, mv ? mv_cost(mx << 2, my << 2) : 0xffffffffu, sad, cst);
        vmask = 0;
#pragma unroll
        for (int s = 0; s < 4; s++) { rd[s] = sad[s] + cst[s]; vmask |= cst[s] != 0xffffffffu ? (1u << s) : 0u; }
    };
    // this lane's offsets inside the small diamond and the two halves of the big diamond (c_small / c_big rows slot, 4 + slot)
    const int sdx = (slot & 1) ? 0 : slot - 1, sdy = (slot & 1) ? slot - 2 : 0;
    const int bdx0 = slot - 2, bdy0 = (slot <= 2) ? -slot : -1, bdx1 = 2 - slot, bdy1 = (slot <= 2) ? slot : 1;

    int bx = 0, by = 0;
    uint32_t bsad = 0, brd = 0, n_probes = 0;
    int mvx = 0, mvy = 0, subx = 0, suby = 0;
    uint32_t best_sad = 0xffffffffu / 8;

    if (a.action & HB_ME_PEL) {
        // ---- origin + extra start points (caller's list, then the parent PU's vector when both components are non-zero)
        int sx[5], sy[5]; bool sv[5];
        sx[0] = min(max(0, xlo), xhi); sy[0] = min(max(0, ylo), yhi); sv[0] = true;
        const int n_start = PRE ? 0 : j.n_start;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int x = PRE ? 0 : j.start[2 * i] >> 2, y = PRE ? 0 : j.start[2 * i + 1] >> 2;
            sx[1 + i] = x; sy[1 + i] = y;
            sv[1 + i] = i < n_start && !(x == 0 && y == 0) && inside(x, y);
        }
        sx[4] = 0; sy[4] = 0; sv[4] = false;
        if (j.has_parent) {
            const int x = j.pmvx >> 2, y = j.pmvy >> 2;
            sx[4] = x; sy[4] = y;
            sv[4] = j.pmvx != 0 && j.pmvy != 0 && !(x == 0 && y == 0) && inside(x, y);
        }
        uint32_t ssad[5], srd[5];
        // the parent's vector rides in the fourth slot of the first round when the caller's list leaves it free (always in the
        // pre-pass): one round instead of two.  The compare order below is unchanged, and a probe's result does not depend on it.
        const bool merged = !sv[3];
        {
            const int cx[4] = { sx[0], sx[1], sx[2], merged ? sx[4] : sx[3] }, cy[4] = { sy[0], sy[1], sy[2], merged ? sy[4] : sy[3] };
            const bool cv[4] = { sv[0], sv[1], sv[2], merged ? sv[4] : sv[3] };
            uint32_t sad[4], rd[4];
            round4(cx, cy, cv, sad, rd);
#pragma unroll
            for (int s = 0; s < 4; s++) { ssad[s] = sad[s]; srd[s] = rd[s]; }
        }
        ssad[4] = ssad[3]; srd[4] = srd[3];
        if (any_pu(sv[4] && !merged)) {                // uniform per PU; PUs that need no extra round idle through it
            const int cx[4] = { sx[4], 0, 0, 0 }, cy[4] = { sy[4], 0, 0, 0 };
            const bool cv[4] = { sv[4] && !merged, false, false, false };
            uint32_t sad[4], rd[4];
            round4(cx, cy, cv, sad, rd);
            if (!merged) { ssad[4] = sad[0]; srd[4] = rd[0]; }
        }
        bx = sx[0]; by = sy[0]; bsad = ssad[0]; brd = srd[0]; n_probes = 1;
        bool skip = bsad == 0;
        if (!skip) {
#pragma unroll
            for (int i = 1; i < 5; i++)
                if (sv[i]) { n_probes++; if (srd[i] < brd) { bsad = ssad[i]; brd = srd[i]; bx = sx[i]; by = sy[i]; } }
            skip = bsad == 0;
        }
        int cx0 = bx, cy0 = by;
        if (any_pu(!skip)) {
            {   // first small diamond, fixed order, centre stays put (:1501-1523)
                uint32_t sad[4], rd[4], vmask;
                round_own(cx0 + sdx, cy0 + sdy, !skip && inside(cx0 + sdx, cy0 + sdy), sad, rd, vmask);
#pragma unroll
                for (int s = 0; s < 4; s++)
                    if ((vmask >> s) & 1u) { n_probes++; if (rd[s] < brd) { bsad = sad[s]; brd = rd[s]; bx = cx0 + c_small[s][0]; by = cy0 + c_small[s][1]; } }
            }
            // rotating big diamond (:1528-1599): dist 2, and 4 when the old centre sat on an axis
            const int end = skip ? 0 : (cx0 != 0 && cy0 != 0) ? 4 : 8;
            int next_start = 0, span = 8;
            cx0 = bx; cy0 = by;
            for (int dist = 2; any_pu(dist < end); dist *= 2) {
                const bool act = dist < end;
                uint32_t sad8[8], rd8[8];
                uint32_t vmask = 0;                           // empty for a PU that only idles through this round
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t sad[4], rd[4], vm;
                    const int mx = cx0 + (h ? bdx1 : bdx0) * dist, my = cy0 + (h ? bdy1 : bdy0) * dist;
                    round_own(mx, my, act && inside(mx, my), sad, rd, vm);
                    vmask |= vm << (4 * h);
#pragma unroll
                    for (int s = 0; s < 4; s++) { sad8[4 * h + s] = sad[s]; rd8[4 * h + s] = rd[s]; }
                }
                if (!beats<8>(rd8, vmask, brd)) {             // nothing can win: the walk only counts the probes it would make
                    n_probes += visited<8>(vmask, next_start, span);
                } else {
                    for (int i = next_start; i < next_start + span; i++) {
                        const int idx = i & 7;
                        if (!((vmask >> idx) & 1u)) continue;
                        n_probes++;
                        const uint32_t rd = pick<8>(rd8, idx);
                        if (rd < brd) {
                            bsad = pick<8>(sad8, idx); brd = rd;
                            bx = cx0 + c_big[idx][0] * dist; by = cy0 + c_big[idx][1] * dist;
                            next_start = (idx - 2 + 8) & 7; span = 5;
                        }
                    }
                }
            }
        }
        // iterated small diamond with rotation until the centre stops moving (:1601-1663)
        cx0 = bx; cy0 = by;
        {
            int next_start = 0, span = 4;
            bool done = false;
            for (;;) {
                uint32_t sad[4], rd[4], vmask;
                round_own(cx0 + sdx, cy0 + sdy, !done && inside(cx0 + sdx, cy0 + sdy), sad, rd, vmask);
                if (!beats<4>(rd, vmask, brd)) {              // the usual last iteration: no neighbour wins, only the probe count moves
                    n_probes += visited<4>(vmask, next_start, span);
                } else {
                    for (int i = next_start; i < next_start + span; i++) {
                        const int idx = i & 3;
                        if (!((vmask >> idx) & 1u)) continue;
                        n_probes++;
                        const uint32_t r = pick<4>(rd, idx);
                        if (r < brd) {
                            bsad = pick<4>(sad, idx); brd = r;
                            bx = cx0 + c_small[idx][0]; by = cy0 + c_small[idx][1];
                            next_start = (idx - 1 + 4) & 3; span = 3;
                        }
                    }
                }
                if (cx0 == bx && cy0 == by) done = true;
                cx0 = bx; cy0 = by;
                if (!any_pu(!done)) break;
            }
        }
        best_sad = bsad;
        mvx = bx << 2; mvy = by << 2;
    }

    if constexpr (PL) {
      if (a.action & HB_ME_HALF) {
        // ---- sub-pel refinement against the picture's quarter-pel planes: a candidate at quarter-pel offset (cx, cy) from the integer
        // winner is the block of plane (cx & 3, cy & 3) at (ix + (cx >> 2), iy + (cy >> 2)); its SAD is taken exactly like an integer
        // probe's (same lanes, same words).  Half-pel: the eight neighbours in the reference's order (s_acMvRefineH_HM, :1035), two
        // rounds of four; quarter-pel: the eight neighbours of the half-pel winner (s_acMvRefineQ, :1062); plain SAD, strict '<'.
        const int ix = mvx >> 2, iy = mvy >> 2;
        uint32_t cur_best = bsad;
        if (!(a.action & HB_ME_PEL)) {
            const int cx[4] = { ix, 0, 0, 0 }, cy[4] = { iy, 0, 0, 0 };
            const bool cv[4] = { true, false, false, false };
            uint32_t sad[4], rd[4];
            round4(cx, cy, cv, sad, rd);
            cur_best = sad[0];
        }
        const uint32_t sp_pitch = static_cast<uint32_t>(a.sp.pitch);
        const uint32_t sp_lane = static_cast<uint32_t>(jy + HB_SUBPEL_OFF + prow0) * sp_pitch + static_cast<uint32_t>(jx + HB_SUBPEL_OFF + pcol);
        const uint32_t sp_step = (L / PPR) * sp_pitch;
        auto plane_of = [&](int cx, int cy) -> const uint8_t * {
            return a.sp.base + static_cast<size_t>(((cy & 3) * 4 + (cx & 3)) - 1) * a.sp.plane_bytes;       // never (0, 0): that is the integer position
        };
        auto plane_off = [&](int cx, int cy) -> uint32_t {
            return sp_lane + static_cast<uint32_t>((iy + (cy >> 2)) * static_cast<int>(sp_pitch) + ix + (cx >> 2));
        };
        auto sad_plane = [&](int cx, int cy) -> uint32_t {
            uint32_t pk[4] = { 0, 0, 0, 0 }, acc;
            const int qx = (ix << 2) + cx, qy = (iy << 2) + cy;
            if constexpr (MEMO == 2) {
                uint32_t v;
                if (memo_get(qx, qy, v)) acc = (l == 0) ? v : 0u;
                else acc = sad_rows(plane_of(cx, cy), plane_off(cx, cy), sp_step, pk);
            } else acc = sad_rows(plane_of(cx, cy), plane_off(cx, cy), sp_step, pk);
            if constexpr (MEMO == 1) { memo_put(pk, qx, qy, true); memo_n += 4; }
            return acc;
        };
        int bidx = 0;
        // the two rounds of a stage run through one copy of the probe (instruction footprint, see k_me_ctu)
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            uint32_t sad[4], cst[4];
            exchange(sad_plane(2 * c_half[1 + 4 * h + slot][0], 2 * c_half[1 + 4 * h + slot][1]), 0u, sad, cst);
#pragma unroll
            for (int s = 0; s < 4; s++)
                if (sad[s] < cur_best) { cur_best = sad[s]; bidx = 1 + 4 * h + s; }
        }
        const int hx = c_half[bidx][0], hy = c_half[bidx][1];
        int sbx = hx * 2, sby = hy * 2;
        if (a.action & HB_ME_QUARTER) {
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                uint32_t sad[4], cst[4];
                exchange(sad_plane(hx * 2 + c_quarter[1 + 4 * h + slot][0], hy * 2 + c_quarter[1 + 4 * h + slot][1]), 0u, sad, cst);
#pragma unroll
                for (int s = 0; s < 4; s++)
                    if (sad[s] < cur_best) { cur_best = sad[s]; sbx = hx * 2 + (h ? quarter_off(5 + s, 0) : quarter_off(1 + s, 0)); sby = hy * 2 + (h ? quarter_off(5 + s, 1) : quarter_off(1 + s, 1)); }
            }
        }
        best_sad = cur_best;
        mvx = (ix << 2) + sbx; mvy = (iy << 2) + sby;
        subx = sbx; suby = sby;
        // ---- the luma prediction of the winner is its block of the plane (of the reference picture itself for an integer vector):
        // the samples hmr_motion_compensation_luma (:1779) produces for this vector.  Slot 0's lanes cover the block.
        if (a.pred.org != nullptr && slot == 0 && j.out >= 0) {
            const bool integer = (sbx | sby) == 0;
            const uint8_t *src = integer ? a.ref.base : plane_of(sbx, sby);
            uint32_t off = integer ? ref_lane + static_cast<uint32_t>(iy * static_cast<int>(rpitch) + ix) : plane_off(sbx, sby);
            const uint32_t step = integer ? ref_step : sp_step;
            const uint32_t sh = (off & 3u) * 8u;
            off &= ~3u;
            uint8_t *dst = a.pred.org + (jy + prow0) * a.pred.pitch + jx + pcol;
#pragma unroll
            for (int k = 0; k < PPL; k++) {
                uint32_t w0, w1, w2;
                ld3(src, off, w0, w1, w2);
                *reinterpret_cast<uint2 *>(dst) = make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
                off += step;
                dst += (L / PPR) * a.pred.pitch;
            }
        }
      }
    } else
    if (a.action & HB_ME_HALF) {
        const int ix = mvx >> 2, iy = mvy >> 2;
        uint32_t cur_best = bsad;
        if (!(a.action & HB_ME_PEL)) {
            const int cx[4] = { ix, 0, 0, 0 }, cy[4] = { iy, 0, 0, 0 };
            const bool cv[4] = { true, false, false, false };
            uint32_t sad[4], rd[4];
            round4(cx, cy, cv, sad, rd);
            cur_best = sad[0];
        }
        // ---- stage the patch: rows iy-4.., columns ix-4..
        // G/N lanes share a row, each takes a run of consecutive words: one misalignment shift per row, no index arithmetic
        {
            constexpr int SPLIT = G / N, WPRW = PS / 4, CH = (WPRW + SPLIT - 1) / SPLIT;
            const int prow = gl / SPLIT, part = gl % SPLIT;
            uint32_t off = static_cast<uint32_t>(jy + a.ref.pad + iy - 4 + prow) * rpitch + static_cast<uint32_t>(jx + a.ref.pad + ix - 4 + part * CH * 4);
            const uint32_t sh = (off & 3u) * 8u;
            off &= ~3u;
#pragma unroll
            for (int r0 = 0; r0 < PROWS; r0 += N) {
                if (r0 + prow < PROWS) {
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(a.ref.base + off);
                    uint32_t *d = reinterpret_cast<uint32_t *>(s_patch + (r0 + prow) * PS) + part * CH;
                    uint32_t lo = __ldg(q);
#pragma unroll
                    for (int k = 0; k < CH; k++) {
                        if (SPLIT == 1 || part * CH + k < WPRW) {
                            const uint32_t hi = __ldg(q + k + 1);
                            d[k] = __funnelshift_r(lo, hi, sh);
                            lo = hi;
                        }
                    }
                }
                off += N * rpitch;
            }
        }
        group_barrier<G>(group, gmask);
        // ---- horizontal 14-bit planes, fractions 0..3: T_f[rho][j] = sum_k taps_f[k] * P[rho][j+k] - 8192 (T_0 = 64 P[rho][j+3] - 8192),
        // stored PAIR-INTERLEAVED: word (q, j) of a plane holds T[2q][j] in its low half and T[2q+1][j] in its high half, so the
        // vertical passes read two taps per shared load and multiply them with one dp2a.  One work item = two rows x four columns
        // x four fractions: u8 samples x s8 taps through dp4a, one 16-byte store per fraction.
        for (int w = gl; w < QROWS * (TS / 4); w += G) {
            const int q = w / (TS / 4), j0 = (w % (TS / 4)) * 4;
            int o[2][4][4];
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const uint32_t *pw = reinterpret_cast<const uint32_t *>(s_patch + (2 * q + rr) * PS + j0);
                const uint32_t w0 = pw[0], w1 = pw[1], w2 = pw[2];
                uint32_t win[8];                                       // win[c] = samples j0+c .. j0+c+3
                win[0] = w0; win[4] = w1;
#pragma unroll
                for (int c = 1; c < 4; c++) { win[c] = __funnelshift_r(w0, w1, 8 * c); win[4 + c] = __funnelshift_r(w1, w2, 8 * c); }
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    o[rr][0][c] = hb_dp4a_us(win[c], htap4(0, 0), -8192);
#pragma unroll
                    for (int f = 1; f < 4; f++) o[rr][f][c] = hb_dp4a_us(win[c + 4], htap4(f, 1), hb_dp4a_us(win[c], htap4(f, 0), -8192));
                }
            }
#pragma unroll
            for (int f = 0; f < 4; f++) {
                uint4 v;
                v.x = __byte_perm(o[0][f][0], o[1][f][0], 0x5410); v.y = __byte_perm(o[0][f][1], o[1][f][1], 0x5410);
                v.z = __byte_perm(o[0][f][2], o[1][f][2], 0x5410); v.w = __byte_perm(o[0][f][3], o[1][f][3], 0x5410);
                *reinterpret_cast<uint4 *>(s_plane + f * PLANE_WORDS + q * TS + j0) = v;
            }
        }
        group_barrier<G>(group, gmask);
        // ---- the current block, transposed (column c at s_cur + c*CS: four vertical neighbours per word), over the patch area
        uint8_t *s_cur = s_patch;
#pragma unroll
        for (int k = 0; k < WPL; k++) {
            const int row = prow0 + (k / 2) * (L / PPR), col = pcol + (k & 1) * 4 + slot;     // slot s stores byte s of every word
            s_cur[col * CS + row] = static_cast<uint8_t>(cur[k] >> (8 * slot));
        }
        if (gl < 8) s_half[group * 8 + gl] = 0;
        group_barrier<G>(group, gmask);

        // one round: SADs of four sub-pel candidates at quarter-pel offsets (qx[s], qy[s]) in [-3,3]^2 from (ix,iy).  A lane filters
        // CPL columns of its slot's candidate top to bottom: per two output rows one shared load and ten dp2a; four clipped
        // samples are packed and compared with four current samples in one instruction.
        auto subpel4 = [&](int cx, int cy, uint32_t (&sad)[4]) {                // (cx, cy): this lane's slot's candidate
            const int fx = cx & 3, fy = cy & 3, cb = cx >> 2, par = (cy >> 2) + 1;
            uint32_t tb[5];
#pragma unroll
            for (int i = 0; i < 5; i++) tb[i] = c_vtab[fy][par][i];
            const int c0 = l * CPL;
            const uint32_t *pl = s_plane + fx * PLANE_WORDS + c0 + cb + 1;
            uint32_t acc = 0;
#pragma unroll
            for (int cc = 0; cc < CPL; cc++) {
                uint32_t win[5];
#pragma unroll
                for (int k = 0; k < 4; k++) win[k] = pl[k * TS + cc];
#pragma unroll
                for (int m4 = 0; m4 < N / 4; m4++) {
                    int v[4];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int m = 2 * m4 + h;
                        win[(m + 4) % 5] = pl[(m + 4) * TS + cc];
                        int sa = kVRound, sb = kVRound;
#pragma unroll
                        for (int i = 0; i < 5; i++) { sa = __dp2a_lo(static_cast<int>(win[(m + i) % 5]), static_cast<int>(tb[i]), sa); sb = __dp2a_hi(static_cast<int>(win[(m + i) % 5]), static_cast<int>(tb[i]), sb); }
                        v[2 * h] = sa >> 12; v[2 * h + 1] = sb >> 12;
                    }
                    acc = hb_sad4_acc(hb_pack_sat_u8x4(v[0], v[1], v[2], v[3]), *reinterpret_cast<const uint32_t *>(s_cur + (c0 + cc) * CS + 4 * m4), acc);
                }
            }
            uint32_t cst[4];
            exchange(acc, 0u, sad, cst);
        };

        // ---- half-pel stage.  The reference builds three planes ((0,2), (2,0), (2,2), :395-438) and reads each of them for
        // two or four candidates shifted by one sample; here a lane walks one COLUMN of the horizontal plane T_0 or T_2 and
        // feeds every filtered sample to all the candidates it belongs to:
        //   T_0 column (N strips): v = V2(T_0)                   -> candidates (0,-2) and (0,+2)
        //   T_2 column (N strips): v = V2(T_2), u = round(T_2)   -> (-2,-2) (+2,-2) (-2,+2) (+2,+2) and (-2,0) (+2,0)
        // Partial SADs go to eight shared counters.  T_0 / T_2 is uniform per warp.
        // c_half order: 1 (0,-1) 2 (0,1) 3 (-1,0) 4 (1,0) 5 (-1,-1) 6 (1,-1) 7 (-1,1) 8 (1,1)
        auto t0_strip = [&](int j) {                                  // plane column j: x = ix - 1 + j, block column j - 1
            uint32_t acc[6] = { 0, 0, 0, 0, 0, 0 };                   // {minus,L} {minus,R} {plus,L} {plus,R} {u,L} {u,R}
            half_strip<N, TS, false>(s_plane + j, s_cur + (j - 1) * CS, s_cur, acc);
            atomicAdd(&s_half[group * 8 + 0], acc[0]); atomicAdd(&s_half[group * 8 + 1], acc[2]);
        };
        auto t2_strip = [&](int j) {                                  // x = ix - 1/2 + j: block columns j - 1 (L) and j (R)
            uint32_t acc[6] = { 0, 0, 0, 0, 0, 0 };
            half_strip<N, TS, true>(s_plane + 2 * PLANE_WORDS + j, s_cur + max(j - 1, 0) * CS, s_cur + j * CS, acc);
            if (j >= 1) { atomicAdd(&s_half[group * 8 + 5], acc[0]); atomicAdd(&s_half[group * 8 + 7], acc[2]); atomicAdd(&s_half[group * 8 + 3], acc[4]); }
            atomicAdd(&s_half[group * 8 + 4], acc[1]); atomicAdd(&s_half[group * 8 + 6], acc[3]); atomicAdd(&s_half[group * 8 + 2], acc[5]);
        };
        if constexpr (G == N) {                                       // small PUs: every lane takes one column of each plane
            t0_strip(gl + 1);
            t2_strip(gl);
        } else {                                                      // large PUs: whole warps take T_0 or T_2 columns (G >= 2N)
            if (gl < N) t0_strip(gl + 1);
            else if (gl < 2 * N) t2_strip(gl - N);
        }
        // the last T_2 column (j = N, right edge of the x = +2 candidates) would cost a whole extra pass of the strip loop for one
        // lane: its N+1 filtered samples are spread over the lanes instead, one direct 8-tap sum each
        for (int rho0 = 0; rho0 <= N; rho0 += G) {
            const int rho = rho0 + gl;
            uint32_t am = 0, ap = 0, au = 0;
            if (rho <= N) {
                auto t2_at = [&](int r) -> int { return reinterpret_cast<const int16_t *>(s_plane + 2 * PLANE_WORDS + (r >> 1) * TS + N)[r & 1]; };
                const int s = hb_luma8<2>(t2_at(rho), t2_at(rho + 1), t2_at(rho + 2), t2_at(rho + 3), t2_at(rho + 4), t2_at(rho + 5), t2_at(rho + 6), t2_at(rho + 7));
                const int v = __vimin_s32_relu((s + kVRound) >> 12, 255);
                if (rho < N) {
                    const int cl = s_cur[(N - 1) * CS + rho];
                    am = __sad(v, cl, 0u);
                    au = __sad(__vimin_s32_relu((t2_at(rho + 4) + 8192 + 32) >> 6, 255), cl, 0u);
                }
                if (rho >= 1) ap = __sad(v, static_cast<int>(s_cur[(N - 1) * CS + rho - 1]), 0u);
            }
            if constexpr (G >= 32) {
                am = __reduce_add_sync(HB_FULL_MASK, am); ap = __reduce_add_sync(HB_FULL_MASK, ap); au = __reduce_add_sync(HB_FULL_MASK, au);
                if (lane == 0 && (am | ap | au)) { atomicAdd(&s_half[group * 8 + 5], am); atomicAdd(&s_half[group * 8 + 7], ap); atomicAdd(&s_half[group * 8 + 3], au); }
            } else if (rho <= N) {
                atomicAdd(&s_half[group * 8 + 5], am); atomicAdd(&s_half[group * 8 + 7], ap); atomicAdd(&s_half[group * 8 + 3], au);
            }
        }
        group_barrier<G>(group, gmask);
        int bidx = 0;
        {
            const uint4 h0 = *reinterpret_cast<const uint4 *>(&s_half[group * 8 + 0]), h1 = *reinterpret_cast<const uint4 *>(&s_half[group * 8 + 4]);
            const uint32_t hv[8] = { h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w };
#pragma unroll
            for (int i = 1; i < 9; i++)                    // candidate 0 is the integer position itself: never smaller
                if (hv[i - 1] < cur_best) { cur_best = hv[i - 1]; bidx = i; }
        }
        const int hx = c_half[bidx][0], hy = c_half[bidx][1];
        int sbx = hx * 2, sby = hy * 2;
        if (a.action & HB_ME_QUARTER) {
#pragma unroll
            for (int h = 0; h < 2; h++) {              // candidate 0 repeats the half-pel winner
                uint32_t sad[4];
                subpel4(hx * 2 + c_quarter[1 + 4 * h + slot][0], hy * 2 + c_quarter[1 + 4 * h + slot][1], sad);
#pragma unroll
                for (int s = 0; s < 4; s++)
                    if (sad[s] < cur_best) { cur_best = sad[s]; sbx = hx * 2 + quarter_off(1 + 4 * h + s, 0); sby = hy * 2 + quarter_off(1 + 4 * h + s, 1); }
            }
        }
        best_sad = cur_best;
        mvx = (ix << 2) + sbx; mvy = (iy << 2) + sby;
        subx = sbx; suby = sby;

        // ---- optionally leave the luma prediction of the winner in the prediction plane: same two-stage samples that
        // hmr_motion_compensation_luma (:1779) produces for this vector, taken from the planes already in shared memory
        if (a.pred.org != nullptr) {
            constexpr int SEGS = G / N, RPS = N / SEGS;            // row segments per column, rows per segment (even)
            const int fx = sbx & 3, fy = sby & 3, cb = sbx >> 2, par = (sby >> 2) + 1;
            uint32_t tb[5];
#pragma unroll
            for (int i = 0; i < 5; i++) tb[i] = c_vtab[fy][par][i];
            const int c = gl % N, r0 = (gl / N) * RPS;
            const uint32_t *pl = s_plane + fx * PLANE_WORDS + (r0 / 2) * TS + c + cb + 1;
            uint8_t *dst = a.pred.org + (jy + r0) * a.pred.pitch + jx + c;
            uint32_t win[5];
#pragma unroll
            for (int k = 0; k < 4; k++) win[k] = pl[k * TS];
#pragma unroll
            for (int m = 0; m < RPS / 2; m++) {
                win[(m + 4) % 5] = pl[(m + 4) * TS];
                int sa = kVRound, sb = kVRound;
#pragma unroll
                for (int i = 0; i < 5; i++) { sa = __dp2a_lo(static_cast<int>(win[(m + i) % 5]), static_cast<int>(tb[i]), sa); sb = __dp2a_hi(static_cast<int>(win[(m + i) % 5]), static_cast<int>(tb[i]), sb); }
                dst[(2 * m) * a.pred.pitch] = static_cast<uint8_t>(__vimin_s32_relu(sa >> 12, 255));
                dst[(2 * m + 1) * a.pred.pitch] = static_cast<uint8_t>(__vimin_s32_relu(sb >> 12, 255));
            }
        }
    }

    if (gl == 0 && j.out >= 0) {
        hb_me_result r;
        r.mv.x = mvx; r.mv.y = mvy; r.subpix.x = subx; r.subpix.y = suby; r.sad = best_sad; r.n_probes = n_probes;
        a.out[j.out] = r;
    }
    win_mvx = mvx; win_mvy = mvy;
}

// ---- one launch per PU size, PUs from a job list (hb_me_search; the pre-pass with per-PU planes or the staged window)
template <int N, bool PL, bool WIN>
__global__ void __launch_bounds__(256, WIN ? 3 : PL ? 4 : ((N <= 16) ? 3 : (N == 64) ? 4 : 3)) k_me(const MeArgs a)
{
    using Cfg = MeCfg<N>;
    constexpr int G = Cfg::G, PUS = Cfg::PUS, NSEG = Cfg::NSEG;
    extern __shared__ __align__(16) uint8_t s_raw[];
    __shared__ uint32_t s_x[((G > 32) ? PUS : 1) * 2 * NSEG * 2];
    __shared__ __align__(16) uint32_t s_half[PUS * 8];                      // SADs of the eight half-pel candidates (c_half order 1..8)

    const int group = threadIdx.x / G, gl = threadIdx.x % G;
    const int job_idx = blockIdx.x * PUS + group;
    // ---- WIN: stage the strip's search window.  Row r of the window = picture row wy0 + r, columns wx0 .. (both clipped to the
    // picture: a probed block never leaves it, :1424-1427); every thread issues the copies of its rows, thread 0 arms the barrier.
    int wx0 = 0, wy0 = 0;
    if constexpr (WIN) {
        __shared__ __align__(8) uint64_t s_bar;
        const hbd_me_job *j0 = a.jobs + blockIdx.x * PUS;              // the first entry of a strip is always a real PU
        const int xa = j0->x, ya = j0->y;
        wx0 = max(xa - 128, 0); wy0 = max(ya - 64, 0);
        const int x_hi = min(xa + PUS * N + 128, a.cur.w), y_hi = min(ya + N + 64, a.cur.h);
        const int wbytes = (x_hi - wx0 + 3 + 15) & ~15, rows = y_hi - wy0;     // + 3: a probe's last word may start up to 3 bytes past its block
        const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&s_bar));
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rows * wbytes) : "memory");
        for (int r = threadIdx.x; r < rows; r += 256) {
            const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(s_raw + r * Cfg::WIN_P));
            const uint8_t *src = a.ref.org + (wy0 + r) * a.ref.pitch + wx0;          // 16-byte aligned: pad, pitch and wx0 are multiples of 16
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"(wbytes), "r"(bar) : "memory");
        }
        asm volatile("{\n .reg .pred p;\n WAITW_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra DONEW_%=;\n bra WAITW_%=;\n DONEW_%=:\n}\n" ::"r"(bar) : "memory");
    }
    if (job_idx >= a.n_jobs) return;                   // the whole group leaves together
    const hbd_me_job *jp = a.jobs + job_idx;
    if (WIN && jp->x < 0) return;                      // padding entry of a strip
    MePu j;
    j.x = jp->x; j.y = jp->y; j.n_amvp = jp->n_amvp;
    j.a0x = jp->amvp[0]; j.a0y = jp->amvp[1]; j.a1x = jp->amvp[2]; j.a1y = jp->amvp[3];
    j.n_start = jp->n_start;
#pragma unroll
    for (int i = 0; i < 6; i++) j.start[i] = jp->start[i];
    j.has_parent = jp->parent >= 0; j.pmvx = 0; j.pmvy = 0;
    if (j.has_parent) { const hb_mv pmv = a.parent[jp->parent].mv; j.pmvx = pmv.x; j.pmvy = pmv.y; }
    j.out = jp->out; j.corr = jp->corr;
    int mvx, mvy;
    me_pu<N, PL, WIN, false>(a, j, group, gl, s_raw, s_x, s_half, wx0, wy0, mvx, mvy);
}

// ---- the pre-pass search as ONE launch: a CTA owns a CTU and searches its 64x64 PU, then its four 32x32, sixteen 16x16 and
// sixty-four 8x8 PUs (two rounds of 32) -- the thread-group shapes of the per-size kernels tile a CTU exactly.  A child's extra start
// point (the parent's winning vector, hmr_motion_inter.c:2613) travels through shared memory; the reference picture's quarter-pel
// planes serve every depth.  PUs outside the picture do not exist; where they share a warp with existing ones (lock-step groups)
// they re-run a neighbour's search without writing anything.
struct MeCtuArgs {
    hbd_plane cur, ref;
    hbd_subpel sp;
    hb_me_result *out[4];
    hbd_plane pred[4];
    const hbd_dyn_params *dyn;
    int action, ctu_cols, ctu_row0;
    int grid_w[4];
};

template <int N, int D>
__device__ __forceinline__ void me_ctu_depth(const MeCtuArgs &c, const int X0, const int Y0, const int round, uint32_t *s_x, int2 *s_mv, MeMemo *memo)
{
    using Cfg = MeCfg<N>;
    constexpr int G = Cfg::G, PUS = Cfg::PUS, SIDE = 64 / N;                   // SIDE x SIDE PUs per CTU
    constexpr int BASE = (D == 0) ? 0 : (D == 1) ? 1 : (D == 2) ? 5 : 21;       // slot of this depth's first PU in s_mv (breadth first)
    constexpr int PBASE = (D <= 1) ? 0 : (D == 2) ? 1 : 5;
    const int group = threadIdx.x / G, gl = threadIdx.x % G;
    int pu = round * PUS + group;                                               // raster index inside the CTU
    const int fw = c.cur.w, fh = c.cur.h;
    auto exists = [&](int p) { return X0 + (p % SIDE) * N + N <= fw && Y0 + (p / SIDE) * N + N <= fh; };
    bool mine = exists(pu);
    if constexpr (G < 32) {
        // lock-step groups: a warp runs as long as one of its PUs exists, the others repeat the first existing one
        const uint32_t vm = __ballot_sync(HB_FULL_MASK, mine);
        if (vm == 0) return;
        if (!mine) pu = round * PUS + (threadIdx.x & ~31) / G + (__ffs(vm) - 1) / G;
    } else {
        if (!mine) return;                                                      // whole warps (a named-barrier group or the CTA) leave together
    }
    const int lx = pu % SIDE, ly = pu / SIDE;
    MePu j;
    j.x = X0 + lx * N; j.y = Y0 + ly * N;
    j.n_amvp = 2; j.a0x = j.a0y = j.a1x = j.a1y = 0; j.n_start = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) j.start[i] = 0;
    j.has_parent = false; j.pmvx = 0; j.pmvy = 0;
    if constexpr (D > 0) {
        const int plx = lx / 2, ply = ly / 2;
        j.has_parent = X0 + plx * 2 * N + 2 * N <= fw && Y0 + ply * 2 * N + 2 * N <= fh;
        if (j.has_parent) { const int2 m = s_mv[PBASE + ply * (SIDE / 2) + plx]; j.pmvx = m.x; j.pmvy = m.y; }
    }
    j.out = mine ? ((Y0 / N) + ly) * c.grid_w[D] + (X0 / N) + lx : -1;
    j.corr = 0.;
    MeArgs a;
    a.cur = c.cur; a.ref = c.ref; a.jobs = nullptr; a.n_jobs = 0; a.parent = nullptr; a.out = c.out[D]; a.action = c.action; a.dyn = c.dyn;
    a.pred = c.pred[D]; a.sp = c.sp;
    int mvx, mvy;
    me_pu<N, true, false, true, (D == 0) ? 1 : 2>(a, j, group, gl, nullptr, s_x, nullptr, 0, 0, mvx, mvy, memo, pu);
    if (D < 3 && gl == 0 && mine) s_mv[BASE + pu] = make_int2(mvx, mvy);
}

__global__ void __launch_bounds__(256, 4) k_me_ctu(const MeCtuArgs c)
{
    constexpr int SX64 = 2 * MeCfg<64>::NSEG * 2, SX32 = MeCfg<32>::PUS * 2 * MeCfg<32>::NSEG * 2;
    __shared__ uint32_t s_x[SX64 > SX32 ? SX64 : SX32];
    __shared__ int2 s_mv[1 + 4 + 16];
    __shared__ MeMemo s_memo;
    const int X0 = (blockIdx.x % c.ctu_cols) * 64, Y0 = (c.ctu_row0 + blockIdx.x / c.ctu_cols) * 64;
    for (int i = threadIdx.x; i < MeMemo::ENTRIES * 32; i += 256) (&s_memo.sum[0][0])[i] = 0;
    if (threadIdx.x < MeMemo::SLOTS / 4) s_memo.idx[threadIdx.x] = 0xffffffffu;
    __syncthreads();
    me_ctu_depth<64, 0>(c, X0, Y0, 0, s_x, s_mv, &s_memo);
    __syncthreads();
    me_ctu_depth<32, 1>(c, X0, Y0, 0, s_x, s_mv, &s_memo);
    __syncthreads();
    me_ctu_depth<16, 2>(c, X0, Y0, 0, s_x, s_mv, &s_memo);
    __syncthreads();
    // the sixty-four 8x8 PUs in two rounds through ONE copy of the code (the kernel's instructions are its largest cache footprint)
#pragma unroll 1
    for (int round = 0; round < 2; round++) me_ctu_depth<8, 3>(c, X0, Y0, round, s_x, s_mv, &s_memo);
}

template <int N> int configure_me()
{
    int e = static_cast<int>(cudaFuncSetAttribute(k_me<N, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MeCfg<N>::SMEM_TOTAL));
    if (!e) e = static_cast<int>(cudaFuncSetAttribute(k_me<N, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MeCfg<N>::WIN_BYTES));
    return e;
}

template <int N> int launch_me(const MeArgs &a, bool window, cudaStream_t s)
{
    const int grid = (a.n_jobs + MeCfg<N>::PUS - 1) / MeCfg<N>::PUS;
    if (a.sp.base && window) k_me<N, true, true><<<grid, 256, MeCfg<N>::WIN_BYTES, s>>>(a);
    else if (a.sp.base) k_me<N, true, false><<<grid, 256, 0, s>>>(a);
    else k_me<N, false, false><<<grid, 256, MeCfg<N>::SMEM_TOTAL, s>>>(a);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace

// opt in to the dynamic shared memory the search kernels need; once per device, outside any stream capture
// PUs of one window strip (= PUs per CTA) for a PU size: the strip-ordered job lists of the pre-pass are built around it
extern "C" int hbk_me_strip_pus(int size) { return size == 64 ? MeCfg<64>::PUS : size == 32 ? MeCfg<32>::PUS : size == 16 ? MeCfg<16>::PUS : MeCfg<8>::PUS; }

extern "C" int hbk_me_configure(void)
{
    int e = configure_me<64>();
    if (!e) e = configure_me<32>();
    if (!e) e = configure_me<16>();
    if (!e) e = configure_me<8>();
    return e;
}

extern "C" int hbk_me_search(const hbd_frame *cur, const hbd_frame *ref, int size, const hbd_me_job *jobs, int n_jobs,
                             const hb_me_result *parent, hb_me_result *out, int action, const hbd_dyn_params *dyn, const hbd_frame *pred_out,
                             const hbd_subpel *sp, int window, void *stream)
{
    if (n_jobs <= 0) return 0;
    MeArgs a;
    memset(&a.pred, 0, sizeof a.pred);
    memset(&a.sp, 0, sizeof a.sp);
    if (sp && (action & HB_ME_HALF)) a.sp = *sp;
    if (pred_out && (action & HB_ME_HALF)) a.pred = pred_out->p[0];
    a.cur = cur->p[0]; a.ref = ref->p[0]; a.jobs = jobs; a.n_jobs = n_jobs; a.parent = parent; a.out = out; a.action = action; a.dyn = dyn;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool win = window != 0 && a.sp.base != nullptr;        // strip-ordered job list (hbk_me_strip_pus) + the picture's sub-pel planes
    switch (size) {
    case 64: return launch_me<64>(a, win, s);
    case 32: return launch_me<32>(a, win, s);
    case 16: return launch_me<16>(a, win, s);
    case 8: return launch_me<8>(a, win, s);
    default: return static_cast<int>(cudaErrorInvalidValue);
    }
}

// the whole search of a picture (or of a band of CTU rows) in one launch: every CTU's PUs of all four sizes, zero predictors, the
// parent's vector as extra start point.  out[d] / pred_out[d]: result table (PU raster of the picture) and prediction picture of depth d.
extern "C" int hbk_me_search_ctus(const hbd_frame *cur, const hbd_frame *ref, const hbd_subpel *sp, hb_me_result *const out[4], const hbd_frame *const pred_out[4],
                                  int action, const hbd_dyn_params *dyn, int ctu_cols, int ctu_row0, int ctu_rows, const int grid_w[4], void *stream)
{
    if (!sp || !sp->base || !(action & HB_ME_HALF) || !dyn) return static_cast<int>(cudaErrorInvalidValue);
    if (ctu_cols <= 0 || ctu_rows <= 0) return 0;
    MeCtuArgs c;
    memset(&c, 0, sizeof c);
    c.cur = cur->p[0]; c.ref = ref->p[0]; c.sp = *sp; c.dyn = dyn; c.action = action; c.ctu_cols = ctu_cols; c.ctu_row0 = ctu_row0;
    for (int d = 0; d < 4; d++) { c.out[d] = out[d]; c.grid_w[d] = grid_w[d]; if (pred_out && pred_out[d]) c.pred[d] = pred_out[d]->p[0]; }
    k_me_ctu<<<ctu_cols * ctu_rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(c);
    return static_cast<int>(cudaGetLastError());
}
