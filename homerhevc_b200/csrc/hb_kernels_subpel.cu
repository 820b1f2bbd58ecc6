// hb_kernels_subpel.cu -- the fifteen quarter-pel planes of a reference picture's luma, built ONCE per picture.
//
// The reference rebuilds its sub-pel planes per PU around that PU's integer winner (hmr_half_pixel_estimation_luma_hm
// hmr_motion_inter.c:395, hmr_quarter_pixel_estimation_luma_hm :442: 13 interpolation passes per PU and depth).  Every sample it
// compares, though, is a function of the reference picture and of the absolute quarter-pel position alone:
//     T_fx[y][x]      = sum_k h_fx[k] * ref[y][x - 3 + k] - 8192                                  (first pass, 14 bit; h_0 = 64 at k = 3)
//     Q_fy,fx[y][x]   = clip255((sum_k v_fy[k] * T_fx[y - 3 + k][x] + 2048 + (8192 << 6)) >> 12)  (second pass)
// -- the two-stage arithmetic of its interpolation functions (:312), which also is what hmr_motion_compensation_luma (:1779)
// produces for a vector with fractions (fx, fy) (one-pass cases included: sum v = 64 makes them the same expression).  So the
// planes Q_fy,fx, (fx, fy) != (0, 0), are computed here for the whole picture and the search kernels only take SADs against them
// (and copy the winner's block out as the luma prediction): 120 MAC per sample and picture instead of ~416.
//
// One CTA = a 64 x 32 tile of all fifteen planes.  The (64+16) x 40 patch of the padded reference it needs arrives in shared memory
// as ONE TMA tile copy (cp.async.bulk.tensor.2d, completion on an mbarrier); the four horizontal planes go to shared memory
// pair-interleaved (one word = rows 2q, 2q+1 of a column) through dp4a, and a thread then walks one column of one horizontal plane
// down the tile, feeding every window of five words to the three vertical filters through dp2a: two output rows of three planes per
// fifteen multiply instructions, the fourth (fy = 0) is a shift.
#include <cuda.h>
#include <cstdlib>
#include "hb_shim.h"
#include "hb_dev_common.cuh"

namespace {

constexpr int TW = 64, TH = 32;                    // output tile
// patch: the 71 x 39 samples a tile depends on (columns x0-3 .. x0+67, rows y0-3 .. y0+35) sit at column PATCH_X of a 96 x 40 box whose
// first column is x0 - 12: that is a multiple of 16 samples from the start of the padded plane (64 bx - 16 + pad, pad = 96), so the
// box starts on a 16-byte boundary of global memory
constexpr int BOX_W = 96, BOX_H = 40, PATCH_X = 9;
constexpr int QR = BOX_H / 2;                      // pair rows of a horizontal plane
constexpr int kVRound = 2048 + (8192 << 6);

__host__ __device__ constexpr int tap0(int f, int k) { return (k < 0 || k > 7) ? 0 : luma_tap(f, k); }
// word i of the five-word window holds rows 2m+2i, 2m+2i+1: output row 2m takes taps (2i, 2i+1), output row 2m+1 taps (2i-1, 2i)
__host__ __device__ constexpr uint32_t vpair(int f, int i) { return pack_s8(tap0(f, 2 * i), tap0(f, 2 * i + 1), tap0(f, 2 * i - 1), tap0(f, 2 * i)); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
                 ::"r"(bar), "r"(parity) : "memory");
}

template <int FY> __device__ __forceinline__ void vfilter2(const uint32_t (&win)[5], int m, int &a, int &b)
{
    int sa = kVRound, sb = kVRound;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const int w = static_cast<int>(win[(m + i) % 5]);
        if (i < 4) sa = __dp2a_lo(w, static_cast<int>(vpair(FY, i)), sa);       // row 2m: taps 0..7 sit in words 0..3
        sb = __dp2a_hi(w, static_cast<int>(vpair(FY, i)), sb);
    }
    a = sa >> 12; b = sb >> 12;                        // the caller's pack clips to 0..255
}

template <bool TMA>
__global__ void __launch_bounds__(256) k_subpel_planes(const __grid_constant__ CUtensorMap tmap, const hbd_plane ref, const hbd_subpel sp, const int tile_row0)
{
    const int tile_y = blockIdx.y + tile_row0;                       // a band of the picture: tile rows tile_row0 .. (the grid's height)
    __shared__ __align__(128) uint8_t s_patch[BOX_H * BOX_W];
    __shared__ __align__(16) uint32_t s_t[4][QR][TW];
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW - HB_SUBPEL_OFF, y0 = tile_y * TH - HB_SUBPEL_OFF;      // picture position of the tile's first sample

    if (TMA) {
        const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&s_bar));
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(s_patch));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BOX_W * BOX_H) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(x0 - 3 - PATCH_X + ref.pad), "r"(y0 - 3 + ref.pad), "r"(bar) : "memory");
        }
        mbar_wait(bar, 0);
    } else {
        for (int wi = tid; wi < BOX_H * (BOX_W / 4); wi += 256) {
            const int r = wi / (BOX_W / 4), c4 = (wi % (BOX_W / 4)) * 4;
            *reinterpret_cast<uint32_t *>(s_patch + r * BOX_W + c4) = __ldg(reinterpret_cast<const uint32_t *>(ref.org + (y0 - 3 + r) * ref.pitch + x0 - 3 - PATCH_X + c4));
        }
        __syncthreads();
    }

    // ---- horizontal planes T_0..T_3, pair-interleaved: one item = two rows x four columns x four fractions
    for (int wi = tid; wi < QR * (TW / 4); wi += 256) {
        const int q = wi / (TW / 4), j0 = (wi % (TW / 4)) * 4;
        int o[2][4][4];
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            // patch column j0 sits one byte into the aligned word at box column PATCH_X - 1 + j0
            const uint32_t *pw = reinterpret_cast<const uint32_t *>(s_patch + (2 * q + rr) * BOX_W + PATCH_X - 1 + j0);
            const uint32_t w0 = pw[0], w1 = pw[1], w2 = pw[2];
            uint32_t win[8];                                       // win[c] = patch samples j0+c .. j0+c+3
            win[3] = w1; win[7] = w2;
#pragma unroll
            for (int c = 0; c < 3; c++) { win[c] = __funnelshift_r(w0, w1, 8 * (c + 1)); win[4 + c] = __funnelshift_r(w1, w2, 8 * (c + 1)); }
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
            *reinterpret_cast<uint4 *>(&s_t[f][q][j0]) = v;
        }
    }
    __syncthreads();

    // ---- vertical pass: thread = (horizontal fraction fx, column c).  A lane ends up with the four vertical fractions of its column (one
    // saturating pack per row: the clip of the second pass); the four lanes of a column quad then swap bytes (4x4 transpose in two shuffle
    // + byte-permute steps) so that lane j holds four neighbouring samples of the plane with vertical fraction j: one 32-bit store per lane
    // and row instead of four byte stores
    const int fx = tid >> 6, c = tid & 63, lane = tid & 31, j = lane & 3;
    const int px = blockIdx.x * TW + (c & ~3);                     // plane column of the quad's first sample
    const bool col_ok = px < sp.w;                                 // whole warps beyond the plane still take part in the shuffles
    const uint32_t *col = &s_t[fx][0][c];
    const int rows = min(TH, sp.h - tile_y * TH);
    const bool plane_ok = col_ok && !(j == 0 && fx == 0);          // (0, 0) is the reference picture itself
    uint8_t *p0 = sp.base + static_cast<size_t>(max(j * 4 + fx - 1, 0)) * sp.plane_bytes + static_cast<size_t>(tile_y * TH) * sp.pitch + px;
    const uint32_t sel0 = (lane & 1) ? 0x3715u : 0x6240u, sel1 = (lane & 2) ? 0x3276u : 0x5410u;
    auto transpose4 = [&](uint32_t w) -> uint32_t {
        const uint32_t a = __byte_perm(w, __shfl_xor_sync(HB_FULL_MASK, w, 1), sel0);
        return __byte_perm(a, __shfl_xor_sync(HB_FULL_MASK, a, 2), sel1);
    };
    uint32_t win[5];
#pragma unroll
    for (int k = 0; k < 4; k++) win[k] = col[k * TW];
#pragma unroll
    for (int m = 0; m < TH / 2; m++) {
        win[(m + 4) % 5] = col[(m + 4) * TW];
        int v[4][2];
        // fy = 0: the first-pass sample itself, rounded: rows 2m+3 (high half of word m+1) and 2m+4 (low half of word m+2)
        v[0][0] = __dp2a_lo(static_cast<int>(win[(m + 1) % 5]), 0x0100, 8192 + 32) >> 6;
        v[0][1] = __dp2a_lo(static_cast<int>(win[(m + 2) % 5]), 0x0001, 8192 + 32) >> 6;
        vfilter2<1>(win, m, v[1][0], v[1][1]);
        vfilter2<2>(win, m, v[2][0], v[2][1]);
        vfilter2<3>(win, m, v[3][0], v[3][1]);
        const uint32_t w0 = transpose4(hb_pack_sat_u8x4(v[0][0], v[1][0], v[2][0], v[3][0]));
        const uint32_t w1 = transpose4(hb_pack_sat_u8x4(v[0][1], v[1][1], v[2][1], v[3][1]));
        if (plane_ok && 2 * m < rows) *reinterpret_cast<uint32_t *>(p0) = w0;
        if (plane_ok && 2 * m + 1 < rows) *reinterpret_cast<uint32_t *>(p0 + sp.pitch) = w1;
        p0 += 2 * static_cast<size_t>(sp.pitch);
    }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn tensor_map_encoder()
{
    static encode_tiled_fn fn = nullptr;
    static int tried = 0;
    if (!tried) {
        tried = 1;
        const char *e = getenv("HB_NO_TMA");
        if (!(e && *e == '1')) {
            void *p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
                fn = reinterpret_cast<encode_tiled_fn>(p);
        }
    }
    return fn;
}

}  // namespace

// 1 when the plane kernel stages its patch with TMA (the default), 0 with plain loads ($HB_NO_TMA=1 or no driver entry point)
extern "C" int hbk_subpel_uses_tma(void) { return tensor_map_encoder() != nullptr; }

// y_lo, y_hi: picture rows [y_lo, y_hi) whose plane samples are wanted (a CTU-row band + the rows its searches can reach); the tile rows
// that hold them are built, the rest of the planes is left as it is
extern "C" int hbk_subpel_planes(const hbd_frame *ref, const hbd_subpel *sp, int y_lo, int y_hi, void *stream)
{
    const hbd_plane &p = ref->p[0];
    const int tiles_y = (sp->h + TH - 1) / TH;
    int t0 = (y_lo + HB_SUBPEL_OFF) / TH, t1 = (y_hi + HB_SUBPEL_OFF + TH - 1) / TH;
    if (y_lo <= 0) t0 = 0;
    t0 = t0 < 0 ? 0 : t0; t1 = t1 > tiles_y ? tiles_y : t1;
    if (y_hi >= p.h) t1 = tiles_y;
    if (t1 <= t0) return 0;
    const dim3 grid((sp->w + TW - 1) / TW, t1 - t0);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    encode_tiled_fn enc = tensor_map_encoder();
    if (enc) {
        // the padded luma plane as a 2-D tensor of bytes: {pitch, rows}; tiles may hang over its right / bottom end (zero filled,
        // never inside the samples a tile's outputs depend on)
        const cuuint64_t dims[2] = { static_cast<cuuint64_t>(p.pitch), static_cast<cuuint64_t>(p.h + 2 * p.pad) };
        const cuuint64_t strides[1] = { static_cast<cuuint64_t>(p.pitch) };
        const cuuint32_t box[2] = { BOX_W, BOX_H }, estr[2] = { 1, 1 };
        const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, p.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return static_cast<int>(cudaErrorInvalidValue);
        k_subpel_planes<true><<<grid, 256, 0, s>>>(tmap, p, *sp, t0);
    } else {
        k_subpel_planes<false><<<grid, 256, 0, s>>>(tmap, p, *sp, t0);
    }
    return static_cast<int>(cudaGetLastError());
}
