// tq_mma_probe.cu -- the tensor-core experiment the north star asks for, on the forward 32x32 transform of the inter T/Q chain
// (k_tq<32>: two 32-point stages per transform unit, hmr_transform.c:54-128 matrices, shifts 4 and 11 for 8-bit video).
//
// Variant A ("imad"): what the product's k_tq<32> does -- a lane owns a row, the HEVC matrix is folded into IMAD immediates through the
//   even/odd recursion (hb_tq_core.cuh), stages exchange rows <-> columns through shared memory.
// Variant B ("mma"):  int8 tensor-core products with SPLIT operands, mma.sync.aligned.m16n8k32 (s32 accumulators):
//   stage 1   Y1^T = T * C^T + (-T) * P^T     the residual never exists: the 8-bit source and prediction blocks are the B operands as they
//                                             lie in memory (four consecutive bytes of a row = one fragment register), T and -T are s8;
//   stage 2   Z = T * Y1, Y1 = 256 * hi + lo  the 16-bit intermediate is split into an unsigned low and a signed high byte plane in
//                                             shared memory, laid out so that stage 2's B fragments are single 32-bit loads.
// Both variants write the same int16 coefficients (checked bit for bit); the program prints their device times.  Instruction counts and
// pipe utilisation come from running it under ncu (profiles/ncu_tq_mma_r02.txt).
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I homerhevc_b200/csrc -o /tmp/tq_mma_probe tools/tq_mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "hb_tq_core.cuh"

static int h_mag(int m)
{
    static const int t[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0 };
    return t[m];
}
static int h_coef(int n, int k, int j)
{
    int m = (k * (32 / n) * (2 * j + 1)) & 127;
    if (m > 64) m = 128 - m;
    return m <= 32 ? h_mag(m) : -h_mag(64 - m);
}

__constant__ int8_t c_T[32][32];

constexpr int S1 = 4, S2 = 11;

// ---------------------------------------------------------------- variant A: IMAD, one warp per unit, lane = row then column
__global__ void __launch_bounds__(256) k_fwd_imad(const uint8_t *cur, const uint8_t *pred, int pitch, const int2 *xy, int n_units, int16_t *out)
{
    __shared__ int16_t s[8][32 * 34];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, u = blockIdx.x * 8 + warp;
    if (u >= n_units) return;
    const int2 p = xy[u];
    const uint4 *c4 = reinterpret_cast<const uint4 *>(cur + (p.y + lane) * pitch + p.x), *p4 = reinterpret_cast<const uint4 *>(pred + (p.y + lane) * pitch + p.x);
    int x[32], y[32];
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const uint4 a = c4[q], b = p4[q];
        const uint32_t aw[4] = { a.x, a.y, a.z, a.w }, bw[4] = { b.x, b.y, b.z, b.w };
#pragma unroll
        for (int w = 0; w < 4; w++)
#pragma unroll
            for (int k = 0; k < 4; k++) x[16 * q + 4 * w + k] = static_cast<int>((aw[w] >> (8 * k)) & 255u) - static_cast<int>((bw[w] >> (8 * k)) & 255u);
    }
    hb_fwd1d<32>(x, y);
    int16_t *t = s[warp];
#pragma unroll
    for (int k = 0; k < 32; k++) t[k * 34 + lane] = static_cast<int16_t>((y[k] + (1 << (S1 - 1))) >> S1);      // [k1][row]
    __syncwarp();
    // second stage: lane = first-stage coefficient k1, transform down the rows
#pragma unroll
    for (int r = 0; r < 32; r++) x[r] = t[lane * 34 + r];
    hb_fwd1d<32>(x, y);
    int16_t *o = out + static_cast<size_t>(u) * 1024;
#pragma unroll
    for (int k = 0; k < 32; k++) o[k * 32 + lane] = static_cast<int16_t>((y[k] + (1 << (S2 - 1))) >> S2);       // [k2][k1]
}

// ---------------------------------------------------------------- variant B: mma.sync m16n8k32, one warp per unit
__device__ __forceinline__ void mma_s8u8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_s8s8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) k_fwd_mma(const uint8_t *cur, const uint8_t *pred, int pitch, const int2 *xy, int n_units, int16_t *out)
{
    // per warp: the stage-1 result as two byte planes [k1][row] (row stride 36 bytes: 32-bit aligned rows, conflict-free fragment loads)
    __shared__ __align__(16) uint8_t s_lo[8][32 * 36], s_hi[8][32 * 36];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, u = blockIdx.x * 8 + warp;
    if (u >= n_units) return;
    // A fragments of T (and of -T), both 16-row tiles: a0 (row g, k 4t..), a1 (row g+8, k 4t..), a2 (row g, k 16+4t..), a3 (row g+8, k 16+4t..)
    uint32_t aT[2][4], aN[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int row = 16 * mt + g + 8 * (i & 1), col = 4 * t + 16 * (i >> 1);
            const uint32_t w = *reinterpret_cast<const uint32_t *>(&c_T[row][col]);
            aT[mt][i] = w;
            aN[mt][i] = __vneg4(w);                          // per-byte negation: -T still fits s8
        }
    const int2 p = xy[u];
    // ---- stage 1: D1[k1][row] = sum_j T[k1][j] * (C[row][j] - P[row][j]);  B fragment of n-tile nt: b0 = bytes j 4t..4t+3 of row 8nt+g, b1 = j 16+4t..
    uint8_t *lo = s_lo[warp], *hi = s_hi[warp];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
        const uint8_t *cr = cur + (p.y + 8 * nt + g) * pitch + p.x + 4 * t, *pr = pred + (p.y + 8 * nt + g) * pitch + p.x + 4 * t;
        const uint32_t c0 = *reinterpret_cast<const uint32_t *>(cr), c1 = *reinterpret_cast<const uint32_t *>(cr + 16);
        const uint32_t p0 = *reinterpret_cast<const uint32_t *>(pr), p1 = *reinterpret_cast<const uint32_t *>(pr + 16);
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            int d[4] = { 0, 0, 0, 0 };
            mma_s8u8(d, aT[mt], c0, c1);
            mma_s8u8(d, aN[mt], p0, p1);
            // d0,d1: k1 = 16mt+g, rows 8nt+2t, +1;  d2,d3: k1 = 16mt+g+8, same rows.  Round, split into bytes, two values per 16-bit store.
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int v0 = (d[2 * h] + (1 << (S1 - 1))) >> S1, v1 = (d[2 * h + 1] + (1 << (S1 - 1))) >> S1;
                const int k1 = 16 * mt + g + 8 * h, r = 8 * nt + 2 * t;
                *reinterpret_cast<uint16_t *>(lo + k1 * 36 + r) = static_cast<uint16_t>((v0 & 255) | ((v1 & 255) << 8));
                *reinterpret_cast<uint16_t *>(hi + k1 * 36 + r) = static_cast<uint16_t>(((v0 >> 8) & 255) | (((v1 >> 8) & 255) << 8));
            }
        }
    }
    __syncwarp();
    // ---- stage 2: Z[k2][k1] = sum_r T[k2][r] * Y1[r][k1];  B fragment of n-tile nt (k1 = 8nt+g): b0 = rows r 4t..4t+3 = four bytes of plane row k1
    int16_t *o = out + static_cast<size_t>(u) * 1024;
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
        const int k1 = 8 * nt + g;
        const uint32_t l0 = *reinterpret_cast<const uint32_t *>(lo + k1 * 36 + 4 * t), l1 = *reinterpret_cast<const uint32_t *>(lo + k1 * 36 + 16 + 4 * t);
        const uint32_t h0 = *reinterpret_cast<const uint32_t *>(hi + k1 * 36 + 4 * t), h1 = *reinterpret_cast<const uint32_t *>(hi + k1 * 36 + 16 + 4 * t);
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            int dl[4] = { 0, 0, 0, 0 }, dh[4] = { 0, 0, 0, 0 };
            mma_s8u8(dl, aT[mt], l0, l1);
            mma_s8s8(dh, aT[mt], h0, h1);
            // d0,d1: k2 = 16mt+g, k1 = 8nt+2t, +1 -> adjacent coefficients of a row: one 32-bit store
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int z0 = ((dh[2 * h] << 8) + dl[2 * h] + (1 << (S2 - 1))) >> S2, z1 = ((dh[2 * h + 1] << 8) + dl[2 * h + 1] + (1 << (S2 - 1))) >> S2;
                *reinterpret_cast<uint32_t *>(o + (16 * mt + g + 8 * h) * 32 + 8 * nt + 2 * t) = (static_cast<uint32_t>(z0) & 0xffffu) | (static_cast<uint32_t>(z1) << 16);
            }
        }
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

int main(int argc, char **argv)
{
    const int W = 1920, H = 1080, pitch = 2048, reps = argc > 1 ? atoi(argv[1]) : 50;
    std::vector<uint8_t> hc(static_cast<size_t>(pitch) * H), hp(hc.size());
    uint32_t seed = 12345;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed >> 24; };
    for (size_t i = 0; i < hc.size(); i++) { hc[i] = static_cast<uint8_t>(rnd()); hp[i] = (i % 7 == 0) ? static_cast<uint8_t>(rnd()) : static_cast<uint8_t>(std::min(255u, hc[i] + (rnd() & 7u))); }
    // a few extreme units: residual +-255 everywhere / alternating (the only place where the 16-bit intermediate truncates)
    for (int r = 0; r < 32; r++) for (int c = 0; c < 64; c++) { hc[r * pitch + c] = (c < 32 || ((r + c) & 1)) ? 255 : 0; hp[r * pitch + c] = (c < 32 || ((r + c) & 1)) ? 0 : 255; }
    std::vector<int2> hxy;
    for (int y = 0; y + 32 <= H; y += 32) for (int x = 0; x + 32 <= W; x += 32) hxy.push_back(make_int2(x, y));
    const int n = static_cast<int>(hxy.size());
    int8_t hT[32][32];
    for (int k = 0; k < 32; k++) for (int j = 0; j < 32; j++) hT[k][j] = static_cast<int8_t>(h_coef(32, k, j));
    CK(cudaMemcpyToSymbol(c_T, hT, sizeof hT));
    uint8_t *dc, *dp; int2 *dxy; int16_t *oa, *ob;
    CK(cudaMalloc(&dc, hc.size())); CK(cudaMalloc(&dp, hp.size())); CK(cudaMalloc(&dxy, sizeof(int2) * n));
    CK(cudaMalloc(&oa, sizeof(int16_t) * 1024 * n)); CK(cudaMalloc(&ob, sizeof(int16_t) * 1024 * n));
    CK(cudaMemcpy(dc, hc.data(), hc.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dp, hp.data(), hp.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dxy, hxy.data(), sizeof(int2) * n, cudaMemcpyHostToDevice));
    const int grid = (n + 7) / 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms[2] = { 0, 0 };
    for (int v = 0; v < 2; v++) {
        for (int i = 0; i < 5; i++) { if (v == 0) k_fwd_imad<<<grid, 256>>>(dc, dp, pitch, dxy, n, oa); else k_fwd_mma<<<grid, 256>>>(dc, dp, pitch, dxy, n, ob); }
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; i++) { if (v == 0) k_fwd_imad<<<grid, 256>>>(dc, dp, pitch, dxy, n, oa); else k_fwd_mma<<<grid, 256>>>(dc, dp, pitch, dxy, n, ob); }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms[v], e0, e1));
        CK(cudaGetLastError());
    }
    std::vector<int16_t> ha(static_cast<size_t>(1024) * n), hb(ha.size());
    CK(cudaMemcpy(ha.data(), oa, ha.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hb.data(), ob, hb.size() * 2, cudaMemcpyDeviceToHost));
    size_t bad = 0, first = 0;
    for (size_t i = 0; i < ha.size(); i++) if (ha[i] != hb[i]) { if (!bad) first = i; bad++; }
    long nz = 0;
    for (size_t i = 0; i < ha.size(); i++) nz += ha[i] != 0;
    printf("{\"units\": %d, \"size\": 32, \"reps\": %d, \"imad_us_per_launch\": %.2f, \"mma_us_per_launch\": %.2f, \"speedup\": %.2f, \"mismatching_coefficients\": %zu, "
           "\"first_mismatch\": %zu, \"nonzero_coefficients\": %ld}\n", n, reps, 1e3 * ms[0] / reps, 1e3 * ms[1] / reps, ms[0] / ms[1], bad, first, nz);
    return bad ? 1 : 0;
}
