// hb_shim.cu -- CUDA runtime wrappers and the small per-call kernels (int16 operands, reference semantics of
// hmr_motion_intra.c:51-186 and hmr_motion_inter.c:261-391, :878-936).  See hb_shim.h.
#include "hb_shim.h"
#include "hb_dev_common.cuh"

// ------------------------------------------------------------------ runtime wrappers
extern "C" int hbc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int hbc_set_device(int dev) { return static_cast<int>(cudaSetDevice(dev)); }
extern "C" int hbc_stream_create(void **stream)
{
    cudaStream_t s;
    const cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    *stream = s;
    return static_cast<int>(e);
}
extern "C" int hbc_stream_destroy(void *stream) { return static_cast<int>(cudaStreamDestroy(static_cast<cudaStream_t>(stream))); }
extern "C" int hbc_stream_sync(void *stream) { return static_cast<int>(cudaStreamSynchronize(static_cast<cudaStream_t>(stream))); }
extern "C" int hbc_malloc(void **p, size_t bytes) { return static_cast<int>(cudaMalloc(p, bytes)); }
extern "C" int hbc_free(void *p) { return static_cast<int>(cudaFree(p)); }
extern "C" int hbc_host_alloc(void **p, size_t bytes) { return static_cast<int>(cudaHostAlloc(p, bytes, cudaHostAllocMapped | cudaHostAllocPortable)); }
extern "C" int hbc_host_free(void *p) { return static_cast<int>(cudaFreeHost(p)); }
extern "C" int hbc_host_devptr(void *host, void **dev) { return static_cast<int>(cudaHostGetDevicePointer(dev, host, 0)); }
extern "C" int hbc_memset_async(void *p, int v, size_t bytes, void *stream) { return static_cast<int>(cudaMemsetAsync(p, v, bytes, static_cast<cudaStream_t>(stream))); }
extern "C" int hbc_h2d_async(void *dst, const void *src, size_t bytes, void *stream)
{
    return static_cast<int>(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
}
extern "C" int hbc_d2h_async(void *dst, const void *src, size_t bytes, void *stream)
{
    return static_cast<int>(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
}
extern "C" int hbc_h2d_2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, void *stream)
{
    return static_cast<int>(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, rows, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
}
extern "C" int hbc_d2d_2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, void *stream)
{
    return static_cast<int>(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, rows, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
}
extern "C" int hbc_d2h_2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, void *stream)
{
    return static_cast<int>(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, rows, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
}
extern "C" int hbc_event_create(void **ev)
{
    cudaEvent_t e;
    const cudaError_t r = cudaEventCreate(&e);
    *ev = e;
    return static_cast<int>(r);
}
extern "C" int hbc_event_destroy(void *ev) { return static_cast<int>(cudaEventDestroy(static_cast<cudaEvent_t>(ev))); }
extern "C" int hbc_event_record(void *ev, void *stream) { return static_cast<int>(cudaEventRecord(static_cast<cudaEvent_t>(ev), static_cast<cudaStream_t>(stream))); }
extern "C" int hbc_event_elapsed(void *a, void *b, float *ms)
{
    cudaError_t e = cudaEventSynchronize(static_cast<cudaEvent_t>(b));
    if (e != cudaSuccess) return static_cast<int>(e);
    return static_cast<int>(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(a), static_cast<cudaEvent_t>(b)));
}
extern "C" int hbc_event_create_notiming(void **ev)
{
    cudaEvent_t e;
    const cudaError_t r = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    *ev = e;
    return static_cast<int>(r);
}
extern "C" int hbc_stream_wait_event(void *stream, void *ev)
{
    return static_cast<int>(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), static_cast<cudaEvent_t>(ev), 0));
}
extern "C" int hbc_graph_begin(void *stream)
{
    return static_cast<int>(cudaStreamBeginCapture(static_cast<cudaStream_t>(stream), cudaStreamCaptureModeThreadLocal));
}
extern "C" int hbc_graph_end(void *stream, void **exec)
{
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(static_cast<cudaStream_t>(stream), &g);
    if (e != cudaSuccess) return static_cast<int>(e);
    cudaGraphExec_t x = nullptr;
    e = cudaGraphInstantiate(&x, g, 0);
    cudaGraphDestroy(g);
    *exec = x;
    return static_cast<int>(e);
}
extern "C" int hbc_graph_launch(void *exec, void *stream) { return static_cast<int>(cudaGraphLaunch(static_cast<cudaGraphExec_t>(exec), static_cast<cudaStream_t>(stream))); }
extern "C" int hbc_graph_destroy(void *exec) { return static_cast<int>(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(exec))); }
extern "C" const char *hbc_error_string(int code) { return cudaGetErrorString(static_cast<cudaError_t>(code)); }

// ------------------------------------------------------------------ per-call pixel kernels
namespace {

// sad / ssd16b of the function table with the EXACT lane arithmetic of the reference's SSE4.2 functions (hmr_sse42_functions_pixel.c:
// sad :330-460, ssd16b :619-745).  On 8-bit video they are the plain sums of hmr_motion_intra.c:51 / :128; the encoder, however, also
// calls them on operands that are not video -- its 64x64 intra mode search predicts with 16-bit arithmetic that wraps
// (homer_loop1_motion_intra, hmr_motion_intra.c:1084 -> :1130) -- and there the SIMD lanes decide the value the mode decision sees:
//   difference: 16-bit wrapping subtraction, |.| with |-32768| = 32768 (psubw, pabsw);
//   sad 4/8/16: the eight 16-bit lanes wrap while they accumulate and while they are folded 8 -> 4 -> 2, the last two add in 32 bits;
//   sad 32    : sixteen lanes (two accumulators) add with unsigned saturation over all rows, then widen;
//   sad 64    : per row, eight lanes add their eight column groups with unsigned saturation, then widen;
//   ssd16b    : squares of the wrapped difference, everything modulo 2^32 (pmaddwd, paddd).
__device__ __forceinline__ uint32_t pc_absdiff16(int a, int b)
{
    const int t = static_cast<int16_t>(a - b);
    return static_cast<uint32_t>(t < 0 ? -t : t);        // 32768 for t = -32768, as pabsw leaves 0x8000
}

__global__ void __launch_bounds__(256) k_pc_sad(const int16_t *a, int as, const int16_t *b, int bs, int n, int squared, uint32_t *out)
{
    __shared__ uint32_t lane[16], total;
    if (threadIdx.x < 16) lane[threadIdx.x] = 0;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    if (squared) {
        uint32_t acc = 0;
        for (int e = threadIdx.x; e < n * n; e += 256) {
            const int t = static_cast<int16_t>(a[(e / n) * as + e % n] - b[(e / n) * bs + e % n]);
            acc += static_cast<uint32_t>(t * t);
        }
        acc = __reduce_add_sync(HB_FULL_MASK, acc);
        if ((threadIdx.x & 31) == 0) atomicAdd(&total, acc);
        __syncthreads();
        if (threadIdx.x == 0) *out = total;
        return;
    }
    if (n == 64) {                                            // 64 rows x 8 lanes: saturate per row, then exact
        uint32_t acc = 0;
        for (int u = threadIdx.x; u < 64 * 8; u += 256) {
            const int r = u >> 3, j = u & 7;
            uint32_t s = 0;
            for (int g = 0; g < 8; g++) s += pc_absdiff16(a[r * as + 8 * g + j], b[r * bs + 8 * g + j]);
            acc += min(s, 65535u);
        }
        acc = __reduce_add_sync(HB_FULL_MASK, acc);
        if ((threadIdx.x & 31) == 0) atomicAdd(&total, acc);
        __syncthreads();
        if (threadIdx.x == 0) *out = total;
        return;
    }
    for (int e = threadIdx.x; e < n * n; e += 256) {
        const int r = e / n, c = e % n;
        const uint32_t d = pc_absdiff16(a[r * as + c], b[r * bs + c]);
        // lane of the element: 4x4 packs two rows per vector; 32x32 keeps columns 0..15 and 16..31 in two accumulators
        const int l = n == 4 ? ((r & 1) * 4 + c) : n == 32 ? ((c >> 4) * 8 + (c & 7)) : (c & 7);
        atomicAdd(&lane[l], d);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t r;
        if (n == 32) {
            r = 0;
            for (int l = 0; l < 16; l++) r += min(lane[l], 65535u);
        } else {
            r = ((lane[0] + lane[4] + lane[2] + lane[6]) & 0xffffu) + ((lane[1] + lane[5] + lane[3] + lane[7]) & 0xffffu);
        }
        *out = r;
    }
}

__global__ void __launch_bounds__(256) k_pc_predict(const int16_t *orig, int os, const int16_t *pred, int ps, int16_t *res, int rs, int n)
{
    for (int e = threadIdx.x; e < n * n; e += 256)
        res[(e / n) * rs + e % n] = static_cast<int16_t>(orig[(e / n) * os + e % n] - pred[(e / n) * ps + e % n]);
}

__global__ void __launch_bounds__(256) k_pc_reconst(const int16_t *pred, int ps, const int16_t *res, int rs, int16_t *dec, int ds, int n)
{
    for (int e = threadIdx.x; e < n * n; e += 256)
        dec[(e / n) * ds + e % n] = static_cast<int16_t>(hb_clip255(static_cast<int>(res[(e / n) * rs + e % n]) + pred[(e / n) * ps + e % n]));
}

// bi-prediction average of two 14-bit predictions (weighted_average_motion, hmr_motion_inter.c:2903)
__global__ void __launch_bounds__(256) k_pc_wavg(const int16_t *s0, const int16_t *s1, int16_t *dst, int w, int h)
{
    for (int e = threadIdx.x; e < w * h; e += 256) dst[e] = static_cast<int16_t>(hb_clip255((static_cast<int>(s0[e]) + s1[e] + 64 + 2 * 8192) >> 7));
}

// one interpolation pass with the reference's four (is_first, is_last) roundings; every value passes through an int16
__global__ void __launch_bounds__(256) k_pc_interp(const int16_t *src, int ss, int org_off, int16_t *dst, int ds, int chroma, int fraction,
                                                   int w, int h, int vertical, int first, int last)
{
    const int16_t *org = src + org_off;
    const int step = vertical ? ss : 1;
    for (int e = threadIdx.x; e < w * h; e += 256) {
        const int r = e / w, c = e % w;
        const int16_t *p = org + r * ss + c;
        int v;
        if (!chroma && fraction == 0) {                         // filter_copy, hmr_motion_inter.c:261
            if (first == last) v = p[0];
            else if (first) v = static_cast<int16_t>(static_cast<int16_t>(p[0] << 6) - 8192);
            else v = hb_clip255(static_cast<int16_t>((p[0] + 8224) >> 6));
        } else {
            int shift = 6, offset;
            if (last) { shift += first ? 0 : 6; offset = (1 << (shift - 1)) + (first ? 0 : (8192 << 6)); }
            else { shift -= first ? 6 : 0; offset = first ? -(8192 << shift) : 0; }
            int sum;
            if (chroma) sum = hb_chroma4_dyn(fraction, p[-step], p[0], p[step], p[2 * step]);
            else sum = hb_luma8_dyn(fraction, p[-3 * step], p[-2 * step], p[-step], p[0], p[step], p[2 * step], p[3 * step], p[4 * step]);
            v = static_cast<int16_t>((sum + offset) >> shift);
            if (last) v = hb_clip255(v);
        }
        dst[r * ds + c] = static_cast<int16_t>(v);
    }
}

}  // namespace

extern "C" int hbk_pc_sad(const int16_t *a, int as, const int16_t *b, int bs, int n, int squared, uint32_t *out, void *stream)
{
    k_pc_sad<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, as, b, bs, n, squared, out);
    return static_cast<int>(cudaGetLastError());
}
extern "C" int hbk_pc_predict(const int16_t *orig, int os, const int16_t *pred, int ps, int16_t *res, int rs, int n, void *stream)
{
    k_pc_predict<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(orig, os, pred, ps, res, rs, n);
    return static_cast<int>(cudaGetLastError());
}
extern "C" int hbk_pc_reconst(const int16_t *pred, int ps, const int16_t *res, int rs, int16_t *dec, int ds, int n, void *stream)
{
    k_pc_reconst<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, ps, res, rs, dec, ds, n);
    return static_cast<int>(cudaGetLastError());
}
extern "C" int hbk_pc_wavg(const int16_t *s0, const int16_t *s1, int16_t *dst, int w, int h, void *stream)
{
    k_pc_wavg<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(s0, s1, dst, w, h);
    return static_cast<int>(cudaGetLastError());
}
extern "C" int hbk_pc_interp(const int16_t *src, int ss, int org_off, int16_t *dst, int ds, int chroma, int fraction, int w, int h,
                             int vertical, int first, int last, void *stream)
{
    k_pc_interp<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, ss, org_off, dst, ds, chroma, fraction, w, h, vertical, first, last);
    return static_cast<int>(cudaGetLastError());
}

// ---- CUDA IPC: another process's device allocations and events (CTU-row bands, one process per GPU)
static_assert(sizeof(cudaIpcMemHandle_t) == 64 && sizeof(cudaIpcEventHandle_t) == 64, "HB_IPC_HANDLE_BYTES");
extern "C" int hbc_ipc_get_mem(void *dev, unsigned char out[64])
{
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, dev);
    if (e == cudaSuccess) memcpy(out, &h, 64);
    return static_cast<int>(e);
}
extern "C" int hbc_ipc_open_mem(const unsigned char in[64], void **dev)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, in, 64);
    return static_cast<int>(cudaIpcOpenMemHandle(dev, h, cudaIpcMemLazyEnablePeerAccess));
}
extern "C" int hbc_ipc_close_mem(void *dev) { return static_cast<int>(cudaIpcCloseMemHandle(dev)); }
extern "C" int hbc_ipc_event_create(void **ev, unsigned char out[64])
{
    cudaEvent_t e;
    cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventInterprocess);
    if (rc != cudaSuccess) return static_cast<int>(rc);
    cudaIpcEventHandle_t h;
    rc = cudaIpcGetEventHandle(&h, e);
    if (rc != cudaSuccess) { cudaEventDestroy(e); return static_cast<int>(rc); }
    memcpy(out, &h, 64);
    *ev = e;
    return 0;
}
extern "C" int hbc_ipc_event_open(const unsigned char in[64], void **ev)
{
    cudaIpcEventHandle_t h;
    memcpy(&h, in, 64);
    cudaEvent_t e;
    const cudaError_t rc = cudaIpcOpenEventHandle(&e, h);
    if (rc == cudaSuccess) *ev = e;
    return static_cast<int>(rc);
}

// rows of up to twelve plane pieces, each from its own source (peer memory over NVLink or local) into this GPU's picture: one CTA per
// row, 16 bytes per thread where the row allows it
struct PullArgs { hbd_pull_span s[12]; };
__global__ void __launch_bounds__(256) k_pull_rows(const PullArgs a)
{
    const hbd_pull_span sp = a.s[blockIdx.y];
    const int row = blockIdx.x;
    if (row >= sp.rows) return;
    const uint8_t *src = sp.src + static_cast<size_t>(row) * sp.src_pitch;
    uint8_t *dst = sp.dst + static_cast<size_t>(row) * sp.dst_pitch;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | static_cast<uintptr_t>(sp.width)) & 15) == 0) {
        for (int i = threadIdx.x; i < sp.width / 16; i += 256) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
    } else {
        for (int i = threadIdx.x; i < sp.width / 4; i += 256) reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(src)[i];
    }
}
extern "C" int hbk_pull_rows(const hbd_pull_span *spans, int n_spans, void *stream)
{
    if (n_spans <= 0) return 0;
    if (n_spans > 12) return static_cast<int>(cudaErrorInvalidValue);
    PullArgs a;
    memset(&a, 0, sizeof a);
    int rows = 0;
    for (int i = 0; i < n_spans; i++) { a.s[i] = spans[i]; rows = spans[i].rows > rows ? spans[i].rows : rows; }
    if (rows <= 0) return 0;
    k_pull_rows<<<dim3(rows, n_spans), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return static_cast<int>(cudaGetLastError());
}

extern "C" int hbc_clear_error(void) { return static_cast<int>(cudaGetLastError()); }

// Can [p, p + bytes) of host memory go up as ONE copy?  Pageable memory: yes, the runtime stages it.  Pinned memory: only when the span lies
// inside one allocation -- adjacent addresses need not be (three separately pinned planes that happen to touch), and a copy across
// allocations is refused.  The allocation's range comes from the driver (cuMemGetAddressRange on the pointer's unified address).
extern "C" int hbc_host_span_one_copy(const void *p, size_t bytes)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return 1; }
    if (at.type == cudaMemoryTypeUnregistered) return 1;
    typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
    static range_fn fn = nullptr;
    static int tried = 0;
    if (!tried) {
        tried = 1;
        void *q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess) fn = reinterpret_cast<range_fn>(q);
        else cudaGetLastError();
    }
    if (!fn || !at.devicePointer) return 0;
    unsigned long long base = 0;
    size_t size = 0;
    const unsigned long long d = reinterpret_cast<unsigned long long>(at.devicePointer);
    if (fn(&base, &size, d) != 0) return 0;
    return d + bytes <= base + size;
}
