// hb_kernels_enc.cu -- block read-back for the CU-granularity encoder calls (hb_encode.c): square blocks of resident 8-bit
// planes are widened into tight int16 blocks, the sample format of the reference's CTU windows (wnd_t, hmr_private.h:660),
// so that one D2H copy returns every block a call produced.
#include "hb_shim.h"
#include "hb_dev_common.cuh"

namespace {

// one CTA per block; four samples per thread and step (blocks are 4..64 wide, x is a multiple of 4: aligned 32-bit loads)
__global__ void __launch_bounds__(256) k_fetch_blocks(const hbd_frame f, const hbd_fetch_job *__restrict__ jobs, int16_t *__restrict__ out)
{
    const hbd_fetch_job j = jobs[blockIdx.x];
    const hbd_plane p = hbd_pick_plane(f, j.comp);
    const int wpr = j.size >> 2, n_words = wpr * j.size;
    int16_t *dst = out + j.off;
    for (int w = threadIdx.x; w < n_words; w += blockDim.x) {
        const int r = w / wpr, c4 = (w % wpr) * 4;
        const uint32_t v = *reinterpret_cast<const uint32_t *>(p.org + (j.y + r) * p.pitch + j.x + c4);
        uint2 o;
        o.x = (v & 0xffu) | ((v & 0xff00u) << 8);
        o.y = ((v >> 16) & 0xffu) | ((v >> 8) & 0xff0000u);
        *reinterpret_cast<uint2 *>(dst + r * j.size + c4) = o;
    }
}

}  // namespace

extern "C" int hbk_fetch_blocks(const hbd_frame *f, const hbd_fetch_job *jobs, int n_jobs, int16_t *out, void *stream)
{
    if (n_jobs <= 0) return 0;
    k_fetch_blocks<<<n_jobs, 256, 0, static_cast<cudaStream_t>(stream)>>>(*f, jobs, out);
    return static_cast<int>(cudaGetLastError());
}
