"""Host-side preparation of the intra mode pre-search: every luma block of the given sizes with the
reference samples taken from the ORIGINAL picture (the usual open-loop pre-selection; the exact closed-loop samples come from
the reconstruction, block by block, on the host).  Layout of a block's 4n+1 samples as in the reference
(hmr_motion_intra.c:189): index 2n = top-left corner, 2n+i above (left to right), 2n-i left (top to bottom); samples outside
the picture are 128."""
import numpy as np


def presearch_jobs(luma, sizes=(32, 16, 8, 4)):
    """luma: (h, w) uint8.  Returns (jobs int32 (n,3) {x, y, size}, adi int16 flat, adi_off int32 (n,))"""
    h, w = luma.shape
    big = max(sizes)
    pad = np.full((h + 1 + 2 * big, w + 1 + 2 * big), 128, np.int16)          # one row / column of context before, 2n after
    pad[1:h + 1, 1:w + 1] = luma
    jobs, adis, offs, off = [], [], [], 0
    for n in sizes:
        ys, xs = np.arange(0, h - n + 1, n), np.arange(0, w - n + 1, n)
        gy, gx = np.meshgrid(ys, xs, indexing="ij")
        gy, gx = gy.reshape(-1), gx.reshape(-1)
        k = len(gy)
        a = np.empty((k, 4 * n + 1), np.int16)
        i = np.arange(0, 2 * n + 1)
        a[:, 2 * n:] = pad[gy[:, None], gx[:, None] + i[None, :]]             # corner + above: row y-1, columns x-1 .. x+2n-1
        j = np.arange(1, 2 * n + 1)
        a[:, 2 * n - j] = pad[gy[:, None] + j[None, :], gx[:, None]]          # left: column x-1, rows y .. y+2n-1
        jobs.append(np.stack([gx, gy, np.full(k, n)], 1).astype(np.int32))
        adis.append(a.reshape(-1))
        offs.append(off + np.arange(k, dtype=np.int64) * (4 * n + 1))
        off += k * (4 * n + 1)
    return np.concatenate(jobs), np.concatenate(adis), np.concatenate(offs).astype(np.int32)
