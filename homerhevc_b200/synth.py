"""Synthetic YUV 4:2:0 clips: frame n is a crop of a fixed blurred block texture panned by
((3n) mod 48, (2n) mod 48), plus per-frame Gaussian noise; U = Y(2x sub)/2 + 64, V = 255 - U.  numpy only."""
import numpy as np

SEEDS = {(1280, 720): 1234, (1920, 1080): 1, (3840, 2160): 2}


def _box(a, k, axis):
    c = np.cumsum(np.pad(a, [(k // 2 + 1, k // 2) if i == axis else (0, 0) for i in range(2)], mode="edge"), axis=axis)
    n = a.shape[axis]
    hi = np.take(c, np.arange(k, k + n), axis=axis)
    lo = np.take(c, np.arange(0, n), axis=axis)
    return (hi - lo) / k


def make_texture(width, height, seed=None):
    seed = SEEDS.get((width, height), 7) if seed is None else seed
    rng = np.random.default_rng(seed)
    th, tw = height + 64, width + 64
    blocks = rng.integers(0, 256, size=((th + 7) // 8, (tw + 7) // 8)).astype(np.float32)
    tex = np.kron(blocks, np.ones((8, 8), np.float32))[:th, :tw]
    return _box(_box(tex, 9, 0), 9, 1).astype(np.float32)


def make_frame(tex, width, height, n, noise=3.0, seed=0):
    """returns (y, u, v) uint8 planes of frame n"""
    ox, oy = (3 * n) % 48, (2 * n) % 48
    rng = np.random.default_rng(seed * 100003 + n)
    y = tex[oy:oy + height, ox:ox + width]
    if noise:
        y = y + rng.normal(0, noise, size=y.shape).astype(np.float32)
    y = np.clip(np.rint(y), 0, 255).astype(np.uint8)
    u = (y[::2, ::2].astype(np.int32) // 2 + 64).astype(np.uint8)
    v = (255 - u.astype(np.int32)).astype(np.uint8)
    return y, u, v


def make_clip(width, height, n_frames, noise=3.0, seed=None):
    tex = make_texture(width, height, seed)
    return [make_frame(tex, width, height, n, noise) for n in range(n_frames)]
