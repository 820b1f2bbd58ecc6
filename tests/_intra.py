"""helpers of the intra reconstruction tests: the transform-unit list of a picture from the per-4x4-unit decisions the reference's encoder left
(tests/_encode.py: encode_and_capture), in coding order, and the oracle's reconstruction of such a list"""
import ctypes as C

import numpy as np

from _oracle import chroma_qp, oracle

HOR, VER, DM = 10, 26, 36
SCAN_HOR, SCAN_VER, SCAN_DIAG = 1, 2, 3        # scan_pyramid index (hmr_private.h:92-94: HOR_SCAN, VER_SCAN, DIAG_SCAN)


def capture_intra_picture(w, h, qp=32, sign_hiding=1, seed=21):
    """the reference encoder's decisions, levels and unfiltered reconstruction of one synthetic I picture (tools/intra_capture.py in its own
    process, on the zero-initialised build of the reference when it exists: the as-is build's pictures depend on stack garbage)"""
    import os, subprocess, sys, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    zinit = os.path.join(root, "oracle", "_ref", "zinit")
    env = dict(os.environ)
    if "HB_REF_DIR" not in env and os.path.exists(os.path.join(zinit, "librefdrv.so")):
        env["HB_REF_DIR"] = zinit
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "cap.npz")
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "intra_capture.py"), str(w), str(h), str(qp), str(sign_hiding), str(seed), out],
                           capture_output=True, text=True, timeout=900, cwd=root, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        z = np.load(out)
        a = {k: z[k] for k in z.files}
    a["recon"] = [a.pop("recon_y"), a.pop("recon_u"), a.pop("recon_v")]
    a["slice_type"], a["slice_qp"], a["seconds"] = int(a["slice_type"]), int(a["slice_qp"]), float(a["seconds"])
    return a


def scan_mode(is_luma, size, mode):
    """find_scan_mode(TRUE, is_luma, size, mode, 0), hmr_tables.c:376: mode dependent for luma 4 / 8 and chroma 4 (and the 2x2 case)"""
    if (is_luma and size in (4, 8)) or (not is_luma and size in (2, 4)):
        return SCAN_HOR if abs(mode - VER) < 5 else (SCAN_VER if abs(mode - HOR) < 5 else SCAN_DIAG)
    return SCAN_DIAG


def intra_tus(a, w, h, chroma_qp_offset=2):
    """(n, 10) int32: comp, x, y, size, mode, qp, scan_mode, and the luma position / size of the quadtree node whose neighbour flags apply.
    Coding order: CTUs in raster order, coding units in z-order, inside a unit the luma transform tree, then U, then V."""
    out = []
    dep, trd, my, mc, qp = a["pred_depth"], a["tr_idx"], a["mode_y"], a["mode_c"], a["qp"]

    def tu_tree(x, y, s, d, cu_x, cu_y, luma, chroma):
        if x >= w or y >= h:
            return
        if int(trd[cu_y // 4, cu_x // 4] if False else trd[y // 4, x // 4]) > d and s > 4:
            for dy in (0, s // 2):
                for dx in (0, s // 2):
                    tu_tree(x + dx, y + dy, s // 2, d + 1, cu_x, cu_y, luma, chroma)
            return
        q = int(qp[y // 4, x // 4])
        m = int(my[y // 4, x // 4])
        luma.append((0, x, y, s, m, q, scan_mode(True, s, m), x, y, s))
        if s >= 8 or ((x & 7) == 0 and (y & 7) == 0):
            # chroma of this leaf; four 4x4 luma units share one 4x4 chroma unit at their 8x8 parent (coded with the first of them)
            cs, nx, ny, ns = (s // 2, x, y, s) if s >= 8 else (4, x, y, 8)
            first_luma_mode = int(my[cu_y // 4, cu_x // 4])
            cm = int(mc[ny // 4, nx // 4])
            cm = first_luma_mode if cm == DM else cm
            qc = chroma_qp(q, chroma_qp_offset)
            sm = scan_mode(False, s // 2, cm)
            for c in (1, 2):
                chroma.append((c, nx // 2, ny // 2, cs, cm, qc, sm, nx, ny, ns))

    def cu_tree(x, y, s, d):
        if x >= w or y >= h:
            return
        if int(dep[y // 4, x // 4]) > d and s > 8:
            for dy in (0, s // 2):
                for dx in (0, s // 2):
                    cu_tree(x + dx, y + dy, s // 2, d + 1)
            return
        luma, chroma = [], []
        tu_tree(x, y, s, 0, x, y, luma, chroma)
        out.extend(luma); out.extend(chroma)

    for cy in range(0, h, 64):
        for cx in range(0, w, 64):
            cu_tree(cx, cy, 64, 0)
    return np.array(out, np.int32).reshape(-1, 10)


class OrcTuOut(C.Structure):
    _fields_ = [("sum", C.c_int32), ("ssd", C.c_uint32), ("ssd_zero", C.c_uint32), ("zeroed", C.c_int32)]


def oracle_intra_recon(planes, w, h, tus, is_islice=1, sign_hiding=1, chroma_weight=1.0):
    """-> (reconstructed planes uint8, levels back to back int16, results)"""
    O = oracle()
    pad = 8
    org = [np.ascontiguousarray(np.pad(p.astype(np.int16), pad, mode="edge")) for p in planes]
    rec = [np.zeros_like(o) for o in org]
    n = len(tus)
    tus = np.ascontiguousarray(tus, np.int32)
    coeff = np.zeros(int((tus[:, 3].astype(np.int64) ** 2).sum()), np.int16)
    res = (OrcTuOut * n)()
    vp = C.c_void_p * 3
    ip = C.c_int * 3
    off = lambda arr: arr.ctypes.data + 2 * (pad * arr.shape[1] + pad)
    O.orc_intra_recon_tus.argtypes = [C.c_void_p, vp, ip, vp, ip, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    O.orc_intra_recon_tus(O.tables, vp(*[off(o) for o in org]), ip(*[o.shape[1] for o in org]), vp(*[off(r) for r in rec]), ip(*[r.shape[1] for r in rec]),
                          w, h, tus.ctypes.data, n, is_islice, sign_hiding, chroma_weight, coeff.ctypes.data, res)
    out = [np.clip(r[pad:-pad, pad:-pad], 0, 255).astype(np.uint8) for r in rec]
    return out, coeff, res


def coeff_wnd_of(tus, coeff, w, h):
    """the levels of a unit list in the reference's ctu->coeff_wnd layout: (n_ctus, 6144)"""
    cols, rows = (w + 63) // 64, (h + 63) // 64
    out = np.zeros((cols * rows, 64 * 64 + 2 * 32 * 32), np.int16)

    def z(ux, uy):
        return sum((((ux >> b) & 1) << (2 * b)) | (((uy >> b) & 1) << (2 * b + 1)) for b in range(4))
    o = 0
    for (c, x, y, s, *_rest) in tus:
        lx, ly = (x * 2, y * 2) if c else (x, y)
        ctu = (ly // 64) * cols + lx // 64
        a = z((lx & 63) // 4, (ly & 63) // 4)
        base = 0 if c == 0 else (4096 if c == 1 else 5120)
        offs = (a << 4) >> 2 if c else a << 4
        out[ctu, base + offs:base + offs + s * s] = coeff[o:o + s * s]
        o += s * s
    return out
