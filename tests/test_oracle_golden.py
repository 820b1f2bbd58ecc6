"""The oracle against golden vectors produced by the UNMODIFIED compiled reference (tests/golden/make_golden.py).
Runs anywhere (no GPU, no /root/reference): this is what pins the oracle on the GPU box."""
import ctypes as C
import os

import numpy as np
import pytest

from _oracle import (OrcMeIn, OrcMeOut, OrcMv, OrcTuOut, aligned_i16, oracle, ptr)

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))


def test_pixel_kernels():
    O = oracle()
    a = aligned_i16(64 * 64); b = aligned_i16(96 * 80); res = aligned_i16(64 * 64)
    a[:], b[:], res[:] = G["pix_a"], G["pix_b"], G["pix_res"]
    i = 0; po = 0
    for n in (4, 8, 16, 32, 64):
        for off in G["pix_off"][:2]:
            assert O.orc_sad(ptr(a), 64, ptr(b, int(off)), 96, n) == G["pix_sad"][i]
            assert O.orc_ssd16b(ptr(a), 64, ptr(b, int(off)), 96, n) == G["pix_ssd"][i]
            i += 1
        r = aligned_i16(64 * 64); d = aligned_i16(64 * 64)
        O.orc_predict(ptr(a), 64, ptr(b), 96, ptr(r), 64, n)
        O.orc_reconst(ptr(b), 96, ptr(res), 64, ptr(d), 64, n)
        assert np.array_equal(r.reshape(64, 64)[:n, :n].reshape(-1), G["pix_predict"][po:po + n * n])
        assert np.array_equal(d.reshape(64, 64)[:n, :n].reshape(-1), G["pix_reconst"][po:po + n * n])
        po += n * n


def test_interpolation():
    O = oracle()
    s8 = aligned_i16(96 * 96); s14 = aligned_i16(96 * 96)
    s8[:], s14[:] = G["int_src8"], G["int_src14"]
    o = 0
    for chroma, frac, first, last, vert, w, h in G["int_cases"]:
        d = aligned_i16(80 * 80)
        f = O.orc_interpolate_chroma if chroma else O.orc_interpolate_luma
        f(ptr(s8 if first else s14, 96 * 8 + 8), 96, ptr(d), 80, int(frac), int(w), int(h), int(vert), int(first), int(last))
        assert np.array_equal(d.reshape(80, 80)[:h, :w].reshape(-1), G["int_out"][o:o + w * h]), (chroma, frac, first, last, vert, w, h)
        o += w * h


def test_transforms():
    O = oracle()
    blk = aligned_i16(64 * 64); blk[:] = G["tx_block"]
    o = 0
    for n, dst in G["tx_cases"]:
        n = int(n)
        c = aligned_i16(1024); co = aligned_i16(1024); bb = aligned_i16(64 * 64)
        O.orc_transform(8, ptr(blk), 64, ptr(c), n, int(dst))
        assert np.array_equal(c[:n * n], G["tx_out"][o:o + n * n]); o += n * n
        co[:n * n] = G["tx_coef"][:n * n]
        O.orc_itransform(8, ptr(bb), 64, ptr(co), n, int(dst))
        assert np.array_equal(bb.reshape(64, 64)[:n, :n].reshape(-1), G["tx_out"][o:o + n * n]); o += n * n


def test_tables():
    O = oracle()
    for key, mode, lgs in (("scan_diag", 3, (2, 3, 4, 5)), ("scan_hor", 1, (2, 3)), ("scan_ver", 2, (2, 3))):
        got = np.concatenate([np.ctypeslib.as_array(O.orc_tables_scan(O.tables, mode, lg), ((1 << lg) ** 2,)) for lg in lgs])
        assert np.array_equal(got, G[key]), key
    qt = []
    for lg in (2, 3, 4, 5):
        for lst in ((0, 3, 4, 5) if lg < 5 else (0, 3)):
            for rem in (0, 3, 5):
                qt.append(np.ctypeslib.as_array(O.orc_tables_quant(O.tables, lg, lst, rem), ((1 << lg) ** 2,)))
                qt.append(np.ctypeslib.as_array(O.orc_tables_dequant(O.tables, lg, lst, rem), ((1 << lg) ** 2,)))
    assert np.array_equal(np.concatenate(qt), G["qtables"])


def test_quant_and_dequant():
    O = oracle()
    so = oo = 0
    for lg, comp, is_intra, isl, sh, qp, scan, esum in G["q_cases"]:
        n2 = (1 << int(lg)) ** 2
        src = aligned_i16(1024); src[:n2] = G["q_src"][so:so + n2]; so += n2
        d = aligned_i16(1024); u = aligned_i16(1024); q = aligned_i16(1024); s = C.c_int(0)
        O.orc_quant(O.tables, ptr(src), ptr(d), ptr(u), int(scan), int(lg), int(comp), int(is_intra), int(isl), int(sh), int(qp) // 6, int(qp) % 6, C.byref(s))
        O.orc_inv_quant(O.tables, ptr(d), ptr(q), int(lg), int(comp), int(is_intra), int(qp) // 6, int(qp) % 6)
        assert s.value == esum
        for arr in (d, u, q):
            assert np.array_equal(arr[:n2], G["q_out"][oo:oo + n2]), (lg, comp, is_intra, isl, sh, qp)
            oo += n2


def _padded(plane, pad):
    return np.ascontiguousarray(np.pad(plane.astype(np.int16), pad, mode="edge"))


def test_motion_estimation_and_compensation():
    O = oracle()
    PAD = 80
    cur, rf = _padded(G["me_cur"], PAD), _padded(G["me_ref"], PAD)
    H, W = G["me_cur"].shape
    S = W + 2 * PAD
    for case, exp in zip(G["me_cases"], G["me_out"]):
        gx, gy, n, qp, avg, action, ns = (int(v) for v in case[:7])
        amvp, st = case[7:11], case[11:17]
        mi = OrcMeIn()
        mi.orig = ptr(cur.reshape(-1), (PAD + gy) * S + PAD + gx); mi.orig_stride = S
        mi.ref = ptr(rf.reshape(-1), (PAD + gy) * S + PAD + gx); mi.ref_stride = S
        mi.gx, mi.gy, mi.size, mi.frame_w, mi.frame_h, mi.range_x, mi.range_y = gx, gy, n, W, H, 128, 64
        mi.n_amvp = 2
        for i in range(2):
            mi.amvp[i] = OrcMv(int(amvp[2 * i]), int(amvp[2 * i + 1]))
        mi.n_start = ns
        for i in range(3):
            mi.start[i] = OrcMv(int(st[2 * i]), int(st[2 * i + 1]))
        mi.qp, mi.avg_dist, mi.action = qp, float(avg), action
        mo = OrcMeOut()
        O.orc_motion_estimation(C.byref(mi), C.byref(mo))
        assert [mo.mv.x, mo.mv.y, mo.subpix.x, mo.subpix.y, mo.sad] == [int(v) for v in exp], (gx, gy, n, action)
    o = 0
    for chroma, gx, gy, n, mvx, mvy in G["mc_cases"]:
        n = int(n)
        p = aligned_i16(64 * 64)
        f = O.orc_mc_chroma if chroma else O.orc_mc_luma
        f(ptr(rf.reshape(-1), (PAD + int(gy)) * S + PAD + int(gx)), S, ptr(p), 64, n, OrcMv(int(mvx), int(mvy)))
        assert np.array_equal(p.reshape(64, 64)[:n, :n].reshape(-1), G["mc_out"][o:o + n * n]), (chroma, gx, gy, n, mvx, mvy)
        o += n * n


def test_inter_tq_chain():
    O = oracle()
    io = oo = 0
    coded = zeroed = 0
    for depth, comp, n, qp, qp_c, isl, sh, avg, esum, essd in G["tq_cases"]:
        n = int(n); n2 = n * n
        orig = aligned_i16(n2); pred = aligned_i16(n2)
        orig[:] = G["tq_in"][io:io + n2]; pred[:] = G["tq_in"][io + n2:io + 2 * n2]; io += 2 * n2
        co = aligned_i16(1024); de = aligned_i16(1024); to = OrcTuOut()
        weight = 1.0 if comp == 0 else 2.0 ** ((int(qp) - int(qp_c)) / 3.0)
        O.orc_encode_inter_tu(O.tables, ptr(orig), n, ptr(pred), n, ptr(co), ptr(de), n, n, int(comp), int(qp) if comp == 0 else int(qp_c),
                              int(isl), int(sh), float(avg), weight, C.byref(to))
        assert (to.sum, to.ssd) == (int(esum), int(essd)), (depth, comp, n, qp)
        assert np.array_equal(co[:n2], G["tq_out"][oo:oo + n2]); oo += n2
        assert np.array_equal(de[:n2], G["tq_out"][oo:oo + n2]); oo += n2
        coded += to.sum > 0; zeroed += to.zeroed
    assert coded > 20 and zeroed > 5


def test_intra_prediction_and_weighted_average():
    """reference-sample smoothing, all 35 predictors (luma; the chroma variants of DC / pure horizontal / vertical) and the
    bi-prediction average against the compiled reference's outputs"""
    O = oracle()
    ia = io = 0
    for k in range(8):                                   # (size, kind) blocks in generation order
        n = (4, 8, 16, 32)[k // 2]
        m = 4 * n + 1
        adi = aligned_i16(m); adi[:] = G["ip_adi"][ia:ia + m]
        flt = aligned_i16(m)
        O.orc_adi_filter(ptr(adi), ptr(flt), n, 1)
        assert np.array_equal(flt, G["ip_flt"][ia:ia + m]), ("filter", n, k % 2)
        ia += m
        for (cn, kind, mode, is_luma) in G["ip_cases"]:
            if cn != n or kind != k % 2:
                continue
            pr = aligned_i16(n * n)
            O.orc_intra_predict(ptr(flt if (kind and is_luma) else adi), n, int(mode), int(is_luma), ptr(pr), n)
            assert np.array_equal(pr.astype(np.uint8), G["ip_pred"][io:io + n * n]) and pr.min() >= 0 and pr.max() <= 255, (n, kind, mode, is_luma)
            io += n * n
    assert io == len(G["ip_pred"])
    a = aligned_i16(64 * 64); b = aligned_i16(64 * 64); d = aligned_i16(64 * 64)
    a[:] = G["wavg_a"]; b[:] = G["wavg_b"]
    O.orc_weighted_average(ptr(a), 64, ptr(b), 64, ptr(d), 64, 64, 64)
    assert np.array_equal(d.astype(np.uint8), G["wavg_out"]) and d.min() >= 0 and d.max() <= 255


def test_sao_statistics():
    from _oracle import oracle_sao_stats
    w, h = 200, 136
    def planes(a):
        return [a[:w * h].reshape(h, w), a[w * h:w * h * 5 // 4].reshape(h // 2, w // 2), a[w * h * 5 // 4:].reshape(h // 2, w // 2)]
    got = oracle_sao_stats(planes(G["sao_rec"]), planes(G["sao_org"]), w, h)
    assert np.array_equal(np.frombuffer(got.tobytes(), np.int32), G["sao_stats"])


def test_deblocking():
    from _oracle import oracle_deblock
    w, h = 192, 136
    def planes(a):
        return [a[:w * h].reshape(h, w), a[w * h:w * h * 5 // 4].reshape(h // 2, w // 2), a[w * h * 5 // 4:].reshape(h // 2, w // 2)]
    got = oracle_deblock(planes(G["dbk_in"]), w, h, G["dbk_bsv"], G["dbk_bsh"], G["dbk_qp"], tuple(int(v) for v in G["dbk_offs"]))
    exp = planes(G["dbk_out"])
    for c in range(3):
        assert np.array_equal(got[c], exp[c]), c
    assert (exp[0] != planes(G["dbk_in"])[0]).sum() > 1000


# ---- round 2: tests/golden/ref_vectors_r02.npz (tests/golden/make_golden_r02.py)
G2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors_r02.npz"))


def _units(prefix):
    return {k: G2[f"{prefix}_{k}"] for k in ("cu", "tu", "intra", "cbf", "qp", "mv")}


def _planes(a, w, h):
    return [a[:w * h].reshape(h, w), a[w * h:w * h * 5 // 4].reshape(h // 2, w // 2), a[w * h * 5 // 4:].reshape(h // 2, w // 2)]


def test_amvp_and_merge_candidates():
    from _oracle import oracle_amvp, oracle_merge
    w, h = 200, 136
    m, jobs = _units("cand"), G2["cand_jobs"]
    assert np.array_equal(oracle_amvp(w, h, m, jobs), G2["cand_amvp"])
    assert np.array_equal(oracle_merge(w, h, m, jobs, 5), G2["cand_merge5"])
    assert np.array_equal(oracle_merge(w, h, m, jobs, 2), G2["cand_merge2"])
    assert (G2["cand_amvp"] != 0).any() and (G2["cand_merge5"][:, 3] != 0).any()


def test_boundary_strengths_of_a_b_picture():
    from _oracle import oracle_deblock_strengths_b
    w, h = 192, 136
    mot = [G2["bsb_" + k] for k in ("ref0", "mv0", "ref1", "mv1", "pic_l0", "pic_l1")]
    gv, gh = oracle_deblock_strengths_b(w, h, _units("bsb"), *mot)
    assert np.array_equal(gv[:h // 4, :w // 4], G2["bsb_ver"][:h // 4, :w // 4]) and np.array_equal(gh[:h // 4, :w // 4], G2["bsb_hor"][:h // 4, :w // 4])
    assert all((G2["bsb_ver"][:h // 4, 2:w // 4:2] == v).sum() > 20 for v in (0, 1, 2))


def test_sao_offset_pass():
    from _oracle import oracle_sao_apply
    w, h = 200, 136
    got = oracle_sao_apply(_planes(G2["saoa_in"], w, h), w, h, G2["saoa_types"], G2["saoa_offs"])
    exp = _planes(G2["saoa_out"], w, h)
    for c in range(3):
        assert np.array_equal(got[c], exp[c]), c
    assert (exp[0] != _planes(G2["saoa_in"], w, h)[0]).sum() > 500


def test_intra_picture_of_the_reference_encoder():
    """the I picture the reference's encoder made of a 192x136 source: from its decisions per transform unit the restatement rebuilds its
    unfiltered reconstruction sample for sample and its levels"""
    from _intra import coeff_wnd_of, oracle_intra_recon
    w, h, qp, sh, seed = (int(v) for v in G2["ipic_cfg"])
    tus = G2["ipic_tus"]
    rec, coeff, _ = oracle_intra_recon(_planes(G2["ipic_src"], w, h), w, h, tus, is_islice=1, sign_hiding=sh)
    exp = _planes(G2["ipic_recon"], w, h)
    for c in range(3):
        assert np.array_equal(rec[c], exp[c]), (c, np.argwhere(rec[c] != exp[c])[:3].tolist())
    covered = coeff_wnd_of(tus, np.ones_like(coeff), w, h) != 0
    assert np.array_equal(coeff_wnd_of(tus, coeff, w, h)[covered], G2["ipic_coeff"][covered])
    assert set(int(s) for s in tus[tus[:, 0] == 0][:, 3]) >= {4, 8, 16}
