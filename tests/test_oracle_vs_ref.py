"""The oracle against the compiled reference run live (oracle/_ref, built from /root/reference by oracle/Makefile):
wider randomised coverage than the committed golden vectors.  Skipped where oracle/_ref does not exist."""
import ctypes as C

import numpy as np
import pytest

from _oracle import (OrcMeIn, OrcMeOut, OrcMv, aligned_i16, have_ref, i32p, make_frame_pair, oracle, ptr, ref, refdrv)

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def test_reference_selects_sse42_table():
    _, D = ref()
    assert D.refdrv_sse_selected(refdrv()) == 1


def test_sad_ssd_sse_lane_arithmetic_on_the_full_int16_range():
    """the encoder also calls sad / ssd16b on operands that are not 8-bit video (the wrapped 16-bit predictions of its 64x64 intra mode
    search, hmr_motion_intra.c:1130): there the SSE4.2 lanes -- wrapping differences, wrapping or saturating 16-bit accumulation -- decide
    the value.  orc_sad_sse / orc_ssd16b_sse restate exactly that; on video-range data they equal the plain sums."""
    O = oracle(); R, _ = ref()
    O.orc_sad_sse.restype = C.c_uint32; O.orc_ssd16b_sse.restype = C.c_uint32
    rng = np.random.default_rng(321)
    differs = 0
    for it in range(600):
        n = int(rng.choice([4, 8, 16, 32, 64]))
        a = aligned_i16(64 * 64); b = aligned_i16(80 * 80)
        kind = it % 4
        if kind == 0:        # anything an int16 can hold, extremes included
            a[:] = rng.integers(-32768, 32768, a.size); b[:] = rng.integers(-32768, 32768, b.size)
            a[::97] = -32768; b[::89] = 32767
        elif kind == 1:      # video against a wrapped prediction (what the 64x64 intra search produces)
            a[:] = rng.integers(0, 256, a.size); b[:] = rng.integers(-30000, 30000, b.size)
        elif kind == 2:      # moderate overflow: lanes wrap / saturate only partly
            a[:] = rng.integers(0, 256, a.size); b[:] = rng.integers(-3000, 3000, b.size)
        else:                # video
            a[:] = rng.integers(0, 256, a.size); b[:] = rng.integers(0, 256, b.size)
        off = int(rng.integers(0, 8))
        sse = R.sse_aligned_sad(ptr(a), 64, ptr(b, off), 80, n)
        assert sse == O.orc_sad_sse(ptr(a), 64, ptr(b, off), 80, n), (kind, n)
        assert R.sse_aligned_ssd16b(ptr(a), 64, ptr(b, off), 80, n) == O.orc_ssd16b_sse(ptr(a), 64, ptr(b, off), 80, n), (kind, n)
        if kind == 3:
            assert sse == O.orc_sad(ptr(a), 64, ptr(b, off), 80, n)
        else:
            differs += sse != O.orc_sad(ptr(a), 64, ptr(b, off), 80, n)
    assert differs > 100          # the lane arithmetic really matters off the video range


def test_pixel_interp_transform_random():
    O = oracle(); R, _ = ref()
    rng = np.random.default_rng(100)
    for it in range(200):
        n = int(rng.choice([4, 8, 16, 32, 64]))
        a = aligned_i16(64 * 64); b = aligned_i16(80 * 80)
        a[:] = rng.integers(0, 256, a.size); b[:] = rng.integers(0, 256, b.size)
        off = int(rng.integers(0, 8))
        assert R.sse_aligned_sad(ptr(a), 64, ptr(b, off), 80, n) == O.orc_sad(ptr(a), 64, ptr(b, off), 80, n)
        assert R.sse_aligned_ssd16b(ptr(a), 64, ptr(b, off), 80, n) == O.orc_ssd16b(ptr(a), 64, ptr(b, off), 80, n)
        # the plain-C twins agree too on 8-bit data
        assert R.sad(ptr(a), 64, ptr(b, off), 80, n) == O.orc_sad(ptr(a), 64, ptr(b, off), 80, n)
    for it in range(800):
        w = int(rng.choice([4, 8, 16, 32, 64])); h = int(rng.choice([4, 8, 9, 16, 17, 32, 64, 72]))
        first, last = [(1, 0), (0, 1), (1, 1), (0, 0)][it % 4]
        vert = int(rng.integers(0, 2))
        src = aligned_i16(96 * 96)
        src[:] = rng.integers(0, 256, src.size) if first else rng.integers(-8192, 8129, src.size)
        for chroma in (0, 1):
            if chroma and w == 64:
                continue
            frac = int(rng.integers(0, 8 if chroma else 4))
            d1 = aligned_i16(80 * 80); d2 = aligned_i16(80 * 80)
            f = (R.sse_interpolate_chroma, O.orc_interpolate_chroma) if chroma else (R.sse_interpolate_luma, O.orc_interpolate_luma)
            f[0](ptr(src, 96 * 8 + 8), 96, ptr(d1), 80, frac, w, h, vert, first, last)
            f[1](ptr(src, 96 * 8 + 8), 96, ptr(d2), 80, frac, w, h, vert, first, last)
            assert np.array_equal(d1.reshape(80, 80)[:h, :w], d2.reshape(80, 80)[:h, :w]), (chroma, frac, w, h, vert, first, last)
    for it in range(300):
        n = int(rng.choice([4, 8, 16, 32])); dst = int(n == 4 and rng.integers(0, 2))
        amp = int(rng.choice([8, 40, 255]))
        blk = aligned_i16(64 * 64); blk[:] = rng.integers(-amp, amp + 1, blk.size)
        c1 = aligned_i16(1024); c2 = aligned_i16(1024); aux = aligned_i16(1024)
        lg = n.bit_length() - 1
        R.sse_transform(8, ptr(blk), ptr(c1), 64, n, n, lg, lg, C.c_uint16(0 if dst else 65535), ptr(aux))
        O.orc_transform(8, ptr(blk), 64, ptr(c2), n, dst)
        assert np.array_equal(c1[:n * n], c2[:n * n])
        co = aligned_i16(1024); co[:n * n] = rng.integers(-32768, 32768, n * n)
        b1 = aligned_i16(64 * 64); b2 = aligned_i16(64 * 64)
        R.sse_itransform(8, ptr(b1), ptr(co), 64, n, n, C.c_uint(0 if dst else 65535), ptr(aux))
        O.orc_itransform(8, ptr(b2), 64, ptr(co), n, dst)
        assert np.array_equal(b1.reshape(64, 64)[:n, :n], b2.reshape(64, 64)[:n, :n])


def test_quant_random_and_c_vs_sse_difference():
    O = oracle(); _, D = ref()
    h = refdrv()
    rng = np.random.default_rng(101)
    differs = 0
    for it in range(1000):
        lg = int(rng.choice([2, 3, 4, 5])); n = 1 << lg
        comp = int(rng.integers(0, 3)) if lg < 5 else 0
        is_intra, isl, sh = (int(v) for v in rng.integers(0, 2, 3))
        qp = int(rng.integers(0, 52)); per, rem = qp // 6, qp % 6
        scan = int(rng.choice([1, 2, 3])) if lg <= 3 else 3
        src = aligned_i16(1024)
        src[:n * n] = rng.laplace(0, [60, 400, 4000][it % 3], n * n).clip(-32768, 32767).astype(np.int16)
        d1 = aligned_i16(1024); d2 = aligned_i16(1024); d3 = aligned_i16(1024); u1 = aligned_i16(1024); u2 = aligned_i16(1024)
        s1 = C.c_int(0); s2 = C.c_int(0); s3 = C.c_int(0)
        D.refdrv_quant(h, ptr(src), ptr(d1), ptr(u1), scan, lg, comp, is_intra, isl, sh, per, rem, C.byref(s1))
        O.orc_quant(O.tables, ptr(src), ptr(d2), ptr(u2), scan, lg, comp, is_intra, isl, sh, per, rem, C.byref(s2))
        assert s1.value == s2.value and np.array_equal(d1[:n * n], d2[:n * n]) and np.array_equal(u1[:n * n], u2[:n * n])
        if not isl:       # the plain-C quant always rounds with 171: it is NOT the target on P slices
            D.refdrv_quant_plainc(h, ptr(src), ptr(d3), scan, lg, comp, is_intra, isl, sh, per, rem, C.byref(s3))
            differs += int(not np.array_equal(d1[:n * n], d3[:n * n]))
        q1 = aligned_i16(1024); q2 = aligned_i16(1024)
        D.refdrv_inv_quant(h, ptr(d1), ptr(q1), lg, comp, is_intra, per, rem)
        O.orc_inv_quant(O.tables, ptr(d1), ptr(q2), lg, comp, is_intra, per, rem)
        assert np.array_equal(q1[:n * n], q2[:n * n])
    assert differs > 50


def test_motion_estimation_random():
    O = oracle(); _, D = ref()
    h = refdrv()
    rng = np.random.default_rng(102)
    W, H, PAD = 416, 240, 80
    S = W + 2 * PAD
    for it in range(300):
        if it % 30 == 0:
            cur, rf = make_frame_pair(rng, W, H, PAD, shift=(int(rng.integers(0, 21)), int(rng.integers(0, 13))), noise=float(rng.choice([0, 1, 3, 8])))
            cur = np.ascontiguousarray(cur); rf = np.ascontiguousarray(rf)
        n = int(rng.choice([8, 16, 32, 64]))
        gx = int(rng.integers(0, (W - n) // n + 1)) * n; gy = int(rng.integers(0, (H - n) // n + 1)) * n
        ob = aligned_i16(64 * 64); ob.reshape(64, 64)[:n, :n] = cur[PAD + gy:PAD + gy + n, PAD + gx:PAD + gx + n]
        amvp = np.array(rng.integers(-40, 41, 4) if it % 3 else [0, 0, 0, 0], np.int32)
        ns = int(rng.integers(0, 4)); st = np.array(rng.integers(-60, 61, 6), np.int32)
        qp = int(rng.integers(20, 45)); avg = float(rng.choice([0., 100., 700., 2500., 5000.])); action = int(rng.choice([7, 7, 7, 3, 1]))
        out = np.zeros(4, np.int32)
        off = (PAD + gy) * S + PAD + gx
        r = D.refdrv_motion_estimation(h, ptr(ob), 64, ptr(rf.reshape(-1), off), S, gx, gy, n, W, H, 2, amvp.ctypes.data_as(i32p), ns,
                                       st.ctypes.data_as(i32p), qp, avg, action, out.ctypes.data_as(i32p))
        mi = OrcMeIn()
        mi.orig = ptr(ob); mi.orig_stride = 64; mi.ref = ptr(rf.reshape(-1), off); mi.ref_stride = S
        mi.gx, mi.gy, mi.size, mi.frame_w, mi.frame_h, mi.range_x, mi.range_y = gx, gy, n, W, H, 128, 64
        mi.n_amvp = 2
        for i in range(2):
            mi.amvp[i] = OrcMv(int(amvp[2 * i]), int(amvp[2 * i + 1]))
        mi.n_start = ns
        for i in range(3):
            mi.start[i] = OrcMv(int(st[2 * i]), int(st[2 * i + 1]))
        mi.qp, mi.avg_dist, mi.action = qp, avg, action
        mo = OrcMeOut()
        O.orc_motion_estimation(C.byref(mi), C.byref(mo))
        assert (r, out[0], out[1], out[2], out[3]) == (mo.sad, mo.mv.x, mo.mv.y, mo.subpix.x, mo.subpix.y), (it, n, gx, gy, action)


def test_lockstep_encode_is_deterministic():
    """the whole-encode golden of SURVEY.md 8c: two lock-step runs give identical bytes and reconstructions"""
    from homerhevc_b200 import synth
    _, D = ref()
    w, h, nf = 192, 128, 3
    clip = synth.make_clip(w, h, nf, seed=11)
    yuv = np.concatenate([np.concatenate([p.reshape(-1) for p in f]) for f in clip])
    outs = []
    for rep in range(2):
        bs = np.zeros(1 << 20, np.uint8); rec = np.zeros(yuv.size, np.uint8); secs = C.c_double(0)
        n = D.refdrv_encode_lockstep(w, h, nf, yuv.ctypes.data_as(C.POINTER(C.c_uint8)), 32, 1, 0, -1, bs.ctypes.data_as(C.POINTER(C.c_uint8)),
                                     bs.size, rec.ctypes.data_as(C.POINTER(C.c_uint8)), None, None, C.byref(secs))
        assert n > 0
        outs.append((bytes(bs[:n]), rec.copy()))
    assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1])


def test_intra_tq_chain_against_reference_calls():
    """orc_encode_intra_tu against the reference's own functions called in the order of encode_intra_cu
    (hmr_motion_intra.c:1037-1069) / the chroma loop (hmr_motion_intra_chroma.c:340-365)"""
    from _oracle import OrcTuOut
    O = oracle(); R, D = ref()
    h = refdrv()
    rng = np.random.default_rng(103)
    coded = 0
    for it in range(400):
        comp = int(rng.integers(0, 3))
        n = int(rng.choice([4, 8, 16, 32] if comp == 0 else [4, 8, 16])); lg = n.bit_length() - 1
        qp = int(rng.integers(15, 45)); isl, sh = (int(v) for v in rng.integers(0, 2, 2))
        scan = int(rng.choice([1, 2, 3])) if n <= 8 else 3
        weight = 1.0 if comp == 0 else float(rng.choice([1.0, 1.2599210498948732, 0.7937005259840998]))
        orig = aligned_i16(n * n); pred = aligned_i16(n * n)
        orig[:] = rng.integers(0, 256, n * n)
        pred[:] = np.clip(orig + np.rint(rng.normal(0, float(rng.choice([1, 4, 15, 50])), n * n)), 0, 255)
        # ---- reference call sequence
        res = aligned_i16(n * n); tc = aligned_i16(1024); lev = aligned_i16(1024); dq = aligned_i16(1024); aux = aligned_i16(1024)
        dec = aligned_i16(n * n); zeros = aligned_i16(64); s = C.c_int(0)
        mode = 10 if comp == 0 else 65535                      # luma hands the intra mode (!= REG_DCT), chroma REG_DCT
        R.sse_aligned_predict(ptr(orig), n, ptr(pred), n, ptr(res), n, n)
        R.sse_transform(8, ptr(res), ptr(tc), n, n, n, lg, lg, C.c_uint16(mode), ptr(aux))
        D.refdrv_quant(h, ptr(tc), ptr(lev), None, scan, lg, comp, 1, isl, sh, qp // 6, qp % 6, C.byref(s))
        if s.value:
            D.refdrv_inv_quant(h, ptr(lev), ptr(dq), lg, comp, 1, qp // 6, qp % 6)
            R.sse_itransform(8, ptr(res), ptr(dq), n, n, n, C.c_uint(mode), ptr(aux))
            R.sse_aligned_reconst(ptr(pred), n, ptr(res), n, ptr(dec), n, n)
        else:
            R.sse_aligned_reconst(ptr(pred), n, ptr(zeros), 0, ptr(dec), n, n)
        ssd = R.sse_aligned_ssd16b(ptr(orig), n, ptr(dec), n, n)
        if comp:
            ssd = int(weight * ssd)
        # ---- oracle
        co2 = aligned_i16(1024); de2 = aligned_i16(n * n); to = OrcTuOut()
        O.orc_encode_intra_tu(O.tables, ptr(orig), n, ptr(pred), n, ptr(co2), ptr(de2), n, n, comp, qp, scan, isl, sh, weight, C.byref(to))
        assert (to.sum, to.ssd) == (s.value, ssd), (comp, n, qp, scan)
        assert np.array_equal(co2[:n * n], lev[:n * n]) and np.array_equal(de2, dec), (comp, n, qp, scan)
        coded += s.value > 0
    assert coded > 50


def _random_adi(rng, n, kind):
    size = 4 * n + 1
    if kind == 0:
        return rng.integers(0, 256, size)
    if kind == 1:
        return np.clip(128 + np.cumsum(rng.integers(-3, 4, size)), 0, 255)
    return np.clip(np.linspace(rng.integers(0, 256), rng.integers(0, 256), size) + rng.integers(-1, 2, size), 0, 255)   # smooth: strong filter


def test_intra_prediction_against_reference():
    """reference-sample smoothing and every intra mode (planar, DC, 33 angular, luma and chroma variants), sizes 4..32,
    through the reference's own table members (create_intra_planar_prediction / create_intra_angular_prediction, adi_filter)"""
    O = oracle(); _, D = ref()
    h = refdrv()
    rng = np.random.default_rng(104)
    strong = 0
    for it in range(240):
        n = int(rng.choice([4, 8, 16, 32]))
        adi = aligned_i16(4 * n + 1 + 8); adi[:4 * n + 1] = _random_adi(rng, n, it % 3)
        f1 = aligned_i16(4 * n + 1 + 8); f2 = aligned_i16(4 * n + 1 + 8)
        D.refdrv_adi_filter(h, ptr(adi), ptr(f1), n)
        O.orc_adi_filter(ptr(adi), ptr(f2), n, 1)
        assert np.array_equal(f1[:4 * n + 1], f2[:4 * n + 1]), ("filter", n, it % 3)
        if n == 32 and it % 3 == 2:
            plain = (adi[:-10].astype(np.int32) + 2 * adi[1:-9] + adi[2:-8] + 2) >> 2
            strong += int(not np.array_equal(plain[:4 * n - 1], f1[1:4 * n]))
        for mode in range(35):
            for is_luma in (1, 0):
                src = f1 if (is_luma and it % 2) else adi
                p1 = aligned_i16(n * n); p2 = aligned_i16(n * n)
                D.refdrv_intra_predict(h, ptr(src), n, mode, is_luma, ptr(p1))
                O.orc_intra_predict(ptr(src), n, mode, is_luma, ptr(p2), n)
                assert np.array_equal(p1, p2), (n, mode, is_luma, it % 3)
    assert strong > 0          # the strong bilinear smoothing branch ran


def test_weighted_average_against_reference():
    """bi-prediction average of two 14-bit predictions (the table's weighted_average_motion), all block shapes, extreme inputs included"""
    O = oracle(); _, D = ref()
    h = refdrv()
    rng = np.random.default_rng(105)
    for it in range(200):
        w = int(rng.choice([4, 8, 16, 32, 64])); hh = int(rng.choice([4, 8, 16, 32, 64]))
        lo, hi = (-8192, 8192) if it % 3 else (-14312, 14248)          # the whole range a first filter stage can produce
        a = aligned_i16(64 * 64); b = aligned_i16(64 * 64)
        a[:] = rng.integers(lo, hi, a.size); b[:] = rng.integers(lo, hi, b.size)
        d1 = aligned_i16(64 * 64); d2 = aligned_i16(64 * 64)
        D.refdrv_weighted_average(h, ptr(a), 64, ptr(b), 64, ptr(d1), 64, hh, w)
        O.orc_weighted_average(ptr(a), 64, ptr(b), 64, ptr(d2), 64, hh, w)
        assert np.array_equal(d1.reshape(64, 64)[:hh, :w], d2.reshape(64, 64)[:hh, :w]), (w, hh, it)


def test_bi_prediction_mc_against_reference():
    """motion compensation with is_bi_predict = 1 (14-bit output of one list), luma and chroma, all fraction combinations"""
    from _oracle import OrcMv, make_frame_pair
    O = oracle(); _, D = ref()
    h = refdrv()
    rng = np.random.default_rng(106)
    plane = aligned_i16(200 * 200); plane[:] = rng.integers(0, 256, plane.size)
    for it in range(300):
        chroma = it % 2
        n = int(rng.choice([4, 8, 16, 32] if chroma else [8, 16, 32, 64]))
        x, y = int(rng.integers(70, 110 - 0)), int(rng.integers(70, 110))
        lim = 7 if chroma else 3
        mvx, mvy = int(rng.integers(-40, 41)), int(rng.integers(-40, 41))
        if it % 7 == 0: mvx &= ~lim
        if it % 5 == 0: mvy &= ~lim
        p1 = aligned_i16(64 * 64); p2 = aligned_i16(64 * 64)
        if x + n + 30 > 200 or y + n + 30 > 200:
            continue
        (D.refdrv_mc_chroma_bi if chroma else D.refdrv_mc_luma_bi)(h, ptr(plane, y * 200 + x), 200, ptr(p1), 64, n, mvx, mvy)
        (O.orc_mc_chroma_ex if chroma else O.orc_mc_luma_ex)(ptr(plane, y * 200 + x), 200, ptr(p2), 64, n, OrcMv(mvx, mvy), 1)
        assert np.array_equal(p1.reshape(64, 64)[:n, :n], p2.reshape(64, 64)[:n, :n]), (chroma, n, mvx, mvy)
        assert p2.reshape(64, 64)[:n, :n].min() < 0 or p2.reshape(64, 64)[:n, :n].max() > 255      # 14-bit values, not samples


def test_sao_statistics_against_reference():
    """SAO statistics (edge classes in four directions + 32 bands) of every CTU and component through the table's get_sao_stats
    versus the restatement: whole CTUs, partial CTUs at the right / bottom edge, a one-CTU-wide picture"""
    from _oracle import make_frame_pair, oracle_sao_stats, ref_sao_stats
    rng = np.random.default_rng(107)
    for (w, h) in ((192, 128), (200, 136), (64, 200), (328, 72)):
        base = [np.clip(rng.normal(128, 40, (hh, ww)), 0, 255) for (ww, hh) in ((w, h), (w // 2, h // 2), (w // 2, h // 2))]
        org = [b.astype(np.uint8) for b in base]
        rec = [np.clip(b + rng.normal(0, 3, b.shape), 0, 255).astype(np.uint8) for b in base]
        a = oracle_sao_stats(rec, org, w, h); b = ref_sao_stats(rec, org, w, h)
        for f in ("eo_diff", "eo_count", "bo_diff", "bo_count"):
            assert np.array_equal(a[f], b[f]), (w, h, f, np.argwhere(a[f] != b[f])[:3])
        assert a["eo_count"].sum() > 0 and (a["eo_count"][:, :, :, 0] > 0).any() and (a["eo_count"][:, :, :, 4] > 0).any()


def test_sao_offset_pass_against_reference():
    """the SAO offset pass (sao_offset_ctu / offset_block) with random per-CTU types and offsets, whole and partial CTUs"""
    from _oracle import oracle_sao_apply, random_sao_params, ref_sao_apply
    rng = np.random.default_rng(108)
    for (w, h) in ((192, 128), (200, 136), (64, 200), (328, 72)):
        src = [np.clip(rng.normal(128, 50, (hh, ww)), 0, 255).astype(np.uint8) for (ww, hh) in ((w, h), (w // 2, h // 2), (w // 2, h // 2))]
        types, offs = random_sao_params(rng, w, h)
        a = oracle_sao_apply(src, w, h, types, offs); b = ref_sao_apply(src, w, h, types, offs)
        for c in range(3):
            assert np.array_equal(a[c], b[c]), (w, h, c, np.argwhere(a[c] != b[c])[:4])
        assert any((a[c] != src[c]).any() for c in range(3))


def test_sao_offset_derivation_against_reference():
    """the library's host half of the SAO decision (hb_sao_derive_offsets: initial offsets, sign rules, the rate-distortion walk,
    the band position, the distortion estimate) against the reference's sao_derive_offsets / sao_get_distortion, every type,
    luma and chroma lambdas from tiny to huge"""
    from _oracle import random_sao_stats, ref_sao_derive
    from homerhevc_b200.lib import sao_derive_offsets
    rng = np.random.default_rng(111)
    st = random_sao_stats(rng, 60)
    n_nonzero = 0
    for i, rec in enumerate(st):
        lam = float(rng.choice([0.5, 7.0, 33.3, 120.0, 900.0, 1e4]))
        for t in range(5):
            off, band, dist = sao_derive_offsets(rec, t, lam)
            roff, rband, rdist = ref_sao_derive(rec, i % 3, t, lam)
            assert np.array_equal(off.astype(np.int32), roff) and dist == rdist, (i, t, lam, off, roff, dist, rdist)
            if t == 4:
                assert band == rband, (i, lam, band, rband)
            n_nonzero += int((off != 0).any())
    assert n_nonzero > 100


def test_deblocking_pixel_stage_against_reference():
    """the reference's own deblocking (its per-CTU function, vertical edges of the whole picture first) on random CU/TU trees,
    modes, cbf, QPs and vectors; the restatement gets the boundary strengths the reference derived and must produce the same
    picture: luma normal / strong filters, chroma filter, QP averaging across CU and CTU borders, partial CTUs"""
    from _oracle import oracle_deblock, oracle_deblock_strengths, random_deblock_case, ref_deblock
    rng = np.random.default_rng(109)
    strong = weak = 0
    for (w, h) in ((192, 136), (128, 128), (200, 72), (64, 200)):
        for rep in range(3):
            m, planes = random_deblock_case(rng, w, h)
            exp, bsv, bsh, offs = ref_deblock(planes, w, h, m)
            assert set(np.unique(bsv)) <= {0, 1, 2} and (bsv == 2).any() and (bsv == 1).any() and (bsh == 1).any()
            obv, obh = oracle_deblock_strengths(m, w, h)
            assert np.array_equal(obv[:h // 4, :w // 4], bsv[:h // 4, :w // 4]) and np.array_equal(obh[:h // 4, :w // 4], bsh[:h // 4, :w // 4]), (w, h, rep)
            got = oracle_deblock(planes, w, h, bsv, bsh, m["qp"], offs)
            for c in range(3):
                assert np.array_equal(got[c], exp[c]), (w, h, rep, c, np.argwhere(got[c] != exp[c])[:4])
            assert (exp[0] != planes[0]).sum() > 500 and (exp[1] != planes[1]).sum() > 20


def test_amvp_candidates_against_reference():
    """get_amvp_candidates (hmr_motion_inter.c:2342) of every 2Nx2N PU of every size on random CU trees with intra / inter units and
    random vectors: left-bottom / left, top-right / top / top-left look-ups across CTU borders, z-order availability, the
    quadtree's neighbour flags, partial CTUs on both picture edges, duplicate removal and zero fill"""
    from _oracle import amvp_jobs, oracle_amvp, random_deblock_case, ref_amvp
    rng = np.random.default_rng(2342)
    seen = set()
    for (w, h) in ((192, 136), (128, 128), (200, 72), (72, 200), (320, 192)):
        for rep in range(3):
            m, _ = random_deblock_case(rng, w, h)
            m["mv"] = rng.integers(-40, 41, m["mv"].shape).astype(np.int16) if rep == 2 else m["mv"]     # per-unit vectors: nothing may depend on CU-constant fields
            jobs = amvp_jobs(w, h)
            exp, got = ref_amvp(w, h, m, jobs), oracle_amvp(w, h, m, jobs)
            bad = np.argwhere((exp != got).any(1))
            assert not len(bad), (w, h, rep, jobs[bad[0, 0]], exp[bad[0, 0]], got[bad[0, 0]])
            for e in exp:
                seen.add((bool(e[0] or e[1]), bool(e[2] or e[3])))
    assert seen == {(False, False), (True, False), (True, True)} or seen == {(False, False), (True, False), (True, True), (False, True)}


def test_merge_candidates_against_reference():
    """get_merge_mvp_candidates (hmr_motion_inter.c:1937) for every 2Nx2N PU: order A1, B1, B0, A0, B2, the pairwise pruning, the
    fifth neighbour only while fewer than four were found, list lengths 1..5 and the zero fill"""
    from _oracle import amvp_jobs, oracle_merge, random_deblock_case, ref_merge
    rng = np.random.default_rng(1937)
    full = 0
    for (w, h) in ((192, 136), (200, 72), (72, 200), (320, 192)):
        for rep in range(3):
            m, _ = random_deblock_case(rng, w, h)
            if rep == 2:
                m["mv"] = rng.integers(-2, 3, m["mv"].shape).astype(np.int16)       # few distinct vectors: the pruning has work to do
            jobs = amvp_jobs(w, h)
            for mx in (5, 3, 1, 2, 4):
                exp, got = ref_merge(w, h, m, jobs, mx), oracle_merge(w, h, m, jobs, mx)
                bad = np.argwhere((exp != got).any((1, 2)))
                assert not len(bad), (w, h, rep, mx, jobs[bad[0, 0]], exp[bad[0, 0]].tolist(), got[bad[0, 0]].tolist())
                if mx == 5:
                    full += int((exp[:, 3] != 0).any(1).sum())
    assert full > 50


def test_boundary_strengths_of_b_pictures_against_reference():
    """the two-list branch of get_boundary_strength_single (hmr_deblocking_filter.c:173-229): uni- and bi-predicted units, lists that name
    the same pictures in opposite order (so equal motion can sit in swapped lists), same-picture pairs, coded / intra neighbours"""
    from _oracle import oracle_deblock_strengths_b, random_b_motion, random_deblock_case, ref_deblock_strengths_b
    rng = np.random.default_rng(173)
    seen = {0: 0, 1: 0, 2: 0}
    for (w, h) in ((192, 136), (128, 128), (200, 72)):
        for rep in range(4):
            m, _ = random_deblock_case(rng, w, h)
            m["cbf"] = (m["cbf"] * (rng.random(m["cbf"].shape) < 0.25)).astype(np.uint8)          # mostly uncoded: the motion rule decides
            n1 = 1 if rep == 3 else 2
            mot = random_b_motion(rng, m, 2, n1)
            ev, eh = ref_deblock_strengths_b(w, h, m, *mot)
            gv, gh = oracle_deblock_strengths_b(w, h, m, *mot)
            assert np.array_equal(gv[:h // 4, :w // 4], ev[:h // 4, :w // 4]) and np.array_equal(gh[:h // 4, :w // 4], eh[:h // 4, :w // 4]), (w, h, rep)
            for v in (0, 1, 2):
                seen[v] += int((ev[:h // 4, 2:w // 4:2] == v).sum())
    assert min(seen.values()) > 50, seen


def test_intra_picture_reconstruction_against_the_reference_encoder():
    """a whole I picture rebuilt by the restatement from the decisions the reference's own encoder made (captured CTU by CTU through
    oracle/ref_hooks.c: CU / TU trees, luma and chroma modes incl. the derived mode, QPs) == the encoder's reconstruction BEFORE its in-loop
    filters, sample for sample in all three planes, and the same levels where the encoder's coeff_wnd holds a unit: reference samples from
    the reconstructed neighbours with the encoder's own availability and padding rules (fill_reference_samples hmr_motion_intra.c:246),
    smoothing rule, 35 predictors, DST / DCT, mode-dependent scans, sign hiding; partial CTUs on both picture edges"""
    from homerhevc_b200 import synth
    from _intra import capture_intra_picture, coeff_wnd_of, intra_tus, oracle_intra_recon
    seen_sizes, seen_modes, n_dm = set(), set(), 0
    for (w, h, qp, sh, seed) in ((192, 136, 32, 1, 21), (200, 72, 24, 1, 5), (328, 200, 38, 0, 9), (320, 192, 30, 1, 3)):
        clip = synth.make_clip(w, h, 1, seed=seed)
        a = capture_intra_picture(w, h, qp, sh, seed)
        assert a["slice_type"] == 2 and (a["pred_mode"][:h // 4, :w // 4] == 1).all()          # I_SLICE, every unit intra
        tus = intra_tus(a, w, h)
        rec, coeff, res = oracle_intra_recon(clip[0], w, h, tus, is_islice=1, sign_hiding=sh)
        for c in range(3):
            bad = np.argwhere(rec[c] != a["recon"][c])
            assert not len(bad), (w, h, qp, c, len(bad), bad[:3].tolist())
        mine = coeff_wnd_of(tus, coeff, w, h)
        covered = coeff_wnd_of(tus, np.ones_like(coeff), w, h) != 0
        assert np.array_equal(mine[covered], a["coeff"][covered]), (w, h, qp)
        seen_sizes |= set(int(s) for s in tus[tus[:, 0] == 0][:, 3]); seen_modes |= set(int(m) for m in tus[:, 4])
        n_dm += int((a["mode_c"][:h // 4, :w // 4] == 36).sum())
    assert seen_sizes == {4, 8, 16, 32} and len(seen_modes) > 15 and n_dm > 0, (seen_sizes, seen_modes, n_dm)
