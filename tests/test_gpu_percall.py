"""Per-call parity: every member of the low-level function table that the library
implements, called through the C ABI with host buffers exactly like the reference's sse_* functions, against the
oracle on the same seeded inputs.  Bit-exact."""
import ctypes as C

import numpy as np
import pytest

import homerhevc_b200 as hb
from homerhevc_b200.lib import UNIT_INFO_DT
from _oracle import aligned_i16, oracle, ptr

pytestmark = pytest.mark.gpu
ll = hb.lowlevel


def test_sad_ssd(ctx):
    O = oracle()
    rng = np.random.default_rng(10)
    for it in range(120):
        n = int(rng.choice([4, 8, 16, 32, 64]))
        a = aligned_i16(64 * 64); b = aligned_i16(160 * 100)
        if it % 3 == 2:      # bi-pred style range (2*orig - pred), still exact
            a[:] = rng.integers(-255, 511, a.size); b[:] = rng.integers(0, 256, b.size)
        else:
            a[:] = rng.integers(0, 256, a.size); b[:] = rng.integers(0, 256, b.size)
        off = int(rng.integers(0, 900))
        assert ll.sad(a, 64, b, 160, n, pred_off=off) == O.orc_sad(ptr(a), 64, ptr(b, off), 160, n)
        assert ll.ssd16b(a, 64, b, 160, n, pred_off=off) == O.orc_ssd16b(ptr(a), 64, ptr(b, off), 160, n)
    # operands that are not video: the reference's 64x64 intra mode search hands sad() wrapped 16-bit predictions
    # (hmr_motion_intra.c:1130), where the SSE4.2 lane arithmetic decides the value -- the drop-ins reproduce it exactly
    O.orc_sad_sse.restype = C.c_uint32; O.orc_ssd16b_sse.restype = C.c_uint32
    for it in range(100):
        n = int(rng.choice([4, 8, 16, 32, 64]))
        a = aligned_i16(64 * 64); b = aligned_i16(160 * 100)
        if it % 2:
            a[:] = rng.integers(-32768, 32768, a.size); b[:] = rng.integers(-32768, 32768, b.size)
            a[::97] = -32768; b[::89] = 32767
        else:
            a[:] = rng.integers(0, 256, a.size); b[:] = rng.integers(-30000, 30000, b.size)
        off = int(rng.integers(0, 900))
        assert ll.sad(a, 64, b, 160, n, pred_off=off) == O.orc_sad_sse(ptr(a), 64, ptr(b, off), 160, n), (it, n)
        assert ll.ssd16b(a, 64, b, 160, n, pred_off=off) == O.orc_ssd16b_sse(ptr(a), 64, ptr(b, off), 160, n), (it, n)
    # pred_stride = 0 against a zero row (hmr_motion_inter.c:94)
    z = np.zeros(256, np.int16)
    a = aligned_i16(64 * 64); a[:] = rng.integers(-255, 256, a.size)
    for n in (4, 8, 16, 32):
        assert ll.ssd16b(a, 64, z, 0, n) == O.orc_ssd16b(ptr(a), 64, ptr(z), 0, n)


def test_predict_reconst(ctx):
    O = oracle()
    rng = np.random.default_rng(11)
    for it in range(60):
        n = int(rng.choice([4, 8, 16, 32, 64]))
        o = aligned_i16(64 * 64); p = aligned_i16(64 * 64)
        o[:] = rng.integers(0, 256, o.size); p[:] = rng.integers(0, 256, p.size)
        r1 = np.zeros(64 * 64, np.int16); r2 = np.zeros(64 * 64, np.int16)
        ll.predict(o, 64, p, 64, r1, 64, n)
        O.orc_predict(ptr(o), 64, ptr(p), 64, ptr(r2), 64, n)
        assert np.array_equal(r1, r2)
        res = aligned_i16(64 * 64); res[:] = rng.integers(-600, 601, res.size)
        d1 = np.zeros(144 * 64, np.int16); d2 = np.zeros(144 * 64, np.int16)
        ll.reconst(p, 64, res, 64, d1, 144, n)
        O.orc_reconst(ptr(p), 64, ptr(res), 64, ptr(d2), 144, n)
        assert np.array_equal(d1, d2)
    z = np.zeros(64, np.int16)       # residual_stride = 0 on an all-zero buffer
    d1 = np.zeros(64 * 64, np.int16); d2 = np.zeros(64 * 64, np.int16)
    ll.reconst(p, 64, z, 0, d1, 64, 32)
    O.orc_reconst(ptr(p), 64, ptr(z), 0, ptr(d2), 64, 32)
    assert np.array_equal(d1, d2)


@pytest.mark.parametrize("chroma", [0, 1])
def test_interpolate(ctx, chroma):
    O = oracle()
    rng = np.random.default_rng(12 + chroma)
    modes = [(1, 0), (0, 1), (1, 1), (0, 0)]
    for it in range(160):
        w = int(rng.choice([4, 8, 16, 32] if chroma else [4, 8, 9, 16, 17, 32, 33, 64, 65]))
        h = int(rng.choice([4, 8, 9, 16, 24, 32, 40, 64, 72]))
        first, last = modes[it % 4]
        vert = int(rng.integers(0, 2))
        frac = int(rng.integers(0, 8 if chroma else 4))
        src = aligned_i16(96 * 96)
        src[:] = rng.integers(0, 256, src.size) if first else rng.integers(-8192, 8129, src.size)
        d1 = np.zeros(80 * 80, np.int16); d2 = np.zeros(80 * 80, np.int16)
        off = 96 * 8 + 8
        if chroma:
            ll.interpolate_chroma(src, 96, d1, 80, frac, w, h, vert, first, last, ref_off=off)
            O.orc_interpolate_chroma(ptr(src, off), 96, ptr(d2), 80, frac, w, h, vert, first, last)
        else:
            ll.interpolate_luma(src, 96, d1, 80, frac, w, h, vert, first, last, ref_off=off)
            O.orc_interpolate_luma(ptr(src, off), 96, ptr(d2), 80, frac, w, h, vert, first, last)
        assert np.array_equal(d1, d2), (chroma, frac, w, h, vert, first, last)


def test_transform_itransform(ctx):
    O = oracle()
    rng = np.random.default_rng(14)
    for it in range(160):
        n = int(rng.choice([4, 8, 16, 32]))
        dst = int(n == 4 and it % 2)
        amp = int(rng.choice([8, 40, 255]))
        blk = aligned_i16(64 * 64); blk[:] = rng.integers(-amp, amp + 1, blk.size)
        c1 = np.zeros(1024, np.int16); c2 = np.zeros(1024, np.int16)
        ll.transform(8, blk, c1, 64, n, mode=0 if dst else hb.REG_DCT)
        O.orc_transform(8, ptr(blk), 64, ptr(c2), n, dst)
        assert np.array_equal(c1[:n * n], c2[:n * n]), ("fwd", n, dst, amp)
        co = np.zeros(1024, np.int16)
        if it % 3 == 0:
            co[:n * n] = c1[:n * n]
        elif it % 3 == 1:
            co[:n * n] = rng.integers(-32768, 32768, n * n)                    # full int16 range: both stages clip
        else:
            co[:n * n] = rng.integers(-32768, 32768, n * n) * (rng.random(n * n) < 0.1)
        b1 = np.zeros(64 * 64, np.int16); b2 = np.zeros(64 * 64, np.int16)
        ll.itransform(8, b1, co, 64, n, mode=0 if dst else hb.REG_DCT)
        O.orc_itransform(8, ptr(b2), 64, ptr(co), n, dst)
        assert np.array_equal(b1, b2), ("inv", n, dst)


def test_quant_inv_quant(ctx):
    O = oracle()
    rng = np.random.default_rng(15)
    n_sbh_changed = 0
    for it in range(400):
        lg = int(rng.choice([2, 3, 4, 5])); n = 1 << lg
        comp = int(rng.integers(0, 3)) if lg < 5 else 0
        is_intra = int(rng.integers(0, 2)); isl = int(rng.integers(0, 2)); sh = int(rng.integers(0, 2))
        qp = int(rng.integers(0, 52)); per, rem = qp // 6, qp % 6
        scan = int(rng.choice([1, 2, 3])) if lg <= 3 else 3
        src = np.zeros(1024, np.int16)
        kind = it % 4
        if kind == 0:
            src[:n * n] = rng.integers(-32768, 32768, n * n)
        elif kind == 1:
            src[:n * n] = rng.laplace(0, 60, n * n).astype(np.int16)
        elif kind == 2:
            src[:n * n] = (rng.laplace(0, 400, n * n) * (rng.random(n * n) < 0.2)).astype(np.int16)
        else:
            src[:n * n] = rng.laplace(0, 2000, n * n).clip(-32768, 32767).astype(np.int16)
        du1 = np.zeros(1024, np.int16); du2 = np.zeros(1024, np.int16)
        d1 = np.zeros(1024, np.int16); d2 = np.zeros(1024, np.int16)
        env = hb.QuantEnv(isl, sh, 6, 8, du1.ctypes.data_as(C.POINTER(C.c_int16)))
        depth = 6 - lg - (comp != 0)
        s1 = ll.quant(env, src, d1, scan, depth, comp, hb.REG_DCT, is_intra, n, per, rem)
        s2 = C.c_int(0)
        O.orc_quant(O.tables, ptr(src), ptr(d2), ptr(du2), scan, lg, comp, is_intra, isl, sh, per, rem, C.byref(s2))
        assert s1 == s2.value, ("sum", lg, comp, is_intra, isl, sh, qp, kind)
        assert np.array_equal(d1[:n * n], d2[:n * n]), ("levels", lg, comp, is_intra, isl, sh, qp, kind, int((d1 != d2).sum()))
        assert np.array_equal(du1[:n * n], du2[:n * n]), ("deltaU", lg, comp, qp, kind)
        if sh:
            d3 = np.zeros(1024, np.int16); s3 = C.c_int(0)
            O.orc_quant(O.tables, ptr(src), ptr(d3), ptr(du2), scan, lg, comp, is_intra, isl, 0, per, rem, C.byref(s3))
            n_sbh_changed += int(not np.array_equal(d3, d2))
        lev = np.zeros(1024, np.int16)
        lev[:n * n] = d1[:n * n] if it % 2 else rng.integers(-32768, 32768, n * n)
        q1 = np.zeros(1024, np.int16); q2 = np.zeros(1024, np.int16)
        ll.inv_quant(env, lev, q1, depth, comp, is_intra, n, per, rem)
        O.orc_inv_quant(O.tables, ptr(lev), ptr(q2), lg, comp, is_intra, per, rem)
        assert np.array_equal(q1[:n * n], q2[:n * n]), ("dequant", lg, comp, is_intra, qp)
    assert n_sbh_changed > 20          # the sign-hiding branch really ran


def test_weighted_average(ctx):
    O = oracle()
    L = hb.load_library()
    rng = np.random.default_rng(18)
    for it in range(60):
        w = int(rng.choice([4, 8, 16, 32, 64])); h = int(rng.choice([4, 8, 16, 32, 64]))
        a = aligned_i16(80 * 64); b = aligned_i16(72 * 64)
        a[:] = rng.integers(-14312, 14248, a.size); b[:] = rng.integers(-8192, 8192, b.size)
        d1 = np.full(96 * 64, 7, np.int16); d2 = np.full(96 * 64, 7, np.int16)
        L.hb_weighted_average_motion(ptr(a), 80, ptr(b), 72, ptr(d1), 96, h, w, 8)
        O.orc_weighted_average(ptr(a), 80, ptr(b), 72, ptr(d2), 96, h, w)
        assert np.array_equal(d1, d2), (w, h)


def test_function_table(ctx):
    """hb_fill_low_level_funcs writes exactly the 10 members it implements with table prototypes (quant/inv_quant and the intra predictors need the adapter)"""
    L = hb.load_library()
    t = hb.LowLevelFuncs()
    L.hb_fill_low_level_funcs(C.byref(t))
    filled = {n for n, _ in t._fields_ if getattr(t, n)}
    assert filled == {"sad", "ssd16b", "predict", "reconst", "interpolate_luma_m_compensation", "weighted_average_motion",
                      "interpolate_chroma_m_compensation", "interpolate_luma_m_estimation", "transform", "itransform"}
    assert t.sad == C.cast(L.hb_sad, C.c_void_p).value


def test_percall_is_thread_safe(ctx):
    """the table is shared by all encoder threads without locks (hmr_encoder_lib.c:1262): concurrent callers, each with its own
    stream and staging area, must get the same answers as a single thread"""
    import threading
    O = oracle()
    rng = np.random.default_rng(17)
    cases = []
    for _ in range(64):
        n = int(rng.choice([4, 8, 16, 32]))
        a = aligned_i16(64 * 64); b = aligned_i16(64 * 64)
        a[:] = rng.integers(0, 256, a.size); b[:] = rng.integers(0, 256, b.size)
        cases.append((n, a, b, O.orc_sad(ptr(a), 64, ptr(b), 64, n), O.orc_ssd16b(ptr(a), 64, ptr(b), 64, n)))
    errors = []

    def worker(k):
        for rep in range(4):
            for i in range(k, len(cases), 8):
                n, a, b, esad, essd = cases[i]
                if ll.sad(a, 64, b, 64, n) != esad or ll.ssd16b(a, 64, b, 64, n) != essd:
                    errors.append((k, i))
                blk = aligned_i16(64 * 64); blk[:] = a - b
                c1 = np.zeros(1024, np.int16); c2 = np.zeros(1024, np.int16)
                ll.transform(8, blk, c1, 64, n)
                O.orc_transform(8, ptr(blk), 64, ptr(c2), n, 0)
                if not np.array_equal(c1, c2):
                    errors.append((k, i, "tx"))

    ths = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors[:5]


def test_batched_api_rejects_bad_arguments(ctx):
    """status-returning entry points validate their jobs instead of launching on garbage"""
    w, h = 128, 64
    f = hb.Frame(ctx, w, h)
    y = np.zeros((h, w), np.uint8); u = np.zeros((h // 2, w // 2), np.uint8)
    f.upload_u8(y, u, u)
    bad_me = hb.MeJob(); bad_me.x, bad_me.y, bad_me.size, bad_me.qp, bad_me.parent = 96, 0, 64, 30, -1      # sticks out of the frame
    with pytest.raises(hb.HbError):
        ctx.me_search(f, f, [bad_me], 100.0)
    odd = hb.MeJob(); odd.x, odd.y, odd.size, odd.qp, odd.parent = 0, 0, 24, 30, -1                            # not a PU size
    with pytest.raises(hb.HbError):
        ctx.me_search(f, f, [odd], 100.0)
    with pytest.raises(hb.HbError):
        ctx.tq_encode(f, f, f, [hb.TuJob(1, 0, 0, 32, 30)], hb.TqParams(0, 1, 0.0, 1.0))                       # no 32x32 chroma TU
    # the T/Q kernels move a unit's rows with 8- / 16-byte accesses: x must be a multiple of min(size, 16)
    for comp, x, size in ((0, 4, 8), (0, 8, 16), (0, 16 + 8, 32), (1, 4, 8), (2, 8, 16)):
        with pytest.raises(hb.HbError):
            ctx.tq_encode(f, f, f, [hb.TuJob(comp, x, 0, size, 30)], hb.TqParams(0, 1, 0.0, 1.0))
        with pytest.raises(hb.HbError):
            ctx.tq_encode_intra(f, f, f, [hb.IntraTuJob(comp, x, 0, size, 30, 3)], 1, 1, 1.0)
    with pytest.raises(hb.HbError):
        hb.Frame(ctx, 100, 60)                                                                                 # not a multiple of 8
    with pytest.raises(hb.HbError):
        hb.Prepass(ctx, w, h, qp=77)
    i16 = np.full((h, w), 300, np.int16); c16 = np.zeros((h // 2, w // 2), np.int16)
    with pytest.raises(hb.HbError):
        f.upload_i16(i16, c16, c16)                                                                            # 8-bit video only
    ok = np.full((h, w), 200, np.int16)
    f.upload_i16(ok, c16, c16)
    assert int(f.download()[0][5, 7]) == 200
    f.close()


@pytest.mark.gpu
def test_finalisation_api_rejects_bad_arguments(ctx):
    """the deblocking, SAO, bi-prediction and merge entry points refuse inconsistent input before any launch"""
    w, h = 128, 64
    a, b, small = hb.Frame(ctx, w, h), hb.Frame(ctx, w, h), hb.Frame(ctx, 64, 64)
    z = np.zeros((h, w), np.uint8); zc = np.zeros((h // 2, w // 2), np.uint8)
    for f in (a, b):
        f.upload_u8(z, zc, zc)
    n_ctus = 2
    types = np.zeros((n_ctus, 3), np.int8); offs = np.zeros((n_ctus, 3, 32), np.int16)
    with pytest.raises(hb.HbError):
        ctx.sao_apply(a, a, types, offs)                                               # classes come from the untouched picture
    with pytest.raises(hb.HbError):
        ctx.sao_apply(a, small, types, offs)
    bad_types = types.copy(); bad_types[1, 2] = 7
    with pytest.raises(hb.HbError):
        ctx.sao_apply(a, b, bad_types, offs)
    with pytest.raises(hb.HbError):
        ctx.sao_stats(a, small)
    narrow = np.zeros((h // 4, w // 4 - 1), np.uint8)
    with pytest.raises(hb.HbError):
        ctx.deblock(a, narrow, narrow, narrow)                                         # maps narrower than the picture
    with pytest.raises(hb.HbError):
        ctx.deblock_units(a, np.zeros((h // 4, w // 4 - 1), UNIT_INFO_DT))
    far = hb.McBiJob(); far.x, far.y, far.size = 0, 0, 16
    far.mv0.x = -4 * 200                                                               # further outside than the padding reaches
    with pytest.raises(hb.HbError):
        ctx.mc_predict_bi(a, b, a, [far])
    cand = hb.McJob(); cand.x, cand.y, cand.size = 0, 0, 16
    with pytest.raises(hb.HbError):
        ctx.merge_eval(a, b, a, b, [cand], 60, 0, hb.TqParams(0, 1, 0.0, 1.0))          # qp out of range
    off_grid = hb.McJob(); off_grid.x, off_grid.y, off_grid.size = 8, 0, 16            # a 16x16 unit off its own grid
    with pytest.raises(hb.HbError):
        ctx.merge_eval(a, b, a, b, [off_grid], 30, 0, hb.TqParams(0, 1, 0.0, 1.0))
    for f in (a, b, small):
        f.close()
