#!/usr/bin/env python
"""Generate tests/golden/ref_vectors.npz: outputs of the UNMODIFIED reference (compiled into oracle/_ref by
oracle/Makefile from /root/reference) on seeded inputs.  The reference ships no golden vectors of its own
(SURVEY.md section 4), so these pin the oracle: tests/test_oracle_golden.py replays the inputs through
oracle/liboracle.so and demands identical outputs, here and on the GPU box (where /root/reference does not exist).

    python tests/golden/make_golden.py          # needs /root/reference (build container only)
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _oracle import aligned_i16, i32p, make_frame_pair, ptr, ref, refdrv  # noqa: E402


def main():
    R, D = ref()
    h = refdrv()
    assert D.refdrv_sse_selected(h) == 1, "the golden must come from the SSE4.2 table the reference selects on x86"
    rng = np.random.default_rng(20260101)
    g = {}

    # ---- a1-a4: SAD / SSD / predict / reconst (hmr_sse42_functions_pixel.c:462/:728/:817/:919)
    a = aligned_i16(64 * 64); b = aligned_i16(96 * 80)
    a[:] = rng.integers(0, 256, a.size); b[:] = rng.integers(0, 256, b.size)
    res = aligned_i16(64 * 64); res[:] = rng.integers(-300, 301, res.size)
    g["pix_a"], g["pix_b"], g["pix_res"] = a.copy(), b.copy(), res.copy()
    offs = rng.integers(0, 600, 10)
    g["pix_off"] = offs
    sad, ssd, pre, rec = [], [], [], []
    for n in (4, 8, 16, 32, 64):
        for off in offs[:2]:
            sad.append(R.sse_aligned_sad(ptr(a), 64, ptr(b, int(off)), 96, n))
            ssd.append(R.sse_aligned_ssd16b(ptr(a), 64, ptr(b, int(off)), 96, n))
        r = aligned_i16(64 * 64); d = aligned_i16(64 * 64)
        R.sse_aligned_predict(ptr(a), 64, ptr(b), 96, ptr(r), 64, n)
        R.sse_aligned_reconst(ptr(b), 96, ptr(res), 64, ptr(d), 64, n)
        pre.append(r.reshape(64, 64)[:n, :n].copy().reshape(-1)); rec.append(d.reshape(64, 64)[:n, :n].copy().reshape(-1))
    g["pix_sad"], g["pix_ssd"] = np.array(sad, np.uint32), np.array(ssd, np.uint32)
    g["pix_predict"], g["pix_reconst"] = np.concatenate(pre), np.concatenate(rec)

    # ---- a5/a6: interpolation passes, all four (is_first, is_last) modes
    src8 = aligned_i16(96 * 96); src14 = aligned_i16(96 * 96)
    src8[:] = rng.integers(0, 256, src8.size); src14[:] = rng.integers(-8192, 8129, src14.size)
    g["int_src8"], g["int_src14"] = src8.copy(), src14.copy()
    cases, outs = [], []
    for chroma in (0, 1):
        for frac in range(8 if chroma else 4):
            for (first, last) in ((1, 0), (0, 1), (1, 1), (0, 0)):
                for vert in (0, 1):
                    w = int(rng.choice([4, 8, 16, 32] if chroma else [8, 16, 32, 64])); hh = int(rng.choice([4, 8, 17, 32, 40]))
                    d = aligned_i16(80 * 80)
                    f = R.sse_interpolate_chroma if chroma else R.sse_interpolate_luma
                    f(ptr(src8 if first else src14, 96 * 8 + 8), 96, ptr(d), 80, frac, w, hh, vert, first, last)
                    cases.append((chroma, frac, first, last, vert, w, hh))
                    outs.append(d.reshape(80, 80)[:hh, :w].copy().reshape(-1))
    g["int_cases"], g["int_out"] = np.array(cases, np.int32), np.concatenate(outs)

    # ---- a11/a12: transforms
    tcases, tout = [], []
    blk = aligned_i16(64 * 64); blk[:] = rng.integers(-255, 256, blk.size)
    g["tx_block"] = blk.copy()
    coefs = rng.integers(-32768, 32768, 1024).astype(np.int16)
    g["tx_coef"] = coefs
    for n in (4, 8, 16, 32):
        for dst in ((0, 1) if n == 4 else (0,)):
            c = aligned_i16(1024); aux = aligned_i16(1024); co = aligned_i16(1024); bb = aligned_i16(64 * 64)
            lg = n.bit_length() - 1
            R.sse_transform(8, ptr(blk), ptr(c), 64, n, n, lg, lg, C.c_uint16(0 if dst else 65535), ptr(aux))
            co[:n * n] = coefs[:n * n]
            R.sse_itransform(8, ptr(bb), ptr(co), 64, n, n, C.c_uint(0 if dst else 65535), ptr(aux))
            tcases.append((n, dst)); tout.append(c[:n * n].copy()); tout.append(bb.reshape(64, 64)[:n, :n].copy().reshape(-1))
    g["tx_cases"], g["tx_out"] = np.array(tcases, np.int32), np.concatenate(tout)

    # ---- tables
    g["scan_diag"] = np.concatenate([np.ctypeslib.as_array(D.refdrv_scan(h, 3, lg), ((1 << lg) ** 2,)).copy() for lg in (2, 3, 4, 5)])
    g["scan_hor"] = np.concatenate([np.ctypeslib.as_array(D.refdrv_scan(h, 1, lg), ((1 << lg) ** 2,)).copy() for lg in (2, 3)])
    g["scan_ver"] = np.concatenate([np.ctypeslib.as_array(D.refdrv_scan(h, 2, lg), ((1 << lg) ** 2,)).copy() for lg in (2, 3)])
    qt = []
    for lg in (2, 3, 4, 5):
        for lst in ((0, 3, 4, 5) if lg < 5 else (0, 3)):
            for rem in (0, 3, 5):
                qt.append(np.ctypeslib.as_array(D.refdrv_quant_table(h, lg, lst, rem), ((1 << lg) ** 2,)).copy())
                qt.append(np.ctypeslib.as_array(D.refdrv_dequant_table(h, lg, lst, rem), ((1 << lg) ** 2,)).copy())
    g["qtables"] = np.concatenate(qt)

    # ---- a13/a14: quant (SSE4.2 rule) + sign hiding, dequant
    qcases, qout = [], []
    qsrc = []
    for it in range(96):
        lg = int(rng.choice([2, 3, 4, 5])); n = 1 << lg
        comp = int(rng.integers(0, 3)) if lg < 5 else 0
        is_intra, isl, sh = (int(v) for v in rng.integers(0, 2, 3))
        qp = int(rng.integers(10, 48))
        scan = int(rng.choice([1, 2, 3])) if lg <= 3 else 3
        src = aligned_i16(1024)
        src[:n * n] = (rng.laplace(0, [40, 300, 2000][it % 3], n * n) * (rng.random(n * n) < [1, .3, 1][it % 3])).clip(-32768, 32767).astype(np.int16)
        d = aligned_i16(1024); u = aligned_i16(1024); q = aligned_i16(1024); s = C.c_int(0)
        D.refdrv_quant(h, ptr(src), ptr(d), ptr(u), scan, lg, comp, is_intra, isl, sh, qp // 6, qp % 6, C.byref(s))
        D.refdrv_inv_quant(h, ptr(d), ptr(q), lg, comp, is_intra, qp // 6, qp % 6)
        qcases.append((lg, comp, is_intra, isl, sh, qp, scan, s.value)); qsrc.append(src[:n * n].copy())
        qout += [d[:n * n].copy(), u[:n * n].copy(), q[:n * n].copy()]
    g["q_cases"], g["q_src"], g["q_out"] = np.array(qcases, np.int32), np.concatenate(qsrc), np.concatenate(qout)

    # ---- a7-a10: motion estimation / compensation on a small synthetic pair
    W, H, PAD = 208, 120, 80
    cur, rf = make_frame_pair(rng, W, H, PAD, shift=(5, 3), noise=3.0)
    cur = np.ascontiguousarray(cur); rf = np.ascontiguousarray(rf)
    g["me_cur"], g["me_ref"] = cur[PAD:-PAD, PAD:-PAD].astype(np.uint8), rf[PAD:-PAD, PAD:-PAD].astype(np.uint8)
    S = W + 2 * PAD
    mecases, meout = [], []
    for it in range(120):
        n = int(rng.choice([8, 16, 32, 64]))
        gx = int(rng.integers(0, (W - n) // n + 1)) * n; gy = int(rng.integers(0, (H - n) // n + 1)) * n
        ob = aligned_i16(64 * 64); ob.reshape(64, 64)[:n, :n] = cur[PAD + gy:PAD + gy + n, PAD + gx:PAD + gx + n]
        amvp = np.array(rng.integers(-40, 41, 4) if it % 3 else [0, 0, 0, 0], np.int32)
        ns = int(rng.integers(0, 4)); st = np.array(rng.integers(-60, 61, 6), np.int32)
        qp = int(rng.integers(20, 45)); avg = float(rng.choice([0., 100., 700., 2500., 5000.])); action = int(rng.choice([7, 7, 7, 3, 1]))
        out = np.zeros(4, np.int32)
        r = D.refdrv_motion_estimation(h, ptr(ob), 64, ptr(rf.reshape(-1), (PAD + gy) * S + PAD + gx), S, gx, gy, n, W, H, 2,
                                       amvp.ctypes.data_as(i32p), ns, st.ctypes.data_as(i32p), qp, avg, action, out.ctypes.data_as(i32p))
        mecases.append([gx, gy, n, qp, int(avg), action, ns] + list(amvp) + list(st)); meout.append(list(out) + [r])
    g["me_cases"], g["me_out"] = np.array(mecases, np.int32), np.array(meout, np.int64)
    mccases, mcout = [], []
    for it in range(60):
        n = int(rng.choice([4, 8, 16, 32])); chroma = it % 2
        gx = int(rng.integers(0, (W - n) // 4 + 1)) * 4; gy = int(rng.integers(0, (H - n) // 4 + 1)) * 4
        mvx, mvy = (int(v) for v in rng.integers(-70, 71, 2))
        p = aligned_i16(64 * 64)
        (D.refdrv_mc_chroma if chroma else D.refdrv_mc_luma)(h, ptr(rf.reshape(-1), (PAD + gy) * S + PAD + gx), S, ptr(p), 64, n, mvx, mvy)
        mccases.append((chroma, gx, gy, n, mvx, mvy)); mcout.append(p.reshape(64, 64)[:n, :n].copy().reshape(-1))
    g["mc_cases"], g["mc_out"] = np.array(mccases, np.int32), np.concatenate(mcout)

    # ---- a15: the inter T/Q chain
    tqcases, tqin, tqout = [], [], []
    for it in range(120):
        depth = int(rng.choice([1, 2, 3, 4])); comp = int(rng.integers(0, 3)) if depth < 4 else 0
        ncu = 64 >> depth; n = ncu if comp == 0 else ncu // 2
        part = int(rng.integers(0, 4 ** depth)); qp = int(rng.integers(18, 46)); isl, sh = (int(v) for v in rng.integers(0, 2, 2))
        avg = float(rng.choice([0., 30., 200., 900., 3000.]))
        orig = aligned_i16(n * n); pred = aligned_i16(n * n)
        orig[:] = rng.integers(0, 256, n * n)
        pred[:] = np.clip(orig + np.rint(rng.normal(0, float(rng.choice([1, 3, 10, 40])), n * n)), 0, 255)
        c1 = aligned_i16(1024); d1 = aligned_i16(1024); s1 = C.c_int(0)
        r = D.refdrv_encode_inter_tu(h, ptr(orig), ptr(pred), depth, part, comp, qp, isl, sh, avg, ptr(c1), ptr(d1), C.byref(s1))
        tqcases.append((depth, comp, n, qp, D.refdrv_chroma_qp(h, qp), isl, sh, int(avg), s1.value, r & 0xffffffff))
        tqin += [orig.copy(), pred.copy()]; tqout += [c1[:n * n].copy(), d1[:n * n].copy()]
    g["tq_cases"], g["tq_in"], g["tq_out"] = np.array(tqcases, np.int64), np.concatenate(tqin), np.concatenate(tqout)

    # ---- 8f item 1: reference-sample smoothing + the 35 intra predictors; bi-prediction average (table members)
    ip_adi, ip_flt, ip_pred, ip_cases = [], [], [], []
    for n in (4, 8, 16, 32):
        for kind in range(2):
            adi = aligned_i16(4 * n + 1 + 8)
            if kind == 0:
                adi[:4 * n + 1] = rng.integers(0, 256, 4 * n + 1)
            else:                                                        # smooth: the strong filter at 32
                adi[:4 * n + 1] = np.clip(np.linspace(rng.integers(20, 90), rng.integers(150, 240), 4 * n + 1) + rng.integers(-1, 2, 4 * n + 1), 0, 255)
            flt = aligned_i16(4 * n + 1 + 8)
            D.refdrv_adi_filter(h, ptr(adi), ptr(flt), n)
            ip_adi.append(adi[:4 * n + 1].copy()); ip_flt.append(flt[:4 * n + 1].copy())
            for mode in range(35):
                for is_luma in ((1, 0) if mode in (1, 10, 26) else (1,)):
                    pr = aligned_i16(n * n)
                    D.refdrv_intra_predict(h, ptr(flt if (kind and is_luma) else adi), n, mode, is_luma, ptr(pr))
                    ip_cases.append((n, kind, mode, is_luma)); ip_pred.append(pr.astype(np.uint8))
    g["ip_adi"], g["ip_flt"], g["ip_pred"], g["ip_cases"] = np.concatenate(ip_adi), np.concatenate(ip_flt), np.concatenate(ip_pred), np.array(ip_cases, np.int32)
    wa, wb = aligned_i16(64 * 64), aligned_i16(64 * 64)
    wa[:] = rng.integers(-14312, 14249, wa.size); wb[:] = rng.integers(-8192, 8193, wb.size)
    wd = aligned_i16(64 * 64)
    D.refdrv_weighted_average(h, ptr(wa), 64, ptr(wb), 64, ptr(wd), 64, 64, 64)
    g["wavg_a"], g["wavg_b"], g["wavg_out"] = wa.copy(), wb.copy(), wd.astype(np.uint8)

    # ---- 8f item 4: SAO statistics of a small picture with partial CTUs (table member get_sao_stats)
    from _oracle import ref_sao_stats
    sw, shh = 200, 136
    base = [np.clip(rng.normal(128, 40, (hh_, ww_)), 0, 255) for (ww_, hh_) in ((sw, shh), (sw // 2, shh // 2), (sw // 2, shh // 2))]
    sao_org = [b_.astype(np.uint8) for b_ in base]
    sao_rec = [np.clip(b_ + rng.normal(0, 3, b_.shape), 0, 255).astype(np.uint8) for b_ in base]
    st = ref_sao_stats(sao_rec, sao_org, sw, shh)
    g["sao_org"] = np.concatenate([p_.reshape(-1) for p_ in sao_org]); g["sao_rec"] = np.concatenate([p_.reshape(-1) for p_ in sao_rec])
    g["sao_stats"] = np.frombuffer(st.tobytes(), np.int32).copy()

    # ---- 8f item 4: deblocking of a small picture by the reference's own per-CTU function, with the strengths it derived
    from _oracle import random_deblock_case, ref_deblock
    dw, dh = 192, 136
    dm, dplanes = random_deblock_case(rng, dw, dh)
    dexp, dbsv, dbsh, doffs = ref_deblock(dplanes, dw, dh, dm)
    g["dbk_in"] = np.concatenate([p_.reshape(-1) for p_ in dplanes]); g["dbk_out"] = np.concatenate([p_.reshape(-1) for p_ in dexp])
    g["dbk_bsv"], g["dbk_bsh"], g["dbk_qp"], g["dbk_offs"] = dbsv, dbsh, dm["qp"], np.array(doffs, np.int32)

    # ---- 8f item 4: the arithmetic half of the SAO decision (sao_derive_offsets + sao_get_distortion) on random statistics
    from _oracle import random_sao_stats, ref_sao_derive
    sd = random_sao_stats(rng, 24)
    sd_lam = rng.choice([0.5, 7.0, 33.3, 120.0, 900.0, 1e4], len(sd))
    sd_off, sd_band, sd_dist = [], [], []
    for i_, rec_ in enumerate(sd):
        for t_ in range(5):
            o_, b_, d_ = ref_sao_derive(rec_, i_ % 3, t_, sd_lam[i_])
            sd_off.append(o_.astype(np.int16)); sd_band.append(b_ if t_ == 4 else 0); sd_dist.append(d_)
    g["saod_stats"] = np.frombuffer(sd.tobytes(), np.int32).copy(); g["saod_lambda"] = sd_lam
    g["saod_off"], g["saod_band"], g["saod_dist"] = np.array(sd_off), np.array(sd_band, np.int32), np.array(sd_dist, np.int64)

    out = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
