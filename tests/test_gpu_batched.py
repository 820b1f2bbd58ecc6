"""Batched jobs on resident frames (include/homer_b200.h section C) against the oracle: motion search
(hmr_motion_estimation), motion compensation, the inter T/Q chain.  Bit-exact."""
import numpy as np
import pytest

import homerhevc_b200 as hb
from homerhevc_b200.lib import Mv
from _frames import HostFrame, clip_pair, oracle_mc, oracle_me, oracle_tu, upload
from _oracle import chroma_qp

pytestmark = pytest.mark.gpu
W, H = 416, 240


def _me_job(x, y, size, qp, amvp, starts, parent=-1):
    j = hb.MeJob()
    j.x, j.y, j.size, j.qp, j.parent = x, y, size, qp, parent
    j.n_amvp = len(amvp)
    for i, (ax, ay) in enumerate(amvp):
        j.amvp[i] = Mv(ax, ay)
    j.n_start = len(starts)
    for i, (sx, sy) in enumerate(starts):
        j.start[i] = Mv(sx, sy)
    return j


@pytest.mark.parametrize("noise,avg_dist,action", [(3.0, 700.0, 7), (0.0, 0.0, 7), (8.0, 5000.0, 7), (3.0, 100.0, 3), (1.0, 2500.0, 1)])
def test_me_search(ctx, noise, avg_dist, action):
    rng = np.random.default_rng(int(noise * 10 + avg_dist))
    cur, ref = clip_pair(W, H, n=3, noise=noise, seed=5)
    fc, fr = upload(ctx, cur, W, H), upload(ctx, ref, W, H)
    jobs, meta = [], []
    for size in (64, 32, 16, 8):
        for _ in range(40 if size > 8 else 80):
            x = int(rng.integers(0, (W - size) // size + 1)) * size
            y = int(rng.integers(0, (H - size) // size + 1)) * size
            qp = int(rng.integers(20, 45))
            amvp = [(0, 0), (0, 0)] if rng.random() < 0.4 else [tuple(int(v) for v in rng.integers(-40, 41, 2)) for _ in range(2)]
            starts = [tuple(int(v) for v in rng.integers(-60, 61, 2)) for _ in range(int(rng.integers(0, 4)))]
            jobs.append(_me_job(x, y, size, qp, amvp, starts))
            meta.append((x, y, size, qp, amvp, starts))
    res = ctx.me_search(fc, fr, jobs, avg_dist, action)
    bad = []
    for r, (x, y, size, qp, amvp, starts) in zip(res, meta):
        o = oracle_me(cur, ref, W, H, x, y, size, qp, amvp, starts, avg_dist, action)
        got = (r.mv.x, r.mv.y, r.subpix.x, r.subpix.y, r.sad, r.n_probes)
        exp = (o.mv.x, o.mv.y, o.subpix.x, o.subpix.y, o.sad, o.n_int_sads)
        if got != exp:
            bad.append(((x, y, size), got, exp))
    assert not bad, bad[:5]
    fc.close(); fr.close()


def test_me_parent_chain(ctx):
    """children take the parent's vector as an extra start only when both components are non-zero (hmr_motion_inter.c:2613)"""
    cur, ref = clip_pair(W, H, n=5, noise=2.0, seed=9)
    fc, fr = upload(ctx, cur, W, H), upload(ctx, ref, W, H)
    parents = [_me_job(x, y, 64, 30, [(0, 0), (0, 0)], []) for y in (0, 64, 128) for x in range(0, 384, 64)]
    pres = ctx.me_search(fc, fr, parents, 700.0)
    kids, meta = [], []
    for i, p in enumerate(parents):
        for (dx, dy) in ((0, 0), (32, 0), (0, 32), (32, 32)):
            kids.append(_me_job(p.x + dx, p.y + dy, 32, 30, [(0, 0), (0, 0)], [], parent=i))
            meta.append((p.x + dx, p.y + dy, i))
    kres = ctx.me_search(fc, fr, kids, 700.0, parent_results=pres)
    for r, (x, y, i) in zip(kres, meta):
        pm = pres[i].mv
        starts = [(pm.x, pm.y)] if (pm.x != 0 and pm.y != 0) else []
        o = oracle_me(cur, ref, W, H, x, y, 32, 30, [(0, 0), (0, 0)], starts, 700.0)
        assert (r.mv.x, r.mv.y, r.sad, r.n_probes) == (o.mv.x, o.mv.y, o.sad, o.n_int_sads)
    fc.close(); fr.close()


def test_mc_predict(ctx):
    rng = np.random.default_rng(21)
    cur, ref = clip_pair(W, H, n=2, noise=3.0, seed=6)
    fr = upload(ctx, ref, W, H)
    pred = hb.Frame(ctx, W, H)
    # non-overlapping PUs on a 64 grid, each cell holds one PU of a random size
    jobs, meta = [], []
    for cy in range(0, H - 63, 64):
        for cx in range(0, W - 63, 64):
            size = int(rng.choice([64, 32, 16, 8]))
            for oy in range(0, 64, size):
                for ox in range(0, 64, size):
                    mvx, mvy = (int(v) for v in rng.integers(-70, 71, 2))
                    if rng.random() < 0.15:
                        mvx &= ~3
                    if rng.random() < 0.15:
                        mvy &= ~3
                    j = hb.McJob(cx + ox, cy + oy, size, Mv(mvx, mvy))
                    jobs.append(j); meta.append((cx + ox, cy + oy, size, mvx, mvy))
    ctx.mc_predict(fr, pred, jobs)
    py, pu, pv = pred.download()
    for (x, y, size, mvx, mvy) in meta:
        assert np.array_equal(py[y:y + size, x:x + size], oracle_mc(ref, 0, x, y, size, mvx, mvy)), ("luma", x, y, size, mvx, mvy)
        c = size // 2
        assert np.array_equal(pu[y // 2:y // 2 + c, x // 2:x // 2 + c], oracle_mc(ref, 1, x // 2, y // 2, c, mvx, mvy)), ("U", x, y, size, mvx, mvy)
        assert np.array_equal(pv[y // 2:y // 2 + c, x // 2:x // 2 + c], oracle_mc(ref, 2, x // 2, y // 2, c, mvx, mvy)), ("V", x, y, size, mvx, mvy)
    fr.close(); pred.close()


@pytest.mark.parametrize("qp,isl,sh,avg_dist", [(32, 0, 1, 700.0), (22, 0, 1, 30.0), (40, 0, 0, 3000.0), (27, 1, 1, 0.0), (36, 0, 1, 200.0)])
def test_tq_encode(ctx, qp, isl, sh, avg_dist):
    rng = np.random.default_rng(qp * 7 + isl)
    cur, ref = clip_pair(W, H, n=4, noise=4.0, seed=8)
    # prediction = reference with a little extra noise, so that some TUs keep levels and some are zeroed out
    noisy = [np.clip(pl.astype(np.int32) + np.rint(rng.normal(0, 2.0, pl.shape)).astype(np.int32), 0, 255).astype(np.uint8)
             for pl in (ref.y, ref.u, ref.v)]
    pred_h = HostFrame(*noisy)
    fc, fp = upload(ctx, cur, W, H), upload(ctx, pred_h, W, H)
    rec = hb.Frame(ctx, W, H)
    qp_c = chroma_qp(qp, 2)
    weight = 2.0 ** ((qp - qp_c) / 3.0)
    jobs, meta = [], []
    # disjoint regions per size so the reconstruction plane can be checked afterwards
    for comp in (0, 1, 2):
        pw, ph = (W, H) if comp == 0 else (W // 2, H // 2)
        bands = [(32, 0), (16, 64), (8, 128), (4, 192)] if comp == 0 else [(16, 0), (8, 48), (4, 80)]
        for size, y0 in bands:
            for y in range(y0, min(y0 + (64 if comp == 0 else 32), ph - size + 1), size):
                for x in range(0, pw - size + 1, size):
                    if rng.random() < 0.5:
                        continue
                    jobs.append(hb.TuJob(comp, x, y, size, qp if comp == 0 else qp_c)); meta.append((comp, x, y, size))
    params = hb.TqParams(isl, sh, avg_dist, weight)
    coeffs, res = ctx.tq_encode(fc, fp, rec, jobs, params)
    ry, ru, rv = rec.download()
    recs = (ry, ru, rv)
    off = 0
    n_keep = n_zeroed = 0
    for r, (comp, x, y, size) in zip(res, meta):
        eco, ede, eo = oracle_tu(cur.block(comp, x, y, size), pred_h.block(comp, x, y, size), size, comp,
                                 qp if comp == 0 else qp_c, isl, sh, avg_dist, 1.0 if comp == 0 else weight)
        got = coeffs[off:off + size * size].reshape(size, size); off += size * size
        assert (r.sum, r.ssd, r.zeroed) == (eo.sum, eo.ssd, eo.zeroed), (comp, x, y, size, (r.sum, r.ssd, r.ssd_zero, r.zeroed), (eo.sum, eo.ssd, eo.ssd_zero, eo.zeroed))
        assert np.array_equal(got, eco), ("levels", comp, x, y, size)
        assert np.array_equal(recs[comp][y:y + size, x:x + size], ede), ("recon", comp, x, y, size)
        n_keep += r.sum > 0; n_zeroed += r.zeroed
    assert n_keep > 10, "test input never produced coded TUs"
    fc.close(); fp.close(); rec.close()


@pytest.mark.parametrize("qp,isl,sh", [(30, 1, 1), (24, 1, 0), (38, 0, 1)])
def test_tq_encode_intra(ctx, qp, isl, sh):
    """the intra T/Q chain after prediction (encode_intra_cu, hmr_motion_intra.c:1023-1069, and the chroma loop): DST for 4x4
    luma, intra quant lists, mode-dependent scans, ssd against the reconstruction"""
    from _frames import oracle_intra_tu
    rng = np.random.default_rng(qp + 100 * isl)
    cur, ref = clip_pair(W, H, n=6, noise=4.0, seed=15)
    # a stand-in intra prediction: the current frame blurred a little (the host's predictors produce the real one)
    blur = [np.clip((pl.astype(np.int32) + np.roll(pl, 1, 0) + np.roll(pl, 1, 1) + np.roll(pl, -1, 1)) // 4 + rng.integers(-12, 13, pl.shape), 0, 255).astype(np.uint8)
            for pl in (cur.y, cur.u, cur.v)]
    pred_h = HostFrame(*blur)
    fc, fp = upload(ctx, cur, W, H), upload(ctx, pred_h, W, H)
    rec = hb.Frame(ctx, W, H)
    qp_c = chroma_qp(qp, 2)
    weight = 2.0 ** ((qp - qp_c) / 3.0)
    jobs, meta = [], []
    for comp in (0, 1, 2):
        pw, ph = (W, H) if comp == 0 else (W // 2, H // 2)
        bands = [(32, 0), (16, 64), (8, 128), (4, 192)] if comp == 0 else [(16, 0), (8, 48), (4, 80)]
        for size, y0 in bands:
            for y in range(y0, min(y0 + (64 if comp == 0 else 32), ph - size + 1), size):
                for x in range(0, pw - size + 1, size):
                    if rng.random() < 0.5:
                        continue
                    scan = int(rng.choice([1, 2, 3])) if size <= 8 else 3
                    jobs.append(hb.IntraTuJob(comp, x, y, size, qp if comp == 0 else qp_c, scan)); meta.append((comp, x, y, size, scan))
    coeffs, res = ctx.tq_encode_intra(fc, fp, rec, jobs, isl, sh, weight)
    recs = rec.download()
    off = 0
    coded = 0
    for r, (comp, x, y, size, scan) in zip(res, meta):
        eco, ede, eo = oracle_intra_tu(cur.block(comp, x, y, size), pred_h.block(comp, x, y, size), size, comp,
                                       qp if comp == 0 else qp_c, scan, isl, sh, 1.0 if comp == 0 else weight)
        got = coeffs[off:off + size * size].reshape(size, size); off += size * size
        assert (r.sum, r.ssd) == (eo.sum, eo.ssd), (comp, x, y, size, scan, (r.sum, r.ssd), (eo.sum, eo.ssd))
        assert np.array_equal(got, eco), ("levels", comp, x, y, size, scan)
        assert np.array_equal(recs[comp][y:y + size, x:x + size], ede), ("recon", comp, x, y, size)
        coded += r.sum > 0
    assert coded > 5
    fc.close(); fp.close(); rec.close()


def test_mc_predict_bi(ctx):
    """bi-prediction: both lists' 14-bit predictions averaged on the GPU == oracle MC (is_bi) x 2 + weighted average"""
    from _frames import oracle_mc_bi
    from homerhevc_b200.lib import Mv
    rng = np.random.default_rng(9)
    _, ref0 = clip_pair(W, H, n=2, noise=2.0, seed=5)
    _, ref1 = clip_pair(W, H, n=5, noise=4.0, seed=6)
    f0, f1 = upload(ctx, ref0, W, H), upload(ctx, ref1, W, H)
    pred = hb.Frame(ctx, W, H)
    jobs, meta = [], []
    for cy in range(0, H - 63, 64):
        for cx in range(0, W - 63, 64):
            size = int(rng.choice([64, 32, 16, 8]))
            for oy in range(0, 64, size):
                for ox in range(0, 64, size):
                    mv0 = [int(v) for v in rng.integers(-70, 71, 2)]; mv1 = [int(v) for v in rng.integers(-70, 71, 2)]
                    if rng.random() < 0.2:
                        mv0[0] &= ~3; mv1[1] &= ~3
                    if rng.random() < 0.1:
                        mv0 = [mv0[0] & ~7, mv0[1] & ~7]
                    jobs.append(hb.McBiJob(cx + ox, cy + oy, size, Mv(*mv0), Mv(*mv1))); meta.append((cx + ox, cy + oy, size, tuple(mv0), tuple(mv1)))
    ctx.mc_predict_bi(f0, f1, pred, jobs)
    py, pu, pv = pred.download()
    for (x, y, size, mv0, mv1) in meta:
        assert np.array_equal(py[y:y + size, x:x + size], oracle_mc_bi(ref0, ref1, 0, x, y, size, mv0, mv1)), ("luma", x, y, size, mv0, mv1)
        c = size // 2
        assert np.array_equal(pu[y // 2:y // 2 + c, x // 2:x // 2 + c], oracle_mc_bi(ref0, ref1, 1, x // 2, y // 2, c, mv0, mv1)), ("U", x, y, size, mv0, mv1)
        assert np.array_equal(pv[y // 2:y // 2 + c, x // 2:x // 2 + c], oracle_mc_bi(ref0, ref1, 2, x // 2, y // 2, c, mv0, mv1)), ("V", x, y, size, mv0, mv1)
    f0.close(); f1.close(); pred.close()


def test_sao_statistics(ctx):
    """SAO statistics of every CTU and component on the GPU == the restatement of sao_get_ctu_stats (whole and partial CTUs)"""
    from _oracle import oracle_sao_stats
    rng = np.random.default_rng(41)
    for (w, h) in ((320, 192), (200, 136), (64, 72)):
        base = [np.clip(rng.normal(128, 45, (hh, ww)), 0, 255) for (ww, hh) in ((w, h), (w // 2, h // 2), (w // 2, h // 2))]
        org = [b.astype(np.uint8) for b in base]
        rec = [np.clip(b + rng.normal(0, 3, b.shape), 0, 255).astype(np.uint8) for b in base]
        fo, fr = hb.Frame(ctx, w, h), hb.Frame(ctx, w, h)
        fo.upload_u8(*org); fr.upload_u8(*rec)
        got = ctx.sao_stats(fo, fr)
        exp = oracle_sao_stats(rec, org, w, h)
        for f in ("eo_diff", "eo_count", "bo_diff", "bo_count"):
            assert np.array_equal(got[f], exp[f]), (w, h, f, np.argwhere(got[f] != exp[f])[:3])
        fo.close(); fr.close()


def test_sao_offset_pass(ctx):
    """the SAO offset pass of a whole picture on the GPU == the restatement of offset_block per CTU, border refreshed"""
    from _oracle import oracle_sao_apply, random_sao_params
    rng = np.random.default_rng(42)
    for (w, h) in ((320, 192), (200, 136), (64, 72)):
        src = [np.clip(rng.normal(128, 50, (hh, ww)), 0, 255).astype(np.uint8) for (ww, hh) in ((w, h), (w // 2, h // 2), (w // 2, h // 2))]
        types, offs = random_sao_params(rng, w, h)
        fs, fd = hb.Frame(ctx, w, h), hb.Frame(ctx, w, h)
        fs.upload_u8(*src)
        ctx.sao_apply(fs, fd, types, offs)
        got = fd.download()
        exp = oracle_sao_apply(src, w, h, types, offs)
        for c in range(3):
            assert np.array_equal(got[c], exp[c]), (w, h, c, np.argwhere(got[c] != exp[c])[:4])
        fs.close(); fd.close()


def test_frame_upload_layouts_and_border(ctx):
    """uploads from contiguous planes (one copy), separate planes and strided views (pitched copies) land identically, the
    replicated border included (checked through a prediction that reaches outside the picture); int16 uploads agree"""
    from homerhevc_b200.lib import Mv
    w, h = 192, 136
    rng = np.random.default_rng(50)
    buf = rng.integers(0, 256, w * h * 3 // 2, dtype=np.uint8)
    y = buf[:w * h].reshape(h, w); u = buf[w * h:w * h * 5 // 4].reshape(h // 2, w // 2); v = buf[w * h * 5 // 4:].reshape(h // 2, w // 2)
    wide = rng.integers(0, 256, (h, w + 40), dtype=np.uint8); wide[:, 8:8 + w] = y
    wide_c = rng.integers(0, 256, (h // 2, w // 2 + 24), dtype=np.uint8); wide_c[:, 4:4 + w // 2] = u
    frames = []
    for planes in ((y, u, v), (y.copy(), u.copy(), v.copy()), (wide[:, 8:8 + w], wide_c[:, 4:4 + w // 2], v)):
        f = hb.Frame(ctx, w, h); f.upload_u8(*planes); frames.append(f)
    f16 = hb.Frame(ctx, w, h); f16.upload_i16(y.astype(np.int16), u.astype(np.int16), v.astype(np.int16)); frames.append(f16)
    ref_dl = frames[0].download()
    assert np.array_equal(ref_dl[0], y) and np.array_equal(ref_dl[1], u) and np.array_equal(ref_dl[2], v)
    # blocks at the four corners predicted from far outside the picture: only border samples contribute
    jobs = [hb.McJob(x, yy, 16, Mv(mx, my)) for (x, yy, mx, my) in ((0, 0, -4 * 40 - 1, -4 * 30 - 2), (w - 16, 0, 4 * 50 + 3, -4 * 20 - 1),
                                                                     (0, h - 24, -4 * 33 - 2, 4 * 45 + 1), (w - 16, h - 24, 4 * 60 + 1, 4 * 55 + 3))]
    outs = []
    for f in frames:
        pred = hb.Frame(ctx, w, h)
        ctx.mc_predict(f, pred, jobs)
        outs.append(pred.download()); pred.close()
    for o in outs[1:]:
        for c in range(3):
            assert np.array_equal(o[c], outs[0][c])
    # and against the oracle on an explicitly edge-padded host frame
    hf = HostFrame(y, u, v)
    for j in jobs:
        assert np.array_equal(outs[0][0][j.y:j.y + 16, j.x:j.x + 16], oracle_mc(hf, 0, j.x, j.y, 16, j.mv.x, j.mv.y)), (j.x, j.y)
    for f in frames:
        f.close()


def test_deblocking(ctx):
    """deblocking of a whole picture on the GPU == the restatement pinned against the reference's own deblocking (CPU suite):
    random strengths on the 8x8 grid (0/1/2, junk off the grid that must be ignored), QPs 10..51, tc / beta offsets"""
    from _oracle import oracle_deblock, random_deblock_case
    rng = np.random.default_rng(43)
    for (w, h, beta_off, tc_off) in ((320, 192, 0, 0), (200, 136, 2, -1), (64, 72, -3, 3)):
        m, planes = random_deblock_case(rng, w, h)
        uh, uw = m["qp"].shape
        bsv = rng.choice([0, 1, 2], (uh, uw), p=[0.3, 0.35, 0.35]).astype(np.uint8); bsh = rng.choice([0, 1, 2], (uh, uw), p=[0.3, 0.35, 0.35]).astype(np.uint8)
        qp = np.repeat(np.repeat(rng.integers(10, 52, (uh // 2 + 1, uw // 2 + 1)), 2, 0), 2, 1)[:uh, :uw].astype(np.uint8)
        f = hb.Frame(ctx, w, h); f.upload_u8(*planes)
        ctx.deblock(f, bsv, bsh, qp, 2, -1, beta_off, tc_off)
        got = f.download()
        exp = oracle_deblock(planes, w, h, bsv, bsh, qp, (2, -1), beta_off, tc_off)
        for c in range(3):
            assert np.array_equal(got[c], exp[c]), (w, h, c, np.argwhere(got[c] != exp[c])[:4])
        assert (got[0] != planes[0]).sum() > 200
        f.close()


def test_deblocking_against_reference_pictures(ctx):
    """the GPU picture equals the picture the reference's own deblocking produces, on the strengths the reference derived"""
    from _oracle import have_ref, random_deblock_case, ref_deblock
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(44)
    for (w, h) in ((192, 136), (128, 128)):
        m, planes = random_deblock_case(rng, w, h)
        exp, bsv, bsh, offs = ref_deblock(planes, w, h, m)
        f = hb.Frame(ctx, w, h); f.upload_u8(*planes)
        ctx.deblock(f, bsv, bsh, m["qp"], offs[0], offs[1])
        got = f.download()
        for c in range(3):
            assert np.array_equal(got[c], exp[c]), (w, h, c, np.argwhere(got[c] != exp[c])[:4])
        # strengths derived on the device from the per-unit mode data
        units = np.zeros(m["qp"].shape, hb.lib.UNIT_INFO_DT)
        units["cu_depth"], units["tu_depth"], units["intra"], units["cbf_luma"], units["qp"] = m["cu"], m["tu"], m["intra"], m["cbf"], m["qp"]
        units["ref_idx"] = np.where(m["intra"] != 0, -1, 0); units["mvx"] = m["mv"][..., 0]; units["mvy"] = m["mv"][..., 1]
        f.upload_u8(*planes)
        gbv, gbh = ctx.deblock_units(f, units, offs[0], offs[1])
        assert np.array_equal(gbv[:h // 4, :w // 4], bsv[:h // 4, :w // 4]) and np.array_equal(gbh[:h // 4, :w // 4], bsh[:h // 4, :w // 4])
        got = f.download()
        for c in range(3):
            assert np.array_equal(got[c], exp[c]), ("units", w, h, c)
        f.close()


def test_merge_candidates(ctx):
    """merge / skip candidate evaluation: per candidate the coded distortion and level sum of its units (oracle MC + oracle T/Q
    chain) and the no-residual distortion exactly as check_rd_cost_merge_2nx2n adds it up (ssd luma + truncated weighted chroma)"""
    from homerhevc_b200.lib import Mv
    rng = np.random.default_rng(45)
    qp, off, avg = 30, 2, 420.0
    cur, ref = clip_pair(W, H, n=3, noise=5.0, seed=29)
    fc, fr = upload(ctx, cur, W, H), upload(ctx, ref, W, H)
    pred, rec = hb.Frame(ctx, W, H), hb.Frame(ctx, W, H)
    cands = []
    for cy in range(0, H - 63, 64):
        for cx in range(0, W - 63, 64):
            size = int(rng.choice([64, 32, 16, 8]))
            for oy in range(0, 64, size):
                for ox in range(0, 64, size):
                    if rng.random() < 0.5:
                        cands.append(hb.McJob(cx + ox, cy + oy, size, Mv(int(rng.integers(-30, 31)), int(rng.integers(-30, 31)))))
    prm = hb.TqParams(0, 1, avg, 1.0)
    got = ctx.merge_eval(fc, fr, pred, rec, cands, qp, off, prm)
    qp_c = chroma_qp(qp, off)
    weight = 2.0 ** ((qp - qp_c) / 3.0)
    n_coded = 0
    for j, g in zip(cands, got):
        dist = ssum = 0
        skip = 0
        ls = 32 if j.size == 64 else j.size
        p_y = oracle_mc(ref, 0, j.x, j.y, j.size, j.mv.x, j.mv.y)
        for yy in range(0, j.size, ls):
            for xx in range(0, j.size, ls):
                _, _, eo = oracle_tu(cur.block(0, j.x + xx, j.y + yy, ls), p_y[yy:yy + ls, xx:xx + ls], ls, 0, qp, 0, 1, avg, 1.0)
                dist += eo.ssd; ssum += eo.sum
        skip += int(((cur.block(0, j.x, j.y, j.size).astype(np.int64) - p_y) ** 2).sum())
        c = j.size // 2
        cs = ls // 2
        for comp in (1, 2):
            p_c = oracle_mc(ref, comp, j.x // 2, j.y // 2, c, j.mv.x, j.mv.y)
            for yy in range(0, c, cs):
                for xx in range(0, c, cs):
                    _, _, eo = oracle_tu(cur.block(comp, j.x // 2 + xx, j.y // 2 + yy, cs), p_c[yy:yy + cs, xx:xx + cs], cs, comp, qp_c, 0, 1, avg, weight)
                    dist += eo.ssd; ssum += eo.sum
            skip += int(weight * float(((cur.block(comp, j.x // 2, j.y // 2, c).astype(np.int64) - p_c) ** 2).sum()))
        assert (int(g["dist_coded"]), int(g["sum"]), int(g["dist_skip"])) == (dist, ssum, skip), (j.x, j.y, j.size, j.mv.x, j.mv.y)
        n_coded += ssum > 0
    assert n_coded > 5 and len(cands) > 30
    fc.close(); fr.close(); pred.close(); rec.close()


@pytest.mark.gpu
def test_amvp_candidates_from_unit_field(ctx):
    """hb_amvp_candidates == the restatement of get_amvp_candidates (itself pinned against the reference, tests/test_oracle_vs_ref.py) for every
    2Nx2N PU of every size on random CU trees, intra / inter units, per-unit vectors, partial CTUs on both edges; argument checks"""
    from homerhevc_b200.lib import UNIT_INFO_DT, HbError
    from _oracle import amvp_jobs, oracle_amvp, oracle_merge, random_deblock_case
    rng = np.random.default_rng(77)
    two = 0
    for (w, h) in ((192, 136), (200, 72), (72, 200), (1280, 720)):
        m, _ = random_deblock_case(rng, w, h)
        m["mv"] = rng.integers(-300, 301, m["mv"].shape).astype(np.int16)
        units = np.zeros(m["cu"].shape, UNIT_INFO_DT)
        units["cu_depth"], units["tu_depth"], units["intra"], units["cbf_luma"], units["qp"] = m["cu"], m["tu"], m["intra"], m["cbf"], m["qp"]
        units["ref_idx"] = np.where(m["intra"] != 0, -1, 0); units["mvx"], units["mvy"] = m["mv"][..., 0], m["mv"][..., 1]
        jobs = amvp_jobs(w, h)
        got = ctx.amvp_candidates(units, w, h, jobs).reshape(-1, 4)
        exp = oracle_amvp(w, h, m, jobs)
        bad = np.argwhere((got != exp).any(1))
        assert not len(bad), (w, h, jobs[bad[0, 0]], got[bad[0, 0]], exp[bad[0, 0]])
        two += int(((exp[:, 2] != 0) | (exp[:, 3] != 0)).sum())
        m["mv"] = (m["mv"] // 150).astype(np.int16)                         # few distinct vectors: the merge pruning has work to do
        units["mvx"], units["mvy"] = m["mv"][..., 0], m["mv"][..., 1]
        for mx in (5, 2):
            gm, em = ctx.merge_candidates(units, w, h, jobs, mx), oracle_merge(w, h, m, jobs, mx)
            bad = np.argwhere((gm != em).any((1, 2)))
            assert not len(bad), ("merge", w, h, mx, jobs[bad[0, 0]], gm[bad[0, 0]].tolist(), em[bad[0, 0]].tolist())
    assert two > 1000
    with pytest.raises(HbError):
        ctx.amvp_candidates(units, w, h, np.array([[4, 0, 8]], np.int32))        # not on the size grid
    with pytest.raises(HbError):
        ctx.amvp_candidates(units, w, h, np.array([[w - 8, 0, 16]], np.int32))   # leaves the picture


def test_search_with_predictors_from_the_unit_field(ctx):
    """hb_me_search_field (AMVP lists derived on the device right before the search) == hb_amvp_candidates + hb_me_search, and the
    predictors do change vectors / costs against the zero-predictor search"""
    from homerhevc_b200.lib import UNIT_INFO_DT
    from _oracle import amvp_jobs, random_deblock_case
    w, h, qp, avg = 192, 136, 30, 500.0
    cur, ref = clip_pair(w, h, n=2, noise=6.0, seed=5)
    fc, fr = upload(ctx, cur, w, h), upload(ctx, ref, w, h)
    rng = np.random.default_rng(8)
    m, _ = random_deblock_case(rng, w, h)
    m["mv"] = rng.integers(-24, 25, m["mv"].shape).astype(np.int16)
    units = np.zeros(m["cu"].shape, UNIT_INFO_DT)
    units["intra"] = m["intra"]; units["ref_idx"] = np.where(m["intra"] != 0, -1, 0); units["mvx"], units["mvy"] = m["mv"][..., 0], m["mv"][..., 1]
    pus = amvp_jobs(w, h)
    lists = ctx.amvp_candidates(units, w, h, pus)
    jobs_zero = [hb.MeJob(int(x), int(y), int(s), qp, 2, (Mv * 2)(Mv(0, 0), Mv(0, 0)), 0, (Mv * 3)(), -1) for (x, y, s) in pus]
    jobs_amvp = [hb.MeJob(int(x), int(y), int(s), qp, 2, (Mv * 2)(Mv(int(l[0][0]), int(l[0][1])), Mv(int(l[1][0]), int(l[1][1]))), 0, (Mv * 3)(), -1)
                 for (x, y, s), l in zip(pus, lists)]
    a = ctx.me_search(fc, fr, jobs_amvp, avg)
    b = ctx.me_search_field(fc, fr, units, jobs_zero, avg)
    z = ctx.me_search(fc, fr, jobs_zero, avg)
    key = lambda r: (r.mv.x, r.mv.y, r.subpix.x, r.subpix.y, r.sad, r.n_probes)
    assert [key(r) for r in a] == [key(r) for r in b]
    assert sum(key(p) != key(q) for p, q in zip(a, z)) > 10
    fc.close(); fr.close()


def test_boundary_strengths_of_b_pictures(ctx):
    """hb_deblock_frame_units_b: strengths of a B picture derived on the device == the restatement of the reference's two-list rule (pinned
    against hmr_deblock_filter_cu in tests/test_oracle_vs_ref.py), and the picture filtered with them == the oracle's pixel stage"""
    from homerhevc_b200.lib import UNIT_INFO_DT, UNIT_L1_DT
    from _oracle import oracle_deblock, oracle_deblock_strengths_b, random_b_motion, random_deblock_case
    rng = np.random.default_rng(229)
    for (w, h) in ((192, 136), (200, 72)):
        m, planes = random_deblock_case(rng, w, h)
        m["cbf"] = (m["cbf"] * (rng.random(m["cbf"].shape) < 0.25)).astype(np.uint8)
        ref0, mv0, ref1, mv1, p0, p1 = random_b_motion(rng, m)
        units = np.zeros(m["cu"].shape, UNIT_INFO_DT); u1 = np.zeros(m["cu"].shape, UNIT_L1_DT)
        units["cu_depth"], units["tu_depth"], units["intra"], units["cbf_luma"], units["qp"] = m["cu"], m["tu"], m["intra"], m["cbf"], m["qp"]
        units["ref_idx"], units["mvx"], units["mvy"] = ref0, mv0[..., 0], mv0[..., 1]
        u1["ref_idx"], u1["mvx"], u1["mvy"] = ref1, mv1[..., 0], mv1[..., 1]
        f = hb.Frame(ctx, w, h); f.upload_u8(*planes)
        bsv, bsh = ctx.deblock_units_b(f, units, u1, p0, p1, 1, -1)
        ev, eh = oracle_deblock_strengths_b(w, h, m, ref0, mv0, ref1, mv1, p0, p1)
        assert np.array_equal(bsv[:h // 4, :w // 4], ev[:h // 4, :w // 4]) and np.array_equal(bsh[:h // 4, :w // 4], eh[:h // 4, :w // 4]), (w, h)
        assert {0, 1, 2} <= set(np.unique(ev[:h // 4, 2:w // 4:2]))
        exp = oracle_deblock(planes, w, h, ev, eh, m["qp"], (1, -1))
        got = f.download()
        for c in range(3):
            assert np.array_equal(got[c], exp[c]), (w, h, c)
        f.close()
