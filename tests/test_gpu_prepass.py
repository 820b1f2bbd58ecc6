"""Frame-level pre-pass (include/homer_b200.h section D) against the oracle chain on a small frame, and
size-independent properties at the benchmark's frame size."""
import numpy as np
import pytest

import homerhevc_b200 as hb
from _frames import clip_pair, oracle_mc, oracle_me, oracle_tu, upload
from _oracle import chroma_qp

pytestmark = pytest.mark.gpu


def _oracle_prepass_me(cur, ref, w, h, qp, avg_dist, action=7):
    """depth by depth, parent vector as extra start when both components are non-zero"""
    out = []
    for d in range(4):
        s = 64 >> d
        gw, gh = ((w + 63) // 64) * (64 // s), ((h + 63) // 64) * (64 // s)
        tab = {}
        for py in range(gh):
            for px in range(gw):
                x, y = px * s, py * s
                if x + s > w or y + s > h:
                    continue
                starts = []
                if d > 0:
                    par = out[d - 1].get((px // 2, py // 2))
                    if par is not None and par.mv.x != 0 and par.mv.y != 0:
                        starts = [(par.mv.x, par.mv.y)]
                tab[(px, py)] = oracle_me(cur, ref, w, h, x, y, s, qp, [(0, 0), (0, 0)], starts, avg_dist, action)
        out.append(tab)
    return out


@pytest.mark.parametrize("w,h,use_graph", [(192, 136, 0), (256, 128, 1), (200, 104, 1)])
def test_prepass_matches_oracle(ctx, w, h, use_graph):
    qp, avg_dist = 32, 650.0
    cur, ref = clip_pair(w, h, n=2, noise=3.0, seed=3)
    fc, fr = upload(ctx, cur, w, h), upload(ctx, ref, w, h)
    pp = hb.Prepass(ctx, w, h, qp=qp, use_graph=use_graph)
    for rep in range(2):                       # twice: the second run replays the captured graph
        pp.run(fc, fr, avg_dist)
    ctx.sync()
    exp_me = _oracle_prepass_me(cur, ref, w, h, qp, avg_dist)
    qp_c = chroma_qp(qp, 2)
    weight = 2.0 ** ((qp - qp_c) / 3.0)
    preds = []
    for d in range(4):
        s = 64 >> d
        got = pp.fetch_me(d)
        gw = ((w + 63) // 64) * (64 // s)
        py_, pu_, pv_ = pp.pred(d).download()
        for idx, r in enumerate(got):
            px, py = idx % gw, idx // gw
            e = exp_me[d].get((px, py))
            if e is None:
                assert r["sad"] == 0xFFFFFFFF
                continue
            assert (r["mvx"], r["mvy"], r["subx"], r["suby"], r["sad"], r["n_probes"]) == (e.mv.x, e.mv.y, e.subpix.x, e.subpix.y, e.sad, e.n_int_sads), (d, px, py)
            x, y = px * s, py * s
            assert np.array_equal(py_[y:y + s, x:x + s], oracle_mc(ref, 0, x, y, s, e.mv.x, e.mv.y)), ("pred Y", d, px, py)
            assert np.array_equal(pu_[y // 2:(y + s) // 2, x // 2:(x + s) // 2], oracle_mc(ref, 1, x // 2, y // 2, s // 2, e.mv.x, e.mv.y)), ("pred U", d, px, py)
            assert np.array_equal(pv_[y // 2:(y + s) // 2, x // 2:(x + s) // 2], oracle_mc(ref, 2, x // 2, y // 2, s // 2, e.mv.x, e.mv.y)), ("pred V", d, px, py)
        preds.append((py_, pu_, pv_))
    n_coded = 0
    for p in range(5):
        d = min(p, 3)
        rec = pp.recon(p).download()
        for comp in range(3):
            t = pp.tu_size(p, comp)
            if not t:
                assert p == 4 and comp > 0
                continue
            xy = pp.tu_xy(p, comp)
            res = pp.fetch_tu(p, comp)
            co = pp.fetch_coeffs(p, comp)
            assert len(xy) == len(res) == len(co) > 0
            for (x, y), r, c in zip(xy, res, co):
                eco, ede, eo = oracle_tu(cur.block(comp, x, y, t), preds[d][comp][y:y + t, x:x + t], t, comp,
                                         qp if comp == 0 else qp_c, 0, 1, avg_dist, 1.0 if comp == 0 else weight)
                assert (r["sum"], r["ssd"], r["zeroed"]) == (eo.sum, eo.ssd, eo.zeroed), (p, comp, x, y)
                assert np.array_equal(c, eco), ("levels", p, comp, x, y)
                assert np.array_equal(rec[comp][y:y + t, x:x + t], ede), ("recon", p, comp, x, y)
                n_coded += r["sum"] > 0
    assert n_coded > 0
    # the packed fetch carries the same bytes as the individual ones
    pin = ctx.pinned(pp.output_bytes())
    assert pp.fetch_all(pin) == pp.output_bytes()
    me0 = pp.fetch_me(0)
    assert bytes(pin[:me0.nbytes]) == me0.tobytes()
    pp.close(); fc.close(); fr.close()


@pytest.mark.parametrize("action", [hb.ME_PEL, hb.ME_PEL | hb.ME_HALF])
def test_prepass_other_search_precisions(ctx, action):
    """integer-only search (luma and chroma predictions from the compensation kernels, nothing fused) and half-pel search:
    vectors, probe counts and predictions of every depth against the oracle"""
    w, h, qp, avg_dist = 192, 136, 32, 650.0
    cur, ref = clip_pair(w, h, n=2, noise=3.0, seed=8)
    fc, fr = upload(ctx, cur, w, h), upload(ctx, ref, w, h)
    pp = hb.Prepass(ctx, w, h, qp=qp, me_action=action)
    pp.run(fc, fr, avg_dist); ctx.sync()
    exp_me = _oracle_prepass_me(cur, ref, w, h, qp, avg_dist, action)
    for d in range(4):
        s = 64 >> d
        got = pp.fetch_me(d)
        gw = ((w + 63) // 64) * (64 // s)
        py_, pu_, pv_ = pp.pred(d).download()
        for idx, r in enumerate(got):
            px, py = idx % gw, idx // gw
            e = exp_me[d].get((px, py))
            if e is None:
                continue
            assert (r["mvx"], r["mvy"], r["sad"], r["n_probes"]) == (e.mv.x, e.mv.y, e.sad, e.n_int_sads), (d, px, py)
            x, y = px * s, py * s
            assert np.array_equal(py_[y:y + s, x:x + s], oracle_mc(ref, 0, x, y, s, e.mv.x, e.mv.y)), ("pred Y", d, px, py)
            assert np.array_equal(pu_[y // 2:(y + s) // 2, x // 2:(x + s) // 2], oracle_mc(ref, 1, x // 2, y // 2, s // 2, e.mv.x, e.mv.y)), ("pred U", d, px, py)
    pp.close(); fc.close(); fr.close()


def test_prepass_quantiser_variants(ctx):
    """the pre-pass T/Q with the I-slice rounding and without sign-data hiding (all unit sizes, the one-thread 4x4 kernel included)"""
    w, h, qp, avg_dist = 128, 72, 27, 120.0
    cur, ref = clip_pair(w, h, n=2, noise=6.0, seed=15)
    fc, fr = upload(ctx, cur, w, h), upload(ctx, ref, w, h)
    for (isl, sh) in ((1, 0), (1, 1), (0, 0)):
        pp = hb.Prepass(ctx, w, h, qp=qp, sign_hiding=sh, is_islice=isl)
        pp.run(fc, fr, avg_dist); ctx.sync()
        qp_c = chroma_qp(qp, 2)
        weight = 2.0 ** ((qp - qp_c) / 3.0)
        n_coded = 0
        for p in range(5):
            d = min(p, 3)
            pred = pp.pred(d).download()
            rec = pp.recon(p).download()
            for comp in range(3):
                t = pp.tu_size(p, comp)
                if not t:
                    continue
                for (x, y), r, c in zip(pp.tu_xy(p, comp), pp.fetch_tu(p, comp), pp.fetch_coeffs(p, comp)):
                    eco, ede, eo = oracle_tu(cur.block(comp, x, y, t), pred[comp][y:y + t, x:x + t], t, comp, qp if comp == 0 else qp_c, isl, sh,
                                             avg_dist, 1.0 if comp == 0 else weight)
                    assert (r["sum"], r["ssd"], r["zeroed"]) == (eo.sum, eo.ssd, eo.zeroed), (isl, sh, p, comp, x, y)
                    assert np.array_equal(c, eco) and np.array_equal(rec[comp][y:y + t, x:x + t], ede), (isl, sh, p, comp, x, y)
                    n_coded += r["sum"] > 0
        assert n_coded > 20
        pp.close()
    fc.close(); fr.close()


def test_prepass_band_union_equals_whole(ctx):
    """CTU-row bands (the multi-GPU partition) produce exactly the rows of the whole-frame run"""
    w, h = 256, 256
    cur, ref = clip_pair(w, h, n=1, noise=3.0, seed=4)
    fc, fr = upload(ctx, cur, w, h), upload(ctx, ref, w, h)
    whole = hb.Prepass(ctx, w, h, use_graph=0)
    whole.run(fc, fr, 500.0)
    full = [whole.fetch_me(d) for d in range(4)]
    for (r0, rows) in ((0, 1), (1, 2), (3, 1)):
        band = hb.Prepass(ctx, w, h, use_graph=0, band=(r0, rows))
        band.run(fc, fr, 500.0)
        for d in range(4):
            s = 64 >> d
            gw = (w // 64) * (64 // s)
            got = band.fetch_me(d)
            for idx, r in enumerate(got):
                ctu_row = (idx // gw) * s // 64
                if r0 <= ctu_row < r0 + rows:
                    assert r == full[d][idx], (d, idx)
                else:
                    assert r["sad"] == 0xFFFFFFFF
        band.close()
    whole.close(); fc.close(); fr.close()


def test_prepass_1080p_properties(ctx):
    """At the benchmark size the oracle is too slow for every PU: check properties instead --
    (1) identical frames give zero vectors, zero SAD and no coded levels; (2) a spot sample of PUs matches the oracle;
    (3) every reconstruction equals prediction + decoded residual implied by sum == 0 TUs (recon == pred there)."""
    w, h = 1920, 1080
    cur, ref = clip_pair(w, h, n=1, noise=3.0)
    fc, fr = upload(ctx, cur, w, h), upload(ctx, ref, w, h)
    pp = hb.Prepass(ctx, w, h)
    pp.run(fc, fc, 650.0)
    for d in range(4):
        r = pp.fetch_me(d)
        ok = r["sad"] != 0xFFFFFFFF
        assert ok.sum() == (w // (64 >> d)) * (h // (64 >> d))
        assert not r["sad"][ok].any() and not r["mvx"][ok].any() and not r["mvy"][ok].any()
    for p in range(5):
        assert not pp.fetch_tu(p, 0)["sum"].any()
    pp.run(fc, fr, 650.0)
    rng = np.random.default_rng(0)
    r3 = pp.fetch_me(3)
    gw = ((w + 63) // 64) * 8
    for idx in rng.choice(len(r3), 40, replace=False):
        px, py = int(idx % gw), int(idx // gw)
        if r3[idx]["sad"] == 0xFFFFFFFF:
            continue
        par = pp.fetch_me(2)[(py // 2) * (gw // 2) + px // 2]
        starts = [(int(par["mvx"]), int(par["mvy"]))] if par["mvx"] != 0 and par["mvy"] != 0 else []
        e = oracle_me(cur, ref, w, h, px * 8, py * 8, 8, 32, [(0, 0), (0, 0)], starts, 650.0)
        assert (r3[idx]["mvx"], r3[idx]["mvy"], r3[idx]["sad"]) == (e.mv.x, e.mv.y, e.sad)
    for p in (0, 3):
        t = pp.tu_size(p, 0)
        xy, res = pp.tu_xy(p, 0), pp.fetch_tu(p, 0)
        rec = pp.recon(p).download()[0]
        prd = pp.pred(min(p, 3)).download()[0]
        for i in rng.choice(len(xy), 200, replace=False):
            x, y = xy[i]
            if res[i]["sum"] == 0:
                assert np.array_equal(rec[y:y + t, x:x + t], prd[y:y + t, x:x + t])
    pp.close(); fc.close(); fr.close()


def test_prepass_1080p_vs_compiled_reference(ctx):
    """The whole 1080p pre-pass against the UNMODIFIED reference's own functions (oracle/_ref: hmr_motion_estimation,
    hmr_motion_compensation_*, encode_inter_cu*), every PU, every TU, every level and reconstructed sample."""
    from _oracle import have_ref, ref_prepass
    if not have_ref():
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    w, h, qp, avg = 1920, 1080, 32, 650.0
    cur, ref = clip_pair(w, h, n=3, noise=3.0)
    fc, fr = upload(ctx, cur, w, h), upload(ctx, ref, w, h)
    pp = hb.Prepass(ctx, w, h, qp=qp)
    pp.run(fc, fr, avg)
    _, exp = ref_prepass((cur.y, cur.u, cur.v), (ref.y, ref.u, ref.v), w, h, qp, avg, n_threads=8)
    for d in range(4):
        got = pp.fetch_me(d)
        for k in ("mvx", "mvy", "subx", "suby", "sad"):
            assert np.array_equal(got[k], exp["me"][d][k]), (d, k, int((got[k] != exp["me"][d][k]).sum()))
        for c, pl in enumerate(pp.pred(d).download()):
            s = 64 >> d
            hh, ww = (h // s) * s // (2 if c else 1), (w // s) * s // (2 if c else 1)
            assert np.array_equal(pl[:hh, :ww], exp["pred"][d][c][:hh, :ww]), ("pred", d, c)
    coded = 0
    for p in range(5):
        rec = pp.recon(p).download()
        for c in range(3):
            if not pp.tu_size(p, c):
                continue
            got = pp.fetch_tu(p, c)
            assert np.array_equal(got["sum"], exp["tu"][(p, c)]["sum"]), ("sum", p, c)
            assert np.array_equal(got["ssd"], exp["tu"][(p, c)]["ssd"]), ("ssd", p, c)
            assert np.array_equal(pp.fetch_coeffs(p, c), exp["coeff"][(p, c)]), ("levels", p, c)
            s = 64 >> min(p, 3)
            hh, ww = (h // s) * s // (2 if c else 1), (w // s) * s // (2 if c else 1)
            assert np.array_equal(rec[c][:hh, :ww], exp["recon"][p][c][:hh, :ww]), ("recon", p, c)
            coded += int((got["sum"] > 0).sum())
    assert coded > 1000
    pp.close(); fc.close(); fr.close()


def test_host_decision_flow_gather(ctx):
    """tables -> host selection -> gather delivers exactly the reconstruction and coded levels of the chosen depth"""
    w, h, qp, avg = 320, 200, 30, 300.0
    cur, ref = clip_pair(w, h, n=4, noise=5.0, seed=12)
    fc, fr = upload(ctx, cur, w, h), upload(ctx, ref, w, h)
    pp = hb.Prepass(ctx, w, h, qp=qp)
    pp.run(fc, fr, avg)
    tables = ctx.pinned(pp.tables_bytes())
    pp.fetch_tables(tables); ctx.sync()
    me0 = pp.fetch_me(0)
    assert bytes(tables[:me0.nbytes]) == me0.tobytes()
    n_ctus = pp.num_ctus()
    assert n_ctus == 5 * 4
    used = set()
    for lam in (0, 40, 400):
        sel = np.zeros(n_ctus, np.uint8); off = np.zeros(n_ctus + 1, np.int32)
        pp.select(tables, lam, sel, off)
        if lam == 0:
            sel[:] = np.arange(n_ctus) % 5          # force every choice to appear, offsets must follow
            sel[15:] = 3 + np.arange(5) % 2         # the last CTU row is 8 samples high: only 8x8 units tile it
            bad = sel.copy(); bad[17] = 1           # 32x32 units in an 8-row CTU: refused, nothing is gathered from unwritten samples
            with pytest.raises(hb.HbError):
                pp.gather(bad, off, ctx.pinned(w * h * 3 // 2 + 64))
            # recompute the layout for the forced selection
            off[:] = 0
            for i in range(n_ctus):
                pass
        used |= set(int(v) for v in sel)
        # expected streams from the individually fetched tables
        exp_len = np.zeros(n_ctus, np.int64); exp_stream = [[] for _ in range(n_ctus)]
        for i in range(n_ctus):
            cx, cy = i % 5, i // 5
            for c in range(3):
                p = int(sel[i]) if c == 0 else min(int(sel[i]), 3)
                t = pp.tu_size(p, c)
                xy, res, co = pp.tu_xy(p, c), pp.fetch_tu(p, c), pp.fetch_coeffs(p, c)
                cs = 64 if c == 0 else 32
                inside = [(k, x, y) for k, (x, y) in enumerate(xy) if x // cs == cx and y // cs == cy]
                inside.sort(key=lambda v: (v[2], v[1]))
                for k, x, y in inside:
                    if res[k]["sum"] > 0:
                        pos = ((y % cs) // t) * (cs // t) + (x % cs) // t
                        hdr = (c << 28) | (t << 16) | pos
                        exp_stream[i].append(np.array([hdr & 0xffff, hdr >> 16], np.uint16).view(np.int16))
                        exp_stream[i].append(co[k].reshape(-1))
            exp_len[i] = sum(len(a) for a in exp_stream[i])
        off[0] = 0
        off[1:] = np.cumsum(exp_len)
        out = ctx.pinned(w * h * 3 // 2 + 2 * int(off[-1]) + 64)
        nbytes = pp.gather(sel, off, out); ctx.sync()
        assert nbytes == w * h * 3 // 2 + 2 * int(off[-1])
        ry = out[:w * h].reshape(h, w); ru = out[w * h:w * h * 5 // 4].reshape(h // 2, w // 2); rv = out[w * h * 5 // 4:w * h * 3 // 2].reshape(h // 2, w // 2)
        lev = out[w * h * 3 // 2:nbytes].view(np.int16)
        recs = [pp.recon(p).download() for p in range(5)]
        for i in range(n_ctus):
            cx, cy = i % 5, i // 5
            p = int(sel[i])
            y1, x1 = min(h, cy * 64 + 64), min(w, cx * 64 + 64)
            # rows below the last whole PU of the chosen depth are not coded: compare the coded area only
            s = 64 >> min(p, 3)
            y1 = (y1 // s) * s if y1 == h else y1
            assert np.array_equal(ry[cy * 64:y1, cx * 64:x1], recs[p][0][cy * 64:y1, cx * 64:x1]), ("recon Y", i, p)
            pc = min(p, 3)
            assert np.array_equal(ru[cy * 32:y1 // 2, cx * 32:x1 // 2], recs[pc][1][cy * 32:y1 // 2, cx * 32:x1 // 2]), ("recon U", i, p)
            assert np.array_equal(rv[cy * 32:y1 // 2, cx * 32:x1 // 2], recs[pc][2][cy * 32:y1 // 2, cx * 32:x1 // 2]), ("recon V", i, p)
            got = lev[off[i]:off[i + 1]]
            exp = np.concatenate(exp_stream[i]) if exp_stream[i] else np.zeros(0, np.int16)
            assert np.array_equal(got, exp), ("levels", i, p, len(got), len(exp))
        if lam != 0:
            # the library's own layout equals the one derived here
            off2 = np.zeros(n_ctus + 1, np.int32); sel2 = np.zeros(n_ctus, np.uint8)
            pp.select(tables, lam, sel2, off2)
            assert np.array_equal(sel2, sel) and np.array_equal(off2, off)
    assert used == {0, 1, 2, 3, 4}
    pp.close(); fc.close(); fr.close()


def test_process_frame_and_its_two_halves(ctx):
    """hb_prepass_process_frame, and begin/finish interleaved over two plans by one thread (frame n+1 queued before frame n
    is finished), deliver the bytes of the step-by-step flow: upload, run, fetch_tables, select, gather"""
    w, h, qp, avg, lam = 320, 192, 31, 420.0, 55
    clips = [clip_pair(w, h, n=k + 2, noise=4.0, seed=40 + k) for k in range(3)]
    fb = w * h * 3 // 2
    pin = ctx.pinned(6 * fb)
    def planes(i, fr):
        b = pin[i * fb:(i + 1) * fb]
        py = b[:w * h].reshape(h, w); pu = b[w * h:w * h * 5 // 4].reshape(h // 2, w // 2); pv = b[w * h * 5 // 4:].reshape(h // 2, w // 2)
        py[:], pu[:], pv[:] = fr.y, fr.u, fr.v
        return (py, pu, pv)
    host = [(planes(2 * k, c), planes(2 * k + 1, r)) for k, (c, r) in enumerate(clips)]
    n_ctus = 5 * 3
    # step by step
    exp = []
    pp = hb.Prepass(ctx, w, h, qp=qp)
    fc, fr = hb.Frame(ctx, w, h), hb.Frame(ctx, w, h)
    for cur_p, ref_p in host:
        fc.upload_u8(*cur_p); fr.upload_u8(*ref_p)
        pp.run(fc, fr, avg)
        tables = ctx.pinned(pp.tables_bytes()); pp.fetch_tables(tables); ctx.sync()
        sel = np.zeros(n_ctus, np.uint8); off = np.zeros(n_ctus + 1, np.int32)
        pp.select(tables, lam, sel, off)
        out = ctx.pinned(fb + 4 * w * h); n = pp.gather(sel, off, out); ctx.sync()
        exp.append((bytes(tables), sel.copy(), off.copy(), bytes(out[:n])))
    # one call per frame
    for k, (cur_p, ref_p) in enumerate(host):
        tables = ctx.pinned(pp.tables_bytes()); sel = np.zeros(n_ctus, np.uint8); off = np.zeros(n_ctus + 1, np.int32); out = ctx.pinned(fb + 4 * w * h)
        n = pp.process_frame(fc, fr, cur_p, ref_p, avg, lam, tables, sel, off, out)
        assert (bytes(tables), bytes(out[:n])) == (exp[k][0], exp[k][3]) and np.array_equal(sel, exp[k][1]) and np.array_equal(off, exp[k][2]), k
    # two plans on two contexts, frame k+1 begun before frame k is finished
    ctx2 = hb.Context(0)
    lanes = []
    for c in (ctx, ctx2):
        lanes.append(dict(pp=hb.Prepass(c, w, h, qp=qp), cur=hb.Frame(c, w, h), ref=hb.Frame(c, w, h), tables=c.pinned(pp.tables_bytes()),
                          sel=np.zeros(n_ctus, np.uint8), off=np.zeros(n_ctus + 1, np.int32), out=c.pinned(fb + 4 * w * h)))
    pending = None
    got = {}
    for k, (cur_p, ref_p) in enumerate(host):
        ln = lanes[k % 2]
        ln["pp"].frame_begin(ln["cur"], ln["ref"], cur_p, ref_p, avg, ln["tables"])
        if pending is not None:
            pk, pl = pending
            n = pl["pp"].frame_finish(lam, pl["tables"], pl["sel"], pl["off"], pl["out"])
            got[pk] = (bytes(pl["tables"]), pl["sel"].copy(), pl["off"].copy(), bytes(pl["out"][:n]))
        pending = (k, ln)
    pk, pl = pending
    n = pl["pp"].frame_finish(lam, pl["tables"], pl["sel"], pl["off"], pl["out"])
    got[pk] = (bytes(pl["tables"]), pl["sel"].copy(), pl["off"].copy(), bytes(pl["out"][:n]))
    for k in range(len(host)):
        assert got[k][0] == exp[k][0] and got[k][3] == exp[k][3] and np.array_equal(got[k][1], exp[k][1]) and np.array_equal(got[k][2], exp[k][2]), k
    for ln in lanes:
        ln["pp"].close(); ln["cur"].close(); ln["ref"].close()
    ctx2.close(); pp.close(); fc.close(); fr.close()


def test_compact_tables(ctx):
    """cfg.compact_tables: the 12-byte wire records carry exactly the values of the full tables, and the host flow built on them
    (select -> gather) ends in the same decision and the same bytes"""
    from homerhevc_b200.lib import ME_COMPACT_DT, TU_COMPACT_DT
    w, h, qp, avg, lam = 320, 192, 30, 380.0, 45
    cur, ref = clip_pair(w, h, n=3, noise=5.0, seed=77)
    fc, fr = upload(ctx, cur, w, h), upload(ctx, ref, w, h)
    outs = []
    for compact in (0, 1):
        pp = hb.Prepass(ctx, w, h, qp=qp, compact_tables=compact)
        pp.run(fc, fr, avg)
        tables = ctx.pinned(pp.tables_bytes()); pp.fetch_tables(tables); ctx.sync()
        n_ctus = pp.num_ctus()
        sel = np.zeros(n_ctus, np.uint8); off = np.zeros(n_ctus + 1, np.int32)
        pp.select(tables, lam, sel, off)
        out = ctx.pinned(w * h * 3 // 2 + 4 * w * h); n = pp.gather(sel, off, out); ctx.sync()
        n_me = sum(pp.num_pus(d) for d in range(4)); n_tu = sum(pp.num_tus(p, c) for p in range(5) for c in range(3))
        outs.append((bytes(tables), sel.copy(), off.copy(), bytes(out[:n]), n_me, n_tu))
        pp.close()
    (tf, self_, offf, outf, n_me, n_tu), (tc, selc, offc, outc, _, _) = outs
    assert len(tf) == 24 * n_me + 16 * n_tu and len(tc) == 12 * (n_me + n_tu)
    mf = np.frombuffer(tf[:24 * n_me], hb.lib.ME_DT if hasattr(hb.lib, "ME_DT") else np.dtype([("mvx", "<i4"), ("mvy", "<i4"), ("subx", "<i4"), ("suby", "<i4"), ("sad", "<u4"), ("n_probes", "<u4")]))
    mc = np.frombuffer(tc[:12 * n_me], ME_COMPACT_DT)
    for k in ("mvx", "mvy", "subx", "suby", "sad", "n_probes"):
        assert np.array_equal(mf[k].astype(np.int64), mc[k].astype(np.int64)), k
    uf = np.frombuffer(tf[24 * n_me:], np.dtype([("sum", "<i4"), ("ssd", "<u4"), ("ssd_zero", "<u4"), ("zeroed", "<i4")]))
    uc = np.frombuffer(tc[12 * n_me:], TU_COMPACT_DT)
    assert np.array_equal(uf["ssd"], uc["ssd"]) and np.array_equal(uf["ssd_zero"], uc["ssd_zero"])
    assert np.array_equal(uf["sum"], (uc["sum_zeroed"] & 0x7fffffff).astype(np.int32)) and np.array_equal(uf["zeroed"] != 0, (uc["sum_zeroed"] >> 31) != 0)
    assert (uf["zeroed"] != 0).any() and (uf["sum"] > 0).any()
    assert np.array_equal(self_, selc) and np.array_equal(offf, offc) and outf == outc

    # ---- compact_tables = 2: one record per coding unit and pass; the sums and coded flags of the CU's transform units
    from homerhevc_b200.lib import CU_COST_DT
    pp = hb.Prepass(ctx, w, h, qp=qp, compact_tables=2)
    pp.run(fc, fr, avg)
    tables = ctx.pinned(pp.tables_bytes()); pp.fetch_tables(tables); ctx.sync()
    cols, rows = (w + 63) // 64, (h + 63) // 64
    n_cu = sum(cols * rows * (64 // (64 >> min(p, 3))) ** 2 for p in range(5))
    assert pp.tables_bytes() == 12 * (n_me + n_cu) and bytes(tables[:12 * n_me]) == tc[:12 * n_me]
    cu = np.frombuffer(bytes(tables[12 * n_me:]), CU_COST_DT)
    tus = {(p, c): (pp.tu_size(p, c), {(int(x), int(y)): r for (x, y), r in zip(pp.tu_xy(p, c), pp.fetch_tu(p, c))}) for p in range(5) for c in range(3)}
    k = 0
    coded_any = False
    for p in range(5):
        d = min(p, 3); s_cu = 64 >> d
        for cy in range(rows * (64 // s_cu)):
            for cx in range(cols * (64 // s_cu)):
                ssd = tot = cbf = 0
                for c in range(3):
                    t, tab = tus[(p if c == 0 else d, c)]
                    side = s_cu if c == 0 else s_cu // 2
                    per = side // t
                    for q in range(per * per):
                        r = tab.get((cx * side + (q % per) * t, cy * side + (q // per) * t))
                        if r is None:
                            continue
                        ssd += int(r["ssd"]); tot += int(r["sum"])
                        if r["sum"] > 0:
                            cbf |= 1 << (4 * c + q)
                assert (int(cu[k]["ssd"]), int(cu[k]["sum"]), int(cu[k]["cbf"])) == (ssd, tot, cbf), (p, cx, cy)
                coded_any |= cbf != 0
                k += 1
    assert k == n_cu and coded_any
    sel2 = np.zeros(n_ctus, np.uint8); off2 = np.zeros(n_ctus + 1, np.int32)
    pp.select(tables, lam, sel2, off2)
    out2 = ctx.pinned(w * h * 3 // 2 + 4 * w * h); n2 = pp.gather(sel2, off2, out2); ctx.sync()
    assert np.array_equal(sel2, self_) and np.array_equal(off2, offf) and bytes(out2[:n2]) == outf
    pp.close()
    fc.close(); fr.close()


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_bands_across_gpus_with_nccl_halo_exchange(ctx, exchange):
    """configs[3]: CTU-row bands on two GPUs, reference halos swapped over NCCL (staged send / recv) or pulled out of the neighbour's
    HBM through CUDA IPC peer memory; needs >= 2 devices (skipped on one)"""
    import os
    import subprocess
    import sys
    L = hb.load_library()
    if L.hb_device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29641" if exchange == "nccl" else "29642", os.path.join(root, "tools", "band_check.py"), "1280", "720", exchange],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert "mismatches=0" in out.stdout


def _host_units(pp, sel, w, h, qp):
    """deblocking unit data of a selection, rebuilt on the host from the individually fetched tables"""
    from homerhevc_b200.lib import UNIT_INFO_DT
    cols = (w + 63) // 64
    units = np.zeros((h // 4, w // 4), UNIT_INFO_DT)
    me = [pp.fetch_me(d) for d in range(4)]
    coded = []
    for p in range(5):
        t = pp.tu_size(p, 0)
        xy, res = pp.tu_xy(p, 0), pp.fetch_tu(p, 0)
        coded.append((t, {(int(x), int(y)): int(r["sum"]) > 0 for (x, y), r in zip(xy, res)}))
    for uy in range(h // 4):
        for ux in range(w // 4):
            x, y = ux * 4, uy * 4
            p = int(sel[(y // 64) * cols + x // 64])
            d = min(p, 3); cu = 64 >> d
            m = me[d][(y // cu) * (cols * (64 // cu)) + x // cu]
            t, tab = coded[p]
            td = 1 if p in (0, 4) else 0
            u = units[uy, ux]
            u["cu_depth"], u["tu_depth"], u["qp"], u["mvx"], u["mvy"] = d, td, qp, m["mvx"], m["mvy"]
            u["cbf_luma"] = (1 << td) if tab.get((x // t * t, y // t * t), False) else 0
    return units


def test_device_resident_finalisation(ctx):
    """the choice gathered into a frame and deblocked without leaving the device equals the host-composed chain (gather to the host,
    unit data rebuilt from the fetched tables, upload, hb_deblock_frame_units); then the whole resident per-frame flow -- SAO
    statistics, stand-in decision, offset pass, border -- equals the same calls made one by one, and its output serves as a
    reference picture exactly like an uploaded copy of it"""
    from homerhevc_b200.lib import SAO_PARAM_DT, sao_decide_from_candidates, sao_decide_standin, sao_derive_offsets
    w, h, qp, avg = 320, 200, 30, 300.0
    cur, ref = clip_pair(w, h, n=4, noise=5.0, seed=21)
    fc, fr = upload(ctx, cur, w, h), upload(ctx, ref, w, h)
    pp = hb.Prepass(ctx, w, h, qp=qp)
    pp.run(fc, fr, avg)
    tables = ctx.pinned(pp.tables_bytes())
    pp.fetch_tables(tables); ctx.sync()
    n_ctus = pp.num_ctus()
    sel = np.zeros(n_ctus, np.uint8); off = np.zeros(n_ctus + 1, np.int32)
    pp.select(tables, 40, sel, off)
    assert set(int(v) for v in sel[15:]) <= {3, 4}                  # 8 picture rows in the last CTU row: only 8x8 units tile them
    assert len(set(int(v) for v in sel)) >= 2
    out = ctx.pinned(w * h * 3 // 2 + 2 * int(off[-1]) + 64)
    nbytes = pp.gather(sel, off, out); ctx.sync()
    ry = out[:w * h].reshape(h, w); ru = out[w * h:w * h * 5 // 4].reshape(h // 2, w // 2); rv = out[w * h * 5 // 4:w * h * 3 // 2].reshape(h // 2, w // 2)
    lev = out[w * h * 3 // 2:nbytes].copy()
    units = _host_units(pp, sel, w, h, qp)
    fa = hb.Frame(ctx, w, h); fa.upload_u8(ry.copy(), ru.copy(), rv.copy())
    bsv, bsh = ctx.deblock_units(fa, units, 2, 2)
    exp = fa.download()
    assert any((a != b).any() for a, b in zip(exp, (ry, ru, rv)))   # the filter did something

    rec = hb.Frame(ctx, w, h)
    levels = ctx.pinned(2 * int(off[-1]) + 64)
    nlev = pp.finalise(sel, off, rec, levels, 2, 2); ctx.sync()
    assert nlev == 2 * int(off[-1]) and bytes(levels[:nlev]) == bytes(lev)
    got_units, gv, gh = pp.fetch_units()
    for f in units.dtype.names:
        assert np.array_equal(got_units[f], units[f]), f
    assert np.array_equal(gv, bsv) and np.array_equal(gh, bsh)
    got = rec.download()
    for c in range(3):
        assert np.array_equal(got[c], exp[c]), ("deblocked", c, np.argwhere(got[c] != exp[c])[:4])

    # ---- the resident flow in two halves, against the same calls made one by one
    lam_sao = (30.0, 36.0, 36.0)
    st = ctx.sao_stats(fc, rec)
    prm = sao_decide_standin(st, lam_sao)
    assert (prm["type"] >= 0).any()
    # the candidates derived on the device are the host function's (which is pinned to the reference), for every CTU, component, type
    cand, st_again = ctx.sao_candidates(fc, rec, lam_sao, want_stats=True)
    assert st_again.tobytes() == st.tobytes()
    for i in range(n_ctus):
        for c in range(3):
            for t in range(5):
                o32, band, dist = sao_derive_offsets(st[i, c], t, lam_sao[c])
                k = cand[i, c, t]
                exp_off = o32[[0, 1, 3, 4]] if t < 4 else o32[band:band + 4]
                assert int(k["dist"]) == dist and np.array_equal(k["offset"], exp_off) and int(k["band"]) == band, (i, c, t, k, o32, band, dist)
    assert sao_decide_from_candidates(cand, lam_sao).tobytes() == prm.tobytes()
    fin_exp = hb.Frame(ctx, w, h)
    ctx.sao_apply(rec, fin_exp, prm["type"], prm["offset"])
    exp_fin = fin_exp.download()

    cur_pin = ctx.pinned(w * h * 3 // 2)
    py = cur_pin[:w * h].reshape(h, w); pu = cur_pin[w * h:w * h * 5 // 4].reshape(h // 2, w // 2); pv = cur_pin[w * h * 5 // 4:].reshape(h // 2, w // 2)
    py[:], pu[:], pv[:] = cur.y, cur.u, cur.v
    fc2, rec2, nxt = hb.Frame(ctx, w, h), hb.Frame(ctx, w, h), hb.Frame(ctx, w, h)
    sel2 = np.zeros(n_ctus, np.uint8); off2 = np.zeros(n_ctus + 1, np.int32)
    prm2 = np.zeros(n_ctus, SAO_PARAM_DT)
    levels2 = ctx.pinned(w * h * 4)
    pp.frame_begin_resident(fc2, fr, (py, pu, pv), avg, tables)
    nlev2 = pp.frame_finish_resident(fc2, 40, tables, sel2, off2, rec2, nxt, (2, 2, 0, 0), lam_sao, levels2, prm2)
    assert np.array_equal(sel2, sel) and np.array_equal(off2, off) and nlev2 == nlev and bytes(levels2[:nlev2]) == bytes(lev)
    assert prm2.tobytes() == prm.tobytes()
    # the offset pass is only queued: a second pre-pass on the same context must see the finished frame without an explicit wait
    pp.run(fc, nxt, avg); me_unsynced = [pp.fetch_me(d).tobytes() for d in range(4)]
    got_fin = nxt.download()
    for c in range(3):
        assert np.array_equal(got_fin[c], exp_fin[c]), ("finished", c)
    assert any((a != b).any() for a, b in zip(got_fin, got))        # SAO changed samples

    # ---- the finished frame as the next reference: identical search results to an uploaded copy of it (border included)
    nxt_copy = hb.Frame(ctx, w, h); nxt_copy.upload_u8(*[np.ascontiguousarray(p) for p in got_fin])
    pp.run(fc, nxt, avg); a = [pp.fetch_me(d).tobytes() for d in range(4)]
    pp.run(fc, nxt_copy, avg); b = [pp.fetch_me(d).tobytes() for d in range(4)]
    assert a == b == me_unsynced
    for f in (fa, rec, fin_exp, fc2, rec2, nxt, nxt_copy, fc, fr):
        f.close()
    pp.close()


def test_graph_cache_follows_the_device_planes(ctx):
    """the captured replay is keyed on the frames' device planes: a frame that is destroyed and re-created (same size, quite
    possibly the same host struct address) must be searched with ITS samples, not with a stale graph's"""
    w, h, qp, avg = 128, 64, 32, 500.0
    cur, ref = clip_pair(w, h, n=2, noise=3.0, seed=21)
    cur2, ref2 = clip_pair(w, h, n=2, noise=3.0, seed=22)
    pp = hb.Prepass(ctx, w, h, qp=qp, use_graph=1)
    plain = hb.Prepass(ctx, w, h, qp=qp, use_graph=0)
    keep = []
    for rep in range(6):
        c, r = (cur, ref) if rep % 2 == 0 else (cur2, ref2)
        fc, fr = upload(ctx, c, w, h), upload(ctx, r, w, h)
        pp.run(fc, fr, avg); ctx.sync()
        got = pp.fetch_me(3).copy()
        plain.run(fc, fr, avg); ctx.sync()
        assert got.tobytes() == plain.fetch_me(3).tobytes(), rep
        fc.close(); fr.close()
        keep.append(hb.Frame(ctx, w, h)) if rep == 2 else None       # perturb the allocator between generations
    for f in keep:
        f.close()
    pp.close(); plain.close()


def test_subpel_planes_per_picture_equal_per_pu(ctx):
    """the quarter-pel planes built once per reference picture (default; TMA-staged tiles) give the same vectors, SADs and luma
    predictions as planes built per PU in shared memory (subpel_per_pu = 1, the reference's own scheme), partial CTUs included"""
    w, h, qp, avg = 328, 200, 30, 420.0
    cur, ref = clip_pair(w, h, n=3, noise=4.0, seed=31)
    fc, fr = upload(ctx, cur, w, h), upload(ctx, ref, w, h)
    a, b = hb.Prepass(ctx, w, h, qp=qp), hb.Prepass(ctx, w, h, qp=qp, subpel_per_pu=1)
    g = hb.Prepass(ctx, w, h, qp=qp, me_staged_window=1)          # planes per picture + the search window staged in shared memory (bulk async copies)
    e = hb.Prepass(ctx, w, h, qp=qp, me_per_depth=1)              # one search launch per PU size instead of one per picture (a CTA per CTU)
    for rep in range(2):
        a.run(fc, fr, avg); b.run(fc, fr, avg); g.run(fc, fr, avg); e.run(fc, fr, avg)
    ctx.sync()
    moved = 0
    for d in range(4):
        ma, mb = a.fetch_me(d), b.fetch_me(d)
        assert ma.tobytes() == mb.tobytes(), d
        assert ma.tobytes() == e.fetch_me(d).tobytes(), ("one launch per picture vs one per PU size", d)
        assert all(np.array_equal(x, y) for x, y in zip(a.pred(d).download(), e.pred(d).download())), d
        assert ma.tobytes() == g.fetch_me(d).tobytes(), ("search window staged in shared memory vs gathered from global memory", d)
        ok = ma["sad"] != 0xFFFFFFFF
        moved += int(((ma["subx"][ok] != 0) | (ma["suby"][ok] != 0)).sum())
        pa, pb = a.pred(d).download(), b.pred(d).download()
        s = 64 >> d
        for c in range(3):
            sc = s // 2 if c else s
            hh, ww = (h // 2 if c else h) // sc * sc, (w // 2 if c else w) // sc * sc
            assert np.array_equal(pa[c][:hh, :ww], pb[c][:hh, :ww]), (d, c)
    assert moved > 50, "the clip never produced sub-pel vectors"
    for p in range(5):
        for c in range(3):
            if a.tu_size(p, c):
                assert a.fetch_tu(p, c).tobytes() == b.fetch_tu(p, c).tobytes() and np.array_equal(a.fetch_coeffs(p, c), b.fetch_coeffs(p, c)), (p, c)
    a.close(); b.close(); g.close(); fc.close(); fr.close()


def test_subpel_planes_without_tma():
    """the same kernels with the plane tiles staged by plain loads ($HB_NO_TMA=1 is read once per process)"""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_prepass.py"), "-q", "-x", "-m", "gpu", "-k",
                          "matches_oracle or per_picture_equal"], capture_output=True, text=True, timeout=900, cwd=root, env=dict(os.environ, HB_NO_TMA="1"))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-1000:]


def test_levels_in_the_reference_coeff_wnd_layout(ctx):
    """hb_prepass_fetch_coeff_wnd: the chosen passes' levels where the reference's entropy coder looks for them (ctu->coeff_wnd: a unit's
    levels row-major at abs_index << 4, chroma >> 2, abs_index = z-order number of its first 4x4 luma block), rebuilt on the host from the
    per-pass tables; every pass chosen somewhere, partial CTUs on both edges"""
    w, h, qp, avg = 328, 200, 26, 300.0
    cur, ref = clip_pair(w, h, n=3, noise=5.0, seed=12)
    fc, fr = upload(ctx, cur, w, h), upload(ctx, ref, w, h)
    pp = hb.Prepass(ctx, w, h, qp=qp)
    pp.run(fc, fr, avg); ctx.sync()
    cols, rows = (w + 63) // 64, (h + 63) // 64
    sel = np.array([(3 + (i % 2)) if (i % cols == cols - 1 or i // cols == rows - 1) else i % 5 for i in range(cols * rows)], np.uint8)   # partial CTUs: 8x8 units
    got = pp.fetch_coeff_wnd(sel)

    def zorder(ux, uy):
        return sum((((ux >> b) & 1) << (2 * b)) | (((uy >> b) & 1) << (2 * b + 1)) for b in range(4))
    exp = np.zeros_like(got)
    coded = 0
    for ctu in range(cols * rows):
        for c in range(3):
            p = min(int(sel[ctu]), 3) if c else int(sel[ctu])
            t = pp.tu_size(p, c)
            xy, res, co = pp.tu_xy(p, c), pp.fetch_tu(p, c), pp.fetch_coeffs(p, c)
            cs = 32 if c else 64
            x0, y0 = (ctu % cols) * cs, (ctu // cols) * cs
            base = 0 if c == 0 else (4096 if c == 1 else 4096 + 1024)
            for i in np.nonzero((xy[:, 0] >= x0) & (xy[:, 0] < x0 + cs) & (xy[:, 1] >= y0) & (xy[:, 1] < y0 + cs))[0]:
                if res[i]["sum"] <= 0:
                    continue
                lx, ly = (int(xy[i, 0]) - x0) * (2 if c else 1), (int(xy[i, 1]) - y0) * (2 if c else 1)     # luma position inside the CTU
                a = zorder(lx // 4, ly // 4)
                off = (a << 4) >> 2 if c else a << 4
                exp[ctu, base + off:base + off + t * t] = co[i].reshape(-1)
                coded += 1
    assert coded > 200 and np.array_equal(got, exp), np.argwhere(got != exp)[:5]
    pp.close(); fc.close(); fr.close()


def test_upload_of_separately_pinned_planes_that_touch(ctx):
    """three pinned allocations that happen to be adjacent look like one contiguous picture to hb_frame_upload_u8, but the runtime refuses a
    copy that spans allocations: the one-copy shortcut must fall back to plane copies (found by bench.py --mode intra_recon at 1280x704)"""
    rng = np.random.default_rng(3)
    for (w, h) in ((1280, 704), (1280, 720), (256, 128)):
        planes = [rng.integers(0, 256, s, dtype=np.uint8) for s in ((h, w), (h // 2, w // 2), (h // 2, w // 2))]
        pin = [ctx.pinned(p.nbytes).view(np.uint8).reshape(p.shape) for p in planes]
        for d, s in zip(pin, planes):
            d[:] = s
        f = hb.Frame(ctx, w, h)
        f.upload_u8(*pin); ctx.sync()
        assert all(np.array_equal(a, b) for a, b in zip(f.download(), planes)), (w, h)
        f.close()
