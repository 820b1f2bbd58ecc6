"""Intra prediction on the GPU against the oracle: all-mode SADs of the mode search, predictions written
into a frame and fed to the intra T/Q chain, and the per-call forms of the two table members."""
import ctypes as C

import numpy as np
import pytest

import homerhevc_b200 as hb
from _frames import clip_pair, upload
from _oracle import aligned_i16, oracle, ptr

pytestmark = pytest.mark.gpu
W, H = 320, 192


def _adi_from_frame(plane, x, y, n, rng):
    """reference samples of a block taken from an already 'reconstructed' plane (here: the frame itself), 128 where missing"""
    adi = np.full(4 * n + 1, 128, np.int16)
    h, w = plane.shape
    for i in range(0, 2 * n + 1):                      # corner + above
        xx, yy = x - 1 + i, y - 1
        if 0 <= xx < w and yy >= 0:
            adi[2 * n + i] = plane[yy, xx]
    for i in range(1, 2 * n + 1):                      # left
        xx, yy = x - 1, y - 1 + i
        if xx >= 0 and 0 <= yy < h:
            adi[2 * n - i] = plane[yy, xx]
    return adi


def test_all_mode_sads_and_predictions(ctx):
    O = oracle()
    rng = np.random.default_rng(31)
    cur, _ = clip_pair(W, H, n=1, noise=3.0, seed=19)
    fc = upload(ctx, cur, W, H)
    pred = hb.Frame(ctx, W, H)
    jobs, adis, meta = [], [], []
    for n, y0 in ((32, 0), (16, 64), (8, 96), (4, 112)):
        for k in range(12):
            x = int(rng.integers(0, (W - n) // n + 1)) * n; y = y0 + int(rng.integers(0, 32 // n if n < 32 else 1)) * n
            if n == 32:
                y = int(rng.integers(0, 2)) * 32
            adi = _adi_from_frame(cur.y, x, y, n, rng)
            if k % 4 == 3:
                adi[:] = np.clip(np.linspace(40, 200, adi.size) + rng.integers(-1, 2, adi.size), 0, 255)     # smooth -> strong filter at 32
            jobs.append(hb.IntraJob(0, x, y, n, -1, -1)); adis.append(adi); meta.append((x, y, n))
    adi_all = np.concatenate(adis).astype(np.int16)
    sads = ctx.intra_run(fc, None, jobs, adi_all)
    for (x, y, n), adi, got in zip(meta, adis, sads):
        exp = (C.c_uint32 * 35)()
        blk = np.ascontiguousarray(cur.y[y:y + n, x:x + n].astype(np.int16)).reshape(-1)
        a = np.ascontiguousarray(adi)
        O.orc_intra_mode_sads(ptr(blk), n, ptr(a), n, exp)
        assert list(got) == list(exp), (x, y, n, [int(m) for m in np.nonzero(np.array(got) != np.array(exp))[0]])
    # predictions into a frame: luma with the search rule, forced raw / smoothed, and chroma (always raw, no edge filters)
    pjobs, padis, pmeta = [], [], []
    for comp, n, yb in ((0, 32, 0), (0, 16, 32), (0, 8, 48), (0, 4, 56), (1, 16, 0), (2, 8, 16), (1, 4, 24)):
        pw = W if comp == 0 else W // 2
        plane = (cur.y, cur.u, cur.v)[comp]
        xs = list(range(0, pw - n + 1, n))
        for mode in range(35):
            x = xs[mode % len(xs)]; y = yb + (mode // len(xs)) * n
            if y + n > (H if comp == 0 else H // 2):
                continue
            adi = _adi_from_frame(plane, x, y, n, rng)
            flt = [-1, 0, 1][mode % 3] if comp == 0 else -1
            pjobs.append(hb.IntraJob(comp, x, y, n, mode, flt)); padis.append(adi); pmeta.append((comp, x, y, n, mode, flt))
    # later jobs may overwrite earlier ones where bands overlap: check each job right after its own launch batch
    for comp_sel in (0, 1, 2):
        sel = [i for i, m in enumerate(pmeta) if m[0] == comp_sel]
        # de-overlap: keep jobs whose rectangles are disjoint
        taken, keep = [], []
        for i in sel:
            _, x, y, n, _, _ = pmeta[i]
            if all(x + n <= tx or tx + tn <= x or y + n <= ty or ty + tn <= y for (tx, ty, tn) in taken):
                taken.append((x, y, n)); keep.append(i)
        ctx.intra_run(None, pred, [pjobs[i] for i in keep], np.concatenate([padis[i] for i in keep]).astype(np.int16))
        got = pred.download()[comp_sel]
        for i in keep:
            comp, x, y, n, mode, flt = pmeta[i]
            a = np.ascontiguousarray(padis[i]); f = np.zeros(4 * n + 1, np.int16)
            O.orc_adi_filter(ptr(a), ptr(f), n, 1)
            use_f = comp == 0 and (O.orc_intra_uses_filtered(n, mode) if flt < 0 else flt)
            exp = np.zeros(n * n, np.int16)
            O.orc_intra_predict(ptr(f if use_f else a), n, mode, int(comp == 0), ptr(exp), n)
            assert np.array_equal(got[y:y + n, x:x + n], exp.reshape(n, n).astype(np.uint8)), (comp, x, y, n, mode, flt)
    fc.close(); pred.close()


def test_percall_intra_predictors(ctx):
    O = oracle()
    L = hb.load_library()
    rng = np.random.default_rng(32)
    for it in range(60):
        n = int(rng.choice([4, 8, 16, 32]))
        adi = aligned_i16(4 * n + 1); adi[:] = rng.integers(0, 256, adi.size)
        for mode in (0, 1, 2, 9, 10, 11, 18, 25, 26, 27, 34, int(rng.integers(2, 35))):
            is_luma = int(rng.integers(0, 2))
            p1 = np.zeros(64 * 64, np.int16); p2 = np.zeros(n * n, np.int16)
            if mode == 0:
                L.hb_create_intra_planar_prediction(ptr(p1), 64, ptr(adi), 4 * n + 1, n, n.bit_length() - 1)
            else:
                L.hb_create_intra_angular_prediction(ptr(p1), 64, ptr(adi), 4 * n + 1, n, mode, is_luma)
            O.orc_intra_predict(ptr(adi), n, mode, is_luma if mode else 1, ptr(p2), n)
            assert np.array_equal(p1.reshape(64, 64)[:n, :n], p2.reshape(n, n)), (n, mode, is_luma)


def test_presearch_against_reference_functions(ctx):
    """every luma block of a small picture, sizes 4..32, reference samples from the original picture: the 35 SADs per block
    equal what the reference's own predictors + sad deliver (oracle/_ref), and the oracle agrees"""
    from _oracle import have_ref, ref_intra_presearch
    from homerhevc_b200.intra_jobs import presearch_jobs
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    w, h = 192, 128
    cur, _ = clip_pair(w, h, n=1, noise=6.0, seed=23)
    jobs, adi, off = presearch_jobs(cur.y)
    assert len(jobs) == sum((w // n) * (h // n) for n in (32, 16, 8, 4))
    fc = upload(ctx, cur, w, h)
    got = ctx.intra_presearch(fc, jobs, adi)
    assert np.array_equal(got, ctx.intra_presearch(fc, hb.lib.presearch_records(jobs), adi, ctx.pinned(len(jobs) * 140).view(np.uint32).reshape(-1, 35)))
    _, exp = ref_intra_presearch(cur.y, jobs, adi, off, n_threads=2)
    bad = np.argwhere(got != exp)
    assert len(bad) == 0, (len(bad), bad[:4], jobs[bad[0][0]])
    fc.close()


def test_intra_picture_reconstruction_wavefront(ctx):
    """hb_intra_reconstruct: whole I pictures rebuilt on the device, dependency level by dependency level, from the decisions the reference's own
    encoder made == the encoder's unfiltered reconstruction (all three planes) and the oracle's levels / sums / distortions unit by unit.
    Partial CTUs on both picture edges, several QPs, sign hiding on and off; the argument checks"""
    from homerhevc_b200 import synth
    from homerhevc_b200.lib import HbError
    from _intra import capture_intra_picture, intra_tus, oracle_intra_recon
    from _oracle import have_ref
    if not have_ref():
        pytest.skip("needs the compiled reference (oracle/_ref)")
    for (w, h, qp, sh, seed) in ((192, 136, 32, 1, 21), (328, 200, 38, 0, 9), (1280, 720, 30, 1, 1234)):
        clip = synth.make_clip(w, h, 1, seed=seed)
        a = capture_intra_picture(w, h, qp, sh, seed)
        tus = intra_tus(a, w, h)
        cur, pred, rec = hb.Frame(ctx, w, h), hb.Frame(ctx, w, h), hb.Frame(ctx, w, h)
        cur.upload_u8(*clip[0])
        for per_level in (1, 0, 0):          # a batch of launches per dependency level; one persistent launch (twice: it must leave no state behind)
            coeffs, res, levels = ctx.intra_reconstruct(cur, pred, rec, tus, 1, sh, 1.0, per_level_launches=per_level)
            got = rec.download()
            for c in range(3):
                bad = np.argwhere(got[c] != a["recon"][c])
                assert not len(bad), (w, h, c, per_level, len(bad), bad[:3].tolist())
            if per_level:
                first = (coeffs.copy(), res.copy())
                rec.upload_u8(*[np.full_like(p, 77) for p in clip[0]])      # the persistent form must not find the answer already there
            else:
                assert np.array_equal(coeffs, first[0]) and res.tobytes() == first[1].tobytes(), per_level
        orec, ocoeff, ores = oracle_intra_recon(clip[0], w, h, tus, 1, sh, 1.0)
        assert np.array_equal(coeffs, ocoeff)
        assert [int(r["sum"]) for r in res] == [r.sum for r in ores] and [int(r["ssd"]) for r in res] == [r.ssd for r in ores]
        assert 2 < levels < len(tus), levels
        print(f"\nintra reconstruction {w}x{h}: {len(tus)} units in {levels} dependency levels")
        for f in (cur, pred, rec):
            f.close()
    cur, pred, rec = hb.Frame(ctx, 64, 64), hb.Frame(ctx, 64, 64), hb.Frame(ctx, 64, 64)
    for bad_unit in ([0, 4, 0, 8, 1, 30, 3, 4, 0, 8], [1, 0, 0, 32, 1, 30, 3, 0, 0, 64], [0, 0, 0, 8, 35, 30, 3, 0, 0, 8], [0, 0, 0, 16, 1, 30, 1, 0, 0, 16]):
        with pytest.raises(HbError):
            ctx.intra_reconstruct(cur, pred, rec, np.array([bad_unit], np.int32))
    for f in (cur, pred, rec):
        f.close()
