#!/usr/bin/env python
"""Generate tests/golden/ref_vectors_r02.npz: outputs of the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/Makefile) for the functions restated in round 2 -- AMVP and merge candidate lists, the boundary strengths of B pictures, the SAO
offset pass, and one whole I picture as the reference ENCODER decided and reconstructed it (oracle/_ref/zinit, captured through
oracle/ref_hooks.c).  tests/test_oracle_golden.py replays the inputs through oracle/liboracle.so and demands identical outputs, here and
where /root/reference does not exist.

    python tests/golden/make_golden_r02.py          # needs /root/reference (build container only)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
from _oracle import (amvp_jobs, random_b_motion, random_deblock_case, random_sao_params, ref_amvp, ref_deblock_strengths_b, ref_merge,  # noqa: E402
                     ref_sao_apply)
from _intra import capture_intra_picture, intra_tus  # noqa: E402

UNIT_KEYS = ("cu", "tu", "intra", "cbf", "qp", "mv")


def main():
    rng = np.random.default_rng(20261017)
    g = {}
    # ---- 8f item 3: get_amvp_candidates / get_merge_mvp_candidates (hmr_motion_inter.c:2342 / :1937) on a random CU tree with partial CTUs
    w, h = 200, 136
    m, _ = random_deblock_case(rng, w, h)
    m["mv"] = rng.integers(-6, 7, m["mv"].shape).astype(np.int16)
    jobs = amvp_jobs(w, h)
    for k in UNIT_KEYS:
        g["cand_" + k] = m[k]
    g["cand_jobs"] = jobs
    g["cand_amvp"] = ref_amvp(w, h, m, jobs)
    g["cand_merge5"] = ref_merge(w, h, m, jobs, 5)
    g["cand_merge2"] = ref_merge(w, h, m, jobs, 2)
    # ---- 8f item 4: boundary strengths of a B picture (get_boundary_strength_single, hmr_deblocking_filter.c:173-229)
    w, h = 192, 136
    m, _ = random_deblock_case(rng, w, h)
    m["cbf"] = (m["cbf"] * (rng.random(m["cbf"].shape) < 0.25)).astype(np.uint8)
    mot = random_b_motion(rng, m, 2, 2)
    bsv, bsh = ref_deblock_strengths_b(w, h, m, *mot)
    for k in UNIT_KEYS:
        g["bsb_" + k] = m[k]
    for k, v in zip(("ref0", "mv0", "ref1", "mv1", "pic_l0", "pic_l1"), mot):
        g["bsb_" + k] = v
    g["bsb_ver"], g["bsb_hor"] = bsv, bsh
    # ---- 8f item 4: SAO offset pass (sao_offset_ctu / offset_block, hmr_sao.c:1210 / :960) with every type incl. off, partial CTUs
    w, h = 200, 136
    planes = [np.clip(rng.normal(128, 40, (hh, ww)), 0, 255).astype(np.uint8) for (ww, hh) in ((w, h), (w // 2, h // 2), (w // 2, h // 2))]
    types, offs = random_sao_params(rng, w, h)
    out = ref_sao_apply(planes, w, h, types, offs)
    g["saoa_in"] = np.concatenate([p.reshape(-1) for p in planes]); g["saoa_out"] = np.concatenate([p.reshape(-1) for p in out])
    g["saoa_types"], g["saoa_offs"] = types, offs
    # ---- 8f item 1 / configs[0]: one I picture of the reference encoder: its decisions per transform unit and its reconstruction BEFORE
    # the in-loop filters (the source picture is homerhevc_b200.synth.make_clip(192, 136, 1, seed=21)[0], stored with the vectors)
    from homerhevc_b200 import synth
    w, h, qp, sh, seed = 192, 136, 32, 1, 21
    a = capture_intra_picture(w, h, qp, sh, seed)
    g["ipic_src"] = np.concatenate([p.reshape(-1) for p in synth.make_clip(w, h, 1, seed=seed)[0]])
    g["ipic_cfg"] = np.array([w, h, qp, sh, seed], np.int32)
    g["ipic_tus"] = intra_tus(a, w, h)
    g["ipic_recon"] = np.concatenate([p.reshape(-1) for p in a["recon"]])
    g["ipic_coeff"] = a["coeff"]
    out_path = os.path.join(HERE, "ref_vectors_r02.npz")
    np.savez_compressed(out_path, **g)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main()
