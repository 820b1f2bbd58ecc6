"""The pin of the oracle on the compiled reference, run again INSIDE the GPU lease: the GPU parity tests compare the kernels with the
restatement (oracle/hb_oracle.c); that is only worth something on a box where the restatement has just been checked against the
unmodified reference (oracle/_ref, built from /root/reference by oracle/Makefile and shipped with the tree).  Same functions as the
CPU suite's tests/test_oracle_vs_ref.py -- the per-function pins the GPU comparisons lean on."""
import pytest

import test_oracle_vs_ref as pin
from _oracle import have_ref

pytestmark = pytest.mark.gpu

PINS = [pin.test_sad_ssd_sse_lane_arithmetic_on_the_full_int16_range, pin.test_pixel_interp_transform_random,
        pin.test_quant_random_and_c_vs_sse_difference, pin.test_motion_estimation_random, pin.test_intra_tq_chain_against_reference_calls,
        pin.test_intra_prediction_against_reference, pin.test_bi_prediction_mc_against_reference, pin.test_sao_statistics_against_reference,
        pin.test_deblocking_pixel_stage_against_reference, pin.test_amvp_candidates_against_reference, pin.test_merge_candidates_against_reference,
        pin.test_boundary_strengths_of_b_pictures_against_reference]


@pytest.mark.parametrize("fn", PINS, ids=[f.__name__[5:] for f in PINS])
def test_oracle_is_pinned_on_this_box(fn):
    if not have_ref():
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    fn()
