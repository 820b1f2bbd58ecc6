"""The drop-in boundary: libhomer_b200.so builds for sm_100a without a GPU, loads, and exports every function that
include/homer_b200.h declares; and the host-side mirror refuses to compute without the CUDA library (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

import homerhevc_b200 as hb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "homer_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", src)) - {"hb_low_level_funcs"})


def test_library_builds_and_exports_every_declared_symbol():
    path = hb.build_library()
    assert os.path.exists(path)
    L = C.CDLL(path)
    names = _declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_sm100a_code_is_embedded():
    out = subprocess.run(["cuobjdump", "--list-elf", hb.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts_match_header():
    # sizes the C side relies on (hb_me_job 17 int32, hb_me_result 6, hb_tu_result 4, 19 function pointers)
    assert C.sizeof(hb.MeJob) == 68 and C.sizeof(hb.MeResult) == 24 and C.sizeof(hb.TuResult) == 16
    assert C.sizeof(hb.McJob) == 20 and C.sizeof(hb.TuJob) == 20 and C.sizeof(hb.LowLevelFuncs) == 19 * C.sizeof(C.c_void_p)
    assert C.sizeof(hb.PrepassCfg) == 48 and C.sizeof(hb.TqParams) == 24
    assert hb.lib.ME_COMPACT_DT.itemsize == 12 and hb.lib.TU_COMPACT_DT.itemsize == 12
    assert hb.lib.SAO_DT.itemsize == 416 and hb.lib.SAO_PARAM_DT.itemsize == 196 and hb.lib.UNIT_INFO_DT.itemsize == 10


def test_no_cpu_fallback_without_a_device():
    L = hb.load_library()
    if L.hb_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(hb.HbError) as e:
        hb.Context(0)
    assert "no CUDA device" in str(e.value)


def test_product_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's baseline legs may use oracle/"""
    pkg = os.path.join(ROOT, "homerhevc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle/" not in txt.replace("no oracle", "") and "hb_oracle" not in txt and "liboracle" not in txt, os.path.join(dirpath, f)


def test_struct_sizes_match_the_c_header(tmp_path):
    """the ctypes / numpy mirrors have the sizes a C compiler gives the structs of include/homer_b200.h"""
    import subprocess
    names = ["hb_me_job", "hb_me_result", "hb_mc_job", "hb_mc_bi_job", "hb_tu_job", "hb_tu_result", "hb_tq_params", "hb_prepass_cfg", "hb_me_result_c",
             "hb_tu_result_c", "hb_sao_stats", "hb_sao_param", "hb_sao_candidate", "hb_unit_info", "hb_deblock_params", "hb_low_level_funcs", "hb_frame_ipc", "hb_row_span", "hb_amvp_job", "hb_amvp_list", "hb_unit_l1", "hb_intra_unit"]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "homer_b200.h"\nint main(void) {\n' +
                   "".join(f'    printf("{n} %zu\\n", sizeof({n}));\n' for n in names) + "    return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    mirror = {"hb_frame_ipc": C.sizeof(hb.FrameIpc), "hb_row_span": C.sizeof(hb.RowSpan), "hb_amvp_job": 12, "hb_amvp_list": 16, "hb_unit_l1": 6, "hb_intra_unit": 40, "hb_me_job": C.sizeof(hb.MeJob), "hb_me_result": C.sizeof(hb.MeResult), "hb_mc_job": C.sizeof(hb.McJob), "hb_mc_bi_job": C.sizeof(hb.McBiJob),
              "hb_tu_job": C.sizeof(hb.TuJob), "hb_tu_result": C.sizeof(hb.TuResult), "hb_tq_params": C.sizeof(hb.TqParams), "hb_prepass_cfg": C.sizeof(hb.PrepassCfg),
              "hb_me_result_c": hb.lib.ME_COMPACT_DT.itemsize, "hb_tu_result_c": hb.lib.TU_COMPACT_DT.itemsize, "hb_sao_stats": hb.lib.SAO_DT.itemsize,
              "hb_sao_param": hb.lib.SAO_PARAM_DT.itemsize, "hb_sao_candidate": hb.lib.SAO_CAND_DT.itemsize, "hb_unit_info": hb.lib.UNIT_INFO_DT.itemsize,
              "hb_deblock_params": 16, "hb_low_level_funcs": C.sizeof(hb.LowLevelFuncs)}
    for n in names:
        assert int(got[n]) == mirror[n], (n, got[n], mirror[n])


def test_committed_ncu_profile_belongs_to_the_committed_kernels():
    """bench.py quotes executed instructions and DRAM traffic from profiles/inst_r02.json only when that profile was taken on exactly the
    kernel sources in the tree (a sha over csrc/*.cu, *.cuh): a kernel edit without a fresh launch list must not go unnoticed"""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("hb_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    with open(os.path.join(root, "profiles", "inst_r02.json")) as f:
        prof = json.load(f)
    assert prof["kernel_source_sha"] == bench.kernel_source_sha(), "profiles/inst_r02.json is stale: re-run tools/gpu_round.sh and commit its launch list"
