"""The C ABI used from plain C99 (examples/prepass_demo.c): compiles against include/homer_b200.h, links the shared library, runs."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_prepass_demo_in_c(ctx):
    exe = os.path.join(ROOT, "build", "prepass_demo")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "prepass_demo.c"), "-o", exe, "-L" + os.path.join(ROOT, "homerhevc_b200"),
                           "-lhomer_b200", "-Wl,-rpath," + os.path.join(ROOT, "homerhevc_b200"), "-lm"])
    out = subprocess.run([exe, "320", "192", "3"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "prepass_demo ok" in out.stdout and out.stdout.count("frame ") == 6 and "resident frame 3:" in out.stdout, out.stdout


def test_e2e_flow_driven_by_c_threads(ctx):
    """examples/e2e_threads.c: bench.py's device-resident e2e flow with pthreads as the host side (two streams in flight per thread)"""
    exe = os.path.join(ROOT, "build", "e2e_threads")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-pthread", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "e2e_threads.c"), "-o", exe, "-L" + os.path.join(ROOT, "homerhevc_b200"),
                           "-lhomer_b200", "-Wl,-rpath," + os.path.join(ROOT, "homerhevc_b200"), "-lm"])
    out = subprocess.run([exe, "416", "240", "3", "8"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "e2e_threads ok" in out.stdout and "48 frames" in out.stdout, out.stdout
