"""The boundary is usable from plain C: include/orb_b200.h is valid C99 (-pedantic), a C program links
against liborbb200.so alone, and without a CUDA device every compute entry point refuses loudly."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    from orb_slam3_b200 import build
    build.build()
    exe = str(tmp_path_factory.mktemp("abi") / "abi_demo")
    libdir = os.path.join(ROOT, "orb_slam3_b200")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "abi_demo.c"), "-L", libdir, "-lorbb200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    return dict(l.split(" ", 1) for l in out.stdout.strip().splitlines())


def test_c99_program_links_and_runs(demo):
    assert demo["version"].startswith("orb_slam3_b200")
    assert demo["orb_create"] == "0" and demo["ham_distance"] == "256"


def test_compute_entry_points_refuse_without_a_device(demo):
    if int(demo["devices"]) > 0:
        pytest.skip("a CUDA device is present")
    assert demo["orb_extract"].startswith("-5 ") and "no CPU path" in demo["orb_extract"]   # ORB_E_NODEVICE
    for k in ("stereo_create", "poseopt_create", "frustum_create"):
        assert demo[k] == "-5", k
