"""The compiled host above the C ABI (musubi_b200/csrc/host/mus_b200_host.cpp): builds against
include/musb200.h and the exported symbols only, and fails loudly where the library does.  Its
run on a GPU (state file in, 40 steps, dump bit-compared with the oracle: tests/host_driver_parity.py,
periodic BGK / MRT, the lid cavity and the channel with a pressure outlet) is the gpu test below."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "musubi_b200", "mus_b200_host")


def _run(*args):
    return subprocess.run([EXE] + list(args), capture_output=True, text=True, timeout=120)


def test_headers_are_valid_c99_and_cxx():
    for hdr in (os.path.join(ROOT, "include", "musb200.h"),
                os.path.join(ROOT, "musubi_b200", "csrc", "host", "treelm_box.h")):
        for cmd in (["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                    ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr]):
            r = subprocess.run(cmd, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr


def test_host_driver_builds_and_explains_itself():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "musubi_b200", "csrc"), "../mus_b200_host"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = _run("--help")
    assert r.returncode == 0 and "usage: mus_b200_host" in r.stdout
    assert _run("--no-such-option").returncode == 2


def test_host_driver_error_behaviour_mirrors_the_library():
    """an identify outside the hot path ends like tem_abort does, before any device is touched
    (MUSB200_ERR_UNSUPPORTED = 4); without a CUDA device musb200_init refuses (no CPU fallback)"""
    r = _run("--relaxation", "cumulant")
    assert r.returncode == 4 and "musb200_scheme_select" in r.stderr
    r = _run("--layout", "d2q9")
    assert r.returncode == 4
    import ctypes
    import musubi_b200._lib as L
    n = ctypes.c_int(0)
    has_gpu = L.lib.musb200_device_count(ctypes.byref(n)) == 0 and n.value > 0
    r = _run("--level", "3", "--steps", "2")
    if has_gpu:
        assert r.returncode == 0 and "MLUPS" in r.stdout, r.stderr
    else:
        assert r.returncode == 2 and "musb200_init" in r.stderr and "failed (code 2)" in r.stderr


@pytest.mark.gpu
def test_host_driver_runs_match_the_oracle_bit_for_bit():
    """the C++ host program drives libmusb200.so through the header alone: initial state from a
    file, 40 steps, restart dump -- four cases, every PDF equal to the oracle's"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_driver_parity.py")], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ndiff=0") == 4 and "host driver parity: OK" in r.stdout
