"""The C-ABI shared library loads on a CPU-only box and exports exactly the symbols that
include/ipp_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from tests.conftest import ROOT

from mripp_b200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "ipp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(ipp_\w+)\s*\(", text)))


def test_header_library_and_binding_agree():
    names = _declared()
    assert len(names) >= 16
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ipp_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.ipp_abi_version() == 12


def test_layout_is_host_only_and_consistent():
    lib = _lib.load()
    out = (ctypes.c_longlong * 10)()
    assert lib.ipp_gp_layout(131, out) == 0
    NP, rows, ld = out[0], out[1], out[2]
    assert NP == 192 and rows == NP + 64 and ld == NP
    assert out[3] == rows * ld and out[4] == NP * ld and out[5] == NP * ld
    assert out[7] == 4 * out[9] and out[6] >= out[8] + 8
    assert lib.ipp_gp_layout(0, out) == -1


def test_bad_arguments_are_rejected_without_a_device():
    lib = _lib.load()
    assert lib.ipp_gram_rbf(None, 0, None, 0, 1.0, 1.0, 0.0, None, 0, None) == -1
    assert lib.ipp_gp_posterior_points(None, None, 0, 1.0, 1.0, 0.0, 1.0, None, 0, None, None, None, None, 0, None) == -1
    assert lib.ipp_gp_posterior_points_tabled(None, None, 0, 1.0, 1.0, 0.0, 1.0, None, 0, None, None, None, None, None, 128,
                                              None, 0.0, 0, None) == -1
    assert lib.ipp_argmax_first(None, 0, None, None, None) == -1


def test_plain_c_consumer_builds_and_runs(tmp_path):
    """include/ipp_b200.h compiles as C, libipp_b200.so links from gcc, and the entry points that need no device
    (version, layout, host-side tree growth) behave (tests/c_abi/consumer.c)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    from mripp_b200 import _lib

    _lib.load()
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "consumer")
    src = os.path.join(ROOT, "tests", "c_abi", "consumer.c")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", libdir,
                    "-lipp_b200", "-lm", f"-Wl,-rpath,{libdir}"], check=True)
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert "c_abi ok" in run.stdout
