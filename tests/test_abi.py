"""The C-ABI library loads and exports every symbol include/smfft.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

import smfft_b200
from smfft_b200 import build as smbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    smbuild.build()
    return smfft_b200.lib()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "smfft.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(smfft_[a-z0-9_]+)\s*\(", hdr)))


def test_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/smfft.h but not exported"
    assert lib.smfft_version() == 200


def test_exports_nothing_from_the_oracle(lib):
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", smfft_b200.lib_path()], capture_output=True, text=True).stdout
    assert "oracle_" not in out and "emu_" not in out
    # the product library contains TMA tensor copies (UTMALDG/UTMASTG are emitted from these PTX ops)
    assert os.path.getsize(smfft_b200.lib_path()) > 1 << 20


def test_argument_errors_do_not_need_a_device(lib):
    ms = ctypes.c_double(0)
    assert lib.smfft_external_benchmark(None, None, 48, 10, 0, 1, ctypes.byref(ms)) != 0
    assert b"wrong FFT length" in lib.smfft_last_error()          # CT:656-658 prints the same words
    assert lib.smfft_external_benchmark(None, None, 1 << 25, 10, 0, 1, ctypes.byref(ms)) != 0
    assert b"wrong FFT length" in lib.smfft_last_error()          # multi-pass transforms stop at 2^24 points
    assert lib.smfft_external_benchmark(None, None, 1 << 15, 10, 0, 0, ctypes.byref(ms)) != 0
    assert b"natural order" in lib.smfft_last_error()             # ... and take natural-order input only
    assert lib.smfft_multiple_benchmark(None, None, 1 << 15, 200, 0, 1, ctypes.byref(ms)) != 0
    assert lib.smfft_multiple_benchmark(None, None, 1024, 99, 0, 1, ctypes.byref(ms)) != 0
    assert ms.value == -1                                          # CT:670-673
    assert lib.smfft_set_option(b"no_such_option", 1) != 0
    assert lib.smfft_set_option(b"twiddle", 1) == 0 and lib.smfft_get_option(b"twiddle") == 1
    assert lib.smfft_set_option(b"twiddle", 0) == 0
    assert lib.smfft_set_option(b"multi_pass_chunk_mib", 64) == 0 and lib.smfft_get_option(b"two_pass_chunk_mib") == 64   # one option, two names
    assert lib.smfft_set_option(b"two_pass_chunk_mib", 1024) == 0 and lib.smfft_set_option(b"multi_pass_chunk_mib", 0) != 0


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("device present")
    with pytest.raises(smfft_b200.SmfftError):
        smfft_b200.FFT_init()
    src = open(os.path.join(ROOT, "smfft_b200", "api.py")).read() + open(os.path.join(ROOT, "smfft_b200", "csrc", "launch.cu")).read()
    assert "oracle" not in src.replace("no CPU fallback", "")
