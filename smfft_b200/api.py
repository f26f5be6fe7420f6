"""ctypes binding of include/smfft.h + thin launchers that keep the reference's names.

Reference interface mirrored (paths relative to the reference root):
  FFT_init                                   SMFFT_CooleyTukey_C2C/FFT-GPU-32bit.cu:576-581
  FFT_external_benchmark(d_input, d_output, FFT_size, nFFTs, inverse, reorder, &time)   ...:583-664
  FFT_multiple_benchmark(...)                                                           ...:666-752
  Stockham / R2C-C2R launchers               SMFFT_Stockham_C2C/...:306-384, SMFFT_Stockham_R2C_C2R/...:396-467
Times are milliseconds and are RETURNED (the C ABI accumulates into *ms like the reference).
"""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SmfftError(RuntimeError):
    pass


def lib_path() -> str:
    # SMFFT_LIB selects an experiment build of the same CUDA library (smfft_b200/build.py, SMFFT_VARIANT)
    return os.environ.get("SMFFT_LIB") or os.path.join(_PKG, "lib", "libsmfft.so")


def lib() -> ctypes.CDLL:
    """Load libsmfft.so (built in-tree by smfft_b200.build).  Fails loudly when it is missing."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise SmfftError(f"{path} not found: run `python -m smfft_b200.build` (no CPU fallback exists)")
        L = ctypes.CDLL(path)
        P, I, LL, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.POINTER(ctypes.c_double)
        sig = {
            "smfft_init": [],
            "smfft_external_benchmark": [P, P, I, LL, I, I, D],
            "smfft_multiple_benchmark": [P, P, I, LL, I, I, D],
            "smfft_exec_c2c": [P, P, I, LL, I, I],
            "smfft_stockham_external_benchmark": [P, P, I, LL, I, D],
            "smfft_stockham_multiple_benchmark": [P, P, I, LL, I, D],
            "smfft_r2c_c2r_external_benchmark": [P, P, I, LL, I, D],
            "smfft_r2c_multiple_benchmark": [P, P, I, LL, D],
            "smfft_exec_r2c_c2r": [P, P, I, LL, I],
            "smfft_c2c_host": [P, P, I, LL, I, I, I, D, D],
            "smfft_r2c_c2r_host": [P, P, I, LL, I, I, D, D],
            "smfft_pipeline_host": [P, P, I, LL, I, I, I, LL, D],
            "smfft_set_option": [ctypes.c_char_p, I],
            "smfft_get_option": [ctypes.c_char_p],
            "smfft_set_stream": [P],
            "smfft_version": [],
        }
        for name, args in sig.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = I
        L.smfft_launch_count.argtypes = []
        L.smfft_launch_count.restype = LL
        L.smfft_last_error.argtypes = []
        L.smfft_last_error.restype = ctypes.c_char_p
        _LIB = L
    return _LIB


def _check(rc: int) -> None:
    if rc != 0:
        raise SmfftError(lib().smfft_last_error().decode() or f"libsmfft returned {rc}")


def _ptr(t) -> int:
    """device (or host) address of a torch tensor / numpy array / raw int"""
    if isinstance(t, int):
        return t
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def _use_current_stream() -> None:
    import torch

    lib().smfft_set_stream(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))


def FFT_init() -> None:
    _check(lib().smfft_init())


def set_option(key: str, value: int) -> None:
    _check(lib().smfft_set_option(key.encode(), int(value)))


def get_option(key: str) -> int:
    return lib().smfft_get_option(key.encode())


def launch_count() -> int:
    return int(lib().smfft_launch_count())


def _timed(fn, *args) -> float:
    ms = ctypes.c_double(0.0)
    _use_current_stream()
    _check(fn(*args, ctypes.byref(ms)))
    return ms.value


def FFT_external_benchmark(d_input, d_output, FFT_size: int, nFFTs: int, inverse: bool, reorder: bool) -> float:
    """One timed launch of the Cooley-Tukey C2C transform; returns milliseconds."""
    return _timed(lib().smfft_external_benchmark, _ptr(d_input), _ptr(d_output), FFT_size, nFFTs, int(inverse), int(reorder))


def FFT_multiple_benchmark(d_input, d_output, FFT_size: int, nFFTs: int, inverse: bool, reorder: bool) -> float:
    return _timed(lib().smfft_multiple_benchmark, _ptr(d_input), _ptr(d_output), FFT_size, nFFTs, int(inverse), int(reorder))


def Stockham_external_benchmark(d_input, d_output, FFT_size: int, nFFTs: int, inverse: bool = True) -> float:
    return _timed(lib().smfft_stockham_external_benchmark, _ptr(d_input), _ptr(d_output), FFT_size, nFFTs, int(inverse))


def Stockham_multiple_benchmark(d_input, d_output, FFT_size: int, nFFTs: int, inverse: bool = True) -> float:
    return _timed(lib().smfft_stockham_multiple_benchmark, _ptr(d_input), _ptr(d_output), FFT_size, nFFTs, int(inverse))


def R2C_C2R_external_benchmark(d_input, d_output, FFT_size: int, nFFTs: int, inverse: int) -> float:
    return _timed(lib().smfft_r2c_c2r_external_benchmark, _ptr(d_input), _ptr(d_output), FFT_size, nFFTs, int(inverse))


def R2C_multiple_benchmark(d_input, d_output, FFT_size: int, nFFTs: int) -> float:
    return _timed(lib().smfft_r2c_multiple_benchmark, _ptr(d_input), _ptr(d_output), FFT_size, nFFTs)


def exec_c2c(d_input, d_output, FFT_size: int, nFFTs: int, inverse: bool, reorder: bool) -> None:
    """Untimed launch on torch's current stream."""
    _use_current_stream()
    _check(lib().smfft_exec_c2c(_ptr(d_input), _ptr(d_output), FFT_size, nFFTs, int(inverse), int(reorder)))


def exec_r2c_c2r(d_input, d_output, FFT_size: int, nFFTs: int, inverse: int) -> None:
    _use_current_stream()
    _check(lib().smfft_exec_r2c_c2r(_ptr(d_input), _ptr(d_output), FFT_size, nFFTs, int(inverse)))


def c2c_host(h_input, h_output, FFT_size: int, nFFTs: int, inverse: bool, reorder: bool, nRuns: int = 1):
    """GPU_smFFT_4elements (CT/FFT-GPU-32bit.cu:827-908): host buffers in, host buffers out.
    Returns (single_ex_time_ms, multi_ex_time_ms)."""
    s, m = ctypes.c_double(0.0), ctypes.c_double(0.0)
    _use_current_stream()
    _check(lib().smfft_c2c_host(_ptr(h_input), _ptr(h_output), FFT_size, nFFTs, int(inverse), int(reorder), nRuns,
                                ctypes.byref(s), ctypes.byref(m)))
    return s.value, m.value


def pipeline_host(h_input, h_output, FFT_size: int, nFFTs: int, inverse: bool = False, reorder: bool = True,
                  mode: int = 0, chunk_ffts: int = 0) -> float:
    """Chunked H2D -> FFT -> D2H pipeline on (pinned) host buffers; returns milliseconds."""
    ms = ctypes.c_double(0.0)
    _check(lib().smfft_pipeline_host(_ptr(h_input), _ptr(h_output), FFT_size, nFFTs, int(inverse), int(reorder), mode,
                                     chunk_ffts, ctypes.byref(ms)))
    return ms.value
