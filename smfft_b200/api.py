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
            "smfft_exec_c2c_stream": [P, P, I, LL, I, I, P],
            "smfft_exec_r2c_c2r_stream": [P, P, I, LL, I, P],
            "smfft_pipeline_release": [],
            "smfft_exec_repeated": [P, P, I, LL, I, I, I, I],
            "smfft_select_report": [ctypes.c_char_p, I],
            "smfft_last_error_code": [],
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
        L.smfft_twiddle_table.argtypes = []
        L.smfft_twiddle_table.restype = ctypes.c_void_p
        _LIB = L
    return _LIB


def _check(rc: int) -> None:
    if rc != 0:
        raise SmfftError(lib().smfft_last_error().decode() or f"libsmfft returned {rc}")


def _ptr(t, need_bytes: int = 0, what: str = "buffer", device: bool = True) -> int:
    """device (or host) address of a torch tensor / numpy array / raw int.  Tensors and arrays are checked before the
    address crosses the C ABI (which sees only a pointer): contiguous, on the right side of the bus, at least
    `need_bytes` long -- a wrong nFFTs would otherwise be a silent out-of-bounds access on the device."""
    if isinstance(t, int):
        return t  # raw address: the caller vouches for it (C-ABI semantics)
    if hasattr(t, "data_ptr"):
        if not t.is_contiguous():
            raise SmfftError(f"{what}: tensor must be contiguous")
        if device and not t.is_cuda:
            raise SmfftError(f"{what}: expected a CUDA tensor")
        if not device and t.is_cuda:
            raise SmfftError(f"{what}: expected a host tensor")
        if str(t.dtype) not in ("torch.float32", "torch.complex64"):
            raise SmfftError(f"{what}: dtype {t.dtype} is not float32 / complex64")
        if t.numel() * t.element_size() < need_bytes:
            raise SmfftError(f"{what}: {t.numel() * t.element_size()} bytes, the call needs {need_bytes}")
        return t.data_ptr()
    if not t.flags["C_CONTIGUOUS"]:
        raise SmfftError(f"{what}: array must be C-contiguous")
    if t.nbytes < need_bytes:
        raise SmfftError(f"{what}: {t.nbytes} bytes, the call needs {need_bytes}")
    return t.ctypes.data


def _current_stream() -> int:
    """torch's current stream, passed PER CALL (nothing library-global is rebound)"""
    import torch

    return torch.cuda.current_stream().cuda_stream


class _on_current_stream:
    """the timed launchers use the calling thread's smfft_set_stream stream: set it for the call, restore the default after"""

    def __enter__(self):
        lib().smfft_set_stream(ctypes.c_void_p(_current_stream()))

    def __exit__(self, *exc):
        lib().smfft_set_stream(None)
        return False


def _c2c_bytes(FFT_size: int, nFFTs: int, reps: int = 1) -> int:
    return max(0, nFFTs // reps) * FFT_size * 8


def _real_bytes(FFT_size: int, nFFTs: int, reps: int = 1) -> int:
    return max(0, nFFTs // reps) * FFT_size * 4


def FFT_init() -> None:
    _check(lib().smfft_init())


def set_option(key: str, value: int) -> None:
    _check(lib().smfft_set_option(key.encode(), int(value)))


def get_option(key: str) -> int:
    return lib().smfft_get_option(key.encode())


def twiddle_table() -> int:
    """device address of the W_16384 table (for smfft::BlockFFT<..., TW_LUT> in user kernels)"""
    p = lib().smfft_twiddle_table()
    if not p:
        raise SmfftError(lib().smfft_last_error().decode())
    return int(p)


def select_report() -> str:
    """the first-use selection's decisions on the current device (smfft_select_report)"""
    buf = ctypes.create_string_buffer(1 << 16)
    lib().smfft_select_report(buf, len(buf))
    return buf.value.decode()


def launch_count() -> int:
    return int(lib().smfft_launch_count())


def _timed(fn, *args) -> float:
    ms = ctypes.c_double(0.0)
    with _on_current_stream():
        _check(fn(*args, ctypes.byref(ms)))
    return ms.value


def FFT_external_benchmark(d_input, d_output, FFT_size: int, nFFTs: int, inverse: bool, reorder: bool) -> float:
    """One timed launch of the Cooley-Tukey C2C transform; returns milliseconds."""
    nb = _c2c_bytes(FFT_size, nFFTs)
    return _timed(lib().smfft_external_benchmark, _ptr(d_input, nb, "d_input"), _ptr(d_output, nb, "d_output"), FFT_size, nFFTs,
                  int(inverse), int(reorder))


def FFT_multiple_benchmark(d_input, d_output, FFT_size: int, nFFTs: int, inverse: bool, reorder: bool) -> float:
    nb = _c2c_bytes(FFT_size, nFFTs, 100)
    return _timed(lib().smfft_multiple_benchmark, _ptr(d_input, nb, "d_input"), _ptr(d_output, nb, "d_output"), FFT_size, nFFTs,
                  int(inverse), int(reorder))


def Stockham_external_benchmark(d_input, d_output, FFT_size: int, nFFTs: int, inverse: bool = True) -> float:
    nb = _c2c_bytes(FFT_size, nFFTs)
    return _timed(lib().smfft_stockham_external_benchmark, _ptr(d_input, nb, "d_input"), _ptr(d_output, nb, "d_output"), FFT_size,
                  nFFTs, int(inverse))


def Stockham_multiple_benchmark(d_input, d_output, FFT_size: int, nFFTs: int, inverse: bool = True) -> float:
    nb = _c2c_bytes(FFT_size, nFFTs, 100)
    return _timed(lib().smfft_stockham_multiple_benchmark, _ptr(d_input, nb, "d_input"), _ptr(d_output, nb, "d_output"), FFT_size,
                  nFFTs, int(inverse))


def R2C_C2R_external_benchmark(d_input, d_output, FFT_size: int, nFFTs: int, inverse: int) -> float:
    nb = _real_bytes(FFT_size, nFFTs)
    return _timed(lib().smfft_r2c_c2r_external_benchmark, _ptr(d_input, nb, "d_input"), _ptr(d_output, nb, "d_output"), FFT_size,
                  nFFTs, int(inverse))


def R2C_multiple_benchmark(d_input, d_output, FFT_size: int, nFFTs: int) -> float:
    nb = _real_bytes(FFT_size, nFFTs, 100)
    return _timed(lib().smfft_r2c_multiple_benchmark, _ptr(d_input, nb, "d_input"), _ptr(d_output, nb, "d_output"), FFT_size, nFFTs)


def exec_c2c(d_input, d_output, FFT_size: int, nFFTs: int, inverse: bool, reorder: bool) -> None:
    """Untimed launch on torch's current stream."""
    nb = _c2c_bytes(FFT_size, nFFTs)
    _check(lib().smfft_exec_c2c_stream(_ptr(d_input, nb, "d_input"), _ptr(d_output, nb, "d_output"), FFT_size, nFFTs, int(inverse),
                                       int(reorder), ctypes.c_void_p(_current_stream())))


def exec_r2c_c2r(d_input, d_output, FFT_size: int, nFFTs: int, inverse: int) -> None:
    nb = _real_bytes(FFT_size, nFFTs)
    _check(lib().smfft_exec_r2c_c2r_stream(_ptr(d_input, nb, "d_input"), _ptr(d_output, nb, "d_output"), FFT_size, nFFTs,
                                           int(inverse), ctypes.c_void_p(_current_stream())))


def exec_repeated(d_input, d_output, FFT_size: int, nFFTs: int, inverse: bool, reorder: bool, mode: int = 0, reps: int = 3) -> None:
    """The FFT_multiple kernels with `reps` in-place repetitions over all nFFTs transforms (value-checkable at reps = 3)."""
    nb = _c2c_bytes(FFT_size, nFFTs) if mode == 0 else _real_bytes(FFT_size, nFFTs)
    with _on_current_stream():
        _check(lib().smfft_exec_repeated(_ptr(d_input, nb, "d_input"), _ptr(d_output, nb, "d_output"), FFT_size, nFFTs, int(inverse),
                                         int(reorder), mode, reps))


def c2c_host(h_input, h_output, FFT_size: int, nFFTs: int, inverse: bool, reorder: bool, nRuns: int = 1):
    """GPU_smFFT_4elements (CT/FFT-GPU-32bit.cu:827-908): host buffers in, host buffers out.
    Returns (single_ex_time_ms, multi_ex_time_ms)."""
    s, m = ctypes.c_double(0.0), ctypes.c_double(0.0)
    nb = _c2c_bytes(FFT_size, nFFTs)
    with _on_current_stream():
        _check(lib().smfft_c2c_host(_ptr(h_input, nb, "h_input", False), _ptr(h_output, nb, "h_output", False), FFT_size, nFFTs,
                                    int(inverse), int(reorder), nRuns, ctypes.byref(s), ctypes.byref(m)))
    return s.value, m.value


def pipeline_host(h_input, h_output, FFT_size: int, nFFTs: int, inverse: bool = False, reorder: bool = True,
                  mode: int = 0, chunk_ffts: int = 0) -> float:
    """Chunked H2D -> FFT -> D2H pipeline on (pinned) host buffers; returns milliseconds."""
    ms = ctypes.c_double(0.0)
    nb = _c2c_bytes(FFT_size, nFFTs) if mode == 0 else _real_bytes(FFT_size, nFFTs)
    _check(lib().smfft_pipeline_host(_ptr(h_input, nb, "h_input", False), _ptr(h_output, nb, "h_output", False), FFT_size, nFFTs,
                                     int(inverse), int(reorder), mode, chunk_ffts, ctypes.byref(ms)))
    return ms.value


def pipeline_release() -> None:
    """free the current device's pipeline buffers / streams (smfft_pipeline_release)"""
    _check(lib().smfft_pipeline_release())
