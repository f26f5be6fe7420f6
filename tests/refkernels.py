"""ctypes access to the reference's own launchers, recompiled UNMODIFIED for sm_100a into oracle/_ref/
(recipe: oracle/Makefile).  Test infrastructure: used by the GPU parity tests, the golden-fixture
generator and bench.py's reported baselines -- never by the product path."""
import ctypes
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
_cache = {}


def available() -> bool:
    return all(os.path.exists(os.path.join(REF, f"libsmfft_ref_{k}.so")) for k in ("ct", "st", "rc"))


def _lib(kind):
    if kind not in _cache:
        _cache[kind] = ctypes.CDLL(os.path.join(REF, f"libsmfft_ref_{kind}.so"), mode=os.RTLD_LOCAL)
    return _cache[kind]


P, I, B, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_bool, ctypes.POINTER(ctypes.c_double)


def ct_external(d_in, d_out, n, nffts, inverse, reorder) -> float:
    """int FFT_external_benchmark(float2*, float2*, int, int, bool, bool, double*)  CT/FFT-GPU-32bit.cu:583"""
    fn = getattr(_lib("ct"), "_Z22FFT_external_benchmarkP6float2S0_iibbPd")
    fn.argtypes = [P, P, I, I, B, B, D]
    ms = ctypes.c_double(0)
    fn(d_in.data_ptr(), d_out.data_ptr(), n, nffts, bool(inverse), bool(reorder), ctypes.byref(ms))
    return ms.value


def ct_multiple(d_in, d_out, n, nffts, inverse, reorder) -> float:
    fn = getattr(_lib("ct"), "_Z22FFT_multiple_benchmarkP6float2S0_iibbPd")
    fn.argtypes = [P, P, I, I, B, B, D]
    ms = ctypes.c_double(0)
    fn(d_in.data_ptr(), d_out.data_ptr(), n, nffts, bool(inverse), bool(reorder), ctypes.byref(ms))
    return ms.value


def st_external(d_in, d_out, n, nffts) -> float:
    """void FFT_external_benchmark(float2*, float2*, int, int, double*)  ST/...:306 (inverse only)"""
    fn = getattr(_lib("st"), "_Z22FFT_external_benchmarkP6float2S0_iiPd")
    fn.argtypes = [P, P, I, I, D]
    ms = ctypes.c_double(0)
    fn(d_in.data_ptr(), d_out.data_ptr(), n, nffts, ctypes.byref(ms))
    return ms.value


def st_multiple(d_in, d_out, n, nffts) -> float:
    fn = getattr(_lib("st"), "_Z22FFT_multiple_benchmarkP6float2S0_iiPd")
    fn.argtypes = [P, P, I, I, D]
    ms = ctypes.c_double(0)
    fn(d_in.data_ptr(), d_out.data_ptr(), n, nffts, ctypes.byref(ms))
    return ms.value


def rc_external(d_in, d_out, n, nffts, inverse) -> float:
    """void FFT_external_benchmark(float*, float*, int, int, int inverse, double*)  RC/...:396"""
    fn = getattr(_lib("rc"), "_Z22FFT_external_benchmarkPfS_iiiPd")
    fn.argtypes = [P, P, I, I, I, D]
    ms = ctypes.c_double(0)
    fn(d_in.data_ptr(), d_out.data_ptr(), n, nffts, int(inverse), ctypes.byref(ms))
    return ms.value


def rc_multiple(d_in, d_out, n, nffts) -> float:
    """void FFT_multiple_benchmark(float*, float*, int, int, double*)  RC/...:435 (forward only)"""
    fn = getattr(_lib("rc"), "_Z22FFT_multiple_benchmarkPfS_iiPd")
    fn.argtypes = [P, P, I, I, D]
    ms = ctypes.c_double(0)
    fn(d_in.data_ptr(), d_out.data_ptr(), n, nffts, ctypes.byref(ms))
    return ms.value
