"""smfft_b200 -- B200-native (sm_100a) shared-memory batched FFT, drop-in for the SMFFT hot path.

Host-side mirror of the reference's launcher interface (same names, argument meaning and error
behaviour as KAdamek/SMFFT's FFT_external_benchmark / FFT_multiple_benchmark, see include/smfft.h)
over the C ABI of smfft_b200/lib/libsmfft.so.  PyTorch is used only for device memory and streams.
There is no CPU fallback: without the CUDA library every call raises.
"""
from .api import (  # noqa: F401
    SmfftError,
    FFT_external_benchmark,
    FFT_multiple_benchmark,
    FFT_init,
    Stockham_external_benchmark,
    Stockham_multiple_benchmark,
    R2C_C2R_external_benchmark,
    R2C_multiple_benchmark,
    exec_c2c,
    exec_r2c_c2r,
    exec_repeated,
    pipeline_release,
    c2c_host,
    pipeline_host,
    set_option,
    get_option,
    launch_count,
    select_report,
    twiddle_table,
    lib,
    lib_path,
)
from .shard import shard_ffts  # noqa: F401
