"""Batch sharding of independent FFTs across the GPUs of one box.

Each FFT is independent and contiguous in memory (row b at offset b*N,
SMFFT_CooleyTukey_C2C/FFT-GPU-32bit.cu:538), so rank g of G owns rows [lo, hi) and no data crosses
GPUs; the only collective in a multi-GPU run is the timing barrier.
"""
from __future__ import annotations


def shard_ffts(n_ffts: int, world_size: int, rank: int, granularity: int = 1):
    """Rows [lo, hi) of the batch owned by `rank`.  Shard boundaries are multiples of `granularity`
    (the reference packs 4 FFTs of 32 / 2 of 64 per CTA, CT/...:588-595; our tiles hold F FFTs)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    if granularity < 1:
        raise ValueError("granularity must be >= 1")
    units = (n_ffts + granularity - 1) // granularity
    base, rem = divmod(units, world_size)
    lo_u = rank * base + min(rank, rem)
    hi_u = lo_u + base + (1 if rank < rem else 0)
    return min(lo_u * granularity, n_ffts), min(hi_u * granularity, n_ffts)
