"""CPU checks of the *kernel sources themselves*: include/smfft/detail/*.cuh and
smfft_b200/csrc/kernels.cuh are compiled with g++ -DSMFFT_EMU and executed by the SIMT emulator in
tests/emu (one fiber per CUDA thread; TMA / mbarrier emulated as an asynchronous queue), then compared
with the FP64 oracle.  This is test infrastructure: the product has no CPU path.  GPU parity proper
is tests/test_gpu_parity.py (-m gpu).  Tolerance: relative L2 <= 1e-5 (north_star)."""
import ctypes

import numpy as np
import pytest

from oracle import oracle_np as O
from tests.emu.build_emu import build

TOL = 1e-5
C2C, R2C, C2R = 0, 1, 2
IO_TMA, IO_LDG = 0, 1


IO_TMA_STG = 2


@pytest.fixture(scope="module")
def emu():
    lib = ctypes.CDLL(build())
    P, I, LL, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.POINTER(ctypes.c_double)
    lib.emu_run.argtypes = [P, P, I, LL, I, I, I, I, I, I, I, D]
    lib.emu_run_alt.argtypes = [P, P, I, LL, I, I, I, I, I, D]
    lib.emu_run_compat.argtypes = [P, P, I, LL, I, I, I, D]
    lib.emu_run_alternate.argtypes = [P, P, I, I, LL, I, I]
    lib.emu_run_compat_ct.argtypes = [P, P, I, LL, I, I, I, D, ctypes.POINTER(ctypes.c_longlong)]
    lib.emu_run_late.argtypes = [P, P, I, LL, I]
    lib.emu_run_dual.argtypes = [P, P, I, I, LL, I, I, I, I, I, D]
    return lib


def run(lib, x, e, mode, direction, reorder, io, tw, reps=1, grid=2, allow_missing=False):
    x = np.ascontiguousarray(x)
    out = np.zeros_like(x)
    bank = ctypes.c_double(0)
    n_ffts = x.size // (1 << e) if x.dtype == np.complex64 else x.size // (2 << e)
    rc = lib.emu_run(x.ctypes.data, out.ctypes.data, e, n_ffts, mode, direction, reorder, io, tw, reps, grid, ctypes.byref(bank))
    if rc != 0 and allow_missing:
        return None, None
    assert rc == 0, "configuration not instantiated in the emulator"
    return out, bank.value


def batch(lib, e, extra=1):
    # a whole number of tiles plus a ragged tail, so the partial-tile path (TMA OOB fill / LDG guards) runs
    return max(3 * lib.emu_tile_points(e) // (1 << e) // 2, 1) + extra


@pytest.mark.parametrize("e", range(5, 13))
@pytest.mark.parametrize("io", [IO_TMA, IO_LDG])
def test_c2c_all_modes(emu, e, io):
    n = 1 << e
    x = O.uniform_c64(batch(emu, e), n)
    for direction in (0, 1):
        for reorder in (1, 0):
            y, bank = run(emu, x, e, C2C, direction, reorder, io, 0)
            assert O.rel_l2(y, O.ct_c2c_fp64(x, bool(direction), bool(reorder))) < TOL
            assert bank == pytest.approx(1.0), "every shared-memory access of the C2C path must be bank-conflict free"


@pytest.mark.parametrize("e", range(5, 13))
def test_c2c_mufu_twiddles(emu, e):
    n = 1 << e
    x = O.uniform_c64(batch(emu, e), n)
    y, _ = run(emu, x, e, C2C, 0, 1, IO_TMA, 1)
    assert O.rel_l2(y, O.ct_c2c_fp64(x, False, True)) < TOL
    y, _ = run(emu, x, e, C2C, 1, 0, IO_LDG, 1)
    assert O.rel_l2(y, O.ct_c2c_fp64(x, True, False)) < TOL


@pytest.mark.parametrize("e", range(5, 13))
def test_permutation_exact_by_one_hot(emu, e):
    """reorder=0 must equal DFT(x o brev_e) for every input slot: delta_p -> W^{brev(p) k}."""
    n = 1 << e
    perm = O.c_reorder_index(n)
    ps = np.unique(np.concatenate([np.arange(min(n, 48)), np.random.default_rng(e).integers(0, n, 16), [n - 1]]))
    x = np.zeros((len(ps), n), np.complex64)
    x[np.arange(len(ps)), ps] = 1
    y, _ = run(emu, x, e, C2C, 0, 0, IO_TMA, 0)
    k = np.arange(n)
    for row, p in enumerate(ps):
        # the source slot is recovered as an exact integer from the phase ramp
        slot = int(np.rint((-np.angle(y[row, 1]) * n / (2 * np.pi)))) % n
        assert slot == perm[p]
        assert np.allclose(y[row], np.exp(-2j * np.pi * perm[p] * k / n), atol=5e-6)
    y1, _ = run(emu, x, e, C2C, 0, 1, IO_TMA, 0)
    for row, p in enumerate(ps):
        assert int(np.rint((-np.angle(y1[row, 1]) * n / (2 * np.pi)))) % n == p


@pytest.mark.parametrize("e", range(5, 13))
@pytest.mark.parametrize("io", [IO_TMA, IO_LDG])
def test_r2c_c2r(emu, e, io):
    n = 2 << e  # real length
    x = O.uniform_f32(batch(emu, e), n)
    y, _ = run(emu, x, e, R2C, 0, 1, io, 0)
    yc = y.view(np.complex64)
    assert O.rel_l2(yc, O.r2c_packed_fp64(x)) < TOL
    h = O.uniform_c64(batch(emu, e), n // 2, seed=5)
    z, _ = run(emu, h, e, C2R, 1, 1, io, 0)
    assert O.rel_l2(z.view(np.float32), O.c2r_packed_fp64(h)) < TOL
    # round trip = N/2 * identity
    back, _ = run(emu, yc, e, C2R, 1, 1, io, 0)
    assert O.rel_l2(back.view(np.float32) / (n / 2), x) < TOL


@pytest.mark.parametrize("e", [5, 8, 10, 12])
def test_multiple_reps_in_place(emu, e):
    """FFT_multiple re-applies the transform in place (CT:563-565); with 3 reps: F^3 x."""
    n = 1 << e
    x = (O.uniform_c64(batch(emu, e), n) / np.float32(n)).astype(np.complex64)
    y, _ = run(emu, x, e, C2C, 0, 1, IO_LDG, 0, reps=3)
    ref = np.fft.fft(np.fft.fft(np.fft.fft(x.astype(np.complex128))))
    assert O.rel_l2(y, ref) < TOL
    y0, _ = run(emu, x, e, C2C, 0, 0, IO_LDG, 0, reps=3)
    ref0 = x.astype(np.complex128)
    for _ in range(3):
        ref0 = O.ct_c2c_fp64(ref0, False, False)
    assert O.rel_l2(y0, ref0) < TOL


@pytest.mark.parametrize("variant", range(8))
def test_generic_pass_machinery_other_shapes(emu, variant):
    """Other radix plans / tile shapes / stage counts through the same templates (R = 4, 8, 32)."""
    n = emu.emu_alt_length(variant)
    x = O.uniform_c64(9, n)
    for direction, reorder, io in ((0, 1, 0), (1, 0, 0), (0, 0, 1), (1, 1, 1)):
        out = np.zeros_like(x)
        rc = emu.emu_run_alt(x.ctypes.data, out.ctypes.data, variant, 9, direction, reorder, io, 0, 2, None)
        assert rc == 0
        assert O.rel_l2(out, O.ct_c2c_fp64(x, bool(direction), bool(reorder))) < TOL


def test_edge_inputs(emu):
    e, n = 10, 1024
    # all-ones -> N * delta_0 ; ramp ; single FFT (one ragged tile) ; zeros
    ones = np.ones((1, n), np.complex64)
    y, _ = run(emu, ones, e, C2C, 0, 1, IO_TMA, 0)
    assert abs(y[0, 0] - n) < 1e-3 and np.max(np.abs(y[0, 1:])) < 1e-3
    ramp = (np.arange(n, dtype=np.float32) / n).astype(np.complex64)[None, :]
    y, _ = run(emu, ramp, e, C2C, 1, 1, IO_LDG, 0)
    assert O.rel_l2(y, O.ct_c2c_fp64(ramp, True, True)) < TOL
    z, _ = run(emu, np.zeros((2, n), np.complex64), e, C2C, 0, 0, IO_TMA, 0)
    assert not z.any()
    # many tiles on one persistent CTA: every pipeline stage is reused several times
    x = O.uniform_c64(14, n)
    y, _ = run(emu, x, e, C2C, 0, 1, IO_TMA, 0, grid=1)
    assert O.rel_l2(y, O.ct_c2c_fp64(x, False, True)) < TOL


@pytest.mark.parametrize("e", range(5, 13))
def test_tma_in_register_out_staging(emu, e):
    """IO_TMA_STG: TMA loads, results stored from registers, refill issued after the first barrier of
    the next tile.  grid=1 makes one persistent CTA walk many tiles so every buffer is reused."""
    n = 1 << e
    x = O.uniform_c64(batch(emu, e, extra=3), n)
    ran = 0
    for direction, reorder in ((0, 1), (0, 0), (1, 1), (1, 0)):
        y, _ = run(emu, x, e, C2C, direction, reorder, IO_TMA_STG, 0, grid=1, allow_missing=True)
        if y is None:
            continue  # shapes with a single tile buffer have no register-output variant
        ran += 1
        assert O.rel_l2(y, O.ct_c2c_fp64(x, bool(direction), bool(reorder))) < TOL
    h = O.uniform_c64(batch(emu, e), n, seed=5)
    z, _ = run(emu, h, e, C2R, 1, 1, IO_TMA_STG, 0, allow_missing=True)
    if z is not None:
        assert O.rel_l2(z.view(np.float32), O.c2r_packed_fp64(h)) < TOL


@pytest.mark.parametrize("e", range(5, 13))
def test_reference_device_api_configuration(emu, e):
    """The configuration behind include/smfft/compat.cuh (do_SMFFT_CT_DIT & co.): 4 points per thread,
    fft_length/4 threads, linear tile in and out, fft_length/N transforms per tile, MUFU twiddles."""
    n = 1 << e
    x = O.uniform_c64(8, n)
    for direction, reorder in ((0, 1), (0, 0), (1, 1), (1, 0)):
        out = np.zeros_like(x)
        assert emu.emu_run_compat(x.ctypes.data, out.ctypes.data, e, 8, 0, direction, reorder, None) == 0
        assert O.rel_l2(out, O.ct_c2c_fp64(x, bool(direction), bool(reorder))) < TOL
    xr = O.uniform_f32(8, 2 * n)
    out = np.zeros((8, n), np.complex64)
    assert emu.emu_run_compat(xr.ctypes.data, out.ctypes.data, e, 8, 1, 0, 1, None) == 0
    assert O.rel_l2(out, O.r2c_packed_fp64(xr)) < TOL
    h = O.uniform_c64(8, n, seed=3)
    out = np.zeros_like(h)
    assert emu.emu_run_compat(h.ctypes.data, out.ctypes.data, e, 8, 2, 1, 1, None) == 0
    assert O.rel_l2(out.view(np.float32), O.c2r_packed_fp64(h)) < TOL


@pytest.mark.parametrize("e", range(5, 13))
def test_reference_device_api_engines(emu, e):
    """compat::ct_dit -- the dispatch behind do_SMFFT_CT_DIT<P> -- and compat::ct_dit_external -- the body of
    SMFFT_DIT_external<P> -- on the emulator: warp-shuffle engines for one-warp tiles and for fft_reorder = 0, Stockham
    passes for natural order above 128 points.  Values vs FP64, three in-place repetitions (the SMFFT_DIT_multiple loop),
    the bank-conflict factor of the shared-memory accesses and the number of warp shuffles per tile."""
    n = 1 << e
    nf = 8
    x = O.uniform_c64(nf, n, seed=e)
    warps = max(1, n // 128)
    for direction, reorder in ((0, 1), (0, 0), (1, 1), (1, 0)):
        want = O.ct_c2c_fp64(x, bool(direction), bool(reorder))
        for reps in (1, 0):                      # 1: tile in shared memory (device function), 0: tile in global memory (external wrapper)
            out = np.zeros_like(x)
            bank = ctypes.c_double(0)
            sh = ctypes.c_longlong(0)
            assert emu.emu_run_compat_ct(x.ctypes.data, out.ctypes.data, e, nf, direction, reorder, reps, ctypes.byref(bank), ctypes.byref(sh)) == 0
            assert O.rel_l2(out, want) < TOL, (e, direction, reorder, reps)
            if not reorder and reps == 1:
                # fft_reorder = 0: every access conflict-free except the final LINEAR store of 32 (2), 2048 (2) and 4096 points (4 wavefronts)
                assert bank.value <= {5: 1.25, 11: 1.17, 12: 1.51}.get(e, 1.0) + 1e-9, (e, bank.value)
            if reorder and e >= 8:
                assert bank.value <= 1.0 + 1e-9        # natural order above 128 points: Stockham passes, LayoutSW4 exchanges, no conflict anywhere
            if e <= 7:
                assert bank.value <= 1.25 + 1e-9       # natural order: the bit-reversed store of 32 / 128 points pays 2 wavefronts
                assert sh.value == ({5: 12, 6: 12, 7: 16}[e] if reorder else {5: 10, 6: 12, 7: 16}[e])  # float shuffles per thread: 6 per 4x4 transposition, 4 per bit swap
        if not reorder:
            # a tile that is only 8-byte aligned takes the 64-bit first read (rotated rows): same values, still conflict-free
            emu.emu_set_compat_misalign(1)
            out = np.zeros_like(x)
            bank = ctypes.c_double(0)
            assert emu.emu_run_compat_ct(x.ctypes.data, out.ctypes.data, e, nf, direction, reorder, 1, ctypes.byref(bank), None) == 0
            emu.emu_set_compat_misalign(0)
            assert O.rel_l2(out, want) < TOL
            assert bank.value <= {5: 1.25, 11: 1.17, 12: 1.51}.get(e, 1.0) + 1e-9, (e, bank.value)
        xs = (x / np.float32(n)).astype(np.complex64)
        out = np.zeros_like(xs)
        assert emu.emu_run_compat_ct(xs.ctypes.data, out.ctypes.data, e, nf, direction, reorder, 3, None, None) == 0
        w = xs.astype(np.complex128)
        for _ in range(3):
            w = O.ct_c2c_fp64(w, bool(direction), bool(reorder))
        assert O.rel_l2(out, w) < TOL
    assert warps >= 1


@pytest.mark.parametrize("e,which", [(7, 0), (7, 1), (8, 0), (8, 1), (9, 0), (9, 1), (10, 0), (10, 1), (12, 2), (5, 2)])
def test_alternate_instances(emu, e, which):
    """The product's alternates -- register-direct shapes A / B for 128..1024 points (cuFFT-shaped: global -> registers ->
    passes -> global, one CTA per tile), the R = 32 plan for 4096 points, one large CTA at 32 points -- i.e. every candidate
    the first-use selection (option "select") may pick: same values as the static table's instances, ragged batch tail."""
    n = 1 << e
    nf = 2 * max(1, 8192 // n) + 3
    x = O.uniform_c64(nf, n, seed=e + which)
    for direction in (0, 1):
        out = np.zeros_like(x)
        assert emu.emu_run_alternate(x.ctypes.data, out.ctypes.data, e, which, nf, direction, 3) == 0, (e, which)
        assert O.rel_l2(out, O.ct_c2c_fp64(x, bool(direction), True)) < TOL, (e, which, direction)


@pytest.mark.parametrize("e,which", [(13, 0), (13, 1), (13, 2), (13, 3), (14, 0), (14, 1), (14, 2), (14, 3), (14, 4), (14, 5)])
def test_8192_and_16384_point_c2c(emu, e, which):
    """8192 and 16384 points (beyond the reference's range): one transform per 64 / 128 KB tile -- R = 32 plan [32,32,8] with
    three buffers, R = 16 plan [16,16,16,4] with ONE buffer -- the tile moved as two / four 256-row TMA boxes; natural order
    and bit-reversed input, both directions, TMA and thread staging, ragged grid; 16384 points also with the results
    leaving from registers while the one buffer is refilled behind the final exchange (which = 4, 5)."""
    n = 1 << e
    nf = 5 if e == 13 else 3
    x = O.uniform_c64(nf, n, seed=which)
    reorder = which & 1
    for direction in (0, 1):
        out = np.zeros_like(x)
        assert emu.emu_run_alternate(x.ctypes.data, out.ctypes.data, e, which, nf, direction, 2) == 0
        assert O.rel_l2(out, O.ct_c2c_fp64(x, bool(direction), bool(reorder))) < TOL, (e, which, direction)


@pytest.mark.parametrize("variant,e,kind", [(0, 12, "c2c_fwd_r"), (1, 12, "c2c_inv_n"), (2, 10, "c2c_fwd_n"), (3, 8, "c2c_fwd_r"),
                                            (4, 10, "r2c"), (5, 11, "c2r"),
                                            (6, 10, "c2c_fwd_r"), (7, 9, "c2c_inv_r"), (8, 10, "r2c"), (9, 7, "r2c"),
                                            (10, 11, "c2r"), (11, 11, "c2r"), (12, 11, "c2r"),
                                            (13, 11, "c2c_fwd_r"), (14, 11, "c2c_inv_r"), (15, 7, "c2c_fwd_r"),
                                            (16, 12, "r2c"), (17, 9, "r2c"), (18, 10, "r2c"),
                                            (19, 12, "c2c_fwd_r"), (20, 12, "c2r"), (21, 10, "c2r"), (22, 10, "c2c_inv_r"),
                                            (23, 13, "r2c"), (24, 13, "c2r")])
def test_late_prefetch_points(emu, variant, e, kind):
    """The next tile's load issued after a later pass (kernel parameter PF): one persistent CTA over many
    tiles; a refill that lands in a buffer still being read shows up as wrong data in the emulator.
    Variants 6-9: the register-direct input path (IO_REG, measured slower than TMA staging, kept as an experiment).
    Variants 10-15: the reversed pass plan (small radix first) and the mirrored C2R head that it enables.
    Variants 16-18 (and 8): mirrored R2C ownership with several butterfly pairs per thread.
    Variants 19-22: reversed plans with a radix-4 first pass and the mirrored C2R head with 2 / 4 butterfly pairs.
    Variants 23-24: 16384 reals on the 8192-point core (three 64 KB buffers, two TMA boxes per tile), mirrored R2C / C2R."""
    n = 1 << e
    nf = max(7 * 4096 // n, 7) + 1
    if kind.startswith("c2c"):
        x = O.uniform_c64(nf, n)
        out = np.zeros_like(x)
        assert emu.emu_run_late(x.ctypes.data, out.ctypes.data, variant, nf, 1) == 0
        assert O.rel_l2(out, O.ct_c2c_fp64(x, "inv" in kind, kind.endswith("r"))) < TOL
    elif kind == "r2c":
        x = O.uniform_f32(nf, 2 * n)
        out = np.zeros((nf, n), np.complex64)
        assert emu.emu_run_late(x.ctypes.data, out.ctypes.data, variant, nf, 1) == 0
        assert O.rel_l2(out, O.r2c_packed_fp64(x)) < TOL
    else:
        h = O.uniform_c64(nf, n, seed=5)
        out = np.zeros_like(h)
        assert emu.emu_run_late(h.ctypes.data, out.ctypes.data, variant, nf, 1) == 0
        assert O.rel_l2(out.view(np.float32), O.c2r_packed_fp64(h)) < TOL


DUAL_SHAPES = [(8, 4), (9, 4), (10, 4), (11, 4), (12, 4), (9, 5), (10, 5), (11, 5)]
IO_NAMES = {IO_TMA: "tma", IO_LDG: "ldg", IO_TMA_STG: "tma_stg"}


def run_dual(lib, x, out, e, b, kind, io, tw=0, reps=1, grid=2, want_bank=True):
    bank = ctypes.c_double(0)
    n_ffts = x.size // (1 << e) if x.dtype == np.complex64 else x.size // (2 << e)
    rc = lib.emu_run_dual(x.ctypes.data, out.ctypes.data, e, b, n_ffts, kind, io, tw, reps, grid,
                          ctypes.byref(bank) if want_bank else None)
    assert rc == 0, f"dual configuration not instantiated: e={e} b={b} kind={kind} io={io} tw={tw} reps={reps}"
    return bank.value


@pytest.mark.parametrize("e,b", DUAL_SHAPES)
def test_dual_lane_c2c(emu, e, b):
    """block_fft_dual.cuh: two transforms per thread in the packed f32x2 lanes (cpair), chunk exchange layout.
    Odd transform counts leave lane 1 of the last pair on TMA zero fill / LDG guards."""
    n = 1 << e
    nf = max(4096 // n, 2) + 1   # one full tile and a ragged one with an odd transform count
    x = O.uniform_c64(nf, n)
    for kind, (direction, reorder) in enumerate(((0, 1), (1, 0), (0, 0), (1, 1))):
        ref = O.ct_c2c_fp64(x, bool(direction), bool(reorder))
        for io in (IO_TMA, IO_TMA_STG, IO_LDG):
            if (kind, io) in ((2, IO_TMA_STG), (3, IO_TMA_STG), (1, IO_LDG), (3, IO_LDG)):
                continue  # not instantiated in the emulator
            out = np.zeros_like(x)
            bank = run_dual(emu, x, out, e, b, kind, io, grid=1 if io == IO_TMA_STG else 2)
            assert O.rel_l2(out, ref) < TOL, (kind, IO_NAMES[io])
            if b == 4 or reorder == 1:  # R = 32: the contiguous first read of the no-reorder transform spans two rows (2-way)
                assert bank == pytest.approx(1.0), f"dual C2C accesses must be bank-conflict free ({kind}, {IO_NAMES[io]}): {bank}"
    out = np.zeros_like(x)
    run_dual(emu, x, out, e, b, 0, IO_TMA, tw=1)
    assert O.rel_l2(out, O.ct_c2c_fp64(x, False, True)) < TOL


@pytest.mark.parametrize("e,b", DUAL_SHAPES)
def test_dual_lane_r2c_c2r(emu, e, b):
    n = 2 << e
    nf = max(4096 // (n // 2), 2) + 1
    x = O.uniform_f32(nf, n)
    h = O.uniform_c64(nf, n // 2, seed=5)
    for io in (IO_TMA, IO_TMA_STG, IO_LDG):
        y = np.zeros((nf, n // 2), np.complex64)
        bank = run_dual(emu, x, y, e, b, 4, io, grid=1 if io == IO_TMA_STG else 2)
        assert O.rel_l2(y, O.r2c_packed_fp64(x)) < TOL, IO_NAMES[io]
        assert bank < 1.2, f"R2C: only the descending partner / mirror runs may conflict ({bank})"
        z = np.zeros_like(h)
        run_dual(emu, h, z, e, b, 5, io, grid=1 if io == IO_TMA_STG else 2)
        assert O.rel_l2(z.view(np.float32), O.c2r_packed_fp64(h)) < TOL, IO_NAMES[io]
    y = np.zeros((nf, n // 2), np.complex64)
    run_dual(emu, x, y, e, b, 4, IO_TMA, tw=1)
    assert O.rel_l2(y, O.r2c_packed_fp64(x)) < TOL
    z = np.zeros_like(h)
    run_dual(emu, h, z, e, b, 5, IO_LDG, tw=1)
    assert O.rel_l2(z.view(np.float32), O.c2r_packed_fp64(h)) < TOL


@pytest.mark.parametrize("e,b", [(8, 4), (10, 4), (11, 4), (10, 5)])
def test_dual_lane_multiple_and_one_hot(emu, e, b):
    n = 1 << e
    x = (O.uniform_c64(5, n) / np.float32(n)).astype(np.complex64)
    y = np.zeros_like(x)
    run_dual(emu, x, y, e, b, 0, IO_LDG, reps=3)
    assert O.rel_l2(y, np.fft.fft(np.fft.fft(np.fft.fft(x.astype(np.complex128))))) < TOL
    ref0 = x.astype(np.complex128)
    for _ in range(3):
        ref0 = O.ct_c2c_fp64(ref0, False, False)
    y0 = np.zeros_like(x)
    run_dual(emu, x, y0, e, b, 2, IO_LDG, reps=3)
    assert O.rel_l2(y0, ref0) < TOL
    xr = (O.uniform_f32(5, 2 * n) / np.float32(n)).astype(np.float32)
    yr = np.zeros((5, n), np.complex64)
    run_dual(emu, xr, yr, e, b, 4, IO_LDG, reps=3)  # R2C re-applied to its own packed output, as the kernel does
    cur = xr
    for _ in range(3):
        cur = O.r2c_packed_fp64(cur.astype(np.float32) if cur.dtype != np.float32 else cur).astype(np.complex64).view(np.float32)
    assert O.rel_l2(yr, cur.view(np.complex64)) < 1e-4
    # exact permutation of the no-reorder transform, both lanes: delta_p -> W^{brev(p) k}
    perm = O.c_reorder_index(n)
    ps = np.unique(np.concatenate([np.arange(min(n, 40)), np.random.default_rng(e).integers(0, n, 9), [n - 1]]))
    oh = np.zeros((len(ps), n), np.complex64)
    oh[np.arange(len(ps)), ps] = 1
    out = np.zeros_like(oh)
    run_dual(emu, oh, out, e, b, 2, IO_TMA)
    for row, p in enumerate(ps):
        assert int(np.rint((-np.angle(out[row, 1]) * n / (2 * np.pi)))) % n == perm[p]


def test_product_flavours(emu):
    """Guards the tuning table: which product instances use the packed add/subtract, the reversed plan and the mirrored
    ownership of the real passes (a silent change of B or of a plan would lose the mirror without failing any numerics)."""
    emu.emu_product_flavour.argtypes = [ctypes.c_int, ctypes.c_int]

    def f(e, mode):
        v = emu.emu_product_flavour(e, mode)
        assert v >= 0
        return {"arith": v & 15, "B": (v >> 4) & 15, "mirror_r2c": (v >> 8) & 1, "mirror_c2r": (v >> 9) & 1, "threads": v >> 12}

    # R2C: mirrored at 4096 reals (R = 16, [16,16,8]) and 8192 reals (R = 32, [32,32,4]), packed add/sub there and at 64 reals
    assert [f(e, R2C)["mirror_r2c"] for e in range(5, 13)] == [0, 0, 0, 0, 0, 0, 1, 1]
    assert f(11, R2C)["B"] == 4 and f(12, R2C)["B"] == 5
    assert [f(e, R2C)["arith"] for e in range(5, 13)] == [2, 0, 0, 0, 0, 0, 2, 2]
    # C2R: packed add/sub everywhere; reversed plan + mirror at 4096 ([8,16,16]) and 8192 reals ([4,32,32])
    assert [f(e, C2R)["arith"] for e in range(5, 13)] == [2, 2, 2, 2, 2, 2, 6, 6]
    assert [f(e, C2R)["mirror_c2r"] for e in range(5, 13)] == [0, 0, 0, 0, 0, 0, 1, 1]
    assert f(11, C2R)["B"] == 4 and f(12, C2R)["B"] == 5
    # register-direct alternates (first-use selection candidates, io = 4 / 5): shapes A and B exist for 128..1024 points only;
    # 256 points is the one size whose STATIC default is register-direct (it wins in every measured regime)
    emu.emu_regdirect_info.argtypes = [ctypes.c_int, ctypes.c_int]
    rd = {e: emu.emu_regdirect_info(e, 0) for e in range(5, 13)}
    assert [rd[e] & 1 for e in range(5, 13)] == [0, 0, 1, 1, 1, 1, 0, 0]            # ON
    assert [(rd[e] >> 2) & 1 for e in range(5, 13)] == [0, 0, 1, 1, 1, 1, 0, 0]     # ON_B
    assert [(rd[e] >> 1) & 1 for e in range(5, 13)] == [0, 0, 0, 1, 0, 0, 0, 0]     # PREFER: 256 points only
    assert [(rd[e] >> 4) & 15 for e in (7, 8, 9, 10)] == [4, 4, 5, 5]               # points per thread like cuFFT's kernels (16, 16, 32, 32)
    rdb = {e: emu.emu_regdirect_info(e, 1) for e in (9, 10)}
    assert all(((v >> 8) & 255) - ((v >> 4) & 15) == 5 for v in rdb.values())       # shape B at 512 / 1024 points: one warp per CTA
    # C2C natural order: R = 32 single-exchange plans at 512 / 1024 points, packed R = 16 at 2048 / 4096 points
    assert [f(e, C2C)["B"] for e in range(5, 13)] == [4, 4, 4, 4, 5, 5, 4, 4]
    assert [f(e, C2C)["arith"] for e in range(5, 13)] == [0, 0, 0, 0, 0, 0, 2, 2]
