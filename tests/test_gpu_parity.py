"""GPU parity tests proper: the CUDA path, called through the C ABI (libsmfft.so), against
  (1) the CPU oracle (oracle/smfft_oracle.c) and the FP64 closed forms, on seeded inputs,
  (2) the committed fixtures produced by the reference's own kernels (tests/golden/ref_*.npz),
  (3) the reference's kernels themselves (oracle/_ref, rebuilt for sm_100a) on the same device buffers,
  (4) size-independent properties at BASELINE.json's full 4 GiB batch.
Tolerance: relative L2 <= 1e-5 (north_star); the reorder permutation must match exactly by integer index.
"""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import oracle_np as O  # noqa: E402
from tests import refkernels as R  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-5
SIZES = [32, 64, 128, 256, 512, 1024, 2048, 4096]


@pytest.fixture(scope="module")
def sm():
    import smfft_b200

    assert os.path.exists(smfft_b200.lib_path()), "libsmfft.so missing: the CUDA extension must be built in-tree"
    smfft_b200.FFT_init()
    yield smfft_b200
    smfft_b200.set_option("io", 0)
    smfft_b200.set_option("twiddle", 0)
    smfft_b200.set_option("quirk_4096", 0)


def to_dev(a):
    a = np.ascontiguousarray(a)
    if np.iscomplexobj(a):
        return torch.from_numpy(a.view(np.float32).reshape(a.shape + (2,))).cuda()
    return torch.from_numpy(a).cuda()


def c64(t):
    return t.cpu().numpy().view(np.complex64).reshape(t.shape[:-1])


def run_c2c(sm, x, inverse, reorder):
    dx = to_dev(x)
    dy = torch.zeros_like(dx)
    sm.exec_c2c(dx, dy, x.shape[-1], x.shape[0], inverse, reorder)
    torch.cuda.synchronize()
    return c64(dy)


@pytest.mark.parametrize("io", [0, 1, 2, 3, 4, 5])   # auto / thread-staged / TMA in+out / TMA in, registers out / register-direct shapes A, B (128..1024 natural, else default)
@pytest.mark.parametrize("tw", [0, 1])
@pytest.mark.parametrize("n", SIZES)
def test_c2c_vs_oracle(sm, n, io, tw):
    sm.set_option("io", io)
    sm.set_option("twiddle", tw)
    nf = 3 * (8192 // n) + 5  # several tiles plus a ragged tail
    x = O.uniform_c64(nf, n)
    for inverse in (False, True):
        for reorder in (True, False):
            y = run_c2c(sm, x, inverse, reorder)
            assert O.rel_l2(y, O.c_ct_c2c(x, inverse, reorder)) < TOL          # the CPU restatement
            assert O.rel_l2(y, O.ct_c2c_fp64(x, inverse, reorder)) < TOL       # FP64 host DFT
    sm.set_option("io", 0)
    sm.set_option("twiddle", 0)


@pytest.mark.parametrize("io", [0, 2, 3])
def test_16384_points_many_tiles_per_cta(sm, io):
    """16384 points with more tiles than SMs, so every persistent CTA refills its ONE 128 KB buffer: through the TMA store
    path (io = 2) and with the results leaving from registers while the same buffer is refilled behind the final exchange
    (io = 3).  Both orders and directions vs FP64; the two stagings agree bit for bit."""
    n, nf = 16384, 2 * 148 + 37
    x = O.uniform_c64(nf, n, seed=77 + io)
    sm.set_option("io", io)
    try:
        for inverse, reorder in ((False, True), (True, False), (False, False)):
            y = run_c2c(sm, x, inverse, reorder)
            assert O.rel_l2(y, O.ct_c2c_fp64(x, inverse, reorder)) < TOL, (io, inverse, reorder)
            if io:
                sm.set_option("io", 5 - io)
                assert np.array_equal(run_c2c(sm, x, inverse, reorder), y)
                sm.set_option("io", io)
    finally:
        sm.set_option("io", 0)


@pytest.mark.parametrize("n", SIZES)
def test_reorder_permutation_exact(sm, n):
    perm = O.c_reorder_index(n)
    x = np.eye(n, dtype=np.complex64)  # delta_p for EVERY p
    k1 = 1
    y0 = run_c2c(sm, x, False, False)
    y1 = run_c2c(sm, x, False, True)
    slot0 = np.rint(-np.angle(y0[:, k1]) * n / (2 * np.pi)).astype(np.int64) % n
    slot1 = np.rint(-np.angle(y1[:, k1]) * n / (2 * np.pi)).astype(np.int64) % n
    assert np.array_equal(slot0, perm)            # no-reorder: DFT of the bit-reversed input
    assert np.array_equal(slot1, np.arange(n))    # reorder: natural order


def test_quirk_4096_switch(sm):
    x = O.uniform_c64(2, 4096)
    sm.set_option("quirk_4096", 1)
    q = run_c2c(sm, x, True, False)
    sm.set_option("quirk_4096", 0)
    assert O.rel_l2(q, O.ct_c2c_fp64(x, False, False)) < TOL     # CT/SM_FFT_parameters.cuh:388
    assert O.rel_l2(run_c2c(sm, x, True, False), O.ct_c2c_fp64(x, True, False)) < TOL


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
@pytest.mark.parametrize("io", [0, 1, 2, 3])
def test_r2c_c2r_vs_oracle(sm, n, io):
    sm.set_option("io", io)
    nf = 2 * (8192 // n) + 3
    x = O.uniform_f32(nf, n)
    dx = to_dev(x)
    dy = torch.zeros((nf, n // 2, 2), dtype=torch.float32, device="cuda")
    sm.exec_r2c_c2r(dx, dy, n, nf, 0)
    torch.cuda.synchronize()
    y = c64(dy)
    assert O.rel_l2(y, O.c_r2c(x)) < TOL
    assert O.rel_l2(y, O.r2c_packed_fp64(x)) < TOL
    dz = torch.zeros_like(dx)
    sm.exec_r2c_c2r(dy, dz, n, nf, 1)
    torch.cuda.synchronize()
    assert O.rel_l2(dz.cpu().numpy() / (n / 2), x) < TOL                 # round trip = N/2 * identity
    h = O.uniform_c64(nf, n // 2, seed=5)
    dh = to_dev(h)
    sm.exec_r2c_c2r(dh, dz, n, nf, 1)
    torch.cuda.synchronize()
    assert O.rel_l2(dz.cpu().numpy(), O.c2r_packed_fp64(h)) < TOL
    sm.set_option("io", 0)


GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.npz")))


@pytest.mark.parametrize("path", GOLDEN or [None])
def test_vs_reference_golden(sm, path):
    if path is None:
        pytest.skip("no reference fixtures committed yet")
    g = np.load(path)
    kind, x, want = str(g["kind"]), g["input"], g["output"]
    if kind == "ct":
        sm.set_option("quirk_4096", 1)   # fixtures come from the reference, quirk included
        got = run_c2c(sm, x, bool(g["inverse"]), bool(g["reorder"]))
        sm.set_option("quirk_4096", 0)
    elif kind == "stockham":
        got = run_c2c(sm, x, True, True)
    else:
        n = x.shape[-1] if kind == "r2c" else 2 * x.shape[-1]
        dx = to_dev(x)
        dy = torch.zeros((x.shape[0], n // 2, 2) if kind == "r2c" else (x.shape[0], n), dtype=torch.float32, device="cuda")
        sm.exec_r2c_c2r(dx, dy, n, x.shape[0], 0 if kind == "r2c" else 1)
        torch.cuda.synchronize()
        got = c64(dy) if kind == "r2c" else dy.cpu().numpy()
    assert O.rel_l2(got, want) < TOL


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n", SIZES)
def test_vs_reference_kernels_live(sm, n):
    """Same device buffers through the reference's FFT_external_benchmark (CT:583) and ours."""
    nf = 4096
    x = O.uniform_c64(nf, n, seed=n)
    dx = to_dev(x)
    for inverse in (False, True):
        for reorder in (True, False):
            if n == 4096 and inverse and not reorder:
                continue  # reference instance runs the forward transform (quirk, covered above)
            dref = torch.zeros_like(dx)
            dour = torch.zeros_like(dx)
            R.ct_external(dx, dref, n, nf, inverse, reorder)
            ms = sm.FFT_external_benchmark(dx, dour, n, nf, inverse, reorder)
            torch.cuda.synchronize()
            assert ms > 0
            assert O.rel_l2(c64(dour), c64(dref)) < TOL
            if reorder:
                assert O.c_ref_compare(c64(dour), c64(dref)) == 0   # the reference's own pass/fail (CT/FFT.c:52-77)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_stockham_and_r2c_vs_reference_kernels_live(sm):
    for n in (256, 1024, 4096):
        x = O.uniform_c64(512, n, seed=3 * n)
        dx = to_dev(x)
        dref, dour = torch.zeros_like(dx), torch.zeros_like(dx)
        R.st_external(dx, dref, n, 512)
        sm.Stockham_external_benchmark(dx, dour, n, 512, True)
        torch.cuda.synchronize()
        assert O.rel_l2(c64(dour), c64(dref)) < TOL
    for n in (512, 2048, 4096):
        xr = O.uniform_f32(512, n, seed=5 * n)
        dx = to_dev(xr)
        dref = torch.zeros((512, n // 2, 2), dtype=torch.float32, device="cuda")
        dour = torch.zeros_like(dref)
        R.rc_external(dx, dref, n, 512, 0)
        sm.R2C_C2R_external_benchmark(dx, dour, n, 512, 0)
        torch.cuda.synchronize()
        assert O.rel_l2(c64(dour), c64(dref)) < TOL
        back_ref, back_our = torch.zeros_like(dx), torch.zeros_like(dx)
        R.rc_external(dref, back_ref, n, 512, 1)
        sm.R2C_C2R_external_benchmark(dref, back_our, n, 512, 1)
        torch.cuda.synchronize()
        assert O.rel_l2(back_our.cpu().numpy(), back_ref.cpu().numpy()) < TOL


def test_multiple_benchmark_contract(sm):
    """FFT_multiple: timing-only (values overflow by design); check the contract and that
    small rep counts of the same code path are exact in the emulator tests.  nFFTs < 100 -> error, *ms = -1."""
    x = torch.rand((100 * 8192, 2), device="cuda")
    y = torch.empty_like(x)
    for n in (32, 1024, 4096):
        ms = sm.FFT_multiple_benchmark(x, y, n, (100 * 8192) // n, False, True)
        assert ms > 0
    with pytest.raises(sm.SmfftError):
        sm.FFT_multiple_benchmark(x, y, 1024, 99, False, True)
    ms = sm.R2C_multiple_benchmark(x, y, 2048, (100 * 8192 * 2) // 2048)
    assert ms > 0


def test_host_drivers(sm):
    n, nf = 1024, 4096
    x = O.uniform_c64(nf, n)
    out = np.zeros_like(x)
    single, multi = sm.c2c_host(x, out, n, nf, False, True, nRuns=2)   # GPU_smFFT_4elements (CT:827-908)
    assert single > 0 and multi > 0
    assert O.rel_l2(out, O.ct_c2c_fp64(x, False, True)) < TOL
    hx = torch.from_numpy(x.view(np.float32).reshape(nf, n, 2)).pin_memory()
    hy = torch.zeros_like(hx).pin_memory()
    ms = sm.pipeline_host(hx, hy, n, nf, False, False, 0, 1000)        # ragged chunks
    assert ms > 0
    assert O.rel_l2(c64(hy), O.ct_c2c_fp64(x, False, False)) < TOL


def test_full_size_properties_4GiB(sm):
    """BASELINE.json configs[1] sizes: 2^29 points (4 GiB in + 4 GiB out).  Size-independent checks:
    inverse(forward(x)) == N x, Parseval on a sample of rows, and a sampled comparison with the oracle."""
    pts = 1 << 29
    gen = torch.Generator(device="cuda")
    gen.manual_seed(20260101)
    x = torch.rand((pts, 2), device="cuda", generator=gen)
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    for n in (32, 1024, 4096):
        nf = pts // n
        sm.exec_c2c(x, y, n, nf, False, True)
        sm.exec_c2c(y, z, n, nf, True, True)
        torch.cuda.synchronize()
        z.div_(n)
        err = (torch.linalg.vector_norm((z - x).double()) / torch.linalg.vector_norm(x.double())).item()
        assert err < TOL
        # sampled rows against the CPU oracle (first, middle, last FFTs of the batch)
        for row0 in (0, nf // 2 - 3, nf - 8):
            xs = c64(x.view(nf, n, 2)[row0:row0 + 8])
            ys = c64(y.view(nf, n, 2)[row0:row0 + 8])
            assert O.rel_l2(ys, O.c_ct_c2c(xs, False, True)) < TOL
        # no-reorder at full size: equals the reorder result of the permuted rows (sampled)
        sm.exec_c2c(x, z, n, nf, False, False)
        torch.cuda.synchronize()
        xs = c64(x.view(nf, n, 2)[nf - 8:])
        assert O.rel_l2(c64(z.view(nf, n, 2)[nf - 8:]), O.ct_c2c_fp64(xs, False, False)) < TOL


# ---- the reference's device API (include/smfft/compat.cuh) used from a "user program" ------------------

@pytest.fixture(scope="module")
def compat():
    import ctypes

    from tests.compat.build_compat import build

    lib = ctypes.CDLL(build())
    P, I = ctypes.c_void_p, ctypes.c_int
    lib.compat_ct_external.argtypes = [P, P, I, I, I, I]
    lib.compat_ct_multiple.argtypes = [P, P, I, I, I, I]
    lib.compat_stockham_external.argtypes = [P, P, I, I]
    lib.compat_r2c_c2r_external.argtypes = [P, P, I, I, I]
    lib.compat_user_convolve_1024.argtypes = [P, P, P, I]
    lib.compat_user_convolve.argtypes = [P, P, P, I, I]
    lib.compat_stockham_multiple.argtypes = [P, P, I, I]
    lib.compat_r2c_multiple.argtypes = [P, P, I, I]
    lib.native_fft_launch.argtypes = [P, P, I, I, I, I, P]
    lib.native_convolve_launch.argtypes = [P, P, P, I, I, I, P]
    return lib


@pytest.mark.parametrize("n", SIZES)
def test_compat_wrapper_kernels_by_reference_names(compat, n):
    """SMFFT_DIT_external<FFT_N_{forward,inverse}[_noreorder]> launched with the reference's grid/block."""
    nf = 64
    x = O.uniform_c64(nf, n, seed=n + 1)
    dx = to_dev(x)
    for inverse in (0, 1):
        for reorder in (1, 0):
            dy = torch.zeros_like(dx)
            assert compat.compat_ct_external(dx.data_ptr(), dy.data_ptr(), n, nf, inverse, reorder) == 0
            torch.cuda.synchronize()
            assert O.rel_l2(c64(dy), O.ct_c2c_fp64(x, bool(inverse), bool(reorder))) < TOL
            assert O.rel_l2(c64(dy), O.c_ct_c2c(x, bool(inverse), bool(reorder))) < TOL
    big = torch.rand((200 * n, 2), device="cuda")
    out = torch.empty_like(big)
    assert compat.compat_ct_multiple(big.data_ptr(), out.data_ptr(), n, 400 if n > 64 else 800, 0, 1) == 0
    torch.cuda.synchronize()


def test_compat_stockham_and_r2c_c2r(compat):
    for n in (256, 512, 1024, 2048, 4096):
        x = O.uniform_c64(32, n, seed=n)
        dx = to_dev(x)
        dy = torch.zeros_like(dx)
        assert compat.compat_stockham_external(dx.data_ptr(), dy.data_ptr(), n, 32) == 0
        torch.cuda.synchronize()
        assert O.rel_l2(c64(dy), O.stockham_c2c_fp64(x, True)) < TOL           # mk6 is inverse-only
    for n in (256, 512, 1024, 2048, 4096):
        xr = O.uniform_f32(32, n, seed=n)
        dx = to_dev(xr)
        dy = torch.zeros((32, n // 2, 2), dtype=torch.float32, device="cuda")
        assert compat.compat_r2c_c2r_external(dx.data_ptr(), dy.data_ptr(), n, 32, 0) == 0
        back = torch.zeros_like(dx)
        assert compat.compat_r2c_c2r_external(dy.data_ptr(), back.data_ptr(), n, 32, 1) == 0
        torch.cuda.synchronize()
        assert O.rel_l2(c64(dy), O.r2c_packed_fp64(xr)) < TOL
        assert O.rel_l2(back.cpu().numpy() / (n / 2), xr) < TOL


def test_compat_device_function_inside_a_user_kernel(compat):
    """load -> do_SMFFT_CT_DIT<forward> -> pointwise filter -> do_SMFFT_CT_DIT<inverse> -> store in ONE
    launch: the use case SMFFT exists for (README.md:2, 10-14)."""
    n, nf = 1024, 48
    x = O.uniform_c64(nf, n, seed=9)
    rng = np.random.default_rng(4)
    h = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    dx, dh = to_dev(x), to_dev(h)
    dy = torch.zeros_like(dx)
    assert compat.compat_user_convolve_1024(dx.data_ptr(), dh.data_ptr(), dy.data_ptr(), nf) == 0
    torch.cuda.synchronize()
    want = np.fft.ifft(np.fft.fft(x.astype(np.complex128), axis=-1) * h.astype(np.complex128), axis=-1)
    assert O.rel_l2(c64(dy), want) < TOL


@pytest.mark.parametrize("prog,args", [("ct", "1024 2000 2 0 1"), ("ct", "32 1001 1 1 1"), ("ct", "4096 300 1 0 1"),
                                       ("st", "2048 500 2"), ("rc", "4096 400 2"), ("rc", "512 1000 1")])
def test_reference_host_programs_run_unmodified(prog, args):
    """The reference's own FFT.c (unmodified, built by oracle/Makefile `refmain`) linked against
    libsmfft_compat.so: its self-check against cuFFT must print PASSED (CT/FFT.c:154-160)."""
    import subprocess

    exe = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", f"FFT_{prog}.exe")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/FFT_*.exe not built")
    r = subprocess.run([exe] + args.split(), capture_output=True, text=True, timeout=300)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out
    assert "FAILED" not in out and "Error" not in out, out
    if prog != "st":  # the Stockham program only self-checks when built with TESTING (ST/debug.h:3, ST/FFT.c:138-144)
        assert "PASSED" in out, out


@pytest.mark.parametrize("n", [32, 256, 1024, 4096])
def test_in_place_and_stream_and_errors(sm, n):
    """d_output == d_input is safe (a tile is fully staged in shared memory before anything is written back);
    launches follow torch's current stream; bad arguments come back as errors, not crashes."""
    nf = 3 * (8192 // n) + 1
    x = O.uniform_c64(nf, n, seed=n + 7)
    d = to_dev(x)
    sm.exec_c2c(d, d, n, nf, False, True)                      # in place
    torch.cuda.synchronize()
    assert O.rel_l2(c64(d), O.ct_c2c_fp64(x, False, True)) < TOL
    side = torch.cuda.Stream()
    d2, out = to_dev(x), torch.zeros_like(d)
    with torch.cuda.stream(side):
        sm.exec_c2c(d2, out, n, nf, True, False)
    side.synchronize()
    assert O.rel_l2(c64(out), O.ct_c2c_fp64(x, True, False)) < TOL
    with pytest.raises(sm.SmfftError):
        sm.exec_c2c(d2, out, 48, nf, False, True)              # wrong FFT length (CT:656-658)
    with pytest.raises(sm.SmfftError):
        sm.exec_c2c(d2.data_ptr() + 8, out, n, nf - 1, False, True)   # misaligned device pointer
    sm.exec_c2c(d2, out, n, 0, False, True)                    # empty batch is a no-op


@pytest.mark.parametrize("argv", [["c2c", "1024", "3000", "2", "0", "1"], ["c2c", "32", "1001", "1", "1", "0"],
                                  ["stockham", "4096", "300", "1"], ["r2c", "2048", "1500", "2"]])
def test_cli_mirror_of_the_reference_programs(argv, capsys):
    """python -m smfft_b200.cli: the reference's FFT.exe argument conventions, seeded data, relative-L2 verdict."""
    from smfft_b200 import cli

    assert cli.main(argv) == 0
    assert "FFT test: PASSED" in capsys.readouterr().out


def test_32GiB_batch_64bit_indexing(sm):
    """BASELINE.json configs[4]: 32 GiB on one GPU = 2^32 points, beyond the reference's 32-bit indexing
    (SURVEY.md 0-9).  Rows sampled from the start, across the 2^31 / 2^32-byte boundaries and the very end."""
    free, _ = torch.cuda.mem_get_info()
    if free < 70 * (1 << 30):
        pytest.skip("needs 64 GiB of free device memory")
    pts = 1 << 32
    gen = torch.Generator(device="cuda")
    gen.manual_seed(7)
    x = torch.empty((pts, 2), device="cuda")
    for i in range(8):  # generate in slices (torch.rand of 2^33 elements at once is fine too, this bounds temp memory)
        x[i * (pts // 8):(i + 1) * (pts // 8)].uniform_(0, 1, generator=gen)
    y = torch.empty_like(x)
    for n, reorder in ((4096, True), (32, False), (1024, True)):
        nf = pts // n
        ms = sm.FFT_external_benchmark(x, y, n, nf, False, reorder)
        torch.cuda.synchronize()
        assert ms > 0
        for row0 in (0, (1 << 28) // n, (1 << 31) // n - 2, (1 << 31) // n + 3, nf // 2 + 5, nf - 4):
            xs = c64(x.view(nf, n, 2)[row0:row0 + 4])
            ys = c64(y.view(nf, n, 2)[row0:row0 + 4])
            assert O.rel_l2(ys, O.ct_c2c_fp64(xs, False, reorder)) < TOL, (n, row0)
    xr = x.view(-1)
    n = 4096
    nf = 2 * pts // n
    sm.exec_r2c_c2r(xr, y, n, nf, 0)
    torch.cuda.synchronize()
    for row0 in (0, nf // 2 + 1, nf - 3):
        xs = xr.view(nf, n)[row0:row0 + 3].cpu().numpy()
        ys = c64(y.view(nf, n // 2, 2)[row0:row0 + 3])
        assert O.rel_l2(ys, O.r2c_packed_fp64(xs)) < TOL
    del x, y
    torch.cuda.empty_cache()


# ---- round 2: the holes the round-1 review named -----------------------------------------------------------------

@pytest.mark.parametrize("n", SIZES)
def test_repeated_path_values_c2c(sm, n):
    """The FFT_multiple kernels (in-place repetitions with their inter-rep barrier, CT:553-572) with THREE repetitions:
    F(F(F(x))) compared with the oracle applied three times -- for reorder, no-reorder, forward and inverse."""
    nf = 2 * (8192 // n) + 3     # several tiles plus a ragged tail
    x = (O.uniform_c64(nf, n, seed=n + 11) / np.float32(n)).astype(np.complex64)
    dx = to_dev(x)
    for inverse in (False, True):
        for reorder in (True, False):
            dy = torch.zeros_like(dx)
            sm.exec_repeated(dx, dy, n, nf, inverse, reorder, 0, 3)
            torch.cuda.synchronize()
            want64, want32 = x.astype(np.complex128), x
            for _ in range(3):
                want64 = O.ct_c2c_fp64(want64, inverse, reorder)
                want32 = O.c_ct_c2c(want32, inverse, reorder)       # the CPU restatement, fp32 like the kernels
            assert O.rel_l2(c64(dy), want64) < TOL, (n, inverse, reorder)
            assert O.rel_l2(c64(dy), want32) < TOL, (n, inverse, reorder)
    with pytest.raises(sm.SmfftError):
        sm.exec_repeated(dx, dy, n, nf, False, True, 0, 7)          # only the 3- and 100-rep instances exist


@pytest.mark.parametrize("n", [64, 256, 512, 1024, 2048, 4096, 8192])
def test_repeated_path_values_r2c(sm, n):
    """FFT_GPU_R2C_C2R_multiple (RC:367-384) re-reads its packed spectrum as reals on every repetition; three of them."""
    nf = 2 * (8192 // n) + 3
    x = (O.uniform_f32(nf, n, seed=n + 13) / np.float32(n)).astype(np.float32)
    dx = to_dev(x)
    dy = torch.zeros_like(dx)
    sm.exec_repeated(dx, dy, n, nf, False, True, 1, 3)
    torch.cuda.synchronize()
    want64, want32 = x.astype(np.float64), x
    for _ in range(3):
        want64 = np.ascontiguousarray(O.r2c_packed_fp64(want64)).view(np.float64).reshape(nf, n)
        want32 = np.ascontiguousarray(O.c_r2c(want32)).view(np.float32).reshape(nf, n)
    got = dy.cpu().numpy()
    assert O.rel_l2(got, want64) < TOL
    assert O.rel_l2(got, want32) < TOL


def test_full_size_properties_4GiB_stockham_and_real(sm):
    """BASELINE.json configs[2] and configs[3] at their full sizes: Stockham inverse on 2^29 points and R2C / C2R on 2^30
    reals (4 GiB), the sizes the bench times.  Sampled rows (first / middle / last transforms of the batch: the persistent
    grid's tail logic) against the oracle and FP64, and the round trips  forward(inverse(x)) = N x,  C2R(R2C(x)) = N/2 x."""
    pts = 1 << 29
    gen = torch.Generator(device="cuda")
    gen.manual_seed(20260102)
    x = torch.rand((pts, 2), device="cuda", generator=gen)
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    for n in (256, 2048, 4096):                                   # Stockham C2C, inverse (the reference's direction, ST:70-78)
        nf = pts // n
        ms = sm.Stockham_external_benchmark(x, y, n, nf, True)
        assert ms > 0
        sm.Stockham_external_benchmark(y, z, n, nf, False)
        torch.cuda.synchronize()
        z.div_(n)
        err = (torch.linalg.vector_norm((z - x).double()) / torch.linalg.vector_norm(x.double())).item()
        assert err < TOL, (n, err)
        for row0 in (0, nf // 2 - 3, nf - 8):
            xs = c64(x.view(nf, n, 2)[row0:row0 + 8])
            ys = c64(y.view(nf, n, 2)[row0:row0 + 8])
            assert O.rel_l2(ys, O.stockham_c2c_fp64(xs, True)) < TOL, (n, row0)
            assert O.rel_l2(ys, O.c_ct_c2c(xs, True, True)) < TOL, (n, row0)
    xr = x.view(-1)                                                # 2^30 reals
    zr = z.view(-1)
    for n in (512, 2048, 4096, 8192):                              # real transform lengths; 4096 is the reference's largest
        nf = 2 * pts // n
        sm.exec_r2c_c2r(xr, y, n, nf, 0)
        sm.exec_r2c_c2r(y, zr, n, nf, 1)
        torch.cuda.synchronize()
        zr.div_(n / 2)
        err = (torch.linalg.vector_norm((zr - xr).double()) / torch.linalg.vector_norm(xr.double())).item()
        assert err < TOL, (n, err)
        for row0 in (0, nf // 2 - 3, nf - 8):
            xs = xr.view(nf, n)[row0:row0 + 8].cpu().numpy()
            ys = c64(y.view(nf, n // 2, 2)[row0:row0 + 8])
            assert O.rel_l2(ys, O.c_r2c(xs)) < TOL, (n, row0)
            assert O.rel_l2(ys, O.r2c_packed_fp64(xs)) < TOL, (n, row0)
        # C2R of an arbitrary packed spectrum at full size (not only of an R2C result): sampled rows vs FP64
        sm.exec_r2c_c2r(x, zr, n, nf, 1)
        torch.cuda.synchronize()
        for row0 in (0, nf - 8):
            hs = c64(x.view(nf, n // 2, 2)[row0:row0 + 8])
            assert O.rel_l2(zr.view(nf, n)[row0:row0 + 8].cpu().numpy(), O.c2r_packed_fp64(hs)) < TOL, (n, row0)
    del x, y, z
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048])
@pytest.mark.parametrize("io", [0, 1, 2, 3, 4, 5])
def test_in_place_every_staging(sm, n, io):
    """d_output == d_input for the sizes and stagings the first in-place test left out: N = 128 (whose default stores
    from registers, IO_TMA_STG), N = 256 (whose default is register-direct) and every staging explicitly.  Safe because a
    tile is completely staged (TMA load, thread copy, or -- register-direct -- every thread's loads consumed by the first pass
    before the first barrier) before any of its results is written, and tiles do not overlap."""
    sm.set_option("io", io)
    nf = 3 * (8192 // n) + 1
    x = O.uniform_c64(nf, n, seed=n + 17)
    for inverse, reorder in ((False, True), (True, False)):
        d = to_dev(x)
        sm.exec_c2c(d, d, n, nf, inverse, reorder)
        torch.cuda.synchronize()
        assert O.rel_l2(c64(d), O.ct_c2c_fp64(x, inverse, reorder)) < TOL, (n, io, inverse, reorder)
    sm.set_option("io", 0)


@pytest.mark.parametrize("n", [64, 256, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("io", [0, 1, 2, 3])
def test_in_place_r2c_c2r(sm, n, io):
    """R2C and C2R in place: input and output of one transform occupy the same bytes (N reals <-> N/2 packed bins)."""
    sm.set_option("io", io)
    nf = 2 * (8192 // n) + 3
    x = O.uniform_f32(nf, n, seed=n + 19)
    d = to_dev(x)
    sm.exec_r2c_c2r(d, d, n, nf, 0)
    torch.cuda.synchronize()
    got = d.cpu().numpy().view(np.complex64).reshape(nf, n // 2)
    assert O.rel_l2(got, O.r2c_packed_fp64(x)) < TOL, (n, io)
    sm.exec_r2c_c2r(d, d, n, nf, 1)
    torch.cuda.synchronize()
    assert O.rel_l2(d.cpu().numpy() / (n / 2), x) < TOL, (n, io)
    sm.set_option("io", 0)


def test_host_threads_concurrently(sm):
    """The C ABI from several host threads at once (one thread per GPU where the box has more than one, else all on
    cuda:0, each on its own stream): per-thread streams and error text, per-device state behind its own lock, the
    pipeline context owned by the device (launch.cu).  Every thread checks its own results against FP64."""
    import threading

    ndev = torch.cuda.device_count()
    nthreads = 4
    errors = []

    def work(tid):
        try:
            dev = tid % ndev
            torch.cuda.set_device(dev)
            stream = torch.cuda.Stream(device=dev)
            for it in range(6):
                n = [32, 256, 1024, 4096][(tid + it) % 4]
                nf = 3 * (8192 // n) + tid + 1
                x = O.uniform_c64(nf, n, seed=100 * tid + it)
                with torch.cuda.stream(stream):
                    dx = to_dev(x).to(f"cuda:{dev}")
                    dy = torch.zeros_like(dx)
                    sm.exec_c2c(dx, dy, n, nf, bool(it & 1), bool(tid & 1))
                    ms = sm.FFT_external_benchmark(dx, dy, n, nf, bool(it & 1), bool(tid & 1))
                stream.synchronize()
                assert ms > 0
                err = O.rel_l2(c64(dy), O.ct_c2c_fp64(x, bool(it & 1), bool(tid & 1)))
                assert err < TOL, (tid, it, n, err)
                # a failing call on this thread must not disturb the others' error state or results
                with pytest.raises(sm.SmfftError, match="wrong FFT length"):
                    sm.exec_c2c(dx, dy, 48, 1, False, True)
                # the host pipeline: its context belongs to the device, concurrent calls on one device serialise
                hx = torch.from_numpy(x.view(np.float32).reshape(nf, n, 2)).pin_memory()
                hy = torch.zeros_like(hx).pin_memory()
                assert sm.pipeline_host(hx, hy, n, nf, False, True, 0, 7) > 0
                assert O.rel_l2(c64(hy), O.ct_c2c_fp64(x, False, True)) < TOL
        except BaseException as ex:  # noqa: BLE001
            errors.append((tid, repr(ex)))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(nthreads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    torch.cuda.set_device(0)
    assert not errors, errors
    sm.pipeline_release()
    sm.pipeline_release()   # idempotent


def test_python_wrappers_check_their_buffers(sm):
    """api.py refuses tensors that do not cover nFFTs * FFT_size elements, are not contiguous or live on the wrong side."""
    x = torch.zeros((100, 1024, 2), device="cuda")
    y = torch.zeros_like(x)
    with pytest.raises(sm.SmfftError, match="needs"):
        sm.exec_c2c(x, y, 1024, 101, False, True)
    with pytest.raises(sm.SmfftError, match="contiguous"):
        sm.exec_c2c(x.transpose(0, 1), y, 1024, 100, False, True)
    with pytest.raises(sm.SmfftError, match="CUDA"):
        sm.exec_c2c(x.cpu(), y, 1024, 100, False, True)
    with pytest.raises(sm.SmfftError, match="fft|FFT"):
        sm.pipeline_host(x.cpu(), y.cpu(), 0, 10)                  # validated before it is used as a divisor


# ---- the native device primitive (include/smfft/device.cuh) used from a user program ------------------------------------

@pytest.mark.parametrize("n", SIZES)
def test_native_device_primitive(sm, compat, n):
    """smfft::BlockFFT<log2 N, dir, FFTS, TW>::load / exec / store inside a user kernel (tests/compat/compat_kernels.cu):
    registers in, registers out, natural order, both directions, MUFU and table twiddles."""
    nf = 64
    x = O.uniform_c64(nf, n, seed=n + 23)
    dx = to_dev(x)
    tw = sm.twiddle_table()
    for inverse in (0, 1):
        want = O.ct_c2c_fp64(x, bool(inverse), True)
        for lut in (0, 1):
            dy = torch.zeros_like(dx)
            assert compat.native_fft_launch(dx.data_ptr(), dy.data_ptr(), n, nf, inverse, lut, tw) == 0
            torch.cuda.synchronize()
            assert O.rel_l2(c64(dy), want) < TOL, (n, inverse, lut)
            assert O.rel_l2(c64(dy), O.c_ct_c2c(x, bool(inverse), True)) < TOL


@pytest.mark.parametrize("n", [256, 1024, 4096])
def test_native_fused_convolution(sm, compat, n):
    """smfft::block_convolve: forward transform, pointwise functor on the registers, inverse transform, one launch -- and the
    same convolution through the drop-in device function (compat) for every size the bench times."""
    nf = 48
    x = O.uniform_c64(nf, n, seed=n + 29)
    rng = np.random.default_rng(n)
    h = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    dx, dh = to_dev(x), to_dev(h)
    want = np.fft.ifft(np.fft.fft(x.astype(np.complex128), axis=-1) * h.astype(np.complex128), axis=-1)
    tw = sm.twiddle_table()
    for lut in (0, 1):
        dy = torch.zeros_like(dx)
        assert compat.native_convolve_launch(dx.data_ptr(), dh.data_ptr(), dy.data_ptr(), n, nf, lut, tw) == 0
        torch.cuda.synchronize()
        assert O.rel_l2(c64(dy), want) < TOL, (n, lut)
    dy = torch.zeros_like(dx)
    assert compat.compat_user_convolve(dx.data_ptr(), dh.data_ptr(), dy.data_ptr(), n, nf) == 0
    torch.cuda.synchronize()
    assert O.rel_l2(c64(dy), want) < TOL


def test_compat_multiple_wrappers_run(compat):
    """FFT_GPU_multiple / FFT_GPU_R2C_C2R_multiple by the reference's names and launch shapes (timing-only kernels)."""
    big = torch.rand((400 * 4096, 2), device="cuda")
    out = torch.empty_like(big)
    for n in (256, 1024, 4096):
        assert compat.compat_stockham_multiple(big.data_ptr(), out.data_ptr(), n, 400) == 0
        assert compat.compat_r2c_multiple(big.data_ptr(), out.data_ptr(), n, 400) == 0
    torch.cuda.synchronize()


def test_first_use_selection(sm):
    """Option "select" = 1: the first large out-of-place call of a transform times the table's instance against its
    alternates on the caller's batch and keeps the fastest; results stay correct whichever wins, in-place and small calls
    are never timed, the report lists every candidate, "select_reset" forgets the decisions."""
    sm.set_option("select_reset", 1)
    sm.set_option("select", 1)
    sm.set_option("select_min_log2_points", 20)
    try:
        before = sm.launch_count()
        for n in (32, 128, 512, 1024, 2048, 4096):
            nf = (1 << 21) // n
            x = O.uniform_c64(nf, n, seed=n + 31)
            dx = to_dev(x)
            dy = torch.zeros_like(dx)
            sm.exec_c2c(dx, dy, n, nf, False, True)          # tunes (where alternates exist), then runs
            torch.cuda.synchronize()
            got = c64(dy)
            rows = [0, nf // 2, nf - 1]
            assert O.rel_l2(got[rows], O.ct_c2c_fp64(x[rows], False, True)) < TOL, n
            dy.zero_()
            sm.exec_c2c(dx, dy, n, nf, False, True)          # served by the selected instance
            torch.cuda.synchronize()
            assert O.rel_l2(c64(dy)[rows], O.ct_c2c_fp64(x[rows], False, True)) < TOL, n
        rep = sm.select_report()
        lines = [ln for ln in rep.splitlines() if ln.strip()]
        assert len(lines) == 6, rep
        assert sum(ln.count(":") >= 2 for ln in lines) >= 5, rep          # 32, 128, 512, 1024, 4096 have alternates (2048 has none)
        assert sm.launch_count() - before > 12 + 5 * 13                   # the timing launches were really made
        # in place and small batches: no timing
        n0 = sm.launch_count()
        d = to_dev(O.uniform_c64((1 << 21) // 256, 256, seed=3))
        sm.exec_c2c(d, d, 256, (1 << 21) // 256, False, True)
        small = to_dev(O.uniform_c64(64, 256, seed=4))
        sm.exec_c2c(small, torch.zeros_like(small), 256, 64, False, True)
        assert sm.launch_count() - n0 == 2
        assert "e 8 " not in sm.select_report()
        sm.set_option("select_reset", 1)
        assert sm.select_report() == ""
    finally:
        sm.set_option("select", 0)
        sm.set_option("select_min_log2_points", 24)
        sm.set_option("select_reset", 1)


@pytest.mark.parametrize("io", [0, 1, 2, 3])
@pytest.mark.parametrize("tw", [0, 1])
@pytest.mark.parametrize("n", [8192, 16384])
def test_c2c_beyond_the_reference(sm, n, io, tw):
    """8192 and 16384 points -- beyond the reference: one transform per 64 / 128 KB shared-memory tile (two /
    four 256-row TMA boxes), R = 32 plan [32,32,8] / R = 16 plan [16,16,16,4].  Both orders, both directions, vs the CPU
    restatement and FP64; ragged batch; in place."""
    sm.set_option("io", io)
    sm.set_option("twiddle", tw)
    nf = 7 if n == 8192 else 5
    x = O.uniform_c64(nf, n, seed=n + io)
    for inverse in (False, True):
        for reorder in (True, False):
            y = run_c2c(sm, x, inverse, reorder)
            assert O.rel_l2(y, O.ct_c2c_fp64(x, inverse, reorder)) < TOL, (io, tw, inverse, reorder)
            assert O.rel_l2(y, O.c_ct_c2c(x, inverse, reorder)) < TOL
    d = to_dev(x)
    sm.exec_c2c(d, d, n, nf, False, True)
    torch.cuda.synchronize()
    assert O.rel_l2(c64(d), O.ct_c2c_fp64(x, False, True)) < TOL
    with pytest.raises(sm.SmfftError):
        sm.FFT_multiple_benchmark(d, d, n, 200, False, True)      # the repeated benchmark stops at 4096 points
    with pytest.raises(sm.SmfftError):
        sm.exec_c2c(d, d, 1 << 25, 1, False, True)              # multi-pass transforms stop at 2^24 points
    sm.set_option("io", 0)
    sm.set_option("twiddle", 0)


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_randomized_configurations(sm, seed):
    """Seeded sweep over what a caller can combine: size (32..4194304), batch count (odd, prime, one, many tiles per CTA), staging
    (io 0..5), twiddle source, direction, order, in place or not, first-use selection on or off -- C2C against the FP64 DFT, and
    R2C / C2R (64..16384 reals) against the packed FP64 forms.  A mismatch names the configuration."""
    rng = np.random.default_rng(seed)
    try:
        for it in range(60):
            n = 1 << int(rng.integers(5, 23))                       # 32 .. 4194304 (two passes from 32768 up, three from 2097152)
            nf = int(rng.choice([1, 2, 3, 7, 31, 127, 149, 331, 1021])) if n >= 2048 else int(rng.choice([1, 3, 17, 257, 1031, 4099]))
            nf = max(1, min(nf, (1 << 22) // n))
            io, tw = int(rng.integers(0, 6)), int(rng.integers(0, 2))
            inverse, reorder, in_place = bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
            reorder = reorder or n >= (1 << 15)                      # the two-pass sizes take natural-order input only
            select = int(rng.integers(0, 4) == 0)
            sm.set_option("two_pass_chunk_mib", int(rng.choice([1, 4, 1024])))
            cfg = dict(n=n, nf=nf, io=io, tw=tw, inverse=inverse, reorder=reorder, in_place=in_place, select=select)
            sm.set_option("io", io)
            sm.set_option("twiddle", tw)
            sm.set_option("select", select)
            sm.set_option("select_min_log2_points", 10)
            x = O.uniform_c64(nf, n, seed=seed * 1000 + it)
            dx = to_dev(x)
            dy = dx if in_place else torch.zeros_like(dx)
            sm.exec_c2c(dx, dy, n, nf, inverse, reorder)
            torch.cuda.synchronize()
            assert O.rel_l2(c64(dy), O.ct_c2c_fp64(x, inverse, reorder)) < TOL, cfg
            if not in_place:
                assert np.array_equal(c64(dx), x), cfg            # the input is left alone
        for it in range(30):
            n = 1 << int(rng.integers(6, 15))                      # real length, 64 .. 16384
            nf = min(int(rng.choice([1, 3, 17, 149, 257, 1031])), (1 << 22) // n)
            io, tw = int(rng.integers(0, 5)), int(rng.integers(0, 2))
            cfg = dict(real_n=n, nf=nf, io=io, tw=tw)
            sm.set_option("io", io)
            sm.set_option("twiddle", tw)
            sm.set_option("select", 0)
            xr = O.uniform_f32(nf, n, seed=seed * 1000 + 500 + it)
            dr = to_dev(xr)
            dc = torch.zeros((nf, n // 2, 2), dtype=torch.float32, device="cuda")
            sm.exec_r2c_c2r(dr, dc, n, nf, 0)
            torch.cuda.synchronize()
            assert O.rel_l2(c64(dc), O.r2c_packed_fp64(xr)) < TOL, cfg
            h = O.uniform_c64(nf, n // 2, seed=seed * 1000 + 700 + it)
            dh = to_dev(h)
            dz = torch.zeros_like(dr)
            sm.exec_r2c_c2r(dh, dz, n, nf, 1)
            torch.cuda.synchronize()
            assert O.rel_l2(dz.cpu().numpy(), O.c2r_packed_fp64(h)) < TOL, cfg
    finally:
        sm.set_option("io", 0)
        sm.set_option("twiddle", 0)
        sm.set_option("select", 0)
        sm.set_option("select_min_log2_points", 24)
        sm.set_option("select_reset", 1)
        sm.set_option("two_pass_chunk_mib", 1024)


@pytest.mark.parametrize("n", [1 << e for e in range(15, 25)])
def test_two_pass_transforms(sm, n):
    """2^15 .. 2^24 points (beyond the reference, SURVEY.md 8f-4 "N > 4096 via multi-pass"): N = N1 N2 in two passes over HBM up
    to 2^20 points, N = N1 N2 N3 in three from 2^21 (csrc/big_fft.cu), strided TMA boxes, twiddles from a three-level
    FP64-rounded table.  Both directions against the FP64 DFT; batches of 1, 5 and 37 (fewer for the largest sizes); chunked
    (scratch smaller than the batch); in place; the host-timed entry point; error contract."""
    try:
        for nf in ((1, 5, 37) if n <= (1 << 18) else (1, 3) if n <= (1 << 22) else (2,)):
            x = O.uniform_c64(nf, n, seed=n % 1000 + nf)
            for inverse in (False, True):
                assert O.rel_l2(run_c2c(sm, x, inverse, True), O.ct_c2c_fp64(x, inverse, True)) < TOL, (n, nf, inverse)
        nb = 11 if n <= (1 << 18) else 3
        x = O.uniform_c64(nb, n, seed=3)
        want = O.ct_c2c_fp64(x, False, True)
        sm.set_option("two_pass_chunk_mib", 1)                       # 1 MiB of scratch: one transform (or a few) per chunk
        assert O.rel_l2(run_c2c(sm, x, False, True), want) < TOL
        d = to_dev(x)
        sm.exec_c2c(d, d, n, nb, False, True)                         # in place, chunked
        torch.cuda.synchronize()
        assert O.rel_l2(c64(d), want) < TOL
        sm.set_option("two_pass_chunk_mib", 1024)
        d = to_dev(x)
        out = torch.zeros_like(d)
        assert sm.FFT_external_benchmark(d, out, n, nb, False, True) > 0
        torch.cuda.synchronize()
        assert O.rel_l2(c64(out), want) < TOL
        assert np.array_equal(c64(d), x)                              # the input is left alone
        # one-hot at p: X[k] = W_N^(p k) -- pins the order of the output exactly
        p = 12345 % n
        hot = np.zeros((1, n), dtype=np.complex64)
        hot[0, p] = 1
        y = run_c2c(sm, hot, False, True)[0]
        k = np.arange(n)
        assert np.max(np.abs(y - np.exp(-2j * np.pi * ((p * k) % n) / n))) < 1e-5
        with pytest.raises(sm.SmfftError, match="natural order"):
            sm.exec_c2c(d, out, n, nb, False, False)                  # no bit-reversed-input flavour for multi-pass sizes
        with pytest.raises(sm.SmfftError):
            sm.FFT_multiple_benchmark(d, out, n, 200, False, True)
    finally:
        sm.set_option("two_pass_chunk_mib", 1024)


def test_two_pass_full_size_round_trip(sm):
    """65536 points on the full 4 GiB batch: inverse(forward(x)) = N x on rows sampled across the batch, and the forward
    spectrum of the first / middle / last transforms against FP64."""
    n = 1 << 16
    nf = (1 << 29) // n
    x = torch.rand((nf, n, 2), device="cuda")
    y = torch.empty_like(x)
    sm.exec_c2c(x, y, n, nf, False, True)
    rows = [0, 1, nf // 2, nf - 2, nf - 1]
    for r in rows:
        xr = c64(x[r:r + 1])
        assert O.rel_l2(c64(y[r:r + 1]), O.ct_c2c_fp64(xr, False, True)) < TOL, r
    sm.exec_c2c(y, y, n, nf, True, True)                              # in place, inverse
    torch.cuda.synchronize()
    for r in rows + list(range(7, nf, nf // 13)):
        assert O.rel_l2(c64(y[r:r + 1]) / n, c64(x[r:r + 1])) < TOL, r


def test_two_pass_through_the_host_pipeline(sm):
    """65536 points through smfft_pipeline_host (pinned host buffers, chunked H2D -> transform -> D2H): every chunk runs both
    passes on the pipeline's own stream with stream-ordered scratch."""
    n, nf = 1 << 16, 41
    x = O.uniform_c64(nf, n, seed=9)
    hx = torch.from_numpy(x.view(np.float32).reshape(nf, n, 2).copy()).pin_memory()
    hy = torch.zeros_like(hx).pin_memory()
    ms = sm.pipeline_host(hx, hy, n, nf, False, True, 0, 8)           # 8 transforms per chunk: six chunks, the last ragged
    assert ms > 0
    got = hy.numpy().view(np.complex64).reshape(nf, n)
    assert O.rel_l2(got, O.ct_c2c_fp64(x, False, True)) < TOL
    sm.pipeline_release()                                             # also returns the two-pass scratch pool and tables
    d = to_dev(x[:3])
    out = torch.zeros_like(d)
    sm.exec_c2c(d, out, n, 3, True, True)                             # ... which are rebuilt on the next call
    torch.cuda.synchronize()
    assert O.rel_l2(c64(out), O.ct_c2c_fp64(x[:3], True, True)) < TOL
