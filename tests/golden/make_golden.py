"""Generates tests/golden/ref_*.npz ON A B200: outputs of the reference's own kernels (oracle/_ref:
unmodified /root/reference sources rebuilt for sm_100a by oracle/Makefile) on seeded inputs.

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy into tests/golden/

The fixtures pin oracle/smfft_oracle.c (tests/test_oracle.py) and the CUDA path (tests/test_gpu_parity.py)
against the real reference; the reference itself ships no fixtures.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle_np as O  # noqa: E402
from tests import refkernels as R  # noqa: E402


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    dev = "cuda"
    ct = {}
    for n in (32, 64, 128, 256, 512, 1024, 2048, 4096):
        nf = 4
        x = O.uniform_c64(nf, n, seed=O.SEED + n)
        dx = torch.from_numpy(x.view(np.float32).reshape(nf, n, 2)).to(dev)
        for inverse in (0, 1):
            for reorder in (1, 0):
                dy = torch.zeros_like(dx)
                R.ct_external(dx, dy, n, nf, inverse, reorder)
                torch.cuda.synchronize()
                y = dy.cpu().numpy().view(np.complex64).reshape(nf, n)
                np.savez(os.path.join(outdir, f"ref_ct_{n}_{'inv' if inverse else 'fwd'}_{'reorder' if reorder else 'noreorder'}.npz"),
                         kind="ct", input=x, output=y, inverse=inverse, reorder=reorder)
    for n in (256, 512, 1024, 2048, 4096):
        nf = 3
        x = O.uniform_c64(nf, n, seed=O.SEED + 7 * n)
        dx = torch.from_numpy(x.view(np.float32).reshape(nf, n, 2)).to(dev)
        dy = torch.zeros_like(dx)
        R.st_external(dx, dy, n, nf)
        torch.cuda.synchronize()
        np.savez(os.path.join(outdir, f"ref_stockham_{n}_inv.npz"), kind="stockham", input=x,
                 output=dy.cpu().numpy().view(np.complex64).reshape(nf, n))
    for n in (512, 1024, 2048, 4096):
        nf = 3
        x = O.uniform_f32(nf, n, seed=O.SEED + 11 * n)
        dx = torch.from_numpy(x).to(dev)
        dy = torch.zeros((nf, n // 2, 2), dtype=torch.float32, device=dev)
        R.rc_external(dx, dy, n, nf, 0)
        torch.cuda.synchronize()
        y = dy.cpu().numpy().view(np.complex64).reshape(nf, n // 2)
        np.savez(os.path.join(outdir, f"ref_r2c_{n}.npz"), kind="r2c", input=x, output=y)
        # C2R on a random half-spectrum with real DC/Nyquist packed in bin 0 (RC/FFT.c:264-283)
        h = O.uniform_c64(nf, n // 2, seed=O.SEED + 13 * n)
        dh = torch.from_numpy(h.view(np.float32).reshape(nf, n // 2, 2)).to(dev)
        dz = torch.zeros((nf, n), dtype=torch.float32, device=dev)
        R.rc_external(dh, dz, n, nf, 1)
        torch.cuda.synchronize()
        np.savez(os.path.join(outdir, f"ref_c2r_{n}.npz"), kind="c2r", input=h, output=dz.cpu().numpy())
    print("golden fixtures written to", outdir, len(os.listdir(outdir)), "files")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__))))
