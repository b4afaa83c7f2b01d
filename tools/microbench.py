"""GPU microbenchmarks used to set the FP64 roofline denominator and to time the dense kernels in isolation.
Run on the GPU box: python tools/microbench.py [sizes...]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import bundle_adjustment_b200 as ba


def dgemm_peak(n=8192, reps=6):
    a = torch.randn(n, n, dtype=torch.float64, device='cuda')
    b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def main():
    out = {'gpu': torch.cuda.get_device_name(0)}
    for n in (4096, 8192):
        out['torch_dgemm_tflops_%d' % n] = dgemm_peak(n)
    print(json.dumps(out), flush=True)
    sizes = [int(x) for x in sys.argv[1:]] or [2048, 4096, 8192, 16384]
    rng = np.random.default_rng(0)
    for n in sizes:
        A = rng.standard_normal((n, 64))
        S = A @ A.T / 64 + np.eye(n) * 2.0
        b = rng.standard_normal((1, n))
        t = time.time()
        Q, x, ms = ba.spd_solve_invert(S, b, invert=True)
        wall = time.time() - t
        r = np.abs(S @ x[0] - b[0]).max()
        rq = np.abs(Q[:, :8].T @ S - np.eye(n)[:8]).max()
        print(json.dumps({'n': n, 'ms_factor': ms[0], 'ms_inverse': ms[1], 'factor_tflops': n ** 3 / 3 / ms[0] / 1e9,
                          'inverse_tflops': 2 * n ** 3 / 3 / ms[1] / 1e9, 'resid_solve': r, 'resid_inv': rq, 'wall_s': wall}), flush=True)


if __name__ == '__main__':
    main()
