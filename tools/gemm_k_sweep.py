#!/usr/bin/env python
"""Tile-product rate against the contraction length K (the distributed Cholesky updates with K = panel width): FP64 DMMA tiles vs
int8 digit products.  python tools/gemm_k_sweep.py  (each arithmetic in its own process: the switch is read once)"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker():
    import bundle_adjustment_b200 as ba
    rng = np.random.default_rng(5)
    out = []
    for K in (4096, 2048, 1024, 512, 1024, 2048, 4096):
        mt = nt = 96
        A = rng.standard_normal((128 * mt, K))
        C = np.zeros((128 * mt, 128 * nt))
        for tri, beta in ((1, 0.0),):
            ba._lib.gemm_tiles(A, A, C, 0, 0, -1.0, beta, tri, 0, reps=1)      # warm-up: scratch allocation, module load
            _, ms = ba._lib.gemm_tiles(A, A, C, 0, 0, -1.0, beta, tri, 0, reps=4)
            flop = 2.0 * (128 * mt) ** 2 * K * 0.5 * (1 + 1 / mt)
            out.append(dict(K=K, tiles=mt * (mt + 1) // 2, ms=ms, tflops=flop / (ms * 1e-3) / 1e12))
    print(json.dumps(out))


if __name__ == '__main__':
    if len(sys.argv) > 1:
        worker()
    else:
        for d in ('0', '8'):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), 'w'], env=dict(os.environ, JAICOV_GEMM_OZAKI=d, JAICOV_OZAKI_MIN_K='128'),
                               capture_output=True, text=True, timeout=900)
            print('digits', d, [l for l in r.stdout.splitlines() if l.startswith('[')][-1] if r.returncode == 0 else r.stderr[-500:])
