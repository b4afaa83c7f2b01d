#!/usr/bin/env python
"""Stand-alone checks of the int8-digit tile products (csrc/ozaki.cu, JAICOV_GEMM_OZAKI) on a B200 -- NOT part of tests/ (the parity
suite covers the path through the adjustment); this is the kernel-level check it had to pass before it became the default
(profiles/r02_ozaki_check.log).

  python tools/ozaki_gpu_check.py            # all stages, every worker in its own process under a timeout
  python tools/ozaki_gpu_check.py gemm|spd|time

Stages
  gemm  jaicov_gemm_tiles on small tile grids, every operand layout / triangular hint / symmetric output, against numpy;
        first with the FP64 tile kernel (checks the entry point itself), then with 8 digits through the tcgen05 path.
        Everything the kernels must never read is NaN.
  spd   jaicov_spd_solve_invert (blocked Cholesky + inverse) with every launch on the digit path, against numpy.
  time  one 8192^3 product and one 16384^2 x 8192 symmetric product: FP64 DMMA kernel vs 8 / 7 digits and the cluster variant (TFLOP/s FP64-equivalent).
The environment variables are read once per process, so every setting runs in a subprocess (JAICOV_OZAKI_MIN_TILES=1 sends
small launches, JAICOV_OZAKI_MIN_K=128 short contractions through the digit path as well).  A hang is cut by the timeout and reported, not retried.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def masked_operands(rng, al, bl, mt, nt, K, kmode):
    """op(A) (128 mt x K) and op(B) (128 nt x K) as the math sees them (zeros where a triangular operand is zero), and the
    stored arrays with NaN wherever the tile kernel must not read."""
    Mr, Nr = 128 * mt, 128 * nt
    A = rng.standard_normal((Mr, K)) * np.exp(rng.uniform(-6, 6, (Mr, 1)))      # rows of very different magnitude
    B = rng.standard_normal((Nr, K)) * np.exp(rng.uniform(-6, 6, (Nr, 1)))
    m, n, k = np.arange(Mr)[:, None], np.arange(Nr)[:, None], np.arange(K)[None, :]
    nanA, nanB = np.zeros_like(A, bool), np.zeros_like(B, bool)
    if kmode == 1:                                   # op(B)'[k][n] lower triangular: k >= n
        B[k < n] = 0.0
        nanB = k < 128 * (n // 128)
    elif kmode == 2:                                 # op(A)[m][k] lower triangular: k <= m
        A[k > m] = 0.0
        nanA = k >= 128 * (m // 128 + 1)
    elif kmode == 3:                                 # both stored [k][.] lower triangular: k >= m, k >= n
        A[k < m] = 0.0
        B[k < n] = 0.0
        nanA = k < 128 * (m // 128)
        nanB = k < 128 * (n // 128)
    As, Bs = A.copy(), B.copy()
    As[nanA] = np.nan
    Bs[nanB] = np.nan
    return A, B, (As if al == 0 else As.T.copy()), (Bs if bl == 0 else Bs.T.copy())


def run_gemm_cases(gemm, tol=1e-13):
    """All operand layouts / triangular hints / symmetric outputs through `gemm(As, Bs, C0, al, bl, alpha, beta, tri, kmode)
    -> C`; returns the list of case records.  Shared by the GPU worker below and by the CPU test that drives the host emulation."""
    rng = np.random.default_rng(11)
    cases = []
    for al in (0, 1):
        for bl in (0, 1):
            for kmode, tri in ((0, 0), (0, 1), (1, 0), (2, 0), (3, 1)):
                if kmode == 3 and not (al == 1 and bl == 1):
                    continue                          # K_MAX_IJ: both operands stored [k][.]
                if kmode == 1 and bl != 1:
                    continue                          # K_B_LOWER: B stored [k][n]
                if kmode == 2 and al != 0:
                    continue                          # K_A_LOWER: A stored [m][k]
                mt = nt = 3
                K = 384 if kmode else 512
                if kmode == 0 and not tri:
                    mt, nt = 3, 2
                A, B, As, Bs = masked_operands(rng, al, bl, mt, nt, K, kmode)
                C0 = rng.standard_normal((128 * mt, 128 * nt))
                for alpha, beta in ((1.0, 0.0), (-1.0, 1.0)):
                    C = gemm(As, Bs, C0, al, bl, alpha, beta, tri, kmode)
                    ref = alpha * (A @ B.T) + beta * C0
                    scale = np.abs(A) @ np.abs(B.T) + np.abs(C0)
                    it, jt = np.arange(128 * mt)[:, None] // 128, np.arange(128 * nt)[None, :] // 128
                    computed = (it >= jt) if tri else np.ones_like(ref, bool)
                    with np.errstate(invalid='ignore'):
                        err = float(np.nanmax(np.abs(C - ref)[computed] / scale[computed]))
                    finite = bool(np.isfinite(C).all())
                    untouched = bool(np.array_equal(C[~computed], C0[~computed]))
                    cases.append(dict(al=al, bl=bl, kmode=kmode, tri=tri, alpha=alpha, beta=beta, err=err, untouched=untouched,
                                      finite=finite, ok=finite and untouched and err < tol))
    return cases


def worker_gemm():
    import bundle_adjustment_b200 as ba
    cases = run_gemm_cases(lambda As, Bs, C0, al, bl, alpha, beta, tri, kmode: ba._lib.gemm_tiles(As, Bs, C0, al, bl, alpha, beta, tri, kmode)[0])
    print(json.dumps(dict(stage='gemm', ok=all(c['ok'] for c in cases), worst=max(c['err'] for c in cases), cases=len(cases),
                          failing=[c for c in cases if not c['ok']][:6])))


def worker_spd():
    import bundle_adjustment_b200 as ba
    rng = np.random.default_rng(12)
    out = []
    for n in (700, 3000):
        G = rng.standard_normal((n, n))
        S = G @ G.T + 0.05 * n * np.eye(n)
        d = 1 / np.sqrt(np.diag(S))
        S = S * d[:, None] * d[None, :]
        b = rng.standard_normal((2, n))
        Q, x, ms = ba.spd_solve_invert(S, b)
        Qr = np.linalg.inv(S)
        sc = np.sqrt(np.diag(Qr))
        out.append(dict(n=n, qxx=float(np.max(np.abs(Q - Qr) / np.outer(sc, sc))), sol=float(np.max(np.abs(x - np.linalg.solve(S, b.T).T))),
                        ms_factor=ms[0], ms_inverse=ms[1]))
    print(json.dumps(dict(stage='spd', ok=all(o['qxx'] < 1e-10 for o in out), runs=out)))


def worker_time():
    import bundle_adjustment_b200 as ba
    rng = np.random.default_rng(13)
    out = []
    for name, mt, nt, K, tri in (('8192^3', 64, 64, 8192, 0), ('syrk 16384^2 x 8192', 128, 128, 8192, 1)):
        A = rng.standard_normal((128 * mt, K))
        B = A if tri else rng.standard_normal((128 * nt, K))
        C = np.zeros((128 * mt, 128 * nt))
        _, ms = ba._lib.gemm_tiles(A, B, C, 0, 0, 1.0, 0.0, tri, 0, reps=3)
        flop = 2.0 * (128 * mt) * (128 * nt) * K * (0.5 * (1 + 1 / mt) if tri else 1.0)
        out.append(dict(case=name, ms=ms, tflops=flop / (ms * 1e-3) / 1e12))
    # the dominant launch of the dense route: M = W'W with W lower triangular, lower output tiles, k >= max(i, j) (LAUUM)
    nt_ = 96
    n = 128 * nt_
    W = np.tril(rng.standard_normal((n, n)))
    C = np.zeros((n, n))
    _, ms = ba._lib.gemm_tiles(W, W, C, 1, 1, 1.0, 0.0, 1, 3, reps=2)
    flop = sum(2.0 * 128 * 128 * (n - 128 * max(i, j)) for i in range(nt_) for j in range(i + 1))
    out.append(dict(case='lauum %d' % n, ms=ms, tflops=flop / (ms * 1e-3) / 1e12))
    print(json.dumps(dict(stage='time', ok=True, runs=out)))


def run_worker(stage, env_extra, timeout):
    env = dict(os.environ)
    for k in ('JAICOV_GEMM_OZAKI', 'JAICOV_OZAKI_MIN_TILES', 'JAICOV_OZAKI_MIN_K', 'JAICOV_OZAKI_CLUSTER'):
        env.pop(k, None)
    env.update(env_extra)
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--worker', stage], env=env, capture_output=True, text=True, timeout=timeout)
        line = [l for l in r.stdout.splitlines() if l.startswith('{')]
        res = json.loads(line[-1]) if line else dict(stage=stage, ok=False, rc=r.returncode, stderr=r.stderr[-600:])
    except subprocess.TimeoutExpired:
        res = dict(stage=stage, ok=False, timeout=timeout)
    res['env'] = env_extra
    res['wall_s'] = round(time.time() - t0, 1)
    print(json.dumps(res), flush=True)
    return res.get('ok', False)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == '--worker':
        {'gemm': worker_gemm, 'spd': worker_spd, 'time': worker_time}[sys.argv[2]]()
        return
    stages = sys.argv[1:] or ['gemm', 'spd', 'time']
    oz = {'JAICOV_GEMM_OZAKI': '8', 'JAICOV_OZAKI_MIN_TILES': '1', 'JAICOV_OZAKI_MIN_K': '128'}
    if 'gemm' in stages:
        if not run_worker('gemm', {}, 300):                       # the entry point itself, FP64 tile kernel
            print('FP64 route failed: fix the harness / entry point first')
            return
        if not run_worker('gemm', oz, 120):
            print('digit path failed on small tile grids: stop here')
            return
        run_worker('gemm', dict(oz, JAICOV_OZAKI_CLUSTER='2'), 120)   # cluster of two CTAs sharing op(A) by TMA multicast
    if 'spd' in stages:
        run_worker('spd', {}, 300)
        if not run_worker('spd', oz, 300):
            return
    if 'time' in stages:
        run_worker('time', {}, 600)
        for s in ('8', '7'):
            run_worker('time', {'JAICOV_GEMM_OZAKI': s}, 600)
        run_worker('time', {'JAICOV_GEMM_OZAKI': '8', 'JAICOV_OZAKI_CLUSTER': '2'}, 600)


if __name__ == '__main__':
    main()
