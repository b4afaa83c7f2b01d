#!/usr/bin/env python
"""Numerics study (CPU, test infrastructure): the blocked Cholesky + inverse schedule of csrc/dense_driver.hpp with every
GEMM computed by the int8-slice emulation of tests/emul/host_backend.cpp (what FP64 products on the INT8 tcgen05 tensor
cores would compute, bit for bit) against the same schedule in FP64, on the preconditioned SPD system M~ = V N V + B~'B~
 of real bundle networks assembled by the oracle.

  python tools/ozaki_study.py [--images 12 --targets 150] [--digits 6 7 8]

Prints, per digit count: largest integer group sum (must stay below 2^31), the correlation-scaled deviation of the
inverse from a long-double reference, next to the deviation of the FP64 schedule itself.
"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_emul():
    L = ctypes.CDLL(os.path.join(ROOT, 'tests', '_build', 'libemul.so'))
    L.emul_spd_solve_invert_ex.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    return L


def scaled_system(scene):
    """M~ = V N V + (B V)'(B V) of the scene's first pass (SPD, unit diagonal up to the datum term), u x u."""
    from oracle.oracle import Oracle, _unpack_upper
    o = Oracle(scene)
    if o.use_centroid:
        o._centroid(False)
    N, nv, V = o.create_normal_equation()
    n, d = o.fp.n, o.fp.d
    K = _unpack_upper(N, n)
    Nuu = K[d:, d:]
    B = K[:d, d:]
    v = V[d:]
    M = Nuu * v[:, None] * v[None, :]
    Bt = B * v[None, :]
    return M + Bt.T @ Bt, nv[d:] * v


def reference_inverse(S):
    """Inverse in extended precision: FP64 inverse refined by two Newton steps X <- X + X (I - S X) in long double."""
    X = np.linalg.inv(S).astype(np.longdouble)
    Sl = S.astype(np.longdouble)
    for _ in range(2):
        R = np.eye(S.shape[0], dtype=np.longdouble) - Sl @ X
        X = X + X @ R
    return X


def run(L, S, rhs, digits, min_tiles=1):
    u = S.shape[0]
    np_ = (u + 127) // 128 * 128
    M = np.eye(np_)
    M[:u, :u] = S
    M = np.tril(M).copy()
    R = np.zeros((128, np_))
    R[0, :u] = rhs
    st = np.zeros(5)
    t0 = time.perf_counter()
    info = L.emul_spd_solve_invert_ex(np_, M.ctypes.data, 1, R.ctypes.data, 1, st.ctypes.data, digits, min_tiles)
    dt = time.perf_counter() - t0
    Q = np.tril(M[:u, :u])
    Q = Q + np.tril(Q, -1).T
    return info, Q, R[0, :u].copy(), st, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', type=int, default=2)
    ap.add_argument('--images', type=int, default=12)
    ap.add_argument('--targets', type=int, default=150)
    ap.add_argument('--digits', type=int, nargs='*', default=[6, 7, 8])
    ap.add_argument('--min-tiles', type=int, default=1)
    args = ap.parse_args()
    from tests.scenes import synthetic_scene
    scene, _ = synthetic_scene(args.config, images=args.images, targets=args.targets)
    S, rhs = scaled_system(scene)
    u = S.shape[0]
    cond = np.linalg.cond(S)
    print('config %d, %d images x %d targets: u = %d, cond(M~) = %.3g' % (args.config, args.images, args.targets, u, cond))
    Xref = reference_inverse(S)
    sc = np.sqrt(np.abs(np.diag(Xref))).astype(np.float64)
    yref = (Xref @ rhs.astype(np.longdouble)).astype(np.float64)
    L = load_emul()

    def report(name, info, Q, y, st, dt):
        eq = float(np.max(np.abs(Q.astype(np.longdouble) - Xref) / np.outer(sc, sc)))
        ey = float(np.max(np.abs(y - yref)) / np.max(np.abs(yref)))
        extra = '' if st[3] == 0 else ', %d launches emulated, max |group sum| = 2^%.1f' % (st[3], np.log2(max(st[4], 1)))
        print('%-22s info %d  Qxx dev (correlation-scaled) %.2e  solution dev %.2e  [%.1f s%s]' % (name, info, eq, ey, dt, extra))
        return eq

    e64 = report('FP64 schedule', *run(L, S, rhs, 0))
    Ql = np.linalg.inv(S)
    print('%-22s         Qxx dev (correlation-scaled) %.2e' % ('LAPACK inv (numpy)', float(np.max(np.abs(Ql - Xref) / np.outer(sc, sc)))))
    for s in args.digits:
        report('int8 slices, s = %d' % s, *run(L, S, rhs, s, args.min_tiles))
    print('parity bar of tests/test_gpu_parity.py: Qxx 1e-8 correlation-scaled, parameters 1e-10')


if __name__ == '__main__':
    main()
