#!/usr/bin/env python
"""Numerics study (CPU, test infrastructure -- it uses the oracle, hence its place under tests/): the blocked Cholesky + inverse schedule of csrc/dense_driver.hpp with every
GEMM computed by the int8-slice emulation of tests/emul/host_backend.cpp (what FP64 products on the INT8 tcgen05 tensor
cores would compute, bit for bit) against the same schedule in FP64, on the preconditioned SPD system M~ = V N V + B~'B~
 of real bundle networks assembled by the oracle.

  python tests/ozaki_study.py [--images 12 --targets 150] [--digits 6 7 8]

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


def structured_study(L, scene, digits):
    """The two big products of the structured route, T = Q'Y' and K^-1[p, p] = P^-1 + Y T, from int8 digit products:
    the bordered preconditioned system K = [[P, Z], [Z', Rb]] (object coordinates first, P block diagonal) is assembled by the
    oracle, the small reduced system is inverted in FP64, and the point block of the inverse is compared with a long-double
    reference -- FP64 products vs emulated digit products."""
    from oracle.oracle import Oracle, _unpack_upper
    o = Oracle(scene)
    if o.use_centroid:
        o._centroid(False)
    N, nv, V = o.create_normal_equation()
    n, d = o.fp.n, o.fp.d
    Kb = _unpack_upper(N, n) * V[:, None] * V[None, :]                 # bordered, Jacobi-scaled (V = 1 on the border)
    pc = np.asarray(o.fp.pt_col).reshape(-1)
    pcols = np.sort(pc[(pc >= 0) & (pc < 2147483647)])
    up = pcols.size
    assert np.array_equal(pcols, np.arange(d, d + up)), 'object coordinates are the leading unknown columns'
    rest = np.concatenate([np.arange(d + up, n), np.arange(0, d)])     # cameras / images, then the datum multipliers
    P, Z, Rb = Kb[np.ix_(pcols, pcols)], Kb[np.ix_(pcols, rest)], Kb[np.ix_(rest, rest)]
    Pinv = np.linalg.inv(P)                                            # block diagonal 3 x 3 (inverted block-wise on the device)
    Y = Pinv @ Z
    Qp = np.linalg.inv(Rb - Z.T @ Y)                                   # Q' = K'^-1, m x m, small
    Kl = Kb.astype(np.longdouble)
    X = np.linalg.inv(Kb).astype(np.longdouble)
    for _ in range(2):
        X = X + X @ (np.eye(n, dtype=np.longdouble) - Kl @ X)
    ref = X[np.ix_(pcols, pcols)]
    sc = np.sqrt(np.abs(np.diag(ref))).astype(np.float64)
    L.emul_gemm_tiles.argtypes = [ctypes.c_int] * 4 + [ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_int64,
                                  ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    pad = lambda a, r, c: np.pad(a, ((0, r - a.shape[0]), (0, c - a.shape[1])))
    Tp, mp = (up + 127) // 128 * 128, (rest.size + 127) // 128 * 128
    Yt = np.ascontiguousarray(pad(Y.T, mp, Tp))                        # stored transposed like the device buffers
    Qpp = np.ascontiguousarray(pad(Qp, mp, mp))
    out = {}
    for s in [0] + list(digits):
        T1 = np.zeros((mp, Tp))                                        # T = Q' Y'   (A = Q'[m][k], B = Yt[k][n])
        L.emul_gemm_tiles(0, 1, mp // 128, Tp // 128, mp, 1.0, 0.0, Qpp.ctypes.data, mp, Yt.ctypes.data, Tp, T1.ctypes.data, Tp, 0, 0, s)
        C = np.zeros((Tp, Tp))                                         # Y T       (A = Yt[k][m], B = T1[k][n]), lower tiles
        L.emul_gemm_tiles(1, 1, Tp // 128, Tp // 128, mp, 1.0, 0.0, Yt.ctypes.data, Tp, T1.ctypes.data, Tp, C.ctypes.data, Tp, 1, 0, s)
        Q = np.tril(C[:up, :up]) + np.tril(C[:up, :up], -1).T + Pinv
        out[s] = float(np.max(np.abs(Q.astype(np.longdouble) - ref) / np.outer(sc, sc)))
    return out, up, rest.size


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', type=int, default=2)
    ap.add_argument('--images', type=int, default=12)
    ap.add_argument('--targets', type=int, default=150)
    ap.add_argument('--digits', type=int, nargs='*', default=[6, 7, 8])
    ap.add_argument('--min-tiles', type=int, default=1)
    ap.add_argument('--structured', action='store_true', help='the structured route\'s two big products instead of the dense schedule')
    args = ap.parse_args()
    from tests.scenes import synthetic_scene
    scene, _ = synthetic_scene(args.config, images=args.images, targets=args.targets)
    if args.structured:
        dev, up, m = structured_study(load_emul(), scene, args.digits)
        print('config %d, %d images x %d targets: %d object-coordinate columns, reduced system %d' % (args.config, args.images, args.targets, up, m))
        for s, e in dev.items():
            print('%-22s point block of Qxx, deviation (correlation-scaled) %.2e' % ('FP64 products' if s == 0 else 'int8 slices, s = %d' % s, e))
        return
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
