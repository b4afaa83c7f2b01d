"""ORACLE (test infrastructure, not the product): the "fast oracle" of SURVEY.md section 7.1 step 2.

The faithful oracle (oracle.py) calls the routines the reference calls -- packed, unblocked dspsv + dsptri, one thread
(MathExtension.java:338-366) -- and cannot finish BASELINE.json's configs[2] (r = 6 000 dense dispersion: dpptrf/dpptri plus a
36 M-term Python loop per pass) or configs[3] (n = 16 220: 4.3e12 flop of level-2 BLAS) inside a test.  This subclass keeps
EVERYTHING of the faithful oracle -- per-observation evaluation and stacking in jaicov_oracle.c, datum rows, preconditioner,
the estimateModel loop, updates, Omega -- and replaces only the two dense solves by blocked LAPACK on all host cores:

* ``MathExtension.solve`` on the bordered indefinite system K = [[0, B], [B', N]] (preconditioned, V = 1 on the border)
  -> Cholesky route of SURVEY.md section 7.3 item 1:  M = N + B'B (SPD),  z = M^-1 n,  G = M^-1 B',  S = B G,
  lambda = S^-1 B z,  y = z - G lambda,  K^-1 = [[I - S^-1, (G S^-1)'], [G S^-1, M^-1 - G S^-1 G']]  with dpotrf / dpotrs / dpotri.
  With d = 0 this is plain Cholesky.  The results are those of dspsv + dsptri up to rounding (both are backward stable;
  forward error ~ eps * cond(VKV));
* ``DirectlyObservedParameterGroup.getWeightMatrix`` (DOPG:67-91, dpptrf + dpptri) -> dpotrf + dpotri on the full matrix, and the
  right-hand-side accumulation of the group as one matrix-vector product instead of the sequential loop.

Pinned by tests/test_fast_oracle.py against the faithful oracle on configs[0] (bundled example) and configs[1] and on networks
with scale bars / observed groups / fixed datum (parameters 1e-12, sigma0^2 1e-11, Qxx 1e-11 correlation-scaled, identical pass
counts).  It is NOT the reference's algorithm, only its result: the CPU baseline of bench.py keeps using the faithful one.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import lapack

from . import lapack_packed as lp
from .oracle import Oracle, _active, lib, pidx


class FastOracle(Oracle):

    # ---- MathExtension.solve replacement (bordered system, blocked) ------------------------------------------------------
    def _solve_packed(self, ap, b, n, invert):
        d = self.fp.d
        L = lib()
        K = np.empty((n, n))
        L.orc_unpack(n, ap.ctypes.data, K.ctypes.data)
        B = np.ascontiguousarray(K[:d, d:])                     # d x u (preconditioned: V = 1 on the border rows)
        M = K[d:, d:]                                           # view; symmetric u x u
        if d:
            if np.any(K[:d, :d] != 0.0):
                raise ValueError('border block of the datum rows is expected to be zero')
            M = M + B.T @ B
        else:
            M = np.ascontiguousarray(M)
        del K
        c, info = lapack.dpotrf(M, lower=1, overwrite_a=1)
        if info > 0:
            raise lp.MatrixNotSPDException()
        if info < 0:
            raise ValueError('dpotrf illegal argument %d' % info)
        rhs = np.empty((n - d, 1 + d), order='F')
        rhs[:, 0] = b[d:]
        rhs[:, 1:] = B.T
        x, info = lapack.dpotrs(c, rhs, lower=1)
        z, G = x[:, 0], x[:, 1:]                                # G = M^-1 B' (u x d)
        if d:
            S = B @ G
            Sinv = np.linalg.inv(0.5 * (S + S.T))
            lam = Sinv @ (B @ z)
            y = z - G @ lam
        else:
            Sinv = np.zeros((0, 0))
            lam = np.zeros(0)
            y = z
        b[:d] = lam
        b[d:] = y
        if not invert:
            return
        Minv, info = lapack.dpotri(c, lower=1, overwrite_c=1)
        if info != 0:
            raise lp.MatrixNotSPDException()
        # dpotri returns the lower triangle (Fortran order): as a C array that is the UPPER triangle of the symmetric inverse
        Q = np.empty((n, n))
        Q[d:, d:] = Minv.T if Minv.flags.f_contiguous else Minv
        del Minv
        if d:
            T = G @ Sinv                                        # u x d
            Q[d:, d:] -= T @ G.T                                # rank-d datum correction (only the upper triangle is read below)
            Q[:d, :d] = np.eye(d) - Sinv
            Q[:d, d:] = T.T
        L.orc_pack(n, Q.ctypes.data, ap.ctypes.data)

    # ---- DirectlyObservedParameterGroup.getWeightMatrix, blocked -----------------------------------------------------------
    def _group_dense_weight(self, g):
        if g.get('Pfull') is None:
            r = len(g['refs'])
            S = np.empty((r, r))
            ap = np.ascontiguousarray(g['dispersion'] * (1.0 / self.sigma2apriori))
            lib().orc_unpack(r, ap.ctypes.data, S.ctypes.data)
            c, info = lapack.dpotrf(S, lower=1, overwrite_a=1)
            if info != 0:
                raise lp.MatrixNotSPDException()
            Pl, info = lapack.dpotri(c, lower=1, overwrite_c=1)
            Pl = np.asarray(Pl)
            P = np.tril(Pl) + np.tril(Pl, -1).T
            g['Pfull'] = P
        return g['Pfull']

    def _stack_group(self, g, N, n):
        if g['dispersion'] is None:
            return super()._stack_group(g, N, n)
        P = self._group_dense_weight(g)
        w, cols = self._group_w(g)
        act = np.array([_active(int(c)) for c in cols])
        ca = cols[act]
        n[ca] += (P[act] @ w)
        lo = np.minimum(ca[:, None], ca[None, :])
        hi = np.maximum(ca[:, None], ca[None, :])
        sel = ca[:, None] <= ca[None, :]
        np.add.at(N, (lo + hi * (hi + 1) // 2)[sel], P[np.ix_(act, act)][sel])

    def _omega_group(self, g, dx):
        if g['dispersion'] is None:
            return super()._omega_group(g, dx)
        P = self._group_dense_weight(g)
        w, cols = self._group_w(g)
        act = np.array([_active(int(c)) for c in cols])
        v = w.copy()
        v[act] -= dx[cols[act]]
        return float(v @ (P @ v))
