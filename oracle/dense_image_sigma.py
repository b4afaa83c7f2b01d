"""ORACLE EXTENSION (test infrastructure, not the product) -- PARITY UNPINNED beyond the block-diagonal case.

Per-image fully populated dispersion matrices of the image coordinates (BASELINE.json north_star (2), configs[3]) do not exist in
the reference: an ``ImageCoordinate`` observation group is always the two rows of one point (camera/ImageCoordinate.java:102-104)
with a 2 x 2 weight (PartialDerivativeFactory.java:296-319).  This subclass extends the oracle the way the reference's own
machinery would treat ONE observation group made of all 2m rows of an image: Jacobian rows from the same per-point evaluation
(jaicov_oracle.c: orc_eval_point, PDF:285-445), weight P = sigma0^2 Sigma^-1 like DirectlyObservedParameterGroup.getWeightMatrix
(DOPG:67-91), stacking N += A'PA, n += A'Pw as in stackNormalEquationSystem (PDF:475-505), Omega += v'Pv (BA:472-491).

Pin: with Sigma = blockdiag([[sx^2, rho sx sy], [., sy^2]]) the group decomposes into the reference's per-point groups, and
tests/test_dense_image_sigma.py requires this oracle to reproduce the FAITHFUL oracle (sigma / rho path) to rounding.
A scene's image carries the matrix as ``image['dispersion']`` (MTJ packed upper, rows x_0, y_0, x_1, y_1, ... in observation order).
"""
from __future__ import annotations

import ctypes

import numpy as np
from scipy.linalg import lapack

from .fast_oracle import FastOracle
from .oracle import _active, eval_point, lib


class DenseImageSigmaOracle(FastOracle):

    def __init__(self, scene, **kw):
        super().__init__(scene, **kw)
        self.dense = {}                       # image index -> dict(P=weight matrix)
        k = 0
        for cam in scene['cameras']:
            for im in cam['images']:
                if im.get('dispersion') is not None:
                    sg = np.asarray(im['dispersion'], float)
                    m2 = 2 * len(im['obj'])
                    assert sg.size == m2 * (m2 + 1) // 2
                    self.dense[k] = {'sigma': sg, 'P': None}
                    idx = np.arange(m2)
                    self.sigma2apriori = min(self.sigma2apriori, float(sg[idx + idx * (idx + 1) // 2].min()))   # DOPG:55-57
                k += 1

    def _weight(self, img):
        e = self.dense[img]
        if e['P'] is None:
            m2 = 2 * int(self.fp.pt_ptr[img + 1] - self.fp.pt_ptr[img])
            S = np.empty((m2, m2))
            ap = np.ascontiguousarray(e['sigma'] * (1.0 / self.sigma2apriori))
            lib().orc_unpack(m2, ap.ctypes.data, S.ctypes.data)
            c, info = lapack.dpotrf(S, lower=1)
            assert info == 0
            Pl, info = lapack.dpotri(c, lower=1)
            Pl = np.asarray(Pl)
            e['P'] = np.tril(Pl) + np.tril(Pl, -1).T
        return e['P']

    def _rows(self, img):
        """Dense Jacobian rows (2m x n) and misclosures of the image at the current values."""
        fp = self.fp
        o0, o1 = int(fp.pt_ptr[img]), int(fp.pt_ptr[img + 1])
        A = np.zeros((2 * (o1 - o0), fp.n))
        w = np.zeros(2 * (o1 - o0))
        for q, j in enumerate(range(o0, o1)):
            cols, a0, a1, wo, _P3 = eval_point(fp, img, j, self.sigma2apriori)
            for c, v0, v1 in zip(cols.tolist(), a0.tolist(), a1.tolist()):
                if _active(c):
                    A[2 * q, c] += v0
                    A[2 * q + 1, c] += v1
            w[2 * q:2 * q + 2] = wo
        return A, w

    def _ranges(self):
        """Observation ranges of the images WITHOUT a dispersion matrix (stacked by the C routine as before)."""
        fp = self.fp
        out = []
        for img in range(fp.nImg):
            if img not in self.dense and fp.pt_ptr[img + 1] > fp.pt_ptr[img]:
                out.append((int(fp.pt_ptr[img]), int(fp.pt_ptr[img + 1])))
        return out

    def create_normal_equation(self):
        fp = self.fp
        if not self.dense:
            return super().create_normal_equation()
        # image points of the ordinary images through the C routine (it takes an observation range), the dense images in numpy
        n = fp.n
        L = lib()
        p = fp.cstruct()
        m_all = fp.m
        # stack the ordinary ranges into scratch arrays first, then let the base class add bars / groups / datum / damping / V
        N0 = np.zeros(n * (n + 1) // 2)
        n0 = np.zeros(n)
        for (a, b) in self._ranges():
            L.orc_stack_image_points(ctypes.byref(p), self.sigma2apriori, N0.ctypes.data, n0.ctypes.data, a, b)
        for img in self.dense:
            A, w = self._rows(img)
            P = self._weight(img)
            PA = P @ A
            N0 += _pack(A.T @ PA)
            n0 += A.T @ (P @ w)
        # base class: everything else, with the image points switched off
        fp.m = 0
        saved = fp.pt_ptr
        fp.pt_ptr = np.zeros_like(saved)
        try:
            N, nv, _V = super().create_normal_equation()
        finally:
            fp.m = m_all
            fp.pt_ptr = saved
        # the base class applied the damping and computed V on a system without image points: redo both on the complete one
        N = N + 0.0
        d = fp.d
        if self.adapted_damping > 0:
            c = np.arange(d, n, dtype=np.int64)
            dg = c + c * (c + 1) // 2
            N0[dg] = N0[dg] + self.adapted_damping * N0[dg]
        N += N0
        nv += n0 if not self.simulation else 0.0
        V = np.empty(n)
        L.orc_preconditioner(n, N.ctypes.data, V.ctypes.data, 2.0 ** -53)
        return N, nv, V

    def get_omega(self, dx):
        fp = self.fp
        if not self.dense:
            return super().get_omega(dx)
        om = 0.0
        # ordinary image points: the C routine sums over ALL observations, so evaluate the dense images' per-point part and remove
        # it is not possible (their sigma / rho are ignored) -- sum range by range in Python instead
        dx = np.ascontiguousarray(dx)
        for img in range(fp.nImg):
            o0, o1 = int(fp.pt_ptr[img]), int(fp.pt_ptr[img + 1])
            if o1 == o0:
                continue
            if img in self.dense:
                A, w = self._rows(img)
                v = w - A @ dx
                om += float(v @ (self._weight(img) @ v))
            else:
                for j in range(o0, o1):
                    cols, a0, a1, wo, P3 = eval_point(fp, img, j, self.sigma2apriori)
                    act = np.array([_active(int(c)) for c in cols])
                    v0 = wo[0] - float(a0[act] @ dx[cols[act]])
                    v1 = wo[1] - float(a1[act] @ dx[cols[act]])
                    om += v0 * (P3[0] * v0 + P3[1] * v1) + v1 * (P3[1] * v0 + P3[2] * v1)
        p = fp.cstruct()
        om += lib().orc_omega_scale_bars(ctypes.byref(p), self.sigma2apriori, dx.ctypes.data)
        for g in fp.groups:
            om += self._omega_group(g, dx)
        return om


def _pack(S):
    n = S.shape[0]
    out = np.empty(n * (n + 1) // 2)
    lib().orc_pack(n, np.ascontiguousarray(S).ctypes.data, out.ctypes.data)
    return out
