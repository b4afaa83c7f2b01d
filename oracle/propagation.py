"""ORACLE (test infrastructure, not the product): covariance propagation of transformed object coordinates.

Restates ``CoordinateTransformationExteriorOrientation.transform``
(/root/reference/JAICOV/src/org/applied_geodesy/adjustment/bundle/tranformation/CoordinateTransformationExteriorOrientation.java:49-121):
every object point seen in an image is carried into the frame of that image's reference image,

    X_trg = X0_trg + R_trg R_src' (X - X0_src)                     (:209-215),   identity for the reference image itself (:141-149)

and the covariance of all transformed coordinates is ``sigma2 * J Qxx J'`` (:110-114).  The reference writes the 45
entries of a Jacobian block out as scalar expressions (:223-279); here the same derivatives are stated in matrix form
(R = Rx(omega) Ry(phi) Rz(kappa), :172-184).  Pin: the reference ships no known answers for this function; golden vectors
were produced by evaluating the reference's own 45 scalar expressions and its transformation formula for seeded random
inputs (tests/golden/make_formula_fixtures.py -> tests/golden/reference_formulas.npz) and this module matches them to
1e-12 (tests/test_reference_formulas.py); in addition the Jacobian is checked against central differences
(tests/test_propagation.py).  The function as a whole -- visibility loops, order and names of the transformed points, column
placement, the final product sigma2 * J Qxx J' (:110-114) in packed upper layout -- is pinned by EXECUTING the reference's transform()
and setPartialDerivations() (tests/golden/make_transform_fixture.py -> reference_transform.npz; tests/test_reference_transform.py:
coordinates 1e-13, covariance 1e-12 correlation-scaled).
"""
from __future__ import annotations

import numpy as np

COL_FIXED = 2147483647


def rotation(om, ph, ka):
    """R and dR/d(omega, phi, kappa) for R = Rx(omega) Ry(phi) Rz(kappa) (:172-184)."""
    so, co, sp, cp, sk, ck = np.sin(om), np.cos(om), np.sin(ph), np.cos(ph), np.sin(ka), np.cos(ka)
    Rx = np.array([[1, 0, 0], [0, co, -so], [0, so, co]]); dRx = np.array([[0, 0, 0], [0, -so, -co], [0, co, -so]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]]); dRy = np.array([[-sp, 0, cp], [0, 0, 0], [-cp, 0, -sp]])
    Rz = np.array([[ck, -sk, 0], [sk, ck, 0], [0, 0, 1]]); dRz = np.array([[-sk, -ck, 0], [ck, -sk, 0], [0, 0, 0]])
    return Rx @ Ry @ Rz, [dRx @ Ry @ Rz, Rx @ dRy @ Rz, Rx @ Ry @ dRz]


def transform_point(X, eo_trg, eo_src):
    """:209-215."""
    RT, _ = rotation(*eo_trg[3:6])
    RS, _ = rotation(*eo_src[3:6])
    return eo_trg[:3] + RT @ (RS.T @ (X - eo_src[:3]))


def jacobian_block(X, eo_trg, eo_src):
    """3 x 15 block with respect to [X0_trg(3), angles_trg(3), X0_src(3), angles_src(3), X(3)] (:223-279)."""
    RT, dRT = rotation(*eo_trg[3:6])
    RS, dRS = rotation(*eo_src[3:6])
    d = X - eo_src[:3]
    q = RS.T @ d
    J = np.zeros((3, 15))
    J[:, 0:3] = np.eye(3)
    for a in range(3):
        J[:, 3 + a] = dRT[a] @ q
        J[:, 9 + a] = RT @ (dRS[a].T @ d)
    J[:, 6:9] = -RT @ RS.T
    J[:, 12:15] = RT @ RS.T
    return J


def propagate(xyz, pt_col, eo_val, eo_col, triples, sigma2, Qxx):
    """triples: iterable of (point, source image, target image); Qxx: dense n x n in reference column numbering.
    Returns (transformed coordinates (R, 3), covariance (3R, 3R))."""
    n = Qxx.shape[0]
    rows = []
    out = []
    J = np.zeros((3 * len(triples), n))
    for i, (p, s, t) in enumerate(triples):
        X = xyz[3 * p:3 * p + 3]
        pc = pt_col[3 * p:3 * p + 3]
        if s == t:
            blk, cols = np.zeros((3, 15)), np.full(15, -1)
            blk[:, 12:15] = np.eye(3)
            cols[12:15] = pc
            out.append(X.copy())
        else:
            eT, eS = eo_val[6 * t:6 * t + 6], eo_val[6 * s:6 * s + 6]
            blk = jacobian_block(X, eT, eS)
            cols = np.concatenate([eo_col[6 * t:6 * t + 6], eo_col[6 * s:6 * s + 6], pc])
            out.append(transform_point(X, eT, eS))
        for k, c in enumerate(cols):
            if 0 <= c < n and c != COL_FIXED:
                J[3 * i:3 * i + 3, c] += blk[:, k]
    return np.array(out), sigma2 * (J @ Qxx @ J.T)
