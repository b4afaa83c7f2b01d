"""ORACLE (test infrastructure, not the product): direct linear transformation, the initial values of one image.

Restates ``DirectLinearTransformation.adjust`` and its helpers
(/root/reference/JAICOV/src/org/applied_geodesy/adjustment/bundle/dlt/DirectLinearTransformation.java:67-352,
 .../dlt/DLTPartialDerivativeFactory.java:40-337) for one image:

* homologous points, coordinate scale ``sqrt(sum |X|^2 / sum |x|^2)``                      (DLT:75-106)
* linear model  x = X b11 + Y b12 + Z b13 + b14 - x (X b31 + Y b32 + Z b33), same for y    (DPF:62-65, :239-337)
* first pass without, later passes with the restrictions as border rows                    (DLT:116-165, DPF:68-236)
* Jacobi preconditioner V_ii = 1/sqrt(N_ii) if N_ii > EPS else 1, dspsv                     (DLT:341-347, :132-134)
* back-scaling and the interior / exterior orientation from the 11 coefficients            (DLT:186-266)

The gradients of the six restrictions are written here in vector form (r1, r2, r3 = rows of the 3 x 3 part of B);
they are the same functions the reference spells out entry by entry.  Pin: the reference ships no known answers for
the DLT; golden vectors were produced by evaluating the reference's own expressions for the six restriction rows and
misclosures (DPF:100-236) and for the expansion (DLT:208-246) on seeded random inputs
(tests/golden/make_formula_fixtures.py -> tests/golden/reference_formulas.npz) and this module matches them to 1e-12
(tests/test_reference_formulas.py); the complete adjust() -- loop logic, passes, return value, coefficients and derived
orientation -- matches DirectLinearTransformation.adjust executed end to end (tests/golden/make_dlt_fixture.py ->
reference_dlt.npz: six images, five restriction sets), and recovers a synthetic pin-hole camera (tests/test_dlt.py).
"""
from __future__ import annotations

import numpy as np

from .lapack_packed import solve_symm_packed

EPS = 2.0 ** -53
SQRT_EPS = np.sqrt(EPS)

# RestrictionType ordinals (DLT:50-57)
IDENTICAL_PRINCIPLE_DISTANCE, ROTATION_WITHOUT_SHEAR, FIXED_PRINCIPLE_DISTANCE_X, FIXED_PRINCIPLE_DISTANCE_Y, \
    FIXED_PRINCIPAL_POINT_X, FIXED_PRINCIPAL_POINT_Y = range(6)


def validate_restrictions(restrictions):
    """DLT:268-277: distinct, insertion-ordered; both fixed principal distances make the identical one redundant."""
    out = list(dict.fromkeys(int(r) for r in restrictions))
    if FIXED_PRINCIPLE_DISTANCE_X in out and FIXED_PRINCIPLE_DISTANCE_Y in out and IDENTICAL_PRINCIPLE_DISTANCE in out:
        out.remove(IDENTICAL_PRINCIPLE_DISTANCE)
    return out


def restriction_row(kind, b, c, x0, y0):
    """Gradient (11) and misclosure of one restriction at the coefficients b (DPF:68-236)."""
    r1, r2, r3 = b[0:3], b[4:7], b[8:11]
    b1, b2, b3 = r1 @ r1, r2 @ r2, r3 @ r3
    bx, by = r1 @ r3, r2 @ r3
    g = np.zeros(11)
    if kind == FIXED_PRINCIPAL_POINT_X:          # x0 = bx / b3
        g[0:3] = r3 / b3
        g[8:11] = r1 / b3 - 2.0 * bx * r3 / b3 ** 2
        w = x0 - bx / b3
    elif kind == FIXED_PRINCIPAL_POINT_Y:        # y0 = by / b3
        g[4:7] = r3 / b3
        g[8:11] = r2 / b3 - 2.0 * by * r3 / b3 ** 2
        w = y0 - by / b3
    elif kind == FIXED_PRINCIPLE_DISTANCE_X:     # c^2 = b1 / b3 - bx^2 / b3^2
        g[0:3] = 2.0 * (r1 * b3 - bx * r3) / b3 ** 2
        g[8:11] = 4.0 * (r3 * bx * bx - 0.5 * b3 * (r3 * b1 + bx * r1)) / b3 ** 3
        w = c * c - b1 / b3 + bx * bx / b3 ** 2
    elif kind == FIXED_PRINCIPLE_DISTANCE_Y:
        g[4:7] = 2.0 * (r2 * b3 - by * r3) / b3 ** 2
        g[8:11] = 4.0 * (r3 * by * by - 0.5 * b3 * (r3 * b2 + by * r2)) / b3 ** 3
        w = c * c - b2 / b3 + by * by / b3 ** 2
    elif kind == IDENTICAL_PRINCIPLE_DISTANCE:   # b3 (b1 - b2) - bx^2 + by^2 = 0
        g[0:3] = 2.0 * (b3 * r1 - bx * r3)
        g[4:7] = -2.0 * (b3 * r2 - by * r3)
        g[8:11] = 2.0 * (r3 * (b1 - b2) - bx * r1 + by * r2)
        w = -b3 * (b1 - b2) + bx * bx - by * by
    elif kind == ROTATION_WITHOUT_SHEAR:         # -b3 (r1 . r2) + bx by = 0
        g[0:3] = -b3 * r2 + by * r3
        g[4:7] = -b3 * r1 + bx * r3
        g[8:11] = -2.0 * r3 * (r1 @ r2) + by * r1 + bx * r2
        w = b3 * (r1 @ r2) - bx * by
    else:
        raise ValueError(kind)
    return g, w


def design_rows(x, y, X, Y, Z):
    """The two rows of A of one point (DPF:262-312); columns b11 b12 b13 b14 b21 b22 b23 b24 b31 b32 b33."""
    a0 = np.array([X, Y, Z, 1.0, 0, 0, 0, 0, -x * X, -x * Y, -x * Z])
    a1 = np.array([0, 0, 0, 0, X, Y, Z, 1.0, -y * X, -y * Y, -y * Z])
    return a0, a1


def expand(b, scale):
    """DLT:186-266: undo the coordinate scale, then (cx, cy, x0, y0) and the exterior orientation."""
    b = b.copy()
    for k in range(11):
        if k not in (3, 7):
            b[k] /= scale
    r1, r2, r3 = b[0:3], b[4:7], b[8:11]
    bb = r3 @ r3
    x0 = (r1 @ r3) / bb
    y0 = (r2 @ r3) / bb
    cx = np.sqrt((r1 @ r1) / bb - x0 * x0)
    cy = np.sqrt((r2 @ r2) / bb - y0 * y0)
    R = np.empty((3, 3))
    R[:, 0] = -(x0 * r3 - r1) / np.sqrt(bb) / cx
    R[:, 1] = -(y0 * r3 - r2) / np.sqrt(bb) / cy
    R[:, 2] = -r3 / np.sqrt(bb)
    if np.linalg.det(R) < 0:
        R = -R
    omega = np.arctan2(-R[1, 2], R[2, 2])
    phi = np.arcsin(R[0, 2])
    kappa = np.arctan2(-R[0, 1], R[0, 0])
    F = np.array([r1, r2, r3])
    t = np.linalg.solve(F, np.array([-b[3], -b[7], -1.0]))
    return b, dict(c=0.5 * (cx + cy), cx=cx, cy=cy, x0=x0, y0=y0, X0=t, omega=omega, phi=phi, kappa=kappa)


def adjust(xy, XYZ, io, restrictions=(), max_iterations=5000):
    """One image.  xy (m, 2), XYZ (m, 3) homologous points, io = (c, x0, y0) of the camera.
    Returns (ok, b (11, back-scaled), derived dict, passes)."""
    xy, XYZ = np.asarray(xy, float), np.asarray(XYZ, float)
    restrictions = validate_restrictions(restrictions)
    m = xy.shape[0]
    if m < 6:
        return False, None, None, 0
    sw, si = float((XYZ ** 2).sum()), float((xy ** 2).sum())
    scale = np.sqrt(sw / si) if si > 0 else 1.0
    Xs = XYZ / scale
    c, x0, y0 = io
    b = np.zeros(11)
    runs = max_iterations - 1
    is_estimated = complete = max_iterations == 0
    is_converge = True
    include = False
    passes = 0
    while True:
        passes += 1
        R = restrictions if include else []
        n = 11 + len(R)
        N = np.zeros((n, n))
        rhs = np.zeros(n)
        for k in range(m):
            a0, a1 = design_rows(xy[k, 0], xy[k, 1], *Xs[k])
            w0 = xy[k, 0] - a0 @ b
            w1 = xy[k, 1] - a1 @ b
            N[:11, :11] += np.outer(a0, a0) + np.outer(a1, a1)
            rhs[:11] += a0 * w0 + a1 * w1
        for j, kind in enumerate(R):
            g, w = restriction_row(kind, b, c, x0, y0)
            N[:11, 11 + j] = g
            N[11 + j, :11] = g
            rhs[11 + j] = w
        dg = np.diag(N)
        V = np.where(dg > EPS, 1.0 / np.sqrt(np.where(dg > EPS, dg, 1.0)), 1.0)
        Np = N * V[:, None] * V[None, :]
        iu = np.triu_indices(n)
        ap = np.zeros(n * (n + 1) // 2)
        ap[iu[0] + iu[1] * (iu[1] + 1) // 2] = Np[iu]
        x = rhs * V
        complete = is_estimated or len(restrictions) == 0
        try:
            solve_symm_packed(ap, x, n, False)
        except Exception:
            return False, None, None, passes
        dx = x * V
        b += dx[:11]
        max_abs = float(np.abs(dx[:11]).max())
        include = True
        if not np.isfinite(max_abs):
            return False, None, None, passes
        elif max_abs <= SQRT_EPS and runs > 0:
            is_estimated = True
        else:
            runs -= 1
            if runs + 1 <= 1:
                if complete:
                    is_converge = False
                is_estimated = True
        if complete:
            break
    bs, derived = expand(b, scale)
    return is_converge, bs, derived, passes
