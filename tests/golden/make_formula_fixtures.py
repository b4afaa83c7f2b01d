"""Generates tests/golden/reference_formulas.npz: golden vectors of the reference's closed-form expressions for

* the Jacobian of the exterior-orientation coordinate transformation
  (tranformation/CoordinateTransformationExteriorOrientation.java:160-279: 45 entries per point + the transformed point), and
* the DLT restrictions (dlt/DLTPartialDerivativeFactory.java:100-236: gradient rows and misclosures of the six
  RestrictionTypes) and the expansion of the 11 coefficients (dlt/DirectLinearTransformation.java:208-246).

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_formula_fixtures.py

How: the Java sources spell these formulas out as plain arithmetic on local `double` variables.  This script parses the
assignment lines and the `J.set(...)` / `NEQ.set(...)` / `neq.set(...)` calls and EVALUATES the right-hand sides for
seeded random inputs -- the reference's own expressions produce the numbers; no expression is restated here and none
is stored: the fixture holds inputs and outputs only.  tests/test_reference_formulas.py checks the oracles
(oracle/propagation.py, oracle/dlt.py) against these vectors.
"""
import math
import os
import re

import numpy as np

REF = '/root/reference/JAICOV/src/org/applied_geodesy/adjustment/bundle'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_formulas.npz')

ENV = {'__builtins__': {}, 'Math': math}
RE_ASSIGN = re.compile(r'^\s*double\s+(\w+)\s*=\s*(.+);\s*$')


def java_eval(expr, ns):
    return eval(expr, ENV, ns)


# ---- coordinate transformation -------------------------------------------------------------------------------------------
EO_NAMES = {'CAMERA_COORDINATE_X': 0, 'CAMERA_COORDINATE_Y': 1, 'CAMERA_COORDINATE_Z': 2, 'CAMERA_OMEGA': 3, 'CAMERA_PHI': 4,
            'CAMERA_KAPPA': 5}
COLUMN_ORDER = ['columnX0Trg', 'columnY0Trg', 'columnZ0Trg', 'columnOmegaTrg', 'columnPhiTrg', 'columnKappaTrg',
                'columnX0Src', 'columnY0Src', 'columnZ0Src', 'columnOmegaSrc', 'columnPhiSrc', 'columnKappaSrc',
                'columnXiSrc', 'columnYiSrc', 'columnZiSrc']


def transformation_vectors(rng, count):
    lines = open(os.path.join(REF, 'tranformation', 'CoordinateTransformationExteriorOrientation.java')).read().splitlines()
    start = next(i for i, l in enumerate(lines) if 'private static ObjectCoordinate setPartialDerivations' in l)
    body = lines[start:]
    else_at = next(i for i, l in enumerate(body) if l.strip() == 'else {')
    branch = body[else_at:]
    re_get = re.compile(r'exteriorOrientation(Trg|Src)\.get\(ParameterType\.(\w+)\)\.getValue\(\)')
    re_jset = re.compile(r'^\s*J\.set\((row[XYZ]),\s*(column\w+),\s*(.+)\);\s*$')
    inputs, jac, xyz = [], [], []
    for _ in range(count):
        eo = {'Trg': np.concatenate([rng.uniform(-3000, 3000, 3), rng.uniform(-math.pi, math.pi, 3)]),
              'Src': np.concatenate([rng.uniform(-3000, 3000, 3), rng.uniform(-math.pi, math.pi, 3)])}
        X = rng.uniform(-1000, 1000, 3)
        ns = {'XiSrc': X[0], 'YiSrc': X[1], 'ZiSrc': X[2]}
        J = np.zeros((3, 15))
        n_set = 0
        for line in branch:
            m = RE_ASSIGN.match(line)
            if m:
                name, expr = m.groups()
                g = re_get.search(expr)
                ns[name] = float(eo[g.group(1)][EO_NAMES[g.group(2)]]) if g else java_eval(expr, ns)
                continue
            m = re_jset.match(line)
            if m:
                row, col, expr = m.groups()
                J['XYZ'.index(row[-1]), COLUMN_ORDER.index(col)] = java_eval(expr, ns)
                n_set += 1
        assert n_set == 45, n_set
        inputs.append(np.concatenate([eo['Trg'], eo['Src'], X]))
        jac.append(J)
        xyz.append([ns['XiTrg'], ns['YiTrg'], ns['ZiTrg']])
    return np.array(inputs), np.array(jac), np.array(xyz)


# ---- DLT restrictions and expansion -----------------------------------------------------------------------------------------
B_NAMES = ['B11', 'B12', 'B13', 'B14', 'B21', 'B22', 'B23', 'B24', 'B31', 'B32', 'B33']
RESTRICTIONS = ['IDENTICAL_PRINCIPLE_DISTANCE', 'ROTATION_WITHOUT_SHEAR', 'FIXED_PRINCIPLE_DISTANCE_X', 'FIXED_PRINCIPLE_DISTANCE_Y',
                'FIXED_PRINCIPAL_POINT_X', 'FIXED_PRINCIPAL_POINT_Y']      # ordinal order, DirectLinearTransformation.java:50-57


def restriction_vectors(rng, count):
    lines = open(os.path.join(REF, 'dlt', 'DLTPartialDerivativeFactory.java')).read().splitlines()
    start = next(i for i, l in enumerate(lines) if 'static void setParameterRestrictions' in l)
    end = next(i for i, l in enumerate(lines) if 'static void addPartialNormalEquationOfDLTParameters' in l)
    body = lines[start:end]
    re_bval = re.compile(r'^\s*(?://\s*)?double\s+(b\d\d)\s*=\s*(B\d\d)\.getValue\(\);')
    re_io = re.compile(r'^\s*double\s+(c|x0|y0)\s*=\s*coefficients\.get\(ParameterType\.(\w+)\)\.getValue\(\);')
    re_case = re.compile(r'^\s*case\s+(\w+):')
    re_nset = re.compile(r'^\s*NEQ\.set\((B\d\d)\.getColumn\(\),\s*rowIndex,\s*(.+)\);\s*$')
    re_rhs = re.compile(r'^\s*neq\.set\(rowIndex\+\+,\s*(.+)\);\s*$')
    inputs, grads, miscl = [], [], []
    for _ in range(count):
        b = rng.normal(size=11)
        c, x0, y0 = rng.uniform(0.5, 3.0), rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2)
        ns = {}
        G = np.zeros((6, 11))
        W = np.zeros(6)
        cur = None
        seen = set()
        for line in body:
            m = re_bval.match(line)
            if m:
                ns[m.group(1)] = float(b[B_NAMES.index(m.group(2))])
                continue
            m = re_io.match(line)
            if m:
                ns[m.group(1)] = {'c': c, 'x0': x0, 'y0': y0}[m.group(1)]
                continue
            m = RE_ASSIGN.match(line)
            if m and '.getValue()' not in m.group(2):
                ns[m.group(1)] = java_eval(m.group(2), ns)
                continue
            m = re_case.match(line)
            if m:
                cur = RESTRICTIONS.index(m.group(1))
                continue
            m = re_nset.match(line)
            if m:
                G[cur, B_NAMES.index(m.group(1))] = java_eval(m.group(2), ns)
                continue
            m = re_rhs.match(line)
            if m:
                W[cur] = java_eval(m.group(1), ns)
                seen.add(cur)
        assert seen == set(range(6)), seen
        inputs.append(np.concatenate([b, [c, x0, y0]]))
        grads.append(G)
        miscl.append(W)
    return np.array(inputs), np.array(grads), np.array(miscl)


def expansion_vectors(rng, count):
    """x0, y0, cx, cy, the nine r_ij BEFORE the determinant flip, detR, and omega / phi / kappa after it (DLT:208-246)."""
    lines = open(os.path.join(REF, 'dlt', 'DirectLinearTransformation.java')).read().splitlines()
    start = next(i for i, l in enumerate(lines) if 'private static void expandUnknownParameters' in l)
    end = next(i for i, l in enumerate(lines) if 'private static RestrictionType[] validateRestrictions' in l)
    body = lines[start:end]
    re_bval = re.compile(r'^\s*double\s+(b\d\d)\s*=\s*coefficients\.get\(ParameterType\.DIRECT_LINEAR_TRANSFORMATION_(B\d\d)\)\.getValue\(\);')
    re_neg = re.compile(r'^\s*(r\d\d)\s*=\s*-\1;')
    inputs, outs = [], []
    for _ in range(count):
        # coefficients of a plausible camera: rows = c R' - x0 r3 ... are not needed, any full-rank B works for the formulas
        b = rng.normal(size=11)
        ns = {}
        in_flip = False
        for line in body:
            m = re_bval.match(line)
            if m:
                ns[m.group(1)] = float(b[B_NAMES.index(m.group(2))])
                continue
            if 'if (detR < 0)' in line:
                in_flip = True
                pre = {k: ns[k] for k in ('r11', 'r12', 'r13', 'r21', 'r22', 'r23', 'r31', 'r32', 'r33')}
                continue
            if in_flip:
                m = re_neg.match(line)
                if m:
                    if ns['detR'] < 0:
                        ns[m.group(1)] = -ns[m.group(1)]
                    continue
                if line.strip() == '}':
                    in_flip = False
                continue
            m = RE_ASSIGN.match(line)
            if m and '.getValue()' not in m.group(2):
                try:
                    ns[m.group(1)] = java_eval(m.group(2), ns)
                except ValueError:            # sqrt of a negative number for this random B: draw again
                    ns = None
                    break
        if ns is None or not all(np.isfinite([ns['cx'], ns['cy'], ns['phi']])):
            continue
        inputs.append(b)
        outs.append([ns['x0'], ns['y0'], ns['cx'], ns['cy']] + [pre[k] for k in sorted(pre)] + [ns['detR'], ns['omega'], ns['phi'], ns['kappa']])
    return np.array(inputs), np.array(outs)


def main():
    rng = np.random.default_rng(20261018)
    t_in, t_jac, t_xyz = transformation_vectors(rng, 12)
    r_in, r_grad, r_w = restriction_vectors(rng, 12)
    e_in, e_out = expansion_vectors(rng, 40)
    np.savez_compressed(OUT, transform_inputs=t_in, transform_jacobian=t_jac, transform_xyz=t_xyz,
                        restriction_inputs=r_in, restriction_gradients=r_grad, restriction_misclosures=r_w,
                        expansion_inputs=e_in, expansion_outputs=e_out)
    print('wrote', OUT, t_jac.shape, r_grad.shape, e_out.shape)


if __name__ == '__main__':
    main()
