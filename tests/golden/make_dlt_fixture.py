"""Generates tests/golden/reference_dlt.npz: direct linear transformations produced by EXECUTING the reference's own
DirectLinearTransformation.adjust (dlt/DirectLinearTransformation.java:67-184) with prepareUnknwonParameters (:279-314),
createNormalEquation (:316-350), updateUnknownParameters (:171-184), expandUnknownParameters (:186-266) and
DLTPartialDerivativeFactory.addPartialNormalEquationOfDLTParameters / setParameterRestrictions
(dlt/DLTPartialDerivativeFactory.java:100-337) on synthetic images with noisy observations.

Run in the build container only (reads /root/reference):
    python tests/golden/make_dlt_fixture.py

Method bodies are transliterated mechanically (make_jacobian_fixture.transliterate) and exec'ed on stub objects.  Local
rewrites, all of them ternaries inside argument lists that the transliterator does not expand: the two in adjust()
(`includeRestrictions ? restrictions : new RestrictionType[0]`, `!includeRestrictions ? numberOfUnknownParameters : n.size()`)
become calls of a two-line helper.  validateRestrictions (:268-277, a LinkedHashSet round trip) is restated.  Third-party
code: MathExtension.solve = LAPACK dspsv (oracle/lapack_packed.py), DenseMatrix.solve = LAPACK dgesv (numpy).
Numbers only are stored.
"""
import math
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_jacobian_fixture as tj  # noqa: E402
import make_lm_fixture as tl  # noqa: E402
import make_normal_equation_fixture as tn  # noqa: E402
from oracle.lapack_packed import MatrixSingularException, solve_symm_packed  # noqa: E402

DLT = os.path.join(tj.REF, 'dlt', 'DirectLinearTransformation.java')
DPF = os.path.join(tj.REF, 'dlt', 'DLTPartialDerivativeFactory.java')
OUT = os.path.join(HERE, 'reference_dlt.npz')
MAXV = tj.MAXV
B = ['DIRECT_LINEAR_TRANSFORMATION_B%d%d' % (i, j) for i in (1, 2, 3) for j in (1, 2, 3, 4)][:11]
ORDER = B + ['PRINCIPAL_POINT_X', 'PRINCIPAL_POINT_Y', 'PRINCIPAL_DISTANCE', 'CAMERA_COORDINATE_X', 'CAMERA_COORDINATE_Y',
             'CAMERA_COORDINATE_Z', 'CAMERA_OMEGA', 'CAMERA_PHI', 'CAMERA_KAPPA']
RESTRICTIONS = ['IDENTICAL_PRINCIPLE_DISTANCE', 'ROTATION_WITHOUT_SHEAR', 'FIXED_PRINCIPLE_DISTANCE_X', 'FIXED_PRINCIPLE_DISTANCE_Y',
                'FIXED_PRINCIPAL_POINT_X', 'FIXED_PRINCIPAL_POINT_Y']


class Param:
    def __init__(self, ptype, value=0.0, column=-1): self.ptype, self.value, self.column = ptype, float(value), column
    def getParameterType(self): return self.ptype
    def getValue(self): return self.value
    def setValue(self, v): self.value = float(v)
    def getColumn(self): return self.column
    def setColumn(self, c): self.column = c


class Coefficients:
    """dlt/DLTCoefficients.java:33-84: a LinkedHashMap of the 20 parameters in this order."""
    def __init__(self, image): self.image, self.params = image, {t: Param(t) for t in ORDER}
    def get(self, t): return self.params[t]
    def getReference(self): return self.image
    def __iter__(self): return iter(self.params.values())


class Obs:
    def __init__(self, v): self.v = float(v)
    def getValue(self): return self.v


class Named:
    def __init__(self, name, xyz): self.name, self.p = name, [Obs(v) for v in xyz]
    def getName(self): return self.name
    def getX(self): return self.p[0]
    def getY(self): return self.p[1]
    def getZ(self): return self.p[2]


class ImageCoordinate:
    def __init__(self, point, xy): self.point, self.p = point, [Obs(xy[0]), Obs(xy[1])]
    def getObjectCoordinate(self): return self.point
    def getX(self): return self.p[0]
    def getY(self): return self.p[1]


class Interior:
    def __init__(self, c, x0, y0, fixed):
        self.c, self.x0, self.y0 = (Param(t, v, MAXV if f else -1) for t, v, f in
                                    zip(('PRINCIPAL_DISTANCE', 'PRINCIPAL_POINT_X', 'PRINCIPAL_POINT_Y'), (c, x0, y0), fixed))

    def getPrincipleDistance(self): return self.c
    def getPrinciplePointX(self): return self.x0
    def getPrinciplePointY(self): return self.y0


class Camera:
    def __init__(self, io): self.io = io
    def getInteriorOrientation(self): return self.io


class Image(list):
    def __init__(self, ident, camera, coords):
        super().__init__(coords)
        self.ident, self.camera = ident, camera

    def getId(self): return self.ident
    def getReference(self): return self.camera


class JMap(dict):
    def containsKey(self, k): return k in self


class DenseMatrix3:
    """new DenseMatrix(double[][]).solve(DenseVector, DenseVector): LAPACK dgesv."""
    def __init__(self, rows): self.a = np.array(rows, float)
    def solve(self, f, t):
        t.v[:] = np.linalg.solve(self.a, f.v)
        return t


class Vec3(tn.DenseVector):
    def __init__(self, arg):
        if isinstance(arg, int):
            super().__init__(arg)
        else:
            self.v = np.array(arg, float)


def build():
    class ME:
        @staticmethod
        def solve(N, n, num_rows, invert):
            solve_symm_packed(N.ap, n.v, int(num_rows), bool(invert))
    g = {'math': tl.JavaMath, 'SQRT_EPS': tl.SQRT_EPS, 'MathExtension': ME, 'MatrixSingularException': MatrixSingularException,
         'JList': tn.JList, 'DenseVector': tn.DenseVector, 'UpperSymmPackMatrix': tn.UpperSymmPackMatrix,
         'UpperSymmBandMatrix': tn.UpperSymmBandMatrix, 'NormalEquationSystem': None, 'numberOfUnknownParameters': 11,
         'maximalNumberOfIterations': 5000}

    def print_stack_trace(e):            # a harness bug must not pass for `adjust() returned false`
        if isinstance(e, (AttributeError, NameError, TypeError, KeyError, IndexError)):
            raise e
    g['printStackTrace'] = print_stack_trace

    class NES(tn.NormalEquationSystem):
        pass
    NES.getMatrix = lambda self: self.N
    NES.getVector = lambda self: self.n
    NES.getPreconditioner = lambda self: self.V
    nes = os.path.join(os.path.dirname(tj.REF), 'NormalEquationSystem.java')
    exec(tj.transliterate(tl.ternaries(tj.method_body(nes, 'public static void applyPrecondition(UpperSymmBandMatrix V')), 'def applyPrecondition3(V, M, m):'), g)
    NES.applyPrecondition = staticmethod(lambda *a: g['applyPrecondition3'](a[0].V, a[0].N, a[0].n) if len(a) == 1 else g['applyPrecondition3'](*a))
    g['NormalEquationSystem'] = NES
    fix = lambda src: src.replace('Constant.EPS', repr(2.0 ** -53))

    # DLTPartialDerivativeFactory
    class DPFNS:
        pass
    exec(tj.transliterate(tj.method_body(DPF, 'static void addPartialNormalEquationOfDLTParameters('),
                          'def addPartialNormalEquationOfDLTParameters(NEQ, neq, coefficients, x, y, X, Y, Z):'), g)
    exec(tj.transliterate(tj.method_body(DPF, 'static void setParameterRestrictions('),
                          'def setParameterRestrictions(NEQ, neq, rowIndex, coefficients, *restrictions):'), g)
    DPFNS.addPartialNormalEquationOfDLTParameters = staticmethod(g['addPartialNormalEquationOfDLTParameters'])
    DPFNS.setParameterRestrictions = staticmethod(g['setParameterRestrictions'])
    g['DLTPartialDerivativeFactory'] = DPFNS
    g['DenseMatrix'] = lambda *a: DenseMatrix3(a[0]) if len(a) == 1 else tn.DenseMatrix(*a)
    # DirectLinearTransformation
    body, joined, buf = tj.method_body(DLT, 'private static void prepareUnknwonParameters('), [], None
    for l in body:                        # the array initialiser `T name[] = new T[] { a, b, c };` spread over several lines -> one statement
        if buf is None and re.search(r'\[\]\s*=\s*new\s+[\w<>\?]+\[\]\s*\{\s*$', l):
            buf = re.sub(r'^\s*[\w<>\?]+\s+(\w+)\[\]\s*=.*$', r'\1 = JList([', l)
            continue
        if buf is not None:
            if l.strip() == '};':
                joined.append(buf + ']);')
                buf = None
            else:
                buf += l.strip()
            continue
        joined.append(l)
    exec(tj.transliterate(tl.ternaries(joined), 'def prepareUnknwonParameters(coefficients):'), g)
    src = fix(tj.transliterate(tl.ternaries(tj.method_body(DLT, 'private static NormalEquationSystem createNormalEquation(')),
                               'def createNormalEquation(coefficients, homologousImageCoordinates, objectCoordinates, scale, *restrictions):'))
    # a Java varargs array handed on to another varargs method stays one array: spread it again
    src = src.replace('numberOfUnknownParameters, coefficients, restrictions)', 'numberOfUnknownParameters, coefficients, *restrictions)')
    exec(src, g)
    exec(tj.transliterate(tj.method_body(DLT, 'private static double updateUnknownParameters('), 'def updateUnknownParameters(coefficients, dx):'), g)
    body = tj.method_body(DLT, 'private static void expandUnknownParameters(')
    body = [l.replace('new DenseMatrix(new double[][] {{b11, b12, b13}, {b21, b22, b23}, {b31, b32, b33}})', 'DenseMatrix([[b11, b12, b13], [b21, b22, b23], [b31, b32, b33]])')
             .replace('new DenseVector(new double[] {-b14, -b24, -1.0})', 'Vec3([-b14, -b24, -1.0])') for l in body]
    g['Vec3'] = Vec3
    exec(tj.transliterate(body, 'def expandUnknownParameters(coefficients, scale):'), g)

    def validate(*restrictions):         # :268-277
        out = list(dict.fromkeys(restrictions))
        if all(r in out for r in ('FIXED_PRINCIPLE_DISTANCE_X', 'FIXED_PRINCIPLE_DISTANCE_Y', 'IDENTICAL_PRINCIPLE_DISTANCE')):
            out.remove('IDENTICAL_PRINCIPLE_DISTANCE')
        return out
    g['validateRestrictions'] = validate
    g['selectRestrictions'] = lambda include, restrictions: restrictions if include else []
    g['selectRows'] = lambda include, n: 11 if not include else n.size()
    body = tj.method_body(DLT, 'public static boolean adjust(')
    body = [l.replace('includeRestrictions ? restrictions : new RestrictionType[0]', '*selectRestrictions(includeRestrictions, restrictions)')
             .replace('!includeRestrictions ? numberOfUnknownParameters : n.size()', 'selectRows(includeRestrictions, n)')
             .replace('restrictions = validateRestrictions(restrictions);', 'restrictions = validateRestrictions(*restrictions);')
             .replace('new RestrictionType[0]', 'JList()') for l in body]
    body = [re.sub(r'throw new MatrixSingularException\(.*\);', 'throw new MatrixSingularException();', l) for l in body]
    # the loop counter lives in a holder object: its post-decrement sits inside an `else if` condition (:162)
    body = [re.sub(r'\bint runs\b', 'R.runs', l) for l in body]
    body = [l.replace('runs-- <= 1', 'R.postDecrement() <= 1') for l in body]
    body = [re.sub(r'(?<![\w\.])runs\b', 'R.runs', l) for l in body]

    class R:
        runs = 0

        @staticmethod
        def postDecrement():
            R.runs -= 1
            return R.runs + 1
    g['R'] = R
    src = tj.transliterate(tl.ternaries(body), 'def adjust(coefficients, objectCoordinates, *restrictions):')
    src = src.replace('raise ValueError()', 'raise MatrixSingularException()')
    exec(src, g)
    return g


def network(noise=0.002, images=6, targets=60, seed=5):
    from tests.scenes import project, synthetic_scene
    sc, truth = synthetic_scene(2, images=images, targets=targets)
    rng = np.random.default_rng(seed)
    io, eo, pts = truth['io'], truth['eo'], truth['points']
    obs = []
    for i in range(images):
        xy, _ = project(io, [], 10.0, eo[i], pts)
        keep = rng.uniform(size=targets) < 0.8
        obs.append((np.nonzero(keep)[0], xy[keep] + rng.normal(0, noise, size=(int(keep.sum()), 2))))
    obs[3] = (obs[3][0][:5], obs[3][1][:5])          # too few points: adjust() returns false (:96-104)
    return truth, obs


SETS = [(), ('IDENTICAL_PRINCIPLE_DISTANCE', 'ROTATION_WITHOUT_SHEAR'),
        ('FIXED_PRINCIPLE_DISTANCE_X', 'FIXED_PRINCIPLE_DISTANCE_Y', 'IDENTICAL_PRINCIPLE_DISTANCE', 'ROTATION_WITHOUT_SHEAR'),
        ('FIXED_PRINCIPAL_POINT_X', 'FIXED_PRINCIPAL_POINT_Y', 'ROTATION_WITHOUT_SHEAR', 'ROTATION_WITHOUT_SHEAR'), ('FIXED_PRINCIPLE_DISTANCE_X',)]


def main():
    g = build()
    truth, obs = network()
    io, pts = truth['io'], truth['points']
    c, x0, y0 = io[2] * 1.001, io[0] + 0.01, io[1] - 0.01
    out = {'pt_ptr': np.concatenate([[0], np.cumsum([len(i) for i, _ in obs])]), 'xy': np.concatenate([x for _, x in obs]),
           'xyz': np.concatenate([pts[i] for i, _ in obs]), 'io': np.array([c, x0, y0])}
    known = JMap({str(k): Named(str(k), pts[k]) for k in range(len(pts))})
    for s, rset in enumerate(SETS):
        res, oks = [], []
        for k, (idx, xy) in enumerate(obs):
            camera = Camera(Interior(c, x0, y0, (False, False, False)))
            image = Image(k + 1, camera, [ImageCoordinate(known[str(i)], p) for i, p in zip(idx, xy)])
            coef = Coefficients(image)
            ok = g['adjust'](coef, known, *rset)
            oks.append(bool(ok))
            res.append([p.getValue() for p in coef])
        out['set%d_restrictions' % s] = np.array([RESTRICTIONS.index(r) for r in rset], np.int64)
        out['set%d_ok' % s] = np.array(oks)
        out['set%d_values' % s] = np.array(res)
        print(rset, oks)
    np.savez_compressed(OUT, **out)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
