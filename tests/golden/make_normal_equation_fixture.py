"""Generates tests/golden/reference_normal_equations.npz: the normal equations N (MTJ packed upper, datum border included)
and n of small networks, produced by EXECUTING the reference's own code path

    BundleAdjustment.prepareUnknownParameters / detectRankDefect                       (columns and rows)
    PartialDerivativeFactory.getPartialDerivativeImageCoordinate                     derivation/PartialDerivativeFactory.java:285-445
    PartialDerivativeFactory.getPartialDerivativeScaleBar                            :210-283
    PartialDerivativeFactory.stackNormalEquationSystem                               :475-505
    BundleAdjustment.addDatumConditionRows                                           BundleAdjustment.java:493-635

    PartialDerivativeFactory.getPartialDerivativeDirectlyObservedParameters          :447-473
    DirectlyObservedParameterGroup.getWeightMatrix                                   parameter/DirectlyObservedParameterGroup.java:67-91
    BundleAdjustment.centroidCoordinates                                             BundleAdjustment.java:115-201

driven by the executed BundleAdjustment.createNormalEquation itself (:789-834: every observation group in insertion
order, the datum rows, the Levenberg-Marquardt damping of the diagonal, the Jacobi preconditioner V), followed by
NormalEquationSystem.applyPrecondition (adjustment/NormalEquationSystem.java:82-91).  MathExtension.inv(UpperSPDPackMatrix) (dpptrf + dpptri, MathExtension.java:304-324) is third-party LAPACK in
the reference; here it is the same LAPACK routine pair out of scipy's OpenBLAS (oracle/lapack_packed.py).  Method bodies are transliterated mechanically (make_jacobian_fixture.transliterate) and exec'ed on stub
objects; the `switch` over the camera's distortion models (:420-444) is replaced by make_jacobian_fixture.apply_models,
which calls the transliterated model factories in the reference's order.  Numbers only are stored.

Run in the build container only (reads /root/reference):
    python tests/golden/make_normal_equation_fixture.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_bookkeeping_fixture as tb  # noqa: E402
import make_jacobian_fixture as tj  # noqa: E402
import make_lm_fixture as tl  # noqa: E402
from oracle.lapack_packed import inv_spd_packed, inv_symm_packed  # noqa: E402

PDF = os.path.join(tj.REF, 'derivation', 'PartialDerivativeFactory.java')
OUT = os.path.join(HERE, 'reference_normal_equations.npz')
MAXV = tj.MAXV


# ---- MTJ stand-ins -----------------------------------------------------------------------------------------------------------------
class DenseVector:
    def __init__(self, n): self.v = np.zeros(n)
    def size(self): return self.v.size
    def get(self, i): return self.v[i]
    def set(self, i, x): self.v[i] = x
    def add(self, *args):
        if len(args) == 2 and not isinstance(args[1], DenseVector):
            self.v[args[0]] += args[1]                      # add(index, value)
        else:
            self.v += args[0] * args[1].v                   # add(alpha, vector): daxpy

    def zero(self): self.v[:] = 0.0


class NormalEquationSystem:
    def __init__(self, N, n, V): self.N, self.n, self.V = N, n, V


class DenseMatrix:
    def __init__(self, r, c): self.v = np.zeros((r, c))
    def numRows(self): return self.v.shape[0]
    def get(self, r, c): return self.v[r, c]
    def set(self, r, c, x): self.v[r, c] = x
    def add(self, r, c, x): self.v[r, c] += x


class UpperSymmBandMatrix:
    """Only ever used with zero off-diagonals (kd = 0): a diagonal."""
    def __init__(self, n, kd): self.d = np.zeros(n)
    def numRows(self): return self.d.size
    def get(self, r, c): return self.d[r] if r == c else 0.0
    def set(self, r, c, x): self.d[r] = x


class UpperSymmPackMatrix:
    """no.uib.cipr.matrix.UpperSymmPackMatrix (MTJ 1.0.4): symmetric, stored column-major packed upper.  get() is symmetric;
    set() and add() act on the stored triangle only -- a write below the diagonal is ignored.  The reference relies on that:
    reduceNormalEquationSystem (BundleAdjustment.java:1325-1340) loops over all (row, column) pairs and each element is
    updated once, through its upper-triangle address."""
    def __init__(self, n): self.n, self.ap = n, np.zeros(n * (n + 1) // 2)
    def numRows(self): return self.n
    def numColumns(self): return self.n
    def get(self, r, c): return self.ap[(r + c * (c + 1) // 2) if r <= c else (c + r * (r + 1) // 2)]

    def set(self, r, c, x):
        if r <= c:
            self.ap[r + c * (c + 1) // 2] = x

    def add(self, r, c, x):
        if r <= c:
            self.ap[r + c * (c + 1) // 2] += x


class UpperSPDPackMatrix(UpperSymmPackMatrix):
    def scale(self, a):
        self.ap *= a
        return self


class MathExtension:
    @staticmethod
    def inv(M):
        if isinstance(M, UpperSPDPackMatrix):
            inv_spd_packed(M.ap, M.n)           # MathExtension.inv(UpperSPDPackMatrix): dpptrf + dpptri, MathExtension.java:304-324
        else:
            inv_symm_packed(M.ap, M.n)          # MathExtension.inv(UpperSymmPackMatrix): dsptrf + dsptri, :403-426
        return M


class Group(list):
    """DirectlyObservedParameterGroup (parameter/DirectlyObservedParameterGroup.java:36-60): observations, optional dispersion."""
    def __init__(self, obs, dispersion):
        super().__init__(obs)
        self.observedParameters = self
        self.sigma2apriori = -1
        self.weightMatrix = None
        if dispersion is not None:
            self.weightMatrix = UpperSPDPackMatrix(len(obs))
            self.weightMatrix.ap[:] = np.asarray(dispersion, float)

    def hasFullyPopulatedWeightMatrix(self): return self.weightMatrix is not None
    def getNumberOfParameters(self): return len(self)


class Centroid:
    def __init__(self): self.p = [VUP('C', 0.0), VUP('C', 0.0), VUP('C', 0.0)]
    def getX(self): return self.p[0]
    def getY(self): return self.p[1]
    def getZ(self): return self.p[2]


class JSet(set):
    pass


class JList(list):
    def __init__(self, arg=()):
        super().__init__(() if isinstance(arg, int) else arg)       # new ArrayList<T>(capacity) | new ArrayList<T>(collection)

    def get(self, i): return self[i]
    def size(self): return len(self)
    def add(self, x): self.append(x)
    def addAll(self, other): self.extend(other)


class GaussMarkovEquations:
    def __init__(self, A, P, w): self.A, self.P, self.w = A, P, w


# ---- object graph with values ----------------------------------------------------------------------------------------------------------
class VUP(tb.UP):
    def __init__(self, ptype, value, ref=None, fixed=False, order=0, poly=None):
        super().__init__(ptype, ref, fixed)
        self.value, self.order, self.poly = float(value), order, poly

    def getValue(self): return self.value
    def setValue(self, v): self.value = v
    def getOrder(self): return self.order
    def getZernikePolynomial(self): return self.poly


class VOP(tb.OP):
    def __init__(self, value, variance):
        super().__init__(variance)
        self.value = float(value)

    def getValue(self): return self.value


class Point(tb.Point):
    def __init__(self, xyz, fixed, datum):
        self.p = [VUP('OBJECT_COORDINATE_' + 'XYZ'[k], xyz[k], self, bool(fixed[k])) for k in range(3)]
        self.seen, self.datum = False, bool(datum)

    def isDatum(self): return self.datum


class ImageCoordinate(list):
    def __init__(self, point, xy, sigma, rho, image):
        super().__init__([VOP(xy[0], sigma[0] ** 2), VOP(xy[1], sigma[1] ** 2)])
        self.point, self.rho, self.image = point, float(rho), image
        point.seen = True

    def getX(self): return self[0]
    def getY(self): return self[1]
    def getObjectCoordinate(self): return self.point
    def getReference(self): return self.image
    def getCorrelationCoefficientXY(self): return self.rho
    def getNumberOfParameters(self): return 2


class Exterior(list):
    def get(self, name): return next(p for p in self if p.ptype == name)


class Image(list):
    def __init__(self, camera, eo_val, eo_fixed):
        super().__init__()
        self.camera = camera
        self.eo = Exterior([VUP(n, v, None, bool(f)) for n, v, f in zip(tj.EO, eo_val, eo_fixed)])

    def getReference(self): return self.camera
    def getExteriorOrientation(self): return self.eo


class Interior(list):
    def getPrinciplePointX(self): return self[0]
    def getPrinciplePointY(self): return self[1]
    def getPrincipleDistance(self): return self[2]


class Camera(list):
    def __init__(self, r0, io_val, io_fixed, coefs):
        super().__init__()
        self.r0 = r0
        self.io = Interior([VUP(n, v, None, bool(f)) for n, v, f in
                            zip(('PRINCIPAL_POINT_X', 'PRINCIPAL_POINT_Y', 'PRINCIPAL_DISTANCE'), io_val, io_fixed)])
        self.coefs = [VUP(tj.COEF_TYPES[t], v, None, bool(f), o, tj.ZernikePoly(o) if t in (161, 162, 163) else None) for (t, o, v, f) in coefs]
        self.by_type = {}
        for (t, _o, _v, _f), p in zip(coefs, self.coefs):
            self.by_type.setdefault(t, []).append(p)
        self.models = [self.coefs]

    def getInteriorOrientation(self): return self.io
    def getDistortionModels(self): return self.models


class ScaleBar(list):
    def __init__(self, a, b, length, sigma):
        super().__init__([VOP(length, sigma * sigma)])
        self.a, self.b = a, b

    def getLength(self): return self[0]
    def getObjectCoordinateA(self): return self.a
    def getObjectCoordinateB(self): return self.b
    def getNumberOfParameters(self): return 1


def build(g):
    """Transliterated PartialDerivativeFactory methods + BundleAdjustment.addDatumConditionRows in namespace g."""
    g.update(UpperSPDPackMatrix=UpperSPDPackMatrix, MathExtension=MathExtension)
    g.update(DenseVector=DenseVector, DenseMatrix=DenseMatrix, UpperSymmBandMatrix=UpperSymmBandMatrix, UpperSymmPackMatrix=UpperSymmPackMatrix,
             JSet=JSet, JList=JList, GaussMarkovEquations=GaussMarkovEquations)

    class CEF:
        @staticmethod
        def getInstance(io, eo, pt):
            ce = tj.Collinearity()
            g['collinearity_init'](ce, io, eo, pt)
            return ce
    g['CollinearityEquationFactory'] = CEF
    g['applyModels'] = lambda camera, ce, columns, A, w: tj.apply_models(g, camera.by_type, camera.r0, ce, columns, A, w)
    body = tj.method_body(PDF, 'private static GaussMarkovEquations getPartialDerivativeImageCoordinate(')
    # the loop with the `switch` over the distortion models (:420-444) -> one call of the dispatcher
    i0 = next(k for k, l in enumerate(body) if l.strip().startswith('for (DistortionModel distortionModel : distortionModels)'))
    depth, i1 = 0, i0
    while True:
        depth += body[i1].count('{') - body[i1].count('}')
        if depth == 0:
            break
        i1 += 1
    body = body[:i0] + ['applyModels(camera, collinearityEquation, columns, A, w);'] + body[i1 + 1:]
    exec(tj.transliterate(body, 'def getPartialDerivativeImageCoordinate(sigma2apriori, NEQ, neq, imageCoordinate):'), g)
    exec(tj.transliterate(tj.method_body(PDF, 'private static GaussMarkovEquations getPartialDerivativeScaleBar('),
                          'def getPartialDerivativeScaleBar(sigma2apriori, NEQ, neq, scaleBar):'), g)
    exec(tj.transliterate(tj.method_body(PDF, 'private static GaussMarkovEquations stackNormalEquationSystem('),
                          'def stackNormalEquationSystem(NEQ, neq, A, P, w, columns, diagonalWeighting):'), g)
    exec(tj.transliterate(tj.method_body(PDF, 'private static GaussMarkovEquations getPartialDerivativeDirectlyObservedParameters('),
                          'def getPartialDerivativeDirectlyObservedParameters(sigma2apriori, NEQ, neq, observedParameterGroup):'), g)
    dopg = os.path.join(tj.REF, 'parameter', 'DirectlyObservedParameterGroup.java')
    exec(tj.transliterate(tj.method_body(dopg, 'public Matrix getWeightMatrix('), 'def getWeightMatrix(self, sigma2apriori):'), g)
    Group.getWeightMatrix = g['getWeightMatrix']
    exec(tj.transliterate(tl.ternaries(tj.method_body(tb.BA, 'private void centroidCoordinates(')), 'def centroidCoordinates(self, invert):'), g)
    tb.Adjustment.centroidCoordinates = g['centroidCoordinates']
    exec(tj.transliterate(tj.method_body(tb.BA, 'private void addDatumConditionRows('), 'def addDatumConditionRows(self, N):'), g)
    tb.Adjustment.addDatumConditionRows = g['addDatumConditionRows']

    class PDF_:                                             # PartialDerivativeFactory.getPartialDerivative, :199-208 (instanceof chain)
        @staticmethod
        def getPartialDerivative(sigma2apriori, NEQ, neq, observations):
            if isinstance(observations, ImageCoordinate):
                return g['getPartialDerivativeImageCoordinate'](sigma2apriori, NEQ, neq, observations)
            if isinstance(observations, ScaleBar):
                return g['getPartialDerivativeScaleBar'](sigma2apriori, NEQ, neq, observations)
            return g['getPartialDerivativeDirectlyObservedParameters'](sigma2apriori, NEQ, neq, observations)
    g.update(PartialDerivativeFactory=PDF_, NormalEquationSystem=NormalEquationSystem)
    fix = lambda src: src.replace('Constant.EPS', repr(2.0 ** -53)).replace('EstimationType.SIMULATION', "'SIMULATION'")
    g['math'] = tl.JavaMath
    exec(fix(tj.transliterate(tl.ternaries(tj.method_body(tb.BA, 'public NormalEquationSystem createNormalEquation(')), 'def createNormalEquation(self):')), g)
    tb.Adjustment.createNormalEquation = g['createNormalEquation']
    nes = os.path.join(os.path.dirname(tj.REF), 'NormalEquationSystem.java')
    exec(tj.transliterate(tl.ternaries(tj.method_body(nes, 'public static void applyPrecondition(UpperSymmBandMatrix V')),
                          'def applyPrecondition(V, M, m):'), g)
    tb.Adjustment.getClass = lambda self: 'BundleAdjustment'


def graph_unprepared(scene):
    """The object graph as a user of the reference would have built it (estimateModel calls prepareUnknownParameters itself)."""
    return graph(scene, prepare=False)


def graph(scene, prepare=True):
    pts = scene['points']
    P = [Point(pts['xyz'][k], pts['fixed'][k], pts['datum'][k]) for k in range(len(pts['xyz']))]
    adj = tb.Adjustment()
    for cam in scene['cameras']:
        camera = Camera(cam['r0'], cam['io_val'], cam['io_fixed'], cam['coefs'])
        for im in cam['images']:
            image = Image(camera, im['eo_val'], im['eo_fixed'])
            for o, xy, sg, rho in zip(im['obj'], np.asarray(im['xy'], float).reshape(-1, 2), np.asarray(im['sigma'], float).reshape(-1, 2), im['rho']):
                image.append(ImageCoordinate(P[int(o)], xy, sg, rho, image))
            camera.append(image)
        adj.cameras.append(camera)
    for (a, b, length, sigma) in scene.get('scale_bars', []):
        adj.scaleBars.add(ScaleBar(P[int(a)], P[int(b)], float(length), float(sigma)))
    images = [im for c in adj.cameras for im in c]

    def target(kind, index, comp):
        if kind == 'point':
            return P[index].p[comp]
        if kind == 'io':
            return adj.cameras[index].io[comp]
        if kind == 'coef':
            return adj.cameras[index].coefs[comp]
        return images[index].eo[comp]
    for grp in scene.get('observed_groups', []):
        obs = [tb.ObsParam(target(*ref), v, val) for ref, v, val in zip(grp['refs'], tb.variances_of(grp), grp['obs'])]
        adj.observedParameterGroups.append(Group(obs, grp.get('dispersion')))
    adj.centroid = Centroid()
    if prepare:
        adj.prepareUnknownParameters()
    return adj, P, images


def run_centroid(scene):
    """centroidCoordinates(false) on the freshly prepared network: the centroid and every shifted value."""
    adj, P, images = graph(scene)
    adj.centroidCoordinates(False)
    return (np.array([p.getValue() for p in adj.centroid.p]), np.array([[q.getValue() for q in pt.p] for pt in P]),
            np.array([[q.getValue() for q in im.eo] for im in images]),
            np.array([o.getValue() for grp in adj.observedParameterGroups for o in grp]))


def run(g, scene, damping=0.0):
    """createNormalEquation() as the first pass of estimateModel() sees it (:207-208: the damping value is taken over in the
    first pass), then applyPrecondition.  Returns N, n, V and the preconditioned N, n."""
    adj, P, images = graph(scene)
    adj.estimationType = 'L2NORM'
    adj.dampingValue, adj.adaptedDampingValue, adj.deriveFirstAdaptedDampingValue = damping, 0.0, damping > 0
    neq = adj.createNormalEquation()
    N, nv, V = neq.N.ap.copy(), neq.n.v.copy(), neq.V.d.copy()
    g['applyPrecondition'](neq.V, neq.N, neq.n)
    return N, nv, V, neq.N.ap.copy(), neq.n.v.copy()


def scenes():
    from tests.scenes import random_scene, synthetic_scene
    yield 'random2_scale_bar', random_scene(2)
    yield 'random0_two_cameras', random_scene(0)
    sc = synthetic_scene(3, images=5, targets=30)[0]         # correlated image coordinates (rho != 0): full 2 x 2 weights
    sc['observed_groups'] = []
    yield 'config3_rho_no_groups', sc
    sc = synthetic_scene(2, images=5, targets=30)[0]         # Zernike + B_i coefficients
    cam = sc['cameras'][0]
    cam['coefs'] = cam['coefs'][:4] + [(131, 1, 1e-4, False)] + cam['coefs'][4:] + [(161, 3, 1e-4, False), (162, 4, 3e-5, False), (163, 5, 2e-5, False)]
    yield 'config2_zernike_bi', sc
    yield 'config3_observed_points_dispersion', synthetic_scene(3, images=5, targets=25)[0]
    yield 'observed_eo_io', tb.observed_eo_io_scene()


def main():
    g = tj.build_functions()
    tb.build_methods()
    build(g)
    out = {}
    for name, sc in scenes():
        N, n, V, Np, npre = run(g, sc)
        out[name + '__N'] = N
        out[name + '__n'] = n
        out[name + '__V'] = V
        out[name + '__N_preconditioned'] = Np
        out[name + '__n_preconditioned'] = npre
        Nd, _, Vd, _, _ = run(g, sc, damping=0.7)
        out[name + '__N_damped'] = Nd
        out[name + '__V_damped'] = Vd
        print(name, N.size, float(np.abs(N).max()), float(np.abs(n).max()))
        try:
            c, xyz, eo, gobs = run_centroid(sc)
            out[name + '__centroid'] = c
            out[name + '__centroid_xyz'] = xyz
            out[name + '__centroid_eo'] = eo
            out[name + '__centroid_obs'] = gobs
        except ValueError:                       # UnsupportedOperationException of BundleAdjustment.java:151: unequal counts
            out[name + '__centroid_refused'] = np.array([1])
    np.savez_compressed(OUT, **out)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
