"""Generates tests/golden/reference_bookkeeping.npz: the integer bookkeeping of the reference -- rows, columns, counts,
rank-defect flags, sigma2apriori -- produced by EXECUTING the reference's own methods

    BundleAdjustment.prepareUnknownParameters     BundleAdjustment.java:667-782
    BundleAdjustment.addUnknownParameter / addObservationGroup                :637-650
    BundleAdjustment.detectRankDefect                                         :836-1042

on the networks the parity tests use (tests/scenes.py: random_scene(0..11), the bundled example, configs 2 and 4 scaled down).

Run in the build container only (reads /root/reference):
    python tests/golden/make_bookkeeping_fixture.py

How: as in make_jacobian_fixture.py the Java method bodies are transliterated to Python text mechanically and exec'ed
against stub objects standing in for the Java object graph (cameras -> images -> image coordinates -> object coordinates,
scale bars, directly observed parameter groups).  RankDefect (defect/RankDefect.java:25-130) is a bag of seven flags and
is restated as a stub.  The fixture stores numbers only.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_jacobian_fixture as tj  # noqa: E402

BA = os.path.join(tj.REF, 'BundleAdjustment.java')
OUT = os.path.join(HERE, 'reference_bookkeeping.npz')
MAXV = tj.MAXV


def cut_group_loops(lines):
    """Drops every block that starts with `for (DirectlyObservedParameterGroup ...` (brace matching)."""
    out, skip = [], 0
    for l in lines:
        if skip == 0 and l.strip().startswith('for (DirectlyObservedParameterGroup'):
            skip = l.count('{') - l.count('}')
            continue
        if skip > 0:
            skip += l.count('{') - l.count('}')
            continue
        out.append(l)
    return out


def pre(lines):
    return lines


# ---- stubs -----------------------------------------------------------------------------------------------------------------------
class OrderedSet:
    def __init__(self): self.d = {}
    def add(self, x): self.d.setdefault(id(x), x)
    def contains(self, x): return id(x) in self.d
    def isEmpty(self): return not self.d
    def size(self): return len(self.d)
    def __iter__(self): return iter(list(self.d.values()))
    def __len__(self): return len(self.d)


class UP:
    """UnknownParameter: value, column (-1 unset, Integer.MAX_VALUE fixed)."""
    def __init__(self, ptype, ref=None, fixed=False):
        self.ptype, self.ref, self.column = ptype, ref, MAXV if fixed else -1

    def getColumn(self): return self.column
    def setColumn(self, c): self.column = c
    def getParameterType(self): return self.ptype
    def getReference(self): return self.ref


class OP:
    """ObservationParameter: variance, row."""
    def __init__(self, variance): self.variance, self.row = float(variance), -1
    def getVariance(self): return self.variance
    def setRow(self, r): self.row = r


class It:
    def __init__(self, has): self.has = has
    def hasNext(self): return self.has


class Point:
    def __init__(self, fixed):
        self.p = [UP('OBJECT_COORDINATE_' + 'XYZ'[k], self, bool(fixed[k])) for k in range(3)]
        self.seen = False          # (the reference's `objectCoordinate.iterator().hasNext()`: seen in at least one image)

    def getX(self): return self.p[0]
    def getY(self): return self.p[1]
    def getZ(self): return self.p[2]
    def iterator(self): return It(self.seen)


class ImageCoordinate(list):
    def __init__(self, point, sx, sy):
        super().__init__([OP(sx * sx), OP(sy * sy)])
        self.point = point
        point.seen = True

    def getX(self): return self[0]
    def getY(self): return self[1]
    def getObjectCoordinate(self): return self.point


class Image(list):
    def __init__(self, coords, eo_fixed):
        super().__init__(coords)
        names = ['CAMERA_COORDINATE_X', 'CAMERA_COORDINATE_Y', 'CAMERA_COORDINATE_Z', 'CAMERA_OMEGA', 'CAMERA_PHI', 'CAMERA_KAPPA']
        self.eo = EO([UP(n, None, bool(f)) for n, f in zip(names, eo_fixed)])

    def getExteriorOrientation(self): return self.eo


class EO(list):
    def get(self, name): return next(p for p in self if p.ptype == name)


class Camera(list):
    def __init__(self, images, io_fixed, coef_fixed):
        super().__init__(images)
        self.io = [UP(n, None, bool(f)) for n, f in zip(('PRINCIPAL_POINT_X', 'PRINCIPAL_POINT_Y', 'PRINCIPAL_DISTANCE'), io_fixed)]
        self.models = [[UP('COEF', None, bool(f)) for f in coef_fixed]]

    def getInteriorOrientation(self): return self.io
    def getDistortionModels(self): return self.models


class ScaleBar(list):
    def __init__(self, a, b, sigma):
        super().__init__([OP(sigma * sigma)])
        self.a, self.b = a, b

    def getLength(self): return self[0]
    def getObjectCoordinateA(self): return self.a
    def getObjectCoordinateB(self): return self.b


class ObsParam:
    """ObservationParameter of a directly observed group: reference = the observed UnknownParameter."""
    def __init__(self, ref, variance, value=0.0): self.ref, self.variance, self.value, self.row = ref, float(variance), float(value), -1
    def getReference(self): return self.ref
    def getParameterType(self): return self.ref.getParameterType()
    def getVariance(self): return self.variance
    def getValue(self): return self.value
    def setValue(self, v): self.value = v
    def setRow(self, r): self.row = r


def group_targets(scene, P, cameras):
    """UnknownParameter objects addressed by the scene's (kind, index, comp) references."""
    images = [im for c in cameras for im in c]

    def target(kind, index, comp):
        if kind == 'point':
            return P[index].p[comp]
        if kind == 'io':
            return cameras[index].io[comp]
        if kind == 'coef':
            return cameras[index].models[0][comp]
        return images[index].eo[comp]
    return target


def variances_of(group):
    r = len(group['refs'])
    if group.get('dispersion') is not None:
        ap = np.asarray(group['dispersion'], float)
        return [ap[k + k * (k + 1) // 2] for k in range(r)]      # DirectlyObservedParameterGroup.java:55-57
    return list(np.asarray(group['var'], float))


class RankDefect:
    """defect/RankDefect.java:25-130."""
    def __init__(self): self.reset()
    def reset(self): self.f = {k: 'NOT_SET' for k in ('tx', 'ty', 'tz', 'rx', 'ry', 'rz', 'm')}
    def _set(self, k, v): self.f[k] = 'FIXED' if v == 'FIXED' else 'FREE'
    def setTranslationX(self, v): self._set('tx', v)
    def setTranslationY(self, v): self._set('ty', v)
    def setTranslationZ(self, v): self._set('tz', v)
    def setRotationX(self, v): self._set('rx', v)
    def setRotationY(self, v): self._set('ry', v)
    def setRotationZ(self, v): self._set('rz', v)
    def setScale(self, v): self._set('m', v)
    def estimateTranslationX(self): return self.f['tx'] == 'FREE'
    def estimateTranslationY(self): return self.f['ty'] == 'FREE'
    def estimateTranslationZ(self): return self.f['tz'] == 'FREE'
    def estimateRotationX(self): return self.f['rx'] == 'FREE'
    def estimateRotationY(self): return self.f['ry'] == 'FREE'
    def estimateRotationZ(self): return self.f['rz'] == 'FREE'
    def estimateScale(self): return self.f['m'] == 'FREE'
    def flags(self): return [self.f[k] == 'FREE' for k in ('tx', 'ty', 'tz', 'rx', 'ry', 'rz', 'm')]
    def getDefect(self): return sum(self.flags())


class Adjustment:
    def __init__(self):
        self.cameras, self.scaleBars, self.observedParameterGroups = [], OrderedSet(), []
        self.objectCoordinates, self.unknownParameters, self.observationGroups = OrderedSet(), OrderedSet(), OrderedSet()
        self.numberOfObservations = self.numberOfUnknownParameters = 0
        self.numberOfInteriorOrientations = self.numberOfDistortionParameters = 0
        self.sigma2apriori = 1.0                      # BundleAdjustment.java:98
        self.rankDefect = RankDefect()


def build_methods():
    class JavaMath:                      # java.lang.Math members the transliterated bodies call through `math.`
        min = staticmethod(min)
        max = staticmethod(max)
    g = {'math': JavaMath, 'OrderedSet': OrderedSet}
    for name, sig, hdr in (('prepareUnknownParameters', 'private void prepareUnknownParameters(', 'def prepareUnknownParameters(self):'),
                           ('addUnknownParameter', 'private void addUnknownParameter(', 'def addUnknownParameter(self, unknownParameter):'),
                           ('addObservationGroup', 'private void addObservationGroup(', 'def addObservationGroup(self, observations):'),
                           ('detectRankDefect', 'private void detectRankDefect(', 'def detectRankDefect(self):')):
        src = tj.transliterate(pre(tj.method_body(BA, sig)), hdr)
        exec(src, g)
        setattr(Adjustment, name, g[name])


def run(scene):
    pts = scene['points']
    P = [Point(pts['fixed'][k]) for k in range(len(pts['xyz']))]
    adj = Adjustment()
    for cam in scene['cameras']:
        images = []
        for im in cam['images']:
            coords = [ImageCoordinate(P[int(o)], s[0], s[1]) for o, s in zip(im['obj'], np.asarray(im['sigma'], float).reshape(-1, 2))]
            images.append(Image(coords, im['eo_fixed']))
        adj.cameras.append(Camera(images, cam['io_fixed'], [c[3] for c in cam['coefs']]))
    for (a, b, _length, sigma) in scene.get('scale_bars', []):
        adj.scaleBars.add(ScaleBar(P[int(a)], P[int(b)], float(sigma)))
    target = group_targets(scene, P, adj.cameras)
    for grp in scene.get('observed_groups', []):
        adj.observedParameterGroups.append([ObsParam(target(*ref), v) for ref, v in zip(grp['refs'], variances_of(grp))])
    adj.prepareUnknownParameters()
    col = lambda plist: np.array([p.getColumn() for p in plist], np.int64)
    return dict(pt_col=np.array([[p.getColumn() for p in q.p] for q in P], np.int64).reshape(-1, 3),
                io_col=np.concatenate([col(c.io) for c in adj.cameras]),
                coef_col=np.concatenate([col(c.models[0]) for c in adj.cameras]) if any(c.models[0] for c in adj.cameras) else np.zeros(0, np.int64),
                eo_col=np.concatenate([col(im.eo) for c in adj.cameras for im in c]),
                group_rows=np.array([o.row for g in adj.observedParameterGroups for o in g], np.int64),
                counts=np.array([adj.numberOfObservations, adj.numberOfUnknownParameters, adj.numberOfInteriorOrientations,
                                 adj.numberOfDistortionParameters, adj.rankDefect.getDefect(), len(adj.objectCoordinates)], np.int64),
                flags=np.array(adj.rankDefect.flags(), bool), sigma2=np.array([adj.sigma2apriori]))


def scenes():
    from tests.scenes import example_scene, random_scene, synthetic_scene
    for seed in range(12):
        yield 'random%d' % seed, random_scene(seed)
    yield 'example', example_scene()
    yield 'config2_small', synthetic_scene(2, images=10, targets=60)[0]
    yield 'config4_fixed_datum', synthetic_scene(4, images=9, targets=70, free_network=False)[0]
    yield 'config3_observed_points', synthetic_scene(3, images=6, targets=40)[0]
    yield 'observed_eo_io', observed_eo_io_scene()
    yield 'observed_omega_only', observed_omega_only_scene()


def observed_eo_io_scene():
    """Free network plus directly observed exterior-orientation and interior-orientation parameters (diagonal weights):
    exercises the CAMERA_* branches of detectRankDefect and a coordinate that enters only through a group."""
    from tests.scenes import synthetic_scene
    sc, truth = synthetic_scene(2, images=7, targets=45)
    eo = truth['eo']
    sc['observed_groups'] = [
        {'refs': [('eo', 1, 0), ('eo', 1, 1), ('eo', 1, 2), ('eo', 1, 3), ('eo', 1, 5)], 'obs': eo[1][[0, 1, 2, 3, 5]] + 0.01,
         'var': np.array([0.01, 0.01, 0.01, 1e-6, 1e-6]), 'dispersion': None},
        {'refs': [('eo', 4, 0), ('eo', 4, 4), ('io', 0, 2), ('coef', 0, 4)], 'obs': np.array([eo[4][0], eo[4][4], 28.8, -2e-4]),
         'var': np.array([0.04, 4e-6, 1e-4, 1e-10]), 'dispersion': None}]
    return sc


def observed_omega_only_scene():
    """Free network with one directly observed rotation angle and one observed X0: a partial datum (d = 5)."""
    from tests.scenes import synthetic_scene
    sc, truth = synthetic_scene(2, images=7, targets=45)
    sc['observed_groups'] = [{'refs': [('eo', 2, 3), ('eo', 3, 0)], 'obs': np.array([truth['eo'][2][3], truth['eo'][3][0]]),
                              'var': np.array([1e-6, 0.01]), 'dispersion': None}]
    return sc


def main():
    build_methods()
    out = {}
    for name, sc in scenes():
        for k, v in run(sc).items():
            out['%s__%s' % (name, k)] = v
        print(name, out[name + '__counts'].tolist(), out[name + '__flags'].astype(int).tolist(), out[name + '__sigma2'][0])
    np.savez_compressed(OUT, **out)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
