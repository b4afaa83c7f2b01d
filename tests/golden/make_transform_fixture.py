"""Generates tests/golden/reference_transform.npz: the covariance propagation of transformed object coordinates produced by EXECUTING
the reference's own CoordinateTransformationExteriorOrientation.transform(...) and setPartialDerivations(...)
(tranformation/CoordinateTransformationExteriorOrientation.java:49-121, :131-283) on small networks -- the visibility loops, the order
and the names of the transformed points, the placement of the 45 partial derivatives in the columns of J, and the final product
sigma2 * J Qxx J' in MTJ's packed upper layout.  (tests/golden/make_formula_fixtures.py pins the 45 expressions one by one; this
script pins the function as a whole.)

Run in the build container only (reads /root/reference):
    python tests/golden/make_transform_fixture.py

The two method bodies are transliterated mechanically (make_jacobian_fixture.transliterate) and exec'ed on stub objects, as in the
other make_*_fixture.py scripts, with three local adaptations that are plain renamings: `imagesToAlign.entrySet()` is handed over as
the attribute `imagesToAlign.entries` (the transliterator iterates over names, not over call results), the array initialiser
`new int[] { row++, row++, row++ }` (:95-99) becomes `threeRows(row); row += 3;`, and Java's String + int concatenation of the point
name (:100) goes through stub string / id types.  Third-party code is stood in for by its definition, not by its loop order: MTJ's
`CoVar.transBmult(J, CJT)` (CJT = CoVar J') and `J.mult(sigma2, CJT, covariance)` (covariance = sigma2 J CJT, written into an
UpperSymmPackMatrix: the upper triangle is kept) are numpy products -- the fixture is compared at 1e-12, not bit for bit.
`LinkedSparseMatrix.set` checks its indices as MTJ does, so a FIXED parameter (column Integer.MAX_VALUE) makes the reference throw:
recorded as `fixed_parameter_throws` (the library's "fixed parameters contribute nothing" goes beyond the reference there).
Numbers and names only are stored.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_jacobian_fixture as tj  # noqa: E402
import make_lm_fixture as tl  # noqa: E402

SRC = '/root/reference/JAICOV/src/org/applied_geodesy/adjustment/bundle/tranformation/CoordinateTransformationExteriorOrientation.java'
OUT = os.path.join(HERE, 'reference_transform.npz')
MAXV = 2147483647
EO_TYPES = ('CAMERA_COORDINATE_X', 'CAMERA_COORDINATE_Y', 'CAMERA_COORDINATE_Z', 'CAMERA_OMEGA', 'CAMERA_PHI', 'CAMERA_KAPPA')


# ---- stubs for the Java object graph and for MTJ ----------------------------------------------------------------------------------
class JStr(str):
    def isBlank(self): return not self.strip()
    def __add__(self, other): return JStr(str.__add__(self, str(other)))
    def __radd__(self, other): return JStr(str(other) + str(self))


class JId(int):
    def __radd__(self, other):              # "text" + id
        return JStr(str(other) + str(int(self)))


class Param:
    def __init__(self, value, column=-1): self.value, self.column = float(value), int(column)
    def getValue(self): return self.value
    def getColumn(self): return self.column
    def setColumn(self, c): self.column = int(c)


class ObjectCoordinate:
    def __init__(self, name, x, y, z):
        self.name, self.p = JStr(name), [Param(x), Param(y), Param(z)]
    def getName(self): return self.name
    def getX(self): return self.p[0]
    def getY(self): return self.p[1]
    def getZ(self): return self.p[2]


class Exterior:
    def __init__(self, values, columns): self.d = {t: Param(v, c) for t, v, c in zip(EO_TYPES, values, columns)}
    def get(self, t): return self.d[t]


class Image:
    def __init__(self, ident, eo, visible): self.ident, self.eo, self.visible = JId(ident), eo, visible
    def getId(self): return self.ident
    def getExteriorOrientation(self): return self.eo
    def get(self, oc): return oc if id(oc) in self.visible else None       # camera/Image.java:77-79: the image coordinate or null


class Entry:
    def __init__(self, k, v): self.k, self.v = k, v
    def getKey(self): return self.k
    def getValue(self): return self.v


class ImageMap:
    def __init__(self, pairs): self.entries = [Entry(k, v) for k, v in pairs]


class JList(list):
    def __init__(self, capacity=0): super().__init__()
    @classmethod
    def of(cls, items):
        out = cls()
        out.extend(items)
        return out
    def add(self, x): self.append(x)
    def size(self): return len(self)


class LinkedSparseMatrix:
    def __init__(self, rows, columns): self.a = np.zeros((int(rows), int(columns)))
    def numRows(self): return self.a.shape[0]
    def numColumns(self): return self.a.shape[1]
    def set(self, r, c, v):
        if not (0 <= r < self.a.shape[0] and 0 <= c < self.a.shape[1]):      # AbstractMatrix.check -> IndexOutOfBoundsException
            raise IndexError('row %d column %d' % (r, c))
        self.a[r, c] = v
    def mult(self, alpha, B, C): C.assign(alpha * (self.a @ B.a))


class UpperSymmPackMatrix:
    def __init__(self, n, full=None): self.n, self.full = int(n), (np.zeros((int(n), int(n))) if full is None else full)
    def numRows(self): return self.n
    def numColumns(self): return self.n
    def transBmult(self, B, C): C.a[:, :] = self.full @ B.a.T             # C = A B'
    def assign(self, M): self.full = np.triu(M) + np.triu(M, 1).T        # writes below the diagonal are ignored: the upper triangle rules
    def packed(self): return np.concatenate([self.full[:c + 1, c] for c in range(self.n)])


class Holder:
    covariance = None
    transformedCoordinates = None


def build():
    g = {'math': tl.JavaMath, 'LinkedSparseMatrix': LinkedSparseMatrix, 'UpperSymmPackMatrix': UpperSymmPackMatrix, 'JList': JList,
         'ObjectCoordinate': ObjectCoordinate, 'threeRows': lambda r: [r, r + 1, r + 2]}
    body = tj.method_body(SRC, 'private static ObjectCoordinate setPartialDerivations(')
    src = tj.transliterate(tl.ternaries(body), 'def setPartialDerivations(name, rows, J, objectCoordinateSrc, exteriorOrientationTrg, exteriorOrientationSrc):')
    exec(src, g)
    body = tj.method_body(SRC, 'public void transform(')
    joined, i = [], 0
    while i < len(body):                     # int rowsInJacobian[] = new int[] { row++, row++, row++, };
        if 'new int[]' in body[i]:
            j = i
            while '};' not in body[j]:
                j += 1
            text = ' '.join(l.strip() for l in body[i:j + 1])
            assert re.sub(r'\s+', '', text) == 'introwsInJacobian[]=newint[]{row++,row++,row++,};', text
            joined += ['Rows rowsInJacobian = threeRows(row);', 'row += 3;']
            i = j + 1
            continue
        joined.append(body[i].replace('imagesToAlign.entrySet()', 'imagesToAlign.entries'))
        i += 1
    exec(tj.transliterate(joined, 'def transform(self, objectCoordinatesToTransform, imagesToAlign, sigma2, CoVar):'), g)
    return g


def network(seed, images, targets, fixed_kappa=False):
    """A small network through the product's own bookkeeping (columns as prepareUnknownParameters assigns them -- pinned against the
    executed reference elsewhere), with visibility gaps; returns the flat arrays both sides are built from."""
    from bundle_adjustment_b200.workloads import flat_problem, synthetic_scene
    scene = synthetic_scene(2, images=images, targets=targets, seed=seed)[0]
    rng = np.random.default_rng(seed)
    for k, im in enumerate(scene['cameras'][0]['images']):
        keep = rng.random(len(im['obj'])) > (0.35 if k % 2 else 0.1)
        keep[:6] = True
        for key in ('obj', 'xy', 'sigma', 'rho'):
            if im.get(key) is not None:
                im[key] = np.asarray(im[key])[keep]
    if fixed_kappa:
        scene['cameras'][0]['images'][1]['eo_fixed'][5] = True
    adj, flat = flat_problem(scene)
    return scene, flat


def run(g, flat, point_ids, align, sigma2, seed):
    n = int(flat['n_unknowns']) + int(np.sum(flat['free_flags']))
    rng = np.random.default_rng(seed + 1000)
    G = rng.standard_normal((n, n))
    Q = (G @ G.T) / n * 1e-2
    xyz, pt_col = np.asarray(flat['xyz']).reshape(-1, 3), np.asarray(flat['pt_col']).reshape(-1, 3)
    eo_val, eo_col = np.asarray(flat['eo_val']).reshape(-1, 6), np.asarray(flat['eo_col']).reshape(-1, 6)
    pt_ptr, obj_idx = np.asarray(flat['pt_ptr']), np.asarray(flat['obj_idx'])
    pts = {}
    for p in range(xyz.shape[0]):
        oc = ObjectCoordinate(str(p), *xyz[p])          # the default names of the generator's points
        for k in range(3):
            oc.p[k].setColumn(pt_col[p, k])
        pts[p] = oc
    imgs = []
    for i in range(eo_val.shape[0]):
        vis = {id(pts[int(p)]) for p in obj_idx[pt_ptr[i]:pt_ptr[i + 1]]}
        imgs.append(Image(i + 1, Exterior(eo_val[i], eo_col[i]), vis))
    self = Holder()
    CoVar = UpperSymmPackMatrix(n, Q)
    g['transform'](self, JList.of(pts[p] for p in point_ids), ImageMap([(imgs[r], JList.of(imgs[i] for i in lst)) for r, lst in align]), sigma2, CoVar)
    names = [str(t.getName()) for t in self.transformedCoordinates]
    out_xyz = np.array([[t.getX().getValue(), t.getY().getValue(), t.getZ().getValue()] for t in self.transformedCoordinates])
    out_col = np.array([[t.getX().getColumn(), t.getY().getColumn(), t.getZ().getColumn()] for t in self.transformedCoordinates])
    packedQ = np.concatenate([Q[:c + 1, c] for c in range(n)])
    return dict(names=np.array(names), xyz=out_xyz, columns=out_col, covariance=self.covariance.packed(), qxx_packed=packedQ,
                in_xyz=xyz, in_pt_col=pt_col, in_eo_val=eo_val, in_eo_col=eo_col)      # the inputs, so that a test can check it rebuilt the same network


CASES = {   # name: (seed, images, targets, points to transform, {reference image: [images]}, sigma2)
    'two_reference_images': (31, 6, 14, [0, 2, 3, 5, 7, 8, 11, 13], [(0, [0, 1, 2]), (3, [4, 5])], 1.7e-7),
    'reference_image_among_its_images_last': (32, 5, 10, [1, 2, 4, 6, 9], [(2, [0, 1, 2])], 1.0),
    'single_pair': (33, 4, 8, [0, 1, 2, 3, 4, 5, 6, 7], [(1, [3])], 2.5e-7),
}


def main():
    g = build()
    out = {}
    for name, (seed, images, targets, point_ids, align, sigma2) in CASES.items():
        _scene, flat = network(seed, images, targets)
        r = run(g, flat, point_ids, align, sigma2, seed)
        assert len(r['names']) >= 5, name
        for k, v in r.items():
            out['%s__%s' % (name, k)] = v
        out['%s__sigma2' % name] = np.float64(sigma2)
        print('%-40s %3d transformed points, covariance %d entries' % (name, len(r['names']), r['covariance'].size))
    # a FIXED exterior-orientation parameter: MTJ's index check makes the reference throw (J.set(row, Integer.MAX_VALUE, ...))
    _scene, flat = network(34, 4, 8, fixed_kappa=True)
    try:
        run(g, flat, [0, 1, 2], [(0, [1])], 1.0, 34)
        threw = False
    except IndexError:
        threw = True
    out['fixed_parameter_throws'] = np.bool_(threw)
    print('fixed parameter -> reference throws:', threw)
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
