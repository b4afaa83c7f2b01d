"""Generates tests/golden/reference_writers.npz: what the reference's two result writers hand to their output libraries, produced by
EXECUTING their own method bodies -- DefaultResultWriter.exportCovarianceInformation / exportCovarianceMatrix
(util/io/writer/DefaultResultWriter.java:67-155) and MatlabResultWriter.export (util/io/writer/MatlabResultWriter.java:52-226) -- on a
small two-camera network with fixed point components, a fixed interior-orientation parameter and fixed distortion coefficients, once with a
cofactor matrix and once without (MatrixInversion.NONE).

Run in the build container only (reads /root/reference):
    python tests/golden/make_writer_fixture.py

Mechanical transliteration as in the other make_*_fixture.py scripts (make_jacobian_fixture.transliterate), with plain renamings where
the transliterator has no rule: `x = counter++;` becomes `x = counter; counter++;`, `(p instanceof PolynomialCoefficient)` becomes
`isPolynomialCoefficient(p)`, and the declared type is dropped from the one declaration that carries a ternary.  The output libraries
are recorders: `PrintWriter.printf(Locale.ENGLISH, format, args...)` stores the format string and the arguments (java.util.Formatter's
number format is a property of the JDK, tested by known answers in tests/test_host_cpp.py), MFL's `Mat5` / `Struct` / `Matrix` store
names, indices and values.  Numbers and names only are stored.
"""
import json
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

W = '/root/reference/JAICOV/src/org/applied_geodesy/util/io/writer/'
OUT = os.path.join(HERE, 'reference_writers.npz')


# ---- recorders for java.io and MFL (us.hebi.matlab.mat) ----------------------------------------------------------------------------
class JStr(str):
    def toLowerCase(self): return JStr(self.lower())
    def __add__(self, other): return JStr(str.__add__(self, str(other)))


class PrintWriter:
    def __init__(self, sink): self.calls, self.closed = [], False
    def printf(self, locale, fmt, *args): self.calls.append((str(fmt), list(args)))
    def println(self): self.calls.append(('%n', []))
    def close(self): self.closed = True


class Struct:
    def __init__(self, rows, cols): self.shape, self.fields = (rows, cols), {}
    def set(self, field, index, value): self.fields.setdefault(str(field), {})[int(index)] = value


class Matrix:
    def __init__(self, rows, cols, kind): self.kind, self.a = kind, np.zeros((rows, cols))
    def setDouble(self, r, c, v): self.a[r, c] = v
    def setInt(self, r, c, v): self.a[r, c] = int(v)
    def setLong(self, r, c, v): self.a[r, c] = int(v)


class MatFile:
    def __init__(self): self.arrays = []
    def addArray(self, name, value): self.arrays.append((str(name), value))


class Mat5:
    written = []
    newStruct = staticmethod(lambda r, c: Struct(r, c))
    newString = staticmethod(lambda s: ('string', str(s)))
    newMatrix = staticmethod(lambda r, c, kind: Matrix(r, c, kind))
    newMatFile = staticmethod(MatFile)
    writeToFile = staticmethod(lambda matFile, f: Mat5.written.append((str(f), matFile)))


class MatlabType:
    Double, Int32, Int64 = 'double', 'int32', 'int64'


class JList(list):
    def __init__(self, capacity=0): super().__init__()
    def add(self, x): self.append(x)
    def get(self, i): return self[i]
    def size(self): return len(self)


# ---- the object graph, from the product's host mirror after prepareUnknownParameters ------------------------------------------------------
class PType:
    def __init__(self, name): self._n = JStr(name)
    def name(self): return self._n


class P:
    def __init__(self, p, poly):
        self.value, self.column, self.ptype, self.poly = float(p.getValue()), int(p.getColumn()), PType(p.getParameterType().name), poly
        self.order = int(p.getOrder()) if poly else None
    def getValue(self): return self.value
    def getColumn(self): return self.column
    def getParameterType(self): return self.ptype
    def getOrder(self): return self.order


class Group(list):
    def getNumberOfParameters(self): return len(self)


class Cam:
    def __init__(self, cam, PolynomialCoefficient):
        self.ident = cam.getId()
        self.io = Group(P(p, False) for p in cam.getInteriorOrientation())
        self.models = [Group(P(p, isinstance(p, PolynomialCoefficient)) for p in m) for m in cam.getDistortionModels()]
    def getId(self): return self.ident
    def getInteriorOrientation(self): return self.io
    def getDistortionModels(self): return self.models


class OC:
    def __init__(self, oc): self.name, self.p = JStr(oc.getName()), [P(oc.getX(), False), P(oc.getY(), False), P(oc.getZ(), False)]
    def getName(self): return self.name
    def getX(self): return self.p[0]
    def getY(self): return self.p[1]
    def getZ(self): return self.p[2]


class Cofactor:
    def __init__(self, Q): self.Q = Q
    def numRows(self): return self.Q.shape[0]
    def numColumns(self): return self.Q.shape[1]
    def get(self, r, c): return float(self.Q[r, c])


class BA:
    def __init__(self, adj, Q, PolynomialCoefficient):
        self.pts = JList(); self.pts.extend(OC(o) for o in adj.getObjectCoordinates())
        self.cams = JList(); self.cams.extend(Cam(c, PolynomialCoefficient) for c in adj.getCameras())
        self.Q, self.adj = (Cofactor(Q) if Q is not None else None), adj
    def getObjectCoordinates(self): return self.pts
    def getCameras(self): return self.cams
    def getCofactorMatrix(self): return self.Q
    def getNumberOfObservations(self): return self.adj.getNumberOfObservations()
    def getNumberOfDatumConditions(self): return self.adj.getNumberOfDatumConditions()
    def getNumberOfUnknownParameters(self): return self.adj.getNumberOfUnknownParameters()
    def getDegreeOfFreedom(self): return self.adj.getDegreeOfFreedom()
    def getVarianceFactorApriori(self): return self.adj.getVarianceFactorApriori()
    def getVarianceFactorAposteriori(self): return 1.25 * self.adj.getVarianceFactorApriori()


class Writer:
    def __init__(self, base): self.base = base
    def getExportPathAndFileBaseName(self): return JStr(self.base)


def pre(lines):
    out = []
    for l in lines:
        m = re.match(r'^(\s*)(\w+) = (\w+)\+\+;\s*$', l)                       # x = counter++;
        if m:
            out += ['%s%s = %s;' % m.groups(), '%s%s++;' % (m.group(1), m.group(3))]
            continue
        l = l.replace('(unknownParameter instanceof PolynomialCoefficient)', 'isPolynomialCoefficient(unknownParameter)')
        l = re.sub(r'^(\s*)List<Integer> (indices = exportDispersionMatrix \?)', r'\1\2', l)
        out.append(l)
    return out


def build():
    g = {'math': tl.JavaMath, 'PrintWriter': PrintWriter, 'BufferedWriter': lambda x: x, 'FileWriter': lambda f: f, 'File': lambda p: JStr(p),
         'Locale': type('Locale', (), {'ENGLISH': 'ENGLISH'}), 'JList': JList, 'Mat5': Mat5, 'MatlabType': MatlabType,
         'isPolynomialCoefficient': lambda p: p.poly,
         'newInteger': lambda v: ('int32', int(v)), 'newLong': lambda v: ('int64', int(v)), 'newDouble': lambda v: ('double', float(v))}
    D = W + 'DefaultResultWriter.java'
    exec(tj.transliterate(pre(tj.method_body(D, 'private List<Integer> exportCovarianceInformation(')),
                          'def exportCovarianceInformation(self, bundleAdjustment, file):'), g)
    exec(tj.transliterate(pre(tj.method_body(D, 'private void exportCovarianceMatrix(')),
                          'def exportCovarianceMatrix(self, bundleAdjustment, indices, file):'), g)
    exec(tj.transliterate(tl.ternaries(pre(tj.method_body(W + 'MatlabResultWriter.java', 'public void export('))),
                          'def exportMatlab(self, bundleAdjustment):'), g)
    return g


def network():
    """Two cameras, 4 images, 14 points; fixed: two point components, a whole point, x0 of the second camera, A3 / Cx of the first."""
    from bundle_adjustment_b200.host import PolynomialCoefficient
    from bundle_adjustment_b200.workloads import build_adjustment, synthetic_scene
    scene = synthetic_scene(2, images=4, targets=14, seed=77, n_cameras=2)[0]
    fixed = np.array(scene['points']['fixed'], bool).reshape(-1, 3)
    fixed[2, 1] = fixed[5, 2] = True
    fixed[9, :] = True
    scene['points']['fixed'] = fixed
    scene['cameras'][1]['io_fixed'] = [True, False, False]
    c0 = scene['cameras'][0]['coefs']
    scene['cameras'][0]['coefs'] = [(t, o, v, f or (t == 121 and o == 3) or t == 141) for (t, o, v, f) in c0]
    adj, _pts = build_adjustment(scene)
    adj._prepare()
    return adj, PolynomialCoefficient


def main():
    g = build()
    adj, PolynomialCoefficient = network()
    n = adj.getNumberOfUnknownParameters() + adj.getNumberOfDatumConditions()
    G = np.random.default_rng(77).standard_normal((n, n))
    Q = G @ G.T * 1e-3
    out = {'Q': Q}
    for tag, cof in (('full', Q), ('none', None)):
        ba = BA(adj, cof, PolynomialCoefficient)
        w = Writer('/tmp/out')
        # DefaultResultWriter.export: the two calls of :57-58, with the PrintWriters recorded
        made = []
        g['PrintWriter'] = lambda sink, made=made: made.append(PrintWriter(sink)) or made[-1]
        indices = g['exportCovarianceInformation'](w, ba, g['File']('/tmp/out.info'))
        g['exportCovarianceMatrix'](w, ba, indices, g['File']('/tmp/out.cxx'))
        assert all(p.closed for p in made)
        info = made[0].calls
        out['%s__default_indices' % tag] = np.array(list(indices), np.int64)
        out['%s__info_format' % tag] = np.array(sorted({c[0] for c in info}))
        out['%s__info_names' % tag] = np.array([str(c[1][0]) for c in info])
        out['%s__info_comp' % tag] = np.array([str(c[1][1]) for c in info])
        out['%s__info_value' % tag] = np.array([c[1][2] for c in info], float)
        out['%s__info_index' % tag] = np.array([c[1][3] for c in info], np.int64)
        out['%s__cxx_written' % tag] = np.bool_(len(made) == 2)
        if len(made) == 2:
            cx = made[1].calls
            out['%s__cxx_format' % tag] = np.array(sorted({c[0] for c in cx}))
            rows, cur = [], []
            for fmt, args in cx:
                if fmt == '%n':
                    rows.append(cur); cur = []
                else:
                    cur.append(args[0])
            assert not cur
            out['%s__cxx_values' % tag] = np.array(rows, float)
        # MatlabResultWriter.export
        Mat5.written.clear()
        g['exportMatlab'](w, ba)
        (path, mf), = Mat5.written
        out['%s__mat_path' % tag] = np.array(path)
        order = [name for name, _ in mf.arrays]
        out['%s__mat_variables' % tag] = np.array(order)
        desc = {}
        for name, v in mf.arrays:
            if isinstance(v, Struct):
                desc[name] = {'shape': list(v.shape), 'fields': {f: [list(v.fields[f][i]) if isinstance(v.fields[f][i], tuple) else v.fields[f][i]
                                                                     for i in sorted(v.fields[f])] for f in v.fields}}
            elif isinstance(v, Matrix):
                out['%s__mat_%s' % (tag, name)] = v.a
                desc[name] = {'matrix': v.kind, 'shape': list(v.a.shape)}
            else:
                desc[name] = list(v)
        out['%s__mat_json' % tag] = np.array(json.dumps(desc))
        print(tag, 'default indices', len(indices), '| .info lines', len(info), '| .cxx', len(made) == 2, '| mat variables', order)
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
