"""Generates tests/golden/reference_jacobian.npz: golden vectors of the reference's per-image-point Jacobian rows and
misclosures (collinearity equations + every distortion model incl. the Zernike models), produced by EXECUTING the
reference's own formulas.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_jacobian_fixture.py

How: the bodies of
    PartialDerivativeFactory.CollinearityEquationFactory(...)            derivation/PartialDerivativeFactory.java:96-193
    DistortionModelFactory.apply(...)                                    derivation/DistortionModelFactory.java:33-101
    {AffinityShear,RadialDistance,RadiallySymmetric,Tangential,Zernike}DistortionModelFactory.apply(...)
are straight-line arithmetic inside simple for / if / else blocks.  `transliterate` turns such a body into Python text
mechanically (strip type names and casts, braces -> indentation, && -> and, Math. -> math., long/long -> //) and the
result is exec'ed against small stub objects that stand in for the Java object graph.  No formula is restated by hand
and no reference text is stored: the fixture holds inputs and outputs only.  What IS restated (structure, not formulas):
the base entries A[., slot] = par_xs_* / par_ys_* and w = observed - (x, y) (PartialDerivativeFactory.java:321-417), the
order in which the models are applied (:420-444), and the Zernike radial polynomial table (parameter/ZernikeCoefficient.java:40-56).
tests/test_reference_formulas.py checks oracle/jaicov_oracle.c (through oracle.oracle.eval_point) against these vectors.
"""
import math
import os
import re

import numpy as np

REF = '/root/reference/JAICOV/src/org/applied_geodesy/adjustment/bundle'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_jacobian.npz')
MAXV = 2147483647

TYPE = r'(?:final\s+)?(?:double|long|int|boolean|[A-Z]\w*(?:\.\w+)?(?:<[^;=()]*>)?)'
RE_DECL = re.compile(r'^' + TYPE + r'\s+(\w+)\s*=\s*(.*)$')


def expr(e):
    e = re.sub(r'\((?:PolynomialCoefficient<\?>|ZernikeCoefficient)\)\s*', '', e)            # casts
    e = e.replace('&&', ' and ').replace('||', ' or ').replace('Integer.MAX_VALUE', str(MAXV)).replace('Math.', 'math.')
    e = re.sub(r'\bParameterType\.(\w+)', r"'\1'", e)
    e = re.sub(r'\bType\.(\w+)', r"'\1'", e)
    e = re.sub(r'\bpj/2\b', '(pj//2)', e)                                                       # long / int
    e = re.sub(r'\bthis\.', 'self.', e)
    e = re.sub(r'\bnew\s+LinkedHashSet<[^>]*>\(\)', 'OrderedSet()', e)
    e = re.sub(r'\bnew\s+HashSet<[^>]*>\(\)', 'JSet()', e)
    e = re.sub(r'\bnew\s+ArrayList<(?:[^<>]|<[^<>]*>)*>\(', 'JList(', e)
    e = re.sub(r'\bnew\s+', '', e)
    e = re.sub(r'\b(\w+)\.printStackTrace\(\)', r'printStackTrace(\1)', e)
    e = re.sub(r'\b(\w+)\.length\b(?!\()', r'len(\1)', e)
    e = re.sub(r'\bnull\b', 'None', e)
    e = re.sub(r'\bBoolean\.TRUE\b', 'True', e)
    e = re.sub(r'\bBoolean\.FALSE\b', 'False', e)
    e = re.sub(r'\bDouble\.isInfinite\(', 'math.isinf(', e)
    e = re.sub(r'\bDouble\.isNaN\(', 'math.isnan(', e)
    e = re.sub(r'\bMatrixInversion\.(\w+)', r"'\1'", e)
    e = re.sub(r'\bEstimationType\.(\w+)', r"'\1'", e)
    e = re.sub(r'\btrue\b', 'True', e)
    e = re.sub(r'\bfalse\b', 'False', e)
    e = re.sub(r'\bCollections\.sort\((\w+)\)', r'\1.sort()', e)
    e = re.sub(r'\bDefectType\.(\w+)', r"'\1'", e)
    e = re.sub(r"\((\w+)\s*\?\s*('\w+')\s*:\s*('\w+')\)", r'(\2 if \1 else \3)', e)                # f(cond ? 'A' : 'B')
    e = re.sub(r'\((?:ImageCoordinate|ScaleBar|ObjectCoordinate|UpperSPDPackMatrix)\)', '', e)
    e = re.sub(r'\(double\)\s*(\w+)', r'float(\1)', e)
    e = e.replace('!', ' not ').replace(' not =', '!=')
    return e


def clean(lines):
    """Comments out, blank lines out, multi-line `if (` conditions joined."""
    out, buf = [], None
    for raw in lines:
        line = re.sub(r'/\*.*?\*/', '', raw)
        line = re.sub(r'//.*$', '', line).strip()
        if not line:
            continue
        if buf is not None:
            buf += ' ' + line
            if buf.count('(') == buf.count(')'):
                out.append(buf)
                buf = None
            continue
        if re.match(r'^(else\s+)?if\s*\(', line) and line.count('(') != line.count(')'):
            buf = line
            continue
        m = re.match(r'^\}\s*while\s*\((.*)\);$', line)                   # do { ... } while (cond);
        if m:
            out += ['__dowhile__(%s);' % m.group(1), '}']
            continue
        m = re.match(r'^while\s*\((.*)\);$', line)                          # the same with the `while` on its own line
        if m and out and out[-1] == '}':
            out[-1:] = ['__dowhile__(%s);' % m.group(1), '}']
            continue
        m = re.match(r'^\}\s*(catch\s*\(.*\)\s*\{|finally\s*\{)$', line)     # } catch (...) {
        if m:
            out += ['}', m.group(1)]
            continue
        out.append(line)
    return out


JAVA_EXCEPTIONS = {'MatrixSingularException': 'MatrixSingularException', 'MatrixNotSPDException': 'MatrixNotSPDException',
                   'IllegalArgumentException': 'ValueError', 'ArrayIndexOutOfBoundsException': 'IndexError', 'Exception': 'Exception',
                   'OutOfMemoryError': 'MemoryError', 'NullPointerException': 'AttributeError', 'IOException': 'OSError',
                   'UnsupportedOperationException': 'ValueError'}


def deswitch(lines):
    """`switch (x) { case A: case B: ...; break; default: ...; }` over ParameterType constants -> if / else-if chain
    (the reference uses no fall-through with statements in between, only grouped labels)."""
    out, i = [], 0
    while i < len(lines):
        m = re.match(r'^switch\s*\((.*)\)\s*\{$', lines[i])
        if not m:
            out.append(lines[i])
            i += 1
            continue
        subject, groups, labels, body, depth = m.group(1).strip(), [], [], [], 1
        i += 1
        while depth > 0:
            l = lines[i]
            i += 1
            if l == '}' and depth == 1:
                depth = 0
                break
            c = re.match(r'^case\s+(\w+)\s*:$', l)
            if depth == 1 and c:
                labels.append(c.group(1))
                continue
            if depth == 1 and l == 'default:':
                labels.append(None)
                continue
            if depth == 1 and l == 'break;':
                groups.append((labels, body))
                labels, body = [], []
                continue
            depth += l.count('{') - l.count('}')
            body.append(l)
        if labels or body:
            groups.append((labels, body))
        first = True
        for labs, stmts in groups:
            stmts = deswitch(stmts)
            if None in labs:
                if not stmts:
                    continue
                out.append('else {' if not first else 'if (true) {')
            else:
                cond = ' || '.join('%s == ParameterType.%s' % (subject, lab) for lab in labs)
                out.append(('if (%s) {' if first else 'else if (%s) {') % cond)
            out += stmts + ['}']
            first = False
    return out


RE_HEAD = re.compile(r'^(?:(?:else\s+)?if\s*\(.*\)|for\s*\(.*\)|else)$')


def bracify(lines):
    """Gives every brace-less if / else / for body its braces (recursively), so that blocks are always explicit."""
    out = []

    def statement(i):
        line = lines[i]
        if RE_HEAD.match(line):
            out.append(line + ' {')
            i = statement(i + 1)
            out.append('}')
            return i
        out.append(line)
        if line.endswith('{'):
            i += 1
            while lines[i] != '}':
                i = statement(i)
            out.append('}')
        return i + 1

    i = 0
    while i < len(lines):
        i = statement(i)
    return out


def transliterate(lines, header):
    """Java block (list of source lines, without the enclosing braces) -> Python function source."""
    out, ind = [header], 1

    def emit(text):
        out.append('    ' * ind + text)

    for line in bracify(deswitch(clean(lines))):
        if line == '}':
            ind -= 1
            continue
        m = re.match(r'^(else\s+)?if\s*\((.*)\)\s*\{$', line)
        if m:
            emit(('elif ' if m.group(1) else 'if ') + expr(m.group(2)) + ':')
            ind += 1
            continue
        if line == 'else {':
            emit('else:')
            ind += 1
            continue
        if line == 'do {':
            emit('while True:')
            ind += 1
            continue
        if line == 'try {':
            emit('try:')
            ind += 1
            continue
        if line == 'finally {':
            emit('finally:')
            ind += 1
            continue
        m = re.match(r'^catch\s*\(([\w\s\|]+?)\s+(\w+)\)\s*\{$', line)
        if m:
            names = [JAVA_EXCEPTIONS[x.strip()] for x in m.group(1).split('|')]
            emit('except (%s,) as %s:' % (', '.join(names), m.group(2)))
            ind += 1
            continue
        m = re.match(r'^__dowhile__\((.*)\);$', line)
        if m:
            emit('if not (%s): break' % expr(m.group(1)))
            continue
        m = re.match(r'^for\s*\(.*\s(\w+)\s*:\s*([\w\.]+)\)\s*\{$', line)
        if m:
            emit('for %s in %s:' % (m.group(1), expr(m.group(2))))
            ind += 1
            continue
        m = re.match(r'^for\s*\(int\s+(\w+)\s*=\s*([\w\s\+\-]+?);\s*\1\s*<\s*([\w\.\(\)]+);\s*\1\+\+\)\s*\{$', line)
        if m:
            emit('for %s in range(%s, %s):' % (m.group(1), m.group(2), expr(m.group(3))))
            ind += 1
            continue
        assert line.endswith(';'), line
        line = line[:-1].strip()
        m = re.match(r'^int\s+(\w+)\s*=\s*(this\.\w+\.\w+\(\))\s*\?\s*(\w+)\+\+\s*:\s*-1$', line)     # x = cond ? counter++ : -1
        if m:
            emit('%s = -1' % m.group(1))
            emit('if %s: %s = %s; %s += 1' % (expr(m.group(2)), m.group(1), m.group(3), m.group(3)))
            continue
        m = re.match(r'^double\s+(\w+)\[\]\s*=\s*new\s+double\[(\w+)\]$', line)
        if m:
            emit('%s = [0.0] * %s' % (m.group(1), m.group(2)))
            continue
        m = re.match(r'^([\w\.]+)\+\+$', line)
        if m:
            emit('%s += 1' % expr(m.group(1)))
            continue
        m = re.match(r'^(.*[\(,]\s*)([\w\.]+)\+\+(\s*[\),].*)$', line)                               # f( counter++ ), f(row, row++, x)
        if m:
            emit(expr(m.group(1) + m.group(2) + m.group(3)))
            emit('%s += 1' % expr(m.group(2)))
            continue
        if line.startswith('throw new '):
            emit('raise ValueError()')
            continue
        m = re.match(r'^(?:double|int|boolean)\s+(\w+\s*=\s*[^,]+(?:,\s*\w+\s*=\s*[^,]+)+)$', line)   # double a = 0, b = 0, c = 0
        if m:
            for part in m.group(1).split(','):
                emit(expr(part.strip()))
            continue
        m = re.match(r'^(\s*[\w\.]+\s*\+=\s*)(.+?)\s*\?\s*1\s*:\s*0$', line)                        # x += cond ? 1 : 0
        if m:
            emit(expr('%s(1 if %s else 0)' % (m.group(1), m.group(2))))
            continue
        m = RE_DECL.match(line)
        if m:
            line = '%s = %s' % (m.group(1), m.group(2))
        emit(expr(line))
    return '\n'.join(out) + '\n'


def method_body(path, signature_start):
    """Lines of the method whose declaration starts with signature_start (brace matching)."""
    lines = open(path).read().splitlines()
    i = next(k for k, l in enumerate(lines) if l.strip().startswith(signature_start))
    depth, body = 0, []
    for l in lines[i:]:
        opens, closes = l.count('{'), l.count('}')
        if depth > 0 and not (depth == 1 and closes > opens and l.strip() == '}'):
            body.append(l)
        depth += opens - closes
        if depth == 0 and body:
            break
    return body


# ---- stubs for the Java object graph ------------------------------------------------------------------------------------------
class Param:
    def __init__(self, ptype, value, column, order=0, poly=None):
        self.ptype, self.value, self.column, self.order, self.poly = ptype, float(value), column, order, poly

    def getParameterType(self): return self.ptype
    def getValue(self): return self.value
    def getColumn(self): return self.column
    def getOrder(self): return self.order
    def getZernikePolynomial(self): return self.poly


class ZernikePoly:
    """parameter/ZernikeCoefficient.java:40-56 (integer arithmetic of the polynomial table)."""

    def __init__(self, order):
        self.n = int(math.ceil((-3 + math.sqrt(9 + 8 * order)) / 2))
        self.m = 2 * order - self.n * (self.n + 2)
        halfnm = (self.n - abs(self.m)) // 2
        self.p = [self.n - 2 * k for k in range(halfnm + 1)]
        self.c = [(1 if k % 2 == 0 else -1) * math.comb(self.n - k, k) * math.comb(self.n - 2 * k, halfnm - k) for k in range(halfnm + 1)]
        self.length = math.sqrt((1 + (1 if self.m != 0 else 0)) * (self.n + 1) / math.pi)

    def getAzimuthalFrequency(self): return self.m
    def getNumberOfRadialTerms(self): return len(self.c)
    def getRadialExponent(self, j): return self.p[j]
    def getRadialCoefficient(self, j): return self.length * self.c[j]


class Triple:
    def __init__(self, a, b, c): self.a, self.b, self.c = a, b, c
    def getX(self): return self.a
    def getY(self): return self.b
    def getZ(self): return self.c
    def getPrinciplePointX(self): return self.a
    def getPrinciplePointY(self): return self.b
    def getPrincipleDistance(self): return self.c


class Exterior:
    def __init__(self, params): self.params = params
    def get(self, name): return self.params[name]


class Model(list):
    def __init__(self, params, r0, typ=None, **named):
        super().__init__(params)
        self.r0, self.typ, self.named = r0, typ, named

    def getR0(self): return self.r0
    def getType(self): return self.typ
    def getBx(self): return self.named['Bx']
    def getBy(self): return self.named['By']
    def getCx(self): return self.named['Cx']
    def getCy(self): return self.named['Cy']


class Rows:
    def __init__(self, n): self.v = np.zeros((2, n))
    def set(self, r, c, x): self.v[r, c] = x
    def add(self, r, c, x): self.v[r, c] += x


class Vec:
    def __init__(self): self.v = np.zeros(2)
    def set(self, r, x): self.v[r] = x
    def add(self, r, x): self.v[r] += x


class Collinearity:
    pass


def build_functions():
    g = {'math': math}
    d = os.path.join(REF, 'derivation')
    src = transliterate(method_body(os.path.join(d, 'PartialDerivativeFactory.java'), 'private CollinearityEquationFactory('),
                        'def collinearity_init(self, interiorOrientation, exteriorOrientation, objectCoordinate):')
    exec(src, g)
    src = transliterate(method_body(os.path.join(d, 'DistortionModelFactory.java'), 'static void apply('),
                        'def chain_apply(collinearityEquation, A, w, deltaX, deltaY, par_deltaX_xs, par_deltaX_ys, par_deltaY_xs, par_deltaY_ys):')
    exec(src, g)

    class DMF:
        apply = staticmethod(g['chain_apply'])
    g['DistortionModelFactory'] = DMF
    hdr = 'def %s(distortionModel, collinearityEquation, columns, A, w):'
    for name, fname, sig in (('affinity', 'AffinityShearDistortionModelFactory.java', 'static void apply('),
                             ('distance', 'RadialDistanceDistortionModelFactory.java', 'static void apply('),
                             ('radial', 'RadiallySymmetricDistortionModelFactory.java', 'static void apply('),
                             ('tangential', 'TangentialDistortionModelFactory.java', 'static void apply('),
                             ('zernike_gradient', 'ZernikeDistortionModelFactory.java', 'static void apply(ZernikeDistortionModel.Gradient')):
        exec(transliterate(method_body(os.path.join(d, fname), sig), hdr % name), g)
    exec(transliterate(method_body(os.path.join(d, 'ZernikeDistortionModelFactory.java'), 'private static void apply(ZernikeDistortionModel distortionModel'),
                       'def zernike_xy(distortionModel, collinearityEquation, columns, A, w, type):'), g)
    return g


EO = ['CAMERA_COORDINATE_X', 'CAMERA_COORDINATE_Y', 'CAMERA_COORDINATE_Z', 'CAMERA_OMEGA', 'CAMERA_PHI', 'CAMERA_KAPPA']
COEF_TYPES = {121: 'RADIAL_POLYNOMIAL_A', 131: 'TANGENTIAL_POLYNOMIAL_B', 132: 'Bx', 133: 'By', 141: 'Cx', 142: 'Cy',
              151: 'DISTANCE_POLYNOMIAL_D', 161: 'ZERNIKE_POLYNOMIAL_X', 162: 'ZERNIKE_POLYNOMIAL_Y', 163: 'ZERNIKE_POLYNOMIAL_Z'}


def apply_models(g, P, r0, ce, cols, A, w):
    """The switch of PartialDerivativeFactory.java:420-444 over Camera.getDistortionModels(), whose order is the ordinal order
    of DistortionModel.Type (camera/Camera.java:47, camera/distortion/DistortionModel.java:29-37).  P: {type id: [Param]}."""
    zero = lambda: Param('none', 0.0, MAXV)
    if 141 in P or 142 in P:
        g['affinity'](Model([], r0, Cx=(P.get(141) or [zero()])[0], Cy=(P.get(142) or [zero()])[0]), ce, cols, A, w)
    if 131 in P or 132 in P or 133 in P:
        g['tangential'](Model(P.get(131, []), r0, Bx=(P.get(132) or [zero()])[0], By=(P.get(133) or [zero()])[0]), ce, cols, A, w)
    if 121 in P:
        g['radial'](Model(P[121], r0), ce, cols, A, w)
    if 151 in P:
        g['distance'](Model(P[151], r0), ce, cols, A, w)
    if 161 in P:
        g['zernike_xy'](Model(P[161], r0, 'ZERNIKE_X'), ce, cols, A, w, 'ZERNIKE_X')
    if 162 in P:
        g['zernike_xy'](Model(P[162], r0, 'ZERNIKE_Y'), ce, cols, A, w, 'ZERNIKE_Y')
    if 163 in P:
        g['zernike_gradient'](Model(P[163], r0), ce, cols, A, w)


def evaluate(g, io, eo, X, r0, coefs, obs):
    """io = (x0, y0, c); coefs = [(type id, order, value)] in slot order 12...; returns A (2, 12 + ncoef), w (2)."""
    ns = 12 + len(coefs)
    pt = Triple(*[Param('OBJ', v, s) for s, v in enumerate(X)])
    inner = Triple(Param('x0', io[0], 3), Param('y0', io[1], 4), Param('c', io[2], 5))
    outer = Exterior({n: Param(n, v, 6 + k) for k, (n, v) in enumerate(zip(EO, eo))})
    ce = Collinearity()
    g['collinearity_init'](ce, inner, outer, pt)
    ce.interiorOrientation, ce.exteriorOrientation, ce.objectCoordinate = inner, outer, pt
    A, w, cols = Rows(ns), Vec(), set()
    w.set(0, obs[0] - ce.x)
    w.set(1, obs[1] - ce.y)
    for s, nm in enumerate(['X', 'Y', 'Z', 'x0', 'y0', 'c', 'X0', 'Y0', 'Z0', 'omega', 'phi', 'kappa']):
        A.set(0, s, getattr(ce, 'par_xs_' + nm))
        A.set(1, s, getattr(ce, 'par_ys_' + nm))
    P = {}
    for k, (t, o, v) in enumerate(coefs):
        P.setdefault(t, []).append(Param(COEF_TYPES[t], v, 12 + k, o, ZernikePoly(o) if t in (161, 162, 163) else None))
    apply_models(g, P, r0, ce, cols, A, w)
    return A.v, w.v


CASES = {
    'pinhole': [],
    'radial_tangential_affinity': [(141, 0, 2e-4), (142, 0, -1e-4), (132, 0, 3e-6), (133, 0, -2e-6), (121, 1, -2e-4), (121, 2, 3e-7), (121, 3, -1e-10)],
    'tangential_polynomial': [(132, 0, 3e-6), (133, 0, -2e-6), (131, 1, 1e-4), (131, 2, -1e-6)],
    'distance': [(121, 1, -2e-4), (151, 1, 0.3), (151, 2, -2e-3), (151, 3, 1e-6)],
    'zernike': [(161, 3, 1e-4), (161, 8, -2e-5), (162, 4, 3e-5), (162, 7, 1e-5), (163, 5, 2e-5), (163, 9, -1e-5), (163, 12, 1e-6)],
}


# ---- datum condition rows (BundleAdjustment.addDatumConditionRows, BundleAdjustment.java:493-635) ---------------------------------
class DatumPoint:
    def __init__(self, xyz, cols, datum):
        self.p = [Param('OBJ', v, int(c)) for v, c in zip(xyz, cols)]
        self.datum = bool(datum)

    def getX(self): return self.p[0]
    def getY(self): return self.p[1]
    def getZ(self): return self.p[2]
    def isDatum(self): return self.datum


class RankDefect:
    def __init__(self, flags): self.f = [bool(x) for x in flags]
    def getDefect(self): return sum(self.f)
    def estimateTranslationX(self): return self.f[0]
    def estimateTranslationY(self): return self.f[1]
    def estimateTranslationZ(self): return self.f[2]
    def estimateRotationX(self): return self.f[3]
    def estimateRotationY(self): return self.f[4]
    def estimateRotationZ(self): return self.f[5]
    def estimateScale(self): return self.f[6]


class Dense:
    def __init__(self, r, c): self.v = np.zeros((r, c))
    def set(self, r, c, x): self.v[r, c] = x
    def get(self, r, c): return self.v[r, c]


class Adjustment:
    def getClass(self): return 'BundleAdjustment'


def datum_vectors():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle.bookkeeping import Bookkeeping
    from tests.scenes import random_scene
    g = {'math': math}
    body = method_body(os.path.join(REF, 'BundleAdjustment.java'), 'private void addDatumConditionRows(')
    exec(transliterate(body, 'def add_datum_rows(self, N):'), g)
    out = {}
    for seed in range(12):
        sc = random_scene(seed)
        bk = Bookkeeping(sc)
        if bk.d == 0:
            continue
        pts = sc['points']
        n = bk.n_unknown + bk.d
        adj = Adjustment()
        adj.rankDefect = RankDefect(bk.defect_free)
        cols = np.asarray(bk.pt_col if hasattr(bk, 'pt_col') else None)
        adj.objectCoordinates = [DatumPoint(pts['xyz'][p], cols.reshape(-1, 3)[p], pts['datum'][p]) for p in bk.oc_order.tolist()]
        N = Dense(bk.d, n)
        g['add_datum_rows'](adj, N)
        out['seed%d' % seed] = N.v
    return out


def main():
    g = build_functions()
    rng = np.random.default_rng(20261019)
    out = {}
    for name, coefs in CASES.items():
        ins, As, ws = [], [], []
        for _ in range(6):
            io = np.array([0.02, 0.06, 28.8]) * (1 + rng.normal(0, 0.01, 3))
            ang = rng.uniform(-math.pi, math.pi, 3)
            X0 = rng.uniform(-500, 500, 3)
            # a point in front of the camera, inside the image format
            so, co, sp, cp, sk, ck = math.sin(ang[0]), math.cos(ang[0]), math.sin(ang[1]), math.cos(ang[1]), math.sin(ang[2]), math.cos(ang[2])
            R = np.array([[cp * ck, -cp * sk, sp], [co * sk + so * sp * ck, co * ck - so * sp * sk, -so * cp],
                          [so * sk - co * sp * ck, so * ck + co * sp * sk, co * cp]])
            local = np.array([rng.uniform(-900, 900), rng.uniform(-600, 600), -rng.uniform(2500, 3500)])
            X = X0 + R @ local
            obs = rng.uniform(-15, 15, 2)
            A, w = evaluate(g, io, np.concatenate([X0, ang]), X, 10.0, coefs, obs)
            ins.append(np.concatenate([io, X0, ang, X, obs]))
            As.append(A)
            ws.append(w)
        out[name + '_inputs'] = np.array(ins)
        out[name + '_A'] = np.array(As)
        out[name + '_w'] = np.array(ws)
        out[name + '_coefs'] = np.array(coefs, float).reshape(-1, 3)
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, {k: v.shape for k, v in out.items() if k.endswith('_A')})
    dat = datum_vectors()
    np.savez_compressed(OUT.replace('reference_jacobian', 'reference_datum_rows'), **dat)
    print('wrote datum rows', {k: v.shape for k, v in dat.items()})


if __name__ == '__main__':
    main()
