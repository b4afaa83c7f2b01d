"""Generates tests/golden/reference_lm_steps.npz: the Levenberg-Marquardt step control of the reference, produced by
EXECUTING BundleAdjustment.updateModel (BundleAdjustment.java:389-448) and updateUnknownParameters (:450-462) on scripted
situations (current damping value, previous Omega, Omega of the shortened step, dx).

Run in the build container only (reads /root/reference):
    python tests/golden/make_lm_fixture.py

The method bodies are transliterated mechanically (make_jacobian_fixture.transliterate); getOmega is scripted (it returns
the Omega value of the scenario), everything else -- step length alpha, accept / reject, the x0.2 / x5 adaptation, the
1/sqrt(eps) clamp, max|dx| bookkeeping, the parameter update -- is the reference's code.  Numbers only are stored.
"""
import math
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_jacobian_fixture as tj  # noqa: E402

BA = os.path.join(tj.REF, 'BundleAdjustment.java')
OUT = os.path.join(HERE, 'reference_lm_steps.npz')
SQRT_EPS = math.sqrt(2.0 ** -53)          # Constant.EPS, Constant.java:61-75; BundleAdjustment.java:77


class JavaMath:
    """java.lang.Math as the transliterated bodies use it: the C library functions plus min / max / abs."""
    min = staticmethod(min)
    max = staticmethod(max)
    abs = staticmethod(abs)


for _k in dir(math):
    if not _k.startswith('_') and not hasattr(JavaMath, _k):
        _v = getattr(math, _k)
        setattr(JavaMath, _k, staticmethod(_v) if callable(_v) else _v)


class Vec:
    def __init__(self, v): self.v = np.array(v, float)
    def scale(self, a): self.v *= a
    def zero(self): self.v[:] = 0.0
    def get(self, i): return self.v[i]
    def size(self): return self.v.size


class Param:
    def __init__(self, value, column): self.value, self.column = float(value), column
    def getColumn(self): return self.column
    def getValue(self): return self.value
    def setValue(self, v): self.value = v


class State:
    def name(self): return 'LEVENBERG_MARQUARDT_STEP'


class Change:
    def __init__(self): self.events = []
    def firePropertyChange(self, name, old, new): self.events.append((old, new))


def ternaries(lines):
    """`lhs = cond ? a : b;` -> explicit if / else (the only ternary form these two methods use)."""
    out = []
    for l in lines:
        m = re.match(r'^(\s*)(?:double\s+|int\s+)?([\w\.]+)\s*=\s*(.+?)\s*\?\s*(.+?)\s*:\s*(.+);\s*$', l)
        if m and '//' not in l.split('?')[0]:
            ind, lhs, cond, a, b = m.groups()
            # one compound statement (so that a brace-less `if (...)` in front of it keeps governing all of it)
            out += ['%sif (true) {' % ind, '%sif (%s) {' % (ind, cond), '%s%s = %s;' % (ind, lhs, a), '%s}' % ind, '%selse {' % ind,
                    '%s%s = %s;' % (ind, lhs, b), '%s}' % ind, '%s}' % ind]
            continue
        # f(a, b, cond ? p : q);  ->  if (cond) { f(a, b, p); } else { f(a, b, q); }
        m = re.match(r'^(\s*)([\w\.]+\((?:.*,\s*)?)([^,?]+?)\s*\?\s*(.+?)\s*:\s*(.+?)\);\s*$', l)
        if m:
            ind, head, cond, a, b = m.groups()
            out += ['%sif (true) {' % ind, '%sif (%s) {' % (ind, cond), '%s%s%s);' % (ind, head, a), '%s}' % ind, '%selse {' % ind,
                    '%s%s%s);' % (ind, head, b), '%s}' % ind, '%s}' % ind]
            continue
        out.append(l)
    return out


def build():
    g = {'math': JavaMath, 'SQRT_EPS': SQRT_EPS}
    fix = lambda src: src.replace('Double.MAX_VALUE', '1.7976931348623157e308').replace('EstimationStateType.LEVENBERG_MARQUARDT_STEP', 'LM_STATE') \
                         .replace('EstimationType.SIMULATION', "'SIMULATION'")
    g['LM_STATE'] = State()
    exec(fix(tj.transliterate(ternaries(tj.method_body(BA, 'private void updateModel(')), 'def updateModel(self, dx, updateCompleteModel):')), g)
    exec(fix(tj.transliterate(tj.method_body(BA, 'private double updateUnknownParameters('), 'def updateUnknownParameters(self, dx):')), g)
    return g


class Adjustment:
    pass


def scenarios(rng, count):
    for k in range(count):
        lam = float(rng.choice([0.0, 1e-3, 0.5, 1.0, 100.0, 3e7, 9e7, 1e8]))
        prev = float(rng.choice([0.0, -1.0, 1.0, 2.5]))
        cur = float(rng.choice([0.5, 1.0, 2.5, 7.0]))
        complete = bool(rng.integers(0, 2))
        dx = rng.normal(0, 1e-3, 6)
        cols = np.array([0, 1, 2147483647, 3, -1, 5])
        vals = rng.normal(0, 10, 6)
        last_valid = float(rng.uniform(0, 1))
        yield lam, prev, cur, complete, dx, cols, vals, last_valid


def main():
    g = build()
    rng = np.random.default_rng(20261020)
    ins, outs = [], []
    for lam, prev, cur, complete, dx, cols, vals, last_valid in scenarios(rng, 400):
        adj = Adjustment()
        adj.adaptedDampingValue, adj.omega, adj.lastValidmaxAbsDx, adj.maxAbsDx = lam, prev, last_valid, 0.0
        adj.estimationType = 'L2NORM'
        adj.change = Change()
        adj.currentEstimationStatus = None
        adj.unknownParameters = [Param(v, int(c)) for v, c in zip(vals, cols)]
        adj.getOmega = lambda dxv, cur=cur: cur
        adj.updateUnknownParameters = lambda dxv, adj=adj: g['updateUnknownParameters'](adj, dxv)
        vec = Vec(dx)
        g['updateModel'](adj, vec, complete)
        ev = adj.change.events[0] if adj.change.events else (-1.0, -1.0)
        ins.append(np.concatenate([[lam, prev, cur, float(complete), last_valid], dx, cols.astype(float), vals]))
        outs.append(np.concatenate([[adj.adaptedDampingValue, adj.omega, adj.maxAbsDx, adj.lastValidmaxAbsDx, ev[0], ev[1]], vec.v,
                                    [p.value for p in adj.unknownParameters]]))
    np.savez_compressed(OUT, inputs=np.array(ins), outputs=np.array(outs))
    print('wrote', OUT, np.array(outs).shape)


if __name__ == '__main__':
    main()
