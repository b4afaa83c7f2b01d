"""ORACLE (test infrastructure, not the product): CPU restatement of JAICOV's adjustment loop.

Follows ``BundleAdjustment.estimateModel`` (BundleAdjustment.java:203-387) pass by pass:
``createNormalEquation`` (:789-834) -> ``applyPrecondition`` (NormalEquationSystem.java:82-91) ->
``MathExtension.solve`` = dspsv [+ dsptri] (MathExtension.java:338-366) -> undo precondition (:273, :297)
-> ``updateModel`` / ``getOmega`` (:389-491) -> convergence test (:327-350), with the centroid shift
(:115-201) and the datum border rows (:493-635).  Per-observation arithmetic is in jaicov_oracle.c
(same directory), the factorisation is reference LAPACK out of scipy's OpenBLAS (lapack_packed.py).

Parity pins:
* tests/test_oracle_golden.py checks this oracle against the AICON report bundled with the reference
  (JAICOV/example/example.htm: S0, IO values, IO standard deviations and correlations, object coordinates) through the
  fixture tests/golden/example_scene.npz;
* tests/test_reference_formulas.py checks it against the REFERENCE'S OWN CODE, executed: the Java method bodies are
  transliterated mechanically and run on stub objects (tests/golden/make_*_fixture.py).  Jacobian rows, bookkeeping,
  datum rows, normal equations, preconditioner, Levenberg-Marquardt control and centroid are identical bit for bit;
  complete adjustments (estimateModel end to end: FULL / NONE / SIMULATION, Levenberg-Marquardt, iteration limit, scale
  bar, observed groups) agree in state, number of passes, every damping value, parameters (1e-13), Omega, sigma0^2 and
  Qxx (1e-11 correlation-scaled).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Levenberg-Marquardt damping (:801-822, :390-426) is restated as well (dampingValue defaults to 0, :96).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

from . import lapack_packed as lp
from .bookkeeping import COL_FIXED, COL_UNSET, Bookkeeping

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EPS = 2.0 ** -53                       # Constant.mashEPS(), Constant.java:61-75
SQRT_EPS = math.sqrt(EPS)              # BundleAdjustment.java:77
MAX_ITER = 5000                        # DefaultValue.java:25

# EstimationStateType ids, EstimationStateType.java:24-42
ERROR_FREE_ESTIMATION, INTERRUPT, SINGULAR_MATRIX, NO_CONVERGENCE, OUT_OF_MEMORY = 1, -1, -2, -4, -7


def build_lib(force=False):
    """Compile jaicov_oracle.c (gcc, -ffp-contract=off) into oracle/libjaicov_oracle.so."""
    src = os.path.join(_HERE, 'jaicov_oracle.c')
    out = os.path.join(_HERE, 'libjaicov_oracle.so')
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared', '-std=gnu11',
                               '-o', out, src, '-lm'])
    return out


class _Problem(ctypes.Structure):
    _fields_ = [
        ('nCam', ctypes.c_int32), ('io_val', ctypes.c_void_p), ('io_col', ctypes.c_void_p), ('r0', ctypes.c_void_p),
        ('coef_ptr', ctypes.c_void_p), ('coef_type', ctypes.c_void_p), ('coef_order', ctypes.c_void_p),
        ('coef_val', ctypes.c_void_p), ('coef_col', ctypes.c_void_p),
        ('nImg', ctypes.c_int32), ('cam_of_img', ctypes.c_void_p), ('eo_val', ctypes.c_void_p),
        ('eo_col', ctypes.c_void_p), ('pt_ptr', ctypes.c_void_p),
        ('m', ctypes.c_int64), ('obj_idx', ctypes.c_void_p), ('xy', ctypes.c_void_p), ('var', ctypes.c_void_p),
        ('rho', ctypes.c_void_p),
        ('nPt', ctypes.c_int32), ('xyz', ctypes.c_void_p), ('pt_col', ctypes.c_void_p),
        ('nBar', ctypes.c_int32), ('bar_a', ctypes.c_void_p), ('bar_b', ctypes.c_void_p),
        ('bar_len', ctypes.c_void_p), ('bar_var', ctypes.c_void_p),
    ]


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build_lib())
        vp, d, i32, i64 = ctypes.c_void_p, ctypes.c_double, ctypes.c_int32, ctypes.c_int64
        L.orc_eval_point.restype = ctypes.c_int
        L.orc_eval_point.argtypes = [ctypes.POINTER(_Problem), i32, i64, d, vp, vp, vp, vp, vp]
        L.orc_stack_image_points.restype = None
        L.orc_stack_image_points.argtypes = [ctypes.POINTER(_Problem), d, vp, vp, i64, i64]
        L.orc_stack_scale_bars.restype = None
        L.orc_stack_scale_bars.argtypes = [ctypes.POINTER(_Problem), d, vp, vp]
        L.orc_omega_image_points.restype = d
        L.orc_omega_image_points.argtypes = [ctypes.POINTER(_Problem), d, vp]
        L.orc_omega_scale_bars.restype = d
        L.orc_omega_scale_bars.argtypes = [ctypes.POINTER(_Problem), d, vp]
        L.orc_apply_precondition.restype = None
        L.orc_apply_precondition.argtypes = [i64, vp, vp, vp]
        L.orc_preconditioner.restype = None
        L.orc_preconditioner.argtypes = [i64, vp, vp, d]
        L.orc_unpack.restype = None
        L.orc_unpack.argtypes = [i64, vp, vp]
        L.orc_pack.restype = None
        L.orc_pack.argtypes = [i64, vp, vp]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


def _i32(a):
    a = np.asarray(a, np.int64)
    return np.ascontiguousarray(a.astype(np.int32))


class FlatProblem:
    """Flat arrays (the layout of include/jaicov_b200.h) built from a raw scene + oracle bookkeeping."""

    def __init__(self, scene, bk: Bookkeeping | None = None):
        self.scene = scene
        self.bk = bk = bk or Bookkeeping(scene)
        cams = scene['cameras']
        self.nCam = len(cams)
        self.io_val = np.ascontiguousarray(np.array([c['io_val'] for c in cams], float).reshape(-1))
        self.io_col = _i32(np.concatenate(bk.io_col) if cams else np.zeros(0))
        self.r0 = np.array([c['r0'] for c in cams], float)
        cp = [0]
        ctype, cord, cval = [], [], []
        for c in cams:
            for (t, o, v, _f) in c['coefs']:
                ctype.append(t); cord.append(o); cval.append(v)
            cp.append(len(ctype))
        self.coef_ptr = np.array(cp, np.int32)
        self.coef_type = np.array(ctype, np.int32)
        self.coef_order = np.array(cord, np.int32)
        self.coef_val = np.array(cval, float)
        self.coef_col = _i32(np.concatenate(bk.coef_col) if cams else np.zeros(0))
        cam_of_img, eo, ptp, obj, xy, var, rho = [], [], [0], [], [], [], []
        for ci, c in enumerate(cams):
            for img in c['images']:
                cam_of_img.append(ci)
                eo.append(np.asarray(img['eo_val'], float))
                o = np.asarray(img['obj'], np.int64)
                obj.append(o)
                xy.append(np.asarray(img['xy'], float).reshape(-1, 2))
                s = np.asarray(img['sigma'], float).reshape(-1, 2)
                var.append(s * s)                       # camera/ImageCoordinate.java:50-51
                rho.append(np.asarray(img['rho'], float).reshape(-1))
                ptp.append(ptp[-1] + o.size)
        self.nImg = len(cam_of_img)
        self.cam_of_img = np.array(cam_of_img, np.int32)
        self.eo_val = np.ascontiguousarray(np.concatenate(eo) if eo else np.zeros(0))
        self.eo_col = _i32(np.concatenate(bk.eo_col) if eo else np.zeros(0))
        self.pt_ptr = np.array(ptp, np.int64)
        self.m = int(ptp[-1])
        self.obj_idx = _i32(np.concatenate(obj) if obj else np.zeros(0))
        self.xy = np.ascontiguousarray(np.concatenate(xy).reshape(-1) if xy else np.zeros(0))
        self.var = np.ascontiguousarray(np.concatenate(var).reshape(-1) if var else np.zeros(0))
        self.rho = np.ascontiguousarray(np.concatenate(rho) if rho else np.zeros(0))
        self.nPt = int(scene['points']['xyz'].shape[0])
        self.xyz = np.ascontiguousarray(np.array(scene['points']['xyz'], float).reshape(-1))
        self.pt_col = _i32(bk.pt_col.reshape(-1))
        self.is_datum = np.ascontiguousarray(np.asarray(scene['points']['datum'], bool))
        bars = scene.get('scale_bars', [])
        self.nBar = len(bars)
        self.bar_a = np.array([b[0] for b in bars], np.int32)
        self.bar_b = np.array([b[1] for b in bars], np.int32)
        self.bar_len = np.array([b[2] for b in bars], float)
        self.bar_var = np.array([float(b[3]) ** 2 for b in bars], float)   # ScaleBar.java:38
        # directly observed groups: (cols, slot views, obs, var|None, dispersion|None)
        self.groups = []
        img_base = 0
        for g in scene.get('observed_groups', []):
            refs = g['refs']
            self.groups.append({
                'refs': refs,
                'obs': np.array(g['obs'], float),
                'var': None if g.get('var') is None else np.array(g['var'], float),
                'dispersion': None if g.get('dispersion') is None else np.array(g['dispersion'], float),
                'P': None,
            })
        self.n = bk.n_unknown + bk.d
        self.d = bk.d

    # access to a referenced unknown: (value array, flat index, column)
    def ref(self, kind, index, comp):
        if kind == 'point':
            return self.xyz, 3 * index + comp, int(self.pt_col[3 * index + comp])
        if kind == 'io':
            return self.io_val, 3 * index + comp, int(self.io_col[3 * index + comp])
        if kind == 'coef':
            k = int(self.coef_ptr[index]) + comp
            return self.coef_val, k, int(self.coef_col[k])
        if kind == 'eo':
            return self.eo_val, 6 * index + comp, int(self.eo_col[6 * index + comp])
        raise KeyError(kind)

    def cstruct(self):
        p = _Problem()
        p.nCam = self.nCam; p.io_val = _ptr(self.io_val); p.io_col = _ptr(self.io_col); p.r0 = _ptr(self.r0)
        p.coef_ptr = _ptr(self.coef_ptr); p.coef_type = _ptr(self.coef_type); p.coef_order = _ptr(self.coef_order)
        p.coef_val = _ptr(self.coef_val); p.coef_col = _ptr(self.coef_col)
        p.nImg = self.nImg; p.cam_of_img = _ptr(self.cam_of_img); p.eo_val = _ptr(self.eo_val)
        p.eo_col = _ptr(self.eo_col); p.pt_ptr = _ptr(self.pt_ptr)
        p.m = self.m; p.obj_idx = _ptr(self.obj_idx); p.xy = _ptr(self.xy); p.var = _ptr(self.var); p.rho = _ptr(self.rho)
        p.nPt = self.nPt; p.xyz = _ptr(self.xyz); p.pt_col = _ptr(self.pt_col)
        p.nBar = self.nBar; p.bar_a = _ptr(self.bar_a); p.bar_b = _ptr(self.bar_b)
        p.bar_len = _ptr(self.bar_len); p.bar_var = _ptr(self.bar_var)
        return p


def _active(c):
    return c >= 0 and c != COL_FIXED


def pidx(r, c):
    return r + c * (c + 1) // 2


class Oracle:
    """One adjustment, reference semantics. ``invert``: 'FULL', 'NONE', 'REDUCED' or 'PRE_ELIMINATION'
    (BundleAdjustment.MatrixInversion, BundleAdjustment.java:65-70)."""

    def __init__(self, scene, invert='FULL', max_iter=MAX_ITER, use_centroid=True, apply_aposteriori=True, damping=0.0,
                 simulation=False):
        self.simulation = bool(simulation)         # EstimationType.SIMULATION, BundleAdjustment.java:830-831, :429-430, :1092
        self.bk = Bookkeeping(scene)
        self.fp = FlatProblem(scene, self.bk)
        self.invert = invert
        self.max_iter = max_iter
        self.use_centroid = use_centroid
        self.apply_aposteriori = apply_aposteriori
        self.sigma2apriori = self.bk.sigma2apriori
        self.damping_value = abs(damping)          # setLevenbergMarquardtDampingValue, BundleAdjustment.java:1189-1191
        self.adapted_damping = 0.0
        self.derive_first_damping = False
        self.last_valid_max_abs_dx = 0.0
        self.lm_steps = []                         # (lastAdaptedDampingValue, adaptedDampingValue, accepted)
        self.centroid = np.zeros(3)
        self.omega = 0.0
        self.max_abs_dx = 0.0
        self.history = []
        self.Qxx = None
        self.iterations = 0
        self.status = 0

    # ---- centroidCoordinates, BundleAdjustment.java:115-201 ---------------------------------------
    def _coord_params(self):
        fp = self.fp
        out = []
        for comp in range(3):
            cols_p = fp.pt_col[comp::3].astype(np.int64)
            idx_p = np.nonzero((cols_p >= 0) & (cols_p != COL_FIXED))[0]
            cols_e = fp.eo_col[comp::6].astype(np.int64)
            idx_e = np.nonzero((cols_e >= 0) & (cols_e != COL_FIXED))[0]
            out.append((idx_p, cols_p[idx_p], idx_e, cols_e[idx_e]))
        return out

    def _centroid(self, invert):
        fp = self.fp
        cp = self._coord_params()
        if not invert:
            cnt = [cp[c][0].size + cp[c][2].size for c in range(3)]
            if not (cnt[0] == cnt[1] == cnt[2] and cnt[0] > 0):
                raise RuntimeError('numbers of coordinate components are un-equal or zero %s' % cnt)
            for c in range(3):
                vals = np.concatenate([fp.xyz[3 * cp[c][0] + c], fp.eo_val[6 * cp[c][2] + c]])
                cols = np.concatenate([cp[c][1], cp[c][3]])
                vals = vals[np.argsort(cols, kind='stable')]      # unknownParameters insertion order
                self.centroid[c] = float(np.cumsum(vals)[-1]) / cnt[c]
        sign = 1.0 if invert else -1.0
        for c in range(3):
            s = sign * self.centroid[c]
            fp.xyz[3 * cp[c][0] + c] += s
            fp.eo_val[6 * cp[c][2] + c] += s
        for g in fp.groups:                                       # :180-200
            for i, (kind, index, comp) in enumerate(g['refs']):
                if (kind == 'point') or (kind == 'eo' and comp < 3):
                    g['obs'][i] += sign * self.centroid[comp if kind == 'point' else comp]

    # ---- directly observed groups -------------------------------------------------------------------
    def _group_weight(self, g):
        """DirectlyObservedParameterGroup.getWeightMatrix, parameter/DirectlyObservedParameterGroup.java:67-91."""
        r = len(g['refs'])
        if g['dispersion'] is None:
            return self.sigma2apriori / g['var']                  # diagonal
        if g['P'] is None:
            ap = np.ascontiguousarray(g['dispersion'] * (1.0 / self.sigma2apriori))
            lp.inv_spd_packed(ap, r)
            g['P'] = ap                                           # packed upper, P = sigma0^2 * Sigma^-1
        return g['P']

    def _group_w(self, g):
        fp = self.fp
        w = np.empty(len(g['refs']))
        cols = np.empty(len(g['refs']), np.int64)
        for i, (kind, index, comp) in enumerate(g['refs']):
            arr, k, col = fp.ref(kind, index, comp)
            w[i] = g['obs'][i] - arr[k]                           # PartialDerivativeFactory.java:468-470
            cols[i] = col
        return w, cols

    def _stack_group(self, g, N, n):
        """getPartialDerivativeDirectlyObservedParameters + stackNormalEquationSystem
        (PartialDerivativeFactory.java:447-505).  A is a selection matrix, so every term of the
        reference's O(r^4) loops with A == 0 adds an exact zero; what remains is
        n[col_a] += sum_colP P[a,colP]*w[colP] (colP ascending) and N[col_a,col_b] += P[a,b] for
        sorted col_a <= col_b -- evaluated here in that order."""
        P = self._group_weight(g)
        w, cols = self._group_w(g)
        r = w.size
        act = np.array([_active(int(c)) for c in cols])
        if g['dispersion'] is None:
            for i in range(r):
                if act[i]:
                    n[cols[i]] += 1.0 * P[i] * w[i]
                    N[pidx(cols[i], cols[i])] += 1.0 * P[i] * 1.0
            return
        # full weight matrix (packed upper, symmetric)
        ii = np.arange(r)
        Pfull = np.empty((r, r))
        iu = np.triu_indices(r)
        Pfull[iu] = P[iu[0] + iu[1] * (iu[1] + 1) // 2]
        Pfull.T[iu] = Pfull[iu]
        for i in range(r):
            if not act[i]:
                continue
            n[cols[i]] = _seq_add(n[cols[i]], Pfull[i] * w)
        ai = ii[act]
        ca = cols[act]
        lo = np.minimum(ca[:, None], ca[None, :])
        hi = np.maximum(ca[:, None], ca[None, :])
        sel = ca[:, None] <= ca[None, :]            # one visit per unordered pair (sorted col_a <= col_b)
        np.add.at(N, (lo + hi * (hi + 1) // 2)[sel], Pfull[np.ix_(ai, ai)][sel])

    def _omega_group(self, g, dx):
        P = self._group_weight(g)
        w, cols = self._group_w(g)
        v = w.copy()
        for i in range(w.size):
            if _active(int(cols[i])):
                v[i] += (-1.0 * dx[cols[i]]) * 1.0
        if g['dispersion'] is None:
            return float(np.dot(v, v * P))
        r = w.size
        iu = np.triu_indices(r)
        Pfull = np.empty((r, r))
        Pfull[iu] = P[iu[0] + iu[1] * (iu[1] + 1) // 2]
        Pfull.T[iu] = Pfull[iu]
        return float(np.dot(v, Pfull @ v))

    # ---- addDatumConditionRows, BundleAdjustment.java:493-635 ------------------------------------
    def datum_rows(self):
        """Returns (d x n) dense border rows B (normalised), built exactly like the reference."""
        fp, bk = self.fp, self.bk
        d = bk.d
        B = np.zeros((d, self.fp.n))
        if d == 0:
            return B
        sel = []
        for p in bk.oc_order.tolist():
            cX, cY, cZ = (int(fp.pt_col[3 * p + c]) for c in range(3))
            if (not fp.is_datum[p]) or cX == COL_FIXED or cY == COL_FIXED or cZ == COL_FIXED:
                continue
            sel.append(p)
        if len(sel) < 3:
            raise ValueError('not enough object points to realise the frame datum')
        x0 = y0 = z0 = 0.0
        for p in sel:
            x0 += fp.xyz[3 * p]; y0 += fp.xyz[3 * p + 1]; z0 += fp.xyz[3 * p + 2]
        cnt = float(len(sel))
        x0, y0, z0 = x0 / cnt, y0 / cnt, z0 / cnt
        tx_f, ty_f, tz_f, rx_f, ry_f, rz_f, sc_f = bk.defect_free
        k = 0
        def nxt(flag):
            nonlocal k
            if flag:
                k += 1
                return k - 1
            return -1
        tx, ty, tz, rx, ry, rz, ms = nxt(tx_f), nxt(ty_f), nxt(tz_f), nxt(rx_f), nxt(ry_f), nxt(rz_f), nxt(sc_f)
        norm = np.zeros(d)
        for p in sel:
            cX, cY, cZ = (int(fp.pt_col[3 * p + c]) for c in range(3))
            x = fp.xyz[3 * p] - x0; y = fp.xyz[3 * p + 1] - y0; z = fp.xyz[3 * p + 2] - z0
            if tx >= 0: B[tx, cX] = 1.0; norm[tx] += 1.0
            if ty >= 0: B[ty, cY] = 1.0; norm[ty] += 1.0
            if tz >= 0: B[tz, cZ] = 1.0; norm[tz] += 1.0
            if rx >= 0: B[rx, cY] = z; B[rx, cZ] = -y; norm[rx] += z * z + y * y
            if ry >= 0: B[ry, cX] = -z; B[ry, cZ] = x; norm[ry] += z * z + x * x
            if rz >= 0: B[rz, cX] = y; B[rz, cY] = -x; norm[rz] += x * x + y * y
            if ms >= 0: B[ms, cX] = x; B[ms, cY] = y; B[ms, cZ] = z; norm[ms] += x * x + y * y + z * z
        for p in sel:
            cX, cY, cZ = (int(fp.pt_col[3 * p + c]) for c in range(3))
            if tx >= 0: B[tx, cX] = B[tx, cX] / math.sqrt(norm[tx])
            if ty >= 0: B[ty, cY] = B[ty, cY] / math.sqrt(norm[ty])
            if tz >= 0: B[tz, cZ] = B[tz, cZ] / math.sqrt(norm[tz])
            if rx >= 0: B[rx, cY] /= math.sqrt(norm[rx]); B[rx, cZ] /= math.sqrt(norm[rx])
            if ry >= 0: B[ry, cX] /= math.sqrt(norm[ry]); B[ry, cZ] /= math.sqrt(norm[ry])
            if rz >= 0: B[rz, cX] /= math.sqrt(norm[rz]); B[rz, cY] /= math.sqrt(norm[rz])
            if ms >= 0:
                B[ms, cX] /= math.sqrt(norm[ms]); B[ms, cY] /= math.sqrt(norm[ms]); B[ms, cZ] /= math.sqrt(norm[ms])
        return B

    # ---- createNormalEquation, BundleAdjustment.java:789-834 --------------------------------------
    def create_normal_equation(self):
        fp = self.fp
        n = fp.n
        N = np.zeros(n * (n + 1) // 2)
        nv = np.zeros(n)
        L = lib()
        p = fp.cstruct()
        L.orc_stack_image_points(ctypes.byref(p), self.sigma2apriori, N.ctypes.data, nv.ctypes.data, 0, fp.m)
        L.orc_stack_scale_bars(ctypes.byref(p), self.sigma2apriori, N.ctypes.data, nv.ctypes.data)
        for g in fp.groups:
            self._stack_group(g, N, nv)
        B = self.datum_rows()
        for k in range(fp.d):                                       # N.set(row, column, value), row < column
            nz = np.nonzero(B[k])[0]
            N[k + nz * (nz + 1) // 2] = B[k, nz]
        if self.derive_first_damping:                               # :801-812
            self.adapted_damping = self.damping_value
            self.derive_first_damping = False
        if self.adapted_damping > 0:                                # :814-822: N_cc += lambda * N_cc for every unknown column
            c = np.arange(fp.d, n, dtype=np.int64)
            dg = c + c * (c + 1) // 2
            N[dg] = N[dg] + self.adapted_damping * N[dg]
        V = np.empty(n)
        L.orc_preconditioner(n, N.ctypes.data, V.ctypes.data, EPS)
        if getattr(self, 'simulation', False):                      # :830-831: n.zero()
            nv[:] = 0.0
        return N, nv, V

    def get_omega(self, dx):
        fp = self.fp
        L = lib()
        p = fp.cstruct()
        dx = np.ascontiguousarray(dx)
        om = L.orc_omega_image_points(ctypes.byref(p), self.sigma2apriori, dx.ctypes.data)
        om += L.orc_omega_scale_bars(ctypes.byref(p), self.sigma2apriori, dx.ctypes.data)
        for g in fp.groups:
            om += self._omega_group(g, dx)
        return om

    # ---- updateUnknownParameters, BundleAdjustment.java:450-462 ----------------------------------
    def update_unknowns(self, dx):
        fp = self.fp
        mx = 0.0
        for vals, cols in ((fp.xyz, fp.pt_col), (fp.io_val, fp.io_col), (fp.coef_val, fp.coef_col), (fp.eo_val, fp.eo_col)):
            c = cols.astype(np.int64)
            act = (c >= 0) & (c < COL_FIXED)
            if act.any():
                dv = dx[c[act]]
                m_ = float(np.max(np.abs(dv)))                      # Math.max propagates NaN
                mx = float('nan') if (math.isnan(m_) or math.isnan(mx)) else max(mx, m_)
                vals[act] += dv
        return mx

    # ---- reduceNormalEquationSystem / extractReducedParameters, BundleAdjustment.java:1197-1453 --------------------
    def _reduction_sets(self):
        """Per image (camera -> image order): active EO columns and the row set S = the camera's active IO + distortion
        columns followed by the active X,Y,Z columns of every point seen in the image (:1205-1236, :1300-1311)."""
        fp = self.fp
        out = []
        for img in range(fp.nImg):
            cam = int(fp.cam_of_img[img])
            E = [int(c) for c in fp.eo_col[6 * img:6 * img + 6] if _active(int(c))]
            S = [int(c) for c in fp.io_col[3 * cam:3 * cam + 3] if _active(int(c))]
            S += [int(c) for c in fp.coef_col[fp.coef_ptr[cam]:fp.coef_ptr[cam + 1]] if _active(int(c))]
            pts = fp.obj_idx[int(fp.pt_ptr[img]):int(fp.pt_ptr[img + 1])]
            pc = fp.pt_col.reshape(-1, 3)[pts].reshape(-1).astype(np.int64)
            S += [int(c) for c in pc if _active(int(c))]
            out.append((np.array(E, np.int64), np.array(S, np.int64)))
        return out

    def num_rows_reduced(self):
        """numRows of the reduced system, :262 (3 * |objectCoordinates| even if some components are fixed)."""
        return self.bk.n_io + self.bk.n_dist + 3 * int(self.bk.oc_order.size) + self.bk.d

    def reduce_normal_equation_system(self, N, n):
        """In place on packed N / n (:1225-1342).  The scalar loops of the reference are evaluated here with dense
        gathers (same arithmetic, different summation order); writes with row > column are dropped like MTJ does."""
        nn = self.fp.n
        D = np.empty((nn, nn))
        lib().orc_unpack(nn, N.ctypes.data, D.ctypes.data)
        pre = self.invert == 'PRE_ELIMINATION'
        for E, S in self._reduction_sets():
            if E.size == 0:
                continue
            N22 = D[np.ix_(E, E)].copy()
            n2 = n[E].copy()
            if pre:
                n[E] = 0.0                                             # :1256-1257
            k = E.size
            ap = _pack_upper(N22)
            lp.inv_symm_packed(ap, k)                                  # MathExtension.inv(UpperSymmPackMatrix), :1259
            N22i = _unpack_upper(ap, k)
            if pre:                                                    # :1261-1298
                D[np.ix_(E, E)] = N22i
                n[E] += N22i @ n2
            N1E = D[np.ix_(S, E)]                                      # N[rowN, EO] : never modified by this image
            T = N1E @ N22i                                             # n12 of every row, :1318-1326
            n[S] += -(T @ n2)                                          # :1329
            upd = T @ N1E.T                                            # :1331-1340
            D[np.ix_(S, S)] -= upd
        iu = np.triu_indices(nn)
        N[iu[0] + iu[1] * (iu[1] + 1) // 2] = D[iu]

    def extract_reduced_parameters(self, N, n):
        """dx2 = inv(N22) n2 - inv(N22) N21 dx1 (:1372-1453); N[EO,EO] holds inv(N22), n[EO] holds inv(N22) n2."""
        nn = self.fp.n
        D = np.empty((nn, nn))
        lib().orc_unpack(nn, N.ctypes.data, D.ctypes.data)
        for E, S in self._reduction_sets():
            if E.size == 0:
                continue
            invN22 = D[np.ix_(E, E)]
            dx2 = n[E].copy()
            N1E = D[np.ix_(S, E)]
            dx2 += -(invN22 @ (N1E.T @ n[S]))
            n[E] = dx2

    # ---- updateModel, BundleAdjustment.java:389-442 (incl. the Levenberg-Marquardt step control) -------------------
    def _update_model(self, dx, update_complete):
        if self.adapted_damping > 0:
            alpha = min(0.25 * self.adapted_damping ** -0.05, 0.75)
            dx *= alpha
            prev = self.omega
            cur = self.get_omega(dx)
            prev = float('inf') if prev <= 0 else prev
            converge = prev >= cur
            self.omega = cur
            last = self.adapted_damping
            if converge:
                self.adapted_damping *= 0.2
            else:
                self.adapted_damping *= 5.0
                if self.adapted_damping > 1.0 / SQRT_EPS:
                    self.adapted_damping = 1.0 / SQRT_EPS
                    self.omega = 0.0
            self.lm_steps.append((last, self.adapted_damping, bool(converge)))
            if not converge:
                self.max_abs_dx = self.last_valid_max_abs_dx
                dx[:] = 0.0
                return
        if update_complete:
            self.omega = 0.0 if getattr(self, 'simulation', False) else self.get_omega(dx)      # :429-430
        self.max_abs_dx = self.update_unknowns(dx)                 # :432
        self.last_valid_max_abs_dx = self.max_abs_dx

    def _solve_packed(self, ap, b, n, invert):
        """MathExtension.solve (MathExtension.java:338-366): dspsv [+ dsptri] on the packed bordered system -- the routines the
        reference calls.  (oracle/fast_oracle.py overrides this one call with a blocked route for sizes these cannot finish.)"""
        lp.solve_symm_packed(ap, b, n, invert)

    # ---- estimateModel, BundleAdjustment.java:203-387 ---------------------------------------------
    def estimate(self):
        fp = self.fp
        L = lib()
        runs = self.max_iter - 1
        is_estimated = estimate_complete = False
        is_converge = True
        self.derive_first_damping = self.damping_value > 0         # :207-208
        self.adapted_damping = 0.0
        self.last_valid_max_abs_dx = 0.0
        if self.max_iter == 0:
            estimate_complete = is_estimated = True
        if self.use_centroid:
            self._centroid(False)
        n = fp.n
        while True:
            self.max_abs_dx = 0.0
            self.iterations = self.max_iter - runs
            N, nv, V = self.create_normal_equation()
            L.orc_apply_precondition(n, V.ctypes.data, N.ctypes.data, nv.ctypes.data)
            estimate_complete = is_estimated
            try:
                if estimate_complete:
                    if self.invert in ('REDUCED', 'PRE_ELIMINATION'):          # :261-267
                        self.reduce_normal_equation_system(N, nv)
                        self._solve_packed(N, nv, self.num_rows_reduced(), True)
                    else:
                        self._solve_packed(N, nv, n, self.invert == 'FULL')
                    L.orc_apply_precondition(n, V.ctypes.data, N.ctypes.data, nv.ctypes.data)
                    self.Qxx = N
                else:
                    if self.invert == 'PRE_ELIMINATION':                       # :283-291
                        self.reduce_normal_equation_system(N, nv)
                        self._solve_packed(N, nv, self.num_rows_reduced(), False)
                        self.extract_reduced_parameters(N, nv)
                    else:
                        self._solve_packed(N, nv, n, False)
                    L.orc_apply_precondition(n, V.ctypes.data, None, nv.ctypes.data)
            except (lp.MatrixSingularException, lp.MatrixNotSPDException, ValueError):
                self.status = SINGULAR_MATRIX
                return self.status
            dx = nv
            self._update_model(dx, estimate_complete)              # :317, :389-442
            self.history.append(self.max_abs_dx)
            if math.isinf(self.max_abs_dx) or math.isnan(self.max_abs_dx):
                self.status = SINGULAR_MATRIX
                return self.status
            elif self.max_abs_dx <= SQRT_EPS and runs > 0 and self.adapted_damping == 0:
                is_estimated = True
            else:
                r = runs
                runs -= 1
                if r <= 1:
                    if estimate_complete:
                        is_converge = False
                    is_estimated = True
            if is_estimated or self.adapted_damping <= SQRT_EPS or runs < self.max_iter * 0.5 + 1:   # :352-353
                self.adapted_damping = 0.0
            if estimate_complete:
                break
        if self.use_centroid:
            self._centroid(True)
        self.status = ERROR_FREE_ESTIMATION if is_converge else NO_CONVERGENCE
        return self.status

    def estimate_and_return_qxx(self):
        self.estimate()
        return self.qxx_dense()

    # ---- getters ------------------------------------------------------------------------------------
    def variance_factor_aposteriori(self):
        """getVarianceFactorAposteriori, BundleAdjustment.java:1090-1093."""
        dof = self.bk.dof
        if dof > 0 and self.omega > 0 and not getattr(self, 'simulation', False) and self.apply_aposteriori:
            return abs(self.omega / float(dof))
        return self.sigma2apriori

    def qxx_dense(self):
        n = self.fp.n
        D = np.empty((n, n))
        lib().orc_unpack(n, self.Qxx.ctypes.data, D.ctypes.data)
        return D


def _pack_upper(A):
    k = A.shape[0]
    iu = np.triu_indices(k)
    ap = np.empty(k * (k + 1) // 2)
    ap[iu[0] + iu[1] * (iu[1] + 1) // 2] = A[iu]
    return ap


def _unpack_upper(ap, k):
    iu = np.triu_indices(k)
    A = np.empty((k, k))
    A[iu] = ap[iu[0] + iu[1] * (iu[1] + 1) // 2]
    A.T[iu] = A[iu]
    return A


def _seq_add(acc, terms):
    for t in terms.tolist():
        acc += t
    return acc


def eval_point(fp: FlatProblem, img: int, j: int, sigma2apriori: float):
    """K1 parity helper: (cols, a0, a1, w, P3) of image point j in slot order (12 + ncoef slots)."""
    L = lib()
    p = fp.cstruct()
    cols = np.zeros(256, np.int32); a0 = np.zeros(256); a1 = np.zeros(256); w = np.zeros(2); P3 = np.zeros(3)
    ns = L.orc_eval_point(ctypes.byref(p), img, j, sigma2apriori, cols.ctypes.data, a0.ctypes.data, a1.ctypes.data,
                          w.ctypes.data, P3.ctypes.data)
    return cols[:ns], a0[:ns], a1[:ns], w, P3
