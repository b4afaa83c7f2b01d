"""ORACLE (test infrastructure, not the product): integer bookkeeping of the reference.

Restates, on a plain "raw scene" dict, what the reference does on its object graph
(paths relative to /root/reference/JAICOV/src/org/applied_geodesy/adjustment/bundle/):

* ``BundleAdjustment.prepareUnknownParameters``   BundleAdjustment.java:667-782
* ``addUnknownParameter`` / ``addObservationGroup`` BundleAdjustment.java:637-650
* ``detectRankDefect``                             BundleAdjustment.java:836-1042
* ``RankDefect.getDefect``                         ../defect/RankDefect.java:35-130

Pinned: rows, columns, counts, rank-defect flags and sigma2apriori equal -- bit for bit -- what the reference's own
methods return when executed on 18 networks incl. directly observed groups (tests/golden/make_bookkeeping_fixture.py ->
tests/golden/reference_bookkeeping.npz, tests/test_reference_formulas.py).

Raw scene (insertion-ordered, exactly what a user of the reference API would have built)::

    {'points':  {'xyz': f64[nPt,3], 'fixed': bool[nPt,3], 'datum': bool[nPt]},
     'cameras': [{'r0': float, 'io_val': [x0,y0,c], 'io_fixed': [b,b,b],
                  'coefs': [(ParameterType id, order, value, fixed)],   # evaluation order
                  'images': [{'eo_val': f64[6], 'eo_fixed': bool[6], 'obj': i32[mi],
                              'xy': f64[mi,2], 'sigma': f64[mi,2], 'rho': f64[mi]}]}],
     'scale_bars': [(a, b, length, sigma)],
     'observed_groups': [{'refs': [(kind, index, comp)], 'obs': f64[r], 'var': f64[r] | None,
                          'dispersion': packed-upper f64[r(r+1)/2] | None}]}

``kind`` is 'point' (index = point, comp 0..2), 'io' (index = camera, comp 0..2 = x0,y0,c),
'coef' (index = camera, comp = position in 'coefs') or 'eo' (index = global image number in
camera->image order, comp 0..5).  Only numpy / integer work happens here.
"""
from __future__ import annotations

import numpy as np

COL_UNSET = -1
COL_FIXED = 2147483647  # Integer.MAX_VALUE, parameter/UnknownParameter.java:27

# ParameterType ids (parameter/ParameterType.java:27-75)
PT_OBJ = (311, 312, 313)   # placeholders; only the X/Y/Z role matters for the defect logic
FREE, FIXED = True, False  # estimateXxx() == True  <=>  DefectType.FREE


def _first_appearance(seq: np.ndarray) -> np.ndarray:
    """Distinct values of seq in order of first appearance (LinkedHashSet insertion order)."""
    if seq.size == 0:
        return seq.astype(np.int64)
    _, first = np.unique(seq, return_index=True)
    return seq[np.sort(first)].astype(np.int64)


class Bookkeeping:
    """Result of prepareUnknownParameters: columns, rows, d, sigma2apriori, ordered sets."""

    def __init__(self, scene):
        pts = scene['points']
        nPt = int(pts['xyz'].shape[0])
        cams = scene['cameras']
        bars = scene.get('scale_bars', [])
        groups = scene.get('observed_groups', [])

        pt_col = np.where(np.asarray(pts['fixed'], bool), COL_FIXED, COL_UNSET).astype(np.int64)
        counter = 0
        unknown_order = []          # list of (kind, index, comp) chunks in insertion order (arrays)
        sigma2 = 1.0                # BundleAdjustment.java:98

        # ---- pass 1: image points, BundleAdjustment.java:670-693 ---------------------------------
        n_obs = 0
        all_obj = []
        for cam in cams:
            for img in cam['images']:
                obj = np.asarray(img['obj'], np.int64)
                img['_row0'] = n_obs            # x row = row0 + 2j, y row = row0 + 2j + 1
                n_obs += 2 * obj.size
                all_obj.append(obj)
                if obj.size:
                    s = np.asarray(img['sigma'], float)
                    sigma2 = min(sigma2, float((s * s).min()))   # addObservationGroup, :637-643
        all_obj = np.concatenate(all_obj) if all_obj else np.zeros(0, np.int64)
        oc_order = _first_appearance(all_obj)   # this.objectCoordinates insertion order
        # addUnknownParameter X,Y,Z per newly seen coordinate (only if column == -1)
        for p in oc_order.tolist() if oc_order.size < 4096 else ():
            for c in range(3):
                if pt_col[p, c] == COL_UNSET:
                    pt_col[p, c] = counter
                    counter += 1
        if oc_order.size >= 4096:   # vectorised equivalent of the loop above
            sub = pt_col[oc_order]                       # (k,3)
            unset = sub == COL_UNSET
            ranks = np.cumsum(unset.ravel()).reshape(unset.shape) - 1
            sub = np.where(unset, ranks + counter, sub)
            counter += int(unset.sum())
            pt_col[oc_order] = sub
        unknown_pts_after_pass1 = oc_order.copy()

        # ---- pass 2: interior orientation + distortion, :695-713 -------------------------------
        n_io = n_dist = 0
        io_cols, coef_cols = [], []
        for cam in cams:
            ic = np.where(np.asarray(cam['io_fixed'], bool), COL_FIXED, COL_UNSET).astype(np.int64)
            for c in range(3):
                if ic[c] == COL_UNSET:
                    n_io += 1
                    ic[c] = counter
                    counter += 1
            io_cols.append(ic)
            cc = np.array([COL_FIXED if f else COL_UNSET for (_, _, _, f) in cam['coefs']], np.int64)
            for k in range(cc.size):
                if cc[k] == COL_UNSET:
                    n_dist += 1
                    cc[k] = counter
                    counter += 1
            coef_cols.append(cc)

        # ---- pass 3: exterior orientations, :715-722 --------------------------------------------
        eo_cols = []
        for cam in cams:
            for img in cam['images']:
                ec = np.where(np.asarray(img['eo_fixed'], bool), COL_FIXED, COL_UNSET).astype(np.int64)
                for c in range(6):
                    if ec[c] == COL_UNSET:
                        ec[c] = counter
                        counter += 1
                eo_cols.append(ec)

        # ---- scale bars, :724-745 ------------------------------------------------------------------
        oc_list = oc_order.tolist()
        oc_set = set(oc_list)
        late_points = []
        bar_rows = []
        for (a, b, length, sigma) in bars:
            bar_rows.append(n_obs)
            n_obs += 1
            for p in (int(a), int(b)):
                if p not in oc_set:
                    oc_set.add(p)
                    oc_list.append(p)
            for p in (int(a), int(b)):
                for c in range(3):
                    if pt_col[p, c] == COL_UNSET:
                        pt_col[p, c] = counter
                        counter += 1
                        late_points.append((p, c))
            sigma2 = min(sigma2, float(sigma) ** 2)

        # ---- directly observed groups, :747-771 ---------------------------------------------------
        group_rows = []
        late_unknowns = []
        for g in groups:
            rows = []
            for (kind, index, comp) in g['refs']:
                if kind == 'point':
                    if index not in oc_set:
                        oc_set.add(index)
                        oc_list.append(index)
                    if pt_col[index, comp] == COL_UNSET:
                        pt_col[index, comp] = counter
                        counter += 1
                        late_points.append((index, comp))
                elif kind == 'io':
                    if io_cols[index][comp] == COL_UNSET:   # cannot happen after pass 2; kept for symmetry
                        io_cols[index][comp] = counter
                        counter += 1
                elif kind == 'coef':
                    if coef_cols[index][comp] == COL_UNSET:
                        coef_cols[index][comp] = counter
                        counter += 1
                elif kind == 'eo':
                    if eo_cols[index][comp] == COL_UNSET:
                        eo_cols[index][comp] = counter
                        counter += 1
                rows.append(n_obs)
                n_obs += 1
            group_rows.append(np.array(rows, np.int64))
            var = g['var'] if g.get('dispersion') is None else _packed_diag(g['dispersion'], len(g['refs']))
            sigma2 = min(sigma2, float(np.min(var)))

        self.n_obs = n_obs
        self.n_unknown = counter
        self.n_io = n_io
        self.n_dist = n_dist
        self.sigma2apriori = sigma2 if sigma2 > 0 else 1.0     # BundleAdjustment.java:221
        self.oc_order = np.array(oc_list, np.int64)            # this.objectCoordinates
        self.bar_rows = np.array(bar_rows, np.int64)
        self.group_rows = group_rows
        self.n_images = len(eo_cols)

        # ---- rank defect, :836-1042 ---------------------------------------------------------------
        self.defect_free = self._detect_rank_defect(scene, pt_col, eo_cols)
        d = int(sum(self.defect_free))
        self.d = d

        # ---- renumber, :776-781 (only parameters in unknownParameters, i.e. real columns) -------
        def shift(a):
            a = np.asarray(a, np.int64)
            return np.where((a >= 0) & (a != COL_FIXED), a + d, a)
        self.pt_col = shift(pt_col)
        self.io_col = [shift(a) for a in io_cols]
        self.coef_col = [shift(a) for a in coef_cols]
        self.eo_col = [shift(a) for a in eo_cols]
        self.dof = self.n_obs - self.n_unknown + d             # BundleAdjustment.java:1080-1082

    # order: tx, ty, tz, rx, ry, rz, scale  (True = FREE = estimated by a datum condition)
    def _detect_rank_defect(self, scene, pt_col, eo_cols):
        bars = scene.get('scale_bars', [])
        groups = scene.get('observed_groups', [])
        has_bars = len(bars) > 0
        tx = ty = tz = rx = ry = rz = FREE
        sc = FIXED if has_bars else FREE
        cx = cy = cz = 0

        def role(kind, comp):
            # ParameterType role of an observed parameter
            if kind == 'point':
                return 'XYZ'[comp]
            if kind == 'eo':
                return ('X', 'Y', 'Z', 'omega', 'phi', 'kappa')[comp]
            return None

        def all_fixed():
            return not (tx or ty or tz or rx or ry or rz or sc)

        # :860-881 angles observed directly fix rotations
        for g in groups:
            for (kind, index, comp) in g['refs']:
                r = role(kind, comp)
                if r == 'omega': rx = FIXED
                elif r == 'phi': ry = FIXED
                elif r == 'kappa': rz = FIXED
                if not rx and not ry and not rz:
                    break

        def rules():
            nonlocal tx, ty, tz, rx, ry, rz, sc
            if tx and cx > 0: tx = FIXED
            if ty and cy > 0: ty = FIXED
            if tz and cz > 0: tz = FIXED
            if (not has_bars) and (cx >= 2 or cy >= 2 or cz >= 2): sc = FIXED
            if rx and cy >= 2 and cz >= 2: rx = FIXED
            if ry and cx >= 2 and cz >= 2: ry = FIXED
            if rz and cx >= 2 and cy >= 2: rz = FIXED
            if cx > 0 and cy > 0 and cz > 0 and ((has_bars and cx + cy + cz >= 6) or ((not has_bars) and cx + cy + cz >= 7)):
                rx = ry = rz = FIXED

        # :883-944 directly observed coordinates
        for g in groups:
            for (kind, index, comp) in g['refs']:
                r = role(kind, comp)
                if r == 'X': cx += 1
                elif r == 'Y': cy += 1
                elif r == 'Z': cz += 1
                elif r == 'omega': rx = FIXED
                elif r == 'phi': ry = FIXED
                elif r == 'kappa': rz = FIXED
                rules()
                if all_fixed():
                    break

        # :946-983 fixed object coordinate components, in objectCoordinates order
        for p in self.oc_order.tolist():
            cx += 1 if pt_col[p, 0] == COL_FIXED else 0
            cy += 1 if pt_col[p, 1] == COL_FIXED else 0
            cz += 1 if pt_col[p, 2] == COL_FIXED else 0
            rules()
            if all_fixed():
                break
        if all_fixed():
            return (tx, ty, tz, rx, ry, rz, sc)

        # :990-1041 fixed exterior orientations
        for ec in eo_cols:
            if rx and ec[3] == COL_FIXED: rx = FIXED
            if ry and ec[4] == COL_FIXED: ry = FIXED
            if rz and ec[5] == COL_FIXED: rz = FIXED
            cx += 1 if ec[0] == COL_FIXED else 0
            cy += 1 if ec[1] == COL_FIXED else 0
            cz += 1 if ec[2] == COL_FIXED else 0
            rules()
            if all_fixed():
                break
        return (tx, ty, tz, rx, ry, rz, sc)


def _packed_diag(ap, r):
    ap = np.asarray(ap, float)
    idx = np.arange(r, dtype=np.int64)
    return ap[idx + idx * (idx + 1) // 2]
