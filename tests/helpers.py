"""Test/bench helpers: build the product's object graph (host mirror API) from a raw scene, the way a reader of the
reference would (util/io/reader/aicon/AICONReportFileReader.java:133-387 builds the same graph from a report)."""
import numpy as np

import bundle_adjustment_b200 as ba

_T = ba.DistortionModel.Type
_MODEL_OF = {141: _T.AFFINITY_AND_SHEAR, 142: _T.AFFINITY_AND_SHEAR, 131: _T.TANGENTIAL_DISTORTION,
             132: _T.TANGENTIAL_DISTORTION, 133: _T.TANGENTIAL_DISTORTION, 121: _T.RADIAL_DISTORTION,
             151: _T.DISTANCE_DISTORTION, 161: _T.ZERNIKE_X, 162: _T.ZERNIKE_Y, 163: _T.ZERNIKE_GRADIENT}
FIXED = 2147483647


def build_adjustment(scene, device=0):
    pts = ba.ObjectCoordinateArray(scene['points'].get('names'), scene['points']['xyz'])
    pts.datum[:] = scene['points']['datum']
    pts.column[np.asarray(scene['points']['fixed'], bool)] = FIXED
    adj = ba.BundleAdjustment(device=device)
    cams, imgs = [], []
    for ci, c in enumerate(scene['cameras']):
        cam = ba.Camera(ci + 1, c['r0'], *sorted({_MODEL_OF[t] for (t, _o, _v, _f) in c['coefs']}))
        for p, v, f in zip(cam.getInteriorOrientation(), c['io_val'], c['io_fixed']):
            p.setValue(v)
            p.setColumn(FIXED if f else -1)
        cparams = []
        for (t, o, v, f) in c['coefs']:
            m = cam.getDistortionModel(_MODEL_OF[t])
            if t == 141: p = m.getCx()
            elif t == 142: p = m.getCy()
            elif t == 132: p = m.getBx()
            elif t == 133: p = m.getBy()
            else: p = m.add(o)
            p.setValue(v)
            p.setColumn(FIXED if f else -1)
            cparams.append(p)
        for ii, im in enumerate(c['images']):
            img = cam.add(len(imgs) + 1)
            for p, v, f in zip(img.getExteriorOrientation(), im['eo_val'], im['eo_fixed']):
                p.setValue(v)
                p.setColumn(FIXED if f else -1)
            img.addAll(pts, im['obj'], im['xy'], im['sigma'], im['rho'])
            imgs.append(img)
        cams.append((cam, cparams))
        adj.add(cam)
    for (a, b, l, s) in scene.get('scale_bars', []):
        adj.add(ba.ScaleBar(pts[int(a)], pts[int(b)], l, s))
    for g in scene.get('observed_groups', []):
        ops = []
        var = g.get('var')
        for i, (kind, index, comp) in enumerate(g['refs']):
            if kind == 'point':
                ref = (pts[index].getX(), pts[index].getY(), pts[index].getZ())[comp]
            elif kind == 'io':
                ref = list(cams[index][0].getInteriorOrientation())[comp]
            elif kind == 'coef':
                ref = cams[index][1][comp]
            else:
                ref = list(imgs[index].getExteriorOrientation())[comp]
            ops.append(ba.ObservationParameter(ref, g['obs'][i], None if var is None else var[i]))
        adj.add(ba.DirectlyObservedParameterGroup(ops, g.get('dispersion')))
    return adj, pts


def flat_problem(scene):
    """Flat C-ABI arrays of a scene through the product's own bookkeeping (host mirror)."""
    adj, _ = build_adjustment(scene)
    flat = adj._prepare()
    return adj, flat
