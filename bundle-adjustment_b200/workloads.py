"""Synthetic close-range networks of BASELINE.json configs[1..4] and the construction of the product's
object graph (host mirror API) from a plain-data scene -- the workload generator of bench.py and of the parity tests.

A *scene* is the plain-data description of an adjustment before bookkeeping: points (xyz, fixed flags, datum flags), cameras
(interior orientation, distortion coefficients, images with exterior orientation and image points), scale bars, directly
observed groups.  ``build_adjustment`` turns it into the object graph a reader of the reference would build
(util/io/reader/aicon/AICONReportFileReader.java:133-387 builds the same graph from a report); ``flat_problem`` runs the
product's own bookkeeping (prepareUnknownParameters / detectRankDefect mirrors) and returns the flat C-ABI arrays.
"""
import numpy as np

# ---- synthetic close-range networks (BASELINE.json configs[1..4]) --------------------------------
IO_TRUTH = dict(c=28.8, x0=0.02, y0=0.06, A=(-1.1e-4, 1.5e-7, -1e-10), B=(6e-6, -9e-6), C=(-7e-5, -3e-5), r0=13.488,
                D=(1e-3, -1e-6, 1e-9))

CONFIGS = {
    2: dict(images=50, targets=500, dist_d=False, rho=False, observed=False),
    3: dict(images=100, targets=2000, dist_d=False, rho=True, observed=True),
    4: dict(images=200, targets=5000, dist_d=True, rho=False, observed=False),
    5: dict(images=500, targets=20000, dist_d=True, rho=False, observed=False),
}


def _angles_from_rotation(R):
    # R as in PartialDerivativeFactory.java:125-135: r13 = sin(phi), r23 = -sin(omega)cos(phi), r33 = cos(omega)cos(phi),
    # r11 = cos(phi)cos(kappa), r12 = -cos(phi)sin(kappa)
    phi = np.arcsin(R[0, 2])
    omega = np.arctan2(-R[1, 2], R[2, 2])
    kappa = np.arctan2(-R[0, 1], R[0, 0])
    return omega, phi, kappa


def _rotation(omega, phi, kappa):
    so, co, sp, cp, sk, ck = np.sin(omega), np.cos(omega), np.sin(phi), np.cos(phi), np.sin(kappa), np.cos(kappa)
    return np.array([[cp * ck, -cp * sk, sp],
                     [co * sk + so * sp * ck, co * ck - so * sp * sk, -so * cp],
                     [so * sk - co * sp * ck, so * ck + co * sp * sk, co * cp]])


def project(io, coefs, r0, eo, X):
    """Forward model of the reference (collinearity + distortion at the undistorted position), vectorised over points."""
    x0, y0, c = io
    R = _rotation(eo[3], eo[4], eo[5])
    dX = X - np.asarray(eo[:3])
    k = dX @ R            # kx, ky, N = R' dX
    N = k[:, 2]
    xs, ys = -c * k[:, 0] / N, -c * k[:, 1] / N
    r2 = xs * xs + ys * ys
    dx = np.zeros_like(xs)
    dy = np.zeros_like(ys)
    by_type = {}
    for (t, o, v, _f) in coefs:
        by_type.setdefault(t, []).append((o, v))
    if 141 in by_type:
        cx, cy = by_type[141][0][1], by_type[142][0][1]
        dx += cx * xs + cy * ys
    if 132 in by_type:
        bx, by = by_type[132][0][1], by_type[133][0][1]
        tx = bx * (r2 + 2 * xs * xs) + by * 2 * xs * ys
        ty = by * (r2 + 2 * ys * ys) + bx * 2 * xs * ys
        s = 1.0 + sum(v * r2 ** o for (o, v) in by_type.get(131, []))
        dx += tx * s
        dy += ty * s
    for (o, v) in by_type.get(121, []):
        f = v * (r2 ** o - (r0 * r0) ** o)
        dx += xs * f
        dy += ys * f
    for (o, v) in by_type.get(151, []):
        f = v * (r2 ** o - (r0 * r0) ** o) / N
        dx += xs * f
        dy += ys * f
    return np.stack([x0 + xs + dx, y0 + ys + dy], axis=1), N


def synthetic_scene(config, images=None, targets=None, seed=None, sigma_img=0.0005, visibility=1.0, free_network=None,
                    n_cameras=1):
    """Deterministic synthetic network (seed = 20260000 + config).  Returns (scene, truth)."""
    cfg = dict(CONFIGS[config])
    I = images or cfg['images']
    T = targets or cfg['targets']
    rng = np.random.Generator(np.random.PCG64(seed if seed is not None else 20260000 + config))
    pts_true = rng.uniform([-1000, -750, -250], [1000, 750, 250], size=(T, 3))
    coefs_true = [(141, 0, IO_TRUTH['C'][0], False), (142, 0, IO_TRUTH['C'][1], False),
                  (132, 0, IO_TRUTH['B'][0], False), (133, 0, IO_TRUTH['B'][1], False),
                  (121, 1, IO_TRUTH['A'][0], False), (121, 2, IO_TRUTH['A'][1], False), (121, 3, IO_TRUTH['A'][2], False)]
    if cfg['dist_d']:
        coefs_true += [(151, 1, IO_TRUTH['D'][0], False), (151, 2, IO_TRUTH['D'][1], False), (151, 3, IO_TRUTH['D'][2], False)]
    io_true = np.array([IO_TRUTH['x0'], IO_TRUTH['y0'], IO_TRUTH['c']])
    elev = np.deg2rad([20.0, 45.0, 70.0])
    cams = []
    per_cam = [[] for _ in range(n_cameras)]
    eo_true = []
    for i in range(I):
        ring = i % 3
        az = 2 * np.pi * (i // 3) / max(1, (I + 2) // 3) + 0.3 * ring
        rad = rng.uniform(2500, 3500)
        pos = rad * np.array([np.cos(elev[ring]) * np.cos(az), np.cos(elev[ring]) * np.sin(az), np.sin(elev[ring])])
        look = rng.uniform(-200, 200, size=3)
        zc = pos - look
        zc /= np.linalg.norm(zc)          # camera z axis points away from the scene
        up = np.array([0.0, 0.0, 1.0]) if abs(zc[2]) < 0.95 else np.array([1.0, 0.0, 0.0])
        xc = np.cross(up, zc); xc /= np.linalg.norm(xc)
        yc = np.cross(zc, xc)
        roll = (i % 4) * np.pi / 2 + rng.uniform(-0.05, 0.05)
        xr = np.cos(roll) * xc + np.sin(roll) * yc
        yr = -np.sin(roll) * xc + np.cos(roll) * yc
        R = np.stack([xr, yr, zc], axis=1)
        om, ph, ka = _angles_from_rotation(R)
        eo_true.append(np.array([pos[0], pos[1], pos[2], om, ph, ka]))
    eo_true = np.array(eo_true)
    # initial values = truth + noise
    pts0 = pts_true + rng.normal(0, 0.5, size=pts_true.shape)
    io0 = io_true * (1 + rng.normal(0, 0.01, size=3))
    coefs0 = [(t, o, v * (1 + rng.normal(0, 0.01)), f) for (t, o, v, f) in coefs_true]
    eo0 = eo_true + np.concatenate([rng.normal(0, 0.5, size=(I, 3)), rng.normal(0, 1e-4, size=(I, 3))], axis=1)
    img_list = [[] for _ in range(n_cameras)]
    for i in range(I):
        xy, N = project(io_true, coefs_true, IO_TRUTH['r0'], eo_true[i], pts_true)
        vis = (np.abs(xy[:, 0]) < 18.0) & (np.abs(xy[:, 1]) < 12.0) & (N < 0)
        if visibility < 1.0:
            vis &= rng.uniform(size=T) < visibility
        idx = np.nonzero(vis)[0].astype(np.int32)
        obs = xy[idx] + rng.normal(0, sigma_img, size=(idx.size, 2))
        rho = rng.uniform(-0.6, 0.6, size=idx.size) if cfg['rho'] else np.zeros(idx.size)
        img_list[i % n_cameras].append({'eo_val': eo0[i].copy(), 'eo_fixed': np.zeros(6, bool), 'obj': idx, 'xy': obs,
                                        'sigma': np.full((idx.size, 2), sigma_img), 'rho': rho})
    cameras = [{'r0': IO_TRUTH['r0'], 'io_val': io0.copy(), 'io_fixed': np.zeros(3, bool), 'coefs': list(coefs0),
                'images': img_list[k]} for k in range(n_cameras)]
    groups = []
    if cfg['observed']:
        r = 3 * T
        sigma_c = 0.05
        G = rng.standard_normal((r, r))
        S = sigma_c ** 2 * (0.2 * np.eye(r) + 0.8 * (G @ G.T) / r)
        Lc = np.linalg.cholesky(S)
        obs = pts_true.reshape(-1) + Lc @ rng.standard_normal(r)
        iu = np.triu_indices(r)
        packed = np.empty(r * (r + 1) // 2)
        packed[iu[0] + iu[1] * (iu[1] + 1) // 2] = S[iu]
        groups.append({'refs': [('point', p, c) for p in range(T) for c in range(3)], 'obs': obs, 'var': None,
                       'dispersion': packed})
    scene = {'points': {'xyz': pts0, 'fixed': np.zeros((T, 3), bool), 'datum': np.ones(T, bool)},
             'cameras': cameras, 'scale_bars': [], 'observed_groups': groups}
    if free_network is False:
        # datum by three fixed (error-free) object points instead of the free-network conditions
        scene['points']['fixed'][:3] = True
        scene['points']['xyz'][:3] = pts_true[:3]
    truth = dict(points=pts_true, io=io_true, coefs=coefs_true, eo=eo_true)
    return scene, truth


def random_scene(seed):
    """Deterministic random network for the randomized parity tests: configuration, number of cameras, size,
    visibility, fixed point components / points / EO components / coefficients and an optional scale bar all vary
    with the seed (datum defects from 0 to 7 occur)."""
    rng = np.random.default_rng(seed)
    cfg = int(rng.choice([2, 4]))
    ncam = int(rng.integers(1, 3))
    images = int(rng.integers(8, 15))
    targets = int(rng.integers(50, 110))
    vis = float(rng.uniform(0.55, 1.0))
    sc = synthetic_scene(cfg, images=images, targets=targets, visibility=vis, seed=1000 + seed, n_cameras=ncam)[0]
    for _ in range(int(rng.integers(0, 4))):
        sc['points']['fixed'][int(rng.integers(targets)), int(rng.integers(3))] = True
    if rng.uniform() < 0.5:
        sc['points']['fixed'][int(rng.integers(targets))] = True
    allimgs = [im for c in sc['cameras'] for im in c['images']]
    if rng.uniform() < 0.7:
        allimgs[int(rng.integers(len(allimgs)))]['eo_fixed'][int(rng.integers(6))] = True
    cam = sc['cameras'][int(rng.integers(ncam))]
    k = int(rng.integers(len(cam['coefs'])))
    cam['coefs'][k] = cam['coefs'][k][:3] + (True,)
    if rng.uniform() < 0.3:
        a, b = 0, 1
        sc['scale_bars'] = [(a, b, float(np.linalg.norm(sc['points']['xyz'][a] - sc['points']['xyz'][b])) + 0.01, 0.02)]
    return sc


# ---- object graph / flat C-ABI arrays of a scene ----------------------------------------------------------------------

from . import host as ba

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
            if im.get('dispersion') is not None:         # extension: fully populated dispersion of the image's coordinates
                img.setDispersion(im['dispersion'])
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
