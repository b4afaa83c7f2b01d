"""direct linear transformation (initial values), dlt/DirectLinearTransformation.java:67-352 and
dlt/DLTPartialDerivativeFactory.java:40-337."""
import numpy as np
import pytest

import bundle_adjustment_b200 as ba
from oracle import dlt as od
from tests.helpers import build_adjustment
from tests.scenes import project, synthetic_scene

ALL_SETS = [(), (od.IDENTICAL_PRINCIPLE_DISTANCE, od.ROTATION_WITHOUT_SHEAR),
            (od.FIXED_PRINCIPLE_DISTANCE_X, od.FIXED_PRINCIPLE_DISTANCE_Y, od.IDENTICAL_PRINCIPLE_DISTANCE, od.ROTATION_WITHOUT_SHEAR),
            (od.FIXED_PRINCIPAL_POINT_X, od.FIXED_PRINCIPAL_POINT_Y, od.ROTATION_WITHOUT_SHEAR, od.ROTATION_WITHOUT_SHEAR),
            (od.FIXED_PRINCIPLE_DISTANCE_X,)]


def _network(noise, images=9, targets=70, seed=3):
    sc, truth = synthetic_scene(2, images=images, targets=targets)
    rng = np.random.default_rng(seed)
    io, eo, pts = truth['io'], truth['eo'], truth['points']
    obs = []
    for i in range(images):
        xy, _ = project(io, [], 10.0, eo[i], pts)          # pin-hole camera: what the DLT models
        keep = rng.uniform(size=targets) < 0.8
        obs.append((np.nonzero(keep)[0], xy[keep] + rng.normal(0, noise, size=(int(keep.sum()), 2))))
    return truth, obs


def test_oracle_recovers_a_pinhole_camera():
    """Known answer: error-free projections of a pin-hole camera give back its interior and exterior orientation."""
    truth, obs = _network(0.0)
    io, eo, pts = truth['io'], truth['eo'], truth['points']
    for R in ALL_SETS:
        for i, (idx, xy) in enumerate(obs[:4]):
            ok, b, d, passes = od.adjust(xy, pts[idx], (io[2], io[0], io[1]), R)
            assert ok and passes == (1 if not R else passes) and passes <= 4
            np.testing.assert_allclose([d['c'], d['x0'], d['y0']], [io[2], io[0], io[1]], rtol=0, atol=1e-9)
            np.testing.assert_allclose(d['X0'], eo[i][:3], rtol=0, atol=1e-8)
            np.testing.assert_allclose([d['omega'], d['phi'], d['kappa']], eo[i][3:], rtol=0, atol=1e-11)


def test_oracle_restriction_gradients():
    """The vector-form gradients (DPF:68-236) are the derivatives of the restriction functions."""
    rng = np.random.default_rng(11)
    b = rng.normal(size=11)
    c, x0, y0 = 1.3, 0.02, 0.06          # (a small c keeps the rounding noise of the central differences below the tolerance)
    for kind in range(6):
        g, w = od.restriction_row(kind, b, c, x0, y0)
        gn = np.zeros(11)
        for k in range(11):
            bp, bm = b.copy(), b.copy()
            bp[k] += 1e-6
            bm[k] -= 1e-6
            gn[k] = -(od.restriction_row(kind, bp, c, x0, y0)[1] - od.restriction_row(kind, bm, c, x0, y0)[1]) / 2e-6
        np.testing.assert_allclose(g, gn, rtol=1e-6, atol=1e-8)     # the border row is -d(misclosure)/db
        assert g[3] == 0.0 and g[7] == 0.0


def test_validate_restrictions():
    assert od.validate_restrictions([1, 1, 0]) == [1, 0]
    assert od.validate_restrictions([2, 0, 3]) == [2, 3]
    assert od.validate_restrictions([0, 2]) == [0, 2]


@pytest.mark.gpu
@pytest.mark.parametrize('restrictions', ALL_SETS)
def test_dlt_batch_matches_oracle(built, restrictions):
    """One launch for all images vs the oracle image by image, noisy observations (several restricted passes)."""
    truth, obs = _network(0.002)
    io, pts = truth['io'], truth['points']
    # image 3 gets only five points: the reference refuses it (DLT:96-104)
    obs[3] = (obs[3][0][:5], obs[3][1][:5])
    pt_ptr = np.concatenate([[0], np.cumsum([len(i) for i, _ in obs])])
    xy = np.concatenate([x for _, x in obs])
    xyz = np.concatenate([pts[i] for i, _ in obs])
    ioc = np.tile([io[2] * 1.001, io[0] + 0.01, io[1] - 0.01], (len(obs), 1))      # camera values used by the FIXED_* restrictions
    out, status, passes = ba._lib.dlt_batch(pt_ptr, xy, xyz, ioc, restrictions)
    worst = 0.0
    for k, (idx, x) in enumerate(obs):
        ok, b, d, p = od.adjust(x, pts[idx], tuple(ioc[k]), restrictions)
        if k == 3:
            assert not ok and status[k] == -1
            continue
        assert ok and status[k] == 1
        assert passes[k] == p
        ref = np.concatenate([b, [d['c'], d['x0'], d['y0']], d['X0'], [d['omega'], d['phi'], d['kappa']]])
        scale = np.maximum(np.abs(ref), 1e-3)
        worst = max(worst, float((np.abs(out[k] - ref) / scale).max()))
    print('DLT', restrictions, 'passes', passes.tolist(), 'worst relative difference', worst)
    assert worst < 1e-8


@pytest.mark.gpu
def test_dlt_host_mirror_initial_values_start_an_adjustment(built):
    """DLTCoefficients / DirectLinearTransformation.adjustAll through the reference-shaped API: the initial exterior
    orientations it delivers are good enough for the bundle adjustment to converge from them."""
    scene, truth = synthetic_scene(2, images=10, targets=80)
    adj, pts = build_adjustment(scene)
    images = [img for cam in adj._cameras for img in cam]
    known = {pts[i].getName(): pts[i] for i in range(len(pts))}
    coefs = [ba.DLTCoefficients(img) for img in images]
    RT = ba.DirectLinearTransformation.RestrictionType
    oks = ba.DirectLinearTransformation.adjustAll(coefs, known, RT.IDENTICAL_PRINCIPLE_DISTANCE, RT.ROTATION_WITHOUT_SHEAR)
    assert all(oks)
    PT = ba.ParameterType
    eo_types = (PT.CAMERA_COORDINATE_X, PT.CAMERA_COORDINATE_Y, PT.CAMERA_COORDINATE_Z, PT.CAMERA_OMEGA, PT.CAMERA_PHI, PT.CAMERA_KAPPA)
    for img, coef, eo_t in zip(images, coefs, truth['eo']):
        got = np.array([coef.get(t).getValue() for t in eo_types])
        assert np.abs(got[:3] - eo_t[:3]).max() < 60.0          # distortion is not modelled by the DLT: metres-level start values
        assert np.abs(np.angle(np.exp(1j * (got[3:] - eo_t[3:])))).max() < 0.05
        for t in eo_types:
            img.getExteriorOrientation().get(t).setValue(coef.get(t).getValue())
    # single-image form, same result as the batch
    c0 = ba.DLTCoefficients(images[0])
    assert ba.DirectLinearTransformation.adjust(c0, known, RT.IDENTICAL_PRINCIPLE_DISTANCE, RT.ROTATION_WITHOUT_SHEAR)
    assert [p.getValue() for p in c0] == [p.getValue() for p in coefs[0]]
    assert adj.estimateModel() == ba.EstimationStateType.ERROR_FREE_ESTIMATION
