"""covariance propagation of transformed object coordinates,
CoordinateTransformationExteriorOrientation.transform (tranformation/CoordinateTransformationExteriorOrientation.java:49-321)."""
import numpy as np
import pytest

import bundle_adjustment_b200 as ba
from oracle import propagation as op
from oracle.oracle import Oracle
from tests.helpers import build_adjustment
from tests.scenes import synthetic_scene


def test_oracle_jacobian_matches_central_differences():
    """The oracle's matrix-form Jacobian (:223-279) is the derivative of the transformation formula (:209-215)."""
    rng = np.random.default_rng(5)
    for _ in range(20):
        X = rng.uniform(-1000, 1000, 3)
        eT = np.concatenate([rng.uniform(-3000, 3000, 3), rng.uniform(-np.pi, np.pi, 3)])
        eS = np.concatenate([rng.uniform(-3000, 3000, 3), rng.uniform(-np.pi, np.pi, 3)])
        J = op.jacobian_block(X, eT, eS)
        p0 = np.concatenate([eT, eS, X])
        Jn = np.zeros((3, 15))
        for k in range(15):
            h = 1e-3 if k % 6 < 3 or k >= 12 else 1e-7
            pp, pm = p0.copy(), p0.copy()
            pp[k] += h
            pm[k] -= h
            Jn[:, k] = (op.transform_point(pp[12:], pp[0:6], pp[6:12]) - op.transform_point(pm[12:], pm[0:6], pm[6:12])) / (2 * h)
        np.testing.assert_allclose(J, Jn, rtol=1e-6, atol=1e-6 * np.abs(J).max())
        # the transformation is a rigid motion: d X_trg / d X is a rotation, d / d X0_src its negative
        np.testing.assert_allclose(J[:, 12:] @ J[:, 12:].T, np.eye(3), atol=1e-14)
        np.testing.assert_array_equal(J[:, 6:9], -J[:, 12:])


def test_oracle_rotation_convention():
    """Entries of R as the reference writes them out (:172-184)."""
    om, ph, ka = 0.3, -0.7, 1.9
    R, _ = op.rotation(om, ph, ka)
    assert R[0, 2] == pytest.approx(np.sin(ph))
    assert R[0, 0] == pytest.approx(np.cos(ph) * np.cos(ka))
    assert R[0, 1] == pytest.approx(-np.cos(ph) * np.sin(ka))
    assert R[1, 2] == pytest.approx(-np.sin(om) * np.cos(ph))
    assert R[2, 2] == pytest.approx(np.cos(om) * np.cos(ph))
    assert R[1, 0] == pytest.approx(np.cos(om) * np.sin(ka) + np.sin(om) * np.sin(ph) * np.cos(ka))
    assert R[2, 1] == pytest.approx(np.sin(om) * np.cos(ka) + np.cos(om) * np.sin(ph) * np.sin(ka))


def _images(adj):
    return [img for cam in adj._cameras for img in cam]


@pytest.mark.gpu
@pytest.mark.parametrize('solver', ['dense', 'structured'])
def test_propagation_matches_oracle(built, solver):
    """sigma2 J Qxx J' on the device-resident Qxx vs the oracle's dense product, through the reference-shaped API."""
    scene = synthetic_scene(2, images=12, targets=80)[0]
    scene['cameras'][0]['images'][3]['eo_fixed'][5] = True          # a fixed kappa: its column contributes nothing
    adj, pts = build_adjustment(scene)
    adj.setSolver({'dense': ba._lib.SOLVER_DENSE, 'structured': ba._lib.SOLVER_STRUCTURED}[solver])
    assert adj.estimateModel() == ba.EstimationStateType.ERROR_FREE_ESTIMATION
    imgs = _images(adj)
    coords = [pts[i] for i in range(0, 40, 2)]
    align = {imgs[0]: [imgs[0], imgs[1], imgs[3]], imgs[5]: [imgs[6], imgs[3]]}
    sigma2 = adj.getVarianceFactorAposteriori()
    t = ba.CoordinateTransformationExteriorOrientation.getInstance()
    t.transform(coords, align, sigma2, adj.getCofactorMatrix())
    C = t.getCovarianceMatrix().toDense()
    got = np.array([[c.getX().getValue(), c.getY().getValue(), c.getZ().getValue()] for c in t.getTransformedCoordinates()])
    # oracle on its own adjusted values and its own Qxx
    o = Oracle(scene)
    assert o.estimate() == 1
    triples = []
    obs = [set(np.asarray(im['obj']).tolist()) for im in scene['cameras'][0]['images']]
    for ref, lst in ((0, [0, 1, 3]), (5, [6, 3])):
        for s in lst:
            for p in range(0, 40, 2):
                if p in obs[s]:
                    triples.append((p, s, ref))
    assert len(triples) == got.shape[0] > 50
    xyz_o, C_o = op.propagate(o.fp.xyz, o.fp.pt_col, o.fp.eo_val, o.fp.eo_col, triples, o.variance_factor_aposteriori(), o.qxx_dense())
    np.testing.assert_allclose(got, xyz_o, rtol=1e-10, atol=1e-9)
    sd = np.sqrt(np.diag(C_o))
    err = (np.abs(C - C_o) / np.outer(sd, sd)).max()
    print('propagation [%s]: %d transformed points, scaled covariance error %.2e' % (solver, len(triples), err))
    assert err <= 1e-8
    assert np.abs(C - C.T).max() == 0.0
    names = [c.getName() for c in t.getTransformedCoordinates()]
    assert names[0] == '%s %s %s' % (pts[0].getName(), imgs[0].getId(), imgs[0].getId())
    # identity rows (source image == reference image): the covariance block is the point's own block of sigma2 Qxx
    Q = adj.getCofactorMatrix().toDense()
    c0 = [pts[0].getX().getColumn(), pts[0].getY().getColumn(), pts[0].getZ().getColumn()]
    np.testing.assert_allclose(C[:3, :3], sigma2 * Q[np.ix_(c0, c0)], rtol=1e-14)


@pytest.mark.gpu
def test_propagation_needs_cofactor_matrix(built):
    scene = synthetic_scene(2, images=8, targets=60)[0]
    adj, pts = build_adjustment(scene)
    adj.setInvertNormalEquation(ba.MatrixInversion.NONE)
    assert adj.estimateModel() == ba.EstimationStateType.ERROR_FREE_ESTIMATION
    with pytest.raises(ba.JaicovError) as e:
        adj._session.propagate_eo_transform([0], [0], [1], 1.0)
    assert e.value.code == ba._lib.NOT_INITIALISED
    with pytest.raises(ba.JaicovError):
        adj._session.propagate_eo_transform([10 ** 6], [0], [1], 1.0, covariance=False)
    xyz, _ = adj._session.propagate_eo_transform([0], [2], [2], 1.0, covariance=False)     # coordinates only: no Qxx needed
    np.testing.assert_array_equal(xyz[0], adj._session.values()[0].reshape(-1, 3)[0])
