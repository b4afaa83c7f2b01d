"""EXTENSION: per-image fully populated dispersion of the image coordinates (north_star (2), BASELINE.json configs[3]).  The reference
cannot express it (camera/ImageCoordinate.java:102-104: a group is the two rows of one point), so parity is UNPINNED except where the
matrix is block diagonal -- there it must reproduce the reference-expressible sigma / rho path, which the first tests require of the
oracle extension (CPU) and of the CUDA path (GPU, against the FAITHFUL oracle).  Genuinely dense matrices are compared with the oracle
extension (oracle/dense_image_sigma.py) and checked through the matrix-free identities of bundle_adjustment_b200.verify."""
import numpy as np
import pytest

from oracle.dense_image_sigma import DenseImageSigmaOracle
from oracle.oracle import Oracle
from tests.scenes import synthetic_scene


def _pack(S):
    n = S.shape[0]
    iu = np.triu_indices(n)
    out = np.empty(n * (n + 1) // 2)
    out[iu[0] + iu[1] * (iu[1] + 1) // 2] = S[iu]
    return out


def scene_with_dispersions(images_dense, common_mode, seed=3, images=8, targets=40):
    """config-2 style network with correlated image coordinates; the listed images get their dispersion as a dense matrix:
    blockdiag(2 x 2 from sigma, rho) + common_mode * sigma^2 * (u u' + v v') with u / v = indicator of the x / y rows
    (a common-mode error of all coordinates of the image; 0 gives the block-diagonal case)."""
    rng = np.random.default_rng(seed)
    sc = synthetic_scene(2, images=images, targets=targets, seed=77)[0]
    k = 0
    for cam in sc['cameras']:
        for im in cam['images']:
            m = len(im['obj'])
            im['rho'] = rng.uniform(-0.5, 0.5, size=m)
            im['sigma'] = im['sigma'] * rng.uniform(0.8, 1.3, size=(m, 2))
            if k in images_dense:
                S = np.zeros((2 * m, 2 * m))
                for q in range(m):
                    sx, sy = im['sigma'][q]
                    S[2 * q, 2 * q], S[2 * q + 1, 2 * q + 1] = sx * sx, sy * sy
                    S[2 * q, 2 * q + 1] = S[2 * q + 1, 2 * q] = im['rho'][q] * sx * sy
                if common_mode:
                    s2 = float(np.mean(im['sigma'] ** 2))
                    u = np.zeros(2 * m); u[0::2] = 1.0
                    v = np.zeros(2 * m); v[1::2] = 1.0
                    S += common_mode * s2 * (np.outer(u, u) + np.outer(v, v))
                im['dispersion'] = _pack(S)
            k += 1
    return sc


def strip(sc):
    for cam in sc['cameras']:
        for im in cam['images']:
            im.pop('dispersion', None)
    return sc


def compare(o, ref, tol_q=1e-9):
    Qa, Qb = o.qxx_dense(), ref.qxx_dense()
    d = ref.fp.d
    sg = np.sqrt(np.abs(np.diag(Qb)))
    sg[:d] = 1.0
    assert np.max(np.abs(Qa - Qb) / np.outer(sg, sg)) <= tol_q
    s2a, s2b = o.variance_factor_aposteriori(), ref.variance_factor_aposteriori()
    assert abs(s2a - s2b) <= 1e-10 * s2b
    for va, vb in ((o.fp.xyz, ref.fp.xyz), (o.fp.io_val, ref.fp.io_val), (o.fp.coef_val, ref.fp.coef_val), (o.fp.eo_val, ref.fp.eo_val)):
        np.testing.assert_allclose(va, vb, rtol=1e-10, atol=1e-10)


def test_block_diagonal_dispersion_reproduces_the_reference_path_on_the_cpu():
    sc = scene_with_dispersions({1, 4, 5}, 0.0)
    o = DenseImageSigmaOracle(sc)
    ref = Oracle(strip(scene_with_dispersions({1, 4, 5}, 0.0)))
    assert o.estimate() == ref.estimate() == 1
    assert len(o.history) == len(ref.history)
    compare(o, ref)


def test_a_common_mode_term_changes_the_result_on_the_cpu():
    a = DenseImageSigmaOracle(scene_with_dispersions({1, 4, 5}, 0.3))
    b = DenseImageSigmaOracle(scene_with_dispersions({1, 4, 5}, 0.0))
    assert a.estimate() == b.estimate() == 1
    Qa, Qb = a.qxx_dense(), b.qxx_dense()
    sg = np.sqrt(np.abs(np.diag(Qb)))
    sg[:a.fp.d] = 1.0
    assert np.max(np.abs(Qa - Qb) / np.outer(sg, sg)) > 1e-3          # measured: 7.6e-3 of a standard deviation
    assert np.max(np.abs(a.fp.eo_val - b.fp.eo_val)) > 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize('common_mode', [0.0, 0.3])
def test_gpu_dense_image_dispersion(built, common_mode):
    """CUDA path through the host mirror (Image.setDispersion -> jaicov_set_image_dispersion).  common_mode = 0: against the FAITHFUL
    oracle on the sigma / rho description of the same network (the reference-expressible case).  0.3: against the oracle extension."""
    import bundle_adjustment_b200 as ba
    from bundle_adjustment_b200 import verify
    from tests.helpers import build_adjustment
    dense = {0, 3, 6}
    adj, pts = build_adjustment(scene_with_dispersions(dense, common_mode))
    assert adj.estimateModel() == ba.EstimationStateType.ERROR_FREE_ESTIMATION
    assert adj.stats.solver_used == ba._lib.SOLVER_DENSE           # a dense dispersion couples the points of its image
    if common_mode == 0.0:
        o = Oracle(strip(scene_with_dispersions(dense, 0.0)))
    else:
        o = DenseImageSigmaOracle(scene_with_dispersions(dense, common_mode))
    assert o.estimate() == 1
    assert adj.stats.iterations == len(o.history)
    Qg, Qo = adj.getCofactorMatrix().toDense(), o.qxx_dense()
    d = o.fp.d
    sg = np.sqrt(np.abs(np.diag(Qo)))
    sg[:d] = 1.0
    errq = np.max(np.abs(Qg - Qo) / np.outer(sg, sg))
    s2g, s2o = adj.getVarianceFactorAposteriori(), o.variance_factor_aposteriori()
    print('dense image dispersion (common mode %.1f): scaled Qxx err %.2e, sigma0^2 rel err %.2e' % (common_mode, errq, abs(s2g - s2o) / s2o))
    assert errq <= 1e-8 and abs(s2g - s2o) <= 1e-8 * s2o
    assert abs(adj.stats.omega - o.omega) <= 1e-8 * o.omega
    xyz_g = adj._session.values()[0]
    np.testing.assert_allclose(xyz_g, o.fp.xyz, rtol=1e-10, atol=1e-9)
    chk = verify.check_pass(adj._session, omega=adj.stats.omega, values_updated=True)
    verify.assert_ok(chk)
    # a pass at the STARTING values: the solution identity as well (nothing updated; at convergence dx and n are rounding noise and
    # their residual is not a meaningful ratio)
    from tests.helpers import flat_problem
    adj0, flat = flat_problem(scene_with_dispersions(dense, common_mode))
    s = ba.Session(sigma2apriori=adj0.getVarianceFactorApriori())
    s.set_problem(flat)
    assert s.iterate(final_pass=True, apply_update=False) == 0
    verify.assert_ok(verify.check_pass(s, omega=s.stats().omega))
    s.close()


@pytest.mark.gpu
def test_gpu_dense_dispersion_on_every_image_midsize(built):
    """Every image of a 12 x 300 network carries a fully populated 600 x 600 dispersion (block diagonal + common mode): the point block
    of N is then dense (every pair of points is coupled through every image), 12 device Cholesky + inverse of the dispersions, the
    tensor-core products T = P Ac at 640 rows.  Against the oracle extension and the matrix-free identities."""
    import bundle_adjustment_b200 as ba
    from bundle_adjustment_b200 import verify
    from tests.helpers import build_adjustment
    dense = set(range(12))
    mk = lambda: scene_with_dispersions(dense, 0.3, images=12, targets=300)
    adj, _pts = build_adjustment(mk())
    assert adj.estimateModel() == ba.EstimationStateType.ERROR_FREE_ESTIMATION
    o = DenseImageSigmaOracle(mk())
    assert o.estimate() == 1 and adj.stats.iterations == len(o.history)
    Qg, Qo = adj.getCofactorMatrix().toDense(), o.qxx_dense()
    sg = np.sqrt(np.abs(np.diag(Qo)))
    sg[:o.fp.d] = 1.0
    errq = np.max(np.abs(Qg - Qo) / np.outer(sg, sg))
    s2g, s2o = adj.getVarianceFactorAposteriori(), o.variance_factor_aposteriori()
    print('dense dispersion on all 12 images (n = %d): scaled Qxx err %.2e, sigma0^2 rel err %.2e, last pass %.2f ms'
          % (o.fp.n, errq, abs(s2g - s2o) / s2o, adj.stats.ms_total))
    assert errq <= 1e-8 and abs(s2g - s2o) <= 1e-8 * s2o
    verify.assert_ok(verify.check_pass(adj._session, omega=adj.stats.omega, values_updated=True))
