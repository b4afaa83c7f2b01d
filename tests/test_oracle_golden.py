"""Pins the CPU oracle on the reference's only known answers: the AICON 3D Studio report bundled with JAICOV
(JAICOV/example/example.htm, read through tests/golden/example_scene.npz) and the survey-time probe values of
SURVEY.md section 8(c)."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.scenes import example_scene


@pytest.fixture(scope='module')
def example(built):
    sc = example_scene()
    o = Oracle(sc)
    status = o.estimate()
    return sc, o, status


def test_bookkeeping_counts(example):
    _, o, _ = example
    # example.htm:31-35,39-42 : n = 19945, u = 1147, b = 6, redundancy 18804
    assert o.bk.n_obs == 19945
    assert o.bk.n_unknown == 1147
    assert o.bk.d == 6
    assert o.bk.dof == 18804
    assert o.bk.defect_free == (True, True, True, True, True, True, False)   # scale bar fixes the scale, BA:849
    assert o.sigma2apriori == 2.5e-7


def test_convergence_history(example):
    _, o, status = example
    assert status == 1
    assert len(o.history) == 4
    np.testing.assert_allclose(o.history[0], 9.451e-5, rtol=2e-4)
    np.testing.assert_allclose(o.history[1], 3.395e-8, rtol=2e-4)
    assert o.history[2] < 1.0536712127723509e-8 and o.history[3] < 1e-11


def test_variance_of_unit_weight_matches_aicon(example):
    _, o, _ = example
    s2 = o.variance_factor_aposteriori()
    np.testing.assert_allclose(o.omega, 0.0030898518985, rtol=1e-9)
    np.testing.assert_allclose(s2, 1.6431886293e-7, rtol=1e-9)
    s0 = 0.0005 * np.sqrt(s2 / o.sigma2apriori)
    assert abs(s0 - 0.000405) < 5e-7                                       # example.htm:31  S0 = 0.000405


def test_interior_orientation_matches_aicon(example):
    _, o, _ = example
    fp = o.fp
    x0, y0, c = fp.io_val
    # example.htm:77-87 (Ck = -c)
    assert abs(c - 28.78507) < 5e-5 and abs(x0 - 1.734892e-2) < 5e-7 and abs(y0 - 5.668731e-2) < 5e-7
    np.testing.assert_allclose(c, 28.785073317293, rtol=1e-10)
    np.testing.assert_allclose(x0, 0.017348775491, rtol=1e-8)
    cx, cy, bx, by, a1, a2, a3 = fp.coef_val
    np.testing.assert_allclose([a1, a2, bx, by], [-1.096069e-4, 1.495660e-7, 5.798428e-6, -8.644540e-6], rtol=2e-5)
    assert a3 == 0.0 and cx == -7.008010e-5 and cy == -3.126270e-5              # fixed ("fest")


def test_io_standard_deviations_and_correlations_match_aicon(example):
    _, o, _ = example
    fp = o.fp
    Q = o.qxx_dense() * o.variance_factor_aposteriori()
    cols = [fp.io_col[2], fp.io_col[0], fp.io_col[1], fp.coef_col[4], fp.coef_col[5], fp.coef_col[2], fp.coef_col[3]]
    sig = np.sqrt(np.diag(Q)[cols])
    aicon = [2.513178e-4, 3.441658e-4, 3.262600e-4, 2.978787e-8, 7.655524e-11, 1.190972e-7, 1.043919e-7]   # :78-85
    np.testing.assert_allclose(sig, aicon, rtol=1e-5)
    C = Q[np.ix_(cols, cols)] / np.outer(sig, sig)
    corr = np.array([[1, 0, 0, 0, 0, 0, 0], [0.240, 1, 0, 0, 0, 0, 0], [-0.555, -0.191, 1, 0, 0, 0, 0],
                     [-0.304, -0.131, 0.206, 1, 0, 0, 0], [0.184, 0.082, -0.127, -0.909, 1, 0, 0],
                     [0.190, 0.939, -0.179, -0.187, 0.097, 1, 0], [-0.376, -0.222, 0.800, 0.302, -0.138, -0.257, 1]])   # :91-97
    sign = np.array([-1, 1, 1, 1, 1, 1, 1])      # AICON reports Ck = -c
    for i in range(7):
        for j in range(i):
            assert abs(C[i, j] * sign[i] * sign[j] - corr[i, j]) < 6e-4


def test_object_point_matches_aicon(example):
    sc, o, _ = example
    i6 = sc['points']['names'].index('6')
    np.testing.assert_allclose(o.fp.xyz[3 * i6:3 * i6 + 3], [573.0039, -49.4291, -121.6922], atol=6e-5)   # example.obc / :1607
    np.testing.assert_allclose(o.fp.xyz[3 * i6:3 * i6 + 3], [573.00385393, -49.42908891, -121.69215192], atol=2e-8)


def test_border_block_of_inverse_vanishes(example):
    _, o, _ = example
    Q = o.qxx_dense()
    assert np.abs(Q[:6, :6]).max() < 1e-12


@pytest.mark.parametrize('mode', ['REDUCED', 'PRE_ELIMINATION'])
def test_reduced_modes_agree_with_full(example, mode):
    """BA:1197-1453 restated: the reduced system's cofactor matrix is the leading block of the full inverse, and the
    parameters / sigma0^2 agree with the FULL run far inside the parity tolerances (ExampleReport itself uses REDUCED)."""
    sc, full, _ = example
    o = Oracle(example_scene(), invert=mode)
    assert o.estimate() == 1 and len(o.history) == 4
    nr = o.num_rows_reduced()
    assert nr == 3 + 4 + 3 * 150 + 6                         # nIO + nDist + 3 |objectCoordinates| + d, BA:262
    Qf, Qr = full.qxx_dense()[:nr, :nr], o.qxx_dense()[:nr, :nr]
    sg = np.sqrt(np.abs(np.diag(Qf)))
    sg[:6] = 1.0
    assert (np.abs(Qr - Qf) / np.outer(sg, sg)).max() < 1e-9
    assert abs(o.variance_factor_aposteriori() / full.variance_factor_aposteriori() - 1) < 1e-11
    assert np.abs(o.fp.xyz - full.fp.xyz).max() < 1e-10 and np.abs(o.fp.eo_val - full.fp.eo_val).max() < 1e-10
