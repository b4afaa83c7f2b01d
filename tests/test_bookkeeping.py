"""Integer bookkeeping (rows, columns, d, dof): the product's host mirror (bundle-adjustment_b200/host.py) against the
oracle's independent restatement (oracle/bookkeeping.py) of BundleAdjustment.prepareUnknownParameters (BA:667-782)
and detectRankDefect (BA:836-1042).  Bit-exact."""
import numpy as np
import pytest

from oracle.bookkeeping import Bookkeeping
from tests.helpers import flat_problem
from tests.scenes import example_scene, synthetic_scene


def compare(scene):
    adj, flat = flat_problem(scene)
    bk = Bookkeeping(scene)
    assert flat['n_unknowns'] == bk.n_unknown
    assert flat['n_observations'] == bk.n_obs
    assert tuple(bool(x) for x in flat['free_flags']) == tuple(bk.defect_free)
    assert adj.getDegreeOfFreedom() == bk.dof
    assert adj.getVarianceFactorApriori() == bk.sigma2apriori
    np.testing.assert_array_equal(flat['pt_col'].reshape(-1, 3), bk.pt_col)
    np.testing.assert_array_equal(flat['io_col'], np.concatenate(bk.io_col))
    np.testing.assert_array_equal(flat['coef_col'], np.concatenate(bk.coef_col))
    np.testing.assert_array_equal(flat['eo_col'], np.concatenate(bk.eo_col))
    return adj, flat, bk


def test_example(built):
    _, flat, bk = compare(example_scene())
    assert bk.d == 6 and bk.n_unknown == 1147


def test_free_network(built):
    sc, _ = synthetic_scene(2, images=8, targets=40)
    _, _, bk = compare(sc)
    assert bk.d == 7


def test_sparse_visibility_first_appearance_order(built):
    sc, _ = synthetic_scene(2, images=9, targets=60, visibility=0.5)
    compare(sc)


def test_observed_points_fix_the_datum(built):
    sc, _ = synthetic_scene(3, images=6, targets=20)
    _, _, bk = compare(sc)
    assert bk.d == 0


def test_fixed_parameters_and_scale_bar(built):
    sc, _ = synthetic_scene(4, images=7, targets=30)
    sc['points']['fixed'][3] = [True, True, True]
    sc['points']['fixed'][5, 2] = True
    sc['cameras'][0]['io_fixed'][0] = True
    sc['cameras'][0]['coefs'][0] = sc['cameras'][0]['coefs'][0][:3] + (True,)
    sc['cameras'][0]['images'][2]['eo_fixed'][3] = True
    sc['scale_bars'] = [(1, 2, 100.0, 0.01)]
    _, _, bk = compare(sc)
    assert bk.d < 7


def test_rank_defect_rules(built):
    # two fully fixed points -> translations, scale fixed; rotations need cntY>=2 and cntZ>=2 etc. (BA:912-937)
    sc, _ = synthetic_scene(2, images=5, targets=20)
    sc['points']['fixed'][0] = True
    sc['points']['fixed'][1] = True
    _, _, bk = compare(sc)
    assert bk.defect_free == (False, False, False, False, False, False, False)
    sc, _ = synthetic_scene(2, images=5, targets=20)
    sc['points']['fixed'][0, 0] = True
    _, _, bk = compare(sc)
    assert bk.defect_free == (False, True, True, True, True, True, True)


def test_two_cameras(built):
    sc, _ = synthetic_scene(2, images=8, targets=30, n_cameras=2)
    compare(sc)


def test_unobserved_point_gets_no_column(built):
    sc, _ = synthetic_scene(2, images=5, targets=20)
    for im in sc['cameras'][0]['images']:
        keep = im['obj'] != 7
        for k in ('obj', 'xy', 'sigma', 'rho'):
            im[k] = im[k][keep]
    _, flat, bk = compare(sc)
    assert (bk.pt_col[7] == -1).all() and flat['is_datum'][7] == 0
