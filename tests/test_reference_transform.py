"""Row f-3 against the EXECUTED reference: tests/golden/reference_transform.npz holds what the reference's own
CoordinateTransformationExteriorOrientation.transform produced (tests/golden/make_transform_fixture.py) -- order and names of the
transformed points, their coordinates, and sigma2 * J Qxx J' in packed upper layout.  Checked here on the CPU: the oracle's product
(oracle/propagation.py) and the host mirror's enumeration of (point, source image, target image) triples, with the device call served by
the oracle.  The CUDA contraction itself is compared with the same oracle in tests/test_propagation.py (GPU)."""
import os

import numpy as np
import pytest

import bundle_adjustment_b200 as ba
from bundle_adjustment_b200.workloads import build_adjustment, synthetic_scene
from oracle import propagation as op

E = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_transform.npz'))
CASES = {   # as in make_transform_fixture.py
    'two_reference_images': (31, 6, 14, [0, 2, 3, 5, 7, 8, 11, 13], [(0, [0, 1, 2]), (3, [4, 5])]),
    'reference_image_among_its_images_last': (32, 5, 10, [1, 2, 4, 6, 9], [(2, [0, 1, 2])]),
    'single_pair': (33, 4, 8, [0, 1, 2, 3, 4, 5, 6, 7], [(1, [3])]),
}


def network(seed, images, targets):
    scene = synthetic_scene(2, images=images, targets=targets, seed=seed)[0]
    rng = np.random.default_rng(seed)
    for k, im in enumerate(scene['cameras'][0]['images']):
        keep = rng.random(len(im['obj'])) > (0.35 if k % 2 else 0.1)
        keep[:6] = True
        for key in ('obj', 'xy', 'sigma', 'rho'):
            if im.get(key) is not None:
                im[key] = np.asarray(im[key])[keep]
    return scene


def unpack(packed, n):
    Q = np.zeros((n, n))
    k = 0
    for c in range(n):
        Q[:c + 1, c] = packed[k:k + c + 1]
        k += c + 1
    return Q + np.triu(Q, 1).T


class OracleSession:
    """Serves jaicov_propagate_eo_transform from the oracle (the device contraction is compared with the same oracle on the GPU)."""

    def __init__(self, flat, Q):
        self.flat, self.Q = flat, Q

    def propagate_eo_transform(self, points, src, trg, sigma2):
        f = self.flat
        xyz, C = op.propagate(np.asarray(f['xyz'], float).ravel(), np.asarray(f['pt_col']).ravel(), np.asarray(f['eo_val'], float).ravel(),
                              np.asarray(f['eo_col']).ravel(), list(zip(points, src, trg)), sigma2, self.Q)
        n = C.shape[0]
        return xyz, np.concatenate([C[:c + 1, c] for c in range(n)])


@pytest.mark.parametrize('name', sorted(CASES))
def test_transform_matches_the_executed_reference(name):
    seed, images, targets, point_ids, align = CASES[name]
    g = lambda k: E['%s__%s' % (name, k)]
    adj, pts = build_adjustment(network(seed, images, targets))
    flat = adj._prepare()
    np.testing.assert_array_equal(np.asarray(flat['pt_col']).reshape(-1, 3), g('in_pt_col'))       # the same network as the fixture's
    np.testing.assert_array_equal(np.asarray(flat['eo_col']).reshape(-1, 6), g('in_eo_col'))
    np.testing.assert_array_equal(np.asarray(flat['eo_val']).reshape(-1, 6), g('in_eo_val'))
    n = int(flat['n_unknowns']) + int(np.sum(flat['free_flags']))
    Q = unpack(g('qxx_packed'), n)
    adj._session = OracleSession(flat, Q)
    imgs = [im for cam in adj.getCameras() for im in cam]
    CoVar = ba.UpperSymmPackMatrix(n, g('qxx_packed'))
    CoVar._adjustment = adj
    t = ba.CoordinateTransformationExteriorOrientation()
    t.transform([pts[p] for p in point_ids], {imgs[r]: [imgs[i] for i in lst] for r, lst in align}, float(g('sigma2')), CoVar)
    got = t.getTransformedCoordinates()
    assert [c.getName() for c in got] == [str(x) for x in g('names')]                              # visibility loops, order, names (:81-105)
    xyz = np.array([[c.getX().getValue(), c.getY().getValue(), c.getZ().getValue()] for c in got])
    np.testing.assert_allclose(xyz, g('xyz'), rtol=1e-13, atol=1e-9)
    cols = np.array([[c.getX().getColumn(), c.getY().getColumn(), c.getZ().getColumn()] for c in got])
    np.testing.assert_array_equal(cols, g('columns'))                                             # rows of J = columns of the result (:146-148)
    cov, ref = np.asarray(t.getCovarianceMatrix().getData()), g('covariance')
    assert cov.shape == ref.shape
    R = 3 * len(got)
    sd = np.sqrt(np.array([ref[c * (c + 1) // 2 + c] for c in range(R)]))
    scale = np.concatenate([sd[:c + 1] * sd[c] for c in range(R)])
    assert np.max(np.abs(cov - ref) / scale) < 1e-12                                               # sigma2 J Qxx J' (:107-114)


def test_reference_throws_for_fixed_parameters():
    """MTJ's index check makes the reference throw when a parameter of the transformation is FIXED (J.set(row, Integer.MAX_VALUE, ..));
    the library's "fixed parameters contribute nothing" (include/jaicov_b200.h) is defined behaviour beyond the reference."""
    assert bool(E['fixed_parameter_throws'])
