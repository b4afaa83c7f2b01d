"""GPU: estimateModel() of the native C++ host mirror (jaicov_host.hpp -> C ABI -> CUDA) against the Python mirror on the same
networks -- same state, sigma0^2, parameters and cofactor matrix (both drive the same library; the two mirrors differ only in how
they number the object points, tests/test_host_cpp.py) -- and against the CPU oracle.
(Ran on a B200 for the first time in round 2: four of five passed as written, the fifth had an invalid scene -- see _scene.)"""
import ctypes

import numpy as np
import pytest

from tests.test_host_cpp import H, Net, _p   # noqa: F401  (H is the fixture)

pytestmark = pytest.mark.gpu


def _scene(name):
    from tests.scenes import random_scene, synthetic_scene
    if name == 'free_network':
        return synthetic_scene(2, images=10, targets=60)[0]            # structured route
    if name == 'scale_bar':
        # (round 1 used random_scene(2) here: it has fixed point components, for which centroidCoordinates refuses -- BA:151 -- in the
        # reference as in both mirrors, so the test could never pass; first seen when the module finally ran on a B200)
        sc = synthetic_scene(2, images=10, targets=60)[0]
        xyz = sc['points']['xyz']
        sc['scale_bars'] = [(0, 1, float(np.linalg.norm(xyz[0] - xyz[1])) + 0.01, 0.02), (5, 9, float(np.linalg.norm(xyz[5] - xyz[9])) - 0.02, 0.05)]
        return sc
    return synthetic_scene(3, images=6, targets=40)[0]                  # observed points with a fully populated dispersion


@pytest.mark.parametrize('name', ['free_network', 'scale_bar', 'observed_points'])
def test_cpp_host_estimate_matches_python_mirror_and_oracle(H, name):
    import bundle_adjustment_b200 as ba
    from oracle.oracle import Oracle
    from tests.helpers import build_adjustment
    adj, pts = build_adjustment(_scene(name))
    state_py = adj.estimateModel()
    net = Net(H, _scene(name))
    state = ctypes.c_int(0)
    net.ok(H.jhost_estimate(net.h, ctypes.byref(state)))
    assert state.value == state_py.getId() == 1
    stats = np.zeros(6)
    has = ctypes.c_int(0)
    n = adj.getNumberOfUnknownParameters() + adj.getNumberOfDatumConditions()
    q = np.zeros(n * (n + 1) // 2)
    net.ok(H.jhost_get_results(net.h, _p(stats), _p(q), ctypes.byref(has)))
    assert has.value == 1
    assert stats[3] == adj.getDegreeOfFreedom() and stats[4] == adj.stats.iterations
    np.testing.assert_allclose(stats[1], adj.getVarianceFactorAposteriori(), rtol=1e-12)
    Qpy = adj.getCofactorMatrix().getData()
    sc = np.sqrt(np.abs(adj.getCofactorMatrix().toDense().diagonal()))
    sc[:adj.getNumberOfDatumConditions()] = 1.0
    iu = np.triu_indices(n)
    scale = (sc[iu[0]] * sc[iu[1]])
    order = iu[0] + iu[1] * (iu[1] + 1) // 2
    assert np.max(np.abs(q[order] - Qpy[order]) / scale) <= 1e-10
    xyz = np.zeros(3 * net.n_pt)
    net.ok(H.jhost_get_columns(net.h, None, None, None, None, _p(xyz), None, None, None))
    np.testing.assert_allclose(xyz.reshape(-1, 3), pts.xyz, rtol=1e-12, atol=1e-9)
    # and the oracle (the parity bars of tests/test_gpu_parity.py)
    orc = Oracle(_scene(name))
    assert orc.estimate() == 1
    s2o = orc.variance_factor_aposteriori()
    assert abs(stats[1] - s2o) <= 1e-8 * s2o
    net.close()


def test_tile_product_stage_entry_point(built):
    """jaicov_gemm_tiles (the FP64 tensor-core tile kernel behind factor / inverse, csrc/dense_kernels.cu: k_gemm) on small tile
    grids: every operand layout, triangular-operand hint and symmetric output, alpha / beta, against numpy -- with NaN wherever
    the kernel must not read.  The same case list passes on the CPU against the host emulation (tests/test_ozaki_emulation.py)."""
    import os
    import sys
    import bundle_adjustment_b200 as ba
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    import ozaki_gpu_check as chk
    cases = chk.run_gemm_cases(lambda As, Bs, C0, al, bl, alpha, beta, tri, kmode: ba._lib.gemm_tiles(As, Bs, C0, al, bl, alpha, beta, tri, kmode)[0])
    assert len(cases) >= 20
    bad = [c for c in cases if not c['ok']]
    assert not bad, bad[:3]


def test_cpp_host_dlt_and_transform_match_python_mirror(H):
    """The callers either side of the path through the C++ mirror (DirectLinearTransformation.adjustAll -> jaicov_dlt_batch,
    CoordinateTransformationExteriorOrientation.transform -> jaicov_propagate_eo_transform) give what the Python mirror gives."""
    import bundle_adjustment_b200 as ba
    from tests.helpers import build_adjustment
    from tests.scenes import synthetic_scene
    scene = synthetic_scene(2, images=6, targets=40)[0]
    adj, pts = build_adjustment(scene)
    images = [img for cam in adj.getCameras() for img in cam]
    # DLT on the initial values, all points known
    coefs = [ba.DLTCoefficients(img) for img in images]
    ok_py = ba.DirectLinearTransformation.adjustAll(coefs, {pts.names[i]: pts[i] for i in range(40)})
    net = Net(H, scene)
    kn = np.arange(40, dtype=np.int32)
    out, ok = np.zeros(20 * net.n_img), np.zeros(net.n_img, np.uint8)
    net.ok(H.jhost_dlt(net.h, 40, _p(kn), 0, None, 1, None, None, None, None, _p(out), _p(ok)))
    assert [bool(v) for v in ok] == ok_py
    vals_py = np.array([[p.getValue() for p in c] for c in coefs])
    np.testing.assert_allclose(out.reshape(-1, 20), vals_py, rtol=1e-12, atol=1e-12)
    # adjustment, then the transformation of five points seen in images 1..3 into the frame of image 0
    assert adj.estimateModel().getId() == 1
    state = ctypes.c_int(0)
    net.ok(H.jhost_estimate(net.h, ctypes.byref(state)))
    assert state.value == 1
    s2 = adj.getVarianceFactorAposteriori()
    t = ba.CoordinateTransformationExteriorOrientation.getInstance()
    t.transform([pts[i] for i in range(5)], {images[0]: images[1:4]}, s2, adj.getCofactorMatrix())
    xyz_py = np.array([[c.getX().getValue(), c.getY().getValue(), c.getZ().getValue()] for c in t.getTransformedCoordinates()])
    cov_py = t.getCovarianceMatrix().getData()
    r = xyz_py.shape[0]
    pts_i, imgs_i = np.arange(5, dtype=np.int32), np.arange(1, 4, dtype=np.int32)
    n_out = ctypes.c_int(0)
    xyz, cov = np.zeros(3 * r), np.zeros(3 * r * (3 * r + 1) // 2)
    net.ok(H.jhost_transform(net.h, 5, _p(pts_i), 0, 3, _p(imgs_i), ctypes.c_double(s2), r, ctypes.byref(n_out), _p(xyz), _p(cov)))
    assert n_out.value == r == 15
    np.testing.assert_allclose(xyz.reshape(-1, 3), xyz_py, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(cov, cov_py, rtol=1e-9, atol=1e-18)
    # DefaultResultWriter: <base>.info and <base>.cxx (sigma0^2 * Qxx of the object coordinates, gathered on the device)
    import os
    import tempfile
    from bundle_adjustment_b200.writers import DefaultResultWriter
    with tempfile.TemporaryDirectory() as tmp:
        DefaultResultWriter(os.path.join(tmp, 'py')).export(adj)
        net.ok(H.jhost_export_default(net.h, os.path.join(tmp, 'cpp').encode()))
        info_py = [l.split('\t') for l in open(os.path.join(tmp, 'py.info'))]
        info_cpp = [l.split('\t') for l in open(os.path.join(tmp, 'cpp.info'))]
        assert [(a[0], a[1], a[3]) for a in info_py] == [(a[0], a[1], a[3]) for a in info_cpp]
        np.testing.assert_allclose([float(a[2]) for a in info_cpp], [float(a[2]) for a in info_py], rtol=1e-12, atol=1e-9)
        C_py, C_cpp = np.loadtxt(os.path.join(tmp, 'py.cxx')), np.loadtxt(os.path.join(tmp, 'cpp.cxx'))
        assert C_py.shape == C_cpp.shape == (120, 120)
        np.testing.assert_allclose(C_cpp, C_py, rtol=1e-9, atol=1e-14)
    net.close()
