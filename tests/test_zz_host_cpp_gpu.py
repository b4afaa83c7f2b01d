"""GPU: estimateModel() of the native C++ host mirror (jaicov_host.hpp -> C ABI -> CUDA) against the Python mirror on the same
networks -- same state, sigma0^2, parameters and cofactor matrix (both drive the same library; the two mirrors differ only in how
they number the object points, tests/test_host_cpp.py) -- and against the CPU oracle.
(Sorted last on purpose: written in a session without GPU access; the CPU tests pin everything this path hands to the library.)"""
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
        sc = random_scene(2)
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
