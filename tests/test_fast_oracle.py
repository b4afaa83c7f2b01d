"""The blocked "fast oracle" (oracle/fast_oracle.py: dpotrf / dpotri on M = N + B'B with the datum identity)
against the faithful oracle (dspsv + dsptri, the routines the reference calls, MathExtension.java:338-366) on everything the
faithful one can finish: it may then stand in for it at the full sizes of BASELINE.json's configs[2] and configs[3]
(tests/test_gpu_fullsize.py)."""
import numpy as np
import pytest

from oracle.fast_oracle import FastOracle
from oracle.oracle import Oracle
from tests.scenes import example_scene, random_scene, synthetic_scene

SCENES = {
    'example_scale_bar': example_scene,                                                    # configs[0], d = 6, one scale bar
    'config2_free_network': lambda: synthetic_scene(2, images=20, targets=120)[0],         # d = 7
    'config3_dense_dispersion': lambda: synthetic_scene(3, images=8, targets=60)[0],       # d = 0, r = 180 fully populated
    'config4_distance_terms': lambda: synthetic_scene(4, images=12, targets=80)[0],
    'fixed_points_datum': lambda: synthetic_scene(2, images=10, targets=60, free_network=False)[0],   # d = 0, plain Cholesky
    'random_11': lambda: random_scene(11),
}


@pytest.mark.parametrize('name', sorted(SCENES))
def test_fast_oracle_equals_faithful_oracle(name):
    kw = dict(use_centroid=False) if name.startswith('random') else {}    # fixed components: the centroid shift refuses (BA:151)
    a, b = Oracle(SCENES[name](), **kw), FastOracle(SCENES[name](), **kw)
    sa, sb = a.estimate(), b.estimate()
    assert sa == sb == 1
    assert a.iterations == b.iterations and len(a.history) == len(b.history)
    Qa, Qb = a.qxx_dense(), b.qxx_dense()
    d = a.fp.d
    sc = np.sqrt(np.abs(np.diag(Qa)))
    sc[:d] = 1.0
    # correlation-scaled, border block included.  Two backward-stable algorithms differ by ~ eps * cond(VKV): 1e-11 on most of
    # these networks, 1.1e-10 on the small config-4 one (cond 2.4e8); the parity bar of the GPU tests is 1e-8
    assert np.max(np.abs(Qa - Qb) / np.outer(sc, sc)) <= 1e-9
    s2a, s2b = a.variance_factor_aposteriori(), b.variance_factor_aposteriori()
    assert abs(s2a - s2b) <= 1e-11 * s2a
    sig = np.sqrt(s2a * np.abs(np.diag(Qa)))
    for va, vb, cols in ((a.fp.xyz, b.fp.xyz, a.fp.pt_col), (a.fp.io_val, b.fp.io_val, a.fp.io_col),
                         (a.fp.coef_val, b.fp.coef_val, a.fp.coef_col), (a.fp.eo_val, b.fp.eo_val, a.fp.eo_col)):
        c = cols.astype(np.int64)
        act = (c >= 0) & (c < 2147483647)
        tol = np.maximum(1e-12 * np.abs(va[act]), 1e-6 * sig[c[act]])      # a millionth of the parameter's own standard deviation
        assert np.all(np.abs(va[act] - vb[act]) <= tol)


def test_fast_oracle_refuses_a_singular_system():
    sc = synthetic_scene(2, images=6, targets=30)[0]
    sc['points']['datum'][:] = False                                       # free network without datum points
    with pytest.raises(ValueError):
        FastOracle(sc).estimate()
