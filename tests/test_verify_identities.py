"""bundle_adjustment_b200.verify on the CPU: the identities it checks (K [lambda; dx] = [0; n], K Qxx e_c = e_c on the preconditioned
system, Omega = w'Pw - n'dx) hold for the ORACLE's results to ~eps * cond, and a corrupted solution / cofactor column / Omega is
caught.  A stand-in session answers the four calls check_pass makes (dx, preconditioner, qxx_block, normal_product) from the oracle's
dense matrices; on the GPU the same function runs against jaicov_normal_product (tests/test_gpu_fullsize.py, bench.py)."""
import numpy as np
import pytest

from bundle_adjustment_b200 import verify
from oracle.oracle import Oracle, lib
from tests.scenes import example_scene, synthetic_scene


class OracleSession:
    """The last final pass of an oracle adjustment, in the shape of bundle_adjustment_b200.Session."""

    def __init__(self, scene):
        o = Oracle(scene)
        assert o.estimate() == 1
        self.o = o
        self.n = n = o.fp.n
        self.flat = {'n_unknowns': n - o.fp.d}
        if o.use_centroid:
            o._centroid(False)                      # the device works on centred values
        N, nv, V = o.create_normal_equation()       # at the adjusted values
        self.K = np.empty((n, n))
        lib().orc_unpack(n, N.ctypes.data, self.K.ctypes.data)
        self.rhs, self.V = nv.copy(), V.copy()
        self.V[:o.fp.d] = 1.0
        self.sol = np.linalg.solve(self.K, self.rhs)
        self.Q = o.qxx_dense()
        self.wpw = o.get_omega(np.zeros(n))
        self.omega = o.get_omega(self.sol)

    def dx(self): return self.sol.copy()
    def preconditioner(self): return self.V.copy()
    def qxx_block(self, r0, r1, c0, c1): return self.Q[r0:r1, c0:c1].copy()

    def normal_product(self, X):
        X = np.atleast_2d(X)
        return X @ self.K, self.rhs.copy(), self.wpw


@pytest.fixture(scope='module', params=['example', 'config2'])
def sess(request):
    return OracleSession(example_scene() if request.param == 'example' else synthetic_scene(2, images=12, targets=70)[0])


def test_identities_hold_for_the_oracle(sess):
    chk = verify.check_pass(sess, omega=sess.omega)
    verify.assert_ok(chk, tol_solve=1e-9, tol_cofactor=1e-9, tol_omega=1e-9)
    assert len(chk['columns']) >= 8 and 0 in chk['columns'] and sess.n - 1 in chk['columns']


def test_corruptions_are_caught(sess):
    good = verify.check_pass(sess, omega=sess.omega)
    d = sess.n - sess.flat['n_unknowns']
    # one entry of dx off by a millionth of its value
    s0 = sess.sol.copy()
    k = d + int(np.argmax(np.abs(s0[d:])))
    sess.sol[k] *= 1 + 1e-6
    with pytest.raises(AssertionError, match='solve_residual'):
        verify.assert_ok(verify.check_pass(sess, omega=sess.omega))
    sess.sol = s0
    # one entry of a sampled cofactor column off by 1e-6 relative
    c = good['columns'][len(good['columns']) // 2]
    q0 = sess.Q[c, c]
    sess.Q[c, c] = q0 * (1 + 1e-6)
    with pytest.raises(AssertionError, match='cofactor_residual'):
        verify.assert_ok(verify.check_pass(sess, omega=sess.omega))
    sess.Q[c, c] = q0
    with pytest.raises(AssertionError, match='omega_rel_diff'):
        verify.assert_ok(verify.check_pass(sess, omega=sess.omega * (1 + 1e-6)))
    verify.assert_ok(verify.check_pass(sess, omega=sess.omega), tol_solve=1e-9, tol_cofactor=1e-9, tol_omega=1e-9)


def test_sample_columns_cover_panel_boundaries_and_ranks():
    cols = verify.sample_columns(63020, 7, world=8)
    assert {0, 7, 63019, 7 + 1023, 7 + 1024, 7 + 127, 7 + 128} <= set(cols.tolist())
    for r in range(8):
        assert 7 + 128 * r + 7 in cols
    assert len(cols) >= 16 and np.all(np.diff(cols) > 0)
    assert verify.sample_columns(5, 0).max() == 4
