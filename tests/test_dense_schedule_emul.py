"""Index logic of the blocked factor/solve/invert schedule (csrc/dense_driver.hpp), run with the TEST-ONLY host
backend tests/emul/host_backend.cpp: recursion splits, triangular k-ranges, in-place strips, out-of-place trtri,
single-launch lauum.  Everything the schedule must never read is poisoned with NaN."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def emul(built):
    L = ctypes.CDLL(os.path.join(ROOT, 'tests', '_build', 'libemul.so'))
    L.emul_spd_solve_invert.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    return L


@pytest.mark.parametrize('nb', [1, 2, 3, 5, 6])
def test_schedule_matches_lapack(emul, nb):
    rng = np.random.default_rng(nb)
    n = 128 * nb
    A = rng.standard_normal((n, n))
    S = A @ A.T + n * np.eye(n)
    dd = 1 / np.sqrt(np.diag(S))
    S = S * dd[:, None] * dd[None, :]
    M = np.tril(S).copy()
    R = np.zeros((128, n))
    R[:4] = rng.standard_normal((4, n))
    R0 = R.copy()
    st = np.zeros(3)
    info = emul.emul_spd_solve_invert(n, M.ctypes.data, 1, R.ctypes.data, 1, st.ctypes.data)
    assert info == 0
    assert not np.isnan(np.tril(M)).any() and not np.isnan(R).any()
    Qi = np.linalg.inv(S)
    np.testing.assert_allclose(np.tril(M), np.tril(Qi), atol=1e-12 * np.abs(Qi).max())
    np.testing.assert_allclose(R[:4], np.linalg.solve(S, R0[:4].T).T, rtol=0, atol=1e-12 * np.abs(R0).max())
    assert st[1] == nb


def test_not_positive_definite_is_reported(emul):
    n = 256
    S = np.eye(n)
    S[200, 200] = -1.0
    M = np.tril(S).copy()
    R = np.zeros((128, n))
    assert emul.emul_spd_solve_invert(n, M.ctypes.data, 0, R.ctypes.data, 0, None) == 201
