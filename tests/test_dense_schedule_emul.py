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


@pytest.mark.parametrize('nb,nranks,pw,merged', [(6, 2, 1, 1), (7, 3, 2, 1), (8, 4, 1, 1), (7, 3, 2, 0), (5, 1, 2, 1)])
def test_distributed_schedule_with_virtual_ranks(emul, nb, nranks, pw, merged):
    """Block-column-cyclic Cholesky (panel broadcasts, look-ahead, one trapezoid update launch per panel) on `nranks`
    virtual ranks (threads, one matrix replica each) followed by every rank's column-tile inverse: all replicas end
    with the same factor, and the gathered column tiles are the inverse."""
    emul.emul_distributed.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_int]
    rng = np.random.default_rng(100 + nb)
    n = 128 * nb
    A = rng.standard_normal((n, n))
    S = A @ A.T + n * np.eye(n)
    dd = 1 / np.sqrt(np.diag(S))
    S = S * dd[:, None] * dd[None, :]
    M = np.tril(S).copy()
    M[np.triu_indices(n, 1)] = np.nan
    for i in range(nb):
        M[i * 128:(i + 1) * 128, i * 128:(i + 1) * 128] = np.tril(S[i * 128:(i + 1) * 128, i * 128:(i + 1) * 128])
    Q = np.full((n, n), np.nan)
    md = ctypes.c_double(0)
    info = emul.emul_distributed(n, M.ctypes.data, nranks, pw, Q.ctypes.data, ctypes.byref(md), merged)
    assert info == 0 and md.value == 0.0
    np.testing.assert_allclose(np.tril(M), np.linalg.cholesky(S), atol=1e-13)
    Qi = np.linalg.inv(S)
    assert not np.isnan(np.tril(Q)).any()
    np.testing.assert_allclose(np.tril(Q), np.tril(Qi), atol=1e-12 * np.abs(Qi).max())


@pytest.mark.parametrize('mt', [1, 2, 7, 8, 9, 24, 61, 127])
@pytest.mark.parametrize('band', [0, 1, 3, 8, 16])
def test_tri_tile_order_is_a_bijection(emul, mt, band):
    """Every tile order of the lower-triangular launches (row by row, band-swizzled) visits each tile (it >= jt) exactly once,
    and inside the swizzled order a band's tiles are contiguous."""
    emul.emul_tri_tile_decode.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    it, jt = ctypes.c_int(), ctypes.c_int()
    seen = []
    for l in range(mt * (mt + 1) // 2):
        emul.emul_tri_tile_decode(l, mt, band, ctypes.byref(it), ctypes.byref(jt))
        assert 0 <= jt.value <= it.value < mt
        seen.append((it.value, jt.value))
    assert len(set(seen)) == len(seen) == mt * (mt + 1) // 2
    if band > 0:
        bands = [i // band for i, _ in seen]
        assert bands == sorted(bands)
    else:
        assert seen == sorted(seen)


@pytest.mark.parametrize('mt,nt', [(1, 1), (7, 3), (8, 8), (13, 64), (64, 5)])
@pytest.mark.parametrize('band', [0, 1, 8, 16])
def test_rect_tile_order_is_a_bijection(emul, mt, nt, band):
    """The band-swizzled order of the full (rectangular) launches of the int8-digit tile kernel visits every tile exactly once."""
    emul.emul_rect_tile_decode.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    seen = set()
    it, jt = ctypes.c_int(0), ctypes.c_int(0)
    for l in range(mt * nt):
        emul.emul_rect_tile_decode(l, mt, nt, band, ctypes.byref(it), ctypes.byref(jt))
        assert 0 <= it.value < mt and 0 <= jt.value < nt
        seen.add((it.value, jt.value))
    assert len(seen) == mt * nt


@pytest.mark.parametrize('nb,nranks,pw,ozaki', [(1, 1, 1, 0), (3, 2, 1, 0), (5, 2, 2, 0), (7, 3, 2, 0), (6, 4, 1, 0), (9, 2, 3, 0), (4, 8, 1, 0), (3, 2, 1, 8),
                                                (17, 8, 2, 0), (13, 4, 3, 0), (11, 8, 1, 0), (10, 3, 16, 0), (5, 3, 2, 8)])   # shapes of config 5 scaled down: more panels than ranks, narrow last panel, one panel wider than the matrix
def test_streamed_owner_only_schedule_with_virtual_ranks(emul, nb, nranks, pw, ozaki):
    """Owner-only storage (DenseSchedule::factor_solve_invert_streamed): every virtual rank holds ONLY its own block-column panels
    (the rest of its matrix is NaN), receives the others through a two-slot window, and still produces the solutions of the
    right-hand sides (identical on every rank) and its column tiles of the inverse -- against numpy."""
    n = 128 * nb
    rng = np.random.default_rng(100 + nb + 7 * nranks + pw)
    G = rng.standard_normal((n, n))
    S = G @ G.T + 0.1 * n * np.eye(n)
    dsc = 1 / np.sqrt(np.diag(S))
    S = S * dsc[:, None] * dsc[None, :]
    R = rng.standard_normal((3, n))
    M = np.ascontiguousarray(S)
    Rio = np.ascontiguousarray(R.copy())
    Q = np.full((n, n), np.nan)
    md, recv = ctypes.c_double(0), ctypes.c_long(0)
    emul.emul_streamed.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.POINTER(ctypes.c_long)]
    info = emul.emul_streamed(n, M.ctypes.data, nranks, pw, 3, Rio.ctypes.data, Q.ctypes.data, ctypes.byref(md), ozaki, ctypes.byref(recv))   # ozaki: every launch through the int8-digit emulation
    assert info == 0
    assert md.value == 0.0                                   # every rank computes the same solutions, bit for bit
    Sinv = np.linalg.inv(S)
    np.testing.assert_allclose(Rio, np.linalg.solve(S, R.T).T, rtol=0, atol=1e-10)
    il = np.tril_indices(n)
    assert np.isfinite(Q[il]).all()
    sc = np.sqrt(np.diag(Sinv))
    assert np.max(np.abs(Q[il] - Sinv[il]) / (sc[il[0]] * sc[il[1]])) < 1e-11
    npan = (nb + pw - 1) // pw
    if nranks > 1:
        assert recv.value <= 2 * npan                        # a non-owner sees every panel at most twice (forward, backward)
