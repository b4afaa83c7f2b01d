"""Numerics of FP64 products built from int8 digit products (experiment JAICOV_GEMM_OZAKI): the
product's blocked Cholesky + inverse schedule (csrc/dense_driver.hpp) runs on the TEST-ONLY host backend with every GEMM
launch computed by the integer emulation of tests/emul/host_backend.cpp -- per-row power-of-two scaling, `s` signed digits
|q| <= 64, exact integer sums per digit-sum group (what an s8 x s8 -> s32 tensor-core accumulator holds), groups beyond
s + 1 dropped, FP64 Horner combination -- and is compared with the FP64 schedule on the same systems."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))            # ozaki_gpu_check (the case list of the GPU checker)

from tests import ozaki_study as oz   # noqa: E402


@pytest.fixture(scope='module')
def emul(built):
    return oz.load_emul()


def deviations(emul, S, rhs, digits):
    Xref = oz.reference_inverse(S)
    sc = np.sqrt(np.abs(np.diag(Xref))).astype(np.float64)
    out = {}
    for s in [0] + list(digits):
        info, Q, y, st, _ = oz.run(emul, S, rhs, s)
        assert info == 0
        out[s] = (float(np.max(np.abs(Q.astype(np.longdouble) - Xref) / np.outer(sc, sc))), st.copy())
    return out


def test_eight_digits_match_the_fp64_schedule_on_a_random_system(emul):
    rng = np.random.default_rng(5)
    n = 300                                           # padded to 384 = 3 tiles: recursion, K_B_LOWER / K_A_LOWER / K_MAX_IJ launches
    A = rng.standard_normal((n, n))
    S = A @ A.T + 0.05 * n * np.eye(n)
    dd = 1 / np.sqrt(np.diag(S))
    S = S * dd[:, None] * dd[None, :]
    dev = deviations(emul, S, rng.standard_normal(n), (5, 8))
    e64, e5, e8 = dev[0][0], dev[5][0], dev[8][0]
    assert dev[8][1][3] > 0                           # launches went through the integer path
    assert dev[8][1][4] < 2 ** 31                     # every integer group sum fits an s32 accumulator
    assert e8 <= 8 * e64 + 1e-15, (e8, e64)           # eight digits: FP64-equivalent
    assert e5 > 100 * e64                             # five digits are visibly short: the test can tell the difference


def test_eight_digits_on_a_bundle_network(emul):
    """M~ = V N V + B~'B~ of a free network (config 2, datum border folded in): the cofactor matrix from int8 digit products
    stays two orders of magnitude inside the 1e-8 parity bar, like the FP64 schedule."""
    from tests.scenes import synthetic_scene
    scene, _ = synthetic_scene(2, images=8, targets=60)
    S, rhs = oz.scaled_system(scene)
    dev = deviations(emul, S, rhs, (8,))
    assert dev[8][0] <= 1e-10 and dev[0][0] <= 1e-10, dev
    assert dev[8][0] <= 8 * dev[0][0] + 1e-14


@pytest.mark.parametrize('digits', [0, 8])
def test_tile_grid_products_with_poisoned_operands(emul, digits):
    """The case list tools/ozaki_gpu_check.py will run on the B200 (every operand layout, triangular hint and symmetric
    output; NaN wherever the kernels must not read; rows of very different magnitude), here through the host backend: the
    FP64 loops and the int8-digit emulation both meet the 1e-13 bar, so a failure on the GPU is the kernel's, not the checker's."""
    import ozaki_gpu_check as chk
    emul.emul_gemm_tiles.argtypes = [ctypes.c_int] * 4 + [ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_int64,
                                     ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int]

    def gemm(As, Bs, C0, al, bl, alpha, beta, tri, kmode):
        As, Bs, C = np.ascontiguousarray(As), np.ascontiguousarray(Bs), C0.copy()
        Mr, K = As.shape if al == 0 else As.shape[::-1]
        Nr = Bs.shape[0] if bl == 0 else Bs.shape[1]
        n = emul.emul_gemm_tiles(al, bl, Mr // 128, Nr // 128, K, alpha, beta, As.ctypes.data, As.shape[1], Bs.ctypes.data, Bs.shape[1],
                                 C.ctypes.data, C.shape[1], int(tri), kmode, digits)
        assert n == (1 if digits else 0)
        return C

    cases = chk.run_gemm_cases(gemm)
    assert len(cases) >= 20
    bad = [c for c in cases if not c['ok']]
    assert not bad, bad[:3]


def test_structured_route_products_from_digit_products(emul):
    """The two big products of the structured route (Q'Y' and Y (Q'Y')) from 8 int8 digits.  Their operands -- rows
    of the inverse of the bordered reduced system -- span many orders of magnitude inside one row, which a per-row digit grid
    alone resolves ~10x worse than FP64 (digits = -8: balancing off); with the contraction index balanced by exact powers of two
    first (the default, also in csrc/ozaki.cu) the point block of the cofactor matrix is as accurate as with FP64 products."""
    from tests.scenes import synthetic_scene
    dev, up, m = oz.structured_study(emul, synthetic_scene(2, images=8, targets=60)[0], (8, -8))
    assert up == 180 and m == 65
    assert dev[0] <= 2e-12 and dev[8] <= 2 * dev[0] + 1e-15, dev
    assert 3 * dev[8] < dev[-8] <= 1e-10, dev          # without the balancing: visibly coarser, still inside the 1e-8 bar


@pytest.mark.parametrize('nb,nranks,pw,merged', [(6, 2, 1, 1), (7, 3, 2, 1), (7, 3, 2, 0)])
def test_distributed_schedule_from_digit_products(emul, nb, nranks, pw, merged):
    """The multi-GPU schedule (block-column-cyclic Cholesky with one trapezoid launch per panel -- column-table launches --, then
    every rank's column-tile inverse with its per-column-tile masks, csrc/dense_driver.hpp) with every eligible launch computed
    from 8 int8 digits: all virtual ranks end with the same factor and the gathered column tiles are the inverse.  This is the
    launch geometry csrc/ozaki.cu accepts beyond the single-GPU one (the kernel itself still has to see a multi-GPU run)."""
    emul.emul_distributed_ex.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_int]
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
    info = emul.emul_distributed_ex(n, M.ctypes.data, nranks, pw, Q.ctypes.data, ctypes.byref(md), merged, 8)
    assert info == 0 and md.value == 0.0
    np.testing.assert_allclose(np.tril(M), np.linalg.cholesky(S), atol=1e-13)
    Qi = np.linalg.inv(S)
    assert not np.isnan(np.tril(Q)).any()
    np.testing.assert_allclose(np.tril(Q), np.tril(Qi), atol=1e-12 * np.abs(Qi).max())
