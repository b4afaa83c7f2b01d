"""The oracles of the widened rows against golden vectors produced by the REFERENCE's own closed-form expressions
(tests/golden/make_formula_fixtures.py parses and evaluates the Java right-hand sides; the fixture holds numbers only):

* oracle/propagation.py: 45 Jacobian entries + transformed point, tranformation/CoordinateTransformationExteriorOrientation.java:160-279
* oracle/dlt.py: restriction gradients / misclosures, dlt/DLTPartialDerivativeFactory.java:100-236, and the expansion of the
  coefficients, dlt/DirectLinearTransformation.java:208-246
"""
import os

import numpy as np
import pytest

from oracle import dlt as od
from oracle import propagation as op

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_formulas.npz'))


def test_transformation_jacobian_matches_reference_expressions():
    assert G['transform_jacobian'].shape == (12, 3, 15)
    for inp, J_ref, xyz_ref in zip(G['transform_inputs'], G['transform_jacobian'], G['transform_xyz']):
        eT, eS, X = inp[0:6], inp[6:12], inp[12:15]
        np.testing.assert_allclose(op.transform_point(X, eT, eS), xyz_ref, rtol=1e-13, atol=1e-9)
        J = op.jacobian_block(X, eT, eS)
        scale = np.abs(J_ref).max()
        np.testing.assert_allclose(J, J_ref, rtol=0, atol=1e-12 * scale)


def test_dlt_restrictions_match_reference_expressions():
    assert G['restriction_gradients'].shape == (12, 6, 11)
    for inp, g_ref, w_ref in zip(G['restriction_inputs'], G['restriction_gradients'], G['restriction_misclosures']):
        b, (c, x0, y0) = inp[:11], inp[11:]
        for kind in range(6):
            g, w = od.restriction_row(kind, b, c, x0, y0)
            np.testing.assert_allclose(g, g_ref[kind], rtol=1e-12, atol=1e-13 * max(1.0, np.abs(g_ref[kind]).max()))
            assert w == pytest.approx(w_ref[kind], rel=1e-12, abs=1e-13)


def test_dlt_expansion_matches_reference_expressions():
    """x0, y0, cx, cy, the determinant rule and omega / phi / kappa (DLT:208-246) for random coefficient sets."""
    flips = 0
    for b, out in zip(G['expansion_inputs'], G['expansion_outputs']):
        x0, y0, cx, cy = out[:4]
        det, omega, phi, kappa = out[13:]
        flips += det < 0
        _, d = od.expand(b, 1.0)
        np.testing.assert_allclose([d['x0'], d['y0'], d['cx'], d['cy']], [x0, y0, cx, cy], rtol=1e-12)
        np.testing.assert_allclose([d['omega'], d['phi'], d['kappa']], [omega, phi, kappa], rtol=1e-11, atol=1e-13)
    assert 0 < flips < len(G['expansion_inputs'])          # both branches of the determinant rule occur


# ---- main path: per-image-point Jacobian rows and misclosures -------------------------------------------------------------------
J = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_jacobian.npz'))
CASES = sorted(k[:-2] for k in J.files if k.endswith('_A'))


@pytest.mark.parametrize('case', CASES)
def test_point_jacobian_matches_reference_formulas(case):
    """oracle/jaicov_oracle.c (K1: collinearity partials PDF:96-193, chain rule DMF:33-101, all five distortion model
    factories incl. the three Zernike models) against vectors produced by executing the reference's own formulas
    (tests/golden/make_jacobian_fixture.py)."""
    from oracle.oracle import FlatProblem, eval_point
    coefs = [(int(t), int(o), float(v), False) for t, o, v in J[case + '_coefs']]
    worst = 0.0
    for inp, A_ref, w_ref in zip(J[case + '_inputs'], J[case + '_A'], J[case + '_w']):
        io, X0, ang, X, obs = inp[0:3], inp[3:6], inp[6:9], inp[9:12], inp[12:14]
        scene = {'points': {'xyz': X.reshape(1, 3).copy(), 'fixed': np.zeros((1, 3), bool), 'datum': np.ones(1, bool)},
                 'cameras': [{'r0': 10.0, 'io_val': io.copy(), 'io_fixed': np.zeros(3, bool), 'coefs': list(coefs),
                              'images': [{'eo_val': np.concatenate([X0, ang]), 'eo_fixed': np.zeros(6, bool), 'obj': np.array([0], np.int32),
                                          'xy': obs.reshape(1, 2).copy(), 'sigma': np.full((1, 2), 0.001), 'rho': np.zeros(1)}]}],
                 'scale_bars': [], 'observed_groups': []}
        fp = FlatProblem(scene)
        cols, a0, a1, w, _ = eval_point(fp, 0, 0, 1e-6)
        assert a0.size == A_ref.shape[1]
        scale = max(np.abs(A_ref).max(), 1e-300)
        # per entry: relative to the entry, with a floor of 1e-13 of the largest entry of the row pair
        tol = 1e-11 * np.abs(A_ref) + 1e-13 * scale
        assert (np.abs(np.stack([a0, a1]) - A_ref) <= tol).all(), (case, np.abs(np.stack([a0, a1]) - A_ref).max())
        np.testing.assert_allclose(w, w_ref, rtol=0, atol=1e-12 * max(1.0, np.abs(obs).max()))
        worst = max(worst, float((np.abs(np.stack([a0, a1]) - A_ref) / (np.abs(A_ref) + 1e-3 * scale)).max()))
    print(case, 'worst relative entry difference', worst)


# ---- datum condition rows ----------------------------------------------------------------------------------------------------------
D = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_datum_rows.npz'))


@pytest.mark.parametrize('key', sorted(D.files, key=lambda k: int(k[4:])))
def test_datum_rows_match_reference_method(key):
    """Oracle.datum_rows against BundleAdjustment.addDatumConditionRows (BA:493-635) executed on the same randomized
    networks (every combination of free translations / rotations / scale that occurs: d = 2 .. 7)."""
    from oracle.oracle import Oracle
    from tests.scenes import random_scene
    o = Oracle(random_scene(int(key[4:])), use_centroid=False)
    B = o.datum_rows()
    assert B.shape == D[key].shape and B.shape[0] == o.bk.d
    np.testing.assert_array_equal(B, D[key])


# ---- integer bookkeeping -------------------------------------------------------------------------------------------------------------
K = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_bookkeeping.npz'))
BK_SCENES = sorted({k.split('__')[0] for k in K.files})


def _bk_scene(name):
    from tests.scenes import example_scene, random_scene, synthetic_scene
    if name.startswith('random'):
        return random_scene(int(name[6:]))
    if name == 'example':
        return example_scene()
    if name == 'config2_small':
        return synthetic_scene(2, images=10, targets=60)[0]
    if name == 'config3_observed_points':
        return synthetic_scene(3, images=6, targets=40)[0]
    if name.startswith('observed_'):
        import importlib.util
        spec = importlib.util.spec_from_file_location('mbf', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'make_bookkeeping_fixture.py'))
        # only the two scene builders are used (pure numpy); the module reads /root/reference lazily, not at import
        mbf = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mbf)
        return mbf.observed_eo_io_scene() if name == 'observed_eo_io' else mbf.observed_omega_only_scene()
    return synthetic_scene(4, images=9, targets=70, free_network=False)[0]


@pytest.mark.parametrize('name', BK_SCENES)
def test_bookkeeping_matches_executed_reference(name):
    """Rows, columns, counts, rank-defect flags and sigma2apriori -- bit-exact -- of (a) the oracle's Bookkeeping and (b) the
    product's host side (bundle-adjustment_b200/host.py: BundleAdjustment._prepare, what fills the C ABI) against the values the
    reference's own prepareUnknownParameters / detectRankDefect produce when executed (tests/golden/make_bookkeeping_fixture.py)."""
    from oracle.bookkeeping import Bookkeeping
    from tests.helpers import flat_problem
    g = lambda k: K['%s__%s' % (name, k)]
    n_obs, n_unknown, n_io, n_dist, d, n_oc = (int(v) for v in g('counts'))
    # (a) oracle
    bk = Bookkeeping(_bk_scene(name))
    assert (bk.n_obs, bk.n_unknown, bk.n_io, bk.n_dist, bk.d, len(bk.oc_order)) == (n_obs, n_unknown, n_io, n_dist, d, n_oc)
    assert [bool(f) for f in bk.defect_free] == g('flags').tolist()
    np.testing.assert_array_equal(np.asarray(bk.pt_col).reshape(-1, 3), g('pt_col'))
    np.testing.assert_array_equal(np.concatenate(bk.io_col), g('io_col'))
    np.testing.assert_array_equal(np.concatenate(bk.coef_col) if g('coef_col').size else np.zeros(0, np.int64), g('coef_col'))
    np.testing.assert_array_equal(np.concatenate(bk.eo_col), g('eo_col'))
    assert bk.sigma2apriori == g('sigma2')[0]
    np.testing.assert_array_equal(np.concatenate([np.asarray(r, np.int64) for r in bk.group_rows]) if len(bk.group_rows) else np.zeros(0, np.int64),
                                  g('group_rows'))
    # (b) host mirror: the flat problem handed to the C ABI
    adj, flat = flat_problem(_bk_scene(name))
    assert (int(flat['n_observations']), int(flat['n_unknowns']), int(np.sum(flat['free_flags']))) == (n_obs, n_unknown, d)
    assert [bool(f) for f in flat['free_flags']] == g('flags').tolist()
    np.testing.assert_array_equal(np.asarray(flat['pt_col']).reshape(-1, 3), g('pt_col'))
    np.testing.assert_array_equal(np.asarray(flat['io_col']), g('io_col'))
    np.testing.assert_array_equal(np.asarray(flat['coef_col']), g('coef_col'))
    np.testing.assert_array_equal(np.asarray(flat['eo_col']), g('eo_col'))
    assert adj.getVarianceFactorApriori() == g('sigma2')[0]
    assert (adj.getNumberOfObservations(), adj.getNumberOfUnknownParameters(), adj.getNumberOfDatumConditions()) == (n_obs, n_unknown, d)


# ---- normal equations ------------------------------------------------------------------------------------------------------------------
NE = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_normal_equations.npz'))


def _ne_scene(name):
    from tests.scenes import random_scene, synthetic_scene
    if name == 'random2_scale_bar':
        return random_scene(2)
    if name == 'random0_two_cameras':
        return random_scene(0)
    if name == 'config3_rho_no_groups':
        sc = synthetic_scene(3, images=5, targets=30)[0]
        sc['observed_groups'] = []
        return sc
    if name == 'config3_observed_points_dispersion':
        return synthetic_scene(3, images=5, targets=25)[0]
    if name == 'observed_eo_io':
        return _bk_scene('observed_eo_io')
    sc = synthetic_scene(2, images=5, targets=30)[0]
    cam = sc['cameras'][0]
    cam['coefs'] = cam['coefs'][:4] + [(131, 1, 1e-4, False)] + cam['coefs'][4:] + [(161, 3, 1e-4, False), (162, 4, 3e-5, False), (163, 5, 2e-5, False)]
    return sc


@pytest.mark.parametrize('name', sorted({k.split('__')[0] for k in NE.files}))
def test_normal_equations_match_executed_reference(name):
    """K2/K3: the oracle's N (packed, datum border included) and n against the reference's own stacking path executed on the
    same network (tests/golden/make_normal_equation_fixture.py): per-point weights incl. correlated image coordinates
    (PDF:306-318), stackNormalEquationSystem (PDF:475-505), scale bars (PDF:210-283), directly observed groups with diagonal
    and fully populated dispersion (PDF:447-473, DOPG:67-91), datum rows (BA:493-635)."""
    from oracle.oracle import Oracle
    import ctypes
    from oracle.oracle import lib
    o = Oracle(_ne_scene(name), use_centroid=False)
    o.derive_first_damping, o.adapted_damping = False, 0.0
    N, n, V = o.create_normal_equation()
    assert N.shape == NE[name + '__N'].shape
    np.testing.assert_array_equal(N, NE[name + '__N'])
    np.testing.assert_array_equal(n, NE[name + '__n'])
    # Jacobi preconditioner (BA:824-828) and NormalEquationSystem.applyPrecondition (NES:82-91)
    np.testing.assert_array_equal(V, NE[name + '__V'])
    lib().orc_apply_precondition(o.fp.n, V.ctypes.data, N.ctypes.data, n.ctypes.data)
    np.testing.assert_array_equal(N, NE[name + '__N_preconditioned'])
    np.testing.assert_array_equal(n, NE[name + '__n_preconditioned'])
    # Levenberg-Marquardt damping of the diagonal in the first pass (BA:801-822)
    o2 = Oracle(_ne_scene(name), use_centroid=False, damping=0.7)
    o2.derive_first_damping, o2.adapted_damping = True, 0.0
    Nd, _, Vd = o2.create_normal_equation()
    np.testing.assert_array_equal(Nd, NE[name + '__N_damped'])
    np.testing.assert_array_equal(Vd, NE[name + '__V_damped'])
    assert o2.adapted_damping == 0.7 and not o2.derive_first_damping


@pytest.mark.parametrize('name', sorted({k.split('__')[0] for k in NE.files}))
def test_centroid_matches_executed_reference(name):
    """centroidCoordinates(false) (BA:115-201): the centroid, every shifted coordinate of points and projection centres, the
    shifted observations of directly observed coordinates -- or the refusal when the component counts differ (BA:151)."""
    from oracle.oracle import Oracle
    o = Oracle(_ne_scene(name), use_centroid=True)
    if name + '__centroid_refused' in NE.files:
        with pytest.raises(RuntimeError):
            o._centroid(False)
        return
    o._centroid(False)
    np.testing.assert_array_equal(np.asarray(o.centroid, float), NE[name + '__centroid'])
    np.testing.assert_array_equal(o.fp.xyz.reshape(-1, 3), NE[name + '__centroid_xyz'])
    np.testing.assert_array_equal(o.fp.eo_val.reshape(-1, 6), NE[name + '__centroid_eo'])
    gobs = np.concatenate([np.asarray(g['obs'], float) for g in o.fp.groups]) if o.fp.groups else np.zeros(0)
    np.testing.assert_array_equal(gobs, NE[name + '__centroid_obs'])


# ---- Levenberg-Marquardt step control and parameter update ------------------------------------------------------------------------------
def test_lm_step_control_matches_executed_reference():
    """Oracle._update_model / update_unknowns against BundleAdjustment.updateModel / updateUnknownParameters (BA:389-462)
    executed on 400 scripted situations (tests/golden/make_lm_fixture.py): step length, accept / reject, x0.2 / x5,
    the 1/sqrt(eps) clamp, max|dx| bookkeeping, the update itself -- all values identical."""
    import types
    from oracle.oracle import Oracle
    L = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_lm_steps.npz'))
    branches = set()
    for inp, out in zip(L['inputs'], L['outputs']):
        lam, prev, cur, complete, last_valid = inp[:5]
        dx, cols, vals = inp[5:11].copy(), inp[11:17].astype(np.int64), inp[17:23].copy()
        o = Oracle.__new__(Oracle)
        o.adapted_damping, o.omega, o.last_valid_max_abs_dx, o.max_abs_dx, o.lm_steps = float(lam), float(prev), float(last_valid), 0.0, []
        # the six scripted parameters spread over the oracle's four parameter stores
        o.fp = types.SimpleNamespace(xyz=vals[0:3], pt_col=cols[0:3], io_val=vals[3:4], io_col=cols[3:4], coef_val=vals[4:5], coef_col=cols[4:5],
                                     eo_val=vals[5:6], eo_col=cols[5:6])
        o.get_omega = lambda d, cur=cur: float(cur)
        o._update_model(dx, bool(complete))
        ev = o.lm_steps[0][:2] if o.lm_steps else (-1.0, -1.0)
        got = np.concatenate([[o.adapted_damping, o.omega, o.max_abs_dx, o.last_valid_max_abs_dx, ev[0], ev[1]], dx, vals])
        np.testing.assert_array_equal(got, out)
        branches.add((lam > 0, bool(o.lm_steps and not o.lm_steps[0][2]), o.adapted_damping == 1.0 / np.sqrt(2.0 ** -53)))
    assert len(branches) >= 4          # plain Gauss-Newton, accepted step, rejected step, clamped damping value


# ---- complete adjustments -----------------------------------------------------------------------------------------------------------------
E = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_estimates.npz'))


def _two_camera_reduced_scene():
    from tests.scenes import synthetic_scene
    sc = synthetic_scene(4, images=6, targets=40, n_cameras=2)[0]
    sc['cameras'][0]['images'][1]['eo_fixed'][4] = True
    return sc


def _estimate_case(name):
    from tests.scenes import random_scene, synthetic_scene
    small = lambda: synthetic_scene(2, images=5, targets=30)[0]
    table = {'config2_full': (small, {}), 'config2_none': (small, dict(invert='NONE')),
             'config2_simulation': (small, dict(simulation=True)), 'config2_lm_1': (small, dict(damping=1.0)),
             'config2_lm_100': (small, dict(damping=100.0)), 'config2_max_iter_3': (small, dict(max_iter=3)),
             'config2_reduced': (small, dict(invert='REDUCED')), 'config2_pre_elimination': (small, dict(invert='PRE_ELIMINATION')),
             'config4_two_cameras_reduced': (_two_camera_reduced_scene, dict(invert='REDUCED')),
             'config4_small': (lambda: synthetic_scene(4, images=6, targets=40)[0], {}),
             'config3_dispersion': (lambda: synthetic_scene(3, images=5, targets=25)[0], {}),
             'observed_eo_io': (lambda: _bk_scene('observed_eo_io'), {}),
             'random2_scale_bar_no_centroid': (lambda: random_scene(2), dict(use_centroid=False)),
             'random2_scale_bar_centroid_refused': (lambda: random_scene(2), dict(use_centroid=True))}
    make, kw = table[name]
    return make(), kw


@pytest.mark.parametrize('name', sorted({k.split('__')[0] for k in E.files}))
def test_complete_adjustment_matches_executed_reference(name):
    """Oracle.estimate() against the reference's estimateModel() executed end to end (tests/golden/make_estimate_fixture.py):
    final state, number of passes, the sequence of Levenberg-Marquardt steps, every adjusted parameter, Omega, sigma0^2
    a posteriori and the packed cofactor matrix.  Both sides call the same LAPACK for the solve, so the comparison is tight:
    identical integers, 1e-13 relative on parameters, 1e-11 on the correlation-scaled Qxx."""
    from oracle.oracle import Oracle
    scene, kw = _estimate_case(name)
    g = lambda k: E['%s__%s' % (name, k)]
    o = Oracle(scene, **kw)
    if int(g('status')[0]) == -999:
        with pytest.raises(RuntimeError):
            o.estimate()
        return
    status = o.estimate()
    assert status == int(g('status')[0])
    assert len(o.history) == int(g('passes')[0])
    lm = g('lm')
    assert len(o.lm_steps) == lm.shape[0]
    for (last, new, _acc), ref in zip(o.lm_steps, lm):
        assert last == ref[0] and new == ref[1]
    fp = o.fp
    for got, ref in ((fp.xyz.reshape(-1, 3), g('xyz')), (fp.io_val.reshape(-1, 3), g('io')), (fp.coef_val, g('coef')), (fp.eo_val.reshape(-1, 6), g('eo'))):
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(o.omega, g('omega')[0], rtol=1e-12)
    np.testing.assert_allclose(o.variance_factor_aposteriori(), g('sigma2')[0], rtol=1e-12)
    assert o.bk.dof == int(g('dof')[0])
    if g('qxx').size:
        Q, Qr = o.Qxx, g('qxx')
        # REDUCED / PRE_ELIMINATION: the cofactor matrix is the leading numRows block (BA:262); what lies behind it are the
        # leftovers of the reduction, which the oracle reproduces as well (compared on the block, where they are defined)
        n = o.fp.n if kw.get('invert', 'FULL') == 'FULL' else int(g('num_rows_reduced')[0])
        if kw.get('invert', 'FULL') != 'FULL':
            assert n == o.num_rows_reduced()
        idx = np.arange(n)
        sd = np.sqrt(np.abs(Qr[idx + idx * (idx + 1) // 2]))
        sd[:o.fp.d] = 1.0
        iu = np.triu_indices(n)
        scale = sd[iu[0]] * sd[iu[1]]
        k = iu[0] + iu[1] * (iu[1] + 1) // 2
        # FULL: same arithmetic up to the Omega sums -> 1e-11.  Reduced modes: the oracle forms the Schur complement with
        # vectorised sums (another summation order than the reference's scalar loops); on these systems (condition ~1e8) that
        # shows at the 1e-9 level, inside the 1e-8 bar the Qxx parity uses everywhere
        tol = 1e-11 if kw.get('invert', 'FULL') == 'FULL' else 1e-8
        assert (np.abs(Q[k] - Qr[k]) / scale).max() < tol


# ---- direct linear transformation, end to end ------------------------------------------------------------------------------------------
def test_dlt_adjust_matches_executed_reference():
    """oracle/dlt.py: adjust() against DirectLinearTransformation.adjust executed (tests/golden/make_dlt_fixture.py): the return
    value, the 11 back-scaled coefficients and the derived interior / exterior orientation of six images (one with too few
    points) for five restriction sets."""
    from oracle import dlt as od
    D = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_dlt.npz'))
    ptr, xy, xyz, io = D['pt_ptr'], D['xy'], D['xyz'], D['io']
    worst = 0.0
    for s in range(5):
        restr = [int(r) for r in D['set%d_restrictions' % s]]
        for k in range(ptr.size - 1):
            sl = slice(int(ptr[k]), int(ptr[k + 1]))
            ok, b, d, _ = od.adjust(xy[sl], xyz[sl], tuple(io), restr)
            assert bool(ok) == bool(D['set%d_ok' % s][k])
            if not ok:
                continue
            ref = D['set%d_values' % s][k]          # DLTCoefficients order: b11..b33, x0, y0, c, X0, Y0, Z0, omega, phi, kappa
            got = np.concatenate([b, [d['x0'], d['y0'], d['c']], d['X0'], [d['omega'], d['phi'], d['kappa']]])
            scale = np.maximum(np.abs(ref), 1e-3)
            worst = max(worst, float((np.abs(got - ref) / scale).max()))
    print('DLT oracle vs executed reference: worst relative difference', worst)
    # coefficients agree to 1e-14; the derived principal point (0.02 against c = 28.8) and projection centre carry the
    # cancellation of their formulas: 2e-11 relative to their own small values
    assert worst < 1e-10
