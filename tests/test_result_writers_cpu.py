"""SURVEY 8 row f-1 on the CPU: index selection, file layout and number format of the two result writers
(util/io/writer/DefaultResultWriter.java:46-155, MatlabResultWriter.java:92-223) with the device session replaced by a stub that
gathers from a known matrix -- the GPU tests check the same files against the oracle's cofactor matrix."""
import numpy as np
import pytest

import bundle_adjustment_b200 as ba
from bundle_adjustment_b200.workloads import build_adjustment, random_scene
from bundle_adjustment_b200.writers import java_format_f


class StubSession:
    """Stands in for the device: jaicov_get_qxx_submatrix = scale * Q[idx][:, idx] (reference column numbers, border included)."""

    def __init__(self, Q):
        self.Q = Q
        self.calls = []

    def qxx_submatrix(self, idx, scale):
        idx = np.asarray(idx)
        self.calls.append((idx.copy(), scale))
        return scale * self.Q[np.ix_(idx, idx)]


def _prepared(seed):
    adj, _pts = build_adjustment(random_scene(seed))
    adj._prepare()
    n = adj.getNumberOfUnknownParameters() + adj.getNumberOfDatumConditions()
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n))
    Q = G @ G.T * 1e-3
    adj._session = StubSession(Q)
    adj._omega = 3.7e-3                       # a posteriori variance factor = omega / dof
    return adj, Q


@pytest.mark.parametrize('seed', [5, 11])
def test_default_result_writer_files(tmp_path, seed):
    adj, Q = _prepared(seed)
    base = str(tmp_path / 'out')
    w = ba.DefaultResultWriter(base)
    w.export(adj)
    # the rule of DefaultResultWriter.java:75-118, restated: estimated components in getObjectCoordinates() order get running indices from 0
    want_idx, want_lines, k = [], [], 0
    for oc in adj.getObjectCoordinates():
        for comp, p in zip('XYZ', (oc.getX(), oc.getY(), oc.getZ())):
            c = p.getColumn()
            if 0 <= c < 2 ** 31 - 1:
                want_idx.append(c)
                ci = k
                k += 1
            else:
                ci = -1
            want_lines.append('%25s\t%5s\t%s\t%10d' % (oc.getName(), comp, java_format_f(p.getValue(), 35, 15), ci))
    assert w.indices == want_idx and len(want_idx) > 10
    assert open(base + '.info').read().splitlines() == want_lines
    assert any(l.endswith('        -1') for l in want_lines) or seed != 5            # seed 5 has fixed components
    s2 = adj.getVarianceFactorAposteriori()
    assert s2 == abs(3.7e-3 / adj.getDegreeOfFreedom())
    idx, scale = adj._session.calls[-1]
    assert list(idx) == want_idx and scale == s2                                     # ONE gather of exactly the exported sub-matrix
    rows = open(base + '.cxx').read().split('\n')
    assert rows[-1] == '' and len(rows) == len(want_idx) + 1
    ref = s2 * Q[np.ix_(want_idx, want_idx)]
    for r, line in enumerate(rows[:-1]):
        assert len(line) == 37 * len(want_idx)                                       # "%+35.15f" + two blanks per entry (:142)
        cells = [line[37 * c:37 * c + 35] for c in range(len(want_idx))]
        assert all(line[37 * c + 35:37 * c + 37] == '  ' for c in range(len(want_idx)))
        assert cells == [java_format_f(v, 35, 15, plus=True) for v in ref[r]]
        assert all(cell.lstrip()[0] in '+-' for cell in cells)
    np.testing.assert_allclose(np.loadtxt(base + '.cxx'), ref, rtol=0, atol=5.1e-16)


def test_default_result_writer_without_cofactor_matrix(tmp_path):
    adj, _Q = _prepared(5)
    adj.setInvertNormalEquation(ba.MatrixInversion.NONE)
    base = str(tmp_path / 'none')
    ba.DefaultResultWriter(base).export(adj)
    assert (tmp_path / 'none.info').exists() and not (tmp_path / 'none.cxx').exists()   # :128-133: no matrix, no file
    assert adj._session.calls == []


def test_matlab_result_writer_variables(tmp_path):
    from scipy.io import loadmat
    adj, Q = _prepared(11)
    base = str(tmp_path / 'adjustment_results')
    w = ba.MatlabResultWriter(base)
    w.export(adj)
    m = loadmat(base + '.mat')
    for key in ('variance_of_unit_weight_prio', 'variance_of_unit_weight_post', 'degree_of_freedom', 'number_of_observations',
                'number_of_unknowns', 'coordinates', 'interior_orientations', 'distortion_parameters', 'dispersion'):
        assert key in m, key
    idx = np.array(w.indices)
    np.testing.assert_array_equal(m['dispersion'], Q[np.ix_(idx, idx)])               # the UNSCALED cofactors (:209-223)
    assert adj._session.calls[-1][1] == 1.0
    # running indices count from 1 (:96-140) over points, then interior orientation, then distortion parameters; -1 = not exported
    flat = lambda a: [int(np.asarray(x).ravel()[0]) for x in np.asarray(a).ravel()]
    cov = [c for f in ('covx', 'covy', 'covz') for c in flat(m['coordinates'][f])]
    cov += flat(m['interior_orientations']['cov']) + flat(m['distortion_parameters']['cov'])
    used = sorted(c for c in cov if c > 0)
    assert used == list(range(1, idx.size + 1))
    assert int(m['number_of_unknowns'].ravel()[0]) == adj.getNumberOfUnknownParameters()
    assert float(m['variance_of_unit_weight_post'].ravel()[0]) == adj.getVarianceFactorAposteriori()


def test_matlab_result_writer_without_cofactor_matrix(tmp_path):
    """MatrixInversion.NONE: no "dispersion" variable, and -- as in the reference (:150-160, :175-187) -- the interior-orientation and
    distortion structs carry no "cov" field at all, while the coordinates keep their running covx / covy / covz (:107-135)."""
    from scipy.io import loadmat
    adj, _Q = _prepared(11)
    adj.setInvertNormalEquation(ba.MatrixInversion.NONE)
    base = str(tmp_path / 'r')
    ba.MatlabResultWriter(base).export(adj)
    m = loadmat(base + '.mat')
    assert 'dispersion' not in m and adj._session.calls == []
    assert 'cov' not in m['interior_orientations'].dtype.names and 'cov' not in m['distortion_parameters'].dtype.names
    assert {'cam_id', 'name', 'value'} <= set(m['interior_orientations'].dtype.names)
    assert {'cam_id', 'name', 'value', 'order'} <= set(m['distortion_parameters'].dtype.names)
    assert {'name', 'X', 'Y', 'Z', 'covx', 'covy', 'covz'} <= set(m['coordinates'].dtype.names)
    assert m['coordinates'].shape[0] == 1 and m['interior_orientations'].shape == (1, 3 * len(adj.getCameras()))


# ---- against what the reference's own writer code produced when executed (tests/golden/make_writer_fixture.py) -----------------------------
def _fixture_network():
    """The network of make_writer_fixture.network()."""
    from bundle_adjustment_b200.workloads import synthetic_scene
    scene = synthetic_scene(2, images=4, targets=14, seed=77, n_cameras=2)[0]
    fixed = np.array(scene['points']['fixed'], bool).reshape(-1, 3)
    fixed[2, 1] = fixed[5, 2] = True
    fixed[9, :] = True
    scene['points']['fixed'] = fixed
    scene['cameras'][1]['io_fixed'] = [True, False, False]
    c0 = scene['cameras'][0]['coefs']
    scene['cameras'][0]['coefs'] = [(t, o, v, f or (t == 121 and o == 3) or t == 141) for (t, o, v, f) in c0]
    adj, _pts = build_adjustment(scene)
    adj._prepare()
    return adj


@pytest.mark.parametrize('tag', ['full', 'none'])
def test_writers_match_the_executed_reference(tmp_path, tag):
    """Row f-1 against tests/golden/reference_writers.npz: the arguments the reference's DefaultResultWriter passes to printf (format
    strings, names, components, values, running indices; every cell of the .cxx matrix) and the variables, struct fields and the
    dispersion matrix its MatlabResultWriter hands to the MAT-file library -- with and without a cofactor matrix."""
    import json
    import os

    from scipy.io import loadmat
    E = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_writers.npz'))
    g = lambda k: E['%s__%s' % (tag, k)]
    adj = _fixture_network()
    Q = E['Q']
    assert Q.shape[0] == adj.getNumberOfUnknownParameters() + adj.getNumberOfDatumConditions()
    adj._session = StubSession(Q)
    if tag == 'none':
        adj.setInvertNormalEquation(ba.MatrixInversion.NONE)
    s2post = 1.25 * adj.getVarianceFactorApriori()                     # what the fixture's stub adjustment reports
    adj.getVarianceFactorAposteriori = lambda: s2post
    # DefaultResultWriter
    base = str(tmp_path / 'd')
    w = ba.DefaultResultWriter(base)
    w.export(adj)
    assert list(w.indices) == list(g('default_indices'))
    assert list(g('info_format')) == ['%25s\t%5s\t%35.15f\t%10d%n']
    want = ['%25s\t%5s\t%s\t%10d' % (n, c, java_format_f(v, 35, 15), i)
            for n, c, v, i in zip(g('info_names'), g('info_comp'), g('info_value'), g('info_index'))]
    assert open(base + '.info').read().splitlines() == want
    assert os.path.exists(base + '.cxx') == bool(g('cxx_written'))
    if bool(g('cxx_written')):
        assert list(g('cxx_format')) == ['%+35.15f  ', '%n']
        ref = g('cxx_values')
        lines = open(base + '.cxx').read().split('\n')
        assert lines[-1] == '' and len(lines) == ref.shape[0] + 1
        for line, row in zip(lines, ref):
            assert line == ''.join(java_format_f(v, 35, 15, plus=True) + '  ' for v in row)
    # MatlabResultWriter
    mbase = str(tmp_path / 'm')
    ba.MatlabResultWriter(mbase).export(adj)
    assert str(E['%s__mat_path' % tag]).endswith('.mat')
    m = loadmat(mbase + '.mat')
    desc = json.loads(str(g('mat_json')))
    assert [k for k in m if not k.startswith('__')] == list(g('mat_variables'))         # the same variables in the same order
    scalar = lambda v: np.asarray(v).ravel()[0]
    for name in ('variance_of_unit_weight_prio', 'variance_of_unit_weight_post', 'degree_of_freedom', 'number_of_observations', 'number_of_unknowns'):
        kind, val = desc[name]
        assert scalar(m[name]) == val and str(m[name].dtype) == {'double': 'float64', 'int32': 'int32'}[kind], name
    for name in ('coordinates', 'interior_orientations', 'distortion_parameters'):
        d = desc[name]
        assert list(m[name].shape) == d['shape'], name                                # 1 x N struct arrays
        assert set(m[name].dtype.names) == set(d['fields']), name                     # "cov" only with a dispersion matrix
        for f, vals in d['fields'].items():
            got = [scalar(x) for x in m[name][f].ravel()]
            assert len(got) == len(vals), (name, f)
            for a, (kind, b) in zip(got, vals):
                assert (str(a) == b) if kind == 'string' else (a == b), (name, f, a, b)
                if kind in ('int32', 'int64'):
                    assert str(np.asarray(a).dtype) == kind, (name, f)
    if 'dispersion' in desc:
        np.testing.assert_array_equal(m['dispersion'], g('mat_dispersion'))
