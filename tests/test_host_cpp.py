"""The native host side (bundle-adjustment_b200/host/jaicov_host.hpp: C++ mirror of the reference's Java classes above the
C ABI) against the reference's own bookkeeping executed (tests/golden/reference_bookkeeping.npz): networks are built through
the mirror's class API (Camera, Image.add, DistortionModel.add, ScaleBar, DirectlyObservedParameterGroup, BundleAdjustment.add)
and prepareUnknownParameters / detectRankDefect must give the same rows, columns, counts, rank-defect flags and sigma0^2 --
bit-exact -- as BundleAdjustment.java:667-782, :836-1042."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K = np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_bookkeeping.npz'))
BK_SCENES = sorted({k.split('__')[0] for k in K.files})
KIND = {'point': 0, 'io': 1, 'coef': 2, 'eo': 3}


@pytest.fixture(scope='module')
def H(built):
    L = ctypes.CDLL(os.path.join(ROOT, 'bundle-adjustment_b200', 'libjaicov_host.so'))
    L.jhost_create.restype = ctypes.c_void_p
    L.jhost_last_error.restype = ctypes.c_char_p
    L.jhost_last_error.argtypes = [ctypes.c_void_p]
    L.jhost_destroy.argtypes = [ctypes.c_void_p]
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Net:
    """A scene dict (tests/scenes.py) handed to the C++ mirror through its flat entry points."""

    def __init__(self, L, scene):
        self.L, self.h = L, ctypes.c_void_p(L.jhost_create())
        pts = scene['points']
        xyz = np.ascontiguousarray(pts['xyz'], np.float64)
        fixed = np.ascontiguousarray(pts['fixed'], np.uint8)
        datum = np.ascontiguousarray(pts['datum'], np.uint8)
        self.n_pt = xyz.shape[0]
        self.ok(L.jhost_add_points(self.h, self.n_pt, _p(xyz), _p(fixed), _p(datum)))
        self.n_img = self.n_coef = 0
        self.n_cam = len(scene['cameras'])
        for c in scene['cameras']:
            ct = np.array([t for (t, _o, _v, _f) in c['coefs']], np.int32)
            co = np.array([o for (_t, o, _v, _f) in c['coefs']], np.int32)
            cv = np.array([v for (_t, _o, v, _f) in c['coefs']], np.float64)
            cf = np.array([f for (_t, _o, _v, f) in c['coefs']], np.uint8)
            idx = ctypes.c_int(-1)
            self.ok(L.jhost_add_camera(self.h, ctypes.c_double(c['r0']), _p(np.ascontiguousarray(c['io_val'], np.float64)),
                                       _p(np.ascontiguousarray(c['io_fixed'], np.uint8)), len(ct), _p(ct), _p(co), _p(cv), _p(cf), ctypes.byref(idx)))
            self.n_coef += len(ct)
            for im in c['images']:
                obj = np.ascontiguousarray(im['obj'], np.int32)
                xy = np.ascontiguousarray(im['xy'], np.float64)
                sg = np.ascontiguousarray(np.broadcast_to(np.asarray(im['sigma'], np.float64), xy.shape))
                rho = np.ascontiguousarray(im['rho'], np.float64)
                self.ok(L.jhost_add_image(self.h, idx.value, _p(np.ascontiguousarray(im['eo_val'], np.float64)),
                                          _p(np.ascontiguousarray(im['eo_fixed'], np.uint8)), ctypes.c_int64(obj.size), _p(obj), _p(xy), _p(sg), _p(rho)))
                self.n_img += 1
        for (a, b, l, s) in scene.get('scale_bars', []):
            self.ok(L.jhost_add_scale_bar(self.h, int(a), int(b), ctypes.c_double(l), ctypes.c_double(s)))
        for g in scene.get('observed_groups', []):
            kind = np.array([KIND[k] for (k, _i, _c) in g['refs']], np.int32)
            index = np.array([i for (_k, i, _c) in g['refs']], np.int32)
            comp = np.array([c for (_k, _i, c) in g['refs']], np.int32)
            obs = np.ascontiguousarray(g['obs'], np.float64)
            var = None if g.get('var') is None else np.ascontiguousarray(g['var'], np.float64)
            disp = None if g.get('dispersion') is None else np.ascontiguousarray(g['dispersion'], np.float64)
            self.ok(L.jhost_add_group(self.h, len(kind), _p(kind), _p(index), _p(comp), _p(obs), _p(var), _p(disp)))

    def ok(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.jhost_last_error(self.h).decode())

    def prepare(self):
        counts, flags, s2 = np.zeros(6, np.int32), np.zeros(7, np.int32), ctypes.c_double(0)
        self.ok(self.L.jhost_prepare(self.h, _p(counts), _p(flags), ctypes.byref(s2)))
        return counts, flags, s2.value

    def columns(self):
        pt, io, cf, eo = (np.zeros(max(n, 1), np.int64) for n in (3 * self.n_pt, 3 * self.n_cam, self.n_coef, 6 * self.n_img))
        self.ok(self.L.jhost_get_columns(self.h, _p(pt), _p(io), _p(cf), _p(eo), None, None, None, None))
        return pt[:3 * self.n_pt].reshape(-1, 3), io[:3 * self.n_cam], cf[:self.n_coef], eo[:6 * self.n_img]

    def close(self):
        self.L.jhost_destroy(self.h)


@pytest.mark.parametrize('name', BK_SCENES)
def test_cpp_host_bookkeeping_matches_executed_reference(H, name):
    from tests.test_reference_formulas import _bk_scene
    g = lambda k: K['%s__%s' % (name, k)]
    net = Net(H, _bk_scene(name))
    counts, flags, s2 = net.prepare()
    assert counts.tolist() == [int(v) for v in g('counts')]          # observations, unknowns, nIO, nDist, d, object points
    assert [bool(f) for f in flags] == g('flags').tolist()
    pt, io, cf, eo = net.columns()
    np.testing.assert_array_equal(pt, g('pt_col'))
    np.testing.assert_array_equal(io, g('io_col'))
    np.testing.assert_array_equal(cf, g('coef_col'))
    np.testing.assert_array_equal(eo, g('eo_col'))
    assert s2 == g('sigma2')[0]
    net.close()


def test_cpp_host_errors_and_no_cpu_path(H):
    """Error behaviour of the mirror (the Java throws IllegalArgumentException) and the loud failure without a device."""
    from tests.scenes import synthetic_scene
    import bundle_adjustment_b200 as ba
    scene = synthetic_scene(2, images=4, targets=20)[0]
    bad = dict(scene, scale_bars=[(0, 1, 100.0, 0.0)])               # sigma = 0 -> variance not positive
    with pytest.raises(RuntimeError, match='variance must be positive'):
        Net(H, bad)
    cam = dict(scene['cameras'][0])
    cam['coefs'] = cam['coefs'] + [(121, 1, 0.0, False)]              # A1 twice
    with pytest.raises(RuntimeError, match='order already exists'):
        Net(H, dict(scene, cameras=[cam]))
    if ba._lib.load().jaicov_device_count() == 0:
        net = Net(H, scene)
        state = ctypes.c_int(0)
        assert H.jhost_estimate(net.h, ctypes.byref(state)) == 0
        assert state.value == -5                                     # NOT_INITIALISED: no sm_100 device, no CPU path
        assert b'sm_100' in H.jhost_last_error(net.h)
        net.close()


def _flat(net):
    L = net.L
    sizes = np.zeros(8, np.int64)
    nul = [None] * 22
    net.ok(L.jhost_get_flat(net.h, _p(sizes), *nul))
    nc, nk, ni, m, npt, nb = (int(v) for v in sizes[:6])
    A = lambda n, t: np.zeros(max(n, 1), t)
    bufs = dict(io_val=A(3 * nc, np.float64), io_col=A(3 * nc, np.int32), coef_ptr=A(nc + 1, np.int32), coef_type=A(nk, np.int32),
                coef_order=A(nk, np.int32), coef_val=A(nk, np.float64), coef_col=A(nk, np.int32), cam_of_img=A(ni, np.int32),
                eo_val=A(6 * ni, np.float64), eo_col=A(6 * ni, np.int32), pt_ptr=A(ni + 1, np.int64), obj_idx=A(m, np.int32),
                xy=A(2 * m, np.float64), var=A(2 * m, np.float64), rho=A(m, np.float64), xyz=A(3 * npt, np.float64),
                pt_col=A(3 * npt, np.int32), is_datum=A(npt, np.uint8), bar_a=A(nb, np.int32), bar_b=A(nb, np.int32),
                free_flags=A(7, np.int32), point_of_scene=A(net.n_pt, np.int32))
    net.ok(L.jhost_get_flat(net.h, _p(sizes), *[_p(b) for b in bufs.values()]))
    n = dict(io_val=3 * nc, io_col=3 * nc, coef_ptr=nc + 1, coef_type=nk, coef_order=nk, coef_val=nk, coef_col=nk, cam_of_img=ni, eo_val=6 * ni,
             eo_col=6 * ni, pt_ptr=ni + 1, obj_idx=m, xy=2 * m, var=2 * m, rho=m, xyz=3 * npt, pt_col=3 * npt, is_datum=npt, bar_a=nb, bar_b=nb,
             free_flags=7, point_of_scene=net.n_pt)
    return {k: v[:n[k]] for k, v in bufs.items()}, sizes


@pytest.mark.parametrize('name', ['example', 'random2', 'random5', 'config3_observed_points', 'observed_eo_io'])
def test_cpp_host_flattens_the_same_problem_as_the_python_mirror(H, name):
    """What the C++ estimateModel() hands to the C ABI is the problem the Python mirror hands over (the one the GPU parity tests run),
    up to the numbering of the object points: the Python mirror lists them in scene order, the C++ mirror in order of first
    appearance -- a relabelling of obj_idx / bar endpoints / point rows, nothing else."""
    from tests.helpers import flat_problem
    from tests.test_reference_formulas import _bk_scene
    if name not in BK_SCENES:
        pytest.skip('scene not in the fixture')
    scene = _bk_scene(name)
    _adj, py = flat_problem(scene)
    net = Net(H, scene)
    cp, sizes = _flat(net)
    pos = cp['point_of_scene']                        # scene point -> flat point of the C++ mirror
    for k in ('io_val', 'io_col', 'coef_ptr', 'coef_type', 'coef_order', 'coef_val', 'coef_col', 'cam_of_img', 'eo_val', 'eo_col', 'pt_ptr',
              'xy', 'var', 'rho', 'free_flags'):
        np.testing.assert_array_equal(cp[k], np.asarray(py[k]).reshape(-1), err_msg=k)
    assert int(sizes[7]) == int(py['n_unknowns'])
    np.testing.assert_array_equal(cp['obj_idx'], pos[np.asarray(py['obj_idx'])])
    np.testing.assert_array_equal(cp['bar_a'], pos[np.asarray(py['bar_a'], np.int64)] if len(py['bar_a']) else cp['bar_a'])
    np.testing.assert_array_equal(cp['bar_b'], pos[np.asarray(py['bar_b'], np.int64)] if len(py['bar_b']) else cp['bar_b'])
    used = pos >= 0
    py_xyz, py_col, py_dat = np.asarray(py['xyz']).reshape(-1, 3), np.asarray(py['pt_col']).reshape(-1, 3), np.asarray(py['is_datum'])
    np.testing.assert_array_equal(cp['xyz'].reshape(-1, 3)[pos[used]], py_xyz[used])
    np.testing.assert_array_equal(cp['pt_col'].reshape(-1, 3)[pos[used]], py_col[used])
    np.testing.assert_array_equal(cp['is_datum'][pos[used]], py_dat[used])
    assert not py_dat[~used].any()                    # points outside the C++ problem take no part in the Python one either
    assert sorted(pos[used].tolist()) == list(range(int(sizes[4])))
    # directly observed groups: targets (points relabelled), observations, variances / the switch to a packed dispersion
    assert int(sizes[6]) == len(py['groups'])
    for gi, pg in enumerate(py['groups']):
        r, hs = ctypes.c_int32(0), ctypes.c_int32(0)
        net.ok(H.jhost_get_flat_group(net.h, gi, ctypes.byref(r), ctypes.byref(hs), None, None, None, None, None))
        assert r.value == len(pg['obs']) and bool(hs.value) == (pg['sigma'] is not None)
        kind, index, comp = (np.zeros(r.value, np.int32) for _ in range(3))
        obs, var = np.zeros(r.value), np.zeros(r.value)
        net.ok(H.jhost_get_flat_group(net.h, gi, ctypes.byref(r), ctypes.byref(hs), _p(kind), _p(index), _p(comp), _p(obs), _p(var)))
        pk, pi, pc = np.asarray(pg['kind']), np.asarray(pg['index']), np.asarray(pg['comp'])
        np.testing.assert_array_equal(kind, pk)
        np.testing.assert_array_equal(index, np.where(pk == 0, pos[pi], pi))
        np.testing.assert_array_equal(comp, pc)
        np.testing.assert_array_equal(obs, np.asarray(pg['obs'], float))
        if pg['var'] is not None:
            np.testing.assert_array_equal(var, np.asarray(pg['var'], float))
    net.close()


def test_cpp_dlt_gathers_the_same_homologous_points_as_the_python_mirror(H, monkeypatch):
    """DirectLinearTransformation (dlt/DirectLinearTransformation.java:78-94, :279-314): what the C++ mirror hands to
    jaicov_dlt_batch -- homologous points matched by name, the camera's interior orientation -- equals what the Python mirror hands
    over, on a network where only some of the object points are known; and there is no CPU path behind it."""
    import bundle_adjustment_b200 as ba
    from tests.helpers import build_adjustment
    from tests.scenes import synthetic_scene
    scene = synthetic_scene(2, images=5, targets=40, visibility=0.7)[0]
    adj, pts = build_adjustment(scene)
    images = [img for cam in adj.getCameras() for img in cam]
    known = [i for i in range(40) if i % 3 != 1]
    captured = {}

    def fake(pt_ptr, xy, xyz, io, restrictions=(), max_iterations=5000, device=0):
        captured.update(pt_ptr=np.asarray(pt_ptr), xy=np.asarray(xy).reshape(-1), xyz=np.asarray(xyz).reshape(-1), io=np.asarray(io).reshape(-1))
        n = len(pt_ptr) - 1
        return np.zeros((n, 20)), -np.ones(n, np.int32), np.zeros(n, np.int32)
    monkeypatch.setattr(ba._lib, 'dlt_batch', fake)
    ba.DirectLinearTransformation.adjustAll([ba.DLTCoefficients(img) for img in images], {pts.names[i]: pts[i] for i in known})
    net = Net(H, scene)
    kn = np.array(known, np.int32)
    m = sum(len(im['obj']) for c in scene['cameras'] for im in c['images'])
    pt_ptr, xy, xyz, io = np.zeros(net.n_img + 1, np.int64), np.zeros(2 * m), np.zeros(3 * m), np.zeros(3 * net.n_img)
    net.ok(H.jhost_dlt(net.h, len(kn), _p(kn), 0, None, 0, _p(pt_ptr), _p(xy), _p(xyz), _p(io), None, None))
    np.testing.assert_array_equal(pt_ptr, captured['pt_ptr'])
    k = int(pt_ptr[-1])
    assert 0 < k < m
    np.testing.assert_array_equal(xy[:2 * k], captured['xy'])
    np.testing.assert_array_equal(xyz[:3 * k], captured['xyz'])
    np.testing.assert_array_equal(io, captured['io'])
    if ba._lib.load().jaicov_device_count() == 0:
        out, ok = np.zeros(20 * net.n_img), np.zeros(net.n_img, np.uint8)
        assert H.jhost_dlt(net.h, len(kn), _p(kn), 0, None, 1, None, None, None, None, _p(out), _p(ok)) == -1
        assert b'jaicov_dlt_batch failed' in H.jhost_last_error(net.h)
        n_out = ctypes.c_int(0)
        pts_i, imgs_i = np.array([0, 1], np.int32), np.array([0, 1], np.int32)
        assert H.jhost_transform(net.h, 2, _p(pts_i), 0, 2, _p(imgs_i), ctypes.c_double(1.0), 0, ctypes.byref(n_out), None, None) == -1
        assert b'no cofactor matrix on the device' in H.jhost_last_error(net.h)
    net.close()


def test_cpp_default_result_writer_info_file_equals_the_python_writer(H, tmp_path):
    """DefaultResultWriter (util/io/writer/DefaultResultWriter.java:67): the .info file of the C++ mirror is byte-identical to the
    Python writer's on a network with fixed components (index -1) -- names, components, values, running indices."""
    from bundle_adjustment_b200.writers import DefaultResultWriter
    from tests.helpers import build_adjustment
    from tests.scenes import random_scene
    scene = random_scene(5)
    adj, _pts = build_adjustment(scene)
    adj._prepare()
    DefaultResultWriter(str(tmp_path / 'py')).export(adj)
    net = Net(H, scene)
    net.prepare()
    net.ok(H.jhost_export_default(net.h, str(tmp_path / 'cpp').encode()))
    a, b = (tmp_path / 'py.info').read_bytes(), (tmp_path / 'cpp.info').read_bytes()
    assert len(a) > 1000 and a == b
    assert b'-1\n' in a                                              # a fixed component is listed without a row / column
    assert not (tmp_path / 'cpp.cxx').exists()                        # no adjustment has run: no cofactor matrix to export
    net.close()


def test_cpp_and_python_mirrors_agree_on_random_networks(H):
    """Beyond the 18 reference-executed networks: 30 random networks (datum defects 0..7, fixed components, scale bars, two
    cameras) get the same columns, counts, rank-defect flags and sigma0^2 from both mirrors."""
    from tests.helpers import flat_problem
    from tests.scenes import random_scene
    defects = set()
    for seed in range(100, 130):
        scene = random_scene(seed)
        adj, flat = flat_problem(scene)
        net = Net(H, scene)
        counts, flags, s2 = net.prepare()
        pt, io, cf, eo = net.columns()
        assert (counts[0], counts[1], counts[4]) == (flat['n_observations'], flat['n_unknowns'], int(np.sum(flat['free_flags']))), seed
        assert list(flags) == list(flat['free_flags']), seed
        np.testing.assert_array_equal(pt.reshape(-1), np.asarray(flat['pt_col']), err_msg=str(seed))
        np.testing.assert_array_equal(io, flat['io_col'])
        np.testing.assert_array_equal(cf, flat['coef_col'])
        np.testing.assert_array_equal(eo, flat['eo_col'])
        assert s2 == adj.getVarianceFactorApriori()
        defects.add(int(counts[4]))
        net.close()
    assert len(defects) >= 3          # the sample really covers different datum situations


def test_writer_number_format_is_javas_not_printfs(H):
    """DefaultResultWriter prints with java.util.Formatter (DefaultResultWriter.java:67,142: "%35.15f", "%+35.15f", Locale.ENGLISH):
    the digits of Double.toString (shortest round-trip) rounded HALF_UP and zero-padded -- not printf's expansion of the binary value.
    Known answers of the Java formatter, and the Python and C++ writers agree on them and on 20 000 random doubles of all magnitudes."""
    from bundle_adjustment_b200.writers import java_format_f
    H.jhost_java_format_f.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    buf = ctypes.create_string_buffer(512)

    def cpp(v, w, p, plus=False):
        n = H.jhost_java_format_f(v, w, p, int(plus), buf, 512)
        assert n >= 0
        return buf.value.decode()

    known = [((0.1, 0, 20, False), '0.10000000000000000000'),          # printf: 0.10000000000000000555
             ((0.15, 0, 1, False), '0.2'),                             # printf: 0.1 (0.1499999...)
             ((1.005, 0, 2, False), '1.01'),                           # printf: 1.00
             ((573.00385393, 35, 15, False), '                573.003853930000000'),     # printf: ...573.003853929999991
             ((2.5e-7, 35, 15, True), '                 +0.000000250000000'),
             ((-0.0, 8, 3, True), '  -0.000'), ((0.0, 0, 3, True), '+0.000'), ((-1e-20, 0, 15, True), '-0.000000000000000'),
             ((9.9996, 0, 3, False), '10.000'), ((0.9995, 0, 3, False), '1.000'), ((1e22, 0, 2, False), '10000000000000000000000.00'),
             ((123456.0, 0, 0, False), '123456'), ((0.5, 0, 0, False), '1'), ((0.4999, 0, 0, False), '0'),
             ((float('nan'), 6, 2, False), '   NaN'), ((float('inf'), 0, 2, True), '+Infinity'), ((float('-inf'), 10, 2, False), ' -Infinity')]
    for args, want in known:
        assert java_format_f(*args) == want, args
        assert cpp(*args) == want, args
    assert '%35.15f' % 573.00385393 != java_format_f(573.00385393, 35, 15)       # the difference is real
    rng = np.random.default_rng(7)
    vals = np.concatenate([rng.standard_normal(8000) * 10.0 ** rng.integers(-18, 6, 8000), rng.standard_normal(4000) * 1e3,
                           np.round(rng.standard_normal(4000) * 100, 3), rng.integers(-10 ** 6, 10 ** 6, 2000) / 2.0 ** 10,
                           np.frombuffer(rng.bytes(16000), dtype=np.float64)])
    vals = vals[np.isfinite(vals) & (np.abs(vals) < 1e30)]
    assert vals.size > 19000
    for v in vals:
        a = java_format_f(v, 35, 15, True)
        assert a == cpp(float(v), 35, 15, True), repr(float(v))
        assert abs(float(a) - v) <= 5.0000001e-16 + 4e-16 * abs(v)       # and it is the value, to the printed precision
    for v in vals[:2000]:
        assert java_format_f(v, 0, 4) == cpp(float(v), 0, 4), repr(float(v))
